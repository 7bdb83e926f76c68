/* Minimal C host for libotn (include/otn.h): the cost, Riemannian gradient and one Hessian-vector product of a 2-site ansatz,
 * and a few batched trust-region iterations, from plain host arrays -- no Python, no torch.
 *
 *   gcc -O2 -I include examples/c_abi_demo.c -L opentn_b200 -lotn -Wl,-rpath,$PWD/opentn_b200 -lm -o c_abi_demo
 *   ./c_abi_demo target_re.bin x0.bin N m K        (raw little-endian doubles: D*D target, m*(4K)*4 isometries)
 *
 * Prints:  cost  |rgrad|_F  <eta,H eta>  cost_after_TR                                                              */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "otn.h"

static double* read_doubles(const char* path, size_t count) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    double* buf = (double*)malloc(count * sizeof(double));
    if (fread(buf, sizeof(double), count, f) != count) { fprintf(stderr, "%s: short read\n", path); exit(2); }
    fclose(f);
    return buf;
}

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, otn_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s target_re.bin x0.bin N m K\n", argv[0]); return 2; }
    const int N = atoi(argv[3]), m = atoi(argv[4]), K = atoi(argv[5]), d = 2;
    size_t D = 1;
    for (int i = 0; i < 2 * N; ++i) D *= d;
    const size_t np = (size_t)(4 * K) * 4, nx = (size_t)m * np;
    double* target = read_doubles(argv[1], D * D);
    double* x = read_doubles(argv[2], nx);

    otn_problem* h = NULL;
    CHECK(otn_create(&h, N, d, m, K, OTN_MODEL_LOCAL, OTN_METRIC_CANONICAL, /*max_batch*/ 1, /*n_targets*/ 1, target, NULL, /*device*/ 0));

    double f = 0.0;
    double* rgrad = (double*)malloc(nx * sizeof(double));
    double* heta = (double*)malloc(nx * sizeof(double));
    CHECK(otn_cost_grad(h, x, 1, &f, rgrad, NULL));
    double g2 = 0.0;
    for (size_t i = 0; i < nx; ++i) g2 += rgrad[i] * rgrad[i];
    /* Hessian-vector product along the gradient direction (a tangent vector), reusing the primal pass of otn_cost_grad */
    CHECK(otn_hvp_cached(h, rgrad, 1, heta));
    double ghg = 0.0;
    CHECK(otn_canonical_dot(x, rgrad, heta, 1, m, 4 * K, 4, &ghg, 0));

    otn_tr_params prm;
    otn_tr_default_params(&prm);
    prm.niter = 3;
    CHECK(otn_tr_run(h, x, 1, &prm, NULL));           /* x is updated in place */
    double f_after = 0.0;
    CHECK(otn_cost(h, x, 1, &f_after));

    printf("%.17g %.17g %.17g %.17g\n", f, sqrt(g2), ghg, f_after);
    otn_destroy(h);
    free(target); free(x); free(rgrad); free(heta);
    return 0;
}
