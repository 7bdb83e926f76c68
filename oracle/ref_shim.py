"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference under a jax->numpy module shim.

The reference (`/root/reference/opentn`) imports `jax`, `cvxpy`, `matplotlib`, `qutip`, `pytenet`,
none of which exist in this image.  Installing the stand-ins below into `sys.modules` lets the
reference's own `transformations`, `optimization`, `stiefel` (minus its three `jax.grad`/`jax.jvp`
call sites: stiefel.py:423,437,450,520) and `trust_region_rcopt` run unmodified on NumPy/SciPy.

This only works where `/root/reference` exists (the build container).  It is used by
`tests/golden/make_golden.py` to generate the committed fixtures and by `tests/test_oracle_vs_reference.py`
(skipped when the reference tree is absent, e.g. on the GPU box).  Nothing in `opentn_b200/` may import it.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("OPENTN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "opentn"))


def _empty(name):
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package so that sub-imports resolve through sys.modules
    return m


def install_shim():
    """Install the stand-in modules (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_otn_shim", False):
        return
    import numpy
    import scipy
    import scipy.linalg

    jax = _empty("jax")
    jax._otn_shim = True
    jnp = types.ModuleType("jax.numpy")
    for k in dir(numpy):
        if not k.startswith("__"):
            setattr(jnp, k, getattr(numpy, k))
    jnp.linalg = numpy.linalg
    jsp = _empty("jax.scipy")
    jsp.linalg = scipy.linalg
    jax.numpy = jnp
    jax.scipy = jsp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    cfg = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.config = cfg

    def _no_autodiff(*a, **k):
        raise NotImplementedError("jax autodiff is not available under the numpy shim")

    jax.grad = jax.jvp = jax.value_and_grad = _no_autodiff
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jnp
    sys.modules["jax.scipy"] = jsp
    sys.modules["jax.scipy.linalg"] = scipy.linalg
    for name in [
        "cvxpy", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "mpl_toolkits",
        "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator", "pytenet", "qutip",
        "qutip.states",
    ]:
        if name not in sys.modules:
            sys.modules[name] = _empty(name)
    # names the reference pulls out of those packages with `from x import y`
    sys.modules["matplotlib.ticker"].MaxNLocator = object
    sys.modules["mpl_toolkits.axes_grid1.inset_locator"].inset_axes = None
    sys.modules["mpl_toolkits.axes_grid1.inset_locator"].mark_inset = None


def load_reference():
    """Return the reference's modules as a namespace (ref.transformations, ref.optimization, ...)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import importlib

    ns = types.SimpleNamespace()
    for mod in ("transformations", "optimization", "stiefel", "trust_region_rcopt"):
        setattr(ns, mod, importlib.import_module(f"opentn.{mod}"))
    return ns
