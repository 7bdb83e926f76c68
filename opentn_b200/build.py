"""Build libotn.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The binary carries the sha256 of its sources (`otn_build_id()`, compiled in with -DOTN_SOURCE_HASH).  `build()` recompiles
whenever that hash differs from the sources on disk, so "the shipped .so corresponds to this source tree" is a checked fact,
not an mtime heuristic (git-ignored binaries travel to the GPU box with the snapshot)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libotn.so")
SOURCES = ["otn_api.cu"]
HEADER = os.path.join(HERE, "..", "include", "otn.h")
MARKER = b"OTN_SOURCE_HASH="
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
              "--cudart=shared"]


def source_files():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl", ".h")))
    return files + [HEADER]


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for path in source_files():
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def binary_hash(path: str = LIB):
    """The source hash compiled into `path` (read from the file, nothing is loaded); None if absent."""
    if not os.path.exists(path):
        return None
    with open(path, "rb") as fh:
        blob = fh.read()
    i = blob.find(MARKER)
    if i < 0:
        return None
    j = blob.find(b"\0", i)
    return blob[i + len(MARKER):j].decode(errors="replace")


def needs_build() -> bool:
    return binary_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    digest = source_hash()
    cmd = ["nvcc"] + NVCC_FLAGS + [f'-DOTN_SOURCE_HASH="{digest}"', "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libotn.so")
    if verbose:
        print(res.stdout + res.stderr)
    if binary_hash() != digest:
        raise RuntimeError("libotn.so does not carry the hash of the sources it was just built from")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print("source hash", source_hash())
