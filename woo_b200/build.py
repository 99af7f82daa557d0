"""Builds libwdem.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.  --fmad=false is part of the
numerical contract (bit parity with the reference's non-FMA x86-64 arithmetic)."""
import os, subprocess, shutil, sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwdem.so")
SOURCES = ["wdem_engine.cu"]
DEPS = ["wdem_engine.cu", "wdem_kernels.cuh", "wdem_domain.cuh", "wdem_comm.h", "wdem_math.cuh", os.path.join("..", "..", "include", "wdem.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-Wno-deprecated-declarations",
              "-diag-suppress", "177"]


def nccl_flags():
    """NCCL for the slab decomposition (halo exchange / migration).  Prefer the copy PyTorch ships (the one torch.distributed loads in the
    same process, so that only one NCCL lives in a rank); fall back to the system library."""
    import importlib.util
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec and spec.submodule_search_locations:
        root = list(spec.submodule_search_locations)[0]
        inc, lib = os.path.join(root, "include"), os.path.join(root, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
            return ["-I", inc, "-L", lib, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + lib]
    return ["-lnccl"]


def nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the woo_b200 engine cannot be built (there is no CPU fallback)")
    return p


# variant "alt": every 3-term reduction order flipped (wdem_math.cuh switches) -> libwdem_alt.so, loaded only by tests/test_order_sensitivity.py
VARIANTS = {"": [], "alt": ["-DWD_SUM3_ORDER=1", "-DWD_TRACE_ORDER=0", "-DWD_MATPROD_ORDER=2"]}


def lib_path(variant=""):
    return LIB if not variant else os.path.join(HERE, "libwdem_%s.so" % variant)


def needs_build(variant=""):
    lib = lib_path(variant)
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, extra=(), variant=""):
    lib = lib_path(variant)
    if not force and not needs_build(variant):
        return lib
    cmd = [nvcc()] + NVCC_FLAGS + VARIANTS[variant] + list(extra) + ["-o", lib] + [os.path.join(CSRC, s) for s in SOURCES] + nccl_flags()
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    cmd += ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, env=env)
    return lib


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else [])
    if "--all" in sys.argv:
        build(force="-f" in sys.argv, verbose=True, variant="alt")
