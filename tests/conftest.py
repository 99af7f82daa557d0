import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a CUDA device skips the GPU tests instead of failing them; with `-m gpu` they are still
    selected (and fail loudly in wdem_create: there is no CPU fallback to hide behind) only if a device exists."""
    if _cuda_available():
        return
    import pytest
    skip = pytest.mark.skip(reason="no CUDA device here (the engine has no CPU fallback); run with -m gpu on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
