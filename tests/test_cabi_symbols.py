"""The C-ABI library loads on a machine without a GPU and exports every function include/wdem.h declares (no compute calls here);
creating an engine without a usable sm_100 device fails loudly — there is no CPU fallback."""
import ctypes as C
import os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "wdem.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wdem_[a-z0-9_]+)\s*\(", src)))


def test_header_and_library_agree():
    from woo_b200 import capi
    names = declared_functions()
    assert len(names) >= 40 and "wdem_step" in names and "wdem_upload_contacts" in names
    L = C.CDLL(capi.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/wdem.h but not exported by libwdem.so: %s" % missing
    assert sorted(capi.EXPORTS) == names, "woo_b200/capi.py EXPORTS and include/wdem.h differ: %s" % sorted(set(capi.EXPORTS) ^ set(names))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="this box has a GPU")
def test_create_fails_loudly_without_a_gpu():
    from woo_b200 import capi
    L = capi.load()
    cfg = capi.WdemConfig(); L.wdem_default_config(C.byref(cfg))
    h = C.c_void_p()
    rc = L.wdem_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in L.wdem_last_error(None)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/wdem.h must compile as C99 (no C++ or torch types in the signatures)."""
    import shutil, subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "include", "wdem.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
