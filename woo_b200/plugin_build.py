"""Builds woo_b200/_woo*.so: the pybind11 host layer (GpuDemLoop + reference-named object model) over libwdem.so's C-ABI."""
import os, subprocess, sys, sysconfig
from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "woo_plugin.cpp")
OUT = os.path.join(HERE, "_woo" + sysconfig.get_config_var("EXT_SUFFIX"))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return os.path.getmtime(SRC) > t or os.path.getmtime(os.path.join(HERE, "..", "include", "wdem.h")) > t


def build(force=False, verbose=False):
    _build.build()
    if not force and not needs_build():
        return OUT
    import pybind11
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           SRC, "-o", OUT, "-L", HERE, "-l:libwdem.so", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
