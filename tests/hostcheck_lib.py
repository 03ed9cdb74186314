"""ctypes handle of sp1_b200/libsp1b200_hostcheck.so: the product's DEVICE arithmetic sources (kb31.cuh, poseidon2.cuh, zc_lower.hpp) compiled
for the host (csrc/hostcheck.cu).  Test support: a separate library, not part of the product's libsp1b200.so."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sp1_b200", "libsp1b200_hostcheck.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) and os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "sp1_b200", "csrc"), "-j8", "all"])
        _lib = C.CDLL(SO)
    return _lib
