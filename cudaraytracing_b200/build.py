"""In-tree build of libcrt.so and the `crt` CLI (nvcc, sm_100a only)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    """Runs `make` in cudaraytracing_b200/ (cross-compiles without a GPU). Returns the .so path."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j4", "all"], stdout=out)
    so = os.path.join(_HERE, "libcrt.so")
    if not os.path.exists(so):
        raise RuntimeError("build did not produce " + so)
    return so
