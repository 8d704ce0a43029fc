"""`torch.ops.tcct_b200.*`: torch custom-op registration (TORCH_LIBRARY) over the C ABI -- csrc_torch/torch_ops.cpp.

    python -m tcct_b200.torch_ops        builds tcct_b200/lib/libtcct_b200_torch.so in-tree (g++ against the torch headers)
    tcct_b200.torch_ops.load()           torch.ops.load_library(...) -> torch.ops.tcct_b200.dice_multi_fwd(...) etc.

The product path binds the same C ABI through ctypes (tcct_b200/_lib.py: no compile-time torch dependency, identical entry points);
this module is the operator-surface binding the north-star names: schema strings, CUDA dispatch key, errors as RuntimeError."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_torch", "torch_ops.cpp")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libtcct_b200_torch.so")


def build(force=False, verbose=False):
    import torch
    from torch.utils import cpp_extension as ce
    from .build import build as build_kernels
    build_kernels()
    stamp = os.path.join(LIBDIR, "torch_ops.stamp")
    sig = "%s|%s|%s" % (os.path.getmtime(SRC), os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "tcct_b200.h")), torch.__version__)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == sig:
        return LIB
    inc = ["-I" + p for p in ce.include_paths(device_type="cuda")] + ["-I" + os.path.join(os.path.dirname(HERE), "include"),
                                                                      "-I" + sysconfig.get_paths()["include"]]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)] + inc + [
        SRC, "-o", LIB, "-L" + LIBDIR, "-ltcct_b200", "-L" + torch_lib, "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
        "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(sig)
    return LIB


def load():
    """Register the ops with this process's torch (idempotent); raises if the shim was not built."""
    import torch
    if not os.path.exists(LIB):
        raise ImportError("tcct_b200: %s is missing - build it with `python -m tcct_b200.torch_ops`" % LIB)
    if not hasattr(torch.ops.tcct_b200, "dice_multi_fwd"):
        import ctypes
        ctypes.CDLL(os.path.join(LIBDIR, "libtcct_b200.so"), mode=ctypes.RTLD_GLOBAL)
        torch.ops.load_library(LIB)
    return torch.ops.tcct_b200


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
