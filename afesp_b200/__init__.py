"""afesp_b200 -- B200-native coupled-cluster engine behind AFESP's host program.

The product is the CUDA library afesp_b200/lib/libafesp_gpu.so (sources in afesp_b200/csrc, C ABI in
include/afesp_gpu.h).  This package is the thin host side: ctypes bindings (capi) and a Python mirror of the
reference's program flow (host) used by the tests and the benchmark.  There is no CPU fallback: importing capi
works anywhere (so the exported symbols can be checked), but opening a handle needs a Blackwell GPU.
"""
from .capi import AfespGpu, AfespError, load_library, LIB_PATH  # noqa: F401
