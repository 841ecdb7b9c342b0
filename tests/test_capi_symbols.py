"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/afesp_gpu.h declares, and its
host-only entry points behave (no GPU needed, no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "afesp_gpu.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(afesp_gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes

    from afesp_b200 import capi

    lib = capi.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert isinstance(getattr(lib, n), ctypes._CFuncPtr)
    assert set(capi.SIGNATURES) | {"afesp_gpu_last_error"} == set(names)


def test_open_fails_loudly_without_a_gpu():
    import torch

    from afesp_b200 import AfespError, AfespGpu

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(AfespError, match="no CUDA device|no CPU fallback"):
        AfespGpu(0)


def test_triples_partition_covers_every_triple_once():
    from afesp_b200 import AfespGpu

    for o in [1, 2, 5, 7, 10, 20]:
        total = o * (o + 1) * (o + 2) // 6
        for nranks in [1, 2, 3, 8]:
            counts = AfespGpu.triples_partition(o, nranks)
            assert sum(counts) == total and max(counts) - min(counts) <= 1
        strict = o * (o - 1) * (o - 2) // 6
        assert sum(AfespGpu.triples_partition(o, 4, True, True)) == strict
        assert sum(AfespGpu.triples_partition(o, 4, False, False)) == o ** 3


def test_library_depends_on_no_vendor_math_or_tensor_library():
    """north_star: no cuBLAS / cuTENSOR / OpenACC runtime behind the boundary -- the contractions are the library's own
    kernels.  The shared object may need only libc / libstdc++ / libm / libgcc (the CUDA runtime is linked statically, NCCL
    is dlopen'ed by name on first use of afesp_gpu_comm_*)."""
    import subprocess

    from afesp_b200 import capi

    needed = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    libs = re.findall(r"\(NEEDED\)\s+Shared library: \[([^\]]+)\]", needed)
    assert libs, needed
    for lib in libs:
        assert re.match(r"(lib(c|m|dl|rt|pthread|stdc\+\+|gcc_s)\.so|ld-linux)", lib), lib
    undefined = subprocess.run(["nm", "-D", "--undefined-only", capi.LIB_PATH], capture_output=True, text=True).stdout.lower()
    for vendor in ("cublas", "cutensor", "cusolver", "cudnn", "acc_", "pgi_", "nvhpc"):
        assert vendor not in undefined, vendor


def test_product_code_never_touches_the_oracle():
    """The oracle (oracle/, tests/_oracle_engine.py, tests/_double/) is test infrastructure: nothing under afesp_b200/, host/,
    include/ or shim/ may import, link or mention it, and the Makefiles of the product build nothing from it."""
    product = []
    for top in ("afesp_b200", "host", "include", "shim"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if "__pycache__" in dirpath or os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
                continue
            product += [os.path.join(dirpath, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h",
                                                                                ".f90", ".inc")) or f == "Makefile"]
    assert len(product) > 25
    for path in product:
        text = open(path, errors="replace").read()
        for needle in ("import oracle", "from oracle", "oracle/", "_oracle_engine", "afesp_gpu_double", "cpu_port", "cpu_ccsd"):
            assert needle not in text, (path, needle)
