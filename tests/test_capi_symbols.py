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
