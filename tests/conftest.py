import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_runs():
    """Cache of full oracle runs on the sample molecules (each a few seconds)."""
    from oracle import afesp_oracle as orc
    from tests._fixtures import load_system

    cache = {}

    def get(name, calc_type=None, **kw):
        key = (name, calc_type, tuple(sorted(kw.items())))
        if key not in cache:
            s = load_system(name, calc_type)
            cache[key] = (s, orc.run(s, **kw))
        return cache[key]

    return get
