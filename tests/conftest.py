import os
import subprocess
import sys
import sysconfig

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


@pytest.fixture(scope="session")
def double_env(tmp_path_factory):
    """TEST DOUBLE of the C ABI for the C++ host (tests/_double/afesp_gpu_double.c: every call goes to the NumPy oracle through an
    embedded interpreter): built into a temporary directory (never in-tree); returns the environment that LD_PRELOADs it
    in front of libafesp_gpu.so.  See tests/test_els_host_flow.py."""
    inc = sysconfig.get_config_var("INCLUDEPY")
    libdir = sysconfig.get_config_var("LIBDIR")
    ver = sysconfig.get_config_var("LDVERSION")
    if not (inc and os.path.exists(os.path.join(inc, "Python.h")) and sysconfig.get_config_var("Py_ENABLE_SHARED")):
        pytest.skip("no embeddable Python (Python.h / libpython) in this environment")
    out = tmp_path_factory.mktemp("double") / "afesp_gpu_test_double.so"
    cmd = ["gcc", "-O1", "-shared", "-fPIC", "-I", inc, os.path.join(ROOT, "tests", "_double", "afesp_gpu_double.c"), "-o",
           str(out), "-L", libdir, f"-lpython{ver}", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cannot build the test double: " + r.stderr[-400:])
    env = dict(os.environ)
    env["LD_PRELOAD"] = str(out)
    env["PYTHONPATH"] = os.pathsep.join([ROOT] + [p for p in sys.path if p.endswith("site-packages")])
    env.pop("AFESP_GPU_OPTIONS", None)
    return env
