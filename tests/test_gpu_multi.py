"""Multi-GPU parity through the C ABI (needs >= 2 devices; skipped on a single-GPU box).

tools/dist_check.py is launched under torchrun with 2 ranks: the rank-sharded AO->MO / CCSD / (T) path must reproduce
the replicated path on the same inputs (packed MO integrals 1e-12, CCSD iteration energies 1e-10 Eh, converged T1/T2
1e-9, (T) sums 1e-10) and leave bit-identical amplitudes on every rank."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("extra", [[], ["--spinorb", "--nbf", "40", "--nocc", "5"]], ids=["spatial", "spinorb"])
def test_sharded_chain_matches_replicated_on_two_gpus(extra):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + (os.getpid() % 300) + (7 if extra else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")] + extra
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["pass"] and line["ranks_bit_identical"], line
