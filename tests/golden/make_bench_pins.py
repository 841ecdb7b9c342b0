"""Pinned energies for bench.py's `parity` block (tests/golden/bench_pinned.json).

    python tests/golden/make_bench_pins.py cpu                 # CPU-only part (here): MP2 from the NumPy oracle, E_CCSD after
                                                               # the first iteration from the CPU port of the reference
    python tests/golden/make_bench_pins.py gpu <bench.json>    # merge the per-step (E_CCSD, e_T) trajectory of a SINGLE-GPU
                                                               # (replicated, unsharded) `bench.py --trajectory` line

The N = 2/4/8 lines of the scaling run are then checked against the single-GPU trajectory (sharded vs replicated), and
every line against the CPU values, at 1e-9 Eh."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PATH = os.path.join(ROOT, "tests", "golden", "bench_pinned.json")


def load():
    return json.load(open(PATH)) if os.path.exists(PATH) else {}


def cpu(nbf=200, nocc=20):
    from oracle import afesp_oracle as orc
    from oracle import cpu_port

    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    mo, Cmo, eps = cpu_port.synthetic_mo_integrals(nbf, nocc)
    V = cpu_port.slices(lib, mo, nbf, nocc)
    o, v = nocc, nbf - nocc
    D1, D2 = orc.denominators(eps, o)
    t2 = np.asfortranarray(V["v_oovv"] / D2)
    t1 = np.zeros((o, v), order="F")
    voovv = np.asarray(V["v_oovv"])
    # MP2 (src/mp2.f90:418-438) = the MP1 CC energy of the spin-free formulation (src/ccsd.f90:1774 with t1 = 0)
    e_mp2 = float(np.sum(voovv * (2.0 * voovv - voovv.transpose(0, 1, 3, 2)) / D2))
    t1n, t2n, _, _ = cpu_port.ccsd_iter(lib, V, eps, t1, t2)
    e1 = orc.restricted_energy(np.asarray(t1n), np.asarray(t2n), voovv)
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key].update({"e_mp2_oracle": e_mp2, "e_ccsd_iter1_cpu_port": e1,
                      "cpu_source": "tests/golden/make_bench_pins.py cpu (NumPy MP2 expression on the factored-form MO "
                                    "integrals; first CCSD iteration through oracle/cpu_ccsd.c)"})
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, pins[key]["e_mp2_oracle"], pins[key]["e_ccsd_iter1_cpu_port"])


def gpu(path):
    pins = load()
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert d["n_gpus"] == 1, "pin the trajectory from a single-GPU run"
    for blk, traj in ((d, d.get("trajectory")), (d.get("target_config") or {}, (d.get("target_config") or {}).get("trajectory"))):
        if not traj:
            continue
        c = blk["config"]
        key = f"nbf{c['nbf']}_nocc{c['nocc']}"
        pins.setdefault(key, {})
        old = pins[key].get("steps", [])
        if len(traj) >= len(old):
            pins[key].update({"e_mp2": blk["energies"]["e_mp2"], "e_mp1": blk["energies"]["e_mp1"], "steps": traj,
                              "source": "single-GPU (replicated) bench.py --trajectory run, merged by make_bench_pins.py gpu"})
        print(key, len(traj), "steps pinned")
    json.dump(pins, open(PATH, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "cpu":
        cpu(*[int(x) for x in sys.argv[2:4]])
    else:
        gpu(sys.argv[2])
