"""The reference's CPU path timed on the host cores (TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `cpu_baseline`
leg and `--impl reference` arm; never imported by the product).

Run as a subprocess -- `python -m oracle.cpu_reference --nbf 200 --nocc 20 --samples 5 --budget 12` -- so that the thread
environment is its own: OMP_NUM_THREADS = every core this process may use (torchrun exports OMP_NUM_THREADS=1),
OMP_WAIT_POLICY=passive (libgomp's spinning workers and OpenBLAS's pthread pool otherwise fight for the same cores
between the reference's alternating OpenMP loops and dgemm calls: 2-10x slower iterations on the sample molecules), and a
libgomp that is initialised with those settings (torch preloads its own copy in the bench process).  Prints one JSON object.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuReference:
    """The reference's CPU path at shape (nbf, nocc) on every host core this process may use, whatever
    OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1): OpenMP loops and OpenBLAS dgemms both.

    Inputs are the same synthetic system the GPU arm runs (MO integrals from the factored form, MP1 amplitudes).  One
    sample = (a) ONE COMPLETE spin-free CCSD iteration as src/ccsd.f90:1040-1312 + 1538-1732 issues it -- every dgemm
    (ladder included) and every reshape in full; the two naive o^3v^3 loop nests (:1170-1182, :1680-1695) in full when
    the time budget allows, else on a slab of their outermost index (>= 1/8), scaled linearly and stated; (b) the
    reference (T) loop (:2152-2233) on one COMPLETE ordered triple per thread, scaled by o^3 / triples; (c) the AO->MO
    quarter transforms (src/mp2.f90:321-387) on a slab of the outermost index (reported as ao2mo_s, not part of `value`,
    which is CCSD iteration + (T) like the GPU arm's)."""

    def __init__(self, nbf, nocc):
        from oracle import cpu_port

        self.cp = cpu_port
        self.lib = cpu_port.load()
        self.threads = cpu_port.set_threads(self.lib, host_cores())
        self.n, self.o, self.v = nbf, nocc, nbf - nocc
        n, o, v = self.n, self.o, self.v
        t0 = time.perf_counter()
        self.mo, self.Cmo, self.eps = cpu_port.synthetic_mo_integrals(n, o)
        try:
            avail = int([ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:
            avail = 32 << 30
        dense = 8 * v ** 4
        self.ncol_d = v if dense <= min(12 << 30, avail // 4) else max(1, int((4 << 30) // (8 * v ** 3)))
        self.V = cpu_port.slices(self.lib, self.mo, n, o, vvvv_cols=(0, self.ncol_d))
        eo, ev = self.eps[:o], self.eps[o:]
        D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
        self.t2 = np.asfortranarray(self.V["v_oovv"] / D2)
        self.t1 = np.zeros((o, v), order="F")
        rng = np.random.default_rng(7)
        self.triples = [tuple(int(x) for x in rng.integers(0, o, 3)) for _ in range(self.threads)]
        self.frac = 1.0 / 8.0      # slab fraction of the two naive loop nests; raised after the first (calibration) sample
        self.setup_s = time.perf_counter() - t0
        self.eri_ao = None

    def sample(self, budget_s):
        cp, lib, n, o, v = self.cp, self.lib, self.n, self.o, self.v
        # (b) (T): one complete ordered triple per thread
        t0 = time.perf_counter()
        _, dt_t = cp.triples(lib, self.t1, self.t2, self.V["v_oovv"], self.V["v_vvov"], self.V["v_oovo"], self.eps,
                             self.triples, True, False)
        t_T = dt_t * (o ** 3) / len(self.triples)
        # (a) complete CCSD iteration; slab fraction from what is left of the budget
        left = max(0.0, budget_s - (time.perf_counter() - t0))
        if hasattr(self, "_naive_full_s"):
            self.frac = float(min(1.0, max(1.0 / 8.0, (left - self._rest_s) / max(self._naive_full_s, 1e-9))))
        bmax = max(1, min(v, int(round(self.frac * v))))
        amax = bmax
        _, _, parts, wall = cp.ccsd_iter(lib, self.V, self.eps, self.t1, self.t2, ring_bmax=bmax, iovov_amax=amax)
        parts = [float(x) for x in parts]
        ladder = parts[4] * v / self.ncol_d
        iovov, ring = parts[1] * v / amax, parts[5] * v / bmax
        rest = parts[0] + parts[2] + parts[3] + parts[6] + parts[7]
        self._naive_full_s, self._rest_s = iovov + ring, rest + parts[4]
        t_iter = rest + ladder + iovov + ring
        return {"ccsd_s_per_iter": t_iter, "t_wall_s": t_T, "value": t_iter + t_T,
                "parts": {"ladder_dgemm": ladder, "ring_loop": ring, "I_ovov_loop": iovov, "other": rest,
                          "triples_sample_s": dt_t, "ccsd_sample_s": wall},
                "slab": {"ring_b": [bmax, v], "iovov_a": [amax, v], "ladder_cols": [self.ncol_d * v, v * v],
                         "triples": [len(self.triples), o ** 3]}}

    def ao2mo(self, budget_s=6.0):
        """AO->MO of src/mp2.f90:321-387 on an l-slab sized for ~budget_s, scaled by n / slab."""
        from afesp_b200 import synthetic

        n = self.n
        if self.eri_ao is None:
            if n > 240:
                return None   # the packed AO integrals alone are 25.7 GB at nbf=400; not sampled
            self.eri_ao, _, _ = synthetic.make(n, self.o)
        _, t1 = self.cp.ao2mo(self.lib, self.eri_ao, self.Cmo, lmax=1, smax=1, want_result=False)
        per_l = float(np.sum(t1[:4]))
        lmax = int(max(1, min(n, budget_s / max(per_l, 1e-6))))
        _, t = self.cp.ao2mo(self.lib, self.eri_ao, self.Cmo, lmax=lmax, smax=lmax, want_result=False)
        return {"ao2mo_s": float(np.sum(t[:4])) * n / lmax, "slab": [lmax, n], "sample_s": float(np.sum(t[:4]))}

    def describe(self, s):
        sl = s["slab"]
        f = lambda a: "in full" if a[0] >= a[1] else f"on {a[0]} of {a[1]} (x{a[1] / a[0]:.2f})"
        return (f"one complete spin-free CCSD iteration as src/ccsd.f90:1040-1312,1538-1732 issues it (all dgemms and "
                f"reshapes in full; ladder dgemm :1669 columns {f(sl['ladder_cols'])}; ring loop nest :1680-1695 outer index "
                f"{f(sl['ring_b'])}; I_ovov loop nest :1170-1182 outer index {f(sl['iovov_a'])}) = "
                f"{s['ccsd_s_per_iter']:.2f} s/iter [{s['parts']['ccsd_sample_s']:.1f} s measured] + reference (T) loop "
                f":2152-2233 on {sl['triples'][0]} complete ordered triples (one per thread) of {sl['triples'][1]} "
                f"(x{sl['triples'][1] / sl['triples'][0]:.0f}) = {s['t_wall_s']:.0f} s [{s['parts']['triples_sample_s']:.1f} s measured]")




def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nbf", type=int, default=200)
    ap.add_argument("--nocc", type=int, default=20)
    ap.add_argument("--samples", type=int, default=3)
    ap.add_argument("--budget", type=float, default=20.0, help="CPU seconds per sample")
    ap.add_argument("--ao2mo-budget", type=float, default=6.0)
    a = ap.parse_args()
    ref = CpuReference(a.nbf, a.nocc)
    print(f"[cpu_reference] {ref.threads} threads, inputs in {ref.setup_s:.1f}s, OpenBLAS {ref.lib._blas_path}", file=sys.stderr, flush=True)
    samples = []
    for s in range(a.samples):
        smp = ref.sample(a.budget)
        smp["describe"] = ref.describe(smp)
        samples.append(smp)
        print(f"[cpu_reference] sample {s}: value {smp['value']:.1f}s (ccsd {smp['ccsd_s_per_iter']:.2f}, T {smp['t_wall_s']:.0f}) "
              f"slab {smp['slab']}", file=sys.stderr, flush=True)
    ao = ref.ao2mo(a.ao2mo_budget) if a.ao2mo_budget > 0 else None
    print(json.dumps({"threads": ref.threads, "setup_s": ref.setup_s, "blas": ref.lib._blas_path, "samples": samples,
                      "ao2mo": ao, "env": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_WAIT_POLICY",
                                                                         "OPENBLAS_NUM_THREADS")}}))


def run_subprocess(nbf, nocc, samples, budget, ao2mo_budget=6.0, timeout=1500):
    """Spawn this module with a clean thread environment; returns the parsed JSON object."""
    cores = host_cores()
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(cores), "OPENBLAS_NUM_THREADS": str(cores), "OMP_WAIT_POLICY": "passive",
                "OMP_DYNAMIC": "false", "MKL_NUM_THREADS": str(cores)})
    for k in ("OMP_PROC_BIND", "OMP_PLACES", "GOMP_CPU_AFFINITY"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "oracle.cpu_reference", "--nbf", str(nbf), "--nocc", str(nocc), "--samples",
                        str(samples), "--budget", str(budget), "--ao2mo-budget", str(ao2mo_budget)], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=sys.stderr, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"oracle.cpu_reference exited with status {r.returncode}")
    return json.loads(r.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    main()
