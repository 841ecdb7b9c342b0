"""Pinned energies for bench.py's `parity` block (tests/golden/bench_pinned.json).

    python tests/golden/make_bench_pins.py cpu                 # CPU-only part (here): MP2 from the NumPy oracle, E_CCSD after
                                                               # the first iteration from the CPU port of the reference
    python tests/golden/make_bench_pins.py cpu_T               # CPU-only, ~15 min: e_T of the first step at nbf=200 through the
                                                               # oracle's BLAS orbit form of the [T] accumulator
    python tests/golden/make_bench_pins.py cpu_mp2 400 40      # CPU-only: MP2 of the target shape from the (ia|jb) block
    python tests/golden/make_bench_pins.py cpu_T_target 400 40      # CPU-only, ~3 h: e_T of the first step at the target shape
    python tests/golden/make_bench_pins.py cpu_ccsd_iter1 400 40    # CPU-only, ~20 min: E_CCSD after the first iteration at the
                                                               # target shape, from the factored integrals
    python tests/golden/make_bench_pins.py cpu_mp1_triples 400 40   # CPU-only: (T) contributions of six single triples on the
                                                               # MP1 amplitudes at the target shape, from the factored integrals
    python tests/golden/make_bench_pins.py cpu_traj 200 20 3   # CPU-only, ~17 min: the first three bench steps (CCSD iteration,
                                                               # DIIS, [T] on the extrapolated amplitudes) entirely on the CPU
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


def cpu_mp2(nbf=400, nocc=40):
    """MP2 of a shape whose packed MO integrals are too large to build on the host (nbf=400: 25.7 GB): only the (ia|jb) block,
    from the factored form -- B_mo(i,a;P) = sum_mn C(i,m) B(mn,P) C(a,n), (ia|jb) = sum_P B_mo(ia,P) B_mo(jb,P) -- then the
    expression of src/mp2.f90:418-438."""
    e_mp2 = mp2_from_factors(nbf, nocc)
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["e_mp2_oracle"] = e_mp2
    pins[key]["cpu_source"] = ("tests/golden/make_bench_pins.py cpu_mp2 (NumPy: (ia|jb) block from the factored form of the "
                               "synthetic integrals, MP2 expression of src/mp2.f90:418-438)")
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, "e_mp2_oracle", e_mp2, "GPU-pinned:", pins[key].get("e_mp2"))


def mp2_from_factors(nbf, nocc):
    from afesp_b200 import synthetic

    B, Cmo, eps = synthetic.make_factors(nbf, nocc)
    n, o, v = nbf, nocc, nbf - nocc
    ii, jj = np.tril_indices(n)
    Bov = np.empty((o * v, B.shape[1]))
    full = np.empty((n, n))
    for P in range(B.shape[1]):
        full[ii, jj] = B[:, P]
        full[jj, ii] = B[:, P]
        Bov[:, P] = (Cmo[:o] @ full @ Cmo[o:].T).ravel()       # (i, a) row-major
    iajb = (Bov @ Bov.T).reshape(o, v, o, v)
    D = eps[:o, None, None, None] + eps[None, None, :o, None] - eps[None, o:, None, None] - eps[None, None, None, o:]
    return float(np.sum(iajb * (2.0 * iajb - iajb.transpose(0, 3, 2, 1)) / D))


def unique_triples(o):
    """Enumeration order of the library's unique (i <= j <= k) list (afesp_b200/csrc/triples.cu: my_triples): with
    afesp_gpu_set_partition(h, r, ntriples) "rank" r owns exactly triple number r."""
    return [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]


def mp1_triples_from_factors(nbf, nocc, picks):
    """(T) contributions of single unique triples on the MP1 amplitudes (t1 = 0, t2 = <ij|ab> / D: the state right after
    afesp_gpu_ccsd_init), everything from the factored form of the synthetic integrals -- so it also works at nbf=400, where
    neither the packed MO integrals (25.7 GB) nor the v_vvov slice (15 GB) are built on the host:
        t2(i,j,a,b) = (ia|jb) / D,   v_vovv(d,k,b,c) = (ck|bd),   v_oovo(r,q,c,l) = (rc|ql),   (pq|rs) = sum_P B(pq,P) B(rs,P).
    Returns the list of e_T orbit contributions (oracle: triples_bracket_T_orbit_form)."""
    from afesp_b200 import synthetic
    from oracle import afesp_oracle as orc

    B, Cmo, eps = synthetic.make_factors(nbf, nocc)
    n, o, v = nbf, nocc, nbf - nocc
    naux = B.shape[1]
    ii, jj = np.tril_indices(n)
    Boo, Bov, Bvv = np.empty((naux, o, o)), np.empty((naux, o, v)), np.empty((naux, v, v))
    full = np.empty((n, n))
    for P in range(naux):
        full[ii, jj] = B[:, P]
        full[jj, ii] = B[:, P]
        m = Cmo @ full @ Cmo.T
        Boo[P], Bov[P], Bvv[P] = m[:o, :o], m[:o, o:], m[o:, o:]
    M = Bov.reshape(naux, o * v)
    iajb = (M.T @ M).reshape(o, v, o, v)
    D1, D2 = orc.denominators(eps, o)
    t2 = np.ascontiguousarray(iajb.transpose(0, 2, 1, 3) / D2)
    del iajb
    Bvv2 = Bvv.reshape(naux, v * v)

    def iv_block(k):      # [d, (b,c)] = (ck|bd) = sum_P Bov[P,k,c] Bvv[P,b,d]
        x = (Bvv2.T @ Bov[:, k, :]).reshape(v, v, v)        # [b, d, c]
        return np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(v, v * v)

    def io_block(r, q):   # [c, l] = (rc|ql) = sum_P Bov[P,r,c] Boo[P,q,l]
        return Bov[:, r, :].T @ Boo[:, q, :]

    return [orc.triples_bracket_T_orbit_form(t2, None, None, eps, triples=[t], iv_block=iv_block, io_block=io_block)
            for t in picks], eps


def ccsd_iter1_from_factors(nbf, nocc, ladder_block=8, log=None, return_state=False):
    """E_CCSD and sum (dT2)^2 after the FIRST spin-free CCSD iteration (input: t1 = 0, t2 = MP1), from the factored form of the
    synthetic integrals, for shapes whose dense v_vvvv slice (nbf=400: 134 GB) rules out both the oracle's general functions
    and the CPU port.  It is oracle.restricted_intermediates + restricted_amplitudes (src/ccsd.f90:1040-1312, 1538-1732) with
    every term that carries a factor t1 dropped, the ladder <ef|ab> = sum_P B(ea,P) B(fb,P) built in slabs of b, and the one
    v_vvov contraction of the T1 equation (:1618-1630) taken through the factors.  Checked against the general oracle
    functions at small shapes (tests/test_cpu_port.py) and against the CPU port at nbf=200 (bench_pinned.json)."""
    import time

    from afesp_b200 import synthetic
    from oracle import afesp_oracle as orc

    t0 = time.perf_counter()
    say = log or (lambda *a: None)
    B, Cmo, eps = synthetic.make_factors(nbf, nocc)
    n, o, v = nbf, nocc, nbf - nocc
    naux = B.shape[1]
    ii, jj = np.tril_indices(n)
    Boo, Bov, Bvv = np.empty((naux, o, o)), np.empty((naux, o, v)), np.empty((naux, v, v))
    full = np.empty((n, n))
    for P in range(naux):
        full[ii, jj] = B[:, P]
        full[jj, ii] = B[:, P]
        m = Cmo @ full @ Cmo.T
        Boo[P], Bov[P], Bvv[P] = m[:o, :o], m[:o, o:], m[o:, o:]
    del B, full
    ein = lambda *a: np.einsum(*a, optimize=True)
    # physicist slices <pq|rs> = (pr|qs)
    v_oovv = ein("Pia,Pjb->ijab", Bov, Bov)
    v_ovov = ein("Pij,Pab->iajb", Boo, Bvv)
    v_oovo = ein("Pia,Pjl->ijal", Bov, Boo)
    v_oooo = ein("Pik,Pjl->ijkl", Boo, Boo)
    D1, D2 = orc.denominators(eps, o)
    t2 = v_oovv / D2
    asym = 2.0 * t2 - t2.transpose(1, 0, 2, 3)
    A = 2.0 * v_oovv - v_oovv.transpose(0, 1, 3, 2)
    say(f"slices + MP1 amplitudes {time.perf_counter() - t0:.0f} s")
    # intermediates with t1 = 0 (c = t2, I_vo = 0)
    I_vv = -ein("mneb,mnea->ba", A, t2)
    I_oo = ein("jmfe,mief->ji", asym, v_oovv)
    I_oooo = v_oooo + ein("klef,ijef->klij", t2, v_oovv)
    I_ovov = v_ovov - 0.5 * ein("mibe,mjae->jbia", v_oovv, t2)
    I_voov = 0.5 * ein("imbe,mjea->bjia", A, t2) - 0.5 * ein("imbe,mjae->bjia", v_oovv, t2) + v_oovv.transpose(3, 0, 1, 2)
    say(f"intermediates {time.perf_counter() - t0:.0f} s")
    # T1 equation: what survives t1 = 0 is  - v_oovo . asym  +  v_vvov . asym   (:1606-1630)
    r1 = -ein("mien,mnea->ia", v_oovo, asym)
    Y = ein("Pme,mief->Pif", Bov, asym)                      # v_vvov(e,f,m,a) = (em|fa) = sum_P Bov(P,m,e) Bvv(P,f,a)
    r1 = r1 + ein("Pif,Pfa->ia", Y, Bvv)
    # T2 equation
    X = ein("ijae,eb->ijab", t2, I_vv) - ein("miba,jm->ijab", t2, I_oo)
    X = X + 0.5 * ein("ijmn,mnab->ijab", I_oooo, t2)
    X = X - ein("mjae,iemb->ijab", t2, I_ovov) - ein("iema,mjeb->ijab", I_ovov, t2) + ein("miea,ejmb->ijab", asym, I_voov)
    say(f"ring terms {time.perf_counter() - t0:.0f} s")
    c2 = np.ascontiguousarray(t2.reshape(o * o, v * v))       # c(ij, ef)
    Bvv_ea = np.ascontiguousarray(Bvv.transpose(1, 2, 0).reshape(v * v, naux))      # [(e,a), P]
    lad = np.empty((o * o, v, v))                             # [ij, a, b]
    for b0 in range(0, v, ladder_block):
        b1 = min(v, b0 + ladder_block)
        W = (Bvv_ea @ Bvv[:, :, b0:b1].reshape(naux, -1)).reshape(v, v, v, b1 - b0)          # [e, a, f, b] = <ef|ab>
        W = np.ascontiguousarray(W.transpose(0, 2, 1, 3)).reshape(v * v, v * (b1 - b0))      # [(e,f), (a,b)]
        lad[:, :, b0:b1] = (c2 @ W).reshape(o * o, v, b1 - b0)
    X = X + 0.5 * lad.reshape(o, o, v, v)
    del lad, W
    say(f"ladder {time.perf_counter() - t0:.0f} s")
    X = X + X.transpose(1, 0, 3, 2) + v_oovv
    t1n, t2n = r1 / D1, X / D2
    e = orc.restricted_energy(t1n, t2n, v_oovv)
    rms = float(np.sum((t2n - t2) ** 2))
    e_mp2 = orc.restricted_energy(np.zeros_like(t1n), t2, v_oovv)
    if return_state:   # amplitudes after the iteration + the factor blocks (for the (T) of the first bench step)
        return float(e), rms, float(e_mp2), {"t1": t1n, "t2": t2n, "Boo": Boo, "Bov": Bov, "Bvv": Bvv, "eps": eps}
    return float(e), rms, float(e_mp2)


def cpu_ccsd_iter1(nbf=400, nocc=40):
    e, rms, e_mp2 = ccsd_iter1_from_factors(nbf, nocc, log=lambda *a: print(*a, flush=True))
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["e_ccsd_iter1_cpu_port"] = e
    pins[key]["iter1_cpu_source"] = ("tests/golden/make_bench_pins.py cpu_ccsd_iter1 (NumPy from the factored integrals: the oracle's "
                                     "spin-free iteration with t1 = 0 on input, ladder integrals built in slabs)")
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, "E_CCSD after iteration 1 (CPU):", e, " rms:", rms, " GPU-pinned step 1:", (pins[key].get("steps") or [[None]])[0][0])


def cpu_T_target(nbf=400, nocc=40, limit=None):
    """e_T of the FIRST bench step at a shape without a CPU port iteration (the target shape): amplitudes after the first
    CCSD iteration from ccsd_iter1_from_factors (the DIIS step that follows has one history entry and leaves them unchanged),
    then the [T] accumulator over all unique triples through oracle/cpu_port.py: orbit_T_fast (dgemm-accumulated products + C
    epilogue; ~0.9 s per orbit at nbf=400 on 8 cores, 11480 orbits: about three hours).  Checkpointed: a rerun resumes."""
    import time

    from oracle import cpu_port

    t0 = time.perf_counter()
    e1, rms, e_mp2, st = ccsd_iter1_from_factors(nbf, nocc, log=lambda *a: print(*a, flush=True), return_state=True)
    o, v = nocc, nbf - nocc
    naux = st["Bvv"].shape[0]
    Bvv2 = st["Bvv"].reshape(naux, v * v)
    Bov, Boo = st["Bov"], st["Boo"]

    def iv_block(k):      # [d, (b,c)] = (ck|bd) = sum_P Bov[P,k,c] Bvv[P,b,d]
        x = (Bvv2.T @ Bov[:, k, :]).reshape(v, v, v)        # [b, d, c]
        return np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(v, v * v)

    v_oovo = np.einsum("Pia,Pjl->ijal", Bov, Boo, optimize=True)
    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    triples = unique_triples(o)
    if limit:
        triples = triples[:int(limit)]
    print(f"E_CCSD after iteration 1: {e1!r}; operands for {len(triples)} orbits ...", flush=True)
    ck = os.path.join(os.path.dirname(PATH), f"_cpu_T_checkpoint_nbf{nbf}.json")
    e_T = cpu_port.orbit_T_fast(lib, np.ascontiguousarray(st["t2"]), iv_block, v_oovo, st["eps"], triples, progress=200,
                                checkpoint=None if limit else (ck, 100))
    print(f"e_T = {e_T!r}  ({time.perf_counter() - t0:.0f} s in all)", flush=True)
    if limit:
        return e_T
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["e_T_step1_cpu_oracle"] = e_T
    pins[key]["cpu_T_source"] = ("tests/golden/make_bench_pins.py cpu_T_target (amplitudes after the first CCSD iteration from the "
                                 "factored integrals; [T] accumulator over all unique triples: dgemm-accumulated products + C "
                                 "epilogue, oracle/cpu_port.py orbit_T_fast)")
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, "e_T_step1_cpu_oracle", e_T, "GPU-pinned step 1:", (pins[key].get("steps") or [[None, None]])[0][1])
    if os.path.exists(ck):
        os.remove(ck)


def cpu_mp1_triples(nbf=400, nocc=40):
    """Pins for bench.py's per-triple (T) check at a shape where no CPU CCSD iteration is affordable (the target shape):
    sixteen unique triples of all three orbit kinds (i=j=k, two equal, all different), spread over the list."""
    o = nocc
    all_t = unique_triples(o)
    picks = [(0, 0, 0), (o // 6, o // 6, o // 2), (o // 8, o // 2, (4 * o) // 5), (o // 3, (2 * o) // 3, (2 * o) // 3),
             (o - 3, o - 2, o - 1), (o - 1, o - 1, o - 1),
             # ten more, spread over the list: the lowest and highest orbitals, neighbours, both two-equal patterns
             (0, 1, 2), (0, 0, o - 1), (0, o - 1, o - 1), (0, o // 2, o - 1), (1, o // 4, o // 2), (o // 5, o // 5 + 1, o // 5 + 2),
             (o // 2, o // 2, o // 2 + 1), (o // 2 - 1, o // 2, o - 2), ((3 * o) // 5, (7 * o) // 10, (9 * o) // 10),
             (o // 10, (3 * o) // 4, (3 * o) // 4)]
    assert len(set(picks)) == len(picks) and all(i <= j <= k < o for i, j, k in picks)
    e, _ = mp1_triples_from_factors(nbf, nocc, picks)
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["mp1_triples"] = {"ntriples": len(all_t), "ijk": [list(t) for t in picks], "ranks": [all_t.index(t) for t in picks],
                                "e_T": [float(x) for x in e],
                                "source": "tests/golden/make_bench_pins.py cpu_mp1_triples (NumPy from the factored integrals: "
                                          "MP1 amplitudes, the [T] accumulator of single (i<=j<=k) orbits through the oracle's "
                                          "BLAS orbit form)"}
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, pins[key]["mp1_triples"])


def cpu_T(nbf=200, nocc=20, limit=None):
    """e_T of the bench's FIRST step on the CPU: amplitudes after one CCSD iteration of the CPU port (the DIIS extrapolation
    that follows has a single history entry and returns them unchanged), then the [T] accumulator over all triples through
    the oracle's BLAS orbit form (oracle/afesp_oracle.py: triples_bracket_T_orbit_form; ~15 CPU-minutes on 8 cores at
    nbf=200).  Pins the headline-shape (T) of every bench line on a CPU computation instead of on a GPU run."""
    import time

    from oracle import afesp_oracle as orc
    from oracle import cpu_port

    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    mo, Cmo, eps = cpu_port.synthetic_mo_integrals(nbf, nocc)
    V = cpu_port.slices(lib, mo, nbf, nocc)
    o = nocc
    D1, D2 = orc.denominators(eps, o)
    t2 = np.asfortranarray(V["v_oovv"] / D2)
    t1 = np.zeros((o, nbf - o), order="F")
    t1n, t2n, _, _ = cpu_port.ccsd_iter(lib, V, eps, t1, t2)
    del V["v_vvvv"]
    triples = None
    if limit:
        triples = [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)][:int(limit)]
    t0 = time.perf_counter()
    e_T = orc.triples_bracket_T_orbit_form(np.ascontiguousarray(t2n), np.ascontiguousarray(V["v_vvov"]),
                                           np.ascontiguousarray(V["v_oovo"]), eps, triples=triples, progress=50)
    print(f"e_T = {e_T!r}  ({time.perf_counter() - t0:.0f} s)")
    if limit:
        return
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["e_T_step1_cpu_oracle"] = e_T
    pins[key]["cpu_T_source"] = ("tests/golden/make_bench_pins.py cpu_T (one CCSD iteration through oracle/cpu_ccsd.c, then the [T] "
                                 "accumulator over all triples through the oracle's BLAS orbit form)")
    json.dump(pins, open(PATH, "w"), indent=1)
    print(key, "e_T_step1_cpu_oracle", e_T, "GPU-pinned step 1:", (pins[key].get("steps") or [[None, None]])[0][1])


def cpu_traj(nbf=200, nocc=20, nsteps=3):
    """The first `nsteps` steps of the bench trajectory on the CPU, in the bench's own order (bench.py step()): one CCSD
    iteration (CPU port of the reference, energy from its output amplitudes), CC-DIIS extrapolation (the oracle's ring of
    depth 8), then the [T] accumulator ON THE EXTRAPOLATED amplitudes (orbit form: oracle/cpu_port.py orbit_T_fast, which equals the
    oracle's NumPy orbit form and the literal loop).  ~2.5 CPU-minutes per step on 8 cores at nbf=200.  Stored as steps_cpu = [[E_CCSD, e_T], ...]."""
    import time

    from oracle import afesp_oracle as orc
    from oracle import cpu_port

    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    mo, Cmo, eps = cpu_port.synthetic_mo_integrals(nbf, nocc)
    V = cpu_port.slices(lib, mo, nbf, nocc)
    o = nocc
    D1, D2 = orc.denominators(eps, o)
    voovv = np.asarray(V["v_oovv"])
    t2 = np.asfortranarray(voovv / D2)
    t1 = np.zeros((o, nbf - o), order="F")
    diis = orc.CCDiis(8, t1.shape, t2.shape)
    vvov, oovo = np.ascontiguousarray(V["v_vvov"]), np.ascontiguousarray(V["v_oovo"])
    nv = nbf - nocc
    ivb = lambda k: np.ascontiguousarray(vvov[:, :, k, :].transpose(2, 1, 0)).reshape(nv, nv * nv)
    uniq = unique_triples(o)
    steps = []
    for it in range(int(nsteps)):
        t0 = time.perf_counter()
        diis.stash(np.asarray(t1), np.asarray(t2))
        t1n, t2n, _, _ = cpu_port.ccsd_iter(lib, V, eps, np.asfortranarray(t1), np.asfortranarray(t2))
        e_cc = orc.restricted_energy(np.asarray(t1n), np.asarray(t2n), voovv)
        t1, t2 = diis.update(np.asarray(t1n), np.asarray(t2n))
        e_T = cpu_port.orbit_T_fast(lib, np.ascontiguousarray(t2), ivb, oovo, eps, uniq)
        steps.append([float(e_cc), float(e_T)])
        print(f"step {it + 1}: E_CCSD {e_cc!r}  e_T {e_T!r}  ({time.perf_counter() - t0:.0f} s)", flush=True)
    pins = load()
    key = f"nbf{nbf}_nocc{nocc}"
    pins.setdefault(key, {})
    pins[key]["steps_cpu"] = steps
    pins[key]["steps_cpu_source"] = ("tests/golden/make_bench_pins.py cpu_traj: per step one CCSD iteration through oracle/cpu_ccsd.c, "
                                     "the oracle's CC-DIIS, the [T] accumulator over all triples in orbit form (oracle/cpu_port.py orbit_T_fast)")
    json.dump(pins, open(PATH, "w"), indent=1)
    for a, b in zip(steps, pins[key].get("steps", [])):
        print("cpu", a, "gpu-pinned", b, "diff", abs(a[0] - b[0]), abs(a[1] - b[1]))


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
    elif sys.argv[1] == "cpu_T_target":
        cpu_T_target(*[int(x) for x in sys.argv[2:5]])
    elif sys.argv[1] == "cpu_ccsd_iter1":
        cpu_ccsd_iter1(*[int(x) for x in sys.argv[2:4]])
    elif sys.argv[1] == "cpu_mp1_triples":
        cpu_mp1_triples(*[int(x) for x in sys.argv[2:4]])
    elif sys.argv[1] == "cpu_mp2":
        cpu_mp2(*[int(x) for x in sys.argv[2:4]])
    elif sys.argv[1] == "cpu_traj":
        cpu_traj(*[int(x) for x in sys.argv[2:5]])
    elif sys.argv[1] == "cpu_T":
        cpu_T(*[int(x) for x in sys.argv[2:5]])
    else:
        gpu(sys.argv[2])
