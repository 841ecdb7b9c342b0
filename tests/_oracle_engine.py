"""TEST DOUBLE -- never imported by the product.

`OracleEngine` offers the step-granular methods of afesp_b200.capi.AfespGpu (ao2mo, mp2_energy, ccsd_init, ccsd_iterate,
ccsd_diis, ccsd_finalize, ccsd_t_spatial, ccsd_t_spinorb ...) on top of the NumPy oracle (oracle/afesp_oracle.py), so that
the host-side driver logic of afesp_b200.host -- the reference's iteration loop, convergence test, DIIS call order, printed
lines and energy assembly (src/main.F90, src/ccsd.f90:223-271, 339-396, 2239-2287) -- can be exercised by the CPU test
suite against the reference's shipped els.out files.  It exists only for `pytest -m "not gpu"`: afesp_b200 has no CPU path,
AfespGpu() fails without a Blackwell device, and the GPU parity tests run the same host code over the real library.
"""
import numpy as np

from oracle import afesp_oracle as orc


class OracleEngine:
    def __init__(self, q1=True, q3a=True, q3b=True, force_symmetry_error=0.0):
        self.q1, self.q3a, self.q3b = q1, q3a, q3b
        self.force_symmetry_error = force_symmetry_error   # > 0: ccsd_init fails the way the library's status 5 does
        self.calls = []          # the order in which the host drove the engine (asserted by the tests)
        self.eri_mo = None
        self.n = 0

    def _log(self, name):
        self.calls.append(name)

    # -- plumbing the host touches
    def close(self):
        self._log("close")

    def last_stage_ms(self):
        return 0.0

    def set_eri_mo(self, nbasis, eri_mo):
        self._log("set_eri_mo")
        self.n = int(nbasis)
        self.eri_mo = np.array(eri_mo, dtype=float)

    def set_partition(self, rank, nranks):
        """afesp_gpu_set_partition: the handle owns the unique (i <= j <= k) triples t with t % nranks == rank, in the
        enumeration order of afesp_b200/csrc/triples.cu (my_triples)."""
        self.partition = (int(rank), int(nranks))

    def release(self, what):
        self._log(f"release:{what}")

    def set_option(self, key, value):
        self._log(f"set_option:{key}")
        switches = {"q1_transposed_foo": "q1", "q3a_truncated_e": "q3a", "q3b_stale_intermediates": "q3b"}
        if key in switches:   # the parity switches of include/afesp_gpu.h
            setattr(self, switches[key], bool(value))

    # -- AO->MO + MP2 (src/mp2.f90:261-449)
    def ao2mo(self, nbasis, eri_ao=None, coeff=None, want_result=True):
        self._log("ao2mo")
        self.n = int(nbasis)
        self.eri_mo = orc.ao2mo_packed(np.asarray(eri_ao), np.asarray(coeff))
        return self.eri_mo.copy() if want_result else None

    def get_eri_mo(self, out=None):
        self._log("get_eri_mo")
        return self.eri_mo.copy()

    def mp2_energy(self, nocc, eps):
        self._log("mp2_energy")
        return orc.mp2_energy(self.eri_mo, np.asarray(eps), int(nocc))

    # -- CCSD (src/ccsd.f90:279-402 / 71-277), one reference iteration per call
    def ccsd_init(self, nocc, restricted, eps, diis_n=8):
        self._log("ccsd_init")
        self.restricted = bool(restricted)
        self.eps = np.asarray(eps, dtype=float)
        self.nocc = int(nocc)
        if self.restricted:
            self.V = orc.spatial_slices(self.eri_mo, self.n, self.nocc)
            self.D1, self.D2 = orc.denominators(self.eps, self.nocc)
            self.vo = self.V["v_oovv"]
            self._energy = orc.restricted_energy
        else:
            asym = orc.spinorb_antisym(self.eri_mo, self.n)
            self.sym_err = orc.spinorb_symmetry_error(asym)
            if self.force_symmetry_error > 0:
                from afesp_b200.capi import AfespError

                self.sym_err = self.force_symmetry_error
                raise AfespError("ccsd_init", 5, "Permutational symmetry of antisymmetrised integrals does not hold")
            self.G = orc.spinorb_slices(asym, 2 * self.nocc)
            self.eps_so = np.repeat(self.eps, 2)
            self.D1, self.D2 = orc.denominators(self.eps_so, 2 * self.nocc)
            self.vo = self.G["oovv"]
            self._energy = orc.spinorb_energy
        self.t1 = np.zeros_like(self.D1)
        self.t2 = self.vo / self.D2
        self.diis = orc.CCDiis(int(diis_n), self.t1.shape, self.t2.shape)
        self.I = None
        self.finalized = False
        e = self._energy(self.t1, self.t2, self.vo)
        rms = float(np.sum(self.t2 ** 2))
        self.t2_old = self.t2.copy()
        return e, rms

    def ccsd_init_info(self):
        return {"symmetry_error": getattr(self, "sym_err", 0.0), "slices_s": 0.0, "check_s": 0.0}

    def ccsd_iterate(self):
        self._log("ccsd_iterate")
        assert not self.finalized
        self.diis.stash(self.t1, self.t2)
        if self.restricted:
            self.I = orc.restricted_intermediates(self.t1, self.t2, self.V)
            self.t1, self.t2 = orc.restricted_amplitudes(self.t1, self.t2, self.V, self.I, self.D1, self.D2)
        else:
            self.t1, self.t2 = orc.spinorb_iteration(self.t1, self.t2, self.G, self.D1, self.D2, q1=self.q1)
        e = self._energy(self.t1, self.t2, self.vo)
        rms = float(np.sum((self.t2 - self.t2_old) ** 2))
        self.t2_old = self.t2.copy()
        self.e_ccsd = e
        return e, rms

    def ccsd_diis(self):
        self._log("ccsd_diis")
        assert not self.finalized
        self.t1, self.t2 = self.diis.update(self.t1, self.t2)

    def ccsd_finalize(self, want_cr=False, want_amplitudes=False, out=None):
        self._log("ccsd_finalize")
        self.finalized = True
        diag = orc.t1_diagnostic(self.t1, 2 * self.nocc) if self.restricted else 0.0
        self.cr = None
        if want_cr:
            if self.q3b:
                ivo, asym = self.I["I_vo"], self.I["asym_t2"]
            else:
                fresh = orc.restricted_intermediates(self.t1, self.t2, self.V)
                ivo, asym = fresh["I_vo"], fresh["asym_t2"]
            self.cr = orc.cr_intermediates(self.t1, self.t2, self.V, ivo, asym, q3a=self.q3a)
        return diag, (self.t1 if want_amplitudes else None), (self.t2 if want_amplitudes else None)

    # -- triples (src/ccsd.f90:2018-2293 / 1812-1922)
    def ccsd_t_spatial(self, paren, renorm, comp_renorm):
        self._log("ccsd_t_spatial")
        assert self.finalized and self.restricted
        cr = self.cr if comp_renorm else (None, None)
        triples = None
        rank, nranks = getattr(self, "partition", (0, 1))
        if nranks > 1:   # this share of the unique triples, each expanded into its distinct orderings (the reference's loop)
            o = self.nocc
            uniq = [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)]
            mine = [t for n_t, t in enumerate(uniq) if n_t % nranks == rank]
            triples = sorted({p for (i, j, k) in mine for p in [(i, j, k), (i, k, j), (j, i, k), (j, k, i), (k, i, j), (k, j, i)]})
        sums = orc.triples_spatial_sums(self.t1, self.t2, self.V["v_oovv"], self.V["v_vvov"], self.V["v_oovo"], self.eps,
                                        paren, renorm, comp_renorm, cr[0], cr[1], triples=triples)
        const = orc.triples_denominator_constant(self.t1, self.t2) if (renorm or comp_renorm) else 0.0
        return np.array(sums), const

    def ccsd_t_spinorb(self):
        self._log("ccsd_t_spinorb")
        assert self.finalized and not self.restricted
        G = self.G
        return orc.triples_spinorb(self.t1, self.t2, G["oovv"], G["vovv"], G["ovoo"], self.eps_so)
