"""TEST DOUBLE (see afesp_gpu_double.c): raw-pointer bridge from the C ABI to tests/_oracle_engine.OracleEngine."""
import ctypes
import os
import traceback

import numpy as np

from tests._oracle_engine import OracleEngine


def _arr(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))


def _npk(n):
    npair = n * (n + 1) // 2
    return npair * (npair + 1) // 2


class Bridge:
    def __init__(self):
        self.eng = OracleEngine()
        self.last_error = ""
        self.n = 0
        # (tests) make the symmetry assertion fire the way a corrupted integral would: status 5 + the reference's message
        self.force_symmetry_error = float(os.environ.get("AFESP_DOUBLE_SYMMETRY_ERROR", "0"))
        log = os.environ.get("AFESP_DOUBLE_CALL_LOG")
        self.log = open(log, "w") if log else None

    def _guard(self, name, fn):
        try:
            if self.log:
                self.log.write(name + "\n")
                self.log.flush()
            return fn() or 0
        except Exception as ex:  # status 1 + text, as the library's guarded() does
            self.last_error = f"{name}: {type(ex).__name__}: {ex}"
            traceback.print_exc()
            return 1

    def set_option(self, key, value):
        return self._guard("set_option", lambda: self.eng.set_option(key, value))

    def ao2mo(self, n, p_eri, p_c, p_out):
        def run():
            self.n = n
            # coeff arrives column-major as C(mo,ao) (sys%canon_coeff): element (mo,ao) at mo + n*ao
            c = _arr(p_c, n * n).reshape((n, n), order="F").copy()
            out = self.eng.ao2mo(n, _arr(p_eri, _npk(n)).copy(), c, want_result=bool(p_out))
            if p_out:
                _arr(p_out, _npk(n))[:] = out
        return self._guard("ao2mo", run)

    def mp2_energy(self, nocc, p_eps, p_e):
        def run():
            _arr(p_e, 1)[0] = self.eng.mp2_energy(nocc, _arr(p_eps, self.n).copy())
        return self._guard("mp2_energy", run)

    def ccsd_init(self, nocc, restricted, p_eps, diis_n, p_e, p_rms):
        def run():
            e, rms = self.eng.ccsd_init(nocc, bool(restricted), _arr(p_eps, self.n).copy(), diis_n)
            _arr(p_e, 1)[0], _arr(p_rms, 1)[0] = e, rms
            if not restricted and self.force_symmetry_error > 0:
                self.eng.sym_err = self.force_symmetry_error
                self.last_error = "Permutational symmetry of antisymmetrised integrals does not hold"
                return 5
        return self._guard("ccsd_init", run)

    def ccsd_init_info(self, p_info):
        def run():
            info = self.eng.ccsd_init_info()
            _arr(p_info, 4)[:] = [info["symmetry_error"], info["slices_s"], info["check_s"], 0.0]
        return self._guard("ccsd_init_info", run)

    def ccsd_iterate(self, p_e, p_rms):
        def run():
            e, rms = self.eng.ccsd_iterate()
            _arr(p_e, 1)[0], _arr(p_rms, 1)[0] = e, rms
        return self._guard("ccsd_iterate", run)

    def ccsd_diis(self):
        return self._guard("ccsd_diis", self.eng.ccsd_diis)

    def ccsd_finalize(self, want_cr, p_diag, p_t1, p_t2):
        def run():
            diag, t1, t2 = self.eng.ccsd_finalize(want_cr=bool(want_cr), want_amplitudes=bool(p_t1 or p_t2))
            _arr(p_diag, 1)[0] = diag
            if p_t1:
                _arr(p_t1, t1.size)[:] = t1.ravel(order="F")
            if p_t2:
                _arr(p_t2, t2.size)[:] = t2.ravel(order="F")
        return self._guard("ccsd_finalize", run)

    def ccsd_t_spatial(self, paren, renorm, comp_renorm, p_sums, p_const):
        def run():
            sums, const = self.eng.ccsd_t_spatial(bool(paren), bool(renorm), bool(comp_renorm))
            _arr(p_sums, 6)[:] = sums
            _arr(p_const, 1)[0] = const
        return self._guard("ccsd_t_spatial", run)

    def ccsd_t_spinorb(self, p_e):
        def run():
            _arr(p_e, 1)[0] = self.eng.ccsd_t_spinorb()
        return self._guard("ccsd_t_spinorb", run)
