/* TEST DOUBLE -- not part of the product, never built by __graft_entry__.build(), never installed next to the library.
 *
 * The functions of include/afesp_gpu.h that host/els_host.cpp calls, implemented by handing every call to the NumPy
 * oracle through an embedded Python interpreter (tests/_double/bridge.py -> tests/_oracle_engine.py).  The CPU test suite
 * builds it into a temporary directory and LD_PRELOADs it in front of libafesp_gpu.so so that the C++ host program's own
 * logic -- readers, RHF, the reference's iteration loop / convergence test / DIIS call order, every printed line, the
 * energy assembly, the error block -- can be compared with the reference's shipped els.out files without a GPU
 * (tests/test_els_host_flow.py).  The product has no CPU path: libafesp_gpu.so fails to open without a Blackwell device.
 */
#include <Python.h>
#include <stdio.h>
#include <string.h>

static PyObject* g_bridge = NULL;
static char g_err[512] = "";

static int call(const char* fn, const char* fmt, ...) {
  if (!g_bridge) { snprintf(g_err, sizeof g_err, "double: not open"); return 1; }
  va_list ap;
  va_start(ap, fmt);
  PyObject* meth = PyObject_GetAttrString(g_bridge, fn);
  PyObject* args = meth ? Py_VaBuildValue(fmt, ap) : NULL;
  va_end(ap);
  PyObject* res = (meth && args) ? PyObject_CallObject(meth, args) : NULL;
  int rc = 2;
  if (res) {
    rc = (int)PyLong_AsLong(res);
    if (rc != 0) {
      PyObject* msg = PyObject_GetAttrString(g_bridge, "last_error");
      const char* text = msg ? PyUnicode_AsUTF8(msg) : NULL;
      snprintf(g_err, sizeof g_err, "%s", text ? text : "double: failure");
      Py_XDECREF(msg);
    }
  } else {
    PyErr_Print();
    snprintf(g_err, sizeof g_err, "double: python exception in %s", fn);
  }
  Py_XDECREF(res); Py_XDECREF(args); Py_XDECREF(meth);
  return rc;
}

#define P(x) ((unsigned long long)(size_t)(x))

int afesp_gpu_open(int device, void** h) {
  (void)device;
  if (!Py_IsInitialized()) Py_Initialize();
  PyObject* mod = PyImport_ImportModule("tests._double.bridge");
  if (!mod) { PyErr_Print(); snprintf(g_err, sizeof g_err, "double: cannot import tests._double.bridge"); return 2; }
  g_bridge = PyObject_CallMethod(mod, "Bridge", NULL);
  Py_DECREF(mod);
  if (!g_bridge) { PyErr_Print(); return 2; }
  *h = (void*)g_bridge;
  return 0;
}
int afesp_gpu_close(void* h) { (void)h; Py_XDECREF(g_bridge); g_bridge = NULL; return 0; }
const char* afesp_gpu_last_error(void* h) { (void)h; return g_err; }
int afesp_gpu_set_option(void* h, const char* key, double value) { (void)h; return call("set_option", "(sd)", key, value); }
int afesp_gpu_ao2mo(void* h, int n, const double* eri_ao, const double* coeff, double* eri_mo) {
  (void)h; return call("ao2mo", "(iKKK)", n, P(eri_ao), P(coeff), P(eri_mo));
}
int afesp_gpu_mp2_energy(void* h, int nocc, const double* eps, double* e) { (void)h; return call("mp2_energy", "(iKK)", nocc, P(eps), P(e)); }
int afesp_gpu_ccsd_init(void* h, int nocc, int restricted, const double* eps, int diis_n, double* e, double* rms) {
  (void)h; return call("ccsd_init", "(iiKiKK)", nocc, restricted, P(eps), diis_n, P(e), P(rms));
}
int afesp_gpu_ccsd_init_info(void* h, double info[4]) { (void)h; return call("ccsd_init_info", "(K)", P(info)); }
int afesp_gpu_ccsd_iterate(void* h, double* e, double* rms) { (void)h; return call("ccsd_iterate", "(KK)", P(e), P(rms)); }
int afesp_gpu_ccsd_diis(void* h) { (void)h; return call("ccsd_diis", "()"); }
int afesp_gpu_ccsd_finalize(void* h, int want_cr, double* t1_diag, double* t1, double* t2) {
  (void)h; return call("ccsd_finalize", "(iKKK)", want_cr, P(t1_diag), P(t1), P(t2));
}
int afesp_gpu_ccsd_t_spatial(void* h, int paren, int renorm, int comp_renorm, double sums[6], double* dconst) {
  (void)h; return call("ccsd_t_spatial", "(iiiKK)", paren, renorm, comp_renorm, P(sums), P(dconst));
}
int afesp_gpu_ccsd_t_spinorb(void* h, double* e_T) { (void)h; return call("ccsd_t_spinorb", "(K)", P(e_T)); }
