#!/bin/bash
# Host-side C/C++ of the product (host/els_host.cpp, host/gint.c) under AddressSanitizer + UBSan, no GPU needed:
#   RHF-level runs of the N2 / F2 / cc-pVTZ directories (readers, namelist, RHF, integral generator with s..f shells, error
#   block) and the complete CRCCSD(T)_spatial / CCSD(T)_spinorb flows over the oracle-backed test double of the C ABI.
# Prints one line per run; any sanitizer report shows up in its stderr tail.  Last result: all clean (round 2).
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W"
gcc -O1 -g -fopenmp -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -c -o gint.o "$ROOT/host/gint.c"
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -o els_host_san "$ROOT/host/els_host.cpp" gint.o \
    "$(gcc -print-file-name=libgomp.so)" -lm -L"$ROOT/afesp_b200/lib" -lafesp_gpu -Wl,-rpath,"$ROOT/afesp_b200/lib"
INC=$(python -c "import sysconfig;print(sysconfig.get_config_var('INCLUDEPY'))")
LIBDIR=$(python -c "import sysconfig;print(sysconfig.get_config_var('LIBDIR'))")
VER=$(python -c "import sysconfig;print(sysconfig.get_config_var('LDVERSION'))")
gcc -O1 -shared -fPIC -I "$INC" "$ROOT/tests/_double/afesp_gpu_double.c" -o double.so -L"$LIBDIR" -lpython"$VER"
ROOT="$ROOT" W="$W" python - <<'P'
import os, pathlib, subprocess, sys, tempfile
ROOT, W = os.environ["ROOT"], os.environ["W"]
sys.path.insert(0, ROOT)
from tests._fixtures import write_sample_dir
from tests.test_gint import _write_tz_dir
asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
base = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:verify_asan_link_order=0", UBSAN_OPTIONS="print_stacktrace=1")
dbl = dict(base, LD_PRELOAD=asan + ":" + os.path.join(W, "double.so"), PYTHONMALLOC="malloc",
           PYTHONPATH=os.pathsep.join([ROOT] + [p for p in sys.path if p.endswith("site-packages")]))
def run(tag, d, env):
    r = subprocess.run([os.path.join(W, "els_host_san"), str(d)], capture_output=True, text=True, env=env, timeout=3600)
    print(tag, "rc", r.returncode, "| stderr:", (r.stderr[-800:] if r.stderr.strip() else "(none)"))
for name, calc in [("n2", "RHF"), ("f2", "UHF")]:
    d = tempfile.mkdtemp(); write_sample_dir(name, d, calc_type=calc); run(f"{name} {calc}", d, base)
d = pathlib.Path(tempfile.mkdtemp()); _write_tz_dir(d, calc_type="RHF"); run("h2o cc-pVTZ RHF (integrals generated)", d, base)
run("empty directory (error block)", tempfile.mkdtemp(), base)
for name, calc in [("n2", None), ("h2o", "CCSD(T)_spinorb")]:
    d = tempfile.mkdtemp(); write_sample_dir(name, d, calc_type=calc); run(f"{name} {calc or 'CRCCSD(T)_spatial'} over the double", d, dbl)
P
