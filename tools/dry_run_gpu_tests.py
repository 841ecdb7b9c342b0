"""Developer aid: run the bodies of the program-level GPU parity tests (tests/test_gpu_parity.py) against the oracle-backed
test double (tests/_oracle_engine.py) on the CPU -- checks the test logic and the host code when no GPU is at hand; the
numerics of the CUDA library are of course not exercised.  python tools/dry_run_gpu_tests.py"""
import sys, pathlib, tempfile, traceback
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import tests.test_gpu_parity as T
from tests._oracle_engine import OracleEngine
from oracle import afesp_oracle as orc
from tests._fixtures import load_system
cache={}
def oracle_runs(name, calc_type=None, **kw):
    key=(name,calc_type,tuple(sorted(kw.items())))
    if key not in cache:
        s=load_system(name,calc_type); cache[key]=(s,orc.run(s,**kw))
    return cache[key]
def run(fn,*a):
    try:
        fn(*a); print('PASS',fn.__name__,a[1:] if len(a)>1 else '')
    except Exception as e:
        print('FAIL',fn.__name__,type(e).__name__,str(e)[:300]); traceback.print_exc(limit=3)
for name in ['n2','f2']:
    run(T.test_crccsd_t_spatial_matches_shipped_els_out, OracleEngine(), name)
    run(T.test_whole_program_output_matches_shipped_els_out, OracleEngine(), name, pathlib.Path(tempfile.mkdtemp()))
for calc in ["CCSD(T)_spatial", "CCSD[T]_spatial", "RCCSD(T)_spatial", "RCCSD[T]_spatial","CRCCSD[T]_spatial"]:
    run(T.test_spatial_calc_types_match_oracle, OracleEngine(), calc, oracle_runs)
run(T.test_spinorbital_ccsd_matches_old_ref_out_with_q1_off, OracleEngine())
run(T.test_spinorbital_ccsd_t_as_coded_matches_oracle, OracleEngine(), oracle_runs)
run(T.test_h2o_cc_pvtz_spinorbital_ccsd_t_matches_reference_els_cpu_out, OracleEngine())
for calc in ["CCSD(T)_spatial", "RCCSD(T)_spatial"]:
    run(T.test_h2o_cc_pvtz_spin_free_matches_oracle, OracleEngine(), calc)
run(T.test_diis_history_deeper_than_eight_matches_oracle, OracleEngine())
# the multi-tile tests, at small stand-in shapes (their CPU sides are the same code at any shape)
for n, o in [(30, 4), (27, 3)]:
    run(T.test_one_bench_step_at_a_multi_tile_shape_matches_the_cpu, OracleEngine(), n, o)
run(T.test_crccsd_t_chain_at_a_multi_tile_shape_matches_the_cpu, OracleEngine(), 24, 4)
run(T.test_single_triple_shares_on_mp1_amplitudes_match_cpu_values_from_the_factored_integrals, OracleEngine())
