"""Stress the TMA GEMM against the cp.async kernel, element-wise, repeated: where do mismatches fall?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu
g = AfespGpu(0)
g.set_option("gemm_force_config", 3)
rng = np.random.default_rng(3)
reps = int(os.environ.get("REPS", 8))
for (M, N, K, beta) in [(144, 13456, 144, 1.0), (144, 13456, 144, 0.0), (128, 13440, 144, 1.0), (400, 16290, 400, 1.0), (384, 16256, 400, 1.0)]:
    A = rng.standard_normal((M, K)) * 1e-3; B = rng.standard_normal((K, N)) * 1e-2; C0 = rng.standard_normal((M, N)) * 1e-3
    Af, Bf, Cf = A.ravel(order="F"), B.ravel(order="F"), C0.ravel(order="F")
    g.set_option("gemm_use_tma", 0)
    ref = g.dgemm_wrapper("N", "N", M, N, K, Af, Bf, Cf, alpha=0.5, beta=beta)
    ref2 = g.dgemm_wrapper("N", "N", M, N, K, Af, Bf, Cf, alpha=0.5, beta=beta)
    g.set_option("gemm_use_tma", 2)
    tot, info = 0, []
    for rep in range(reps):
        out = g.dgemm_wrapper("N", "N", M, N, K, Af, Bf, Cf, alpha=0.5, beta=beta)
        d = np.abs(out - ref).reshape((M, N), order="F")
        bad = np.argwhere(d > 1e-15)
        tot += len(bad)
        if len(bad):
            mt = sorted(set((bad[:, 0] // 64).tolist())); rows = sorted(set((bad[:, 0] % 64).tolist()))
            info.append((len(bad), "mtiles", mt, "rows-in-tile", rows[:4], rows[-2:], "ntiles", len(set((bad[:, 1] // 64).tolist())), "maxerr %.1e" % d.max()))
    print(M, N, K, "beta", beta, "cpasync self-consistent", bool(np.array_equal(ref, ref2)), "bad total", tot, info[:3], flush=True)
