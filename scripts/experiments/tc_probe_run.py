"""GPU experiment: tcgen05 probe -- correctness and MMA-phase cycle counts at small N."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vadc_b200
e = vadc_b200.Engine(max_streams=4)
rng = np.random.default_rng(0)
for n, k in ((16, 16), (16, 128), (32, 128), (64, 128), (128, 128), (256, 128)):
    a = rng.standard_normal((128, k)).astype(np.float32)
    b = rng.standard_normal((n, k)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    for ns in ((1, 2) if n > 64 else (1, 2, 3)):
        d, _ = e.stage_tc_gemm(a, b, nsplit=ns)
        _, c1 = e.stage_tc_gemm(a, b, nsplit=ns, reps=1)
        _, c101 = e.stage_tc_gemm(a, b, nsplit=ns, reps=101)
        nmma = (k // 16) * (1, 3, 6)[ns - 1]
        print("N=%3d K=%3d nsplit=%d  max|err|=%.3e  (rel to sqrt(K): %.2e)  cycles/rep=%.0f  per-MMA=%.1f" % (
            n, k, ns, np.abs(d - ref).max(), np.abs(d - ref).max() / np.sqrt(k), (c101 - c1) / 100.0, (c101 - c1) / 100.0 / nmma), flush=True)
