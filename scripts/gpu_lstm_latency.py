"""GPU: serial latency of the exact path's two LSTM kernels on ONE stream (the bound of BASELINE cfg1 / cfg4)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200

e = vadc_b200.Engine()
rng = np.random.default_rng(1)
for B in (200, 3000):
    x = np.maximum(rng.standard_normal((B, 7, 64)).astype(np.float32), 0)
    for wave in (True, False):
        e.stage_exact_lstm(x, wave=wave)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            out = e.stage_exact_lstm(x, wave=wave)
            best = min(best, time.perf_counter() - t)
        print("B=%d chunks (%d steps x 2 layers) %s: %.2f ms = %.2f us per step (both layers)" % (B, B * 7, "wavefront kernel " if wave else "multi-stream kernel", best * 1e3, best * 1e6 / (B * 7)), flush=True)
