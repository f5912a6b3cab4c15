"""GPU: time of ONE stream on the default (faithful) engine -- BASELINE cfg1 (625 chunks) and a 3000-chunk stream."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200
for S, N in ((1, 625), (1, 3000), (8, 1000), (64, 400), (128, 200)):
    pcm = np.stack([vadc_b200.synth_pcm(1 + s, N * 1536) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S)
    e.run_streams(pcm)
    best = 1e9
    for _ in range(3):
        e.reset()
        t = time.perf_counter()
        e.run_streams(pcm)
        best = min(best, time.perf_counter() - t)
    e.set_profiling(True); e.reset(); e.run_streams(pcm); ms, n = e.last_timing(); e.close()
    print("S=%d N=%d: %.1f ms = %.0f x realtime; stages %s" % (S, N, best * 1e3, S * N * 0.096 / best, " ".join("%.1f" % v for v in ms.values())), flush=True)
