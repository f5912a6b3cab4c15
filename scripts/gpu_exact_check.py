"""GPU: the exact (reference rounding sequence) path at large stream counts -- bits against the oracle, then stage times.
   python scripts/gpu_exact_check.py > gpurun_out/exact_check.log"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def check(S, N, window, sample):
    base = [vadc_b200.synth_pcm(900 + 17 * i, 1536 * N) for i in range(16)]
    pcm = np.stack([np.roll(base[s % 16], 1536 * ((s // 16) % N)) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, layer_mode=vadc_b200.LAYERS_FAITHFUL, window_chunks=window)
    t = time.perf_counter()
    probs, out2 = e.run_streams(pcm, want_out2=True)
    dt = time.perf_counter() - t
    e.close()
    nbad, worst = 0, 0.0
    for s in sample:
        ref = Oracle().run_pcm(pcm[s])
        nbad += int((bits(out2[s]) != bits(ref)).sum())
        worst = max(worst, float(np.abs(out2[s] - ref).max()))
    print("S=%d N=%d window=%d: %d sampled streams, differing values %d, max |d| %.3e (%.1f ms)" % (S, N, window, len(sample), nbad, worst, dt * 1e3), flush=True)
    return nbad


bad = 0
bad += check(75, 20, 7, range(0, 75, 5))
bad += check(149, 12, 0, [0, 1, 74, 147, 148])
bad += check(300, 30, 11, range(0, 300, 23))
bad += check(1185, 9, 4, [0, 147, 148, 295, 296, 1036, 1183, 1184])
bad += check(4096, 24, 0, range(0, 4096, 293))
print("RESULT", "ok" if bad == 0 else "DIFFERENCES")
