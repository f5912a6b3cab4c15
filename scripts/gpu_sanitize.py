"""Compact workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the exact path in both of its mappings, the
device segmenter, and the opt-in fast family, on small shapes; results are compared with the oracle so that a sanitizer run is also a
parity run.   compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

fast = len(sys.argv) > 1 and sys.argv[1] == "fast"
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
ok = True
for S, N in ((2, 12), (160, 10), (800, 2)):
    pcm = np.stack([vadc_b200.synth_pcm(10 + (s % 6), N * 1536) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S)
    e.segments_configure()
    segs, counts, probs = e.run_streams_segments(pcm, end_of_stream=True, want_probs=True)
    e.close()
    for s in (0, S - 1):
        ref = Oracle().run_pcm(pcm[s])[:, 1]
        same = np.array_equal(bits(probs[s]), bits(ref))
        ok &= same
        print("exact S=%d N=%d stream %d: bit-identical %s" % (S, N, s, same), flush=True)
if fast:
    S, N = 64, 6
    pcm = np.stack([vadc_b200.synth_pcm(30 + s, N * 1536) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, stft_mode=vadc_b200.STFT_HYBRID, lstm_mode=vadc_b200.LSTM_TENSOR, layer_mode=vadc_b200.LAYERS_TENSOR)
    p = e.run_streams(pcm)
    e.close()
    err = max(float(np.abs(p[s] - Oracle().run_pcm(pcm[s])[:, 1]).max()) for s in (0, 63))
    ok &= err <= 1e-4
    print("fast family S=%d N=%d: max |d| %.2e" % (S, N, err), flush=True)
print("RESULT", "ok" if ok else "MISMATCH")
