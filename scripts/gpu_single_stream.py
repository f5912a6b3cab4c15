"""On-GPU: cfg4-shaped run -- ONE long stream (serial LSTM scan), host PCM in, probabilities + segments out."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200
N = int(os.environ.get("NCHUNKS", "37500"))       # 1 hour
base = vadc_b200.synth_pcm(77, 1536 * 3750)
pcm = np.tile(base, N // 3750 + 1)[: N * 1536][None, :]
e = vadc_b200.Engine(max_streams=1)
e.run_streams(pcm[:, : 1536 * 100])
e.reset()
t0 = time.time()
p = e.run_streams(pcm)
dt = time.time() - t0
tm, nl = e.last_timing()
print("single stream, %d chunks (%.1f h): wall %.3f s, device %.1f ms, %d launches -> %.0f x realtime" % (N, N * 0.096 / 3600, dt, tm["total"], nl, N * 0.096 / dt), flush=True)
e.set_profiling(1); e.reset(); e.run_streams(pcm[:, : 1536 * 5000]); tm, nl = e.last_timing()
print({k: round(v, 2) for k, v in tm.items()})
