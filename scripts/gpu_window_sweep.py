"""On-GPU: device-resident throughput of one bench-shaped step (4096 streams x 125 chunks) as a function of the window size."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200
S, N = 4096, 125
base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(8)]
pcm = np.stack([np.roll(base[s % 8].reshape(N, 1536), (s // 8) % N, axis=0).reshape(-1) for s in range(S)])
for w in [int(v) for v in os.environ.get("WINDOWS", "0,25,32,42,63,125").split(",")]:
    e = vadc_b200.Engine(max_streams=S, window_chunks=w)
    d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm)
    for it in range(2):
        e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
    e.sync()
    e.timer_start()
    for it in range(5):
        e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
    ms = e.timer_stop() / 5
    print("window %3d chunks: %.3f ms per step -> %.2f M chunks/s" % (w, ms, S * N / ms / 1e3), flush=True)
    e.close()
