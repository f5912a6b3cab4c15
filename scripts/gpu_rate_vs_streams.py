"""GPU: device-resident rate (chunks per ms) of one call against the number of streams, around the points where the engine changes
kernel mapping. python scripts/gpu_rate_vs_streams.py [N chunks]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(8)]
sizes = (1, 8, 16, 64, 74, 75, 100, 128, 148, 149, 200, 296, 297, 400, 512, 1023, 1024, 2048, 4096)
e = vadc_b200.Engine(max_streams=max(sizes))
for S in sizes:
    pcm = np.stack([base[s % 8] for s in range(S)])
    d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm)
    best = 1e9
    for _ in range(3):
        e.reset()
        e.timer_start()
        e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
        best = min(best, e.timer_stop())
    e.set_profiling(True)
    e.reset(); e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs); e.sync()
    ms, nl = e.last_timing()
    e.set_profiling(False)
    e.device_free(d_pcm); e.device_free(d_probs)
    print("S=%5d: %8.3f ms  %8.1f chunks/ms  stages stft %.3f enc %.3f lstm %.3f (%d launches)" % (S, best, S * N / best, ms["stft"], ms["layer1"] + ms["layer2"] + ms["layer3"] + ms["layer4"], ms["lstm0"] + ms["lstm1_decoder"], nl), flush=True)
