"""GPU: per-stage device times of one engine configuration on S streams x N chunks of device-resident PCM.
   python scripts/gpu_stage_times.py S N [layer_mode] [stft_mode] [lstm_mode]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kw = {}
if len(sys.argv) > 3: kw["layer_mode"] = int(sys.argv[3])
if len(sys.argv) > 4: kw["stft_mode"] = int(sys.argv[4])
if len(sys.argv) > 5: kw["lstm_mode"] = int(sys.argv[5])
e = vadc_b200.Engine(max_streams=S, **kw)
base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(8)]
pcm = np.stack([np.roll(base[s % 8].reshape(N, 1536), (s // 8) % N, axis=0).reshape(-1) for s in range(S)])
d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
e.h2d(d_pcm, pcm)
for prof in (False, True):
    e.set_profiling(prof)
    for it in range(3):
        e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
        e.sync()
    ms, n = e.last_timing()
    print("S=%d N=%d %s profiling=%d total %.3f ms, %d launches, %.3f M chunks/s = %.0f x realtime; stages %s" % (
        S, N, kw, prof, ms["total"], n, S * N / ms["total"] / 1e3, S * N * 0.096 / ms["total"] * 1e3,
        " ".join("%s=%.3f" % (k, v) for k, v in ms.items())), flush=True)
