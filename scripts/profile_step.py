"""Small fixed workload for ncu: S streams x N chunks, `iters` calls (device-resident PCM)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
e = vadc_b200.Engine(max_streams=S, stft_mode=int(os.environ.get('STFT_MODE', '0')), layer_mode=int(os.environ.get('LAYER_MODE', '0')))
base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(8)]
pcm = np.stack([np.roll(base[s % 8].reshape(N, 1536), (s // 8) % N, axis=0).reshape(-1) for s in range(S)])
d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
e.h2d(d_pcm, pcm)
for it in range(iters):
    e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
    e.sync()
ms, n = e.last_timing()
print("S=%d N=%d total %.3f ms, %d launches, %.2f M chunks/s" % (S, N, ms["total"], n, S * N / ms["total"] / 1e3))
