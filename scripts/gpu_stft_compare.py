"""On-GPU: time the STFT kernels (modes 0 = fft8, 2 = warp-per-frame, 3 = tensor) on the bench-shaped window and print
their accuracy against the oracle on a long stream."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

S, N = 4096, 40
nb = 8
base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(nb)]
pcm2 = np.stack([np.roll(base[s % nb].reshape(N, 1536), (s // nb) % N, axis=0).reshape(-1) for s in range(S)])
long_pcm = vadc_b200.synth_pcm(4242, 1536 * 3000)
o = Oracle(); ref_long = o.run_pcm(long_pcm)
x_long = (long_pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)[:600]
o.reset(); st_long = o.run_stages(x_long)
for mode in [int(m) for m in os.environ.get('MODES', '0,2').split(',')]:
    e = vadc_b200.Engine(max_streams=S, stft_mode=mode, stft_k_rel=float(os.environ.get('K_REL', '0')))
    d_pcm = e.device_alloc(pcm2.nbytes); d_probs = e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm2)
    e.set_profiling(1)
    for it in range(3):
        e.reset(); e.run_streams_device(d_pcm, pcm2.shape[1], S, N, d_probs); e.sync()
    tm, nl = e.last_timing()
    e.set_profiling(0)
    for it in range(3):
        e.reset(); e.run_streams_device(d_pcm, pcm2.shape[1], S, N, d_probs); e.sync()
    tm0, _ = e.last_timing()
    e.reset(); e.stft_stats(reset=True)
    p, out2 = e.run_streams(long_pcm[None, :], want_out2=True)
    tot, ex = e.stft_stats(reset=True)
    nrm = e.stage_stft_norm(x_long)[0]; mg = e.stage_stft_magnitude(x_long)
    print("mode", mode, "norm max|d| %.2e mean|d| %.2e, mag max rel %.2e |" % (float(np.abs(nrm - st_long["norm"]).max()), float(np.abs(nrm - st_long["norm"]).mean()),
          float((np.abs(mg - st_long["stft"]) / np.maximum(st_long["stft"], 1e-30)).max())), end=" ")
    print( {k: round(v, 3) for k, v in tm.items()}, "unprofiled total %.3f ms -> %.2f M chunks/s" % (tm0["total"], S * N / tm0["total"] / 1e3),
          "| long stream max|dp| %.2e, exact bins %.3f %%" % (float(np.abs(out2[0] - ref_long).max()), 100.0 * ex / tot), flush=True)
    e.close()
