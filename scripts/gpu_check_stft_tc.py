"""On-GPU check of the tensor-core STFT kernel: magnitudes vs the oracle (exact), flagged-bin statistics, timing."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

def main():
    o = Oracle()
    B = 41
    pcm = vadc_b200.synth_pcm(3, 1536 * B)
    x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
    st = o.run_stages(x)
    for name, mode in (("tensor", vadc_b200.STFT_HYBRID_TENSOR), ("fft", vadc_b200.STFT_HYBRID_FFT)):
        e = vadc_b200.Engine(max_streams=4096, stft_mode=mode)
        mag = e.stage_stft_magnitude(x)
        ref = st["stft"]
        rel = np.abs(mag - ref) / np.maximum(ref, 1e-30)
        tot, ex = e.stft_stats(reset=True)
        print(name, "magnitude: max abs diff %.3e  max rel diff %.3e  exact-equal bins %.1f %%  flagged %d of %d (%.3f %%)" % (
            np.abs(mag - ref).max(), rel.max(), 100.0 * np.mean(mag == ref), ex, tot, 100.0 * ex / max(tot, 1)), flush=True)
        frame_norm = np.sqrt((ref ** 2).sum(axis=1, keepdims=True))
        print("   worst |d|/||frame spectrum|| = %.3e" % (np.abs(mag - ref) / np.maximum(frame_norm, 1e-30)).max(), flush=True)
        norm, logmag = e.stage_stft_norm(x)
        print("   norm max diff %.3e" % np.abs(norm - st["norm"]).max(), flush=True)
        out = e.run_chunks(x)
        print("   run_chunks vs oracle %.3e" % np.abs(out - st["out"]).max(), flush=True)
        S, N = 4096, 20
        base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(4)]
        pcm2 = np.stack([np.roll(base[s % 4].reshape(N, 1536), (s // 4) % N, axis=0).reshape(-1) for s in range(S)])
        d_pcm = e.device_alloc(pcm2.nbytes); d_probs = e.device_alloc(S * N * 4)
        e.h2d(d_pcm, pcm2)
        e.set_profiling(1)
        for it in range(3):
            e.reset(); e.run_streams_device(d_pcm, pcm2.shape[1], S, N, d_probs); e.sync()
        tm, nl = e.last_timing()
        p = np.zeros((S, N), np.float32); e.d2h(p, d_probs)
        print("  ", {k: round(v, 3) for k, v in tm.items()}, "%.2f M chunks/s" % (S * N / tm["total"] / 1e3), flush=True)
        worst = 0
        for s in (0, 1, 2, 3, 4095):
            o.reset(); ref2 = o.run_pcm(pcm2[s]); worst = max(worst, float(np.abs(p[s] - ref2[:, 1]).max()))
        print("   multi-stream vs oracle max diff %.3e" % worst, flush=True)
        tot, ex = e.stft_stats(reset=True)
        print("   flagged %.3f %%" % (100.0 * ex / max(tot, 1)))
        e.close()
main()
