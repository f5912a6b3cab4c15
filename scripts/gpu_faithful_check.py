"""GPU: quick look at the faithful path -- bit comparison with the oracle on a few shapes and its speed next to the fast paths.
Run under gpurun:  python scripts/gpu_faithful_check.py > gpurun_out/faithful.log"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def compare(tag, out2, pcm):
    worst, nbad = 0.0, 0
    for s in range(pcm.shape[0]):
        ref = Oracle().run_pcm(pcm[s])
        d = bits(out2[s]) != bits(ref)
        nbad += int(d.sum())
        worst = max(worst, float(np.abs(out2[s] - ref).max()))
        if d.any():
            first = int(np.argwhere(d.any(axis=1))[0, 0])
            print("   stream %d: first differing chunk %d of %d" % (s, first, len(ref)))
    print("%-44s differing values %d, max |d| %.3e" % (tag, nbad, worst), flush=True)
    return nbad


def timed(e, pcm, reps=2):
    e.run_streams(pcm)
    best = 1e9
    for _ in range(reps):
        e.reset()
        t = time.perf_counter()
        e.run_streams(pcm)
        best = min(best, time.perf_counter() - t)
    return best


def main():
    bad = 0
    pcm = vadc_b200.synth_pcm(4242, 1536 * 625)[None, :]
    e = vadc_b200.Engine()
    bad += compare("1 stream x 625 (cfg1), default engine", e.run_streams(pcm, want_out2=True)[1], pcm)
    e.close()
    pcm = np.stack([vadc_b200.synth_pcm(600 + 7 * s, 90 * 1536) for s in range(8)])
    e = vadc_b200.Engine(max_streams=8, window_chunks=17)
    bad += compare("8 streams x 90, windows of 17", e.run_streams(pcm, want_out2=True)[1], pcm)
    e.close()
    pcm = np.stack([vadc_b200.synth_pcm(50000 + 13 * i, 3000 * 1536) for i in (3, 11, 17, 23)])
    e = vadc_b200.Engine(max_streams=4)
    bad += compare("4 streams x 3000 (long silences)", e.run_streams(pcm, want_out2=True)[1], pcm)
    e.close()
    e = vadc_b200.Engine(max_streams=4, lstm_mode=vadc_b200.LSTM_FP32)
    compare("   same, fp32 fast kernels (for scale)", e.run_streams(pcm, want_out2=True)[1], pcm)
    e.close()
    # speed: x realtime = streams * chunks * 0.096 s / wall
    for S, N in ((1, 3000), (8, 1000), (64, 400)):
        pcm = np.stack([vadc_b200.synth_pcm(1 + s, N * 1536) for s in range(S)])
        row = []
        for name, kw in (("faithful", {}), ("fp32+exact stft", dict(lstm_mode=vadc_b200.LSTM_FP32, stft_mode=vadc_b200.STFT_EXACT)),
                         ("fast", dict(stft_mode=vadc_b200.STFT_HYBRID, lstm_mode=vadc_b200.LSTM_FP32))):
            e = vadc_b200.Engine(max_streams=S, **kw)
            t = timed(e, pcm)
            e.set_profiling(True)
            e.reset()
            e.run_streams(pcm)
            ms, _ = e.last_timing()
            e.close()
            row.append("%s %.0fx (%.1f ms; stages %s)" % (name, S * N * 0.096 / t, t * 1e3, " ".join("%.1f" % v for v in list(ms.values())[:9])))
        print("S=%d N=%d: " % (S, N) + " | ".join(row), flush=True)
    print("RESULT", "ok" if bad == 0 else "DIFFERENCES")


if __name__ == "__main__":
    main()
