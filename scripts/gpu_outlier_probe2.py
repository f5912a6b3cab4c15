"""On-GPU: exact-STFT (AUTO, small batch) path on the worst stream: is the normalized spectrogram bit-identical? where does the state drift?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle
N = 1850
pcm = vadc_b200.synth_pcm(50000 + 13 * 17, 1536 * 3000)[: N * 1536]
x = (pcm.astype(np.float32) / np.float32(32768)).reshape(N, 1536)
o = Oracle()
sl = slice(1500, 1840)
o.reset(); o.run_chunks(x[:1500]); st0 = o.state.copy()
stg = o.run_stages(x[sl]); st1 = o.state.copy()
e = vadc_b200.Engine(max_streams=1, stft_mode=1)
norm, logmag = e.stage_stft_norm(x[sl])
d = np.abs(norm - stg["norm"]).reshape(norm.shape[0], -1).max(axis=1)
print("norm: chunks with any difference %d of %d; max %.2e; typical nonzero %.2e" % (int((d > 0).sum()), d.size, d.max(), float(np.median(d[d > 0])) if (d > 0).any() else 0))
mag = e.stage_stft_magnitude(x[sl])
print("magnitudes bit-identical:", np.array_equal(mag, stg["stft"]))
l = e.stage_pipeline(x[sl])
for a, k in zip(l, ("l1", "l2", "l3", "l4")):
    dd = a - stg[k]
    print("  %s max|d| %.2e  mean d %.2e (bias)  mean|d| %.2e" % (k, float(np.abs(dd).max()), float(dd.mean()), float(np.abs(dd).mean())))
# state drift: run the GPU from chunk 1500 with the oracle's state up to 1840 and compare states
e.set_state(st0[:128], st0[128:])
e.run_chunks(x[sl])
h, c = e.get_state()
dh, dc = np.abs(h.reshape(-1) - st1[:128]), np.abs(c.reshape(-1) - st1[128:])
print("state after 340 chunks from the oracle's state: max|dh| %.2e max|dc| %.2e; |c| max %.1f; worst c unit %d (layer %d) c=%.3f" %
      (dh.max(), dc.max(), float(np.abs(st1[128:]).max()), int(dc.argmax()) % 64, int(dc.argmax()) // 64, float(st1[128:][dc.argmax()])))
# same with the LSTM tap fed the ORACLE's l4 sequence: isolates the LSTM kernel
lo, hn, cn = e.stage_lstm(np.transpose(stg["l4"], (0, 2, 1)), h0=st0[:128], c0=st0[128:])
print("LSTM tap on the oracle's l4: max|dc| %.2e max|dh| %.2e seq %.2e" % (float(np.abs(cn.reshape(-1) - st1[128:]).max()), float(np.abs(hn.reshape(-1) - st1[:128]).max()), float(np.abs(lo - stg["lstm"]).max())))
