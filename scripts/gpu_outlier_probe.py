"""On-GPU: where does the largest |dp| of a long stream come from? Stream seed 50221 (the worst of scripts/gpu_parity_sweep.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle
N = 3000
pcm = vadc_b200.synth_pcm(50000 + 13 * 17, 1536 * N)
x = (pcm.astype(np.float32) / np.float32(32768)).reshape(N, 1536)
o = Oracle()
ref = np.zeros((N, 2), np.float32); states = {}
o.reset()
for n0 in range(0, N, 10):
    states[n0] = o.state.copy()
    ref[n0:n0 + 10] = o.run_chunks(x[n0:n0 + 10])
for mode in (1, 0):
    e = vadc_b200.Engine(max_streams=1, stft_mode=mode)
    got = e.run_streams(pcm[None, :], want_out2=True)[1][0]
    d = np.abs(got - ref).max(axis=1)
    w = int(d.argmax())
    print("stft_mode %d: worst |dp| %.2e at chunk %d; p_ref %s p_gpu %s; chunks with |dp| > 1e-4: %s" % (mode, d.max(), w, ref[w], got[w], np.nonzero(d > 1e-4)[0][:12]))
    print("   |dp| around it:", " ".join("%.1e" % v for v in d[w - 6:w + 6]))
    # restart 20 chunks before the worst chunk from the ORACLE's state: is the error local or accumulated?
    n0 = (w - 20) // 10 * 10
    st = states[n0]
    e.set_state(st[:128], st[128:])
    loc = e.run_chunks(x[n0:w + 5])
    dl = np.abs(loc - ref[n0:w + 5]).max(axis=1)
    print("   restarted from the oracle state at chunk %d: worst |dp| %.2e at chunk %d" % (n0, dl.max(), n0 + int(dl.argmax())))
    # stage errors on those chunks
    o.state[:] = st
    stg = o.run_stages(x[n0:w + 5])
    l = e.stage_pipeline(x[n0:w + 5])
    print("   encoder stage errors (production kernels) l1..l4:", " ".join("%.1e" % float(np.abs(a - stg[k]).max()) for a, k in zip(l, ("l1", "l2", "l3", "l4"))),
          "| l4 magnitude %.1f" % float(np.abs(stg["l4"]).max()))
    lo, hn, cn = e.stage_lstm(np.transpose(stg["l4"], (0, 2, 1)), h0=st[:128], c0=st[128:])
    print("   LSTM tap fed with the oracle's l4 and state: seq err %.1e" % float(np.abs(lo - stg["lstm"]).max()))
    e.close()
