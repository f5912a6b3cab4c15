"""On-GPU: worst |dp| vs the oracle over many long single streams (different seeds and signal kinds), default engine settings."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle
N = int(os.environ.get("NCHUNKS", "3000"))
S = int(os.environ.get("NSTREAMS", "32"))
pcm = np.stack([vadc_b200.synth_pcm(50000 + 13 * s, 1536 * N, kind=(s % 3 if s % 8 == 7 else 0)) for s in range(S)])
o = Oracle()
cache = os.path.join(ROOT, "scripts", "_sweep_refs.npy")        # optional: oracle outputs computed beforehand (scripts are run where GPU time is metered)
refs = np.load(cache) if os.path.exists(cache) and N == 3000 and S <= 32 else None
e = vadc_b200.Engine(max_streams=S, layer_mode=int(os.environ.get('LAYER_MODE', '0')), lstm_mode=int(os.environ.get('LSTM_MODE', '0')),
                     stft_mode=int(os.environ.get('STFT_MODE', '0')), stft_k_rel=float(os.environ.get('K_REL', '0')))
p, out2 = e.run_streams(pcm, want_out2=True)
errs, same = [], 0
for s in range(S):
    if refs is not None:
        ref = refs[s]
    else:
        o.reset(); ref = o.run_pcm(pcm[s])
    errs.append(float(np.abs(out2[s] - ref).max()))
    same += vadc_b200.segments_text(p[s]) == o.segments_text(ref[:, 1])
errs = np.array(errs)
print("streams %d x %d chunks: worst |dp| %.2e, median %.2e, > 1e-4: %d, segments identical %d/%d" % (S, N, errs.max(), np.median(errs), int((errs > 1e-4).sum()), same, S))
print(" ".join("%.1e" % v for v in errs))
