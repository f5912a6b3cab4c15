"""On-GPU: worst |dp| against the oracle as a function of the hybrid STFT threshold k_rel, on several long single streams
(the LSTM carries STFT errors forward, so long streams are the hard case)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

N = int(os.environ.get("NCHUNKS", "3000"))
seeds = [4242, 7, 99, 1234, 31337, 2718]
pcm = np.stack([vadc_b200.synth_pcm(s, 1536 * N, kind=(2 if i == 5 else 0)) for i, s in enumerate(seeds)])
o = Oracle()
refs = []
for i in range(len(seeds)):
    o.reset(); refs.append(o.run_pcm(pcm[i]))
for k in [float(v) for v in os.environ.get("KS", "5e-4,1e-3,1.5e-3,2e-3,3e-3,4e-3").split(",")]:
    e = vadc_b200.Engine(max_streams=len(seeds), stft_k_rel=abs(k), stft_mode=(1 if k < 0 else int(os.environ.get('STFT_MODE', '0'))),
                         layer_mode=int(os.environ.get('LAYER_MODE', '0')), lstm_mode=int(os.environ.get('LSTM_MODE', '0')))   # k < 0: exact STFT
    e.stft_stats(reset=True)
    p, out2 = e.run_streams(pcm, want_out2=True)
    tot, ex = e.stft_stats(reset=True)
    errs = [float(np.abs(out2[i] - refs[i]).max()) for i in range(len(seeds))]
    segs = sum(vadc_b200.segments_text(p[i]) == o.segments_text(refs[i][:, 1]) for i in range(len(seeds)))
    marg = min(float(np.abs(refs[i][:, 1] - 0.5).min()) for i in range(len(seeds)))
    print("k_rel %.1e exact %.3f %%  max|dp| per stream %s  worst %.2e  segments identical %d/%d (closest p to 0.5: %.1e)" %
          (k, 100.0 * ex / max(tot, 1), " ".join("%.1e" % v for v in errs), max(errs), segs, len(seeds), marg), flush=True)
    e.close()
