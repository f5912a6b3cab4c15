"""GPU: the exact path's encoder stage by stage against the oracle's stage dumps (bits)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 23
pcm = vadc_b200.synth_pcm(4242, 1536 * B)
x = (pcm.astype(np.float32) / np.float32(32768.0)).reshape(-1, 1536)
ref = Oracle().run_stages(x)
e = vadc_b200.Engine(max_streams=1, layer_mode=vadc_b200.LAYERS_FAITHFUL)
y1, l1, l2, l3, l4 = e.stage_exact_pipeline(x)
for name, got in (("l1", l1), ("l2", l2), ("l3", l3), ("l4", l4)):
    r = ref[name]
    d = got.view(np.uint32) != r.view(np.uint32)
    print("%s: differing %d of %d, max |d| %.3e, chunks with differences %s" % (name, int(d.sum()), d.size, float(np.abs(got - r).max()),
          sorted(set(np.argwhere(d)[:, 0].tolist()))[:12]))
    if d.any():
        i = tuple(np.argwhere(d)[0])
        print("   first at", i, "got", got[i], "ref", r[i])
print("y1 sample", y1[0, :2, :4])
e2 = vadc_b200.Engine(max_streams=1, layer_mode=vadc_b200.LAYERS_FP32, lstm_mode=vadc_b200.LSTM_FP32)
cb = e2.stage_layer_tap(0, 0, 1, ref["norm"])  # [B,T,C] conv_block output of the first layer on the fp32 kernels
cb = np.asarray(cb).reshape(B, 25, 16).transpose(0, 2, 1)
print("y1 vs fp32 conv_block tap: max |d| %.3e" % float(np.abs(cb - y1).max()), "worst at", np.unravel_index(np.argmax(np.abs(cb - y1)), cb.shape))
print(cb[0, :2, :4])
l1f = e2.stage_layer(0, ref["norm"])
print("fp32 l1 vs oracle: %.3e" % float(np.abs(np.asarray(l1f).reshape(ref["l1"].shape) - ref["l1"]).max()))
