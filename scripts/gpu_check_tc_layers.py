"""On-GPU check of the tensor-core layer kernels: per-layer error vs the oracle and timing vs the FP32 kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

def main():
    o = Oracle()
    B = 40
    pcm = vadc_b200.synth_pcm(3, 1536 * B)
    x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
    st = o.run_stages(x)
    et = vadc_b200.Engine(max_streams=4096, layer_mode=vadc_b200.LAYERS_TENSOR)
    ef = vadc_b200.Engine(max_streams=4096, layer_mode=vadc_b200.LAYERS_FP32)
    for layer, (ki, ko) in enumerate((("norm", "l1"), ("l1", "l2"), ("l2", "l3"), ("l3", "l4"))):
        gt, gf = et.stage_layer(layer, st[ki]), ef.stage_layer(layer, st[ko if False else ki])
        print("layer", layer, "tensor err %.3e  fp32 err %.3e  scale %.2f" % (np.abs(gt - st[ko]).max(), np.abs(gf - st[ko]).max(), np.abs(st[ko]).max()), flush=True)
    S, N = 4096, 20
    base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(4)]
    pcm2 = np.stack([np.roll(base[s % 4].reshape(N, 1536), (s // 4) % N, axis=0).reshape(-1) for s in range(S)])
    for name, e in (("tensor", et), ("fp32", ef)):
        d_pcm = e.device_alloc(pcm2.nbytes); d_probs = e.device_alloc(S * N * 4)
        e.h2d(d_pcm, pcm2)
        e.set_profiling(1)
        for it in range(3):
            e.reset(); e.run_streams_device(d_pcm, pcm2.shape[1], S, N, d_probs); e.sync()
        tm, nl = e.last_timing()
        p = np.zeros((S, N), np.float32); e.d2h(p, d_probs)
        print(name, {k: round(v, 3) for k, v in tm.items()}, "%.2f M chunks/s" % (S * N / tm["total"] / 1e3), flush=True)
        if name == "tensor": pt = p
        else: print("max |p_tensor - p_fp32| =", float(np.abs(pt - p).max()))
    worst = 0
    for s in (0, 1, 2, 3, 4095):
        o.reset(); ref = o.run_pcm(pcm2[s]); worst = max(worst, float(np.abs(pt[s] - ref[:, 1]).max()))
    print("tensor path vs oracle max diff", worst)
main()
