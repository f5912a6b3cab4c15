"""CPU experiment (numpy): how much precision do the LSTM gate contractions need?
Runs the decoder LSTM + head on the oracle's exact encoder output with the gate GEMV evaluated in
(a) fp32, (b) single TF32, (c) 3xTF32 (hi/lo split, 3 products), (d) bf16x3 (6 products), and
reports max |p - oracle|. Justifies the precision choice of a tensor-core LSTM."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle, WEIGHTS
from testtensor_io import load_testtensor

def tf32(x, rz=False):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    if rz:
        return (u & np.uint32(0xFFFFE000)).view(np.float32)
    r = (u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)     # round-to-nearest (ties away), like cvt.rna.tf32
    return r.view(np.float32)

def bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> 16) & 1)) & np.uint32(0xFFFF0000)
    return r.view(np.float32)

def mm(W, v, mode):
    # W [256,128], v [128]; products in float64 then summed (models a wide accumulator), result fp32
    W64 = W.astype(np.float64)
    if mode == "fp32":
        return (W @ v).astype(np.float32)
    if mode == "tf32":
        return (tf32(W).astype(np.float64) @ tf32(v).astype(np.float64)).astype(np.float32)
    if mode in ("3xtf32", "2xtf32_wonly", "3xtf32_rz"):
        rz = mode.endswith("rz")
        Wh = tf32(W, rz); Wl = tf32(W - Wh, rz); vh = tf32(v, rz); vl = tf32(v - vh, rz)
        acc = Wh.astype(np.float64) @ vh.astype(np.float64) + Wl.astype(np.float64) @ vh.astype(np.float64)
        if mode != "2xtf32_wonly":
            acc = acc + Wh.astype(np.float64) @ vl.astype(np.float64)
        return acc.astype(np.float32)
    if mode in ("bf16x3", "bf16x2"):
        W1 = bf16(W); W2 = bf16(W - W1); W3 = bf16(W - W1 - W2)
        v1 = bf16(v); v2 = bf16(v - v1); v3 = bf16(v - v1 - v2)
        f = lambda a, b: a.astype(np.float64) @ b.astype(np.float64)
        if mode == "bf16x2":
            return (f(W1, v1) + f(W1, v2) + f(W2, v1)).astype(np.float32)
        return (f(W1, v1) + f(W1, v2) + f(W2, v1) + f(W1, v3) + f(W2, v2) + f(W3, v1)).astype(np.float32)
    raise ValueError(mode)

def sig(x): return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)

def run(l4, W, b, dw, db, mode):
    h = np.zeros((2, 64), np.float32); c = np.zeros((2, 64), np.float32)
    out = np.zeros(l4.shape[0], np.float32)
    for n in range(l4.shape[0]):
        acc = np.zeros(2, np.float32)
        for t in range(7):
            x = l4[n, :, t]
            for l in range(2):
                z = mm(W[l], np.concatenate([x, h[l]]), mode) + b[l]
                i, f, g, o = sig(z[:64]), sig(z[64:128]), np.tanh(z[128:192]), sig(z[192:])
                c[l] = f * c[l] + i * g
                h[l] = o * np.tanh(c[l])
                x = h[l]
            acc += dw @ np.maximum(x, 0)
        out[n] = sig(acc / np.float32(7) + db)[1]
    return out

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    w = list(load_testtensor(WEIGHTS).values())
    W, b, dw, db = w[95], w[96], w[97].reshape(2, 64), w[98]
    o = Oracle()
    for seed in (3, 11):
        pcm = vadc_b200.synth_pcm(seed, 1536 * N)
        x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
        o.reset()
        st = o.run_stages(x)
        ref = st["out"][:, 1]
        for mode in ("fp32", "3xtf32", "3xtf32_rz", "2xtf32_wonly", "bf16x3", "bf16x2", "tf32"):
            p = run(st["l4"], W, b, dw, db, mode)
            print("seed %d %-14s max|dp| = %.3e" % (seed, mode, np.abs(p - ref).max()), flush=True)

if __name__ == "__main__":
    main()
