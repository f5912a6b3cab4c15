"""CPU experiment (numpy): how much precision do the encoder's dense contractions need?
Re-evaluates the four transformer layers (from the oracle's exact normalized spectrogram) with every
1x1-conv / linear contraction computed as (a) fp32, (b) bf16x2 split (3 partial products: hi*hi + hi*lo
+ lo*hi, wide accumulation -- the tcgen05 scheme of vadc_b200/csrc/tc_common.cuh), (c) single bf16,
(d) fp16x2 split (same 3 products with fp16 terms: 22 significant bits),
then runs the exact fp32 LSTM + decoder on top and reports max |p - oracle| of the speech probability.
Depthwise conv, attention scores / AV, softmax, layer norm stay fp32 in every mode (they stay on the
CUDA cores in the kernel). Justifies the precision choice of the tensor-core layer kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle, WEIGHTS
from testtensor_io import load_testtensor


def bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> 16) & 1)) & np.uint32(0xFFFF0000)
    return r.view(np.float32)


def mm(x, W, mode):
    """x [..., K] @ W[N, K]^T"""
    if mode == "fp32":
        return (x @ W.T).astype(np.float32)
    f = lambda a, b: a.astype(np.float64) @ b.astype(np.float64).T
    if mode == "fp16x2":
        h = lambda a: np.asarray(a, np.float32).astype(np.float16).astype(np.float32)
        x1, W1 = h(x), h(W)
        x2, W2 = h(x - x1), h(W - W1)
        return (f(x1, W1) + f(x1, W2) + f(x2, W1)).astype(np.float32)
    x1, W1 = bf16(x), bf16(W)
    if mode == "bf16":
        return f(x1, W1).astype(np.float32)
    x2, W2 = bf16(x - x1), bf16(W - W1)
    if mode == "bf16x2":
        return (f(x1, W1) + f(x1, W2) + f(x2, W1)).astype(np.float32)
    raise ValueError(mode)


def layer_norm(x, w, b):
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    return ((x - m) / np.sqrt(v + np.float32(1e-5)) * w + b).astype(np.float32)


def layer(x, w, stride, proj, mode):
    """x [B, Cin, T] -> [B, C, Tout] (transformer.c:237-295)"""
    i = 0
    def nxt():
        nonlocal i
        i += 1
        return w[i - 1]
    dw_w, dw_b, pw_w, pw_b = nxt(), nxt(), nxt(), nxt()
    if proj:
        pj_w, pj_b = nxt(), nxt()
    qkv_w, qkv_b, ao_w, ao_b, n1w, n1b, f1w, f1b, f2w, f2b, n2w, n2b, cv_w, cv_b, bn_w, bn_b, bn_m, bn_v = [nxt() for _ in range(18)]
    B, Cin, T = x.shape
    C = pw_w.shape[0]
    xp = np.pad(x, ((0, 0), (0, 0), (2, 2)))
    d = sum(xp[:, :, k:k + T] * dw_w[:, 0, k][None, :, None] for k in range(5)) + dw_b[None, :, None]
    d = np.maximum(d, 0).astype(np.float32)
    xt, dt = x.transpose(0, 2, 1), d.transpose(0, 2, 1)                 # [B, T, Cin]
    y = mm(dt, pw_w[:, :, 0], mode) + pw_b
    y = y + (mm(xt, pj_w[:, :, 0], mode) + pj_b if proj else xt)
    u = np.maximum(y, 0).astype(np.float32)                            # [B, T, C]
    qkv = mm(u, qkv_w, mode) + qkv_b
    D = C // 2
    o = np.zeros_like(u)
    for h in range(2):
        q, k, v = (qkv[:, :, p * C + h * D: p * C + (h + 1) * D] for p in range(3))
        s = np.einsum("bkd,bqd->bkq", k, q) * np.float32(1.0 / np.sqrt(D))   # rows = K positions (transformer.c:101-120)
        s = np.exp(s - s.max(-1, keepdims=True))
        a = s / s.sum(-1, keepdims=True)
        o[:, :, h * D:(h + 1) * D] = np.einsum("bkq,bqd->bkd", a, v)
    a = mm(o, ao_w, mode) + ao_b
    u1 = layer_norm(u + a, n1w, n1b)
    f = mm(np.maximum(mm(u1, f1w, mode) + f1b, 0), f2w, mode) + f2b
    u2 = layer_norm(u1 + f, n2w, n2b)
    z = mm(u2[:, ::stride], cv_w[:, :, 0], mode) + cv_b
    z = (z - bn_m) / np.sqrt(bn_v + np.float32(1e-5)) * bn_w + bn_b
    return np.maximum(z, 0).astype(np.float32).transpose(0, 2, 1)


def sig(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def lstm_decoder(l4, W, b, dw, db):
    h = np.zeros((2, 64), np.float32); c = np.zeros((2, 64), np.float32)
    out = np.zeros(l4.shape[0], np.float32)
    for n in range(l4.shape[0]):
        acc = np.zeros(2, np.float32)
        for t in range(7):
            x = l4[n, :, t]
            for l in range(2):
                z = W[l] @ np.concatenate([x, h[l]]) + b[l]
                i, f, g, o = sig(z[:64]), sig(z[64:128]), np.tanh(z[128:192]), sig(z[192:])
                c[l] = f * c[l] + i * g
                h[l] = o * np.tanh(c[l])
                x = h[l]
            acc += dw @ np.maximum(x, 0)
        out[n] = sig(acc / np.float32(7) + db)[1]
    return out


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    w = list(load_testtensor(WEIGHTS).values())
    spec = ((1, 25, 2, True), (25, 49, 2, True), (49, 71, 1, False), (71, 95, 1, True))
    o = Oracle()
    for seed in (3, 11, 29):
        pcm = vadc_b200.synth_pcm(seed, 1536 * N)
        x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
        o.reset()
        st = o.run_stages(x)
        ref = st["out"][:, 1]
        for mode in ("fp32", "fp16x2", "bf16x2", "bf16"):
            for which in ("all", "l2-4"):
                a = st["norm"]
                errs = []
                for li, (lo, hi, stride, proj) in enumerate(spec):
                    m = mode if (which == "all" or li > 0) else "fp32"
                    a = layer(a, w[lo:hi], stride, proj, m)
                    errs.append(float(np.abs(a - st["l%d" % (li + 1)]).max()))
                p = lstm_decoder(a, w[95], w[96], w[97].reshape(2, 64), w[98])
                print("seed %2d %-7s %-5s layer max err %s   max|dp| = %.3e" % (seed, mode, which, " ".join("%.1e" % e for e in errs), np.abs(p - ref).max()), flush=True)
                if mode == "fp32":
                    break


if __name__ == "__main__":
    main()
