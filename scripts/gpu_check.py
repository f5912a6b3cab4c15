"""Quick on-GPU sanity run: per-stage parity vs the oracle restatement + a first timing."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200
from oracle_lib import Oracle

def md(a, b): return float(np.abs(a - b).max())

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    mode = int(os.environ.get("STFT_MODE", "0"))
    e = vadc_b200.Engine(max_streams=max(S, 64), stft_mode=mode, stft_k_rel=float(os.environ.get("K_REL", "0")))
    print(e.info())
    o = Oracle()
    pcm = vadc_b200.synth_pcm(3, 1536 * 40)
    x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
    st = o.run_stages(x)
    mag = e.stage_stft_magnitude(x)
    print("stft magnitude max diff", md(mag, st["stft"]), "bit-exact", np.array_equal(mag, st["stft"]), "max rel diff", float((np.abs(mag - st["stft"]) / np.maximum(st["stft"], 1e-30)).max()))
    print("stft stats (total, exact)", e.stft_stats(reset=True))
    norm, logmag = e.stage_stft_norm(x)
    print("norm max diff", md(norm, st["norm"]))
    print("stage_norm(mag) diff", md(e.stage_norm(st["stft"]), st["norm"]))
    l = e.stage_encoder(st["norm"])
    for i, k in enumerate(("l1", "l2", "l3", "l4")): print("encoder", k, md(l[i], st[k]))
    l = e.stage_pipeline(x)
    for i, k in enumerate(("l1", "l2", "l3", "l4")): print("pipeline", k, md(l[i], st[k]))
    for i, (k_in, k_out) in enumerate((("norm", "l1"), ("l1", "l2"), ("l2", "l3"), ("l3", "l4"))):
        print("layer", i, md(e.stage_layer(i, st[k_in]), st[k_out]))
    lo, hn, cn = e.stage_lstm(np.transpose(st["l4"], (0, 2, 1)))
    print("lstm seq", md(lo, st["lstm"]), "hn", md(hn.reshape(-1), o.state[:128]), "cn", md(cn.reshape(-1), o.state[128:]))
    print("decoder", md(e.stage_decoder(np.transpose(st["lstm"], (0, 2, 1))), st["out"]))
    out = e.run_chunks(x)
    print("run_chunks vs oracle", md(out, st["out"]), out[:3, 1], st["out"][:3, 1])
    # multi-stream: S streams built from a few base streams with different chunk shifts
    nb = 4
    base = [vadc_b200.synth_pcm(100 + i, 1536 * N) for i in range(nb)]
    pcm2 = np.stack([np.roll(base[s % nb].reshape(N, 1536), (s // nb) % N, axis=0).reshape(-1) for s in range(S)])
    e.reset()
    t0 = time.time()
    probs, out2 = e.run_streams(pcm2, want_out2=True)
    dt = time.time() - t0
    tm, nl = e.last_timing()
    print("run_streams S=%d N=%d: wall %.3fs device %.3f ms launches %d -> %.0f x realtime" % (S, N, dt, tm["total"], nl, S * N * 0.096 / (tm["total"] / 1e3)))
    worst = 0
    for s in list(range(0, min(S, 8))) + [S - 1]:
        o.reset()
        ref = o.run_pcm(pcm2[s])
        worst = max(worst, md(out2[s], ref))
    print("multi-stream vs oracle max diff", worst)
    tot, ex = e.stft_stats(reset=True)
    print("hybrid stft: %d bins, %d exact (%.3f %%)" % (tot, ex, 100.0 * ex / max(tot, 1)))
    # device-resident timing with per-stage profile
    d_pcm = e.device_alloc(pcm2.nbytes); d_probs = e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm2)
    for prof in (0, 1):
        e.set_profiling(prof)
        for it in range(3):
            e.reset()
            e.run_streams_device(d_pcm, pcm2.shape[1], S, N, d_probs); e.sync()
        tm, nl = e.last_timing()
        print("device-resident prof=%d:" % prof, {k: round(v, 3) for k, v in tm.items()}, "launches", nl,
              "-> %.0f x realtime, %.2f M chunks/s" % (S * N * 0.096 / (tm["total"] / 1e3), S * N / tm["total"] / 1e3))
    p2 = np.zeros((S, N), np.float32); e.d2h(p2, d_probs)
    print("device path == host path:", np.array_equal(p2, probs))

main()
