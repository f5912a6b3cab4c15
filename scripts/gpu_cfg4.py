"""BASELINE cfg4 for real: ONE 10-hour synthetic recording (375 000 chunks, 1.15 GB of s16le) through the native CLI on the GPU and through
the unmodified reference CLI (oracle/_ref/vadc_linux) on a host core: stdout must be byte-identical (timestamps bit-exact), both wall
times are recorded.   python scripts/gpu_cfg4.py [hours] > gpurun_out/cfg4.json
The recording is NOT split into segments: the decoder LSTM's state carries across the whole file (a warm-up overlap of even
98 s leaves 6.5e-3 of error), so the scan is serial and bit-exact; the stateless front end (STFT + encoder) is what the GPU batches."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vadc_b200

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
nchunks = int(hours * 3600 * 16000 / 1536)
path = "/tmp/cfg4_%dh.s16le" % int(hours)
t = time.perf_counter()
PIECE = 37500                                      # one hour per piece, different seeds: speech bursts, pauses and long silences
with open(path, "wb") as f:
    done = 0
    while done < nchunks:
        n = min(PIECE, nchunks - done)
        vadc_b200.synth_pcm(777000 + done // PIECE, n * 1536).tofile(f)
        done += n
t_synth = time.perf_counter() - t
cli = os.path.join(ROOT, "vadc_b200", "vadc_b200_cli")
ref = os.path.join(ROOT, "oracle", "_ref", "vadc_linux")
res = {"workload": "cfg4: one %.1f-hour synthetic 16 kHz s16le recording = %d chunks, stdin -> segment timestamps on stdout" % (hours, nchunks),
       "bytes": os.path.getsize(path), "synth_seconds": t_synth}
outs = {}
for name, exe, args in (("b200_cli", cli, []), ("b200_cli_batch_1536", cli, ["--batch", "1536"]), ("reference_cli", ref, [])):
    if not os.path.exists(exe):
        res[name] = {"error": "missing " + exe}
        continue
    with open(path, "rb") as f:
        t = time.perf_counter()
        r = subprocess.run([exe] + args, stdin=f, capture_output=True)
        dt = time.perf_counter() - t
    outs[name] = r.stdout
    res[name] = {"seconds": dt, "x_realtime": nchunks * 0.096 / dt, "segments": r.stdout.count(b"\n"), "returncode": r.returncode}
if "reference_cli" in outs:
    for k in ("b200_cli", "b200_cli_batch_1536"):
        if k in outs:
            res[k]["stdout_identical_to_reference"] = outs[k] == outs["reference_cli"]
os.remove(path)
print(json.dumps(res))
