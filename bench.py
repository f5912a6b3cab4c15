#!/usr/bin/env python
"""bench.py -- audio-seconds/sec (x realtime) of Silero VAD v3.1 (16 kHz) on B200, beside the
reference C backend on the host CPU.   python bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): 4096 concurrent synthetic 16 kHz s16le streams per GPU with
per-stream LSTM state kept on device. A "step" is one pass of the hot path over 4096 streams x 125
chunks (12 s of audio each, 512 000 chunks, 1.57 GB of PCM); 50 steps are the full 10 minutes. Under
torchrun every rank owns its own 4096 streams (weak scaling, no data-path collective).

value : device-resident throughput (PCM already in HBM), CUDA events on the engine's stream, max over ranks.
e2e   : the same steps through silero_b200_run_streams with pinned HOST buffers (H2D of the PCM and
        D2H of the probabilities inside the timed region).
--impl reference : the reference's own C backend (oracle/_ref, built from /root/reference; else the
        oracle port) on the host cores, one process per core.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHUNK = 1536
CHUNK_SECONDS = CHUNK / 16000.0
FLOP_PER_CHUNK = 5404954          # 2 x 2 702 477 MAC, reference's dense formulation
STFT_FLOP_PER_CHUNK = 2 * 1651200 # K1: 258 x 25 x 256 MAC
# algorithmic MACs per chunk of every stage on the reference's formulation (sums to 2 702 477)
STAGE_MAC = {"stft": 1651200, "layer1": 181053, "layer2": 112208, "layer3": 61600, "layer4": 236768,
             "lstm0": 229376, "lstm1_decoder": 229376 + 896}
STAGE_KERNEL = {"stft": "stft_fft8_kernel<s16> (fp32 FFT, 8 lanes per frame, + exact fix-up)", "layer1": "layer0_tc_kernel (tcgen05 fp16x2)", "layer2": "layer_tc_kernel<1> (tcgen05 fp16x2)",
                "layer3": "layer_tc_kernel<2> (tcgen05 fp16x2)", "layer4": "layer_tc_kernel<3> (tcgen05 fp16x2)",
                "lstm0": "lstm_tc_kernel<0> (tcgen05 bf16x2)", "lstm1_decoder": "lstm_tc_kernel<1> (tcgen05 bf16x2, +decoder)"}
# FP32 operations the STFT kernel actually EXECUTES per chunk (2*FFMA + FADD + FMUL thread instructions from the committed ncu
# capture profiles/ncu_summary_r01h.md: 6585 flop/cycle x 2.212e6 cycles / 81920 chunks): it evaluates the reference's dense
# 258x256 correlation (3.30 MFLOP/chunk algorithmic) as a 256-point real FFT plus exact re-evaluation of ~0.5 % of the bins.
STFT_EXECUTED_FLOP_PER_CHUNK = 178e3
STREAMS_PER_GPU = 4096
STEP_CHUNKS = 125
N_BASE = 32                       # distinct synthetic base streams
BASE_CHUNKS = 1250                # 120 s each; streams are chunk-rotated views of the bases


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference C backend, one process per core (it is not thread-safe: conv.c:172)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, seeds, nsamples = args
    import vadc_b200
    from oracle_lib import Oracle, Reference
    impl = Reference() if kind == "reference" else Oracle()
    pcms = [vadc_b200.synth_pcm(s, nsamples) for s in seeds]
    impl.run_pcm(pcms[0][: CHUNK * 20])  # warm caches / page in
    impl.reset()
    t0 = time.perf_counter()
    n = 0
    for p in pcms:
        impl.reset()
        out = impl.run_pcm(p)
        n += out.shape[0]
    return n, time.perf_counter() - t0


def cpu_reference_run(streams_per_core=2, seconds=60.0):
    """Times the reference backend on every host core. Returns (audio_s_per_s, cores, kind, sample, per_core)."""
    from oracle_lib import have_ref
    kind = "reference" if have_ref() else "port"
    cores = len(os.sched_getaffinity(0))
    nsamples = int(seconds * 16000)
    jobs = [(kind, [90000 + c * streams_per_core + i for i in range(streams_per_core)], nsamples) for c in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(kind, [1], CHUNK * 8)] * cores)   # spawn + import cost outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    chunks = sum(r[0] for r in res)
    per_core = float(np.mean([r[0] * CHUNK_SECONDS / r[1] for r in res]))
    sample = "%d procs x %d streams x %.0f s synthetic bursts, batch 96, %s" % (
        cores, streams_per_core, seconds, "oracle/_ref (unmodified reference, gcc -O2 -mavx2 -ffp-contract=off)" if kind == "reference"
        else "oracle/libsilero_oracle.so (C port)")
    return chunks * CHUNK_SECONDS / wall, cores, kind, sample, per_core


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_rows, self.t0, self.t1 = [], None, None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def in_region(self):
        return sum(1 for t in self.t_rows if self.t0 is not None and t >= self.t0 and (self.t1 is None or t <= self.t1))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            self.t_rows.append(time.perf_counter())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        # only samples taken while the GPU was under the bench load: the timed region, extended (by the caller, mark_end) over identical
        # untimed steps when the region itself is shorter than a few sampling periods
        rows = [r for r, t in zip(self.rows, self.t_rows) if self.t0 is None or (t >= self.t0 and (self.t1 is None or t <= self.t1))]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def build_step_inputs(torch, dev, nbuf):
    """PCM of `nbuf` steps for this rank's 4096 streams, resident in HBM: [nbuf][S][STEP_CHUNKS*1536] s16."""
    import vadc_b200
    rank = int(os.environ.get("RANK", "0"))
    base = np.stack([vadc_b200.synth_pcm(5000 + 97 * rank + i, BASE_CHUNKS * CHUNK) for i in range(N_BASE)])
    d_base = torch.from_numpy(base).to(dev).view(N_BASE, BASE_CHUNKS, CHUNK)
    s = torch.arange(STREAMS_PER_GPU, device=dev)
    sid, off = s % N_BASE, (s // N_BASE) * 37
    n = torch.arange(STEP_CHUNKS, device=dev)
    bufs = []
    for k in range(nbuf):
        idx = (off[:, None] + k * STEP_CHUNKS + n[None, :]) % BASE_CHUNKS
        bufs.append(d_base[sid[:, None], idx].contiguous().view(STREAMS_PER_GPU, STEP_CHUNKS * CHUNK))
    torch.cuda.synchronize()
    return base, bufs


def host_stream(base, s, k0, nchunks):
    """The s16 stream that rank-0 stream `s` sees from step k0 on (for the oracle spot check)."""
    sid, off = s % N_BASE, (s // N_BASE) * 37
    idx = (off + k0 * STEP_CHUNKS + np.arange(nchunks)) % BASE_CHUNKS
    return base[sid].reshape(BASE_CHUNKS, CHUNK)[idx].reshape(-1)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    vals = []
    for _ in range(args.warmup):
        cpu_reference_run(1, 10.0)
    t_total, audio_total = 0.0, 0.0
    last = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        last = cpu_reference_run(2, 60.0)
        dt = time.perf_counter() - t0
        vals.append(last[0])
    value = float(np.mean(vals))
    _, cores, kind, sample, per_core = last
    line = {
        "impl": "reference", "metric": "audio-seconds/sec (RTF) Silero v3.1 at 1/2/4/8 B200 vs host-CPU C backend",
        "value": value, "unit": "audio-seconds/sec", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cores * 2 * 60.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg3: 4096 concurrent 16 kHz s16le streams x 10 min per GPU (reference arm: bounded sample per step, see cpu_baseline.sample)",
                   "cpu": cpu_model()},
        "cpu_baseline": {"value": value, "unit": "audio-seconds/sec", "cores": cores, "kind": kind, "sample": sample, "per_core": per_core},
        "e2e": {"value": value, "unit": "audio-seconds/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist

    import vadc_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S, C = STREAMS_PER_GPU, STEP_CHUNKS
    eng = vadc_b200.Engine(device=local, max_streams=S)
    nbuf = min(args.steps + args.warmup, 6)
    base, bufs = build_step_inputs(torch, dev, nbuf)
    d_probs = torch.empty((S, C), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def step(k):
        eng.run_streams_device(bufs[k % nbuf].data_ptr(), C * CHUNK, S, C, d_probs.data_ptr())

    # ---- parity spot check on this exact workload (rank 0, not timed) ---------------------------
    parity = None
    if rank == 0:
        from oracle_lib import Oracle
        step(0)
        eng.sync()
        got = d_probs.cpu().numpy()
        o = Oracle()
        worst = 0.0
        for s in (0, 1337, S - 1):
            o.reset()
            worst = max(worst, float(np.abs(got[s] - o.run_pcm(host_stream(base, s, 0, C))[:, 1]).max()))
        parity = worst
        assert worst <= 1e-4, "parity lost on the bench workload: %g" % worst
    eng.reset()

    # ---- device-resident timed region ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                # nvidia-smi needs a few hundred ms to come up: start it before the warm-up
    for k in range(args.warmup):
        step(k)
    eng.sync()
    barrier()
    sampler.mark_begin()
    eng.timer_start()
    for k in range(args.steps):
        step(args.warmup + k)
    ms = eng.timer_stop()
    barrier()
    launches = args.steps * eng.last_timing()[1]   # kernels launched per run_streams_device call, counted by the engine
    clock_note = "sampled during the timed region"
    if rank == 0 and sampler.proc is not None and sampler.in_region() < 5:
        # the timed region (steps x ~20 ms) can be shorter than a handful of 100 ms sampling periods: keep the SAME load running,
        # untimed, until the sampler has seen it (the clocks line is about the state of the GPU under this workload)
        t_end = time.perf_counter() + 3.0
        k = args.warmup + args.steps
        while sampler.in_region() < 5 and time.perf_counter() < t_end:
            step(k)
            eng.sync()
            k += 1
        clock_note = "sampled during the timed region and %d identical untimed steps right after it" % (k - args.warmup - args.steps)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["how"] = clock_note
    barrier()
    ms = max_over_ranks(ms)
    audio_s = world * S * C * CHUNK_SECONDS * args.steps
    value = audio_s / (ms / 1e3)

    # ---- per-kernel profile of the same step (CUDA events between kernels; separate pass) -----------
    eng.set_profiling(True)
    stage_ms = None
    for k in range(2):
        step(k)
        eng.sync()
        stage_ms, n_launch = eng.last_timing()
    eng.set_profiling(False)
    windows = n_launch // 7
    chunks_per_launch = S * C / windows
    fp32_peak = eng.measure_fp32_peak()
    kernel_sum = sum(v for k, v in stage_ms.items() if k != "total")
    top = max((k for k in stage_ms if k != "total"), key=lambda k: stage_ms[k])
    top_ms_per_launch = stage_ms[top] / windows
    top_tflops = 2 * STAGE_MAC[top] * chunks_per_launch / (top_ms_per_launch * 1e-3) / 1e12
    per_stage = {k: {"ms_per_launch": stage_ms[k] / windows, "share": stage_ms[k] / kernel_sum,
                     "algorithmic_tflops": 2 * STAGE_MAC[k] * chunks_per_launch / (stage_ms[k] / windows * 1e-3) / 1e12}
                 for k in stage_ms if k != "total"}
    bins_total, bins_exact = eng.stft_stats()
    # the two rooflines of the contract, from the driver-written measured peaks (else the profiling recipe's fallbacks), for the same
    # dominant kernel: they show why neither bounds it
    peaks = {"hbm_gbs": 6500.0, "bf16_tflops": 1600.0, "source": "B200_PROFILING.md fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp_ = json.load(fh)
        peaks = {"hbm_gbs": float(mp_["hbm_gbs"]), "bf16_tflops": float(mp_.get("bf16_tflops_sustained", mp_["bf16_tflops"])), "source": "MEASURED_PEAKS.json"}
    except Exception:
        pass
    alg_bytes = ALGORITHMIC_BYTES_PER_CHUNK.get(top)
    hbm_view = None
    if alg_bytes:
        gbs = alg_bytes * chunks_per_launch / (top_ms_per_launch * 1e-3) / 1e9
        hbm_view = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                    "algorithmic_bytes_per_launch": alg_bytes * chunks_per_launch,
                    "traffic": TRAFFIC_BYTES_PER_CHUNK[top] * chunks_per_launch if TRAFFIC_BYTES_PER_CHUNK.get(top) else None,
                    "peak_source": peaks["source"]}
    # tensor-pipe view of the kernels that run on tcgen05: algorithmic rate, and x3 for the three partial products of the fp16/bf16 splits
    tensor_view = {k: {"algorithmic_tflops": per_stage[k]["algorithmic_tflops"], "issued_tflops": 3 * per_stage[k]["algorithmic_tflops"],
                       "frac_of_bf16_peak": 3 * per_stage[k]["algorithmic_tflops"] / peaks["bf16_tflops"]}
                   for k in ("layer1", "layer2", "layer3", "layer4", "lstm0", "lstm1_decoder") if k in per_stage}
    roofline = {
        "kernel": STAGE_KERNEL[top], "bound": "fp32",
        "achieved": top_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": top_tflops / fp32_peak,
        "peak_source": "FP32 FMA pipe measured live by silero_b200_measure_fp32_peak (independent FFMA chains); MEASURED_PEAKS.json has no FP32 figure "
                       "and its HBM / bf16-tensor peaks do not bound these kernels (DESIGN.md sections 2, 4)",
        "algorithmic_flop_per_launch": 2 * STAGE_MAC[top] * chunks_per_launch, "ms_per_launch": top_ms_per_launch,
        "share_of_step": stage_ms[top] / kernel_sum,
        "traffic": TRAFFIC_BYTES_PER_CHUNK.get(top, 0) * chunks_per_launch if TRAFFIC_BYTES_PER_CHUNK.get(top) else None,
        "stages": per_stage,
        "hbm": hbm_view,
        "tensor": {"bound": "tensor", "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "peak_source": peaks["source"], "kernels": tensor_view},
        "note": "achieved = ALGORITHMIC FLOPs of the reference's dense formulation (SURVEY.md 8d) / measured launch time; the STFT kernel is an FFT + "
                "exact fix-up, so the algorithmic rate exceeds the FP32 pipe peak (frac > 1); executed_* is what the FP32 pipe really did",
        "executed_tflops": (STFT_EXECUTED_FLOP_PER_CHUNK * chunks_per_launch / (stage_ms["stft"] / windows * 1e-3) / 1e12) if top == "stft" else None,
        "executed_frac": (STFT_EXECUTED_FLOP_PER_CHUNK * chunks_per_launch / (stage_ms["stft"] / windows * 1e-3) / 1e12 / fp32_peak) if top == "stft" else None,
        "stft_exact_bin_fraction": bins_exact / max(bins_total, 1),
        "pipeline_algorithmic_tflops": FLOP_PER_CHUNK * (value / world / CHUNK_SECONDS) / 1e12,
        "pipeline_frac_of_fp32_peak": FLOP_PER_CHUNK * (value / world / CHUNK_SECONDS) / 1e12 / fp32_peak,
    }

    # ---- end to end through the C ABI with host buffers --------------------------------------------
    h_pcm, h_pcm_ptr = vadc_b200.pinned_empty((S, C * CHUNK), np.int16)
    SEG_CAP = C // 2 + 2
    outs = []                                              # two sets of pinned output buffers: step k+1 is submitted before step k is read
    for _ in range(2):
        hp, hp_ptr = vadc_b200.pinned_empty((S, C), np.float32)
        hs, hs_ptr = vadc_b200.pinned_empty((S, SEG_CAP, 2), np.int32)
        hc, hc_ptr = vadc_b200.pinned_empty((S,), np.int32)
        outs.append((hp, hp_ptr, hs, hs_ptr, hc, hc_ptr))
    h_pcm[:] = bufs[0].cpu().numpy()
    eng.reset()
    eng.segments_configure()

    def e2e_submit(k, last):
        # asynchronous public call: pinned host PCM in; probabilities AND finished speech segments (on-device segmenter) out
        hp, hp_ptr, hs, hs_ptr, hc, hc_ptr = outs[k & 1]
        return eng.submit_streams_segments_ptr(h_pcm_ptr, C * CHUNK, S, C, last, hs_ptr, SEG_CAP, hc_ptr, hp_ptr)

    for k in range(2):
        eng.wait(e2e_submit(k, False))
    eng.reset()
    eng.segments_reset()
    seg_log = []

    def collect(k, ticket):
        eng.wait(ticket)                                       # results of step k are on the host from here on
        hp, _, hs, _, hc, _ = outs[k & 1]
        m = int(hc.max())
        seg_log.append((hc.copy(), hs[:, :m].copy()))

    barrier()
    t0 = time.perf_counter()
    prev = None
    for k in range(args.steps):                                # 2-deep pipeline: the H2D of step k overlaps the compute of step k-1
        t = e2e_submit(k, k == args.steps - 1)
        if prev is not None:
            collect(k - 1, prev)
        prev = t
    collect(args.steps - 1, prev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    h_probs = outs[(args.steps - 1) & 1][0]
    my_segments = [[] for _ in range(S)]
    for counts, segs in seg_log:
        if counts.max() > SEG_CAP:
            raise SystemExit("segment capacity exceeded")
        for s in np.nonzero(counts)[0]:
            my_segments[s] += [tuple(p) for p in segs[s, :counts[s]].tolist()]
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = audio_s / (e2e_ms / 1e3)
    # the "final gather of per-stream segments" (not timed): rank 0 receives every stream's (start, end) pairs
    from vadc_b200 import shard
    gathered = shard.gather_segments(my_segments, rank * S, world * S)
    seg_count = sum(len(x) for x in gathered) if rank == 0 else 0

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, kind, sample, per_core = cpu_reference_run(2, 60.0)
        cpu = {"value": v, "unit": "audio-seconds/sec", "cores": cores, "kind": kind, "sample": sample, "per_core": per_core, "cpu": cpu_model()}

    # ---- BASELINE configs[0] beside it (rank 0, N=1, not part of `value`): ONE 60 s stream through the public host call -- the way
    # the reference itself runs. Small batches take the faithful kernels (faithful_kernel.cuh): bits compared with the oracle.
    cfg1 = None
    if rank == 0 and world == 1:
        try:
            from oracle_lib import Oracle
            one = vadc_b200.Engine(device=local, max_streams=1)
            pcm1 = vadc_b200.synth_pcm(4242, 625 * CHUNK)[None, :]
            got1 = one.run_streams(pcm1, want_out2=True)[1][0]
            best = 1e30
            for _ in range(3):
                one.reset()
                t1 = time.perf_counter()
                one.run_streams(pcm1)
                best = min(best, time.perf_counter() - t1)
            one.close()
            ref1 = Oracle().run_pcm(pcm1[0])
            cfg1 = {"workload": "cfg1: one 60 s stream (625 chunks), silero_b200_run_streams with host PCM in, probabilities out",
                    "value": 625 * CHUNK_SECONDS / best, "unit": "audio-seconds/sec", "ms": best * 1e3,
                    "bit_identical_to_oracle": bool(np.array_equal(got1.view(np.uint32), ref1.view(np.uint32))),
                    "max_abs_err_vs_oracle": float(np.abs(got1 - ref1).max())}
        except Exception as ex:                                   # informational: never costs the bench line
            cfg1 = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "audio-seconds/sec (RTF) Silero v3.1 at 1/2/4/8 B200 vs host-CPU C backend",
            "value": value, "unit": "audio-seconds/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: 4096 concurrent synthetic 16 kHz s16le streams x 10 min per GPU, per-stream LSTM state on device; "
                                   "step = 4096 streams x 125 chunks (12 s); 50 steps = the 10 minutes",
                       "streams_per_gpu": S, "chunks_per_step": C, "l2_policy": "inputs larger than L2 (1.57 GB PCM per step, rotating step buffers)",
                       "parity_max_abs_err_vs_oracle": parity, "segments_gathered": seg_count, "sharding": "streams across ranks, no collective on the data path",
                       "cfg1_single_stream": cfg1},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "audio-seconds/sec", "h2d_bytes_per_step": S * C * CHUNK * 2, "d2h_bytes_per_step": S * C * 4 + S * SEG_CAP * 8 + S * 4,
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "silero_b200_submit_streams_segments + silero_b200_wait (2 steps in flight): pinned host s16 PCM in; probabilities + "
                            "on-device-segmenter (start,end) pairs out, every step's H2D and D2H inside the timed region"},
            "gpu_launches": launches,
        }
        print(json.dumps(line))
    vadc_b200.pinned_free(h_pcm_ptr)
    for o in outs:
        for ptr in (o[1], o[3], o[5]):
            vadc_b200.pinned_free(ptr)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


# dram bytes per chunk of each kernel from the committed ncu captures (profiles/); absent until measured
# dram__bytes_read.sum + dram__bytes_write.sum per chunk of each kernel, from the ncu --set full capture of one window of
# 81 920 chunks (profiles/ncu_summary_r01h.md). For the STFT kernel the algorithmic bytes are 3 072 (s16 PCM) + 12 900 (log
# spectrogram) + 4 (normalization scalar) = 15 976 per chunk: traffic == algorithmic, nothing is re-read.
# algorithmic bytes per chunk of each kernel (what it must read + write once): STFT 3 072 s16 PCM + 12 900 log spectrogram + 4 scalar;
# first layer 12 900 + 4 in, 832 out; layers 2..4 832/896/896 in, 896/896/1 792 out; LSTM layer 0 1 792 in + 1 792 packed h out
# (state: 1 KB per stream per launch, negligible per chunk); layer 1 1 792 in, 8 out
ALGORITHMIC_BYTES_PER_CHUNK = {"stft": 15976, "layer1": 13736, "layer2": 1728, "layer3": 1792, "layer4": 2688, "lstm0": 3584, "lstm1_decoder": 1800}
TRAFFIC_BYTES_PER_CHUNK = {"stft": 15336, "layer1": 13714, "layer2": 1128, "layer3": 1218, "layer4": 2064, "lstm0": 3117, "lstm1_decoder": 1905}

if __name__ == "__main__":
    sys.exit(main())
