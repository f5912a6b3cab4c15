"""Per-kernel facts of one window from an .ncu-rep (read here, no GPU needed) as JSON for bench.py and DESIGN.md:
launch duration, executed FP32 thread instructions (FADD, FMUL, FFMA), DRAM bytes, all also per chunk.
usage: python scripts/ncu_export.py gpurun_out/prof_X.ncu-rep <chunks per launch> profiles/opmix_X.json"""
import csv, io, json, subprocess, sys

rep, chunks, dst = sys.argv[1], float(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, body = rows[0], rows[1], rows[2:]


def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    for pre, k in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9), ("second", 1.0)):
        if u.startswith(pre):
            return v * k
    return v * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


# Packed FP32 instructions (FADD2 / FMUL2 / FFMA2, sm_100) are not in the op_fadd / op_fmul / op_ffma thread-instruction metrics: take
# their warp-level execution counts from the source page (x 32 lanes; the kernels that use them run full warps there)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
packed = []   # per launch, in report order: {opcode: warp-level count}
cur = None
for r in srows:
    if "Instructions Executed" in r:
        iS, iE = r.index("Source"), r.index("Instructions Executed")
        cur = {"FADD2": 0, "FMUL2": 0, "FFMA2": 0}
        packed.append(cur)
        continue
    if cur is None or len(r) <= iE or not r[iE].strip().isdigit():
        continue
    toks = r[iS].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    if op in cur:
        cur[op] += int(r[iE])

kernels = {}
for li, r in enumerate(body):
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
    cyc = val(r, "smsp__cycles_elapsed.max") if "smsp__cycles_elapsed.max" in hdr else val(r, "sm__cycles_elapsed.max")
    fadd = val(r, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed") * cyc
    fmul = val(r, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed") * cyc
    ffma = val(r, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed") * cyc
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    # (the source page lists every launch twice)
    pk = packed[2 * li] if 2 * li < len(packed) else {"FADD2": 0, "FMUL2": 0, "FFMA2": 0}
    packed_flop = 32.0 * (2 * pk["FADD2"] + 2 * pk["FMUL2"] + 4 * pk["FFMA2"])
    k = {"seconds_under_ncu": val(r, "gpu__time_duration.sum"), "fadd": fadd, "fmul": fmul, "ffma": ffma,
         "fadd2_warp_instructions": pk["FADD2"], "fmul2_warp_instructions": pk["FMUL2"], "ffma2_warp_instructions": pk["FFMA2"],
         "executed_flop_per_chunk": (fadd + fmul + 2 * ffma + packed_flop) / chunks, "dram_bytes": dram, "dram_bytes_per_chunk": dram / chunks,
         "warp_instructions_per_chunk": val(r, "smsp__inst_executed.sum") / chunks,
         "fma_pipe_pct": val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "registers": val(r, "launch__registers_per_thread")}
    kernels.setdefault(name, k)  # first instance of each kernel
json.dump({"report": rep, "chunks_per_launch": chunks, "kernels": kernels}, open(dst, "w"), indent=1)
for n, k in kernels.items():
    print("%-28s %8.3f ms  executed %9.0f flop/chunk  dram %8.0f B/chunk  fma pipe %4.1f %%" % (n, k["seconds_under_ncu"] * 1e3, k["executed_flop_per_chunk"], k["dram_bytes_per_chunk"], k["fma_pipe_pct"]))
