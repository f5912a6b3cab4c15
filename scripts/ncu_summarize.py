"""Summarise an .ncu-rep (read here, no GPU needed) + a launch list into profiles/ncu_summary_<tag>.md.
usage: python scripts/ncu_summarize.py gpurun_out/prof_X.ncu-rep gpurun_out/launches_X.csv profiles/ncu_summary_X.md "<title>" """
import csv, io, subprocess, sys, collections

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sass__inst_executed_register_spilling", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]
STALL_PREFIX = "smsp__pcsamp_warps_issue_stalled_"


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, launches, dst, title = sys.argv[1:5]
    md = ["# " + title, ""]
    # ---- launch shares ------------------------------------------------------------------------------
    tot = collections.OrderedDict()
    with open(launches) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r["Metric Value"].replace(",", "")) / 1e6
    s = sum(v[1] for v in tot.values())
    md += ["## Launch shares (`%s`, ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)" % launches.split("/")[-1], "",
           "| kernel | launches | avg ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in tot.items():
        md.append("| `%s` | %d | %.3f | %.1f %% |" % (k, n, ms / n, 100 * ms / s))
    md.append("")
    # ---- per-kernel metrics --------------------------------------------------------------------------
    hdr, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        name = r[col["Kernel Name"]].split("(")[0]
        md += ["## `%s`  grid %s block %s" % (name, r[col["Grid Size"]], r[col["Block Size"]]), "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in col:
                md.append("| %s | %s | %s |" % (m, r[col[m]], units[col[m]]))
        stalls = sorted(((float(r[i].replace(",", "") or 0), h[len(STALL_PREFIX):]) for h, i in col.items()
                         if h.startswith(STALL_PREFIX) and not h.endswith("_not_issued")), reverse=True)[:7]
        md.append("| top stall samples | %s | |" % ", ".join("%s %d" % (n, v) for v, n in stalls))
        md.append("")
    open(dst, "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
