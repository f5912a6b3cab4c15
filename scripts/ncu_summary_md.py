"""Per-kernel table (markdown) of one ncu --set full capture. usage: python scripts/ncu_summary_md.py <rep> <out.md> "<title>" "<command>" """
import csv, io, subprocess, sys
rep, dst, title, cmd = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
M = [('gpu__time_duration.sum', 'ms'), ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
     ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'), ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU %'),
     ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU %'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU %'),
     ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 %'),
     ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem wavefronts %'),
     ('smsp__inst_executed.sum', 'warp instr'), ('dram__bytes_read.sum', 'DRAM rd'), ('dram__bytes_write.sum', 'DRAM wr')]
md = ["# " + title, "", "Command: `" + cmd + "`", "", "| kernel | " + " | ".join(n for _, n in M) + " | issue util % |", "|---|" + "---|" * (len(M) + 1)]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
    vals = []
    for k, n in M:
        i = hdr.index(k)
        f = float(r[i].replace(',', ''))
        vals.append("%.3f" % f if n == 'ms' else "%.3g %s" % (f, units[i]) if 'DRAM' in n else "%.3g" % f)
    inst = float(r[hdr.index('smsp__inst_executed.sum')].replace(',', ''))
    cyc = float(r[hdr.index('smsp__cycles_elapsed.max')].replace(',', ''))
    nsm = float(r[hdr.index('launch__grid_size')].replace(',', '')) if False else 148
    vals.append("%.0f" % (100 * inst / (4 * nsm) / cyc))
    md.append("| `%s` | " % name + " | ".join(vals) + " |")
open(dst, "w").write("\n".join(md) + "\n")
print("\n".join(md))
