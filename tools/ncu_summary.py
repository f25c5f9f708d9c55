"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + SASS opcode mix + stall reasons.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__cycles_active.avg"]
for w in want:
    if w in h:
        i = h.index(w)
        print("%-75s %s  [%s]" % (w, [r[i][:60] for r in rows[2:]], rows[1][i]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hh = rows[hdr[which]]
end = hdr[which + 1] - 1 if len(hdr) > which + 1 else len(rows)
body = rows[hdr[which] + 1:end]
ie, isrc, ist = hh.index("Instructions Executed"), hh.index("Source"), hh.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in body if r[ie].isdigit())
print("static instr:", len(body), "dynamic warp-instr:", tot)
op, stall = collections.Counter(), collections.Counter()
for r in body:
    if not r[ie].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    name = m.group(2).split(".")[0] if m else "?"
    op[name] += int(r[ie])
    stall[name] += int(r[ist]) if r[ist].isdigit() else 0
ts = max(1, sum(stall.values()))
for k, v in op.most_common(22):
    print("  %-10s %6.2f%% instr   %6.2f%% stall samples" % (k, 100 * v / tot, 100 * stall[k] / ts))
for name in ["stall_long_sb", "stall_lg", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_mio", "stall_branch_resolving",
             "stall_no_inst", "stall_dispatch", "stall_barrier", "stall_membar", "stall_drain"]:
    if name in hh:
        i = hh.index(name)
        print("  %-24s %d" % (name, sum(int(r[i]) for r in body if r[i].isdigit())))
