"""Turns an .ncu-rep (gpurun_out/, scratch) into the text summary committed under profiles/:
the `--page details` sections plus the raw counters the roofline numbers come from and a
warp-stall sample table per barrier-delimited code segment (from `--page source`)."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(*a):
    return subprocess.run(["ncu", "-i", rep] + list(a), capture_output=True, text=True).stdout


lines = ["# %s" % rep, "", "## raw counters", ""]
rows = list(csv.reader(run("--page", "raw", "--csv").splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
lines.append("kernel: %s  grid %s block %s" % (d["Kernel Name"][1], d["Grid Size"][1], d["Block Size"][1]))
for k in KEYS:
    if k in d:
        lines.append("%-90s %12s %s" % (k, d[k][1], d[k][0]))
for h, (u, v) in d.items():
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        try:
            if float(v) >= 0.1:
                lines.append("%-90s %12s" % (h, v))
        except ValueError:
            pass
lines += ["", "## warp-stall samples per code segment (segments end at BAR.SYNC / EXIT)", ""]
rows = list(csv.reader(run("--page", "source", "--csv").splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
seg0, acc, nd, nf, agg = 0, 0, 0, 0, {s: 0 for s in stalls}
for i, r in enumerate(data):
    src = r[ix["Source"]].strip()
    acc += int(r[ix["# Samples"]])
    for s in stalls:
        agg[s] += int(r[ix[s]])
    nd += "DMMA" in src
    nf += any(t in src for t in ("DFMA", "DMUL", "DADD"))
    if "BAR.SYNC" in src or "EXIT" in src or i == len(data) - 1:
        if acc > tot * 0.004:
            top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
            st = sum(agg.values()) or 1
            lines.append("instr %5d-%5d  samples %8d (%5.1f%%)  DMMA %4d  FP64 %4d  top stalls: %s" % (
                seg0, i, acc, 100.0 * acc / tot, nd, nf,
                ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / st) for k, v in top)))
        seg0, acc, nd, nf, agg = i + 1, 0, 0, 0, {s: 0 for s in stalls}
lines += ["", "## ncu --page details", "", run("--page", "details")]
open(out, "w").write("\n".join(lines))
print("wrote", out)
