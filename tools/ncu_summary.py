"""Turns ncu reports / launch lists under gpurun_out/ into the short text summaries kept in profiles/."""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def summary(rep):
    hdr, units, data = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in data:
        lines.append(f"kernel: {r[col['Kernel Name']][:90]}   (launch id {r[col['ID']]})")
        for k in KEYS:
            if k in col:
                lines.append(f"  {k:78s} {r[col[k]]:>16s} {units[col[k]]}")
    return "\n".join(lines)


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki][:70]; v = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"{'kernel':72s} {'n':>4s} {'total ms':>10s} {'share':>7s}"]
    for k, a in agg.items():
        out.append(f"{k:72s} {a[0]:4d} {a[1]/1e6:10.3f} {a[1]/tot*100:6.1f}%")
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(summary(sys.argv[1]))
