"""Turns gpurun_out/*.ncu-rep / launch CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_v5.csv profiles/r01_launch_shares_greedy_step.csv "<note>"
    python profiles/summarize.py kernel   gpurun_out/prof_tc_v5.ncu-rep profiles/r01_ncu_knn_tc_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__mem_tensor_cycles_active", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct", "sm__inst_executed.avg.per_cycle_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime"]


def launches(src, dst, note):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        agg[r[kn]].append(v)
    tot = sum(sum(v) for v in agg.values())
    out = ["# " + note, "# ncu --metrics gpu__time_duration.sum --clock-control none ; serialised cold-cache times: compare SHARES",
           "kernel,launches,avg_us,total_ms,share_pct"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append('"%s",%d,%.1f,%.2f,%.1f' % (k.split("(")[0], len(v), sum(v) / len(v), sum(v) / 1000, 100 * sum(v) / tot))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full --clock-control none --import-source on ; source: %s" % src]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append("== %s" % name[:120])
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS):
                out.append("%-90s %-10s %s" % (h, units[i], r[i]))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:80]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3])
