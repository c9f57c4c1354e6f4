#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and launch CSVs into the small text summaries committed under profiles/.

    python profiles/summarize.py rep  gpurun_out/prof_x.ncu-rep  profiles/r1_ncu_x.txt
    python profiles/summarize.py list gpurun_out/launches.csv    profiles/r1_launches_x.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none; cold-cache, serialised replays)\n" % path)
        for n, r in enumerate(rows[2:]):
            f.write("\n## launch %d\n" % n)
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("%-86s %s %s\n" % (k, r[i], units[i]))
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")])
                wr = float(r[hdr.index("dram__bytes_write.sum")])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
                dur = float(r[hdr.index("gpu__time_duration.sum")])
                du = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
                f.write("%-86s %.4g bytes, %.1f GB/s under ncu\n" % ("traffic = dram read + write", tot, tot / (dur * du) / 1e9))
            except Exception:
                pass


def lst(path, out):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# %s: every launch of the command, ncu --metrics gpu__time_duration.sum --clock-control none\n" % path)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("%-100s %6s %12s %7s %10s\n" % ("kernel", "n", "total_us", "share", "avg_us"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-100s %6d %12.1f %7.3f %10.1f\n" % (k[:100], v[0], v[1] / 1e3, v[1] / tot, v[1] / v[0] / 1e3))
        is_mine = lambda k: "fq::" in k or k.startswith("fq") or "hist_" in k or "kl_" in k
        mine = sum(v[1] for k, v in agg.items() if is_mine(k))
        f.write("\nshare of libfq_b200 kernels in the whole process: %.3f\n" % (mine / tot))
        f.write("\n# libfq_b200 kernels only.  bench.py's TIMED region launches nothing else (the torch/cuDNN kernels above\n"
                "# belong to the untimed set-up that produces the layer inputs), so these are the shares to compare with\n"
                "# the CUDA-event breakdown in the bench JSON.\n")
        f.write("%-100s %6s %12s %7s %10s\n" % ("kernel", "n", "total_us", "share", "avg_us"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if is_mine(k):
                f.write("%-100s %6d %12.1f %7.3f %10.1f\n" % (k[:100], v[0], v[1] / 1e3, v[1] / mine, v[1] / v[0] / 1e3))


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
