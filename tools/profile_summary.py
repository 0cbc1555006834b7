#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.txt
  python tools/profile_summary.py full     gpurun_out/prof_X.ncu-rep profiles/X_ncu_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list: {src}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':72s} {'n':>5s} {'total_ms':>10s} {'share':>7s} {'avg_us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {v[0]:5d} {v[1] / 1e3:10.3f} {v[1] / tot:7.3f} {v[1] / v[0]:10.1f}\n")
        f.write(f"total {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print(open(dst).read())


def _raw_rows(src):
    """rows of `ncu --page raw --csv`: from a report, or from a csv the GPU session already exported (large reports
    are not brought back)."""
    if src.endswith(".csv"):
        return list(csv.reader(open(src)))
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(src, dst):
    rows = _raw_rows(src)
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none capture: {src} (one column per captured launch)\n")
        names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void stc::", "").replace("stc::", "") for r in data]
        f.write("metric," + ",".join(f"{i}:{n}" for i, n in enumerate(names)) + "\n")
        for k in KEYS:
            if k in idx:
                f.write(f"{k} [{units[idx[k]]}]," + ",".join(r[idx[k]] for r in data) + "\n")
    print(open(dst).read())


def traffic(src, dst, batch="4096"):
    """profiles/traffic.json: mean dram bytes (read + write) per launch of each captured kernel kind."""
    import json
    rows = _raw_rows(src)
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kinds = {"tc_conv_bwd_dx_kernel": "tc_conv_bwd_dx", "tc_conv_fwd": "tc_conv_fwd", "tc_conv_bwd_dw": "tc_conv_bwd_dw",
             "tc_support_kernel": "tc_support", "tc_outer_kernel": "tc_outer"}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in data:
        name = r[idx["Kernel Name"]]
        kind = next((v for k, v in kinds.items() if k in name), None)
        if kind is None:
            continue
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[idx["gpu__time_duration.sum"]]]
        agg[kind][0] += 1
        agg[kind][1] += b
        agg[kind][2] += t
    d = {"source": "ncu --set full --clock-control none capture " + src.split("/")[-1], "batch": int(batch),
         "kernels": {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "mean_us_under_ncu": v[2] / v[0]}
                     for k, v in agg.items()}}
    with open(dst, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
