#!/usr/bin/env python
"""Condense ncu artefacts brought back in gpurun_out/ into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv        > profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof_pair.ncu-rep   > profiles/rNN_k_pair.txt

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (count, total, share).
`kernel` prints the headline metrics of one `--set full` capture (duration, registers, occupancy, pipe
utilisation, DRAM traffic, stall breakdown) and the hottest SASS lines with their dominant stall reason.
Needs the `ncu` CLI (no GPU) for the `kernel` mode.
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki]
        name = name.replace("void ", "").replace("<unnamed>::", "")
        agg[name[:70]][0] += 1
        agg[name[:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of device time "
          "(ncu serialised, cold-cache: compare shares, not absolutes)")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>12s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 5e-4:
            continue
        print(f"{k:70s} {v[0]:5d} {v[1] / 1e3:12.1f} {v[1] / tot:7.3f}")


def kernel(rep, top=28):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    for vals in raw[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:140])
        for k in RAW_KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>16s} {u[k]}")
        stalls = sorted(((float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and
                         h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")), reverse=True)
        print("  stall reasons (warps stalled per issue-active cycle):")
        for v, h in stalls[:8]:
            print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
    src = ncu_csv(rep, "source")
    # the source page holds one table per captured kernel; take the first
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[start]
    i_s, i_i, i_src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    data = []
    for r in src[start + 1:]:
        if len(r) != len(h) or r[0] == "Address":
            break
        data.append(r)
    tot = sum(int(r[i_s]) for r in data) or 1
    toti = sum(int(r[i_i]) for r in data)
    print(f"  SASS: {len(data)} instructions, {toti} warp-instructions executed, {tot} stall samples; hottest lines:")
    for r in sorted(data, key=lambda r: -int(r[i_s]))[:top]:
        st = max(((int(r[i]), h[i]) for i in stall_cols), default=(0, ""))
        print(f"    {100 * int(r[i_s]) / tot:5.1f}%  exec={int(r[i_i]):>11d}  {r[i_src].strip()[:58]:58s} {st[1]}")


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "kernel"):
        sys.exit(__doc__)
    (launches if sys.argv[1] == "launches" else kernel)(sys.argv[2])
