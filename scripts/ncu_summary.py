#!/usr/bin/env python
"""Condense an `ncu --set full` report into the markdown summary kept under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [launches.csv] > profiles/rNN_X.md

Per kernel: duration, DRAM bytes (the `traffic` figure of bench.py's roofline object), issue /
pipe utilisation, occupancy limiters, shared-memory wavefronts and bank conflicts, warp stall
reasons, branch efficiency.  The optional launch list (`--metrics gpu__time_duration.sum`) is
folded into a per-kernel share-of-step table.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid size (CTAs)"),
    ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__shared_mem_per_block_static", "static smem / CTA"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: smem (CTAs/SM)"),
    ("launch__occupancy_limit_warps", "occupancy limit: warps (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch efficiency (uniform targets) %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "pipe ADU %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "pipe uniform %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` on one B200 (see "
          "scripts/gpu_profile.sh); per-launch values, cold cache, serialised launches.\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(f"## `{name}`\n")
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                v = r[col[k]]
                f = fnum(v)
                if f is not None:
                    v = f"{f:,.2f}" if abs(f) < 1e6 and f != int(f) else f"{f:,.0f}"
                print(f"| {label} | {v} {units[col[k]]} |")
        rd, wr = fnum(r[col["dram__bytes_read.sum"]]), fnum(r[col["dram__bytes_write.sum"]])
        dur = fnum(r[col["gpu__time_duration.sum"]])
        mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        urd, uwr = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
        udur = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[col["gpu__time_duration.sum"]]]
        if rd is not None and wr is not None and dur:
            tot = rd * mult.get(urd, 1.0) + wr * mult.get(uwr, 1.0)
            print(f"| **DRAM traffic (read+write)** | {tot / 1e6:,.2f} MB -> {tot / (dur * udur) / 1e9:,.0f} GB/s |")
        stalls = []
        for h, i in col.items():
            if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                f = fnum(r[i])
                if f is not None and f >= 0.05:
                    stalls.append((f, h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
        stalls.sort(reverse=True)
        print("\nWarp stall reasons (warps stalled per issue-active cycle, >= 0.05): " +
              ", ".join(f"{n} {f:.2f}" for f, n in stalls) + "\n")
    if len(sys.argv) > 2:
        agg, order = defaultdict(list), []
        for r in csv.DictReader(l for l in open(sys.argv[2]) if l.startswith('"')):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            n = r["Kernel Name"].split("(")[0].replace("void ", "")
            if n not in agg:
                order.append(n)
            agg[n].append(float(r["Metric Value"]))
        ours = [n for n in order if "gpc" in n]
        tot = sum(sum(agg[n]) for n in ours)
        print(f"## launch list `{sys.argv[2].split('/')[-1]}` (gpu__time_duration.sum, ns)\n")
        print("| kernel | launches | mean ns | share of the step (our kernels) |")
        print("|---|---|---|---|")
        for n in order:
            v = agg[n]
            share = f"{100 * sum(v) / tot:.1f} %" if n in ours and tot else "(torch fill, setup)"
            print(f"| `{n}` | {len(v)} | {sum(v) / len(v):,.0f} | {share} |")


if __name__ == "__main__":
    main()
