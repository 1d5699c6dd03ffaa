#!/usr/bin/env python3
"""Compact text summary of an .ncu-rep (raw page): per kernel launch the metrics this project's roofline
story needs — duration, issue-slot and FP32-pipe utilisation, warp execution efficiency, occupancy, stall
reasons, L2 / DRAM traffic.   usage: tools/ncu_summary.py file.ncu-rep [> profiles/xxx.txt]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy % (active)"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (elapsed)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst (of 32)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy %"),
    ("sm__inst_executed_pipe_fma.sum", "inst pipe fma"), ("sm__inst_executed_pipe_fmaheavy.sum", "inst pipe fmaheavy"),
    ("sm__inst_executed_pipe_fmalite.sum", "inst pipe fmalite"),
    ("sm__inst_executed_pipe_alu.sum", "inst pipe alu"), ("sm__inst_executed_pipe_xu.sum", "inst pipe xu (MUFU)"),
    ("sm__inst_executed_pipe_lsu.sum", "inst pipe lsu"), ("sm__inst_executed_pipe_uniform.sum", "inst pipe uniform"),
    ("sm__inst_executed_pipe_cbu.sum", "inst pipe cbu (branch)"), ("sm__inst_executed_pipe_adu.sum", "inst pipe adu"),
    ("smsp__inst_executed_op_branch.sum", "branch inst"),
    ("sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "thread FFMA"), ("sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "thread FMUL"),
    ("sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "thread FADD"),
    ("smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "thread FP32 inst"),
    ("lts__t_sectors.sum", "L2 sectors (x 32 B)"), ("lts__t_sectors.sum.per_second", "L2 sectors per second"),
    ("lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "L2 sector throughput % of peak"),
    ("SM_B.TriageCompute.l1tex__t_sectors.sum", "L1 sectors (x 32 B)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        get = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else None
        print("== %s  [launch id %s]" % (r[hdr.index("Kernel Name")], r[hdr.index("ID")]))
        for k, label in KEYS:
            v = get(k)
            if v:
                print("   %-40s %s %s" % (label, v[0], v[1]))
        # achieved L2 / L1 bandwidth in GB/s (sectors are 32 B), the evidence behind "L2-resident, not HBM-bound"
        def num(k):
            v = get(k)
            try:
                return float(v[0].replace(",", "")), v[1]
            except Exception:
                return None, None
        dur, du = num("gpu__time_duration.sum")
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(du or "", None)
        l2s, _ = num("lts__t_sectors.sum")
        l1s, _ = num("SM_B.TriageCompute.l1tex__t_sectors.sum")
        dr, dru = num("dram__bytes_read.sum")
        if dur and scale:
            if l2s:
                print("   %-40s %.1f GB/s" % ("L2 achieved bandwidth", l2s * 32 / (dur * scale) / 1e9))
            if l1s:
                print("   %-40s %.1f GB/s" % ("L1 achieved bandwidth (sectors)", l1s * 32 / (dur * scale) / 1e9))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
               h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   top stall reasons (warps stalled per issue / latency ratios):")
        for v, h in stalls[:8]:
            print("      %8.3f  %s" % (v, h.replace("smsp__average_", "")))
        print()


if __name__ == "__main__":
    main()
