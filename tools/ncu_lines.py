#!/usr/bin/env python3
"""Attribute the executed warp instructions of one profiled kernel to CUDA source lines / inlined functions.

ncu's SASS source page gives per-instruction execution counts; nvdisasm -g gives the source line (and
inlining chain) of every instruction of the same cubin.  Joined by instruction order.
usage: tools/ncu_lines.py report.ncu-rep <kernel mangled-name substring> <ncu kernel regex> [topN]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200/lib/libptcuda.so")],
                   cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines = out.split("\n")
    res = []
    inside = False
    cur = ("?", 0, "")
    for ln in lines:
        if ln.startswith("//--------------------- .text."):
            inside = mangled in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (os.path.basename(m.group(1)), int(m.group(2)), " <- ".join("%s:%s" % (os.path.basename(a), b) for a, b in inl))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            res.append(cur)
    return res


def main():
    rep, mangled, regex = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ie = hdr.index("Instructions Executed")
    ite = hdr.index("Thread Instructions Executed")
    ist = hdr.index("Warp Stall Sampling (All Samples)")
    data = [r for r in rows[2:] if len(r) > ie and r[0].startswith("0x")]
    # the csv may list the kernel twice; keep the first pass (addresses strictly increasing)
    first = []
    last = -1
    for r in data:
        a = int(r[0], 16)
        if a <= last:
            break
        first.append(r)
        last = a
    loc = sass_lines(mangled)
    if len(loc) != len(first):
        print("warning: %d SASS instructions in cubin vs %d in report" % (len(loc), len(first)))
    n = min(len(loc), len(first))
    per_line = {}
    per_thr = {}
    per_stall = {}
    tot = tot_thr = tot_stall = 0
    for k in range(n):
        c = int(first[k][ie])
        tot += c
        key = loc[k][:2]
        per_line[key] = per_line.get(key, 0) + c
        th = int(first[k][ite] or 0); st = int(first[k][ist] or 0)
        per_thr[key] = per_thr.get(key, 0) + th
        per_stall[key] = per_stall.get(key, 0) + st
        tot_thr += th; tot_stall += st
    srcs = {}
    print("total executed warp instructions: %d   avg active threads %.2f   stall samples %d" % (tot, tot_thr / max(tot, 1), tot_stall))
    print(" warp-inst%  thr/inst  stall%   file:line")
    for (f, l), c in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        if f not in srcs:
            p = os.path.join(ROOT, "opencl_montecarlo_path_tracing_b200/csrc", f)
            srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
        print("%6.2f%%  %5.1f  %6.2f%%  %-18s %4d  %s" % (100.0 * c / tot, per_thr[(f, l)] / max(c, 1), 100.0 * per_stall[(f, l)] / max(tot_stall, 1), f, l, text))


if __name__ == "__main__":
    main()
