#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report.

`ncu --page source --csv` lists SASS instructions with their counters but (from the CLI) no CUDA-C correlation;
`nvdisasm -g` of the cubin inside the library carries the line table.  The two listings hold the same
instructions in the same order, so they are joined by offset.

    python tools/ncu_lines.py gpurun_out/TAG_source.csv.gz deflate_kernel [--top 40] [--ranges a-b,c-d ...]
"""
import csv
import gzip
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(path, kernel):
    op = gzip.open if path.endswith(".gz") else open
    rows, cur, hdr = [], None, None
    with op(path, "rt", newline="") as f:
        for rec in csv.reader(f):
            if not rec:
                continue
            if rec[0] == "Kernel Name":
                cur = rec[1]
                hdr = None
                continue
            if rec[0] == "Address":
                hdr = rec
                continue
            if cur and kernel in cur and hdr:
                rows.append(dict(zip(hdr, rec)))
    return rows


def line_table(kernel, lib):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for fn in sorted(os.listdir(tmp)):
        if not fn.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, fn)], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        out, on, loc = [], False, ("?", 0)
        for ln in txt.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                on = kernel in ln
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((int(m.group(1), 16), loc, m.group(2).strip()))
        if out:
            return out
    raise SystemExit("kernel %s not found in %s" % (kernel, lib))


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    top = 40
    ranges = []
    lib = os.path.join(ROOT, "slow5tools_b200", "libslow5b200.so")
    a = sys.argv[3:]
    while a:
        if a[0] == "--top":
            top = int(a[1]); a = a[2:]
        elif a[0] == "--lib":
            lib = a[1]; a = a[2:]
        elif a[0] == "--ranges":
            for part in a[1].split(","):
                f, _, r = part.partition(":")
                lo, _, hi = r.partition("-")
                ranges.append((f, int(lo), int(hi)))
            a = a[2:]
        else:
            raise SystemExit("unknown arg " + a[0])
    rows = sass_rows(path, kernel)
    tab = line_table(kernel, lib)
    if len(rows) != len(tab):
        print("warning: %d profiled instructions vs %d in the cubin (different build?)" % (len(rows), len(tab)))
    inst, samp, nsass = defaultdict(int), defaultdict(int), defaultdict(int)
    tot_i = tot_s = 0
    for r, (_, loc, _) in zip(rows, tab):
        i = int(r.get("Instructions Executed", "0") or 0)
        s = int(r.get("# Samples", "0") or 0)
        inst[loc] += i; samp[loc] += s; nsass[loc] += 1
        tot_i += i; tot_s += s
    print("%s: %d SASS, %.4g warp instructions, %d stall samples" % (kernel, len(rows), tot_i, tot_s))
    if ranges:
        for f, lo, hi in ranges:
            ii = sum(v for (ff, l), v in inst.items() if ff.startswith(f) and lo <= l <= hi)
            ss = sum(v for (ff, l), v in samp.items() if ff.startswith(f) and lo <= l <= hi)
            print("  %-28s %5d-%-5d inst %6.2f%%  samples %6.2f%%" % (f, lo, hi, 100.0 * ii / tot_i, 100.0 * ss / max(1, tot_s)))
    print("  top lines by stall samples:")
    for loc, s in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
        print("  %-28s:%-5d sass %4d  inst %6.2f%%  samples %6.2f%%" % (loc[0], loc[1], nsass[loc], 100.0 * inst[loc] / tot_i,
                                                                  100.0 * s / max(1, tot_s)))


if __name__ == "__main__":
    main()
