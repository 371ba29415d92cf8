#!/usr/bin/env python
"""Joins an ncu report's per-SASS-instruction execution counts with nvdisasm line info and prints the hottest source
lines of one kernel.  Usage: tools/sass_lines.py <report.ncu-rep> <kernel-substring> [instance] [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kname = sys.argv[1], sys.argv[2]
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    sel = [k for k in range(len(idx) - 1) if kname in rows[idx[k]][1]]
    k = sel[inst]
    h = rows[idx[k] + 1]
    ia, ie, it, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    body = [r for r in rows[idx[k] + 2 : idx[k + 1]] if len(r) > ie and r[ie].isdigit()]
    mangled = None
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "niagara_b200", "libniagara_cull.so")], cwd=tmp, capture_output=True)
    cubin = os.path.join(tmp, "nvc_kernels.sm_100a.cubin")
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # split into functions
    funcs = defaultdict(list)
    cur, line, last_frame, chain_open = None, None, None, False

    def helper(fl):
        f, n = fl
        if f == "nvc_filter.cuh":
            return n < 93
        if f == "nvc_math.cuh":
            return n < 50
        return f not in ("nvc_kernels.cu", "nvc_math.cuh", "nvc_tma.cuh", "nvc_cook.cuh", "nvc_math2.cuh")

    for l in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', l)
        if m:
            # -gi prints the inlining chain innermost first, one line per frame; an instruction is attributed to the first
            # frame that is not a one-line helper (nvf_fma / nvf_rcp / half conversion / CUDA headers)
            here = (os.path.basename(m.group(1)), int(m.group(2)))
            if not chain_open:
                line, chain_open = None, True
            if line is None and not helper(here):
                line = here
            if line is None and m.group(4):
                up = (os.path.basename(m.group(4)), int(m.group(5)))
                if not helper(up):
                    line = up
            last_frame = here
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur:
            funcs[cur].append((line or last_frame, m.group(2).strip()))
            chain_open = False
    # pick the function whose instruction count matches
    cand = [f for f, ins in funcs.items() if len(ins) == len(body)]
    want = rows[idx[k]][1]
    print("kernel:", want, "sass:", len(body), "candidates:", cand)
    fn = [c for c in cand if ("Lb1" in c) == ("(bool)1" in want.split("<")[1].split(",")[0] + want)]
    fn = (fn or cand)[0]
    ins = funcs[fn]
    tot = sum(int(r[ie]) for r in body)
    per_line = Counter()
    per_line_ops = defaultdict(Counter)
    samples = Counter()
    for (line, text), r in zip(ins, body):
        per_line[line] += int(r[ie])
        samples[line] += int(r[isamp]) if r[isamp].isdigit() else 0
        toks = text.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        per_line_ops[line][op.split(".")[0]] += int(r[ie])
    src = {}
    print("total warp instructions:", tot, " stall samples:", sum(samples.values()))
    for line, n in per_line.most_common(top):
        if line and line[0] not in src:
            for d in ("niagara_b200/csrc",):
                p = os.path.join(ROOT, d, line[0])
                if os.path.exists(p):
                    src[line[0]] = open(p).read().splitlines()
        text = src.get(line[0], [""] * 100000)[line[1] - 1].strip() if line else ""
        ops = " ".join("%s:%d" % (o, c * 1000 // max(n, 1)) for o, c in per_line_ops[line].most_common(4))
        print("%5.1f%% smp %4.1f%% %s:%d  %s   [%s]" % (100.0 * n / tot, 100.0 * samples[line] / max(1, sum(samples.values())), line[0] if line else "?", line[1] if line else 0, text[:110], ops))


if __name__ == "__main__":
    main()
