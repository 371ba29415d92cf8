#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from an ncu report captured with --import-source on (page `source`): warp-level
executed-instruction counts grouped by opcode and by class.  Usage: tools/opcode_mix.py <report.ncu-rep> <kernel-substring> [instance] [top-opcodes]"""
import csv
import re
import subprocess
import sys
from collections import Counter

CLASSES = [
    ("fp32 arithmetic", r"^(FMUL|FADD|FFMA|FMNMX|FSEL|FSETP|FCHK|MUFU|FRND|F2I|I2F|I2FP|F2F|HADD2|HFMA2|HMUL2|F2FP|FSET)"),
    ("integer / logic", r"^(IMAD|IADD3|VIADD|LOP3|SHF|LEA|ISETP|SEL|PRMT|POPC|VIMNMX|IABS|FLO|BREV|PLOP3|MOV|CS2R|S2R|IMNMX|SGXT|BMSK|P2R|R2P|VABSDIFF)"),
    ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR|VOTEU|REDUX|LDCU)"),
    ("memory", r"^(LDG|STG|LDS|STS|LDL|STL|LDC|ATOMG|ATOMS|RED|ATOM|LD|ST|CCTL|MEMBAR|ERRBAR|UBLKCP)"),
    ("warp / sync", r"^(SHFL|VOTE|MATCH|BAR|WARPSYNC|NANOSLEEP|DEPBAR|BSSY|BSYNC|BMOV|SYNCS)"),
    ("control", r"^(BRA|BRX|JMP|CALL|RET|EXIT|NOP|BREAK|BPT|YIELD|ENDCOLLECTIVE|ACQBULK|KILL)"),
]


def main():
    rep, kname = sys.argv[1], sys.argv[2]
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    sel = [k for k in range(len(idx) - 1) if kname in rows[idx[k]][1]]
    k = sel[inst]
    h = rows[idx[k] + 1]
    isrc, iexec, ithr = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ops, thr = Counter(), Counter()
    for r in rows[idx[k] + 2 : idx[k + 1]]:
        if len(r) <= iexec or not r[iexec].isdigit():
            continue
        text = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip())
        op = text.split()[0].split(".")[0] if text else "?"
        ops[op] += int(r[iexec])
        thr[op] += int(r[ithr]) if r[ithr].isdigit() else 0
    total, tthr = sum(ops.values()), sum(thr.values())
    print("kernel: %s   warp instructions %d   thread instructions %d (%.1f active lanes / warp instruction)" % (rows[idx[k]][1], total, tthr, tthr / max(total, 1)))
    cls = Counter()
    for op, n in ops.items():
        for name, pat in CLASSES:
            if re.match(pat, op):
                cls[name] += n
                break
        else:
            cls["other (" + op + ")"] += n
    print("\nby class:")
    for name, n in cls.most_common():
        print("  %-22s %12d  %5.1f%%" % (name, n, 100.0 * n / total))
    print("\nby opcode:")
    for op, n in ops.most_common(top):
        print("  %-10s %12d  %5.1f%%" % (op, n, 100.0 * n / total))


if __name__ == "__main__":
    main()
