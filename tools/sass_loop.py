#!/usr/bin/env python3
"""Static instruction mix of the sweep loop(s) of search_kernel, from cuobjdump -sass.

usage: tools/sass_loop.py opal_b200/csrc/build/kernels_R17.o [flavor=3] [arith=Packed16] [chain]

For every innermost loop that holds the DPX recurrence (the wavefront step) prints how many instructions of each
pipe class the straight-line body holds.  Classes follow the B200 measurements in profiles/README.md:
the DPX / min-max / logic / compare / select / shift instructions issue on the 16-lane ALU pipe (the
bottleneck of this kernel), IMAD* / VIADD on the FMA pipe, loads / stores / shuffles on the LSU.
"""
import re
import subprocess
import sys
from collections import Counter

ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "ISETP", "SEL", "LOP3", "PRMT", "SHF", "IADD3", "LEA", "PLOP3", "VABSDIFF", "IABS", "FLO", "POPC", "BREV", "SGXT", "BMSK", "ICMP", "IMNMX", "VOTE", "P2R", "R2P")
FMA = ("IMAD", "VIADD", "MOV", "FMUL", "FADD", "FFMA")
LSU = ("LDS", "LDG", "STG", "STS", "SHFL", "LD", "ST", "ATOMG", "LDC", "LDCU", "REDUX", "MATCH")


def classify(op):
    base = op.split(".")[0]
    if base == "VIADD" and ".16x2" in op:
        return "FMA(VIADD.16x2)"
    if base in ALU:
        return "ALU"
    if base in FMA:
        return "FMA"
    if base in LSU:
        return "LSU"
    return "CTRL"


def main():
    obj = sys.argv[1]
    flavor = sys.argv[2] if len(sys.argv) > 2 else "3"
    arith = sys.argv[3] if len(sys.argv) > 3 else "Packed16"
    chained = len(sys.argv) > 4 and sys.argv[4] == "chain"
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if f"ELi{flavor}ENS_" not in name or arith not in name:
            continue
        if ("ELb1E" in name) != chained:  # the chained-pass variants are listed with a fourth argument "chain"
            continue
        ins = []  # (addr, pred, opcode, text)
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Za-z0-9_.]+)\s*([^;]*);", f):
            ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
        addr_index = {a: i for i, (a, _, _, _) in enumerate(ins)}
        loops = []
        for i, (a, pred, op, txt) in enumerate(ins):
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", txt)
                if t and int(t.group(1), 16) <= a and int(t.group(1), 16) in addr_index:
                    loops.append((addr_index[int(t.group(1), 16)], i))
        print(name)
        for lo, hi in loops:
            body = ins[lo:hi + 1]
            def is_sweep(b):
                return sum(1 for _, _, op, _ in b if op.startswith("VIADDMNMX")) >= 8
            if not is_sweep(body):
                continue
            if any((l2 > lo or h2 < hi) and l2 >= lo and h2 <= hi and is_sweep(ins[l2:h2 + 1]) for l2, h2 in loops):
                continue
            cls = Counter()
            ops = Counter()
            for _, _, op, _ in body:
                c = classify(op)
                cls[c] += 1
                ops[(c, op.split(".")[0] + (".16x2" if "16x2" in op else ""))] += 1
            print(f"  loop 0x{body[0][0]:x}-0x{body[-1][0]:x}: {len(body)} instructions  " + "  ".join(f"{k}={v}" for k, v in sorted(cls.items())))
            for (c, op), n in sorted(ops.items()):
                print(f"      {c:16s} {op:20s} {n}")


if __name__ == "__main__":
    main()
