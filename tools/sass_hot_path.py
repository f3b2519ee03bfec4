#!/usr/bin/env python3
"""Instruction mix of the HOT PATH of search_kernel's sweep loop, from cuobjdump -sass (static, no GPU needed).

usage: tools/sass_hot_path.py <object> [flavor=2] [arith=Packed16] [maxthreads=384] [chain=0]

The sweep loop is the innermost loop holding the DPX recurrence.  Blocks that a forward branch jumps over and that hold
no VIADDMNMX (rare events, the boundary paths of the other pass kinds) are listed as "cold" with their size and are not
counted; everything else between the loop head and the back edge is the hot path.  The executed mix measured by ncu
(profiles/README.md) is the reference this estimate was checked against."""
import re
import subprocess
import sys
from collections import Counter

ALU = ("VIADDMNMX", "VIMNMX", "VIMNMX3", "ISETP", "SEL", "LOP3", "PRMT", "SHF", "IADD3", "LEA", "PLOP3", "VABSDIFF", "IABS", "FLO", "POPC",
       "BREV", "SGXT", "BMSK", "ICMP", "IMNMX", "VOTE", "P2R", "R2P")
FMA = ("IMAD", "VIADD", "MOV", "FMUL", "FADD", "FFMA")
LSU = ("LDS", "LDG", "STG", "STS", "SHFL", "LD", "ST", "ATOMG", "LDC", "LDCU", "REDUX", "MATCH")


def classify(op):
    base = op.split(".")[0]
    if base in ALU:
        return "ALU"
    if base in FMA:
        return "FMA"
    if base in LSU:
        return "LSU"
    return "CTRL"


def main():
    obj = sys.argv[1]
    flavor = sys.argv[2] if len(sys.argv) > 2 else "2"
    arith = sys.argv[3] if len(sys.argv) > 3 else "Packed16"
    maxt = sys.argv[4] if len(sys.argv) > 4 else "384"
    chain = sys.argv[5] if len(sys.argv) > 5 else "0"
    import os
    dump = bool(os.environ.get("DUMP"))
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0]
        if f"ELi{flavor}ENS_" not in name or arith not in name or f"ELi{maxt}ELb{chain}E" not in name:
            continue
        ins = []
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Za-z0-9_.]+)\s*([^;]*);", f):
            ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
        index = {a: i for i, (a, _, _, _) in enumerate(ins)}

        def target(i):
            t = re.search(r"0x([0-9a-f]+)", ins[i][3])
            return index.get(int(t.group(1), 16)) if t else None

        loops = [(target(i), i) for i in range(len(ins)) if ins[i][2].startswith("BRA") and target(i) is not None and target(i) <= i]
        loops = [(lo, hi) for lo, hi in loops if sum(1 for k in range(lo, hi) if ins[k][2].startswith("VIADDMNMX")) >= 8]
        inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
        print(name)
        for lo, hi in inner:
            hot, cold = Counter(), []
            ops = Counter()
            i = lo
            while i <= hi:
                a, pred, op, txt = ins[i]
                hot[classify(op)] += 1
                if dump and not op.startswith(("VIADDMNMX", "LDS.128", "IMAD.IADD", "VIADD.16x2")):
                    print(f"      {a:06x} {pred:6s} {op} {txt}")
                ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("VIADD.", "IMAD.", "LDS.", "VIMNMX")) and "." in op else "")] += 1
                t = target(i) if op.startswith("BRA") else None
                if t is not None and i < t <= hi + 1 and not any(ins[k][2].startswith("VIADDMNMX") for k in range(i + 1, t)) and t - i > 1:
                    cold.append((t - i - 1, ins[i + 1][2] + " " + ins[i + 1][3][:40]))
                    i = t
                    continue
                i += 1
            total = sum(hot.values())
            print(f"  loop of {hi - lo + 1} instructions; hot path {total}: " + "  ".join(f"{k}={v}" for k, v in sorted(hot.items())))
            print("    " + "  ".join(f"{k}={v}" for k, v in ops.most_common()))
            for n, first in cold:
                print(f"    cold block, {n} instructions, starts with {first}")


if __name__ == "__main__":
    main()
