#!/usr/bin/env python
"""Per-basic-block opcode histogram of one kernel in a `cuobjdump -sass` listing.
usage: cuobjdump -sass x.o > x.sass; python tools/sass_blocks.py x.sass <substring of mangled name> [--pipes]"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("Function : ")
fn = [f for f in txt[1:] if sys.argv[2] in f.split("\n")[0]][0]
ins = []
for line in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
targets = set()
for a, s in ins:
    m = re.search(r"\b(BRA|BRX|JMP|BSSY|CALL)\S*\s.*?(0x[0-9a-f]+)", s)
    if m and s.split()[0].lstrip("@!UP0123456789 ").startswith("BRA") or (m and "BRA" in s):
        targets.add(int(m.group(2), 16))
ALU = ("IADD3", "LOP3", "SHF", "ISETP", "SEL", "PRMT", "LEA", "IABS", "IMNMX", "FLO", "POPC", "BMSK", "SGXT", "IADD", "VIADD", "PLOP3", "MOV", "SHL", "SHR", "VIMNMX", "I2I")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
blocks, cur, start = [], collections.Counter(), ins[0][0]
def flush(end):
    global cur, start
    if sum(cur.values()):
        blocks.append((start, end, cur))
    cur = collections.Counter()
for a, s in ins:
    if a in targets:
        flush(a); start = a
    op = s.split()
    o = op[1] if op[0].startswith("@") else op[0]
    cur[o] += 1
    if o.startswith(("BRA", "EXIT", "BRX", "RET")):
        flush(a + 16); start = a + 16
flush(ins[-1][0] + 16)
tot = collections.Counter()
for s, e, c in blocks:
    n = sum(c.values())
    alu = sum(v for k, v in c.items() if k.split(".")[0] in ALU)
    fma = sum(v for k, v in c.items() if k.split(".")[0] in FMA)
    print(f"block {s:#06x}-{e:#06x}: {n:5d} instr  alu~{alu} fma~{fma} other {n-alu-fma}")
    if n > 40:
        g = collections.Counter()
        for k, v in c.items():
            g[k.split(".")[0] + ("." + k.split(".")[1] if k.startswith("IMAD.") or k.startswith("IADD3.") else "")] += v
        print("     ", dict(g.most_common(18)))
