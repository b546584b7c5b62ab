"""SASS excerpt of the headline kernel out of the shipped library: the bulk-copy (TMA engine, 1-D) and mbarrier
instructions, the work-counter atomic, and a static opcode histogram.
usage: python tools/sass_excerpt.py > profiles/polymul_sass_r2.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "libsafecrypto_b200", "libscgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pick = [f for f in funcs[1:] if "k_polymul_w32" in f.split("\n")[0] and "ArFqELi9ELi0ELb1ELb1ELb0" in f.split("\n")[0]]
f = pick[0]
name = f.split("\n")[0]
ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
hist = collections.Counter(i.split(".")[0] for i in ins)
print("# profiles/polymul_sass_r2.txt -- cuobjdump -sass of libscgpu.so, kernel k_polymul_w32<ArFq, 9, FQ_POLYMUL, TMA=true, BM=true, CHK=false>")
print("# (sm_100a; the headline kernel of bench.py: base multiplication, no range vote).  %d SASS instructions." % len(ins))
print("# TMA engine = UBLKCP (cp.async.bulk, 1-D rows), completion through mbarrier = SYNCS.*; no constant-bank operands exist")
print("# on this architecture: every kernel-parameter constant arrives by LDC (register) or LDCU (uniform register).")
print("# static opcode histogram: " + ", ".join("%s %d" % kv for kv in hist.most_common(26)))
print()
print("\t\tFunction : " + name)
keep = re.compile(r"UBLKCP|SYNCS|ATOMG|LDCU\.128|FENCE|ELECT|I2FP")
shown = 0
for line in f.split("\n")[1:]:
    if re.search(r"/\*[0-9a-f]{4}\*/", line) and keep.search(line):
        print(line.rstrip())
        shown += 1
# one base multiplication worth of arithmetic, for the reader: the first 40 instructions after the first I2FP
lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
for i, l in enumerate(lines):
    if "I2FP" in l:
        print("...\n# arithmetic around the first int->float conversion of the base multiplication:")
        for m in lines[i:i + 40]:
            print(m.rstrip())
        break
