"""Summarise an ncu source-page CSV: instruction mix by opcode and top stall sites.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_mix.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
ex, src, smp = ci["Instructions Executed"], ci["Source"], ci["# Samples"]
ops, samples = collections.Counter(), collections.Counter()
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
for r in body:
    if len(r) <= max(ex, src, smp):
        continue
    toks = r[src].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    parts = op.split(".")
    key = parts[0]
    if key in ("IMAD", "LDS", "STS", "LDG", "STG", "SHF", "ISETP", "LEA") and len(parts) > 1:
        key += "." + parts[1]
    ops[key] += int(r[ex])
    samples[key] += int(r[smp])
tot, stot = sum(ops.values()), sum(samples.values())
print("warp-instructions executed: %d, stall samples: %d" % (tot, stot))
for k, v in ops.most_common(28):
    print("  %-14s %12d  %5.1f%% of instr  %5.1f%% of samples" % (k, v, 100.0 * v / tot, 100.0 * samples[k] / max(stot, 1)))
print("top sampled instructions:")
for r in sorted([x for x in body if len(x) > smp], key=lambda r: -int(r[smp]))[:14]:
    print("  %6s  %s" % (r[smp], r[src].strip()[:90]))
