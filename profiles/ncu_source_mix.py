"""Instruction mix and hot basic blocks from an `ncu --page source --csv` export (SASS view with executed counts and
stall samples).  usage: python profiles/ncu_source_mix.py source.csv [min_Minstr_per_block]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
floor = float(sys.argv[2]) * 1e6 if len(sys.argv) > 2 else 5e6
print(rows[0][1] if len(rows[0]) > 1 else rows[0])


def opcode(src):
    m = re.match(r"\s*(@!?U?P[T\d]+\s+)?([A-Z0-9_]+(?:\.(?:64|128))?)", src)
    op = m.group(2) if m else "?"
    return op if op.startswith(("LDS", "STS")) else op.split(".")[0]


tot, smp = collections.Counter(), collections.Counter()
for r in data:
    tot[opcode(r[iS])] += int(r[iE])
    smp[opcode(r[iS])] += int(r[iSm])
T, S = sum(tot.values()), sum(smp.values())
fp64 = sum(v for k, v in tot.items() if k in ("DFMA", "DMUL", "DADD"))
print(f"warp instructions {T}  (DFMA+DMUL+DADD {fp64} = {100 * fp64 / T:.1f} %)  stall samples {S}  static {len(data)}")
for k, v in tot.most_common(24):
    print(f"  {k:12s} {v:12d} {100 * v / T:5.1f} %   samples {100 * smp[k] / S:5.1f} %")
print("hot blocks (runs of equal executed count):")
runs, cur = [], None
for idx, r in enumerate(data):
    c = int(r[iE])
    if cur is None or cur[0] != c:
        cur = [c, idx, idx]
        runs.append(cur)
    cur[2] = idx
for c, a, b in runs:
    n = b - a + 1
    if c * n < floor:
        continue
    ops = collections.Counter(opcode(r[iS]) for r in data[a:b + 1])
    s = sum(int(r[iSm]) for r in data[a:b + 1])
    print(f"  lines {a:5d}-{b:5d}  executed {c:9d} x {n:4d} = {c * n / 1e6:7.1f} M  samples {100 * s / S:5.1f} %  "
          + " ".join(f"{k}:{v}" for k, v in ops.most_common(8)))
