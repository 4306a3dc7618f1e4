"""Where the warp-instructions and the stall samples of a kernel sit: consecutive SASS instructions with the same
execution count are one region (a loop body, a per-tile prologue, ...).
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
       python profiles/tools/sass_regions.py src.csv <kernel name substring> [min share in %]"""
import collections
import csv
import re
import sys

want = sys.argv[2] if len(sys.argv) > 2 else ""
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
kernels, hdr, name = {}, None, None   # (the first launch of every kernel name)
seen = collections.Counter()
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "Kernel Name":
        seen[r[1]] += 1
        name = r[1] if seen[r[1]] == 1 else None
        if name:
            kernels[name] = []
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and name and len(r) > 6:
        kernels[name].append(r)
ix, isamp, ith = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")


def op(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s.strip())
    return s.split()[0].split(".")[0]


for name, rows in kernels.items():
    if want not in name:
        continue
    tot = sum(int(r[ix]) for r in rows)
    tots = sum(int(r[isamp]) for r in rows)
    print(f"{name}: {len(rows)} SASS instructions, {tot} warp-instructions executed, {tots} stall samples")
    i = 0
    while i < len(rows):
        c = int(rows[i][ix])
        j, ops, smp, th = i, [], 0, 0.0
        while j < len(rows) and abs(int(rows[j][ix]) - c) <= 0.03 * max(c, 1):
            ops.append(op(rows[j][1]))
            smp += int(rows[j][isamp])
            th += float(rows[j][ith])
            j += 1
        if c * (j - i) > min_share / 100 * tot:
            h = collections.Counter(ops)
            print(f"  [{i:5d}-{j - 1:5d}] {j - i:4d} instr x {c:8d} = {c * (j - i) / tot * 100:5.1f} % of the instructions, "
                  f"{smp / max(tots, 1) * 100:5.1f} % of the samples, {th / (j - i):4.1f} lanes :: "
                  + " ".join(f"{k}{v}" for k, v in h.most_common(8)))
        i = j
