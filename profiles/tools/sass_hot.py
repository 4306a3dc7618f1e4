"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`:
top SASS instructions by stall samples, shared-memory wavefronts and a per-opcode summary."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def num(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError):
        return 0.0


tot_s = sum(num(r, "# Samples") for r in data)
tot_i = sum(num(r, "Instructions Executed") for r in data)
tot_w = sum(num(r, "L1 Wavefronts Shared") for r in data)
print(f"samples {tot_s:.0f}  warp-instructions {tot_i:.0f}  shared wavefronts {tot_w:.0f} "
      f"(ideal {sum(num(r, 'L1 Wavefronts Shared Ideal') for r in data):.0f})")
print("--- top by samples")
for k, r in sorted(enumerate(data), key=lambda kr: -num(kr[1], "# Samples"))[:top]:
    st = {h[6:]: num(r, h) for h in hdr if h.startswith("stall_") and "(" not in h and num(r, h) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{k:5d} {num(r, '# Samples'):7.0f} {100 * num(r, '# Samples') / tot_s:5.1f}%  inst {num(r, 'Instructions Executed'):9.0f} "
          f"shw {num(r, 'L1 Wavefronts Shared'):9.0f}  {r[ix['Source']][:70]:70s} {st}")
print("--- shared-memory wavefronts by instruction")
for k, r in sorted(enumerate(data), key=lambda kr: -num(kr[1], "L1 Wavefronts Shared"))[:15]:
    print(f"{k:5d} shw {num(r, 'L1 Wavefronts Shared'):9.0f} ideal {num(r, 'L1 Wavefronts Shared Ideal'):9.0f} inst {num(r, 'Instructions Executed'):9.0f}  {r[ix['Source']][:80]}")
ops = Counter()
for r in data:
    src = r[ix["Source"]].split()
    op = next((w for w in src if not w.startswith("@") and not w.startswith("/*")), "?").split(".")[0]
    ops[op] += num(r, "Instructions Executed")
print("--- warp-instructions by opcode")
print(", ".join(f"{o} {v / tot_i * 100:.1f}%" for o, v in ops.most_common(25)))
