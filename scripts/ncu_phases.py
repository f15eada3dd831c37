"""Split an `ncu --page source --csv` SASS export into regions delimited by BAR.SYNC and print, per region, the
share of stall samples, the stall reasons and the opcode mix.  Usage: python scripts/ncu_phases.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
names = ["stall_wait", "stall_math", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_selected", "stall_not_selected",
         "stall_branch_resolving", "stall_dispatch", "stall_mio", "stall_no_inst"]
cols = {n: hdr.index(n) for n in names}
new = lambda: {"samples": 0, "n": 0, "ops": {}, "st": {k: 0 for k in cols}}
reg, cur, tot = [], new(), 0
for r in rows:
    if len(r) <= max(cols.values()) or not r[isamp].isdigit():
        continue
    s = r[ia].strip()
    t = s.split()
    op = (t[1] if s.startswith("@") else t[0]).split(".")[0]
    smp, ex = int(r[isamp]), int(r[iex] or 0)
    cur["samples"] += smp; cur["n"] += 1; tot += smp
    cur["ops"][op] = cur["ops"].get(op, 0) + ex
    for k, i in cols.items():
        cur["st"][k] += int(r[i] or 0)
    if "BAR.SYNC" in s or s.startswith("EXIT"):
        reg.append(cur); cur = new()
reg.append(cur)
for i, g in enumerate(reg):
    if g["n"] == 0: continue
    ops = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:10]
    print(f"region {i}: samples {g['samples']} ({100 * g['samples'] / max(tot, 1):.1f}%)  sass instrs {g['n']}  ",
          {k[6:]: v for k, v in g["st"].items() if v > 0.03 * g["samples"]})
    print("     ", ops)
