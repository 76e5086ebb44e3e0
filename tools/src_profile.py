"""Summarise `ncu --page source --csv` of one kernel: stall samples by opcode, by stall reason, and the
hottest SASS instructions (developer tool).  usage: src_profile.py rep.ncu-rep kernel_regex [top]"""
import collections, csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[start - 1][:2] if start else "")
hdr, data = rows[start], []
for r in rows[start + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r, k: int(r[ix[k]] or 0)
tot = sum(S(r, "# Samples") for r in data)
print("total samples", tot, "static instrs", len(data), "executed warp-instrs", sum(S(r, "Instructions Executed") for r in data))
by, cnt, ex = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    op = r[ix["Source"]].split()
    o = (op[0] if not op[0].startswith("@") else op[1]).split(".")[0]
    by[o] += S(r, "# Samples"); cnt[o] += 1; ex[o] += S(r, "Instructions Executed")
for o, s in by.most_common(top):
    print(f"{o:10s} samples {s:7d} {100*s/max(tot,1):5.1f}%  static {cnt[o]:4d}  executed {ex[o]}")
for k in hdr:
    if k.startswith("stall_") and "Not Issued" not in k:
        s = sum(S(r, k) for r in data)
        if s * 100 > tot:
            print(f"{k:28s} {s:7d} {100*s/max(tot,1):5.1f}%")
print("hottest instructions:")
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:top]:
    reasons = sorted(((S(r, k), k[6:]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
    print(f"  {S(r,'# Samples'):6d}  {r[ix['Source']].strip()[:70]:70s} {reasons}")
