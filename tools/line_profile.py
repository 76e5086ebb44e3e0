"""Join the per-SASS-instruction stall samples of an ncu report (--page source) with the line table of the cubin
(nvdisasm -g) and print samples / executed instructions per CUDA source line (developer tool).
usage: line_profile.py rep.ncu-rep kernel_regex object.o mangled_kernel_symbol source_file [top]"""
import collections, csv, io, re, subprocess, sys, os, tempfile

rep, pat, obj, sym, src = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], []
for r in rows[start + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for l in sass:
    if l.startswith(".text."):
        active = l.startswith(".text." + sym + ":")
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
print("sass instrs in report", len(data), "in cubin", len(lines))
n = min(len(data), len(lines))
samp, exe = collections.Counter(), collections.Counter()
for i in range(n):
    samp[lines[i]] += int(data[i][ix["# Samples"]] or 0)
    exe[lines[i]] += int(data[i][ix["Instructions Executed"]] or 0)
tot = sum(samp.values())
text = open(src).read().splitlines()
for (f, ln), s in samp.most_common(top):
    t = text[ln - 1].strip()[:100] if f == os.path.basename(src) and ln <= len(text) else ""
    print(f"{100*s/tot:5.1f}%  exec {exe[(f, ln)]:11d}  {f}:{ln:4d}  {t}")
