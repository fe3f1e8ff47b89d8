"""Aggregate an ncu SASS source page (--page source --csv) by CUDA source line, using nvdisasm -g line info
of the cubin the kernel lives in.  usage: sass_by_line.py <src.csv> <cubin> <mangled-kernel-substring> [top]"""
import csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(src_csv)))
# first kernel block only
hdr = rows[1]
iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
sass = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > iI:
        sass.append((int(r[iI] or 0), int(r[iS] or 0), r[iSrc]))
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith("\t.text.") and kname in l) if any(l.startswith("\t.text.") and kname in l for l in dis) else None
if start is None:
    start = next(i for i, l in enumerate(dis) if ".section" in l and ".text." in l and kname in l)
lines = []
cur = ("?", 0)
inl = ""
for l in dis[start + 1:]:
    if ".section" in l and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); inl = m.group(3)
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not l.strip().startswith("/*") is False:
        pass
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        lines.append((cur, inl, m2.group(2)))
print("sass in report", len(sass), "sass in cubin", len(lines))
agg = {}
n = min(len(sass), len(lines))
for (ins, smp, txt), (loc, inl, t2) in zip(sass[:n], lines[:n]):
    a = agg.setdefault(loc, [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += 1
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-inst", tot, "samples", ts)
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/terrarium.jl_b200/csrc/{f}").read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% stall  n={a[2]:4d} {f}:{ln}: {text}")
