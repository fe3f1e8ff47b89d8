"""Aggregate an ncu SASS source page (--page source --csv) by the OUTERMOST CUDA source line of each instruction's inline
chain (nvdisasm -gi of the cubin the kernel lives in), optionally restricted to instructions executed at least
`minfrac` x the most executed one (the hot loop).
usage: sass_by_outer_line.py <src.csv> <cubin> <mangled-kernel-substring> [top] [minfrac]"""
import csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
minfrac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
sass = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > iI:
        sass.append((int(r[iI] or 0), int(r[iS] or 0), r[iSrc]))
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if ".text." in l and kname in l and (l.startswith("\t.text.") or ".section" in l))
lines, chain, fresh = [], [("?", 0)], True
for l in dis[start + 1:]:
    if ".section" in l and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if fresh: chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))   # innermost frame first, outermost last
        continue
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        # the frame just inside the kernel body's call of the `iterate` lambda if there is one, else the outermost frame
        cur = chain[-2] if len(chain) >= 2 and chain[-1][0] == chain[-2][0] else chain[-1]
        lines.append((cur, m2.group(2)))
        fresh = True
print("sass in report", len(sass), "sass in cubin", len(lines))
mx = max(s[0] for s in sass)
agg, ops = {}, {}
for (ins, smp, txt), (loc, t2) in zip(sass, lines):
    if ins < minfrac * mx:
        continue
    a = agg.setdefault(loc, [0, 0, 0]); a[0] += ins; a[1] += smp; a[2] += 1
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("warp-inst", tot, "samples", ts, "per most-executed instruction:", tot / mx)
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/terrarium.jl_b200/csrc/{f}").read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print(f"{a[0]/mx:6.1f} inst {100*a[0]/tot:5.1f}%  {100*a[1]/max(ts,1):5.1f}% stall  n={a[2]:4d} {f}:{ln}: {text}")
