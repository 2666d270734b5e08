#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an .ncu-rep (captured with --import-source on, code built with -lineinfo).

    python profiles/ncu_by_line.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name substring> [top N]

ncu's CSV source page is SASS-only, so the line table comes from `nvdisasm -g` on the same cubin (cuobjdump -xelf) and
the two instruction streams are joined by offset."""
import collections, csv, io, re, subprocess, sys

rep, kre, cubin, fsub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# first kernel instance only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
inst = []
for r in rows[start + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    inst.append(dict(zip(hdr, r)))
base = int(inst[0]["Address"], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for l in dis:
    if l.startswith(".text."):
        infn = fsub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for d in inst:
    off = int(d["Address"], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    ie, te, smp = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])
    a[0] += ie; a[1] += te; a[2] += smp; a[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += smp
print(f"# {rep} kernel~{kre}: {tot[0]:,} warp-instructions, {tot[1] / max(tot[0], 1):.2f} threads/instr, {tot[2]:,} stall samples")
print(f"{'file:line':>22} {'inst%':>7} {'thr/inst':>8} {'samples%':>8} {'#sass':>6}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{key[0] + ':' + str(key[1]):>22} {100 * a[0] / tot[0]:7.2f} {a[1] / max(a[0], 1):8.2f} {100 * a[2] / max(tot[2], 1):8.2f} {a[3]:6d}")
