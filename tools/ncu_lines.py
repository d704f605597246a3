#!/usr/bin/env python
"""Per-source-line view of one ncu report: warp-instructions executed and stall samples, SASS in address order, each instruction
labelled with the line of `main_file` it was last preceded by (inlined helpers are charged to their call site's neighbourhood).
    python tools/ncu_lines.py rep.ncu-rep dmk_fd_mma.cuh [users]"""
import collections, csv, io, subprocess, sys
rep, main_file = sys.argv[1], sys.argv[2]
users = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; hdr = None; sass = []; cur_line = None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iA = 2; iS = 3; iSm = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    if r[0] != "": cur_line = int(r[0]); continue
    if r[iA].startswith("0x"):
        sass.append((int(r[iA], 16), r[iS].strip(), int(r[iI] or 0), int(r[iSm] or 0), cur_file, cur_line))
sass.sort()
seen = {}
for rec in sass:                      # an instruction can be listed under several lines of its inline stack: keep the main file's
    if rec[0] not in seen or (rec[4] == main_file and seen[rec[0]][4] != main_file): seen[rec[0]] = rec
sass = [seen[a] for a in sorted(seen)]
tot_i = sum(s[2] for s in sass); tot_s = sum(s[3] for s in sass)
agg = collections.OrderedDict(); label = 0
for a, ins, n, smp, f, l in sass:
    if f == main_file: label = l
    k = label
    e = agg.setdefault(k, [0, 0, collections.Counter()])
    e[0] += n; e[1] += smp; e[2][ins.split()[1].split(".")[0] if ins.startswith("@") else ins.split()[0].split(".")[0]] += n
print(f"total warp-instructions {tot_i} ({tot_i / users:.0f} per user), samples {tot_s}")
src = open(f"deepmimo_b200/csrc/{main_file}").read().split("\n")
for k in sorted(agg):
    n, smp, ops = agg[k]
    if n < tot_i * 0.004 and smp < tot_s * 0.004: continue
    top = ",".join(f"{o}:{c / users:.0f}" for o, c in ops.most_common(4))
    print(f"{k:4d} {n / users:8.1f}/user {100 * n / tot_i:5.1f}%  samples {100 * smp / max(tot_s, 1):5.1f}%  [{top}]  {src[k - 1].strip()[:90] if k else ''}")
