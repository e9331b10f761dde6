"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by enclosing function of rr_kernels.inl.
    python tools/ncu_by_function.py gpurun_out/prof_src.csv
Prints executed warp instructions, stall samples and shared-memory wavefront excess per function."""
import bisect
import collections
import csv
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "brax_rodent_run_b200", "csrc", "rr_kernels.inl")).read().split("\n")
starts = []
for i, line in enumerate(src, 1):
    m = re.match(r"RR_DEV(?:_MEMBER)? [\w:<> ,\*&]*?(\w+)\(", line)
    if m:
        starts.append((i, m.group(1)))
lines_ = [s[0] for s in starts]


def func_of(line):
    k = bisect.bisect_right(lines_, line) - 1
    return starts[k][1] if k >= 0 else "?"


hdr = None
cur_file = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
per_line = collections.defaultdict(lambda: [0, 0])
cur_line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r
        ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
        cw, cwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0] != "":
        cur_line = int(r[0])
        continue  # source row = sum of its SASS rows
    key = func_of(cur_line) if cur_file == "rr_kernels.inl" else cur_file
    def num(x):
        try: return float(x)
        except Exception: return 0.0
    a = agg[key]
    a[0] += num(r[ci]); a[1] += num(r[cs]); a[2] += num(r[cw]); a[3] += num(r[cwi])
    if cur_file == "rr_kernels.inl":
        per_line[cur_line][0] += num(r[ci]); per_line[cur_line][1] += num(r[cs])
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot_i:.3e}  samples {tot_s:.0f}")
print(f"{'function':28s} {'inst%':>7s} {'samples%':>9s} {'cyc/inst(rel)':>13s} {'smem wavefronts / ideal':>24s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[0] == 0 and a[1] == 0: continue
    print(f"{k:28s} {100*a[0]/tot_i:7.2f} {100*a[1]/tot_s:9.2f} {(a[1]/tot_s)/(a[0]/tot_i+1e-12):13.2f} {a[2]/(a[3]+1e-9):24.2f}")
if len(sys.argv) > 2:
    print("\nhottest source lines (by stall samples):")
    for ln, (i, s) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2])]:
        print(f"{ln:5d} {100*i/tot_i:6.2f}% inst {100*s/tot_s:6.2f}% samples | {src[ln-1].strip()[:110]}")
