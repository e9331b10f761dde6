"""Per-function and per-line stall-reason breakdown of an `ncu --page source --csv --print-source sass,cuda` export.
    python tools/ncu_stalls.py gpurun_out/prof_src.csv [top_lines]
"""
import bisect, collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
srcp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "brax_rodent_run_b200", "csrc", "rr_kernels.inl")
src = open(srcp).read().split("\n")
starts = []
for i, line in enumerate(src, 1):
    m = re.match(r"RR_DEV(?:_MEMBER|_NOINLINE)? [\w:<> ,\*&]*?(\w+)\(", line)
    if m: starts.append((i, m.group(1)))
lines_ = [s[0] for s in starts]
def func_of(n):
    k = bisect.bisect_right(lines_, n) - 1
    return starts[k][1] if k >= 0 else "?"
hdr = None; cur_file = None; cur_line = None
REASONS = ["stall_barrier","stall_branch_resolving","stall_dispatch","stall_lg","stall_long_sb","stall_math","stall_mio","stall_no_inst","stall_not_selected","stall_selected","stall_short_sb","stall_wait","stall_sleep","stall_misc"]
agg = collections.defaultdict(lambda: collections.Counter()); per_line = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; idx = {}
        for k, name in enumerate(hdr):
            if name not in idx: idx[name] = k
        continue
    if hdr is None or r[0] == "Function Name": continue
    if r[0] != "": cur_line = int(r[0]); continue
    key = func_of(cur_line) if cur_file == "rr_kernels.inl" else cur_file
    def num(name):
        try: return float(r[idx[name]])
        except Exception: return 0.0
    a = agg[key]; a["inst"] += num("Instructions Executed"); a["samples"] += num("# Samples")
    a["wf"] += num("L1 Wavefronts Shared"); 
    sass = r[3] if len(r) > 3 else ""
    if re.search(r"\b(LDS|STS|SHFL|LDSM|ATOMS)\b", sass): a["mio"] += num("Instructions Executed")
    if "SHFL" in sass: a["shfl"] += num("Instructions Executed")
    for q in REASONS: a[q] += num(q)
    if cur_file == "rr_kernels.inl":
        pl = per_line[cur_line]; pl["inst"] += num("Instructions Executed"); pl["samples"] += num("# Samples")
        for q in REASONS: pl[q] += num(q)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samples"] for a in agg.values())
print(f"total warp instructions {ti:.3e} samples {ts:.0f}")
short = [q.replace("stall_", "")[:8] for q in REASONS]
print(f"{'function':22s} {'inst%':>6s} {'smp%':>6s} {'mio%i':>6s} {'shfl%i':>6s} " + " ".join(f"{s:>8s}" for s in short))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:28]:
    if a["samples"] < 0.002 * ts: continue
    print(f"{k:22s} {100*a['inst']/ti:6.2f} {100*a['samples']/ts:6.2f} {100*a['mio']/max(a['inst'],1):6.1f} {100*a['shfl']/max(a['inst'],1):6.1f} " + " ".join(f"{100*a[q]/ts:8.2f}" for q in REASONS))
tot = collections.Counter()
for a in agg.values():
    for q in REASONS: tot[q] += a[q]
    tot["mio"] += a["mio"]; tot["shfl"] += a["shfl"]
print(f"{'TOTAL':22s} {100.0:6.2f} {100.0:6.2f} {100*tot['mio']/ti:6.1f} {100*tot['shfl']/ti:6.1f} " + " ".join(f"{100*tot[q]/ts:8.2f}" for q in REASONS))
print("\nhottest lines:")
for ln, pl in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    best = sorted(REASONS, key=lambda q: -pl[q])[:3]
    print(f"{ln:5d} {100*pl['inst']/ti:5.2f}%i {100*pl['samples']/ts:5.2f}%s  " + " ".join(f"{q.replace('stall_','')}={100*pl[q]/max(pl['samples'],1):.0f}%" for q in best) + f" | {src[ln-1].strip()[:90]}")
