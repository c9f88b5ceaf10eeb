#!/usr/bin/env python3
"""Join an ncu source-page CSV (SASS view) with nvdisasm line info: instructions executed and
stall samples per device function and per source line.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin-object-name e.g. ag_board> [top_n]
env NCU_LINES_STALL=long_sb (or any stall_* suffix): also list the instructions with the most samples
of that stall reason (the instruction WAITING, with its source line)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, kre, obj = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "aprilgrid-rs_b200", "lib", "libaprilgrid_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(obj + ".") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
# several launches of the same kernel may be in the report: keep the first (or NCU_LINES_LAUNCH = k)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k_launch = int(os.environ.get("NCU_LINES_LAUNCH", "0"))
end = starts[k_launch + 1] if len(starts) > k_launch + 1 else len(rows)
hdr, data = rows[starts[k_launch] + 1], [r for r in rows[starts[k_launch] + 2:end] if len(r) == len(rows[starts[k_launch] + 1])]
iS, iA, iI, iT = (hdr.index(k) for k in ("# Samples", "Address", "Instructions Executed", "Thread Instructions Executed"))
iSrc = hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][iA], 16)
# nvdisasm: find the kernel's section, map offset -> (function label, file, line)
in_kernel, func, loc, mp = False, None, ("?", 0), {}
for ln in sass.splitlines():
    m = re.match(r"^\.text\.(\S+):", ln)
    if m:
        in_kernel = re.search(kre, m.group(1)) is not None
        func = "<kernel>"
        continue
    if not in_kernel:
        continue
    m = re.match(r"^(\$?[_A-Za-z][\w\$]*):", ln)
    if m and not m.group(1).startswith(".L"):
        name = m.group(1)
        func = name.split("$")[-1] if "$" in name else "<kernel>"
        continue
    m = re.match(r'^\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        mp[int(m.group(1), 16)] = (func, loc)
by_func, by_line = defaultdict(lambda: [0, 0, 0, 0]), defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
stall_tot = defaultdict(int)
per_inst = []
for r in data:
    off = int(r[iA], 16) - base
    f, l = mp.get(off, ("?", ("?", 0)))
    v = [int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0), 0]
    bar = sum(int(r[i] or 0) for i, h in stall_cols if h == "stall_barrier")
    v[3] = bar
    for agg in (by_func[f], by_line[(f, l)], tot):
        for k in range(4):
            agg[k] += v[k]
    for i, h in stall_cols:
        stall_tot[h] += int(r[i] or 0)
    if os.environ.get("NCU_LINES_STALL"):
        want = "stall_" + os.environ["NCU_LINES_STALL"]
        for i, h in stall_cols:
            if h == want and int(r[i] or 0):
                per_inst.append((int(r[i]), f, l, r[iSrc].strip(), int(r[iI] or 0)))
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
print("total: samples %d (barrier %d)  warp-inst %d  thread-inst %d  avg active threads %.1f" %
      (tot[0], tot[3], tot[1], tot[2], tot[2] / max(tot[1], 1)))
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * c / max(tot[0], 1))
                           for h, c in sorted(stall_tot.items(), key=lambda x: -x[1])[:8]))
print("\nper device function (samples excl. barrier | warp-inst):")
for f, v in sorted(by_func.items(), key=lambda x: -x[1][1]):
    print("  %-34s samples %6.2f%%  nonbar %6.2f%%  inst %6.2f%%  (%d)  thr/inst %.1f" %
          (demangle(f)[-34:], 100.0 * v[0] / tot[0], 100.0 * (v[0] - v[3]) / max(tot[0] - tot[3], 1),
           100.0 * v[1] / tot[1], v[1], v[2] / max(v[1], 1)))
print("\ntop source lines by warp-inst:")
for (f, l), v in sorted(by_line.items(), key=lambda x: -x[1][1])[:topn]:
    print("  %-24s %s:%-5d inst %5.2f%%  nonbar-samples %5.2f%%  thr/inst %.1f" %
          (demangle(f)[-24:], l[0], l[1], 100.0 * v[1] / tot[1], 100.0 * (v[0] - v[3]) / max(tot[0] - tot[3], 1),
           v[2] / max(v[1], 1)))
if per_inst:
    want = os.environ["NCU_LINES_STALL"]
    n_all = sum(p[0] for p in per_inst)
    print("\ninstructions waiting with stall_%s (%d samples = %.1f%% of all):" % (want, n_all, 100.0 * n_all / tot[0]))
    by_l = defaultdict(int)
    for c, f, l, src, ex in per_inst:
        by_l[(f, l)] += c
    for (f, l), c in sorted(by_l.items(), key=lambda x: -x[1])[:topn]:
        print("  %-24s %s:%-5d %5.1f%% of that stall" % (demangle(f)[-24:], l[0], l[1], 100.0 * c / n_all))
    print("  -- by instruction:")
    for c, f, l, src, ex in sorted(per_inst, key=lambda x: -x[0])[:topn]:
        print("  %5.1f%%  %s:%-5d exec %-8d %s" % (100.0 * c / n_all, l[0], l[1], ex, src[:80]))
