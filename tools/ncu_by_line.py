#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line / function using nvdisasm line info.

usage: ncu_by_line.py <report.ncu-rep> <object-or-so-with-the-kernel> <mangled-kernel-substring> [top]
"""
import csv, re, subprocess, sys, tempfile, os, collections

rep, obj, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
lineof = {}
for cb in cubins:
    dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, None
    for ln in dis:
        if ln.startswith("\t.section\t.text."):
            inside = ksub in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            lineof[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci = {n: H.index(n) for n in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for r in rows[hdr + 1:]:
    if len(r) <= ci["# Samples"]:
        continue
    try:
        addr = int(r[ci["Address"]], 16) if not r[ci["Address"]].isdigit() else int(r[ci["Address"]])
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    key = lineof.get(off, (None, ""))[0]
    vals = []
    for n in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"):
        try:
            vals.append(float(r[ci[n]] or 0))
        except ValueError:
            vals.append(0)
    for k in range(5):
        agg[key][k] += vals[k]
        tot[k] += vals[k]
print("total inst %.3e thread-inst %.3e samples %d smem wavefronts %.3e ideal %.3e" % tuple(tot))
# read source text
srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_kmanip_b200", "csrc")
text = {}
for f in os.listdir(srcdir):
    p = os.path.join(srcdir, f)
    if os.path.isfile(p):
        text[f] = open(p).read().splitlines()
print("%-22s %7s %7s %7s %6s  %s" % ("file:line", "inst%", "samp%", "thr/in", "bankx", "source"))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    f, l = key if key else ("?", 0)
    s = text.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in text and l > 0 else ""
    print("%-22s %7.2f %7.2f %7.1f %6.2f  %s" % (f"{f}:{l}", 100 * v[0] / tot[0], 100 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1), v[3] / max(v[4], 1), s))

# ---- per-function totals (line ranges of KM_TPL definitions in km_sim.cuh)
import bisect
defs = []
for f in ("km_sim.cuh", "km_launch.cuh"):
    for n, ln in enumerate(text.get(f, []), 1):
        m_ = re.match(r"^(?:KM_TPL|template <[^>]*>) (?:KM_FN|KM_HD|__global__|__device__)[^(]*?(\w+)\(", ln)
        if m_:
            defs.append((f, n, m_.group(1)))
byf = collections.defaultdict(lambda: [0, 0, 0])
for key, v in agg.items():
    if not key:
        name = "?"
    else:
        f, l = key
        cands = [d for d in defs if d[0] == f and d[1] <= l]
        name = cands[-1][2] if cands else f
    byf[name][0] += v[0]; byf[name][1] += v[1]; byf[name][2] += v[2]
print("\n%-28s %7s %7s %7s" % ("function (incl. inlined helpers by file)", "inst%", "samp%", "thr/in"))
for name, v in sorted(byf.items(), key=lambda kv: -kv[1][2]):
    print("%-28s %7.2f %7.2f %7.1f" % (name, 100 * v[0] / tot[0], 100 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))
