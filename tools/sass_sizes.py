#!/usr/bin/env python
"""Static SASS instruction count per device function of one kernel (nvdisasm labels). usage: sass_sizes.py <obj> <kernel-substr>"""
import collections, os, re, subprocess, sys, tempfile
obj, ksub = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt = collections.OrderedDict()
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    sec = cur = None
    for ln in subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sec = m.group(1); cur = sec
            continue
        if not (sec and ksub in sec):
            continue
        m = re.match(r"^(\$?[_A-Za-z0-9$]+):", ln)
        if m:
            lab = m.group(1)
            if lab.startswith("$") or lab.startswith("_ZN"):
                cur = lab
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            cnt[cur] = cnt.get(cur, 0) + 1
tot = sum(cnt.values())
print("total instructions", tot, "=", tot * 16 // 1024, "KiB")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    name = k.split("$_ZN2km")[-1] if "$_ZN2km" in k else k
    print("%6d  %s" % (v, name[:100]))
