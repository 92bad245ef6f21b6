#!/usr/bin/env python
"""Static SASS instruction count per source line of one device function (nvdisasm -g). usage: sass_by_line.py <obj> <function-substr> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, fsub = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt = collections.Counter()
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    cur = None; infn = False
    for ln in subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"^(\$?[_A-Za-z0-9$]+):", ln)
        if m:
            lab = m.group(1)
            if lab.startswith("$_Z") or lab.startswith("_Z"):
                infn = fsub in lab.split("$")[-1]
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            cnt[cur] += 1
print("total", sum(cnt.values()))
for k, v in cnt.most_common(top):
    print("%6d  %s:%d" % (v, k[0], k[1]) if k else "%6d  ?" % v)
