#!/usr/bin/env python3
"""Static SASS instruction count per source line of one file, attributing inlined callees to their call site.
usage: tools/sass_by_line.py <lib.so> <kernel-substring> <source-file-name> [min_count]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
minc = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt = collections.Counter()
inside = False
site = None
in_group = False
chain = []
for l in sass.splitlines():
    if l.startswith("//-") and ".text." in l:
        inside = kern in l
        continue
    if not inside:
        continue
    if "//## File" in l:
        if not in_group:                    # a new group of consecutive //## lines = one inline chain, inner -> outer
            chain, in_group = [], True
        chain += re.findall(r'"([^"]+)", line (\d+)', l)
        site = None
        for f, ln in chain:                 # keep the outermost entry inside `src`
            if f.endswith(src):
                site = int(ln)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        in_group = False
        cnt[site] += 1
lines = open([os.path.join(r, src) for r, _, fs in os.walk(os.path.dirname(os.path.abspath(so))) if src in fs][0]).read().splitlines()
print("total SASS instructions in kernel:", sum(cnt.values()))
for k in sorted(cnt, key=lambda x: (x is None, x)):
    if cnt[k] >= minc:
        print("%5s %5d  %s" % (k, cnt[k], lines[k - 1].strip()[:110] if k else "(other files)"))
