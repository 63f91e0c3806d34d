#!/usr/bin/env python
"""Extract one kernel's SASS from a .so (cuobjdump -sass) and tally opcodes.
usage: sass_fn.py <lib.so> <mangled-name substring> [--dump]"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
out, on = [], False
for l in txt:
    if "Function :" in l:
        on = pat in l
        if on:
            print(l.strip())
    elif on:
        out.append(l)
ops = collections.Counter()
n = 0
for l in out:
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m:
        ops[m.group(3)] += 1
        n += 1
print("instructions:", n)
print(" ".join(f"{k}:{v}" for k, v in ops.most_common()))
if "--dump" in sys.argv:
    for l in out:
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
