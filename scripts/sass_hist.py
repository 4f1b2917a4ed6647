#!/usr/bin/env python
"""Opcode histogram per kernel of a cubin / shared library (cuobjdump -sass): evidence for instruction-mix claims.
usage: sass_hist.py lib.so [regex-of-kernel-names] -> one line per kernel: total, then the top opcodes"""
import collections, re, subprocess, sys
lib = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, hist = None, None
def flush():
    if name and (pat is None or pat.search(name)):
        tot = sum(hist.values())
        top = " ".join(f"{k}={v}" for k, v in hist.most_common(22))
        print(f"{name}\n  total={tot} {top}")
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush()
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and hist is not None:
        hist[m.group(1)] += 1
flush()
