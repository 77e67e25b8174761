#!/usr/bin/env python
"""Registers / spills per kernel from `make` output (-Xptxas -v).  usage: make -C toppra_b200/csrc 2>&1 | python scripts/ptxas_report.py [regex]"""
import re
import subprocess
import sys

pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
name = None
rows = []
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("tb::(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        spill = ""
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = "stack %s spill st %s ld %s" % m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        if pat is None or pat.search(name):
            rows.append((name, int(m.group(1)), spill))
        name = None
for r in rows:
    print("%-70s %4d regs  %s" % r)
