#!/usr/bin/env python
"""Registers / spills per kernel from a `-Xptxas -v` log:  python scripts/ptxas_summary.py pcetk_b200/csrc/pce_mc.ptxas.log [filter]"""
import re
import subprocess
import sys

text = open(sys.argv[1]).read()
pattern = sys.argv[2] if len(sys.argv) > 2 else ""
entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", text)
names = subprocess.run(["c++filt"], input="\n".join(e[0] for e in entries), capture_output=True, text=True).stdout.split("\n")
for (mangled, stack, st, ld, regs), name in sorted(zip(entries, names), key=lambda x: x[1]):
    name = re.sub(r"\(anonymous namespace\)::|\(.*\)$|void ", "", name)
    if pattern in name:
        print("%-60s regs %3s  stack %3s  spill st/ld %s/%s" % (name, regs, stack, st, ld))
