#!/usr/bin/env python
"""Static size of the trial loop of one k_mc_* variant: builds nothing, reads an object file.

usage: sass_loop.py OBJECT SUBSTRING [--dump]
Finds the function whose mangled name contains SUBSTRING, the largest backward-branch region that contains MUFU.EX2 (the
Metropolis screen), and prints instruction counts by opcode class for that region."""
import collections
import re
import subprocess
import sys


def disassemble(obj):
    return subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout


def loop_region(text, key):
    """[(address, instruction text)] of the trial / round loop of the function whose name contains `key`."""
    functions = re.split(r"\n\s*Function : ", text)
    body = next(f for f in functions if key in f.split("\n", 1)[0])
    instr = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            instr.append((int(m.group(1), 16), m.group(2).strip()))
    address = {a: i for i, (a, _) in enumerate(instr)}
    best = None
    for i, (a, s) in enumerate(instr):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)*(?:0x)([0-9a-f]+)", s)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in address:
            j = address[int(m.group(1), 16)]
            if any("MUFU.EX2" in t for _, t in instr[j:i + 1]) and (best is None or i - j > best[1] - best[0]):
                best = (j, i)
    j, i = best
    return instr[j:i + 1]


def main():
    obj, key = sys.argv[1], sys.argv[2]
    region = loop_region(disassemble(obj), key)
    classes = collections.Counter()
    for _, s in region:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        classes[op.split(".")[0]] += 1
    print("%s: loop %#x..%#x, %d instructions" % (key, region[0][0], region[-1][0], len(region)))
    print("  " + ", ".join("%s %d" % kv for kv in classes.most_common(14)))
    if "--dump" in sys.argv:
        for a, s in region:
            print("%05x  %s" % (a, s))


if __name__ == "__main__":
    main()
