#!/usr/bin/env python
"""Executed warp instructions (and stall samples) per CUDA source line of one ncu report: the per-line view that
profiles/README.md quotes.   python scripts/ncu_source_lines.py report.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys


def main():
    report = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    path, header, out, total, samples = "", None, [], 0, 0
    for row in rows:
        if len(row) == 2 and row[0] == "File Path":
            path = row[1].split("/")[-1]
            continue
        if len(row) > 8 and row[0] == "Line No":
            header = row
            continue
        if header is None or len(row) < len(header) or not row[0].isdigit():
            continue
        index = {h: i for i, h in enumerate(header)}
        try:
            count = int(row[index["Instructions Executed"]] or 0)
            stall = int(row[index["# Samples"]] or 0)
        except ValueError:
            continue
        total += count
        samples += stall
        out.append((count, stall, path, row[0], row[1].strip()[:140]))
    print("total warp instructions %d, stall samples %d" % (total, samples))
    for count, stall, path, line, text in sorted(out, reverse=True)[:top]:
        print("%6.2f%% inst %6.2f%% samples  %s:%s  %s" % (100.0 * count / max(total, 1), 100.0 * stall / max(samples, 1), path, line, text))


if __name__ == "__main__":
    main()
