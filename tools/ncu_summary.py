#!/usr/bin/env python
"""`ncu -i file.ncu-rep --page raw --csv` -> a three-column `metric,unit,value` summary (what profiles/*_summary.csv hold).
usage: python tools/ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(head, units, vals):
            w.writerow([h, u, v])


if __name__ == "__main__":
    main()
