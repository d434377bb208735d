#!/usr/bin/env python
"""Convert a vector file of the reference's format (one vector per line: id <delim> x0 <delim> x1 ...; SURVEY.md App. E,
vector_reader.hpp:55-85) into the binary columnar file the engine ingests (include/crx.h, "input ingest").

    python tools/csv_to_columnar.py INPUT.csv [OUTPUT.crxcol] [--delimiter ,] [--skip-lines 0]

The default output name is INPUT.csv.crxcol: the drop-in VectorReader (include/crx/lib/in_out/vector_reader.hpp) picks a
file of that name up when it is at least as new as the text file.  Every value is float(token) -- the correctly rounded
double, i.e. what the reference's stod returns -- so a run over the converted file sees bit-identical vectors.  Pure host
code (no GPU): run it once per input file."""
import argparse
import struct
import sys

import numpy as np

MAGIC = b"CRXCOL1\0"


def convert(src, dst, delimiter=",", skip_lines=0):
    ids, rows = [], []
    with open(src, "r", newline="") as f:
        for _ in range(skip_lines):
            f.readline()
        for line in f:
            line = line.rstrip("\n").replace("\r", "")       # vector_reader.hpp:75-76
            head, _, rest = line.partition(delimiter)         # :77-78 (no delimiter: the whole line is id AND body there)
            if not _:
                rest = line
            toks = rest.split(delimiter)
            if toks and toks[-1] == "":                       # getline-based split drops one trailing empty token
                toks.pop()
            ids.append(head)
            rows.append([float(t) for t in toks])
    n = len(rows)
    d = len(rows[0]) if n else 0
    if any(len(r) != d for r in rows) or d < 1:
        raise SystemExit("%s: rows of different lengths (or empty) cannot be stored as columns" % src)
    X = np.asarray(rows, np.float64)
    blob = b"".join(i.encode() + b"\0" for i in ids)
    with open(dst, "wb") as o:
        o.write(MAGIC + struct.pack("<qiiq", n, d, 0, len(blob)))
        o.write(blob + b"\0" * ((8 - len(blob) % 8) % 8))
        o.write(np.ascontiguousarray(X.T).tobytes())
    return n, d


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--delimiter", default=",")
    ap.add_argument("--skip-lines", type=int, default=0, help="metadata lines in front of the vectors (strt_line - 1)")
    a = ap.parse_args()
    n, d = convert(a.src, a.dst or a.src + ".crxcol", a.delimiter, a.skip_lines)
    print("wrote %d vectors x %d coordinates" % (n, d), file=sys.stderr)
