#!/usr/bin/env python3
"""Builds tests/golden/matviz_pressure_system.npz from the only golden-vector-like fixture the reference
ships: devDocs/matviz/{data,vin,vout}.txt -- a dump of one real 64x64 pressure system (402 fluid rows
`idx:diag iNeg iPos jNeg jPos`, matrix scale 0.109227), its right-hand side and the pressure the
reference's solver returned for it. Run in the container that has /root/reference."""
import os
import re
import sys

import numpy as np

SRC = os.path.join(os.environ.get("FS2D_REFERENCE_ROOT", "/root/reference"), "devDocs", "matviz")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "matviz_pressure_system.npz")


def floats(path):
    vals = []
    for line in open(path):
        line = line.strip()
        if not line or "=" in line:
            continue
        vals.append(float(line))
    return np.array(vals, np.float64)


def main():
    row = re.compile(r"(\d+):(-?\d+\.*\d*) (-?\d+\.*\d*) (-?\d+\.*\d*) (-?\d+\.*\d*) (-?\d+\.*\d*)")
    size = None
    rows = []
    for line in open(os.path.join(SRC, "data.txt")):
        m = row.search(line)
        if m:
            rows.append([float(g) for g in m.groups()])
        elif "," in line and "=" not in line:
            size = [int(v) for v in line.split(",")]
    rows = np.array(rows)
    np.savez_compressed(OUT, size=np.array(size), index=rows[:, 0].astype(np.int64), diag=rows[:, 1], i_neg=rows[:, 2],
                        i_pos=rows[:, 3], j_neg=rows[:, 4], j_pos=rows[:, 5], vin=floats(os.path.join(SRC, "vin.txt")),
                        vout=floats(os.path.join(SRC, "vout.txt")))
    print("wrote", OUT, "rows", len(rows), "size", size)


if __name__ == "__main__":
    sys.exit(main())
