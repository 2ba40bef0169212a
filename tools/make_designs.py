#!/usr/bin/env python3
"""Regenerate the packaged spherical-design table from the reference's node files.

Reads the nine `ssTTT.NNN.txt` files the reference's SphericalDesign class loads at run time
(/root/reference/Quadratures/SphericalDesign.cpp:12-24, one "x y z" row per point in `%.16e`)
and writes `boltzmann-fourier-spectral-method_b200/data/spherical_designs.json`:

    {"<N>": {"degree": t, "antipodal": true|false, "xyz_hex": [[hx, hy, hz], ...]}, ...}

Coordinates are stored as C99 hex floats (`float.hex()`), i.e. bit-exact doubles, so the
product's host side does not depend on /root/reference (which does not exist on the GPU box).
`antipodal` records whether row i + N/2 == -row i bit-for-bit; the CUDA path folds such
designs to N/2 transformed directions ("antipodal folding").

Run from the repo root in the build container:  python tools/make_designs.py
"""
import glob
import json
import os
import re
import sys

REF = os.environ.get("BFSM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "boltzmann-fourier-spectral-method_b200", "data", "spherical_designs.json")


def main():
    files = sorted(glob.glob(os.path.join(REF, "Quadratures", "ss*.txt")))
    if not files:
        sys.exit(f"no design files under {REF}/Quadratures")
    table = {}
    for path in files:
        m = re.match(r"ss(\d+)\.(\d+)\.txt", os.path.basename(path))
        degree, n = int(m.group(1)), int(m.group(2))
        rows = []
        with open(path) as fh:
            for line in fh:
                parts = line.split()
                if len(parts) == 3:
                    rows.append([float(v) for v in parts])
        assert len(rows) == n, (path, len(rows), n)
        half = n // 2
        antipodal = n % 2 == 0 and all(
            rows[i + half][c] == -rows[i][c] for i in range(half) for c in range(3))
        table[str(n)] = {
            "degree": degree,
            "antipodal": bool(antipodal),
            "xyz_hex": [[v.hex() for v in r] for r in rows],
        }
        print(f"{os.path.basename(path)}: N={n} degree={degree} antipodal={antipodal}")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        json.dump(table, fh, indent=0, separators=(",", ":"))
        fh.write("\n")
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
