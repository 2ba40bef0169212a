"""Host-side quadrature setup, mirroring the reference's Quadratures/ classes.

`GaussLegendreQuadrature(n, a, b)` follows Quadratures/GaussLegendre.hpp:7-31 (nodes ascending,
x = B -/+ A*xhat, w = A*what with A=(b-a)/2, B=(a+b)/2) without GSL: the rule on [-1,1] is
found by Newton iteration on P_n in extended precision.  `SphericalDesign(N)` follows
Quadratures/SphericalDesign.cpp:6-50 (equal weights 4*pi/N) but reads the node table packaged in
data/spherical_designs.json (bit-exact copies of the nine ssTTT.NNN.txt tables, see
tools/make_designs.py) or, if a directory is given, the reference's text format.

Method names (getNodes, getWeights, getx, ...) are the reference's (AbstractQuadrature.hpp:8-47,
AbstractSphericalQuadratures.hpp:11-61).
"""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_DESIGN_JSON = os.path.join(HERE, "data", "spherical_designs.json")

pi = 3.14159265358979323846  # Utilities/constants.hpp:7

#: point count -> design degree (file ssTTT.NNN.txt), SphericalDesign.cpp:12-21
DESIGN_SIZES = {6: 3, 12: 5, 32: 7, 48: 9, 70: 11, 94: 13, 120: 15, 156: 17, 192: 19}


def _legendre_rule(n):
    """Non-negative nodes (ascending) and weights of the n-point rule on [-1,1]."""
    ld = np.longdouble
    m = (n + 1) // 2
    xs = np.zeros(m, dtype=np.float64)
    ws = np.zeros(m, dtype=np.float64)
    for k in range(1, m + 1):
        z = ld(math.cos(math.pi * (k - 0.25) / (n + 0.5)))
        pp = ld(1)
        for _ in range(100):
            p0, p1 = ld(1), z
            for j in range(2, n + 1):
                p0, p1 = p1, ((2 * j - 1) * z * p1 - (j - 1) * p0) / j
            pp = n * (z * p1 - p0) / (z * z - 1)
            dz = p1 / pp
            z = z - dz
            if abs(dz) < ld(1e-19):
                break
        p0, p1 = ld(1), z
        for j in range(2, n + 1):
            p0, p1 = p1, ((2 * j - 1) * z * p1 - (j - 1) * p0) / j
        pp = n * (z * p1 - p0) / (z * z - 1)
        w = 2 / ((1 - z * z) * pp * pp)
        if n % 2 == 1 and k == m:
            z = ld(0)
        xs[m - k] = float(z)
        ws[m - k] = float(w)
    return xs, ws


class GaussLegendreQuadrature:
    """n-point Gauss-Legendre rule on [a, b], nodes ascending (GaussLegendre.hpp:10-24)."""

    def __init__(self, n_points, a, b):
        n = int(n_points)
        if n <= 0:
            raise ValueError("Number of points must be a positive integer")
        xs, ws = _legendre_rule(n)
        A = (b - a) / 2
        B = (a + b) / 2
        nodes = np.empty(n)
        weights = np.empty(n)
        for i in range(n):
            if n % 2 == 1:
                k = i - n // 2
                if k < 0:
                    nodes[i] = B - A * xs[-k]
                    weights[i] = A * ws[-k]
                else:
                    nodes[i] = B + A * xs[k]
                    weights[i] = A * ws[k]
            elif i < n // 2:
                k = n // 2 - 1 - i
                nodes[i] = B - A * xs[k]
                weights[i] = A * ws[k]
            else:
                k = i - n // 2
                nodes[i] = B + A * xs[k]
                weights[i] = A * ws[k]
        self.nodes = nodes
        self.weights = weights

    def getNodes(self):
        return self.nodes

    def getWeights(self):
        return self.weights

    def getNumberOfPoints(self):
        return len(self.weights)


class SphericalQuadrature:
    """Generic spherical rule (AbstractSphericalQuadratures.hpp:11-61)."""

    def __init__(self, x, y, z, weights):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)
        if not (len(self.x) == len(self.y) == len(self.z) == len(self.weights)):
            raise ValueError("x, y, z and weights must have the same length")

    def getx(self):
        return self.x

    def gety(self):
        return self.y

    def getz(self):
        return self.z

    def getWeights(self):
        return self.weights

    def getNumberOfPoints(self):
        return len(self.weights)

    def is_antipodal(self):
        """True if every node has a bit-exact antipode with equal weight (enables folding)."""
        pts = {(a, b, c): w for a, b, c, w in zip(self.x, self.y, self.z, self.weights)}
        return len(pts) == len(self.x) and all(
            pts.get((-a, -b, -c)) == w for (a, b, c), w in pts.items())


_design_cache = None


def _packaged_designs():
    global _design_cache
    if _design_cache is None:
        with open(_DESIGN_JSON) as fh:
            _design_cache = json.load(fh)
    return _design_cache


class SphericalDesign(SphericalQuadrature):
    """Spherical t-design with N points, equal weights 4*pi/N (SphericalDesign.cpp:6-50).

    `design_dir` (or $BFSM_DESIGN_DIR) may point at a directory holding the reference's
    ssTTT.NNN.txt files; otherwise the packaged table is used.
    """

    def __init__(self, N, design_dir=None):
        N = int(N)
        if N <= 0:
            raise ValueError("Number of points N must be a positive integer")
        if N not in DESIGN_SIZES:
            raise ValueError("Invalid value of N")
        design_dir = design_dir or os.environ.get("BFSM_DESIGN_DIR")
        if design_dir:
            fname = os.path.join(design_dir, f"ss{DESIGN_SIZES[N]:03d}.{N:03d}.txt")
            if not os.path.exists(fname):
                raise RuntimeError("Could not open file " + fname)
            rows = [[float(v) for v in line.split()] for line in open(fname) if line.strip()]
        else:
            rows = [[float.fromhex(v) for v in r] for r in _packaged_designs()[str(N)]["xyz_hex"]]
        if len(rows) != N:
            raise RuntimeError(f"design table for N={N} has {len(rows)} rows")
        xyz = np.asarray(rows, dtype=np.float64)
        super().__init__(xyz[:, 0], xyz[:, 1], xyz[:, 2], np.full(N, (4 * pi) / N))
        self.N = N
