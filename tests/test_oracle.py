"""CPU tests that PIN the oracle (no GPU): the C restatement (oracle/bfsm_oracle.c) against
 (1) the golden vectors produced by the unmodified reference operator (tests/golden/),
 (2) the known answers published in the reference's Results/ files,
 (3) an independent NumPy (pocketfft) restatement,
 (4) when present, oracle/_ref/libbfsm_ref.so itself."""
import json
import os

import numpy as np
import pytest

from helpers import inp, make_input, oracle_args, quadrature, rel_linf
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VECTORS = np.load(os.path.join(GOLDEN, "reference_q_vectors.npz"))
with open(os.path.join(GOLDEN, "bkw_known_answers.json")) as fh:
    KNOWN = json.load(fh)

# oracle noise floor measured in the survey: <= 1.4e-14 between builds; allow 1e-13
ORACLE_TOL = 1e-13


def _cases():
    for key in VECTORS.files:
        if key.endswith("_Q"):
            nv, r, s, kind = key[:-2].split("_")
            yield int(nv[2:]), int(r[1:]), int(s[1:]), kind


@pytest.mark.parametrize("Nv,n_r,n_s,kind", list(_cases()))
def test_port_matches_reference_golden_vectors(port_oracle, Nv, n_r, n_s, kind):
    gl, sd = quadrature(n_r, n_s)
    # the quadrature the reference built (GSL stand-in) vs our GSL-free host class
    assert np.abs(VECTORS[f"Nv{Nv}_r{n_r}_s{n_s}_rho"] - gl.getNodes()).max() < 4e-15
    assert np.abs(VECTORS[f"Nv{Nv}_r{n_r}_s{n_s}_wr"] - gl.getWeights()).max() < 4e-15
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), make_input(kind, Nv))
    assert rel_linf(Q, VECTORS[f"Nv{Nv}_r{n_r}_s{n_s}_{kind}_Q"]) < ORACLE_TOL


@pytest.mark.parametrize("kind", ["maxmix", "noise"])
def test_port_matches_reference_with_the_full_192_point_design(port_oracle, kind):
    """One more link of the chain reference -> port -> GPU at the flagship size: the C port against the
    UNMODIFIED reference operator at 64^3 with ALL 192 directions of ss019.192 and one radius (the
    reference's six batch arrays need 4.8 GB there; the full 32-radius list would need 154.6 GB, which
    is why the flagship golden itself, port_q_cfg4.npz, has to come from the port)."""
    Nv, n_r, n_s = 64, 1, 192
    G = np.load(os.path.join(GOLDEN, "reference_q_64cubed.npz"))
    gl, sd = quadrature(n_r, n_s)
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), make_input(kind, Nv)).reshape(Nv, Nv, Nv)
    key, st = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}", int(G["stride"])
    qmax = float(G[key + "_max"])
    assert np.abs(Q[::st, ::st, ::st] - G[key + "_Qsub"]).max() / qmax < ORACLE_TOL
    assert np.abs(Q.sum(axis=(1, 2)) - G[key + "_plane_sum"]).max() / (qmax * Nv * Nv) < ORACLE_TOL
    assert np.abs((Q * Q).sum(axis=(1, 2)) - G[key + "_plane_sumsq"]).max() / (qmax ** 2 * Nv * Nv) < ORACLE_TOL


@pytest.mark.parametrize("kind", ["maxmix", "noise"])
def test_port_matches_reference_at_64_cubed(port_oracle, kind):
    """64^3 is where the pipelined plane kernel runs; the GPU parity tests there check against the C
    port, so the port itself is pinned at that size against the UNMODIFIED reference operator:
    every second point per axis elementwise, per-x-plane sums of Q and Q^2 for the rest
    (tests/golden/reference_q_64cubed.npz, written by make_golden.py --only-64)."""
    Nv, n_r, n_s = 64, 2, 12
    G = np.load(os.path.join(GOLDEN, "reference_q_64cubed.npz"))
    gl, sd = quadrature(n_r, n_s)
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), make_input(kind, Nv)).reshape(Nv, Nv, Nv)
    key, st = f"Nv{Nv}_r{n_r}_s{n_s}_{kind}", int(G["stride"])
    qmax = float(G[key + "_max"])
    assert np.abs(Q[::st, ::st, ::st] - G[key + "_Qsub"]).max() / qmax < ORACLE_TOL
    assert np.abs(Q.sum(axis=(1, 2)) - G[key + "_plane_sum"]).max() / (qmax * Nv * Nv) < ORACLE_TOL
    assert np.abs((Q * Q).sum(axis=(1, 2)) - G[key + "_plane_sumsq"]).max() / (qmax ** 2 * Nv * Nv) < ORACLE_TOL
    assert abs(np.abs(Q).max() - qmax) / qmax < ORACLE_TOL


@pytest.mark.parametrize("shape", [(32, 64, 16), (16, 24, 12), (8, 16, 32)])
@pytest.mark.parametrize("kind", ["maxmix", "noise"])
def test_port_matches_live_reference_on_non_cubic_grids(port_oracle, shape, kind):
    """The reference interface carries independent Nvx, Nvy, Nvz with per-axis mode tables
    (FFTWBoltzmannOperator.hpp:30-36, .cpp:46-57) although its drivers only run cubes.  The C port is
    pinned on non-cubic grids (one with axes that are not powers of two) against the UNMODIFIED
    reference operator run live (build container only) and against the NumPy restatement; the GPU
    general-grid path is then tested against the port."""
    n_r, n_s = 3, 12
    gl, sd = quadrature(n_r, n_s)
    f = make_input(kind, shape)
    Qp = port_oracle.collide(shape, *oracle_args(gl, sd), f)
    Qn = O.numpy_collide(shape, *oracle_args(gl, sd), f)
    assert rel_linf(Qp, Qn) < ORACLE_TOL
    if O.reference_available():
        ref = O.ReferenceOperator(shape, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                                  a=0.0, b=inp.R_SUPPORT, threads=1)
        Qr = ref(f)
        ref.close()
        assert rel_linf(Qp, Qr) < ORACLE_TOL


@pytest.mark.parametrize("kind", ["bkw", "maxmix", "noise"])
def test_port_matches_numpy_restatement(port_oracle, kind):
    Nv, n_r, n_s = 16, 4, 12
    gl, sd = quadrature(n_r, n_s)
    f = make_input(kind, Nv)
    Qp = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    Qn = O.numpy_collide((Nv,) * 3, *oracle_args(gl, sd), f)
    assert rel_linf(Qp, Qn) < ORACLE_TOL


def test_fft_standin_matches_numpy(port_oracle):
    rng = np.random.default_rng(7)
    for n in (8, 16, 32):
        a = rng.standard_normal((n, n, n)) + 1j * rng.standard_normal((n, n, n))
        ref = np.fft.fftn(a)
        assert np.abs(port_oracle.fft3(a, -1) - ref).max() / np.abs(ref).max() < 1e-14
        back = np.fft.ifftn(a) * a.size
        assert np.abs(port_oracle.fft3(a, +1) - back).max() / np.abs(back).max() < 1e-14
    # non power-of-two fallback
    a = rng.standard_normal((6, 5, 12)) + 0j
    assert np.abs(port_oracle.fft3(a, -1) - np.fft.fftn(a)).max() < 1e-11


@pytest.mark.parametrize("entry", [e for e in KNOWN["published"] if e["Nv"] == 32],
                         ids=lambda e: f"Nv{e['Nv']}_Ns{e['N_sigma']}")
def test_port_reproduces_published_bkw_errors(port_oracle, entry):
    """Results/maxwell_bkw_fftw_atomics.txt:19-21 and :371-373 (printed to 9 digits)."""
    Nv, n_r, n_s = entry["Nv"], entry["N_r"], entry["N_sigma"]
    gl, sd = quadrature(n_r, n_s)
    f, Q_exact = inp.bkw(Nv)
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    l1, l2, linf = inp.error_norms(Q, Q_exact, Nv)
    assert abs(l1 - entry["L1"]) <= 1e-8 * entry["L1"]
    assert abs(l2 - entry["L2"]) <= 1e-8 * entry["L2"]
    assert abs(linf - entry["Linf"]) <= 1e-8 * entry["Linf"]


@pytest.mark.parametrize("entry", [e for e in KNOWN["published"] if e["Nv"] == 64],
                         ids=lambda e: f"Nv{e['Nv']}_Ns{e['N_sigma']}")
def test_port_reproduces_published_bkw_errors_at_64_cubed(port_oracle, entry):
    """Results/maxwell_bkw_fftw_atomics.txt:195-197 and :547-549.  At 64^3 the error against the exact
    BKW derivative is ~1e-10, i.e. rounding noise of the sum over 768 / 2048 pairs is visible from the
    4th digit on (the reference's own runs differ there), so the bar is north_star's
    "within 1 %"; observed here: 2e-4 (L1), 1e-6 (L2), 1e-5 (Linf)."""
    Nv, n_r, n_s = entry["Nv"], entry["N_r"], entry["N_sigma"]
    gl, sd = quadrature(n_r, n_s)
    f, Q_exact = inp.bkw(Nv)
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    l1, l2, linf = inp.error_norms(Q, Q_exact, Nv)
    assert abs(l1 - entry["L1"]) <= 1e-2 * entry["L1"]
    assert abs(l2 - entry["L2"]) <= 1e-2 * entry["L2"]
    assert abs(linf - entry["Linf"]) <= 1e-2 * entry["Linf"]


@pytest.mark.parametrize("entry", [e for e in KNOWN["recomputed"] if e["Nv"] == 16],
                         ids=lambda e: f"Nv{e['Nv']}_Nr{e['N_r']}_Ns{e['N_sigma']}")
def test_port_reproduces_recomputed_bkw_errors(port_oracle, entry):
    Nv, n_r, n_s = entry["Nv"], entry["N_r"], entry["N_sigma"]
    gl, sd = quadrature(n_r, n_s)
    f, Q_exact = inp.bkw(Nv)
    Q = port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f)
    l1, l2, linf = inp.error_norms(Q, Q_exact, Nv)
    for got, key in ((l1, "L1"), (l2, "L2"), (linf, "Linf")):
        assert abs(got - entry[key]) <= 1e-9 * entry[key]


def test_shard_sum_equals_full_gain_spectrum(port_oracle):
    """Q_gain_hat is a plain sum over pairs: any partition of the pair list sums to the full
    spectrum, and finishing the summed spectrum gives the full Q (exactness of the multi-GPU
    split)."""
    Nv, n_r, n_s = 16, 4, 6
    gl, sd = quadrature(n_r, n_s)
    args = oracle_args(gl, sd)
    f = make_input("noise", Nv)
    P = n_r * n_s
    full = port_oracle.gain_hat((Nv,) * 3, *args, f, 0, P)
    cuts = [0, 5, 11, 17, P]
    parts = sum(port_oracle.gain_hat((Nv,) * 3, *args, f, a, b) for a, b in zip(cuts, cuts[1:]))
    assert np.abs(parts - full).max() / np.abs(full).max() < 1e-14
    Q_full = port_oracle.collide((Nv,) * 3, *args, f)
    Q_sum = port_oracle.finish((Nv,) * 3, args[0], args[1], args[6], args[7], args[8], f, parts)
    assert rel_linf(Q_sum, Q_full) < ORACLE_TOL


def test_hermitian_part_of_gain_spectrum_gives_same_Q(port_oracle):
    """The CUDA path transforms only Re(g1*g2), i.e. returns the Hermitian part of the reference's
    Q_gain_hat; that must not change Q (beta1 real and even, only Re(Q_gain) is kept)."""
    Nv, n_r, n_s = 16, 4, 6
    gl, sd = quadrature(n_r, n_s)
    args = oracle_args(gl, sd)
    f = make_input("noise", Nv)
    full = port_oracle.gain_hat((Nv,) * 3, *args, f, 0, n_r * n_s)
    herm = O.hermitian_part(full)
    assert np.abs(herm - full).max() / np.abs(full).max() > 1e-6  # noise input: really different
    Q_a = port_oracle.finish((Nv,) * 3, args[0], args[1], args[6], args[7], args[8], f, full)
    Q_b = port_oracle.finish((Nv,) * 3, args[0], args[1], args[6], args[7], args[8], f, herm)
    assert rel_linf(Q_b, Q_a) < ORACLE_TOL


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref not built")
def test_port_matches_live_reference_build(port_oracle):
    Nv, n_r, n_s = 16, 8, 12
    ref = O.ReferenceOperator(Nv, n_r, n_s, inp.GAMMA_MAXWELL, inp.B_GAMMA_MAXWELL, inp.L_DOMAIN,
                              a=0.0, b=inp.R_SUPPORT, threads=2)
    gl, sd = quadrature(n_r, n_s)
    f = make_input("noise", Nv)
    assert rel_linf(port_oracle.collide((Nv,) * 3, *oracle_args(gl, sd), f), ref(f)) < ORACLE_TOL
    rho, w_r, sx, sy, sz, sw = ref.quadrature()
    assert np.array_equal(sx, sd.getx()) and np.array_equal(sy, sd.gety()) and np.array_equal(sz, sd.getz())
    assert np.abs(sw - sd.getWeights()).max() == 0.0
    ref.close()
