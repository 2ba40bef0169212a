"""oracle/oracle.py -- TEST INFRASTRUCTURE: Python access to the three parity checkers.

  numpy_collide / numpy_gain_hat   NumPy (pocketfft) restatement of
                                   FFTWBoltzmannOperator.cpp:147-334 -- an FFT implementation
                                   independent of oracle/shim/
  PortOracle                       ctypes binding of oracle/libbfsm_oracle.so (bfsm_oracle.c,
                                   the streaming plain-C restatement, "port")
  ReferenceOperator                ctypes binding of oracle/_ref/libbfsm_ref.so (the UNMODIFIED
                                   reference CPU operator compiled from /root/reference, "reference")

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
The product path (boltzmann-fourier-spectral-method_b200/) never does.
"""
import contextlib
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "libbfsm_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libbfsm_ref.so")
REF_DESIGNS = os.path.join(HERE, "_ref", "designs")

pi = 3.14159265358979323846  # Utilities/constants.hpp:7
_dp = ctypes.POINTER(ctypes.c_double)


def build(which=("port", "ref")):
    """Run oracle/Makefile (compiles the checkers; does not use them)."""
    subprocess.run(["make", "-C", HERE, "-s"] + list(which), check=True)


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


# ----------------------------------------------------------------------------- numpy
def sincc(x):
    """FFTWBoltzmannOperator.hpp:17-21"""
    eps = 2.220446049250313e-16
    return np.sin(x + eps) / (x + eps)


def mode_table(n):
    """FFTWBoltzmannOperator.cpp:50-57"""
    return np.concatenate([np.arange(0, n // 2), np.arange(-(n // 2), 0)])


def numpy_gain_hat(shape, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L, f, pairs=None):
    """Q_gain_hat restricted to `pairs` (iterable of b = r*N_s + s; default all) and f_hat.

    Steps 0-5 of the reference, cpp:168-276, one pair at a time."""
    nx, ny, nz = shape
    N = nx * ny * nz
    fft_scale = 1.0 / N
    lx, ly, lz = mode_table(nx), mode_table(ny), mode_table(nz)
    LX, LY, LZ = np.meshgrid(lx, ly, lz, indexing="ij")
    norm_l = np.sqrt((LX * LX + LY * LY + LZ * LZ).astype(np.float64))
    f = np.asarray(f, dtype=np.float64).reshape(shape)
    f_hat = np.fft.fftn(f)
    n_s = len(sx)
    Qg_hat = np.zeros(shape, dtype=np.complex128)
    if pairs is None:
        pairs = range(len(rho) * n_s)
    beta1_cache = {}
    for b in pairs:
        r, s = divmod(b, n_s)
        theta = -(pi / (2 * L)) * rho[r] * (LX * sx[s] + LY * sy[s] + LZ * sz[s])
        alpha = np.exp(1j * theta)
        a1 = np.fft.ifftn(fft_scale * alpha * f_hat) * N          # unnormalised backward
        a2 = np.fft.ifftn(fft_scale * np.conj(alpha) * f_hat) * N
        prod_hat = np.fft.fftn(a1 * a2)
        weight = fft_scale * w_r[r] * w_s[s] * rho[r] ** (gamma + 2)
        if r not in beta1_cache:
            beta1_cache[r] = 4 * pi * b_gamma * sincc(pi * rho[r] * norm_l / (2 * L))
        Qg_hat += weight * beta1_cache[r] * prod_hat
    return Qg_hat, f_hat


def numpy_finish(shape, rho, w_r, gamma, b_gamma, L, f, f_hat, Qg_hat):
    """Steps 6-7 of the reference, cpp:281-330."""
    nx, ny, nz = shape
    N = nx * ny * nz
    fft_scale = 1.0 / N
    lx, ly, lz = mode_table(nx), mode_table(ny), mode_table(nz)
    LX, LY, LZ = np.meshgrid(lx, ly, lz, indexing="ij")
    norm_l = np.sqrt((LX * LX + LY * LY + LZ * LZ).astype(np.float64))
    beta2 = np.zeros(shape)
    for r in range(len(rho)):
        beta2 += 16 * pi * pi * b_gamma * w_r[r] * rho[r] ** (gamma + 2) * sincc(pi * rho[r] * norm_l / L)
    Q_gain = np.fft.ifftn(Qg_hat) * N
    h = np.fft.ifftn(fft_scale * beta2 * f_hat) * N
    f = np.asarray(f, dtype=np.float64).reshape(shape)
    return (Q_gain.real - (h * f).real).copy()


def numpy_collide(shape, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L, f):
    Qg_hat, f_hat = numpy_gain_hat(shape, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L, f)
    return numpy_finish(shape, rho, w_r, gamma, b_gamma, L, f, f_hat, Qg_hat)


def hermitian_part(Qhat):
    """(Qhat(l) + conj(Qhat(-l)))/2, the spectrum of Re(IFFT(Qhat)): what the CUDA path's
    bfsm_gain_hat returns (it transforms only Re(g1*g2), see include/bfsm_b200.h)."""
    rev = Qhat
    for ax in range(Qhat.ndim):
        rev = np.roll(np.flip(rev, axis=ax), 1, axis=ax)
    return 0.5 * (Qhat + np.conj(rev))


# ----------------------------------------------------------------------------- port (C)
class PortOracle:
    def __init__(self):
        if not os.path.exists(PORT_LIB):
            build(("port",))
        self.lib = ctypes.CDLL(PORT_LIB)
        self.lib.bfsm_oracle_collide.restype = ctypes.c_int
        self.lib.bfsm_oracle_gain_hat.restype = ctypes.c_int
        self.lib.bfsm_oracle_finish.restype = ctypes.c_int
        self.lib.bfsm_oracle_fft3.restype = ctypes.c_int
        self.lib.bfsm_oracle_gauss_legendre.restype = ctypes.c_int
        self.lib.bfsm_oracle_max_threads.restype = ctypes.c_int

    def set_threads(self, n):
        self.lib.bfsm_oracle_set_threads(ctypes.c_int(int(n)))

    def max_threads(self):
        return int(self.lib.bfsm_oracle_max_threads())

    def gauss_legendre(self, n, a, b):
        x, w = np.empty(n), np.empty(n)
        rc = self.lib.bfsm_oracle_gauss_legendre(ctypes.c_int(n), ctypes.c_double(a),
                                                 ctypes.c_double(b), _p(x), _p(w))
        assert rc == 0
        return x, w

    def _quad_args(self, rho, w_r, sx, sy, sz, w_s):
        arrs = [_c(a) for a in (rho, w_r, sx, sy, sz, w_s)]
        return arrs

    def collide(self, shape, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L, f):
        rho, w_r, sx, sy, sz, w_s = self._quad_args(rho, w_r, sx, sy, sz, w_s)
        f = _c(f).ravel()
        Q = np.empty_like(f)
        rc = self.lib.bfsm_oracle_collide(
            ctypes.c_int(shape[0]), ctypes.c_int(shape[1]), ctypes.c_int(shape[2]),
            ctypes.c_int(len(rho)), _p(rho), _p(w_r), ctypes.c_int(len(sx)), _p(sx), _p(sy), _p(sz),
            _p(w_s), ctypes.c_double(gamma), ctypes.c_double(b_gamma), ctypes.c_double(L), _p(f), _p(Q))
        assert rc == 0
        return Q.reshape(shape)

    def gain_hat(self, shape, rho, w_r, sx, sy, sz, w_s, gamma, b_gamma, L, f, pair_begin, pair_end):
        rho, w_r, sx, sy, sz, w_s = self._quad_args(rho, w_r, sx, sy, sz, w_s)
        f = _c(f).ravel()
        out = np.empty(2 * f.size)
        rc = self.lib.bfsm_oracle_gain_hat(
            ctypes.c_int(shape[0]), ctypes.c_int(shape[1]), ctypes.c_int(shape[2]),
            ctypes.c_int(len(rho)), _p(rho), _p(w_r), ctypes.c_int(len(sx)), _p(sx), _p(sy), _p(sz),
            _p(w_s), ctypes.c_double(gamma), ctypes.c_double(b_gamma), ctypes.c_double(L), _p(f),
            ctypes.c_int(pair_begin), ctypes.c_int(pair_end), _p(out))
        assert rc == 0
        return out.view(np.complex128).reshape(shape)

    def finish(self, shape, rho, w_r, gamma, b_gamma, L, f, Qhat):
        rho, w_r = _c(rho), _c(w_r)
        f = _c(f).ravel()
        qh = np.ascontiguousarray(np.asarray(Qhat, dtype=np.complex128)).ravel().view(np.float64)
        Q = np.empty_like(f)
        rc = self.lib.bfsm_oracle_finish(
            ctypes.c_int(shape[0]), ctypes.c_int(shape[1]), ctypes.c_int(shape[2]),
            ctypes.c_int(len(rho)), _p(rho), _p(w_r), ctypes.c_double(gamma),
            ctypes.c_double(b_gamma), ctypes.c_double(L), _p(f), _p(qh), _p(Q))
        assert rc == 0
        return Q.reshape(shape)

    def fft3(self, a, sign):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.complex128))
        out = np.empty_like(a)
        rc = self.lib.bfsm_oracle_fft3(
            ctypes.c_int(a.shape[0]), ctypes.c_int(a.shape[1]), ctypes.c_int(a.shape[2]),
            ctypes.c_int(sign), _p(a.view(np.float64)), _p(out.view(np.float64)))
        assert rc == 0
        return out


# ----------------------------------------------------------------------------- reference
@contextlib.contextmanager
def _quiet_stdout():
    """The reference's initialize() prints "Failed to import wisdom ..." on std::cout
    (FFTWBoltzmannOperator.cpp:60-62); keep it off our stdout (bench.py prints one JSON line)."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def reference_available():
    return os.path.exists(REF_LIB)


class ReferenceOperator:
    """The reference's BoltzmannOperator<FFTW_Backend> (unmodified sources) via ref_capi.cpp."""

    def __init__(self, Nv, n_gl, n_sph, gamma, b_gamma, L, a=0.0, b=10.0, threads=None):
        if not reference_available():
            raise FileNotFoundError(REF_LIB)
        os.environ.setdefault("BFSM_REF_DESIGN_DIR", REF_DESIGNS)
        lib = ctypes.CDLL(REF_LIB)
        lib.bfsm_ref_create.restype = ctypes.c_void_p
        lib.bfsm_ref_last_error.restype = ctypes.c_char_p
        lib.bfsm_ref_apply_timed.restype = ctypes.c_double
        lib.bfsm_ref_max_threads.restype = ctypes.c_int
        self.lib = lib
        if threads:
            lib.bfsm_ref_set_threads(ctypes.c_int(int(threads)))
        # Nv: one size, or (Nvx, Nvy, Nvz) -- the reference class takes three (FFTWBoltzmannOperator.hpp:30-36)
        self.shape = (Nv, Nv, Nv) if np.isscalar(Nv) else tuple(int(n) for n in Nv)
        self.n_gl, self.n_sph = n_gl, n_sph
        with _quiet_stdout():
            self.h = lib.bfsm_ref_create(
                ctypes.c_int(self.shape[0]), ctypes.c_int(self.shape[1]), ctypes.c_int(self.shape[2]),
                ctypes.c_int(n_gl),
                ctypes.c_double(a), ctypes.c_double(b), ctypes.c_int(n_sph), ctypes.c_double(gamma),
                ctypes.c_double(b_gamma), ctypes.c_double(L))
        if not self.h:
            raise RuntimeError("reference operator: " + lib.bfsm_ref_last_error().decode())

    def max_threads(self):
        return int(self.lib.bfsm_ref_max_threads())

    def quadrature(self):
        g = [np.empty(self.n_gl) for _ in range(2)]
        s = [np.empty(self.n_sph) for _ in range(4)]
        self.lib.bfsm_ref_quadrature(ctypes.c_void_p(self.h), *[_p(a) for a in g + s])
        return g[0], g[1], s[0], s[1], s[2], s[3]

    def __call__(self, f, timed=False):
        f = _c(f).ravel()
        Q = np.empty_like(f)
        if timed:
            t = self.lib.bfsm_ref_apply_timed(ctypes.c_void_p(self.h), _p(f), _p(Q))
            return Q.reshape(self.shape), float(t)
        self.lib.bfsm_ref_apply(ctypes.c_void_p(self.h), _p(f), _p(Q))
        return Q.reshape(self.shape)

    def close(self):
        if self.h:
            self.lib.bfsm_ref_destroy(ctypes.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
