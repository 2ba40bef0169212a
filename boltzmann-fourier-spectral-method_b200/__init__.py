"""B200-native (sm_100a) fast Fourier-spectral Boltzmann collision operator Q(f,f).

A drop-in for ONE hot path of i3s93/Boltzmann-Fourier-Spectral-Method:
`BoltzmannOperator<Backend>::computeCollision` (Collisions/FFTWBoltzmannOperator.cpp:147-334).
The arithmetic lives in csrc/ (hand-written fp64 CUDA kernels behind the C ABI of
include/bfsm_b200.h); this package is the Python host-side mirror of the reference's
operator / quadrature interface.  Importable as `bfsm_b200` (see bfsm_b200.py at the repo root).
"""
from .quadratures import GaussLegendreQuadrature, SphericalDesign, SphericalQuadrature, pi
from .operator import BoltzmannOperatorB200
from . import inputs

__all__ = [
    "GaussLegendreQuadrature", "SphericalDesign", "SphericalQuadrature", "pi",
    "BoltzmannOperatorB200", "inputs",
]
