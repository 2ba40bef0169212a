"""Import alias: the package directory is named after the reference repository
(`boltzmann-fourier-spectral-method_b200`, not a valid Python identifier), so
`import bfsm_b200` loads it through importlib and re-exports its contents."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_pkg = importlib.import_module("boltzmann-fourier-spectral-method_b200")
sys.modules.setdefault("bfsm_b200_pkg", _pkg)

GaussLegendreQuadrature = _pkg.GaussLegendreQuadrature
SphericalDesign = _pkg.SphericalDesign
SphericalQuadrature = _pkg.SphericalQuadrature
BoltzmannOperatorB200 = _pkg.BoltzmannOperatorB200
inputs = _pkg.inputs
pi = _pkg.pi
package = _pkg


def submodule(name):
    """e.g. submodule('distributed'), submodule('_capi'), submodule('build')"""
    return importlib.import_module("boltzmann-fourier-spectral-method_b200." + name)
