"""In-tree build of the CUDA shared library (sm_100a only).

`nvcc` cross-compiles without a GPU, so this runs in the build container; the resulting
`csrc/libbfsm_b200.so` is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libbfsm_b200.so")
SOURCES = ["bfsm_capi.cu"]
HEADERS = ["bfsm_fft.cuh", "bfsm_kernels.cuh", "bfsm_pencil_reg.cuh", "bfsm_fused.cuh", "bfsm_plane_r32.cuh", "bfsm_tmem.cuh",
           "bfsm_cluster.cuh", "bfsm_general.cuh", "bfsm_aux.cuh",
           os.path.join("..", "..", "include", "bfsm_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libbfsm_b200.so. Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-ldl"]
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper; let nvcc pick the system g++
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB
