"""CPU tests of the host-side logic and of the C-ABI boundary (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import bfsm_b200 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = B.submodule("_capi")
D = B.submodule("distributed")


def test_library_loads_and_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "bfsm_b200.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(bfsm_[a-z_0-9]+)\s*\(", header))
    declared -= {"bfsm_plan", "bfsm_plan_info"}
    assert declared, "no declarations parsed"
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in bfsm_b200.h but not exported"
    assert set(capi.EXPORTS) == declared
    assert lib.bfsm_version() == 100


def test_integration_notes_name_every_entry_point():
    with open(os.path.join(ROOT, "include", "bfsm_b200.h")) as fh:
        declared = set(re.findall(r"\b(bfsm_[a-z_0-9]+)\s*\(", fh.read())) - {"bfsm_plan", "bfsm_plan_info"}
    with open(os.path.join(ROOT, "INTEGRATION.md")) as fh:
        notes = fh.read()
    missing = [d for d in sorted(declared) if d not in notes]
    assert not missing, missing


def test_constants_mirror_the_header():
    with open(os.path.join(ROOT, "include", "bfsm_b200.h")) as fh:
        header = fh.read()
    for name in ("BFSM_HOST_PIPE_DEPTH", "BFSM_UNIQUE_ID_BYTES"):
        value = int(re.search(r"#define\s+%s\s+(\d+)" % name, header).group(1))
        assert getattr(capi, name) == value, name


def test_plan_info_mirror_matches_the_header_field_by_field():
    """The ctypes mirror of bfsm_plan_info must list the header's fields in the same order with the
    same C types (the library fills the struct through a plain pointer)."""
    with open(os.path.join(ROOT, "include", "bfsm_b200.h")) as fh:
        header = fh.read()
    body = re.search(r"typedef struct \{([^{}]*)\} bfsm_plan_info;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = []
    for ctype, names in re.findall(r"\b(int|long long)\s+([a-z_0-9, ]+);", body):
        for name in names.split(","):
            declared.append((name.strip(), ctype))
    ctype_of = {ctypes.c_int: "int", ctypes.c_longlong: "long long"}
    mirrored = [(name, ctype_of[t]) for name, t in capi.PlanInfo._fields_]
    assert mirrored == declared


def test_plan_options_mirror_matches_the_header_field_by_field():
    """Same for bfsm_plan_options (the Python harness fills it, the library reads it)."""
    with open(os.path.join(ROOT, "include", "bfsm_b200.h")) as fh:
        header = fh.read()
    body = re.search(r"typedef struct \{([^{}]*)\} bfsm_plan_options;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = []
    for name, dim in re.findall(r"\bint\s+([a-z_0-9]+)(?:\[(\d+)\])?;", body):
        declared.append((name, int(dim) if dim else 1))
    mirrored = [(name, 1 if t is ctypes.c_int else t._length_) for name, t in capi.PlanOptions._fields_]
    assert mirrored == declared
    opts = capi.PlanOptions()
    capi.load().bfsm_plan_options_init(ctypes.byref(opts))
    assert opts.struct_size == ctypes.sizeof(capi.PlanOptions) and opts.side_stream == 1


def test_library_reads_no_environment_variables():
    """Tuning goes through bfsm_plan_options; the library sources must not call getenv (a stray
    variable used to change the product path silently).  The statically linked CUDA runtime does
    reference getenv, so the check is on our sources, not on the symbol table."""
    csrc = os.path.join(ROOT, "boltzmann-fourier-spectral-method_b200", "csrc")
    for name in os.listdir(csrc):
        if name.endswith((".cu", ".cuh", ".hpp", ".h")):
            with open(os.path.join(csrc, name)) as fh:
                assert "getenv" not in fh.read(), name


@pytest.mark.parametrize("pairs_local,pair_lo,n_dir,chunk,seg", [
    (3072, 0, 96, 384, 24), (384, 768, 96, 384, 24), (752, 0, 47, 1024, 12), (3072, 0, 96, 7, 24),
    (1024, 512, 96, 384, 5), (24, 0, 3, 24, 1), (48, 0, 6, 1024, 24), (100, 37, 16, 33, 6), (0, 0, 8, 16, 4)])
def test_work_units_of_the_register_x_stage_tile_the_pair_list(pairs_local, pair_lo, n_dir, chunk, seg):
    """bfsm_debug_units runs the host code that cuts a shard's pair list into work units: the units
    tile [0, pairs_local) in order, none crosses a launch (chunk) or a radius boundary or exceeds
    seg_pairs, and within a radius the slots are 0, 1, 2, ... (each (slot, radius) written once)."""
    lib = capi.load()
    cap = max(1, pairs_local)
    out = (ctypes.c_int * (4 * cap))()
    n = lib.bfsm_debug_units(pairs_local, pair_lo, n_dir, chunk, seg, out, cap)
    assert 0 <= n <= cap
    units = [tuple(out[4 * k:4 * k + 4]) for k in range(n)]
    pos, seen = 0, {}
    r_first = pair_lo // n_dir
    for p0, p1, r, slot in units:
        assert p0 == pos and p0 < p1 <= p0 + seg
        assert p0 // chunk == (p1 - 1) // chunk
        assert (pair_lo + p0) // n_dir - r_first == r == (pair_lo + p1 - 1) // n_dir - r_first
        assert slot == seen.get(r, 0)
        seen[r] = slot + 1
        pos = p1
    assert pos == pairs_local
    assert lib.bfsm_debug_units(10, 0, 0, 4, 2, None, 0) == -capi.BFSM_ERR_INVALID


@pytest.mark.parametrize("n,n_items,n_ctas", [(64, 384, 148), (64, 12, 148), (64, 1, 67), (64, 7, 148),
                                             (32, 752, 296), (16, 24, 592), (64, 383, 132)])
def test_plane_kernel_work_split_covers_every_entry_once_and_is_balanced(n, n_items, n_ctas):
    """The gain plane kernels split the flat (plane, item) list so that every CTA gets an equal share of
    the n regular planes AND of the 3 costlier Nyquist planes (one contiguous range per CTA left the
    owners of the Nyquist planes 10 % behind).  bfsm_debug_plane_work runs the kernels' range arithmetic
    and item walker on the host."""
    lib = capi.load()
    seen = np.zeros((n + 3, n_items), dtype=np.int32)
    per_cta = []
    for cta in range(n_ctas):
        cap = (n + 3) * n_items
        planes = (ctypes.c_int * cap)()
        items = (ctypes.c_int * cap)()
        cnt = lib.bfsm_debug_plane_work(n, n_items, n_ctas, cta, planes, items, cap)
        assert 0 <= cnt <= cap
        pl, it = np.frombuffer(planes, dtype=np.int32)[:cnt], np.frombuffer(items, dtype=np.int32)[:cnt]
        assert ((0 <= pl) & (pl < n + 3) & (0 <= it) & (it < n_items)).all()
        np.add.at(seen, (pl, it), 1)
        per_cta.append((int((pl < n).sum()), int((pl >= n).sum())))
        # regular entries come first, each class is walked in flat-list order
        flat = pl.astype(np.int64) * n_items + it
        assert (np.diff(flat) > 0).all()
    assert (seen == 1).all()
    reg = [a for a, _ in per_cta]
    nyq = [b for _, b in per_cta]
    assert max(reg) - min(reg) <= 1 and max(nyq) - min(nyq) <= 1
    assert lib.bfsm_debug_plane_work(n, 0, n_ctas, 0, None, None, 0) == -capi.BFSM_ERR_INVALID


@pytest.mark.parametrize("n,n_items,n_groups", [(32, 752, 1184), (64, 384, 296), (64, 384, 444), (32, 1, 1184),
                                                (64, 5, 1), (32, 3, 2), (64, 192, 7), (32, 50, 1000)])
def test_radix32_plane_kernel_work_split(n, n_items, n_groups):
    """k_plane_gain_r32: every (plane, item) entry exactly once; one contiguous range per group; the Nyquist
    planes are served by their own groups (no group mixes the two classes unless it is alone); shares inside
    a class differ by at most one entry and the two classes are balanced by cost (a Nyquist entry ~1.3 x)."""
    lib = capi.load()
    lib.bfsm_debug_plane_work_r32.restype = ctypes.c_int
    seen = np.zeros((n + 3, n_items), dtype=np.int32)
    reg, nyq = [], []
    cap = (n + 3) * n_items
    for grp in range(n_groups):
        planes = (ctypes.c_int * cap)()
        items = (ctypes.c_int * cap)()
        cnt = lib.bfsm_debug_plane_work_r32(n, n_items, n_groups, grp, planes, items, cap)
        assert 0 <= cnt <= cap
        pl, it = np.frombuffer(planes, dtype=np.int32)[:cnt], np.frombuffer(items, dtype=np.int32)[:cnt]
        np.add.at(seen, (pl, it), 1)
        flat = pl.astype(np.int64) * n_items + it
        assert (np.diff(flat) == 1).all()
        if n_groups > 1:
            assert (pl < n).all() or (pl >= n).all()
        if cnt and (pl < n).all():
            reg.append(cnt)
        elif cnt:
            nyq.append(cnt)
    assert (seen == 1).all()
    if n_groups > 1:
        assert nyq, "somebody must own the Nyquist planes"
        if reg and len(reg) + len(nyq) == n_groups:  # no idle groups: both classes evenly cut
            assert max(reg) - min(reg) <= 1 and max(nyq) - min(nyq) <= 1
        if n * n_items >= 20 * n_groups and n_groups >= 64:
            assert 0.7 <= (1.3 * max(nyq)) / max(reg) <= 1.4
    assert lib.bfsm_debug_plane_work_r32(16, 4, 4, 0, None, None, 0) == -capi.BFSM_ERR_UNSUPPORTED


def test_single_slot_condition_matches_a_brute_force_check():
    """BFSM_ALIGNED_SLOTS=1 lets all CTA rows of the pencil / Nyquist kernels share one partial-sum
    slot when every row's share of every launch starts at a radius boundary.  The host predicate is
    checked against a direct enumeration: no radius may be touched by two rows of the same launch."""
    lib = capi.load()

    def brute(pairs_local, pair_lo, n_dir, chunk, groups):
        for c0 in range(0, pairs_local, chunk):
            nc = min(chunk, pairs_local - c0)
            G = min(groups, nc)
            owner = {}
            for g in range(G):
                for q in range(nc * g // G, nc * (g + 1) // G):
                    r = (pair_lo + c0 + q) // n_dir
                    if owner.setdefault(r, g) != g:
                        return False
        return True

    cases = [(3072, 0, 96, 384, 4), (384, 768, 96, 384, 4), (752, 0, 47, 1024, 8), (752, 0, 47, 1024, 4),
             (256, 0, 16, 1024, 8), (768, 0, 24, 1024, 8), (3072, 0, 96, 576, 4), (3072, 0, 96, 7, 4),
             (1024, 512, 96, 384, 4), (24, 0, 3, 24, 16), (48, 0, 6, 1024, 8)]
    rng = np.random.default_rng(7)
    for _ in range(200):
        n_dir = int(rng.integers(1, 50))
        cases.append((int(rng.integers(1, 400)), int(rng.integers(0, 300)), n_dir,
                      int(rng.integers(1, 200)), int(rng.integers(1, 9))))
    n_true = 0
    for c in cases:
        got = lib.bfsm_debug_shares_aligned(*c)
        # the predicate may be conservative (alignment is sufficient, not necessary), never optimistic
        if got:
            assert brute(*c), c
            n_true += 1
    assert lib.bfsm_debug_shares_aligned(3072, 0, 96, 384, 4) == 1      # cfg 4, one GPU
    assert lib.bfsm_debug_shares_aligned(384, 1152, 96, 384, 4) == 1    # cfg 4, shard 3 of 8
    assert lib.bfsm_debug_shares_aligned(752, 0, 47, 1024, 8) == 1      # cfg 5 cell
    assert lib.bfsm_debug_shares_aligned(3072, 0, 96, 576, 4) == 0
    assert n_true >= 8


def _create(lib, nv=(16, 16, 16), n_r=2, n_s=6, shard=(0, 1), L=1.0, null_rho=False):
    dp = ctypes.POINTER(ctypes.c_double)
    rho = np.array([1.0, 2.0][:n_r] + [1.0] * max(0, n_r - 2))
    w = np.ones(max(n_r, 1))
    sd = B.SphericalDesign(6)
    plan = ctypes.c_void_p()
    rc = lib.bfsm_plan_create(
        ctypes.byref(plan), nv[0], nv[1], nv[2], n_r,
        None if null_rho else rho.ctypes.data_as(dp), w.ctypes.data_as(dp), n_s,
        sd.getx().ctypes.data_as(dp), sd.gety().ctypes.data_as(dp), sd.getz().ctypes.data_as(dp),
        sd.getWeights().ctypes.data_as(dp), 0.0, 1.0, L, 0, shard[0], shard[1], 0)
    return rc, plan


def test_plan_create_argument_errors_are_reported_not_fatal():
    lib = capi.load()
    rc, _ = _create(lib, null_rho=True)
    assert rc == capi.BFSM_ERR_INVALID and b"NULL" in lib.bfsm_last_error()
    rc, _ = _create(lib, n_r=0)
    assert rc == capi.BFSM_ERR_INVALID
    rc, _ = _create(lib, L=-1.0)
    assert rc == capi.BFSM_ERR_INVALID
    rc, _ = _create(lib, shard=(3, 2))
    assert rc == capi.BFSM_ERR_INVALID
    rc, _ = _create(lib, nv=(16, 16, 33))        # odd size: the mode tables assume even sizes
    assert rc == capi.BFSM_ERR_UNSUPPORTED and b"not supported" in lib.bfsm_last_error()
    rc, _ = _create(lib, nv=(256, 16, 16))       # longer than the general path's 128
    assert rc == capi.BFSM_ERR_UNSUPPORTED
    assert lib.bfsm_plan_destroy(None) == capi.BFSM_OK


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load()
    rc, _ = _create(lib)
    assert rc == capi.BFSM_ERR_CUDA
    assert b"no CPU fallback" in lib.bfsm_last_error()
    gl = B.GaussLegendreQuadrature(4, 0.0, 10.0)
    op = B.BoltzmannOperatorB200(gl, B.SphericalDesign(6), 16, 16, 16, 0.0, 1.0, 1.0)
    with pytest.raises(RuntimeError):
        op.initialize()
    with pytest.raises(RuntimeError):
        op.computeCollision(np.zeros(4096), np.zeros(4096))


def test_cpp_driver_fails_loudly_without_a_device():
    """The C++ operator class (include/B200BoltzmannOperator.hpp) follows the reference's CUDA
    backend on errors: message on stderr and a non-zero exit status -- never a silent CPU result."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    exe = os.path.join(ROOT, "boltzmann-fourier-spectral-method_b200", "drivers", "build", "maxwell_bkw_b200")
    designs = os.path.join(ROOT, "oracle", "_ref", "designs")
    if not (os.path.exists(exe) and os.path.isdir(designs)):
        pytest.skip("driver or design files not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([exe, "--Nv", "16", "--Ns", "6", "-t", "1", "--design-dir", designs],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
    assert "no CUDA device available (there is no CPU fallback)" in (out.stdout + out.stderr)


def test_gauss_legendre_against_numpy_and_exactness():
    for n in (1, 2, 5, 8, 16, 32, 64):
        gl = B.GaussLegendreQuadrature(n, 0.0, 10.0)
        x, w = np.polynomial.legendre.leggauss(n)
        assert np.abs(gl.getNodes() - (5 + 5 * x)).max() < 5e-14
        assert np.abs(gl.getWeights() - 5 * w).max() < 5e-14
        assert np.all(np.diff(gl.getNodes()) > 0)            # ascending, GaussLegendre.hpp:19-21
        assert gl.getNumberOfPoints() == n
        # exact for polynomials up to degree 2n-1
        k = 2 * n - 1
        assert abs((gl.getWeights() * gl.getNodes() ** k).sum() - 10.0 ** (k + 1) / (k + 1)) \
            <= 1e-13 * 10.0 ** (k + 1)
    with pytest.raises(ValueError):
        B.GaussLegendreQuadrature(0, 0.0, 1.0)


def test_spherical_designs():
    for n, degree in {6: 3, 12: 5, 32: 7, 48: 9, 70: 11, 94: 13, 120: 15, 156: 17, 192: 19}.items():
        sd = B.SphericalDesign(n)
        assert sd.getNumberOfPoints() == n
        r = np.sqrt(sd.getx() ** 2 + sd.gety() ** 2 + sd.getz() ** 2)
        assert np.abs(r - 1).max() < 1e-14
        assert np.all(sd.getWeights() == (4 * B.pi) / n)      # SphericalDesign.cpp:48
        assert sd.is_antipodal()
        # a t-design integrates odd monomials to zero and x^2 to 4 pi / 3
        assert abs((sd.getWeights() * sd.getx() ** 3).sum()) < 1e-13
        assert abs((sd.getWeights() * sd.getx() ** 2).sum() - 4 * B.pi / 3) < 1e-13
    with pytest.raises(ValueError):
        B.SphericalDesign(0)                                   # SphericalDesign.cpp:7-9
    with pytest.raises(ValueError):
        B.SphericalDesign(7)                                   # SphericalDesign.cpp:22-23


def test_design_directory_loader(tmp_path):
    sd = B.SphericalDesign(6)
    with open(tmp_path / "ss003.006.txt", "w") as fh:
        for x, y, z in zip(sd.getx(), sd.gety(), sd.getz()):
            fh.write(f"  {x: .16e}  {y: .16e}  {z: .16e}\n")
    sd2 = B.SphericalDesign(6, design_dir=str(tmp_path))
    assert np.array_equal(sd2.getx(), sd.getx()) and np.array_equal(sd2.getz(), sd.getz())
    with pytest.raises(RuntimeError):
        B.SphericalDesign(12, design_dir=str(tmp_path))        # SphericalDesign.cpp:29-31


def test_non_antipodal_quadrature_detected():
    q = B.SphericalQuadrature([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0])
    assert not q.is_antipodal()
    q2 = B.SphericalQuadrature([1.0, -1.0], [0.0, -0.0], [0.0, -0.0], [1.0, 2.0])
    assert not q2.is_antipodal()                               # unequal weights


def test_shard_ranges_partition_the_work_list():
    for total in (0, 1, 7, 48, 3072, 6144):
        for count in (1, 2, 3, 4, 8):
            edges = [D.shard_range(total, i, count) for i in range(count)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_bkw_input_is_a_normalised_density():
    Nv = 32
    f, Q = B.inputs.bkw(Nv)
    _, dv = B.inputs.velocity_axis(Nv)
    assert abs(f.sum() * dv ** 3 - 1.0) < 1e-6      # unit mass
    assert abs(Q.sum() * dv ** 3) < 1e-8            # collisions conserve mass
    assert f.min() >= 0


def test_rk4_integrator_with_cpu_stand_in():
    """The integrator is backend agnostic: with dQ/dt = -f it must reproduce exp(-t) to RK4 order."""
    I = B.submodule("integrate")
    f0 = np.array([1.0, 2.0, -3.0])
    f, steps, evals = I.rk4_numpy(lambda x: -x, f0.copy(), 0.0, 1.0, 0.1)
    assert steps == 10 and evals == 40
    assert np.abs(f - f0 * np.exp(-1.0)).max() < 5e-6
    # exact BKW solution: unit mass, positive for t > 6 ln(5/2)
    fb = I.bkw_exact(32, 6.0)
    _, dv = B.inputs.velocity_axis(32)
    assert abs(fb.sum() * dv ** 3 - 1.0) < 1e-6 and fb.min() >= 0
    assert np.allclose(I.bkw_exact(16, 6.5), B.inputs.bkw(16)[0], rtol=1e-14, atol=0)   # same formula, different op order
