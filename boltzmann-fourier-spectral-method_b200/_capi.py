"""ctypes binding of include/bfsm_b200.h (the C ABI of the CUDA library).

There is no fallback: if the shared library is missing, `load()` raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libbfsm_b200.so")

BFSM_OK, BFSM_ERR_INVALID, BFSM_ERR_UNSUPPORTED, BFSM_ERR_CUDA, BFSM_ERR_NOMEM, BFSM_ERR_COMM = range(6)
BFSM_UNIQUE_ID_BYTES = 128
BFSM_HOST_PIPE_DEPTH = 4  # steps bfsm_collide_host_async keeps in flight
BFSM_FLAG_NO_FOLD = 1
BFSM_FLAG_NO_PACK = 2
BFSM_FLAG_GENERAL = 4

#: every symbol include/bfsm_b200.h declares
EXPORTS = (
    "bfsm_version", "bfsm_last_error", "bfsm_plan_create", "bfsm_plan_destroy", "bfsm_collide",
    "bfsm_collide_host", "bfsm_gain_hat", "bfsm_finish", "bfsm_plan_get_info",
    "bfsm_plan_set_chunk", "bfsm_collide_profiled", "bfsm_sync", "bfsm_device_malloc",
    "bfsm_device_free", "bfsm_copy_to_device", "bfsm_copy_to_host", "bfsm_measure_fp64_peak",
    "bfsm_debug_plane_work", "bfsm_debug_plane_work_r32", "bfsm_debug_shares_aligned", "bfsm_debug_fail_lane_alloc", "bfsm_debug_units",
    "bfsm_plan_options_init", "bfsm_plan_create_ex", "bfsm_comm_unique_id", "bfsm_comm_init_rank",
    "bfsm_comm_init_all", "bfsm_comm_adopt", "bfsm_comm_destroy", "bfsm_collide_sharded",
    "bfsm_collide_sharded_group", "bfsm_collide_partial", "bfsm_vec_axpby", "bfsm_moments",
    "bfsm_collide_host_async", "bfsm_collide_host_flush",
)

KCLASS_NAMES = ("forward", "plane_gain", "pencil_gain", "accum", "final", "nyquist")


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int), ("n_r", ctypes.c_int), ("n_s", ctypes.c_int),
        ("folded", ctypes.c_int), ("packed", ctypes.c_int), ("pairs_total", ctypes.c_int), ("pairs_local", ctypes.c_int),
        ("chunk_pairs", ctypes.c_int), ("launches_per_cell", ctypes.c_int),
        ("scratch_bytes", ctypes.c_longlong), ("plane_kernel", ctypes.c_int),
        ("partial_slots", ctypes.c_int), ("pencil_kernel", ctypes.c_int),
        ("batch_lanes_used", ctypes.c_int), ("gain_pipeline", ctypes.c_int),
        ("ny", ctypes.c_int), ("nz", ctypes.c_int), ("general", ctypes.c_int),
        ("batch_group_cells", ctypes.c_int),
    ]


class PlanOptions(ctypes.Structure):
    """Mirror of bfsm_plan_options (include/bfsm_b200.h); fill with options_from_env() or by hand
    after bfsm_plan_options_init."""
    _fields_ = [
        ("struct_size", ctypes.c_int), ("chunk_pairs", ctypes.c_int), ("pencil_kernel", ctypes.c_int),
        ("seg_pairs", ctypes.c_int), ("plane_kernel", ctypes.c_int),
        ("side_stream", ctypes.c_int), ("batch_lanes", ctypes.c_int), ("gain_ctas", ctypes.c_int),
        ("gain_pipeline", ctypes.c_int), ("fused_sub_pairs", ctypes.c_int), ("fused_ring", ctypes.c_int),
        ("fused_pencil_ctas", ctypes.c_int), ("fused_nyq_ctas", ctypes.c_int),
        ("pencil_groups", ctypes.c_int), ("reserved", ctypes.c_int * 2),
    ]


#: test/tuning override: environment variable -> options field.  The LIBRARY reads no environment
#: variables; only this Python harness does, so that tools/ab_plane.py and the variant tests can
#: select kernels per process.
ENV_OPTIONS = {
    "BFSM_CHUNK_PAIRS": "chunk_pairs", "BFSM_PENCIL_KERNEL": "pencil_kernel",
    "BFSM_SEG_PAIRS": "seg_pairs", "BFSM_PLANE_KERNEL": "plane_kernel",
    "BFSM_SIDE_STREAM": "side_stream",
    "BFSM_BATCH_LANES": "batch_lanes", "BFSM_GAIN_CTAS": "gain_ctas",
    "BFSM_GAIN_PIPELINE": "gain_pipeline", "BFSM_FUSED_SUB_PAIRS": "fused_sub_pairs",
    "BFSM_FUSED_RING": "fused_ring", "BFSM_FUSED_PENCIL_CTAS": "fused_pencil_ctas",
    "BFSM_FUSED_NYQ_CTAS": "fused_nyq_ctas", "BFSM_PENCIL_GROUPS": "pencil_groups",
}


def default_options(**overrides):
    """bfsm_plan_options with the library defaults, BFSM_* environment overrides, then `overrides`."""
    opts = PlanOptions()
    load().bfsm_plan_options_init(ctypes.byref(opts))
    for env, field in ENV_OPTIONS.items():
        val = os.environ.get(env)
        if val not in (None, ""):
            setattr(opts, field, int(val))
    for field, val in overrides.items():
        if val is not None:
            setattr(opts, field, int(val))
    return opts


class BfsmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"bfsm error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load csrc/libbfsm_b200.so and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    vp = ctypes.c_void_p
    lib.bfsm_version.restype = ctypes.c_int
    lib.bfsm_last_error.restype = ctypes.c_char_p
    lib.bfsm_plan_create.restype = ctypes.c_int
    lib.bfsm_plan_create.argtypes = [
        ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, dp, dp, ctypes.c_int, dp, dp, dp, dp,
        ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
    lib.bfsm_plan_create_ex.restype = ctypes.c_int
    lib.bfsm_plan_create_ex.argtypes = lib.bfsm_plan_create.argtypes + [ctypes.POINTER(PlanOptions)]
    lib.bfsm_plan_options_init.restype = None
    lib.bfsm_plan_options_init.argtypes = [ctypes.POINTER(PlanOptions)]
    lib.bfsm_comm_unique_id.restype = ctypes.c_int
    lib.bfsm_comm_unique_id.argtypes = [ctypes.c_char_p]
    lib.bfsm_comm_init_rank.restype = ctypes.c_int
    lib.bfsm_comm_init_rank.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.bfsm_comm_init_all.restype = ctypes.c_int
    lib.bfsm_comm_init_all.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.bfsm_comm_adopt.restype = ctypes.c_int
    lib.bfsm_comm_adopt.argtypes = [ctypes.POINTER(vp), vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.bfsm_comm_destroy.restype = ctypes.c_int
    lib.bfsm_comm_destroy.argtypes = [vp]
    lib.bfsm_collide_sharded.restype = ctypes.c_int
    lib.bfsm_collide_sharded.argtypes = [vp, vp, vp, vp, vp]
    lib.bfsm_collide_sharded_group.restype = ctypes.c_int
    lib.bfsm_collide_sharded_group.argtypes = [ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                               ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    lib.bfsm_collide_partial.restype = ctypes.c_int
    lib.bfsm_collide_partial.argtypes = [vp, vp, vp, vp]
    lib.bfsm_collide_host_async.restype = ctypes.c_int
    lib.bfsm_collide_host_async.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
    lib.bfsm_collide_host_flush.restype = ctypes.c_int
    lib.bfsm_collide_host_flush.argtypes = [vp]
    lib.bfsm_vec_axpby.restype = ctypes.c_int
    lib.bfsm_vec_axpby.argtypes = [ctypes.c_int, vp, ctypes.c_double, vp, ctypes.c_double, vp,
                                   ctypes.c_ulonglong, vp]
    lib.bfsm_moments.restype = ctypes.c_int
    lib.bfsm_moments.argtypes = [vp, vp, ctypes.c_int, vp, vp]
    lib.bfsm_debug_units.restype = ctypes.c_int
    lib.bfsm_debug_units.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.bfsm_debug_fail_lane_alloc.restype = ctypes.c_int
    lib.bfsm_debug_fail_lane_alloc.argtypes = [vp, ctypes.c_int]
    lib.bfsm_plan_destroy.restype = ctypes.c_int
    lib.bfsm_plan_destroy.argtypes = [vp]
    lib.bfsm_collide.restype = ctypes.c_int
    lib.bfsm_collide.argtypes = [vp, vp, vp, ctypes.c_int, vp]
    lib.bfsm_collide_host.restype = ctypes.c_int
    lib.bfsm_collide_host.argtypes = [vp, vp, vp, ctypes.c_int, vp]
    lib.bfsm_gain_hat.restype = ctypes.c_int
    lib.bfsm_gain_hat.argtypes = [vp, vp, vp, vp]
    lib.bfsm_finish.restype = ctypes.c_int
    lib.bfsm_finish.argtypes = [vp, vp, vp, vp, vp]
    lib.bfsm_plan_get_info.restype = ctypes.c_int
    lib.bfsm_plan_get_info.argtypes = [vp, ctypes.POINTER(PlanInfo)]
    lib.bfsm_collide_profiled.restype = ctypes.c_int
    lib.bfsm_collide_profiled.argtypes = [vp, vp, vp, vp, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_int)]
    lib.bfsm_measure_fp64_peak.restype = ctypes.c_int
    lib.bfsm_measure_fp64_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.bfsm_plan_set_chunk.restype = ctypes.c_int
    lib.bfsm_plan_set_chunk.argtypes = [vp, ctypes.c_int]
    ip = ctypes.POINTER(ctypes.c_int)
    lib.bfsm_debug_shares_aligned.restype = ctypes.c_int
    lib.bfsm_debug_shares_aligned.argtypes = [ctypes.c_int] * 5
    lib.bfsm_debug_plane_work.restype = ctypes.c_int
    lib.bfsm_debug_plane_work.argtypes = [ctypes.c_int] * 4 + [ip, ip, ctypes.c_int]
    lib.bfsm_debug_plane_work_r32.restype = ctypes.c_int
    lib.bfsm_debug_plane_work_r32.argtypes = [ctypes.c_int] * 4 + [ip, ip, ctypes.c_int]
    _lib = lib
    return lib


def check(rc):
    if rc != BFSM_OK:
        msg = load().bfsm_last_error()
        raise BfsmError(rc, msg.decode() if msg else "")


def measure_fp64_peak(device=0):
    """Peak DFMA/s of the FP64 pipe, measured on the device (bfsm_measure_fp64_peak)."""
    out = ctypes.c_double()
    check(load().bfsm_measure_fp64_peak(int(device), ctypes.byref(out)))
    return out.value
