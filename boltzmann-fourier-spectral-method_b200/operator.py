"""Python mirror of the reference operator interface for the B200 backend.

`BoltzmannOperatorB200` has the reference's surface (Collisions/AbstractCollisionOperator.hpp:7-26,
ctor of Collisions/CUDABoltzmannOperator.hpp:48-54): construct with the two quadrature objects,
the grid sizes and (gamma, b_gamma, L); `initialize()` once; then `computeCollision(Q, f_in)` /
`op(Q, f_in)` any number of times.  All arithmetic happens in csrc/libbfsm_b200.so through the C
ABI (include/bfsm_b200.h); torch is used only for device memory and streams.

Pointer convention: torch CUDA tensors are passed as device pointers (the CUDA backend's
convention, CUDABoltzmannOperator.cu:119-134); numpy arrays / CPU tensors are host pointers
(the FFTW backend's convention) and go through bfsm_collide_host (H2D + evaluate + D2H).
"""
import ctypes

import numpy as np

from . import _capi


def _torch():
    import torch
    return torch


class BoltzmannOperatorB200:
    def __init__(self, gl_quadrature, spherical_quadrature, Nvx, Nvy, Nvz, gamma, b_gamma, L,
                 device=None, shard_index=0, shard_count=1, fold=True, pack=True, options=None, general=False):
        # like the reference constructors: store arguments only
        self.gl_quadrature = gl_quadrature
        self.spherical_quadrature = spherical_quadrature
        self.Nvx, self.Nvy, self.Nvz = int(Nvx), int(Nvy), int(Nvz)
        self.gamma, self.b_gamma, self.L = float(gamma), float(b_gamma), float(L)
        self.device = device
        self.shard_index, self.shard_count = int(shard_index), int(shard_count)
        self.fold = bool(fold)
        self.pack = bool(pack)
        #: force the general-grid path even on a cubic 16/32/64 grid (tests)
        self.general = bool(general)
        #: dict of bfsm_plan_options fields (chunk_pairs, pencil_kernel, ...); None = defaults
        self.options = dict(options or {})
        self._plan = None
        self._lib = None

    # ------------------------------------------------------------------ lifecycle
    def initialize(self):
        if self._plan is not None:
            return
        lib = _capi.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("B200 backend needs a CUDA device (there is no CPU fallback)")
        dev = self.device
        if dev is None:
            dev = torch.cuda.current_device()
        elif not isinstance(dev, int):
            idx = torch.device(dev).index
            dev = torch.cuda.current_device() if idx is None else idx   # 'cuda' = the current device
        self.device = int(dev)

        def arr(v):
            return np.ascontiguousarray(np.asarray(v, dtype=np.float64))

        rho, w_r = arr(self.gl_quadrature.getNodes()), arr(self.gl_quadrature.getWeights())
        sx, sy, sz = (arr(self.spherical_quadrature.getx()), arr(self.spherical_quadrature.gety()),
                      arr(self.spherical_quadrature.getz()))
        w_s = arr(self.spherical_quadrature.getWeights())
        dp = ctypes.POINTER(ctypes.c_double)
        plan = ctypes.c_void_p()
        opts = _capi.default_options(**self.options)
        rc = lib.bfsm_plan_create_ex(
            ctypes.byref(plan), self.Nvx, self.Nvy, self.Nvz,
            len(rho), rho.ctypes.data_as(dp), w_r.ctypes.data_as(dp),
            len(sx), sx.ctypes.data_as(dp), sy.ctypes.data_as(dp), sz.ctypes.data_as(dp),
            w_s.ctypes.data_as(dp), self.gamma, self.b_gamma, self.L, self.device,
            self.shard_index, self.shard_count,
            (0 if self.fold else _capi.BFSM_FLAG_NO_FOLD) | (0 if self.pack else _capi.BFSM_FLAG_NO_PACK)
            | (_capi.BFSM_FLAG_GENERAL if self.general else 0), ctypes.byref(opts))
        _capi.check(rc)
        self._plan = plan
        self._lib = lib

    def close(self):
        if self._plan is not None and self._lib is not None:
            self._lib.bfsm_plan_destroy(self._plan)
        self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getBackendName(self):
        return "B200"

    # ------------------------------------------------------------------ helpers
    @property
    def grid_size(self):
        return self.Nvx * self.Nvy * self.Nvz

    def info(self):
        self._require()
        info = _capi.PlanInfo()
        _capi.check(self._lib.bfsm_plan_get_info(self._plan, ctypes.byref(info)))
        return {name: getattr(info, name) for name, _ in info._fields_}

    def set_chunk(self, chunk_pairs):
        self._require()
        _capi.check(self._lib.bfsm_plan_set_chunk(self._plan, int(chunk_pairs)))

    def _require(self):
        if self._plan is None:
            raise RuntimeError("initialize() has not been called")

    def _stream(self, stream):
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        return ctypes.c_void_p(stream.cuda_stream)

    def _check_dev(self, t, numel, name):
        torch = _torch()
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise TypeError(f"{name} must be a CUDA tensor")
        if t.dtype != torch.float64 or not t.is_contiguous():
            raise TypeError(f"{name} must be a contiguous float64 tensor")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on {t.device}, the plan on cuda:{self.device}")
        if t.numel() < numel:
            raise ValueError(f"{name} has {t.numel()} elements, needs {numel}")

    # ------------------------------------------------------------------ hot path
    def computeCollision(self, Q, f_in, stream=None, n_cells=None):
        """Q <- Q(f_in, f_in).  Q and f_in hold `n_cells` consecutive N-point grids."""
        self._require()
        torch = _torch()
        N = self.grid_size
        if isinstance(f_in, torch.Tensor) and f_in.is_cuda:
            if n_cells is None:
                n_cells = f_in.numel() // N
                if n_cells < 1 or f_in.numel() != n_cells * N:
                    raise ValueError(f"f_in has {f_in.numel()} elements: not a positive multiple of "
                                     f"the grid size {N}")
            if n_cells < 0:
                raise ValueError("n_cells must be >= 0")   # an explicit 0 is an empty batch: no-op
            self._check_dev(f_in, n_cells * N, "f_in")
            self._check_dev(Q, n_cells * N, "Q")
            rc = self._lib.bfsm_collide(self._plan, ctypes.c_void_p(Q.data_ptr()),
                                        ctypes.c_void_p(f_in.data_ptr()), int(n_cells),
                                        self._stream(stream))
            _capi.check(rc)
            return Q
        # host pointers (numpy arrays or CPU tensors)
        f_np = f_in.numpy() if isinstance(f_in, torch.Tensor) else np.asarray(f_in)
        q_np = Q.numpy() if isinstance(Q, torch.Tensor) else Q
        if not isinstance(q_np, np.ndarray):
            raise TypeError("Q must be a numpy array or tensor")
        for a, name in ((f_np, "f_in"), (q_np, "Q")):
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{name} must be a C-contiguous float64 array")
        if n_cells is None:
            n_cells = f_np.size // N
        if f_np.size != n_cells * N or q_np.size != n_cells * N:
            raise ValueError("f_in / Q size does not match n_cells * Nvx*Nvy*Nvz")
        with torch.cuda.device(self.device):
            rc = self._lib.bfsm_collide_host(self._plan, ctypes.c_void_p(q_np.ctypes.data),
                                             ctypes.c_void_p(f_np.ctypes.data), int(n_cells),
                                             self._stream(stream))
        _capi.check(rc)
        return Q

    def __call__(self, Q, f_in, **kw):
        return self.computeCollision(Q, f_in, **kw)

    def profile(self, Q, f_in, stream=None):
        """One evaluation with CUDA events around every launch group.

        Returns {class name: (milliseconds, launch groups)} (see BFSM_KCLASS_* in bfsm_b200.h)."""
        self._require()
        self._check_dev(f_in, self.grid_size, "f_in")
        self._check_dev(Q, self.grid_size, "Q")
        n = len(_capi.KCLASS_NAMES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int * n)()
        _capi.check(self._lib.bfsm_collide_profiled(
            self._plan, ctypes.c_void_p(Q.data_ptr()), ctypes.c_void_p(f_in.data_ptr()),
            self._stream(stream), ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_capi.KCLASS_NAMES)}

    # ------------------------------------------------------------------ streaming host buffers
    def submit_host(self, Q, f_in, comm=None, n_cells=None, stream=None):
        """Pipelined evaluation with HOST buffers (numpy arrays or CPU tensors, ideally pinned):
        enqueues H2D copy, evaluation and D2H copy of this step and returns; the copies of neighbouring
        steps overlap the kernels (bfsm_collide_host_async).  Q is valid after flush_host(), or once
        BFSM_HOST_PIPE_DEPTH (4) further steps have been submitted.  `comm`: NcclCommunicator for a sharded plan."""
        self._require()
        torch = _torch()
        f_np = f_in.numpy() if isinstance(f_in, torch.Tensor) else np.asarray(f_in)
        q_np = Q.numpy() if isinstance(Q, torch.Tensor) else Q
        for a, name in ((f_np, "f_in"), (q_np, "Q")):
            if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{name} must be a C-contiguous float64 host array")
        N = self.grid_size
        if n_cells is None:
            n_cells = f_np.size // N
        if n_cells < 1 or f_np.size != n_cells * N or q_np.size != n_cells * N:
            raise ValueError("f_in / Q size does not match n_cells * Nvx*Nvy*Nvz")
        with torch.cuda.device(self.device):
            _capi.check(self._lib.bfsm_collide_host_async(
                self._plan, comm.handle if comm is not None else None, ctypes.c_void_p(q_np.ctypes.data),
                ctypes.c_void_p(f_np.ctypes.data), int(n_cells), self._stream(stream)))

    def flush_host(self):
        """Wait for every step submitted with submit_host()."""
        self._require()
        _capi.check(self._lib.bfsm_collide_host_flush(self._plan))

    # ------------------------------------------------------------------ diagnostics
    def moments(self, g, n_cells=None, stream=None):
        """Per-cell velocity moments of the grids in the CUDA tensor `g` (f or Q): a (n_cells, 5) CUDA
        tensor of dv^3 sum_v g (1, vx, vy, vz, |v|^2/2).  For f: density, momentum, energy density; for
        Q(f,f): the conservation defects (mass, momentum and energy of the collision term vanish)."""
        self._require()
        torch = _torch()
        N = self.grid_size
        if n_cells is None:
            n_cells = g.numel() // N
        self._check_dev(g, n_cells * N, "g")
        out = torch.empty((n_cells, 5), dtype=torch.float64, device=g.device)
        _capi.check(self._lib.bfsm_moments(self._plan, ctypes.c_void_p(g.data_ptr()), int(n_cells),
                                           ctypes.c_void_p(out.data_ptr()), self._stream(stream)))
        return out

    # ------------------------------------------------------------------ multi-GPU (collective in C)
    def collide_partial(self, Q_partial, f_in, stream=None):
        """Q_partial <- this pair shard's partial Q (shard 0 carries the loss term): the sum over all
        shards is Q(f,f).  For callers that own the exchange step."""
        self._require()
        self._check_dev(f_in, self.grid_size, "f_in")
        self._check_dev(Q_partial, self.grid_size, "Q_partial")
        _capi.check(self._lib.bfsm_collide_partial(self._plan, ctypes.c_void_p(Q_partial.data_ptr()),
                                                   ctypes.c_void_p(f_in.data_ptr()), self._stream(stream)))
        return Q_partial

    def collide_sharded(self, Q, f_in, comm, stream=None):
        """Q <- Q(f_in, f_in) with the pair shards summed by ONE ncclAllReduce issued inside the C
        library (bfsm_collide_sharded); `comm` is a `distributed.NcclCommunicator`."""
        self._require()
        self._check_dev(f_in, self.grid_size, "f_in")
        self._check_dev(Q, self.grid_size, "Q")
        _capi.check(self._lib.bfsm_collide_sharded(self._plan, comm.handle, ctypes.c_void_p(Q.data_ptr()),
                                                   ctypes.c_void_p(f_in.data_ptr()), self._stream(stream)))
        return Q

    # ------------------------------------------------------------------ multi-GPU halves
    def gain_hat(self, Qhat, f_in, stream=None):
        """Qhat (2*N doubles, complex interleaved) <- this shard's partial gain spectrum."""
        self._require()
        self._check_dev(f_in, self.grid_size, "f_in")
        self._check_dev(Qhat, 2 * self.grid_size, "Qhat")
        _capi.check(self._lib.bfsm_gain_hat(self._plan, ctypes.c_void_p(Qhat.data_ptr()),
                                            ctypes.c_void_p(f_in.data_ptr()), self._stream(stream)))
        return Qhat

    def finish(self, Q, Qhat, f_in, stream=None):
        """Q <- Re(IFFT3(Qhat)) - loss(f_in); call after gain_hat on the same f_in."""
        self._require()
        self._check_dev(f_in, self.grid_size, "f_in")
        self._check_dev(Qhat, 2 * self.grid_size, "Qhat")
        self._check_dev(Q, self.grid_size, "Q")
        _capi.check(self._lib.bfsm_finish(self._plan, ctypes.c_void_p(Q.data_ptr()),
                                          ctypes.c_void_p(Qhat.data_ptr()),
                                          ctypes.c_void_p(f_in.data_ptr()), self._stream(stream)))
        return Q
