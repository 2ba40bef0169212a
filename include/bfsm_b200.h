/*
 * bfsm_b200.h -- C ABI of the B200-native (sm_100a) fast Fourier-spectral Boltzmann
 * collision operator Q(f,f).
 *
 * This is the drop-in boundary for ONE hot path of i3s93/Boltzmann-Fourier-Spectral-Method:
 *   BoltzmannOperator<Backend>::computeCollision(double* Q, const double* f_in)
 *     reference CPU:  Collisions/FFTWBoltzmannOperator.cpp:147-334   (parity oracle)
 *     reference GPU:  Collisions/CUDABoltzmannOperator.cu:119-220 + BoltzmannCUDAKernels.cu:4-177
 * A C++ class `BoltzmannOperator<B200_Backend>` (include/B200BoltzmannOperator.hpp) wraps these
 * entry points behind the reference's AbstractCollisionOperator interface
 * (Collisions/AbstractCollisionOperator.hpp:7-26); Python binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, a BFSM_ERR_* code
 *     otherwise, and never calls exit(); bfsm_last_error() returns the message (thread local).
 *   - grids are row-major (i*Nvy + j)*Nvz + k, z fastest, real fp64, N = Nvx*Nvy*Nvz values
 *     (FFTWBoltzmannOperator.cpp:173).  Cubic grids Nvx=Nvy=Nvz in {16, 32, 64} run on the tuned
 *     kernels; every other combination of even sizes from 4 to 128 per axis (non-cubic, 128, sizes
 *     that are not powers of two) runs on a slower general path; odd or larger sizes are rejected
 *     with BFSM_ERR_UNSUPPORTED at plan creation.
 *   - `*_dev` pointers are device pointers on the plan's device (the CUDA backend's convention,
 *     CUDABoltzmannOperator.cu:119-134); `*_host` pointers are host pointers (the FFTW backend's).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer
 *     calls are asynchronous on that stream; host-pointer calls return when the result is ready.
 *   - Q may alias f_in (both reference backends tolerate it: FFTWBoltzmannOperator.cpp:168-180).
 *   - a plan is not re-entrant (it owns its scratch memory), like the reference operators.
 *   - there is NO CPU fallback: without a CUDA device plan creation fails.
 */
#ifndef BFSM_B200_H
#define BFSM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BFSM_VERSION 100 /* 0.1.0 */

enum {
    BFSM_OK = 0,
    BFSM_ERR_INVALID = 1,     /* bad argument (NULL pointer, non-positive size, bad shard) */
    BFSM_ERR_UNSUPPORTED = 2, /* grid shape not supported by this build */
    BFSM_ERR_CUDA = 3,        /* CUDA runtime error, message in bfsm_last_error() */
    BFSM_ERR_NOMEM = 4,
    BFSM_ERR_COMM = 5         /* NCCL missing or a collective failed, message in bfsm_last_error() */
};

/* plan flags */
#define BFSM_FLAG_NO_FOLD 1u /* transform every (r,sigma) pair even if the design is antipodal */
#define BFSM_FLAG_NO_PACK 2u /* two 3-D transforms per pair (g1 and g2) instead of the Hermitian-
                                packed single transform + Nyquist-plane correction */
#define BFSM_FLAG_GENERAL 4u /* use the general-grid path even for a cubic 16/32/64 grid (tests) */

typedef struct bfsm_plan bfsm_plan;
typedef struct bfsm_comm bfsm_comm; /* NCCL communicator wrapper, see bfsm_comm_* below */

int bfsm_version(void);
const char *bfsm_last_error(void);

/*
 * Replaces BoltzmannOperator<Backend>::BoltzmannOperator(...) + initialize()
 *   (FFTWBoltzmannOperator.hpp:30-36, .cpp:14-70; CUDABoltzmannOperator.hpp:48-54, .cu:28-115).
 *
 *   rho[n_r], w_r[n_r]      Gauss-Legendre nodes/weights on [0,R]  (GaussLegendre.hpp:10-24)
 *   sx,sy,sz,w_s[n_s]       spherical quadrature nodes/weights     (SphericalDesign.cpp:38-48)
 *   gamma, b_gamma, L       collision-kernel exponent/constant, domain half-width
 *   device                  CUDA device ordinal
 *   shard_index/shard_count this plan evaluates shard `shard_index` of the (r,sigma) pair list
 *                           split into `shard_count` contiguous, equal (+-1) parts; 0/1 = all pairs.
 *
 * The work list is the N_r x N_sigma pair list; if every direction has a bit-exact antipode with
 * equal weight (true for all nine reference ssXXX.NNN.txt files) only one of each antipodal pair
 * is transformed with doubled weight -- pairs (r,s) and (r,-s) give bitwise identical g1*g2.
 */
int bfsm_plan_create(bfsm_plan **out, int nvx, int nvy, int nvz, int n_r, const double *rho,
                     const double *w_r, int n_s, const double *sx, const double *sy,
                     const double *sz, const double *w_s, double gamma, double b_gamma, double L,
                     int device, int shard_index, int shard_count, unsigned flags);

/*
 * Tuning options of a plan (all optional: bfsm_plan_create uses the defaults).  Fill the struct with
 * bfsm_plan_options_init() first, change what you need, pass it to bfsm_plan_create_ex().  The library
 * reads NO environment variables; the Python test harness maps BFSM_* variables onto this struct.
 */
typedef struct {
    int struct_size;       /* = sizeof(bfsm_plan_options), set by bfsm_plan_options_init */
    int chunk_pairs;       /* pairs per launch of the gain kernels; 0 = heuristic */
    int pencil_kernel;     /* x stage (packed mode): 0 = default, 1 = staged through a shared-memory ring
                              filled by cp.async (LDGSTS), 2 = register resident (no shared memory, LDG +
                              SHFL), 3 = the staged kernel with the ring filled by the TMA unit (one
                              cp.async.bulk.tensor per tile + mbarrier) */
    int seg_pairs;         /* pairs per work unit of the register-resident x stage and of the Nyquist
                              accumulate (one partial slot per unit of a radius); 0 = heuristic */
    int plane_kernel;      /* (y,z) stage (packed mode): 0 = default, 1 = k_plane_gain3 (every warp runs
                              all three stages), 2 = k_plane_gain_ws (warp-specialised pipeline, 64^3),
                              3 = k_plane_gain_r32 (radix-32 register transforms, two stages and one
                              shared-memory exchange per plane; 64^3 and 32^3), 4 = the same with each
                              thread's entries of the fhat plane kept in tensor memory (tcgen05.st/ld)
                              instead of registers: more warps per SM */
    int side_stream;       /* 1: Nyquist accumulate on an internal side stream (default), 0: in line */
    int batch_lanes;       /* cells kept in flight by bfsm_collide(n_cells > 1): 1..4, 0 = default (4); 1 also
                              switches the cell-group path of the 32^3 grids off */
    int gain_ctas;         /* persistent CTAs of k_plane_gain3; 0 = SMs x occupancy */
    int gain_pipeline;     /* 0 = default, 1 = one plane + one x kernel per chunk (hybrid grids through
                              HBM), 2 = fused persistent kernel (64^3 packed mode only): plane, Nyquist and
                              x roles on disjoint SMs, hybrid grids through an L2-resident ring (opt-in:
                              measured slower than pipeline 1), 3 = thread-block-cluster
                              kernel (32^3 packed mode only): eight CTAs own a pair, hybrid grids handed
                              over through distributed shared memory, no global scratch */
    int fused_sub_pairs;   /* fused kernel: pairs per hand-over (sub-chunk); 0 = default */
    int fused_ring;        /* fused kernel: ring slots (sub-chunks in flight); 0 = default (2) */
    int fused_pencil_ctas; /* fused kernel: CTAs (= SMs) of the x role; 0 = default */
    int fused_nyq_ctas;    /* fused kernel: CTAs of the Nyquist-plane role, a multiple of 3; 0 = default */
    int pencil_groups;     /* staged x stage (pencil_kernel 1 / 3): CTA rows a chunk's pairs are split over
                              (grid = tiles x rows; one partial slot per row unless every row's share starts
                              at a radius boundary); 0 = default */
    int reserved[2];
} bfsm_plan_options;

void bfsm_plan_options_init(bfsm_plan_options *opts);

/* bfsm_plan_create with explicit options (NULL = defaults). */
int bfsm_plan_create_ex(bfsm_plan **out, int nvx, int nvy, int nvz, int n_r, const double *rho,
                        const double *w_r, int n_s, const double *sx, const double *sy,
                        const double *sz, const double *w_s, double gamma, double b_gamma, double L,
                        int device, int shard_index, int shard_count, unsigned flags,
                        const bfsm_plan_options *opts);

/* Replaces ~BoltzmannOperator() (FFTWBoltzmannOperator.cpp:338-363). NULL is accepted. */
int bfsm_plan_destroy(bfsm_plan *plan);

/*
 * Replaces computeCollision / operator() (FFTWBoltzmannOperator.cpp:147-334,
 * CUDABoltzmannOperator.cu:119-220) for `n_cells` consecutive N-sized grids:
 * Q_dev[c*N + idx] = Q(f,f)[idx] of cell c.  Requires shard_count == 1.
 */
int bfsm_collide(bfsm_plan *plan, double *Q_dev, const double *f_dev, int n_cells, void *stream);

/* Same with host buffers: H2D copy of f, evaluation, D2H copy of Q, stream synchronised. */
int bfsm_collide_host(bfsm_plan *plan, double *Q_host, const double *f_host, int n_cells,
                      void *stream);

/*
 * Pipelined host-pointer evaluation for callers that stream many evaluations: returns as soon as step k
 * is enqueued -- H2D copy of f on a copy stream, kernels on `stream`, D2H copy of Q on a second copy
 * stream, BFSM_HOST_PIPE_DEPTH staging slots -- so that the copies of neighbouring steps run under the
 * kernels.  The call blocks only until the step submitted BFSM_HOST_PIPE_DEPTH calls earlier has delivered its Q
 * (a caller therefore cycles through at least that many host Q buffers); bfsm_collide_host_flush()
 * waits for everything outstanding.  Host buffers should be page-locked and must stay untouched until
 * their step is complete.  With a sharded plan pass the rank's communicator (one cell per call, every
 * rank submits the same sequence); NULL otherwise.
 */
#define BFSM_HOST_PIPE_DEPTH 4
int bfsm_collide_host_async(bfsm_plan *plan, bfsm_comm *comm_or_null, double *Q_host, const double *f_host,
                            int n_cells, void *stream);
int bfsm_collide_host_flush(bfsm_plan *plan);

/*
 * Multi-GPU split of one evaluation (one cell).  Step 1 on every rank: this shard's partial
 * gain spectrum  Qhat_dev[2*N] (complex, re/im interleaved) = sum over the shard's pairs of
 * W_rs * beta1_r(|l|) * FFT3(Re(g1*g2))   (FFTWBoltzmannOperator.cpp:191-276 restricted to the
 * shard).  Step 2: the caller sums Qhat over ranks (one ncclAllReduce of 2*N doubles).  Step 3:
 * bfsm_finish() = loss term, inverse transform and combine (cpp:281-330).  bfsm_finish must be
 * called with the same f_dev as the preceding bfsm_gain_hat on this plan (it reuses FFT3(f)).
 *
 * Only Re(g1*g2) is transformed: beta1 is real and even in l, so Im(g1*g2) only feeds
 * Im(Q_gain), which the reference discards (cpp:326).  Hence Qhat is the Hermitian part of the
 * reference's Q_gain_hat; Q is unchanged.
 */
int bfsm_gain_hat(bfsm_plan *plan, double *Qhat_dev, const double *f_dev, void *stream);
int bfsm_finish(bfsm_plan *plan, double *Q_dev, const double *Qhat_dev, const double *f_dev,
                void *stream);

/*
 * Pair-sharded evaluation with the collective under the ABI ("N_r x N_sigma pairs
 * shard over 2/4/8 GPUs, followed by one NCCL allreduce").  A communicator wraps an ncclComm_t; NCCL is
 * bound at run time (dlopen), so the library has no link-time NCCL dependency.
 *
 *   one process per GPU:  rank 0 calls bfsm_comm_unique_id() and sends the 128 bytes to the other ranks
 *                         (any side channel); every rank calls bfsm_comm_init_rank();
 *   one process, n GPUs:  bfsm_comm_init_all() (ncclCommInitAll), then bfsm_collide_sharded_group();
 *   existing ncclComm_t:  bfsm_comm_adopt() (not destroyed by bfsm_comm_destroy).
 *
 * bfsm_collide_sharded: plan k of W (shard_index = k, shard_count = W = communicator size) evaluates its
 * shard, transforms its partial gain spectrum back to physical space (linear, so the partial Q's add
 * up), rank 0 subtracts the loss term, and ONE ncclAllReduce(sum) of N real doubles, in place in Q_dev,
 * leaves Q(f,f) on every rank.  Asynchronous on `stream`.
 */
#define BFSM_UNIQUE_ID_BYTES 128
int bfsm_comm_unique_id(unsigned char *id /* [BFSM_UNIQUE_ID_BYTES] */);
int bfsm_comm_init_rank(bfsm_comm **out, const unsigned char *id, int n_ranks, int rank, int device);
int bfsm_comm_init_all(bfsm_comm **comms /* [n_devices] */, int n_devices, const int *devices);
int bfsm_comm_adopt(bfsm_comm **out, void *nccl_comm, int n_ranks, int rank, int device);
int bfsm_comm_destroy(bfsm_comm *comm);
int bfsm_collide_sharded(bfsm_plan *plan, bfsm_comm *comm, double *Q_dev, const double *f_dev, void *stream);
int bfsm_collide_sharded_group(int n_ranks, bfsm_plan **plans, bfsm_comm **comms, double **Q_dev,
                               const double **f_dev, void **streams /* may be NULL */);
/* The shard's partial Q without the collective (sum over all shards = Q(f,f)); for callers that own the
 * exchange step. */
int bfsm_collide_partial(bfsm_plan *plan, double *Q_partial_dev, const double *f_dev, void *stream);

/* cudaStreamSynchronize(stream) on the plan's device. */
int bfsm_sync(bfsm_plan *plan, void *stream);

/* Device-memory helpers so that C/C++ hosts can link this library alone (no CUDA headers):
 * thin wrappers over cudaMalloc / cudaFree / cudaMemcpy on device ordinal `device`. */
int bfsm_device_malloc(int device, void **ptr, unsigned long long bytes);
int bfsm_device_free(int device, void *ptr);
int bfsm_copy_to_device(int device, void *dst_dev, const void *src_host, unsigned long long bytes);
int bfsm_copy_to_host(int device, void *dst_host, const void *src_dev, unsigned long long bytes);

/*
 * Callers' helpers (the steps either side of the path):
 *   bfsm_vec_axpby  out = a x + b y on `n` device doubles (out may alias x or y): the stage update of an
 *                   explicit time integrator, so that RK stages never leave the device;
 *   bfsm_moments    per-cell velocity moments of `n_cells` consecutive grids g (f or Q):
 *                   moments_dev[5 c .. 5 c + 4] = dv^3 sum_v g (1, vx, vy, vz, |v|^2/2) on the plan's
 *                   grid v_i = -L + dv/2 + i dv (maxwell_bkw_fftw.cpp:62-71).  Of f: density, momentum,
 *                   energy; of Q(f,f): the conservation defects (all five vanish analytically).
 */
int bfsm_vec_axpby(int device, double *out_dev, double a, const double *x_dev, double b, const double *y_dev,
                   unsigned long long n, void *stream);
int bfsm_moments(bfsm_plan *plan, const double *g_dev, int n_cells, double *moments_dev, void *stream);

/* Introspection (used by bench.py for the roofline arithmetic and by tests). */
typedef struct {
    int n;                   /* points per axis */
    int n_r, n_s;            /* quadrature sizes as given */
    int folded;              /* 1 if antipodal folding is active */
    int packed;              /* 1 if Hermitian packing (one transform per pair) is active */
    int pairs_total;         /* size of the (possibly folded) work list, all shards */
    int pairs_local;         /* pairs this plan transforms */
    int chunk_pairs;         /* pairs per gain-kernel launch */
    int launches_per_cell;   /* kernel launches issued per evaluated cell */
    long long scratch_bytes; /* device memory owned by the plan */
    int plane_kernel;        /* gain plane kernel in use: 0 k_plane_gain (4-pass, unpacked mode),
                                1 k_plane_gain3, 2 k_plane_gain_ws (warp-specialised pipeline, 64^3),
                                3 k_plane_gain_r32 (radix-32, two stages), 4 the same with the fhat line
                                in tensor memory */
    int partial_slots;       /* partial-sum slots of S_r summed per evaluation */
    int pencil_kernel;       /* x stage in use: 0 k_pencil_gain (unpacked mode), 1 k_pencil_gain_async
                                (LDGSTS-filled ring), 2 k_pencil_gain_reg (register resident), 3
                                k_pencil_gain_async with a TMA-filled ring */
    int batch_lanes_used;    /* lanes the last bfsm_collide(n_cells > 1) ran on (1 before any) */
    int gain_pipeline;       /* 1 = plane + x kernel per chunk, 2 = fused persistent kernel, 3 = cluster,
                                0 = general-grid path */
    int ny, nz;              /* points along y and z (`n` is the x size) */
    int general;             /* 1 if the general-grid path (bfsm_general.cuh) serves this plan */
    int batch_group_cells;   /* cells per kernel launch the last bfsm_collide(n_cells > 1) used: > 0 = the
                                cell-group path (32^3: every kernel has a cell dimension; batch_lanes_used
                                then counts the groups in flight), 0 = one cell per launch on lanes */
} bfsm_plan_info;

int bfsm_plan_get_info(const bfsm_plan *plan, bfsm_plan_info *info);

/*
 * Measurement aid (bench.py): one evaluation of one cell with a CUDA-event pair recorded on
 * `stream` around every launch group; returns after synchronising the stream.
 * ms_by_class[c] / launches_by_class[c] (arrays of BFSM_KCLASS_COUNT) receive the summed
 * device time and the number of bracketed launch groups of each kernel class.
 */
enum {
    BFSM_KCLASS_FORWARD = 0,     /* k_plane<REAL> + k_pencil_fwd : fhat = FFT3(f)          */
    BFSM_KCLASS_PLANE_GAIN = 1,  /* k_plane_gain* : phase multiply + 2-D inverse FFT (y,z)  */
    BFSM_KCLASS_PENCIL_GAIN = 2, /* k_pencil_gain : inverse FFT (x) + product + accumulate  */
    BFSM_KCLASS_ACCUM = 3,       /* k_plane<REAL> + k_pencil_accum : Qhat = sum_r ...       */
    BFSM_KCLASS_FINAL = 4,       /* k_plane<FINAL> + k_pencil_final : loss + combine        */
    BFSM_KCLASS_NYQUIST = 5,     /* k_extract_nyq + k_plane_nyq + k_nyq_accum (packed mode) */
    BFSM_KCLASS_COUNT = 6
};
int bfsm_collide_profiled(bfsm_plan *plan, double *Q_dev, const double *f_dev, void *stream,
                          double *ms_by_class, int *launches_by_class);

/* Measurement aid: peak rate of the FP64 pipe on `device` (independent DFMA chains, no memory
 * traffic), in fused multiply-adds per second; 2x that is the usual FLOP/s figure. */
int bfsm_measure_fp64_peak(int device, double *dfma_per_second);

/* Test aid (no device needed): the (plane, item) entries that CTA `cta` of `n_ctas` walks in the gain
 * plane kernels for a launch of `n_items` pairs on an n^3 grid -- planes 0..n-1 are the regular
 * (y,z) planes, n..n+2 the three Nyquist planes.  Writes up to `capacity` entries and returns the
 * CTA's entry count (negative BFSM_ERR_* code on bad arguments). */
int bfsm_debug_plane_work(int n, int n_items, int n_ctas, int cta, int *planes, int *items,
                          int capacity);
/* Same for k_plane_gain_r32 (n = 32 or 64): the entries group `group` of `n_groups` walks -- one
 * contiguous range of the flat list per group, the three Nyquist planes on their own groups. */
int bfsm_debug_plane_work_r32(int n, int n_items, int n_groups, int group, int *planes, int *items,
                              int capacity);

/* Test aid (no device needed): 1 if, for every launch of `chunk` pairs over a shard of `pairs_local`
 * pairs starting at global pair `pair_lo`, the `groups` equal shares of the launch all start at a
 * radius boundary of the r-major pair list (n_dir pairs per radius) -- the condition under which the
 * staged x stage and the Nyquist accumulate share one partial-sum slot per kernel. */
int bfsm_debug_shares_aligned(int pairs_local, int pair_lo, int n_dir, int chunk, int groups);

/* Test aid (no device needed): the work units the register-resident x stage cuts a shard's pair list
 * into (see bfsm_plan_options.seg_pairs): out[4k .. 4k+3] = {first pair, one past the last pair, local
 * radius index, partial slot} of unit k, pairs counted from the shard's first pair.  Every unit lies
 * inside one launch (chunk) and one radius.  Writes up to `capacity` units, returns the unit count
 * (negative BFSM_ERR_* code on bad arguments). */
int bfsm_debug_units(int pairs_local, int pair_lo, int n_dir, int chunk, int seg_pairs, int *out,
                     int capacity);

/* Test aid: makes the allocation of batch lanes >= `first_failing_lane` fail with BFSM_ERR_NOMEM
 * (0 = off), to exercise the "run with the lanes that exist" path of bfsm_collide(n_cells > 1). */
int bfsm_debug_fail_lane_alloc(bfsm_plan *plan, int first_failing_lane);

/* Tuning knob: pairs per launch of the gain kernels (0 = default heuristic). */
int bfsm_plan_set_chunk(bfsm_plan *plan, int chunk_pairs);

#ifdef __cplusplus
}
#endif

#endif /* BFSM_B200_H */
