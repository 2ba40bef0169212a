// bfsm_capi.cu -- plan management, host-side table construction and kernel launches behind
// the C ABI declared in include/bfsm_b200.h.  No torch types, no cuFFT, no CPU fallback.
#include "../../include/bfsm_b200.h"
#include "bfsm_kernels.cuh"
#include "bfsm_pencil_reg.cuh"
#include "bfsm_plane_r32.cuh"
#include "bfsm_fused.cuh"
#include "bfsm_cluster.cuh"
#include "bfsm_general.cuh"
#include "bfsm_aux.cuh"

#include <dlfcn.h>
#include <nccl.h> // types and enums only: the library is loaded with dlopen at first use

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

using namespace bfsm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                     \
            return fail(BFSM_ERR_CUDA, buf_);                                                 \
        }                                                                                     \
    } while (0)

const long double PI_L = 3.14159265358979323846264338327950288L;
const double PI_D = 3.14159265358979323846; // Utilities/constants.hpp:7

// FFTWBoltzmannOperator.hpp:17-21
double sincc(double x)
{
    const double eps = 2.220446049250313e-16;
    return std::sin(x + eps) / (x + eps);
}

} // namespace

struct bfsm_plan {
    int N = 0, n_r = 0, n_s = 0, device = 0;
    double gamma = 0, b_gamma = 0, L = 0;
    int folded = 0;
    int packed = 1;   // Hermitian packing: one 3-D transform per pair + Nyquist-plane correction
    int pencil_kernel = 2; // x stage (packed): 1 = staged cp.async ring, 2 = register resident (units)
    int plane_ws = 0;      // warp-specialised pipelined plane kernel (packed mode, N = 64)
    int plane_r32 = 0;     // radix-32 two-stage plane kernel (packed mode, N = 64 or 32)
    int use_side = 1;      // run k_nyq_accum on an internal side stream (overlaps the pencil kernel)
    // Staged x stage and Nyquist accumulate: CTA row gy owns share gy of a chunk's pairs and, in general,
    // partial slot gy of S.  When every share of every launch starts at a radius boundary no two rows
    // touch the same (radius, tile), so one slot per kernel is enough (decided by update_slot_layout).
    int one_slot_pencil = 0;
    // Register-resident x stage: the pair list is cut into work units (<= seg_pairs pairs of one radius
    // inside one launch); unit k of a radius owns partial slot k of S, written once with plain stores.
    int seg_pairs = 0, unit_slots = 0, uniform_w = 0;
    PencilUnit *units = nullptr;  // device copy of h_units
    int *slots_of_r = nullptr;    // [n_r_local] partial slots radius r uses
    std::vector<PencilUnit> h_units;
    std::vector<int> chunk_unit_first; // [n_chunks + 1] first unit of every launch
    std::vector<double> h_pair_w;
    // fused persistent gain kernel (64^3 packed mode): roles, sub-chunk size, ring depth
    int fused = 0, fused_K = 0, fused_D = 0, fused_NQ = 0, fused_NN = 0;
    int *sync_flags = nullptr;    // [2 * n_sub] ready / consumed counters of the fused kernel
    int cluster = 0;              // 32^3 packed mode: cluster/DSMEM gain kernel (no hybrid scratch)

    int S_slots_capacity = 0;     // partial slots S was allocated for
    int chunk_capacity = 0;       // pairs the per-chunk scratch (hyb, uvw) was allocated for
    bfsm_plan_options opt;
    // general path (non-cubic / not 16-32-64 grids): per-axis sizes and twiddle tables
    int general = 0, nx = 0, ny = 0, nz = 0;
    cplx *gen_tw[3] = {nullptr, nullptr, nullptr};
    int n_dir = 0, pair_lo = 0; // pairs per radius; global index of this shard's first pair
    cudaStream_t side = nullptr;
    cudaEvent_t ev_plane[2] = {nullptr, nullptr}, ev_nyq[2] = {nullptr, nullptr};
    int pairs_total = 0, pairs_local = 0;
    int n_r_local = 0;
    int M = 0; // |l|^2 table length
    int chunk = 0;
    int gy = 1;       // persistent CTAs of k_plane_gain
    int G = 1;        // cross-CTA pair groups in k_pencil_gain (= number of S partials)
    int sm_count = 148;
    int shard_index = 0, shard_count = 1;
    long long scratch_bytes = 0;

    // device tables
    cplx *tw = nullptr;       // [N] exp(+2 pi i t/N)
    cplx *phase = nullptr;    // [pairs_local][3][N]
    cplx *zpm = nullptr;      // [pairs_local][N]  (Re+Im, Re-Im) of the z phase (k_plane_gain_ws)
    cplx *zpm_r32 = nullptr;  // same with the entries k = 3 mod 4 negated at N = 64 (k_plane_gain_r32)
    int *pair_r = nullptr;    // [pairs_local] local radius index
    double *pair_w = nullptr; // [pairs_local] spherical weight (x2 when folded)
    int *r_end = nullptr;     // [n_r_local] one past the last local pair of that radius
    double *coef = nullptr;   // [n_r_local][M]
    double *beta2 = nullptr;  // [M]
    // device scratch
    cplx *fhat = nullptr;  // [N^3]
    cplx *tmp = nullptr;   // [max(2, n_r_local)][N^3]  hybrid scratch of the single-shot stages
    cplx *hyb = nullptr;   // [(packed ? 1 : 2)*chunk][N^3]
    CUtensorMap hyb_tmap;  // hyb as the 3-D tensor [pair*N + x][y][2 z] of doubles (TMA-filled x stage)
    double *S = nullptr;   // [x-stage slots + Nyquist slots][n_r_local][N^3]
    cplx *nyq = nullptr;   // [3][N][N] Nyquist planes of fhat (packed mode)
    cplx *uvw = nullptr;   // [2][chunk][3][N][N] (packed mode, double buffered)
    cplx *qhat = nullptr;  // [N^3]
    double *stage_f = nullptr, *stage_q = nullptr; // host-pointer entry point staging
    size_t stage_cells = 0;
    // pipelined host-pointer entry point (bfsm_collide_host_async): BFSM_HOST_PIPE_DEPTH staging slots,
    // copies on their own streams so that the H2D of later steps and the D2H of earlier ones run under
    // the kernels of step k.  (Two slots were enough on one GPU; with a collective in every step the
    // host-side jitter of ANY rank stalls all of them, and a deeper queue absorbs it.)
    struct HostPipe {
        static constexpr int DEPTH = BFSM_HOST_PIPE_DEPTH;
        double *f[DEPTH] = {}, *q[DEPTH] = {};
        size_t cells = 0;
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t in_done[DEPTH] = {}, comp_done[DEPTH] = {}, out_done[DEPTH] = {};
        bool used[DEPTH] = {};
        unsigned long long submitted = 0;
    } pipe;
    std::vector<void *> allocs;

    // Batch mode (n_cells > 1): up to n_lanes cells are kept in flight on "lanes", each with its own
    // per-evaluation scratch and stream pair, so that one cell's small single-shot kernels and
    // kernel tails overlap the other cell's gain kernels.  Lane 0 is the scratch above (and is
    // what single-cell calls use, on the caller's stream); lanes 1.. are allocated on first use.
    // Measured at 32^3 / 16 x 94 (cfg 5): 2653 (1 lane), 3384 (2), 3768 (3), 3897 (4) cells/s.
    struct Lane {
        cplx *fhat = nullptr, *tmp = nullptr, *hyb = nullptr, *nyq = nullptr, *uvw = nullptr,
             *qhat = nullptr;
        CUtensorMap hyb_tmap;
        double *S = nullptr;
        int *sync_flags = nullptr;
        cudaStream_t main = nullptr, side = nullptr;
        cudaEvent_t ev_plane[2] = {nullptr, nullptr}, ev_nyq[2] = {nullptr, nullptr}, done = nullptr;
        bool allocated = false;
    };
    static constexpr int MAX_LANES = 4;
    Lane lanes[MAX_LANES];
    // Batch mode, cell groups (32^3-class grids with the radix-32 plane kernel and the register x stage):
    // every kernel of the path takes a cell dimension, so ONE launch serves up to `cells` cells -- the
    // launch ramps, kernel tails and the latency-bound single-shot kernels are paid once per group
    // instead of once per cell.  Two groups alternate (own scratch and streams) so that one group's
    // tail overlaps the other group's head.
    struct Group {
        cplx *fhat = nullptr, *nyq = nullptr, *tmp = nullptr, *qhat = nullptr, *hyb = nullptr, *uvw = nullptr;
        double *S = nullptr;
        cudaStream_t main = nullptr, side = nullptr;
        cudaEvent_t ev_plane[2] = {nullptr, nullptr}, ev_nyq[2] = {nullptr, nullptr}, done = nullptr;
        int cells = 0;       // capacity in cells
        int slots = 0;       // partial slots per cell S was sized for
        bool allocated = false;
    };
    static constexpr int MAX_GROUPS = 2, GROUP_CELLS = 8;
    Group groups[MAX_GROUPS];
    int group_cells_last = 0; // cells per launch the last batch call used (0 = lane path)
    cudaEvent_t ev_fork = nullptr;
    int n_lanes = 4; // cells kept in flight in batch mode (1 = strictly sequential)
    int lane_fail_from = 0; // test hook: allocation of lanes >= this index fails (0 = off)
    int lanes_used_last = 1; // lanes the last batch call actually ran on

    // optional per-kernel-class timing (bfsm_collide_profiled)
    bool profiling = false, profile_failed = false;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    size_t events_used = 0;
};

namespace {

size_t grid_points(const bfsm_plan *p) { return (size_t)p->nx * p->ny * p->nz; }

int dev_alloc(bfsm_plan *p, void **out, size_t bytes)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        char b[256];
        snprintf(b, sizeof b, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return fail(e == cudaErrorMemoryAllocation ? BFSM_ERR_NOMEM : BFSM_ERR_CUDA, b);
    }
    p->allocs.push_back(q);
    p->scratch_bytes += (long long)bytes;
    *out = q;
    return BFSM_OK;
}

template <class T> int upload(bfsm_plan *p, T **out, const std::vector<T> &h)
{
    int rc = dev_alloc(p, (void **)out, h.size() * sizeof(T));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return BFSM_OK;
}

// ---- per-N launch geometry ------------------------------------------------------------
template <int N> struct Launch;
template <> struct Launch<64> { static constexpr int TG = 256, GROUPS = 2, MINB = 1, PG = 4, PMINB = 2, G = 1, CHUNK = 384, PSTAGES = 3; };
template <> struct Launch<32> { static constexpr int TG = 128, GROUPS = 2, MINB = 2, PG = 8, PMINB = 2, G = 4, CHUNK = 1024, PSTAGES = 3; };
template <> struct Launch<16> { static constexpr int TG = 64, GROUPS = 2, MINB = 4, PG = 8, PMINB = 2, G = 8, CHUNK = 1024, PSTAGES = 3; };

template <int N> size_t plane_gain_smem()
{
    using Lc = Launch<N>;
    return sizeof(cplx) * ((size_t)N * N + (size_t)Lc::GROUPS * N * Geo<N>::ROW +
                           (size_t)Lc::GROUPS * 2 * 3 * N);
}
template <int N> size_t plane_smem() { return sizeof(cplx) * (size_t)N * Geo<N>::ROW; }
template <int N> size_t pencil_gain_smem()
{
    return sizeof(cplx) * (size_t)Launch<N>::PG * 2 * N * TZ;
}

template <int N> size_t plane_gain3_smem()
{
    return sizeof(cplx) * ((size_t)(1 + Launch<N>::GROUPS) * N * (N + 1) +
                           (size_t)Launch<N>::GROUPS * 2 * 3 * N);
}

// radix-32 plane kernel: groups per CTA and CTAs per SM, with the fhat line in registers (TM = false,
// 8 warps per SM at 255 registers) or in tensor memory (TM = true, 11-12 warps per SM at 168 registers)
template <int N, bool TM> struct R32Launch;
template <> struct R32Launch<64, false> { static constexpr int G = 1, MINB = 2; };
template <> struct R32Launch<64, true> { static constexpr int G = 1, MINB = 3; };
template <> struct R32Launch<32, false> { static constexpr int G = 8, MINB = 1; };
template <> struct R32Launch<32, true> { static constexpr int G = 11, MINB = 1; };

template <int N, bool TM> int configure_plane_r32()
{
    using Rl = R32Launch<N, TM>;
    CUDA_TRY(cudaFuncSetAttribute(k_plane_gain_r32<N, Rl::G, Rl::MINB, TM>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)plane_r32_smem<N, Rl::G>()));
    return BFSM_OK;
}

template <int N> size_t plane_ws_smem()
{
    return sizeof(cplx) * ((size_t)3 * N * (N + 1) + (size_t)2 * 4 * N + (size_t)N);
}

template <int N> size_t pencil_async_smem()
{
    return sizeof(cplx) * (size_t)Launch<N>::PG * Launch<N>::PSTAGES * N * TZ;
}

// register-resident x stage: 8 warps per CTA, two CTAs per SM (128 registers)
constexpr int PR_WARPS = 8, PR_MINB = 2;

template <int N> int configure_kernels()
{
    using Lc = Launch<N>;
    CUDA_TRY(cudaFuncSetAttribute(k_plane_gain3<N, Lc::GROUPS, Lc::MINB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)plane_gain3_smem<N>()));
    if constexpr (N == 64) {
        CUDA_TRY(cudaFuncSetAttribute(k_plane_gain_ws<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)plane_ws_smem<N>()));
        CUDA_TRY(cudaFuncSetAttribute(k_gain_fused<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)plane_ws_smem<N>()));
        CUDA_TRY(cudaFuncSetAttribute(k_gain_fused<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)plane_ws_smem<N>()));
    }
    if constexpr (N == 64 || N == 32) {
        if (int rc = configure_plane_r32<N, false>()) return rc;
        if (int rc = configure_plane_r32<N, true>()) return rc;
    }
    if constexpr (N == 32) {
        CUDA_TRY(cudaFuncSetAttribute(k_gain_cluster<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cluster_smem_bytes<N>()));
        CUDA_TRY(cudaFuncSetAttribute(k_gain_cluster<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cluster_smem_bytes<N>()));
    }
    CUDA_TRY(cudaFuncSetAttribute(k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pencil_async_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pencil_async_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pencil_async_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pencil_async_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_plane_gain<N, Lc::TG, Lc::GROUPS, Lc::MINB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)plane_gain_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_pencil_gain<N, Lc::PG, Lc::PMINB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pencil_gain_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_plane<N, -1, PLANE_REAL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem<N>()));
    CUDA_TRY(cudaFuncSetAttribute(k_plane<N, +1, PLANE_FINAL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem<N>()));
    return BFSM_OK;
}

// ---- optional CUDA-event bracket around one launch (same stream as the kernel) -----------
// Profiling must never change the result of a call: an event that cannot be created or recorded
// switches the profile off for this evaluation (bfsm_collide_profiled then reports the failure).
cudaEvent_t prof_event(bfsm_plan *p)
{
    if (p->events_used == p->event_pool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) {
            p->profile_failed = true;
            return nullptr;
        }
        p->event_pool.push_back(e);
    }
    return p->event_pool[p->events_used++];
}
struct ProfSpan {
    bfsm_plan *p;
    cudaStream_t st;
    int cls;
    cudaEvent_t a = nullptr;
    ProfSpan(bfsm_plan *p_, cudaStream_t st_, int cls_) : p(p_), st(st_), cls(cls_)
    {
        if (p->profiling) {
            a = prof_event(p);
            if (!a || cudaEventRecord(a, st) != cudaSuccess) p->profile_failed = true;
        }
    }
    ~ProfSpan()
    {
        if (p->profiling) {
            cudaEvent_t b = prof_event(p);
            if (!a || !b || cudaEventRecord(b, st) != cudaSuccess) p->profile_failed = true;
            else p->spans.push_back({cls, a, b});
        }
    }
};

// ---- partial-slot layout ---------------------------------------------------------------------
// true if, for every launch of `chunk` pairs, the `groups` shares [nc*g/G, nc*(g+1)/G) all start at a
// radius boundary of the r-major pair list (pair = r*n_dir + d)
bool shares_start_at_radius_boundaries(int pairs_local, int pair_lo, int n_dir, int chunk, int groups)
{
    if (n_dir <= 0 || chunk <= 0) return false;
    for (int c0 = 0; c0 < pairs_local; c0 += chunk) {
        const int nc = std::min(chunk, pairs_local - c0);
        const int G = std::min(groups, nc);
        for (int g = 1; g < G; ++g) {
            const int b = (int)(((long long)nc * g) / G);
            if ((pair_lo + c0 + b) % n_dir != 0) return false;
        }
    }
    return true;
}
// Cuts a shard's pair list (pairs_local pairs starting at global pair pair_lo, n_dir pairs per radius,
// r-major) into the work units of the register-resident x stage: every launch (chunk) is cut at radius
// boundaries, every radius segment into pieces of at most seg_pairs pairs; piece k of a radius
// (counted over the whole shard) owns partial slot k.  Returns the slot count.
int cut_units(int pairs_local, int pair_lo, int n_dir, int chunk, int seg_pairs,
              std::vector<PencilUnit> &units, std::vector<int> &chunk_first, std::vector<int> &slots_of_r)
{
    units.clear();
    chunk_first.clear();
    const int r_first = (pairs_local > 0) ? pair_lo / n_dir : 0;
    const int r_last = (pairs_local > 0) ? (pair_lo + pairs_local - 1) / n_dir : -1;
    slots_of_r.assign(std::max(1, r_last - r_first + 1), 0);
    for (int c0 = 0; c0 < pairs_local; c0 += chunk) {
        chunk_first.push_back((int)units.size());
        const int c1 = std::min(pairs_local, c0 + chunk);
        int q = c0;
        while (q < c1) {
            const int r = (pair_lo + q) / n_dir - r_first;            // local radius index
            const int r_stop = std::min(c1, (r_first + r + 1) * n_dir - pair_lo);
            const int len = r_stop - q, pieces = (len + seg_pairs - 1) / seg_pairs;
            for (int k = 0; k < pieces; ++k) {
                PencilUnit u;
                u.p0 = q + (int)(((long long)len * k) / pieces);
                u.p1 = q + (int)(((long long)len * (k + 1)) / pieces);
                u.r = r;
                u.slot = slots_of_r[r]++;
                units.push_back(u);
            }
            q = r_stop;
        }
    }
    chunk_first.push_back((int)units.size());
    int slots = 0;
    for (int v : slots_of_r) slots = std::max(slots, v);
    return slots;
}
int build_units(bfsm_plan *p, std::vector<int> &slots_of_r)
{
    // fused kernel: one launch over all pairs, units = the radius pieces of every sub-chunk
    if (p->fused)
        return cut_units(p->pairs_local, p->pair_lo, p->n_dir, p->fused_K, p->fused_K, p->h_units,
                         p->chunk_unit_first, slots_of_r);
    return cut_units(p->pairs_local, p->pair_lo, p->n_dir, p->chunk, p->seg_pairs, p->h_units,
                     p->chunk_unit_first, slots_of_r);
}

void update_slot_layout(bfsm_plan *p)
{
    const bool staged = p->packed && (p->pencil_kernel == 1 || p->pencil_kernel == 3) && !p->fused && !p->cluster;
    auto aligned = [&](int groups) {
        return shares_start_at_radius_boundaries(p->pairs_local, p->pair_lo, p->n_dir, p->chunk, groups);
    };
    p->one_slot_pencil = (staged && p->G > 1 && aligned(p->G)) ? 1 : 0;
}
bool pencil_units_active(const bfsm_plan *p)
{
    return p->packed && (p->pencil_kernel == 2 || p->fused || p->cluster);
}
bool units_needed(const bfsm_plan *p) { return p->packed != 0; } // the Nyquist accumulate always walks units
// hybrid grids / Nyquist-field sets the per-launch scratch holds
size_t hyb_grids(const bfsm_plan *p)
{
    if (p->fused) return (size_t)p->fused_D * p->fused_K;
    if (p->cluster) return 0; // the hybrid grids live in the clusters' shared memory
    return (size_t)(p->packed ? 1 : 2) * p->chunk_capacity;
}
size_t uvw_sets(const bfsm_plan *p) { return p->fused ? (size_t)std::max(1, p->pairs_local) : (size_t)2 * p->chunk_capacity; }
int fused_subs(const bfsm_plan *p) { return (p->pairs_local + p->fused_K - 1) / std::max(1, p->fused_K); }
int pencil_slots(const bfsm_plan *p)
{
    if (pencil_units_active(p)) return p->unit_slots;
    return p->one_slot_pencil ? 1 : p->G;
}
int nyq_slots(const bfsm_plan *p) { return p->packed ? p->unit_slots : 0; } // one per work unit of a radius

// (re)builds everything that depends on the chunk size: unit table, slot layout, S capacity
int relayout(bfsm_plan *p)
{
    const size_t N3 = grid_points(p);
    if (units_needed(p)) {
        std::vector<int> slots_of_r;
        p->unit_slots = build_units(p, slots_of_r);
        if (p->units) { cudaFree(p->units); p->units = nullptr; }
        if (p->slots_of_r) { cudaFree(p->slots_of_r); p->slots_of_r = nullptr; }
        const size_t nu = std::max<size_t>(1, p->h_units.size());
        CUDA_TRY(cudaMalloc((void **)&p->units, nu * sizeof(PencilUnit)));
        CUDA_TRY(cudaMalloc((void **)&p->slots_of_r, slots_of_r.size() * sizeof(int)));
        if (!p->h_units.empty())
            CUDA_TRY(cudaMemcpy(p->units, p->h_units.data(), p->h_units.size() * sizeof(PencilUnit),
                                cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->slots_of_r, slots_of_r.data(), slots_of_r.size() * sizeof(int),
                            cudaMemcpyHostToDevice));
    }
    update_slot_layout(p);
    const int need = std::max(1, pencil_slots(p) + nyq_slots(p));
    if (need > p->S_slots_capacity) {
        if (p->S) {
            cudaFree(p->S);
            p->scratch_bytes -= (long long)sizeof(double) * N3 * p->S_slots_capacity * std::max(1, p->n_r_local);
            p->S = nullptr;
        }
        const size_t bytes = sizeof(double) * N3 * (size_t)need * std::max(1, p->n_r_local);
        cudaError_t e = cudaMalloc((void **)&p->S, bytes);
        if (e != cudaSuccess) {
            p->S_slots_capacity = 0;
            return fail(e == cudaErrorMemoryAllocation ? BFSM_ERR_NOMEM : BFSM_ERR_CUDA,
                        "cudaMalloc of the partial-sum slots failed");
        }
        p->scratch_bytes += (long long)bytes;
        p->S_slots_capacity = need;
    }
    return BFSM_OK;
}

// ---- one evaluation, split in the two halves the multi-GPU path needs -------------------

// f -> fhat (scaled by 1/N^3) -> partial gain spectrum of this shard
template <int N> int run_gain_hat(bfsm_plan *p, cplx *qhat_out, const double *f, cudaStream_t st)
{
    using Lc = Launch<N>;
    constexpr size_t N3 = (size_t)N * N * N;
    constexpr int TILES = N * N / TZ;
    constexpr int TGP = Geo<N>::B * TZ;

    // forward transform of f (cpp:168-186)
    {
        ProfSpan ps(p, st, BFSM_KCLASS_FORWARD);
        k_plane<N, -1, PLANE_REAL><<<dim3(N, 1), N * Geo<N>::B, plane_smem<N>(), st>>>(
            f, 1, 0, nullptr, nullptr, nullptr, p->tw, p->tmp);
        // (packed mode: the kernel also files the three Nyquist planes of fhat into p->nyq)
        k_pencil_fwd<N><<<TILES, TGP, 0, st>>>(p->tmp, p->tw, 1.0 / (double)N3, p->fhat,
                                               p->packed ? p->nyq : nullptr);
    }

    // gain: S_r = sum_sigma w Re(g1 g2), kept as partial slots that the next stage sums in fixed order.
    // The register-resident x stage writes each of its slots exactly once (plain stores); the other
    // accumulating kernels add into theirs, which are cleared first.
    const bool units = pencil_units_active(p);
    const int pslots = pencil_slots(p), nslots = nyq_slots(p);
    const size_t slot_stride = (size_t)p->n_r_local * N3;
    double *S2 = p->S + (size_t)pslots * slot_stride; // Nyquist slots (written once per unit, never cleared)
    if (!units) CUDA_TRY(cudaMemsetAsync(p->S, 0, sizeof(double) * (size_t)pslots * slot_stride, st));
    if (p->fused) {
        if constexpr (N == 64) {
            const int n_sub = fused_subs(p);
            CUDA_TRY(cudaMemsetAsync(p->sync_flags, 0, sizeof(int) * 2 * (size_t)n_sub, st));
            FusedParams fp;
            fp.n_nyq = p->fused_NN;
            fp.n_pencil = p->fused_NQ;
            fp.n_plane = p->sm_count - fp.n_nyq - fp.n_pencil;
            fp.pair0 = 0;
            fp.n_pairs = p->pairs_local;
            fp.n_units = (int)p->h_units.size();
            fp.rs.ready = p->sync_flags;
            fp.rs.consumed = p->sync_flags + n_sub;
            fp.rs.n_sub = n_sub;
            fp.rs.ring = p->fused_D;
            fp.rs.sub_pairs = p->fused_K;
            fp.rs.n_pairs = p->pairs_local;
            fp.rs.tiles = PencilGeo<N>::WT;
            const cplx *fhat = p->fhat, *phase = p->phase, *zpm = p->zpm, *tw = p->tw, *nyqp = p->nyq;
            cplx *ring = p->hyb, *uvw = p->uvw;
            const double *pw = p->pair_w;
            const PencilUnit *un = p->units;
            double *Sp = p->S;
            int nrl = p->n_r_local;
            void *args[] = {&fp, &fhat, &phase, &zpm, &tw, &ring, &nyqp, &pw, &uvw, &un, &Sp, &nrl};
            {
                ProfSpan ps(p, st, BFSM_KCLASS_PLANE_GAIN);
                const void *fn = p->uniform_w ? (const void *)k_gain_fused<N, true> : (const void *)k_gain_fused<N, false>;
                CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(fp.n_plane + fp.n_nyq + fp.n_pencil),
                                                     dim3(FUSED_THREADS), args, plane_ws_smem<N>(), st));
            }
            {
                ProfSpan ps(p, st, BFSM_KCLASS_NYQUIST);
                constexpr int NYQ_TILES = (N / 16) * (N / 16) * (N / 16);
                k_nyq_accum<N><<<dim3(NYQ_TILES, (int)p->h_units.size()), 256, 0, st>>>(
                    p->uvw, 0, p->units, S2, p->n_r_local);
            }
        }
    }
    int ci = 0;
    bool nyq_pending[2] = {false, false};
    const bool side = p->packed && p->use_side && p->side;
    for (int c0 = 0; c0 < p->pairs_local && !p->fused; c0 += p->chunk, ++ci) {
        const int nc = std::min(p->chunk, p->pairs_local - c0);
        const int items = p->packed ? nc : 2 * nc;
        const int ub = ci & 1; // uvw buffer of this chunk
        cplx *uvw = p->packed ? p->uvw + (size_t)ub * 3 * N * N * p->chunk_capacity : nullptr;
        if (side && nyq_pending[ub]) { // k_nyq_accum of chunk ci-2 must be done with uvw[ub]
            CUDA_TRY(cudaStreamWaitEvent(st, p->ev_nyq[ub], 0));
            nyq_pending[ub] = false;
        }
        // persistent: one CTA per SM slot, never more CTAs than (plane, item-pair) units
        int ctas = std::min(p->gy, std::max(1, (N * items) / Lc::GROUPS));
        {
            ProfSpan ps(p, st, BFSM_KCLASS_PLANE_GAIN);
            if (p->packed && p->plane_r32 && !p->cluster) {
                if constexpr (N == 64 || N == 32) {
                    auto launch = [&](auto tm) {
                        constexpr bool TM = decltype(tm)::value;
                        using Rl = R32Launch<N, TM>;
                        const int groups = (N + 3) * items; // never more groups than list entries
                        const int grid = std::max(1, std::min(p->sm_count * Rl::MINB, (groups + Rl::G - 1) / Rl::G));
                        k_plane_gain_r32<N, Rl::G, Rl::MINB, TM>
                            <<<grid, Rl::G * R32Geo<N>::GT, plane_r32_smem<N, Rl::G>(), st>>>(
                                p->fhat, p->phase, p->zpm_r32, p->hyb, c0, items, p->nyq, p->pair_w, uvw);
                    };
                    if (p->plane_r32 == 2) launch(std::true_type{});
                    else launch(std::false_type{});
                }
            } else if (N == 64 && p->packed && p->plane_ws) {
                if constexpr (N == 64) {
                    const int grid = std::min(p->sm_count, (N + 3) * items);
                    k_plane_gain_ws<N><<<grid, 384, plane_ws_smem<N>(), st>>>(
                        p->fhat, p->phase, p->zpm, p->tw, p->hyb, c0, items, p->nyq, p->pair_w, uvw);
                }
            } else if (p->packed)
                k_plane_gain3<N, Lc::GROUPS, Lc::MINB>
                    <<<ctas, 4 * N * Lc::GROUPS, plane_gain3_smem<N>(), st>>>(
                        p->fhat, p->phase, p->tw, p->hyb, c0, items, p->nyq, p->pair_w, uvw,
                        p->cluster ? 2 : 3); // cluster mode: only the Nyquist planes
            else
                k_plane_gain<N, Lc::TG, Lc::GROUPS, Lc::MINB>
                    <<<ctas, Lc::TG * Lc::GROUPS, plane_gain_smem<N>(), st>>>(
                        p->fhat, p->phase, p->tw, p->hyb, c0, items);
        }
        if (p->packed) {
            // exact correction for the Nyquist planes: S2_r += sum_s Re(Y_s^2).  FP64-only work,
            // issued on the side stream so that it overlaps the memory-bound pencil kernel.
            cudaStream_t ns = side ? p->side : st;
            if (side) {
                CUDA_TRY(cudaEventRecord(p->ev_plane[ub], st));
                CUDA_TRY(cudaStreamWaitEvent(ns, p->ev_plane[ub], 0));
            }
            {
                ProfSpan ps(p, ns, BFSM_KCLASS_NYQUIST);
                constexpr int NYQ_TILES = (N / 16) * (N / 16) * (N / 16);
                const int u0 = p->chunk_unit_first[ci], nu = p->chunk_unit_first[ci + 1] - u0;
                k_nyq_accum<N><<<dim3(NYQ_TILES, nu), 256, 0, ns>>>(uvw, c0, p->units + u0, S2, p->n_r_local);
            }
            if (side) {
                CUDA_TRY(cudaEventRecord(p->ev_nyq[ub], ns));
                nyq_pending[ub] = true;
            }
        }
        const int G = std::min(p->G, nc);
        {
            ProfSpan ps(p, st, BFSM_KCLASS_PENCIL_GAIN);
            if (p->cluster) {
                if constexpr (N == 32) {
                    const int u0 = p->chunk_unit_first[ci], nu = p->chunk_unit_first[ci + 1] - u0;
                    const int n_clusters = std::max(1, std::min(nu, p->sm_count / CL_SIZE));
                    if (p->uniform_w)
                        k_gain_cluster<N, true><<<CL_SIZE * n_clusters, CL_THREADS, cluster_smem_bytes<N>(), st>>>(
                            p->fhat, p->phase, p->tw, p->units + u0, nu, p->pair_w, p->S, p->n_r_local);
                    else
                        k_gain_cluster<N, false><<<CL_SIZE * n_clusters, CL_THREADS, cluster_smem_bytes<N>(), st>>>(
                            p->fhat, p->phase, p->tw, p->units + u0, nu, p->pair_w, p->S, p->n_r_local);
                }
            } else if (units) {
                const int u0 = p->chunk_unit_first[ci], nu = p->chunk_unit_first[ci + 1] - u0;
                const dim3 grid(PencilGeo<N>::WT / PR_WARPS, nu);
                if (p->uniform_w)
                    k_pencil_gain_reg<N, true, PR_WARPS, PR_MINB><<<grid, PR_WARPS * 32, 0, st>>>(
                        p->hyb, c0, p->units + u0, p->pair_w, p->S, p->n_r_local);
                else
                    k_pencil_gain_reg<N, false, PR_WARPS, PR_MINB><<<grid, PR_WARPS * 32, 0, st>>>(
                        p->hyb, c0, p->units + u0, p->pair_w, p->S, p->n_r_local);
            } else if (p->packed && p->pencil_kernel == 3 && p->one_slot_pencil)
                k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, true, true>
                    <<<dim3(TILES, G), Lc::PG * TGP, pencil_async_smem<N>(), st>>>(
                        p->hyb, p->tw, p->pair_r, p->pair_w, p->r_end, p->S, c0, nc, p->n_r_local, p->hyb_tmap);
            else if (p->packed && p->pencil_kernel == 3)
                k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, false, true>
                    <<<dim3(TILES, G), Lc::PG * TGP, pencil_async_smem<N>(), st>>>(
                        p->hyb, p->tw, p->pair_r, p->pair_w, p->r_end, p->S, c0, nc, p->n_r_local, p->hyb_tmap);
            else if (p->packed && p->one_slot_pencil)
                k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB, true>
                    <<<dim3(TILES, G), Lc::PG * TGP, pencil_async_smem<N>(), st>>>(
                        p->hyb, p->tw, p->pair_r, p->pair_w, p->r_end, p->S, c0, nc, p->n_r_local, p->hyb_tmap);
            else if (p->packed)
                k_pencil_gain_async<N, Lc::PG, Lc::PSTAGES, Lc::PMINB>
                    <<<dim3(TILES, G), Lc::PG * TGP, pencil_async_smem<N>(), st>>>(
                        p->hyb, p->tw, p->pair_r, p->pair_w, p->r_end, p->S, c0, nc, p->n_r_local, p->hyb_tmap);
            else
                k_pencil_gain<N, Lc::PG, Lc::PMINB>
                    <<<dim3(TILES, G), Lc::PG * TGP, pencil_gain_smem<N>(), st>>>(
                        p->hyb, p->tw, p->pair_r, p->pair_w, p->r_end, p->S, c0, nc, p->n_r_local);
        }
    }
    for (int ub = 0; ub < 2; ++ub)
        if (side && nyq_pending[ub]) CUDA_TRY(cudaStreamWaitEvent(st, p->ev_nyq[ub], 0));

    // Qhat = sum_r coef_r(|l|^2) FFT3(S_r)   (cpp:249-273)
    {
        ProfSpan ps(p, st, BFSM_KCLASS_ACCUM);
        if (p->n_r_local > 0) {
            k_plane<N, -1, PLANE_REAL><<<dim3(N, p->n_r_local), N * Geo<N>::B, plane_smem<N>(), st>>>(
                p->S, pslots, slot_stride, nullptr, nullptr, nullptr, p->tw, p->tmp,
                units ? p->slots_of_r : nullptr, S2, nslots, p->packed ? p->slots_of_r : nullptr);
        }
        k_pencil_accum<N><<<TILES, AccumGeo<N>::RS * TGP, 0, st>>>(p->tmp, p->tw, p->coef, p->n_r_local, p->M, qhat_out);
    }
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

// loss term + inverse transform + combine (cpp:281-330); needs p->fhat of the same f.
// with_loss = false: Q = Re(IFFT3(qhat)) only (a pair shard's partial gain, see bfsm_collide_sharded).
template <int N>
int run_finish(bfsm_plan *p, double *Q, const cplx *qhat, const double *f, cudaStream_t st, bool with_loss = true)
{
    constexpr int TILES = N * N / TZ;
    constexpr int TGP = Geo<N>::B * TZ;
    {
        ProfSpan ps(p, st, BFSM_KCLASS_FINAL);
        k_plane<N, +1, PLANE_FINAL><<<dim3(N, with_loss ? 2 : 1), N * Geo<N>::B, plane_smem<N>(), st>>>(
            nullptr, 0, 0, qhat, p->fhat, p->beta2, p->tw, p->tmp);
        if (with_loss) k_pencil_final<N, true><<<TILES, TGP, 0, st>>>(p->tmp, p->tw, f, Q);
        else k_pencil_final<N, false><<<TILES, TGP, 0, st>>>(p->tmp, p->tw, f, Q);
    }
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

template <int N> int launches_per_cell(const bfsm_plan *p)
{
    if (p->fused) return 2 + 1 + 1 + (p->n_r_local > 0 ? 1 : 0) + 1 + 2; // fwd, fused, nyq accum, accum, final
    const int chunks = (p->pairs_local + p->chunk - 1) / p->chunk;
    return 2 + (p->packed ? 3 * chunks : 2 * chunks) + (p->n_r_local > 0 ? 1 : 0) + 1 + 2;
}

// ---- batch mode: groups of cells per launch ------------------------------------------------------
bool group_path_ok(const bfsm_plan *p)
{
    return !p->general && p->packed && p->plane_r32 && !p->fused && !p->cluster && p->pencil_kernel == 2 &&
           p->N == 32 && p->shard_count == 1 && p->pairs_local > 0;
}
void group_free(bfsm_plan *p, int k)
{
    bfsm_plan::Group &G = p->groups[k];
    void *bufs[] = {G.fhat, G.nyq, G.tmp, G.qhat, G.hyb, G.uvw, G.S};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (G.main) cudaStreamDestroy(G.main);
    if (G.side) cudaStreamDestroy(G.side);
    for (int j = 0; j < 2; ++j) {
        if (G.ev_plane[j]) cudaEventDestroy(G.ev_plane[j]);
        if (G.ev_nyq[j]) cudaEventDestroy(G.ev_nyq[j]);
    }
    if (G.done) cudaEventDestroy(G.done);
    G = bfsm_plan::Group();
}
void groups_free(bfsm_plan *p)
{
    for (int k = 0; k < bfsm_plan::MAX_GROUPS; ++k) group_free(p, k);
}
// scratch of group k for `cells` cells per launch; BFSM_ERR_NOMEM leaves the group unallocated
int group_alloc(bfsm_plan *p, int k, int cells)
{
    bfsm_plan::Group &G = p->groups[k];
    const int slots = std::max(1, pencil_slots(p) + nyq_slots(p));
    if (G.allocated && G.cells >= cells && G.slots >= slots) return BFSM_OK;
    group_free(p, k);
    const size_t N = (size_t)p->N, N3 = N * N * N, C = (size_t)cells;
    const size_t nr = (size_t)std::max(1, p->n_r_local);
    auto need = [&](void **q, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(q, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            *q = nullptr;
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? BFSM_ERR_NOMEM : BFSM_ERR_CUDA,
                        "cudaMalloc for a batch cell group failed");
        }
        return BFSM_OK;
    };
    int rc = BFSM_OK;
    if ((rc = need((void **)&G.fhat, sizeof(cplx) * N3 * C)) || (rc = need((void **)&G.qhat, sizeof(cplx) * N3 * C)) ||
        (rc = need((void **)&G.nyq, sizeof(cplx) * 3 * N * N * C)) ||
        (rc = need((void **)&G.tmp, sizeof(cplx) * N3 * std::max<size_t>(2, nr) * C)) ||
        (rc = need((void **)&G.hyb, sizeof(cplx) * N3 * (size_t)p->chunk * C)) ||
        (rc = need((void **)&G.uvw, sizeof(cplx) * 3 * N * N * 2 * (size_t)p->chunk * C)) ||
        (rc = need((void **)&G.S, sizeof(double) * N3 * (size_t)slots * nr * C))) {
        std::string keep = g_err;
        group_free(p, k);
        g_err = keep;
        return rc;
    }
    bool ok = cudaStreamCreateWithFlags(&G.main, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&G.side, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&G.done, cudaEventDisableTiming) == cudaSuccess;
    for (int j = 0; j < 2 && ok; ++j)
        ok = cudaEventCreateWithFlags(&G.ev_plane[j], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&G.ev_nyq[j], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        group_free(p, k);
        return fail(BFSM_ERR_CUDA, "stream / event creation for a batch cell group failed");
    }
    G.cells = cells;
    G.slots = slots;
    G.allocated = true;
    return BFSM_OK;
}

// One evaluation of `cg` consecutive cells (f, Q: cg x N^3) with every kernel launched once per group
// (run_gain_hat + run_finish with a cell dimension; the per-cell arithmetic and summation order are
// those of the single-cell path, so the results are bitwise the same).
template <int N> int run_group(bfsm_plan *p, bfsm_plan::Group &G, double *Q, const double *f, int cg)
{
    if constexpr (N != 32) {
        return fail(BFSM_ERR_UNSUPPORTED, "cell groups are built for 32^3 grids");
    } else {
        constexpr size_t N3 = (size_t)N * N * N;
        constexpr int TILES = N * N / TZ;
        constexpr int TGP = Geo<N>::B * TZ;
        cudaStream_t st = G.main;
        const int nr = p->n_r_local;
        // forward transforms (cpp:168-186)
        k_plane<N, -1, PLANE_REAL><<<dim3(N, 1, cg), N * Geo<N>::B, plane_smem<N>(), st>>>(
            f, 1, 0, nullptr, nullptr, nullptr, p->tw, G.tmp, nullptr, nullptr, 0, nullptr, N3, N3);
        k_pencil_fwd<N><<<dim3(TILES, cg), TGP, 0, st>>>(G.tmp, p->tw, 1.0 / (double)N3, G.fhat, G.nyq);

        const int pslots = pencil_slots(p), nslots = nyq_slots(p);
        const size_t slot_stride = (size_t)nr * N3;
        const size_t S_cell_stride = (size_t)(pslots + nslots) * slot_stride;
        double *S2 = G.S + (size_t)pslots * slot_stride;
        int ci = 0;
        bool nyq_pending[2] = {false, false};
        for (int c0 = 0; c0 < p->pairs_local; c0 += p->chunk, ++ci) {
            const int nc = std::min(p->chunk, p->pairs_local - c0);
            const int items = cg * nc;
            const int ub = ci & 1;
            cplx *uvw = G.uvw + (size_t)ub * 3 * N * N * p->chunk * G.cells;
            if (nyq_pending[ub]) {
                CUDA_TRY(cudaStreamWaitEvent(st, G.ev_nyq[ub], 0));
                nyq_pending[ub] = false;
            }
            auto launch = [&](auto tm) {
                constexpr bool TM = decltype(tm)::value;
                using Rl = R32Launch<N, TM>;
                const int groups = (N + 3) * items;
                const int grid = std::max(1, std::min(p->sm_count * Rl::MINB, (groups + Rl::G - 1) / Rl::G));
                k_plane_gain_r32<N, Rl::G, Rl::MINB, TM>
                    <<<grid, Rl::G * R32Geo<N>::GT, plane_r32_smem<N, Rl::G>(), st>>>(
                        G.fhat, p->phase, p->zpm_r32, G.hyb, c0, items, G.nyq, p->pair_w, uvw, nc);
            };
            if (p->plane_r32 == 2) launch(std::true_type{});
            else launch(std::false_type{});
            const int u0 = p->chunk_unit_first[ci], nu = p->chunk_unit_first[ci + 1] - u0;
            CUDA_TRY(cudaEventRecord(G.ev_plane[ub], st));
            CUDA_TRY(cudaStreamWaitEvent(G.side, G.ev_plane[ub], 0));
            constexpr int NYQ_TILES = (N / 16) * (N / 16) * (N / 16);
            k_nyq_accum<N><<<dim3(NYQ_TILES, nu, cg), 256, 0, G.side>>>(uvw, c0, p->units + u0, S2, nr, nc,
                                                                        S_cell_stride);
            CUDA_TRY(cudaEventRecord(G.ev_nyq[ub], G.side));
            nyq_pending[ub] = true;
            const dim3 grid(PencilGeo<N>::WT / PR_WARPS, nu, cg);
            if (p->uniform_w)
                k_pencil_gain_reg<N, true, PR_WARPS, PR_MINB><<<grid, PR_WARPS * 32, 0, st>>>(
                    G.hyb, c0, p->units + u0, p->pair_w, G.S, nr, nc, S_cell_stride);
            else
                k_pencil_gain_reg<N, false, PR_WARPS, PR_MINB><<<grid, PR_WARPS * 32, 0, st>>>(
                    G.hyb, c0, p->units + u0, p->pair_w, G.S, nr, nc, S_cell_stride);
        }
        for (int ub = 0; ub < 2; ++ub)
            if (nyq_pending[ub]) CUDA_TRY(cudaStreamWaitEvent(st, G.ev_nyq[ub], 0));
        // Qhat = sum_r coef_r FFT3(S_r)   (cpp:249-273)
        if (nr > 0)
            k_plane<N, -1, PLANE_REAL><<<dim3(N, nr, cg), N * Geo<N>::B, plane_smem<N>(), st>>>(
                G.S, pslots, slot_stride, nullptr, nullptr, nullptr, p->tw, G.tmp, p->slots_of_r, S2, nslots,
                p->slots_of_r, S_cell_stride, (size_t)nr * N3);
        k_pencil_accum<N><<<dim3(TILES, cg), AccumGeo<N>::RS * TGP, 0, st>>>(G.tmp, p->tw, p->coef, nr, p->M, G.qhat);
        // loss term, inverse transforms, combine (cpp:281-330)
        k_plane<N, +1, PLANE_FINAL><<<dim3(N, 2, cg), N * Geo<N>::B, plane_smem<N>(), st>>>(
            nullptr, 0, 0, G.qhat, G.fhat, p->beta2, p->tw, G.tmp, nullptr, nullptr, 0, nullptr, 0, 2 * N3);
        k_pencil_final<N, true><<<dim3(TILES, cg), TGP, 0, st>>>(G.tmp, p->tw, f, Q);
        CUDA_TRY(cudaGetLastError());
        return BFSM_OK;
    }
}

// ---- general path (bfsm_general.cuh) -------------------------------------------------------
int gen_blocks(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }

// in-place 3-D transform of `n_arr` consecutive arrays
template <int SIGN> int gen_fft3(bfsm_plan *p, cplx *data, int n_arr, cudaStream_t st)
{
    const int n[3] = {p->nx, p->ny, p->nz};
    const long long inner[3] = {(long long)p->ny * p->nz, p->nz, 1};
    const long long total = (long long)p->nx * p->ny * p->nz * n_arr;
    for (int ax = 2; ax >= 0; --ax) {
        const int len = n[ax];
        int lg = -1;
        if ((len & (len - 1)) == 0) { lg = 0; while ((1 << lg) < len) ++lg; }
        const long long lines = total / len;
        const size_t smem = sizeof(cplx) * GEN_TL * (len + 1) * (lg >= 0 ? 1 : 2);
        k_gen_fft_axis<SIGN><<<(unsigned)((lines + GEN_TL - 1) / GEN_TL), 256, smem, st>>>(
            data, len, inner[ax], lines, p->gen_tw[ax], lg);
    }
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

int gen_gain_hat(bfsm_plan *p, cplx *qhat_out, const double *f, cudaStream_t st)
{
    const size_t n = (size_t)p->nx * p->ny * p->nz;
    int rc;
    // fhat = FFT3(f) / N   (cpp:168-186; the 1/N of cpp:162 folded in)
    k_gen_r2c<<<gen_blocks(n), 256, 0, st>>>(f, 1.0 / (double)n, p->fhat, n);
    if ((rc = gen_fft3<-1>(p, p->fhat, 1, st))) return rc;
    CUDA_TRY(cudaMemsetAsync(p->S, 0, sizeof(double) * n * std::max(1, p->n_r_local), st));
    for (int c0 = 0; c0 < p->pairs_local; c0 += p->chunk) {
        const int nc = std::min(p->chunk, p->pairs_local - c0);
        k_gen_phase<<<gen_blocks(n * nc), 256, 0, st>>>(p->fhat, p->phase, c0, nc, p->nx, p->ny, p->nz, p->hyb);
        if ((rc = gen_fft3<+1>(p, p->hyb, 2 * nc, st))) return rc;
        k_gen_prod_acc<<<gen_blocks(n), 256, 0, st>>>(p->hyb, p->pair_r, p->pair_w, c0, nc, n, p->S);
    }
    // Qhat = sum_r coef_r(|l|^2) FFT3(S_r), radii in batches of the scratch capacity
    const int cap = 2 * p->chunk;
    if (p->n_r_local == 0) CUDA_TRY(cudaMemsetAsync(qhat_out, 0, sizeof(cplx) * n, st));
    for (int r0 = 0; r0 < p->n_r_local; r0 += cap) {
        const int na = std::min(cap, p->n_r_local - r0);
        k_gen_real_batch<<<gen_blocks(n * na), 256, 0, st>>>(p->S + (size_t)r0 * n, p->hyb, n * na);
        if ((rc = gen_fft3<-1>(p, p->hyb, na, st))) return rc;
        k_gen_coef_acc<<<gen_blocks(n), 256, 0, st>>>(p->hyb, p->coef, p->M, r0, na, p->nx, p->ny, p->nz,
                                                      r0 == 0 ? 1 : 0, qhat_out);
    }
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

int gen_finish(bfsm_plan *p, double *Q, const cplx *qhat, const double *f, cudaStream_t st, bool with_loss)
{
    const size_t n = (size_t)p->nx * p->ny * p->nz;
    k_gen_final_in<<<gen_blocks(n), 256, 0, st>>>(qhat, p->fhat, p->beta2, p->nx, p->ny, p->nz, p->tmp);
    int rc = gen_fft3<+1>(p, p->tmp, with_loss ? 2 : 1, st);
    if (rc) return rc;
    k_gen_final_out<<<gen_blocks(n), 256, 0, st>>>(p->tmp, f, Q, n, with_loss ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

#define DISPATCH_N(p, CALL)                                           \
    switch ((p)->N) {                                                 \
    case 16: { constexpr int N_ = 16; return CALL; }                  \
    case 32: { constexpr int N_ = 32; return CALL; }                  \
    case 64: { constexpr int N_ = 64; return CALL; }                  \
    default: return fail(BFSM_ERR_UNSUPPORTED, "unsupported grid size"); \
    }

int do_gain_hat(bfsm_plan *p, cplx *qhat, const double *f, cudaStream_t st)
{
    if (p->general) return gen_gain_hat(p, qhat, f, st);
    DISPATCH_N(p, run_gain_hat<N_>(p, qhat, f, st));
}
int do_finish(bfsm_plan *p, double *Q, const cplx *qhat, const double *f, cudaStream_t st, bool with_loss = true)
{
    if (p->general) return gen_finish(p, Q, qhat, f, st, with_loss);
    DISPATCH_N(p, run_finish<N_>(p, Q, qhat, f, st, with_loss));
}
int do_group(bfsm_plan *p, bfsm_plan::Group &G, double *Q, const double *f, int cg)
{
    DISPATCH_N(p, run_group<N_>(p, G, Q, f, cg));
}
int do_configure(bfsm_plan *p) { DISPATCH_N(p, configure_kernels<N_>()); }
int do_launch_count(const bfsm_plan *p)
{
    if (p->general) {
        const int chunks = (p->pairs_local + p->chunk - 1) / std::max(1, p->chunk);
        const int rb = (p->n_r_local + 2 * p->chunk - 1) / std::max(1, 2 * p->chunk);
        return 4 + 5 * chunks + 5 * rb + 5;
    }
    DISPATCH_N(p, launches_per_cell<N_>(p));
}

// ---- TMA descriptor of a hybrid scratch buffer ----------------------------------------------
// hyb[pair][x][y][z] (complex doubles) as a rank-3 tensor of doubles: dim0 = 2 N (z, re/im interleaved),
// dim1 = N (y), dim2 = N * pairs (pair * N + x); box = one x-stage tile (2 TZ, 1, N).  The encoder is a
// driver entry point fetched through the runtime, so the library does not link libcuda.
int make_hyb_tmap(const bfsm_plan *p, const cplx *hyb, size_t grids, CUtensorMap *out)
{
    std::memset(out, 0, sizeof *out);
    if (!(p->packed && p->pencil_kernel == 3) || grids == 0) return BFSM_OK;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess)
        return fail(BFSM_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t N = (cuuint64_t)p->N;
    const cuuint64_t dims[3] = {2 * N, N, N * (cuuint64_t)grids};
    const cuuint64_t strides[2] = {N * sizeof(cplx), N * N * sizeof(cplx)};
    const cuuint32_t box[3] = {2 * TZ, 1, (cuuint32_t)N};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)hyb, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[96];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return fail(BFSM_ERR_CUDA, b);
    }
    return BFSM_OK;
}

// ---- batch lanes --------------------------------------------------------------------------
void lane_save(bfsm_plan *p, int k)
{
    bfsm_plan::Lane &L = p->lanes[k];
    L.fhat = p->fhat; L.tmp = p->tmp; L.hyb = p->hyb; L.nyq = p->nyq; L.uvw = p->uvw;
    L.qhat = p->qhat; L.S = p->S; L.side = p->side; L.sync_flags = p->sync_flags; L.hyb_tmap = p->hyb_tmap;
    for (int j = 0; j < 2; ++j) { L.ev_plane[j] = p->ev_plane[j]; L.ev_nyq[j] = p->ev_nyq[j]; }
}
void lane_activate(bfsm_plan *p, int k)
{
    const bfsm_plan::Lane &L = p->lanes[k];
    p->fhat = L.fhat; p->tmp = L.tmp; p->hyb = L.hyb; p->nyq = L.nyq; p->uvw = L.uvw;
    p->qhat = L.qhat; p->S = L.S; p->side = L.side; p->sync_flags = L.sync_flags; p->hyb_tmap = L.hyb_tmap;
    for (int j = 0; j < 2; ++j) { p->ev_plane[j] = L.ev_plane[j]; p->ev_nyq[j] = L.ev_nyq[j]; }
}
void lane_free(bfsm_plan *p, int k)
{
    bfsm_plan::Lane &L = p->lanes[k];
    if (!L.allocated) return;
    void *bufs[] = {L.fhat, L.tmp, L.hyb, L.nyq, L.uvw, L.qhat, L.S, L.sync_flags};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (L.side) cudaStreamDestroy(L.side);
    for (int j = 0; j < 2; ++j) {
        if (L.ev_plane[j]) cudaEventDestroy(L.ev_plane[j]);
        if (L.ev_nyq[j]) cudaEventDestroy(L.ev_nyq[j]);
    }
    cudaStream_t keep_main = L.main;
    cudaEvent_t keep_done = L.done;
    L = bfsm_plan::Lane();
    L.main = keep_main;
    L.done = keep_done;
}
void lanes_free(bfsm_plan *p)
{
    for (int k = 1; k < bfsm_plan::MAX_LANES; ++k) lane_free(p, k);
}
int lane_alloc(bfsm_plan *p, int k);
// Makes lanes 0 .. want-1 usable (lane 0 is the plan's own scratch) and returns how many are: fewer
// than `want` when device memory runs out -- the batch then runs with the lanes that exist.
// `fail_from` (test hook, <= 0 = off) makes the allocation of lane `fail_from` and above fail.
int lanes_prepare(bfsm_plan *p, int want, int *usable)
{
    *usable = 1;
    if (!p->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    if (!p->lanes[0].main) {
        for (int k = 0; k < bfsm_plan::MAX_LANES; ++k) {
            CUDA_TRY(cudaStreamCreateWithFlags(&p->lanes[k].main, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&p->lanes[k].done, cudaEventDisableTiming));
        }
    }
    lane_save(p, 0); // lane 0 == the plan's own scratch
    for (int k = 1; k < want; ++k) {
        int rc = (p->lane_fail_from > 0 && k >= p->lane_fail_from)
                     ? fail(BFSM_ERR_NOMEM, "batch lane allocation failure injected by the test hook")
                     : lane_alloc(p, k);
        if (rc == BFSM_ERR_NOMEM) { // not enough device memory for another lane: run with fewer
            cudaGetLastError();
            break;
        }
        if (rc) return rc;
        *usable = k + 1;
    }
    return BFSM_OK;
}
int lane_alloc(bfsm_plan *p, int k)
{
    const int N = p->N;
    const size_t N3 = (size_t)N * N * N;
    bfsm_plan::Lane &L = p->lanes[k];
    if (L.allocated) return BFSM_OK;
    L.allocated = true; // from here on lane_free() releases whatever was obtained
    auto need = [&](void **q, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(q, bytes ? bytes : 16);
        if (e != cudaSuccess) {
            *q = nullptr;
            char b[200];
            snprintf(b, sizeof b, "cudaMalloc(%zu bytes) for batch lane %d failed: %s", bytes, k,
                     cudaGetErrorString(e));
            return fail(e == cudaErrorMemoryAllocation ? BFSM_ERR_NOMEM : BFSM_ERR_CUDA, b);
        }
        return BFSM_OK;
    };
    auto give_up = [&](int rc) {
        std::string keep = g_err;
        lane_free(p, k);
        g_err = keep;
        return rc;
    };
    int rc = BFSM_OK;
    const int nr = std::max(1, p->n_r_local);
    if ((rc = need((void **)&L.fhat, sizeof(cplx) * N3))) return give_up(rc);
    if ((rc = need((void **)&L.qhat, sizeof(cplx) * N3))) return give_up(rc);
    if ((rc = need((void **)&L.tmp, sizeof(cplx) * N3 * std::max(2, p->n_r_local)))) return give_up(rc);
    if ((rc = need((void **)&L.hyb, sizeof(cplx) * N3 * hyb_grids(p)))) return give_up(rc);
    if ((rc = make_hyb_tmap(p, L.hyb, hyb_grids(p), &L.hyb_tmap))) return give_up(rc);
    if (p->fused && (rc = need((void **)&L.sync_flags, sizeof(int) * 2 * (size_t)std::max(1, fused_subs(p)))))
        return give_up(rc);
    if ((rc = need((void **)&L.S, sizeof(double) * N3 * (size_t)std::max(1, p->S_slots_capacity) * nr)))
        return give_up(rc);
    if (p->packed) {
        if ((rc = need((void **)&L.nyq, sizeof(cplx) * 3 * N * N))) return give_up(rc);
        if ((rc = need((void **)&L.uvw, sizeof(cplx) * 3 * N * N * uvw_sets(p)))) return give_up(rc);
        if (p->use_side) {
            if (cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking) != cudaSuccess)
                return give_up(fail(BFSM_ERR_CUDA, "cudaStreamCreate failed"));
            for (int j = 0; j < 2; ++j) {
                if (cudaEventCreateWithFlags(&L.ev_plane[j], cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&L.ev_nyq[j], cudaEventDisableTiming) != cudaSuccess)
                    return give_up(fail(BFSM_ERR_CUDA, "cudaEventCreate failed"));
            }
        }
    }
    return BFSM_OK;
}

struct GuardDevice {
    int prev = -1;
    bool ok = true;
    explicit GuardDevice(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~GuardDevice()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

} // namespace

// =========================================================================== C ABI

extern "C" int bfsm_version(void) { return BFSM_VERSION; }

extern "C" const char *bfsm_last_error(void) { return g_err.c_str(); }

extern "C" void bfsm_plan_options_init(bfsm_plan_options *o)
{
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->struct_size = (int)sizeof *o;
    o->side_stream = 1;
}

extern "C" int bfsm_plan_create(bfsm_plan **out, int nvx, int nvy, int nvz, int n_r,
                                const double *rho, const double *w_r, int n_s, const double *sx,
                                const double *sy, const double *sz, const double *w_s,
                                double gamma, double b_gamma, double L, int device,
                                int shard_index, int shard_count, unsigned flags)
{
    return bfsm_plan_create_ex(out, nvx, nvy, nvz, n_r, rho, w_r, n_s, sx, sy, sz, w_s, gamma, b_gamma, L,
                               device, shard_index, shard_count, flags, nullptr);
}

extern "C" int bfsm_plan_create_ex(bfsm_plan **out, int nvx, int nvy, int nvz, int n_r,
                                   const double *rho, const double *w_r, int n_s, const double *sx,
                                   const double *sy, const double *sz, const double *w_s,
                                   double gamma, double b_gamma, double L, int device,
                                   int shard_index, int shard_count, unsigned flags,
                                   const bfsm_plan_options *opts)
{
    if (!out) return fail(BFSM_ERR_INVALID, "out is NULL");
    bfsm_plan_options opt;
    bfsm_plan_options_init(&opt);
    if (opts) {
        if (opts->struct_size != (int)sizeof opt)
            return fail(BFSM_ERR_INVALID, "bfsm_plan_options: struct_size mismatch (call bfsm_plan_options_init)");
        opt = *opts;
    }
    if (opt.chunk_pairs < 0 || opt.seg_pairs < 0 || opt.gain_ctas < 0 ||
        opt.batch_lanes < 0 || opt.pencil_kernel < 0 || opt.pencil_kernel > 3 || opt.plane_kernel < 0 ||
        opt.plane_kernel > 4 || opt.pencil_groups < 0 || opt.gain_pipeline < 0 || opt.gain_pipeline > 3)
        return fail(BFSM_ERR_INVALID, "bfsm_plan_options: field out of range");
    *out = nullptr;
    if (!rho || !w_r || !sx || !sy || !sz || !w_s)
        return fail(BFSM_ERR_INVALID, "quadrature pointer is NULL");
    if (n_r <= 0 || n_s <= 0) return fail(BFSM_ERR_INVALID, "quadrature sizes must be positive");
    if (nvx <= 0 || nvy <= 0 || nvz <= 0) return fail(BFSM_ERR_INVALID, "grid sizes must be positive");
    if (!(L > 0.0)) return fail(BFSM_ERR_INVALID, "L must be positive");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
        return fail(BFSM_ERR_INVALID, "shard_index/shard_count out of range");
    // cubic 16 / 32 / 64: the tuned kernels.  Any other even sizes from 4 to 128 per axis (non-cubic,
    // not a power of two, 128): the general path (bfsm_general.cuh).
    const bool tuned = (nvx == nvy && nvy == nvz && (nvx == 16 || nvx == 32 || nvx == 64)) &&
                       !(flags & BFSM_FLAG_GENERAL);
    for (int n : {nvx, nvy, nvz})
        if (n < 4 || n > GEN_MAXLEN || (n & 1)) {
            char b[200];
            snprintf(b, sizeof b, "grid %dx%dx%d not supported: every axis must be even and between 4 and %d "
                     "(the mode tables of FFTWBoltzmannOperator.cpp:50-57 assume even sizes)", nvx, nvy, nvz, GEN_MAXLEN);
            return fail(BFSM_ERR_UNSUPPORTED, b);
        }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BFSM_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(BFSM_ERR_INVALID, "device ordinal out of range");
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");

    bfsm_plan *p = new bfsm_plan;
    p->N = nvx;
    p->nx = nvx; p->ny = nvy; p->nz = nvz;
    p->general = tuned ? 0 : 1;
    p->n_r = n_r;
    p->n_s = n_s;
    p->device = device;
    p->gamma = gamma;
    p->b_gamma = b_gamma;
    p->L = L;
    p->shard_index = shard_index;
    p->shard_count = shard_count;
    p->packed = (flags & BFSM_FLAG_NO_PACK) ? 0 : 1;
    const int N = p->N;
    const size_t N3 = (size_t)nvx * nvy * nvz;
    const int nax[3] = {nvx, nvy, nvz};
    const int PT = nvx + nvy + nvz; // entries of a pair's phase table [ex | ey | ez]

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p->sm_count = prop.multiProcessorCount;

    // ---- antipodal folding: partner[s] = index of -sigma_s with equal weight, or -1
    std::vector<int> rep; // directions actually transformed
    std::vector<double> rep_w;
    {
        std::vector<int> partner(n_s, -1);
        bool all = !(flags & BFSM_FLAG_NO_FOLD) && (n_s % 2 == 0);
        if (all) {
            for (int s = 0; s < n_s && all; ++s) {
                if (partner[s] >= 0) continue;
                int found = -1;
                for (int t = 0; t < n_s; ++t) {
                    if (t == s || partner[t] >= 0) continue;
                    if (sx[t] == -sx[s] && sy[t] == -sy[s] && sz[t] == -sz[s] && w_s[t] == w_s[s]) {
                        found = t;
                        break;
                    }
                }
                if (found < 0) {
                    all = false;
                } else {
                    partner[s] = found;
                    partner[found] = s;
                }
            }
        }
        p->folded = all ? 1 : 0;
        for (int s = 0; s < n_s; ++s) {
            if (p->folded) {
                if (partner[s] > s) { // keep the lower index of each antipodal couple
                    rep.push_back(s);
                    rep_w.push_back(2.0 * w_s[s]);
                }
            } else {
                rep.push_back(s);
                rep_w.push_back(w_s[s]);
            }
        }
    }
    const int n_dir = (int)rep.size();
    p->pairs_total = n_r * n_dir;
    const int lo = (int)(((long long)p->pairs_total * shard_index) / shard_count);
    const int hi = (int)(((long long)p->pairs_total * (shard_index + 1)) / shard_count);
    p->pairs_local = hi - lo;
    p->n_dir = n_dir;
    p->pair_lo = lo;

    // ---- per-pair tables (work list is r-major: pair = r*n_dir + d)
    const int r_first = (p->pairs_local > 0) ? lo / n_dir : 0;
    const int r_last = (p->pairs_local > 0) ? (hi - 1) / n_dir : -1;
    p->n_r_local = r_last - r_first + 1;
    std::vector<int> h_pair_r(std::max(p->pairs_local, 1));
    std::vector<double> h_pair_w(std::max(p->pairs_local, 1));
    std::vector<int> h_r_end(std::max(p->n_r_local, 1), 0);
    std::vector<cplx> h_phase((size_t)std::max(p->pairs_local, 1) * PT);
    for (int q = 0; q < p->pairs_local; ++q) {
        const int pair = lo + q;
        const int r = pair / n_dir, d = pair % n_dir, s = rep[d];
        h_pair_r[q] = r - r_first;
        h_pair_w[q] = rep_w[d];
        h_r_end[r - r_first] = q + 1;
        // theta = -(pi/(2L)) * rho_r * (l . sigma)  (cpp:209-210), separable in the three axes
        const long double c = -(PI_L / (2.0L * (long double)L)) * (long double)rho[r];
        const double sig[3] = {sx[s], sy[s], sz[s]};
        size_t off = (size_t)q * PT;
        for (int ax = 0; ax < 3; ++ax) {
            for (int t = 0; t < nax[ax]; ++t) {
                const int mode = t < nax[ax] / 2 ? t : t - nax[ax]; // cpp:50-57
                const long double th = c * (long double)mode * (long double)sig[ax];
                h_phase[off + t] = make_double2((double)cosl(th), (double)sinl(th));
            }
            off += nax[ax];
        }
    }

    // ---- beta1 / beta2 tables indexed by the integer |l|^2 (cpp:252-265, 281-299)
    p->M = (nvx / 2) * (nvx / 2) + (nvy / 2) * (nvy / 2) + (nvz / 2) * (nvz / 2) + 1;
    const double fft_scale = 1.0 / (double)N3; // cpp:162
    std::vector<double> h_coef((size_t)std::max(p->n_r_local, 1) * p->M, 0.0);
    std::vector<double> h_beta2(p->M, 0.0);
    for (int m = 0; m < p->M; ++m) {
        const double norm_l = std::sqrt((double)m);
        for (int rl = 0; rl < p->n_r_local; ++rl) {
            const int r = r_first + rl;
            const double beta1 = 4 * PI_D * b_gamma * sincc(PI_D * rho[r] * norm_l / (2 * L));
            h_coef[(size_t)rl * p->M + m] = fft_scale * w_r[r] * std::pow(rho[r], gamma + 2) * beta1;
        }
        double b2 = 0.0;
        for (int r = 0; r < n_r; ++r)
            b2 += 16 * PI_D * PI_D * b_gamma * w_r[r] * std::pow(rho[r], gamma + 2) *
                  sincc(PI_D * rho[r] * norm_l / L);
        h_beta2[m] = b2; // fft_scale is already folded into fhat
    }
    std::vector<cplx> h_tw(N);
    for (int t = 0; t < N; ++t) {
        const long double a = 2.0L * PI_L * (long double)t / (long double)N;
        h_tw[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    if (p->general) {
        // ---- general path: per-axis twiddle tables, plain scratch, nothing else
        p->opt = opt;
        p->packed = 0;
        p->uniform_w = 0;
        p->chunk = p->chunk_capacity = std::max(1, std::min(opt.chunk_pairs > 0 ? opt.chunk_pairs : 8,
                                                             std::max(1, p->pairs_local)));
        int rc = BFSM_OK;
        auto bailg = [&](int code) {
            std::string keep = g_err;
            bfsm_plan_destroy(p);
            g_err = keep;
            return code;
        };
        for (int ax = 0; ax < 3; ++ax) {
            std::vector<cplx> tw(nax[ax]);
            for (int t = 0; t < nax[ax]; ++t) {
                const long double a = 2.0L * PI_L * (long double)t / (long double)nax[ax];
                tw[t] = make_double2((double)cosl(a), (double)sinl(a));
            }
            if ((rc = upload(p, &p->gen_tw[ax], tw))) return bailg(rc);
        }
        if ((rc = upload(p, &p->phase, h_phase))) return bailg(rc);
        if ((rc = upload(p, &p->pair_r, h_pair_r))) return bailg(rc);
        if ((rc = upload(p, &p->pair_w, h_pair_w))) return bailg(rc);
        if ((rc = upload(p, &p->r_end, h_r_end))) return bailg(rc);
        if ((rc = upload(p, &p->coef, h_coef))) return bailg(rc);
        if ((rc = upload(p, &p->beta2, h_beta2))) return bailg(rc);
        if ((rc = dev_alloc(p, (void **)&p->fhat, sizeof(cplx) * N3))) return bailg(rc);
        if ((rc = dev_alloc(p, (void **)&p->qhat, sizeof(cplx) * N3))) return bailg(rc);
        if ((rc = dev_alloc(p, (void **)&p->tmp, sizeof(cplx) * N3 * 2))) return bailg(rc);
        if ((rc = dev_alloc(p, (void **)&p->hyb, sizeof(cplx) * N3 * 2 * (size_t)p->chunk))) return bailg(rc);
        {
            const size_t bytes = sizeof(double) * N3 * std::max(1, p->n_r_local);
            cudaError_t e = cudaMalloc((void **)&p->S, bytes);
            if (e != cudaSuccess)
                return bailg(fail(e == cudaErrorMemoryAllocation ? BFSM_ERR_NOMEM : BFSM_ERR_CUDA,
                                  "cudaMalloc of the per-radius sums failed"));
            p->scratch_bytes += (long long)bytes;
            p->S_slots_capacity = 1;
        }
        p->n_lanes = 1;
        CUDA_TRY(cudaFuncSetAttribute(k_gen_fft_axis<+1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(2 * sizeof(cplx) * GEN_TL * (GEN_MAXLEN + 1))));
        CUDA_TRY(cudaFuncSetAttribute(k_gen_fft_axis<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(2 * sizeof(cplx) * GEN_TL * (GEN_MAXLEN + 1))));
        *out = p;
        return BFSM_OK;
    }

    // ---- launch geometry
    p->opt = opt;
    int dflt_chunk = (N == 64) ? Launch<64>::CHUNK : (N == 32) ? Launch<32>::CHUNK : Launch<16>::CHUNK;
    p->chunk = opt.chunk_pairs > 0 ? opt.chunk_pairs : dflt_chunk;
    p->chunk = std::min(p->chunk, std::max(1, p->pairs_local));
    p->chunk_capacity = p->chunk;
    p->G = (N == 64) ? Launch<64>::G : (N == 32) ? Launch<32>::G : Launch<16>::G;
    if (opt.pencil_groups > 0) p->G = std::min(opt.pencil_groups, 64);
    {
        const int occ = (N == 64) ? Launch<64>::MINB : (N == 32) ? Launch<32>::MINB : Launch<16>::MINB;
        p->gy = opt.gain_ctas > 0 ? opt.gain_ctas : p->sm_count * occ;
    }
    // x stage: at 64^3 every variant is HBM bound; the staged one interferes less with the side-stream
    // Nyquist accumulate (174 vs 170 evals/s) and its TMA-filled ring spends no LSU issue slots on the
    // copies (bitwise equal to the LDGSTS fill, same speed); at 32^3 / 16^3 the register-resident kernel
    // is 15 % faster (profiles/r02_ab64_tma.log, r02_ab32_tma.log)
    p->pencil_kernel = opt.pencil_kernel > 0 ? opt.pencil_kernel : (N == 64 ? 3 : 2);
    p->plane_ws = (N == 64 && opt.plane_kernel != 1) ? 1 : 0;
    // 32^3: the radix-32 kernel with the fhat line in tensor memory is the default: 0.094 ms per 752-pair
    // launch against 0.115 (line in registers) and 0.142 (k_plane_gain3), profiles/r02_ab32_r32c.log; at
    // 64^3 it is opt-in (3.9 ms per evaluation against 3.2 for the pipelined kernel, r02_ab64_r32c.log)
    p->plane_r32 = ((N == 64 || N == 32) && p->packed && opt.plane_kernel >= 3) ? opt.plane_kernel - 2
                   : (N == 32 && p->packed && opt.plane_kernel == 0)            ? 2
                                                                                : 0;
    if (opt.plane_kernel >= 3 && !p->plane_r32)
        return (delete p, fail(BFSM_ERR_UNSUPPORTED, "plane_kernel = 3 / 4 (radix-32 plane kernel) needs a 64^3 or 32^3 grid in packed mode"));
    // work units of the register-resident x stage: a quarter of a radius' directions, at most 24 pairs
    p->seg_pairs = opt.seg_pairs > 0 ? opt.seg_pairs : std::max(1, std::min(24, (n_dir + 3) / 4));
    p->n_lanes = std::min((int)bfsm_plan::MAX_LANES, opt.batch_lanes > 0 ? opt.batch_lanes : 4);
    p->use_side = opt.side_stream ? 1 : 0;
    p->uniform_w = 1;
    for (double w : rep_w)
        if (w != rep_w[0]) p->uniform_w = 0;
    // fused persistent gain kernel: 64^3, packed mode, at least one pair, a GPU with room for every role
    p->fused = (N == 64 && p->packed && opt.gain_pipeline == 2 && p->pairs_local > 0) ? 1 : 0;
    if (opt.gain_pipeline == 2 && !p->fused && !(N == 64 && p->packed))
        return (delete p, fail(BFSM_ERR_UNSUPPORTED, "gain_pipeline = 2 (fused kernel) needs a 64^3 grid in packed mode"));
    p->cluster = (N == 32 && p->packed && opt.gain_pipeline == 3 && p->pairs_local > 0) ? 1 : 0;
    if (opt.gain_pipeline == 3 && !(N == 32 && p->packed))
        return (delete p, fail(BFSM_ERR_UNSUPPORTED, "gain_pipeline = 3 (cluster kernel) needs a 32^3 grid in packed mode"));
    if (p->fused) {
        p->fused_K = opt.fused_sub_pairs > 0 ? opt.fused_sub_pairs : 12;
        p->fused_K = std::min(p->fused_K, p->pairs_local);
        p->fused_D = opt.fused_ring > 0 ? opt.fused_ring : 2;
        p->fused_NN = opt.fused_nyq_ctas > 0 ? opt.fused_nyq_ctas : 6;
        p->fused_NQ = opt.fused_pencil_ctas > 0 ? opt.fused_pencil_ctas : (p->sm_count * 38) / 148;
        if (p->fused_NN % 3 != 0 || p->fused_NQ < 1 || p->fused_NN + p->fused_NQ >= p->sm_count)
            return (delete p, fail(BFSM_ERR_INVALID, "fused kernel: role sizes do not fit the device"));
        p->chunk = p->chunk_capacity = std::max(1, p->pairs_local); // one launch covers the shard
    }

    int rc = BFSM_OK;
    auto bail = [&](int code) {
        std::string keep = g_err;
        bfsm_plan_destroy(p);
        g_err = keep;
        return code;
    };
    if ((rc = upload(p, &p->tw, h_tw))) return bail(rc);
    if ((rc = upload(p, &p->phase, h_phase))) return bail(rc);
    {
        // the packed multiplier only needs Re+Im and Re-Im of the z phase: tabulated once here
        // (the same two IEEE additions the kernels would do per element and pair)
        std::vector<cplx> h_zpm((size_t)std::max(p->pairs_local, 1) * N);
        for (int q = 0; q < p->pairs_local; ++q)
            for (int t = 0; t < N; ++t) {
                const cplx ez = h_phase[((size_t)q * 3 + 2) * N + t];
                h_zpm[(size_t)q * N + t] = make_double2(ez.x + ez.y, ez.x - ez.y);
            }
        if ((rc = upload(p, &p->zpm, h_zpm))) return bail(rc);
        if (p->plane_r32) {
            // k_plane_gain_r32 at N = 64: lane 1 of a line feeds its radix-32 with (-1)^a in[2a+1]
            if (N == 64)
                for (int q = 0; q < p->pairs_local; ++q)
                    for (int t = 3; t < N; t += 4) {
                        cplx &z = h_zpm[(size_t)q * N + t];
                        z = make_double2(-z.x, -z.y);
                    }
            if ((rc = upload(p, &p->zpm_r32, h_zpm))) return bail(rc);
        }
    }
    if ((rc = upload(p, &p->pair_r, h_pair_r))) return bail(rc);
    if ((rc = upload(p, &p->pair_w, h_pair_w))) return bail(rc);
    if ((rc = upload(p, &p->r_end, h_r_end))) return bail(rc);
    if ((rc = upload(p, &p->coef, h_coef))) return bail(rc);
    if ((rc = upload(p, &p->beta2, h_beta2))) return bail(rc);
    if ((rc = dev_alloc(p, (void **)&p->fhat, sizeof(cplx) * N3))) return bail(rc);
    if ((rc = dev_alloc(p, (void **)&p->qhat, sizeof(cplx) * N3))) return bail(rc);
    if ((rc = dev_alloc(p, (void **)&p->tmp, sizeof(cplx) * N3 * std::max(2, p->n_r_local))))
        return bail(rc);
    if ((rc = dev_alloc(p, (void **)&p->hyb, sizeof(cplx) * N3 * hyb_grids(p)))) return bail(rc);
    if ((rc = make_hyb_tmap(p, p->hyb, hyb_grids(p), &p->hyb_tmap))) return bail(rc);
    if (p->fused &&
        (rc = dev_alloc(p, (void **)&p->sync_flags, sizeof(int) * 2 * (size_t)std::max(1, fused_subs(p)))))
        return bail(rc);
    if (p->packed) {
        if ((rc = dev_alloc(p, (void **)&p->nyq, sizeof(cplx) * 3 * N * N))) return bail(rc);
        if ((rc = dev_alloc(p, (void **)&p->uvw, sizeof(cplx) * 3 * N * N * uvw_sets(p)))) return bail(rc);
        if (p->use_side) {
            if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess)
                return bail(fail(BFSM_ERR_CUDA, "cudaStreamCreate failed"));
            for (int k = 0; k < 2; ++k) {
                if (cudaEventCreateWithFlags(&p->ev_plane[k], cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&p->ev_nyq[k], cudaEventDisableTiming) != cudaSuccess)
                    return bail(fail(BFSM_ERR_CUDA, "cudaEventCreate failed"));
            }
        }
    }
    if ((rc = relayout(p))) return bail(rc); // unit table, slot layout, partial-sum slots S
    if ((rc = do_configure(p))) return bail(rc);
    *out = p;
    return BFSM_OK;
}

extern "C" int bfsm_plan_destroy(bfsm_plan *p)
{
    if (!p) return BFSM_OK;
    GuardDevice guard(p->device);
    for (void *q : p->allocs) cudaFree(q);
    if (p->S) cudaFree(p->S);
    if (p->units) cudaFree(p->units);
    if (p->slots_of_r) cudaFree(p->slots_of_r);
    for (cudaEvent_t e : p->event_pool) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        if (p->ev_plane[k]) cudaEventDestroy(p->ev_plane[k]);
        if (p->ev_nyq[k]) cudaEventDestroy(p->ev_nyq[k]);
    }
    if (p->side) cudaStreamDestroy(p->side);
    lanes_free(p);
    groups_free(p);
    for (int k = 0; k < bfsm_plan::MAX_LANES; ++k) {
        if (p->lanes[k].main) cudaStreamDestroy(p->lanes[k].main);
        if (p->lanes[k].done) cudaEventDestroy(p->lanes[k].done);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->stage_f) cudaFree(p->stage_f);
    if (p->stage_q) cudaFree(p->stage_q);
    for (int b = 0; b < bfsm_plan::HostPipe::DEPTH; ++b) {
        if (p->pipe.f[b]) cudaFree(p->pipe.f[b]);
        if (p->pipe.q[b]) cudaFree(p->pipe.q[b]);
        if (p->pipe.in_done[b]) cudaEventDestroy(p->pipe.in_done[b]);
        if (p->pipe.comp_done[b]) cudaEventDestroy(p->pipe.comp_done[b]);
        if (p->pipe.out_done[b]) cudaEventDestroy(p->pipe.out_done[b]);
    }
    if (p->pipe.h2d) cudaStreamDestroy(p->pipe.h2d);
    if (p->pipe.d2h) cudaStreamDestroy(p->pipe.d2h);
    delete p;
    return BFSM_OK;
}

// Host mirror of the work split of the gain plane kernels (k_plane_gain_ws / k_plane_gain3): the
// (plane, item) entries CTA `cta` of `n_ctas` walks, in order.  Same range arithmetic as the kernels
// (first an equal share of the n regular planes, then of the 3 Nyquist planes n..n+2) and the very
// ItemWalk the pipelined kernel steps with; lets the CPU test suite check that every entry is
// visited exactly once and that the shares are balanced.
extern "C" int bfsm_debug_plane_work(int n, int n_items, int n_ctas, int cta, int *planes,
                                     int *items, int capacity)
{
    if (n <= 0 || n_items <= 0 || n_ctas <= 0 || cta < 0 || cta >= n_ctas || capacity < 0 ||
        (capacity > 0 && (!planes || !items)))
        return -fail(BFSM_ERR_INVALID, "bfsm_debug_plane_work: bad argument");
    // the very walker the pipelined kernel steps with (LaunchWalk<N>; its arithmetic does not depend on N
    // beyond the plane count, k_plane_gain3 uses the same ranges)
    auto run = [&](auto walk) {
        walk.init(n_items, 0, cta, n_ctas);
        for (int k = 0; k < walk.cnt && k < capacity; ++k) {
            planes[k] = walk.i;
            items[k] = walk.dst_item;
            walk.next();
        }
        return walk.cnt;
    };
    if (n == 64) return run(LaunchWalk<64>());
    if (n == 32) return run(LaunchWalk<32>());
    if (n == 16) return run(LaunchWalk<16>());
    return -fail(BFSM_ERR_UNSUPPORTED, "bfsm_debug_plane_work: n must be 16, 32 or 64");
}

// Same for the radix-32 plane kernel (R32Walk: one contiguous range per group, the Nyquist entries on
// their own groups).
extern "C" int bfsm_debug_plane_work_r32(int n, int n_items, int n_groups, int group, int *planes,
                                         int *items, int capacity)
{
    if (n_items <= 0 || n_groups <= 0 || group < 0 || group >= n_groups || capacity < 0 ||
        (capacity > 0 && (!planes || !items)))
        return -fail(BFSM_ERR_INVALID, "bfsm_debug_plane_work_r32: bad argument");
    auto run = [&](auto walk) {
        walk.init(n_items, group, n_groups);
        for (int k = 0; k < walk.cnt && k < capacity; ++k) {
            planes[k] = walk.i;
            items[k] = walk.it;
            walk.next();
        }
        return walk.cnt;
    };
    if (n == 64) return run(R32Walk<64>());
    if (n == 32) return run(R32Walk<32>());
    return -fail(BFSM_ERR_UNSUPPORTED, "bfsm_debug_plane_work_r32: n must be 32 or 64");
}

extern "C" int bfsm_debug_shares_aligned(int pairs_local, int pair_lo, int n_dir, int chunk, int groups)
{
    return shares_start_at_radius_boundaries(pairs_local, pair_lo, n_dir, chunk, groups) ? 1 : 0;
}

extern "C" int bfsm_debug_units(int pairs_local, int pair_lo, int n_dir, int chunk, int seg_pairs,
                                int *out, int capacity)
{
    if (pairs_local < 0 || pair_lo < 0 || n_dir <= 0 || chunk <= 0 || seg_pairs <= 0 || capacity < 0 ||
        (capacity > 0 && !out))
        return -fail(BFSM_ERR_INVALID, "bfsm_debug_units: bad argument");
    std::vector<PencilUnit> units;
    std::vector<int> chunk_first, slots_of_r;
    cut_units(pairs_local, pair_lo, n_dir, chunk, seg_pairs, units, chunk_first, slots_of_r);
    for (size_t k = 0; k < units.size() && (int)k < capacity; ++k) {
        out[4 * k + 0] = units[k].p0;
        out[4 * k + 1] = units[k].p1;
        out[4 * k + 2] = units[k].r;
        out[4 * k + 3] = units[k].slot;
    }
    return (int)units.size();
}

extern "C" int bfsm_debug_fail_lane_alloc(bfsm_plan *p, int first_failing_lane)
{
    if (!p) return fail(BFSM_ERR_INVALID, "plan is NULL");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaDeviceSynchronize());
    lanes_free(p);
    p->lane_fail_from = first_failing_lane;
    return BFSM_OK;
}

extern "C" int bfsm_plan_set_chunk(bfsm_plan *p, int chunk_pairs)
{
    if (!p) return fail(BFSM_ERR_INVALID, "plan is NULL");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    const int N = p->N;
    const size_t N3 = (size_t)N * N * N;
    int dflt = (N == 64) ? Launch<64>::CHUNK : (N == 32) ? Launch<32>::CHUNK : Launch<16>::CHUNK;
    int c = chunk_pairs > 0 ? chunk_pairs : dflt;
    c = std::min(c, std::max(1, p->pairs_local));
    if (p->fused) return BFSM_OK; // the fused kernel covers the shard in one launch; see fused_sub_pairs
    if (p->general) { // plain scratch: any chunk up to the allocated capacity
        p->chunk = std::max(1, std::min(chunk_pairs > 0 ? chunk_pairs : p->chunk_capacity, p->chunk_capacity));
        return BFSM_OK;
    }
    CUDA_TRY(cudaDeviceSynchronize());
    lanes_free(p); // re-allocated lazily with the new chunk size
    groups_free(p);
    if (c > p->chunk_capacity) {
        // grow the per-chunk scratch (never shrunk: the capacity, not the current chunk, is tracked)
        auto regrow = [&](void **slot, size_t bytes_new, size_t bytes_old) -> int {
            void *q = nullptr;
            CUDA_TRY(cudaMalloc(&q, bytes_new));
            for (auto &a : p->allocs)
                if (a == *slot) a = q;
            cudaFree(*slot);
            *slot = q;
            p->scratch_bytes += (long long)bytes_new - (long long)bytes_old;
            return BFSM_OK;
        };
        const size_t per = sizeof(cplx) * N3 * (p->packed ? 1 : 2);
        int rc = regrow((void **)&p->hyb, per * c, per * p->chunk_capacity);
        if (rc) return rc;
        if ((rc = make_hyb_tmap(p, p->hyb, (size_t)(p->packed ? 1 : 2) * c, &p->hyb_tmap))) return rc;
        if (p->packed) {
            const size_t pern = sizeof(cplx) * 2 * 3 * N * N;
            rc = regrow((void **)&p->uvw, pern * c, pern * p->chunk_capacity);
            if (rc) return rc;
        }
        p->chunk_capacity = c;
    }
    p->chunk = c;
    return relayout(p);
}

extern "C" int bfsm_plan_get_info(const bfsm_plan *p, bfsm_plan_info *info)
{
    if (!p || !info) return fail(BFSM_ERR_INVALID, "NULL argument");
    info->n = p->N;
    info->n_r = p->n_r;
    info->n_s = p->n_s;
    info->folded = p->folded;
    info->packed = p->packed;
    info->pairs_total = p->pairs_total;
    info->pairs_local = p->pairs_local;
    info->chunk_pairs = p->chunk;
    info->launches_per_cell = do_launch_count(p);
    info->scratch_bytes = p->scratch_bytes;
    info->plane_kernel = !p->packed ? 0 : (p->plane_r32 && !p->cluster) ? 2 + p->plane_r32 : (p->N == 64 && p->plane_ws) ? 2 : 1;
    info->pencil_kernel = !p->packed ? 0 : p->pencil_kernel;
    info->batch_lanes_used = p->lanes_used_last;
    info->gain_pipeline = p->general ? 0 : (p->fused ? 2 : (p->cluster ? 3 : 1));
    info->ny = p->ny;
    info->nz = p->nz;
    info->general = p->general;
    info->batch_group_cells = p->group_cells_last;
    if (p->general) info->plane_kernel = info->pencil_kernel = -1;
    info->partial_slots = pencil_slots(p) + nyq_slots(p);
    return BFSM_OK;
}

extern "C" int bfsm_gain_hat(bfsm_plan *p, double *Qhat_dev, const double *f_dev, void *stream)
{
    if (!p || !Qhat_dev || !f_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    return do_gain_hat(p, reinterpret_cast<cplx *>(Qhat_dev), f_dev, (cudaStream_t)stream);
}

extern "C" int bfsm_finish(bfsm_plan *p, double *Q_dev, const double *Qhat_dev, const double *f_dev,
                           void *stream)
{
    if (!p || !Q_dev || !Qhat_dev || !f_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    return do_finish(p, Q_dev, reinterpret_cast<const cplx *>(Qhat_dev), f_dev, (cudaStream_t)stream);
}

extern "C" int bfsm_collide(bfsm_plan *p, double *Q_dev, const double *f_dev, int n_cells, void *stream)
{
    if (!p || !Q_dev || !f_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (n_cells < 0) return fail(BFSM_ERR_INVALID, "n_cells must be >= 0");
    if (p->shard_count != 1)
        return fail(BFSM_ERR_INVALID,
                    "bfsm_collide needs an unsharded plan; use bfsm_collide_sharded (or bfsm_collide_partial / "
                    "bfsm_gain_hat + your own reduction)");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    const size_t N3 = grid_points(p);
    cudaStream_t st = (cudaStream_t)stream;
    p->lanes_used_last = 1;
    p->group_cells_last = 0;
    if (n_cells >= 2 && p->n_lanes >= 2 && !p->profiling && group_path_ok(p) && p->lane_fail_from <= 0) {
        // cell groups: up to GROUP_CELLS cells per launch, two groups in flight
        const int cg_max = std::min((int)bfsm_plan::GROUP_CELLS, n_cells);
        const int n_groups = (n_cells > cg_max) ? bfsm_plan::MAX_GROUPS : 1;
        int usable = 0;
        for (int k = 0; k < n_groups; ++k) {
            int rc = group_alloc(p, k, cg_max);
            if (rc == BFSM_ERR_NOMEM) break; // run with the groups that exist (or fall back to the lanes)
            if (rc) return rc;
            usable = k + 1;
        }
        if (usable > 0) {
            if (!p->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(p->ev_fork, st));
            for (int k = 0; k < usable; ++k) CUDA_TRY(cudaStreamWaitEvent(p->groups[k].main, p->ev_fork, 0));
            int gi = 0;
            for (int c = 0; c < n_cells; c += cg_max, ++gi) {
                const int cg = std::min(cg_max, n_cells - c);
                int rc = do_group(p, p->groups[gi % usable], Q_dev + (size_t)c * N3, f_dev + (size_t)c * N3, cg);
                if (rc) return rc;
            }
            for (int k = 0; k < usable; ++k) {
                CUDA_TRY(cudaEventRecord(p->groups[k].done, p->groups[k].main));
                CUDA_TRY(cudaStreamWaitEvent(st, p->groups[k].done, 0));
            }
            p->group_cells_last = cg_max;
            p->lanes_used_last = usable;
            return BFSM_OK;
        }
    }
    if (n_cells >= 2 && p->n_lanes >= 2 && !p->profiling) {
        // only as many lanes as there are cells, and only those that could be allocated
        int nl = 1;
        int rc = lanes_prepare(p, std::min(p->n_lanes, n_cells), &nl);
        if (rc) return rc;
        p->lanes_used_last = nl;
        CUDA_TRY(cudaEventRecord(p->ev_fork, st));
        for (int k = 0; k < nl; ++k) CUDA_TRY(cudaStreamWaitEvent(p->lanes[k].main, p->ev_fork, 0));
        for (int c = 0; c < n_cells && !rc; ++c) {
            const int k = c % nl;
            lane_activate(p, k);
            cudaStream_t ls = p->lanes[k].main;
            rc = do_gain_hat(p, p->qhat, f_dev + (size_t)c * N3, ls);
            if (!rc) rc = do_finish(p, Q_dev + (size_t)c * N3, p->qhat, f_dev + (size_t)c * N3, ls);
        }
        lane_activate(p, 0);
        for (int k = 0; k < nl; ++k) {
            CUDA_TRY(cudaEventRecord(p->lanes[k].done, p->lanes[k].main));
            CUDA_TRY(cudaStreamWaitEvent(st, p->lanes[k].done, 0));
        }
        return rc;
    }
    for (int c = 0; c < n_cells; ++c) {
        int rc = do_gain_hat(p, p->qhat, f_dev + (size_t)c * N3, st);
        if (rc) return rc;
        rc = do_finish(p, Q_dev + (size_t)c * N3, p->qhat, f_dev + (size_t)c * N3, st);
        if (rc) return rc;
    }
    return BFSM_OK;
}

extern "C" int bfsm_collide_host(bfsm_plan *p, double *Q_host, const double *f_host, int n_cells,
                                 void *stream)
{
    if (!p || !Q_host || !f_host) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (n_cells < 0) return fail(BFSM_ERR_INVALID, "n_cells must be >= 0");
    if (n_cells == 0) return BFSM_OK;
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    const size_t N3 = grid_points(p);
    const size_t bytes = sizeof(double) * N3 * (size_t)n_cells;
    if (p->stage_cells < (size_t)n_cells) {
        if (p->stage_f) cudaFree(p->stage_f);
        if (p->stage_q) cudaFree(p->stage_q);
        p->stage_f = p->stage_q = nullptr;
        p->stage_cells = 0;
        CUDA_TRY(cudaMalloc((void **)&p->stage_f, bytes));
        CUDA_TRY(cudaMalloc((void **)&p->stage_q, bytes));
        p->stage_cells = (size_t)n_cells;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(p->stage_f, f_host, bytes, cudaMemcpyHostToDevice, st));
    int rc = bfsm_collide(p, p->stage_q, p->stage_f, n_cells, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(Q_host, p->stage_q, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return BFSM_OK;
}

// Pipelined host-pointer evaluation: returns once the step is enqueued (and the step submitted two
// calls earlier has delivered its Q); bfsm_collide_host_flush waits for everything outstanding.
extern "C" int bfsm_collide_host_async(bfsm_plan *p, bfsm_comm *cm, double *Q_host, const double *f_host,
                                       int n_cells, void *stream)
{
    if (!p || !Q_host || !f_host) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (n_cells < 0) return fail(BFSM_ERR_INVALID, "n_cells must be >= 0");
    if (n_cells == 0) return BFSM_OK;
    if (p->shard_count > 1 && (n_cells != 1 || !cm))
        return fail(BFSM_ERR_INVALID, "a sharded plan evaluates one cell per call and needs a communicator");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    bfsm_plan::HostPipe &hp = p->pipe;
    const size_t N3 = grid_points(p);
    const size_t bytes = sizeof(double) * N3 * (size_t)n_cells;
    if (!hp.h2d) {
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp.d2h, cudaStreamNonBlocking));
        for (int b = 0; b < bfsm_plan::HostPipe::DEPTH; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&hp.in_done[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hp.comp_done[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hp.out_done[b], cudaEventDisableTiming));
        }
    }
    if (hp.cells < (size_t)n_cells) {
        CUDA_TRY(cudaDeviceSynchronize());
        for (int b = 0; b < bfsm_plan::HostPipe::DEPTH; ++b) {
            if (hp.f[b]) cudaFree(hp.f[b]);
            if (hp.q[b]) cudaFree(hp.q[b]);
            hp.f[b] = hp.q[b] = nullptr;
            hp.used[b] = false;
        }
        hp.cells = 0;
        for (int b = 0; b < bfsm_plan::HostPipe::DEPTH; ++b) {
            CUDA_TRY(cudaMalloc((void **)&hp.f[b], bytes));
            CUDA_TRY(cudaMalloc((void **)&hp.q[b], bytes));
        }
        hp.cells = (size_t)n_cells;
    }
    const int b = (int)(hp.submitted % bfsm_plan::HostPipe::DEPTH);
    cudaStream_t st = (cudaStream_t)stream;
    if (hp.used[b]) {
        // slot b was used by the step submitted DEPTH calls ago: its Q must have reached the host (that
        // also means its kernels are done with f[b] and q[b])
        CUDA_TRY(cudaEventSynchronize(hp.out_done[b]));
    }
    CUDA_TRY(cudaMemcpyAsync(hp.f[b], f_host, bytes, cudaMemcpyHostToDevice, hp.h2d));
    CUDA_TRY(cudaEventRecord(hp.in_done[b], hp.h2d));
    CUDA_TRY(cudaStreamWaitEvent(st, hp.in_done[b], 0));
    int rc = (p->shard_count > 1) ? bfsm_collide_sharded(p, cm, hp.q[b], hp.f[b], stream)
                                  : bfsm_collide(p, hp.q[b], hp.f[b], n_cells, stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(hp.comp_done[b], st));
    CUDA_TRY(cudaStreamWaitEvent(hp.d2h, hp.comp_done[b], 0));
    CUDA_TRY(cudaMemcpyAsync(Q_host, hp.q[b], bytes, cudaMemcpyDeviceToHost, hp.d2h));
    CUDA_TRY(cudaEventRecord(hp.out_done[b], hp.d2h));
    hp.used[b] = true;
    ++hp.submitted;
    return BFSM_OK;
}

extern "C" int bfsm_collide_host_flush(bfsm_plan *p)
{
    if (!p) return fail(BFSM_ERR_INVALID, "plan is NULL");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    for (int b = 0; b < bfsm_plan::HostPipe::DEPTH; ++b)
        if (p->pipe.used[b]) CUDA_TRY(cudaEventSynchronize(p->pipe.out_done[b]));
    return BFSM_OK;
}

extern "C" int bfsm_collide_profiled(bfsm_plan *p, double *Q_dev, const double *f_dev, void *stream,
                                     double *ms_by_class, int *launches_by_class)
{
    if (!p || !Q_dev || !f_dev || !ms_by_class || !launches_by_class)
        return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    p->profiling = true;
    p->profile_failed = false;
    p->spans.clear();
    p->events_used = 0;
    int rc = bfsm_collide(p, Q_dev, f_dev, 1, stream);
    p->profiling = false;
    if (rc) return rc;
    if (p->profile_failed) return fail(BFSM_ERR_CUDA, "a profiling event could not be created or recorded");
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < BFSM_KCLASS_COUNT; ++c) {
        ms_by_class[c] = 0.0;
        launches_by_class[c] = 0;
    }
    for (const auto &sp : p->spans) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms_by_class[sp.cls] += (double)ms;
        launches_by_class[sp.cls] += 1;
    }
    return BFSM_OK;
}

extern "C" int bfsm_sync(bfsm_plan *p, void *stream)
{
    if (!p) return fail(BFSM_ERR_INVALID, "plan is NULL");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return BFSM_OK;
}

extern "C" int bfsm_device_malloc(int device, void **ptr, unsigned long long bytes)
{
    if (!ptr) return fail(BFSM_ERR_INVALID, "ptr is NULL");
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
    return BFSM_OK;
}

extern "C" int bfsm_device_free(int device, void *ptr)
{
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaFree(ptr));
    return BFSM_OK;
}

extern "C" int bfsm_copy_to_device(int device, void *dst_dev, const void *src_host,
                                   unsigned long long bytes)
{
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaMemcpy(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice));
    return BFSM_OK;
}

extern "C" int bfsm_copy_to_host(int device, void *dst_host, const void *src_dev,
                                 unsigned long long bytes)
{
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    CUDA_TRY(cudaMemcpy(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost));
    return BFSM_OK;
}

// =========================================================================== multi-GPU (NCCL)
// The path's one exchange step: every rank evaluates its shard of the (r, sigma)
// pair list, turns its partial gain spectrum into a partial Q in physical space (the inverse transform
// is linear), rank 0 subtracts the loss term, and ONE ncclAllReduce of N^3 real doubles (2 MiB at
// 64^3, half of the complex spectrum) leaves Q(f,f) on every rank -- nothing runs after the collective.
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a PyTorch process that is the copy torch
// already loaded, in a plain C++ host the system library.
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    auto sym = [&](const char *name) { return dlsym(h, name); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllReduce ||
        !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
        return nullptr;
    api.handle = h;
    return &api;
}

#define NCCL_TRY(api, expr)                                                                    \
    do {                                                                                       \
        ncclResult_t r_ = (expr);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            char buf_[512];                                                                    \
            snprintf(buf_, sizeof buf_, "%s failed: %s", #expr, (api)->GetErrorString(r_));     \
            return fail(BFSM_ERR_COMM, buf_);                                                  \
        }                                                                                      \
    } while (0)

} // namespace

struct bfsm_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0, device = 0;
    bool owned = true;
};

static_assert(BFSM_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" int bfsm_comm_unique_id(unsigned char *id)
{
    if (!id) return fail(BFSM_ERR_INVALID, "id is NULL");
    NcclApi *api = nccl_api();
    if (!api) return fail(BFSM_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded");
    ncclUniqueId uid;
    NCCL_TRY(api, api->GetUniqueId(&uid));
    std::memcpy(id, uid.internal, BFSM_UNIQUE_ID_BYTES);
    return BFSM_OK;
}

extern "C" int bfsm_comm_init_rank(bfsm_comm **out, const unsigned char *id, int n_ranks, int rank, int device)
{
    if (!out || !id) return fail(BFSM_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(BFSM_ERR_INVALID, "rank / n_ranks out of range");
    NcclApi *api = nccl_api();
    if (!api) return fail(BFSM_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded");
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, BFSM_UNIQUE_ID_BYTES);
    ncclComm_t c = nullptr;
    NCCL_TRY(api, api->CommInitRank(&c, n_ranks, uid, rank));
    bfsm_comm *cm = new bfsm_comm;
    cm->comm = c; cm->n_ranks = n_ranks; cm->rank = rank; cm->device = device;
    *out = cm;
    return BFSM_OK;
}

extern "C" int bfsm_comm_init_all(bfsm_comm **comms, int n_devices, const int *devices)
{
    if (!comms || !devices || n_devices < 1) return fail(BFSM_ERR_INVALID, "bad argument");
    NcclApi *api = nccl_api();
    if (!api) return fail(BFSM_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded");
    std::vector<ncclComm_t> raw(n_devices, nullptr);
    NCCL_TRY(api, api->CommInitAll(raw.data(), n_devices, devices));
    for (int k = 0; k < n_devices; ++k) {
        bfsm_comm *cm = new bfsm_comm;
        cm->comm = raw[k]; cm->n_ranks = n_devices; cm->rank = k; cm->device = devices[k];
        comms[k] = cm;
    }
    return BFSM_OK;
}

extern "C" int bfsm_comm_adopt(bfsm_comm **out, void *nccl_comm, int n_ranks, int rank, int device)
{
    if (!out || !nccl_comm) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(BFSM_ERR_INVALID, "rank / n_ranks out of range");
    if (!nccl_api()) return fail(BFSM_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded");
    bfsm_comm *cm = new bfsm_comm;
    cm->comm = (ncclComm_t)nccl_comm; cm->n_ranks = n_ranks; cm->rank = rank; cm->device = device;
    cm->owned = false;
    *out = cm;
    return BFSM_OK;
}

extern "C" int bfsm_comm_destroy(bfsm_comm *cm)
{
    if (!cm) return BFSM_OK;
    NcclApi *api = nccl_api();
    if (cm->owned && cm->comm && api) {
        GuardDevice guard(cm->device);
        api->CommDestroy(cm->comm);
    }
    delete cm;
    return BFSM_OK;
}

namespace {
int check_shard(const bfsm_plan *p, const bfsm_comm *cm)
{
    if (!p || !cm) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (p->shard_count != cm->n_ranks || p->shard_index != cm->rank)
        return fail(BFSM_ERR_INVALID, "plan shard (index, count) does not match the communicator (rank, size)");
    if (p->device != cm->device) return fail(BFSM_ERR_INVALID, "plan and communicator live on different devices");
    return BFSM_OK;
}
// this rank's partial Q: gain of its pair shard in physical space, minus the loss term on rank 0
int sharded_partial(bfsm_plan *p, double *Q, const double *f, cudaStream_t st)
{
    int rc = do_gain_hat(p, p->qhat, f, st);
    if (rc) return rc;
    return do_finish(p, Q, p->qhat, f, st, /*with_loss=*/p->shard_index == 0);
}
} // namespace

extern "C" int bfsm_collide_sharded(bfsm_plan *p, bfsm_comm *cm, double *Q_dev, const double *f_dev, void *stream)
{
    if (!Q_dev || !f_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    int rc = check_shard(p, cm);
    if (rc) return rc;
    NcclApi *api = nccl_api();
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = sharded_partial(p, Q_dev, f_dev, st))) return rc;
    const size_t n = grid_points(p);
    NCCL_TRY(api, api->AllReduce(Q_dev, Q_dev, n, ncclDouble, ncclSum, cm->comm, st));
    return BFSM_OK;
}

// One host thread driving every rank (ncclCommInitAll): all local work is enqueued first, the
// collectives are issued as one NCCL group.
extern "C" int bfsm_collide_sharded_group(int n_ranks, bfsm_plan **plans, bfsm_comm **comms, double **Q_dev,
                                          const double **f_dev, void **streams)
{
    if (n_ranks < 1 || !plans || !comms || !Q_dev || !f_dev) return fail(BFSM_ERR_INVALID, "bad argument");
    NcclApi *api = nccl_api();
    if (!api) return fail(BFSM_ERR_COMM, "NCCL (libnccl.so.2) could not be loaded");
    for (int k = 0; k < n_ranks; ++k) {
        int rc = check_shard(plans[k], comms[k]);
        if (rc) return rc;
        if (!Q_dev[k] || !f_dev[k]) return fail(BFSM_ERR_INVALID, "NULL buffer");
        GuardDevice guard(plans[k]->device);
        if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
        if ((rc = sharded_partial(plans[k], Q_dev[k], f_dev[k], streams ? (cudaStream_t)streams[k] : nullptr)))
            return rc;
    }
    NCCL_TRY(api, api->GroupStart());
    for (int k = 0; k < n_ranks; ++k) {
        const size_t n = grid_points(plans[k]);
        ncclResult_t r = api->AllReduce(Q_dev[k], Q_dev[k], n, ncclDouble, ncclSum, comms[k]->comm,
                                        streams ? (cudaStream_t)streams[k] : nullptr);
        if (r != ncclSuccess) {
            api->GroupEnd();
            return fail(BFSM_ERR_COMM, std::string("ncclAllReduce failed: ") + api->GetErrorString(r));
        }
    }
    NCCL_TRY(api, api->GroupEnd());
    return BFSM_OK;
}

// Pair-shard partial without the collective: for callers that own the exchange step (tests emulating
// ranks on one GPU, other communication libraries).  Sum of Q_partial over all shards = Q(f,f).
extern "C" int bfsm_collide_partial(bfsm_plan *p, double *Q_partial_dev, const double *f_dev, void *stream)
{
    if (!p || !Q_partial_dev || !f_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    return sharded_partial(p, Q_partial_dev, f_dev, (cudaStream_t)stream);
}

// ---- callers' helpers: integrator update and per-cell moments ---------------------------------
extern "C" int bfsm_vec_axpby(int device, double *out_dev, double a, const double *x_dev, double b,
                              const double *y_dev, unsigned long long n, void *stream)
{
    if (!out_dev || !x_dev || !y_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    if (n == 0) return BFSM_OK;
    const int blocks = (int)std::min<unsigned long long>((n + 255) / 256, 148ull * 8);
    k_axpby<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, a, x_dev, b, y_dev, (size_t)n);
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

extern "C" int bfsm_moments(bfsm_plan *p, const double *g_dev, int n_cells, double *moments_dev, void *stream)
{
    if (!p || !g_dev || !moments_dev) return fail(BFSM_ERR_INVALID, "NULL argument");
    if (n_cells < 0) return fail(BFSM_ERR_INVALID, "n_cells must be >= 0");
    GuardDevice guard(p->device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    if (n_cells == 0) return BFSM_OK;
    k_moments<512><<<n_cells, 512, 0, (cudaStream_t)stream>>>(g_dev, p->nx, p->ny, p->nz, p->L, moments_dev);
    CUDA_TRY(cudaGetLastError());
    return BFSM_OK;
}

// ---- FP64 pipe peak (measurement aid): independent DFMA chains, no memory traffic ------------
namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double seed)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + 1e-3 * (threadIdx.x + k);
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.6789) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // keep the chains alive
}
} // namespace

extern "C" int bfsm_measure_fp64_peak(int device, double *dfma_per_second)
{
    if (!dfma_per_second) return fail(BFSM_ERR_INVALID, "NULL argument");
    GuardDevice guard(device);
    if (!guard.ok) return fail(BFSM_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0 + rep);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double n = (double)blocks * threads * (double)iters * 64.0; // DFMA per launch
        if (rep > 0) best = std::max(best, n / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *dfma_per_second = best;
    return BFSM_OK;
}
