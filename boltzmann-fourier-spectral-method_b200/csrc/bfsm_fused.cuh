// bfsm_fused.cuh -- the whole gain loop of one evaluation as ONE persistent, spatially partitioned
// kernel (N = 64, packed mode).
//
// The stand-alone kernels hand the hybrid grid Z(i; y, z) of every pair (4 MiB at 64^3) from the plane
// stage to the x stage through HBM, one kernel after the other: 8.4 MB of DRAM traffic per pair, the
// LSU-bound plane kernel and the HBM-bound x kernel never overlap.  Here the SMs are split into roles
// (one CTA per SM, cooperative launch, so every CTA is resident):
//
//   CTAs [0, n_plane)                 plane role: the warp-specialised plane pipeline of
//                                     k_plane_gain_ws over the 64 regular (y,z) planes
//   CTAs [n_plane, n_plane + n_nyq)   Nyquist role: the same pipeline over the 3 Nyquist planes of
//                                     fhat (-> uvw, consumed by k_nyq_accum after this kernel)
//   the remaining CTAs                pencil role: the register-resident x stage (bfsm_pencil_reg.cuh),
//                                     8 warps per CTA at 232 registers (setmaxnreg), the loads of the
//                                     next pair in flight while the current one is transformed
//
// The pair list of the launch is cut into SUB-CHUNKS of `sub_pairs` pairs.  The plane role writes
// sub-chunk s into slot s % ring of a ring of hybrid grids sized to stay inside the 126 MB L2; when a
// plane CTA has issued its last store of sub-chunk s it publishes it (fence + ready[s] += 1); a pencil
// warp starts on sub-chunk s when ready[s] == n_plane, and adds the pairs it has finished to
// consumed[s]; a plane CTA may overwrite a ring slot once consumed[s - ring] has reached
// pairs x warp tiles.  All waits are on strictly older sub-chunks, so the schedule cannot deadlock as
// long as all CTAs are co-resident -- which the cooperative launch guarantees.
//
// Every plane CTA keeps working on the same one or two planes in every sub-chunk (its share of the
// plane-major entry list is the same every time) and walks its share back and forth (snake order), so
// that the fhat plane cached in the registers of its S1 warpgroup changes at most once per sub-chunk.
#pragma once
#include "bfsm_kernels.cuh"
#include "bfsm_pencil_reg.cuh"

namespace bfsm {

struct FusedParams {
    int n_plane, n_nyq, n_pencil; // role sizes (n_nyq is a multiple of 3)
    int pair0, n_pairs;           // plan-local pair range of this launch
    int n_units;                  // pencil work units of this launch (each inside one sub-chunk)
    RingSync rs;
};

// Regular planes: sub-chunk s holds N x K_s entries (plane-major); CTA `cta` owns the same fraction
// [cta/n, (cta+1)/n) of every sub-chunk, walked upwards in even and downwards in odd sub-chunks.
template <int N> struct FusedPlaneWalk {
    int cta, n_ctas, K, n_pairs, pair0, ring, n_sub;
    int s, Ks, idx, left, dir;
    int cnt;
    int i, pair, dst_item, sub;
    __host__ __device__ __forceinline__ void range(int s_, int &lo, int &hi, int &ks) const
    {
        ks = min(K, n_pairs - s_ * K);
        const long long tot = (long long)N * ks;
        lo = (int)((tot * cta) / n_ctas);
        hi = (int)((tot * (cta + 1)) / n_ctas);
    }
    __host__ __device__ __forceinline__ void enter(int s_)
    {
        // first sub-chunk >= s_ in which this CTA has entries
        for (s = s_; s < n_sub; ++s) {
            int lo, hi;
            range(s, lo, hi, Ks);
            if (hi > lo) {
                left = hi - lo;
                dir = (s & 1) ? -1 : +1;
                idx = dir > 0 ? lo : hi - 1;
                load();
                return;
            }
        }
        left = 0;
        sub = n_sub;
    }
    __host__ __device__ __forceinline__ void load()
    {
        i = idx / Ks;
        const int it = idx - i * Ks;
        pair = pair0 + s * K + it;
        dst_item = (s % ring) * K + it;
        sub = s;
    }
    __host__ __device__ __forceinline__ void init(int cta_, int n_ctas_, int K_, int n_pairs_, int pair0_,
                                                  int ring_)
    {
        cta = cta_; n_ctas = n_ctas_; K = K_; n_pairs = n_pairs_; pair0 = pair0_; ring = ring_;
        n_sub = (n_pairs + K - 1) / K;
        cnt = 0;
        for (int t = 0; t < n_sub; ++t) {
            int lo, hi, ks;
            range(t, lo, hi, ks);
            cnt += hi - lo;
        }
        enter(0);
    }
    __host__ __device__ __forceinline__ void next()
    {
        if (--left > 0) { idx += dir; load(); }
        else enter(s + 1);
    }
};

// Nyquist planes: CTA (q, h) of 3 x c owns plane N + q and the items [K_s h / c, K_s (h+1) / c) of every
// sub-chunk; uvw holds the whole launch (no ring: it is read after the kernel).
template <int N> struct FusedNyqWalk {
    int q, h, c, K, n_pairs, pair0, n_sub;
    int s, it, it_hi;
    int cnt;
    int i, pair, dst_item, sub;
    __host__ __device__ __forceinline__ void range(int s_, int &lo, int &hi) const
    {
        const int ks = min(K, n_pairs - s_ * K);
        lo = (ks * h) / c;
        hi = (ks * (h + 1)) / c;
    }
    __host__ __device__ __forceinline__ void enter(int s_)
    {
        for (s = s_; s < n_sub; ++s) {
            range(s, it, it_hi);
            if (it_hi > it) { load(); return; }
        }
        sub = n_sub;
    }
    __host__ __device__ __forceinline__ void load()
    {
        i = N + q;
        pair = pair0 + s * K + it;
        dst_item = s * K + it;
        sub = s;
    }
    __host__ __device__ __forceinline__ void init(int cta, int n_ctas, int K_, int n_pairs_, int pair0_)
    {
        q = cta % 3; h = cta / 3; c = n_ctas / 3; K = K_; n_pairs = n_pairs_; pair0 = pair0_;
        n_sub = (n_pairs + K - 1) / K;
        cnt = 0;
        for (int t = 0; t < n_sub; ++t) {
            int lo, hi;
            range(t, lo, hi);
            cnt += hi - lo;
        }
        enter(0);
    }
    __host__ __device__ __forceinline__ void next()
    {
        if (++it < it_hi) load();
        else enter(s + 1);
    }
};

constexpr int FUSED_THREADS = 384, FUSED_WARPS = FUSED_THREADS / 32;

template <int N, bool UNIFORM_W>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_gain_fused(FusedParams fp, const cplx *__restrict__ fhat, const cplx *__restrict__ phase,
             const cplx *__restrict__ zpm, const cplx *__restrict__ twtab, cplx *__restrict__ ring_buf,
             const cplx *__restrict__ nyq, const double *__restrict__ pair_w, cplx *__restrict__ uvw,
             const PencilUnit *__restrict__ units, double *__restrict__ S, int n_r_local)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr size_t N2 = (size_t)N * N, N3 = N2 * N;
    const int cta = blockIdx.x;
    if (cta < fp.n_plane) {
        FusedPlaneWalk<N> walk;
        walk.init(cta, fp.n_plane, fp.rs.sub_pairs, fp.n_pairs, fp.pair0, fp.rs.ring);
        plane_ws_pipeline<N, true>(walk, smem_raw, fhat, phase, zpm, twtab, ring_buf, nyq, pair_w, uvw, fp.rs);
    } else if (cta < fp.n_plane + fp.n_nyq) {
        FusedNyqWalk<N> walk;
        walk.init(cta - fp.n_plane, fp.n_nyq, fp.rs.sub_pairs, fp.n_pairs, fp.pair0);
        plane_ws_pipeline<N, false>(walk, smem_raw, fhat, phase, zpm, twtab, ring_buf, nyq, pair_w, uvw, fp.rs);
    } else {
        // =========================== pencil role =================================================
        // warps 0-7 work with 232 registers each (two pairs of a warp tile in registers: one in flight,
        // one being transformed), warps 8-11 hand their registers over and leave
        constexpr int P_WARPS = 8;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (warp >= P_WARPS) {
            reg_dealloc<40>();
            return;
        }
        reg_alloc<232>();
        constexpr int WT = PencilGeo<N>::WT;
        const int n_warps = fp.n_pencil * P_WARPS;
        const int w0 = (cta - fp.n_plane - fp.n_nyq) * P_WARPS + warp;
        const long long total = (long long)fp.n_units * WT;
        const int K = fp.rs.sub_pairs;
        for (long long g = w0; g < total; g += n_warps) {
            const PencilUnit un = units[g / WT];
            PencilLane<N> L;
            L.init(lane, (int)(g % WT));
            const int rel = un.p0 - fp.pair0, s = rel / K;
            // relaxed polls, one fence once the sub-chunk is there (an acquire load per poll would
            // invalidate L1 every time: CCTL.IVALL)
            while (ld_relaxed(fp.rs.ready + s) < fp.n_plane) __nanosleep(64);
            __threadfence();
            const cplx *src = ring_buf + ((size_t)(s % fp.rs.ring) * K + (rel - s * K)) * N3 + L.first_entry();
            double acc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 0.0;
            const int np = un.p1 - un.p0;
            auto wq = [&](int q) { return UNIFORM_W ? 1.0 : __ldg(&pair_w[un.p0 + q]); };
            // two register buffers in rotation; loads past the end of the unit are clamped to its last
            // pair (a redundant L2 hit) so that no load is conditional
            auto at = [&](int q) { return src + (size_t)min(q, np - 1) * N3; };
            cplx va[16], vb[16];
            pencil_reg_load<N>(at(0), va);
            for (int q = 0; q < np; q += 2) {
                pencil_reg_load<N>(at(q + 1), vb);
                pencil_reg_compute<N, UNIFORM_W>(va, L.c, L.d, L.tw, wq(q), acc);
                if (q + 1 >= np) break;
                pencil_reg_load<N>(at(q + 2), va);
                pencil_reg_compute<N, UNIFORM_W>(vb, L.c, L.d, L.tw, wq(q + 1), acc);
            }
            // the ring slot may be overwritten once every warp has reported its pairs
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(fp.rs.consumed + s, np);
            const double scale = UNIFORM_W ? __ldg(&pair_w[un.p0]) : 1.0;
            double *Sr = S + ((size_t)un.slot * n_r_local + un.r) * N3 + (size_t)L.y * N + L.z;
#pragma unroll
            for (int j = 0; j < 16; ++j) Sr[(size_t)pencil_out_x<N>(j, L.c, L.d) * N2] = acc[j] * scale;
        }
    }
}

} // namespace bfsm
