// bfsm_plane_r32.cuh -- (y,z) stage of the gain loop with RADIX-32 register transforms: two stages and
// ONE shared-memory exchange per plane instead of three stages and two exchanges (k_plane_gain_ws /
// k_plane_gain3).  Replaces compute_alpha_times_f_hat + the (y,z) part of the batched inverse cuFFTs
// (CUDABoltzmannOperator.cu:144-164) like those kernels; same inputs, same hybrid-grid layout.
//
// Why: the three-stage kernels are bound by the shared-memory data path, not by HBM or FP64 (ncu:
// l1tex data pipe 77 %, 3100 wavefronts per 64 x 64 plane: 4 shared-memory accesses + the global
// store per element + the z phase table).  A length-32 line fits the registers of one thread
// (Dft<32>, 64 data registers), so:
//
//   N = 32: a WARP owns a plane item.  Stage A: lane j holds row j of the fhat plane in registers for
//           as long as the plane does not change, applies the real multiplier, radix-32 along z, stores
//           the row.  __syncwarp.  Stage B: lane z loads column z, radix-32 along y, natural-order global
//           store (512 contiguous bytes per warp and row).  No block-level barrier at all; per element
//           1 STS + 1 LDS + 1 STG (k_plane_gain3: 5 shared-memory accesses + 1 STG).
//   N = 64: a group of 128 threads owns a plane item; lanes l and l ^ 16 of a warp share a line: lane
//           bit h = 0 transforms the even entries, h = 1 the odd ones (radix-32 each), and the last
//           radix-2 step runs ACROSS the two lanes: each lane sends 16 of its 32 values with SHFL
//           (4 x 32 bit per value) and ends up with outputs 16h..16h+15 and 32+16h..32+16h+15.
//           Per element 1 STS + 1 LDS + 1 STG + 2 SHFL.32 (a SHFL.32 costs a quarter of an LDS.128 on
//           the same pipe: profiles/r02_microbench.log) -- about 2200 wavefronts per plane.
//
// Cross-lane step without a send-side select: lane h = 1 feeds its radix-32 with (-1)^a in[2a+1], which
// rotates its outputs by 16 (slot r holds Y1[(r+16) mod 32]); both lanes then SEND slot 16+i and KEEP
// slot i.  The signs cost nothing: along z they are folded into the (Re+Im, Re-Im) table of the z phase
// on the host (entries k = 3 mod 4 negated, `zpm_r32`), along y into the per-row multiplier of stage A
// (rows j = 3 mod 4 negated).  Lane h owns outputs z1 = 16h + i:
//     out[z1]      = Y0[z1] + W64^z1 Y1[z1]        W64^(16h+i) = i^h W64^i  (compile-time W64^i)
//     out[z1 + 32] = Y0[z1] - W64^z1 Y1[z1]
//
// Every group walks its own contiguous share of the flat (plane, item) list (R32Walk), groups never
// synchronise with each other.
#pragma once
#include "bfsm_kernels.cuh"
#include "bfsm_tmem.cuh"

namespace bfsm {

template <int N> struct R32Geo {
    static_assert(N == 32 || N == 64, "radix-32 plane kernel: N = 32 or 64");
    static constexpr int GT = N * N / 32;          // threads of a group: one radix-32 unit per thread and stage
    static constexpr int PITCH = N + 1;            // row pitch of the plane buffer (== 16 bytes mod 128)
    static constexpr int SMEM_CPLX = N * PITCH + 2 * 4 * N; // plane buffer + two phase-table slots
};

// Work list of one group: the flat (plane, item) list has N*n_items regular entries followed by 3*n_items
// Nyquist-plane entries (about 1.3 x the cost each: their multiplier is computed per element).  Groups
// [0, gA) share the regular entries and groups [gA, n_groups) the Nyquist entries, in proportion to cost,
// every group ONE contiguous range -- with a thousand groups per launch an equal share of both classes
// for everybody (LaunchWalk) would make every group reload two or three planes for a couple of entries.
template <int N> struct R32Walk {
    int i, it, n_items, cnt;
    __host__ __device__ __forceinline__ void init(int n_items_, int grp, int n_groups)
    {
        n_items = n_items_;
        const long long totA = (long long)N * n_items, totB = (long long)3 * n_items;
        int gB = (int)(((long long)n_groups * 39 + (10 * N + 39) / 2) / (10 * N + 39)); // 3 * 1.3 : N
        if (gB < 1) gB = 1;
        if (gB > (int)totB) gB = (int)totB;
        long long lo, hi;
        if (n_groups < 2 || totB == 0) { // a single group walks the whole list
            lo = grp == 0 ? 0 : totA + totB;
            hi = totA + totB;
        } else {
            if (gB > n_groups - 1) gB = n_groups - 1;
            const int gA = n_groups - gB;
            if (grp < gA) { lo = (totA * grp) / gA; hi = (totA * (grp + 1)) / gA; }
            else { lo = totA + (totB * (grp - gA)) / gB; hi = totA + (totB * (grp - gA + 1)) / gB; }
        }
        cnt = (int)(hi - lo);
        i = (int)(lo / n_items);
        it = (int)(lo % n_items);
    }
    __host__ __device__ __forceinline__ void next()
    {
        if (++it == n_items) { it = 0; ++i; }
    }
};

template <int N, int G> constexpr size_t plane_r32_smem() { return sizeof(cplx) * (size_t)G * R32Geo<N>::SMEM_CPLX; }

template <int N> __device__ __forceinline__ void r32_group_sync(int g)
{
    if constexpr (N == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(R32Geo<N>::GT) : "memory");
}

// One cross-lane radix-2 butterfly of a 64-point line (see the header): `keep` = slot I, `send` = slot
// 16+I of this lane's rotated radix-32 result; returns out[16h+I] and out[16h+I+32].
// I is a compile-time constant at every call site (unrolled loops).
__device__ __forceinline__ void r2_cross(const int I, const cplx keep, const cplx send, const bool h,
                                         const double sg, cplx &o_lo, cplx &o_hi)
{
    cplx recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
    const cplx Pk = h ? recv : keep; // Y0[16h+I]
    const cplx Mk = h ? keep : recv; // Y1[16h+I]
    cplx T = Mk;
    if (I != 0) T = cmul(Mk, w64<+1>(I));
    // multiply by i^h: (tr, ti) = h ? (-T.y, T.x) : (T.x, T.y); sg = h ? -1 : +1 carries the sign
    const double tr = h ? T.y : T.x, ti = h ? T.x : T.y;
    o_lo = make_double2(fma(sg, tr, Pk.x), Pk.y + ti);
    o_hi = make_double2(fma(-sg, tr, Pk.x), Pk.y - ti);
}

// TM: keep the thread's entries of the fhat plane in tensor memory (bfsm_tmem.cuh) instead of 64
// registers -- 168 registers per thread are then enough without spills, i.e. 12 warps per SM instead of 8.
template <int N, int G, int MINB, bool TM>
__global__ void __launch_bounds__(G * R32Geo<N>::GT, MINB)
k_plane_gain_r32(const cplx *__restrict__ fhat, const cplx *__restrict__ phase,
                 const cplx *__restrict__ zpm, cplx *__restrict__ hyb, int pair0, int n_items,
                 const cplx *__restrict__ nyq, const double *__restrict__ pair_w,
                 cplx *__restrict__ uvw, int pairs_per_cell = 0)
{
    // Batch of cells: the launch's n_items = cells x pairs_per_cell; item `it` belongs to cell it / pairs_per_cell
    // (own fhat and Nyquist planes) and to pair pair0 + it % pairs_per_cell (phase tables, weight); the
    // hybrid grids and Nyquist fields of the launch are stored by item index.  pairs_per_cell = 0: one cell.
    const int ppc = pairs_per_cell > 0 ? pairs_per_cell : n_items;
    constexpr int GT = R32Geo<N>::GT, PITCH = R32Geo<N>::PITCH, H = N / 2;
    constexpr bool X2 = (N == 64); // two lanes share a line
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / GT, t = threadIdx.x % GT;
    cplx *buf = reinterpret_cast<cplx *>(smem_raw) + (size_t)g * R32Geo<N>::SMEM_CPLX; // N x PITCH
    cplx *phs = buf + N * PITCH;                                                        // 2 x 4N

    const int lane = t & 31;
    const bool h = X2 ? ((lane >> 4) & 1) != 0 : false;
    const int hb = h ? 1 : 0;
    const int line = X2 ? 16 * (t >> 5) + (lane & 15) : lane; // row in stage A, column in stage B
    const double sg = h ? -1.0 : 1.0;
    // rows that feed an odd-indexed input of lane 1's radix-32 along y are stored negated
    const double ysign = (X2 && (line & 3) == 3) ? -1.0 : 1.0;

    R32Walk<N> wk;
    wk.init(n_items, blockIdx.x * G + g, gridDim.x * G);
    const int cnt = wk.cnt;

    auto stage_phase = [&](int pair_src, int slot_dst) {
        const cplx *src = phase + (size_t)pair_src * 3 * N;
        const cplx *srz = zpm + (size_t)pair_src * N;
        cplx *d = phs + slot_dst * 4 * N;
#pragma unroll
        for (int e = t; e < 3 * N; e += GT) cp_async16(d + e, src + e);
#pragma unroll
        for (int e = t; e < N; e += GT) cp_async16(d + 3 * N + e, srz + e);
    };

    // this thread's entries of the current fhat plane, fr[a] = plane[line][X2 ? 2a+h : a]: registers, or
    // 128 words of the thread's TMEM lane (4 words per entry)
    cplx fr[TM ? 1 : 32];
    constexpr unsigned TM_COLS_USED = 128u * ((G * GT / 32 + 3) / 4);
    constexpr unsigned TM_COLS = TM_COLS_USED <= 128 ? 128 : TM_COLS_USED <= 256 ? 256 : 512;
    static_assert(!TM || TM_COLS_USED <= 512, "tensor memory: at most 16 warps per CTA");
    __shared__ unsigned tmem_slot;
    unsigned taddr = 0;
    if constexpr (TM) {
        if (threadIdx.x < 32) tmem_alloc(&tmem_slot, TM_COLS);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        const unsigned wcta = threadIdx.x >> 5;
        taddr = tmem_slot + (((wcta & 3u) * 32u) << 16) + (wcta >> 2) * 128u;
    }
    // entries 8c .. 8c+7 of the cached line
    auto load_chunk = [&](int c, cplx (&f8)[8]) {
        if constexpr (TM) {
            unsigned w[32];
            tmem_ld32(taddr + 32u * c, w);
            tmem_wait_ld32(w);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                f8[e] = make_double2(__hiloint2double((int)w[4 * e + 1], (int)w[4 * e]),
                                     __hiloint2double((int)w[4 * e + 3], (int)w[4 * e + 2]));
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f8[e] = fr[8 * c + e];
        }
    };
    int cur_plane = -1, slot = 0;
    if (cnt > 0) stage_phase(pair0 + wk.it % ppc, 0);
    cp_async_commit();
    cp_async_wait<0>();
    r32_group_sync<N>(g);

    for (int n = 0; n < cnt; ++n) {
        const int i = wk.i, dst_item = wk.it, cell = wk.it / ppc, pair = pair0 + wk.it % ppc;
        const int plane_key = cell * (N + 3) + i;
        wk.next(); // next entry: its tables go to the other slot while this one is computed

        if (plane_key != cur_plane) {
            // new plane: coalesced copy into the (free) plane buffer, then every thread picks its line
            // (all N*N/GT = 32 asynchronous 16-byte copies of a thread in flight at once: one round trip)
            const cplx *srcp = (i < N) ? fhat + ((size_t)cell * N + i) * N * N
                                       : nyq + ((size_t)cell * 3 + (i - N)) * N * N;
#pragma unroll
            for (int e = t; e < N * N; e += GT) cp_async16(&buf[(e / N) * PITCH + (e % N)], &srcp[e]);
            cp_async_commit();
            cp_async_wait<0>();
            r32_group_sync<N>(g);
            if constexpr (TM) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    unsigned w[32];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const cplx f = buf[line * PITCH + (X2 ? 2 * (8 * c + e) + hb : 8 * c + e)];
                        w[4 * e] = (unsigned)__double2loint(f.x);
                        w[4 * e + 1] = (unsigned)__double2hiint(f.x);
                        w[4 * e + 2] = (unsigned)__double2loint(f.y);
                        w[4 * e + 3] = (unsigned)__double2hiint(f.y);
                    }
                    tmem_st32(taddr + 32u * c, w);
                }
                tmem_wait_st();
            } else {
#pragma unroll
                for (int a = 0; a < 32; ++a) fr[a] = buf[line * PITCH + (X2 ? 2 * a + hb : a)];
            }
            r32_group_sync<N>(g);
            cur_plane = plane_key;
        }
        if (n + 1 < cnt) stage_phase(pair0 + wk.it % ppc, slot ^ 1);
        cp_async_commit();
        const cplx *P = phs + slot * 4 * N;

        // ---------------- stage A: real multiplier, radix-32 along z (+ cross-lane radix-2), row store
        cplx v[32];
        {
            const int j = line;
            if (i < N) {
                // m_H = A (Z.x+Z.y) + B (Z.x-Z.y), see k_plane_gain3; row 3 of P holds the two sums
                const cplx exi = P[i], eyj = P[N + j];
                const cplx X = cmul(exi, eyj);
                const cplx ext = (i == H) ? exi : make_double2(exi.x, -exi.y);
                const cplx eyt = (j == H) ? eyj : make_double2(eyj.x, -eyj.y);
                const cplx Xt = cmul(ext, eyt);
                const double hs = 0.5 * ysign;
                const double cA = hs * (X.x + Xt.x), cB = hs * (X.y - Xt.y);
                const double nA = hs * (X.x - Xt.y), nB = hs * (X.y + Xt.x);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    cplx f8[8];
                    load_chunk(c, f8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int a = 8 * c + e;
                        const int k = X2 ? 2 * a + hb : a;
                        const cplx zz = P[3 * N + k];
                        const bool ny = X2 ? ((a == H / 2) && !h) : (a == H); // k == H
                        const double m = (ny ? nA : cA) * zz.x + (ny ? nB : cB) * zz.y;
                        v[a] = make_double2(m * f8[e].x, m * f8[e].y);
                    }
                }
            } else {
                // Nyquist plane q = i - N: fixed axis q, free axes (axA rows, axB columns)
                const int nq = i - N;
                const int axA = (nq == 0) ? 1 : 0, axB = (nq == 2) ? 1 : 2;
                const cplx efix = P[nq * N + H];
                const double sw = 0.5 * ysign * sqrt(__ldg(&pair_w[pair]));
                const cplx ea = P[axA * N + j];
                const cplx eat = (j == H) ? ea : make_double2(ea.x, -ea.y);
                const cplx fa = cmul(efix, ea), fat = cmul(efix, eat);
                const bool zero_row = (nq >= 1) && (j == H);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    cplx f8[8];
                    load_chunk(c, f8);
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int a = 8 * c + e8;
                        const int k = X2 ? 2 * a + hb : a;
                        const cplx eb = P[axB * N + k];
                        const cplx ebt = (k == H) ? eb : make_double2(eb.x, -eb.y);
                        const cplx e = cmul(fa, eb), et = cmul(fat, ebt);
                        double n2 = sw * ((e.x - et.x) + (e.y + et.y));
                        if (zero_row || (nq == 2 && k == H)) n2 = 0.0;
                        if (X2 && (a & 1) && h) n2 = -n2; // rotation of lane 1's radix-32 (the table carries it above)
                        v[a] = make_double2(n2 * f8[e8].x, n2 * f8[e8].y);
                    }
                }
            }
            Dft<32, +1>::run(v);
            cplx *row = buf + j * PITCH;
            if constexpr (X2) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    cplx lo, hi;
                    r2_cross(q, v[dft32_reg(q)], v[dft32_reg(16 + q)], h, sg, lo, hi);
                    row[16 * hb + q] = lo;
                    row[16 * hb + q + 32] = hi;
                }
            } else {
#pragma unroll
                for (int z = 0; z < 32; ++z) row[z] = v[dft32_reg(z)];
            }
        }
        r32_group_sync<N>(g);

        // ---------------- stage B: radix-32 along y (+ cross-lane radix-2), natural-order global store
        {
            const cplx *col = buf + line;
#pragma unroll
            for (int a = 0; a < 32; ++a) v[a] = col[(X2 ? 2 * a + hb : a) * PITCH];
            Dft<32, +1>::run(v);
            cplx *dst = ((i < N) ? hyb + ((size_t)dst_item * N + i) * N * N
                                 : uvw + ((size_t)dst_item * 3 + (i - N)) * N * N) + line;
            if constexpr (X2) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    cplx lo, hi;
                    r2_cross(q, v[dft32_reg(q)], v[dft32_reg(16 + q)], h, sg, lo, hi);
                    dst[(16 * hb + q) * N] = lo;
                    dst[(16 * hb + q + 32) * N] = hi;
                }
            } else {
#pragma unroll
                for (int y = 0; y < 32; ++y) dst[y * N] = v[dft32_reg(y)];
            }
        }
        cp_async_wait<0>();
        r32_group_sync<N>(g); // next tables visible; the plane buffer is free again
        slot ^= 1;
    }
    if constexpr (TM) {
        tmem_fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tmem_dealloc(tmem_slot, TM_COLS);
    }
}

} // namespace bfsm
