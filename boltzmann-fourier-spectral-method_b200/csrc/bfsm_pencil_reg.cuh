// bfsm_pencil_reg.cuh -- the x stage of the gain path with NO shared memory: every x line lives in
// the registers of 1, 2 or 4 lanes of a warp.
//
//   hybrid grid Z(i; y, z) of pair q  (i = x-spectral index, written by the plane stage)
//     --> inverse x-FFT --> Z_H(x, y, z) --> acc(x,y,z) += w_q (Re^2 - Im^2)        (cpp:229-246)
//
// LANES = N/16 lanes share the line (y, z).  Lane s loads the 16 entries i = LANES a + s straight
// from global memory (no staging: LDG.128, L1 bypassed), does a radix-16 DFT in registers (all
// twiddles compile-time constants) and then log2(LANES) radix-2 steps ACROSS lanes, decimation in
// time: before a step every lane of a couple holds the same 2M frequencies k of its own
// sub-sequence; the couple swaps halves (4 SHFL per complex value) and each lane finishes
//
//   X[k], X[k + P] = Z_0[k] +- W_2P^k Z_1[k]        for its half of the k set (P = length so far).
//
// A warp covers 32/LANES consecutive z of one y (every load instruction reads LANES runs of at
// least 128 bytes) and walks the pairs [p0, p1) of one WORK UNIT, all of one radius, with the 16
// accumulators per lane in registers.  At the end it STORES them into partial slot `slot` of S_r --
// every (slot, radius, line) is written by exactly one warp, so there is no read-modify-write, no
// atomics, nothing to clear, and the result does not depend on scheduling.
//
// Shared-memory-path cost per element: one LDG.128 + at most one 4-SHFL exchange, against LDGSTS +
// LDS + STS + LDS of the staged kernel (k_pencil_gain_async), which turns LSU bound as soon as it is
// no longer HBM bound (measured: STS.128 + LDS.128 cost 8 cycles per warp, a SHFL 1 cycle, on the
// same pipe -- tools/microbench.cu).
#pragma once
#include "bfsm_fft.cuh"

namespace bfsm {

struct PencilUnit {
    int p0, p1; // plan-local pair range, one radius
    int r;      // local radius index
    int slot;   // partial slot of S_r this unit owns
};

__device__ __forceinline__ cplx ld_cg(const cplx *p)
{
    cplx v;
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ cplx shfl_xor_c(cplx a, int mask)
{
    return make_double2(__shfl_xor_sync(0xffffffffu, a.x, mask), __shfl_xor_sync(0xffffffffu, a.y, mask));
}

__device__ __forceinline__ cplx mul_i(cplx a) { return make_double2(-a.y, a.x); }

template <int N> struct PencilGeo {
    static constexpr int LANES = N / 16;        // lanes per x line
    static constexpr int ZW = 32 / LANES;       // z values (lines) per warp ...
    static constexpr int ZRUN = ZW < N ? ZW : N; // ... of which this many are consecutive in z
    static constexpr int YW = ZW / ZRUN;        // rows per warp (2 for N = 16)
    static constexpr int WT = N * N / ZW;       // warp tiles per grid
};

// x-index of accumulator `j` of the lane with sub-sequence bits (c, d); see pencil_reg_pair
template <int N> __device__ __forceinline__ int pencil_out_x(int j, int c, int d)
{
    if (N == 16) return j;
    if (N == 32) return 8 * c + (j >> 1) + 16 * (j & 1);
    return 8 * c + 4 * d + (j >> 2) + 16 * ((j >> 1) & 1) + 32 * (j & 1);
}

// Loads this lane's 16 entries of one pair: `src` points at its first entry (i = its sub-sequence
// index) of the pair's hybrid grid.
template <int N> __device__ __forceinline__ void pencil_reg_load(const cplx *__restrict__ src, cplx (&v)[16])
{
    constexpr int LANES = PencilGeo<N>::LANES;
    constexpr size_t N2 = (size_t)N * N;
#pragma unroll
    for (int a = 0; a < 16; ++a) v[a] = ld_cg(src + (size_t)(LANES * a) * N2);
}

// Transforms the 16 loaded entries and accumulates.  c, d are the lane's sub-sequence bits
// (i = LANES a + 2 c + d for N = 64, 2 a + c for N = 32), tw[m] = W_64^(8c + 4d + m) (N = 64 only).
template <int N, bool UNIFORM_W>
__device__ __forceinline__ void pencil_reg_compute(cplx (&v)[16], int c, int d, const cplx (&tw)[4],
                                                   double w, double (&acc)[16])
{
    constexpr int LANES = PencilGeo<N>::LANES;
    Dft<16, +1>::run(v);
    auto add = [&](int j, cplx o) {
        if (UNIFORM_W) {
            acc[j] = fma(o.x, o.x, acc[j]);
            acc[j] = fma(-o.y, o.y, acc[j]);
        } else {
            acc[j] = fma(w, o.x * o.x - o.y * o.y, acc[j]);
        }
    };
    // step over c (couple = lanes 8*(LANES/2) apart): lane c finishes k1 in [8c, 8c + 8):
    // Z[8c + m], Z[8c + m + 16] = Y_0[k1] +- W_32^k1 Y_1[k1],  W_32^(8c + m) = i^c W_32^m
    auto step_c = [&](int m, cplx &zlo, cplx &zhi) {
        const cplx lo = v[dft16_reg(m)], hi = v[dft16_reg(8 + m)];
        const cplx keep = c ? hi : lo;
        const cplx recv = shfl_xor_c(c ? lo : hi, LANES == 4 ? 8 : 16);
        const cplx y0 = c ? recv : keep, y1 = c ? keep : recv;
        cplx t = cmul(y1, w64<+1>(2 * m));
        if (c) t = mul_i(t);
        zlo = cadd(y0, t);
        zhi = csub(y0, t);
    };
    if constexpr (LANES == 1) {
#pragma unroll
        for (int k = 0; k < 16; ++k) add(k, v[dft16_reg(k)]);
    } else if constexpr (LANES == 2) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            cplx zlo, zhi;
            step_c(m, zlo, zhi);
            add(2 * m, zlo);
            add(2 * m + 1, zhi);
        }
    } else {
        // step over d (couple = lanes 16 apart): with z[j] = Z_d[k(j)], k(j) = 8c + (j>>1) + 16 (j&1)
        // (z[2m], z[2m+1] come from step_c(m)), lane d finishes j in [8d, 8d + 8):
        //   k = 8c + 4d + (jj>>1) + 16 (jj&1),  W_64^k = i^(jj&1) tw[jj>>1];  acc[2jj], acc[2jj+1] <- X[k], X[k+32]
        // Interleaved so that only four z values are alive at a time: jj = 2mm, 2mm+1 need z[2mm],
        // z[2mm+1] (step_c(mm)) and z[8+2mm], z[8+2mm+1] (step_c(4+mm)).
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            cplx zl[2], zh[2];
            step_c(mm, zl[0], zl[1]);
            step_c(4 + mm, zh[0], zh[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int jj = 2 * mm + e;
                const cplx lo = zl[e], hi = zh[e];
                const cplx keep = d ? hi : lo;
                const cplx recv = shfl_xor_c(d ? lo : hi, 16);
                const cplx y0 = d ? recv : keep, y1 = d ? keep : recv;
                cplx t = cmul(y1, tw[jj >> 1]);
                if (jj & 1) t = mul_i(t);
                add(2 * jj, cadd(y0, t));
                add(2 * jj + 1, csub(y0, t));
            }
        }
    }
}

// One pair of one warp tile: load, transform, accumulate.
template <int N, bool UNIFORM_W>
__device__ __forceinline__ void pencil_reg_pair(const cplx *__restrict__ src, int c, int d,
                                                const cplx (&tw)[4], double w, double (&acc)[16])
{
    cplx v[16];
    pencil_reg_load<N>(src, v);
    pencil_reg_compute<N, UNIFORM_W>(v, c, d, tw, w, acc);
}

// Lane geometry shared by the stand-alone kernel and the fused kernel's pencil role.
template <int N> struct PencilLane {
    int c, d, y, z;
    cplx tw[4];
    __device__ __forceinline__ void init(int lane, int wt)
    {
        using G = PencilGeo<N>;
        constexpr int ZG = N / G::ZRUN; // warp tiles per row (N >= 32), rows per tile otherwise
        const int zz = lane & (G::ZRUN - 1);
        if (N == 64) { d = lane >> 4; c = (lane >> 3) & 1; }
        else if (N == 32) { d = 0; c = lane >> 4; }
        else { d = 0; c = 0; }
        if (N == 16) { y = wt * 2 + (lane >> 4); z = zz; }
        else { y = wt / ZG; z = (wt % ZG) * G::ZRUN + zz; }
#pragma unroll
        for (int m = 0; m < 4; ++m) tw[m] = w64<+1>(0); // overwritten below for N = 64
        if (N == 64) {
            // W_64^(8c + 4d + m): 8c + 4d in {0, 4, 8, 12}
            const int base = 8 * c + 4 * d;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const cplx w0 = w64<+1>(m), w4 = w64<+1>(4 + m), w8 = w64<+1>(8 + m), w12 = w64<+1>(12 + m);
                tw[m] = base == 0 ? w0 : base == 4 ? w4 : base == 8 ? w8 : w12;
            }
        }
    }
    __device__ __forceinline__ size_t first_entry() const
    {
        constexpr size_t N2 = (size_t)N * N;
        return (size_t)(N == 64 ? 2 * c + d : c) * N2 + (size_t)y * N + z;
    }
};

// grid (WT / WARPS warp-tile groups, n_units[, cells]), block WARPS*32.
template <int N, bool UNIFORM_W, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_pencil_gain_reg(const cplx *__restrict__ hyb, int hyb_pair0, const PencilUnit *__restrict__ units,
                  const double *__restrict__ pair_w, double *__restrict__ S, int n_r_local,
                  int pairs_per_cell = 0, size_t S_cell_stride = 0)
{
    constexpr size_t N2 = (size_t)N * N, N3 = N2 * N;
    // batch of cells (grid.z): cell c's hybrid grids of the launch follow those of cell c-1
    hyb += (size_t)blockIdx.z * pairs_per_cell * N3;
    S += (size_t)blockIdx.z * S_cell_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PencilLane<N> L;
    L.init(lane, blockIdx.x * WARPS + warp);
    const PencilUnit un = units[blockIdx.y];

    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;

    const cplx *src = hyb + (size_t)(un.p0 - hyb_pair0) * N3 + L.first_entry();
    for (int q = un.p0; q < un.p1; ++q, src += N3) {
        const double w = UNIFORM_W ? 1.0 : __ldg(&pair_w[q]);
        pencil_reg_pair<N, UNIFORM_W>(src, L.c, L.d, L.tw, w, acc);
    }
    const double scale = UNIFORM_W ? __ldg(&pair_w[un.p0]) : 1.0;
    double *Sr = S + ((size_t)un.slot * n_r_local + un.r) * N3 + (size_t)L.y * N + L.z;
#pragma unroll
    for (int j = 0; j < 16; ++j) Sr[(size_t)pencil_out_x<N>(j, L.c, L.d) * N2] = acc[j] * scale;
}

} // namespace bfsm
