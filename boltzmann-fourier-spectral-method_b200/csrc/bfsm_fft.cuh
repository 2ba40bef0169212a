// bfsm_fft.cuh -- fp64 FFT building blocks for the sm_100a collision kernels.
//
// A length-N line (N = A*B) is transformed in place in shared memory by two
// register-radix passes (decimation in frequency):
//   pass 1, one unit per column b:   v[a] = x[B*a + b]  -> radix-A DFT -> * W_N^(b*k1)
//                                     -> stored at position B*k1 + b
//   pass 2, one unit per k1:          v[b] = pos[B*k1+b] -> radix-B DFT
//                                     -> X[k1 + A*k2] left at position B*k1 + k2
// so natural frequency k = k1 + A*k2 lives at position B*k1 + k2.  A (y,z) plane is an
// N x ROW array of complex doubles, position p of a row at column p + (p>>3)
// (one pad slot per 8 elements) and ROW chosen so that both the row-wise and the
// column-wise passes are free of shared-memory bank conflicts for 16-byte accesses.
//
// The packed gain kernel (k_plane_gain3) uses a different plane scheme on the same radix
// butterflies: three register stages (radix-R, 4x4 block, radix-R; N = 4R) over a plane with
// row pitch N+1 and columns in natural order -- see bfsm_kernels.cuh.
//
// No cuFFT, no tensor cores: the whole path is fp64 SIMT (DFMA/DADD) + LDS/STS.
#pragma once
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched through the runtime)
#include <cuda_runtime.h>

namespace bfsm {

typedef double2 cplx;

template <int N> struct Geo;
template <> struct Geo<64> { static constexpr int A = 8, B = 8, ROW = 73; };
template <> struct Geo<32> { static constexpr int A = 8, B = 4, ROW = 42; };
template <> struct Geo<16> { static constexpr int A = 4, B = 4, ROW = 18; };

__device__ __forceinline__ int padk(int k) { return k + (k >> 3); }

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b)
{
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by SIGN*i
template <int SIGN> __device__ __forceinline__ cplx mul_si(cplx a)
{
    return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// Natural-order in-place DFTs with kernel exp(SIGN * 2 pi i n k / R).
template <int SIGN> __device__ __forceinline__ void dft4(cplx &x0, cplx &x1, cplx &x2, cplx &x3)
{
    const cplx c0 = cadd(x0, x2), c1 = csub(x0, x2);
    const cplx c2 = cadd(x1, x3), c3 = mul_si<SIGN>(csub(x1, x3));
    x0 = cadd(c0, c2);
    x2 = csub(c0, c2);
    x1 = cadd(c1, c3);
    x3 = csub(c1, c3);
}

template <int R, int SIGN> struct Dft;

template <int SIGN> struct Dft<4, SIGN> {
    static __device__ __forceinline__ void run(cplx (&v)[4]) { dft4<SIGN>(v[0], v[1], v[2], v[3]); }
};

template <int SIGN> struct Dft<8, SIGN> {
    static __device__ __forceinline__ void run(cplx (&v)[8])
    {
        constexpr double h = 0.70710678118654752440;
        constexpr double s = (double)SIGN;
        cplx a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
        cplx a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
        cplx a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
        cplx a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
        // odd half times W8^n, W8 = exp(SIGN*i*pi/4)
        a5 = make_double2((a5.x - s * a5.y) * h, (a5.y + s * a5.x) * h);
        a6 = mul_si<SIGN>(a6);
        a7 = make_double2((-a7.x - s * a7.y) * h, (-a7.y + s * a7.x) * h);
        dft4<SIGN>(a0, a1, a2, a3); // X[0], X[2], X[4], X[6]
        dft4<SIGN>(a4, a5, a6, a7); // X[1], X[3], X[5], X[7]
        v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
        v[1] = a4; v[3] = a5; v[5] = a6; v[7] = a7;
    }
};

// Radix-16 as 4 x 4 with compile-time W16 twiddles.  The result is left TRANSPOSED:
// X[k1 + 4*k2] is in v[4*k1 + k2]; use dft16_reg(K) to find the register of frequency K.
__host__ __device__ constexpr int dft16_reg(int K) { return 4 * (K % 4) + K / 4; }

template <int SIGN> struct Dft<16, SIGN> {
    static __device__ __forceinline__ cplx mulw(cplx a, double wr, double wi)
    {
        return make_double2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
    static __device__ __forceinline__ void run(cplx (&v)[16])
    {
        constexpr double s = (double)SIGN;
        constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        constexpr double h = 0.70710678118654752440;
        // pass 1: radix-4 over a for each residue b (elements v[4a+b]); y[k1][b] -> v[4*k1+b]
#pragma unroll
        for (int b = 0; b < 4; ++b) dft4<SIGN>(v[b], v[4 + b], v[8 + b], v[12 + b]);
        // twiddles W16^(b*k1)
        v[4 * 1 + 1] = mulw(v[4 * 1 + 1], c1, s * s1);                                   // W^1
        v[4 * 1 + 2] = make_double2((v[6].x - s * v[6].y) * h, (v[6].y + s * v[6].x) * h); // W^2
        v[4 * 1 + 3] = mulw(v[4 * 1 + 3], s1, s * c1);                                   // W^3
        v[4 * 2 + 1] = make_double2((v[9].x - s * v[9].y) * h, (v[9].y + s * v[9].x) * h); // W^2
        v[4 * 2 + 2] = mul_si<SIGN>(v[4 * 2 + 2]);                                       // W^4
        v[4 * 2 + 3] = make_double2((-v[11].x - s * v[11].y) * h, (-v[11].y + s * v[11].x) * h); // W^6
        v[4 * 3 + 1] = mulw(v[4 * 3 + 1], s1, s * c1);                                   // W^3
        v[4 * 3 + 2] = make_double2((-v[14].x - s * v[14].y) * h, (-v[14].y + s * v[14].x) * h); // W^6
        v[4 * 3 + 3] = mulw(v[4 * 3 + 3], -c1, -s * s1);                                 // W^9
        // pass 2: radix-4 over b for each k1; X[k1 + 4*k2] -> v[4*k1 + k2]
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4<SIGN>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    }
};

// exp(SIGN * 2 pi i m / 64) for a compile-time-foldable m (callers unroll their loops): quarter-wave
// table of cos(2 pi r / 64), r = 0..16, long-double literals.
template <int SIGN> __device__ __forceinline__ cplx w64(int m)
{
    const double C[17] = {1.0, 0.995184726672196886245, 0.980785280403230449126,
                          0.956940335732208864936, 0.923879532511286756128, 0.881921264348355029713,
                          0.831469612302545237079, 0.773010453362736960811, 0.707106781186547524401,
                          0.634393284163645498215, 0.555570233019602224743, 0.471396736825997648556,
                          0.382683432365089771728, 0.290284677254462367636, 0.195090322016128267848,
                          0.0980171403295606019942, 0.0};
    const int q = (m & 63) >> 4, r = m & 15;
    const double c = C[r], s = C[16 - r];
    const double wr = q == 0 ? c : q == 1 ? -s : q == 2 ? -c : s;
    const double wi = q == 0 ? s : q == 1 ? c : q == 2 ? -s : -c;
    return make_double2(wr, SIGN > 0 ? wi : -wi);
}

// Radix-32 as 4 x 8, every twiddle a compile-time constant.  Input natural (v[n]); the result is
// left TRANSPOSED: X[k1 + 4*k2] is in v[8*k1 + k2]; use dft32_reg(K).
__host__ __device__ constexpr int dft32_reg(int K) { return 8 * (K % 4) + K / 4; }

template <int SIGN> struct Dft<32, SIGN> {
    static __device__ __forceinline__ void run(cplx (&v)[32])
    {
        // pass 1: radix-4 over n1 for each n2 (elements v[8 n1 + n2]); y[k1][n2] -> v[8 k1 + n2]
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) dft4<SIGN>(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);
        // twiddles W32^(n2 k1) = W64^(2 n2 k1)
#pragma unroll
        for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
            for (int n2 = 1; n2 < 8; ++n2) v[8 * k1 + n2] = cmul(v[8 * k1 + n2], w64<SIGN>(2 * n2 * k1));
        // pass 2: radix-8 over n2 for each k1; X[k1 + 4 k2] -> v[8 k1 + k2]
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) Dft<8, SIGN>::run(reinterpret_cast<cplx(&)[8]>(v[8 * k1]));
    }
};

// Register holding frequency K after Dft<R>::run (natural for R = 4, 8; transposed for 16, 32).
template <int R> __host__ __device__ constexpr int dft_reg(int K)
{
    return R == 32 ? dft32_reg(K) : R == 16 ? dft16_reg(K) : K;
}

// Pass-1 twiddles of the calling thread: tw[k1-1] = W_N^(SIGN*b*k1), k1 = 1..A-1.
// twtab[t] = exp(+2 pi i t / N), t in [0,N), tabulated on the host in long double.
template <int N, int SIGN>
__device__ __forceinline__ void load_twiddles(cplx (&tw)[Geo<N>::A - 1], const cplx *__restrict__ twtab,
                                              int b)
{
#pragma unroll
    for (int k1 = 1; k1 < Geo<N>::A; ++k1) {
        const cplx w = __ldg(&twtab[(b * k1) & (N - 1)]);
        tw[k1 - 1] = make_double2(w.x, SIGN > 0 ? w.y : -w.y);
    }
}

// ------------------------------------------------------------------ plane (y,z) passes
// `buf` is an N x ROW padded plane in shared memory, `tg` the thread's index inside
// its group of TG threads.  TG must be a multiple of B so that a thread's column
// residue b = tg % B is the same for every unit it owns (twiddles live in registers).

// z pass 1: `in(j,k)` supplies element (row j, spectral/physical column k).
template <int N, int SIGN, int TG, class In>
__device__ __forceinline__ void z1_pass(cplx *buf, const cplx (&tw)[Geo<N>::A - 1], int tg, In in)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, ROW = Geo<N>::ROW;
    static_assert(TG % B == 0, "TG must be a multiple of B");
#pragma unroll
    for (int u0 = 0; u0 < N * B; u0 += TG) {
        const int u = u0 + tg;
        if ((N * B) % TG != 0 && u >= N * B) break;
        const int j = u / B, b = u % B;
        cplx v[A];
#pragma unroll
        for (int a = 0; a < A; ++a) v[a] = in(j, B * a + b);
        Dft<A, SIGN>::run(v);
        cplx *row = buf + j * ROW;
        row[padk(b)] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < A; ++k1) row[padk(B * k1 + b)] = cmul(v[k1], tw[k1 - 1]);
    }
}

// z pass 2 (in place).
template <int N, int SIGN, int TG> __device__ __forceinline__ void z2_pass(cplx *buf, int tg)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, ROW = Geo<N>::ROW;
#pragma unroll
    for (int u0 = 0; u0 < N * A; u0 += TG) {
        const int u = u0 + tg;
        if ((N * A) % TG != 0 && u >= N * A) break;
        const int j = u / A, k1 = u % A;
        cplx *row = buf + j * ROW;
        cplx v[B];
#pragma unroll
        for (int b = 0; b < B; ++b) v[b] = row[padk(B * k1 + b)];
        Dft<B, SIGN>::run(v);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) row[padk(B * k1 + k2)] = v[k2];
    }
}

// y pass 1 (in place); unit = (column position q, row residue b) with b = u % B so the
// twiddles are the same registers as in z pass 1.
template <int N, int SIGN, int TG>
__device__ __forceinline__ void y1_pass(cplx *buf, const cplx (&tw)[Geo<N>::A - 1], int tg)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, ROW = Geo<N>::ROW;
#pragma unroll
    for (int u0 = 0; u0 < N * B; u0 += TG) {
        const int u = u0 + tg;
        if ((N * B) % TG != 0 && u >= N * B) break;
        const int b = u % B, q = u / B;
        cplx *col = buf + padk(q);
        cplx v[A];
#pragma unroll
        for (int a = 0; a < A; ++a) v[a] = col[(B * a + b) * ROW];
        Dft<A, SIGN>::run(v);
        col[b * ROW] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < A; ++k1) col[(B * k1 + b) * ROW] = cmul(v[k1], tw[k1 - 1]);
    }
}

// y pass 2: results leave shared memory through `out(y, z, value)` in natural order;
// consecutive threads own consecutive natural z, so global stores coalesce.
template <int N, int SIGN, int TG, class Out>
__device__ __forceinline__ void y2_pass(const cplx *buf, int tg, Out out)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, ROW = Geo<N>::ROW;
#pragma unroll
    for (int u0 = 0; u0 < N * A; u0 += TG) {
        const int u = u0 + tg;
        if ((N * A) % TG != 0 && u >= N * A) break;
        const int z = u % N, k1 = u / N;
        const int q = B * (z % A) + z / A; // position holding natural index z
        const cplx *col = buf + padk(q);
        cplx v[B];
#pragma unroll
        for (int b = 0; b < B; ++b) v[b] = col[(B * k1 + b) * ROW];
        Dft<B, SIGN>::run(v);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) out(k1 + A * k2, z, v[k2]);
    }
}

// ------------------------------------------------------------------ x pencil passes
// A tile is N x TZ (x by 8 consecutive z) complex doubles, dense in shared memory;
// TGP = B*TZ threads, thread (b = tg / TZ, z = tg % TZ) in pass 1.
constexpr int TZ = 8;

template <int N, int SIGN, class In>
__device__ __forceinline__ void x1_pass(cplx *sm, const cplx (&tw)[Geo<N>::A - 1], int tg, In in)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B;
    const int z = tg % TZ, b = tg / TZ;
    cplx v[A];
#pragma unroll
    for (int a = 0; a < A; ++a) v[a] = in(B * a + b, z);
    Dft<A, SIGN>::run(v);
    sm[b * TZ + z] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < A; ++k1) sm[(B * k1 + b) * TZ + z] = cmul(v[k1], tw[k1 - 1]);
}

// Number of pass-2 units a thread of a B*TZ group owns.
template <int N> struct X2 { static constexpr int UNITS = Geo<N>::A / Geo<N>::B; };

// x pass 2, unit m of the thread: fills v[k2] = X[k1 + A*k2]; returns k1.
template <int N, int SIGN>
__device__ __forceinline__ int x2_unit(const cplx *sm, int tg, int m, cplx (&v)[Geo<N>::B])
{
    constexpr int B = Geo<N>::B;
    const int u = tg + m * (B * TZ);
    const int z = u % TZ, k1 = u / TZ;
#pragma unroll
    for (int b = 0; b < B; ++b) v[b] = sm[(B * k1 + b) * TZ + z];
    Dft<B, SIGN>::run(v);
    return k1;
}

// Fourier mode of index t (FFTWBoltzmannOperator.cpp:50-57): 0..N/2-1, -N/2..-1.
template <int N> __device__ __forceinline__ int mode_of(int t) { return t < N / 2 ? t : t - N; }

// x pass 1 reading its inputs from the same dense tile it writes (in place): thread (b,z)
// reads rows B*a+b and writes rows B*k1+b -- the same set of rows, so no hazard inside the pass.
template <int N, int SIGN>
__device__ __forceinline__ void x1_pass_inplace(cplx *sm, const cplx (&tw)[Geo<N>::A - 1], int tg)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B;
    const int z = tg % TZ, b = tg / TZ;
    cplx v[A];
#pragma unroll
    for (int a = 0; a < A; ++a) v[a] = sm[(B * a + b) * TZ + z];
    Dft<A, SIGN>::run(v);
    sm[b * TZ + z] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < A; ++k1) sm[(B * k1 + b) * TZ + z] = cmul(v[k1], tw[k1 - 1]);
}

// 16-byte asynchronous global -> shared copy (LDGSTS) and its group bookkeeping
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory");
}

// ---- bulk asynchronous copies (TMA, SASS: UBLKCP) completing on an mbarrier ---------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(a), "r"(parity) : "memory");
}
// `bytes` (a multiple of 16) from global to shared memory, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
// One TMA tensor copy: the box at coordinates (c0, c1, c2) of the 3-D tensor map -> shared memory
// (SASS: UTMALDG), completion counted on `bar`.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const void *tmap, int c0, int c1, int c2,
                                            unsigned long long *bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(d), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void group_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

} // namespace bfsm
