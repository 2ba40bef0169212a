// bfsm_kernels.cuh -- the sm_100a kernels of the fused FFT-collision path.
//
// One evaluation Q(f,f) on an N^3 grid (reference: FFTWBoltzmannOperator.cpp:147-334):
//
//   f --k_plane<REAL>--> Fh --k_pencil_fwd--> fhat/N^3                     (cpp:168-186)
//   for chunks of (r,sigma) pairs:                                          (cpp:191-250)
//       plane kernel : fhat plane x separable phase -> 2-D inverse FFT (y,z) -> hybrid scratch
//                      (unpacked mode: of both alpha1*fhat and conj(alpha1)*fhat).  Three
//                      implementations: k_plane_gain_ws (64^3 packed mode, default: warp-specialised
//                      3-stage pipeline), k_plane_gain3 (3 register stages, every warp does all of
//                      them), k_plane_gain (4 passes; the only one for unpacked mode)
//       pencil kernel: inverse FFT along x, Re(g1*g2) * w_pair accumulated in registers over the
//                      chunk's pairs, per-r flush into S_r (k_pencil_gain_async / k_pencil_gain)
//   S_r --k_plane<REAL>--> Ph_r --k_pencil_accum--> Qhat = sum_r coef_r(|l|^2) FFT3(S_r)
//                                                                           (cpp:249-273)
//   Qhat, beta2*fhat --k_plane<FINAL>--> H --k_pencil_final--> Q = Re(Qg) - Re(h) f
//                                                                           (cpp:281-330)
//
// Differences from the reference's loop nest, all exact up to fp64 rounding:
//   * the forward transform is linear, so it is applied once per radius r to
//     S_r = sum_sigma w_sigma Re(g1 g2) instead of once per pair (3P+4 -> 2P+N_r+4 FFTs);
//   * beta1 is real and even in l, so only Re(g1 g2) can reach Re(Q_gain) (the only part
//     the reference keeps, cpp:326);
//   * exp(i theta(l)) = ex[i] ey[j] ez[k] (theta is linear in l): three N-entry tables per pair;
//   * beta1(r,|l|) and beta2(|l|) depend on l only through the integer |l|^2: tabulated;
//   * 1/N^3 (a power of two, hence exact) is folded into fhat and the tables.
//
// PACKED mode (default; f is real, so fhat is Hermitian): write e^{i theta} = c + i s.  Then
//   g1 = C + iD, g2 = C - iD with C = IFFT3(c fhat), D = IFFT3(s fhat), and
//   Re(g1 g2) = Re(Z_H^2) + Re(Y^2),   Z_H = IFFT3(m_H fhat),  Y = IFFT3(n fhat)
// with REAL multipliers m_H = c_even + s_odd, n = c_odd + s_even (even/odd under index reversal
// l -> -l mod N).  n vanishes off the three Nyquist planes, so
//   Y(v) = (-1)^vx U(vy,vz) + (-1)^vy V(vx,vz) + (-1)^vz W(vx,vy)
// is assembled from three 2-D transforms per pair (the plane kernels' "planes" N, N+1, N+2) and its
// square is accumulated by k_nyq_accum.  ONE 3-D transform per pair instead of two, exact for every real input
// (validated against the unpacked path and the oracle with non-band-limited noise input).
#pragma once
#include "bfsm_fft.cuh"
#include "bfsm_pencil_reg.cuh"

namespace bfsm {

// ---------------------------------------------------------------------------------------
// k_plane_gain (UNPACKED mode, BFSM_FLAG_NO_PACK): persistent grid (about one CTA per SM slot),
// block GROUPS*TG.  The flat work list (plane i, item it) with it = 2*pair + array is split evenly
// over the CTAs; a CTA walks its range plane by plane (reloading the fhat plane when it changes)
// and its GROUPS groups take the items of a plane alternately.  Four passes (z1, z2, y1, y2) over a
// padded plane, see bfsm_fft.cuh.
// Shared memory: fhat plane (N*N) | GROUPS padded planes (N*ROW) | GROUPS x 2 phase slots (3N).
// ---------------------------------------------------------------------------------------
template <int N, int TG, int GROUPS, int MINB>
__global__ void __launch_bounds__(TG *GROUPS, MINB)
k_plane_gain(const cplx *__restrict__ fhat, const cplx *__restrict__ phase,
             const cplx *__restrict__ twtab, cplx *__restrict__ hyb, int pair0, int n_items)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, ROW = Geo<N>::ROW;
    static_assert(TG >= 3 * N, "phase staging needs 3N threads per group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *fpl = reinterpret_cast<cplx *>(smem_raw);
    cplx *bufs = fpl + N * N;
    cplx *phs = bufs + GROUPS * N * ROW;

    const int g = threadIdx.x / TG, tg = threadIdx.x % TG;
    cplx *buf = bufs + g * N * ROW;
    cplx *myph = phs + g * 2 * 3 * N;

    cplx tw[A - 1];
    load_twiddles<N, +1>(tw, twtab, tg % B);

    // flat work list: index = plane * n_items + item; this CTA owns [w_lo, w_hi)
    const long long total = (long long)N * n_items;
    const int w_lo = (int)((total * blockIdx.x) / gridDim.x);
    const int w_hi = (int)((total * (blockIdx.x + 1)) / gridDim.x);

    int w = w_lo;
    while (w < w_hi) {
        const int i = w / n_items;                       // plane of this segment
        const int it_lo = w - i * n_items;
        int it_hi = n_items;
        if ((long long)(i + 1) * n_items > w_hi) it_hi = w_hi - i * n_items;

        __syncthreads(); // every group is done with the previous plane
        {
            const cplx *srcp = fhat + (size_t)i * N * N;
            for (int t = threadIdx.x; t < N * N; t += TG * GROUPS) fpl[t] = srcp[t];
        }
        const int first = it_lo + g;
        if (first < it_hi && tg < 3 * N)
            myph[tg] = __ldg(&phase[(size_t)(pair0 + (first >> 1)) * 3 * N + tg]);
        __syncthreads();

        int slot = 0;
        for (int it = first; it < it_hi; it += GROUPS, slot ^= 1) {
            const int arr = it & 1;
            const cplx *P = myph + slot * 3 * N;
            const bool have_next = (it + GROUPS < it_hi) && (tg < 3 * N);
            cplx nxt = make_double2(0.0, 0.0);
            if (have_next) nxt = __ldg(&phase[(size_t)(pair0 + ((it + GROUPS) >> 1)) * 3 * N + tg]);
            const cplx exi = P[i];

            // z pass 1 with the phase-weighted load fused in (cpp:198-225):
            // A1 = e^{i theta} fhat, A2 = e^{-i theta} fhat, theta separable in (i,j,k).
#pragma unroll
            for (int u0 = 0; u0 < N * B; u0 += TG) {
                const int u = u0 + tg;
                const int j = u / B, b = u % B;
                const cplx exy = cmul(exi, P[N + j]);
                cplx v[A];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const int k = B * a + b;
                    const cplx e = cmul(exy, P[2 * N + k]);
                    const cplx f = fpl[j * N + k];
                    v[a] = arr ? cmulc(f, e) : cmul(f, e);
                }
                Dft<A, +1>::run(v);
                cplx *row = buf + j * ROW;
                row[padk(b)] = v[0];
#pragma unroll
                for (int k1 = 1; k1 < A; ++k1) row[padk(B * k1 + b)] = cmul(v[k1], tw[k1 - 1]);
            }
            group_sync(1 + g, TG);
            z2_pass<N, +1, TG>(buf, tg);
            group_sync(1 + g, TG);
            if (have_next) myph[(slot ^ 1) * 3 * N + tg] = nxt;
            y1_pass<N, +1, TG>(buf, tw, tg);
            group_sync(1 + g, TG);
            cplx *dst = hyb + ((size_t)it * N + i) * N * N;
            y2_pass<N, +1, TG>(buf, tg, [&](int y, int z, cplx val) { dst[y * N + z] = val; });
            group_sync(1 + g, TG);
        }
        w = i * n_items + it_hi;
    }
}

// ---------------------------------------------------------------------------------------
// k_plane_gain3 (packed mode): same persistent work split as k_plane_gain, but the (y,z) plane
// transform uses THREE register stages and only TWO shared-memory exchanges (N = 4*R):
//   S1  thread (row j, residue b):   radix-R along z over k = 4a+b, phase-weighted load fused in
//   S2  thread (k1, row residue b'): 4x4 block -- z twiddle W_N^(b k1), radix-4 along z (b),
//                                    radix-4 along y (rows R a'+b'), y twiddle W_N^(b' k1')
//   S3  thread (k1', column slot):   radix-R along y over rows R k1'+b', natural-order store
// Column k1 + R*k2 of a row holds natural z (S1 writes residue b of frequency k1 to column
// k1 + R*b); row R*k1'+b' feeds natural y = k1' + 4*k2'.  Rows have pitch N+1 elements (== 16 bytes
// mod 128); S1 and S2 map consecutive lanes to consecutive rows, S3 to consecutive columns, so all
// 16-byte accesses are bank-conflict free, global stores are contiguous per quarter-warp, and the
// z phase table is read as a warp broadcast.  Per element: 5 shared-memory accesses + 1 global
// store (the 4-pass version needs 8 + 1).
// ---------------------------------------------------------------------------------------
template <int N, int GROUPS, int MINB>
__global__ void __launch_bounds__(4 * N *GROUPS, MINB)
k_plane_gain3(const cplx *__restrict__ fhat, const cplx *__restrict__ phase,
              const cplx *__restrict__ twtab, cplx *__restrict__ hyb, int pair0, int n_items,
              const cplx *__restrict__ nyq, const double *__restrict__ pair_w,
              cplx *__restrict__ uvw, int parts = 3)
{
    constexpr int R = N / 4, TG = 4 * N, PITCH = N + 1, H = N / 2, NPL = N + 3;
    static_assert(TG >= 3 * N, "phase staging needs 3N threads per group");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *fpl = reinterpret_cast<cplx *>(smem_raw);          // N x PITCH
    cplx *bufs = fpl + N * PITCH;                             // GROUPS x (N x PITCH)
    cplx *phs = bufs + GROUPS * N * PITCH;                    // GROUPS x 2 x 3N

    const int g = threadIdx.x / TG, tg = threadIdx.x % TG;
    cplx *buf = bufs + g * N * PITCH;
    cplx *myph = phs + g * 2 * 3 * N;

    // S1 role: row j1, residue b1.  S2 role: k1 = s2k, b' = s2b (threads < R*R).  S3: slot, k1'.
    const int j1 = tg % N, b1 = tg / N;
    const int s2b = tg % R, s2k = (tg / R) % R;
    const bool s2_active = tg < R * R;
    const int s3slot = tg % N, s3k = tg / N;

    // S2 twiddles: wz[b-1] = W_N^(b*k1), wy[k-1] = W_N^(b'*k), b,k = 1..3 (inverse transform: +)
    cplx wz[3], wy[3];
#pragma unroll
    for (int m = 1; m < 4; ++m) {
        wz[m - 1] = __ldg(&twtab[(m * s2k) & (N - 1)]);
        wy[m - 1] = __ldg(&twtab[(m * s2b) & (N - 1)]);
    }

    // Work split: the flat list index = plane * n_items + item has two classes of entries, the N
    // regular planes and the 3 (costlier) Nyquist planes; every CTA takes an equal share of EACH
    // class, so that all CTAs finish together (with one contiguous range per CTA the owners of the
    // Nyquist planes ran 10 % longer than the rest -- ncu sm__cycles_active max vs avg).
    // parts: bit 0 = the N regular planes, bit 1 = the 3 Nyquist planes (the cluster kernel does the
    // regular planes itself and leaves only the Nyquist planes to this kernel)
    for (int part = 0; part < 2; ++part) {
    if (!((parts >> part) & 1)) continue;
    const long long total = (long long)(part == 0 ? N : NPL - N) * n_items;
    const int base = part == 0 ? 0 : N * n_items;
    const int w_lo = base + (int)((total * blockIdx.x) / gridDim.x);
    const int w_hi = base + (int)((total * (blockIdx.x + 1)) / gridDim.x);

    int w = w_lo;
    while (w < w_hi) {
        const int i = w / n_items;
        const int it_lo = w - i * n_items;
        int it_hi = n_items;
        if ((long long)(i + 1) * n_items > w_hi) it_hi = w_hi - i * n_items;

        __syncthreads();
        {
            const cplx *srcp = (i < N) ? fhat + (size_t)i * N * N : nyq + (size_t)(i - N) * N * N;
            for (int t = threadIdx.x; t < N * N; t += TG * GROUPS)
                fpl[(t / N) * PITCH + (t % N)] = srcp[t];
        }
        const int first = it_lo + g;
        if (first < it_hi && tg < 3 * N)
            myph[tg] = __ldg(&phase[(size_t)(pair0 + first) * 3 * N + tg]);
        __syncthreads();

        int slot = 0;
        for (int it = first; it < it_hi; it += GROUPS, slot ^= 1) {
            const cplx *P = myph + slot * 3 * N;
            const bool have_next = (it + GROUPS < it_hi) && (tg < 3 * N);
            cplx nxt = make_double2(0.0, 0.0);
            if (have_next) nxt = __ldg(&phase[(size_t)(pair0 + it + GROUPS) * 3 * N + tg]);

            // ---------------- S1: phase-weighted load + radix-R along z
            {
                const int j = j1, b = b1;
                cplx v[R];
                const cplx *frow = fpl + j * PITCH;
                if (i < N) {
                    // m_H = ((Re E + Im E) + (Re Et - Im Et))/2 with E = X*Z, X = ex[i] ey[j],
                    // Z = ez[k], and Et = E(-l) = Xt*conj(Z) (Xt*Z at the Nyquist column k == H),
                    // Xt = ex~[i] ey~[j] (conjugates except at the Nyquist index).  Expanding:
                    //   m_H = A (Z.x+Z.y) + B (Z.x-Z.y),   A = (X.x+Xt.x)/2, B = (X.y-Xt.y)/2
                    //   k == H:                            A = (X.x-Xt.y)/2, B = (X.y+Xt.x)/2
                    // -- one divergence-free formula for interior and Nyquist rows alike.
                    const cplx exi = P[i], eyj = P[N + j];
                    const cplx X = cmul(exi, eyj);
                    const cplx ext = (i == H) ? exi : make_double2(exi.x, -exi.y);
                    const cplx eyt = (j == H) ? eyj : make_double2(eyj.x, -eyj.y);
                    const cplx Xt = cmul(ext, eyt);
                    const double cA = 0.5 * (X.x + Xt.x), cB = 0.5 * (X.y - Xt.y);
                    const double nA = 0.5 * (X.x - Xt.y), nB = 0.5 * (X.y + Xt.x);
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int k = 4 * a + b;
                        const cplx ez = P[2 * N + k];
                        const double zp = ez.x + ez.y, zm = ez.x - ez.y;
                        const bool ny = (a == R / 2) && (b == 0); // k == H
                        const double m = (ny ? nA : cA) * zp + (ny ? nB : cB) * zm;
                        const cplx f = frow[k];
                        v[a] = make_double2(m * f.x, m * f.y);
                    }
                } else {
                    // Nyquist plane q = i - N: fixed axis q, free axes (axA rows, axB columns)
                    const int nq = i - N;
                    const int axA = (nq == 0) ? 1 : 0, axB = (nq == 2) ? 1 : 2;
                    const cplx efix = P[nq * N + H];
                    const double sw = 0.5 * sqrt(__ldg(&pair_w[pair0 + it]));
                    const cplx ea = P[axA * N + j];
                    const cplx eat = (j == H) ? ea : make_double2(ea.x, -ea.y);
                    const cplx fa = cmul(efix, ea), fat = cmul(efix, eat);
                    const bool zero_row = (nq >= 1) && (j == H);
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int k = 4 * a + b;
                        const cplx eb = P[axB * N + k];
                        const cplx ebt = (k == H) ? eb : make_double2(eb.x, -eb.y);
                        const cplx e = cmul(fa, eb), et = cmul(fat, ebt);
                        double n = sw * ((e.x - et.x) + (e.y + et.y));
                        if (zero_row || (nq == 2 && k == H)) n = 0.0;
                        const cplx f = frow[k];
                        v[a] = make_double2(n * f.x, n * f.y);
                    }
                }
                Dft<R, +1>::run(v);
                cplx *row = buf + j * PITCH + R * b;
#pragma unroll
                for (int k1 = 0; k1 < R; ++k1) row[k1] = v[dft_reg<R>(k1)];
            }
            group_sync(1 + g, TG);
            if (have_next) myph[(slot ^ 1) * 3 * N + tg] = nxt;

            // ---------------- S2: 4 x 4 block, z twiddle + radix-4 (z), radix-4 (y) + y twiddle
            if (s2_active) {
                cplx e[4][4]; // [a' (row R a'+b')][b (column k1 + R b)]
                cplx *blk = buf + s2b * PITCH + s2k;
#pragma unroll
                for (int ap = 0; ap < 4; ++ap)
#pragma unroll
                    for (int b = 0; b < 4; ++b) e[ap][b] = blk[(R * ap) * PITCH + R * b];
#pragma unroll
                for (int ap = 0; ap < 4; ++ap) {
#pragma unroll
                    for (int b = 1; b < 4; ++b) e[ap][b] = cmul(e[ap][b], wz[b - 1]);
                    dft4<+1>(e[ap][0], e[ap][1], e[ap][2], e[ap][3]); // -> z = k1 + R*k2 in e[ap][k2]
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    dft4<+1>(e[0][c], e[1][c], e[2][c], e[3][c]);     // -> k1' in e[k1'][c]
#pragma unroll
                    for (int kp = 1; kp < 4; ++kp) e[kp][c] = cmul(e[kp][c], wy[kp - 1]);
                }
#pragma unroll
                for (int kp = 0; kp < 4; ++kp)
#pragma unroll
                    for (int c = 0; c < 4; ++c) blk[(R * kp) * PITCH + R * c] = e[kp][c];
            }
            group_sync(1 + g, TG);

            // ---------------- S3: radix-R along y, natural-order store
            {
                cplx v[R];
                const cplx *col = buf + (R * s3k) * PITCH + s3slot;
#pragma unroll
                for (int bp = 0; bp < R; ++bp) v[bp] = col[bp * PITCH];
                Dft<R, +1>::run(v);
                const int z = s3slot; // columns are in natural z order
                cplx *dst = (i < N) ? hyb + ((size_t)it * N + i) * N * N
                                    : uvw + ((size_t)it * 3 + (i - N)) * N * N;
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) dst[(s3k + 4 * k2) * N + z] = v[dft_reg<R>(k2)];
            }
            group_sync(1 + g, TG);
        }
        w = i * n_items + it_hi;
    }
    } // part
}

// ---------------------------------------------------------------------------------------
// k_plane_gain_ws (packed mode, N = 64): the three stages of k_plane_gain3 as a WARP-SPECIALISED
// PIPELINE.  A CTA is three warpgroups of 128 threads, one per stage S1, S2, S3;
// every warpgroup walks the CTA's whole item list and item n lives in plane buffer n % 3 on its
// way through the stages:
//
//   S1 warpgroup: keeps its entries of the fhat plane IN REGISTERS for as long as the plane does
//       not change (no per-item re-read of fhat from shared memory), apply the real multiplier m_H,
//       radix-16 along z, store rows                                       -> full1[buf]
//   S2 warpgroup: 4 x 4 blocks in place (both twiddles)                    -> full2[buf]
//   S3 warpgroup: radix-16 along y, natural-order global store             -> empty[buf]
//
// Why: in k_plane_gain3 all warps of a group are in the same stage, so the SM alternates between
// LDS/STS bursts and FP64 butterflies (ncu: l1tex data pipe 64 %, FP64 pipe 52 %, both idle a third
// of the time); here every SM sub-partition holds a warp of each stage, so that the LSU and the
// FP64 pipe always have a customer, and the fhat re-read (one of six 16-byte accesses per element)
// is gone.  Hand-offs are named barriers used as producer/consumer pairs (bar.arrive by the
// producers, bar.sync by the consumers): ids 1-3 full1, 4-6 full2, 7-9 empty, 10 is the S1
// warpgroups' own barrier (phase-table and plane staging).  S3 releases a buffer only after the
// global stores that depend on its loads, so a buffer cannot be overwritten under an outstanding
// LDS.  Registers move between the warpgroups with setmaxnreg:
//   384 threads launched at 168 -> S1 240 | S2 152 | S3 112      (2 units per S1 thread)
// (two S1 warpgroups at 512 threads and three other register splits were measured 2.3-5 % slower:
// profiles/r02_ab64_next.log).
// The arithmetic is that of k_plane_gain3 (same formulas, same operation order; the results agree
// to the last bit or two -- the compiler contracts a few multiply-adds differently).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync_n(int id, int n)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive_n(int id, int n)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
template <int REGS> __device__ __forceinline__ void reg_alloc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS> __device__ __forceinline__ void reg_dealloc()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}

// Walks a CTA's entries of the flat (plane, item) list: `leftA` entries of a first contiguous range,
// then a second range starting at (iB, itB).
struct ItemWalk {
    int i, it, n_items, leftA, iB, itB;
    __host__ __device__ __forceinline__ void next()
    {
        if (--leftA == 0) { i = iB; it = itB; }
        else if (++it == n_items) { it = 0; ++i; }
    }
};

// register targets (setmaxnreg) of the three warpgroups; 384 threads launched at 168
struct WsRegs { static constexpr int S1 = 240, S2 = 152, S3 = 112; };

// Work list of one CTA of the stand-alone pipelined plane kernel: one launch = one "sub-chunk".
// A walk names, for the current entry, the plane `i` (0..N-1 regular, N..N+2 Nyquist), the plan-local
// pair `pair` (phase tables, weight), the destination slot `dst_item` (plane index into hyb: dst_item*N + i;
// into uvw: dst_item*3 + i-N) and the sub-chunk `sub` it belongs to (hand-over granularity of the
// fused kernel).
template <int N> struct LaunchWalk {
    ItemWalk w;
    int cnt, pair0;
    int i, pair, dst_item, sub;
    __host__ __device__ __forceinline__ void init(int n_items, int pair0_, int cta, int n_ctas)
    {
        // an equal share of the N regular planes, then an equal share of the 3 costlier Nyquist planes
        const long long totA = (long long)N * n_items, totB = (long long)3 * n_items;
        const int a_lo = (int)((totA * cta) / n_ctas), a_hi = (int)((totA * (cta + 1)) / n_ctas);
        const int b_lo = (int)((totB * cta) / n_ctas), b_hi = (int)((totB * (cta + 1)) / n_ctas);
        const int cntA = a_hi - a_lo;
        cnt = cntA + (b_hi - b_lo);
        pair0 = pair0_;
        w.n_items = n_items;
        w.iB = N + b_lo / n_items;
        w.itB = b_lo % n_items;
        if (cntA > 0) { w.i = a_lo / n_items; w.it = a_lo % n_items; w.leftA = cntA; }
        else          { w.i = w.iB;           w.it = w.itB;          w.leftA = -1; }
        sub = 0;
        load();
    }
    __host__ __device__ __forceinline__ void load() { i = w.i; pair = pair0 + w.it; dst_item = w.it; }
    __host__ __device__ __forceinline__ void next() { w.next(); load(); }
};

// Hand-over of the fused kernel (plane role -> pencil role through an L2-resident ring of sub-chunks);
// unused (null counters) by the stand-alone plane kernel.
struct RingSync {
    int *ready;          // [n_sub] plane CTAs that finished sub-chunk s
    int *consumed;       // [n_sub] pairs x warp tiles of sub-chunk s the pencil role is done with
    int n_sub, ring;     // sub-chunks in this launch, ring slots
    int sub_pairs;       // pairs per sub-chunk (the last one may be shorter)
    int n_pairs;         // pairs in this launch
    int tiles;           // warp tiles per pair (expected consumed count = pairs of the sub-chunk x tiles)
};

__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------
// The warp-specialised plane pipeline, shared by k_plane_gain_ws and by the plane role of
// k_gain_fused.  Called by all 384 threads of a CTA; `walk0` is the CTA's work list.
// SYNC: hand the sub-chunks over through `rs` (S3 waits for the ring slot before the first store of a
// sub-chunk and publishes the sub-chunk after its last store -- also for sub-chunks in which this CTA
// has no entry).
// ---------------------------------------------------------------------------------------
template <int N, bool SYNC, class Walk>
__device__ __forceinline__ void
plane_ws_pipeline(const Walk &walk0, unsigned char *smem_raw, const cplx *__restrict__ fhat,
                  const cplx *__restrict__ phase, const cplx *__restrict__ zpm,
                  const cplx *__restrict__ twtab, cplx *__restrict__ hyb,
                  const cplx *__restrict__ nyq, const double *__restrict__ pair_w,
                  cplx *__restrict__ uvw, const RingSync &rs)
{
    constexpr int R = N / 4, GT = 128, PITCH = N + 1, H = N / 2, NBUF = 3;
    constexpr int T1 = GT;                // threads of stage S1
    constexpr int U1 = (4 * N) / T1;      // S1 units (row, residue) per thread and item
    constexpr int U = (4 * N) / GT;       // S2 / S3 units per thread and item
    constexpr int BAR_FULL1 = 1, BAR_FULL2 = 4, BAR_EMPTY = 7, BAR_S1 = 10, BAR_S3 = 11;
    constexpr int REG_S1 = WsRegs::S1, REG_S2 = WsRegs::S2, REG_S3 = WsRegs::S3;
    static_assert(REG_S1 + REG_S2 + REG_S3 <= 3 * 168, "register targets exceed what the launch allocates");
    static_assert(N == 64 && R == 16 && U == 2, "the pipelined plane kernel is written for N = 64");
    cplx *bufs = reinterpret_cast<cplx *>(smem_raw);          // NBUF x (N x PITCH)
    cplx *phs = bufs + NBUF * N * PITCH;                      // 2 x 4N (S1's phase tables: ex, ey, ez, zpm)
    cplx *tws = phs + 2 * 4 * N;                              // N twiddles exp(+2 pi i t/N)

    const int wg = threadIdx.x / GT;
    const int cnt = walk0.cnt;

    if (threadIdx.x < N) tws[threadIdx.x] = __ldg(&twtab[threadIdx.x]);
    __syncthreads();

    if (wg == 0) {
        // =========================== S1: phase-weighted fhat, radix-R along z ===================
        reg_alloc<REG_S1>();
        const int ts = threadIdx.x;            // 0 .. T1-1
        const int j = ts % N, b0 = ts / N;     // unit u: row j, residue b0 + (T1/N) u
        cplx fr[U1][R];                        // fr[u][a] = plane[j][4a + b_u]
        int cur_plane = -1;
        // the phase table of item n+1 is copied (cp.async) into the other slot while item n is computed
        auto stage_phase = [&](int pair_src, int slot_dst) {
            const cplx *src = phase + (size_t)pair_src * 3 * N;
            cplx *dstp = phs + slot_dst * 4 * N;
            if (ts < 3 * N) cp_async16(dstp + ts, src + ts);
            if (ts + GT < 3 * N) cp_async16(dstp + ts + GT, src + ts + GT);
            // fourth row: (Re+Im, Re-Im) of the z phase, taken by the last N threads
            if (ts >= T1 - N) cp_async16(dstp + 3 * N + (ts - (T1 - N)), zpm + (size_t)pair_src * N + (ts - (T1 - N)));
        };
        Walk wk = walk0;                       // plane and item of the current list entry
        if (cnt > 0) stage_phase(wk.pair, 0);
        cp_async_commit();
        cp_async_wait<0>();
        bar_sync_n(BAR_S1, T1);
        int buf_id = 0, slot = 0;
        for (int n = 0; n < cnt; ++n) {
            const int i = wk.i, pair = wk.pair;
            wk.next();                         // next entry: its phase table goes to the other slot
            if (n + 1 < cnt) stage_phase(wk.pair, slot ^ 1);
            cp_async_commit();
            cplx *buf = bufs + buf_id * N * PITCH;
            if (n >= NBUF) bar_sync_n(BAR_EMPTY + buf_id, T1 + GT); // S3 is done with item n - NBUF
            if (i != cur_plane) {
                // new plane: coalesced copy into the (free) pipeline buffer, then every thread picks
                // its entries -- the strided direct load costs ~10 us per plane change
                const cplx *srcp = (i < N) ? fhat + (size_t)i * N * N : nyq + (size_t)(i - N) * N * N;
#pragma unroll 8
                for (int e = ts; e < N * N; e += T1) buf[(e / N) * PITCH + (e % N)] = __ldg(&srcp[e]);
                bar_sync_n(BAR_S1, T1);
#pragma unroll
                for (int u = 0; u < U1; ++u)
#pragma unroll
                    for (int a = 0; a < R; ++a) fr[u][a] = buf[j * PITCH + 4 * a + b0 + (T1 / N) * u];
                bar_sync_n(BAR_S1, T1); // all entries are in registers before anybody overwrites buf
                cur_plane = i;
            }
            const cplx *P = phs + slot * 4 * N;

            if (i < N) {
                // m_H = A (Z.x+Z.y) + B (Z.x-Z.y), see k_plane_gain3; row 3 of P holds the two sums
                const cplx exi = P[i], eyj = P[N + j];
                const cplx X = cmul(exi, eyj);
                const cplx ext = (i == H) ? exi : make_double2(exi.x, -exi.y);
                const cplx eyt = (j == H) ? eyj : make_double2(eyj.x, -eyj.y);
                const cplx Xt = cmul(ext, eyt);
                const double cA = 0.5 * (X.x + Xt.x), cB = 0.5 * (X.y - Xt.y);
                const double nA = 0.5 * (X.x - Xt.y), nB = 0.5 * (X.y + Xt.x);
#pragma unroll
                for (int u = 0; u < U1; ++u) {
                    const int b = b0 + (T1 / N) * u;
                    cplx v[R];
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int k = 4 * a + b;
                        const cplx zz = P[3 * N + k];
                        const double zp = zz.x, zm = zz.y;
                        const bool ny = (a == R / 2) && (b == 0); // k == H
                        const double m = (ny ? nA : cA) * zp + (ny ? nB : cB) * zm;
                        const cplx f = fr[u][a];
                        v[a] = make_double2(m * f.x, m * f.y);
                    }
                    Dft<R, +1>::run(v);
                    cplx *row = buf + j * PITCH + R * b;
#pragma unroll
                    for (int k1 = 0; k1 < R; ++k1) row[k1] = v[dft_reg<R>(k1)];
                }
            } else {
                // Nyquist plane q = i - N: fixed axis q, free axes (axA rows, axB columns)
                const int nq = i - N;
                const int axA = (nq == 0) ? 1 : 0, axB = (nq == 2) ? 1 : 2;
                const cplx efix = P[nq * N + H];
                const double sw = 0.5 * sqrt(__ldg(&pair_w[pair]));
                const cplx ea = P[axA * N + j];
                const cplx eat = (j == H) ? ea : make_double2(ea.x, -ea.y);
                const cplx fa = cmul(efix, ea), fat = cmul(efix, eat);
                const bool zero_row = (nq >= 1) && (j == H);
#pragma unroll
                for (int u = 0; u < U1; ++u) {
                    const int b = b0 + (T1 / N) * u;
                    cplx v[R];
#pragma unroll
                    for (int a = 0; a < R; ++a) {
                        const int k = 4 * a + b;
                        const cplx eb = P[axB * N + k];
                        const cplx ebt = (k == H) ? eb : make_double2(eb.x, -eb.y);
                        const cplx e = cmul(fa, eb), et = cmul(fat, ebt);
                        double n2 = sw * ((e.x - et.x) + (e.y + et.y));
                        if (zero_row || (nq == 2 && k == H)) n2 = 0.0;
                        const cplx f = fr[u][a];
                        v[a] = make_double2(n2 * f.x, n2 * f.y);
                    }
                    Dft<R, +1>::run(v);
                    cplx *row = buf + j * PITCH + R * b;
#pragma unroll
                    for (int k1 = 0; k1 < R; ++k1) row[k1] = v[dft_reg<R>(k1)];
                }
            }
            bar_arrive_n(BAR_FULL1 + buf_id, T1 + GT);
            cp_async_wait<0>();
            bar_sync_n(BAR_S1, T1); // next table visible; everyone is done reading the current one
            slot ^= 1;
            buf_id = (buf_id + 1 == NBUF) ? 0 : buf_id + 1;
        }
        // absorb S3's releases of the last items so that every barrier ends balanced
        for (int n = (cnt > NBUF ? cnt - NBUF : 0); n < cnt; ++n)
            bar_sync_n(BAR_EMPTY + n % NBUF, T1 + GT);
    } else if (wg == 1) {
        // =========================== S2: 4 x 4 blocks in place ===================================
        reg_dealloc<REG_S2>();
        const int t = threadIdx.x - GT;
        // unit q = t + 128 u: b' = q % R = t % R, k1 = q / R = t / R + 8 u
        const int s2b = t % R;
        cplx wy[3];
#pragma unroll
        for (int m = 1; m < 4; ++m) wy[m - 1] = tws[(m * s2b) & (N - 1)];
        int buf_id = 0;
        for (int n = 0; n < cnt; ++n) {
            cplx *buf = bufs + buf_id * N * PITCH;
            bar_sync_n(BAR_FULL1 + buf_id, T1 + GT);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s2k = t / R + (GT / R) * u;
                cplx e[4][4]; // [a' (row R a'+b')][b (column k1 + R b)]
                cplx *blk = buf + s2b * PITCH + s2k;
#pragma unroll
                for (int ap = 0; ap < 4; ++ap)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) e[ap][bb] = blk[(R * ap) * PITCH + R * bb];
                cplx wz[3];
#pragma unroll
                for (int m = 1; m < 4; ++m) wz[m - 1] = tws[(m * s2k) & (N - 1)];
#pragma unroll
                for (int ap = 0; ap < 4; ++ap) {
#pragma unroll
                    for (int bb = 1; bb < 4; ++bb) e[ap][bb] = cmul(e[ap][bb], wz[bb - 1]);
                    dft4<+1>(e[ap][0], e[ap][1], e[ap][2], e[ap][3]);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    dft4<+1>(e[0][c], e[1][c], e[2][c], e[3][c]);
#pragma unroll
                    for (int kp = 1; kp < 4; ++kp) e[kp][c] = cmul(e[kp][c], wy[kp - 1]);
                }
#pragma unroll
                for (int kp = 0; kp < 4; ++kp)
#pragma unroll
                    for (int c = 0; c < 4; ++c) blk[(R * kp) * PITCH + R * c] = e[kp][c];
            }
            bar_arrive_n(BAR_FULL2 + buf_id, 2 * GT);
            buf_id = (buf_id + 1 == NBUF) ? 0 : buf_id + 1;
        }
    } else {
        // =========================== S3: radix-R along y, natural-order store ====================
        reg_dealloc<REG_S3>();
        const int t = threadIdx.x - 2 * GT;
        // unit q = t + 128 u: column slot q % N = t % N, k1' = q / N = t / N + 2 u
        const int s3slot = t % N;
        int buf_id = 0;
        Walk wk = walk0;
        // fused hand-over: sub-chunks [0, published) have been published by this CTA; `cur_sub` is the
        // sub-chunk whose ring slot this CTA may write
        int published = 0, cur_sub = -1;
        auto publish_upto = [&](int s_end) { // all stores of sub-chunks < s_end are issued
            if (published < s_end) {
                __threadfence();
                bar_sync_n(BAR_S3, GT);
                if (t == 0)
                    for (int s = published; s < s_end; ++s) atomicAdd(rs.ready + s, 1);
                published = s_end;
            }
        };
        for (int n = 0; n < cnt; ++n) {
            const cplx *buf = bufs + buf_id * N * PITCH;
            if (SYNC && wk.sub != cur_sub) {
                publish_upto(wk.sub);
                cur_sub = wk.sub;
                if (cur_sub >= rs.ring) {
                    // the ring slot still holds sub-chunk cur_sub - ring: wait until it has been read
                    const int old = cur_sub - rs.ring;
                    const int old_pairs = min(rs.sub_pairs, rs.n_pairs - old * rs.sub_pairs);
                    if (t == 0) {
                        while (ld_relaxed(rs.consumed + old) < old_pairs * rs.tiles) __nanosleep(100);
                        __threadfence();
                    }
                    bar_sync_n(BAR_S3, GT);
                }
            }
            cplx *dst = (wk.i < N) ? hyb + ((size_t)wk.dst_item * N + wk.i) * N * N
                                   : uvw + ((size_t)wk.dst_item * 3 + (wk.i - N)) * N * N;
            wk.next();
            bar_sync_n(BAR_FULL2 + buf_id, 2 * GT);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int s3k = t / N + (GT / N) * u;
                cplx v[R];
                const cplx *col = buf + (R * s3k) * PITCH + s3slot;
#pragma unroll
                for (int bp = 0; bp < R; ++bp) v[bp] = col[bp * PITCH];
                Dft<R, +1>::run(v);
#pragma unroll
                for (int k2 = 0; k2 < R; ++k2) dst[(s3k + 4 * k2) * N + s3slot] = v[dft_reg<R>(k2)];
            }
            bar_arrive_n(BAR_EMPTY + buf_id, T1 + GT); // after the stores that consumed the loads
            buf_id = (buf_id + 1 == NBUF) ? 0 : buf_id + 1;
        }
        if (SYNC) publish_upto(rs.n_sub);
    }
}

template <int N>
__global__ void __launch_bounds__(3 * 128, 1)
k_plane_gain_ws(const cplx *__restrict__ fhat, const cplx *__restrict__ phase,
                const cplx *__restrict__ zpm, const cplx *__restrict__ twtab,
                cplx *__restrict__ hyb, int pair0, int n_items,
                const cplx *__restrict__ nyq, const double *__restrict__ pair_w,
                cplx *__restrict__ uvw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LaunchWalk<N> walk;
    walk.init(n_items, pair0, blockIdx.x, gridDim.x);
    RingSync none = {};
    plane_ws_pipeline<N, false>(walk, smem_raw, fhat, phase, zpm, twtab, hyb, nyq, pair_w, uvw, none);
}

// ---------------------------------------------------------------------------------------
// k_pencil_gain (UNPACKED mode): grid (N*N/TZ tiles, G), block PG*(B*TZ).  CTA (tile, gy) owns chunk
// pairs [lo,hi) = share gy of the chunk; its PG groups take them round-robin.  Each group:
// inverse x-FFT of g1' and g2' pencils, acc += w * Re(g1 g2) (cpp:233-246 + linearity of
// the forward FFT).  At every change of radius r the PG partial sums are reduced through
// shared memory in fixed order and added to S[gy][r] -- no atomics, deterministic.
// ---------------------------------------------------------------------------------------
template <int N, int PG, int MINB>
__global__ void __launch_bounds__(PG *Geo<N>::B *TZ, MINB)
k_pencil_gain(const cplx *__restrict__ hyb, const cplx *__restrict__ twtab,
              const int *__restrict__ pair_r, const double *__restrict__ pair_w,
              const int *__restrict__ r_end, double *__restrict__ S, int pair0, int n_pairs_chunk,
              int n_r_local)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B;
    constexpr int TGP = B * TZ, UNITS = X2<N>::UNITS, TILE = N * TZ;
    constexpr size_t N3 = (size_t)N * N * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx(*sm)[2][TILE] = reinterpret_cast<cplx(*)[2][TILE]>(smem_raw); // [PG][2][TILE]

    const int g = threadIdx.x / TGP, tg = threadIdx.x % TGP;
    const int y = blockIdx.x / (N / TZ), zg = blockIdx.x % (N / TZ);
    const int G = gridDim.y;
    const int lo = (int)(((long long)n_pairs_chunk * blockIdx.y) / G);
    const int hi = (int)(((long long)n_pairs_chunk * (blockIdx.y + 1)) / G);

    cplx tw[A - 1];
    load_twiddles<N, +1>(tw, twtab, tg / TZ);

    double acc[UNITS][B];
#pragma unroll
    for (int m = 0; m < UNITS; ++m)
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) acc[m][k2] = 0.0;

    const size_t tile_off = (size_t)y * N + zg * TZ;
    int p = lo;
    while (p < hi) {
        const int r = pair_r[pair0 + p];
        int seg_end = r_end[r] - pair0;
        if (seg_end > hi) seg_end = hi;
        for (int q = p + g; q < seg_end; q += PG) {
            const double w = pair_w[pair0 + q];
            const cplx *g1 = hyb + (size_t)(2 * q) * N3 + tile_off;
            const cplx *g2 = g1 + N3;
            x1_pass<N, +1>(sm[g][0], tw, tg, [&](int x, int z) { return g1[(size_t)x * N * N + z]; });
            x1_pass<N, +1>(sm[g][1], tw, tg, [&](int x, int z) { return g2[(size_t)x * N * N + z]; });
            group_sync(1 + g, TGP);
#pragma unroll
            for (int m = 0; m < UNITS; ++m) {
                cplx v0[B], v1[B];
                x2_unit<N, +1>(sm[g][0], tg, m, v0);
                x2_unit<N, +1>(sm[g][1], tg, m, v1);
#pragma unroll
                for (int k2 = 0; k2 < B; ++k2)
                    acc[m][k2] += w * (v0[k2].x * v1[k2].x - v0[k2].y * v1[k2].y);
            }
            group_sync(1 + g, TGP);
        }
        // flush this radius: fixed-order reduction over the PG groups
        __syncthreads();
        double *red = reinterpret_cast<double *>(&sm[0][0][0]); // PG*TILE doubles <= smem size
#pragma unroll
        for (int m = 0; m < UNITS; ++m)
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) {
                red[g * TILE + (tg + m * TGP) * B + k2] = acc[m][k2];
                acc[m][k2] = 0.0;
            }
        __syncthreads();
        double *Sr = S + ((size_t)blockIdx.y * n_r_local + r) * N3 + tile_off;
        for (int e = threadIdx.x; e < TILE; e += PG * TGP) {
            const int x = e / TZ, z = e % TZ;
            const int k1 = x % A, k2 = x / A;
            const int src = (k1 * TZ + z) * B + k2;
            double s = 0.0;
#pragma unroll
            for (int gg = 0; gg < PG; ++gg) s += red[gg * TILE + src];
            Sr[(size_t)x * N * N + z] += s;
        }
        __syncthreads();
        p = seg_end;
    }
}

// ---------------------------------------------------------------------------------------
// k_pencil_gain_async (packed mode): same work split and flush as k_pencil_gain, but each group
// streams its tiles through a STAGES-deep ring of shared-memory slots, so the global loads of the
// next pairs are in flight while the current pair is transformed.  Per pair and group: wait for
// the oldest slot, x pass 1 in place, x pass 2, acc += w (Re^2 - Im^2); two group barriers per pair.
//
// TMA = true: the ring is filled by the TMA unit -- a tile is the box (TZ z, 1 y, N x) of the hybrid
// scratch seen as the 3-D tensor [pair*N + x][y][z] (8 KiB, rows of 128 contiguous bytes N^2 elements
// apart); ONE thread of the group issues ONE cp.async.bulk.tensor (SASS UTMALDG) per tile and the
// slot's mbarrier counts the bytes in -- no LSU issue slots and no address arithmetic for the copy.
// TMA = false (default): cp.async (LDGSTS) fill, 8 copies per thread and tile.  Measured at 64^3 the
// x stage is HBM bound either way (profiles/r02_ab64_tma.log).
// ---------------------------------------------------------------------------------------
// ONE_SLOT: every CTA row gy accumulates into partial slot 0 -- legal when the gy shares of the chunk
// start at radius boundaries (then no two CTAs touch the same (radius, tile)); the host checks that.
template <int N, int PG, int STAGES, int MINB, bool ONE_SLOT = false, bool BULK = false>
__global__ void __launch_bounds__(PG *Geo<N>::B *TZ, MINB)
k_pencil_gain_async(const cplx *__restrict__ hyb, const cplx *__restrict__ twtab,
                    const int *__restrict__ pair_r, const double *__restrict__ pair_w,
                    const int *__restrict__ r_end, double *__restrict__ S, int pair0,
                    int n_pairs_chunk, int n_r_local, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B;
    constexpr int TGP = B * TZ, UNITS = X2<N>::UNITS, TILE = N * TZ;
    constexpr int CP_PER_THREAD = TILE / TGP; // 16-byte copies per thread and tile (LDGSTS fill)
    constexpr size_t N3 = (size_t)N * N * N;
    extern __shared__ __align__(128) unsigned char smem_ring[]; // TMA destinations: 128-byte aligned
    cplx(*ring)[STAGES][TILE] = reinterpret_cast<cplx(*)[STAGES][TILE]>(smem_ring); // [PG][STAGES][TILE]
    __shared__ __align__(8) unsigned long long full[PG][STAGES]; // TMA: "tile landed" barriers

    const int g = threadIdx.x / TGP, tg = threadIdx.x % TGP;
    const int y = blockIdx.x / (N / TZ), zg = blockIdx.x % (N / TZ);
    const int G = gridDim.y;
    const int lo = (int)(((long long)n_pairs_chunk * blockIdx.y) / G);
    const int hi = (int)(((long long)n_pairs_chunk * (blockIdx.y + 1)) / G);

    if (BULK) {
        if (threadIdx.x < PG * STAGES) mbar_init(&full[0][0] + threadIdx.x, 1);
        mbar_init_fence();
        __syncthreads();
    }

    cplx tw[A - 1];
    load_twiddles<N, +1>(tw, twtab, tg / TZ);

    double acc[UNITS][B];
#pragma unroll
    for (int m = 0; m < UNITS; ++m)
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) acc[m][k2] = 0.0;

    const size_t tile_off = (size_t)y * N + zg * TZ;
    // tile element e = x*TZ + z  <->  global x*N*N + z
    auto issue = [&](int q, int slot) {
        const cplx *src = hyb + (size_t)q * N3 + tile_off;
        cplx *dst = ring[g][slot];
        if (BULK) {
            if (tg == 0) {
                // coordinates in doubles / rows: (2 z0, y, pair * N)
                mbar_expect_tx(&full[g][slot], TILE * sizeof(cplx));
                tma_load_3d(dst, &tmap, 2 * zg * TZ, y, q * N, &full[g][slot]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < CP_PER_THREAD; ++c) {
                const int e = tg + c * TGP;
                cp_async16(dst + e, src + (size_t)(e / TZ) * N * N + (e % TZ));
            }
        }
    };
    unsigned used = 0; // BULK: tiles consumed so far by this group (slot = used % STAGES, parity from used / STAGES)

    int p = lo;
    // the pair list is r-major, so the radius segments of [lo,hi) are consecutive radii: only the
    // first radius index is looked up, and the end of the NEXT segment is fetched one segment ahead
    // (two dependent global loads per segment and one per pair used to sit on the critical path:
    // ncu long_scoreboard 29 % of the stall samples)
    int r = (lo < hi) ? __ldg(&pair_r[pair0 + lo]) : 0;
    int r_end_cur = (lo < hi) ? __ldg(&r_end[r]) : 0;
    while (p < hi) {
        int seg_end = r_end_cur - pair0;
        if (seg_end > hi) seg_end = hi;
        if (seg_end < hi) r_end_cur = __ldg(&r_end[r + 1]); // consumed at the next segment
        // this group's pairs: q_n = p + g + n*PG
        const int n_mine = (seg_end - p - g + PG - 1) / PG;
        // BULK: the ring slots are used round robin over the whole kernel (`used`), so that every
        // barrier's phase parity is (use count / STAGES) & 1
        const unsigned base = used;
#pragma unroll
        for (int s0 = 0; s0 < STAGES - 1; ++s0) {
            if (s0 < n_mine) issue(p + g + s0 * PG, BULK ? (base + s0) % STAGES : s0);
            if (!BULK) cp_async_commit();
        }
        for (int n = 0; n < n_mine; ++n) {
            const int q = p + g + n * PG;
            const double w = __ldg(&pair_w[pair0 + q]); // in flight during the wait and x pass 1
            const int slot = BULK ? (int)((base + n) % STAGES) : n % STAGES;
            if (BULK) {
                mbar_wait(&full[g][slot], ((base + n) / STAGES) & 1); // pair n has landed
                fence_proxy_async();           // my generic writes to the slot freed below come first
                group_sync(1 + g, TGP);        // everybody is done with pair n-1: its slot is free again
            } else {
                cp_async_wait<STAGES - 2>();   // this thread's copies for pair n have landed
                group_sync(1 + g, TGP);        // ... and everybody else's; slot (n-1) is free again
            }
            if (n + STAGES - 1 < n_mine)
                issue(q + (STAGES - 1) * PG, BULK ? (int)((base + n + STAGES - 1) % STAGES) : (n + STAGES - 1) % STAGES);
            if (!BULK) cp_async_commit();
            cplx *sm = ring[g][slot];
            x1_pass_inplace<N, +1>(sm, tw, tg);
            group_sync(1 + g, TGP);
#pragma unroll
            for (int m = 0; m < UNITS; ++m) {
                cplx v0[B];
                x2_unit<N, +1>(sm, tg, m, v0);
#pragma unroll
                for (int k2 = 0; k2 < B; ++k2)
                    acc[m][k2] += w * (v0[k2].x * v0[k2].x - v0[k2].y * v0[k2].y);
            }
        }
        used += (unsigned)n_mine;
        if (!BULK) cp_async_wait<0>();
        // flush this radius: fixed-order reduction over the PG groups
        if (BULK) fence_proxy_async(); // the reduction buffer below aliases ring slots the TMA unit wrote
        __syncthreads();
        double *red = reinterpret_cast<double *>(smem_ring); // PG*TILE doubles <= ring size
#pragma unroll
        for (int m = 0; m < UNITS; ++m)
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) {
                red[g * TILE + (tg + m * TGP) * B + k2] = acc[m][k2];
                acc[m][k2] = 0.0;
            }
        __syncthreads();
        double *Sr = S + ((size_t)(ONE_SLOT ? 0 : blockIdx.y) * n_r_local + r) * N3 + tile_off;
        for (int e = threadIdx.x; e < TILE; e += PG * TGP) {
            const int x = e / TZ, z = e % TZ;
            const int k1 = x % A, k2 = x / A;
            const int src = (k1 * TZ + z) * B + k2;
            double s = 0.0;
#pragma unroll
            for (int gg = 0; gg < PG; ++gg) s += red[gg * TILE + src];
            Sr[(size_t)x * N * N + z] += s;
        }
        if (BULK) fence_proxy_async(); // generic writes to `red` precede the next segment's bulk copies
        __syncthreads();
        p = seg_end;
        ++r;
    }
}

// ---------------------------------------------------------------------------------------
// Generic single-shot plane kernel: grid (N planes, n_items[, cells]), block TG = N*B.
//   PLANE_REAL : in = sum_{g<n_partials} src_real[g*partial_stride + item*N^3 + ...] (imag 0)
//   PLANE_FINAL: item 0: in = Qhat;  item 1: in = beta2[|l|^2] * fhat   (cpp:281-299)
// out: dst[item][plane][y][z] in natural order.
// ---------------------------------------------------------------------------------------
enum { PLANE_REAL = 0, PLANE_FINAL = 1 };

template <int N, int SIGN, int MODE>
__global__ void __launch_bounds__(N *Geo<N>::B)
k_plane(const double *__restrict__ src_real, int n_partials, size_t partial_stride,
        const cplx *__restrict__ qhat, const cplx *__restrict__ fhat,
        const double *__restrict__ beta2, const cplx *__restrict__ twtab, cplx *__restrict__ dst,
        const int *__restrict__ n_partials_item = nullptr, const double *__restrict__ src_real2 = nullptr,
        int n_partials2 = 0, const int *__restrict__ n_partials2_item = nullptr,
        size_t src_cell_stride = 0, size_t dst_cell_stride = 0)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, TG = N * B;
    constexpr size_t N3 = (size_t)N * N * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *buf = reinterpret_cast<cplx *>(smem_raw);
    const int i = blockIdx.x, item = blockIdx.y, tg = threadIdx.x;
    // batch of cells (grid.z): every cell has its own source, spectrum and destination
    const size_t cell = blockIdx.z;
    if (MODE == PLANE_REAL) {
        src_real += cell * src_cell_stride;
        if (src_real2) src_real2 += cell * src_cell_stride;
    } else {
        qhat += cell * N3;
        fhat += cell * N3;
    }
    dst += cell * dst_cell_stride;

    cplx tw[A - 1];
    load_twiddles<N, SIGN>(tw, twtab, tg % B);

    if (MODE == PLANE_REAL) {
        // fixed-order sum of the partial slots: `n_partials` (or this item's own count) slots of
        // src_real, then n_partials2 (or this item's own count) slots of src_real2
        const double *s = src_real + (size_t)item * N3 + (size_t)i * N * N;
        const double *s2 = src_real2 + (size_t)item * N3 + (size_t)i * N * N;
        const int np = n_partials_item ? __ldg(&n_partials_item[item]) : n_partials;
        const int np2 = n_partials2_item ? __ldg(&n_partials2_item[item]) : n_partials2;
        // z pass 1 with the slot loop OUTSIDE the element loop: the A loads of a slot are independent, so a
        // thread has A (x2, unrolled) loads in flight instead of one (the per-element accumulation loop made
        // the kernel a chain of 40 round trips: 171 us for 450 MB at 64^3).  Same summation order per element.
        static_assert(TG == N * B, "one z-pass-1 unit per thread");
        const int j = tg / B, b = tg % B;
        double accv[A];
#pragma unroll
        for (int a = 0; a < A; ++a) accv[a] = 0.0;
        const double *sj = s + j * N + b, *sj2 = s2 + j * N + b;
#pragma unroll 2
        for (int gq = 0; gq < np; ++gq) {
#pragma unroll
            for (int a = 0; a < A; ++a) accv[a] += sj[(size_t)gq * partial_stride + B * a];
        }
#pragma unroll 2
        for (int gq = 0; gq < np2; ++gq) {
#pragma unroll
            for (int a = 0; a < A; ++a) accv[a] += sj2[(size_t)gq * partial_stride + B * a];
        }
        cplx v[A];
#pragma unroll
        for (int a = 0; a < A; ++a) v[a] = make_double2(accv[a], 0.0);
        Dft<A, SIGN>::run(v);
        cplx *row = buf + j * Geo<N>::ROW;
        row[padk(b)] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < A; ++k1) row[padk(B * k1 + b)] = cmul(v[k1], tw[k1 - 1]);
    } else {
        const size_t off = (size_t)i * N * N;
        const int li = mode_of<N>(i);
        if (item == 0) {
            z1_pass<N, SIGN, TG>(buf, tw, tg, [&](int j, int k) { return qhat[off + j * N + k]; });
        } else {
            z1_pass<N, SIGN, TG>(buf, tw, tg, [&](int j, int k) {
                const int lj = mode_of<N>(j), lk = mode_of<N>(k);
                const double b2 = __ldg(&beta2[li * li + lj * lj + lk * lk]);
                const cplx f = fhat[off + j * N + k];
                return make_double2(b2 * f.x, b2 * f.y);
            });
        }
    }
    __syncthreads();
    z2_pass<N, SIGN, TG>(buf, tg);
    __syncthreads();
    y1_pass<N, SIGN, TG>(buf, tw, tg);
    __syncthreads();
    cplx *d = dst + (size_t)item * N3 + (size_t)i * N * N;
    y2_pass<N, SIGN, TG>(buf, tg, [&](int y, int z, cplx val) { d[y * N + z] = val; });
}

// ---------------------------------------------------------------------------------------
// Single-shot pencil kernels: grid N*N/TZ tiles, block B*TZ.
// ---------------------------------------------------------------------------------------

// fhat[i][j][k] = scale * FFT_x(Fh[.][j][k])           (forward transform of f, cpp:186)
template <int N>
__global__ void __launch_bounds__(Geo<N>::B *TZ)
k_pencil_fwd(const cplx *__restrict__ Fh, const cplx *__restrict__ twtab, double scale,
             cplx *__restrict__ fhat, cplx *__restrict__ nyq = nullptr)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, UNITS = X2<N>::UNITS;
    __shared__ __align__(16) cplx sm[N * TZ];
    const int tg = threadIdx.x;
    const int j = blockIdx.x / (N / TZ), kg = blockIdx.x % (N / TZ);
    const size_t off = (size_t)j * N + kg * TZ;
    // batch of cells (grid.y)
    Fh += (size_t)blockIdx.y * N * N * N;
    fhat += (size_t)blockIdx.y * N * N * N;
    if (nyq) nyq += (size_t)blockIdx.y * 3 * N * N;
    cplx tw[A - 1];
    load_twiddles<N, -1>(tw, twtab, tg / TZ);
    x1_pass<N, -1>(sm, tw, tg, [&](int x, int z) { return Fh[off + (size_t)x * N * N + z]; });
    __syncthreads();
#pragma unroll
    for (int m = 0; m < UNITS; ++m) {
        cplx v[B];
        const int k1 = x2_unit<N, -1>(sm, tg, m, v);
        const int z = (tg + m * B * TZ) % TZ;
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            const int i = k1 + A * k2;
            const cplx o = make_double2(v[k2].x * scale, v[k2].y * scale);
            fhat[off + (size_t)i * N * N + z] = o;
            if (nyq) { // the three Nyquist planes of fhat (packed mode), see k_extract_nyq
                constexpr int H = N / 2;
                const int kk = kg * TZ + z;
                if (i == H) nyq[(size_t)j * N + kk] = o;
                if (j == H) nyq[(size_t)N * N + (size_t)i * N + kk] = o;
                if (kk == H) nyq[(size_t)2 * N * N + (size_t)i * N + j] = o;
            }
        }
    }
}

// Qhat[l] = sum_r coef[r][|l|^2] * FFT_x(Ph_r)[l]      (cpp:252-273, register reduction over r)
// A tile is only B*TZ threads wide, and the loop over the radii is a chain of small dependent steps: with
// one group per CTA the kernel ran one or two warps per SM and was bound by instruction latency (32^3,
// 16 radii: 21.7 us for 8 MiB).  So RS groups per CTA take the radii r = g, g+RS, ... concurrently, each
// with its own double-buffered tile (one barrier per radius), the entries and coefficients of a group's
// next radius are loaded while the current one is transformed, and the groups' sums are added in the
// fixed order g = 0..RS-1 through shared memory (deterministic).
template <int N> struct AccumGeo { static constexpr int RS = (N == 64) ? 2 : 4; };

template <int N>
__global__ void __launch_bounds__(AccumGeo<N>::RS *Geo<N>::B *TZ)
k_pencil_accum(const cplx *__restrict__ Ph, const cplx *__restrict__ twtab,
               const double *__restrict__ coef, int n_r_local, int M, cplx *__restrict__ Qhat)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, UNITS = X2<N>::UNITS, RS = AccumGeo<N>::RS, TG = B * TZ;
    constexpr size_t N3 = (size_t)N * N * N;
    static_assert(UNITS * B * TG <= 2 * N * TZ, "the reduction reuses a group's tile buffers");
    __shared__ __align__(16) cplx sm[RS][2][N * TZ];
    const int g = threadIdx.x / TG, tg = threadIdx.x % TG;
    const int j = blockIdx.x / (N / TZ), kg = blockIdx.x % (N / TZ);
    const size_t off = (size_t)j * N + kg * TZ;
    // batch of cells (grid.y): cell c reads Ph[c][r], writes Qhat[c]
    Ph += (size_t)blockIdx.y * n_r_local * N3;
    Qhat += (size_t)blockIdx.y * N3;
    cplx tw[A - 1];
    load_twiddles<N, -1>(tw, twtab, tg / TZ);

    cplx acc[UNITS][B];
    int msq[UNITS][B];
#pragma unroll
    for (int m = 0; m < UNITS; ++m) {
        const int u = tg + m * TG;
        const int z = u % TZ, k1 = u / TZ;
        const int lj = mode_of<N>(j), lk = mode_of<N>(kg * TZ + z);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            const int li = mode_of<N>(k1 + A * k2);
            msq[m][k2] = li * li + lj * lj + lk * lk;
            acc[m][k2] = make_double2(0.0, 0.0);
        }
    }
    // this thread's pass-1 entries x = B a + b (b = tg / TZ) and coefficients of one radius
    const int zt = tg % TZ, bt = tg / TZ;
    cplx nxt[A];
    double cnx[UNITS][B];
    auto fetch = [&](int r) {
        const cplx *src = Ph + (size_t)r * N3 + off + zt;
#pragma unroll
        for (int a = 0; a < A; ++a) nxt[a] = __ldg(&src[(size_t)(B * a + bt) * N * N]);
#pragma unroll
        for (int m = 0; m < UNITS; ++m)
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) cnx[m][k2] = __ldg(&coef[(size_t)r * M + msq[m][k2]]);
    };
    auto sync_group = [&]() {
        if constexpr (TG == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(TG) : "memory");
    };
    if (g < n_r_local) fetch(g);
    int it = 0;
    for (int r = g; r < n_r_local; r += RS, ++it) {
        cplx cur[A];
        double cc[UNITS][B];
#pragma unroll
        for (int a = 0; a < A; ++a) cur[a] = nxt[a];
#pragma unroll
        for (int m = 0; m < UNITS; ++m)
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) cc[m][k2] = cnx[m][k2];
        if (r + RS < n_r_local) fetch(r + RS);
        cplx *tile = sm[g][it & 1];
        // x pass 1 (x1_pass with the operands already in registers)
        Dft<A, -1>::run(cur);
        tile[bt * TZ + zt] = cur[0];
#pragma unroll
        for (int k1 = 1; k1 < A; ++k1) tile[(B * k1 + bt) * TZ + zt] = cmul(cur[k1], tw[k1 - 1]);
        sync_group(); // this tile is complete; the group's other buffer was read before this barrier
#pragma unroll
        for (int m = 0; m < UNITS; ++m) {
            cplx v[B];
            x2_unit<N, -1>(tile, tg, m, v);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) {
                acc[m][k2].x += cc[m][k2] * v[k2].x;
                acc[m][k2].y += cc[m][k2] * v[k2].y;
            }
        }
    }
    // fixed-order sum over the groups: groups 1.. park their sums in their own tile buffers
    __syncthreads();
    cplx *park = &sm[g][0][0];
    if (g > 0) {
#pragma unroll
        for (int m = 0; m < UNITS; ++m)
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) park[(m * B + k2) * TG + tg] = acc[m][k2];
    }
    __syncthreads();
    if (g == 0) {
#pragma unroll
        for (int m = 0; m < UNITS; ++m) {
            const int u = tg + m * TG;
            const int z = u % TZ, k1 = u / TZ;
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) {
                cplx t = acc[m][k2];
#pragma unroll
                for (int gg = 1; gg < RS; ++gg) {
                    const cplx o = sm[gg][0][(m * B + k2) * TG + tg];
                    t.x += o.x;
                    t.y += o.y;
                }
                Qhat[off + (size_t)(k1 + A * k2) * N * N + z] = t;
            }
        }
    }
}

// Q = Re(IFFT_x H0) - Re(IFFT_x H1) * f                 (cpp:304-330)
// WITH_LOSS = false: Q = Re(IFFT_x H0) only -- the partial gain of one pair shard; the shards' Q are
// summed by the all-reduce and exactly one rank adds the loss term (bfsm_collide_sharded).
template <int N, bool WITH_LOSS>
__global__ void __launch_bounds__(Geo<N>::B *TZ)
k_pencil_final(const cplx *__restrict__ H, const cplx *__restrict__ twtab,
               const double *f, double *Q)
{
    constexpr int A = Geo<N>::A, B = Geo<N>::B, UNITS = X2<N>::UNITS;
    constexpr size_t N3 = (size_t)N * N * N;
    __shared__ __align__(16) cplx sm[WITH_LOSS ? 2 : 1][N * TZ];
    const int tg = threadIdx.x;
    const int y = blockIdx.x / (N / TZ), zg = blockIdx.x % (N / TZ);
    const size_t off = (size_t)y * N + zg * TZ;
    // batch of cells (grid.y): cell c reads H[c][0..1], f[c], writes Q[c]
    H += (size_t)blockIdx.y * (WITH_LOSS ? 2 : 1) * N3;
    f += (size_t)blockIdx.y * N3;
    Q += (size_t)blockIdx.y * N3;
    cplx tw[A - 1];
    load_twiddles<N, +1>(tw, twtab, tg / TZ);
    x1_pass<N, +1>(sm[0], tw, tg, [&](int x, int z) { return H[off + (size_t)x * N * N + z]; });
    if (WITH_LOSS)
        x1_pass<N, +1>(sm[WITH_LOSS ? 1 : 0], tw, tg,
                       [&](int x, int z) { return H[N3 + off + (size_t)x * N * N + z]; });
    __syncthreads();
#pragma unroll
    for (int m = 0; m < UNITS; ++m) {
        cplx v0[B], v1[B];
        const int k1 = x2_unit<N, +1>(sm[0], tg, m, v0);
        if (WITH_LOSS) x2_unit<N, +1>(sm[WITH_LOSS ? 1 : 0], tg, m, v1);
        const int z = (tg + m * B * TZ) % TZ;
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            const size_t idx = off + (size_t)(k1 + A * k2) * N * N + z;
            if (WITH_LOSS) {
                const double fv = f[idx]; // read before the (possibly aliased) write below
                Q[idx] = v0[k2].x - v1[k2].x * fv;
            } else {
                Q[idx] = v0[k2].x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Nyquist-plane correction of the PACKED mode.
// ---------------------------------------------------------------------------------------

// nyq[q][a][b]: the three Nyquist planes of fhat, q = 0: fhat(H,a,b), 1: fhat(a,H,b), 2: fhat(a,b,H) --
// filed by k_pencil_fwd as it writes fhat.

// k_nyq_accum: S2[slot][r] = sum_s Re(Y_s^2) over the pairs of one work unit,
//   Y_s = (-1)^x U_s(y,z) + (-1)^y V_s(x,z) + (-1)^z W_s(x,y)     (sqrt(w_s) already folded into U,V,W).
// grid (tiles of 16^3 outputs, work units of the launch); block 256 threads, each owning the 2 x 2 x 4
// brick x in {bx,bx+1}, y in {by,by+1}, z in {zq, zq+4, zq+8, zq+12} (even bx, by: the x/y signs are
// compile-time; the z sign is a per-thread constant).  A work unit is a run of pairs of ONE radius
// (the same unit table as the register-resident x stage, bfsm_pencil_reg.cuh); CTA (tile, unit) stores
// its sums into partial slot `slot` of that radius -- every (slot, radius, tile) is written exactly
// once, so there is nothing to clear and no read-modify-write.  The 16 x 16 slices of U, V, W stream
// through a 3-stage cp.async ring: one barrier per pair.
template <int N>
__global__ void __launch_bounds__(256, 2)
k_nyq_accum(const cplx *__restrict__ uvw, int uvw_pair0, const PencilUnit *__restrict__ units,
            double *__restrict__ S2, int n_r_local, int pairs_per_cell = 0, size_t S_cell_stride = 0)
{
    // batch of cells (grid.z): cell c's fields of the launch follow those of cell c-1
    uvw += (size_t)blockIdx.z * pairs_per_cell * 3 * N * N;
    S2 += (size_t)blockIdx.z * S_cell_stride;
    constexpr int NT = 16, TPD = N / NT, PADR = NT + 2, STAGES = 3;
    constexpr size_t N3 = (size_t)N * N * N;
    __shared__ __align__(16) cplx ring[STAGES][3][NT][PADR]; // [stage][U|V|W][row][col]
    const int t = threadIdx.x;
    const int tx0 = (blockIdx.x / (TPD * TPD)) * NT, ty0 = ((blockIdx.x / TPD) % TPD) * NT,
              tz0 = (blockIdx.x % TPD) * NT;
    const int zq = t % 4, by = 2 * ((t / 4) % 8), bx = 2 * (t / 32);
    const double sz = (zq & 1) ? -1.0 : 1.0;
    const int la = t / NT, lb = t % NT; // element of the 16 x 16 slices this thread stages
    const PencilUnit un = units[blockIdx.y];
    const int n_mine = un.p1 - un.p0;

    // U[y][z], V[x][z], W[x][y] tiles of pair un.p0 + n -> ring slot
    auto issue = [&](int n, int slot) {
        const cplx *base = uvw + (size_t)(un.p0 - uvw_pair0 + n) * 3 * N * N;
        cp_async16(&ring[slot][0][la][lb], base + (size_t)(ty0 + la) * N + tz0 + lb);
        cp_async16(&ring[slot][1][la][lb], base + (size_t)N * N + (size_t)(tx0 + la) * N + tz0 + lb);
        cp_async16(&ring[slot][2][la][lb], base + (size_t)2 * N * N + (size_t)(tx0 + la) * N + ty0 + lb);
    };

    double acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0;
#pragma unroll
    for (int s0 = 0; s0 < STAGES - 1; ++s0) {
        if (s0 < n_mine) issue(s0, s0);
        cp_async_commit();
    }
    for (int n = 0; n < n_mine; ++n) {
        cp_async_wait<STAGES - 2>();
        __syncthreads(); // pair n landed for everybody; slot of pair n-1 is free again
        if (n + STAGES - 1 < n_mine) issue(n + STAGES - 1, (n + STAGES - 1) % STAGES);
        cp_async_commit();
        const cplx(*sU)[PADR] = ring[n % STAGES][0];
        const cplx(*sV)[PADR] = ring[n % STAGES][1];
        const cplx(*sW)[PADR] = ring[n % STAGES][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const cplx w0 = sW[bx + a][by + b];
                // fold the x sign into everything: Y^2 is even in Y, so use (-1)^x Y
                const double wr = (a ? -sz : sz) * w0.x, wi = (a ? -sz : sz) * w0.y;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const cplx uu = sU[by + b][zq + 4 * c], vv = sV[bx + a][zq + 4 * c];
                    // (-1)^x Y = U + (-1)^(x+y) V + (-1)^(x+z) W
                    const double yr = uu.x + ((a ^ b) ? -vv.x : vv.x) + wr;
                    const double yi = uu.y + ((a ^ b) ? -vv.y : vv.y) + wi;
                    acc[a][b][c] = fma(yr, yr, acc[a][b][c]);
                    acc[a][b][c] = fma(-yi, yi, acc[a][b][c]);
                }
            }
    }
    cp_async_wait<0>();
    double *Sr = S2 + ((size_t)un.slot * n_r_local + un.r) * N3;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                Sr[((size_t)(tx0 + bx + a) * N + ty0 + by + b) * N + tz0 + zq + 4 * c] = acc[a][b][c];
}

} // namespace bfsm
