// bfsm_general.cuh -- the collision path for grids the tuned kernels are not written for: independent
// Nvx, Nvy, Nvz (the reference interface carries three sizes, FFTWBoltzmannOperator.hpp:30-36, with
// per-axis mode tables, .cpp:46-57), any even size from 4 to 128 per axis, powers of two or not.
//
// Same algorithm and the same exact restructurings as the tuned path except the Hermitian packing
// (antipodal folding, one forward transform per radius, separable phase tables, beta tables by |l|^2),
// built from plain pieces: a batched in-place 1-D transform along one axis (radix-2 in shared memory for
// powers of two, a direct DFT otherwise) and a few pointwise kernels.  It moves ~10x the bytes of the
// tuned path per pair and is meant for coverage, not for the headline sizes (cubic 16/32/64 never come
// here).  No atomics: every accumulation has one owner thread and a fixed order.
#pragma once
#include "bfsm_fft.cuh"

namespace bfsm {

constexpr int GEN_TL = 8;        // lines per CTA of the axis transform
constexpr int GEN_MAXLEN = 128;  // longest axis

__device__ __forceinline__ int gen_mode(int t, int n) { return t < n / 2 ? t : t - n; } // cpp:50-57

// In-place transform along one axis of `n_lines` lines: line L = (outer, in) with in = L % inner,
// element t at data[(outer * len + t) * inner + in].  z axis: inner = 1; y: inner = nz; x: inner = ny nz
// (a batch of arrays only adds outer indices).  tw[t] = exp(+2 pi i t / len); log2len < 0: not a power
// of two (direct DFT).
template <int SIGN>
__global__ void __launch_bounds__(256)
k_gen_fft_axis(cplx *__restrict__ data, int len, long long inner, long long n_lines,
               const cplx *__restrict__ tw, int log2len)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *sm = reinterpret_cast<cplx *>(smem_raw);           // [GEN_TL][len + 1]
    cplx *sm2 = sm + GEN_TL * (len + 1);                      // second buffer (direct DFT only)
    const int pitch = len + 1;
    const long long L0 = (long long)blockIdx.x * GEN_TL;
    const int nl = (int)min((long long)GEN_TL, n_lines - L0);
    auto gidx = [&](int l, int t) {
        const long long L = L0 + l, outer = L / inner, in = L % inner;
        return (size_t)((outer * len + t) * inner + in);
    };
    for (int e = threadIdx.x; e < GEN_TL * len; e += blockDim.x) {
        int l, t;
        if (inner == 1) { l = e / len; t = e % len; }         // contiguous lines: t fastest
        else            { t = e / GEN_TL; l = e % GEN_TL; }   // strided lines: neighbouring lines fastest
        if (l < nl) {
            int dst = t;
            if (log2len >= 0) dst = (int)(__brev((unsigned)t) >> (32 - log2len)); // bit-reversed placement
            sm[l * pitch + dst] = data[gidx(l, t)];
        }
    }
    __syncthreads();
    const cplx *res = sm;
    if (log2len >= 0) {
        // decimation in time, log2(len) radix-2 stages in place
        for (int s = 0; s < log2len; ++s) {
            const int half = 1 << s, step = len >> (s + 1);
            for (int e = threadIdx.x; e < GEN_TL * (len / 2); e += blockDim.x) {
                const int l = e / (len / 2), b = e % (len / 2);
                const int k = b & (half - 1), t0 = ((b >> s) << (s + 1)) + k;
                cplx w = tw[k * step];
                if (SIGN < 0) w.y = -w.y;
                const cplx a = sm[l * pitch + t0], c = cmul(sm[l * pitch + t0 + half], w);
                sm[l * pitch + t0] = cadd(a, c);
                sm[l * pitch + t0 + half] = csub(a, c);
            }
            __syncthreads();
        }
    } else {
        for (int e = threadIdx.x; e < GEN_TL * len; e += blockDim.x) {
            const int l = e / len, k = e % len;
            cplx acc = make_double2(0.0, 0.0);
            int ph = 0; // (k * t) mod len
            for (int t = 0; t < len; ++t) {
                cplx w = tw[ph];
                if (SIGN < 0) w.y = -w.y;
                const cplx v = sm[l * pitch + t];
                acc.x += v.x * w.x - v.y * w.y;
                acc.y += v.x * w.y + v.y * w.x;
                ph += k;
                if (ph >= len) ph -= len;
            }
            sm2[l * pitch + k] = acc;
        }
        __syncthreads();
        res = sm2;
    }
    for (int e = threadIdx.x; e < GEN_TL * len; e += blockDim.x) {
        int l, t;
        if (inner == 1) { l = e / len; t = e % len; }
        else            { t = e / GEN_TL; l = e % GEN_TL; }
        if (l < nl) data[gidx(l, t)] = res[l * pitch + t];
    }
}

// c[idx] = (scale * f[idx], 0)                                                      (cpp:168-180)
__global__ void k_gen_r2c(const double *__restrict__ f, double scale, cplx *__restrict__ c, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        c[i] = make_double2(scale * f[i], 0.0);
}

// g[2c] = E fhat, g[2c+1] = conj(E) fhat for the pairs pair0 + c, c < n_chunk; E = ex[i] ey[j] ez[k]
// from the pair's separable table phase[pair][nx + ny + nz]                          (cpp:198-225)
__global__ void k_gen_phase(const cplx *__restrict__ fhat, const cplx *__restrict__ phase, int pair0, int n_chunk,
                            int nx, int ny, int nz, cplx *__restrict__ g)
{
    const size_t n = (size_t)nx * ny * nz;
    const int pt = nx + ny + nz;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * n_chunk; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / n);
        const size_t idx = e % n;
        const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((size_t)nz * ny));
        const cplx *P = phase + (size_t)(pair0 + c) * pt;
        const cplx E = cmul(cmul(P[i], P[nx + j]), P[nx + ny + k]);
        const cplx f = fhat[idx];
        g[(size_t)(2 * c) * n + idx] = cmul(f, E);
        g[(size_t)(2 * c + 1) * n + idx] = cmulc(f, E);
    }
}

// S[r(pair)][idx] += w_pair Re(g1 g2), pairs of the chunk in order (one owner thread per point)  (cpp:233-246)
__global__ void k_gen_prod_acc(const cplx *__restrict__ g, const int *__restrict__ pair_r,
                               const double *__restrict__ pair_w, int pair0, int n_chunk, size_t n,
                               double *__restrict__ S)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
        for (int c = 0; c < n_chunk; ++c) {
            const cplx a = g[(size_t)(2 * c) * n + idx], b = g[(size_t)(2 * c + 1) * n + idx];
            S[(size_t)pair_r[pair0 + c] * n + idx] += pair_w[pair0 + c] * (a.x * b.x - a.y * b.y);
        }
}

// c[a][idx] = (S[r0 + a][idx], 0) for a < n_arr
__global__ void k_gen_real_batch(const double *__restrict__ S, cplx *__restrict__ c, size_t n_total)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (size_t)gridDim.x * blockDim.x)
        c[i] = make_double2(S[i], 0.0);
}

// Qhat[idx] (+)= sum_a coef[r0 + a][|l|^2] P[a][idx]                                (cpp:252-273)
__global__ void k_gen_coef_acc(const cplx *__restrict__ P, const double *__restrict__ coef, int M, int r0, int n_arr,
                               int nx, int ny, int nz, int first, cplx *__restrict__ Qhat)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((size_t)nz * ny));
        const int li = gen_mode(i, nx), lj = gen_mode(j, ny), lk = gen_mode(k, nz);
        const int m = li * li + lj * lj + lk * lk;
        cplx q = first ? make_double2(0.0, 0.0) : Qhat[idx];
        for (int a = 0; a < n_arr; ++a) {
            const double cf = coef[(size_t)(r0 + a) * M + m];
            const cplx v = P[(size_t)a * n + idx];
            q.x += cf * v.x;
            q.y += cf * v.y;
        }
        Qhat[idx] = q;
    }
}

// H[0] = Qhat, H[1] = beta2[|l|^2] fhat                                             (cpp:281-299)
__global__ void k_gen_final_in(const cplx *__restrict__ Qhat, const cplx *__restrict__ fhat,
                               const double *__restrict__ beta2, int nx, int ny, int nz, cplx *__restrict__ H)
{
    const size_t n = (size_t)nx * ny * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((size_t)nz * ny));
        const int li = gen_mode(i, nx), lj = gen_mode(j, ny), lk = gen_mode(k, nz);
        const double b2 = beta2[li * li + lj * lj + lk * lk];
        const cplx f = fhat[idx];
        H[idx] = Qhat[idx];
        H[n + idx] = make_double2(b2 * f.x, b2 * f.y);
    }
}

// Q = Re(H0) - Re(H1) f   (with_loss)   |   Q = Re(H0)                              (cpp:314-330)
__global__ void k_gen_final_out(const cplx *__restrict__ H, const double *f, double *Q, size_t n, int with_loss)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double fv = f[i]; // read before the (possibly aliased) write
        Q[i] = with_loss ? H[i].x - H[n + i].x * fv : H[i].x;
    }
}

} // namespace bfsm
