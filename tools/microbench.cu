// microbench.cu -- pipe/bandwidth probes that decide kernel design questions on B200 (tuning aid).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
//   ./tools/microbench            (prints one JSON object per probe)
//
// Probes (every kernel is independent: no kernel waits on another):
//   fp64   : DFMA alone, DMMA (mma.sync m8n8k4 / m16n8k8 f64) alone, and both in the same CTA on
//            different warps -- do the FP64 tensor and SIMT pipes run concurrently on sm_100a?
//   lsu    : LDS.128+STS.128 alone, SHFL.32 alone, both together -- do shuffles have their own
//            throughput or do they share the shared-memory data path?
//   l2     : write a region then read it back, region sizes below and above the 126 MB L2, plus a
//            split grid (half the CTAs write region A while the others read region B)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// ------------------------------------------------------------------------------ fp64 pipes
// mode bit 0: warps with (warp & 1) == 0 (or all, if mode == 1) run DFMA chains
// mode bit 1: the other warps (or all, if mode == 2) run DMMA chains
template <int SHAPE>
__global__ void __launch_bounds__(256) k_fp64(double *out, int iters, int mode)
{
    const int warp = threadIdx.x >> 5;
    const bool do_fma = (mode == 1) || (mode == 3 && (warp & 1) == 0);
    const bool do_mma = (mode == 2) || (mode == 3 && (warp & 1) == 1);
    double s = 0.0;
    if (do_fma) {
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
        const double m = 1.0000001, c = 1e-7;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += a[k];
    }
    if (do_mma) {
        if (SHAPE == 0) {
            // m8n8k4: A 1 reg, B 1 reg, C 2 regs; 256 FMA per warp instruction
            double c0[8], c1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { c0[k] = 0.0; c1[k] = 0.0; }
            const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s += c0[k] + c1[k];
        } else {
            // m16n8k8: A 4 regs, B 2 regs, C 4 regs; 1024 FMA per warp instruction
            double c[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[k][j] = 0.0;
            const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                                     : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) s += c[k][j];
        }
    }
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------ LSU probes
// mode 1: STS.128 + LDS.128 pairs (conflict free); mode 2: SHFL.32 x 4; mode 3: even warps 1, odd warps 2
__global__ void __launch_bounds__(256) k_lsu(double *out, int iters, int mode)
{
    __shared__ double2 sm[256 * 4];
    const int warp = threadIdx.x >> 5;
    const bool do_smem = (mode == 1) || (mode == 3 && (warp & 1) == 0);
    const bool do_shfl = (mode == 2) || (mode == 3 && (warp & 1) == 1);
    double2 v = make_double2(threadIdx.x, 1.0);
    if (do_smem) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sm[threadIdx.x + 256 * u] = v;
                __syncwarp();
                const double2 w = sm[(threadIdx.x ^ 1) + 256 * u];
                v.x += w.y;
                v.y += w.x;
            }
        }
    }
    if (do_shfl) {
        int a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a = __shfl_xor_sync(0xffffffffu, a, 1) + 1;
                b = __shfl_xor_sync(0xffffffffu, b, 2) + 1;
                c = __shfl_xor_sync(0xffffffffu, c, 3) + 1;
                d = __shfl_xor_sync(0xffffffffu, d, 1) + 1;
            }
        }
        v.x += a + b + c + d;
    }
    if (v.x == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = v.x;
}

// ------------------------------------------------------------------------------ L2 probes
__global__ void __launch_bounds__(256) k_write(double2 *p, size_t n, int cta_lo, int cta_n)
{
    const int cta = blockIdx.x - cta_lo;
    if (cta < 0 || cta >= cta_n) return;
    const double2 v = make_double2(1.0, 2.0);
    for (size_t i = (size_t)cta * blockDim.x + threadIdx.x; i < n; i += (size_t)cta_n * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(256) k_read(const double2 *p, size_t n, double *out, int cta_lo, int cta_n)
{
    const int cta = blockIdx.x - cta_lo;
    if (cta < 0 || cta >= cta_n) return;
    double s = 0.0;
    for (size_t i = (size_t)cta * blockDim.x + threadIdx.x; i < n; i += (size_t)cta_n * blockDim.x) {
        double2 v;
        asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + i));
        s += v.x + v.y;
    }
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// half the grid writes region A, the other half reads region B (written by the previous launch)
__global__ void __launch_bounds__(256) k_split(double2 *wr, const double2 *rd, size_t n, double *out, int n_writers)
{
    if ((int)blockIdx.x < n_writers) {
        const double2 v = make_double2(1.0, 2.0);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)n_writers * blockDim.x)
            wr[i] = v;
    } else {
        const int cta = blockIdx.x - n_writers, n_readers = gridDim.x - n_writers;
        double s = 0.0;
        for (size_t i = (size_t)cta * blockDim.x + threadIdx.x; i < n; i += (size_t)n_readers * blockDim.x) {
            double2 v;
            asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(rd + i));
            s += v.x + v.y;
        }
        if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

// MLP-rich read: 8 independent 16-byte loads per thread and iteration (per-SM L2 read bandwidth probe)
__global__ void __launch_bounds__(1024) k_read8(const double2 *p, size_t n, double *out)
{
    double s = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        double2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "l"(p + i + k * stride));
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y;
    }
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int r = 0; r < reps; ++r) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
    return ms / reps;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * 4096 * 256));

    // ---- fp64 pipes: 4 CTAs x 256 threads per SM
    {
        const int blocks = sms * 4, iters = 2048;
        const double fma_per_thread = (double)iters * 64.0;
        for (int shape = 0; shape < 2; ++shape) {
            float t1 = time_ms([&] { if (shape == 0) k_fp64<0><<<blocks, 256>>>(out, iters, 1); else k_fp64<1><<<blocks, 256>>>(out, iters, 1); }, 3);
            float t2 = time_ms([&] { if (shape == 0) k_fp64<0><<<blocks, 256>>>(out, iters, 2); else k_fp64<1><<<blocks, 256>>>(out, iters, 2); }, 3);
            float t3 = time_ms([&] { if (shape == 0) k_fp64<0><<<blocks, 256>>>(out, iters, 3); else k_fp64<1><<<blocks, 256>>>(out, iters, 3); }, 3);
            const double n_fma = (double)blocks * 256 * fma_per_thread;
            const double mma_fma_per_warp = (double)iters * 16.0 * (shape == 0 ? 256.0 : 1024.0);
            const double n_mma = (double)blocks * 8 * mma_fma_per_warp;
            printf("{\"probe\": \"fp64\", \"mma_shape\": \"%s\", \"dfma_alone_tfma_s\": %.2f, \"dmma_alone_tfma_s\": %.2f, "
                   "\"mixed_ms\": %.4f, \"dfma_half_alone_ms\": %.4f, \"dmma_half_alone_ms\": %.4f, "
                   "\"note\": \"mixed: even warps DFMA, odd warps DMMA; if mixed_ms ~ max(halves) the pipes are independent, if ~ sum they share\"}\n",
                   shape == 0 ? "m8n8k4" : "m16n8k8", n_fma / (t1 * 1e-3) / 1e12, n_mma / (t2 * 1e-3) / 1e12, t3,
                   t1 / 2, t2 / 2);
        }
    }
    // ---- LSU
    {
        const int blocks = sms * 4, iters = 4096;
        float t1 = time_ms([&] { k_lsu<<<blocks, 256>>>(out, iters, 1); }, 3);
        float t2 = time_ms([&] { k_lsu<<<blocks, 256>>>(out, iters, 2); }, 3);
        float t3 = time_ms([&] { k_lsu<<<blocks, 256>>>(out, iters, 3); }, 3);
        const double clk = prop.clockRate * 1e3; // Hz (nominal)
        const double warp_instr = (double)blocks / sms * 8 * iters * 4; // per SM, per kind
        printf("{\"probe\": \"lsu\", \"smem_pair_ms\": %.4f, \"shfl_ms\": %.4f, \"mixed_ms\": %.4f, "
               "\"cyc_per_sts128_lds128_pair_per_sm\": %.2f, \"cyc_per_shfl32_per_sm\": %.2f, "
               "\"note\": \"cycles at the nominal %.0f MHz; mixed = even warps smem, odd warps shfl (half the work of each)\"}\n",
               t1, t2, t3, t1 * 1e-3 * clk / warp_instr, t2 * 1e-3 * clk / (warp_instr * 4), clk / 1e6);
    }
    // ---- L2
    {
        const size_t sizes_mib[] = {16, 32, 48, 64, 96, 128, 256, 1024};
        double2 *buf, *buf2;
        CK(cudaMalloc(&buf, (size_t)1024 << 20));
        CK(cudaMalloc(&buf2, (size_t)1024 << 20));
        const int blocks = sms * 8;
        for (size_t mib : sizes_mib) {
            const size_t n = (mib << 20) / sizeof(double2);
            float tw = 0, tr = 0;
            // alternate write / read of the same region so that reads find what the writes left in L2
            cudaEvent_t e[3];
            for (auto &x : e) CK(cudaEventCreate(&x));
            const int reps = 10;
            for (int r = 0; r < reps + 1; ++r) {
                CK(cudaEventRecord(e[0]));
                k_write<<<blocks, 256>>>(buf, n, 0, blocks);
                CK(cudaEventRecord(e[1]));
                k_read<<<blocks, 256>>>(buf, n, out, 0, blocks);
                CK(cudaEventRecord(e[2]));
                CK(cudaEventSynchronize(e[2]));
                float a, b;
                CK(cudaEventElapsedTime(&a, e[0], e[1]));
                CK(cudaEventElapsedTime(&b, e[1], e[2]));
                if (r > 0) { tw += a; tr += b; }
            }
            tw /= reps; tr /= reps;
            // split grid: 60 % writers, 40 % readers, regions swap every launch
            const int n_writers = blocks * 3 / 5;
            float ts = time_ms([&] {
                k_split<<<blocks, 256>>>(buf, buf2, n, out, n_writers);
                k_split<<<blocks, 256>>>(buf2, buf, n, out, n_writers);
            }, 5) / 2;
            printf("{\"probe\": \"l2\", \"region_MiB\": %zu, \"write_GBs\": %.0f, \"read_back_GBs\": %.0f, "
                   "\"split_write_plus_read_GBs\": %.0f}\n",
                   mib, (double)(mib << 20) / (tw * 1e-3) / 1e9, (double)(mib << 20) / (tr * 1e-3) / 1e9,
                   2.0 * (double)(mib << 20) / (ts * 1e-3) / 1e9);
        }
    }
    // ---- per-SM L2 read bandwidth: n_sm CTAs of 1024 threads (one per SM) stream an L2-resident 48 MiB
    {
        double2 *buf;
        const size_t bytes = (size_t)48 << 20, n = bytes / sizeof(double2);
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 0, bytes));
        const int counts[] = {8, 19, 38, 50, 74, 110, 148, 296};
        for (int c : counts) {
            // warm L2, then time
            k_read8<<<c, 1024>>>(buf, n, out);
            float t = time_ms([&] { k_read8<<<c, 1024>>>(buf, n, out); }, 10);
            printf("{\"probe\": \"l2_read_per_sm\", \"ctas_1024thr\": %d, \"total_GBs\": %.0f, \"per_cta_GBs\": %.1f}\n", c,
                   (double)bytes / (t * 1e-3) / 1e9, (double)bytes / (t * 1e-3) / 1e9 / c);
        }
    }
    return 0;
}
