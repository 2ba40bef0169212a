// fp64_ilp.cu -- FP64 pipe probe: DADD/DFMA throughput of one SM sub-partition as a function of the
// warps resident on it and of the independent dependency chains per warp (ILP).  Answers "how many warps
// does a register-heavy radix-16/32 butterfly kernel need before the FP64 pipe is saturated?".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_ilp tools/fp64_ilp.cu ; tools/fp64_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool FMA> __global__ void k(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = FMA ? fma(v[i], a, b) : v[i] + b;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, bool FMA> void run(double *out, int sms, int warps_per_smsp)
{
    const int iters = 4000 / ILP * 4;
    const int threads = 32 * 4 * warps_per_smsp; // one CTA per SM, warps spread over the 4 sub-partitions
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP, FMA><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP, FMA><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)sms * threads * iters * 8.0 * ILP; // thread instructions
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // warp instructions per cycle and sub-partition at the nominal clock
    const double wipc = inst / 32.0 / (sms * 4.0) / (ms * 1e-3 * clk * 1e3);
    printf("{\"op\": \"%s\", \"warps_per_smsp\": %d, \"ilp\": %d, \"ms\": %.4f, \"tinst_per_s\": %.3e, \"warp_inst_per_clk_smsp\": %.3f}\n",
           FMA ? "dfma" : "dadd", warps_per_smsp, ILP, ms, inst / (ms * 1e-3), wipc);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, 8);
    for (int w : {1, 2, 3, 4, 6, 8}) {
        run<1, false>(out, sms, w);
        run<2, false>(out, sms, w);
        run<4, false>(out, sms, w);
        run<8, false>(out, sms, w);
        run<16, false>(out, sms, w);
        run<4, true>(out, sms, w);
        run<16, true>(out, sms, w);
    }
    return 0;
}
