// bfsm_aux.cuh -- small device helpers around the collision path: the vector update of an explicit
// time integrator (stages stay on the device) and per-cell velocity moments (what a transport solver
// reads off f, and the conservation check of Q: mass, momentum and energy of Q(f,f) vanish).
#pragma once
#include <cuda_runtime.h>

namespace bfsm {

// out = a x + b y  (out may alias x or y)
__global__ void __launch_bounds__(256) k_axpby(double *out, double a, const double *x, double b, const double *y,
                                               size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + b * y[i];
}

// One CTA per cell: m[cell] = dvx dvy dvz * sum_v g(v) * (1, vx, vy, vz, |v|^2 / 2) on the grid
// v_i = -L + dv/2 + i dv per axis (maxwell_bkw_fftw.cpp:62-71), summed in a fixed order (deterministic).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_moments(const double *__restrict__ g, int nx, int ny, int nz,
                                                     double L, double *__restrict__ m)
{
    __shared__ double red[5][THREADS];
    const size_t N3 = (size_t)nx * ny * nz;
    const double dx = 2.0 * L / nx, dy = 2.0 * L / ny, dz = 2.0 * L / nz;
    const double *gc = g + (size_t)blockIdx.x * N3;
    double s[5] = {0, 0, 0, 0, 0};
    for (size_t idx = threadIdx.x; idx < N3; idx += THREADS) {
        const int k = (int)(idx % nz), j = (int)((idx / nz) % ny), i = (int)(idx / ((size_t)nz * ny));
        const double vx = -L + 0.5 * dx + i * dx, vy = -L + 0.5 * dy + j * dy, vz = -L + 0.5 * dz + k * dz;
        const double v = gc[idx];
        s[0] += v;
        s[1] += v * vx;
        s[2] += v * vy;
        s[3] += v * vz;
        s[4] += v * 0.5 * (vx * vx + vy * vy + vz * vz);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) red[q][threadIdx.x] = s[q];
    __syncthreads();
    for (int h = THREADS / 2; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h)
#pragma unroll
            for (int q = 0; q < 5; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x < 5) m[(size_t)blockIdx.x * 5 + threadIdx.x] = red[threadIdx.x][0] * dx * dy * dz;
}

} // namespace bfsm
