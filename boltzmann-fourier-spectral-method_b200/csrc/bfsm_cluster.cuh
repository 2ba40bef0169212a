// bfsm_cluster.cuh -- the gain path of a 32^3 grid with NO global scratch: one thread-block CLUSTER of
// eight CTAs owns a whole (r, sigma) pair (32^3 complex doubles = 512 KiB, spread over the shared
// memory of eight SMs) and takes it from the phase multiply to the accumulated product
//
//   fhat(i,j,k) --(phase, (y,z) inverse FFT)--> Z(i; y, z) --DSMEM transpose--> Z(.; y, z) lines
//               --(x inverse FFT)--> Z_H(x,y,z) --> acc(x,y,z) += w (Re^2 - Im^2)       (cpp:198-246)
//
//   plane stage   CTA c of the cluster owns the x-spectral planes i = 4c .. 4c+3.  Two groups of 128
//                 threads run the three register stages of k_plane_gain3 (radix-8 along z | 4 x 4 block |
//                 radix-8 along y) on two planes each; their entries of fhat stay in REGISTERS for the
//                 whole kernel.  The last stage does not store to global memory: thread (y = s3k + 4 k2,
//                 z) sends its eight results to the eight CTAs of the cluster -- CTA k2 receives rows
//                 y = 4 k2 .. 4 k2 + 3 of every plane (st.shared::cluster, 512-byte runs per warp).
//   hand-over     one split cluster barrier per pair (barrier.cluster.arrive.release / wait.acquire); the
//                 receive buffer is double buffered, so the plane stage of pair n+1 runs while ...
//   x stage       ... 256 more threads of every CTA transform the 128 lines (4 rows x 32 z) it received:
//                 the register-resident x stage of bfsm_pencil_reg.cuh reading shared memory instead of
//                 global memory, accumulators in registers over all pairs of a work unit, stored once
//                 into the unit's partial slot of S_r.
//
// The hybrid grids never exist in global memory; per pair a CTA moves 56 KiB through DSMEM instead of
// 1 MiB through L2/HBM per pair and grid.  The three Nyquist planes of the packed mode are left to
// k_plane_gain3 (restricted to them) + k_nyq_accum: they are 3/35 of the plane work and would unbalance
// the eight CTAs.
#pragma once
#include "bfsm_kernels.cuh"
#include "bfsm_pencil_reg.cuh"

namespace bfsm {

constexpr int CL_SIZE = 8;          // CTAs per cluster (portable maximum)
// Role sizes.  Measured at 32^3 / 16 x 94 (752 pairs, 18 clusters; profiles/r02_ab32_cluster*.log):
//   2 plane groups (two planes each) + 256 x threads, 512 threads at 128 registers   0.386 ms  <- kept
//   4 plane groups + 128 x threads (two passes), setmaxnreg 88 / 128                 0.405 ms
//   4 plane groups + 256 x threads, setmaxnreg 72 / 96 (spills)                      0.462 ms
// against 0.216 ms for k_plane_gain3 + k_pencil_gain_reg through global memory.  With the remote stores
// and the cluster barriers compiled out the kernel still takes 0.28 ms: both roles are latency bound
// (one instruction per ~15 cycles per warp, 16-20 warps per SM), the DSMEM transfer (64 KiB per pair
// and CTA, ~16 B/clk) and the barrier add ~0.09 ms each.  See DESIGN.md.
constexpr int CL_PLANE_GROUPS = 2;  // plane groups per CTA (two planes each per pair)
constexpr int CL_X_THREADS = 256;   // x-stage threads per CTA (one pass over the 128 lines of a pair)
constexpr int CL_THREADS = CL_PLANE_GROUPS * 128 + CL_X_THREADS;

template <int N> constexpr size_t cluster_smem_bytes()
{
    // plane buffers | phase slots | twiddles | receive buffers
    return sizeof(cplx) * ((size_t)CL_PLANE_GROUPS * N * (N + 1) + (size_t)CL_PLANE_GROUPS * 2 * 3 * N + N +
                           (size_t)2 * N * (N / CL_SIZE) * N);
}

__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(unsigned addr, cplx v)
{
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void cluster_arrive()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
// arrive without the release fence: for threads that wrote nothing another CTA will read
__device__ __forceinline__ void cluster_arrive_relaxed()
{
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = CL_SIZE x n_clusters (cluster q walks the units q, q + n_clusters, ...), block CL_THREADS.
template <int N, bool UNIFORM_W>
__global__ void __cluster_dims__(CL_SIZE, 1, 1) __launch_bounds__(CL_THREADS, 1)
k_gain_cluster(const cplx *__restrict__ fhat, const cplx *__restrict__ phase, const cplx *__restrict__ twtab,
               const PencilUnit *__restrict__ units, int n_units, const double *__restrict__ pair_w,
               double *__restrict__ S, int n_r_local)
{
    static_assert(N == 32, "the cluster kernel is written for 32^3 grids (one pair = 8 x 64 KiB)");
    constexpr int R = N / 4, TG = 4 * N, PITCH = N + 1, H = N / 2;
    constexpr int YL = N / CL_SIZE;                 // rows of every plane a CTA receives (4)
    constexpr int PL = YL / CL_PLANE_GROUPS;        // planes per plane group and pair (2)
    constexpr size_t N2 = (size_t)N * N, N3 = N2 * N;
    constexpr int RB = N * YL * N;                  // elements of one receive buffer: [i][yl][z]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *pbufs = reinterpret_cast<cplx *>(smem_raw);            // groups x (N x PITCH)
    cplx *phs = pbufs + CL_PLANE_GROUPS * N * PITCH;              // groups x 2 x 3N
    cplx *tws = phs + CL_PLANE_GROUPS * 2 * 3 * N;                // N twiddles
    cplx *rbuf = tws + N;                                         // 2 x RB

    const unsigned crank = cluster_ctarank();
    const int n_clusters = gridDim.x / CL_SIZE, cluster_id = blockIdx.x / CL_SIZE;
    if (threadIdx.x < N) tws[threadIdx.x] = __ldg(&twtab[threadIdx.x]);
    __syncthreads();

    if (threadIdx.x < CL_PLANE_GROUPS * TG) {
        // =========================== plane stage ==================================================
        const int g = threadIdx.x / TG, tg = threadIdx.x % TG;
        cplx *buf = pbufs + g * N * PITCH;
        cplx *myph = phs + g * 2 * 3 * N;
        const int j1 = tg % N, b1 = tg / N;                    // S1: row, residue
        const int s2b = tg % R, s2k = (tg / R) % R;            // S2 (threads < R*R)
        const bool s2_active = tg < R * R;
        const int s3slot = tg % N, s3k = tg / N;               // S3: column z, k1'
        cplx wz[3], wy[3];
#pragma unroll
        for (int m = 1; m < 4; ++m) {
            wz[m - 1] = tws[(m * s2k) & (N - 1)];
            wy[m - 1] = tws[(m * s2b) & (N - 1)];
        }
        // this thread's entries of the group's planes, for the whole kernel
        cplx fr[PL][R];
#pragma unroll
        for (int pl = 0; pl < PL; ++pl) {
            const int i = YL * crank + PL * g + pl;
#pragma unroll
            for (int a = 0; a < R; ++a) fr[pl][a] = __ldg(&fhat[(size_t)i * N2 + j1 * N + 4 * a + b1]);
        }
        // receive buffers of the eight CTAs (S3 sends result k2 to CTA k2), element offset of this thread
        unsigned rb[CL_SIZE];
        const unsigned rbuf_s = (unsigned)__cvta_generic_to_shared(rbuf);
#pragma unroll
        for (int k = 0; k < CL_SIZE; ++k) rb[k] = map_to_rank(rbuf_s, k);

        int slot = 0, n = 0; // phase-table slot, running pair count (receive buffer = n & 1)
        bool first = true;
        for (int u = cluster_id; u < n_units; u += n_clusters) {
            const PencilUnit un = units[u];
            for (int q = un.p0; q < un.p1; ++q, ++n, slot ^= 1) {
                if (first) { // the very first table; later ones are prefetched below
                    if (tg < 3 * N) myph[tg] = __ldg(&phase[(size_t)q * 3 * N + tg]);
                    group_sync(1 + g, TG);
                    first = false;
                }
                const cplx *P = myph + slot * 3 * N;
                // next pair of this cluster (same unit, or the first pair of its next unit)
                int qn = q + 1;
                if (qn == un.p1) qn = (u + n_clusters < n_units) ? units[u + n_clusters].p0 : -1;
                cplx nxt = make_double2(0.0, 0.0);
                if (qn >= 0 && tg < 3 * N) nxt = __ldg(&phase[(size_t)qn * 3 * N + tg]);
#pragma unroll
                for (int pl = 0; pl < PL; ++pl) {
                    const int i = YL * crank + PL * g + pl;
                    // ---------------- S1: phase-weighted fhat (real multiplier m_H), radix-R along z
                    {
                        const cplx exi = P[i], eyj = P[N + j1];
                        const cplx X = cmul(exi, eyj);
                        const cplx ext = (i == H) ? exi : make_double2(exi.x, -exi.y);
                        const cplx eyt = (j1 == H) ? eyj : make_double2(eyj.x, -eyj.y);
                        const cplx Xt = cmul(ext, eyt);
                        const double cA = 0.5 * (X.x + Xt.x), cB = 0.5 * (X.y - Xt.y);
                        const double nA = 0.5 * (X.x - Xt.y), nB = 0.5 * (X.y + Xt.x);
                        cplx v[R];
#pragma unroll
                        for (int a = 0; a < R; ++a) {
                            const cplx ez = P[2 * N + 4 * a + b1];
                            const double zp = ez.x + ez.y, zm = ez.x - ez.y;
                            const bool ny = (a == R / 2) && (b1 == 0); // k == H
                            const double m = (ny ? nA : cA) * zp + (ny ? nB : cB) * zm;
                            v[a] = make_double2(m * fr[pl][a].x, m * fr[pl][a].y);
                        }
                        Dft<R, +1>::run(v);
                        cplx *row = buf + j1 * PITCH + R * b1;
#pragma unroll
                        for (int k1 = 0; k1 < R; ++k1) row[k1] = v[dft_reg<R>(k1)];
                    }
                    group_sync(1 + g, TG);
                    // ---------------- S2: 4 x 4 block, z twiddle + radix-4 (z), radix-4 (y) + y twiddle
                    if (s2_active) {
                        cplx e[4][4];
                        cplx *blk = buf + s2b * PITCH + s2k;
#pragma unroll
                        for (int ap = 0; ap < 4; ++ap)
#pragma unroll
                            for (int b = 0; b < 4; ++b) e[ap][b] = blk[(R * ap) * PITCH + R * b];
#pragma unroll
                        for (int ap = 0; ap < 4; ++ap) {
#pragma unroll
                            for (int b = 1; b < 4; ++b) e[ap][b] = cmul(e[ap][b], wz[b - 1]);
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
                    group_sync(1 + g, TG);
                    // ---------------- S3: radix-R along y; result y = s3k + 4 k2 goes to CTA k2, row s3k
                    {
                        cplx v[R];
                        const cplx *col = buf + (R * s3k) * PITCH + s3slot;
#pragma unroll
                        for (int bp = 0; bp < R; ++bp) v[bp] = col[bp * PITCH];
                        Dft<R, +1>::run(v);
                        const unsigned off = (unsigned)(sizeof(cplx) * ((size_t)(n & 1) * RB + ((size_t)i * YL + s3k) * N + s3slot));
#pragma unroll
                        for (int k2 = 0; k2 < R; ++k2) st_cluster(rb[k2] + off, v[dft_reg<R>(k2)]);
                    }
                    group_sync(1 + g, TG);
                }
                if (qn >= 0 && tg < 3 * N) myph[(slot ^ 1) * 3 * N + tg] = nxt;
                group_sync(1 + g, TG);
                // pair n is complete in every receive buffer once all threads of the cluster got here
                cluster_arrive();
                cluster_wait();
            }
        }
    } else {
        // =========================== x stage =======================================================
        const int tx = threadIdx.x - CL_PLANE_GROUPS * TG;
        const int lane = tx & 31, wx = tx >> 5;
        const int c = lane >> 4, zz = lane & 15;
        const int z = (wx & 1) * 16 + zz;                   // warp = 16 lines of one row
        constexpr int PASSES = (YL * N) / (CL_X_THREADS / 2);
        cplx tw_unused[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) tw_unused[m] = make_double2(1.0, 0.0);
        int n = 0;
        for (int u = cluster_id; u < n_units; u += n_clusters) {
            const PencilUnit un = units[u];
            double acc[PASSES][16];
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps)
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[ps][k] = 0.0;
            for (int q = un.p0; q < un.p1; ++q, ++n) {
                // my own results of pair n-1 are consumed; wait for pair n of the whole cluster
                cluster_arrive_relaxed();
                cluster_wait();
                const double w = UNIFORM_W ? 1.0 : __ldg(&pair_w[q]);
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int yl = (CL_X_THREADS / 64) * ps + (wx >> 1);
                    const cplx *src = rbuf + (size_t)(n & 1) * RB + ((size_t)c * YL + yl) * N + z;
                    cplx v[16];
#pragma unroll
                    for (int a = 0; a < 16; ++a) v[a] = src[(size_t)(2 * a) * YL * N];
                    pencil_reg_compute<N, UNIFORM_W>(v, c, 0, tw_unused, w, acc[ps]);
                }
            }
            const double scale = UNIFORM_W ? __ldg(&pair_w[un.p0]) : 1.0;
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) {
                const int y = YL * crank + (CL_X_THREADS / 64) * ps + (wx >> 1);
                double *Sr = S + ((size_t)un.slot * n_r_local + un.r) * N3 + (size_t)y * N + z;
#pragma unroll
                for (int j = 0; j < 16; ++j) Sr[(size_t)pencil_out_x<N>(j, c, 0) * N2] = acc[ps][j] * scale;
            }
        }
    }
}

} // namespace bfsm
