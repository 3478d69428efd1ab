// mg_kernels.cu -- geometric multigrid V-cycle used as the preconditioner of the on-device CG solve (the solve that
// replaces np.linalg.solve on the BC-reduced system, model.py:1028-1033, 1291, 1335).
//
// The Jacobi-preconditioned CG needs O(N) iterations on an N x N mesh (cond(K) ~ h^-2: 338 642 iterations for the 140
// solves of a 256 x 256 inclusion run); with a V(2,2) cycle the count is mesh-independent.  Everything is specific to
// the structured Q4 mesh of this path:
//   * every level is a (NX/2^l) x (NY/2^l) node grid with the same 9-node x 2-dof stencil, so the level matrices use
//     the closed-form CSR layout of csr_row_start (no rowptr / col arrays are read or stored);
//   * prolongation P is bilinear interpolation per displacement component (weights 1, 1/2, 1/4), restriction is P^T,
//     the coarse operators are Galerkin products A_c = P^T A P (they inherit whatever tangent field the elements
//     currently have -- plastic zones, inclusions -- without re-discretisation);
//   * Dirichlet dofs are handled by masks: masked rows and columns of a level act as removed; a coarse dof is masked
//     when the coincident fine dof is (boundary edges coincide on all levels);
//   * the smoother is damped Jacobi (symmetric, so the V-cycle with equal pre- and post-smoothing is an SPD
//     preconditioner), the coarsest level does a fixed number of the same sweeps.
//   * the level matrices are stored in FP32 (the finest one is a converted copy of the model's FP64 K, refreshed at
//     every solve): the cycle is a preconditioner, its matrix traffic is what it costs, and CG's own residuals and
//     products stay FP64; vectors and accumulations are FP64.
// All kernels are one thread per row / dof, no atomics: the V-cycle is deterministic.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "csr_stencil.cuh"
#include "mg.h"
#include "plf_comm.h"

namespace plf {

// ---------------------------------------------------------------------------------------------------------------
// row helpers
// ---------------------------------------------------------------------------------------------------------------
struct RowGeom {
    int j, k, c;
    int64_t pos;   // first CSR entry of the row
};

__device__ __forceinline__ RowGeom row_geom(int NX, int NY, int64_t row)
{
    RowGeom g;
    const int NnY = NY + 1;
    const int64_t n = row >> 1;
    g.c = (int)(row & 1);
    g.j = (int)(n / NnY);
    g.k = (int)(n - (int64_t)g.j * NnY);
    g.pos = csr_row_start(NX, NY, row);
    return g;
}

// y_row = sum_j A[row][j] x[j] over the stencil of the row (x holds zeros at masked dofs)
__device__ __forceinline__ double stencil_row_dot(int NX, int NY, const RowGeom &g, const float *__restrict__ val,
                                                  const double *__restrict__ x)
{
    const int NnY = NY + 1;
    int64_t pos = g.pos;
    double s = 0.;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
        if (g.j + dj < 0 || g.j + dj > NX) continue;
#pragma unroll
        for (int dk = -1; dk <= 1; dk++) {
            if (g.k + dk < 0 || g.k + dk > NY) continue;
            const int64_t m = (int64_t)(g.j + dj) * NnY + (g.k + dk);
            const float2 a = __ldg(reinterpret_cast<const float2 *>(val + pos));
            const double2 xv = __ldg(reinterpret_cast<const double2 *>(x + 2 * m));
            s = fma((double)a.x, xv.x, s);
            s = fma((double)a.y, xv.y, s);
            pos += 2;
        }
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// setup kernels
// ---------------------------------------------------------------------------------------------------------------
// joff (all transfer kernels): the coarse grid starts at fine node column joff -- 0 everywhere except between levels 0 and
// 1 of a strip with a left halo column (distributed solve): that column is a ghost, masked, and not part of the hierarchy
__global__ void __launch_bounds__(256) mg_mask_inject_kernel(int NXc, int NYc, int NYf, const unsigned char *__restrict__ mask_f,
                                                             unsigned char *__restrict__ mask_c, int joff)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nc = 2ll * (NXc + 1) * (NYc + 1);
    if (row >= nc) return;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int J = (int)(n / (NYc + 1)), K = (int)(n - (int64_t)J * (NYc + 1));
    const int64_t nf = (int64_t)(2 * J + joff) * (NYf + 1) + 2 * K;
    mask_c[row] = mask_f[2 * nf + c];
}

// 1 / diag per row (0 on masked rows)
// FP32 copy of the model's K for the finest level
__global__ void __launch_bounds__(256) mg_to_float_kernel(int64_t n, const double *__restrict__ a, float *__restrict__ b)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = (float)a[i];
}

__global__ void __launch_bounds__(256) mg_dinv_kernel(int NX, int NY, const float *__restrict__ val,
                                                      const unsigned char *__restrict__ mask, double *__restrict__ dinv)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nd = 2ll * (NX + 1) * (NY + 1);
    if (row >= nd) return;
    if (mask[row]) { dinv[row] = 0.; return; }
    const RowGeom g = row_geom(NX, NY, row);
    // position of the diagonal: neighbours before (0,0) in the (dj, dk) order
    int before = 0;
    if (g.j > 0) before += 1 + (g.k > 0) + (g.k < NY);
    if (g.k > 0) before += 1;
    const double d = (double)val[g.pos + 2 * before + g.c];
    dinv[row] = d != 0. ? 1. / d : 0.;
}

// Galerkin coarse operator A_c = P^T A_f P for one coarse row per thread.  (a, b): offset of the fine row's node from
// the coincident fine node (2J, 2K); (dj, dk): stencil offset of the fine column; the fine column at offset
// o = a + dj from 2J belongs to coarse column J + o/2 (o even) or to J + (o-1)/2 and J + (o+1)/2 with weight 1/2 each.
__global__ void __launch_bounds__(128) mg_rap_kernel(int NXc, int NYc, const float *__restrict__ val_f,
                                                     const unsigned char *__restrict__ mask_f, float *__restrict__ val_c, int joff)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nc = 2ll * (NXc + 1) * (NYc + 1);
    if (row >= nc) return;
    const int NXf = 2 * NXc + joff, NYf = 2 * NYc, NnYf = NYf + 1;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int J = (int)(n / (NYc + 1)), K = (int)(n - (int64_t)J * (NYc + 1));
    double acc[3][3][2];
#pragma unroll
    for (int x = 0; x < 3; x++)
#pragma unroll
        for (int y = 0; y < 3; y++) acc[x][y][0] = acc[x][y][1] = 0.;
#pragma unroll
    for (int a = -1; a <= 1; a++) {
        const int fj = 2 * J + a + joff;
        if (fj < joff || fj > NXf) continue;
#pragma unroll
        for (int b = -1; b <= 1; b++) {
            const int fk = 2 * K + b;
            if (fk < 0 || fk > NYf) continue;
            const int64_t frow = 2 * ((int64_t)fj * NnYf + fk) + c;
            if (mask_f[frow]) continue;
            const double wi = (a == 0 ? 1. : 0.5) * (b == 0 ? 1. : 0.5);
            int64_t pos = csr_row_start(NXf, NYf, frow);
#pragma unroll
            for (int dj = -1; dj <= 1; dj++) {
                if (fj + dj < 0 || fj + dj > NXf) continue;
#pragma unroll
                for (int dk = -1; dk <= 1; dk++) {
                    if (fk + dk < 0 || fk + dk > NYf) continue;
                    const int64_t m = (int64_t)(fj + dj) * NnYf + (fk + dk);
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(val_f + pos));
                    pos += 2;
                    const double vx = mask_f[2 * m] ? 0. : wi * (double)v.x;
                    const double vy = mask_f[2 * m + 1] ? 0. : wi * (double)v.y;
                    const int ox = a + dj, oy = b + dk;   // compile-time after unrolling: -2 .. 2
                    // x targets
                    const int x0 = (ox >= 0 ? ox : ox - 1) / 2;          // floor(ox / 2)
                    const int nxt = (ox & 1) ? 2 : 1;
                    const double wx = (ox & 1) ? 0.5 : 1.;
                    const int y0 = (oy >= 0 ? oy : oy - 1) / 2;
                    const int nyt = (oy & 1) ? 2 : 1;
                    const double wy = (oy & 1) ? 0.5 : 1.;
#pragma unroll
                    for (int tx = 0; tx < nxt; tx++)
#pragma unroll
                        for (int ty = 0; ty < nyt; ty++) {
                            const double w = wx * wy;
                            acc[x0 + tx + 1][y0 + ty + 1][0] = fma(w, vx, acc[x0 + tx + 1][y0 + ty + 1][0]);
                            acc[x0 + tx + 1][y0 + ty + 1][1] = fma(w, vy, acc[x0 + tx + 1][y0 + ty + 1][1]);
                        }
                }
            }
        }
    }
    int64_t pos = csr_row_start(NXc, NYc, row);
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
        if (J + dj < 0 || J + dj > NXc) continue;
#pragma unroll
        for (int dk = -1; dk <= 1; dk++) {
            if (K + dk < 0 || K + dk > NYc) continue;
            *reinterpret_cast<float2 *>(val_c + pos) = make_float2((float)acc[dj + 1][dk + 1][0], (float)acc[dj + 1][dk + 1][1]);
            pos += 2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cycle kernels
// ---------------------------------------------------------------------------------------------------------------
// damped Jacobi sweep: xout = xin + omega D^-1 (b - A xin)   (first = 1: xin = 0, no matrix product)
__global__ void __launch_bounds__(256) mg_jacobi_kernel(int NX, int NY, const float *__restrict__ val,
                                                        const double *__restrict__ dinv, const double *__restrict__ b,
                                                        const double *__restrict__ xin, double *__restrict__ xout, double omega,
                                                        int first, int64_t row0, int64_t row1)
{
    // [row0, row1): the rows this launch updates (all of them, or the rows a rank owns when level 0 is shared among ranks)
    const int64_t row = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row1) return;
    const double di = dinv[row];   // 0 on masked rows: they stay 0
    if (first) { xout[row] = omega * di * b[row]; return; }
    const RowGeom g = row_geom(NX, NY, row);
    const double s = stencil_row_dot(NX, NY, g, val, xin);
    xout[row] = di == 0. ? 0. : fma(omega * di, b[row] - s, xin[row]);
}

// r = b - A x (0 on masked rows)
__global__ void __launch_bounds__(256) mg_residual_kernel(int NX, int NY, const float *__restrict__ val,
                                                          const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                                          const double *__restrict__ x, double *__restrict__ r, int64_t row0,
                                                          int64_t row1)
{
    const int64_t row = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row1) return;
    if (mask[row]) { r[row] = 0.; return; }
    const RowGeom g = row_geom(NX, NY, row);
    r[row] = b[row] - stencil_row_dot(NX, NY, g, val, x);
}

// b_c = P^T r_f
__global__ void __launch_bounds__(256) mg_restrict_kernel(int NXc, int NYc, const double *__restrict__ r_f,
                                                          const unsigned char *__restrict__ mask_c, double *__restrict__ b_c, int joff,
                                                          int64_t row0, int64_t row1)
{
    const int64_t row = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row1) return;
    if (mask_c[row]) { b_c[row] = 0.; return; }
    const int NXf = 2 * NXc + joff, NYf = 2 * NYc, NnYf = NYf + 1;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int J = (int)(n / (NYc + 1)), K = (int)(n - (int64_t)J * (NYc + 1));
    double s = 0.;
#pragma unroll
    for (int a = -1; a <= 1; a++) {
        const int fj = 2 * J + a + joff;
        if (fj < joff || fj > NXf) continue;
#pragma unroll
        for (int b = -1; b <= 1; b++) {
            const int fk = 2 * K + b;
            if (fk < 0 || fk > NYf) continue;
            const double w = (a == 0 ? 1. : 0.5) * (b == 0 ? 1. : 0.5);
            s = fma(w, r_f[2 * ((int64_t)fj * NnYf + fk) + c], s);
        }
    }
    b_c[row] = s;
}

// x_f += P x_c (masked fine dofs stay 0)
__global__ void __launch_bounds__(256) mg_prolong_add_kernel(int NXf, int NYf, const double *__restrict__ x_c,
                                                             const unsigned char *__restrict__ mask_f, double *__restrict__ x_f, int joff,
                                                             int64_t row0, int64_t row1)
{
    const int64_t row = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row1) return;
    if (mask_f[row]) return;
    const int NYc = NYf / 2, NnYc = NYc + 1;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int p = (int)(n / (NYf + 1)) - joff, q = (int)(n - (int64_t)(p + joff) * (NYf + 1));
    if (p < 0) return;   // ghost column left of the hierarchy (masked anyway)
    const int J0 = p >> 1, K0 = q >> 1;
    const int nj = (p & 1) ? 2 : 1, nk = (q & 1) ? 2 : 1;
    const double w = ((p & 1) ? 0.5 : 1.) * ((q & 1) ? 0.5 : 1.);
    double s = 0.;
    for (int tj = 0; tj < nj; tj++)
        for (int tk = 0; tk < nk; tk++) s += x_c[2 * ((int64_t)(J0 + tj) * NnYc + (K0 + tk)) + c];
    x_f[row] += w * s;
}

// ---------------------------------------------------------------------------------------------------------------
// fused tail of the V-cycle: every level from `l0` down to the coarsest one and back up in ONE thread block, all level
// vectors in shared memory (ncu launch list of a 256 x 256 solve: 86 % of the GPU time sat in ~5 us launches of the
// kernels above on levels with a few thousand dofs or less -- 45 of the 68 launches of a V-cycle).  Same arithmetic in
// the same order as the per-level kernels, so the fused cycle returns bit-identical vectors.  Level matrices are read
// through the read-only path (a 32 x 32 level is 150 KB of FP32: L1 / L2 resident); a masked row is one with dinv == 0.
// ---------------------------------------------------------------------------------------------------------------
struct TailLevel {
    int NX, NY, ndof, off;        // off: start of the level's slice in each shared-memory vector
    const float *val;
    const double *dinv;
};
struct TailArgs {
    TailLevel lv[MG_MAX_LEVELS];
    int nl;                       // fused levels (lv[0] is the finest of them)
    int nu, coarse_sweeps, total; // total: sum of ndof over the fused levels
    double omega;
    const double *b_in;           // right-hand side of lv[0] (written by the restriction from the level above)
    double *x_out;                // solution of lv[0] (read by the prolongation to the level above)
};

__device__ __forceinline__ double tail_row_dot(const TailLevel &L, int row, const double *__restrict__ x)
{
    const int NnY = L.NY + 1, n = row >> 1;
    const int j = n / NnY, k = n - j * NnY;
    int64_t pos = csr_row_start(L.NX, L.NY, row);
    double s = 0.;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
        if (j + dj < 0 || j + dj > L.NX) continue;
#pragma unroll
        for (int dk = -1; dk <= 1; dk++) {
            if (k + dk < 0 || k + dk > L.NY) continue;
            const int m = (j + dj) * NnY + (k + dk);
            const float2 a = __ldg(reinterpret_cast<const float2 *>(L.val + pos));
            s = fma((double)a.x, x[2 * m], s);
            s = fma((double)a.y, x[2 * m + 1], s);
            pos += 2;
        }
    }
    return s;
}

// damped Jacobi sweeps on one level; returns the buffer holding the result (the per-level path arranges for it to be x;
// here the caller just follows the pointer)
__device__ __forceinline__ double *tail_smooth(const TailLevel &L, int sweeps, bool from_zero, double om, const double *dinv,
                                               const double *b, double *x, double *t)
{
    double *cur = x, *oth = t;
    int s0 = 0;
    if (from_zero) {
        for (int r = threadIdx.x; r < L.ndof; r += blockDim.x) cur[r] = om * dinv[r] * b[r];
        __syncthreads();
        s0 = 1;
    }
    for (int s = s0; s < sweeps; s++) {
        for (int r = threadIdx.x; r < L.ndof; r += blockDim.x) {
            const double di = dinv[r];
            const double a = tail_row_dot(L, r, cur);
            oth[r] = di == 0. ? 0. : fma(om * di, b[r] - a, cur[r]);
        }
        __syncthreads();
        double *h = cur; cur = oth; oth = h;
    }
    return cur;
}

__device__ __forceinline__ void tail_cycle(const TailArgs &A, double *sm)
{
    double *s_x = sm, *s_t = sm + A.total, *s_b = sm + 2 * A.total, *s_r = sm + 3 * A.total, *s_d = sm + 4 * A.total;
    __shared__ double *s_res[MG_MAX_LEVELS];   // where each level's current iterate lives (x or t slice)
    const double om = A.omega;
    for (int l = 0; l < A.nl; l++) {
        const TailLevel &L = A.lv[l];
        for (int r = threadIdx.x; r < L.ndof; r += blockDim.x) s_d[L.off + r] = L.dinv[r];
    }
    for (int r = threadIdx.x; r < A.lv[0].ndof; r += blockDim.x) s_b[r] = A.b_in[r];
    __syncthreads();
    // down
    for (int l = 0; l < A.nl - 1; l++) {
        const TailLevel &L = A.lv[l], &C = A.lv[l + 1];
        double *x = tail_smooth(L, A.nu, true, om, s_d + L.off, s_b + L.off, s_x + L.off, s_t + L.off);
        if (threadIdx.x == 0) s_res[l] = x;
        for (int r = threadIdx.x; r < L.ndof; r += blockDim.x)
            s_r[L.off + r] = s_d[L.off + r] == 0. ? 0. : s_b[L.off + r] - tail_row_dot(L, r, x);
        __syncthreads();
        // b_c = P^T r_f
        const int NnYf = L.NY + 1, NnYc = C.NY + 1;
        for (int r = threadIdx.x; r < C.ndof; r += blockDim.x) {
            double acc = 0.;
            if (s_d[C.off + r] != 0.) {
                const int n = r >> 1, c = r & 1, J = n / NnYc, K = n - J * NnYc;
#pragma unroll
                for (int a = -1; a <= 1; a++) {
                    const int fj = 2 * J + a;
                    if (fj < 0 || fj > L.NX) continue;
#pragma unroll
                    for (int bb = -1; bb <= 1; bb++) {
                        const int fk = 2 * K + bb;
                        if (fk < 0 || fk > L.NY) continue;
                        const double w = (a == 0 ? 1. : 0.5) * (bb == 0 ? 1. : 0.5);
                        acc = fma(w, s_r[L.off + 2 * (fj * NnYf + fk) + c], acc);
                    }
                }
            }
            s_b[C.off + r] = acc;
        }
        __syncthreads();
    }
    {
        const TailLevel &L = A.lv[A.nl - 1];
        double *x = tail_smooth(L, A.coarse_sweeps, true, om, s_d + L.off, s_b + L.off, s_x + L.off, s_t + L.off);
        if (threadIdx.x == 0) s_res[A.nl - 1] = x;
        __syncthreads();
    }
    // up
    for (int l = A.nl - 2; l >= 0; l--) {
        const TailLevel &L = A.lv[l], &C = A.lv[l + 1];
        double *xf = s_res[l];
        const double *xc = s_res[l + 1];
        const int NnYf = L.NY + 1, NnYc = C.NY + 1;
        for (int r = threadIdx.x; r < L.ndof; r += blockDim.x) {
            if (s_d[L.off + r] == 0.) continue;
            const int n = r >> 1, c = r & 1, p = n / NnYf, q = n - p * NnYf;
            const int J0 = p >> 1, K0 = q >> 1, nj = (p & 1) ? 2 : 1, nk = (q & 1) ? 2 : 1;
            const double w = ((p & 1) ? 0.5 : 1.) * ((q & 1) ? 0.5 : 1.);
            double acc = 0.;
            for (int tj = 0; tj < nj; tj++)
                for (int tk = 0; tk < nk; tk++) acc += xc[2 * ((J0 + tj) * NnYc + (K0 + tk)) + c];
            xf[r] += w * acc;
        }
        __syncthreads();
        double *oth = (xf == s_x + L.off) ? s_t + L.off : s_x + L.off;
        double *x = tail_smooth(L, A.nu, false, om, s_d + L.off, s_b + L.off, xf, oth);
        __syncthreads();
        if (threadIdx.x == 0) s_res[l] = x;
        __syncthreads();
    }
    const double *x0 = s_res[0];
    for (int r = threadIdx.x; r < A.lv[0].ndof; r += blockDim.x) A.x_out[r] = x0[r];
}

__global__ void __launch_bounds__(1024, 1) mg_tail_kernel(const TailArgs A)
{
    extern __shared__ double sm[];
    tail_cycle(A, sm);
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative middle of the V-cycle: the levels between the bandwidth-bound fine ones and the single-block tail
// (at 1024 x 1024: 256^2, 128^2, 64^2) are a few microseconds of work per step each; as separate kernels a cycle spends
// its time between launches.  One cooperative grid walks these levels down, lets block 0 run the tail in shared
// memory, and walks back up, with a grid-wide barrier after every step.  Level vectors live in global memory (L2) and
// are read with ld.global.cg: another block wrote them earlier in this very kernel, L1 may hold a stale line.
// Arithmetic and order are those of the per-level kernels (bit-identical results).
// ---------------------------------------------------------------------------------------------------------------
struct MidLevel {
    int NX, NY, ndof, pad;
    const float *val;
    const double *dinv;
    double *x, *t, *b, *r;
};
struct CoopArgs {
    MidLevel lv[MG_MAX_LEVELS];
    int nmid;                 // levels run by the whole grid; lv[nmid] (x, b only) is the first tail level when has_tail
    int has_tail;
    TailArgs tail;
};

__device__ __forceinline__ double mid_row_dot(const MidLevel &L, int row, const double *x)
{
    const int NnY = L.NY + 1, n = row >> 1;
    const int j = n / NnY, k = n - j * NnY;
    int64_t pos = csr_row_start(L.NX, L.NY, row);
    double s = 0.;
#pragma unroll
    for (int dj = -1; dj <= 1; dj++) {
        if (j + dj < 0 || j + dj > L.NX) continue;
#pragma unroll
        for (int dk = -1; dk <= 1; dk++) {
            if (k + dk < 0 || k + dk > L.NY) continue;
            const int m = (j + dj) * NnY + (k + dk);
            const float2 a = __ldg(reinterpret_cast<const float2 *>(L.val + pos));
            const double2 xv = __ldcg(reinterpret_cast<const double2 *>(x + 2 * m));
            s = fma((double)a.x, xv.x, s);
            s = fma((double)a.y, xv.y, s);
            pos += 2;
        }
    }
    return s;
}

__device__ __forceinline__ void coop_barrier(unsigned int *bar, unsigned int &epoch)
{
    // monotonic arrival counter (zeroed by the host before the launch); all blocks are co-resident (cooperative launch)
    __syncthreads();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(bar) < epoch) { }
        __threadfence();
    }
    __syncthreads();
}

// result buffer of `sweeps` Jacobi sweeps (same ping-pong as the host-driven smoother: from zero the first sweep writes
// x when `sweeps` is odd, else t; otherwise the sweeps start from x)
__device__ __forceinline__ double *mid_smooth(const MidLevel &L, int sweeps, bool from_zero, double om, double *x_start,
                                              unsigned int *bar, unsigned int &epoch)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    double *cur, *oth;
    int s0 = 0;
    if (from_zero) {
        cur = (sweeps & 1) ? L.x : L.t;
        oth = (cur == L.x) ? L.t : L.x;
        for (int r = gtid; r < L.ndof; r += gsz) __stcg(cur + r, om * __ldg(L.dinv + r) * __ldcg(L.b + r));
        coop_barrier(bar, epoch);
        s0 = 1;
    } else {
        cur = x_start;
        oth = (cur == L.x) ? L.t : L.x;
    }
    for (int s = s0; s < sweeps; s++) {
        for (int r = gtid; r < L.ndof; r += gsz) {
            const double di = __ldg(L.dinv + r);
            const double a = mid_row_dot(L, r, cur);
            __stcg(oth + r, di == 0. ? 0. : fma(om * di, __ldcg(L.b + r) - a, __ldcg(cur + r)));
        }
        coop_barrier(bar, epoch);
        double *h = cur; cur = oth; oth = h;
    }
    return cur;
}

__global__ void __launch_bounds__(1024, 1) mg_coop_kernel(const CoopArgs A, unsigned int *bar)
{
    extern __shared__ double sm[];
    unsigned int epoch = 0;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const double om = A.tail.omega;
    const int nu = A.tail.nu;
    // down
    for (int l = 0; l < A.nmid; l++) {
        const MidLevel &L = A.lv[l];
        const MidLevel &C = A.lv[l + 1];
        double *x = mid_smooth(L, nu, true, om, nullptr, bar, epoch);   // ends in L.x (nu is even)
        for (int r = gtid; r < L.ndof; r += gsz)
            __stcg(L.r + r, __ldg(L.dinv + r) == 0. ? 0. : __ldcg(L.b + r) - mid_row_dot(L, r, x));
        coop_barrier(bar, epoch);
        const int NnYf = L.NY + 1, NnYc = C.NY + 1;
        for (int r = gtid; r < C.ndof; r += gsz) {
            double acc = 0.;
            if (__ldg(C.dinv + r) != 0.) {
                const int n = r >> 1, c = r & 1, J = n / NnYc, K = n - J * NnYc;
#pragma unroll
                for (int a = -1; a <= 1; a++) {
                    const int fj = 2 * J + a;
                    if (fj < 0 || fj > L.NX) continue;
#pragma unroll
                    for (int bb = -1; bb <= 1; bb++) {
                        const int fk = 2 * K + bb;
                        if (fk < 0 || fk > L.NY) continue;
                        const double w = (a == 0 ? 1. : 0.5) * (bb == 0 ? 1. : 0.5);
                        acc = fma(w, __ldcg(L.r + 2 * (fj * NnYf + fk) + c), acc);
                    }
                }
            }
            __stcg(C.b + r, acc);
        }
        coop_barrier(bar, epoch);
    }
    if (A.has_tail) {
        if (blockIdx.x == 0) tail_cycle(A.tail, sm);
    } else {
        // no tail: lv[nmid] is the coarsest level and is smoothed by the grid
        const MidLevel &L = A.lv[A.nmid];
        double *x = mid_smooth(L, A.tail.coarse_sweeps, true, om, nullptr, bar, epoch);
        if (x != L.x) {
            for (int r = gtid; r < L.ndof; r += gsz) __stcg(L.x + r, __ldcg(x + r));
        }
    }
    coop_barrier(bar, epoch);
    // up
    for (int l = A.nmid - 1; l >= 0; l--) {
        const MidLevel &L = A.lv[l];
        const MidLevel &C = A.lv[l + 1];
        const int NnYf = L.NY + 1, NnYc = C.NY + 1;
        for (int r = gtid; r < L.ndof; r += gsz) {
            if (__ldg(L.dinv + r) == 0.) continue;
            const int n = r >> 1, c = r & 1, p = n / NnYf, q = n - p * NnYf;
            const int J0 = p >> 1, K0 = q >> 1, nj = (p & 1) ? 2 : 1, nk = (q & 1) ? 2 : 1;
            const double w = ((p & 1) ? 0.5 : 1.) * ((q & 1) ? 0.5 : 1.);
            double acc = 0.;
            for (int tj = 0; tj < nj; tj++)
                for (int tk = 0; tk < nk; tk++) acc += __ldcg(C.x + 2 * ((J0 + tj) * NnYc + (K0 + tk)) + c);
            __stcg(L.x + r, __ldcg(L.x + r) + w * acc);
        }
        coop_barrier(bar, epoch);
        mid_smooth(L, nu, false, om, L.x, bar, epoch);   // even number of sweeps: back in L.x
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

void mg_free(MgHierarchy &H)
{
    for (size_t l = 0; l < H.lv.size(); l++) {
        MgLevel &L = H.lv[l];
        cudaFree(L.val);
        if (l > 0) { cudaFree(L.mask); cudaFree(L.x); cudaFree(L.b); }
        cudaFree(L.dinv); cudaFree(L.r); cudaFree(L.t);
    }
    H.lv.clear();
    cudaFree(H.d_bar);
    H.d_bar = nullptr;
}

cudaError_t mg_build(MgHierarchy &H, int NX, int NY, int j0)
{
    mg_free(H);
    H.j0 = j0;
    int nx = NX, ny = NY;
    for (;;) {
        MgLevel L;
        L.NX = nx; L.NY = ny;
        L.ndof = 2ll * (nx + 1) * (ny + 1);
        L.nnz = 4ll * (3 * nx + 1) * (3 * ny + 1);
        H.lv.push_back(L);
        // level 1 covers the node columns [j0, NX] of level 0 (j0 = 1: the left halo column of a strip stays outside)
        const int nxe = H.lv.size() == 1 ? nx - j0 : nx;
        if ((nxe & 1) || (ny & 1) || nxe < 4 || ny < 4 || (int)H.lv.size() >= MG_MAX_LEVELS) break;
        nx = nxe >> 1; ny >>= 1;
    }
    for (size_t l = 0; l < H.lv.size(); l++) {
        MgLevel &L = H.lv[l];
        cudaError_t e;
        const size_t nb = (size_t)L.ndof * sizeof(double);
        if ((e = cudaMalloc((void **)&L.val, (size_t)L.nnz * sizeof(float))) != cudaSuccess) return e;
        if (l > 0) {
            if ((e = cudaMalloc((void **)&L.mask, (size_t)L.ndof)) != cudaSuccess) return e;
            if ((e = cudaMalloc((void **)&L.x, nb)) != cudaSuccess) return e;
            if ((e = cudaMalloc((void **)&L.b, nb)) != cudaSuccess) return e;
        }
        if ((e = cudaMalloc((void **)&L.dinv, nb)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&L.r, nb)) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&L.t, nb)) != cudaSuccess) return e;
    }
    // fused tail: the first level whose vectors (5 per level, all coarser levels included) fit one block's shared memory
    H.tail_first = -1;
    for (size_t l = 1; l < H.lv.size(); l++) {   // level 0 takes CG's own vectors: never part of the tail
        int64_t total = 0;
        for (size_t m = l; m < H.lv.size(); m++) total += H.lv[m].ndof;
        if (total * 5 * (int64_t)sizeof(double) <= 200 * 1024) { H.tail_first = (int)l; break; }
    }
    if (H.tail_first >= 0) {
        cudaError_t e = cudaFuncSetAttribute(mg_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(mg_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
    }
    // cooperative middle: the first level (>= 1) with at most 160 k dofs (256 x 256: ~2 rows per thread of a 64-block grid)
    H.coop_first = -1;
    for (size_t l = 1; l < H.lv.size(); l++)
        if (H.lv[l].ndof <= 160 * 1024) { H.coop_first = (int)l; break; }
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        H.coop_blocks = std::max(1, std::min(64, sms));
        cudaError_t e = cudaMalloc((void **)&H.d_bar, 64);
        if (e != cudaSuccess) return e;
        cudaMemset(H.d_bar, 0, 64);
    }
    return cudaSuccess;
}

// masks, Galerkin operators and inverse diagonals for the current fine matrix (called once per linear solve)
cudaError_t mg_setup(const LaunchCtx &lc, MgHierarchy &H, const double *val_fine, const unsigned char *mask_fine, int64_t *launches)
{
    if (H.lv.empty()) return cudaErrorInvalidValue;
    mg_to_float_kernel<<<std::min<unsigned>(grid_for(H.lv[0].nnz, 256), 148u * 16u), 256, 0, lc.stream>>>(H.lv[0].nnz, val_fine, H.lv[0].val);
    (*launches)++;
    H.lv[0].mask = const_cast<unsigned char *>(mask_fine);
    for (size_t l = 0; l < H.lv.size(); l++) {
        MgLevel &L = H.lv[l];
        if (l > 0) {
            const MgLevel &F = H.lv[l - 1];
            const int joff = l == 1 ? H.j0 : 0;
            mg_mask_inject_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, F.NY, F.mask, L.mask, joff);
            mg_rap_kernel<<<grid_for(L.ndof, 128), 128, 0, lc.stream>>>(L.NX, L.NY, F.val, F.mask, L.val, joff);
            *launches += 2;
        }
        mg_dinv_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.mask, L.dinv);
        (*launches)++;
    }
    return cudaGetLastError();
}

// levels [0, n) of the V-cycle that run as one kernel launch per step (the rest: cooperative kernel / single-block tail)
static int per_level_kernels(const MgHierarchy &H)
{
    const int nl = (int)H.lv.size();
    const int top = (H.fuse_tail && H.tail_first > 0) ? H.tail_first : nl;
    const int mid = (H.fuse_tail >= 2 && H.coop_first > 0 && H.coop_first < top) ? H.coop_first : -1;
    return mid > 0 ? mid : std::min(top, nl - 1);
}

void mg_plan_sharing(const MgHierarchy &H, int world, int rank, MgDist &md)
{
    static const long long min_dofs = getenv("PLF_SHARE_DOFS") ? atoll(getenv("PLF_SHARE_DOFS")) : 1000000ll;
    const int nper = per_level_kernels(H), NX = H.lv[0].NX;
    auto col_a = [&](int r) { return (int)((int64_t)NX * r / world); };
    auto col_b = [&](int r) { return (int)((int64_t)NX * (r + 1) / world) + (r == world - 1 ? 1 : 0); };
    int S = 1;
    while (S < nper && S < 8 && H.lv[S].ndof >= min_dofs) {
        bool wide = true;   // every rank keeps >= 2 node columns on level S
        for (int r = 0; r < world; r++) {
            int a = col_a(r), b = col_b(r);
            for (int l = 0; l < S; l++) { a = (a + 1) / 2; b = (b + 1) / 2; }
            wide = wide && b - a >= 2;
        }
        if (!wide) break;
        S++;
    }
    md.nshared = S;
    md.a[0] = col_a(rank); md.b[0] = col_b(rank);
    for (int l = 1; l < S; l++) { md.a[l] = (md.a[l - 1] + 1) / 2; md.b[l] = (md.b[l - 1] + 1) / 2; }
    const MgLevel &G = H.lv[S];   // first replicated level: its right-hand side is gathered
    for (int r = 0; r < world; r++) {
        int a = col_a(r), b = col_b(r);
        for (int l = 0; l < S; l++) { a = (a + 1) / 2; b = (b + 1) / 2; }
        md.goff[r] = 2ll * a * (G.NY + 1);
        md.cnt[r] = 2ll * b * (G.NY + 1) - md.goff[r];
    }
}

// z = V(r): one V(nu, nu) cycle from a zero initial guess.  r and z are fine-level vectors (z must not alias r).
// dist != nullptr: the finest dist->nshared levels are SHARED among the ranks of a replicated solve -- every rank holds the
// global vectors and the global hierarchy, but on a shared level it smooths / forms the residual / restricts / prolongs only
// the rows of the node columns it owns and swaps one ghost column with each neighbour after every step whose output the
// next stencil reads; the right-hand side of the first level below is gathered from all ranks and the rest runs on every
// rank alike.  Every row sees the arithmetic of the single-GPU cycle in the same order: the cycle is bit-identical.
cudaError_t mg_vcycle(const LaunchCtx &lc, MgHierarchy &H, const double *r_fine, double *z_fine, int64_t *launches, const MgDist *dist)
{
    const int nl = (int)H.lv.size();
    const double om = H.omega;
    H.lv[0].b = const_cast<double *>(r_fine);
    H.lv[0].x = z_fine;
    bool comm_failed = false;
    const int S = dist ? dist->nshared : 0;   // shared levels
    auto rows = [&](const MgLevel &L, int l, int64_t &r0, int64_t &r1) {   // rows a launch on level l works on
        r0 = 0; r1 = L.ndof;
        if (l < S) { r0 = 2ll * dist->a[l] * (L.NY + 1); r1 = 2ll * dist->b[l] * (L.NY + 1); }
    };
    auto swap_ghosts = [&](const MgLevel &L, int l, double *v) {
        if (l < S && comm_halo_exchange_cols(*dist->comm, v, L.NY, dist->a[l], dist->b[l], L.NX, lc.stream)) comm_failed = true;
    };
    auto jacobi = [&](MgLevel &L, int l, const double *in, double *out, int first) {
        int64_t r0, r1;
        rows(L, l, r0, r1);
        mg_jacobi_kernel<<<grid_for(r1 - r0, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.dinv, L.b, in, out, om, first, r0, r1);
        (*launches)++;
    };
    auto smooth = [&](MgLevel &L, int l, int sweeps, bool from_zero, bool last_needs_ghosts) {
        // the result ends in L.x; sweeps alternate between L.x and L.t (from zero the first sweep has no input and writes the
        // buffer that makes the last one land in L.x; otherwise the input is L.x and the number of sweeps is even)
        double *cur = L.x, *oth = L.t;
        int s0 = 0;
        if (from_zero) {
            cur = (sweeps & 1) ? L.x : L.t;
            oth = (cur == L.x) ? L.t : L.x;
            jacobi(L, l, nullptr, cur, 1);
            if (sweeps > 1 || last_needs_ghosts) swap_ghosts(L, l, cur);
            s0 = 1;
        }
        for (int s = s0; s < sweeps; s++) {
            jacobi(L, l, cur, oth, 0);
            std::swap(cur, oth);
            if (s + 1 < sweeps || last_needs_ghosts) swap_ghosts(L, l, cur);
        }
    };
    auto down = [&](int l) {
        MgLevel &L = H.lv[l], &C = H.lv[l + 1];
        smooth(L, l, H.nu, true, true);
        int64_t r0, r1;
        rows(L, l, r0, r1);
        mg_residual_kernel<<<grid_for(r1 - r0, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.mask, L.b, L.x, L.r, r0, r1);
        swap_ghosts(L, l, L.r);
        int64_t c0 = 0, c1 = C.ndof;
        if (l < S) { c0 = 2ll * ((dist->a[l] + 1) / 2) * (C.NY + 1); c1 = 2ll * ((dist->b[l] + 1) / 2) * (C.NY + 1); }
        mg_restrict_kernel<<<grid_for(c1 - c0, 256), 256, 0, lc.stream>>>(C.NX, C.NY, L.r, C.mask, C.b, l == 0 ? H.j0 : 0, c0, c1);
        *launches += 2;
        if (l == S - 1 && comm_gather_ranges(*dist->comm, C.b + dist->goff[dist->comm->rank], C.b, dist->goff, dist->cnt, lc.stream))
            comm_failed = true;
    };
    auto up = [&](int l) {
        MgLevel &L = H.lv[l], &C = H.lv[l + 1];
        int64_t r0, r1;
        rows(L, l, r0, r1);
        if (l < S) {   // the ghost columns too (the coarse solution is complete, or has its own ghosts): no exchange needed here
            r0 = 2ll * std::max(dist->a[l] - 1, 0) * (L.NY + 1);
            r1 = 2ll * std::min(dist->b[l] + 1, L.NX + 1) * (L.NY + 1);
        }
        mg_prolong_add_kernel<<<grid_for(r1 - r0, 256), 256, 0, lc.stream>>>(L.NX, L.NY, C.x, L.mask, L.x, l == 0 ? H.j0 : 0, r0, r1);
        (*launches)++;
        smooth(L, l, H.nu, false, l >= 1 && l < S);   // a shared finer level prolongs from this level's ghost columns as well
    };
    // levels [0, mid) run as per-level kernels, [mid, top) in the cooperative grid kernel, [top, nl) in the fused
    // single-block tail (inside the cooperative kernel when there is one)
    int top = (H.fuse_tail && H.tail_first > 0) ? H.tail_first : nl;
    const int mid = (H.fuse_tail >= 2 && H.coop_first > 0 && H.coop_first < top) ? H.coop_first : -1;
    if (mid > 0) {
        for (int l = 0; l < mid; l++) down(l);
        CoopArgs A;
        memset(&A, 0, sizeof A);
        A.nmid = std::min(top, nl - 1) - mid;          // the coarsest level is never a "mid" level: it has no level below
        A.has_tail = top < nl ? 1 : 0;
        for (int l = mid; l <= mid + A.nmid && l < nl; l++) {
            const MgLevel &L = H.lv[l];
            A.lv[l - mid] = MidLevel{L.NX, L.NY, (int)L.ndof, 0, L.val, L.dinv, L.x, L.t, L.b, L.r};
        }
        int off = 0;
        A.tail.nl = 0;
        if (A.has_tail) {
            A.tail.nl = nl - top;
            for (int l = top; l < nl; l++) {
                const MgLevel &L = H.lv[l];
                A.tail.lv[l - top] = TailLevel{L.NX, L.NY, (int)L.ndof, off, L.val, L.dinv};
                off += (int)L.ndof;
            }
            A.tail.b_in = H.lv[top].b; A.tail.x_out = H.lv[top].x;
        }
        A.tail.total = off; A.tail.nu = H.nu; A.tail.coarse_sweeps = H.coarse_sweeps; A.tail.omega = om;
        cudaMemsetAsync(H.d_bar, 0, sizeof(unsigned int), lc.stream);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeCooperative;
        attr.val.cooperative = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)H.coop_blocks);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = 5 * (size_t)off * sizeof(double);
        cfg.stream = lc.stream;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, mg_coop_kernel, A, H.d_bar);
        if (e != cudaSuccess) return e;
        (*launches)++;
        for (int l = mid - 1; l >= 0; l--) up(l);
        return comm_failed ? cudaErrorUnknown : cudaGetLastError();
    }
    for (int l = 0; l < std::min(top, nl - 1); l++) down(l);
    if (top < nl) {
        TailArgs A;
        int off = 0;
        A.nl = nl - top;
        for (int l = top; l < nl; l++) {
            const MgLevel &L = H.lv[l];
            A.lv[l - top] = TailLevel{L.NX, L.NY, (int)L.ndof, off, L.val, L.dinv};
            off += (int)L.ndof;
        }
        A.total = off; A.nu = H.nu; A.coarse_sweeps = H.coarse_sweeps; A.omega = om;
        A.b_in = H.lv[top].b; A.x_out = H.lv[top].x;
        mg_tail_kernel<<<1, 1024, 5 * (size_t)off * sizeof(double), lc.stream>>>(A);
        (*launches)++;
    } else {
        smooth(H.lv[nl - 1], nl - 1, H.coarse_sweeps, true, false);
    }
    for (int l = std::min(top, nl - 1) - 1; l >= 0; l--) up(l);
    return comm_failed ? cudaErrorUnknown : cudaGetLastError();
}

}  // namespace plf
