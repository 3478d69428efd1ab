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
#include <vector>

#include "csr_stencil.cuh"
#include "mg.h"

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
__global__ void __launch_bounds__(256) mg_mask_inject_kernel(int NXc, int NYc, int NYf, const unsigned char *__restrict__ mask_f,
                                                             unsigned char *__restrict__ mask_c)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nc = 2ll * (NXc + 1) * (NYc + 1);
    if (row >= nc) return;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int J = (int)(n / (NYc + 1)), K = (int)(n - (int64_t)J * (NYc + 1));
    const int64_t nf = (int64_t)(2 * J) * (NYf + 1) + 2 * K;
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
                                                     const unsigned char *__restrict__ mask_f, float *__restrict__ val_c)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nc = 2ll * (NXc + 1) * (NYc + 1);
    if (row >= nc) return;
    const int NXf = 2 * NXc, NYf = 2 * NYc, NnYf = NYf + 1;
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
        const int fj = 2 * J + a;
        if (fj < 0 || fj > NXf) continue;
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
                                                        int first)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nd = 2ll * (NX + 1) * (NY + 1);
    if (row >= nd) return;
    const double di = dinv[row];   // 0 on masked rows: they stay 0
    if (first) { xout[row] = omega * di * b[row]; return; }
    const RowGeom g = row_geom(NX, NY, row);
    const double s = stencil_row_dot(NX, NY, g, val, xin);
    xout[row] = di == 0. ? 0. : fma(omega * di, b[row] - s, xin[row]);
}

// r = b - A x (0 on masked rows)
__global__ void __launch_bounds__(256) mg_residual_kernel(int NX, int NY, const float *__restrict__ val,
                                                          const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                                          const double *__restrict__ x, double *__restrict__ r)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nd = 2ll * (NX + 1) * (NY + 1);
    if (row >= nd) return;
    if (mask[row]) { r[row] = 0.; return; }
    const RowGeom g = row_geom(NX, NY, row);
    r[row] = b[row] - stencil_row_dot(NX, NY, g, val, x);
}

// b_c = P^T r_f
__global__ void __launch_bounds__(256) mg_restrict_kernel(int NXc, int NYc, const double *__restrict__ r_f,
                                                          const unsigned char *__restrict__ mask_c, double *__restrict__ b_c)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nc = 2ll * (NXc + 1) * (NYc + 1);
    if (row >= nc) return;
    if (mask_c[row]) { b_c[row] = 0.; return; }
    const int NXf = 2 * NXc, NYf = 2 * NYc, NnYf = NYf + 1;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int J = (int)(n / (NYc + 1)), K = (int)(n - (int64_t)J * (NYc + 1));
    double s = 0.;
#pragma unroll
    for (int a = -1; a <= 1; a++) {
        const int fj = 2 * J + a;
        if (fj < 0 || fj > NXf) continue;
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
                                                             const unsigned char *__restrict__ mask_f, double *__restrict__ x_f)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nf = 2ll * (NXf + 1) * (NYf + 1);
    if (row >= nf) return;
    if (mask_f[row]) return;
    const int NYc = NYf / 2, NnYc = NYc + 1;
    const int64_t n = row >> 1;
    const int c = (int)(row & 1);
    const int p = (int)(n / (NYf + 1)), q = (int)(n - (int64_t)p * (NYf + 1));
    const int J0 = p >> 1, K0 = q >> 1;
    const int nj = (p & 1) ? 2 : 1, nk = (q & 1) ? 2 : 1;
    const double w = ((p & 1) ? 0.5 : 1.) * ((q & 1) ? 0.5 : 1.);
    double s = 0.;
    for (int tj = 0; tj < nj; tj++)
        for (int tk = 0; tk < nk; tk++) s += x_c[2 * ((int64_t)(J0 + tj) * NnYc + (K0 + tk)) + c];
    x_f[row] += w * s;
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
}

cudaError_t mg_build(MgHierarchy &H, int NX, int NY)
{
    mg_free(H);
    int nx = NX, ny = NY;
    for (;;) {
        MgLevel L;
        L.NX = nx; L.NY = ny;
        L.ndof = 2ll * (nx + 1) * (ny + 1);
        L.nnz = 4ll * (3 * nx + 1) * (3 * ny + 1);
        H.lv.push_back(L);
        if ((nx & 1) || (ny & 1) || nx < 4 || ny < 4 || (int)H.lv.size() >= 16) break;
        nx >>= 1; ny >>= 1;
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
            mg_mask_inject_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, F.NY, F.mask, L.mask);
            mg_rap_kernel<<<grid_for(L.ndof, 128), 128, 0, lc.stream>>>(L.NX, L.NY, F.val, F.mask, L.val);
            *launches += 2;
        }
        mg_dinv_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.mask, L.dinv);
        (*launches)++;
    }
    return cudaGetLastError();
}

// z = V(r): one V(nu, nu) cycle from a zero initial guess.  r and z are fine-level vectors (z must not alias r).
cudaError_t mg_vcycle(const LaunchCtx &lc, MgHierarchy &H, const double *r_fine, double *z_fine, int64_t *launches)
{
    const int nl = (int)H.lv.size();
    const double om = H.omega;
    H.lv[0].b = const_cast<double *>(r_fine);
    H.lv[0].x = z_fine;
    auto smooth = [&](MgLevel &L, int sweeps, bool from_zero) {
        // result must end in L.x; sweeps alternate between L.x and L.t
        double *cur = L.x, *oth = L.t;
        int first = from_zero ? 1 : 0;
        // choose the starting buffer so that the last sweep writes L.x
        // (from zero: the first sweep has no input; otherwise the input is L.x)
        if (from_zero) {
            double *out = (sweeps & 1) ? L.x : L.t;
            mg_jacobi_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.dinv, L.b, nullptr, out, om, 1);
            (*launches)++;
            cur = out; oth = (out == L.x) ? L.t : L.x;
            for (int s = 1; s < sweeps; s++) {
                mg_jacobi_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.dinv, L.b, cur, oth, om, 0);
                (*launches)++;
                std::swap(cur, oth);
            }
        } else {
            // even number of sweeps keeps the result in L.x
            for (int s = 0; s < sweeps; s++) {
                mg_jacobi_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.dinv, L.b, cur, oth, om, 0);
                (*launches)++;
                std::swap(cur, oth);
            }
        }
        (void)first;
    };
    for (int l = 0; l < nl - 1; l++) {
        MgLevel &L = H.lv[l], &C = H.lv[l + 1];
        smooth(L, H.nu, true);
        mg_residual_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, L.val, L.mask, L.b, L.x, L.r);
        mg_restrict_kernel<<<grid_for(C.ndof, 256), 256, 0, lc.stream>>>(C.NX, C.NY, L.r, C.mask, C.b);
        *launches += 2;
    }
    smooth(H.lv[nl - 1], H.coarse_sweeps, true);
    for (int l = nl - 2; l >= 0; l--) {
        MgLevel &L = H.lv[l], &C = H.lv[l + 1];
        mg_prolong_add_kernel<<<grid_for(L.ndof, 256), 256, 0, lc.stream>>>(L.NX, L.NY, C.x, L.mask, L.x);
        (*launches)++;
        smooth(L, H.nu, false);
    }
    return cudaGetLastError();
}

}  // namespace plf
