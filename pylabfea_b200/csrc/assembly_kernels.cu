// assembly_kernels.cu -- K3: element stiffness integration B^T C B (Element.calc_Kel, model.py:365-370) fused with
// the assembly of the global stiffness (Model.setupK, model.py:954-977) into CSR, for the structured Q4 mesh.
//
// COMPILED WITH -fmad=false: only the explicit fma() calls fuse.  numpy evaluates (B^T @ C) @ B entry by entry as
// the chain s = fma(a_k, b_k, s) over ascending k from s = 0; with the structural zeros of B skipped (they leave
// the chain unchanged) that is exactly what the code below does, so Kel is bit-identical to the reference.
//
// One thread owns one CSR row (one dof).  It visits the <= 4 elements around its node in ascending element index
// -- the order in which setupK adds them -- computes the one row of each element's Kel it needs (rows of Kel are
// independent, so nothing is computed twice across threads), and accumulates into the <= 18 entries of its row
// held in registers.  Every CSR value is written exactly once and no Kel ever touches HBM:
//   HBM traffic = 288 B tangent read + 4 B descriptor per element, 8 B per non-zero written (= 288 B per element).
#include "plf_device.cuh"
#include "plf_host.h"

namespace plf {

__constant__ GeomClass c_geom[PLF_MAX_GEOM];

cudaError_t assembly_upload_geom(const GeomClass *h, int n, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_geom, h, sizeof(GeomClass) * n, 0, cudaMemcpyHostToDevice, st);
}

// CSR pattern of the structured mesh: row 2n+c has the dofs of all nodes within +-1 in (j,k), ascending.
__global__ void csr_pattern_kernel(int NX, int NY, int32_t *__restrict__ rowptr, int32_t *__restrict__ col)
{
    const int64_t Ndof = 2ll * (NX + 1) * (NY + 1);
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > Ndof) return;
    const int NnY = NY + 1;
    const int64_t colw = 4ll * (3 * NY + 1);   // entries per unit of "x neighbours" in one node column
    auto start = [&](int64_t row) -> int64_t {
        if (row >= Ndof) {
            int64_t sx = 3ll * (NX + 1) - 2;   // sum over all node columns of the x-neighbour count
            return colw * sx;
        }
        int64_t n = row >> 1;
        int c = (int)(row & 1);
        int j = (int)(n / NnY), k = (int)(n % NnY);
        int nxc = 1 + (j > 0) + (j < NX), nyc = 1 + (k > 0) + (k < NY);
        int64_t pj = (j == 0) ? 0 : 3ll * j - 1;
        int64_t pk = (k == 0) ? 0 : 3ll * k - 1;
        return colw * pj + 4ll * nxc * pk + (int64_t)c * 2 * nxc * nyc;
    };
    const int64_t s0 = start(r);
    rowptr[r] = (int32_t)s0;
    if (r == Ndof) return;
    const int64_t n = r >> 1;
    const int j = (int)(n / NnY), k = (int)(n % NnY);
    int64_t pos = s0;
    for (int dj = -1; dj <= 1; dj++) {
        if (j + dj < 0 || j + dj > NX) continue;
        for (int dk = -1; dk <= 1; dk++) {
            if (k + dk < 0 || k + dk > NY) continue;
            int64_t mnode = (int64_t)(j + dj) * NnY + (k + dk);
            col[pos++] = (int32_t)(2 * mnode);
            col[pos++] = (int32_t)(2 * mnode + 1);
        }
    }
}

// One row (dof `c` of local node `a`) of the element stiffness: kel[8].
template <bool PS>
__device__ __forceinline__ void kel_row(const GeomClass &gc, const double *__restrict__ C, int a, int c, double *kel)
{
    const int row1 = c;   // ux rows use C row 0, uy rows use C row 1
    const double c10 = C[6 * row1 + 0], c11 = C[6 * row1 + 1], c12 = C[6 * row1 + 2], c15 = C[6 * row1 + 5];
    const double c50 = C[30], c51 = C[31], c52 = C[32], c55 = C[35];
    double c20 = 0., c21 = 0., c22 = 0., c25 = 0.;
    if (PS) { c20 = C[12]; c21 = C[13]; c22 = C[14]; c25 = C[17]; }
    double ksum[8];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const double b1 = c ? gc.q[g][a] : gc.p[g][a];
        const double b5 = c ? gc.p[g][a] : gc.q[g][a];
        // T = row (2a+c) of B^T C, only the columns that meet a non-zero row of B
        double t0 = b1 * c10, t1 = b1 * c11, t2 = b1 * c12, t5 = b1 * c15;
        if (PS) {
            const double b2 = gc.r[g][2 * a + c];
            t0 = fma(b2, c20, t0); t1 = fma(b2, c21, t1); t2 = fma(b2, c22, t2); t5 = fma(b2, c25, t5);
        }
        t0 = fma(b5, c50, t0); t1 = fma(b5, c51, t1); t2 = fma(b5, c52, t2); t5 = fma(b5, c55, t5);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            double kx = t0 * gc.p[g][m];
            double ky = t1 * gc.q[g][m];
            if (PS) {
                kx = fma(t2, gc.r[g][2 * m], kx);
                ky = fma(t2, gc.r[g][2 * m + 1], ky);
            }
            kx = fma(t5, gc.q[g][m], kx);
            ky = fma(t5, gc.p[g][m], ky);
            if (g == 0) { ksum[2 * m] = kx; ksum[2 * m + 1] = ky; }
            else { ksum[2 * m] = ksum[2 * m] + kx; ksum[2 * m + 1] = ksum[2 * m + 1] + ky; }
        }
    }
#pragma unroll
    for (int x = 0; x < 8; x++) kel[x] = gc.jacw * ksum[x];
}

template <bool PS>
__global__ void __launch_bounds__(128) assemble_rows_kernel(int NX, int NY, const int32_t *__restrict__ elem_desc,
                                                            const double *__restrict__ elstiff,
                                                            const int32_t *__restrict__ rowptr, double *__restrict__ val)
{
    const int64_t Ndof = 2ll * (NX + 1) * (NY + 1);
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Ndof) return;
    const int NnY = NY + 1;
    const int64_t n = r >> 1;
    const int c = (int)(r & 1);
    const int j = (int)(n / NnY), k = (int)(n % NnY);

    double acc[9][2];
#pragma unroll
    for (int s = 0; s < 9; s++) acc[s][0] = acc[s][1] = 0.;

#pragma unroll
    for (int ddj = 0; ddj < 2; ddj++) {
#pragma unroll
        for (int ddk = 0; ddk < 2; ddk++) {
            const int ej = j - 1 + ddj, ek = k - 1 + ddk;   // ascending element index e = ej*NY + ek
            if (ej < 0 || ej >= NX || ek < 0 || ek >= NY) continue;
            const int64_t e = (int64_t)ej * NY + ek;
            const int cls = __ldg(elem_desc + e) >> 16;
            const int a = 2 * (1 - ddj) + (1 - ddk);   // local index of this node in element e
            double kel[8];
            kel_row<PS>(c_geom[cls], elstiff + e * 36, a, c, kel);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int sj = ddj + (m >> 1), sk = ddk + (m & 1);   // stencil slot of local node m, in 0..2
                acc[3 * sj + sk][0] = acc[3 * sj + sk][0] + kel[2 * m];
                acc[3 * sj + sk][1] = acc[3 * sj + sk][1] + kel[2 * m + 1];
            }
        }
    }
    int64_t pos = rowptr[r];
#pragma unroll
    for (int sj = 0; sj < 3; sj++) {
        if (j + sj - 1 < 0 || j + sj - 1 > NX) continue;
#pragma unroll
        for (int sk = 0; sk < 3; sk++) {
            if (k + sk - 1 < 0 || k + sk - 1 > NY) continue;
            // entries come in (ux, uy) pairs and row starts are even: one 128-bit store per pair
            *reinterpret_cast<double2 *>(val + pos) = make_double2(acc[3 * sj + sk][0], acc[3 * sj + sk][1]);
            pos += 2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-per-thread with staged stores (production).  Same loads and arithmetic as assemble_rows_kernel; a block owns
// 64 consecutive nodes of one node column, whose 128 CSR rows form one contiguous run of values.  The rows are
// collected in shared memory (128-bit, conflict-free) and written back with fully coalesced 128-bit stores, so L2
// receives whole 32 B sectors instead of two partial writes per sector.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SN = 64;   // nodes per block

template <bool PS>
__global__ void __launch_bounds__(2 * SN) assemble_rows_staged_kernel(int NX, int NY, const int32_t *__restrict__ elem_desc,
                                                                      const double *__restrict__ elstiff,
                                                                      const int32_t *__restrict__ rowptr, double *__restrict__ val)
{
    __shared__ double2 s_out[SN * 2 * 9];
    const int tid = threadIdx.x;
    const int NnY = NY + 1;
    const int j = blockIdx.y, k0 = blockIdx.x * SN;
    const int c = tid & 1, k = k0 + (tid >> 1);
    const bool active = k <= NY;
    const int kmax = min(k0 + SN - 1, NY);
    const int64_t r_first = 2 * ((int64_t)j * NnY + k0);
    const int64_t r_last = 2 * ((int64_t)j * NnY + kmax) + 1;
    const int32_t begin = __ldg(rowptr + r_first), end = __ldg(rowptr + r_last + 1);

    if (active) {
        double acc[9][2];
#pragma unroll
        for (int s = 0; s < 9; s++) acc[s][0] = acc[s][1] = 0.;
#pragma unroll
        for (int ddj = 0; ddj < 2; ddj++) {
#pragma unroll
            for (int ddk = 0; ddk < 2; ddk++) {
                const int ej = j - 1 + ddj, ek = k - 1 + ddk;   // ascending element index e = ej*NY + ek
                if (ej < 0 || ej >= NX || ek < 0 || ek >= NY) continue;
                const int64_t e = (int64_t)ej * NY + ek;
                const int cls = __ldg(elem_desc + e) >> 16;
                const int a = 2 * (1 - ddj) + (1 - ddk);
                double kel[8];
                kel_row<PS>(c_geom[cls], elstiff + e * 36, a, c, kel);
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int sj = ddj + (m >> 1), sk = ddk + (m & 1);
                    acc[3 * sj + sk][0] = acc[3 * sj + sk][0] + kel[2 * m];
                    acc[3 * sj + sk][1] = acc[3 * sj + sk][1] + kel[2 * m + 1];
                }
            }
        }
        const int64_t r = 2 * ((int64_t)j * NnY + k) + c;
        int pos2 = (__ldg(rowptr + r) - begin) >> 1;
#pragma unroll
        for (int sj = 0; sj < 3; sj++) {
            if (j + sj - 1 < 0 || j + sj - 1 > NX) continue;
#pragma unroll
            for (int sk = 0; sk < 3; sk++) {
                if (k + sk - 1 < 0 || k + sk - 1 > NY) continue;
                s_out[pos2++] = make_double2(acc[3 * sj + sk][0], acc[3 * sj + sk][1]);
            }
        }
    }
    __syncthreads();
    const int n2 = (end - begin) >> 1;
    double2 *dst = reinterpret_cast<double2 *>(val + begin);
    for (int i = tid; i < n2; i += 2 * SN) dst[i] = s_out[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Tiled variant (experimental, slower: kept for the record, see profiles/).  The row-per-thread kernel above reads the 288 B AoS tangents and writes its 144 B rows
// with a stride between lanes, so every warp-level access touches 16-36 cache lines and the L1 wavefront rate
// (one 128 B line per cycle per SM) becomes the limit (ncu: long_scoreboard + lg_throttle, 42 % of the HBM roofline).
// Here a block of 256 threads owns a tile of 4 node columns x 32 node rows:
//   1. the 5 x 33 element tangents around the tile are loaded with coalesced 128-bit loads (each element column of
//      the tile is one contiguous run in memory) into shared memory, record stride 37 doubles (conflict-free),
//   2. every thread computes its CSR row exactly as above, reading the tangents from shared memory,
//   3. the rows are staged in shared memory in CSR order (one contiguous run per node column) and written with
//      coalesced 128-bit stores.
// Arithmetic and summation order are unchanged -> CSR(K) stays bit-identical to the reference.
// ---------------------------------------------------------------------------------------------------------------
static int g_tiled = 2;   // 0: row per thread, direct stores; 1: shared-memory tile; 2: row per thread, staged stores
constexpr int TJ = 4, TK = 32;                 // node tile
constexpr int TEJ = TJ + 1, TEK = TK + 1;      // element tile (one halo column / row on the low side)
constexpr int ESTRIDE = 37;

template <bool PS>
__global__ void __launch_bounds__(TJ * TK * 2) assemble_tile_kernel(int NX, int NY, const int32_t *__restrict__ elem_desc,
                                                                    const double *__restrict__ elstiff,
                                                                    const int32_t *__restrict__ rowptr, double *__restrict__ val)
{
    extern __shared__ double smem[];                       // TEJ*TEK*ESTRIDE doubles, reused for the output staging
    __shared__ int s_cls[TEJ * TEK];
    const int tid = threadIdx.x;
    const int NnY = NY + 1;
    const int j0 = blockIdx.x * TJ, k0 = blockIdx.y * TK;

    // 1. coalesced load of the element tile
    for (int cj = 0; cj < TEJ; cj++) {
        const int ej = j0 - 1 + cj;
        if (ej < 0 || ej >= NX) continue;
        const int ek_lo = max(k0 - 1, 0), ek_hi = min(k0 + TK - 1, NY - 1);
        const int nel = ek_hi - ek_lo + 1;
        if (nel <= 0) continue;
        const int64_t e0 = (int64_t)ej * NY + ek_lo;
        const double2 *src = reinterpret_cast<const double2 *>(elstiff + e0 * 36);
        const int lbase = cj * TEK + (ek_lo - (k0 - 1));
        for (int i = tid; i < nel * 18; i += TJ * TK * 2) {
            const double2 v = __ldg(src + i);
            const int el = i / 18, w = i - el * 18;
            double *dst = smem + (lbase + el) * ESTRIDE + 2 * w;
            dst[0] = v.x;
            dst[1] = v.y;
        }
        for (int i = tid; i < nel; i += TJ * TK * 2) s_cls[lbase + i] = __ldg(elem_desc + e0 + i) >> 16;
    }
    __syncthreads();

    // 2. one CSR row per thread
    const int c = tid & 1, kk = (tid >> 1) & (TK - 1), jj = tid >> 6;
    const int j = j0 + jj, k = k0 + kk;
    const bool active = (j <= NX) && (k <= NY);
    double acc[9][2];
#pragma unroll
    for (int s = 0; s < 9; s++) acc[s][0] = acc[s][1] = 0.;
    if (active) {
#pragma unroll
        for (int ddj = 0; ddj < 2; ddj++) {
#pragma unroll
            for (int ddk = 0; ddk < 2; ddk++) {
                const int ej = j - 1 + ddj, ek = k - 1 + ddk;   // ascending element index e = ej*NY + ek
                if (ej < 0 || ej >= NX || ek < 0 || ek >= NY) continue;
                const int le = (jj + ddj) * TEK + (kk + ddk);
                const int a = 2 * (1 - ddj) + (1 - ddk);
                double kel[8];
                kel_row<PS>(c_geom[s_cls[le]], smem + le * ESTRIDE, a, c, kel);
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int sj = ddj + (m >> 1), sk = ddk + (m & 1);
                    acc[3 * sj + sk][0] = acc[3 * sj + sk][0] + kel[2 * m];
                    acc[3 * sj + sk][1] = acc[3 * sj + sk][1] + kel[2 * m + 1];
                }
            }
        }
    }
    __syncthreads();   // all reads of the tangent tile are done: reuse the memory for the rows

    // 3. stage the rows in CSR order (per node column: rows 2*(j*NnY + k0) ... are one contiguous run) and store
    const bool col_ok = (j <= NX);
    const int kmax = min(k0 + TK - 1, NY);                 // last node row of this tile
    int32_t col_begin = 0, col_end = 0;
    if (col_ok) {
        const int64_t r_first = 2 * ((int64_t)j * NnY + k0);
        const int64_t r_last = 2 * ((int64_t)j * NnY + kmax) + 1;
        col_begin = rowptr[r_first];
        col_end = rowptr[r_last + 1];
    }
    double *s_out = smem + jj * (TK * 2 * 18);             // <= 64 rows x 18 entries per node column
    if (active) {
        const int64_t r = 2 * ((int64_t)j * NnY + k) + c;
        int pos = rowptr[r] - col_begin;
#pragma unroll
        for (int sj = 0; sj < 3; sj++) {
            if (j + sj - 1 < 0 || j + sj - 1 > NX) continue;
#pragma unroll
            for (int sk = 0; sk < 3; sk++) {
                if (k + sk - 1 < 0 || k + sk - 1 > NY) continue;
                s_out[pos] = acc[3 * sj + sk][0];
                s_out[pos + 1] = acc[3 * sj + sk][1];
                pos += 2;
            }
        }
    }
    __syncthreads();
    if (col_ok) {
        const int n2 = (col_end - col_begin) >> 1;          // entries come in (ux, uy) pairs; col_begin is even
        double2 *dst = reinterpret_cast<double2 *>(val + col_begin);
        const double2 *src = reinterpret_cast<const double2 *>(s_out);
        for (int i = tid & 63; i < n2; i += 64) dst[i] = src[i];
    }
}

// Element.Kel for every element (plf_get_kel): thread per (element, row).
template <bool PS>
__global__ void __launch_bounds__(128) kel_kernel(int64_t Nel, const int32_t *__restrict__ elem_desc,
                                                  const double *__restrict__ elstiff, double *__restrict__ Kel)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Nel * 8) return;
    const int64_t e = t >> 3;
    const int i = (int)(t & 7);
    const int cls = elem_desc[e] >> 16;
    double kel[8];
    kel_row<PS>(c_geom[cls], elstiff + e * 36, i >> 1, i & 1, kel);
#pragma unroll
    for (int x = 0; x < 8; x++) Kel[e * 64 + 8 * i + x] = kel[x];
}

cudaError_t launch_csr_pattern(const LaunchCtx &lc, int NX, int NY, int32_t *rowptr, int32_t *col, int64_t *launches)
{
    const int64_t Ndof = 2ll * (NX + 1) * (NY + 1);
    csr_pattern_kernel<<<(unsigned)((Ndof + 1 + 255) / 256), 256, 0, lc.stream>>>(NX, NY, rowptr, col);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_assemble(const LaunchCtx &lc, int NX, int NY, int planestress, const int32_t *elem_desc,
                            const double *elstiff, const int32_t *rowptr, double *val, int64_t *launches)
{
    if (g_tiled == 2) {
        const dim3 grid((NY + 1 + SN - 1) / SN, NX + 1);
        if (planestress)
            assemble_rows_staged_kernel<true><<<grid, 2 * SN, 0, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
        else
            assemble_rows_staged_kernel<false><<<grid, 2 * SN, 0, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
    } else if (g_tiled == 1) {
        static bool attr_set = false;
        const int smem = TEJ * TEK * ESTRIDE * (int)sizeof(double);
        if (!attr_set) {
            cudaFuncSetAttribute(assemble_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(assemble_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set = true;
        }
        const dim3 grid((NX + 1 + TJ - 1) / TJ, (NY + 1 + TK - 1) / TK);
        if (planestress)
            assemble_tile_kernel<true><<<grid, TJ * TK * 2, smem, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
        else
            assemble_tile_kernel<false><<<grid, TJ * TK * 2, smem, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
    } else {
        const int64_t Ndof = 2ll * (NX + 1) * (NY + 1);
        const unsigned grid = (unsigned)((Ndof + 127) / 128);
        if (planestress)
            assemble_rows_kernel<true><<<grid, 128, 0, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
        else
            assemble_rows_kernel<false><<<grid, 128, 0, lc.stream>>>(NX, NY, elem_desc, elstiff, rowptr, val);
    }
    (*launches)++;
    return cudaGetLastError();
}

void assembly_set_tiled(int on) { g_tiled = on; }

cudaError_t launch_kel(const LaunchCtx &lc, int64_t Nel, int planestress, const int32_t *elem_desc, const double *elstiff,
                       double *Kel, int64_t *launches)
{
    const unsigned grid = (unsigned)((Nel * 8 + 127) / 128);
    if (planestress)
        kel_kernel<true><<<grid, 128, 0, lc.stream>>>(Nel, elem_desc, elstiff, Kel);
    else
        kel_kernel<false><<<grid, 128, 0, lc.stream>>>(Nel, elem_desc, elstiff, Kel);
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace plf
