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
#include "csr_stencil.cuh"
#include "plf_device.cuh"
#include "plf_host.h"

namespace plf {

__constant__ GeomClass c_geom[PLF_MAX_GEOM];

cudaError_t assembly_upload_geom(const GeomClass *h, int n, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_geom, h, sizeof(GeomClass) * n, 0, cudaMemcpyHostToDevice, st);
}

// Row starts come in closed form (csr_stencil.cuh) instead of rowptr loads: a dependent load at the top and at the
// bottom of every thread was 17 % of the assembly kernel's stall samples.

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
// The row needs 8 (plane strain) or 12 (plane stress) of the 36 tangent entries; they are fetched with 128-bit loads
// (entry pairs [6c, 6c+1], [6c+4, 6c+5], [30, 31], [34, 35]: all 16 B aligned in the 288 B record).  With scalar
// loads the kernel was bound by L1 wavefronts (every warp-level load touches 16 different records), not by HBM.
template <bool PS>
struct CRow {   // the tangent entries one Kel row needs
    double c10, c11, c12, c15, c50, c51, c52, c55, c20, c21, c22, c25;
};

template <bool PS>
__device__ __forceinline__ CRow<PS> load_crow(const double *__restrict__ C, int c)
{
    CRow<PS> r;
    const double2 ra = __ldg(reinterpret_cast<const double2 *>(C + 6 * c));       // ux rows use C row 0, uy rows use C row 1
    const double2 rb = __ldg(reinterpret_cast<const double2 *>(C + 6 * c + 4));
    const double2 r5a = __ldg(reinterpret_cast<const double2 *>(C + 30));
    const double2 r5b = __ldg(reinterpret_cast<const double2 *>(C + 34));
    r.c10 = ra.x; r.c11 = ra.y; r.c15 = rb.y;
    r.c50 = r5a.x; r.c51 = r5a.y; r.c55 = r5b.y;
    r.c12 = r.c52 = r.c20 = r.c21 = r.c22 = r.c25 = 0.;
    if (PS) {
        r.c12 = __ldg(C + 6 * c + 2);
        r.c52 = __ldg(C + 32);
        const double2 r2a = __ldg(reinterpret_cast<const double2 *>(C + 12));
        const double2 r2b = __ldg(reinterpret_cast<const double2 *>(C + 14));
        const double2 r2c = __ldg(reinterpret_cast<const double2 *>(C + 16));
        r.c20 = r2a.x; r.c21 = r2a.y; r.c22 = r2b.x; r.c25 = r2c.y;
    }
    return r;
}

template <bool PS>
__device__ __forceinline__ void kel_row_from(const GeomClass &gc, const CRow<PS> &R, int a, int c, double *kel);

template <bool PS>
__device__ __forceinline__ void kel_row(const GeomClass &gc, const double *__restrict__ C, int a, int c, double *kel)
{
    const CRow<PS> R = load_crow<PS>(C, c);
    kel_row_from<PS>(gc, R, a, c, kel);
}

template <bool PS>
__device__ __forceinline__ void kel_row_from(const GeomClass &gc, const CRow<PS> &R, int a, int c, double *kel)
{
    const double c10 = R.c10, c11 = R.c11, c12 = R.c12, c15 = R.c15, c50 = R.c50, c51 = R.c51, c52 = R.c52, c55 = R.c55;
    const double c20 = R.c20, c21 = R.c21, c22 = R.c22, c25 = R.c25;
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

// UNIFORM: the mesh has a single geometry class -> no descriptor loads.  MINB: blocks per SM asked of the register
// allocator (5 measured faster than 6, which spills).  All four elements' operands are requested before the first one is used (indices clamped into the mesh,
// so the loads are unconditional; contributions of elements outside the mesh are skipped): four independent memory
// round trips per thread instead of four dependent ones.
template <bool PS, bool UNIFORM, int MINB, bool REVERSE = true>
__global__ void __launch_bounds__(2 * SN, MINB) assemble_rows_staged_kernel(int NX, int NY, int j_first,
                                                                            const int32_t *__restrict__ elem_desc,
                                                                            const double *__restrict__ elstiff,
                                                                            double *__restrict__ val)
{
    __shared__ double2 s_out[SN * 2 * 9];
    const int tid = threadIdx.x;
    const int NnY = NY + 1;
    // node column j (a launch covers columns j_first ...), highest column first: the return mapping that ran just before
    // wrote the tangents of the highest element columns last, so they are still L2-resident when this kernel starts
    // (measured 2-3 % of the kernel at 1024^2)
    const int j = j_first + (REVERSE ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y), k0 = blockIdx.x * SN;
    const int c = tid & 1, k = k0 + (tid >> 1);
    const bool active = k <= NY;
    const int kmax = min(k0 + SN - 1, NY);
    const int64_t r_first = 2 * ((int64_t)j * NnY + k0);
    const int64_t r_last = 2 * ((int64_t)j * NnY + kmax) + 1;
    const int64_t begin = csr_row_start(NX, NY, r_first), end = csr_row_start(NX, NY, r_last + 1);

    if (active) {
        double acc[9][2];
#pragma unroll
        for (int s = 0; s < 9; s++) acc[s][0] = acc[s][1] = 0.;
        CRow<PS> R[4];
        int cls[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int ej = min(max(j - 1 + (q >> 1), 0), NX - 1), ek = min(max(k - 1 + (q & 1), 0), NY - 1);
            const int64_t e = (int64_t)ej * NY + ek;
            cls[q] = UNIFORM ? 0 : (__ldg(elem_desc + e) >> 16);
            R[q] = load_crow<PS>(elstiff + e * 36, c);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int ddj = q >> 1, ddk = q & 1;
            const int ej = j - 1 + ddj, ek = k - 1 + ddk;   // ascending element index e = ej*NY + ek
            if (ej < 0 || ej >= NX || ek < 0 || ek >= NY) continue;
            const int a = 2 * (1 - ddj) + (1 - ddk);
            double kel[8];
            kel_row_from<PS>(c_geom[cls[q]], R[q], a, c, kel);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int sj = ddj + (m >> 1), sk = ddk + (m & 1);
                acc[3 * sj + sk][0] = acc[3 * sj + sk][0] + kel[2 * m];
                acc[3 * sj + sk][1] = acc[3 * sj + sk][1] + kel[2 * m + 1];
            }
        }
        const int64_t r = 2 * ((int64_t)j * NnY + k) + c;
        int pos2 = (int)((csr_row_start(NX, NY, r) - begin) >> 1);
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
    const int n2 = (int)((end - begin) >> 1);
    double2 *dst = reinterpret_cast<double2 *>(val + begin);
    for (int i = tid; i < n2; i += 2 * SN) dst[i] = s_out[i];
}

// LaunchCtx::assembly_tiled -- 0: row per thread, direct stores; 2: row per thread, staged stores (production)
// (a shared-memory tiled variant -- coalesced tile loads of the tangents, 3 barriers -- measured 0.27 ms against
//  0.15 ms for the staged row kernel at 1024^2 and was removed)

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

// first CSR entry of node column j (host side of csr_row_start)
int64_t csr_column_start(int NX, int NY, int j)
{
    const int64_t colw = 4ll * (3 * NY + 1);
    if (j > NX) return colw * (3ll * (NX + 1) - 2);
    return colw * ((j == 0) ? 0 : 3ll * j - 1);
}

// node columns [j_first, j_last) only (plf_step_host assembles a column range as soon as its elements are updated)
cudaError_t launch_assemble_columns(const LaunchCtx &lc, int NX, int NY, int planestress, int ngeom, int j_first, int j_last,
                                    const int32_t *elem_desc, const double *elstiff, double *val, int64_t *launches)
{
    if (j_last <= j_first) return cudaSuccess;
    const dim3 grid((NY + 1 + SN - 1) / SN, j_last - j_first);
#define ASM(PS, U, M) assemble_rows_staged_kernel<PS, U, M><<<grid, 2 * SN, 0, lc.stream>>>(NX, NY, j_first, elem_desc, elstiff, val)
    const bool uni = (ngeom == 1);
    if (planestress) { if (uni) ASM(true, true, 4); else ASM(true, false, 4); }
    else { if (uni) ASM(false, true, 5); else ASM(false, false, 5); }
#undef ASM
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_assemble(const LaunchCtx &lc, int NX, int NY, int planestress, int ngeom, const int32_t *elem_desc,
                            const double *elstiff, const int32_t *rowptr, double *val, int64_t *launches)
{
    if (lc.assembly_tiled != 0) {
        const dim3 grid((NY + 1 + SN - 1) / SN, NX + 1);
#define ASM(PS, U, M)                                                      \
    do {                                                                   \
        trace_kernel(lc, "assemble_rows_staged_kernel", PS, U, M, 1);      \
        assemble_rows_staged_kernel<PS, U, M><<<grid, 2 * SN, 0, lc.stream>>>(NX, NY, 0, elem_desc, elstiff, val); \
    } while (0)
        const bool uni = (ngeom == 1);
        if (planestress) { if (uni) ASM(true, true, 4); else ASM(true, false, 4); }
        else { if (uni) ASM(false, true, 5); else ASM(false, false, 5); }
#undef ASM
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
