// plf_host.h -- host-side declarations shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/plfb200.h"

namespace plf {

struct DevMat;

// Geometry class of an element: everything Element.__init__ / calc_Bmat (model.py:262-363, 439-512) derive from
// (lx, ly, thick[, material for plane stress]).  B_g of Gauss point g is described by its non-zeros:
//   B[0][2l] = p[g][l],  B[1][2l+1] = q[g][l],  B[5][2l] = q[g][l],  B[5][2l+1] = p[g][l],  B[2][i] = r[g][i]
struct GeomClass {
    double p[4][4];
    double q[4][4];
    double r[4][8];     // plane-stress row 2 (model.py:498-501), zero for plane strain
    double jacw;        // Jac * wght = 4 * Vel * 1 (model.py:316-322, 340)
    double vel;         // element volume
    double bsum[6][8];  // sum_g B_g: the centroid strain operator of Element.deps / eps_t (model.py:387-411)
};

struct LaunchCtx {
    cudaStream_t stream;
    int sm_count;
    int passa_direct = 0;  // 1: pass A writes its records straight from registers (256-bit stores) instead of staging them
    int max_fm = 4;        // cap of the single-material fast mode (hill_fast.cuh)
    int k1_only = 0;       // diagnostics: 1 = launch pass A only
    int passb_recur = 1;   // 1: planar isotropic J2 loop of pass B with scalar recurrences on the carried chain (hill_fast.cuh)
    int passa_bulk = 0;      // 1: pass A moves its records with cp.async.bulk (TMA engine: measured 6 % slower), 0: cp.async ring + staged stores
    int assembly_tiled = 2;  // 2: row per thread with staged coalesced stores (production); 0: direct stores
    char *trace = nullptr;   // when set (256 bytes): the launchers append the names of the kernel variants they launch, ';'-separated
};

// appends "name<a, b, c>;" to lc.trace (the spelling ncu prints), used by plf_last_kernels
void trace_kernel(const LaunchCtx &lc, const char *name, int a, int b, int c, int d = -1);

// response_kernels.cu
cudaError_t response_init();
cudaError_t response_upload_materials(const DevMat *h, int n, cudaStream_t st);
// fast mode the Hill / J2 materials of a material set share (hill_fast.cuh; 0: general form), capped by max_fm
int response_fast_mode(const DevMat *hmats, int nmat, int max_fm);
cudaError_t launch_response(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t n, const int32_t *mat_id,
                            const double *sig, const double *epl, const double *deps, double *fy, double *sig_out,
                            double *depl, double *ctan, int32_t *nsteps, int32_t *worklist, unsigned int *work_count,
                            unsigned long long *queue, int compact, int64_t *launches);
// compact results for the host: elastic bit per element (words), upper triangles of the other tangents in element order
cudaError_t launch_pack(const LaunchCtx &lc, int64_t n, const int32_t *mat_id, int single, const double *ctan, uint32_t *words,
                        int32_t *block_counts, int32_t *block_offsets, int64_t *total, double *ct21, int64_t *launches);
cudaError_t launch_prim(const LaunchCtx &lc, const DevMat *hmats, int op, int mid, int64_t n, const double *sig,
                        const double *epl, const double *deps, const double *ld, double *out, int64_t *launches);

// assembly_kernels.cu
cudaError_t assembly_upload_geom(const GeomClass *h, int n, cudaStream_t st);
cudaError_t launch_csr_pattern(const LaunchCtx &lc, int NX, int NY, int32_t *rowptr, int32_t *col, int64_t *launches);
cudaError_t launch_assemble(const LaunchCtx &lc, int NX, int NY, int planestress, int ngeom, const int32_t *elem_desc,
                            const double *elstiff, const int32_t *rowptr, double *val, int64_t *launches);
int64_t csr_column_start(int NX, int NY, int j);
cudaError_t launch_assemble_columns(const LaunchCtx &lc, int NX, int NY, int planestress, int ngeom, int j_first, int j_last,
                                    const int32_t *elem_desc, const double *elstiff, double *val, int64_t *launches);
cudaError_t launch_kel(const LaunchCtx &lc, int64_t Nel, int planestress, const int32_t *elem_desc, const double *elstiff,
                       double *Kel, int64_t *launches);

}  // namespace plf
