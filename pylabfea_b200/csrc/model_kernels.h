// model_kernels.h -- launchers of model_kernels.cu
#pragma once
#include <algorithm>
#include "plf_host.h"
#include "plf_comm.h"

namespace plf {

struct DevMat;
struct MgHierarchy;

// sweep statistics block in device scalars (mirrored to pinned host memory after every sweep)
struct SweepStats {
    unsigned int change;            // any tangent replaced (model.py:1355)
    unsigned int nonconv;           // elements with f > yf_tolerance * 1.0001 (1361)
    double max_hh;
    unsigned long long counts[4];   // elastic, 1-step, 50-sub-step, radial corrections
    unsigned int work_count;        // deferred (50-sub-step) elements of the compacted pass
    unsigned int pad;
    unsigned long long queue;       // SVC dynamic element queue
};

// CUDA graph of one chunk of CG iterations, owned by the context whose buffers it captured
struct CgGraphCache {
    cudaGraphExec_t exec = nullptr;
    int64_t n = 0;
    const void *key[6] = {nullptr};
    int chunk = 0;
    cudaStream_t stream = nullptr;
    int64_t launches_per_chunk = 0;
};

cudaError_t modelk_init();
cudaError_t modelk_upload_materials(const DevMat *h, int n, cudaStream_t st);
cudaError_t modelk_upload_geom(const GeomClass *h, int n, cudaStream_t st);
cudaError_t launch_sweep(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t Nel, int NY, int nit, const int32_t *desc,
                         const double *sig, const double *epl, const double *du, double *elstiff, double *res_sig,
                         double *res_depl, double *fyld, int32_t *nsteps, unsigned char *has_res, int32_t *worklist,
                         void *d_stats, void *h_stats, int compact, int64_t *launches, int64_t stat_first = 0);
cudaError_t launch_commit(const LaunchCtx &lc, int64_t Nel, int64_t Ndof, int NY, const int32_t *desc, double *u, const double *du,
                          double *f, const int32_t *rowptr, const int32_t *col, const double *val, const double *elstiff,
                          double *sig, double *eps, double *epl, const double *res_sig, const double *res_depl,
                          const unsigned char *has_res, int64_t *launches);
cudaError_t launch_global_sums(const LaunchCtx &lc, int64_t Nel, const int32_t *desc, const double *sig, const double *eps,
                               const double *epl, double *partial, unsigned int *counter, double *out, int64_t *launches,
                               int64_t e_first = 0);
cudaError_t launch_scf(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t Nel, int NY, const int32_t *desc, const double *sig,
                       const double *epl, const double *du, const double *elstiff, const double *d_sld, double *hval,
                       unsigned char *mult, double *partial, unsigned int *counter, double *d_out, double *h_out, double *scf,
                       int64_t *launches, int64_t e_first = 0, Comm *comm = nullptr);
cudaError_t launch_cg(const LaunchCtx &lc, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                      const unsigned char *fixed, const double *b, double *x, double *r, double *z, double *p, double *q,
                      double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc, double rtol, int max_iter,
                      int check_every, int *iters, double *relres, int64_t *launches, MgHierarchy *mg /* nullable: Jacobi */,
                      int warm /* 1: start from the previous content of x */, int NX, int NY /* structured mesh: the
                      matrix product uses the closed-form CSR layout; NX = 0: generic CSR kernel */, CgGraphCache &g_cg,
                      int mg_setup_needed = 1 /* 0: the hierarchy was set up for this very K and mask */);
cudaError_t launch_spmv(const LaunchCtx &lc, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val, const double *x,
                        const double *rhs, double *y, const unsigned char *masked, int mode, int64_t *launches);
void cg_graph_reset(CgGraphCache &g);   // drops the cached graph (call before freeing the buffers it captured)
cudaError_t launch_cg_dist(const LaunchCtx &lc, Comm &comm, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                           const unsigned char *mask, const double *b, double *x, double *r, double *z, double *p, double *q,
                           double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc, double rtol,
                           int max_iter, int check_every, int *iters, double *relres, int64_t *launches, int warm, int NX, int NY,
                           MgHierarchy *mg = nullptr /* local V-cycle as preconditioner; nullptr: Jacobi */, int mg_setup_needed = 1);
struct MgDist;
cudaError_t launch_cg_hybrid(const LaunchCtx &lc, Comm &comm, int a_col, int b_col, int64_t n, const int32_t *rowptr, const int32_t *col,
                             const double *val, const unsigned char *mask, const double *b, double *x, double *r, double *z,
                             double *p, double *q, double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc,
                             double rtol, int max_iter, int *iters, double *relres, int64_t *launches, int warm, int NX, int NY,
                             MgHierarchy &mg, const MgDist &md, int mg_setup_needed, const unsigned char *fixed);
cudaError_t launch_ghost_mask(const LaunchCtx &lc, int64_t n, int NY, int own_first, int own_end, const unsigned char *fixed,
                              unsigned char *mask, int64_t *launches);
cudaError_t launch_merge_du(const LaunchCtx &lc, int64_t n, const double *x, const double *du_bc, const unsigned char *fixed, double *du,
                            int64_t *launches);
cudaError_t launch_bc_scatter(const LaunchCtx &lc, int n, const int32_t *dof, const unsigned char *fx, const double *vals,
                              unsigned char *fixed, double *du, double *w, double *fext, int64_t *launches);
cudaError_t launch_edge_sums(const LaunchCtx &lc, int NX, int NY, int c0, int c1, const double *u, const double *f, double *out,
                             int64_t *launches);

}  // namespace plf
