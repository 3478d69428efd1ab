// mg.h -- geometric multigrid preconditioner of the on-device CG solve (mg_kernels.cu)
#pragma once
#include <vector>
#include "plf_host.h"

namespace plf {

constexpr int MG_MAX_LEVELS = 16;

struct MgLevel {
    int NX = 0, NY = 0;
    int64_t ndof = 0, nnz = 0;
    float *val = nullptr;             // level matrix, FP32, closed-form CSR layout (level 0: converted copy of the model's K)
    unsigned char *mask = nullptr;    // 1 = Dirichlet dof (level 0: the solver's mask, not owned)
    double *dinv = nullptr;           // 1 / diag, 0 on masked rows
    double *x = nullptr, *b = nullptr, *r = nullptr, *t = nullptr;   // level 0: x and b are the caller's z and r
};

struct MgHierarchy {
    std::vector<MgLevel> lv;
    int nu = 2;                // pre- and post-smoothing sweeps (even)
    int coarse_sweeps = 10;    // damped Jacobi sweeps on the coarsest level (40 measured 15 % slower per solve, 4 % fewer iterations)
    double omega = 0.7;
    int j0 = 0;                // first node column of level 0 that belongs to the hierarchy (1: left halo column of a strip)
    int tail_first = -1;       // first level of the fused single-block tail of the cycle (-1: none fits)
    int fuse_tail = 2;         // option "mg_fuse_tail": 0 every level as its own kernels, 1 fused single-block tail, 2 (default)
                               // additionally the cooperative grid kernel for the levels between the fine ones and the tail
    int coop_first = -1;       // first level run by the cooperative kernel (-1: none)
    int coop_blocks = 64;
    unsigned int *d_bar = nullptr;   // arrival counter of the cooperative kernel's grid barrier
    int levels() const { return (int)lv.size(); }
};

cudaError_t mg_build(MgHierarchy &H, int NX, int NY, int j0 = 0);
void mg_free(MgHierarchy &H);
cudaError_t mg_setup(const LaunchCtx &lc, MgHierarchy &H, const double *val_fine, const unsigned char *mask_fine, int64_t *launches);
struct Comm;
// the finest `nshared` levels of the hierarchy shared among the ranks of a replicated solve (mg_vcycle)
struct MgDist {
    Comm *comm = nullptr;
    int nshared = 1;
    int a[8] = {}, b[8] = {};         // node columns [a[l], b[l]) of level l this rank owns (a[l+1] = ceil(a[l] / 2), b alike)
    int64_t goff[64], cnt[64];        // per rank: start and length (doubles) of its part of the right-hand side of level nshared
};
// how many levels to share: level 0, and every further level that is large enough to repay its ghost exchanges, runs as
// per-level kernels and leaves every rank at least two node columns; fills md.nshared, md.a / md.b (from a[0], b[0]), goff, cnt
void mg_plan_sharing(const MgHierarchy &H, int world, int rank, MgDist &md);
cudaError_t mg_vcycle(const LaunchCtx &lc, MgHierarchy &H, const double *r_fine, double *z_fine, int64_t *launches,
                      const MgDist *dist = nullptr);

}  // namespace plf
