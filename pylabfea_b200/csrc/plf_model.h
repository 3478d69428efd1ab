// plf_model.h -- the context object behind the C-ABI handle and the device-resident model state.
#pragma once
#include <string>
#include <vector>

#include "plf_device.cuh"
#include "plf_host.h"
#include "model_kernels.h"

// device buffers of one batch of independent stress updates (AoS records, see DESIGN.md "data layout")
namespace plf { struct MgHierarchy; struct Comm; }

struct plf_solve_state;   // plf_solve.cu: boundary description and staging buffers of the load-increment controller

struct plf_batch {
    int64_t n = 0, cap = 0;
    int32_t *mat_id = nullptr;
    double *sig = nullptr, *epl = nullptr, *deps = nullptr;                       // in:  (n,6) each
    double *fy = nullptr, *sig_out = nullptr, *depl = nullptr, *ctan = nullptr;   // out: (n), (n,6), (n,6), (n,36)
    int32_t *nsteps = nullptr;
    int32_t *worklist = nullptr;   // compaction list of 50-sub-step elements
    // plf_step_host_packed (allocated on first use): packed upper triangles, elastic bits, block counts / offsets, totals
    double *ct21 = nullptr;
    uint32_t *el_words = nullptr;
    int32_t *pk_counts = nullptr, *pk_offsets = nullptr;
    int64_t *pk_totals = nullptr;      // device, one per pipeline chunk
    int64_t *h_pk_totals = nullptr;    // pinned mirror
    bool mat_from_mesh = false;        // mat_id holds the mesh's material ids (multi-material plf_step_host)
};

struct plf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;            // copy streams of plf_step_host (created on first use)
    std::vector<cudaEvent_t> ev_step;                        // per-chunk events of plf_step_host
    std::vector<cudaEvent_t> ev_pool;
    std::string err;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    int64_t mem_bytes = 0;
    char dev_name[128] = {0};
    int64_t launches = 0;
    unsigned char *d_flush = nullptr;
    int flush_val = 0;
    unsigned char *d_scalars = nullptr;   // 4 KB of device scratch scalars (counters, reductions)
    unsigned char *h_scalars = nullptr;   // pinned mirror
    int compact = 1;                      // 1: two-pass (compacted) return mapping, 0: unsplit
    int passa_direct = 0;                 // pass A: shared-memory staged coalesced stores (0, fastest measured) or per-thread 256-bit stores (1)
    int max_fm = 4;                       // cap of the single-material fast mode (hill_fast.cuh)
    int k1_only = 0;                      // diagnostics: 1 = launch pass A only
    int passb_recur = 1;                  // pass B, planar isotropic J2: recurrence form of the sub-step loop (0: direct form; env PLF_PASSB_RECUR)

    // materials
    int nmat = 0, planestress = 0;
    plf::DevMat hmats[PLF_MAX_MATERIALS];
    std::vector<plf_material> mats;
    double *d_sv[PLF_MAX_MATERIALS] = {nullptr};
    double *d_dual[PLF_MAX_MATERIALS] = {nullptr};
    plf_batch scratch;

    // mesh (Model.mesh, model.py:758-952)
    bool mesh_ready = false, K_ready = false;
    int NX = 0, NY = 0;
    int64_t Nel = 0, Nnode = 0, Ndof = 0, nnz = 0;
    double dy = 0., thick = 0.;
    std::vector<double> dx;
    std::vector<int32_t> h_desc;   // (geometry class << 16) | material id, per element
    int ngeom = 0;
    plf::GeomClass geom[PLF_MAX_GEOM];
    double geom_B[PLF_MAX_GEOM][4][48];   // full B matrices per class (plf_get_bmat)
    double volume = 0.;

    int32_t *d_desc = nullptr;
    double *d_elstiff = nullptr;                                                  // Element.elstiff (Nel,36)
    double *d_sig = nullptr, *d_eps = nullptr, *d_epl = nullptr;                  // Element.sig/eps/epl (Nel,6)
    double *d_res_sig = nullptr, *d_res_depl = nullptr;                           // Element.res_sig/res_depl
    double *d_fyld = nullptr;                                                     // f list (model.py:1345)
    int32_t *d_nsteps = nullptr, *d_branch = nullptr;
    unsigned char *d_has_res = nullptr;                                           // el.res_sig is not None
    double *d_u = nullptr, *d_du = nullptr, *d_f = nullptr;                       // Model.u/du/f (Ndof)
    int32_t *d_rowptr = nullptr, *d_col = nullptr;
    double *d_val = nullptr;                                                      // CSR(K)
    int32_t *d_worklist = nullptr;
    // solver work vectors
    double *d_r = nullptr, *d_z = nullptr, *d_p = nullptr, *d_q = nullptr, *d_b = nullptr, *d_diag = nullptr, *d_w = nullptr;
    unsigned char *d_fixed = nullptr;
    double *d_bc_du = nullptr, *d_bc_f = nullptr;   // calc_BC's du and force vector of the last plf_solve_linear (d_w: its w)
    bool bc_valid = false;
    plf::MgHierarchy *mg = nullptr;   // multigrid preconditioner of plf_solve_linear (built with the mesh)
    int solver_mg = 1;                // 1: V-cycle preconditioner when the mesh coarsens at least once, 0: Jacobi
    int solver_warm = 1;              // CG starts from the previous du (0: from zero)
    int mg_nu = 2, mg_coarse_sweeps = 10;   // smoothing sweeps per level (even) / on the coarsest level
    int mg_fuse_tail = 2;             // coarse levels of the V-cycle in one fused kernel (mg_kernels.cu)
    double mg_omega = 0.7;            // Jacobi damping
    plf::Comm *comm = nullptr;       // NCCL communicator + strip neighbourhood (plf_comm.cu); nullptr: single GPU
    unsigned char *d_cgmask = nullptr;   // fixed | ghost-column mask of the distributed CG
    double *d_partial = nullptr;   // block partial sums
    int n_partial = 0;
    plf::CgGraphCache cg_graph;    // captured chunk of CG iterations over this context's buffers
    int assembly_tiled = 2;        // option "assembly_tiled"
    int passa_bulk = 0;            // option "passa_bulk" (1: cp.async.bulk variant of pass A; measured slower)
    // The kernels read materials and geometry classes from __constant__ banks, which are per device, not per context.
    // tables_dirty: this context's host tables changed since it last uploaded them; plf_bind() (head of every launch
    // sequence) re-uploads when the banks hold another context's tables or stale ones.
    bool tables_dirty = true;
    plf_solve_state *solve = nullptr;   // set by plf_solve_begin
    // set on the GLOBAL context of a replicated solve while level 0 of its solver is shared among the ranks (solver_hybrid):
    // the communicator of the strip context and the global node columns [share_a, share_b) this rank owns
    plf::Comm *share_comm = nullptr;
    int share_a = 0, share_b = 0;
    int solver_hybrid = 2;              // option (on the strip context): share the finest levels + CG of the replicated solve among
                                        // the ranks: 0 never, 1 always, 2 when the exchanges run through peer mailboxes (Comm::p2p)
                                        // (same iterates; ~12 NCCL calls per iteration: pays only with low-latency exchanges)
    plf_ctx *replica = nullptr;         // plf_comm_set_replica: context holding the GLOBAL mesh, solves the gathered system (not owned)
    // change tracking, so that an unchanged K is neither re-assembled nor re-coarsened: tangent_epoch moves whenever
    // element tangents change (plf_set_tangent, a sweep that replaced one), K_serial with every assembly, mask_epoch
    // with every new set of boundary-condition vectors
    char trace[256] = {0};    // kernel variants launched since the last plf_last_kernels (when trace_on)
    int trace_on = 0;
    int skip_unchanged = 1;   // option: 0 re-assembles / re-coarsens even when nothing changed (A/B, tests)
    int64_t tangent_epoch = 1, K_tangent_epoch = 0, K_serial = 0, mask_epoch = 0, mg_K_serial = -1, mg_mask_epoch = -1;
};

int fail(plf_ctx *c, int code, const std::string &msg);
// Makes this context the owner of the device's __constant__ material / geometry banks (uploads its tables if another
// context used the device in between or the tables changed).  Returns PLF_OK or PLF_ECUDA.
int plf_bind(plf_ctx *ctx);
int comm_host_allreduce(plf_ctx *ctx, double *vals, int n, int op);   // plf_comm.cu: no-op without a communicator
void solve_state_free(plf_ctx *ctx);            // plf_solve.cu
void solve_state_invalidate_bc(plf_ctx *ctx);   // plf_solve.cu: the device BC vectors were overwritten by the caller
int model_reset_state(plf_ctx *ctx);            // zero u, f, element state; elstiff = CV (first Model.solve, model.py:1214-1227)
int model_solver_buffers(plf_ctx *ctx);         // work vectors of plf_solve_linear
// reduced solve with the boundary-condition vectors already on the device (d_fixed, d_bc_du, d_w, d_bc_f)
int model_solve_device_bc(plf_ctx *ctx, double rtol, int max_iter, int *iters, double *relres);
// Counters of the batch return mapping (launch_response: two-ended work list + finished blocks, SVC element queue) inside the
// context's 4 KB scalar page.  Their own slot: the page's first 256 bytes are the SweepStats of plf_sweep, and the counters
// must read zero when a batch starts (zeroed at context creation, re-zeroed by the last block of every launch) -- a sweep
// followed by a batch call on the same context used to find a branch count in them and overran the work list.
inline unsigned int *batch_counters(plf_ctx *c) { return (unsigned int *)(c->d_scalars + 2048); }
inline unsigned long long *batch_queue(plf_ctx *c) { return (unsigned long long *)(c->d_scalars + 2064); }

inline plf::LaunchCtx make_lc(const plf_ctx *c)
{
    plf::LaunchCtx lc{c->stream, c->sm_count, c->passa_direct, c->max_fm, c->k1_only, c->passb_recur, c->passa_bulk, c->assembly_tiled};
    lc.trace = c->trace_on ? const_cast<char *>(c->trace) : nullptr;
    return lc;
}
void model_free(plf_ctx *ctx);

void model_apply_mg_options(plf_ctx *ctx);   // copies mg_nu / mg_coarse_sweeps / mg_omega into the hierarchy
