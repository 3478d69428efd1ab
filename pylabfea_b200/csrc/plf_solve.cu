// plf_solve.cu -- the load-increment controller behind Model.solve (model.py:1208-1450) on the library side of the
// C-ABI: boundary-condition vectors (calc_BC, 1070-1206), the plan / scale / halve / advance rules of the incremented
// boundary values, the sequence predictor -> calc_scf -> inner passes -> commit -> calc_global, and Model.bcval.
// Host code; every O(Nel) / O(Ndof) step is one of the launch sequences of plf_model.cu.
//
// Behaviour is the reference's (same increments, same passes); the structure is not: the three incremented boundary
// conditions (right edge, top edge, node set) are `Channel`s with a target, the value reached so far, the step of the
// current increment and the limit the halving rule refers to, and the controller drives them through plan(), scale(),
// halve() and advance().  With a communicator (plf_comm) the same controller runs on every rank of a strip partition:
// sweep flags, calc_scf moments, the calc_global sums and the boundary sums are combined over the ranks, so all ranks
// take identical decisions.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "model_kernels.h"
#include "plf_comm.h"
#include "plf_model.h"

using namespace plf;

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PLF_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

namespace {

// One incremented boundary condition (model.py:1271-1288): `target` is what the user asked for (Model.bcr / bct / bcn),
// `reached` what the committed increments add up to (bcr0 / bct0 / bcn0).
struct Channel {
    double target[2], reached[2], limit[2], step[2];
    // The reference binds the node-set step and its limit to ONE array (dbcn = max_dbcn, model.py:1285), so halving the
    // step moves the limit with it; right and top get a scaled copy (1297-1298) and keep their limit.
    bool limit_follows_step = false;

    void plan(int min_step, int il)   // 1272-1288: what is left, over the increments still wanted
    {
        for (int k = 0; k < 2; k++) {
            double d = target[k] - reached[k];
            if (min_step > 0) d /= (double)std::max(1, min_step - il);
            limit[k] = step[k] = d;
        }
    }
    void scale(double s)   // 1297-1298
    {
        for (int k = 0; k < 2; k++) step[k] = limit[k] * s;
    }
    void halve()   // 1310-1330: half the step, at least 5 % of the limit, at most what is left
    {
        for (int k = 0; k < 2; k++) {
            const double left = target[k] - reached[k], half = step[k] * 0.5, floor5 = 0.05 * limit[k];
            step[k] = (limit[k] >= 0.) ? std::max(floor5, std::min(left, half)) : std::min(floor5, std::max(left, half));
            if (limit_follows_step) limit[k] = step[k];
        }
    }
    void advance()   // 1398, 1402, 1406
    {
        for (int k = 0; k < 2; k++) reached[k] += step[k];
    }
    bool pending() const   // 1399-1409
    {
        for (int k = 0; k < 2; k++)
            if (std::fabs(reached[k] - target[k]) > 1.e-6 && std::fabs(target[k]) > 1.e-9) return true;
        return false;
    }
};

struct BcEntry {
    unsigned char fixed = 0;
    double du = 0., w = 0., fext = 0.;
};

// Sparse restatement of calc_BC on this context's node grid.  In the reference every displacement value applied to a
// dof multiplies column i of K into df (even when the dof was already removed by an earlier edge: the sum of all values
// is `w`), the FIRST value wins for du, and a differing later value only warns.
struct BcBuilder {
    std::unordered_map<int32_t, size_t> slot;
    std::vector<int32_t> dof;
    std::vector<BcEntry> ent;
    int n_warn = 0;
    std::string first_warn;

    BcEntry &at(int32_t d)
    {
        auto it = slot.find(d);
        if (it != slot.end()) return ent[it->second];
        slot.emplace(d, dof.size());
        dof.push_back(d);
        ent.emplace_back();
        return ent.back();
    }
    void disp(int64_t node, int k, double val, const char *who, int64_t node_global)
    {
        BcEntry &e = at((int32_t)(2 * node + k));
        if (!e.fixed) {
            e.fixed = 1;
            e.du = val;
        } else if (e.du != val) {
            if (n_warn++ == 0) {
                char buf[200];
                snprintf(buf, sizeof buf, "Inconsistent BC at %s node %lld (%g vs %g).", who, (long long)node_global, e.du, val);
                first_warn = buf;
            }
        }
        e.w += val;
    }
    void force(int64_t node, int k, double val) { at((int32_t)(2 * node + k)).fext += val; }
};

}   // namespace

// what plf_solve_begin leaves on the context for the increments that follow
struct plf_solve_state {
    plf_bc bc;
    std::vector<int32_t> noset;
    double sld[6];
    bool nonlin = false;
    // boundary values the device vectors were built from (the passes of an increment mostly repeat them)
    bool bc_built = false;
    double built[6];
    // staging buffers of the scatter
    int cap = 0;
    int32_t *d_dof = nullptr;
    unsigned char *d_fx = nullptr;
    double *d_vals = nullptr;
};

void solve_state_invalidate_bc(plf_ctx *ctx)
{
    if (ctx->solve) ctx->solve->bc_built = false;
}

void solve_state_free(plf_ctx *ctx)
{
    if (!ctx->solve) return;
    cudaFree(ctx->solve->d_dof); cudaFree(ctx->solve->d_fx); cudaFree(ctx->solve->d_vals);
    delete ctx->solve;
    ctx->solve = nullptr;
}

// position of the partition inside the global mesh (single GPU: the mesh itself)
struct GridView {
    int NX, NY, col0, NXg;       // local element columns, first local column's global index, global element columns
    double x0, lenx, leny;        // x of local node column 0, global extent
    bool has_left, has_right;
    int own0, own1;               // owned node columns [own0, own1)
};

static GridView grid_view(const plf_ctx *ctx)
{
    GridView g;
    g.NX = ctx->NX; g.NY = ctx->NY;
    double lx = 0.;
    for (double v : ctx->dx) lx += v;
    g.leny = ctx->dy * ctx->NY;
    const bool dist = ctx->comm && ctx->comm->active() && ctx->comm->strip_set && ctx->comm->global_set;
    if (dist) {
        g.col0 = ctx->comm->first_col_global; g.NXg = ctx->comm->NX_global;
        g.x0 = ctx->comm->x_offset; g.lenx = ctx->comm->lenx_global;
        g.own0 = ctx->comm->own_first; g.own1 = ctx->comm->own_end;
    } else {
        g.col0 = 0; g.NXg = ctx->NX; g.x0 = 0.; g.lenx = lx; g.own0 = 0; g.own1 = ctx->NX + 1;
    }
    g.has_left = (g.col0 == 0);
    g.has_right = (g.col0 + g.NX == g.NXg);
    return g;
}

// calc_BC (model.py:1070-1206) -> device vectors d_fixed, d_bc_du, d_w, d_bc_f
static int build_bc_vectors(plf_ctx *ctx, const double *dr, const double *dt, const double *dn, plf_increment_report *rep)
{
    plf_solve_state &S = *ctx->solve;
    const double now[6] = {dr[0], dr[1], dt[0], dt[1], dn[0], dn[1]};
    if (S.bc_built && ctx->bc_valid && !memcmp(now, S.built, sizeof now)) return PLF_OK;
    const plf_bc &bc = S.bc;
    const GridView g = grid_view(ctx);
    const int NnY = g.NY + 1;
    BcBuilder B;
    B.dof.reserve(4 * (size_t)(g.NX + g.NY + 2));
    B.ent.reserve(4 * (size_t)(g.NX + g.NY + 2));
    auto gnode = [&](int j, int k) { return (int64_t)(g.col0 + j) * NnY + k; };
    // left edge, static (1099-1105)
    if (g.has_left)
        for (int k = 0; k < 2; k++)
            if (bc.left_disp[k])
                for (int i = 0; i < NnY; i++) B.disp(i, k, bc.left[k], "left", gnode(0, i));
    // bottom edge, static (1111-1122)
    for (int k = 0; k < 2; k++)
        if (bc.bot_disp[k])
            for (int j = 0; j <= g.NX; j++) B.disp((int64_t)j * NnY, k, bc.bot[k], "bottom", gnode(j, 0));
    // right edge (1126-1151): displacement, or a force shared 1 / (NnodeY - 1) per node with half weight at the corners
    if (g.has_right)
        for (int k = 0; k < 2; k++) {
            if (bc.right_disp[k]) {
                for (int i = 0; i < NnY; i++) B.disp((int64_t)g.NX * NnY + i, k, dr[k], "right", gnode(g.NX, i));
            } else {
                for (int i = 0; i < NnY; i++) {
                    double share = 1. / (double)(NnY - 1);
                    const double y = i * ctx->dy;
                    if (y < 1.e-3 || y > g.leny - 1.e-3) share *= 0.5;
                    B.force((int64_t)g.NX * NnY + i, k, dr[k] * share);
                }
            }
        }
    // top edge (1156-1179)
    {
        std::vector<double> xs(g.NX + 1);
        xs[0] = g.x0;
        for (int j = 0; j < g.NX; j++) xs[j + 1] = xs[j] + ctx->dx[j];
        for (int k = 0; k < 2; k++) {
            if (bc.top_disp[k]) {
                for (int j = 0; j <= g.NX; j++) B.disp((int64_t)j * NnY + g.NY, k, dt[k], "top", gnode(j, g.NY));
            } else {
                for (int j = 0; j <= g.NX; j++) {
                    double share = 1. / (double)g.NXg;   // NnodeX - 1
                    if (xs[j] < 1.e-3 || xs[j] > g.lenx - 1.e-3) share *= 0.5;
                    B.force((int64_t)j * NnY + g.NY, k, dt[k] * share);
                }
            }
        }
    }
    // node set (1183-1205)
    for (int k = 0; k < 2 && !S.noset.empty(); k++)
        for (int32_t ng : S.noset) {
            const int jg = ng / NnY, kk = ng - jg * NnY, j = jg - g.col0;
            if (j < 0 || j > g.NX) continue;   // another rank's node
            if (bc.node_disp[k]) B.disp((int64_t)j * NnY + kk, k, dn[k], "node set", ng);
            else B.force((int64_t)j * NnY + kk, k, dn[k]);
        }
    if (rep && B.n_warn) {
        rep->n_warnings += B.n_warn;
        if (!rep->warning[0]) snprintf(rep->warning, sizeof rep->warning, "%s", B.first_warn.c_str());
    }
    // scatter into the zeroed dense vectors
    const int n = (int)B.dof.size();
    if (n > S.cap) {
        cudaFree(S.d_dof); cudaFree(S.d_fx); cudaFree(S.d_vals);
        S.d_dof = nullptr; S.d_fx = nullptr; S.d_vals = nullptr;
        S.cap = n + n / 4 + 64;
        CK(cudaMalloc((void **)&S.d_dof, (size_t)S.cap * sizeof(int32_t)));
        CK(cudaMalloc((void **)&S.d_fx, (size_t)S.cap));
        CK(cudaMalloc((void **)&S.d_vals, 3 * (size_t)S.cap * sizeof(double)));
    }
    std::vector<unsigned char> fx(n);
    std::vector<double> vals(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        fx[i] = B.ent[i].fixed;
        vals[i] = B.ent[i].du;
        vals[(size_t)n + i] = B.ent[i].w;
        vals[2 * (size_t)n + i] = B.ent[i].fext;
    }
    int rc = model_solver_buffers(ctx);
    if (rc != PLF_OK) return rc;
    const size_t D = (size_t)ctx->Ndof;
    CK(cudaMemsetAsync(ctx->d_fixed, 0, D, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bc_du, 0, D * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_w, 0, D * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bc_f, 0, D * sizeof(double), ctx->stream));
    if (n > 0) {
        CK(cudaMemcpyAsync(S.d_dof, B.dof.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(S.d_fx, fx.data(), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(S.d_vals, vals.data(), 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        LaunchCtx lc = make_lc(ctx);
        CK(launch_bc_scatter(lc, n, S.d_dof, S.d_fx, S.d_vals, ctx->d_fixed, ctx->d_bc_du, ctx->d_w, ctx->d_bc_f, &ctx->launches));
    }
    CK(cudaStreamSynchronize(ctx->stream));   // the host staging vectors go out of scope
    ctx->mask_epoch++;   // the multigrid hierarchy carries the Dirichlet mask: rebuilt at the next solve
    ctx->bc_valid = true;
    S.bc_built = true;
    memcpy(S.built, now, sizeof now);
    return PLF_OK;
}

extern "C" size_t plf_sizeof_increment_report(void) { return sizeof(plf_increment_report); }

extern "C" int plf_solve_begin(plf_ctx *ctx, const plf_bc *bc, int reset)
{
    if (!ctx) return PLF_EINVAL;
    if (!ctx->mesh_ready) return fail(ctx, PLF_ESTATE, "plf_solve_begin: attributes for mesh not set, but required by solver");
    if (!bc || (bc->n_noset > 0 && !bc->noset)) return fail(ctx, PLF_EINVAL, "plf_solve_begin: bc is NULL or its node set is missing");
    if (!ctx->solve) ctx->solve = new plf_solve_state();
    plf_solve_state &S = *ctx->solve;
    S.bc = *bc;
    S.noset.assign(bc->noset, bc->noset + std::max(0, bc->n_noset));
    S.bc.noset = nullptr;
    S.bc_built = false;   // BC types / node sets may differ from the previous Model.solve call
    const int64_t nn_global = (int64_t)(grid_view(ctx).NXg + 1) * (ctx->NY + 1);
    for (int32_t n : S.noset)
        if (n < 0 || n >= nn_global) return fail(ctx, PLF_EINVAL, "plf_solve_begin: node set index out of range");
    S.nonlin = false;
    for (int i = 0; i < ctx->nmat; i++) S.nonlin |= (ctx->hmats[i].kind != PLF_MAT_ELASTIC);
    // loading direction for the yield-point search of ML yield functions in calc_scf (1245-1258)
    double *s = S.sld;
    for (int k = 0; k < 6; k++) s[k] = 0.;
    auto sgn = [](double v) { return v > 0. ? 1. : (v < 0. ? -1. : 0.); };
    if (std::fabs(bc->right[0]) > 1.e-6) s[0] = sgn(bc->right[0]);
    if (std::fabs(bc->top[1]) > 1.e-6) s[1] = sgn(bc->top[1]);
    if (std::fabs(bc->right[1]) > 1.e-6) s[5] = sgn(bc->right[1]);
    if (std::fabs(bc->top[0]) > 1.e-6) s[5] = sgn(bc->top[0]);
    if (std::sqrt(s[0] * s[0] + s[1] * s[1] + s[5] * s[5]) < 1.e-3) s[0] = 1.;   // the reference warns here (1255-1258)
    int rc;
    if (reset && (rc = model_reset_state(ctx)) != PLF_OK) return rc;
    if (ctx->replica) {
        // the global solver context gets the same boundary description (its vectors are built on the global node grid)
        plf_ctx *g = ctx->replica;
        if (!g->solve) g->solve = new plf_solve_state();
        g->solve->bc = S.bc;
        g->solve->noset = S.noset;
        g->solve->bc_built = false;
        g->solve->nonlin = S.nonlin;
    }
    return plf_setupK(ctx);   // 1242
}

// replicated solve (plf_comm_set_replica): gather the owned K rows of every rank into the global context, solve there,
// bring the strip's part of du back
static int solve_replicated(plf_ctx *ctx, const double *dr, const double *dt, const double *dn, double rtol, int max_iter, int *it,
                            double *rr, plf_increment_report *rep)
{
    plf_ctx *g = ctx->replica;
    Comm &c = *ctx->comm;
    const int NY = ctx->NY, NXg = c.NX_global;
    int64_t goff[64], cnt[64];
    if (c.world > 64) return fail(ctx, PLF_EUNSUPPORTED, "replicated solve: at most 64 ranks");
    for (int r = 0; r < c.world; r++) {   // owned node columns of rank r: [NXg r / world, NXg (r+1) / world) (+ the last column)
        const int a = (int)((int64_t)NXg * r / c.world), b = (int)((int64_t)NXg * (r + 1) / c.world) + (r == c.world - 1 ? 1 : 0);
        goff[r] = csr_column_start(NXg, NY, a);
        cnt[r] = csr_column_start(NXg, NY, b) - goff[r];
    }
    const int64_t lo = csr_column_start(ctx->NX, NY, c.own_first);
    if (csr_column_start(ctx->NX, NY, c.own_end) - lo != cnt[c.rank])
        return fail(ctx, PLF_ESTATE, "replicated solve: the strip does not match the global partition");
    CK(cudaSetDevice(ctx->device));
    // the global context works on its own stream: order it behind this context's assembly and ahead of what follows
    CK(cudaStreamSynchronize(ctx->stream));
    static const bool timing = getenv("PLF_TIMING") != nullptr;   // diagnostics: where a replicated solve spends its time
    static double t_gather = 0., t_solve = 0.;
    static int n_calls = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (comm_gather_ranges(c, ctx->d_val + lo, g->d_val, goff, cnt, g->stream))
        return fail(ctx, PLF_ECUDA, std::string("replicated solve: ") + comm_error());
    if (timing) CK(cudaStreamSynchronize(g->stream));
    const auto t1 = std::chrono::steady_clock::now();
    g->K_ready = true;
    g->K_tangent_epoch = 0;
    g->K_serial++;
    int rc = build_bc_vectors(g, dr, dt, dn, rep);
    if (rc != PLF_OK) { ctx->err = g->err; return rc; }
    // every rank solves the gathered system; with solver_hybrid (2 = when the exchanges run through peer mailboxes: the
    // ranks agreed on c.p2p, so all take the same branch) the CG vectors and the finest levels of the V-cycle are split over the
    // node columns the ranks own (the levels below stay replicated)
    const bool share = (ctx->solver_hybrid == 1 || (ctx->solver_hybrid == 2 && c.p2p)) && c.world > 1;
    g->share_comm = share ? &c : nullptr;
    g->share_a = (int)((int64_t)NXg * c.rank / c.world);
    g->share_b = (int)((int64_t)NXg * (c.rank + 1) / c.world) + (c.rank == c.world - 1 ? 1 : 0);
    rc = model_solve_device_bc(g, rtol, max_iter, it, rr);
    g->share_comm = nullptr;
    if (timing) {
        const auto t2 = std::chrono::steady_clock::now();
        t_gather += std::chrono::duration<double>(t1 - t0).count();
        t_solve += std::chrono::duration<double>(t2 - t1).count();
        if (++n_calls % 50 == 0 && c.rank == 0)
            fprintf(stderr, "[plf timing] %d replicated solves: gather %.3f s, BC + solve %.3f s\n", n_calls, t_gather, t_solve);
    }
    if (rc != PLF_OK) { ctx->err = g->err; return rc; }
    CK(cudaMemcpyAsync(ctx->d_du, g->d_du + 2 * (size_t)(NY + 1) * (size_t)c.first_col_global, (size_t)ctx->Ndof * sizeof(double),
                       cudaMemcpyDeviceToDevice, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    return PLF_OK;
}

// Model.bcval sums of the four edges, combined over the ranks of a partition
static int edge_values(plf_ctx *ctx, double edge[4][4])
{
    const GridView g = grid_view(ctx);
    LaunchCtx lc = make_lc(ctx);
    double *d_out = (double *)(ctx->d_scalars + 1536), *h = (double *)(ctx->h_scalars + 1536);
    CK(launch_edge_sums(lc, g.NX, g.NY, g.own0, g.own1, ctx->d_u, ctx->d_f, d_out, &ctx->launches));
    CK(cudaMemcpyAsync(h, d_out, 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!g.has_left) for (int q = 0; q < 4; q++) h[q] = 0.;
    if (!g.has_right) for (int q = 0; q < 4; q++) h[4 + q] = 0.;
    if (ctx->comm && ctx->comm->active()) {
        int rc = plf_comm_allreduce(ctx, h, 16);
        if (rc != PLF_OK) return rc;
    }
    const double cnt[4] = {(double)(g.NY + 1), (double)(g.NY + 1), (double)(g.NXg + 1), (double)(g.NXg + 1)};
    for (int e = 0; e < 4; e++) {
        edge[e][0] = h[4 * e] / cnt[e];
        edge[e][1] = h[4 * e + 1] / cnt[e];
        edge[e][2] = h[4 * e + 2];
        edge[e][3] = h[4 * e + 3];
    }
    return PLF_OK;
}

extern "C" int plf_boundary_values(plf_ctx *ctx, double edge[4][4])
{
    if (!ctx || !edge) return PLF_EINVAL;
    if (!ctx->mesh_ready) return fail(ctx, PLF_ESTATE, "plf_boundary_values: attributes for mesh not set");
    if (int rc = plf_bind(ctx)) return rc;
    return edge_values(ctx, edge);
}

extern "C" int plf_load_increment(plf_ctx *ctx, plf_load_state *st, int min_step, double rtol, int max_iter,
                                  plf_increment_report *rep)
{
    if (!ctx || !st || !rep) return PLF_EINVAL;
    if (!ctx->mesh_ready || !ctx->solve) return fail(ctx, PLF_ESTATE, "plf_load_increment: call plf_solve_begin first");
    plf_solve_state &S = *ctx->solve;
    memset(rep, 0, sizeof *rep);
    Channel right, top, node;
    for (int k = 0; k < 2; k++) {
        right.target[k] = S.bc.right[k]; right.reached[k] = st->reached_right[k];
        top.target[k] = S.bc.top[k]; top.reached[k] = st->reached_top[k];
        node.target[k] = S.bc.node[k]; node.reached[k] = st->reached_node[k];
    }
    node.limit_follows_step = true;
    const bool has_set = !S.noset.empty() || S.bc.n_noset > 0;
    const int il = st->il;
    right.plan(min_step, il);
    top.plan(min_step, il);
    node.plan(has_set ? min_step : 0, il);
    if (!has_set) node.step[0] = node.step[1] = node.limit[0] = node.limit[1] = 0.;

    int rc, it = 0;
    double rr = 0.;
    auto solve = [&](int slot) -> int {
        int r;
        if (ctx->replica) {
            r = solve_replicated(ctx, right.step, top.step, node.step, rtol, max_iter, &it, &rr, rep);
        } else {
            r = build_bc_vectors(ctx, right.step, top.step, node.step, rep);
            if (r != PLF_OK) return r;
            r = model_solve_device_bc(ctx, rtol, max_iter, &it, &rr);
        }
        rep->solves++;
        rep->cg_iterations += it;
        rep->pass_cg_iters[slot] = it;
        rep->max_relres = std::max(rep->max_relres, rr);
        return r;
    };
    // elastic predictor: the final solution of a linear model (1290-1291)
    if ((rc = solve(0)) != PLF_OK) return rc;
    rep->scale_bc = 1.;
    int nit = 0;
    if (S.nonlin) {
        if (il < 10 && (rc = plf_calc_scf(ctx, S.sld, &rep->scale_bc)) != PLF_OK) return rc;   // 1296
        right.scale(rep->scale_bc);
        top.scale(rep->scale_bc);
        bool change = true, conv = false;
        while ((change || !conv) && nit <= 15) {   // 1305
            if (il < 6 && nit > 1) {               // 1308
                right.halve();
                top.halve();
                if (has_set) node.halve();
            }
            if ((rc = plf_setupK(ctx)) != PLF_OK) return rc;   // 1333
            if ((rc = solve(1 + nit)) != PLF_OK) return rc;    // 1334-1335
            int32_t ch = 0, cv = 0;
            if ((rc = plf_sweep(ctx, nit, &ch, &cv, &rep->pass_max_hh[nit], rep->pass_counts[nit])) != PLF_OK) return rc;   // 1338-1361
            change = ch != 0;
            conv = cv != 0;
            rep->pass_change[nit] = ch;
            rep->pass_conv[nit] = cv;
            if (!conv) st->nconv++;
            nit++;
        }
    }
    rep->passes = nit;
    // commit (1383-1392): f += K du with the K of the last solve (still in the CSR memory), element state update
    if ((rc = plf_commit(ctx)) != PLF_OK) return rc;
    right.advance();
    top.advance();
    if (has_set) node.advance();
    st->il = il + 1;
    st->more = (right.pending() || top.pending() || (has_set && node.pending())) ? 1 : 0;
    for (int k = 0; k < 2; k++) {
        st->reached_right[k] = right.reached[k]; st->reached_top[k] = top.reached[k]; st->reached_node[k] = node.reached[k];
        rep->step_right[k] = right.step[k]; rep->step_top[k] = top.step[k]; rep->step_node[k] = node.step[k];
    }
    // calc_global (1415, 1473-1511)
    if ((rc = plf_reduce_global(ctx, rep->glob_sig, rep->glob_eps, rep->glob_epl)) != PLF_OK) return rc;
    return edge_values(ctx, rep->edge);
}
