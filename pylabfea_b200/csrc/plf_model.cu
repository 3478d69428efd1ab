// plf_model.cu -- C-ABI entry points for the device-resident model: mesh, element tangents, Kel, setupK -> CSR,
// the sweep / commit / calc_global / calc_scf steps of Model.solve, and the CG solve of the BC-reduced system.
// Host code only (compiled with -ffp-contract=off): the B matrices are computed here with the reference's
// operation order so they are bit-identical to Element.calc_Bmat (model.py:439-512).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "model_kernels.h"
#include "mg.h"
#include "plf_comm.h"
#include "plf_device.cuh"
#include "plf_host.h"
#include "plf_model.h"

using namespace plf;

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PLF_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

template <class T>
static int dalloc(plf_ctx *ctx, T **p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) return PLF_OK;
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    return PLF_OK;
}

void model_apply_mg_options(plf_ctx *ctx)
{
    if (!ctx->mg) return;
    ctx->mg->nu = std::max(2, ctx->mg_nu);
    ctx->mg->coarse_sweeps = std::max(1, ctx->mg_coarse_sweeps);
    ctx->mg->omega = ctx->mg_omega;
    ctx->mg->fuse_tail = ctx->mg_fuse_tail;
    cg_graph_reset(ctx->cg_graph);   // the captured V-cycle has the old sweep counts baked in
}

void model_free(plf_ctx *c)
{
    cg_graph_reset(c->cg_graph);
    if (c->mg) { mg_free(*c->mg); delete c->mg; c->mg = nullptr; }
    cudaFree(c->d_desc); cudaFree(c->d_elstiff); cudaFree(c->d_sig); cudaFree(c->d_eps); cudaFree(c->d_epl);
    cudaFree(c->d_res_sig); cudaFree(c->d_res_depl); cudaFree(c->d_fyld); cudaFree(c->d_nsteps); cudaFree(c->d_branch);
    cudaFree(c->d_has_res); cudaFree(c->d_u); cudaFree(c->d_du); cudaFree(c->d_f); cudaFree(c->d_rowptr); cudaFree(c->d_col);
    cudaFree(c->d_val); cudaFree(c->d_worklist); cudaFree(c->d_r); cudaFree(c->d_z); cudaFree(c->d_p); cudaFree(c->d_q);
    cudaFree(c->d_b); cudaFree(c->d_diag); cudaFree(c->d_w); cudaFree(c->d_fixed); cudaFree(c->d_partial);
    cudaFree(c->d_bc_du); cudaFree(c->d_bc_f); cudaFree(c->d_cgmask); c->d_cgmask = nullptr;
    if (c->comm) { c->comm->strip_set = false; c->comm->global_set = false; }   // a new mesh needs a new plf_comm_set_strip
    solve_state_free(c);
    c->tangent_epoch++; c->K_tangent_epoch = 0; c->mg_K_serial = -1; c->mg_mask_epoch = -1;
    c->d_bc_du = c->d_bc_f = nullptr; c->bc_valid = false;
    c->d_desc = nullptr; c->d_elstiff = c->d_sig = c->d_eps = c->d_epl = c->d_res_sig = c->d_res_depl = c->d_fyld = nullptr;
    c->d_nsteps = c->d_branch = nullptr; c->d_has_res = nullptr; c->d_u = c->d_du = c->d_f = nullptr;
    c->d_rowptr = c->d_col = nullptr; c->d_val = nullptr; c->d_worklist = nullptr;
    c->d_r = c->d_z = c->d_p = c->d_q = c->d_b = c->d_diag = c->d_w = nullptr; c->d_fixed = nullptr; c->d_partial = nullptr;
    c->mesh_ready = c->K_ready = false;
}

// ---------------------------------------------------------------------------------------------------------------
// geometry classes: Element.__init__ (model.py:316-348) + calc_Bmat (439-512)
// ---------------------------------------------------------------------------------------------------------------
static void make_geom(const DevMat &m, int planestress, double lx, double ly, double thick, GeomClass &g, double B4[4][48])
{
    memset(&g, 0, sizeof g);
    const double cpos = std::sqrt(1. / 3.);
    g.vel = lx * ly * thick;          // 316
    g.jacw = (g.vel * 4.) * 1.;       // 322, 340, 370: Jac * wght
    for (int i = 0; i < 4; i++) {
        const double sx = ((i / 2) % 2) ? -1. : 1.;   // (-1) ** int(i / 2)
        const double sy = (i % 2) ? -1. : 1.;         // (-1) ** i
        const double x = 0.5 * (1. + sx * cpos) * lx;
        const double y = 0.5 * (1. + sy * cpos) * ly;
        const double xi1 = 2. * x / lx - 1.;
        const double xi2 = 2. * y / ly - 1.;
        const double hxm = 0.125 * (1. - xi1) / ly;
        const double hym = 0.125 * (1. - xi2) / lx;
        const double hxp = 0.125 * (1. + xi1) / ly;
        const double hyp = 0.125 * (1. + xi2) / lx;
        const double p[4] = {-hym, -hyp, hym, hyp};
        const double q[4] = {-hxm, hxm, -hxp, hxp};
        double *B = B4[i];
        memset(B, 0, 48 * sizeof(double));
        for (int l = 0; l < 4; l++) {
            g.p[i][l] = p[l];
            g.q[i][l] = q[l];
            B[0 * 8 + 2 * l] = p[l];
            B[1 * 8 + 2 * l + 1] = q[l];
            B[5 * 8 + 2 * l] = q[l];
            B[5 * 8 + 2 * l + 1] = p[l];
        }
        if (planestress) {
            // hh = CV @ B (numpy matmul = ascending-k fma chain); B[2,:] = -nu*(hh[0,:]+hh[1,:])/E   (498-501)
            for (int j = 0; j < 8; j++) {
                double h0 = 0., h1 = 0.;
                for (int k = 0; k < 6; k++) {
                    h0 = std::fma(m.CV[k], B[8 * k + j], h0);
                    h1 = std::fma(m.CV[6 + k], B[8 * k + j], h1);
                }
                const double v = -m.nu * (h0 + h1) / m.E;
                g.r[i][j] = v;
                B[2 * 8 + j] = v;
            }
        }
    }
    // centroid operator: sum_g wght * B_g, accumulated in Gauss point order like Element.deps (395-398)
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 8; c++) {
            double s = 0.;
            for (int i = 0; i < 4; i++) s += B4[i][8 * r + c];
            g.bsum[r][c] = s;
        }
}

extern "C" int plf_mesh_structured(plf_ctx *ctx, int NX, int NY, const double *dx, double dy, double thick, const int32_t *elem_mat)
{
    if (!ctx) return PLF_EINVAL;
    if (ctx->nmat == 0) return fail(ctx, PLF_ESTATE, "plf_mesh_structured: call plf_set_materials first (Model.assign before Model.mesh)");
    if (NX < 1 || NY < 1 || !dx || !elem_mat || !(dy > 0.) || !(thick > 0.))
        return fail(ctx, PLF_EINVAL, "plf_mesh_structured: bad arguments");
    const int64_t Nel = (int64_t)NX * NY, Nnode = (int64_t)(NX + 1) * (NY + 1), Ndof = 2 * Nnode;
    const int64_t nnz = 4ll * (3ll * NX + 1) * (3ll * NY + 1);
    if (nnz > 2147483647ll) return fail(ctx, PLF_EUNSUPPORTED, "plf_mesh_structured: CSR index exceeds int32");
    CK(cudaSetDevice(ctx->device));
    model_free(ctx);
    ctx->NX = NX; ctx->NY = NY; ctx->Nel = Nel; ctx->Nnode = Nnode; ctx->Ndof = Ndof; ctx->nnz = nnz;
    ctx->dy = dy; ctx->thick = thick;
    ctx->dx.assign(dx, dx + NX);
    // geometry classes: unique (dx, material if plane stress)
    ctx->ngeom = 0;
    ctx->h_desc.resize((size_t)Nel);
    struct Key { double dx; int mat; };
    std::vector<Key> keys;
    double vol = 0.;
    for (int j = 0; j < NX; j++) {
        if (!(dx[j] > 0.)) return fail(ctx, PLF_EINVAL, "plf_mesh_structured: dx must be positive");
        for (int k = 0; k < NY; k++) {
            const int64_t e = (int64_t)j * NY + k;
            const int mid = elem_mat[e];
            if (mid < 0 || mid >= ctx->nmat) return fail(ctx, PLF_EINVAL, "plf_mesh_structured: material index out of range");
            const int kmat = ctx->planestress ? mid : 0;
            int cls = -1;
            for (size_t c = 0; c < keys.size(); c++)
                if (keys[c].dx == dx[j] && keys[c].mat == kmat) { cls = (int)c; break; }
            if (cls < 0) {
                if ((int)keys.size() >= PLF_MAX_GEOM)
                    return fail(ctx, PLF_EUNSUPPORTED, "plf_mesh_structured: more than " + std::to_string(PLF_MAX_GEOM) +
                                                           " distinct element geometries");
                cls = (int)keys.size();
                keys.push_back({dx[j], kmat});
                make_geom(ctx->hmats[kmat], ctx->planestress, dx[j], dy, thick, ctx->geom[cls], ctx->geom_B[cls]);
            }
            ctx->h_desc[(size_t)e] = (cls << 16) | mid;
            vol += ctx->geom[cls].vel;
        }
    }
    ctx->ngeom = (int)keys.size();
    ctx->volume = vol;
    int rc;
    const size_t E = (size_t)Nel, D = (size_t)Ndof;
    if ((rc = dalloc(ctx, &ctx->d_desc, E)) || (rc = dalloc(ctx, &ctx->d_elstiff, 36 * E)) || (rc = dalloc(ctx, &ctx->d_sig, 6 * E)) ||
        (rc = dalloc(ctx, &ctx->d_eps, 6 * E)) || (rc = dalloc(ctx, &ctx->d_epl, 6 * E)) || (rc = dalloc(ctx, &ctx->d_res_sig, 6 * E)) ||
        (rc = dalloc(ctx, &ctx->d_res_depl, 6 * E)) || (rc = dalloc(ctx, &ctx->d_fyld, E)) || (rc = dalloc(ctx, &ctx->d_nsteps, E)) ||
        (rc = dalloc(ctx, &ctx->d_has_res, E)) || (rc = dalloc(ctx, &ctx->d_worklist, E)) || (rc = dalloc(ctx, &ctx->d_u, D)) ||
        (rc = dalloc(ctx, &ctx->d_du, D)) || (rc = dalloc(ctx, &ctx->d_f, D)) || (rc = dalloc(ctx, &ctx->d_rowptr, D + 1)) ||
        (rc = dalloc(ctx, &ctx->d_col, (size_t)nnz)) || (rc = dalloc(ctx, &ctx->d_val, (size_t)nnz)))
        return rc;
    CK(cudaMemcpyAsync(ctx->d_desc, ctx->h_desc.data(), E * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->tables_dirty = true;   // new geometry classes: uploaded by the next plf_bind
    LaunchCtx lc = make_lc(ctx);
    CK(launch_csr_pattern(lc, NX, NY, ctx->d_rowptr, ctx->d_col, &ctx->launches));
    ctx->mg = new MgHierarchy();
    CK(mg_build(*ctx->mg, NX, NY));
    model_apply_mg_options(ctx);
    ctx->mesh_ready = true;
    ctx->K_ready = false;
    return model_reset_state(ctx);
}

// first Model.solve (model.py:1214-1227): u = f = 0, zero element state, elstiff = CV
int model_reset_state(plf_ctx *ctx)
{
    const size_t E = (size_t)ctx->Nel, D = (size_t)ctx->Ndof;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_sig, 0, 6 * E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_eps, 0, 6 * E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_epl, 0, 6 * E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_res_sig, 0, 6 * E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_res_depl, 0, 6 * E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_fyld, 0, E * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_nsteps, 0, E * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_has_res, 0, E, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_u, 0, D * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_du, 0, D * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_f, 0, D * sizeof(double), ctx->stream));
    return plf_set_tangent(ctx, nullptr);
}

#define NEED_MESH(name) \
    if (!ctx) return PLF_EINVAL; \
    if (!ctx->mesh_ready) return fail(ctx, PLF_ESTATE, std::string(name) + ": attributes for mesh not set (call plf_mesh_structured)"); \
    if (int brc_ = plf_bind(ctx)) return brc_

extern "C" int plf_mesh_sizes(const plf_ctx *ctx, int64_t *Nel, int64_t *Nnode, int64_t *Ndof, int64_t *nnz)
{
    if (!ctx || !ctx->mesh_ready) return PLF_ESTATE;
    if (Nel) *Nel = ctx->Nel;
    if (Nnode) *Nnode = ctx->Nnode;
    if (Ndof) *Ndof = ctx->Ndof;
    if (nnz) *nnz = ctx->nnz;
    return PLF_OK;
}

extern "C" int plf_get_bmat(const plf_ctx *ctx, int64_t e, double *B4)
{
    if (!ctx || !ctx->mesh_ready || !B4 || e < 0 || e >= ctx->Nel) return PLF_EINVAL;
    memcpy(B4, ctx->geom_B[ctx->h_desc[(size_t)e] >> 16], 4 * 48 * sizeof(double));
    return PLF_OK;
}

extern "C" int plf_set_tangent(plf_ctx *ctx, const double *elstiff)
{
    NEED_MESH("plf_set_tangent");
    CK(cudaSetDevice(ctx->device));
    const size_t E = (size_t)ctx->Nel;
    if (elstiff) {
        CK(cudaMemcpyAsync(ctx->d_elstiff, elstiff, 36 * E * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        std::vector<double> h(36 * E);
        for (size_t e = 0; e < E; e++) memcpy(&h[36 * e], ctx->hmats[ctx->h_desc[e] & 0xffff].CV, 36 * sizeof(double));
        CK(cudaMemcpyAsync(ctx->d_elstiff, h.data(), 36 * E * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->K_ready = false;
    ctx->tangent_epoch++;
    return PLF_OK;
}

extern "C" int plf_get_tangent(plf_ctx *ctx, double *elstiff)
{
    NEED_MESH("plf_get_tangent");
    if (!elstiff) return PLF_EINVAL;
    CK(cudaMemcpyAsync(elstiff, ctx->d_elstiff, 36 * (size_t)ctx->Nel * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_get_kel(plf_ctx *ctx, double *Kel)
{
    NEED_MESH("plf_get_kel");
    if (!Kel) return PLF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, 64 * (size_t)ctx->Nel * sizeof(double)));
    LaunchCtx lc = make_lc(ctx);
    cudaError_t e = launch_kel(lc, ctx->Nel, ctx->planestress, ctx->d_desc, ctx->d_elstiff, d, &ctx->launches);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Kel, d, 64 * (size_t)ctx->Nel * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CK(e);
    return PLF_OK;
}

extern "C" int plf_setupK(plf_ctx *ctx)
{
    NEED_MESH("plf_setupK");
    // no element tangent changed since this K was assembled from them: the CSR values are already the result
    if (ctx->K_ready && ctx->K_tangent_epoch == ctx->tangent_epoch && ctx->skip_unchanged) return PLF_OK;
    LaunchCtx lc = make_lc(ctx);
    CK(launch_assemble(lc, ctx->NX, ctx->NY, ctx->planestress, ctx->ngeom, ctx->d_desc, ctx->d_elstiff, ctx->d_rowptr, ctx->d_val, &ctx->launches));
    ctx->K_ready = true;
    ctx->K_tangent_epoch = ctx->tangent_epoch;
    ctx->K_serial++;
    return PLF_OK;
}

extern "C" int plf_setupK_from_batch(plf_ctx *ctx, const plf_batch *b)
{
    NEED_MESH("plf_setupK_from_batch");
    if (!b || b->n != ctx->Nel) return fail(ctx, PLF_EINVAL, "plf_setupK_from_batch: batch must hold exactly Nel elements");
    LaunchCtx lc = make_lc(ctx);
    CK(launch_assemble(lc, ctx->NX, ctx->NY, ctx->planestress, ctx->ngeom, ctx->d_desc, b->ctan, ctx->d_rowptr, ctx->d_val, &ctx->launches));
    ctx->K_ready = true;
    ctx->K_tangent_epoch = 0;   // not assembled from Element.elstiff
    ctx->K_serial++;
    return PLF_OK;
}

// material ids of the mesh elements (descriptor low half) into the batch, once: plf_step_host on mixed meshes
static int batch_materials_from_mesh(plf_ctx *ctx, plf_batch *b)
{
    if (b->mat_from_mesh) return PLF_OK;
    if (ctx->nmat == 1) {
        CK(cudaMemsetAsync(b->mat_id, 0, (size_t)b->n * sizeof(int32_t), ctx->stream));
    } else {
        std::vector<int32_t> ids((size_t)b->n);
        for (size_t e = 0; e < ids.size(); e++) ids[e] = ctx->h_desc[e] & 0xffff;
        CK(cudaMemcpyAsync(b->mat_id, ids.data(), ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    b->mat_from_mesh = true;
    return PLF_OK;
}

// One inner pass of the load-step loop with HOST buffers (model.py:1338-1361 + setupK): element states in, response
// results and CSR(K) values out.  The mesh is cut into chunks of element columns; chunk c is uploaded on one copy
// stream while chunk c-1 is updated and its node columns are assembled on the compute stream and chunk c-2 drains
// to the host on a second copy stream, so that the PCIe link is busy in both directions for the whole call
// (host buffers should be pinned, plf_host_alloc).  Any output pointer may be NULL.
extern "C" int plf_step_host(plf_ctx *ctx, plf_batch *b, const double *sig, const double *epl, const double *deps, double *fy,
                             double *sig_out, double *depl, double *ctan, int32_t *nsteps, double *kval)
{
    NEED_MESH("plf_step_host");
    if (!b || b->n != ctx->Nel || !sig || !epl || !deps)
        return fail(ctx, PLF_EINVAL, "plf_step_host: batch must hold exactly Nel elements; sig, epl, deps are required");
    if (ctx->nmat == 0) return fail(ctx, PLF_ESTATE, "plf_step_host: no materials set");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    }
    const int NX = ctx->NX, NY = ctx->NY;
    const int nchunk = std::max(1, std::min(NX, 8));
    while ((int)ctx->ev_step.size() < 2 * nchunk + 1) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_step.push_back(e);
    }
    // everything queued earlier on the compute stream (previous steps reading / writing the batch) comes first
    CK(cudaEventRecord(ctx->ev_step[2 * nchunk], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_step[2 * nchunk], 0));
    CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_step[2 * nchunk], 0));
    {
        int rcm = batch_materials_from_mesh(ctx, b);
        if (rcm != PLF_OK) return rcm;
    }
    LaunchCtx lc = make_lc(ctx);
    for (int c = 0; c < nchunk; c++) {
        const int ja = (int)((int64_t)NX * c / nchunk), jb = (int)((int64_t)NX * (c + 1) / nchunk);   // element columns
        const int64_t e0 = (int64_t)ja * NY, ne = (int64_t)(jb - ja) * NY;
        const size_t o6 = 6 * (size_t)e0, n6 = 6 * (size_t)ne * sizeof(double);
        CK(cudaMemcpyAsync(b->sig + o6, sig + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(b->epl + o6, epl + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(b->deps + o6, deps + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->ev_step[2 * c], ctx->s_in));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_step[2 * c], 0));
        CK(launch_response(lc, ctx->hmats, ctx->nmat, ne, b->mat_id + e0, b->sig + o6, b->epl + o6, b->deps + o6, b->fy + e0,
                           b->sig_out + o6, b->depl + o6, b->ctan + 36 * (size_t)e0, b->nsteps + e0, b->worklist + e0,
                           batch_counters(ctx), batch_queue(ctx), ctx->compact,
                           &ctx->launches));
        // node columns whose elements are all updated now: ja .. jb-1 (column jb needs the next chunk), plus NX at the end
        const int na = ja, nb = (c == nchunk - 1) ? NX + 1 : jb;
        CK(launch_assemble_columns(lc, NX, NY, ctx->planestress, ctx->ngeom, na, nb, ctx->d_desc, b->ctan, ctx->d_val,
                                   &ctx->launches));
        CK(cudaEventRecord(ctx->ev_step[2 * c + 1], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_step[2 * c + 1], 0));
        if (fy) CK(cudaMemcpyAsync(fy + e0, b->fy + e0, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
        if (nsteps) CK(cudaMemcpyAsync(nsteps + e0, b->nsteps + e0, (size_t)ne * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out));
        if (sig_out) CK(cudaMemcpyAsync(sig_out + o6, b->sig_out + o6, n6, cudaMemcpyDeviceToHost, ctx->s_out));
        if (depl) CK(cudaMemcpyAsync(depl + o6, b->depl + o6, n6, cudaMemcpyDeviceToHost, ctx->s_out));
        if (ctan) CK(cudaMemcpyAsync(ctan + 36 * (size_t)e0, b->ctan + 36 * (size_t)e0, 36 * (size_t)ne * sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->s_out));
        if (kval) {
            const int64_t v0 = csr_column_start(NX, NY, na), v1 = csr_column_start(NX, NY, nb);
            CK(cudaMemcpyAsync(kval + v0, ctx->d_val + v0, (size_t)(v1 - v0) * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
        }
    }
    ctx->K_ready = true;
    ctx->K_tangent_epoch = 0;
    ctx->K_serial++;
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

// plf_step_host with compact results (include/plfb200.h).  Same three-stream pipeline over chunks of element columns; all
// uploads and all kernel work are enqueued first, then the host follows the chunks as they finish: the number of packed
// tangents of a chunk (device-side scan) decides the size of its download.
extern "C" int plf_step_host_packed(plf_ctx *ctx, plf_batch *b, const double *sig, const double *epl, const double *deps, double *fy,
                                    double *sig_out, double *depl, int32_t *nsteps, double *ct21, uint32_t *elastic_words,
                                    int64_t *n_packed, double *kval)
{
    NEED_MESH("plf_step_host_packed");
    if (!b || b->n != ctx->Nel || !sig || !epl || !deps || !ct21 || !elastic_words || !n_packed)
        return fail(ctx, PLF_EINVAL, "plf_step_host_packed: batch must hold exactly Nel elements; sig, epl, deps, ct21, elastic_words, "
                                     "n_packed are required");
    if (ctx->nmat == 0) return fail(ctx, PLF_ESTATE, "plf_step_host_packed: no materials set");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    }
    const int NX = ctx->NX, NY = ctx->NY;
    // chunks of element columns must start on a bitmap word (32 elements): the largest chunk count <= 8 that does, else one
    int nchunk = 1;
    for (int c = std::min(NX, 8); c > 1; c--) {
        bool ok = true;
        for (int q = 1; q < c && ok; q++) ok = (((int64_t)NX * q / c) * NY) % 32 == 0;
        if (ok) { nchunk = c; break; }
    }
    while ((int)ctx->ev_step.size() < 2 * nchunk + 1) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_step.push_back(e);
    }
    const size_t N = (size_t)b->n;
    if (!b->ct21) {
        const size_t nblk = (N + 255) / 256 + 8;
        CK(cudaMalloc((void **)&b->ct21, 21 * N * sizeof(double)));
        CK(cudaMalloc((void **)&b->el_words, ((N + 31) / 32 + 8) * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&b->pk_counts, nblk * sizeof(int32_t)));
        CK(cudaMalloc((void **)&b->pk_offsets, nblk * sizeof(int32_t)));
        CK(cudaMalloc((void **)&b->pk_totals, 64 * sizeof(int64_t)));
        CK(cudaMallocHost((void **)&b->h_pk_totals, 64 * sizeof(int64_t)));
    }
    CK(cudaEventRecord(ctx->ev_step[2 * nchunk], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_step[2 * nchunk], 0));
    CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_step[2 * nchunk], 0));
    int rc = batch_materials_from_mesh(ctx, b);
    if (rc != PLF_OK) return rc;
    LaunchCtx lc = make_lc(ctx);
    struct Chunk { int ja, jb, na, nb; int64_t e0, ne; };
    std::vector<Chunk> ch(nchunk);
    for (int c = 0; c < nchunk; c++) {
        Chunk &k = ch[c];
        k.ja = (int)((int64_t)NX * c / nchunk);
        k.jb = (int)((int64_t)NX * (c + 1) / nchunk);
        k.e0 = (int64_t)k.ja * NY;
        k.ne = (int64_t)(k.jb - k.ja) * NY;
        k.na = k.ja;
        k.nb = (c == nchunk - 1) ? NX + 1 : k.jb;
    }
    for (int c = 0; c < nchunk; c++) {   // all uploads
        const Chunk &k = ch[c];
        const size_t o6 = 6 * (size_t)k.e0, n6 = 6 * (size_t)k.ne * sizeof(double);
        CK(cudaMemcpyAsync(b->sig + o6, sig + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(b->epl + o6, epl + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(b->deps + o6, deps + o6, n6, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->ev_step[2 * c], ctx->s_in));
    }
    const bool single = ctx->nmat == 1;
    for (int c = 0; c < nchunk; c++) {   // all kernel work
        const Chunk &k = ch[c];
        const size_t o6 = 6 * (size_t)k.e0;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_step[2 * c], 0));
        CK(launch_response(lc, ctx->hmats, ctx->nmat, k.ne, b->mat_id + k.e0, b->sig + o6, b->epl + o6, b->deps + o6, b->fy + k.e0,
                           b->sig_out + o6, b->depl + o6, b->ctan + 36 * (size_t)k.e0, b->nsteps + k.e0, b->worklist + k.e0,
                           batch_counters(ctx), batch_queue(ctx), ctx->compact,
                           &ctx->launches));
        CK(launch_pack(lc, k.ne, b->mat_id + k.e0, single ? 1 : 0, b->ctan + 36 * (size_t)k.e0, b->el_words + (k.e0 >> 5),
                       b->pk_counts + (k.e0 >> 8) + c, b->pk_offsets + (k.e0 >> 8) + c, b->pk_totals + c, b->ct21 + 21 * (size_t)k.e0,
                       &ctx->launches));
        CK(cudaMemcpyAsync(b->h_pk_totals + c, b->pk_totals + c, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(launch_assemble_columns(lc, NX, NY, ctx->planestress, ctx->ngeom, k.na, k.nb, ctx->d_desc, b->ctan, ctx->d_val,
                                   &ctx->launches));
        CK(cudaEventRecord(ctx->ev_step[2 * c + 1], ctx->stream));
    }
    ctx->K_ready = true;
    ctx->K_tangent_epoch = 0;
    ctx->K_serial++;
    int64_t packed = 0;
    for (int c = 0; c < nchunk; c++) {   // downloads, in the order the chunks finish
        const Chunk &k = ch[c];
        const size_t o6 = 6 * (size_t)k.e0, n6 = 6 * (size_t)k.ne * sizeof(double);
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_step[2 * c + 1], 0));
        if (fy) CK(cudaMemcpyAsync(fy + k.e0, b->fy + k.e0, (size_t)k.ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
        if (nsteps) CK(cudaMemcpyAsync(nsteps + k.e0, b->nsteps + k.e0, (size_t)k.ne * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_out));
        if (sig_out) CK(cudaMemcpyAsync(sig_out + o6, b->sig_out + o6, n6, cudaMemcpyDeviceToHost, ctx->s_out));
        if (depl) CK(cudaMemcpyAsync(depl + o6, b->depl + o6, n6, cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaMemcpyAsync(elastic_words + (k.e0 >> 5), b->el_words + (k.e0 >> 5), (size_t)((k.ne + 31) / 32) * sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, ctx->s_out));
        if (kval) {
            const int64_t v0 = csr_column_start(NX, NY, k.na), v1 = csr_column_start(NX, NY, k.nb);
            CK(cudaMemcpyAsync(kval + v0, ctx->d_val + v0, (size_t)(v1 - v0) * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_out));
        }
        CK(cudaEventSynchronize(ctx->ev_step[2 * c + 1]));   // the chunk's count is in pinned memory now
        const int64_t cnt = b->h_pk_totals[c];
        if (cnt > 0)
            CK(cudaMemcpyAsync(ct21 + 21 * (size_t)packed, b->ct21 + 21 * (size_t)k.e0, 21 * (size_t)cnt * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->s_out));
        packed += cnt;
    }
    *n_packed = packed;
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_get_csr(plf_ctx *ctx, int32_t *rowptr, int32_t *col, double *val)
{
    NEED_MESH("plf_get_csr");
    if (val && !ctx->K_ready) return fail(ctx, PLF_ESTATE, "plf_get_csr: call plf_setupK first");
    if (rowptr) CK(cudaMemcpyAsync(rowptr, ctx->d_rowptr, ((size_t)ctx->Ndof + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (col) CK(cudaMemcpyAsync(col, ctx->d_col, (size_t)ctx->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (val) CK(cudaMemcpyAsync(val, ctx->d_val, (size_t)ctx->nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// fields
// ---------------------------------------------------------------------------------------------------------------
static int field_ptr(plf_ctx *ctx, int field, void **p, size_t *bytes)
{
    const size_t E = (size_t)ctx->Nel, D = (size_t)ctx->Ndof;
    switch (field) {
    case PLF_SIG: *p = ctx->d_sig; *bytes = 6 * E * 8; break;
    case PLF_EPS: *p = ctx->d_eps; *bytes = 6 * E * 8; break;
    case PLF_EPL: *p = ctx->d_epl; *bytes = 6 * E * 8; break;
    case PLF_RES_SIG: *p = ctx->d_res_sig; *bytes = 6 * E * 8; break;
    case PLF_RES_DEPL: *p = ctx->d_res_depl; *bytes = 6 * E * 8; break;
    case PLF_U: *p = ctx->d_u; *bytes = D * 8; break;
    case PLF_DU: *p = ctx->d_du; *bytes = D * 8; break;
    case PLF_F: *p = ctx->d_f; *bytes = D * 8; break;
    case PLF_FYLD: *p = ctx->d_fyld; *bytes = E * 8; break;
    case PLF_NSTEPS: *p = ctx->d_nsteps; *bytes = E * 4; break;
    default: return fail(ctx, PLF_EINVAL, "unknown field id");
    }
    return PLF_OK;
}

extern "C" int plf_get_field(plf_ctx *ctx, int field, void *host)
{
    NEED_MESH("plf_get_field");
    if (!host) return PLF_EINVAL;
    void *p; size_t bytes;
    int rc = field_ptr(ctx, field, &p, &bytes);
    if (rc != PLF_OK) return rc;
    CK(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_set_field(plf_ctx *ctx, int field, const void *host)
{
    NEED_MESH("plf_set_field");
    if (!host) return PLF_EINVAL;
    void *p; size_t bytes;
    int rc = field_ptr(ctx, field, &p, &bytes);
    if (rc != PLF_OK) return rc;
    CK(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (field == PLF_RES_SIG) CK(cudaMemsetAsync(ctx->d_has_res, 1, (size_t)ctx->Nel, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// non-linear loop pieces
// ---------------------------------------------------------------------------------------------------------------
static int ensure_partials(plf_ctx *ctx)
{
    const int need = ctx->sm_count * 8 * 18 + 64;
    const int64_t spmv_blocks = (ctx->Ndof * 8 + 255) / 256 + 1;
    const int64_t want = std::max<int64_t>(need, spmv_blocks * 2);
    if (ctx->d_partial && ctx->n_partial >= want) return PLF_OK;
    int rc = dalloc(ctx, &ctx->d_partial, (size_t)want);
    if (rc != PLF_OK) return rc;
    ctx->n_partial = (int)want;
    return PLF_OK;
}

// layout of the 4 KB scalar scratch: [0,256) SweepStats | [256,512) CgScalars | [512,516) reduction counter |
// [576,1024) small vectors (sld, global sums, scf moments)
// strip partition with its position in the global mesh set (plf_comm_set_partition): one rank of a distributed Model.solve
static inline bool is_partitioned(const plf_ctx *ctx) { return ctx->comm && ctx->comm->active() && ctx->comm->global_set; }
// elements of the halo column (updated here for the K rows of the first owned node column, owned and counted by the left rank)
static inline int64_t halo_elements(const plf_ctx *ctx) { return is_partitioned(ctx) ? (int64_t)ctx->comm->own_first * ctx->NY : 0; }

static inline unsigned int *counter_ptr(plf_ctx *ctx) { return (unsigned int *)(ctx->d_scalars + 512); }

extern "C" int plf_sweep(plf_ctx *ctx, int nit, int32_t *change, int32_t *conv, double *max_hh, int64_t counts[4])
{
    NEED_MESH("plf_sweep");
    for (int i = 0; i < ctx->nmat; i++)
        if (ctx->hmats[i].tresca && (ctx->hmats[i].kind == PLF_MAT_HILL || ctx->hmats[i].kind == PLF_MAT_HILL3))
            // material.py:824-825: the first plastic step of a Tresca material raises in the reference
            return fail(ctx, PLF_EUNSUPPORTED, "plf_sweep: analytical gradient for Tresca not implemented");
    LaunchCtx lc = make_lc(ctx);
    CK(launch_sweep(lc, ctx->hmats, ctx->nmat, ctx->Nel, ctx->NY, nit, ctx->d_desc, ctx->d_sig, ctx->d_epl, ctx->d_du, ctx->d_elstiff,
                    ctx->d_res_sig, ctx->d_res_depl, ctx->d_fyld, ctx->d_nsteps, ctx->d_has_res, ctx->d_worklist, ctx->d_scalars,
                    ctx->h_scalars, ctx->compact, &ctx->launches, halo_elements(ctx)));
    const SweepStats *st = (const SweepStats *)ctx->h_scalars;
    if (st->change) ctx->tangent_epoch++;   // local tangents (halo column included) changed
    double v[6] = {st->change ? 1. : 0., (double)st->nonconv, (double)st->counts[0], (double)st->counts[1], (double)st->counts[2],
                   (double)st->counts[3]};
    double mh = st->max_hh;
    if (is_partitioned(ctx)) {   // every rank must take the same decisions (model.py:1305, 1355-1361)
        int rc = comm_host_allreduce(ctx, v, 6, 0);
        if (rc == PLF_OK) rc = comm_host_allreduce(ctx, &mh, 1, 1);
        if (rc != PLF_OK) return rc;
    }
    if (change) *change = v[0] > 0. ? 1 : 0;
    if (conv) *conv = v[1] == 0. ? 1 : 0;
    if (max_hh) *max_hh = mh;
    if (counts) for (int i = 0; i < 4; i++) counts[i] = (int64_t)v[2 + i];
    // K_ready stays set: the CSR memory still holds the stiffness the current increment was solved with, which is
    // what plf_commit must use for f += K du (model.py:1384 uses the K assembled before the last sweep).
    return PLF_OK;
}

extern "C" int plf_commit(plf_ctx *ctx)
{
    NEED_MESH("plf_commit");
    if (!ctx->K_ready) return fail(ctx, PLF_ESTATE, "plf_commit: the stiffness the increment was solved with is not assembled");
    LaunchCtx lc = make_lc(ctx);
    CK(launch_commit(lc, ctx->Nel, ctx->Ndof, ctx->NY, ctx->d_desc, ctx->d_u, ctx->d_du, ctx->d_f, ctx->d_rowptr, ctx->d_col, ctx->d_val,
                     ctx->d_elstiff, ctx->d_sig, ctx->d_eps, ctx->d_epl, ctx->d_res_sig, ctx->d_res_depl, ctx->d_has_res,
                     &ctx->launches));
    if (ctx->comm && ctx->comm->active() && ctx->comm->strip_set &&
        comm_halo_exchange(*ctx->comm, ctx->d_f, ctx->NX, ctx->NY, ctx->stream))   // rows of ghost columns are not complete here
        return fail(ctx, PLF_ECUDA, std::string("plf_commit: ") + comm_error());
    return PLF_OK;
}

extern "C" int plf_reduce_global(plf_ctx *ctx, double *sig6, double *eps6, double *epl6)
{
    NEED_MESH("plf_reduce_global");
    int rc = ensure_partials(ctx);
    if (rc != PLF_OK) return rc;
    LaunchCtx lc = make_lc(ctx);
    double *d_out = (double *)(ctx->d_scalars + 576);
    CK(launch_global_sums(lc, ctx->Nel, ctx->d_desc, ctx->d_sig, ctx->d_eps, ctx->d_epl, ctx->d_partial, counter_ptr(ctx), d_out,
                          &ctx->launches, halo_elements(ctx)));
    double *h = (double *)(ctx->h_scalars + 576);
    CK(cudaMemcpyAsync(h, d_out, 18 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // model.py:1508-1511: divide by the model volume lenx*leny*thick
    double lenx = 0.;
    for (double v : ctx->dx) lenx += v;
    if (is_partitioned(ctx)) {
        rc = comm_host_allreduce(ctx, h, 18, 0);
        if (rc != PLF_OK) return rc;
        lenx = ctx->comm->lenx_global;
    }
    const double Vm = lenx * (ctx->dy * ctx->NY) * ctx->thick;
    for (int k = 0; k < 6; k++) {
        if (sig6) sig6[k] = h[k] / Vm;
        if (eps6) eps6[k] = h[6 + k] / Vm;
        if (epl6) epl6[k] = h[12 + k] / Vm;
    }
    return PLF_OK;
}

extern "C" int plf_calc_scf(plf_ctx *ctx, const double *sld, double *scf)
{
    NEED_MESH("plf_calc_scf");
    if (!sld || !scf) return PLF_EINVAL;
    int rc = ensure_partials(ctx);
    if (rc != PLF_OK) return rc;
    LaunchCtx lc = make_lc(ctx);
    double *d_sld = (double *)(ctx->d_scalars + 576);
    double *d_out = (double *)(ctx->d_scalars + 640);
    CK(cudaMemcpyAsync(d_sld, sld, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // hval / mult reuse the fyld array and the worklist bytes (both are rewritten by the next sweep)
    CK(launch_scf(lc, ctx->hmats, ctx->nmat, ctx->Nel, ctx->NY, ctx->d_desc, ctx->d_sig, ctx->d_epl, ctx->d_du, ctx->d_elstiff, d_sld,
                  ctx->d_fyld, (unsigned char *)ctx->d_worklist, ctx->d_partial, counter_ptr(ctx), d_out,
                  (double *)(ctx->h_scalars + 640), scf, &ctx->launches, halo_elements(ctx), is_partitioned(ctx) ? ctx->comm : nullptr));
    return PLF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// linear solve
// ---------------------------------------------------------------------------------------------------------------
int model_solver_buffers(plf_ctx *ctx)
{
    const size_t D = (size_t)ctx->Ndof;
    int rc;
    if (!ctx->d_r) {
        if ((rc = dalloc(ctx, &ctx->d_r, D)) || (rc = dalloc(ctx, &ctx->d_z, D)) || (rc = dalloc(ctx, &ctx->d_p, D)) ||
            (rc = dalloc(ctx, &ctx->d_q, D)) || (rc = dalloc(ctx, &ctx->d_b, D)) || (rc = dalloc(ctx, &ctx->d_diag, D)) ||
            (rc = dalloc(ctx, &ctx->d_w, D)) || (rc = dalloc(ctx, &ctx->d_fixed, D)) || (rc = dalloc(ctx, &ctx->d_bc_du, D)) ||
            (rc = dalloc(ctx, &ctx->d_bc_f, D)))
            return rc;
    }
    return ensure_partials(ctx);
}

extern "C" int plf_solve_linear(plf_ctx *ctx, const uint8_t *fixed, const double *du_bc, const double *w, const double *f_ext,
                                double rtol, int max_iter, int *iters, double *relres)
{
    NEED_MESH("plf_solve_linear");
    const bool reuse = !fixed && !du_bc && !w && !f_ext;   // all NULL: the boundary-condition vectors of the previous call
    if (!reuse && (!fixed || !du_bc || !w || !f_ext)) return fail(ctx, PLF_EINVAL, "plf_solve_linear: NULL argument");
    if (reuse && !ctx->bc_valid) return fail(ctx, PLF_ESTATE, "plf_solve_linear: no boundary-condition vectors to re-use");
    if (!ctx->K_ready) return fail(ctx, PLF_ESTATE, "plf_solve_linear: call plf_setupK first");
    CK(cudaSetDevice(ctx->device));
    const size_t D = (size_t)ctx->Ndof;
    int rc;
    if ((rc = model_solver_buffers(ctx)) != PLF_OK) return rc;
    if (!reuse) {
        // the BC vectors stay on the device (d_fixed, d_w, d_bc_f, d_bc_du): the Newton-like passes of an increment call
        // again with the same boundary conditions and then skip these 4 host-to-device copies
        CK(cudaMemcpyAsync(ctx->d_fixed, fixed, D, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_w, w, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_bc_f, f_ext, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_bc_du, du_bc, D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->bc_valid = true;
        ctx->mask_epoch++;
        solve_state_invalidate_bc(ctx);
    }
    return model_solve_device_bc(ctx, rtol, max_iter, iters, relres);
}

int model_solve_device_bc(plf_ctx *ctx, double rtol, int max_iter, int *iters, double *relres)
{
    if (!ctx->K_ready) return fail(ctx, PLF_ESTATE, "plf_solve_linear: call plf_setupK first");
    if (!ctx->bc_valid) return fail(ctx, PLF_ESTATE, "plf_solve_linear: no boundary-condition vectors on the device");
    if (int brc = plf_bind(ctx)) return brc;
    int rc;
    if ((rc = model_solver_buffers(ctx)) != PLF_OK) return rc;
    const size_t D = (size_t)ctx->Ndof;
    LaunchCtx lc = make_lc(ctx);
    const bool dist = ctx->comm && ctx->comm->active();
    if (dist && !ctx->comm->strip_set) return fail(ctx, PLF_ESTATE, "plf_solve_linear: call plf_comm_set_strip after plf_mesh_structured");
    const unsigned char *mask = ctx->d_fixed;
    if (dist) {
        // rows of node columns owned by a neighbour are not this rank's equations (and incomplete here): masked like fixed dofs
        if (!ctx->d_cgmask && (rc = dalloc(ctx, &ctx->d_cgmask, D))) return rc;
        CK(launch_ghost_mask(lc, ctx->Ndof, ctx->NY, ctx->comm->own_first, ctx->comm->own_end, ctx->d_fixed, ctx->d_cgmask, &ctx->launches));
        mask = ctx->d_cgmask;
    }
    // b = f_ext - K w on the free dofs (calc_BC, model.py:1099-1205)
    CK(launch_spmv(lc, ctx->Ndof, ctx->d_rowptr, ctx->d_col, ctx->d_val, ctx->d_w, ctx->d_bc_f, ctx->d_b, mask, 2, &ctx->launches));
    int it = 0;
    double rr = 0.;
    double *x = ctx->d_du;   // CG iterate lives in du; fixed dofs are merged at the end
    if (dist) {
        const bool dmg = ctx->solver_mg && ctx->comm->mg_all;   // the same on every rank
        const bool fresh = dmg && ctx->skip_unchanged && ctx->mg_K_serial == ctx->K_serial && ctx->mg_mask_epoch == ctx->mask_epoch;
        if (dmg) { ctx->mg_K_serial = ctx->K_serial; ctx->mg_mask_epoch = ctx->mask_epoch; }
        cudaError_t ce = launch_cg_dist(lc, *ctx->comm, ctx->Ndof, ctx->d_rowptr, ctx->d_col, ctx->d_val, mask, ctx->d_b, x, ctx->d_r,
                                        ctx->d_z, ctx->d_p, ctx->d_q, ctx->d_diag, ctx->d_partial, counter_ptr(ctx), ctx->d_scalars + 256,
                                        ctx->h_scalars + 256, rtol, max_iter, dmg ? 4 : 50, &it, &rr, &ctx->launches, ctx->solver_warm,
                                        ctx->NX, ctx->NY, dmg ? ctx->mg : nullptr, fresh ? 0 : 1);
        if (ce != cudaSuccess)
            return fail(ctx, PLF_ECUDA, std::string("plf_solve_linear (distributed): ") +
                                            (ce == cudaErrorUnknown && comm_error()[0] ? comm_error() : cudaGetErrorString(ce)));
        CK(launch_merge_du(lc, ctx->Ndof, x, ctx->d_bc_du, ctx->d_fixed, ctx->d_du, &ctx->launches));
        if (comm_halo_exchange(*ctx->comm, ctx->d_du, ctx->NX, ctx->NY, ctx->stream))   // ghost columns of du from their owners
            return fail(ctx, PLF_ECUDA, std::string("plf_solve_linear (distributed): ") + comm_error());
        CK(cudaStreamSynchronize(ctx->stream));
        if (iters) *iters = it;
        if (relres) *relres = rr;
        if (!(rr <= rtol))
            return fail(ctx, PLF_ENOCONV, "plf_solve_linear: distributed CG stopped after " + std::to_string(it) +
                                              " iterations with relative residual " + std::to_string(rr));
        return PLF_OK;
    }
    // multigrid V-cycle preconditioner when the mesh coarsens at least once (mesh-independent iteration counts), else Jacobi
    const bool use_mg = ctx->solver_mg && ctx->mg && ctx->mg->levels() >= 2;
    if (ctx->share_comm && use_mg) {
        // replicated solve, level 0 and the CG vectors shared among the ranks (launch_cg_hybrid)
        Comm &c = *ctx->share_comm;
        const int a = ctx->share_a, bcol = ctx->share_b;
        if (!ctx->d_cgmask && (rc = dalloc(ctx, &ctx->d_cgmask, D))) return rc;
        CK(launch_ghost_mask(lc, ctx->Ndof, ctx->NY, a, bcol, ctx->d_fixed, ctx->d_cgmask, &ctx->launches));
        MgDist md;
        md.comm = &c;
        mg_plan_sharing(*ctx->mg, c.world, c.rank, md);
        if (md.a[0] != a || md.b[0] != bcol) return fail(ctx, PLF_ESTATE, "plf_solve_linear (shared levels): column ranges disagree");
        const bool fresh = ctx->skip_unchanged && ctx->mg_K_serial == ctx->K_serial && ctx->mg_mask_epoch == ctx->mask_epoch;
        ctx->mg_K_serial = ctx->K_serial; ctx->mg_mask_epoch = ctx->mask_epoch;
        cudaError_t ce = launch_cg_hybrid(lc, c, a, bcol, ctx->Ndof, ctx->d_rowptr, ctx->d_col, ctx->d_val, ctx->d_cgmask, ctx->d_b, x, ctx->d_r,
                                          ctx->d_z, ctx->d_p, ctx->d_q, ctx->d_diag, ctx->d_partial, counter_ptr(ctx), ctx->d_scalars + 256,
                                          ctx->h_scalars + 256, rtol, max_iter, &it, &rr, &ctx->launches, ctx->solver_warm, ctx->NX, ctx->NY,
                                          *ctx->mg, md, fresh ? 0 : 1, ctx->d_fixed);
        if (ce != cudaSuccess)
            return fail(ctx, PLF_ECUDA, std::string("plf_solve_linear (shared level 0): ") +
                                            (ce == cudaErrorUnknown && comm_error()[0] ? comm_error() : cudaGetErrorString(ce)));
        CK(launch_merge_du(lc, ctx->Ndof, x, ctx->d_bc_du, ctx->d_fixed, ctx->d_du, &ctx->launches));
        CK(cudaStreamSynchronize(ctx->stream));
        if (iters) *iters = it;
        if (relres) *relres = rr;
        if (!(rr <= rtol))
            return fail(ctx, PLF_ENOCONV, "plf_solve_linear: CG (shared level 0) stopped after " + std::to_string(it) +
                                              " iterations with relative residual " + std::to_string(rr));
        return PLF_OK;
    }
    // same K and same Dirichlet mask as at the previous solve: the Galerkin hierarchy is still valid
    const bool mg_fresh = use_mg && ctx->skip_unchanged && ctx->mg_K_serial == ctx->K_serial && ctx->mg_mask_epoch == ctx->mask_epoch;
    if (use_mg) { ctx->mg_K_serial = ctx->K_serial; ctx->mg_mask_epoch = ctx->mask_epoch; }
    CK(launch_cg(lc, ctx->Ndof, ctx->d_rowptr, ctx->d_col, ctx->d_val, ctx->d_fixed, ctx->d_b, x, ctx->d_r, ctx->d_z, ctx->d_p, ctx->d_q,
                 ctx->d_diag, ctx->d_partial, counter_ptr(ctx), ctx->d_scalars + 256, ctx->h_scalars + 256, rtol, max_iter,
                 use_mg ? 4 : 50, &it, &rr, &ctx->launches, use_mg ? ctx->mg : nullptr, ctx->solver_warm, ctx->NX, ctx->NY, ctx->cg_graph, mg_fresh ? 0 : 1));
    CK(launch_merge_du(lc, ctx->Ndof, x, ctx->d_bc_du, ctx->d_fixed, ctx->d_du, &ctx->launches));
    CK(cudaStreamSynchronize(ctx->stream));
    if (iters) *iters = it;
    if (relres) *relres = rr;
    if (!(rr <= rtol))   // also catches NaN (break-down on an indefinite tangent)
        return fail(ctx, PLF_ENOCONV, "plf_solve_linear: CG stopped after " + std::to_string(it) +
                                          " iterations with relative residual " + std::to_string(rr));
    return PLF_OK;
}
