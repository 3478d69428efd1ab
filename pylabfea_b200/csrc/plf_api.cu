// plf_api.cu -- the C-ABI of libplfb200.so (include/plfb200.h): context, materials, mesh, device buffers and the
// launch sequences.  Host code only; all arithmetic that must be bit-identical to numpy (B matrices) uses plain
// IEEE double operations in the reference's order (this file is compiled with -ffp-contract=off) and std::fma
// where numpy's matmul fuses.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "model_kernels.h"
#include "plf_device.cuh"
#include "plf_host.h"
#include "plf_model.h"

using namespace plf;

static std::string g_create_error;

int fail(plf_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

// ---------------------------------------------------------------------------------------------------------------
// __constant__ bank ownership.  c_mat / c_geom (response_kernels.cu, model_kernels.cu, assembly_kernels.cu) exist once
// per device; several contexts (a Model, the default context behind Material.calc_*, a second Model) share them.  The
// context that launches next must find ITS tables there: plf_bind() is called at the head of every launch sequence.
// ---------------------------------------------------------------------------------------------------------------
static const int kMaxDevices = 64;
static plf_ctx *g_bank_owner[kMaxDevices] = {nullptr};
static std::mutex g_bank_mutex;

int plf_bind(plf_ctx *ctx)
{
    if (!ctx) return PLF_EINVAL;
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    const int dev = ctx->device;
    cudaError_t e = cudaSetDevice(dev);
    if (e != cudaSuccess) return fail(ctx, PLF_ECUDA, std::string("plf_bind: ") + cudaGetErrorString(e));
    plf_ctx *prev = g_bank_owner[dev];
    if (prev == ctx && !ctx->tables_dirty) return PLF_OK;
    // kernels of the previous owner (asynchronous entry points: plf_batch_run, plf_setupK) may still read the banks
    if (prev && prev != ctx) cudaStreamSynchronize(prev->stream);
    if (prev == ctx) cudaStreamSynchronize(ctx->stream);
    if (ctx->nmat > 0) {
        if ((e = response_upload_materials(ctx->hmats, ctx->nmat, ctx->stream)) != cudaSuccess ||
            (e = modelk_init()) != cudaSuccess ||
            (e = modelk_upload_materials(ctx->hmats, ctx->nmat, ctx->stream)) != cudaSuccess)
            return fail(ctx, PLF_ECUDA, std::string("plf_bind (materials): ") + cudaGetErrorString(e));
    }
    if (ctx->ngeom > 0) {
        if ((e = assembly_upload_geom(ctx->geom, ctx->ngeom, ctx->stream)) != cudaSuccess ||
            (e = modelk_upload_geom(ctx->geom, ctx->ngeom, ctx->stream)) != cudaSuccess)
            return fail(ctx, PLF_ECUDA, std::string("plf_bind (geometry): ") + cudaGetErrorString(e));
    }
    g_bank_owner[dev] = ctx;
    ctx->tables_dirty = false;
    return PLF_OK;
}

static void bank_release(plf_ctx *ctx)
{
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    if (ctx->device >= 0 && ctx->device < kMaxDevices && g_bank_owner[ctx->device] == ctx) g_bank_owner[ctx->device] = nullptr;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, PLF_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

template <class T>
static int dev_alloc(plf_ctx *ctx, T **p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) return PLF_OK;
    CK(cudaMalloc((void **)p, n * sizeof(T)));
    return PLF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
extern "C" const char *plf_version(void) { return "plfb200 0.1.0 (sm_100a)"; }
extern "C" size_t plf_sizeof_material(void) { return sizeof(plf_material); }

extern "C" int plf_create(int device, plf_ctx **out)
{
    plf_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, PLF_EINVAL, "plf_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PLF_ECUDA, std::string("plf_create: no CUDA device (") + cudaGetErrorString(e) +
                                            "); this library has no CPU fallback");
    if (device < 0 || device >= ndev || device >= kMaxDevices) return fail(nullptr, PLF_EINVAL, "plf_create: device index out of range");
    CK(cudaSetDevice(device));
    ctx = new plf_ctx();
    ctx->device = device;
    if (const char *ev = getenv("PLF_PASSB_RECUR")) ctx->passb_recur = atoi(ev);   // A/B default of the option of the same name
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, PLF_ECUDA, cudaGetErrorString(e)); }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->mem_bytes = (int64_t)prop.totalGlobalMem;
    snprintf(ctx->dev_name, sizeof ctx->dev_name, "%s", prop.name);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = response_init()) != cudaSuccess || (e = cudaMalloc((void **)&ctx->d_scalars, 4096)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&ctx->h_scalars, 4096)) != cudaSuccess) {
        std::string msg = std::string("plf_create: ") + cudaGetErrorString(e) +
                          " (is this an sm_100a device? the library carries sm_100a code only)";
        delete ctx;
        return fail(nullptr, PLF_ECUDA, msg);
    }
    cudaMemset(ctx->d_scalars, 0, 4096);
    *out = ctx;
    return PLF_OK;
}

static void free_batch_bufs(plf_batch &b)
{
    cudaFree(b.mat_id); cudaFree(b.sig); cudaFree(b.epl); cudaFree(b.deps); cudaFree(b.fy); cudaFree(b.sig_out);
    cudaFree(b.depl); cudaFree(b.ctan); cudaFree(b.nsteps); cudaFree(b.worklist);
    cudaFree(b.ct21); cudaFree(b.el_words); cudaFree(b.pk_counts); cudaFree(b.pk_offsets); cudaFree(b.pk_totals);
    if (b.h_pk_totals) cudaFreeHost(b.h_pk_totals);
    b = plf_batch();
}

extern "C" void plf_destroy(plf_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bank_release(ctx);
    for (int i = 0; i < PLF_MAX_MATERIALS; i++) { cudaFree(ctx->d_sv[i]); cudaFree(ctx->d_dual[i]); }
    free_batch_bufs(ctx->scratch);
    model_free(ctx);
    plf_comm_destroy(ctx);
    cudaFree(ctx->d_flush);
    cudaFree(ctx->d_scalars);
    cudaFreeHost(ctx->h_scalars);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t ev : ctx->ev_pool) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev_step) if (ev) cudaEventDestroy(ev);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *plf_last_error(const plf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int plf_device_info(const plf_ctx *ctx, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                               int64_t *mem_bytes)
{
    if (!ctx) return PLF_EINVAL;
    if (name && name_len > 0) snprintf(name, name_len, "%s", ctx->dev_name);
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return PLF_OK;
}

extern "C" int plf_timer_start(plf_ctx *ctx)
{
    if (!ctx) return PLF_EINVAL;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return PLF_OK;
}

extern "C" int plf_timer_stop(plf_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return PLF_EINVAL;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return PLF_OK;
}

extern "C" int plf_event_record(plf_ctx *ctx, int idx)
{
    if (!ctx || idx < 0 || idx >= 4096) return PLF_EINVAL;
    if ((int)ctx->ev_pool.size() <= idx) {
        size_t old = ctx->ev_pool.size();
        ctx->ev_pool.resize(idx + 1, nullptr);
        for (size_t i = old; i < ctx->ev_pool.size(); i++) CK(cudaEventCreate(&ctx->ev_pool[i]));
    }
    CK(cudaEventRecord(ctx->ev_pool[idx], ctx->stream));
    return PLF_OK;
}

extern "C" int plf_event_elapsed(plf_ctx *ctx, int idx0, int idx1, float *ms)
{
    if (!ctx || !ms || idx0 < 0 || idx1 < 0 || idx0 >= (int)ctx->ev_pool.size() || idx1 >= (int)ctx->ev_pool.size())
        return PLF_EINVAL;
    CK(cudaEventSynchronize(ctx->ev_pool[idx1]));
    CK(cudaEventElapsedTime(ms, ctx->ev_pool[idx0], ctx->ev_pool[idx1]));
    return PLF_OK;
}

extern "C" int plf_synchronize(plf_ctx *ctx)
{
    if (!ctx) return PLF_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_flush_l2(plf_ctx *ctx)
{
    if (!ctx) return PLF_EINVAL;
    const size_t bytes = 256ull << 20;
    if (!ctx->d_flush) CK(cudaMalloc((void **)&ctx->d_flush, bytes));
    ctx->flush_val ^= 0x5a;
    CK(cudaMemsetAsync(ctx->d_flush, ctx->flush_val, bytes, ctx->stream));
    return PLF_OK;
}

extern "C" int64_t plf_launch_count(const plf_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int plf_last_kernels(plf_ctx *ctx, char *buf, int len)
{
    if (!ctx || !buf || len < 1) return PLF_EINVAL;
    snprintf(buf, (size_t)len, "%s", ctx->trace);
    ctx->trace[0] = 0;
    ctx->trace_on = 1;   // recording starts with the first query
    return PLF_OK;
}

extern "C" int plf_pinned_alloc(size_t bytes, void **out)
{
    if (!out) return PLF_EINVAL;
    return cudaMallocHost(out, bytes) == cudaSuccess ? PLF_OK : PLF_ECUDA;
}

extern "C" int plf_pinned_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? PLF_OK : PLF_ECUDA; }

// ---------------------------------------------------------------------------------------------------------------
// materials
// ---------------------------------------------------------------------------------------------------------------
// Element.CV (model.py:270-304)
static void element_CV(const plf_material &m, int planestress, double *CV)
{
    if (!planestress) { memcpy(CV, m.CV, 36 * sizeof(double)); return; }
    memset(CV, 0, 36 * sizeof(double));
    double hh = m.E / (1 - m.nu * m.nu);
    double C12 = m.nu * hh, C11 = hh;
    CV[0] = C11; CV[1] = C12; CV[6] = C12; CV[7] = C11; CV[35] = m.C44;
}

// compliance of the radial-return correction (material.py:315-320)
static void compliance(const double *CV, double *SV)
{
    memset(SV, 0, 36 * sizeof(double));
    const int n = CV[14] > 1. ? 3 : 2;
    double a[3][6];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { a[i][j] = CV[6 * i + j]; a[i][n + j] = (i == j) ? 1. : 0.; }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(a[c][j], a[p][j]);
        double piv = a[c][c];
        if (piv == 0.) return;
        for (int j = 0; j < 2 * n; j++) a[c][j] /= piv;
        for (int r = 0; r < n; r++) if (r != c) {
            double f = a[r][c];
            for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) SV[6 * i + j] = a[i][n + j];
    for (int i = 3; i < 6; i++) if (CV[7 * i] > 1.) SV[7 * i] = 1. / CV[7 * i];
}

extern "C" int plf_set_materials(plf_ctx *ctx, int nmat, const plf_material *mats, int planestress)
{
    if (!ctx) return PLF_EINVAL;
    if (nmat < 1 || nmat > PLF_MAX_MATERIALS || !mats)
        return fail(ctx, PLF_EINVAL, "plf_set_materials: 1 <= nmat <= " + std::to_string(PLF_MAX_MATERIALS));
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < nmat; i++) {
        const plf_material &m = mats[i];
        DevMat &d = ctx->hmats[i];
        memset(&d, 0, sizeof d);
        if (m.kind < PLF_MAT_ELASTIC || m.kind > PLF_MAT_SVC3) return fail(ctx, PLF_EINVAL, "plf_set_materials: unknown kind");
        // the sdim = 3 SVC runs on the SVC kernels with another feature map / gradient / equivalent stress (SvcModel, svc_princ)
        d.kind = (m.kind == PLF_MAT_SVC3) ? PLF_MAT_SVC : m.kind;
        d.svc_princ = (m.kind == PLF_MAT_SVC3) ? 1 : 0;
        d.svc_nf = 6;
        d.tresca = m.tresca;
        d.lhs = m.lhs;
        d.dev_only = m.dev_only;
        d.E = m.E; d.nu = m.nu;
        d.sy = m.sy; d.khard = m.khard;
        for (int k = 0; k < 6; k++) d.hill[k] = (m.kind == PLF_MAT_ELASTIC) ? 1. : m.hill[k];
        for (int k = 0; k < 3; k++) {
            d.d[k] = (m.kind == PLF_MAT_ELASTIC) ? 0. : m.d[k];
            d.d3[k] = m.lhs ? d.d[k] : d.d[k] / 3.;   // material.py:829-832
        }
        element_CV(m, planestress, d.CV);
        compliance(d.CV, d.SV);
        d.cv_block = 1;
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
                if (!((r < 3 && c < 3) || r == c) && d.CV[6 * r + c] != 0.) d.cv_block = 0;
        d.has_d = (d.d[0] != 0. || d.d[1] != 0. || d.d[2] != 0.) ? 1 : 0;
        d.cv_sym = 1;
        for (int r = 0; r < 6; r++)
            for (int c = r + 1; c < 6; c++)
                if (d.CV[6 * r + c] != d.CV[6 * c + r]) d.cv_sym = 0;
        // isotropic elasticity in full 3-d form (C11 - C12 = 2 C44): enables hill_passB_iso for J2 (hill_fast.cuh, FM 3)
        {
            const double *C = d.CV;
            const double C11 = C[0], C12 = C[1], C44 = C[21];
            d.cv_iso = (d.cv_block && d.cv_sym && C[7] == C11 && C[14] == C11 && C[2] == C12 && C[8] == C12 && C[28] == C44 &&
                        C[35] == C44 && C44 > 0. && fabs(C11 - C12 - 2. * C44) <= 1e-13 * fabs(C11)) ? 1 : 0;
            d.iso_ihh = 1. / (3. * C44 + d.khard);
        }
        memcpy(d.bar_m1, m.barlat_m1, sizeof d.bar_m1);
        memcpy(d.bar_m2, m.barlat_m2, sizeof d.bar_m2);
        d.bar_exp = m.barlat_exp;
        d.bar_exp_int = (m.barlat_exp >= 2. && m.barlat_exp <= 64. && m.barlat_exp == std::floor(m.barlat_exp)) ? (int)m.barlat_exp : 0;
        if (m.kind == PLF_MAT_SVC || m.kind == PLF_MAT_SVC3) {
            if (m.nsv < 1 || !m.svc_sv || !m.svc_dual) return fail(ctx, PLF_EINVAL, "plf_set_materials: SVC table missing");
            const int np = (m.nsv + 127) / 128 * 128;   // four support vectors per lane and loop trip (svc_decision_lane)
            const int nfeat = m.svc_nfeat == 0 ? 6 : m.svc_nfeat;
            if (nfeat != 6 && !(nfeat == 15 && m.kind == PLF_MAT_SVC))
                return fail(ctx, PLF_EUNSUPPORTED, "plf_set_materials: SVC feature vectors with 6 (stress) or 15 (stress + work "
                                                   "hardening, sdim = 6) entries are supported; texture features are not");
            if (nfeat == 15 && !(m.svc_scale_wh > 0.)) return fail(ctx, PLF_EINVAL, "plf_set_materials: svc_scale_wh must be positive");
            const int nf = nfeat == 15 ? 12 : 6;
            if ((size_t)np * (nf + 1) * sizeof(double) > 227 * 1024)
                return fail(ctx, PLF_EUNSUPPORTED, "plf_set_materials: SVC table exceeds 227 KB of shared memory (" +
                                                       std::to_string(227 * 1024 / (8 * (nf + 1)) / 128 * 128) + " support vectors at most)");
            d.nsv = m.nsv; d.nsv_pad = np; d.svc_nf = nf;
            d.gamma = m.svc_gamma; d.rho = m.svc_rho; d.scale_seq = m.svc_scale_seq; d.inv_scale_seq = 1. / m.svc_scale_seq;
            if (nf == 12) {
                d.inv_scale_wh = 1. / m.svc_scale_wh;
                d.khard_fac = m.svc_scale_seq / m.svc_scale_wh;
                // khard stays as the host set it: the value Material.khard has before the first gradient evaluation
            } else {
                d.khard = 0.;   // material.py:812-814: the SVC gradient resets khard (no work-hardening features)
            }
            std::vector<double> soa((size_t)nf * np, 0.), dual(np, 0.);
            for (int s = 0; s < m.nsv; s++) {
                for (int k = 0; k < nf; k++) soa[(size_t)k * np + s] = m.svc_sv[(size_t)nfeat * s + k];
                double fold = 1.;
                if (nfeat == 15) {
                    // features 12..14 (accumulated strain, maximum stress, flag) are always 0 here: their share of
                    // exp(-gamma |x - sv|^2) is a constant factor of the support vector's weight
                    double r = 0.;
                    for (int k = 12; k < 15; k++) r += m.svc_sv[(size_t)nfeat * s + k] * m.svc_sv[(size_t)nfeat * s + k];
                    fold = std::exp(-m.svc_gamma * r);
                }
                dual[s] = m.svc_dual[s] * fold;
            }
            // padded support vectors: dual = 0 so they add exactly 0 to every sum
            int rc;
            if ((rc = dev_alloc(ctx, &ctx->d_sv[i], soa.size())) != PLF_OK) return rc;
            if ((rc = dev_alloc(ctx, &ctx->d_dual[i], dual.size())) != PLF_OK) return rc;
            CK(cudaMemcpy(ctx->d_sv[i], soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(ctx->d_dual[i], dual.data(), dual.size() * sizeof(double), cudaMemcpyHostToDevice));
            d.sv = ctx->d_sv[i];
            d.dual = ctx->d_dual[i];
        }
    }
    ctx->nmat = nmat;
    ctx->planestress = planestress ? 1 : 0;
    ctx->mats.assign(mats, mats + nmat);
    ctx->tables_dirty = true;   // uploaded to the __constant__ banks by the next plf_bind
    ctx->mesh_ready = false;   // geometry classes depend on the materials for plane stress
    return PLF_OK;
}

extern "C" int plf_get_element_CV(const plf_ctx *ctx, int id, double *CV)
{
    if (!ctx || !CV || id < 0 || id >= ctx->nmat) return PLF_EINVAL;
    memcpy(CV, ctx->hmats[id].CV, 36 * sizeof(double));
    return PLF_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------------------------------------------
static int batch_reserve(plf_ctx *ctx, plf_batch &b, int64_t n)
{
    if (n <= b.cap) { b.n = n; return PLF_OK; }
    free_batch_bufs(b);
    int rc;
    size_t N = (size_t)n;
    if ((rc = dev_alloc(ctx, &b.mat_id, N)) || (rc = dev_alloc(ctx, &b.sig, 6 * N)) || (rc = dev_alloc(ctx, &b.epl, 6 * N)) ||
        (rc = dev_alloc(ctx, &b.deps, 6 * N)) || (rc = dev_alloc(ctx, &b.fy, N)) || (rc = dev_alloc(ctx, &b.sig_out, 6 * N)) ||
        (rc = dev_alloc(ctx, &b.depl, 6 * N)) || (rc = dev_alloc(ctx, &b.ctan, 36 * N)) || (rc = dev_alloc(ctx, &b.nsteps, N)) ||
        (rc = dev_alloc(ctx, &b.worklist, N)))
        return rc;
    b.cap = b.n = n;
    return PLF_OK;
}

extern "C" int plf_batch_create(plf_ctx *ctx, int64_t n, plf_batch **out)
{
    if (!ctx || !out || n < 0) return PLF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    plf_batch *b = new plf_batch();
    int rc = batch_reserve(ctx, *b, std::max<int64_t>(n, 1));
    if (rc != PLF_OK) { free_batch_bufs(*b); delete b; return rc; }
    b->n = n;
    *out = b;
    return PLF_OK;
}

extern "C" void plf_batch_destroy(plf_ctx *ctx, plf_batch *b)
{
    if (!b) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    free_batch_bufs(*b);
    delete b;
}

extern "C" int plf_batch_upload(plf_ctx *ctx, plf_batch *b, const int32_t *mat_id, const double *sig, const double *epl,
                                const double *deps)
{
    if (!ctx || !b || !sig || !epl || !deps) return PLF_EINVAL;
    const size_t N = (size_t)b->n;
    b->mat_from_mesh = false;
    if (mat_id) {
        for (size_t e = 0; e < N; e++)
            if (mat_id[e] < 0 || mat_id[e] >= std::max(ctx->nmat, 1)) return fail(ctx, PLF_EINVAL, "plf_batch_upload: material id out of range");
    }
    if (mat_id) CK(cudaMemcpyAsync(b->mat_id, mat_id, N * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemsetAsync(b->mat_id, 0, N * sizeof(int32_t), ctx->stream));
    CK(cudaMemcpyAsync(b->sig, sig, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b->epl, epl, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b->deps, deps, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return PLF_OK;
}

extern "C" int plf_batch_run(plf_ctx *ctx, plf_batch *b)
{
    if (!ctx || !b) return PLF_EINVAL;
    if (ctx->nmat == 0) return fail(ctx, PLF_ESTATE, "plf_batch_run: no materials set");
    for (int i = 0; i < ctx->nmat; i++)
        if (ctx->hmats[i].tresca && (ctx->hmats[i].kind == PLF_MAT_HILL || ctx->hmats[i].kind == PLF_MAT_HILL3))
            // material.py:824-825: the first plastic step of a Tresca material raises in the reference
            return fail(ctx, PLF_EUNSUPPORTED, "response: analytical gradient for Tresca not implemented");
    if (int brc = plf_bind(ctx)) return brc;
    LaunchCtx lc = make_lc(ctx);
    CK(launch_response(lc, ctx->hmats, ctx->nmat, b->n, b->mat_id, b->sig, b->epl, b->deps, b->fy, b->sig_out, b->depl,
                       b->ctan, b->nsteps, b->worklist, batch_counters(ctx),
                       batch_queue(ctx), ctx->compact, &ctx->launches));
    return PLF_OK;
}

extern "C" int plf_batch_download(plf_ctx *ctx, plf_batch *b, double *fy, double *sig_out, double *depl, double *ctan,
                                  int32_t *nsteps)
{
    if (!ctx || !b) return PLF_EINVAL;
    const size_t N = (size_t)b->n;
    if (fy) CK(cudaMemcpyAsync(fy, b->fy, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sig_out) CK(cudaMemcpyAsync(sig_out, b->sig_out, 6 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (depl) CK(cudaMemcpyAsync(depl, b->depl, 6 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctan) CK(cudaMemcpyAsync(ctan, b->ctan, 36 * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (nsteps) CK(cudaMemcpyAsync(nsteps, b->nsteps, N * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_set_option(plf_ctx *ctx, const char *key, int value)
{
    if (!ctx || !key) return PLF_EINVAL;
    if (!strcmp(key, "compact")) { ctx->compact = value; return PLF_OK; }
    if (!strcmp(key, "assembly_tiled")) { ctx->assembly_tiled = value; return PLF_OK; }
    if (!strcmp(key, "passa_direct")) { ctx->passa_direct = value; return PLF_OK; }
    if (!strcmp(key, "passa_bulk")) { ctx->passa_bulk = value; return PLF_OK; }
    if (!strcmp(key, "max_fm")) { ctx->max_fm = value; return PLF_OK; }
    if (!strcmp(key, "k1_only")) { ctx->k1_only = value; return PLF_OK; }
    if (!strcmp(key, "passb_recur")) { ctx->passb_recur = value; return PLF_OK; }
    if (!strcmp(key, "solver_mg")) { ctx->solver_mg = value; return PLF_OK; }
    if (!strcmp(key, "solver_warm")) { ctx->solver_warm = value; return PLF_OK; }
    if (!strcmp(key, "skip_unchanged")) { ctx->skip_unchanged = value; return PLF_OK; }
    if (!strcmp(key, "solver_hybrid")) { ctx->solver_hybrid = value; return PLF_OK; }
    if (!strcmp(key, "mg_nu") || !strcmp(key, "mg_coarse_sweeps") || !strcmp(key, "mg_omega_permille") || !strcmp(key, "mg_fuse_tail")) {
        // multigrid tuning knobs (take effect at the next plf_mesh_structured / solve; the CG graph is rebuilt)
        if (!strcmp(key, "mg_nu")) ctx->mg_nu = value & ~1;
        else if (!strcmp(key, "mg_coarse_sweeps")) ctx->mg_coarse_sweeps = value;
        else if (!strcmp(key, "mg_fuse_tail")) ctx->mg_fuse_tail = value;
        else ctx->mg_omega = value * 1e-3;
        model_apply_mg_options(ctx);
        return PLF_OK;
    }
    return fail(ctx, PLF_EINVAL, std::string("plf_set_option: unknown key ") + key);
}

extern "C" int plf_response_batch(plf_ctx *ctx, int64_t n, const int32_t *mat_id, const double *sig, const double *epl,
                                  const double *deps, double *fy, double *sig_out, double *depl, double *ctan, int32_t *nsteps)
{
    if (!ctx || n < 0 || !sig || !epl || !deps) return fail(ctx, PLF_EINVAL, "plf_response_batch: bad arguments");
    if (ctx->nmat == 0) return fail(ctx, PLF_ESTATE, "plf_response_batch: no materials set");
    if (n == 0) return PLF_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = batch_reserve(ctx, ctx->scratch, n);
    if (rc != PLF_OK) return rc;
    if ((rc = plf_batch_upload(ctx, &ctx->scratch, mat_id, sig, epl, deps)) != PLF_OK) return rc;
    if ((rc = plf_batch_run(ctx, &ctx->scratch)) != PLF_OK) return rc;
    return plf_batch_download(ctx, &ctx->scratch, fy, sig_out, depl, ctan, nsteps);
}

enum { OP_SEQ = 0, OP_YF = 1, OP_FGRAD = 2, OP_EPLDOT = 3, OP_CTAN = 4, OP_MLFULL = 5 };

static int prim_call(plf_ctx *ctx, int op, int mat, int64_t n, const double *sig, const double *epl, const double *deps,
                     const double *ld, double *out, int out_w)
{
    if (!ctx || n < 0 || !sig || !out) return fail(ctx, PLF_EINVAL, "batched constitutive call: bad arguments");
    if (mat < 0 || mat >= ctx->nmat) return fail(ctx, PLF_EINVAL, "batched constitutive call: material id out of range");
    const DevMat &m = ctx->hmats[mat];
    if ((op == OP_FGRAD || op == OP_EPLDOT || op == OP_CTAN) && m.tresca)
        // material.py:824-825: the reference raises ValueError for Tresca.  (It also raises for Barlat, 822-823; the
        // Barlat gradient is implemented here -- barlat_seq_grad -- because BASELINE config 5 needs the return mapping.)
        return fail(ctx, PLF_EUNSUPPORTED, "calc_fgrad: analytical gradient for Tresca not implemented");
    if (n == 0) return PLF_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = batch_reserve(ctx, ctx->scratch, n);
    if (rc != PLF_OK) return rc;
    plf_batch &b = ctx->scratch;
    const size_t N = (size_t)n;
    CK(cudaMemcpyAsync(b.sig, sig, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (epl) CK(cudaMemcpyAsync(b.epl, epl, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (deps) CK(cudaMemcpyAsync(b.deps, deps, 6 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double *d_ld = nullptr;
    if (ld) {
        d_ld = (double *)(ctx->d_scalars + 64);
        CK(cudaMemcpyAsync(d_ld, ld, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int brc = plf_bind(ctx)) return brc;
    LaunchCtx lc = make_lc(ctx);
    CK(launch_prim(lc, ctx->hmats, op, mat, n, b.sig, epl ? b.epl : nullptr, deps ? b.deps : nullptr, d_ld, b.ctan,
                   &ctx->launches));
    CK(cudaMemcpyAsync(out, b.ctan, (size_t)out_w * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_calc_seq_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, double *seq)
{
    return prim_call(ctx, OP_SEQ, mat, n, sig, nullptr, nullptr, nullptr, seq, 1);
}
extern "C" int plf_calc_yf_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *yf)
{
    return prim_call(ctx, OP_YF, mat, n, sig, epl, nullptr, nullptr, yf, 1);
}
extern "C" int plf_calc_fgrad_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *fgrad)
{
    return prim_call(ctx, OP_FGRAD, mat, n, sig, epl, nullptr, nullptr, fgrad, 6);
}
extern "C" int plf_ml_full_yf_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, const double *ld,
                                    double *yf)
{
    return prim_call(ctx, OP_MLFULL, mat, n, sig, epl, nullptr, ld, yf, 1);
}
extern "C" int plf_epl_dot_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, const double *deps,
                                 double *pdot)
{
    if (!deps) return fail(ctx, PLF_EINVAL, "plf_epl_dot_batch: deps is NULL");
    return prim_call(ctx, OP_EPLDOT, mat, n, sig, epl, deps, nullptr, pdot, 6);
}
extern "C" int plf_ctan_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *ctan)
{
    return prim_call(ctx, OP_CTAN, mat, n, sig, epl, nullptr, nullptr, ctan, 36);
}
