// model_kernels.cu -- device-resident load-step loop of Model.solve (model.py:979-1450):
//   sweep   : per plastic element deps = B du, Material.response, convergence measure, tangent replacement
//             (model.py:1338-1361), thread-per-element (Hill/J2) or warp-per-element (SVC)
//   commit  : state update (1383-1392), calc_global element sums (1500-1511), calc_scf (1036-1067)
//   solver  : CSR SpMV + Jacobi-preconditioned CG on the BC-reduced system (1028-1033, 1291, 1335)
// All reductions are deterministic (fixed-order partial sums, last-block-done final pass).
#include "csr_stencil.cuh"
#include "hill_fast.cuh"
#include "plf_device.cuh"
#include "plf_host.h"
#include "model_kernels.h"
#include "mg.h"

namespace plf {

__constant__ DevMat c_mat[PLF_MAX_MATERIALS];
__constant__ GeomClass c_geom[PLF_MAX_GEOM];

cudaError_t modelk_upload_materials(const DevMat *h, int n, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_mat, h, sizeof(DevMat) * n, 0, cudaMemcpyHostToDevice, st);
}
cudaError_t modelk_upload_geom(const GeomClass *h, int n, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_geom, h, sizeof(GeomClass) * n, 0, cudaMemcpyHostToDevice, st);
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
// element strain from nodal values: sum_g B_g u_e = bsum u_e (Element.deps / eps_t, model.py:387-411)
__device__ __forceinline__ void element_strain(const GeomClass &gc, int NY, int64_t e, const double *__restrict__ u, double *eps)
{
    const int64_t j = e / NY, k = e - j * NY;
    const int64_t n1 = j * (NY + 1) + k, n3 = n1 + NY + 1;
    double ue[8];
    const double2 a0 = __ldg(reinterpret_cast<const double2 *>(u + 2 * n1));
    const double2 a1 = __ldg(reinterpret_cast<const double2 *>(u + 2 * n1 + 2));
    const double2 b0 = __ldg(reinterpret_cast<const double2 *>(u + 2 * n3));
    const double2 b1 = __ldg(reinterpret_cast<const double2 *>(u + 2 * n3 + 2));
    ue[0] = a0.x; ue[1] = a0.y; ue[2] = a1.x; ue[3] = a1.y; ue[4] = b0.x; ue[5] = b0.y; ue[6] = b1.x; ue[7] = b1.y;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        if (r == 3 || r == 4) { eps[r] = 0.; continue; }
        double s = 0.;
#pragma unroll
        for (int c = 0; c < 8; c++) s += gc.bsum[r][c] * ue[c];
        eps[r] = s;
    }
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    // v >= 0: IEEE bit patterns of non-negative doubles order like unsigned integers
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// analytic yield models that share the thread-per-element kernels: Hill-6p / J2 (HillModel), Yld2004-18p (BarlatModel)
// and the sdim = 3 principal-stress formulation (HillPrincModel)
template <class F>
__device__ __forceinline__ void with_model(const DevMat &m, F f)
{
    if (m.kind == PLF_MAT_BARLAT) {
        BarlatModel ym(m);
        f(ym);
    } else if (m.kind == PLF_MAT_HILL3) {
        HillPrincModel ym(m);
        f(ym);
    } else {
        HillModel ym(m);
        f(ym);
    }
}

struct ElemStats {
    bool active = false, changed = false, nonconv = false;
    double hh = 0.;
    int br = 0;
};

__device__ __forceinline__ void stats_atomics(const ElemStats &es, SweepStats *st)
{
    if (!es.active) return;
    if (es.changed) atomicOr(&st->change, 1u);
    if (es.nonconv) atomicAdd(&st->nonconv, 1u);
    atomic_max_nonneg(&st->max_hh, es.hh);
    atomicAdd(&st->counts[es.br & 3], 1ull);
    if (es.br & 4) atomicAdd(&st->counts[3], 1ull);
}

// block-aggregated statistics: a handful of atomics per block instead of per element
__device__ __forceinline__ void stats_block(const ElemStats &es, SweepStats *st)
{
    __shared__ double s_hh[32];
    const int c_chg = __syncthreads_or(es.active && es.changed);
    const int c_non = __syncthreads_count(es.active && es.nonconv);
    const int c0 = __syncthreads_count(es.active && (es.br & 3) == 0);
    const int c1 = __syncthreads_count(es.active && (es.br & 3) == 1);
    const int c2 = __syncthreads_count(es.active && (es.br & 3) == 2);
    const int c3 = __syncthreads_count(es.active && (es.br & 4));
    double h = es.active ? es.hh : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h = fmax(h, __shfl_xor_sync(0xffffffffu, h, o));
    if ((threadIdx.x & 31) == 0) s_hh[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)((blockDim.x + 31) >> 5); w++) h = fmax(h, s_hh[w]);
        if (c_chg) atomicOr(&st->change, 1u);
        if (c_non) atomicAdd(&st->nonconv, (unsigned)c_non);
        if (h > 0.) atomic_max_nonneg(&st->max_hh, h);
        if (c0) atomicAdd(&st->counts[0], (unsigned long long)c0);
        if (c1) atomicAdd(&st->counts[1], (unsigned long long)c1);
        if (c2) atomicAdd(&st->counts[2], (unsigned long long)c2);
        if (c3) atomicAdd(&st->counts[3], (unsigned long long)c3);
    }
}

// what Model.solve does with the result of one response call (model.py:1342-1357)
__device__ __forceinline__ ElemStats sweep_finish(const DevMat &m, int64_t e, int nit, const double *sig_new, const double *depl,
                                             const double *gs, double fyv, int ns, int br, const double *epl_old,
                                             double *__restrict__ elstiff, double *__restrict__ res_sig,
                                             double *__restrict__ res_depl, double *__restrict__ fyld,
                                             int32_t *__restrict__ nsteps, unsigned char *__restrict__ has_res)
{
    ElemStats es;
    es.active = true;
    es.br = br;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        reinterpret_cast<double2 *>(res_sig + e * 6)[k] = make_double2(sig_new[2 * k], sig_new[2 * k + 1]);
        reinterpret_cast<double2 *>(res_depl + e * 6)[k] = make_double2(depl[2 * k], depl[2 * k + 1]);
    }
    has_res[e] = 1;
    nsteps[e] = ns;
    const double f = fyv / (m.sy + eps_eq6(epl_old) * m.khard);   // 1345
    fyld[e] = f;
    double *C = elstiff + e * 36;
    double old[36], hh = 0.;
#pragma unroll
    for (int k = 0; k < 18; k++) {
        double2 v = reinterpret_cast<const double2 *>(C)[k];
        old[2 * k] = v.x; old[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) {
            double d = old[6 * i + j] - gs[sym_idx(i, j)];
            hh += d * d;
        }
    hh = sqrt(hh);   // Frobenius norm, 1346-1347
    if (hh > PLF_DSTIFF_TOL) {
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
                double g0 = gs[sym_idx(i, j)], g1 = gs[sym_idx(i, j + 1)];
                if (nit >= 15) { g0 = 0.5 * (g0 + old[6 * i + j]); g1 = 0.5 * (g1 + old[6 * i + j + 1]); }   // 1350-1353
                reinterpret_cast<double2 *>(C)[(6 * i + j) >> 1] = make_double2(g0, g1);
            }
        es.changed = true;
    }
    es.nonconv = !(f <= PLF_YF_TOLERANCE * 1.0001);   // 1361
    es.hh = hh;
    return es;
}

// ---------------------------------------------------------------------------------------------------------------
// sweep, thread per element (Hill / J2).  GATHER: compacted second pass over the deferred elements.
// ---------------------------------------------------------------------------------------------------------------
//   FM: fast mode of the model's Hill / J2 materials (response_fast_mode: 0 general form, 4 / 8 J2 on isotropic elasticity with
//   the direct / recurrence form of the planar sub-step loop, 6 Hill on isotropic elasticity, 7 block-structured CV) -- the same
//   device functions as the batch kernels of response_kernels.cu, so a sweep and plf_response_batch agree bit for bit
template <int MODE, bool GATHER, int FM>
__global__ void __launch_bounds__(128, (GATHER && FM != 0) ? 3 : 1) sweep_hill_kernel(int64_t n, int NY, int nit, const int32_t *__restrict__ desc,
                                                         const double *__restrict__ sig, const double *__restrict__ epl,
                                                         const double *__restrict__ du, double *__restrict__ elstiff,
                                                         double *__restrict__ res_sig, double *__restrict__ res_depl,
                                                         double *__restrict__ fyld, int32_t *__restrict__ nsteps,
                                                         unsigned char *__restrict__ has_res, const int32_t *__restrict__ index,
                                                         int32_t *__restrict__ worklist, SweepStats *st, int64_t stat_first)
{
    // MODE 1 / !GATHER: pass A over all elements; MODE 0 / GATHER: persistent pass B over the work list (length read
    // from device memory: no host round trip between the passes); MODE 0 / !GATHER: unsplit response_core.
    __shared__ double s_xc[(GATHER && FM != 0) ? 12 * 128 : 1];   // correction sums + deps_r / 50 of the fast sub-step loops
    const int64_t total = GATHER ? (int64_t)st->work_count : n;
    const int64_t stride = GATHER ? (int64_t)gridDim.x * blockDim.x : total + blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {
        const int64_t t = t0 + threadIdx.x;
        ElemStats es;
        bool defer = false;
        int64_t e = -1;
        if (t < total) {
            e = GATHER ? (int64_t)index[t] : t;
            const int d = __ldg(desc + e);
            const DevMat &m = c_mat[d & 0xffff];
            if (m.kind == PLF_MAT_ELASTIC) {
                fyld[e] = 0.;   // 1358-1359
            } else if (m.kind == PLF_MAT_HILL) {
                double s[6], ep[6], de[6], dp[6], gs[21], fyv;
                int ns = 0, br;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double2 a = __ldg(reinterpret_cast<const double2 *>(sig + e * 6) + k);
                    double2 b = __ldg(reinterpret_cast<const double2 *>(epl + e * 6) + k);
                    s[2 * k] = a.x; s[2 * k + 1] = a.y; ep[2 * k] = b.x; ep[2 * k + 1] = b.y;
                }
                element_strain(c_geom[d >> 16], NY, e, du, de);
                if (MODE == 1) {
                    double ca[6], tc, tw;
                    br = (FM == 7) ? hillblk_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                       : (FM == 6) ? hilliso_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                       : (FM == 4 || FM == 8) ? iso_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                                              : hill_passA<0>(m, s, ep, de, dp, ca, tc, tw, fyv);
#pragma unroll
                    for (int i = 0; i < 6; i++)
#pragma unroll
                        for (int j = i; j < 6; j++) gs[sym_idx(i, j)] = m.CV[6 * i + j] * tc - (ca[i] * ca[j]) * tw;   // as pass A of the batch path
                } else if (GATHER) {
                    double epd[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) epd[k] = ep[k];
                    if (FM == 0) {
                        WRegs W;
                        br = hill_passB<0>(m, s, epd, de, W, gs, fyv);
                    } else if (planar_state(s, ep, de)) {   // every state of a 2-d model that started from planar fields
                        double *xc = s_xc + threadIdx.x;
                        br = (FM == 7) ? hill_passB_blk2d(m, s, epd, de, xc, 128, gs, fyv)
                           : (FM == 6) ? hill_passB_hilliso2d(m, s, epd, de, xc, 128, gs, fyv)
                           : (FM == 8) ? hill_passB_iso2d<true>(m, s, epd, de, xc, 128, gs, fyv)
                                       : hill_passB_iso2d<false>(m, s, epd, de, xc, 128, gs, fyv);
                    } else if (FM == 4 || FM == 8) {
                        br = hill_passB_iso(m, s, epd, de, s_xc + threadIdx.x, 128, gs, fyv);
                    } else {
                        WRegs W;
                        br = hill_passB<1>(m, s, epd, de, W, gs, fyv);
                    }
#pragma unroll
                    for (int k = 0; k < 6; k++) dp[k] = epd[k] - ep[k];
                    ns = PLF_MAXIT - 1;
                } else {
                    HillModel ym(m);
                    br = response_core<0>(ym, m, s, ep, de, dp, gs, fyv, ns);
                }
                if (br == -1) {
                    defer = true;
                } else {
                    es = sweep_finish(m, e, nit, s, dp, gs, fyv, ns, br, ep, elstiff, res_sig, res_depl, fyld, nsteps, has_res);
                }
            }
        }
        if (e < stat_first) es.active = false;   // halo elements of a strip partition: updated here, counted by their owner
        if (MODE == 1) warp_append(defer, (int32_t)e, worklist, &st->work_count);
        stats_block(es, st);
    }
}

// sweep, thread per element, for the yield models with eigen-decompositions (Yld2004-18p, sdim = 3): unsplit
// response_core, FP64-bound
template <class Model, int KIND, int MODE>
__global__ void __launch_bounds__(128, 4) sweep_model_kernel(int64_t n, int NY, int nit, const int32_t *__restrict__ desc,
                                                           const double *__restrict__ sig, const double *__restrict__ epl,
                                                           const double *__restrict__ du, double *__restrict__ elstiff,
                                                           double *__restrict__ res_sig, double *__restrict__ res_depl,
                                                           double *__restrict__ fyld, int32_t *__restrict__ nsteps,
                                                           unsigned char *__restrict__ has_res, int32_t *__restrict__ worklist,
                                                           SweepStats *st, int64_t stat_first)
{
    // MODE 0: unsplit; MODE 1: pass over all elements, 50-sub-step elements deferred to the work list; MODE 2: persistent
    // pass over the list (length read from device memory)
    const int64_t total = (MODE == 2) ? (int64_t)st->work_count : n;
    const int64_t stride = (MODE == 2) ? (int64_t)gridDim.x * blockDim.x : total + blockDim.x;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {
        const int64_t t = t0 + threadIdx.x;
        ElemStats es;
        bool defer = false;
        int64_t e = -1;
        if (t < total) {
            e = (MODE == 2) ? (int64_t)worklist[t] : t;
            const int d = __ldg(desc + e);
            const DevMat &m = c_mat[d & 0xffff];
            if (m.kind == KIND) {
                double s[6], ep[6], de[6], dp[6], gs[21], fyv;
                int ns = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) { s[k] = __ldg(sig + e * 6 + k); ep[k] = __ldg(epl + e * 6 + k); }
                element_strain(c_geom[d >> 16], NY, e, du, de);
                Model ym(m);
                const int br = response_core<MODE == 1 ? 1 : 0>(ym, m, s, ep, de, dp, gs, fyv, ns);
                if (br < 0) defer = true;
                else es = sweep_finish(m, e, nit, s, dp, gs, fyv, ns, br, ep, elstiff, res_sig, res_depl, fyld, nsteps, has_res);
            }
        }
        if (e < stat_first) es.active = false;
        if (MODE == 1) warp_append(defer, (int32_t)e, worklist, &st->work_count);
        stats_block(es, st);
    }
}

// sweep, warp per element, one SVC material (SV table in shared memory)
constexpr int SVCW = 16;   // warps per block of the SVC kernels (one block per SM, <= 128 registers)

__global__ void __launch_bounds__(SVCW * 32, 1) sweep_svc_kernel(int64_t n, int NY, int nit, int mid, const int32_t *__restrict__ desc,
                                                        const double *__restrict__ sig, const double *__restrict__ epl,
                                                        const double *__restrict__ du, double *__restrict__ elstiff,
                                                        double *__restrict__ res_sig, double *__restrict__ res_depl,
                                                        double *__restrict__ fyld, int32_t *__restrict__ nsteps,
                                                        unsigned char *__restrict__ has_res, SweepStats *st, int64_t stat_first)
{
    extern __shared__ double smem[];
    const DevMat &m = c_mat[mid];
    const int np = m.nsv_pad;
    const int nf = m.svc_nf;
    for (int i = threadIdx.x; i < nf * np; i += blockDim.x) smem[i] = m.sv[i];
    for (int i = threadIdx.x; i < np; i += blockDim.x) smem[nf * np + i] = m.dual[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    SvcModel ym(m, smem, smem + nf * np, lane);
    for (;;) {
        unsigned long long e0 = 0;
        if (lane == 0) e0 = atomicAdd(&st->queue, 1ull);
        const int64_t e = (int64_t)__shfl_sync(0xffffffffu, e0, 0);
        if (e >= n) break;
        const int d = __ldg(desc + e);
        if ((d & 0xffff) != mid) continue;
        double s[6], ep[6], de[6], dp[6], gs[21], fyv;
        int ns;
#pragma unroll
        for (int k = 0; k < 6; k++) { s[k] = __ldg(sig + e * 6 + k); ep[k] = __ldg(epl + e * 6 + k); }
        element_strain(c_geom[d >> 16], NY, e, du, de);
        const int br = response_core<0>(ym, m, s, ep, de, dp, gs, fyv, ns);
        if (lane == 0) {
            ElemStats es = sweep_finish(m, e, nit, s, dp, gs, fyv, ns, br, ep, elstiff, res_sig, res_depl, fyld, nsteps, has_res);
            if (e >= stat_first) stats_atomics(es, st);
        }
    }
}

cudaError_t modelk_init()
{
    return cudaFuncSetAttribute(sweep_svc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_sweep(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t Nel, int NY, int nit, const int32_t *desc,
                         const double *sig, const double *epl, const double *du, double *elstiff, double *res_sig,
                         double *res_depl, double *fyld, int32_t *nsteps, unsigned char *has_res, int32_t *worklist,
                         void *d_stats, void *h_stats, int compact, int64_t *launches, int64_t stat_first)
{
    SweepStats *st = (SweepStats *)d_stats;
    cudaError_t e = cudaMemsetAsync(st, 0, sizeof(SweepStats), lc.stream);
    if (e != cudaSuccess) return e;
    bool any_hill = false;
    for (int i = 0; i < nmat; i++) any_hill |= (hmats[i].kind == PLF_MAT_HILL || hmats[i].kind == PLF_MAT_ELASTIC);
    const unsigned grid = (unsigned)((Nel + 127) / 128);
    if (any_hill) {
        if (compact) {
            int fm = response_fast_mode(hmats, nmat, lc.max_fm);
            if (fm == 4 && lc.passb_recur) fm = 8;
            if (fm != 4 && fm != 8 && fm != 6 && fm != 7) fm = 0;
            const unsigned bgrid = (unsigned)std::min<int64_t>((int64_t)lc.sm_count * 8, (int64_t)grid);
#define SWEEP_HILL(F)                                                                                                                  \
    do {                                                                                                                               \
        sweep_hill_kernel<1, false, F><<<grid, 128, 0, lc.stream>>>(Nel, NY, nit, desc, sig, epl, du, elstiff, res_sig, res_depl,     \
                                                                     fyld, nsteps, has_res, nullptr, worklist, st, stat_first);        \
        sweep_hill_kernel<0, true, F><<<bgrid, 128, 0, lc.stream>>>(0, NY, nit, desc, sig, epl, du, elstiff, res_sig, res_depl, fyld, \
                                                                     nsteps, has_res, worklist, nullptr, st, stat_first);              \
    } while (0)
            if (fm == 8) SWEEP_HILL(8);
            else if (fm == 4) SWEEP_HILL(4);
            else if (fm == 6) SWEEP_HILL(6);
            else if (fm == 7) SWEEP_HILL(7);
            else SWEEP_HILL(0);
#undef SWEEP_HILL
            *launches += 2;
        } else {
            sweep_hill_kernel<0, false, 0><<<grid, 128, 0, lc.stream>>>(Nel, NY, nit, desc, sig, epl, du, elstiff, res_sig, res_depl,
                                                                      fyld, nsteps, has_res, nullptr, nullptr, st, stat_first);
            (*launches)++;
        }
    }
    bool any_barlat = false, any_hill3 = false;
    for (int i = 0; i < nmat; i++) {
        any_barlat |= (hmats[i].kind == PLF_MAT_BARLAT);
        any_hill3 |= (hmats[i].kind == PLF_MAT_HILL3);
    }
    const unsigned mgrid = (unsigned)std::min<int64_t>((int64_t)lc.sm_count * 8, (int64_t)grid);
#define SWEEP_MODEL(MODEL, KIND)                                                                                                       \
    do {                                                                                                                               \
        if (compact) {                                                                                                                 \
            if ((e = cudaMemsetAsync(&st->work_count, 0, sizeof(unsigned int), lc.stream)) != cudaSuccess) return e;                   \
            sweep_model_kernel<MODEL, KIND, 1><<<grid, 128, 0, lc.stream>>>(Nel, NY, nit, desc, sig, epl, du, elstiff, res_sig, res_depl, \
                                                                            fyld, nsteps, has_res, worklist, st, stat_first);          \
            sweep_model_kernel<MODEL, KIND, 2><<<mgrid, 128, 0, lc.stream>>>(Nel, NY, nit, desc, sig, epl, du, elstiff, res_sig,        \
                                                                             res_depl, fyld, nsteps, has_res, worklist, st, stat_first); \
            *launches += 2;                                                                                                            \
        } else {                                                                                                                       \
            sweep_model_kernel<MODEL, KIND, 0><<<grid, 128, 0, lc.stream>>>(Nel, NY, nit, desc, sig, epl, du, elstiff, res_sig, res_depl, \
                                                                            fyld, nsteps, has_res, nullptr, st, stat_first);           \
            (*launches)++;                                                                                                             \
        }                                                                                                                              \
    } while (0)
    if (any_barlat) SWEEP_MODEL(BarlatModel, PLF_MAT_BARLAT);
    if (any_hill3) SWEEP_MODEL(HillPrincModel, PLF_MAT_HILL3);
#undef SWEEP_MODEL
    for (int i = 0; i < nmat; i++) {
        if (hmats[i].kind != PLF_MAT_SVC) continue;
        e = cudaMemsetAsync(&st->queue, 0, sizeof(unsigned long long), lc.stream);
        if (e != cudaSuccess) return e;
        int blocks = lc.sm_count;
        int64_t need = (Nel + SVCW - 1) / SVCW;
        if (blocks > need) blocks = (int)need;
        sweep_svc_kernel<<<blocks, SVCW * 32, sizeof(double) * (size_t)(hmats[i].svc_nf + 1) * (size_t)hmats[i].nsv_pad, lc.stream>>>(
            Nel, NY, nit, i, desc, sig, epl, du, elstiff, res_sig, res_depl, fyld, nsteps, has_res, st, stat_first);
        (*launches)++;
    }
    e = cudaMemcpyAsync(h_stats, st, sizeof(SweepStats), cudaMemcpyDeviceToHost, lc.stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(lc.stream);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// commit (model.py:1383-1392): element part.  u += du and f += K du are done by the vector kernels below.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) commit_elements_kernel(int64_t Nel, int NY, const int32_t *__restrict__ desc,
                                                              const double *__restrict__ u_new, const double *__restrict__ du,
                                                              const double *__restrict__ elstiff, double *__restrict__ sig,
                                                              double *__restrict__ eps, double *__restrict__ epl,
                                                              const double *__restrict__ res_sig,
                                                              const double *__restrict__ res_depl,
                                                              const unsigned char *__restrict__ has_res)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Nel) return;
    const int d = desc[e];
    const GeomClass &gc = c_geom[d >> 16];
    const DevMat &m = c_mat[d & 0xffff];
    double s[6], ep[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { s[k] = sig[e * 6 + k]; ep[k] = epl[e * 6 + k]; }
    if (has_res[e]) {   // 1389-1391
#pragma unroll
        for (int k = 0; k < 6; k++) { ep[k] += res_depl[e * 6 + k]; s[k] = res_sig[e * 6 + k]; }
    } else {            // 1386-1388: el.epl += el.depl(); el.sig += el.dsig()
        double de[6], ds[6], C[36];
        element_strain(gc, NY, e, du, de);
#pragma unroll
        for (int k = 0; k < 36; k++) C[k] = elstiff[e * 36 + k];
        if (m.kind == PLF_MAT_HILL || m.kind == PLF_MAT_BARLAT || m.kind == PLF_MAT_HILL3) {
            // Element.depl (model.py:424-437): Mat.epl_dot(sig, epl, CV, deps)
            with_model(m, [&](const auto &ym) {
                double cd[6], st[6], a[6], ca[6];
                matvec6(m.CV, de, cd);
#pragma unroll
                for (int k = 0; k < 6; k++) st[k] = s[k] + cd[k];
                const double sflow = m.sy + eps_eq6(ep) * m.khard;
                if (ym.yf(st, sflow) > PLF_YF_TOLERANCE) {
                    ym.fgrad(s, a);
                    matvec6(m.CV, a, ca);
                    double hh = m.khard, acd = 0.;
#pragma unroll
                    for (int k = 0; k < 6; k++) { hh += a[k] * ca[k]; acd += a[k] * cd[k]; }
                    const double lam = acd / hh;
#pragma unroll
                    for (int k = 0; k < 6; k++) ep[k] += lam * a[k];
                }
            });
        }
        matvec6(C, de, ds);
#pragma unroll
        for (int k = 0; k < 6; k++) s[k] += ds[k];
    }
    double et[6];
    element_strain(gc, NY, e, u_new, et);   // 1392
#pragma unroll
    for (int k = 0; k < 6; k++) { sig[e * 6 + k] = s[k]; epl[e * 6 + k] = ep[k]; eps[e * 6 + k] = et[k]; }
}

// ---------------------------------------------------------------------------------------------------------------
// deterministic reductions: per-block partials in fixed slots, summed in fixed order by the last block to finish
// ---------------------------------------------------------------------------------------------------------------
// Returns true in exactly one thread of the grid (thread 0 of the last block) after it has written out[0..NV).
template <int NV>
__device__ __forceinline__ bool block_reduce_to_partials(double *v, double *partial, unsigned int *counter, double *out,
                                                         int mode /*0 sum*/)
{
    __shared__ double sh[NV][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[i][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double x = 0.;
            for (int w = 0; w < nwarp; w++) x += sh[i][w];
            partial[(size_t)blockIdx.x * NV + i] = x;
        }
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // last block: every thread sums a strided slice in ascending order, then a fixed tree over the threads
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = 0.;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) x += __ldcg(partial + (size_t)b * NV + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[i][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double x = 0.;
            for (int w = 0; w < nwarp; w++) x += sh[i][w];
            out[i] = x;
        }
        *counter = 0;
        return true;
    }
    (void)mode;
    return false;
}

// calc_global element sums (model.py:1500-1507): sum_e Vel * {sig, eps, epl}
__global__ void __launch_bounds__(256) global_sums_kernel(int64_t Nel, const int32_t *__restrict__ desc,
                                                          const double *__restrict__ sig, const double *__restrict__ eps,
                                                          const double *__restrict__ epl, double *partial,
                                                          unsigned int *counter, double *out, int64_t e_first)
{
    double v[18];
#pragma unroll
    for (int i = 0; i < 18; i++) v[i] = 0.;
    for (int64_t e = e_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < Nel; e += (int64_t)gridDim.x * blockDim.x) {
        const double vel = c_geom[desc[e] >> 16].vel;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            v[k] += sig[e * 6 + k] * vel;
            v[6 + k] += eps[e * 6 + k] * vel;
            v[12 + k] += epl[e * 6 + k] * vel;
        }
    }
    block_reduce_to_partials<18>(v, partial, counter, out, 0);
}

// calc_scf (model.py:1036-1067), pass 1: per element scaling candidate hh with multiplicity (0, 1 or 2)
__global__ void __launch_bounds__(128) scf_elements_kernel(int64_t Nel, int NY, const int32_t *__restrict__ desc,
                                                           const double *__restrict__ sig, const double *__restrict__ epl,
                                                           const double *__restrict__ du, const double *__restrict__ elstiff,
                                                           double *__restrict__ hval, unsigned char *__restrict__ mult)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Nel) return;
    const int d = desc[e];
    const DevMat &m = c_mat[d & 0xffff];
    mult[e] = 0;
    hval[e] = 0.;
    if (m.kind == PLF_MAT_ELASTIC || m.kind == PLF_MAT_SVC) return;   // SVC elements: scf_svc_kernel
    double de[6], ds[6], C[36], s[6], ep[6];
    element_strain(c_geom[d >> 16], NY, e, du, de);
#pragma unroll
    for (int k = 0; k < 36; k++) C[k] = elstiff[e * 36 + k];
    matvec6(C, de, ds);
    with_model(m, [&](const auto &ym) {
        const double sref = ym.seq(ds);   // Stress(el.dsig()).seq(el.Mat), 1040
        if (!(sref > 0.1)) return;
#pragma unroll
        for (int k = 0; k < 6; k++) { s[k] = sig[e * 6 + k]; ep[k] = epl[e * 6 + k]; }
        const double sflow = m.sy + eps_eq6(ep) * m.khard;
        const double yf0 = ym.yf(s, sflow);
        if (yf0 < PLF_SPLIT_FY0) {
            hval[e] = fmin(1., -yf0 / sref);   // appended twice (1050 and 1054)
            mult[e] = 2;
        } else {
            hval[e] = fmin(1., sqrt(1.5) * sflow / sref);
            mult[e] = 1;
        }
    });
}

__global__ void __launch_bounds__(SVCW * 32, 1) scf_svc_kernel(int64_t Nel, int NY, int mid, const int32_t *__restrict__ desc,
                                                      const double *__restrict__ sig, const double *__restrict__ epl,
                                                      const double *__restrict__ du, const double *__restrict__ elstiff,
                                                      const double *__restrict__ sld, double *__restrict__ hval,
                                                      unsigned char *__restrict__ mult)
{
    extern __shared__ double smem[];
    const DevMat &m = c_mat[mid];
    const int np = m.nsv_pad;
    const int nf = m.svc_nf;
    for (int i = threadIdx.x; i < nf * np; i += blockDim.x) smem[i] = m.sv[i];
    for (int i = threadIdx.x; i < np; i += blockDim.x) smem[nf * np + i] = m.dual[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SvcModel ym(m, smem, smem + nf * np, lane);
    double ld[6];
#pragma unroll
    for (int k = 0; k < 6; k++) ld[k] = sld[k];
    for (int64_t e = (int64_t)blockIdx.x * SVCW + warp; e < Nel; e += (int64_t)gridDim.x * SVCW) {
        const int d = desc[e];
        if ((d & 0xffff) != mid) continue;
        double de[6], ds[6], C[36], s[6], ep[6];
        element_strain(c_geom[d >> 16], NY, e, du, de);
#pragma unroll
        for (int k = 0; k < 36; k++) C[k] = elstiff[e * 36 + k];
        matvec6(C, de, ds);
        const double sref = ym.seq(ds);
        double h = 0.;
        unsigned char mu = 0;
        if (sref > 0.1) {
#pragma unroll
            for (int k = 0; k < 6; k++) { s[k] = sig[e * 6 + k]; ep[k] = epl[e * 6 + k]; }
            const double sflow = m.sy + eps_eq6(ep) * m.khard;
            ym.set_epl(ep);
            double yf0 = ym.yf(s, sflow);
            if (yf0 < PLF_SPLIT_FY0) {
                yf0 = ym.ml_full_yf(s, sflow, ld);   // 1045-1048
                h = fmin(1., -yf0 / sref);
                mu = 2;
            } else {
                h = fmin(1., sqrt(1.5) * sflow / sref);
                mu = 1;
            }
        }
        if (lane == 0) { hval[e] = h; mult[e] = mu; }
    }
}

// pass 2/3: moments.  mode 0: {count, sum, min-as-negated-max}; mode 1: sum mult * (h - mean)^2
__global__ void __launch_bounds__(256) scf_moments_kernel(int64_t Nel, const double *__restrict__ hval,
                                                          const unsigned char *__restrict__ mult, double mean, int mode,
                                                          double *partial, unsigned int *counter, double *out, double *minval,
                                                          int64_t e_first)
{
    double v[2] = {0., 0.};
    double mn = 2.;
    for (int64_t e = e_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < Nel; e += (int64_t)gridDim.x * blockDim.x) {
        const double mu = (double)mult[e];
        if (mu == 0.) continue;
        const double h = hval[e];
        if (mode == 0) { v[0] += mu; v[1] += mu * h; mn = fmin(mn, h); }
        else { double dlt = h - mean; v[0] += mu * dlt * dlt; }
    }
    if (mode == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        // min over [0,1] values: atomicMin on the bit pattern of a non-negative double
        if ((threadIdx.x & 31) == 0 && mn <= 1.)
            atomicMin(reinterpret_cast<unsigned long long *>(minval), (unsigned long long)__double_as_longlong(fmax(mn, 0.)));
    }
    block_reduce_to_partials<2>(v, partial, counter, out, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// vector kernels and CG
// ---------------------------------------------------------------------------------------------------------------
// y (op)= A x for the CSR stiffness, 8 lanes per row.  mode 0: y = A x ; 1: y += A x ; 2: y = rhs - A x
// masked != nullptr: rows with masked[r] != 0 give y = 0 (mode 0/2).  dot != nullptr: partial sums of x . y.
__global__ void __launch_bounds__(256) spmv_kernel(int64_t nrows, const int32_t *__restrict__ rowptr,
                                                   const int32_t *__restrict__ col, const double *__restrict__ val,
                                                   const double *__restrict__ x, const double *__restrict__ rhs,
                                                   double *__restrict__ y, const unsigned char *__restrict__ masked, int mode,
                                                   double *partial, unsigned int *counter, double *dot_out)
{
    // grid-stride over groups of 32 rows per block (8 lanes per row): a capped grid keeps the number of partial sums
    // (and of atomics in the final reduction) small for the launch-bound CG loop
    const int sub = threadIdx.x & 7;
    double xy = 0.;
    for (int64_t r0 = (int64_t)blockIdx.x * 32; r0 < nrows; r0 += (int64_t)gridDim.x * 32) {
        const int64_t r = r0 + (threadIdx.x >> 3);
        double s = 0.;
        if (r < nrows) {
            const int32_t b = rowptr[r], en = rowptr[r + 1];
            for (int32_t i = b + sub; i < en; i += 8) s += __ldg(val + i) * __ldg(x + __ldg(col + i));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (r < nrows && sub == 0) {
            double out;
            const bool mk = masked && masked[r];
            if (mode == 0) out = mk ? 0. : s;
            else if (mode == 1) out = y[r] + s;
            else out = mk ? 0. : rhs[r] - s;
            y[r] = out;
            if (partial) xy += x[r] * out;
        }
    }
    if (partial) {
        double v[1] = {xy};
        block_reduce_to_partials<1>(v, partial, counter, dot_out, 0);
    }
}

// q = A p (masked) and p.q for the structured mesh: one thread per row on the closed-form CSR layout (csr_stencil.cuh),
// matrix pairs and vector pairs as 128-bit loads, no rowptr / col traffic.  The CSR kernel above reads 12 B per
// non-zero and was the slowest kernel of a multigrid-preconditioned iteration (129 us at 1024^2).
__global__ void __launch_bounds__(256) stencil_spmv_dot_kernel(int NX, int NY, const double *__restrict__ val,
                                                               const double *__restrict__ x, double *__restrict__ y,
                                                               const unsigned char *__restrict__ masked, double *partial,
                                                               unsigned int *counter, double *dot_out, int64_t row0, int64_t row1)
{
    // [row0, row1): all rows, or the rows of the node columns a rank owns (replicated solve with a shared level 0)
    const int NnY = NY + 1;
    double xy = 0.;
    for (int64_t row = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < row1; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = row >> 1;
        const int j = (int)(n / NnY), k = (int)(n - (int64_t)j * NnY);
        int64_t pos = csr_row_start(NX, NY, row);
        double s = 0.;
#pragma unroll
        for (int dj = -1; dj <= 1; dj++) {
            if (j + dj < 0 || j + dj > NX) continue;
#pragma unroll
            for (int dk = -1; dk <= 1; dk++) {
                if (k + dk < 0 || k + dk > NY) continue;
                const int64_t m = (int64_t)(j + dj) * NnY + (k + dk);
                const double2 a = __ldg(reinterpret_cast<const double2 *>(val + pos));
                const double2 xv = __ldg(reinterpret_cast<const double2 *>(x + 2 * m));
                s = fma(a.x, xv.x, s);
                s = fma(a.y, xv.y, s);
                pos += 2;
            }
        }
        const double out = masked[row] ? 0. : s;
        y[row] = out;
        xy += x[row] * out;
    }
    double v[1] = {xy};
    block_reduce_to_partials<1>(v, partial, counter, dot_out, 0);
}

// CG scalars in device memory.  rz is double-buffered by iteration parity so that no kernel has to swap it:
// step with parity P reads rz[P] (r.z before the step) and writes rz[1-P].
struct CgScalars {
    double pq, rz[2], tmp[2], rr;
    double target, iters;   // target = (rtol |b|)^2: once rr <= target the iteration freezes; iters counts real steps
};

// x += alpha p ; r -= alpha q ; z = r / diag ; r.z and r.r   (alpha = rz / pq computed by every thread)
// MG: the preconditioner is the multigrid V-cycle (mg_kernels.cu), applied between this kernel and cg_rz_kernel:
// z and r.z are not touched here.
// DIST: the sums of this rank stay in sc->tmp; they are all-reduced and published by cg_publish_kernel.
template <int P, bool MG, bool DIST = false>
__global__ void __launch_bounds__(256) cg_update_kernel(int64_t n, double *__restrict__ x, double *__restrict__ r,
                                                        double *__restrict__ z, const double *__restrict__ p,
                                                        const double *__restrict__ q, const double *__restrict__ diag,
                                                        const unsigned char *__restrict__ fixed, CgScalars *sc, double *partial,
                                                        unsigned int *counter)
{
    // converged (or exactly solved: small systems reach r = 0 in a few steps) -> no further update, no 0/0
    const bool live = sc->rr > sc->target && sc->pq != 0.;
    const double alpha = live ? sc->rz[P] / sc->pq : 0.;
    double v[2] = {0., 0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (fixed[i]) continue;
        const double xi = x[i] + alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        x[i] = xi; r[i] = ri;
        if (!MG) {
            const double zi = ri / diag[i];
            z[i] = zi;
            v[0] += ri * zi;
        }
        v[1] += ri * ri;
    }
    if (block_reduce_to_partials<2>(v, partial, counter, sc->tmp, 0) && !DIST) {
        // the one finalising thread: publish r.z for the next step, r.r, and count the step if it was a real one
        if (!MG) sc->rz[1 - P] = sc->tmp[0];
        if (live) { sc->rr = sc->tmp[1]; sc->iters += 1.; }
    }
}

// distributed CG: what the finalising thread of cg_update_kernel does, after sc->tmp has been summed over the ranks
template <int P>
__global__ void cg_publish_kernel(CgScalars *sc)
{
    const bool live = sc->rr > sc->target && sc->pq != 0.;
    sc->rz[1 - P] = sc->tmp[0];
    if (live) { sc->rr = sc->tmp[1]; sc->iters += 1.; }
}

// mask of the distributed CG: Dirichlet dofs and the dofs of node columns owned by a neighbour rank
__global__ void __launch_bounds__(256) cg_ghost_mask_kernel(int64_t n, int NnY, int own_first, int own_end,
                                                            const unsigned char *__restrict__ fixed, unsigned char *__restrict__ mask)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)((i >> 1) / NnY);
        mask[i] = (fixed[i] || j < own_first || j >= own_end) ? 1 : 0;
    }
}

// r.z after the V-cycle -> rz[SLOT]
// DIST: this rank's sum stays in sc->tmp[0]; it is all-reduced (together with r.r in tmp[1]) and published by
// cg_publish_kernel
template <int SLOT, bool DIST = false>
__global__ void __launch_bounds__(256) cg_rz_kernel(int64_t n, const double *__restrict__ r, const double *__restrict__ z,
                                                    const unsigned char *__restrict__ fixed, CgScalars *sc, double *partial,
                                                    unsigned int *counter)
{
    double v[1] = {0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!fixed[i]) v[0] += r[i] * z[i];
    if (block_reduce_to_partials<1>(v, partial, counter, sc->tmp, 0) && !DIST) sc->rz[SLOT] = sc->tmp[0];
}

// distributed start with the multigrid preconditioner: r.z of z = V(r), summed over the ranks, becomes rz[0]
__global__ void cg_publish_init_kernel(CgScalars *sc) { sc->rz[0] = sc->tmp[0]; }

// beta = rz_new / rz_old ; p = z + beta p
template <int P>
__global__ void __launch_bounds__(256) cg_p_kernel(int64_t n, double *__restrict__ p, const double *__restrict__ z,
                                                   const CgScalars *sc)
{
    const double beta = (sc->rr <= sc->target || sc->rz[P] == 0.) ? 0. : sc->rz[1 - P] / sc->rz[P];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = z[i] + beta * p[i];
}

// init: r = b (masked), z = r / diag, p = z, x = 0 ; r.z -> rz[0], r.r -> rr ; target = rtol^2 r.r ; iters = 0
__global__ void __launch_bounds__(256) cg_init_kernel(int64_t n, const double *__restrict__ b, double *__restrict__ x,
                                                      double *__restrict__ r, double *__restrict__ z, double *__restrict__ p,
                                                      const double *__restrict__ diag, const unsigned char *__restrict__ fixed,
                                                      CgScalars *sc, double rtol, double *partial, unsigned int *counter)
{
    double v[2] = {0., 0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double ri = fixed[i] ? 0. : b[i];
        double zi = fixed[i] ? 0. : ri / diag[i];
        x[i] = 0.; r[i] = ri; z[i] = zi; p[i] = zi;
        v[0] += ri * zi;
        v[1] += ri * ri;
    }
    if (block_reduce_to_partials<2>(v, partial, counter, sc->tmp, 0)) {
        sc->rz[0] = sc->tmp[0];
        sc->rz[1] = 0.;
        sc->rr = sc->tmp[1];
        sc->target = rtol * rtol * sc->tmp[1];
        sc->iters = 0.;
        sc->pq = 0.;
    }
}

__global__ void __launch_bounds__(256) diag_kernel(int64_t nrows, const int32_t *__restrict__ rowptr,
                                                   const int32_t *__restrict__ col, const double *__restrict__ val,
                                                   double *__restrict__ diag)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double d = 1.;
    for (int32_t i = rowptr[r]; i < rowptr[r + 1]; i++)
        if (col[i] == r) d = val[i];
    diag[r] = (d != 0.) ? d : 1.;
}

// du = x on free dofs, du_bc on fixed dofs
__global__ void __launch_bounds__(256) merge_du_kernel(int64_t n, const double *__restrict__ x, const double *__restrict__ du_bc,
                                                       const unsigned char *__restrict__ fixed, double *__restrict__ du)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        du[i] = fixed[i] ? du_bc[i] : x[i];
}

__global__ void __launch_bounds__(256) axpy_kernel(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}

static inline unsigned spmv_grid(int64_t nrows, int sm)
{
    const int64_t g = (nrows + 31) / 32;   // 32 rows per block and pass
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm * 8));
}

static inline unsigned vec_grid(int64_t n, int sm)
{
    int64_t g = (n + 255) / 256;
    int64_t cap = (int64_t)sm * 8;
    return (unsigned)std::max<int64_t>(1, std::min(g, cap));
}

cudaError_t launch_commit(const LaunchCtx &lc, int64_t Nel, int64_t Ndof, int NY, const int32_t *desc, double *u, const double *du,
                          double *f, const int32_t *rowptr, const int32_t *col, const double *val, const double *elstiff,
                          double *sig, double *eps, double *epl, const double *res_sig, const double *res_depl,
                          const unsigned char *has_res, int64_t *launches)
{
    // u += du ; f += K du (1383-1384)
    axpy_kernel<<<vec_grid(Ndof, lc.sm_count), 256, 0, lc.stream>>>(Ndof, 1., du, u);
    spmv_kernel<<<spmv_grid(Ndof, lc.sm_count), 256, 0, lc.stream>>>(Ndof, rowptr, col, val, du, nullptr, f, nullptr, 1, nullptr,
                                                                          nullptr, nullptr);
    commit_elements_kernel<<<(unsigned)((Nel + 127) / 128), 128, 0, lc.stream>>>(Nel, NY, desc, u, du, elstiff, sig, eps, epl, res_sig,
                                                                                 res_depl, has_res);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_global_sums(const LaunchCtx &lc, int64_t Nel, const int32_t *desc, const double *sig, const double *eps,
                               const double *epl, double *partial, unsigned int *counter, double *out, int64_t *launches,
                               int64_t e_first)
{
    unsigned grid = vec_grid(Nel, lc.sm_count);
    global_sums_kernel<<<grid, 256, 0, lc.stream>>>(Nel, desc, sig, eps, epl, partial, counter, out, e_first);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_scf(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t Nel, int NY, const int32_t *desc, const double *sig,
                       const double *epl, const double *du, const double *elstiff, const double *d_sld, double *hval,
                       unsigned char *mult, double *partial, unsigned int *counter, double *d_out /*>= 4 doubles*/,
                       double *h_out /*pinned*/, double *scf, int64_t *launches, int64_t e_first, Comm *comm)
{
    scf_elements_kernel<<<(unsigned)((Nel + 127) / 128), 128, 0, lc.stream>>>(Nel, NY, desc, sig, epl, du, elstiff, hval, mult);
    (*launches)++;
    for (int i = 0; i < nmat; i++) {
        if (hmats[i].kind != PLF_MAT_SVC) continue;
        int blocks = lc.sm_count;
        int64_t need = (Nel + SVCW - 1) / SVCW;
        if (blocks > need) blocks = (int)need;
        cudaFuncSetAttribute(scf_svc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        scf_svc_kernel<<<blocks, SVCW * 32, sizeof(double) * (size_t)(hmats[i].svc_nf + 1) * (size_t)hmats[i].nsv_pad, lc.stream>>>(Nel, NY, i, desc, sig, epl, du, elstiff,
                                                                                                  d_sld, hval, mult);
        (*launches)++;
    }
    unsigned grid = vec_grid(Nel, lc.sm_count);
    const double two = 2.;
    cudaMemcpyAsync(d_out + 3, &two, sizeof(double), cudaMemcpyHostToDevice, lc.stream);   // min accumulator
    scf_moments_kernel<<<grid, 256, 0, lc.stream>>>(Nel, hval, mult, 0., 0, partial, counter, d_out, d_out + 3, e_first);
    (*launches)++;
    if (comm) {   // count and sum over all ranks, minimum over all ranks: every rank then derives the same factor
        if (comm_allreduce_op(*comm, d_out, 2, 0, lc.stream) || comm_allreduce_op(*comm, d_out + 3, 1, 2, lc.stream)) return cudaErrorUnknown;
    }
    cudaError_t e = cudaMemcpyAsync(h_out, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, lc.stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
    const double cnt = h_out[0], sum = h_out[1], mn = h_out[3];
    if (cnt == 0.) { *scf = 1.; return cudaSuccess; }   // sc_list = [1.] (1056): std = 0 -> min = 1
    const double mean = sum / cnt;
    scf_moments_kernel<<<grid, 256, 0, lc.stream>>>(Nel, hval, mult, mean, 1, partial, counter, d_out, d_out + 3, e_first);
    (*launches)++;
    if (comm && comm_allreduce_op(*comm, d_out, 1, 0, lc.stream)) return cudaErrorUnknown;
    if ((e = cudaMemcpyAsync(h_out, d_out, sizeof(double), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
    const double sd = sqrt(h_out[0] / cnt);   // np.std (population)
    double s = (sd < 0.1) ? mn : fmax(1.e-3, mean - sd);   // 1057-1062
    if (s < 1.e-3) s = 1.e-3;                               // 1063-1066
    *scf = s;
    return cudaGetLastError();
}

// Jacobi-preconditioned CG on the masked operator.  b (rhs on free dofs) must already be in d_b.
// One iteration = 3 launches (SpMV + p.q, update + r.z + r.r, direction); all scalars stay on the device.  A chunk of
// `check_every` iterations is captured once into a CUDA graph (launch-bound loop: the kernels of a 10^5-dof system
// run for a few microseconds each) and replayed; the host only reads the residual after every chunk.  Iterations
// after convergence inside a chunk are frozen on the device (alpha = beta = 0), so the result does not depend on
// the chunk length.
// warm start: x0 = previous solution on the free dofs (0 on the fixed ones)
__global__ void __launch_bounds__(256) cg_mask_x_kernel(int64_t n, double *__restrict__ x, const unsigned char *__restrict__ fixed)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (fixed[i] || !isfinite(x[i])) x[i] = 0.;
}

// warm start: r = b - A x0 (q holds the masked product A x0), z = r / diag, p = z; r.z -> rz[0], r.r -> rr;
// the convergence target stays relative to |b|: target = rtol^2 b.b
__global__ void __launch_bounds__(256) cg_init_warm_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ q,
                                                           double *__restrict__ r, double *__restrict__ z, double *__restrict__ p,
                                                           const double *__restrict__ diag, const unsigned char *__restrict__ fixed,
                                                           CgScalars *sc, double rtol, double *partial, unsigned int *counter)
{
    double v[3] = {0., 0., 0.};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = fixed[i] ? 0. : b[i];
        const double ri = fixed[i] ? 0. : bi - q[i];
        const double zi = fixed[i] ? 0. : ri / diag[i];
        r[i] = ri; z[i] = zi; p[i] = zi;
        v[0] += ri * zi;
        v[1] += ri * ri;
        v[2] += bi * bi;
    }
    __shared__ double s_out[3];
    if (block_reduce_to_partials<3>(v, partial, counter, s_out, 0)) {
        sc->rz[0] = s_out[0];
        sc->rz[1] = 0.;
        sc->rr = s_out[1];
        sc->tmp[1] = s_out[2];   // b.b, read by the host as the reference norm
        sc->target = rtol * rtol * s_out[2];
        sc->iters = 0.;
        sc->pq = 0.;
    }
}


cudaError_t launch_cg(const LaunchCtx &lc, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                      const unsigned char *fixed, const double *b, double *x, double *r, double *z, double *p, double *q,
                      double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc, double rtol, int max_iter,
                      int check_every, int *iters, double *relres, int64_t *launches, MgHierarchy *mg, int warm, int NX,
                      int NY, CgGraphCache &g_cg, int mg_setup_needed)
{
    CgScalars *sc = (CgScalars *)d_sc;
    CgScalars *hs = (CgScalars *)h_sc;
    const unsigned vg = vec_grid(n, lc.sm_count);
    const unsigned sg = spmv_grid(n, lc.sm_count);
    cudaError_t e;
    check_every = std::max(2, check_every & ~1);   // even: the rz ping-pong returns to parity 0 at chunk end
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(n, rowptr, col, val, diag);
    if (warm) {
        // x keeps the previous solution (the Newton-like passes of an increment solve nearly the same system again)
        cg_mask_x_kernel<<<vg, 256, 0, lc.stream>>>(n, x, fixed);
        spmv_kernel<<<sg, 256, 0, lc.stream>>>(n, rowptr, col, val, x, nullptr, q, fixed, 0, nullptr, nullptr, nullptr);
        cg_init_warm_kernel<<<vg, 256, 0, lc.stream>>>(n, b, q, r, z, p, diag, fixed, sc, rtol, partial, counter);
        *launches += 4;
    } else {
        cg_init_kernel<<<vg, 256, 0, lc.stream>>>(n, b, x, r, z, p, diag, fixed, sc, rtol, partial, counter);
        *launches += 2;
    }
    if (mg) {
        // z = V(r), r.z, p = z with the multigrid preconditioner (the Jacobi z / p / r.z of cg_init_kernel are replaced)
        if (mg_setup_needed && (e = mg_setup(lc, *mg, val, fixed, launches)) != cudaSuccess) return e;
        if ((e = mg_vcycle(lc, *mg, r, z, launches)) != cudaSuccess) return e;
        cg_rz_kernel<0><<<vg, 256, 0, lc.stream>>>(n, r, z, fixed, sc, partial, counter);
        (*launches)++;
        if ((e = cudaMemcpyAsync(p, z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, lc.stream)) != cudaSuccess) return e;
    }
    if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
    const double bb = warm ? hs->tmp[1] : hs->rr;   // |b|^2: the residual norm is reported relative to it
    *iters = 0;
    *relres = 0.;
    if (bb == 0.) {
        if (warm) cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), lc.stream);   // b = 0: the solution is 0
        return cudaSuccess;
    }
    const double target = hs->target;
    if (warm && !(hs->rr > target)) {   // the previous solution already satisfies the tolerance
        *relres = sqrt(hs->rr / bb);
        return cudaSuccess;
    }

    int64_t chunk_launches = 0;
    auto enqueue_chunk = [&]() {
        for (int c = 0; c < check_every; c += 2) {
            if (NX > 0) stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, fixed, partial, counter, &sc->pq, 0, n);
            else spmv_kernel<<<sg, 256, 0, lc.stream>>>(n, rowptr, col, val, p, nullptr, q, fixed, 0, partial, counter, &sc->pq);
            if (mg) {
                cg_update_kernel<0, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, fixed, sc, partial, counter);
                mg_vcycle(lc, *mg, r, z, &chunk_launches);
                cg_rz_kernel<1><<<vg, 256, 0, lc.stream>>>(n, r, z, fixed, sc, partial, counter);
                chunk_launches++;
            } else {
                cg_update_kernel<0, false><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, fixed, sc, partial, counter);
            }
            cg_p_kernel<0><<<vg, 256, 0, lc.stream>>>(n, p, z, sc);
            if (NX > 0) stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, fixed, partial, counter, &sc->pq, 0, n);
            else spmv_kernel<<<sg, 256, 0, lc.stream>>>(n, rowptr, col, val, p, nullptr, q, fixed, 0, partial, counter, &sc->pq);
            if (mg) {
                cg_update_kernel<1, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, fixed, sc, partial, counter);
                mg_vcycle(lc, *mg, r, z, &chunk_launches);
                cg_rz_kernel<0><<<vg, 256, 0, lc.stream>>>(n, r, z, fixed, sc, partial, counter);
                chunk_launches++;
            } else {
                cg_update_kernel<1, false><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, fixed, sc, partial, counter);
            }
            cg_p_kernel<1><<<vg, 256, 0, lc.stream>>>(n, p, z, sc);
            chunk_launches += 6;
        }
    };
    const void *key[6] = {NX > 0 ? (const void *)val : (const void *)rowptr, val, x, fixed, d_sc,
                          mg ? (const void *)mg : (const void *)partial};
    bool cached = g_cg.exec && g_cg.n == n && g_cg.chunk == check_every && g_cg.stream == lc.stream;
    for (int i = 0; i < 6 && cached; i++) cached = (g_cg.key[i] == key[i]);
    if (!cached) {
        if (g_cg.exec) { cudaGraphExecDestroy(g_cg.exec); g_cg.exec = nullptr; }
        cudaGraph_t graph = nullptr;
        if ((e = cudaStreamBeginCapture(lc.stream, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return e;
        enqueue_chunk();
        if ((e = cudaStreamEndCapture(lc.stream, &graph)) != cudaSuccess) return e;
        e = cudaGraphInstantiate(&g_cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { g_cg.exec = nullptr; return e; }
        g_cg.n = n; g_cg.chunk = check_every; g_cg.stream = lc.stream; g_cg.launches_per_chunk = chunk_launches;
        for (int i = 0; i < 6; i++) g_cg.key[i] = key[i];
    }
    int it = 0;
    while (it < max_iter) {
        if ((e = cudaGraphLaunch(g_cg.exec, lc.stream)) != cudaSuccess) return e;
        *launches += g_cg.launches_per_chunk;
        it += check_every;
        if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
        *relres = sqrt(hs->rr / bb);
        if (!(hs->rr > target)) break;   // also stops on NaN
    }
    *iters = (int)hs->iters;
    return cudaGetLastError();
}

// Strip-partitioned Jacobi-preconditioned CG.  Every rank runs the same kernel sequence on its local
// node grid; `mask` hides the Dirichlet dofs and the ghost node columns, so all sums run over owned free dofs only.  Per
// iteration: ghost columns of p from the neighbours, q = K p on the owned rows (complete rows, see shard.py), p.q summed
// over the ranks, update, (r.z, r.r) summed over the ranks, direction.  All scalars stay on the device and are identical
// on every rank after each all-reduce, so every rank takes the same convergence decisions.  Enqueued eagerly (NCCL calls
// between the kernels); the host reads the residual every `check_every` iterations.
cudaError_t launch_cg_dist(const LaunchCtx &lc, Comm &comm, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val,
                           const unsigned char *mask, const double *b, double *x, double *r, double *z, double *p, double *q,
                           double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc, double rtol,
                           int max_iter, int check_every, int *iters, double *relres, int64_t *launches, int warm, int NX, int NY,
                           MgHierarchy *mg, int mg_setup_needed)
{
    // mg != nullptr: the preconditioner is this rank's own V-cycle on its strip with the ghost columns masked like
    // Dirichlet dofs -- a non-overlapping additive Schwarz method whose subdomain solves are multigrid cycles: symmetric
    // positive definite, no communication inside the preconditioner, and far fewer iterations than Jacobi (the count grows
    // mildly with the number of strips).  Per iteration: one halo column per neighbour and two all-reduces (p.q; r.z and r.r).
    CgScalars *sc = (CgScalars *)d_sc;
    CgScalars *hs = (CgScalars *)h_sc;
    const unsigned vg = vec_grid(n, lc.sm_count);
    cudaError_t e;
#define PLF_NC(call) do { if (call) return cudaErrorUnknown; } while (0)
    check_every = std::max(2, check_every & ~1);
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(n, rowptr, col, val, diag);
    if (warm) {
        cg_mask_x_kernel<<<vg, 256, 0, lc.stream>>>(n, x, mask);
        PLF_NC(comm_halo_exchange(comm, x, NX, NY, lc.stream));
        spmv_kernel<<<spmv_grid(n, lc.sm_count), 256, 0, lc.stream>>>(n, rowptr, col, val, x, nullptr, q, mask, 0, nullptr, nullptr, nullptr);
        cg_init_warm_kernel<<<vg, 256, 0, lc.stream>>>(n, b, q, r, z, p, diag, mask, sc, rtol, partial, counter);
        *launches += 4;
    } else {
        cg_init_kernel<<<vg, 256, 0, lc.stream>>>(n, b, x, r, z, p, diag, mask, sc, rtol, partial, counter);
        *launches += 2;
    }
    // pq (= 0), rz[2], tmp[2], rr, target: every one of them is a sum over dofs (target = rtol^2 b.b)
    PLF_NC(comm_allreduce_sum(comm, reinterpret_cast<double *>(sc), 7, lc.stream));
    if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
    const double bb = warm ? hs->tmp[1] : hs->rr;
    *iters = 0;
    *relres = 0.;
    if (bb == 0.) {
        if (warm) cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), lc.stream);
        return cudaSuccess;
    }
    const double target = hs->target;
    if (warm && !(hs->rr > target)) {
        *relres = sqrt(hs->rr / bb);
        return cudaSuccess;
    }
    if (mg) {
        if (mg_setup_needed && (e = mg_setup(lc, *mg, val, mask, launches)) != cudaSuccess) return e;
        if ((e = mg_vcycle(lc, *mg, r, z, launches)) != cudaSuccess) return e;
        cg_rz_kernel<0, true><<<vg, 256, 0, lc.stream>>>(n, r, z, mask, sc, partial, counter);
        PLF_NC(comm_allreduce_sum(comm, sc->tmp, 1, lc.stream));
        cg_publish_init_kernel<<<1, 1, 0, lc.stream>>>(sc);
        *launches += 2;
        if ((e = cudaMemcpyAsync(p, z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, lc.stream)) != cudaSuccess) return e;
    }
    int it = 0;
    while (it < max_iter) {
        for (int c = 0; c < check_every; c += 2) {
            PLF_NC(comm_halo_exchange(comm, p, NX, NY, lc.stream));
            stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, mask, partial, counter, &sc->pq, 0, n);
            PLF_NC(comm_allreduce_sum(comm, &sc->pq, 1, lc.stream));
            if (mg) {
                cg_update_kernel<0, true, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, mask, sc, partial, counter);
                if ((e = mg_vcycle(lc, *mg, r, z, launches)) != cudaSuccess) return e;
                cg_rz_kernel<1, true><<<vg, 256, 0, lc.stream>>>(n, r, z, mask, sc, partial, counter);
                (*launches)++;
            } else {
                cg_update_kernel<0, false, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, mask, sc, partial, counter);
            }
            PLF_NC(comm_allreduce_sum(comm, sc->tmp, 2, lc.stream));
            cg_publish_kernel<0><<<1, 1, 0, lc.stream>>>(sc);
            cg_p_kernel<0><<<vg, 256, 0, lc.stream>>>(n, p, z, sc);
            PLF_NC(comm_halo_exchange(comm, p, NX, NY, lc.stream));
            stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, mask, partial, counter, &sc->pq, 0, n);
            PLF_NC(comm_allreduce_sum(comm, &sc->pq, 1, lc.stream));
            if (mg) {
                cg_update_kernel<1, true, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, mask, sc, partial, counter);
                if ((e = mg_vcycle(lc, *mg, r, z, launches)) != cudaSuccess) return e;
                cg_rz_kernel<0, true><<<vg, 256, 0, lc.stream>>>(n, r, z, mask, sc, partial, counter);
                (*launches)++;
            } else {
                cg_update_kernel<1, false, true><<<vg, 256, 0, lc.stream>>>(n, x, r, z, p, q, diag, mask, sc, partial, counter);
            }
            PLF_NC(comm_allreduce_sum(comm, sc->tmp, 2, lc.stream));
            cg_publish_kernel<1><<<1, 1, 0, lc.stream>>>(sc);
            cg_p_kernel<1><<<vg, 256, 0, lc.stream>>>(n, p, z, sc);
            *launches += 8;
        }
        it += check_every;
        if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
        *relres = sqrt(hs->rr / bb);
        if (!(hs->rr > target)) break;
    }
#undef PLF_NC
    *iters = (int)hs->iters;
    return cudaGetLastError();
}

// CG of the replicated solve with level 0 shared among the ranks (plf_comm_set_replica, option solver_hybrid).  Every rank
// holds the GLOBAL matrix, hierarchy and vectors, but updates only the rows of the node columns [a, b) it owns: per iteration
// one ghost column of p per neighbour, q = K p and p.q on the owned rows, all-reduce, update of x and r on the owned rows,
// the V-cycle (mg_vcycle with MgDist: its level 0 split the same way, the levels below on every rank alike), all-reduce of
// (r.z, r.r), direction.  Same iterates as the single-GPU solver up to the summation order of the dot products.  `mask` =
// fixed | not owned, so the vector kernels of the single-GPU solver apply unchanged to the owned range.
cudaError_t launch_cg_hybrid(const LaunchCtx &lc, Comm &comm, int a_col, int b_col, int64_t n, const int32_t *rowptr, const int32_t *col,
                             const double *val, const unsigned char *mask, const double *b, double *x, double *r, double *z,
                             double *p, double *q, double *diag, double *partial, unsigned int *counter, void *d_sc, void *h_sc,
                             double rtol, int max_iter, int *iters, double *relres, int64_t *launches, int warm, int NX, int NY,
                             MgHierarchy &mg, const MgDist &md, int mg_setup_needed, const unsigned char *fixed)
{
    CgScalars *sc = (CgScalars *)d_sc;
    CgScalars *hs = (CgScalars *)h_sc;
    const int64_t r0 = 2ll * a_col * (NY + 1), r1 = 2ll * b_col * (NY + 1), m = r1 - r0;   // owned rows
    const unsigned vg = vec_grid(m, lc.sm_count);
    cudaError_t e;
#define PLF_NC(call) do { if (call) return cudaErrorUnknown; } while (0)
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, lc.stream>>>(n, rowptr, col, val, diag);
    if (warm) {
        cg_mask_x_kernel<<<vg, 256, 0, lc.stream>>>(m, x + r0, mask + r0);
        PLF_NC(comm_halo_exchange_cols(comm, x, NY, a_col, b_col, NX, lc.stream));
        stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, x, q, mask, partial, counter, &sc->pq, r0, r1);
        cg_init_warm_kernel<<<vg, 256, 0, lc.stream>>>(m, b + r0, q + r0, r + r0, z + r0, p + r0, diag + r0, mask + r0, sc, rtol, partial, counter);
        *launches += 4;
    } else {
        cg_init_kernel<<<vg, 256, 0, lc.stream>>>(m, b + r0, x + r0, r + r0, z + r0, p + r0, diag + r0, mask + r0, sc, rtol, partial, counter);
        *launches += 2;
    }
    PLF_NC(comm_allreduce_sum(comm, reinterpret_cast<double *>(sc), 7, lc.stream));
    if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
    const double bb = warm ? hs->tmp[1] : hs->rr;
    *iters = 0;
    *relres = 0.;
    if (bb == 0.) {
        if (warm) cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), lc.stream);
        return cudaSuccess;
    }
    const double target = hs->target;
    if (warm && !(hs->rr > target)) {
        *relres = sqrt(hs->rr / bb);
        return cudaSuccess;
    }
    // the hierarchy is the global one on every rank (masks, Galerkin products, inverse diagonals from the gathered K)
    if (mg_setup_needed && (e = mg_setup(lc, mg, val, fixed, launches)) != cudaSuccess) return e;
    if ((e = mg_vcycle(lc, mg, r, z, launches, &md)) != cudaSuccess) return e;
    cg_rz_kernel<0, true><<<vg, 256, 0, lc.stream>>>(m, r + r0, z + r0, mask + r0, sc, partial, counter);
    PLF_NC(comm_allreduce_sum(comm, sc->tmp, 1, lc.stream));
    cg_publish_init_kernel<<<1, 1, 0, lc.stream>>>(sc);
    *launches += 2;
    if ((e = cudaMemcpyAsync(p + r0, z + r0, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, lc.stream)) != cudaSuccess) return e;
    int it = 0;
    const int check_every = 4;
    while (it < max_iter) {
        for (int c = 0; c < check_every; c += 2) {
            PLF_NC(comm_halo_exchange_cols(comm, p, NY, a_col, b_col, NX, lc.stream));
            stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, mask, partial, counter, &sc->pq, r0, r1);
            PLF_NC(comm_allreduce_sum(comm, &sc->pq, 1, lc.stream));
            cg_update_kernel<0, true, true><<<vg, 256, 0, lc.stream>>>(m, x + r0, r + r0, z + r0, p + r0, q + r0, diag + r0, mask + r0, sc, partial, counter);
            if ((e = mg_vcycle(lc, mg, r, z, launches, &md)) != cudaSuccess) return e;
            cg_rz_kernel<1, true><<<vg, 256, 0, lc.stream>>>(m, r + r0, z + r0, mask + r0, sc, partial, counter);
            PLF_NC(comm_allreduce_sum(comm, sc->tmp, 2, lc.stream));
            cg_publish_kernel<0><<<1, 1, 0, lc.stream>>>(sc);
            cg_p_kernel<0><<<vg, 256, 0, lc.stream>>>(m, p + r0, z + r0, sc);
            PLF_NC(comm_halo_exchange_cols(comm, p, NY, a_col, b_col, NX, lc.stream));
            stencil_spmv_dot_kernel<<<vg, 256, 0, lc.stream>>>(NX, NY, val, p, q, mask, partial, counter, &sc->pq, r0, r1);
            PLF_NC(comm_allreduce_sum(comm, &sc->pq, 1, lc.stream));
            cg_update_kernel<1, true, true><<<vg, 256, 0, lc.stream>>>(m, x + r0, r + r0, z + r0, p + r0, q + r0, diag + r0, mask + r0, sc, partial, counter);
            if ((e = mg_vcycle(lc, mg, r, z, launches, &md)) != cudaSuccess) return e;
            cg_rz_kernel<0, true><<<vg, 256, 0, lc.stream>>>(m, r + r0, z + r0, mask + r0, sc, partial, counter);
            PLF_NC(comm_allreduce_sum(comm, sc->tmp, 2, lc.stream));
            cg_publish_kernel<1><<<1, 1, 0, lc.stream>>>(sc);
            cg_p_kernel<1><<<vg, 256, 0, lc.stream>>>(m, p + r0, z + r0, sc);
            *launches += 10;
        }
        it += check_every;
        if ((e = cudaMemcpyAsync(hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, lc.stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(lc.stream)) != cudaSuccess) return e;
        *relres = sqrt(hs->rr / bb);
        if (!(hs->rr > target)) break;
    }
#undef PLF_NC
    *iters = (int)hs->iters;
    // ghost columns of the solution from their owners: the strip of du a rank copies out covers its halo / ghost columns
    if (comm_halo_exchange_cols(comm, x, NY, a_col, b_col, NX, lc.stream)) return cudaErrorUnknown;
    return cudaGetLastError();
}

cudaError_t launch_ghost_mask(const LaunchCtx &lc, int64_t n, int NY, int own_first, int own_end, const unsigned char *fixed,
                              unsigned char *mask, int64_t *launches)
{
    cg_ghost_mask_kernel<<<vec_grid(n, lc.sm_count), 256, 0, lc.stream>>>(n, NY + 1, own_first, own_end, fixed, mask);
    (*launches)++;
    return cudaGetLastError();
}

void cg_graph_reset(CgGraphCache &g_cg)
{
    if (g_cg.exec) cudaGraphExecDestroy(g_cg.exec);
    g_cg = CgGraphCache();
}

cudaError_t launch_spmv(const LaunchCtx &lc, int64_t n, const int32_t *rowptr, const int32_t *col, const double *val, const double *x,
                        const double *rhs, double *y, const unsigned char *masked, int mode, int64_t *launches)
{
    spmv_kernel<<<spmv_grid(n, lc.sm_count), 256, 0, lc.stream>>>(n, rowptr, col, val, x, rhs, y, masked, mode, nullptr, nullptr,
                                                                       nullptr);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_merge_du(const LaunchCtx &lc, int64_t n, const double *x, const double *du_bc, const unsigned char *fixed, double *du,
                            int64_t *launches)
{
    merge_du_kernel<<<vec_grid(n, lc.sm_count), 256, 0, lc.stream>>>(n, x, du_bc, fixed, du);
    (*launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// boundary-condition vectors and boundary sums of the load-increment controller (plf_solve.cu)
// ---------------------------------------------------------------------------------------------------------------
// calc_BC (model.py:1070-1206) touches boundary dofs only: the host builds the entries, this kernel scatters them
// into the (zeroed) dense vectors the solver reads.
__global__ void __launch_bounds__(256) bc_scatter_kernel(int n, const int32_t *__restrict__ dof, const unsigned char *__restrict__ fx,
                                                         const double *__restrict__ vals /* [3][n]: du, w, fext */,
                                                         unsigned char *__restrict__ fixed, double *__restrict__ du,
                                                         double *__restrict__ w, double *__restrict__ fext)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t d = dof[i];
    fixed[d] = fx[i];
    du[d] = vals[i];
    w[d] = vals[n + i];
    fext[d] = vals[2 * n + i];
}

// Model.bcval (model.py:1452-1471) sums for the four edges: block b = edge (left, right, bottom, top), out[4 b + {0..3}] =
// sum ux, sum uy, sum fx, sum fy over the edge's nodes (bottom / top: node columns [c0, c1) only -- the columns a rank
// owns).  Fixed summation order.
__global__ void __launch_bounds__(256) edge_sums_kernel(int NX, int NY, int c0, int c1, const double *__restrict__ u,
                                                        const double *__restrict__ f, double *__restrict__ out)
{
    __shared__ double sh[4][8];
    const int edge = blockIdx.x, NnY = NY + 1;
    const int cnt = edge < 2 ? NnY : (c1 - c0);
    double v[4] = {0., 0., 0., 0.};
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        int64_t node;
        if (edge == 0) node = i;
        else if (edge == 1) node = (int64_t)NX * NnY + i;
        else if (edge == 2) node = (int64_t)(c0 + i) * NnY;
        else node = (int64_t)(c0 + i) * NnY + NY;
        const double2 uu = *reinterpret_cast<const double2 *>(u + 2 * node);
        const double2 ff = *reinterpret_cast<const double2 *>(f + 2 * node);
        v[0] += uu.x; v[1] += uu.y; v[2] += ff.x; v[3] += ff.y;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[q][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double x = 0.;
        for (int w = 0; w < 8; w++) x += sh[threadIdx.x][w];
        out[4 * edge + threadIdx.x] = x;
    }
}

cudaError_t launch_bc_scatter(const LaunchCtx &lc, int n, const int32_t *dof, const unsigned char *fx, const double *vals,
                              unsigned char *fixed, double *du, double *w, double *fext, int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    bc_scatter_kernel<<<(n + 255) / 256, 256, 0, lc.stream>>>(n, dof, fx, vals, fixed, du, w, fext);
    (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_edge_sums(const LaunchCtx &lc, int NX, int NY, int c0, int c1, const double *u, const double *f, double *out,
                             int64_t *launches)
{
    edge_sums_kernel<<<4, 256, 0, lc.stream>>>(NX, NY, c0, c1, u, f, out);
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace plf
