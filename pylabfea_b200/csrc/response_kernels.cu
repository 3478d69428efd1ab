// response_kernels.cu -- K1/K2: batched Material.response (material.py:207-346) and the batched user-facing
// constitutive calls (calc_seq / calc_yf / calc_fgrad / ML_full_yf / epl_dot / C_tan) for sm_100a.
//
// K1 (Hill-6p / J2): one thread per element.  The 6-component sig / epl / deps records are AoS in HBM (48 B per
//    element); a block stages its 128 records through shared memory with 128-bit loads so that every HBM
//    transaction is a fully used 128 B line, and stages the 36-double tangent the same way on the way out.
// K2 (SVC): one warp per element, support-vector table resident in shared memory (SoA), dynamic element queue.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "hill_fast.cuh"
#include "plf_device.cuh"
#include "plf_host.h"

namespace plf {

__constant__ DevMat c_mat[PLF_MAX_MATERIALS];

cudaError_t response_upload_materials(const DevMat *h, int n, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_mat, h, sizeof(DevMat) * n, 0, cudaMemcpyHostToDevice, st);
}

constexpr int RB = 128;       // threads (= elements) per block, K1
constexpr int CT_STRIDE = 37; // padded shared-memory stride of one tangent record (odd -> conflict free)

__device__ __forceinline__ void expand_sym(const double *gs, double *out, int stride1)
{
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) out[(6 * i + j) * stride1] = gs[sym_idx(i, j)];
}

// ------------------------------------------------------------------------------------------------------------
// K1: thread-per-element return mapping.
//   response_passA_kernel  : staged AoS I/O; elastic + single-step elements finish, 50-sub-step elements are
//                            appended (warp-aggregated) to the work list
//   response_passB_kernel  : persistent; runs the 50 sub-steps for the listed elements (gathered I/O); reads the
//                            list length from device memory, so no host round trip sits between A and B
//   response_generic_kernel: the unsplit reference formulation (response_core), "compact = 0"
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_in(const double *__restrict__ sig, const double *__restrict__ epl,
                                         const double *__restrict__ deps, int64_t base, int cnt, double *s_sig, double *s_epl,
                                         double *s_dep, double *s, double *ep, double *de)
{
    const int tid = threadIdx.x;
    // coalesced 128-bit loads: cnt*3 double2 per array
    const double2 *g0 = reinterpret_cast<const double2 *>(sig + base * 6);
    const double2 *g1 = reinterpret_cast<const double2 *>(epl + base * 6);
    const double2 *g2 = reinterpret_cast<const double2 *>(deps + base * 6);
    for (int i = tid; i < cnt * 3; i += RB) {
        reinterpret_cast<double2 *>(s_sig)[i] = __ldg(g0 + i);
        reinterpret_cast<double2 *>(s_epl)[i] = __ldg(g1 + i);
        reinterpret_cast<double2 *>(s_dep)[i] = __ldg(g2 + i);
    }
    __syncthreads();
    if (tid < cnt) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double2 a = reinterpret_cast<double2 *>(s_sig)[3 * tid + k];
            double2 b = reinterpret_cast<double2 *>(s_epl)[3 * tid + k];
            double2 c = reinterpret_cast<double2 *>(s_dep)[3 * tid + k];
            s[2 * k] = a.x; s[2 * k + 1] = a.y;
            ep[2 * k] = b.x; ep[2 * k + 1] = b.y;
            de[2 * k] = c.x; de[2 * k + 1] = c.y;
        }
    }
}

__device__ __forceinline__ void stage_out(int64_t base, int cnt, const double *s_sig, const double *s_epl, const double *s_ct,
                                          const unsigned char *s_done, double *__restrict__ sig_out,
                                          double *__restrict__ depl_out, double *__restrict__ ctan)
{
    const int tid = threadIdx.x;
    // coalesced 128-bit stores of the staged records (skipping deferred / foreign elements)
    for (int i = tid; i < cnt * 3; i += RB) {
        int el = i / 3;
        if (!s_done[el]) continue;
        if (sig_out) reinterpret_cast<double2 *>(sig_out + base * 6)[i] = reinterpret_cast<const double2 *>(s_sig)[i];
        if (depl_out) reinterpret_cast<double2 *>(depl_out + base * 6)[i] = reinterpret_cast<const double2 *>(s_epl)[i];
    }
    if (ctan) {
        double2 *g = reinterpret_cast<double2 *>(ctan + base * 36);
        for (int i = tid; i < cnt * 18; i += RB) {
            int el = i / 18, k2 = i - el * 18;
            if (!s_done[el]) continue;
            const double *p = s_ct + CT_STRIDE * el + 2 * k2;
            g[i] = make_double2(p[0], p[1]);
        }
    }
}

// SINGLE: the batch uses one material -> c_mat[0] is addressed with immediate constant-bank operands (no registers
// hold material parameters, no indexed constant loads)
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 256-bit global store (sm_100: STG.E.ENL2.256): one full 32 B sector per thread, so a thread can write its own
// 288 B tangent record (9 sectors, 32 B aligned) straight from registers without partial-sector traffic
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// prefetch of one input tile (3 x RB records of 48 B) into shared memory
__device__ __forceinline__ void tile_prefetch(int64_t n, int64_t tile, double *d, const double *__restrict__ sig,
                                              const double *__restrict__ epl, const double *__restrict__ deps)
{
    const int64_t b0 = tile * RB;
    const int c = (int)min((int64_t)RB, n - b0);
    const double2 *g0 = reinterpret_cast<const double2 *>(sig + b0 * 6);
    const double2 *g1 = reinterpret_cast<const double2 *>(epl + b0 * 6);
    const double2 *g2 = reinterpret_cast<const double2 *>(deps + b0 * 6);
    for (int i = threadIdx.x; i < c * 3; i += RB) {
        cp_async16(reinterpret_cast<double2 *>(d) + i, g0 + i);
        cp_async16(reinterpret_cast<double2 *>(d + RB * 6) + i, g1 + i);
        cp_async16(reinterpret_cast<double2 *>(d + 2 * RB * 6) + i, g2 + i);
    }
    cp_async_commit();
}

// Pass A on one tile whose inputs are resident in shared memory (s_in = [3][RB*6]).
//   DIRECT = false: results are staged through shared memory (tangent in s_ct, stride 37) and written with coalesced
//                   128-bit stores by the whole block
//   DIRECT = true : every thread writes its own records from registers (tangent: 9 x 256-bit stores); no tangent
//                   staging buffer, no copy loop, one barrier less -- ~300 fewer instructions per element
template <bool SINGLE, int FM, bool DIRECT>
__device__ __forceinline__ void passA_tile(int64_t n, int64_t tile, double *s_in, double *s_ct, unsigned char *s_done,
                                           const int32_t *__restrict__ mat_id, double *__restrict__ fy,
                                           double *__restrict__ sig_out, double *__restrict__ depl_out,
                                           double *__restrict__ ctan, int32_t *__restrict__ nsteps, int32_t *worklist,
                                           unsigned int *work_count)
{
    constexpr int FA = (FM == 6 || FM == 7) ? 1 : (FM >= 3 ? 2 : FM);   // (modes >= 3 have their own pass A functions)
    const int tid = threadIdx.x;
    double *s_sig = s_in, *s_epl = s_in + RB * 6, *s_dep = s_in + 2 * RB * 6;
    const int64_t base = tile * RB;
    const int cnt = (int)min((int64_t)RB, n - base);
    double s[6], ep[6], de[6];
    if (tid < cnt) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double2 a = reinterpret_cast<double2 *>(s_sig)[3 * tid + k];
            double2 b = reinterpret_cast<double2 *>(s_epl)[3 * tid + k];
            double2 c = reinterpret_cast<double2 *>(s_dep)[3 * tid + k];
            s[2 * k] = a.x; s[2 * k + 1] = a.y;
            ep[2 * k] = b.x; ep[2 * k + 1] = b.y;
            de[2 * k] = c.x; de[2 * k + 1] = c.y;
        }
    }
    bool done = false, defer = false, general = false;
    const int64_t e = base + tid;
    if (tid < cnt) {
        const DevMat &m = c_mat[SINGLE ? 0 : mat_id[e]];
        if (FM == 4 || FM == 6 || FM == 7) general = !planar_state(s, ep, de);   // planar states go to the front of the list
        if (m.kind == PLF_MAT_ELASTIC || m.kind == PLF_MAT_HILL) {
            double dp[6], ca[6], tc = 1., tw = 0., fyv = 0.;
            int br = 0;
            if (m.kind == PLF_MAT_ELASTIC) {
                // Model.solve never calls response for elastic materials (model.py:1341); elastic update for completeness
                double t[6];
                cv_mul(m, de, t);
#pragma unroll
                for (int k = 0; k < 6; k++) { s[k] += t[k]; dp[k] = 0.; ca[k] = 0.; }
            } else {
                if (FM == 5) {   // diagnostics only (k1_only = 2): no constitutive math, same data movement
                    br = 1;
#pragma unroll
                    for (int k = 0; k < 6; k++) { dp[k] = ep[k]; ca[k] = de[k]; }
                    fyv = s[0];
                } else
                br = (FM == 7) ? hillblk_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                   : (FM == 6) ? hilliso_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                   : (FM >= 3) ? iso_passA(m, s, ep, de, dp, ca, tc, tw, fyv) : hill_passA<FA>(m, s, ep, de, dp, ca, tc, tw, fyv);
            }
            if (br < 0) {
                defer = true;
            } else {
                done = true;
                if (fy) fy[e] = fyv;
                if (nsteps) nsteps[e] = 0;
                if (DIRECT) {
                    if (sig_out)
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            reinterpret_cast<double2 *>(sig_out + e * 6)[k] = make_double2(s[2 * k], s[2 * k + 1]);
                    if (depl_out)
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            reinterpret_cast<double2 *>(depl_out + e * 6)[k] = make_double2(dp[2 * k], dp[2 * k + 1]);
                    if (ctan) {
#pragma unroll
                        for (int g = 0; g < 9; g++) {
                            double v[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int idx = 4 * g + q, i = idx / 6, j = idx - 6 * i;
                                v[q] = m.CV[idx] * tc - (ca[i] * ca[j]) * tw;   // (ca_i ca_j) commutes: exactly symmetric like the reference's kron
                            }
                            st256(ctan + e * 36 + 4 * g, v[0], v[1], v[2], v[3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 6; k++) { s_sig[6 * tid + k] = s[k]; s_epl[6 * tid + k] = dp[k]; }
                    double *o = s_ct + CT_STRIDE * tid;
#pragma unroll
                    for (int i = 0; i < 6; i++)
#pragma unroll
                        for (int j = i; j < 6; j++) {
                            // ca_i ca_j first: the dyad is exactly symmetric, as np.kron(ca, ca) is in C_tan (material.py:1083-1085)
                            const double pw = (ca[i] * ca[j]) * tw;
                            o[6 * i + j] = m.CV[6 * i + j] * tc - pw;
                            if (j > i) o[6 * j + i] = m.CV[6 * j + i] * tc - pw;
                        }
                }
            }
        }
    }
    warp_append2(defer, general, (int32_t)e, worklist, n, work_count);
    if (!DIRECT) {
        s_done[tid] = done;
        __syncthreads();
        stage_out(base, cnt, s_sig, s_epl, s_ct, s_done, sig_out, depl_out, ctan);
    }
}

// Pass B for one deferred element: gathered 128-bit loads, the 50 sub-steps, direct stores.
//   xc: this thread's 6 shared-memory words (stride RB) for the radial-return correction sums of FM 3
template <bool SINGLE, int FM, class WAcc>
__device__ __forceinline__ void passB_element(int64_t e, WAcc &W, double *xc, const int32_t *__restrict__ mat_id,
                                              const double *__restrict__ sig, const double *__restrict__ epl,
                                              const double *__restrict__ deps, double *__restrict__ fy,
                                              double *__restrict__ sig_out, double *__restrict__ depl_out,
                                              double *__restrict__ ctan, int32_t *__restrict__ nsteps)
{
    const DevMat &m = c_mat[SINGLE ? 0 : mat_id[e]];
    double s[6], ep[6], de[6], dp[6], gs[21], fyv;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double2 a = __ldg(reinterpret_cast<const double2 *>(sig + e * 6) + k);
        double2 b = __ldg(reinterpret_cast<const double2 *>(epl + e * 6) + k);
        double2 c = __ldg(reinterpret_cast<const double2 *>(deps + e * 6) + k);
        s[2 * k] = a.x; s[2 * k + 1] = a.y;
        ep[2 * k] = b.x; ep[2 * k + 1] = b.y;
        de[2 * k] = c.x; de[2 * k + 1] = c.y;
    }
    if (FM == 7) hill_passB_blk2d(m, s, ep, de, xc, RB, gs, fyv);
    else if (FM == 6) hill_passB_hilliso2d(m, s, ep, de, xc, RB, gs, fyv);
    else if (FM == 4) hill_passB_iso2d<false>(m, s, ep, de, xc, RB, gs, fyv);
    else if (FM == 8) hill_passB_iso2d<true>(m, s, ep, de, xc, RB, gs, fyv);   // scalar recurrences on the carried chain
    else if (FM == 3) hill_passB_iso(m, s, ep, de, xc, RB, gs, fyv);
    else hill_passB<FM>(m, s, ep, de, W, gs, fyv);
    // ep now holds epl + depl; the increment is recovered against the (L2-resident) input record
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double2 b = __ldcg(reinterpret_cast<const double2 *>(epl + e * 6) + k);   // a real re-load, not a live register
        dp[2 * k] = ep[2 * k] - b.x;
        dp[2 * k + 1] = ep[2 * k + 1] - b.y;
    }
    if (fy) fy[e] = fyv;
    if (nsteps) nsteps[e] = PLF_MAXIT - 1;
    if (sig_out)
#pragma unroll
        for (int k = 0; k < 3; k++) reinterpret_cast<double2 *>(sig_out + e * 6)[k] = make_double2(s[2 * k], s[2 * k + 1]);
    if (depl_out)
#pragma unroll
        for (int k = 0; k < 3; k++) reinterpret_cast<double2 *>(depl_out + e * 6)[k] = make_double2(dp[2 * k], dp[2 * k + 1]);
    if (ctan)
#pragma unroll
        for (int g = 0; g < 9; g++) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int idx = 4 * g + q, i = idx / 6, j = idx - 6 * i;
                v[q] = gs[sym_idx(i, j)];
            }
            st256(ctan + e * 36 + 4 * g, v[0], v[1], v[2], v[3]);
        }
}

// Persistent, software-pipelined pass A: while a block computes and writes tile t, the 3 x 6 KB input records of its
// next tile are already in flight (cp.async into the second shared-memory buffer), so HBM reads overlap the math and
// the write-back instead of being exposed at the top of every block (ncu of the non-pipelined version:
// long_scoreboard on the staging loads dominated, 39 % DRAM throughput).
template <bool SINGLE, int FM, bool DIRECT>
__global__ void __launch_bounds__(RB, 3) response_passA_kernel(int64_t n, const int32_t *__restrict__ mat_id,
                                                               const double *__restrict__ sig, const double *__restrict__ epl,
                                                               const double *__restrict__ deps, double *__restrict__ fy,
                                                               double *__restrict__ sig_out, double *__restrict__ depl_out,
                                                               double *__restrict__ ctan, int32_t *__restrict__ nsteps,
                                                               int32_t *__restrict__ worklist, unsigned int *__restrict__ work_count)
{
    extern __shared__ double smem[];   // [2][3][RB*6] input buffers (+ [RB*CT_STRIDE] tangent staging when !DIRECT)
    double *s_ct = smem + 2 * 3 * RB * 6;
    __shared__ unsigned char s_done[DIRECT ? 1 : RB];
    const int64_t ntiles = (n + RB - 1) / RB;
    int buf = 0;
    if ((int64_t)blockIdx.x < ntiles) tile_prefetch(n, blockIdx.x, smem, sig, epl, deps);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int64_t nxt = tile + gridDim.x;
        if (nxt < ntiles) {
            tile_prefetch(n, nxt, smem + (buf ^ 1) * (3 * RB * 6), sig, epl, deps);
            cp_async_wait<1>();   // everything but the newest group (the next tile) has landed
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        passA_tile<SINGLE, FM, DIRECT>(n, tile, smem + buf * (3 * RB * 6), s_ct, s_done, mat_id, fy, sig_out, depl_out, ctan,
                                       nsteps, worklist, work_count);
        __syncthreads();   // input / staging buffers are free again (the next prefetch overwrites this input buffer)
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pass A with bulk asynchronous copies (the TMA engine's linear mode, SASS: UBLKCP): one elected thread moves a tile's
// three 6 KB input records into shared memory with cp.async.bulk + mbarrier transaction counts, and its results -- the
// 36 KB tangent tile and the two 6 KB stress / plastic-strain-increment tiles, all contiguous in HBM -- leave with three
// cp.async.bulk stores.  The other 127 threads issue no global memory instruction for the records at all (the cp.async
// ring needed 9 LDGSTS per thread and tile, the staged write-back 21 LDS + 21 STG).  Tangent records of deferred
// elements are written as whatever the staging buffer holds: pass B, which starts after this grid has completed,
// overwrites them.  (First version: all three result tiles as bulk stores from the in-place buffers; every trip then
// began with a wait for the previous stores to drain before the next load could be issued -- 8 % slower than the
// cp.async kernel.  Now only the 36 KB tangent tile is a bulk store and its drain overlaps the next trip's math.)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <bool SINGLE, int FM>
__global__ void __launch_bounds__(RB, 3) response_passA_bulk_kernel(int64_t n, const int32_t *__restrict__ mat_id,
                                                                    const double *__restrict__ sig, const double *__restrict__ epl,
                                                                    const double *__restrict__ deps, double *__restrict__ fy,
                                                                    double *__restrict__ sig_out, double *__restrict__ depl_out,
                                                                    double *__restrict__ ctan, int32_t *__restrict__ nsteps,
                                                                    int32_t *__restrict__ worklist, unsigned int *__restrict__ work_count)
{
    constexpr int FA = (FM == 6 || FM == 7) ? 1 : (FM >= 3 ? 2 : FM);
    extern __shared__ __align__(128) double smem[];   // [2][3][RB*6] input records (outputs sig / depl in place) | [RB*36] tangents
    double *s_ct = smem + 2 * 3 * RB * 6;
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ unsigned char s_done[RB];
    const int tid = threadIdx.x;
    const int64_t ntiles = (n + RB - 1) / RB;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto load_tile = [&](int64_t tile, int b) {   // thread 0 only
        const int64_t b0 = tile * RB;
        const unsigned bytes = (unsigned)min((int64_t)RB, n - b0) * 48u;
        double *d = smem + b * (3 * RB * 6);
        mbar_expect_tx(&s_bar[b], 3u * bytes);
        bulk_load(d, sig + b0 * 6, bytes, &s_bar[b]);
        bulk_load(d + RB * 6, epl + b0 * 6, bytes, &s_bar[b]);
        bulk_load(d + 2 * RB * 6, deps + b0 * 6, bytes, &s_bar[b]);
    };
    if (tid == 0 && (int64_t)blockIdx.x < ntiles) load_tile(blockIdx.x, 0);
    unsigned phase[2] = {0u, 0u};
    int buf = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        // the other input buffer was last read by the generic-proxy write-back of the tile before (finished at the
        // barrier that closed the previous trip), so its refill needs no wait on the bulk stores
        if (tid == 0) {
            const int64_t nxt = tile + gridDim.x;
            if (nxt < ntiles) load_tile(nxt, buf ^ 1);
        }
        mbar_wait(&s_bar[buf], phase[buf]);
        phase[buf] ^= 1u;
        double *s_sig = smem + buf * (3 * RB * 6), *s_epl = s_sig + RB * 6, *s_dep = s_sig + 2 * RB * 6;
        const int64_t base = tile * RB;
        const int cnt = (int)min((int64_t)RB, n - base);
        const int64_t e = base + tid;
        bool defer = false, general = false, done = false;
        double s[6], dp[6], ca[6], tc = 1., tw = 0.;
        int mat_index = 0;
        if (tid < cnt) {
            double ep[6], de[6];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double2 a = reinterpret_cast<double2 *>(s_sig)[3 * tid + k];
                const double2 b = reinterpret_cast<double2 *>(s_epl)[3 * tid + k];
                const double2 c = reinterpret_cast<double2 *>(s_dep)[3 * tid + k];
                s[2 * k] = a.x; s[2 * k + 1] = a.y; ep[2 * k] = b.x; ep[2 * k + 1] = b.y; de[2 * k] = c.x; de[2 * k + 1] = c.y;
            }
            mat_index = SINGLE ? 0 : mat_id[e];
            const DevMat &m = c_mat[mat_index];
            if (FM == 4 || FM == 6 || FM == 7) general = !planar_state(s, ep, de);
            if (m.kind == PLF_MAT_ELASTIC || m.kind == PLF_MAT_HILL) {
                double fyv = 0.;
                int br = 0;
                if (m.kind == PLF_MAT_ELASTIC) {
                    double t[6];
                    cv_mul(m, de, t);
#pragma unroll
                    for (int k = 0; k < 6; k++) { s[k] += t[k]; dp[k] = 0.; ca[k] = 0.; }
                } else {
                    br = (FM == 7) ? hillblk_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                       : (FM == 6) ? hilliso_passA(m, s, ep, de, dp, ca, tc, tw, fyv)
                       : (FM >= 3) ? iso_passA(m, s, ep, de, dp, ca, tc, tw, fyv) : hill_passA<FA>(m, s, ep, de, dp, ca, tc, tw, fyv);
                }
                if (br < 0) {
                    defer = true;
                } else {
                    done = true;
                    if (fy) fy[e] = fyv;
                    if (nsteps) nsteps[e] = 0;
                }
            }
        }
        warp_append2(defer, general, (int32_t)e, worklist, n, work_count);
        // the tangent tile of the previous trip left through a bulk store: its shared-memory image must have been read
        // before it is overwritten (the wait overlapped this trip's load and math)
        if (tid == 0) bulk_wait_read();
        __syncthreads();
        s_done[tid] = done;
        if (done) {
            const DevMat &m = c_mat[mat_index];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                reinterpret_cast<double2 *>(s_sig)[3 * tid + k] = make_double2(s[2 * k], s[2 * k + 1]);
                reinterpret_cast<double2 *>(s_epl)[3 * tid + k] = make_double2(dp[2 * k], dp[2 * k + 1]);
            }
            double2 *o = reinterpret_cast<double2 *>(s_ct + 36 * tid);
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j < 6; j += 2)
                    o[3 * i + (j >> 1)] = make_double2(m.CV[6 * i + j] * tc - (ca[i] * ca[j]) * tw,
                                                       m.CV[6 * i + j + 1] * tc - (ca[i] * ca[j + 1]) * tw);
        }
        fence_async_smem();     // generic-proxy writes to shared memory -> visible to the bulk-copy engine
        __syncthreads();
        if (tid == 0 && ctan) {
            bulk_store(ctan + base * 36, s_ct, (unsigned)cnt * 288u);
            bulk_commit();
        }
        // stress and plastic-strain increment: 2 x 6 KB, coalesced 128-bit stores by the whole block
        for (int i = tid; i < cnt * 3; i += RB) {
            if (!s_done[i / 3]) continue;
            if (sig_out) reinterpret_cast<double2 *>(sig_out + base * 6)[i] = reinterpret_cast<const double2 *>(s_sig)[i];
            if (depl_out) reinterpret_cast<double2 *>(depl_out + base * 6)[i] = reinterpret_cast<const double2 *>(s_epl)[i];
        }
        __syncthreads();        // this input buffer is refilled at the top of the next trip
    }
    if (tid == 0) bulk_wait_all();   // shared memory must outlive the last stores
}

// Pass B: persistent kernel over the two-ended work list.  (Variants that kept the 21-entry tangent accumulator in
// shared memory for 4 blocks per SM, or ran at 2 blocks per SM without register cap, measured slower and were removed.)
template <int MINB, bool SINGLE, int FM>
__global__ void __launch_bounds__(RB, MINB) response_passB_kernel(const int32_t *__restrict__ mat_id, const double *__restrict__ sig,
                                                            const double *__restrict__ epl, const double *__restrict__ deps,
                                                            double *__restrict__ fy, double *__restrict__ sig_out,
                                                            double *__restrict__ depl_out, double *__restrict__ ctan,
                                                            int32_t *__restrict__ nsteps, const int32_t *__restrict__ worklist,
                                                            int64_t nlist, unsigned int *count)
{
    // count: [0] front class, [1] back class of the two-ended work list, [2] blocks that have finished
    __shared__ double s_xc[FM >= 3 ? 12 * RB : 1];   // radial-return correction sums + deps_r / 50 of the isotropic modes
    // launched with programmatic stream serialization: the blocks are resident while pass A drains and wait here for
    // its completion (and the visibility of its work list), which hides the launch latency between the two passes
    cudaGridDependencySynchronize();
    const unsigned int n0 = count[0], n1 = count[1];
    for (unsigned int t = blockIdx.x * RB + threadIdx.x; t < n0; t += gridDim.x * RB) {
        const int64_t e = worklist[t];
        WRegs W;
        passB_element<SINGLE, FM>(e, W, s_xc + threadIdx.x, mat_id, sig, epl, deps, fy, sig_out, depl_out, ctan, nsteps);
    }
    if (FM == 4 || FM == 6 || FM == 7 || FM == 8) {
        // the non-planar elements (back of the list) with the general loop of the material class
        for (unsigned int t = blockIdx.x * RB + threadIdx.x; t < n1; t += gridDim.x * RB) {
            const int64_t e = worklist[nlist - 1 - (int64_t)t];
            WRegs W;
            passB_element<SINGLE, ((FM == 4 || FM == 8) ? 3 : 1)>(e, W, s_xc + threadIdx.x, mat_id, sig, epl, deps, fy, sig_out, depl_out, ctan,
                                                     nsteps);
        }
    }
    // the last block to finish zeroes the counters for the next launch (saves a memset node per step)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(count + 2, 1u) == gridDim.x - 1) {
            count[0] = 0; count[1] = 0; count[2] = 0;
        }
    }
}

__global__ void __launch_bounds__(RB) response_generic_kernel(int64_t n, const int32_t *__restrict__ mat_id,
                                                              const double *__restrict__ sig, const double *__restrict__ epl,
                                                              const double *__restrict__ deps, double *__restrict__ fy,
                                                              double *__restrict__ sig_out, double *__restrict__ depl_out,
                                                              double *__restrict__ ctan, int32_t *__restrict__ nsteps)
{
    extern __shared__ double smem[];
    double *s_sig = smem, *s_epl = smem + RB * 6, *s_dep = smem + 2 * RB * 6, *s_ct = smem + 3 * RB * 6;
    __shared__ unsigned char s_done[RB];
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * RB;
    const int cnt = (int)min((int64_t)RB, n - base);
    double s[6], ep[6], de[6];
    stage_in(sig, epl, deps, base, cnt, s_sig, s_epl, s_dep, s, ep, de);
    bool done = false;
    const int64_t e = base + tid;
    if (tid < cnt) {
        const DevMat &m = c_mat[mat_id ? mat_id[e] : 0];
        if (m.kind == PLF_MAT_ELASTIC || m.kind == PLF_MAT_HILL) {
            double dp[6], gs[21], fyv = 0.;
            int ns = 0;
            if (m.kind == PLF_MAT_ELASTIC) {
                double t[6];
                matvec6(m.CV, de, t);
#pragma unroll
                for (int k = 0; k < 6; k++) { s[k] += t[k]; dp[k] = 0.; }
#pragma unroll
                for (int i = 0; i < 6; i++)
#pragma unroll
                    for (int j = i; j < 6; j++) gs[sym_idx(i, j)] = m.CV[6 * i + j];
            } else {
                HillModel ym(m);
                response_core<0>(ym, m, s, ep, de, dp, gs, fyv, ns);
            }
            done = true;
            if (fy) fy[e] = fyv;
            if (nsteps) nsteps[e] = ns;
#pragma unroll
            for (int k = 0; k < 6; k++) { s_sig[6 * tid + k] = s[k]; s_epl[6 * tid + k] = dp[k]; }
            expand_sym(gs, s_ct + CT_STRIDE * tid, 1);
        }
    }
    s_done[tid] = done;
    __syncthreads();
    stage_out(base, cnt, s_sig, s_epl, s_ct, s_done, sig_out, depl_out, ctan);
}

// ------------------------------------------------------------------------------------------------------------
// K2: warp-per-element return mapping for one SVC material, SV table in shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int SVC_WARPS = 16;   // one block per SM (the SV table fills its shared memory).  With the decision / gradient loops out of line
                                // 16 warps at <= 128 registers beat 8 at 255 by 6 % (before that: 0.55 M against 0.79 M updates/s)

__device__ __forceinline__ void svc_load_table(const DevMat &m, double *s_tab)
{
    const int np = m.nsv_pad;
    const int nf = m.svc_nf;
    for (int i = threadIdx.x; i < nf * np; i += blockDim.x) s_tab[i] = m.sv[i];
    for (int i = threadIdx.x; i < np; i += blockDim.x) s_tab[nf * np + i] = m.dual[i];
    __syncthreads();
}

__global__ void __launch_bounds__(SVC_WARPS * 32, 1) response_svc_kernel(int64_t n, int mid, const int32_t *__restrict__ mat_id,
                                                                      const double *__restrict__ sig, const double *__restrict__ epl,
                                                                      const double *__restrict__ deps, double *__restrict__ fy,
                                                                      double *__restrict__ sig_out, double *__restrict__ depl_out,
                                                                      double *__restrict__ ctan, int32_t *__restrict__ nsteps,
                                                                      unsigned long long *__restrict__ queue)
{
    extern __shared__ double smem[];
    const DevMat &m = c_mat[mid];
    svc_load_table(m, smem);
    const int lane = threadIdx.x & 31;
    SvcModel ym(m, smem, smem + m.svc_nf * m.nsv_pad, lane);
    for (;;) {
        unsigned long long e0 = 0;
        if (lane == 0) e0 = atomicAdd(queue, 1ull);
        const int64_t e = (int64_t)__shfl_sync(0xffffffffu, e0, 0);
        if (e >= n) break;
        if (mat_id && mat_id[e] != mid) continue;
        double s[6], ep[6], de[6], dp[6], gs[21], fyv;
        int ns;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            s[k] = __ldg(sig + e * 6 + k);
            ep[k] = __ldg(epl + e * 6 + k);
            de[k] = __ldg(deps + e * 6 + k);
        }
        response_core<0>(ym, m, s, ep, de, dp, gs, fyv, ns);
        if (lane == 0) {
            if (fy) fy[e] = fyv;
            if (nsteps) nsteps[e] = ns;
            if (sig_out)
#pragma unroll
                for (int k = 0; k < 6; k++) sig_out[e * 6 + k] = s[k];
            if (depl_out)
#pragma unroll
                for (int k = 0; k < 6; k++) depl_out[e * 6 + k] = dp[k];
            if (ctan) expand_sym(gs, ctan + e * 36, 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// batched user-facing calls.  op: 0 calc_seq, 1 calc_yf, 2 calc_fgrad, 3 epl_dot, 4 C_tan
// ------------------------------------------------------------------------------------------------------------
enum { OP_SEQ = 0, OP_YF = 1, OP_FGRAD = 2, OP_EPLDOT = 3, OP_CTAN = 4, OP_MLFULL = 5 };

template <class Model>
__device__ __forceinline__ void prim_eval(const Model &ym, const DevMat &m, int op, const double *s, const double *ep,
                                          const double *de, const double *ld, double *out, bool writer, int64_t e)
{
    ym.set_epl(ep);   // ML yield functions with work-hardening features read the plastic strain (material.py:2342-2346)
    if (op == OP_SEQ) {
        double v = ym.seq(s);
        if (writer) out[e] = v;
    } else if (op == OP_YF) {
        double sflow = m.sy + eps_eq6(ep) * m.khard;
        double v = ym.yf(s, sflow);
        if (writer) out[e] = v;
    } else if (op == OP_MLFULL) {
        double sflow = m.sy + eps_eq6(ep) * m.khard;
        double v = ym.full_yf(s, sflow);
        if (writer) out[e] = v;
    } else if (op == OP_FGRAD) {
        double a[6];
        ym.fgrad(s, a);
        if (writer)
#pragma unroll
            for (int k = 0; k < 6; k++) out[e * 6 + k] = a[k];
    } else {
        // epl_dot (material.py:1009-1055) / C_tan (1057-1086) with Cel = Element.CV
        double a[6], ca[6], cd[6], st[6];
        matvec6(m.CV, de, cd);
#pragma unroll
        for (int k = 0; k < 6; k++) st[k] = s[k] + cd[k];
        double sflow = m.sy + eps_eq6(ep) * m.khard;
        double yfun = (op == OP_EPLDOT) ? ym.yf(st, sflow) : 1.;
        ym.fgrad(s, a);
        matvec6(m.CV, a, ca);
        double hh = ym.khard(), acd = 0.;
#pragma unroll
        for (int k = 0; k < 6; k++) { hh += a[k] * ca[k]; acd += a[k] * cd[k]; }
        if (!writer) return;
        if (op == OP_EPLDOT) {
            double lam = (yfun <= PLF_YF_TOLERANCE) ? 0. : acd / hh;
#pragma unroll
            for (int k = 0; k < 6; k++) out[e * 6 + k] = lam * a[k];
        } else {
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j < 6; j++) out[e * 36 + 6 * i + j] = m.CV[6 * i + j] - ca[i] * ca[j] / hh;
        }
    }
}

__global__ void __launch_bounds__(128) prim_hill_kernel(int op, int mid, int64_t n, const double *__restrict__ sig,
                                                        const double *__restrict__ epl, const double *__restrict__ deps,
                                                        double *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const DevMat &m = c_mat[mid];
    double s[6], ep[6] = {0, 0, 0, 0, 0, 0}, de[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        s[k] = sig[e * 6 + k];
        if (epl) ep[k] = epl[e * 6 + k];
        if (deps) de[k] = deps[e * 6 + k];
    }
    HillModel ym(m);
    prim_eval(ym, m, op, s, ep, de, nullptr, out, true, e);
}

// Yield models that need eigen-decompositions -- Barlat Yld2004-18p (BarlatModel, BASELINE config 5) and the sdim = 3
// principal-stress formulation (HillPrincModel) -- run the unsplit response_core, thread per element (FP64-bound, not
// tuned).
template <class Model>
__global__ void __launch_bounds__(128) prim_model_kernel(int op, int mid, int64_t n, const double *__restrict__ sig,
                                                          const double *__restrict__ epl, const double *__restrict__ deps,
                                                          double *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const DevMat &m = c_mat[mid];
    double s[6], ep[6] = {0, 0, 0, 0, 0, 0}, de[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        s[k] = sig[e * 6 + k];
        if (epl) ep[k] = epl[e * 6 + k];
        if (deps) de[k] = deps[e * 6 + k];
    }
    Model ym(m);
    prim_eval(ym, m, op, s, ep, de, nullptr, out, true, e);
}

// MODE 0: the whole return mapping per thread.  MODE 1 / 2: the two passes of the compacted form (as for Hill): pass 1
// finishes elastic and single-step elements and appends the ones whose trial step fails the tolerance to the work list
// (one atomic per warp); pass 2, a persistent grid over the list, runs the 50 sub-steps for them.  Without it a warp with
// one 50-sub-step element runs ~50 times longer than its neighbours (22 % of the Barlat bench states are such elements).
template <class Model, int MODE>
__global__ void __launch_bounds__(128, 4) response_model_kernel(int64_t n, int mid, const int32_t *__restrict__ mat_id,
                                                              const double *__restrict__ sig, const double *__restrict__ epl,
                                                              const double *__restrict__ deps, double *__restrict__ fy,
                                                              double *__restrict__ sig_out, double *__restrict__ depl_out,
                                                              double *__restrict__ ctan, int32_t *__restrict__ nsteps,
                                                              int32_t *__restrict__ worklist, unsigned int *__restrict__ count)
{
    const int64_t total = (MODE == 2) ? (int64_t)count[0] : n;
    const int64_t stride = (MODE == 2) ? (int64_t)gridDim.x * blockDim.x : total + blockDim.x;   // MODE 0 / 1: one trip
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {
        const int64_t t = t0 + threadIdx.x;
        bool defer = false;
        int64_t e = -1;
        if (t < total) {
            e = (MODE == 2) ? (int64_t)worklist[t] : t;
            if (!(mat_id && mat_id[e] != mid)) {
                const DevMat &m = c_mat[mid];
                double s[6], ep[6], de[6], dp[6], gs[21], fyv;
                int ns;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(sig + e * 6) + k);
                    const double2 b = __ldg(reinterpret_cast<const double2 *>(epl + e * 6) + k);
                    const double2 c = __ldg(reinterpret_cast<const double2 *>(deps + e * 6) + k);
                    s[2 * k] = a.x; s[2 * k + 1] = a.y; ep[2 * k] = b.x; ep[2 * k + 1] = b.y; de[2 * k] = c.x; de[2 * k + 1] = c.y;
                }
                Model ym(m);
                const int br = response_core<MODE == 1 ? 1 : 0>(ym, m, s, ep, de, dp, gs, fyv, ns);
                if (br < 0) {
                    defer = true;
                } else {
                    if (fy) fy[e] = fyv;
                    if (nsteps) nsteps[e] = ns;
                    if (sig_out)
#pragma unroll
                        for (int k = 0; k < 3; k++) reinterpret_cast<double2 *>(sig_out + e * 6)[k] = make_double2(s[2 * k], s[2 * k + 1]);
                    if (depl_out)
#pragma unroll
                        for (int k = 0; k < 3; k++) reinterpret_cast<double2 *>(depl_out + e * 6)[k] = make_double2(dp[2 * k], dp[2 * k + 1]);
                    if (ctan)
#pragma unroll
                        for (int g = 0; g < 9; g++) {
                            double v[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int idx = 4 * g + q, i = idx / 6, j = idx - 6 * i;
                                v[q] = gs[sym_idx(i, j)];
                            }
                            st256(ctan + e * 36 + 4 * g, v[0], v[1], v[2], v[3]);
                        }
                }
            }
        }
        if (MODE == 1) warp_append(defer, (int32_t)e, worklist, count);
    }
    if (MODE == 2) {   // the last block to finish zeroes the counter for the next launch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(count + 2, 1u) == gridDim.x - 1) { count[0] = 0; count[1] = 0; count[2] = 0; }
        }
    }
}

__global__ void __launch_bounds__(SVC_WARPS * 32, 1) prim_svc_kernel(int op, int mid, int64_t n, const double *__restrict__ sig,
                                                                  const double *__restrict__ epl, const double *__restrict__ deps,
                                                                  const double *__restrict__ ld, double *__restrict__ out)
{
    extern __shared__ double smem[];
    const DevMat &m = c_mat[mid];
    svc_load_table(m, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SvcModel ym(m, smem, smem + m.svc_nf * m.nsv_pad, lane);
    for (int64_t e = (int64_t)blockIdx.x * SVC_WARPS + warp; e < n; e += (int64_t)gridDim.x * SVC_WARPS) {
        double s[6], ep[6] = {0, 0, 0, 0, 0, 0}, de[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            s[k] = sig[e * 6 + k];
            if (epl) ep[k] = epl[e * 6 + k];
            if (deps) de[k] = deps[e * 6 + k];
        }
        if (op == OP_MLFULL) {
            double ldv[6];
            if (ld)
#pragma unroll
                for (int k = 0; k < 6; k++) ldv[k] = ld[k];
            double sflow = m.sy + eps_eq6(ep) * m.khard;
            ym.set_epl(ep);
            double v = ym.ml_full_yf(s, sflow, ld ? ldv : nullptr);
            if (lane == 0) out[e] = v;
        } else {
            prim_eval(ym, m, op, s, ep, de, nullptr, out, lane == 0, e);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// compact form of the return-mapping results for the host (plf_step_host_packed): the tangent of an element that took the
// elastic branch IS Element.CV (bit for bit) and carries no information, and every grad_stiff is symmetric, so only the
// upper triangles (21 doubles) of the other elements cross PCIe, in element order, next to one "elastic" bit per element.
//   flags : bit per element + number of non-elastic elements per block of 256
//   scan  : exclusive prefix over the block counts (one block; a chunk of the pipeline has <= 64 k blocks)
//   write : upper triangles to their compacted slots
// ------------------------------------------------------------------------------------------------------------
constexpr int PK = 256;

__device__ __forceinline__ bool tangent_is_cv(const double *__restrict__ ct, const double *cv)
{
    bool same = true;
#pragma unroll
    for (int g = 0; g < 18; g++) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(ct) + g);
        same = same && v.x == cv[2 * g] && v.y == cv[2 * g + 1];
    }
    return same;
}

__global__ void __launch_bounds__(PK) pack_flags_kernel(int64_t n, const int32_t *__restrict__ mat_id, int single,
                                                        const double *__restrict__ ctan, uint32_t *__restrict__ words,
                                                        int32_t *__restrict__ block_counts)
{
    const int64_t e = (int64_t)blockIdx.x * PK + threadIdx.x;
    bool elastic = true;   // padding lanes count as elastic
    if (e < n) elastic = tangent_is_cv(ctan + e * 36, c_mat[single ? 0 : mat_id[e]].CV);
    const unsigned bal = __ballot_sync(0xffffffffu, elastic);
    if ((threadIdx.x & 31) == 0 && e < n) words[e >> 5] = bal;
    const int cnt = __syncthreads_count(!elastic);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) pack_scan_kernel(int nblocks, const int32_t *__restrict__ counts, int32_t *__restrict__ offsets,
                                                         int64_t *__restrict__ total)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = s_carry + ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < nblocks) offsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(PK) pack_write_kernel(int64_t n, const double *__restrict__ ctan, const uint32_t *__restrict__ words,
                                                        const int32_t *__restrict__ block_offsets, double *__restrict__ ct21)
{
    __shared__ int s_w[PK / 32];
    const int64_t e = (int64_t)blockIdx.x * PK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned el = (e - lane < n) ? words[(e - lane) >> 5] : 0xffffffffu;
    const unsigned plastic = ~el;
    if (lane == 0) s_w[warp] = __popc(plastic);
    __syncthreads();
    int off = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; w++) off += s_w[w];
    if (e < n && ((plastic >> lane) & 1u)) {
        const int64_t slot = off + __popc(plastic & ((1u << lane) - 1u));
        double *o = ct21 + slot * 21;
        const double *c = ctan + e * 36;
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) o[k++] = __ldg(c + 6 * i + j);
    }
}

cudaError_t launch_pack(const LaunchCtx &lc, int64_t n, const int32_t *mat_id, int single, const double *ctan, uint32_t *words,
                        int32_t *block_counts, int32_t *block_offsets, int64_t *total, double *ct21, int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    const int nb = (int)((n + PK - 1) / PK);
    pack_flags_kernel<<<nb, PK, 0, lc.stream>>>(n, mat_id, single, ctan, words, block_counts);
    pack_scan_kernel<<<1, 1024, 0, lc.stream>>>(nb, block_counts, block_offsets, total);
    pack_write_kernel<<<nb, PK, 0, lc.stream>>>(n, ctan, words, block_offsets, ct21);
    *launches += 3;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------
void trace_kernel(const LaunchCtx &lc, const char *name, int a, int b, int c, int d)
{
    if (!lc.trace) return;
    const size_t used = strlen(lc.trace);
    if (used > 200) return;
    if (d >= 0) snprintf(lc.trace + used, 256 - used, "%s<%d, %d, %d, %d>;", name, a, b, c, d);
    else snprintf(lc.trace + used, 256 - used, "%s<%d, %d, %d>;", name, a, b, c);
}

static size_t k1_smem() { return sizeof(double) * (3 * RB * 6 + RB * CT_STRIDE); }
// double-buffered inputs (+ tangent staging for the non-direct variants)
static size_t k1a_smem(bool direct) { return sizeof(double) * (2 * 3 * RB * 6 + (direct ? 0 : RB * CT_STRIDE)); }
static size_t k1a_bulk_smem() { return sizeof(double) * (2 * 3 * RB * 6 + RB * 36); }
static size_t svc_smem(const DevMat &m) { return sizeof(double) * (size_t)(m.svc_nf + 1) * (size_t)m.nsv_pad; }

template <class K>
static cudaError_t big_smem(K kernel, size_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t response_init()
{
    cudaError_t e;
    // only the staged (non-direct) variants need more than the default 48 KB of dynamic shared memory
    if ((e = big_smem(response_passA_kernel<false, 0, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<false, 4, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<false, 6, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<false, 7, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 0, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 1, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 2, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 3, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 4, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 5, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 6, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_kernel<true, 7, false>, k1a_smem(false))) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<false, 0>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 0>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 1>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 2>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 3>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 4>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 6>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_passA_bulk_kernel<true, 7>, k1a_bulk_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_generic_kernel, k1_smem())) != cudaSuccess) return e;
    if ((e = big_smem(response_svc_kernel, 227 * 1024)) != cudaSuccess) return e;
    return big_smem(prim_svc_kernel, 227 * 1024);
}

// Fast mode of a single-material Hill batch (hill_fast.cuh): 1 = block-structured symmetric CV without hydrostatic
// term, 2 = additionally J2, 3 = additionally isotropic elasticity, 4 = 3 with planar states sorted to their own
// loop, 6 = Hill on isotropic elasticity with its planar loop.  max_fm caps it (A/B measurements, tests).
static int fast_mode_of(const DevMat &h, int max_fm)
{
    if (h.kind != PLF_MAT_HILL || !h.cv_block || !h.cv_sym || h.has_d) return 0;
    int fm = 2;
    for (int k = 0; k < 6; k++)
        if (h.hill[k] != 1.) fm = 1;
    if (fm == 2 && h.cv_iso) fm = 4;
    if (fm == 1 && h.cv_iso && max_fm >= 4) return 6;   // Hill on isotropic elasticity: its own planar mode
    if (!h.cv_iso && max_fm >= 4) return 7;              // plane-stress / cubic CV: block-CV planar mode (J2 and Hill)
    return std::min(fm, max_fm);
}
// Several materials (an FE model: plastic matrix + elastic inclusion, examples/inclusion.py): the elastic ones take pass A's
// elastic branch whatever the mode, so the batch runs in a fast mode when all its Hill / J2 materials fall into the same
// class; only the planar modes (4, 6, 7) are instantiated for that case.
static int fast_mode(const DevMat *hmats, int nmat, int max_fm)
{
    if (nmat == 1) return fast_mode_of(hmats[0], max_fm);
    int fm = -1;
    for (int i = 0; i < nmat; i++) {
        if (hmats[i].kind == PLF_MAT_ELASTIC) continue;
        const int f = fast_mode_of(hmats[i], max_fm);
        if (fm >= 0 && f != fm) return 0;
        fm = f;
    }
    return (fm == 4 || fm == 6 || fm == 7) ? fm : 0;
}

int response_fast_mode(const DevMat *hmats, int nmat, int max_fm) { return fast_mode(hmats, nmat, max_fm); }

// Launches the return mapping over n elements (device pointers).  scratch: worklist[n] int32 (all -1) + counters.
cudaError_t launch_response(const LaunchCtx &lc, const DevMat *hmats, int nmat, int64_t n, const int32_t *mat_id,
                            const double *sig, const double *epl, const double *deps, double *fy, double *sig_out,
                            double *depl, double *ctan, int32_t *nsteps, int32_t *worklist, unsigned int *work_count,
                            unsigned long long *queue, int compact, int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    bool any_hill = false;
    for (int i = 0; i < nmat; i++)
        if (hmats[i].kind == PLF_MAT_HILL || hmats[i].kind == PLF_MAT_ELASTIC) any_hill = true;
    const unsigned grid = (unsigned)((n + RB - 1) / RB);
    if (any_hill) {
        const bool single = (nmat == 1);
        const bool direct = lc.passa_direct != 0;
        const int fm = fast_mode(hmats, nmat, lc.max_fm);
        const unsigned pgrid = (unsigned)std::min<int64_t>((int64_t)lc.sm_count * 3, (int64_t)grid);
        if (compact && worklist) {
            // work_count[0..2] are zero here: set at context creation and re-zeroed by the last block of pass B
            // persistent: 3 blocks per SM, each walking tiles with a double-buffered cp.async prefetch
#define PASSA(S, F)                                                                                                              \
    do {                                                                                                                         \
        if (lc.passa_bulk && (F) != 5) trace_kernel(lc, "response_passA_bulk_kernel", S, F, -1 + 1);                             \
        else trace_kernel(lc, "response_passA_kernel", S, F, direct ? 1 : 0);                                                    \
        if (lc.passa_bulk && (F) != 5)                                                                                           \
            response_passA_bulk_kernel<S, ((F) == 5 || !(S)) ? 0 : (F)><<<pgrid, RB, k1a_bulk_smem(), lc.stream>>>(              \
                n, mat_id, sig, epl, deps, fy, sig_out, depl, ctan, nsteps, worklist, work_count);                               \
        else if (direct)                                                                                                              \
            response_passA_kernel<S, F, true><<<pgrid, RB, k1a_smem(true), lc.stream>>>(n, mat_id, sig, epl, deps, fy, sig_out,  \
                                                                                        depl, ctan, nsteps, worklist, work_count); \
        else                                                                                                                     \
            response_passA_kernel<S, F, false><<<pgrid, RB, k1a_smem(false), lc.stream>>>(n, mat_id, sig, epl, deps, fy, sig_out, \
                                                                                          depl, ctan, nsteps, worklist, work_count); \
    } while (0)
            if (lc.k1_only == 2) PASSA(true, 5);
            else if (!single && fm == 7) PASSA(false, 7);
            else if (!single && fm == 6) PASSA(false, 6);
            else if (!single && fm == 4) PASSA(false, 4);
            else if (!single) PASSA(false, 0);
            else if (fm == 7) PASSA(true, 7);
            else if (fm == 6) PASSA(true, 6);
            else if (fm == 4) PASSA(true, 4);
            else if (fm == 3) PASSA(true, 3);
            else if (fm == 2) PASSA(true, 2);
            else if (fm == 1) PASSA(true, 1);
            else PASSA(true, 0);
#undef PASSA
            if (lc.k1_only) {   // diagnostics: time pass A alone (2: without the constitutive math)
                cudaMemsetAsync(work_count, 0, 3 * sizeof(unsigned int), lc.stream);
                (*launches)++;
                return cudaGetLastError();
            }
            const int64_t bgrid = std::min<int64_t>((int64_t)lc.sm_count * 8, (int64_t)grid);
            cudaLaunchAttribute pdl_attr;
            pdl_attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl_attr.val.programmaticStreamSerializationAllowed = 1;
            cudaLaunchConfig_t pdl_cfg = {};
            pdl_cfg.gridDim = dim3((unsigned)bgrid);
            pdl_cfg.blockDim = dim3(RB);
            pdl_cfg.dynamicSmemBytes = 0;
            pdl_cfg.stream = lc.stream;
            pdl_cfg.attrs = &pdl_attr;
            pdl_cfg.numAttrs = 1;
            const int32_t *wl_c = worklist;
#define PASSB(M, S, F)                                                                                                         \
    do {                                                                                                                       \
        trace_kernel(lc, "response_passB_kernel", M, S, F);                                                                    \
        cudaLaunchKernelEx(&pdl_cfg, response_passB_kernel<M, S, F>, mat_id, sig, epl, deps, fy, sig_out, depl, ctan, nsteps, \
                           wl_c, n, work_count);                                                                               \
    } while (0)
            if (single) {
                if (fm == 7) PASSB(3, true, 7);
                else if (fm == 6) PASSB(3, true, 6);
                else if (fm == 4 && lc.passb_recur) PASSB(3, true, 8);   // FM 4 with the recurrence form of the planar loop
                else if (fm == 4) PASSB(3, true, 4);   // planar elements first, then whatever is not planar
                else if (fm == 3) PASSB(3, true, 3);
                else if (fm == 2) PASSB(3, true, 2);
                else if (fm == 1) PASSB(3, true, 1);
                else PASSB(3, true, 0);
            } else {
                if (fm == 7) PASSB(3, false, 7);
                else if (fm == 6) PASSB(3, false, 6);
                else if (fm == 4 && lc.passb_recur) PASSB(3, false, 8);
                else if (fm == 4) PASSB(3, false, 4);
                else PASSB(2, false, 0);
            }
#undef PASSB
            *launches += 2;
        } else {
            response_generic_kernel<<<grid, RB, k1_smem(), lc.stream>>>(n, mat_id, sig, epl, deps, fy, sig_out, depl, ctan, nsteps);
            (*launches)++;
        }
    }
    for (int i = 0; i < nmat; i++) {
        const int32_t *mid_p = nmat > 1 ? mat_id : nullptr;
        if (hmats[i].kind == PLF_MAT_BARLAT || hmats[i].kind == PLF_MAT_HILL3) {
            const bool bar = hmats[i].kind == PLF_MAT_BARLAT;
            if (compact && worklist) {
                const unsigned bgrid = (unsigned)std::min<int64_t>((int64_t)lc.sm_count * 8, (int64_t)grid);
                if (bar) {
                    response_model_kernel<BarlatModel, 1><<<grid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl, ctan,
                                                                                      nsteps, worklist, work_count);
                    response_model_kernel<BarlatModel, 2><<<bgrid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl, ctan,
                                                                                       nsteps, worklist, work_count);
                } else {
                    response_model_kernel<HillPrincModel, 1><<<grid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl,
                                                                                         ctan, nsteps, worklist, work_count);
                    response_model_kernel<HillPrincModel, 2><<<bgrid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl,
                                                                                          ctan, nsteps, worklist, work_count);
                }
                *launches += 2;
            } else {
                if (bar)
                    response_model_kernel<BarlatModel, 0><<<grid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl, ctan,
                                                                                      nsteps, nullptr, nullptr);
                else
                    response_model_kernel<HillPrincModel, 0><<<grid, RB, 0, lc.stream>>>(n, i, mid_p, sig, epl, deps, fy, sig_out, depl,
                                                                                         ctan, nsteps, nullptr, nullptr);
                (*launches)++;
            }
        }
    }
    for (int i = 0; i < nmat; i++) {
        if (hmats[i].kind != PLF_MAT_SVC) continue;
        cudaMemsetAsync(queue, 0, sizeof(unsigned long long), lc.stream);
        int blocks = lc.sm_count;
        int64_t need = (n + SVC_WARPS - 1) / SVC_WARPS;
        if (blocks > need) blocks = (int)need;
        response_svc_kernel<<<blocks, SVC_WARPS * 32, svc_smem(hmats[i]), lc.stream>>>(n, i, nmat > 1 ? mat_id : nullptr, sig,
                                                                                       epl, deps, fy, sig_out, depl, ctan,
                                                                                       nsteps, queue);
        (*launches)++;
    }
    return cudaGetLastError();
}

cudaError_t launch_prim(const LaunchCtx &lc, const DevMat *hmats, int op, int mid, int64_t n, const double *sig,
                        const double *epl, const double *deps, const double *ld, double *out, int64_t *launches)
{
    if (n <= 0) return cudaSuccess;
    if (hmats[mid].kind == PLF_MAT_SVC) {
        int blocks = lc.sm_count;
        int64_t need = (n + SVC_WARPS - 1) / SVC_WARPS;
        if (blocks > need) blocks = (int)need;
        prim_svc_kernel<<<blocks, SVC_WARPS * 32, svc_smem(hmats[mid]), lc.stream>>>(op, mid, n, sig, epl, deps, ld, out);
    } else if (hmats[mid].kind == PLF_MAT_BARLAT) {
        prim_model_kernel<BarlatModel><<<(unsigned)((n + 127) / 128), 128, 0, lc.stream>>>(op, mid, n, sig, epl, deps, out);
    } else if (hmats[mid].kind == PLF_MAT_HILL3) {
        prim_model_kernel<HillPrincModel><<<(unsigned)((n + 127) / 128), 128, 0, lc.stream>>>(op, mid, n, sig, epl, deps, out);
    } else {
        prim_hill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, lc.stream>>>(op, mid, n, sig, epl, deps, out);
    }
    (*launches)++;
    return cudaGetLastError();
}

}  // namespace plf
