// hill_fast.cuh -- the production return mapping for Hill-6p / J2 materials (thread per element), split into
//   pass A : elastic predictor, elastic/plastic split, ONE trial step.  Elastic and single-step elements finish
//            here; elements that need the 50-sub-step integration are deferred to a compacted work list.
//   pass B : the 50 forward-Euler sub-steps (material.py:295-344) for the deferred elements only -- dense, no
//            divergence against cheap elements.
// Same algorithm as response_core (plf_device.cuh), restructured for instruction count:
//   * CV deps_r is hoisted out of the sub-step loop (deps_r is constant there)
//   * one reciprocal square root per sub-step yields seq(sig) and 1/seq(sig) for the next gradient
//   * yield-function comparisons are done on squared quantities (no sqrt); the value is formed once at the end
//   * rcp/rsqrt: hardware seed (rcp.approx / rsqrt.approx, 2^-22) + 2 Newton steps = full double precision (~1 ulp)
//   * CV with the [3x3 normal block | diagonal shear] structure (every material made by Material.elasticity from
//     E,nu or C11,C12,C44, plane stress or strain) uses a 12-instruction product instead of 36
// Results agree with response_core / the reference to ~1e-13 relative; branch decisions can differ only when a
// yield-function value sits within rounding of its threshold.
#pragma once
#include "plf_device.cuh"

namespace plf {

__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // two Newton steps: y <- y (1.5 - 0.5 x y^2)
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.);
    y = fma(y, e, y);
    e = fma(-x, y, 1.);
    y = fma(y, e, y);
    return y;
}

// One-Newton-step versions for the inside of the 50-sub-step loops of the isotropic modes: the hardware seed is good
// to 2^-22, one step gives ~1e-13 relative -- the quantity only scales the plastic correction of a sub-step (itself
// ~1e-2 of the stress), and the sub-step loop is bound by exactly this dependent FP64 chain.
__device__ __forceinline__ double fast_rsqrt1(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-0.5 * x * y, y, 0.5);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rcp1(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.);
    return fma(y, e, y);
}

// o = CV v
__device__ __forceinline__ void cv_mul(const DevMat &m, const double *v, double *o)
{
    if (m.cv_block) {
        const double *C = m.CV;
        o[0] = C[0] * v[0] + C[1] * v[1] + C[2] * v[2];
        o[1] = C[6] * v[0] + C[7] * v[1] + C[8] * v[2];
        o[2] = C[12] * v[0] + C[13] * v[1] + C[14] * v[2];
        o[3] = C[21] * v[3];
        o[4] = C[28] * v[4];
        o[5] = C[35] * v[5];
    } else {
        matvec6(m.CV, v, o);
    }
}

// 0.5 * I2 of material.py:655-661 and the hydrostatic term I1 (650)
__device__ __forceinline__ double hill_q(const DevMat &m, const double *s)
{
    const double d01 = s[0] - s[1], d12 = s[1] - s[2], d20 = s[2] - s[0];
    return 0.5 * (m.hill[0] * (d01 * d01) + m.hill[1] * (d12 * d12) + m.hill[2] * (d20 * d20)) +
           3. * (m.hill[3] * (s[3] * s[3]) + m.hill[4] * (s[4] * s[4]) + m.hill[5] * (s[5] * s[5]));
}
__device__ __forceinline__ double hill_I1(const DevMat &m, const double *s)
{
    return m.has_d ? (s[0] * m.d[0] + s[1] * m.d[1] + s[2] * m.d[2]) / 3. : 0.;
}
// seq(s) > thr, without a square root
__device__ __forceinline__ bool hill_seq_gt(const DevMat &m, const double *s, double thr)
{
    const double t = thr - hill_I1(m, s);
    return t < 0. || hill_q(m, s) > t * t;
}
// seq and 1/seq with one reciprocal square root
__device__ __forceinline__ void hill_seq_inv(const DevMat &m, const double *s, double &seq, double &iseq)
{
    const double q = hill_q(m, s);
    if (q > 0.) {
        const double rs = fast_rsqrt(q);
        seq = q * rs;
        iseq = rs;
    } else {
        seq = 0.;
        iseq = 1. / 0.;
    }
    if (m.has_d) {
        seq += hill_I1(m, s);
        iseq = 1. / seq;
    }
}
__device__ __forceinline__ void hill_grad(const DevMat &m, const double *s, double iseq, double *a)
{
    const double p = (s[0] + s[1] + s[2]) * (1. / 3.);   // multiplication instead of the IEEE division routine
    const double s0 = s[0] - p, s1 = s[1] - p, s2 = s[2] - p;
    const double h0 = m.hill[0], h1 = m.hill[1], h2 = m.hill[2];
    const double i2 = 0.5 * iseq;
    a[0] = ((h0 + h2) * s0 - h0 * s1 - h2 * s2) * i2 + m.d3[0];
    a[1] = ((h1 + h0) * s1 - h0 * s0 - h1 * s2) * i2 + m.d3[1];
    a[2] = ((h2 + h1) * s2 - h2 * s0 - h1 * s1) * i2 + m.d3[2];
    a[3] = 3. * m.hill[3] * s[3] * iseq;
    a[4] = 3. * m.hill[4] * s[4] * iseq;
    a[5] = 3. * m.hill[5] * s[5] * iseq;
}
// fy = seq - (sy + khard * eps_eq(eplt)) > toler, without a square root
__device__ __forceinline__ bool hill_fy_gt(const DevMat &m, double seq, const double *eplt, double toler)
{
    const double lhs = seq - m.sy - toler;
    if (m.khard == 0.) return lhs > 0.;
    if (lhs <= 0.) return false;
    const double a = eplt[0] * eplt[0] + eplt[1] * eplt[1] + eplt[2] * eplt[2];
    const double b = eplt[3] * eplt[3] + eplt[4] * eplt[4] + eplt[5] * eplt[5];
    const double e2 = (2. / 3.) * (a + 0.5 * b);
    return lhs * lhs > m.khard * m.khard * e2;
}

// gradient of the quadratic form q = hill_q: q(s + c) = q(s) + s . hill_qgrad(c) + q(c)
__device__ __forceinline__ void hill_qgrad(const DevMat &m, const double *c, double *g)
{
    const double d01 = c[0] - c[1], d12 = c[1] - c[2], d20 = c[2] - c[0];
    g[0] = m.hill[0] * d01 - m.hill[2] * d20;
    g[1] = m.hill[1] * d12 - m.hill[0] * d01;
    g[2] = m.hill[2] * d20 - m.hill[1] * d12;
    g[3] = 6. * m.hill[3] * c[3];
    g[4] = 6. * m.hill[4] * c[4];
    g[5] = 6. * m.hill[5] * c[5];
}

// accumulator of the symmetric tangent sum: 21 registers
struct WRegs {
    double w[21];
    __device__ __forceinline__ double get(int k) const { return w[k]; }
    __device__ __forceinline__ void set(int k, double v) { w[k] = v; }
};

struct HillStep {
    double sflow0, toler;   // get_sflow(epl), yf_tolerance * get_sflow(epl)   (material.py:243)
    double dr[6], cdr[6];   // remaining strain increment and CV @ dr
    double st_scal;
    bool split;
};

// fast-mode helpers (defined below, see "Fast modes")
template <int FM> __device__ __forceinline__ void cv_mul_t(const DevMat &m, const double *v, double *o);
template <int FM> __device__ __forceinline__ void seq_inv_t(const DevMat &m, const double *s, double &seq, double &iseq);
template <int FM> __device__ __forceinline__ void grad_t(const DevMat &m, const double *s, double iseq, double *a);
template <int FM>
__device__ __forceinline__ bool seq_gt_t(const DevMat &m, const double *s, double thr)
{
    if (FM == 0) return hill_seq_gt(m, s, thr);
    if (FM >= 2) {
        const double p = (s[0] + s[1] + s[2]) * (1. / 3.);
        const double s0 = s[0] - p, s1 = s[1] - p, s2 = s[2] - p;
        const double q = 1.5 * (s0 * s0 + s1 * s1 + s2 * s2) + 3. * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
        return thr < 0. || q > thr * thr;
    }
    return thr < 0. || hill_q(m, s) > thr * thr;
}

// Elastic predictor and elastic/plastic split (material.py:243-274).  Returns true when the step is purely elastic
// (s updated, fy set).  Otherwise s is the stress at the start of the plastic part and P describes that part.
template <int FM>
__device__ __forceinline__ bool hill_prelude(const DevMat &m, double *s, const double *ep, const double *de, HillStep &P, double &fy)
{
    P.sflow0 = m.sy + (m.khard != 0. ? eps_eq6(ep) * m.khard : 0.);
    P.toler = PLF_YF_TOLERANCE * P.sflow0;
    double cd[6], st[6];
    cv_mul_t<FM>(m, de, cd);
#pragma unroll
    for (int k = 0; k < 6; k++) st[k] = s[k] + cd[k];
    if (!seq_gt_t<FM>(m, st, P.sflow0 + P.toler) ) {
        // fy1 < toler would be the exact test; seq == sflow0 + toler to the last bit is treated as elastic too
        double seq, iseq;
        seq_inv_t<FM>(m, st, seq, iseq);
        fy = seq - P.sflow0;
        if (fy < P.toler) {
#pragma unroll
            for (int k = 0; k < 6; k++) s[k] = st[k];
            return true;
        }
    }
    P.st_scal = 1.;
    P.split = false;
    double seq0, iseq0;
    seq_inv_t<FM>(m, s, seq0, iseq0);
    const double fy0 = seq0 - P.sflow0;   // 259
    if (fy0 < PLF_SPLIT_FY0) {            // 260-270
        P.split = true;
        double seqd, iseqd;
        seq_inv_t<FM>(m, cd, seqd, iseqd);
        P.st_scal += fy0 / seqd;
        const double c = 1. - P.st_scal;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double del = de[k] * c;
            P.dr[k] = de[k] - del;
        }
        double cdel[6], del[6];
#pragma unroll
        for (int k = 0; k < 6; k++) del[k] = de[k] * c;
        cv_mul_t<FM>(m, del, cdel);
#pragma unroll
        for (int k = 0; k < 6; k++) s[k] += cdel[k];
        cv_mul_t<FM>(m, P.dr, P.cdr);
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) { P.dr[k] = de[k]; P.cdr[k] = cd[k]; }
    }
    return false;
}

// Pass A.  Returns 0 (elastic), 1 (plastic, single step) or -1 (needs the 50-sub-step loop: deferred).
// Tangent is returned in factored form: ct_ij = CV_ij * tc - ca_i ca_j * tw   (elastic: tc = 1, tw = 0).
template <int FM>
__device__ __forceinline__ int hill_passA(const DevMat &m, double *s, const double *ep, const double *de, double *dp, double *ca,
                                          double &tc, double &tw, double &fy)
{
    HillStep P;
#pragma unroll
    for (int k = 0; k < 6; k++) { dp[k] = 0.; ca[k] = 0.; }
    tc = 1.; tw = 0.;
    if (hill_prelude<FM>(m, s, ep, de, P, fy)) return 0;
    // trial step with the full remaining increment (277-285) == the only step when nsteps = 1 (295-308)
    double st[6], a[6], seq, iseq;
#pragma unroll
    for (int k = 0; k < 6; k++) st[k] = s[k] + P.cdr[k];
    const bool yplus = seq_gt_t<FM>(m, st, P.sflow0 + PLF_YF_TOLERANCE);   // epl_dot 1031-1041
    seq_inv_t<FM>(m, s, seq, iseq);
    grad_t<FM>(m, s, iseq, a);
    cv_mul_t<FM>(m, a, ca);
    double hh = m.khard, acd = 0., cadr = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) { hh += a[k] * ca[k]; acd += a[k] * P.cdr[k]; }
    cadr = acd;   // symmetric CV (see hill_passB)
    if (FM == 0 && !m.cv_sym) {
        cadr = 0.;
#pragma unroll
        for (int k = 0; k < 6; k++) cadr += ca[k] * P.dr[k];
    }
    const double ihh = fast_rcp(hh);
    const double lam = yplus ? acd * ihh : 0.;
    const double q = cadr * ihh;
    double eplt[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        dp[k] = lam * a[k];
        eplt[k] = ep[k] + dp[k];
        st[k] = s[k] + (P.cdr[k] - ca[k] * q);
    }
    seq_inv_t<FM>(m, st, seq, iseq);
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(eplt) * m.khard : 0.));
    if (fy > P.toler) return -1;   // 288-291: subdivide
#pragma unroll
    for (int k = 0; k < 6; k++) s[k] = st[k];
    // grad_stiff = CV (1 - st_scal) [split] + (CV - ca ca^T / hh) st_scal   (269, 343 with nsteps = 1)
    tc = (P.split ? (1. - P.st_scal) : 0.) + P.st_scal;
    tw = ihh * P.st_scal;
    return 1;
}

// Pass B: the 50 sub-steps for an element that pass A deferred.
//   epd : in = epl at the start of the step, out = epl + depl (the caller recovers depl = epd - epl; keeping only the
//         running sum live saves 12 registers in the loop)
//   W   : accumulator (WRegs) for sum_k ca ca^T / hh (+ correction blocks)
//   gs  : out, symmetric 21.
// Returns branch code 2 (| 4 when a radial-return correction fired).
// Fast modes of the sub-step loop, chosen on the host for single-material batches (the flags are compile-time, so the
// uniform branches and the dead general-case code disappear from the loop):
//   FM 0 : general (any CV, Drucker / LHS terms)
//   FM 1 : block-structured symmetric CV, no hydrostatic term            (DevMat.cv_block && cv_sym && !has_d)
//   FM 2 : FM 1 and all Hill parameters equal 1 (J2): q = 1.5 |dev s|^2 + 3 |shear|^2, a = 1.5 dev(s) / seq
//   FM 3 : FM 2 and isotropic elasticity (DevMat.cv_iso): iso_passA / hill_passB_iso below
//   FM 4 : FM 3 with planar states sorted to hill_passB_iso2d (the non-planar rest runs the FM 3 loop)
//   FM 8 : FM 4 with the recurrence form of the planar loop (hill_passB_iso2d<true>; pass B only, option passb_recur)
//   FM 6 : FM 1 and isotropic elasticity, planar states sorted to hill_passB_hilliso2d (rest: the FM 1 loop)
//   FM 7 : FM 1 / 2 on a block-structured CV that is not isotropic (plane stress, cubic), planar states sorted to
//          hill_passB_blk2d (rest: the FM 1 loop)
template <int FM>
__device__ __forceinline__ void cv_mul_t(const DevMat &m, const double *v, double *o)
{
    if (FM > 0) {
        const double *C = m.CV;
        o[0] = C[0] * v[0] + C[1] * v[1] + C[2] * v[2];
        o[1] = C[6] * v[0] + C[7] * v[1] + C[8] * v[2];
        o[2] = C[12] * v[0] + C[13] * v[1] + C[14] * v[2];
        o[3] = C[21] * v[3];
        o[4] = C[28] * v[4];
        o[5] = C[35] * v[5];
    } else {
        cv_mul(m, v, o);
    }
}
// seq(s), 1/seq(s) and the gradient direction for the NEXT sub-step share the deviator
template <int FM>
__device__ __forceinline__ void seq_inv_t(const DevMat &m, const double *s, double &seq, double &iseq)
{
    if (FM >= 2) {
        const double p = (s[0] + s[1] + s[2]) * (1. / 3.);
        const double s0 = s[0] - p, s1 = s[1] - p, s2 = s[2] - p;
        const double q = 1.5 * (s0 * s0 + s1 * s1 + s2 * s2) + 3. * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]);
        const double rs = q > 0. ? fast_rsqrt(q) : 1. / 0.;
        seq = q > 0. ? q * rs : 0.;
        iseq = rs;
    } else if (FM == 1) {
        const double q = hill_q(m, s);
        const double rs = q > 0. ? fast_rsqrt(q) : 1. / 0.;
        seq = q > 0. ? q * rs : 0.;
        iseq = rs;
    } else {
        hill_seq_inv(m, s, seq, iseq);
    }
}
template <int FM>
__device__ __forceinline__ void grad_t(const DevMat &m, const double *s, double iseq, double *a)
{
    if (FM >= 2) {
        const double p = (s[0] + s[1] + s[2]) * (1. / 3.);
        const double c15 = 1.5 * iseq, c3 = 3. * iseq;
        a[0] = (s[0] - p) * c15; a[1] = (s[1] - p) * c15; a[2] = (s[2] - p) * c15;
        a[3] = s[3] * c3; a[4] = s[4] * c3; a[5] = s[5] * c3;
    } else {
        hill_grad(m, s, iseq, a);
    }
}

template <int FM, class WAcc>
__device__ __forceinline__ int hill_passB(const DevMat &m, double *s, double *epd, const double *de, WAcc &W, double *gs, double &fy)
{
    HillStep P;
    hill_prelude<FM>(m, s, epd, de, P, fy);   // cannot be elastic: pass A deferred this element
    const double inv = 1. / (double)PLF_MAXIT;
#pragma unroll
    for (int k = 0; k < 6; k++) P.dr[k] = P.dr[k] / (double)PLF_MAXIT;   // 290
    cv_mul_t<FM>(m, P.dr, P.cdr);
#pragma unroll
    for (int k = 0; k < 21; k++) W.set(k, 0.);
    // q(s + cdr) = q(s) + s . gq + qc: the sub-step's elastic-trial yield check without re-evaluating the form
    double gq[6];
    hill_qgrad(m, P.cdr, gq);
    const double qc = hill_q(m, P.cdr);
    double seq, iseq;
    seq_inv_t<FM>(m, s, seq, iseq);
    bool corrected = false;
    const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
    const double ythr2 = ythr * ythr;
#pragma unroll 1
    for (int it = 0; it < PLF_MAXIT; it++) {
        double a[6], ca[6], dd[6];
        bool yplus;
        if (FM == 0 && m.has_d) {
            double st[6];
#pragma unroll
            for (int k = 0; k < 6; k++) st[k] = s[k] + P.cdr[k];
            yplus = hill_seq_gt(m, st, ythr);
        } else {
            double qst = seq * seq + qc;   // seq^2 = q(s) when there is no hydrostatic term
#pragma unroll
            for (int k = 0; k < 6; k++) qst = fma(s[k], gq[k], qst);
            yplus = qst > ythr2;
        }
        grad_t<FM>(m, s, iseq, a);
        cv_mul_t<FM>(m, a, ca);
        double hh = m.khard, acd = 0.;
#pragma unroll
        for (int k = 0; k < 6; k++) { hh = fma(a[k], ca[k], hh); acd = fma(a[k], P.cdr[k], acd); }
        // (CV a) . deps_r == a . (CV deps_r) for a symmetric CV: the second dot product (and deps_r itself) is only
        // needed for a non-symmetric user-supplied stiffness
        double cadr = acd;
        if (FM == 0 && !m.cv_sym) {
            cadr = 0.;
#pragma unroll
            for (int k = 0; k < 6; k++) cadr = fma(ca[k], P.dr[k], cadr);
        }
        const double ihh = fast_rcp(hh);
        const double lam = yplus ? acd * ihh : 0.;
        const double q = cadr * ihh;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            dd[k] = lam * a[k];
            s[k] += P.cdr[k] - ca[k] * q;
        }
        seq_inv_t<FM>(m, s, seq, iseq);
        double eplt[6];
#pragma unroll
        for (int k = 0; k < 6; k++) eplt[k] = epd[k] + dd[k];
        if (hill_fy_gt(m, seq, eplt, P.toler)) {
            // radial-return correction (310-342), rare: exact arithmetic path
            corrected = true;
            const double fyv = seq - (m.sy + eps_eq6(eplt) * m.khard);
            double ex[6], t[6], x[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                ex[k] = s[k] * fyv / seq;
                s[k] -= ex[k];
            }
            matvec6(m.SV, ex, t);
#pragma unroll
            for (int k = 0; k < 6; k++) eplt[k] += t[k];
            lstsq_3x6(P.dr, ex, x);
            W.set(sym_idx(0, 0), W.get(sym_idx(0, 0)) + x[0]);
            W.set(sym_idx(0, 1), W.get(sym_idx(0, 1)) + x[5]);
            W.set(sym_idx(0, 2), W.get(sym_idx(0, 2)) + x[4]);
            W.set(sym_idx(1, 1), W.get(sym_idx(1, 1)) + x[1]);
            W.set(sym_idx(1, 2), W.get(sym_idx(1, 2)) + x[3]);
            W.set(sym_idx(2, 2), W.get(sym_idx(2, 2)) + x[2]);
            seq_inv_t<FM>(m, s, seq, iseq);
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double ci = ca[i] * ihh;
#pragma unroll
            for (int j = i; j < 6; j++) W.set(sym_idx(i, j), fma(ci, ca[j], W.get(sym_idx(i, j))));
        }
#pragma unroll
        for (int k = 0; k < 6; k++) epd[k] = eplt[k];
    }
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(epd) * m.khard : 0.));
    const double f1 = P.split ? (1. - P.st_scal) : 0.;
    const double f2 = P.st_scal * inv;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double c = m.CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)PLF_MAXIT * c - W.get(sym_idx(i, j))) * f2;
        }
    return 2 | (corrected ? 4 : 0);
}

// sqrt through the fast reciprocal square root (x >= 0; ~1 ulp)
__device__ __forceinline__ double fast_sqrt(double x) { return x * fast_rsqrt(x + 1e-300); }
__device__ __forceinline__ double eps_eq6_fast(const double *e)
{
    const double a = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    const double b = e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
    return fast_sqrt((2. / 3.) * (a + 0.5 * b));
}

// sqrt(q) and 1/sqrt(q) of q >= 0 without a branch (q = 0: seq = 0 and a finite rs; a stress-free state has no flow
// direction, the reference divides by zero there)
__device__ __forceinline__ void seq_rs(double q, double &seq, double &rs)
{
    rs = fast_rsqrt(q + 1e-300);   // the addend vanishes in rounding for any q > 1e-284
    seq = q * rs;
}
// inside the sub-step loops: one Newton step (see fast_rsqrt1)
__device__ __forceinline__ void seq_rs1(double q, double &seq, double &rs)
{
    rs = fast_rsqrt1(q + 1e-300);
    seq = q * rs;
}

// ------------------------------------------------------------------------------------------------------------
// FM 3: J2 yield function on isotropic elasticity (C11 - C12 = 2 C44 = 2G; BASELINE configs 2 and the 1024^2 target).
// With n = (dev(s)_0..2, s_3..5) and r = 1/seq the flow direction is a = r (1.5 n_i | 3 n_k), and because a is
// deviatoric   CV a = 3G r n   for ALL six components, so that
//     hh  = a.CV a + khard = 3G + khard                      (a constant: no reciprocal in the loop)
//     acd = a.(CV deps_r) = r ncd,   ncd = 1.5 n_i dc_i + 3 n_k dc_k,  dc = dev/shear part of CV deps_r
//     sig += CV deps_r - CV a (acd / hh)      <=>  n_k += dc_k - beta n_k,  p += pc,  beta = 3G ncd / (hh seq^2)
//     q(sig + CV deps_r) = seq^2 + q(CV deps_r) + 2 ncd       (elastic-trial check of epl_dot, no extra dot product)
//     sum CV a (CV a)^T / hh = sum w n n^T,  w = (3G)^2 / (hh seq^2); n_2 = -n_0 - n_1, so only the 15 entries over
//     the index set {0,1,3,4,5} are accumulated and the index-2 row is recovered at the end.
// The same reference algorithm (50 forward-Euler sub-steps, material.py:295-344), ~85 FP64 instructions per
// sub-step instead of ~160.  Radial-return corrections (rare) are added through 6 shared-memory words per thread.
// ------------------------------------------------------------------------------------------------------------
// Elastic predictor and elastic/plastic split for the isotropic modes (QF 2: J2, QF 1: Hill-6p form): hill_prelude with
//   * CV (c deps) = c (CV deps): the split re-uses CV deps instead of two more matrix products,
//   * the ratio fy0 / seq(CV deps) through the reciprocal square root that seq comes from,
//   * sqrt of eps_eq through fast_rsqrt (no IEEE sqrt / division sequences on the path).
template <int QF = 2>
__device__ __forceinline__ bool iso_prelude(const DevMat &m, double *s, const double *ep, const double *de, HillStep &P, double &fy)
{
    P.sflow0 = m.sy + (m.khard != 0. ? eps_eq6_fast(ep) * m.khard : 0.);
    P.toler = PLF_YF_TOLERANCE * P.sflow0;
    double cd[6], st[6];
    cv_mul_t<QF>(m, de, cd);
#pragma unroll
    for (int k = 0; k < 6; k++) st[k] = s[k] + cd[k];
    if (!seq_gt_t<QF>(m, st, P.sflow0 + P.toler)) {
        double seq, iseq;
        seq_inv_t<QF>(m, st, seq, iseq);
        fy = seq - P.sflow0;
        if (fy < P.toler) {
#pragma unroll
            for (int k = 0; k < 6; k++) s[k] = st[k];
            return true;
        }
    }
    P.st_scal = 1.;
    P.split = false;
    double seq0, iseq0;
    seq_inv_t<QF>(m, s, seq0, iseq0);
    const double fy0 = seq0 - P.sflow0;   // 259
    if (fy0 < PLF_SPLIT_FY0) {            // 260-270
        P.split = true;
        double seqd, iseqd;
        seq_inv_t<QF>(m, cd, seqd, iseqd);
        P.st_scal += fy0 * iseqd;
        const double c = 1. - P.st_scal;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            s[k] = fma(c, cd[k], s[k]);
            P.dr[k] = de[k] - de[k] * c;
            P.cdr[k] = cd[k] - cd[k] * c;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) { P.dr[k] = de[k]; P.cdr[k] = cd[k]; }
    }
    return false;
}

// Pass A for the isotropic J2 modes (same contract as hill_passA): the trial step written with the identities of
// hill_passB_iso -- CV a = 3G r n, hh = 3G + khard -- so that no gradient, no matrix product and no reciprocal is
// evaluated; roughly half the FP64 instructions of hill_passA<2>.
__device__ __forceinline__ int iso_passA(const DevMat &m, double *s, const double *ep, const double *de, double *dp, double *ca,
                                         double &tc, double &tw, double &fy)
{
    HillStep P;
#pragma unroll
    for (int k = 0; k < 6; k++) { dp[k] = 0.; ca[k] = 0.; }
    tc = 1.; tw = 0.;
    if (iso_prelude(m, s, ep, de, P, fy)) return 0;
    const double G3 = 3. * m.CV[21];
    const double ihh = m.iso_ihh;
    const double pc = (P.cdr[0] + P.cdr[1] + P.cdr[2]) * (1. / 3.);
    const double dc[6] = {P.cdr[0] - pc, P.cdr[1] - pc, P.cdr[2] - pc, P.cdr[3], P.cdr[4], P.cdr[5]};
    const double qc = 1.5 * (dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]) + 3. * (dc[3] * dc[3] + dc[4] * dc[4] + dc[5] * dc[5]);
    const double p = (s[0] + s[1] + s[2]) * (1. / 3.);
    const double n[6] = {s[0] - p, s[1] - p, s[2] - p, s[3], s[4], s[5]};
    const double q2 = 1.5 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 3. * (n[3] * n[3] + n[4] * n[4] + n[5] * n[5]);
    double seq, rs;
    seq_rs(q2, seq, rs);
    const double ncd = 1.5 * (n[0] * dc[0] + n[1] * dc[1] + n[2] * dc[2]) + 3. * (n[3] * dc[3] + n[4] * dc[4] + n[5] * dc[5]);
    const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
    const bool yplus = fma(2., ncd, q2 + qc) > ythr * ythr;   // epl_dot 1031-1041 on the elastic trial stress
    const double g = ncd * (rs * rs);
    const double beta = g * (G3 * ihh);
    const double gam = yplus ? g * (1.5 * ihh) : 0.;
    double eplt[6], nt[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        dp[k] = (k < 3 ? gam : gam + gam) * n[k];
        eplt[k] = ep[k] + dp[k];
        nt[k] = fma(-beta, n[k], n[k] + dc[k]);
    }
    const double q2t = 1.5 * (nt[0] * nt[0] + nt[1] * nt[1] + nt[2] * nt[2]) + 3. * (nt[3] * nt[3] + nt[4] * nt[4] + nt[5] * nt[5]);
    fy = fast_sqrt(q2t) - (m.sy + (m.khard != 0. ? eps_eq6_fast(eplt) * m.khard : 0.));
    if (fy > P.toler) return -1;   // 288-291: subdivide
    const double pt = p + pc;
    s[0] = nt[0] + pt; s[1] = nt[1] + pt; s[2] = nt[2] + pt; s[3] = nt[3]; s[4] = nt[4]; s[5] = nt[5];
    const double cf = G3 * rs;
#pragma unroll
    for (int k = 0; k < 6; k++) ca[k] = cf * n[k];
    // grad_stiff = CV (1 - st_scal) [split] + (CV - ca ca^T / hh) st_scal   (269, 343 with nsteps = 1)
    tc = (P.split ? (1. - P.st_scal) : 0.) + P.st_scal;
    tw = ihh * P.st_scal;
    return 1;
}

__device__ __forceinline__ int hill_passB_iso(const DevMat &m, double *s, double *epd, const double *de, double *xc, int xstride,
                                              double *gs, double &fy)
{
    HillStep P;
    iso_prelude(m, s, epd, de, P, fy);   // cannot be elastic: pass A deferred this element
    const double inv = 1. / (double)PLF_MAXIT;
#pragma unroll
    for (int k = 0; k < 6; k++) P.dr[k] *= inv;   // 290 (x * 0.02 instead of x / 50: 1 ulp, saves six IEEE divisions)
    cv_mul_t<2>(m, P.dr, P.cdr);
    const double G3 = 3. * m.CV[21];
    const double ihh = m.iso_ihh;
    const double bfac = G3 * ihh, wfac = G3 * G3 * ihh;
    const double pc = (P.cdr[0] + P.cdr[1] + P.cdr[2]) * (1. / 3.);
    double dc[6] = {P.cdr[0] - pc, P.cdr[1] - pc, P.cdr[2] - pc, P.cdr[3], P.cdr[4], P.cdr[5]};
    const double qc = 1.5 * (dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]) + 3. * (dc[3] * dc[3] + dc[4] * dc[4] + dc[5] * dc[5]);
    double p = (s[0] + s[1] + s[2]) * (1. / 3.);
    double n[6] = {s[0] - p, s[1] - p, s[2] - p, s[3], s[4], s[5]};
    double q2 = 1.5 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 3. * (n[3] * n[3] + n[4] * n[4] + n[5] * n[5]);
    double rs, seq;
    seq_rs(q2, seq, rs);
    const double c0 = m.sy + P.toler, K2 = m.khard * m.khard;
    // reduced accumulator over the index set {0,1,3,4,5}: W[a][b], a <= b
    double W00 = 0., W01 = 0., W03 = 0., W04 = 0., W05 = 0., W11 = 0., W13 = 0., W14 = 0., W15 = 0., W33 = 0., W34 = 0.,
           W35 = 0., W44 = 0., W45 = 0., W55 = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) xc[k * xstride] = 0.;
    bool corrected = false;
    const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
    const double ythr2 = ythr * ythr;
#pragma unroll 1
    for (int it = 0; it < PLF_MAXIT; it++) {
        const double ncd = 1.5 * (n[0] * dc[0] + n[1] * dc[1] + n[2] * dc[2]) + 3. * (n[3] * dc[3] + n[4] * dc[4] + n[5] * dc[5]);
        const bool yplus = fma(2., ncd, q2 + qc) > ythr2;
        const double r2 = rs * rs;
        const double g = ncd * r2;
        const double beta = g * bfac;
        const double gam = yplus ? 1.5 * (g * ihh) : 0.;
        const double gam2 = gam + gam;
        const double w = r2 * wfac;
        const double t0 = w * n[0], t1 = w * n[1], t3 = w * n[3], t4 = w * n[4], t5 = w * n[5];
        W00 = fma(t0, n[0], W00); W01 = fma(t0, n[1], W01); W03 = fma(t0, n[3], W03); W04 = fma(t0, n[4], W04); W05 = fma(t0, n[5], W05);
        W11 = fma(t1, n[1], W11); W13 = fma(t1, n[3], W13); W14 = fma(t1, n[4], W14); W15 = fma(t1, n[5], W15);
        W33 = fma(t3, n[3], W33); W34 = fma(t3, n[4], W34); W35 = fma(t3, n[5], W35);
        W44 = fma(t4, n[4], W44); W45 = fma(t4, n[5], W45);
        W55 = fma(t5, n[5], W55);
        double eplt[6];
        eplt[0] = fma(gam, n[0], epd[0]); eplt[1] = fma(gam, n[1], epd[1]); eplt[2] = fma(gam, n[2], epd[2]);
        eplt[3] = fma(gam2, n[3], epd[3]); eplt[4] = fma(gam2, n[4], epd[4]); eplt[5] = fma(gam2, n[5], epd[5]);
#pragma unroll
        for (int k = 0; k < 6; k++) n[k] = fma(-beta, n[k], n[k] + dc[k]);
        p += pc;
        q2 = 1.5 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 3. * (n[3] * n[3] + n[4] * n[4] + n[5] * n[5]);
        seq_rs(q2, seq, rs);
        const double e2s = (2. / 3.) * ((eplt[0] * eplt[0] + eplt[1] * eplt[1] + eplt[2] * eplt[2]) +
                                        0.5 * (eplt[3] * eplt[3] + eplt[4] * eplt[4] + eplt[5] * eplt[5]));
        const double lhs = seq - c0;
        if (lhs > 0. && lhs * lhs > K2 * e2s) {   // fy > toler (hill_fy_gt), one rarely taken branch
            // radial-return correction (310-342), rare: exact arithmetic path on the full stress vector
            corrected = true;
            const double fyv = seq - (m.sy + eps_eq6(eplt) * m.khard);
            double sf[6] = {n[0] + p, n[1] + p, n[2] + p, n[3], n[4], n[5]};
            double ex[6], t[6], x[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                ex[k] = sf[k] * fyv / seq;
                sf[k] -= ex[k];
            }
            matvec6(m.SV, ex, t);
#pragma unroll
            for (int k = 0; k < 6; k++) eplt[k] += t[k];
            lstsq_3x6(P.dr, ex, x);
#pragma unroll
            for (int k = 0; k < 6; k++) xc[k * xstride] += x[k];
            p = (sf[0] + sf[1] + sf[2]) * (1. / 3.);
            n[0] = sf[0] - p; n[1] = sf[1] - p; n[2] = sf[2] - p; n[3] = sf[3]; n[4] = sf[4]; n[5] = sf[5];
            q2 = 1.5 * (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 3. * (n[3] * n[3] + n[4] * n[4] + n[5] * n[5]);
            seq_rs(q2, seq, rs);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) epd[k] = eplt[k];
    }
    s[0] = n[0] + p; s[1] = n[1] + p; s[2] = n[2] + p; s[3] = n[3]; s[4] = n[4]; s[5] = n[5];
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(epd) * m.khard : 0.));
    // full symmetric sum: index 2 from the trace relation n_2 = -n_0 - n_1, plus the correction blocks
    double Wf[21];
    Wf[sym_idx(0, 0)] = W00; Wf[sym_idx(0, 1)] = W01; Wf[sym_idx(0, 3)] = W03; Wf[sym_idx(0, 4)] = W04; Wf[sym_idx(0, 5)] = W05;
    Wf[sym_idx(1, 1)] = W11; Wf[sym_idx(1, 3)] = W13; Wf[sym_idx(1, 4)] = W14; Wf[sym_idx(1, 5)] = W15;
    Wf[sym_idx(3, 3)] = W33; Wf[sym_idx(3, 4)] = W34; Wf[sym_idx(3, 5)] = W35;
    Wf[sym_idx(4, 4)] = W44; Wf[sym_idx(4, 5)] = W45; Wf[sym_idx(5, 5)] = W55;
    Wf[sym_idx(0, 2)] = -W00 - W01;
    Wf[sym_idx(1, 2)] = -W01 - W11;
    Wf[sym_idx(2, 2)] = W00 + 2. * W01 + W11;
    Wf[sym_idx(2, 3)] = -W03 - W13;
    Wf[sym_idx(2, 4)] = -W04 - W14;
    Wf[sym_idx(2, 5)] = -W05 - W15;
    if (corrected) {
        Wf[sym_idx(0, 0)] += xc[0 * xstride];
        Wf[sym_idx(0, 1)] += xc[5 * xstride];
        Wf[sym_idx(0, 2)] += xc[4 * xstride];
        Wf[sym_idx(1, 1)] += xc[1 * xstride];
        Wf[sym_idx(1, 2)] += xc[3 * xstride];
        Wf[sym_idx(2, 2)] += xc[2 * xstride];
    }
    const double f1 = P.split ? (1. - P.st_scal) : 0.;
    const double f2 = P.st_scal * inv;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double c = m.CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)PLF_MAXIT * c - Wf[sym_idx(i, j)]) * f2;
        }
    return 2 | (corrected ? 4 : 0);
}

// ------------------------------------------------------------------------------------------------------------
// FM 3, planar states: sig, epl, deps with zero 23 and 13 components (every state a 2-d pyLabFEA model produces:
// Element.deps has no out-of-plane shear, model.py:387-398).  Block-structured CV keeps such a state planar through
// the whole update, so the loop of hill_passB_iso shrinks to the components {0,1,2,5}: 6 accumulator entries over
// the index set {0,1,5}, ~65 FP64 instructions per sub-step, ~100 registers.  Pass A sorts deferred elements into
// planar (front of the work list) and general (back).  dr: the six words of deps_r / 50 are parked in shared memory
// next to the correction sums (only the rare radial-return correction reads them).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool planar_state(const double *s, const double *ep, const double *de)
{
    return s[3] == 0. && s[4] == 0. && ep[3] == 0. && ep[4] == 0. && de[3] == 0. && de[4] == 0.;
}

// radial-return correction of hill_passB_iso2d (material.py:310-342) on st = {n0, n1, n2, n5, p, eplt0, eplt1, eplt2,
// eplt5, seq}; xc: [0..5] correction sums, [6..11] deps_r / 50 (shared memory, stride xstride)
static __device__ __noinline__ void iso2d_correct(const DevMat &m, double *st, double *xc, int xstride)
{
    const double p = st[4], seq = st[9];
    double eplt[6] = {st[5], st[6], st[7], 0., 0., st[8]};
    const double fyv = seq - (m.sy + eps_eq6(eplt) * m.khard);
    double sf[6] = {st[0] + p, st[1] + p, st[2] + p, 0., 0., st[3]};
    double ex[6], t[6], x[6], dr[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ex[k] = sf[k] * fyv / seq;
        sf[k] -= ex[k];
        dr[k] = xc[(6 + k) * xstride];
    }
    matvec6(m.SV, ex, t);
    lstsq_3x6(dr, ex, x);
#pragma unroll
    for (int k = 0; k < 6; k++) xc[k * xstride] += x[k];
    const double pn = (sf[0] + sf[1] + sf[2]) * (1. / 3.);
    st[0] = sf[0] - pn; st[1] = sf[1] - pn; st[2] = sf[2] - pn; st[3] = sf[5]; st[4] = pn;
    st[5] = eplt[0] + t[0]; st[6] = eplt[1] + t[1]; st[7] = eplt[2] + t[2]; st[8] = eplt[5] + t[5];
}

// RC (fast mode 8, option passb_recur): because n' = (1 - beta) n + dc, the two scalars a sub-step needs follow exact
// recurrences -- n'.dc = (1 - beta) n.dc + dc.dc and q' = (1 - beta)((1 - beta) q + 2 n.dc) + dc.dc (weighted products
// as in q) -- so the loop-carried dependency shrinks from  n -> q (4 ops) -> rsqrt -> r^2 -> g -> beta -> n  to
// q -> rcp -> beta -> q'; the reciprocal for the next sub-step is issued before the yield check of this one.  The vector
// update, the tangent sums and the check hang off that chain.  q and n.dc are re-derived from n after a radial-return
// correction and at the end; they differ from the direct evaluation by ~1e-14 relative after 50 steps.
template <bool RC = false>
__device__ __forceinline__ int hill_passB_iso2d(const DevMat &m, double *s, double *epd, const double *de, double *xc, int xstride,
                                                double *gs, double &fy)
{
    double st_scal, f1;
    double pc, dc0, dc1, dc2, dc5, qc, c0, ythr2;
    {
        HillStep P;
        iso_prelude(m, s, epd, de, P, fy);   // cannot be elastic: pass A deferred this element
#pragma unroll
        for (int k = 0; k < 6; k++) {
            P.dr[k] *= 1. / (double)PLF_MAXIT;   // 290
            xc[(6 + k) * xstride] = P.dr[k];
            xc[k * xstride] = 0.;
        }
        cv_mul_t<2>(m, P.dr, P.cdr);
        pc = (P.cdr[0] + P.cdr[1] + P.cdr[2]) * (1. / 3.);
        dc0 = P.cdr[0] - pc; dc1 = P.cdr[1] - pc; dc2 = P.cdr[2] - pc; dc5 = P.cdr[5];
        qc = 1.5 * (dc0 * dc0 + dc1 * dc1 + dc2 * dc2) + 3. * (dc5 * dc5);
        c0 = m.sy + P.toler;
        const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
        ythr2 = ythr * ythr;
        st_scal = P.st_scal;
        f1 = P.split ? (1. - P.st_scal) : 0.;
    }
    const double G3 = 3. * m.CV[21];
    const double ihh = m.iso_ihh;
    const double bfac = G3 * ihh, wfac = G3 * G3 * ihh, gfac = 1.5 * ihh, K2 = m.khard * m.khard;
    double p = (s[0] + s[1] + s[2]) * (1. / 3.);
    double n0 = s[0] - p, n1 = s[1] - p, n2 = s[2] - p, n5 = s[5];
    double e0 = epd[0], e1 = epd[1], e2 = epd[2], e5 = epd[5];
    double q2 = 1.5 * (n0 * n0 + n1 * n1 + n2 * n2) + 3. * (n5 * n5);
    double rs, seq;
    seq_rs(q2, seq, rs);
    double W00 = 0., W01 = 0., W05 = 0., W11 = 0., W15 = 0., W55 = 0.;
    bool corrected = false;
    double ncd = fma(1.5, fma(n2, dc2, fma(n1, dc1, n0 * dc0)), 3. * (n5 * dc5));
    double r2 = rs * rs;
#pragma unroll 1
    for (int it = 0; it < PLF_MAXIT; it++) {
        if (!RC) {
            ncd = fma(1.5, fma(n2, dc2, fma(n1, dc1, n0 * dc0)), 3. * (n5 * dc5));
            r2 = rs * rs;
        }
        const bool yplus = fma(2., ncd, q2 + qc) > ythr2;
        const double g = ncd * r2;
        const double beta = g * bfac;
        const double gam = yplus ? g * gfac : 0.;
        const double gam2 = gam + gam;
        const double w = r2 * wfac;
        const double t0 = w * n0, t1 = w * n1, t5 = w * n5;
        W00 = fma(t0, n0, W00); W01 = fma(t0, n1, W01); W05 = fma(t0, n5, W05);
        W11 = fma(t1, n1, W11); W15 = fma(t1, n5, W15); W55 = fma(t5, n5, W55);
        double f0 = fma(gam, n0, e0), f1e = fma(gam, n1, e1), f2 = fma(gam, n2, e2), f5 = fma(gam2, n5, e5);
        n0 = fma(-beta, n0, n0 + dc0); n1 = fma(-beta, n1, n1 + dc1); n2 = fma(-beta, n2, n2 + dc2); n5 = fma(-beta, n5, n5 + dc5);
        p += pc;
        if (RC) {
            const double omb = 1. - beta;
            q2 = fma(omb, fma(omb, q2, ncd + ncd), qc);
            ncd = fma(omb, ncd, qc);
            r2 = fast_rcp1(q2);   // for the next sub-step: issued ahead of the yield check below
            seq = q2 * fast_rsqrt1(q2 + 1e-300);
        } else {
            q2 = 1.5 * (n0 * n0 + n1 * n1 + n2 * n2) + 3. * (n5 * n5);
            seq_rs1(q2, seq, rs);
        }
        const double e2s = (2. / 3.) * ((f0 * f0 + f1e * f1e + f2 * f2) + 0.5 * (f5 * f5));
        const double lhs = seq - c0;
        if (lhs > 0. && lhs * lhs > K2 * e2s) {
            // radial-return correction (310-342), rare: out of line, state handed over through local memory so that
            // its temporaries do not take registers from the loop
            corrected = true;
            double st[10] = {n0, n1, n2, n5, p, f0, f1e, f2, f5, seq};
            iso2d_correct(m, st, xc, xstride);
            n0 = st[0]; n1 = st[1]; n2 = st[2]; n5 = st[3]; p = st[4];
            f0 = st[5]; f1e = st[6]; f2 = st[7]; f5 = st[8];
            q2 = 1.5 * (n0 * n0 + n1 * n1 + n2 * n2) + 3. * (n5 * n5);
            seq_rs(q2, seq, rs);
            if (RC) {
                ncd = fma(1.5, fma(n2, dc2, fma(n1, dc1, n0 * dc0)), 3. * (n5 * dc5));
                r2 = rs * rs;
            }
        }
        e0 = f0; e1 = f1e; e2 = f2; e5 = f5;
    }
    if (RC) q2 = 1.5 * (n0 * n0 + n1 * n1 + n2 * n2) + 3. * (n5 * n5);
    seq_rs(q2, seq, rs);   // full-precision seq for the returned yield-function value
    s[0] = n0 + p; s[1] = n1 + p; s[2] = n2 + p; s[3] = 0.; s[4] = 0.; s[5] = n5;
    epd[0] = e0; epd[1] = e1; epd[2] = e2; epd[5] = e5;   // epd[3], epd[4] stay zero
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(epd) * m.khard : 0.));
    double Wf[21];
#pragma unroll
    for (int k = 0; k < 21; k++) Wf[k] = 0.;
    Wf[sym_idx(0, 0)] = W00; Wf[sym_idx(0, 1)] = W01; Wf[sym_idx(0, 5)] = W05;
    Wf[sym_idx(1, 1)] = W11; Wf[sym_idx(1, 5)] = W15; Wf[sym_idx(5, 5)] = W55;
    Wf[sym_idx(0, 2)] = -W00 - W01;
    Wf[sym_idx(1, 2)] = -W01 - W11;
    Wf[sym_idx(2, 2)] = W00 + 2. * W01 + W11;
    Wf[sym_idx(2, 5)] = -W05 - W15;
    if (corrected) {
        Wf[sym_idx(0, 0)] += xc[0 * xstride];
        Wf[sym_idx(0, 1)] += xc[5 * xstride];
        Wf[sym_idx(0, 2)] += xc[4 * xstride];
        Wf[sym_idx(1, 1)] += xc[1 * xstride];
        Wf[sym_idx(1, 2)] += xc[3 * xstride];
        Wf[sym_idx(2, 2)] += xc[2 * xstride];
    }
    const double f2s = st_scal * (1. / (double)PLF_MAXIT);
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double c = m.CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)PLF_MAXIT * c - Wf[sym_idx(i, j)]) * f2s;
        }
    return 2 | (corrected ? 4 : 0);
}

// ------------------------------------------------------------------------------------------------------------
// FM 6: Hill-6p yield function (no Drucker / LHS term) on isotropic elasticity (BASELINE config 3), planar states
// sorted to their own loop like FM 4.  With g = grad q(s) (hill_qgrad: q is the quadratic form, seq = sqrt(q)) the
// flow direction is a = g / (2 seq); g is deviatoric in its normal part, so
//     CV a = G ghat / seq,   ghat = (g_0, g_1, g_2, g_3/2, g_4/2, g_5/2)
//     hh   = a.CV a + khard = (G / (2 q)) (Gam + kap q),   Gam = g.ghat,  kap = 2 khard / G
//     sig += CV deps_r - CV a (a.CV deps_r / hh)   <=>   s += cdr - beta ghat,   beta = (g.cdr) D,  D = 1 / (Gam + kap q)
//     depl = gam g,  gam = beta / (2G)  (when the elastic trial stress is outside the yield surface)
//     CV a (CV a)^T / hh = 2G D ghat ghat^T;  q(s + cdr) = q(s) + g.cdr + q(cdr)
// so a sub-step needs one reciprocal for the update and one reciprocal square root for the yield check, and the
// accumulated tangent sum lives on the index set {0,1,5} (ghat_2 = -ghat_0 - ghat_1).  ~80 FP64 instructions per
// sub-step against ~170 of the general loop.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hilliso_passA(const DevMat &m, double *s, const double *ep, const double *de, double *dp, double *ca,
                                             double &tc, double &tw, double &fy)
{
    HillStep P;
#pragma unroll
    for (int k = 0; k < 6; k++) { dp[k] = 0.; ca[k] = 0.; }
    tc = 1.; tw = 0.;
    if (iso_prelude<1>(m, s, ep, de, P, fy)) return 0;
    const double G = m.CV[21];
    double g[6], gh[6];
    hill_qgrad(m, s, g);
    gh[0] = g[0]; gh[1] = g[1]; gh[2] = g[2]; gh[3] = 0.5 * g[3]; gh[4] = 0.5 * g[4]; gh[5] = 0.5 * g[5];
    const double q = hill_q(m, s), qc = hill_q(m, P.cdr);
    double Gam = 0., gcd = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) { Gam = fma(g[k], gh[k], Gam); gcd = fma(g[k], P.cdr[k], gcd); }
    const double D = fast_rcp(fma(2. * m.khard / G, q, Gam));
    const double beta = gcd * D;
    const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
    const bool yplus = (q + gcd) + qc > ythr * ythr;   // epl_dot 1031-1041 on the elastic trial stress
    const double gam = yplus ? beta * (0.5 / G) : 0.;
    double eplt[6], st[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        dp[k] = gam * g[k];
        eplt[k] = ep[k] + dp[k];
        st[k] = fma(-beta, gh[k], s[k] + P.cdr[k]);
    }
    fy = fast_sqrt(hill_q(m, st)) - (m.sy + (m.khard != 0. ? eps_eq6_fast(eplt) * m.khard : 0.));
    if (fy > P.toler) return -1;   // 288-291: subdivide
#pragma unroll
    for (int k = 0; k < 6; k++) { s[k] = st[k]; ca[k] = gh[k]; }
    // grad_stiff = CV (1 - st_scal) [split] + (CV - 2G D ghat ghat^T) st_scal   (269, 343 with nsteps = 1)
    tc = (P.split ? (1. - P.st_scal) : 0.) + P.st_scal;
    tw = 2. * G * D * P.st_scal;
    return 1;
}

// radial-return correction of the planar Hill loop (material.py:310-342) on st = {s0, s1, s2, s5, eplt0, eplt1, eplt2,
// eplt5, seq}; xc as in iso2d_correct
static __device__ __noinline__ void hilliso2d_correct(const DevMat &m, double *st, double *xc, int xstride)
{
    const double seq = st[8];
    double eplt[6] = {st[4], st[5], st[6], 0., 0., st[7]};
    const double fyv = seq - (m.sy + eps_eq6(eplt) * m.khard);
    double sf[6] = {st[0], st[1], st[2], 0., 0., st[3]};
    double ex[6], t[6], x[6], dr[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ex[k] = sf[k] * fyv / seq;
        sf[k] -= ex[k];
        dr[k] = xc[(6 + k) * xstride];
    }
    matvec6(m.SV, ex, t);
    lstsq_3x6(dr, ex, x);
#pragma unroll
    for (int k = 0; k < 6; k++) xc[k * xstride] += x[k];
    st[0] = sf[0]; st[1] = sf[1]; st[2] = sf[2]; st[3] = sf[5];
    st[4] = eplt[0] + t[0]; st[5] = eplt[1] + t[1]; st[6] = eplt[2] + t[2]; st[7] = eplt[5] + t[5];
}

__device__ __forceinline__ int hill_passB_hilliso2d(const DevMat &m, double *s, double *epd, const double *de, double *xc, int xstride,
                                                    double *gs, double &fy)
{
    double st_scal, f1;
    double c0k, c1k, c2k, c5k, qc, c0, ythr2;   // CV deps_r / 50 on {0,1,2,5}
    {
        HillStep P;
        iso_prelude<1>(m, s, epd, de, P, fy);   // cannot be elastic: pass A deferred this element
#pragma unroll
        for (int k = 0; k < 6; k++) {
            P.dr[k] *= 1. / (double)PLF_MAXIT;   // 290
            xc[(6 + k) * xstride] = P.dr[k];
            xc[k * xstride] = 0.;
        }
        cv_mul_t<1>(m, P.dr, P.cdr);
        c0k = P.cdr[0]; c1k = P.cdr[1]; c2k = P.cdr[2]; c5k = P.cdr[5];
        qc = hill_q(m, P.cdr);
        c0 = m.sy + P.toler;
        const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
        ythr2 = ythr * ythr;
        st_scal = P.st_scal;
        f1 = P.split ? (1. - P.st_scal) : 0.;
    }
    const double G = m.CV[21];
    const double kap = 2. * m.khard / G, G2 = 2. * G, iG2 = 0.5 / G, K2 = m.khard * m.khard;
    const double h0 = m.hill[0], h1 = m.hill[1], h2 = m.hill[2], h5x6 = 6. * m.hill[5];
    double s0 = s[0], s1 = s[1], s2 = s[2], s5 = s[5];
    double e0 = epd[0], e1 = epd[1], e2 = epd[2], e5 = epd[5];
    // g = grad q(s), q = q(s): evaluated at the end of a sub-step for the yield check and re-used by the next one
    double u0 = h0 * (s0 - s1), u1 = h1 * (s1 - s2), u2 = h2 * (s2 - s0);
    double g0 = u0 - u2, g1 = u1 - u0, g2 = u2 - u1, g5 = h5x6 * s5;
    double q = 0.5 * (u0 * (s0 - s1) + u1 * (s1 - s2) + u2 * (s2 - s0) + g5 * s5);
    double W00 = 0., W01 = 0., W05 = 0., W11 = 0., W15 = 0., W55 = 0.;
    double seq = 0.;
    bool corrected = false;
#pragma unroll 1
    for (int it = 0; it < PLF_MAXIT; it++) {
        const double gh5 = 0.5 * g5;
        const double Gam = fma(g5, gh5, fma(g2, g2, fma(g1, g1, g0 * g0)));
        const double gcd = fma(g5, c5k, fma(g2, c2k, fma(g1, c1k, g0 * c0k)));
        const bool yplus = (q + gcd) + qc > ythr2;
        const double D = fast_rcp1(fma(kap, q, Gam));
        const double beta = gcd * D;
        const double gam = yplus ? beta * iG2 : 0.;
        const double w = G2 * D;
        const double t0 = w * g0, t1 = w * g1, t5 = w * gh5;
        W00 = fma(t0, g0, W00); W01 = fma(t0, g1, W01); W05 = fma(t0, gh5, W05);
        W11 = fma(t1, g1, W11); W15 = fma(t1, gh5, W15); W55 = fma(t5, gh5, W55);
        double f0 = fma(gam, g0, e0), f1e = fma(gam, g1, e1), f2 = fma(gam, g2, e2), f5 = fma(gam, g5, e5);
        s0 = fma(-beta, g0, s0 + c0k); s1 = fma(-beta, g1, s1 + c1k); s2 = fma(-beta, g2, s2 + c2k); s5 = fma(-beta, gh5, s5 + c5k);
        const double d01 = s0 - s1, d12 = s1 - s2, d20 = s2 - s0;
        u0 = h0 * d01; u1 = h1 * d12; u2 = h2 * d20;
        g0 = u0 - u2; g1 = u1 - u0; g2 = u2 - u1; g5 = h5x6 * s5;
        q = 0.5 * fma(g5, s5, fma(u2, d20, fma(u1, d12, u0 * d01)));
        seq = q * fast_rsqrt1(q + 1e-300);
        const double e2s = (2. / 3.) * ((f0 * f0 + f1e * f1e + f2 * f2) + 0.5 * (f5 * f5));
        const double lhs = seq - c0;
        if (lhs > 0. && lhs * lhs > K2 * e2s) {
            corrected = true;
            double st[9] = {s0, s1, s2, s5, f0, f1e, f2, f5, seq};
            hilliso2d_correct(m, st, xc, xstride);
            s0 = st[0]; s1 = st[1]; s2 = st[2]; s5 = st[3];
            f0 = st[4]; f1e = st[5]; f2 = st[6]; f5 = st[7];
            u0 = h0 * (s0 - s1); u1 = h1 * (s1 - s2); u2 = h2 * (s2 - s0);
            g0 = u0 - u2; g1 = u1 - u0; g2 = u2 - u1; g5 = h5x6 * s5;
            q = 0.5 * (u0 * (s0 - s1) + u1 * (s1 - s2) + u2 * (s2 - s0) + g5 * s5);
            seq = fast_sqrt(q);
        }
        e0 = f0; e1 = f1e; e2 = f2; e5 = f5;
    }
    seq = fast_sqrt(q);   // full-precision seq for the returned yield-function value
    s[0] = s0; s[1] = s1; s[2] = s2; s[3] = 0.; s[4] = 0.; s[5] = s5;
    epd[0] = e0; epd[1] = e1; epd[2] = e2; epd[5] = e5;   // epd[3], epd[4] stay zero
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(epd) * m.khard : 0.));
    double Wf[21];
#pragma unroll
    for (int k = 0; k < 21; k++) Wf[k] = 0.;
    Wf[sym_idx(0, 0)] = W00; Wf[sym_idx(0, 1)] = W01; Wf[sym_idx(0, 5)] = W05;
    Wf[sym_idx(1, 1)] = W11; Wf[sym_idx(1, 5)] = W15; Wf[sym_idx(5, 5)] = W55;
    Wf[sym_idx(0, 2)] = -W00 - W01;
    Wf[sym_idx(1, 2)] = -W01 - W11;
    Wf[sym_idx(2, 2)] = W00 + 2. * W01 + W11;
    Wf[sym_idx(2, 5)] = -W05 - W15;
    if (corrected) {
        Wf[sym_idx(0, 0)] += xc[0 * xstride];
        Wf[sym_idx(0, 1)] += xc[5 * xstride];
        Wf[sym_idx(0, 2)] += xc[4 * xstride];
        Wf[sym_idx(1, 1)] += xc[1 * xstride];
        Wf[sym_idx(1, 2)] += xc[3 * xstride];
        Wf[sym_idx(2, 2)] += xc[2 * xstride];
    }
    const double f2s = st_scal * (1. / (double)PLF_MAXIT);
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double c = m.CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)PLF_MAXIT * c - Wf[sym_idx(i, j)]) * f2s;
        }
    return 2 | (corrected ? 4 : 0);
}

// ------------------------------------------------------------------------------------------------------------
// FM 7: Hill-6p / J2 (no Drucker / LHS term) on any block-structured symmetric CV that is not isotropic in the 3-d sense
// -- the plane-stress CV of Element.__init__ (model.py:274-283, zero third row and column: most 2-d examples of the
// reference), cubic C11 / C12 / C44 -- with planar states sorted to their own loop.  The identities of FM 6 without the
// isotropy:  a = g / (2 seq),  CV a = cg / (2 seq) with cg = CV g,
//     hh = a.CV a + khard = (Gc + 4 khard q) / (4 q),   Gc = g.cg
//     sig += CV deps_r - CV a (a.CV deps_r / hh)   <=>   s += cdr - beta cg,   beta = (g.cdr) D,  D = 1 / (Gc + 4 khard q)
//     depl = beta g (when the elastic trial stress is outside the yield surface),   CV a (CV a)^T / hh = D cg cg^T
// One reciprocal and one reciprocal square root per sub-step; the tangent sum lives on {0,1,2,5} (10 entries).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hillblk_passA(const DevMat &m, double *s, const double *ep, const double *de, double *dp, double *ca,
                                             double &tc, double &tw, double &fy)
{
    HillStep P;
#pragma unroll
    for (int k = 0; k < 6; k++) { dp[k] = 0.; ca[k] = 0.; }
    tc = 1.; tw = 0.;
    if (iso_prelude<1>(m, s, ep, de, P, fy)) return 0;
    double g[6], cg[6];
    hill_qgrad(m, s, g);
    cv_mul_t<1>(m, g, cg);
    const double q = hill_q(m, s), qc = hill_q(m, P.cdr);
    double Gc = 0., gcd = 0.;
#pragma unroll
    for (int k = 0; k < 6; k++) { Gc = fma(g[k], cg[k], Gc); gcd = fma(g[k], P.cdr[k], gcd); }
    const double D = fast_rcp(fma(4. * m.khard, q, Gc));
    const double beta = gcd * D;
    const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
    const bool yplus = (q + gcd) + qc > ythr * ythr;   // epl_dot 1031-1041 on the elastic trial stress
    const double gam = yplus ? beta : 0.;
    double eplt[6], st[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        dp[k] = gam * g[k];
        eplt[k] = ep[k] + dp[k];
        st[k] = fma(-beta, cg[k], s[k] + P.cdr[k]);
    }
    fy = fast_sqrt(hill_q(m, st)) - (m.sy + (m.khard != 0. ? eps_eq6_fast(eplt) * m.khard : 0.));
    if (fy > P.toler) return -1;   // 288-291: subdivide
#pragma unroll
    for (int k = 0; k < 6; k++) { s[k] = st[k]; ca[k] = cg[k]; }
    // grad_stiff = CV (1 - st_scal) [split] + (CV - D cg cg^T) st_scal   (269, 343 with nsteps = 1)
    tc = (P.split ? (1. - P.st_scal) : 0.) + P.st_scal;
    tw = D * P.st_scal;
    return 1;
}

__device__ __forceinline__ int hill_passB_blk2d(const DevMat &m, double *s, double *epd, const double *de, double *xc, int xstride,
                                                double *gs, double &fy)
{
    double st_scal, f1;
    double c0k, c1k, c2k, c5k, qc, c0, ythr2;   // CV deps_r / 50 on {0,1,2,5}
    {
        HillStep P;
        iso_prelude<1>(m, s, epd, de, P, fy);   // cannot be elastic: pass A deferred this element
#pragma unroll
        for (int k = 0; k < 6; k++) {
            P.dr[k] *= 1. / (double)PLF_MAXIT;   // 290
            xc[(6 + k) * xstride] = P.dr[k];
            xc[k * xstride] = 0.;
        }
        cv_mul_t<1>(m, P.dr, P.cdr);
        c0k = P.cdr[0]; c1k = P.cdr[1]; c2k = P.cdr[2]; c5k = P.cdr[5];
        qc = hill_q(m, P.cdr);
        c0 = m.sy + P.toler;
        const double ythr = P.sflow0 + PLF_YF_TOLERANCE;
        ythr2 = ythr * ythr;
        st_scal = P.st_scal;
        f1 = P.split ? (1. - P.st_scal) : 0.;
    }
    const double *C = m.CV;
    const double C00 = C[0], C01 = C[1], C02 = C[2], C11 = C[7], C12 = C[8], C22 = C[14], C55 = C[35];
    const double kh4 = 4. * m.khard, K2 = m.khard * m.khard;
    const double h0 = m.hill[0], h1 = m.hill[1], h2 = m.hill[2], h5x6 = 6. * m.hill[5];
    double s0 = s[0], s1 = s[1], s2 = s[2], s5 = s[5];
    double e0 = epd[0], e1 = epd[1], e2 = epd[2], e5 = epd[5];
    double u0 = h0 * (s0 - s1), u1 = h1 * (s1 - s2), u2 = h2 * (s2 - s0);
    double g0 = u0 - u2, g1 = u1 - u0, g2 = u2 - u1, g5 = h5x6 * s5;
    double q = 0.5 * (u0 * (s0 - s1) + u1 * (s1 - s2) + u2 * (s2 - s0) + g5 * s5);
    double W00 = 0., W01 = 0., W02 = 0., W05 = 0., W11 = 0., W12 = 0., W15 = 0., W22 = 0., W25 = 0., W55 = 0.;
    double seq = 0.;
    bool corrected = false;
#pragma unroll 1
    for (int it = 0; it < PLF_MAXIT; it++) {
        const double cg0 = fma(C02, g2, fma(C01, g1, C00 * g0));
        const double cg1 = fma(C12, g2, fma(C11, g1, C01 * g0));
        const double cg2 = fma(C22, g2, fma(C12, g1, C02 * g0));
        const double cg5 = C55 * g5;
        const double Gc = fma(g5, cg5, fma(g2, cg2, fma(g1, cg1, g0 * cg0)));
        const double gcd = fma(g5, c5k, fma(g2, c2k, fma(g1, c1k, g0 * c0k)));
        const bool yplus = (q + gcd) + qc > ythr2;
        const double D = fast_rcp1(fma(kh4, q, Gc));
        const double beta = gcd * D;
        const double gam = yplus ? beta : 0.;
        const double t0 = D * cg0, t1 = D * cg1, t2 = D * cg2, t5 = D * cg5;
        W00 = fma(t0, cg0, W00); W01 = fma(t0, cg1, W01); W02 = fma(t0, cg2, W02); W05 = fma(t0, cg5, W05);
        W11 = fma(t1, cg1, W11); W12 = fma(t1, cg2, W12); W15 = fma(t1, cg5, W15);
        W22 = fma(t2, cg2, W22); W25 = fma(t2, cg5, W25); W55 = fma(t5, cg5, W55);
        double f0 = fma(gam, g0, e0), f1e = fma(gam, g1, e1), f2 = fma(gam, g2, e2), f5 = fma(gam, g5, e5);
        s0 = fma(-beta, cg0, s0 + c0k); s1 = fma(-beta, cg1, s1 + c1k); s2 = fma(-beta, cg2, s2 + c2k); s5 = fma(-beta, cg5, s5 + c5k);
        const double d01 = s0 - s1, d12 = s1 - s2, d20 = s2 - s0;
        u0 = h0 * d01; u1 = h1 * d12; u2 = h2 * d20;
        g0 = u0 - u2; g1 = u1 - u0; g2 = u2 - u1; g5 = h5x6 * s5;
        q = 0.5 * fma(g5, s5, fma(u2, d20, fma(u1, d12, u0 * d01)));
        seq = q * fast_rsqrt1(q + 1e-300);
        const double e2s = (2. / 3.) * ((f0 * f0 + f1e * f1e + f2 * f2) + 0.5 * (f5 * f5));
        const double lhs = seq - c0;
        if (lhs > 0. && lhs * lhs > K2 * e2s) {
            corrected = true;
            double st[9] = {s0, s1, s2, s5, f0, f1e, f2, f5, seq};
            hilliso2d_correct(m, st, xc, xstride);
            s0 = st[0]; s1 = st[1]; s2 = st[2]; s5 = st[3];
            f0 = st[4]; f1e = st[5]; f2 = st[6]; f5 = st[7];
            u0 = h0 * (s0 - s1); u1 = h1 * (s1 - s2); u2 = h2 * (s2 - s0);
            g0 = u0 - u2; g1 = u1 - u0; g2 = u2 - u1; g5 = h5x6 * s5;
            q = 0.5 * (u0 * (s0 - s1) + u1 * (s1 - s2) + u2 * (s2 - s0) + g5 * s5);
            seq = fast_sqrt(q);
        }
        e0 = f0; e1 = f1e; e2 = f2; e5 = f5;
    }
    seq = fast_sqrt(q);   // full-precision seq for the returned yield-function value
    s[0] = s0; s[1] = s1; s[2] = s2; s[3] = 0.; s[4] = 0.; s[5] = s5;
    epd[0] = e0; epd[1] = e1; epd[2] = e2; epd[5] = e5;   // epd[3], epd[4] stay zero
    fy = seq - (m.sy + (m.khard != 0. ? eps_eq6(epd) * m.khard : 0.));
    double Wf[21];
#pragma unroll
    for (int k = 0; k < 21; k++) Wf[k] = 0.;
    Wf[sym_idx(0, 0)] = W00; Wf[sym_idx(0, 1)] = W01; Wf[sym_idx(0, 2)] = W02; Wf[sym_idx(0, 5)] = W05;
    Wf[sym_idx(1, 1)] = W11; Wf[sym_idx(1, 2)] = W12; Wf[sym_idx(1, 5)] = W15;
    Wf[sym_idx(2, 2)] = W22; Wf[sym_idx(2, 5)] = W25; Wf[sym_idx(5, 5)] = W55;
    if (corrected) {
        Wf[sym_idx(0, 0)] += xc[0 * xstride];
        Wf[sym_idx(0, 1)] += xc[5 * xstride];
        Wf[sym_idx(0, 2)] += xc[4 * xstride];
        Wf[sym_idx(1, 1)] += xc[1 * xstride];
        Wf[sym_idx(1, 2)] += xc[3 * xstride];
        Wf[sym_idx(2, 2)] += xc[2 * xstride];
    }
    const double f2s = st_scal * (1. / (double)PLF_MAXIT);
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            const double c = m.CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)PLF_MAXIT * c - Wf[sym_idx(i, j)]) * f2s;
        }
    return 2 | (corrected ? 4 : 0);
}

}  // namespace plf
