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

// accumulator of the symmetric tangent sum: 21 registers, or 21 shared-memory words per thread (stride = block size)
struct WRegs {
    double w[21];
    __device__ __forceinline__ double get(int k) const { return w[k]; }
    __device__ __forceinline__ void set(int k, double v) { w[k] = v; }
};
struct WShared {
    double *p;
    int stride;
    __device__ __forceinline__ double get(int k) const { return p[k * stride]; }
    __device__ __forceinline__ void set(int k, double v) { p[k * stride] = v; }
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
    if (FM == 2) {
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
//   W   : accumulator policy (WRegs / WShared) for sum_k ca ca^T / hh (+ correction blocks)
//   gs  : out, symmetric 21.
// Returns branch code 2 (| 4 when a radial-return correction fired).
// Fast modes of the sub-step loop, chosen on the host for single-material batches (the flags are compile-time, so the
// uniform branches and the dead general-case code disappear from the loop):
//   FM 0 : general (any CV, Drucker / LHS terms)
//   FM 1 : block-structured symmetric CV, no hydrostatic term            (DevMat.cv_block && cv_sym && !has_d)
//   FM 2 : FM 1 and all Hill parameters equal 1 (J2): q = 1.5 |dev s|^2 + 3 |shear|^2, a = 1.5 dev(s) / seq
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
    if (FM == 2) {
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
    if (FM == 2) {
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

}  // namespace plf
