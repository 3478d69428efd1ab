// plf_device.cuh -- device-side constitutive math shared by the sm_100a kernels.
//
// FP64 throughout.  One "yield model" policy per material family provides seq / yf / full_yf / fgrad; the
// return-mapping core (response_core) is written once against that policy:
//   HillModel : thread-per-element, analytic Hill-6p / J2 (+ Drucker / LHS term)   material.py:576-676, 816-858
//   SvcModel  : warp-per-element, RBF support-vector decision function + gradient     material.py:398-405, 765-815
//               with the table in shared memory and __shfl_xor reductions; ML_full_yf (414-516) incl. Brent.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/plfb200.h"

namespace plf {

struct DevMat {
    int32_t kind, nsv, dev_only, tresca, lhs, nsv_pad;
    int32_t cv_block;   // CV = [3x3 normal block | diagonal shear] (cv_mul, hill_fast.cuh)
    int32_t has_d;      // any Drucker / LHS weight non-zero
    int32_t cv_sym;     // CV is exactly symmetric: (CV a) . deps == a . (CV deps)
    int32_t cv_iso;     // cv_block && cv_sym && C11 - C12 == 2 C44 (to 1e-13): isotropic elasticity, plane strain / 3-d
    int32_t svc_princ;  // SVC trained on (seq_J2, polar angle) of the principal stresses: the sdim = 3 ML yield function
    double sy, khard;
    double hill[6];
    double d[3];     // I1 weights (calc_seq 645-650)
    double d3[3];    // constant part of the gradient (calc_fgrad 829-832)
    double CV[36];   // Element.CV (plane-strain or plane-stress variant)
    double SV[36];   // compliance used by the radial-return correction (material.py:315-320)
    double E, nu;
    double iso_ihh;   // 1 / (3 C44 + khard): the constant hardening denominator of the isotropic J2 modes
    double gamma, rho, scale_seq, inv_scale_seq;
    const double *sv;    // device, SoA [6][nsv_pad]
    const double *dual;  // device, [nsv_pad] (zero padded)
    double bar_m1[36], bar_m2[36], bar_exp;
    int32_t bar_exp_int;   // barlat_exp when it is a whole number in [2, 64] (Yld2004-18p: 6 for bcc, 8 for fcc), else 0
    int32_t svc_nf;        // feature columns of the device SV table: 6 (stress), 12 (stress + plastic strain: work-hardening SVC)
    double inv_scale_wh;   // 1 / scale_wh (material.py:2343)
    double khard_fac;      // scale_seq / scale_wh: khard = max(0, -sum_k dK/dx[6 + k] * khard_fac) (material.py:808-814)
};

// symmetric 6x6 packing: (i<=j) -> index
__host__ __device__ constexpr int sym_idx(int i, int j)
{
    return (i <= j) ? (i * 6 - (i * (i - 1)) / 2 + (j - i)) : (j * 6 - (j * (j - 1)) / 2 + (i - j));
}

__device__ __forceinline__ double eps_eq6(const double *e)
{
    // basic.py:328-360
    double a = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    double b = e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
    return sqrt(2. * (a + 0.5 * b) / 3.);
}

__device__ __forceinline__ void matvec6(const double *__restrict__ C, const double *v, double *o)
{
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < 6; k++) s += C[6 * i + k] * v[k];
        o[i] = s;
    }
}

// eigenvalues of a symmetric 3x3 in Voigt order (11,22,33,23,13,12): cyclic Jacobi (Barlat, material.py:695-696)
__device__ inline void sym3_eig(const double *v, double *w)
{
    double a00 = v[0], a11 = v[1], a22 = v[2], a12 = v[3], a02 = v[4], a01 = v[5];
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = fabs(a01) + fabs(a02) + fabs(a12);
        double dia = fabs(a00) + fabs(a11) + fabs(a22);
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
        // rotate (0,1)
        if (a01 != 0.) {
            double th = (a11 - a00) / (2. * a01);
            double t = copysign(1., th) / (fabs(th) + sqrt(th * th + 1.));
            double c = rsqrt(t * t + 1.), s = t * c;
            double n00 = a00 - t * a01, n11 = a11 + t * a01;
            double n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
            a00 = n00; a11 = n11; a02 = n02; a12 = n12; a01 = 0.;
        }
        if (a02 != 0.) {
            double th = (a22 - a00) / (2. * a02);
            double t = copysign(1., th) / (fabs(th) + sqrt(th * th + 1.));
            double c = rsqrt(t * t + 1.), s = t * c;
            double n00 = a00 - t * a02, n22 = a22 + t * a02;
            double n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
            a00 = n00; a22 = n22; a01 = n01; a12 = n12; a02 = 0.;
        }
        if (a12 != 0.) {
            double th = (a22 - a11) / (2. * a12);
            double t = copysign(1., th) / (fabs(th) + sqrt(th * th + 1.));
            double c = rsqrt(t * t + 1.), s = t * c;
            double n11 = a11 - t * a12, n22 = a22 + t * a12;
            double n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
            a11 = n11; a22 = n22; a01 = n01; a02 = n02; a12 = 0.;
        }
    }
    w[0] = a00; w[1] = a11; w[2] = a22;
}

// ------------------------------------------------------------------------------------------------------------
// Hill-6p / J2 equivalent stress (material.py:650-673) and its gradient (826-845)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double hill_seq(const DevMat &m, const double *s)
{
    if (m.tresca) {   // material.py:630-632: max - min principal stress, no hydrostatic term
        double w[3];
        sym3_eig(s, w);
        return fmax(w[0], fmax(w[1], w[2])) - fmin(w[0], fmin(w[1], w[2]));
    }
    double I1 = (s[0] * m.d[0] + s[1] * m.d[1] + s[2] * m.d[2]) / 3.;
    double d01 = s[0] - s[1], d12 = s[1] - s[2], d20 = s[2] - s[0];
    double I2 = m.hill[0] * (d01 * d01) + m.hill[1] * (d12 * d12) + m.hill[2] * (d20 * d20) +
                6. * m.hill[3] * (s[3] * s[3]) + 6. * m.hill[4] * (s[4] * s[4]) + 6. * m.hill[5] * (s[5] * s[5]);
    return sqrt(0.5 * I2) + I1;
}

__device__ __forceinline__ double j2_seq(const double *s)
{
    // calc_seq for an elastic material: h = 1, d = 0 (material.py:640-643)
    double d01 = s[0] - s[1], d12 = s[1] - s[2], d20 = s[2] - s[0];
    double I2 = d01 * d01 + d12 * d12 + d20 * d20 + 6. * (s[3] * s[3]) + 6. * (s[4] * s[4]) + 6. * (s[5] * s[5]);
    return sqrt(0.5 * I2);
}

// |d|^n for a whole exponent by binary powering (a handful of multiplications instead of a ~150-instruction pow())
__device__ __forceinline__ double powi_abs(double ad, int n)
{
    // the exponents of Yld2004-18p in practice (a = 8 for fcc, 6 for bcc) and a - 1 for the gradient: straight-line code
    const double x2 = ad * ad, x4 = x2 * x2;
    if (n == 8) return x4 * x4;
    if (n == 7) return x4 * (x2 * ad);
    if (n == 6) return x4 * x2;
    if (n == 5) return x4 * ad;
    double r = 1., b = ad;
    while (n) {
        if (n & 1) r *= b;
        b *= b;
        n >>= 1;
    }
    return r;
}

// x^(1/n), x >= 0: square roots when n is a power of two (n = 8: three of them)
__device__ __forceinline__ double root_n(double x, int n, double a)
{
    if (n > 0 && (n & (n - 1)) == 0) {
        while (n > 1) { x = sqrt(x); n >>= 1; }
        return x;
    }
    return pow(x, 1. / a);
}

static __device__ __noinline__ double barlat_seq(const DevMat &m, const double *sv)
{
    // material.py:678-702
    double p = (sv[0] + sv[1] + sv[2]) / 3.;
    double sd[6] = {sv[0] - p, sv[1] - p, sv[2] - p, sv[3], sv[4], sv[5]};
    double st1[6], st2[6], p1[3], p2[3];
    matvec6(m.bar_m1, sd, st1);
    matvec6(m.bar_m2, sd, st2);
    sym3_eig(st1, p1);
    sym3_eig(st2, p2);
    double s = 0.;
    const int n = m.bar_exp_int;
    if (n) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) s += powi_abs(fabs(p1[i] - p2[j]), n);
        return root_n(0.25 * s, n, m.bar_exp);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) s += pow(fabs(p1[i] - p2[j]), m.bar_exp);
    return pow(0.25 * s, 1. / m.bar_exp);
}

// eigenvalues w[e] and orthonormal eigenvectors V[i][e] of a symmetric 3x3 given in Voigt order: cyclic Jacobi
__device__ inline void sym3_eigv(const double *v, double *w, double V[3][3])
{
    double A[3][3] = {{v[0], v[5], v[4]}, {v[5], v[1], v[3]}, {v[4], v[3], v[2]}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1. : 0.;
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.) continue;
                const double th = (A[q][q] - A[p][p]) / (2. * A[p][q]);
                const double t = copysign(1., th) / (fabs(th) + sqrt(th * th + 1.));
                const double c = rsqrt(t * t + 1.), sn = t * c;
                const int r = 3 - p - q;
                const double apq = A[p][q], arp = A[r][p], arq = A[r][q];
                A[p][p] -= t * apq;
                A[q][q] += t * apq;
                A[p][q] = A[q][p] = 0.;
                A[r][p] = A[p][r] = c * arp - sn * arq;
                A[r][q] = A[q][r] = sn * arp + c * arq;
                for (int k = 0; k < 3; k++) {
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - sn * vq;
                    V[k][q] = sn * vp + c * vq;
                }
            }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

// Gradient of the Yld2004-18p equivalent stress (material.py:678-702) with respect to the Voigt stress.  The reference
// has no analytic gradient for Barlat (calc_fgrad raises, material.py:822-823); this is the chain rule through
//   seq = (phi / 4)^(1/a),  phi = sum_ij |S'_i - S''_j|^a,  S' / S'' = eigenvalues of L' dev(sig) / L'' dev(sig),
// with d(lambda_e) / d(Voigt k) = (v1^2, v2^2, v3^2, 2 v2 v3, 2 v1 v3, 2 v1 v2) of eigenvector e (phi is a symmetric
// function of the eigenvalues, so any orthonormal eigenbasis gives the gradient, also for repeated eigenvalues).
// Returns seq.
static __device__ __noinline__ double barlat_seq_grad(const DevMat &m, const double *sv, double *g)
{
    const double p = (sv[0] + sv[1] + sv[2]) / 3.;
    const double sd[6] = {sv[0] - p, sv[1] - p, sv[2] - p, sv[3], sv[4], sv[5]};
    double st1[6], st2[6], p1[3], p2[3], V1[3][3], V2[3][3];
    matvec6(m.bar_m1, sd, st1);
    matvec6(m.bar_m2, sd, st2);
    sym3_eigv(st1, p1, V1);
    sym3_eigv(st2, p2, V2);
    const double a = m.bar_exp;
    double phi = 0., g1[3] = {0., 0., 0.}, g2[3] = {0., 0., 0.};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double d = p1[i] - p2[j], ad = fabs(d);
            const double pw = m.bar_exp_int ? powi_abs(ad, m.bar_exp_int - 1) : pow(ad, a - 1.);
            phi += pw * ad;
            const double sg = copysign(pw, d);   // |d|^(a-1) sign(d)
            g1[i] += sg;
            g2[j] -= sg;
        }
    const double seq = m.bar_exp_int ? root_n(0.25 * phi, m.bar_exp_int, a) : pow(0.25 * phi, 1. / a);
    double E1[6] = {0., 0., 0., 0., 0., 0.}, E2[6] = {0., 0., 0., 0., 0., 0.};
    for (int e = 0; e < 3; e++) {
        const double x = V1[0][e], y = V1[1][e], z = V1[2][e];
        E1[0] += g1[e] * x * x; E1[1] += g1[e] * y * y; E1[2] += g1[e] * z * z;
        E1[3] += g1[e] * 2. * y * z; E1[4] += g1[e] * 2. * x * z; E1[5] += g1[e] * 2. * x * y;
        const double x2 = V2[0][e], y2 = V2[1][e], z2 = V2[2][e];
        E2[0] += g2[e] * x2 * x2; E2[1] += g2[e] * y2 * y2; E2[2] += g2[e] * z2 * z2;
        E2[3] += g2[e] * 2. * y2 * z2; E2[4] += g2[e] * 2. * x2 * z2; E2[5] += g2[e] * 2. * x2 * y2;
    }
    double h[6];
    for (int k = 0; k < 6; k++) {
        double t = 0.;
        for (int r = 0; r < 6; r++) t += m.bar_m1[6 * r + k] * E1[r] + m.bar_m2[6 * r + k] * E2[r];   // L'^T E1 + L''^T E2
        h[k] = t;
    }
    const double hm = (h[0] + h[1] + h[2]) / 3.;
    const double f = seq / phi;   // d seq / d phi = seq / (a phi); the factor a of d|d|^a cancels it
    g[0] = f * (h[0] - hm); g[1] = f * (h[1] - hm); g[2] = f * (h[2] - hm);
    g[3] = f * h[3]; g[4] = f * h[4]; g[5] = f * h[5];
    return seq;
}

__device__ __forceinline__ void hill_fgrad(const DevMat &m, const double *s, double seq, double *a)
{
    double p = (s[0] + s[1] + s[2]) / 3.;
    double s0 = s[0] - p, s1 = s[1] - p, s2 = s[2] - p;
    double h0 = m.hill[0], h1 = m.hill[1], h2 = m.hill[2];
    double i2 = 1. / (2. * seq), i1 = 1. / seq;
    a[0] = ((h0 + h2) * s0 - h0 * s1 - h2 * s2) * i2 + m.d3[0];
    a[1] = ((h1 + h0) * s1 - h0 * s0 - h1 * s2) * i2 + m.d3[1];
    a[2] = ((h2 + h1) * s2 - h2 * s0 - h1 * s1) * i2 + m.d3[2];
    a[3] = 3. * m.hill[3] * s[3] * i1;
    a[4] = 3. * m.hill[4] * s[4] * i1;
    a[5] = 3. * m.hill[5] * s[5] * i1;
}

struct HillModel {
    static constexpr bool is_ml = false;
    const DevMat &m;
    __device__ explicit HillModel(const DevMat &mm) : m(mm) {}
    __device__ __forceinline__ void set_epl(const double *) const {}
    __device__ __forceinline__ double khard() const { return m.khard; }
    __device__ __forceinline__ double seq(const double *s) const
    {
        return m.kind == PLF_MAT_ELASTIC ? j2_seq(s) : hill_seq(m, s);
    }
    __device__ __forceinline__ double yf(const double *s, double sflow) const { return seq(s) - sflow; }
    __device__ __forceinline__ double full_yf(const double *s, double sflow) const { return seq(s) - sflow; }
    __device__ __forceinline__ void fgrad(const double *s, double *a) const { hill_fgrad(m, s, hill_seq(m, s), a); }
};

// sig_princ (basic.py:107-179) of a Voigt stress: principal values in the reference's order.  The reference sorts
// the eigenvalues that numpy.linalg.eig returns by "major force axes": iev[i] = index of the eigenvector with the largest
// component along coordinate axis i, then the AXIS indices i are listed by ascending iev and used to index the
// eigenvalues (basic.py:156-170).  Here the eigen-decomposition is cyclic Jacobi, whose natural order keeps eigenvalue e
// with diagonal entry e (out of line: response_core calls it from a dozen sites); for stresses whose principal axes are close to the coordinate axes (every state of the
// reference's own sdim = 3 tests) both give sp = (s11, s22, s33)-like ordering.  For strongly rotated states the
// reference's result depends on LAPACK's eigenvalue order and is not reproduced bit for bit.
static __device__ __noinline__ void princ_ordered(const double *s, double *sp)
{
    double w[3], V[3][3];
    sym3_eigv(s, w, V);
    int iev[3];
    for (int i = 0; i < 3; i++) {
        int best = 0;
        double vb = fabs(V[i][0]);
        for (int e = 1; e < 3; e++)
            if (fabs(V[i][e]) > vb) { vb = fabs(V[i][e]); best = e; }
        iev[i] = best;
    }
    int j[3] = {0, 1, 2}, n = 0;
    for (int v = 0; v < 3; v++)
        for (int i = 0; i < 3; i++)
            if (iev[i] == v && n < 3) j[n++] = i;
    sp[0] = w[j[0]]; sp[1] = w[j[1]]; sp[2] = w[j[2]];
}

// sdim = 3 formulation (material.py:663-673, 836-838, 1045-1048): 3-parameter Hill / J2 equivalent stress on the ordered
// principal stresses; the flow direction is the gradient with respect to the principal stresses written into the
// normal Voigt components (shear components zero), exactly as epl_dot / C_tan build it.
struct HillPrincModel {
    static constexpr bool is_ml = false;
    const DevMat &m;
    __device__ explicit HillPrincModel(const DevMat &mm) : m(mm) {}
    __device__ __forceinline__ void set_epl(const double *) const {}
    __device__ __forceinline__ double khard() const { return m.khard; }
    __device__ __forceinline__ double seq_sp(const double *sp, const double *nrm) const
    {
        if (m.tresca) return fmax(sp[0], fmax(sp[1], sp[2])) - fmin(sp[0], fmin(sp[1], sp[2]));
        const double d12 = sp[0] - sp[1], d23 = sp[1] - sp[2], d31 = sp[2] - sp[0];
        const double I2 = 0.5 * (m.hill[0] * (d12 * d12) + m.hill[1] * (d23 * d23) + m.hill[2] * (d31 * d31));
        const double I1 = (nrm[0] * m.d[0] + nrm[1] * m.d[1] + nrm[2] * m.d[2]) / 3.;   // 650: Voigt normal components
        return sqrt(I2) + I1;
    }
    __device__ __forceinline__ double seq(const double *s) const
    {
        double sp[3];
        princ_ordered(s, sp);
        return seq_sp(sp, s);
    }
    __device__ __forceinline__ double yf(const double *s, double sflow) const { return seq(s) - sflow; }
    __device__ __forceinline__ double full_yf(const double *s, double sflow) const { return seq(s) - sflow; }
    __device__ __forceinline__ void fgrad(const double *s, double *a) const
    {
        // calc_fgrad(sig_princ(sig)[0]): a (3,) principal stress vector, seq = calc_seq(sp) with I1 from sp (607-611)
        double sp[3];
        princ_ordered(s, sp);
        const double sq = seq_sp(sp, sp);
        const double p = (sp[0] + sp[1] + sp[2]) / 3.;
        const double s0 = sp[0] - p, s1 = sp[1] - p, s2 = sp[2] - p;
        const double h0 = m.hill[0], h1 = m.hill[1], h2 = m.hill[2];
        a[0] = ((h0 + h2) * s0 - h0 * s1 - h2 * s2) / (2. * sq) + m.d3[0];
        a[1] = ((h1 + h0) * s1 - h0 * s0 - h1 * s2) / (2. * sq) + m.d3[1];
        a[2] = ((h2 + h1) * s2 - h2 * s0 - h1 * s1) / (2. * sq) + m.d3[2];
        a[3] = a[4] = a[5] = 0.;
    }
};

// Yld2004-18p yield model (thread per element): equivalent stress of the reference, analytic gradient from the chain
// rule (the reference raises in calc_fgrad for Barlat, so its return mapping cannot run; BASELINE config 5 asks for it)
struct BarlatModel {
    static constexpr bool is_ml = false;
    const DevMat &m;
    __device__ explicit BarlatModel(const DevMat &mm) : m(mm) {}
    __device__ __forceinline__ void set_epl(const double *) const {}
    __device__ __forceinline__ double khard() const { return m.khard; }
    __device__ __forceinline__ double seq(const double *s) const { return barlat_seq(m, s); }
    __device__ __forceinline__ double yf(const double *s, double sflow) const { return barlat_seq(m, s) - sflow; }
    __device__ __forceinline__ double full_yf(const double *s, double sflow) const { return barlat_seq(m, s) - sflow; }
    __device__ __forceinline__ void fgrad(const double *s, double *a) const { barlat_seq_grad(m, s, a); }
};

// ------------------------------------------------------------------------------------------------------------
// RBF support-vector yield function, one warp per element.  All lanes hold the same element state; the sum over
// support vectors is strided over lanes and combined with __shfl_xor (every lane ends with the full sum).
// Shared-memory table: sv[6][nsv_pad] (SoA: lane-consecutive, conflict free) + dual[nsv_pad] (zero padded).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// warp-aggregated append to a work list: one atomic per warp.  Must be called by all 32 lanes.
__device__ __forceinline__ void warp_append(bool pred, int32_t item, int32_t *list, unsigned int *count)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = item;
}

// Two-ended work list: class-0 items grow from the front (count[0]), class-1 items from the back (count[1]) of the
// same n-slot buffer -- the two classes are consumed by different kernels.  Must be called by all 32 lanes.
__device__ __forceinline__ void warp_append2(bool pred, bool back, int32_t item, int32_t *list, int64_t n, unsigned int *count)
{
    const unsigned m0 = __ballot_sync(0xffffffffu, pred && !back), m1 = __ballot_sync(0xffffffffu, pred && back);
    if ((m0 | m1) == 0) return;
    const int lane = threadIdx.x & 31;
    if (m0) {
        const int leader = __ffs(m0) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m0));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred && !back) list[base + __popc(m0 & ((1u << lane) - 1u))] = item;
    }
    if (m1) {
        const int leader = __ffs(m1) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(count + 1, (unsigned)__popc(m1));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred && back) list[n - 1 - (int64_t)(base + __popc(m1 & ((1u << lane) - 1u)))] = item;
    }
}

// exp(x) for x <= 0 with an instruction-level-parallel evaluation: magic-number rounding of x log2(e), two-constant
// Cody-Waite reduction, degree-13 Taylor polynomial in Estrin form (dependency depth 4 instead of the 11 of a Horner
// chain), exponent add.  Relative error < 3e-16 on |f| <= ln(2)/2; arguments below -700 return 0 (true value < 1e-304).
__device__ __forceinline__ double exp_neg(double x)
{
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int n = __double2loint(t);
    const double nd = t - 6755399441055744.0;
    double f = fma(nd, -6.93147180369123816490e-01, x);
    f = fma(nd, -1.90821492927058770002e-10, f);
    const double f2 = f * f, f4 = f2 * f2, f8 = f4 * f4;
    const double a0 = fma(f, 1., 1.);
    const double a1 = fma(f, 1.6666666666666666e-01, 0.5);
    const double a2 = fma(f, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double a3 = fma(f, 1.9841269841269841e-04, 1.3888888888888889e-03);
    const double a4 = fma(f, 2.7557319223985893e-06, 2.4801587301587302e-05);
    const double a5 = fma(f, 2.5052108385441720e-08, 2.7557319223985888e-07);
    const double a6 = fma(f, 1.6059043836821613e-10, 2.0876756987868100e-09);
    const double b0 = fma(a1, f2, a0), b1 = fma(a3, f2, a2), b2 = fma(a5, f2, a4);
    const double d0 = fma(b1, f4, b0), d1 = fma(a6, f4, b2);
    const double p = fma(d1, f8, d0);
    const double r = __longlong_as_double(__double_as_longlong(p) + ((long long)n << 52));
    return x < -700. ? 0. : r;
}

// sum_i dual_i exp(-gamma |x - sv_i|^2) over the support vectors of this lane (i = lane, lane + 32, ...), four support
// vectors per trip: four independent distance / exp chains in flight.  Out of line on purpose: inside the return-mapping
// core the loop competed with ~200 live registers and was scheduled as one dependent chain (ncu: 15 % of the FP64 pipe,
// stall reason "wait"); as a function of its own it gets its own register allocation.  np is a multiple of 128.
template <int NF>
static __device__ __noinline__ double svc_decision_lane(const double *__restrict__ s_sv, const double *__restrict__ s_dual, int np,
                                                       double gamma, const double *__restrict__ xin, int lane)
{
    double x[NF];
#pragma unroll
    for (int c = 0; c < NF; c++) x[c] = xin[c];
    double acc[4] = {0., 0., 0., 0.};
    const double ng = -gamma;
    for (int i = lane; i < np; i += 128) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i + 32 * u;
            double q0 = 0., q1 = 0., q2 = 0.;
#pragma unroll
            for (int c = 0; c < NF; c += 3) {
                const double t0 = x[c] - s_sv[c * np + k], t1 = x[c + 1] - s_sv[(c + 1) * np + k], t2 = x[c + 2] - s_sv[(c + 2) * np + k];
                q0 = fma(t0, t0, q0); q1 = fma(t1, t1, q1); q2 = fma(t2, t2, q2);
            }
            acc[u] = fma(s_dual[k], exp_neg(ng * ((q0 + q1) + q2)), acc[u]);
        }
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// the same loop for the gradient: g_c = sum_i dual_i exp(-gamma r_i) (-2 gamma) (x_c - sv_ic), two support vectors per trip
template <int NF>
static __device__ __noinline__ void svc_gradient_lane(const double *__restrict__ s_sv, const double *__restrict__ s_dual, int np,
                                                     double gamma, const double *__restrict__ xin, int lane, double *g)
{
    double x[NF], ga[NF], gb[NF];
#pragma unroll
    for (int c = 0; c < NF; c++) { x[c] = xin[c]; ga[c] = 0.; gb[c] = 0.; }
    const double ng = -gamma, m2g = -2. * gamma;
    for (int i = lane; i < np; i += 64) {
        const int k = i + 32;
        double a[NF], b[NF], ra0 = 0., ra1 = 0., rb0 = 0., rb1 = 0.;
#pragma unroll
        for (int c = 0; c < NF; c += 2) {
            a[c] = x[c] - s_sv[c * np + i]; a[c + 1] = x[c + 1] - s_sv[(c + 1) * np + i];
            b[c] = x[c] - s_sv[c * np + k]; b[c + 1] = x[c + 1] - s_sv[(c + 1) * np + k];
            ra0 = fma(a[c], a[c], ra0); ra1 = fma(a[c + 1], a[c + 1], ra1);
            rb0 = fma(b[c], b[c], rb0); rb1 = fma(b[c + 1], b[c + 1], rb1);
        }
        const double wa = s_dual[i] * exp_neg(ng * (ra0 + ra1)) * m2g, wb = s_dual[k] * exp_neg(ng * (rb0 + rb1)) * m2g;
#pragma unroll
        for (int c = 0; c < NF; c++) { ga[c] = fma(wa, a[c], ga[c]); gb[c] = fma(wb, b[c], gb[c]); }
    }
#pragma unroll
    for (int c = 0; c < NF; c++) g[c] = ga[c] + gb[c];
}

struct SvcModel {
    static constexpr bool is_ml = true;
    const DevMat &m;
    const double *s_sv;    // shared, [svc_nf][nsv_pad]
    const double *s_dual;  // shared, [nsv_pad]
    int lane;
    // Work-hardening SVC (material.py:2342-2346, 808-814): the plastic strain is part of the feature vector, and every
    // gradient evaluation sets Material.khard from the plastic-strain part of the gradient.  The reference keeps khard on the
    // Material object (one value for all elements, left behind by whichever element was visited last); here every element
    // starts from the khard the host set and carries its own value through its return mapping.
    mutable const double *cur_epl;   // plastic strain the next yield-function / gradient evaluation refers to
    mutable double kh;
    __device__ SvcModel(const DevMat &mm, const double *sv, const double *dual, int l)
        : m(mm), s_sv(sv), s_dual(dual), lane(l), cur_epl(nullptr), kh(mm.svc_nf == 12 ? mm.khard : 0.) {}
    __device__ __forceinline__ void set_epl(const double *e) const { cur_epl = e; }
    // material.py:812-814: without work-hardening features every gradient call resets khard to 0
    __device__ __forceinline__ double khard() const { return kh; }
    // sdim = 3 (svc_princ): calc_seq is the 3-parameter form on the ordered principal stresses (material.py:663-673)
    __device__ __forceinline__ double seq(const double *s) const { return m.svc_princ ? HillPrincModel(m).seq(s) : hill_seq(m, s); }

    // create_scaled_input, sdim = 3 (material.py:2331-2334): (sig_eq_j2 / scale_seq - 1, sig_polar_ang / pi) of the
    // principal stresses (basic.py:30-104; a_vec, b_vec of basic.py:20-23); table columns 2..5 are zero
    __device__ __forceinline__ void features_sp(const double *sp, double *x) const
    {
        const double d12 = sp[0] - sp[1], d23 = sp[1] - sp[2], d31 = sp[2] - sp[0];
        x[0] = sqrt(0.5 * (d12 * d12 + d23 * d23 + d31 * d31)) / m.scale_seq - 1.;
        const double hyd = (sp[0] + sp[1] + sp[2]) / 3.;
        const double dv[3] = {sp[0] - hyd, sp[1] - hyd, sp[2] - hyd};
        double vn = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
        if (vn < 1.e-4) vn = 1.;
        const double r15 = sqrt(1.5), r2 = sqrt(2.);
        const double dsa = dv[0] / vn * (1. / r15) + dv[1] / vn * (-0.5 / r15) + dv[2] / vn * (-0.5 / r15);
        const double dsb = dv[1] / vn * (0.5 * r2) + dv[2] / vn * (-0.5 * r2);
        x[1] = atan2(dsb, dsa) / 3.141592653589793;
        x[2] = x[3] = x[4] = x[5] = 0.;
    }

    __device__ __forceinline__ void features(const double *s, double *x) const
    {
        if (m.svc_princ) {   // a Voigt stress goes through sig_princ first (basic.py:49-52, 86-89)
            double sp[3];
            princ_ordered(s, sp);
            features_sp(sp, x);
            return;
        }
        // create_scaled_input, sdim = 6 (material.py:2336-2340)
        double p = m.dev_only ? (s[0] + s[1] + s[2]) / 3. : 0.;
        x[0] = (s[0] - p) / m.scale_seq; x[1] = (s[1] - p) / m.scale_seq; x[2] = (s[2] - p) / m.scale_seq;
        x[3] = s[3] / m.scale_seq; x[4] = s[4] / m.scale_seq; x[5] = s[5] / m.scale_seq;
        if (m.svc_nf == 12) {
            // work-hardening features (2342-2346): epl / scale_wh; accumulated strain, maximum stress and the flag are 0 on
            // this path (the defaults of calc_yf, 348-350) and are folded into the dual coefficients on the host
#pragma unroll
            for (int k = 0; k < 6; k++) x[6 + k] = cur_epl ? cur_epl[k] * m.inv_scale_wh : 0.;
        }
    }
    // decision_function (material.py:405): sum_i dual_i exp(-gamma |x - sv_i|^2) + rho
    __device__ __forceinline__ double decision_x(const double *x) const
    {
        const double part = (m.svc_nf == 12) ? svc_decision_lane<12>(s_sv, s_dual, m.nsv_pad, m.gamma, x, lane)
                                             : svc_decision_lane<6>(s_sv, s_dual, m.nsv_pad, m.gamma, x, lane);
        return warp_sum(part) + m.rho;
    }
    __device__ __forceinline__ double yf(const double *s, double /*sflow*/) const
    {
        double x[12];
        features(s, x);
        return decision_x(x);
    }
    // gradient of the decision function w.r.t. stress (material.py:768-778, 802-807)
    __device__ __forceinline__ void fgrad(const double *s, double *a) const
    {
        if (m.svc_princ) { fgrad_princ(s, a); return; }
        double x[12], g[12];
        features(s, x);
        if (m.svc_nf == 12) {
            svc_gradient_lane<12>(s_sv, s_dual, m.nsv_pad, m.gamma, x, lane, g);
            double hk = 0.;
#pragma unroll
            for (int k = 0; k < 6; k++) hk -= warp_sum(g[6 + k]) * m.khard_fac;   // 808-809
            kh = fmax(hk, 0.);                                                    // 812-814: no strain softening
        } else {
            svc_gradient_lane<6>(s_sv, s_dual, m.nsv_pad, m.gamma, x, lane, g);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) a[k] = warp_sum(g[k]) / m.scale_seq;
    }

    // sdim = 3 (material.py:780-805, called with sig_princ(sig)[0] by epl_dot / C_tan, 1045-1048, 1078-1080):
    // fgrad = Jac(sp) @ [1, dK/dx[1], 0] -- the first entry is the constant 1 of the reference, not dK/dx[0].  Jacobian
    // columns: d seq / d sp = 3 dev / vn with vn = sqrt(1.5) |dev|, and Re(-i ((a + i b) / (sp.a + i sp.b) - dseqds / vn)) =
    // (b (sp.a) - a (sp.b)) / |sp.a + i sp.b|^2; the all-ones matrix when vn <= 0.1.  Shear components of the flow
    // direction are zero.
    __device__ __forceinline__ void fgrad_princ(const double *s, double *a) const
    {
        const int np = m.nsv_pad;
        double sp[3], x[6], g1 = 0.;
        princ_ordered(s, sp);
        features_sp(sp, x);
        for (int i = lane; i < np; i += 32) {
            const double t0 = x[0] - s_sv[i], t1 = x[1] - s_sv[np + i];
            const double r = fma(t1, t1, t0 * t0);
            g1 = fma(s_dual[i] * exp(-m.gamma * r) * (-2. * m.gamma), t1, g1);
        }
        g1 = warp_sum(g1);
        const double hyd = (sp[0] + sp[1] + sp[2]) / 3.;
        const double dv[3] = {sp[0] - hyd, sp[1] - hyd, sp[2] - hyd};
        const double vn = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]) * sqrt(1.5);
        if (vn > 0.1) {
            const double r15 = sqrt(1.5), r2 = sqrt(2.);
            const double av[3] = {1. / r15, -0.5 / r15, -0.5 / r15}, bv[3] = {0., 0.5 * r2, -0.5 * r2};
            const double dsa = sp[0] * av[0] + sp[1] * av[1] + sp[2] * av[2];
            const double dsb = sp[0] * bv[0] + sp[1] * bv[1] + sp[2] * bv[2];
            const double n2 = dsa * dsa + dsb * dsb;
#pragma unroll
            for (int k = 0; k < 3; k++) a[k] = 3. * dv[k] / vn + ((bv[k] * dsa - av[k] * dsb) / n2) * g1;
        } else {
            a[0] = a[1] = a[2] = 1. + g1;
        }
        a[3] = a[4] = a[5] = 0.;
    }

    // find_yloc_scalar (material.py:547-574)
    __device__ __forceinline__ double yloc(double xs, const double *xu) const
    {
        double x[12];
#pragma unroll
        for (int k = 0; k < 6; k++) x[k] = xs * xu[k];
#pragma unroll
        for (int k = 0; k < 6; k++) x[6 + k] = (m.svc_nf == 12 && cur_epl) ? cur_epl[k] * m.inv_scale_wh : 0.;
        return decision_x(x);
    }

    // ML_full_yf (material.py:414-516).  ld == nullptr: loading direction = sig / seq.
    __device__ double ml_full_yf(const double *sig, double sflow, const double *ld) const
    {
        double seq = this->seq(sig);
        if (seq < 0.01 && ld == nullptr) return seq - 0.85 * sflow;
        double su[6];
        if (ld == nullptr) {
#pragma unroll
            for (int k = 0; k < 6; k++) su[k] = sig[k] / seq;
        } else {
            double hh = 0.;
#pragma unroll
            for (int k = 0; k < 6; k++) hh += ld[k] * ld[k];
            hh = sqrt(hh);
            if (hh < 1.e-3) {
#pragma unroll
                for (int k = 0; k < 6; k++) su[k] = 0.;
                su[0] = sqrt(1.5);
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) su[k] = ld[k] * sqrt(1.5) / hh;
            }
        }
        // features of the unit stress: x(xs * su) = xs * x(su) is NOT used (division by scale_seq after the
        // product, as the reference does): keep xs*su then scale.
        double x0 = sflow;
        if (su[0] * su[1] < -1.e-5) x0 *= (m.tresca ? 0.4 : 0.5);
        double x1 = x0;
        auto f = [&](double xs) {
            double s[6];
#pragma unroll
            for (int k = 0; k < 6; k++) s[k] = xs * su[k];
            return yf(s, sflow);
        };
        while (f(x0) >= 0. && x0 > 0.01) x0 *= 0.98;
        while (f(x1) < 0. && x1 < 5. * sflow) x1 *= 1.02;
        double f0 = f(x0), f1 = f(x1);
        if (f0 * f1 > 0.) return seq - 0.85 * sflow;
        // scipy.optimize.root_scalar(method='brentq', xtol=1e-5): rtol = 4 eps, maxiter = 100
        const double xtol = 1.e-5, rtol = 4. * 2.220446049250313e-16;
        double xpre = x0, xcur = x1, xblk = 0., fpre = f0, fcur = f1, fblk = 0., spre = 0., scur = 0.;
        double xs = 0.;
        bool conv = false;
        if (fpre == 0.) { xs = xpre; conv = true; }
        else if (fcur == 0.) { xs = xcur; conv = true; }
        else {
            for (int it = 0; it < 100; it++) {
                if (fpre != 0. && fcur != 0. && (signbit(fpre) != signbit(fcur))) {
                    xblk = xpre; fblk = fpre;
                    spre = scur = xcur - xpre;
                }
                if (fabs(fblk) < fabs(fcur)) {
                    xpre = xcur; xcur = xblk; xblk = xpre;
                    fpre = fcur; fcur = fblk; fblk = fpre;
                }
                double delta = (xtol + rtol * fabs(xcur)) / 2.;
                double sbis = (xblk - xcur) / 2.;
                if (fcur == 0. || fabs(sbis) < delta) { xs = xcur; conv = true; break; }
                if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
                    double stry;
                    if (xpre == xblk) {
                        stry = -fcur * (xcur - xpre) / (fcur - fpre);
                    } else {
                        double dpre = (fpre - fcur) / (xpre - xcur);
                        double dblk = (fblk - fcur) / (xblk - xcur);
                        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
                    }
                    if (2. * fabs(stry) < fmin(fabs(spre), 3. * fabs(sbis) - delta)) { spre = scur; scur = stry; }
                    else { spre = sbis; scur = sbis; }
                } else {
                    spre = sbis; scur = sbis;
                }
                xpre = xcur; fpre = fcur;
                if (fabs(scur) > delta) xcur += scur;
                else xcur += (sbis > 0. ? delta : -delta);
                fcur = f(xcur);
            }
            if (!conv) xs = xcur;
        }
        if (conv && xs < 4. * sflow) return seq - xs * this->seq(su);
        return seq - 0.85 * sflow;
    }
    __device__ __forceinline__ double full_yf(const double *s, double sflow) const { return ml_full_yf(s, sflow, nullptr); }
};

// ------------------------------------------------------------------------------------------------------------
// radial-return correction helpers (material.py:310-342)
// ------------------------------------------------------------------------------------------------------------
// minimum-norm least-squares solution of the 3x6 system of material.py:328-332 (np.linalg.lstsq, rcond=None):
// x = A^T (A A^T)^+ b, pseudo-inverse from the Jacobi eigen-decomposition of the 3x3 Gram matrix.
static __device__ __noinline__ void lstsq_3x6(const double *d /*deps_r*/, const double *b /*dsig[0:3]*/, double *x)
{
    const double A[3][6] = {{d[0], 0., 0., 0., d[2], d[1]}, {0., d[1], 0., d[2], 0., d[0]}, {0., 0., d[2], d[1], d[0], 0.}};
    double G[3][3], V[3][3] = {{1., 0., 0.}, {0., 1., 0.}, {0., 0., 1.}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.;
            for (int k = 0; k < 6; k++) s += A[i][k] * A[j][k];
            G[i][j] = s;
        }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = fabs(G[0][1]) + fabs(G[0][2]) + fabs(G[1][2]);
        double dia = fabs(G[0][0]) + fabs(G[1][1]) + fabs(G[2][2]);
        if (off <= 1e-300 || off <= 1e-20 * dia) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (G[p][q] == 0.) continue;
                double theta = (G[q][q] - G[p][p]) / (2. * G[p][q]);
                double t = copysign(1., theta) / (fabs(theta) + sqrt(theta * theta + 1.));
                double c = 1. / sqrt(t * t + 1.), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double gkp = G[k][p], gkq = G[k][q];
                    G[k][p] = c * gkp - s * gkq; G[k][q] = s * gkp + c * gkq;
                }
                for (int k = 0; k < 3; k++) {
                    double gpk = G[p][k], gqk = G[q][k];
                    G[p][k] = c * gpk - s * gqk; G[q][k] = s * gpk + c * gqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = fmax(G[0][0], fmax(G[1][1], G[2][2]));
    double smax = lmax > 0. ? sqrt(lmax) : 0.;
    double cut = 2.220446049250313e-16 * 6. * smax;
    double y[3] = {0., 0., 0.};
    for (int e = 0; e < 3; e++) {
        double lam = G[e][e];
        double sv = lam > 0. ? sqrt(lam) : 0.;
        if (sv <= cut || sv == 0.) continue;
        double vb = V[0][e] * b[0] + V[1][e] * b[1] + V[2][e] * b[2];
        for (int i = 0; i < 3; i++) y[i] += V[i][e] * vb / lam;
    }
    for (int k = 0; k < 6; k++) x[k] = A[0][k] * y[0] + A[1][k] * y[1] + A[2][k] * y[2];
}

// ------------------------------------------------------------------------------------------------------------
// Material.response (material.py:207-346) for one element.
//
// The tangent of a sub-step is Ct = CV - ca ca^T / hh (C_tan, 1083-1085), so the stress update is
// sig += CV deps_r - ca (ca . deps_r) / hh, and the returned average tangent is
//   grad_stiff = CV (1 - st_scal) [split only] + (nsteps CV - W) st_scal / nsteps,  W = sum_steps (ca ca^T / hh + Ct_corr)
// with W symmetric (21 registers instead of two 6x6 matrices).
//
// branch codes (for the sweep statistics): 0 elastic, 1 plastic single step, 2 plastic 50 sub-steps; bit 2 set when
// at least one radial-return correction fired.
// MODE 0: everything; MODE 1: return -1 instead of running the 50-sub-step loop (deferred to the compacted pass);
// ------------------------------------------------------------------------------------------------------------
template <int MODE, class Model>
__device__ __forceinline__ int response_core(const Model &ym, const DevMat &m, double *sig /*in/out*/, const double *epl,
                                             const double *deps, double *depl /*out*/, double *gs /*out, 21 sym*/,
                                             double &fy_out, int &nsteps_out)
{
    const double *CV = m.CV;
    const double sflow0 = m.sy + eps_eq6(epl) * ym.khard();   // get_sflow(epl) with the khard in effect
    const double toler = PLF_YF_TOLERANCE * (m.sy + eps_eq6(epl) * m.khard);   // 243 (uses Material.khard as set)
    double dsig[6], st[6];
#pragma unroll
    for (int k = 0; k < 6; k++) depl[k] = 0.;
    matvec6(CV, deps, dsig);   // 244
#pragma unroll
    for (int k = 0; k < 6; k++) st[k] = sig[k] + dsig[k];
    ym.set_epl(epl);   // yield function and gradient refer to the plastic strain at the start of the step unless stated otherwise
    double fy1 = ym.full_yf(st, sflow0);   // 249-252
    nsteps_out = 0;
    if (fy1 < toler) {   // 253-256 purely elastic
#pragma unroll
        for (int k = 0; k < 6; k++) sig[k] = st[k];
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) gs[sym_idx(i, j)] = CV[6 * i + j];
        fy_out = fy1;
        return 0;
    }
    double st_scal = 1.;
    double deps_r[6];
    bool split = false;
    double fy0 = ym.yf(sig, sflow0);   // 259
    if (fy0 < PLF_SPLIT_FY0) {         // 260-270
        split = true;
        // ML: distance to the yield surface, called without epl (263-265): sflow = sy
        if (Model::is_ml) {
            ym.set_epl(nullptr);   // 265: ML_full_yf(sig) without epl
            fy0 = ym.full_yf(sig, m.sy);
            ym.set_epl(epl);
        }
        st_scal += fy0 / ym.seq(dsig);
        double del[6], t[6];
#pragma unroll
        for (int k = 0; k < 6; k++) del[k] = deps[k] * (1. - st_scal);
        matvec6(CV, del, t);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            sig[k] += t[k];
            deps_r[k] = deps[k] - del[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) deps_r[k] = deps[k];
    }

    // one plastic (sub-)step; returns fy at its end.  W accumulates ca ca^T / hh (+ correction block).
    double W[21];
#pragma unroll
    for (int k = 0; k < 21; k++) W[k] = 0.;
    double ddepl[6];
    bool corrected = false;

    auto plastic_step = [&](double *s, const double *dr, bool commit) -> double {
        double cd[6], a[6], ca[6], eplt[6];
        matvec6(CV, dr, cd);
#pragma unroll
        for (int k = 0; k < 6; k++) st[k] = s[k] + cd[k];
        double yfun = ym.yf(st, sflow0);   // epl_dot 1031
        ym.fgrad(s, a);                    // calc_fgrad (shared by epl_dot 1050 and C_tan 1082)
        matvec6(CV, a, ca);
        double hh = ym.khard(), acd = 0., cadr = 0.;   // the hardening rate in effect after this gradient evaluation
#pragma unroll
        for (int k = 0; k < 6; k++) {
            hh += a[k] * ca[k];
            acd += a[k] * cd[k];
            cadr += ca[k] * dr[k];
        }
        double ihh = 1. / hh;
        double lam = (yfun <= PLF_YF_TOLERANCE) ? 0. : acd * ihh;   // 1041-1054
        double q = cadr * ihh;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            ddepl[k] = lam * a[k];
            eplt[k] = epl[k] + depl[k] + ddepl[k];
            st[k] = s[k] + (cd[k] - ca[k] * q);   // sig + t_stiff @ deps_r
        }
        double sfl = m.sy + eps_eq6(eplt) * ym.khard();
        ym.set_epl(eplt);   // 283, 304: the yield function at the end of the step sees the updated plastic strain
        double fy = ym.full_yf(st, sfl);
        ym.set_epl(epl);
        if (!commit) return fy;
#pragma unroll
        for (int k = 0; k < 6; k++) s[k] = st[k];
        if (fy > toler) {
            // 310-342: scale the stress back radially, balance with plastic strain, correct the tangent
            corrected = true;
            double sq = ym.seq(s);
            double ex[6], t[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                ex[k] = s[k] * fy / sq;
                s[k] -= ex[k];
            }
            matvec6(m.SV, ex, t);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                ddepl[k] += t[k];
                eplt[k] = epl[k] + depl[k] + ddepl[k];
            }
            double x[6];
            lstsq_3x6(dr, ex, x);
            W[sym_idx(0, 0)] += x[0]; W[sym_idx(0, 1)] += x[5]; W[sym_idx(0, 2)] += x[4];
            W[sym_idx(1, 1)] += x[1]; W[sym_idx(1, 2)] += x[3]; W[sym_idx(2, 2)] += x[2];
            sfl = m.sy + eps_eq6(eplt) * ym.khard();
            ym.set_epl(eplt);
            fy = ym.full_yf(s, sfl);
            ym.set_epl(epl);
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 6; j++) W[sym_idx(i, j)] += ca[i] * ca[j] * ihh;
#pragma unroll
        for (int k = 0; k < 6; k++) depl[k] += ddepl[k];
        return fy;
    };

    // trial step with the full remaining increment (277-285)
    fy1 = plastic_step(sig, deps_r, false);
    int nsteps = 1;
    if (fy1 > toler) {   // 288-293
        if (MODE == 1) return -1;
        nsteps = PLF_MAXIT;
#pragma unroll
        for (int k = 0; k < 6; k++) deps_r[k] /= (double)PLF_MAXIT;
    }
    for (int it = 0; it < nsteps; it++) fy1 = plastic_step(sig, deps_r, true);   // 295-344

    const double f1 = split ? (1. - st_scal) : 0.;
    const double f2 = st_scal / (double)nsteps;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
            double c = CV[6 * i + j];
            gs[sym_idx(i, j)] = c * f1 + ((double)nsteps * c - W[sym_idx(i, j)]) * f2;
        }
    fy_out = fy1;
    nsteps_out = nsteps - 1;   // msg['nsteps'] = niter (345)
    return (nsteps == 1 ? 1 : 2) | (corrected ? 4 : 0);
}

}  // namespace plf
