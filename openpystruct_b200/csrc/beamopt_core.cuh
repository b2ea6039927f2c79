// Per-beam arithmetic of the fused optimisation iteration: stiffness assembly, block LDL^T of the
// banded Euler-Bernoulli system, solve, element end forces (FP64), then loss / frozen-M,V gradient /
// Adam / clamp in FP32 with torch's CPU operation order.
//
// Replaces, per epoch (reference file:line):
//   setup_model + analyze         OpenPyStruct_BeamOpt_training_SingleCore.py:89-124, 176-182
//   eleResponse(e,'forces')[1|2]  SingleCore:189-190
//   loss                          SingleCore:195-199
//   backward / Adam / LR / clamp  SingleCore:202-208
//
// The stiffness K(I) is block tridiagonal with 2x2 node blocks (half bandwidth 3 over the
// (uy, theta) DOFs).  One thread owns one beam; its factor lives in on-chip storage laid out
// [slot][thread] (bank-conflict free, see BeamStore).  Constrained uy DOFs keep their row/column as
// an identity row (arithmetic on the free DOFs is that of OpenSees' `Plain` handler).
//
// Every floating-point operation below is exactly one IEEE rounding: the translation unit is built
// with -fmad=false and all fused operations are written as fma()/fmaf().  This header has no CUDA
// intrinsics so that tests/hostsim can compile the very same arithmetic with g++ (debug aid only).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OPS_HD __host__ __device__ __forceinline__
#else
#define OPS_HD inline
#endif

namespace ops {

// Launch-wide constants, derived on the host from OpsBeamOptParams (fp32 ones are the values torch
// obtains when a Python double meets an fp32 tensor).
struct BeamConsts {
    int nn, n, max_forces, max_epochs, patience, early_stop, zero_last_node;
    double E, udl, tol;
    float I0f, E2, Gf, kf, am, as_, epsf, clampf, w1, b2f, omb2f, adam_epsf;
};

// Per-beam storage: `d` holds 5 doubles per node, `f` holds I, m, v (3 x n floats).  Element k of
// a logical array sits at base[k * stride]; the base pointers are already offset by the thread.
//   node i, slots 5i..5i+2 : pivot block S_i = L D L^T as (l, 1/d0, 1/d1) during factor/solve; afterwards
//                            slot 5i   = float2 {M_i, V_i}   (fp32 element end forces)
//                            slot 5i+1 = float2 {d_i, q_i}   (energy densities of the loss)
//   node i, slots 5i+3,5i+4: g_i (condensed load) -> overwritten by the solution (uy_i, theta_i)
struct BeamStore {
    double *d;
    float *f;
    long stride;
    OPS_HD double &D(int slot) const { return d[(long)slot * stride]; }
    OPS_HD float &F(int slot) const { return f[(long)slot * stride]; }
    OPS_HD float *pairf(int slot) const { return reinterpret_cast<float *>(&d[(long)slot * stride]); }
};

// Per-beam scalars kept in registers.
template <int MAXF>
struct BeamInputs {
    double Le, invLe;
    double ka, kb, kc2, kc4;     // element stiffness entries per unit I: 12E/Le^3, 6E/Le^2, 2E/Le, 4E/Le
    double wl, mfe;              // w*Le, w*Le^2/12
    int fnode[MAXF];
    double fval[MAXF];
};

template <int MAXF>
OPS_HD void beam_geometry(const BeamConsts &k, double L, BeamInputs<MAXF> &in)
{
    const double Le = L / (double)k.n;
    const double s3 = k.E / (Le * Le * Le);
    in.Le = Le;
    in.invLe = 1.0 / Le;
    in.ka = 12.0 * s3;
    in.kb = 6.0 * Le * s3;
    in.kc2 = 2.0 * Le * Le * s3;
    in.kc4 = 4.0 * Le * Le * s3;
    in.wl = k.udl * Le;
    in.mfe = k.udl * Le * Le / 12.0;
}

template <int MAXF>
OPS_HD double nodal_point_load(const BeamInputs<MAXF> &in, int i)
{
    double f = 0.0;
#pragma unroll
    for (int j = 0; j < MAXF; ++j) f += (in.fnode[j] == i) ? in.fval[j] : 0.0;
    return f;
}

// ---------------------------------------------------------------------------------------------
// FP64: factor + solve + end forces.  `fixed(i)` tells whether uy_i is constrained.
// I(e) returns the element inertia as double.  Returns 0, or 1 when a pivot block is not SPD.
// ---------------------------------------------------------------------------------------------
template <int MAXF, class FixedFn, class InertiaFn>
OPS_HD int factor_forward(const BeamConsts &k, const BeamInputs<MAXF> &in, const BeamStore &st,
                          FixedFn fixed, InertiaFn I)
{
    const int n = k.n;
    int bad = 0;
    // carried from node i-1: pivot block factors (S = L D L^T: l, 1/d0, 1/d1), condensed load,
    // element i-1 entries, mask
    double l = 0.0, r0 = 0.0, r1 = 0.0, g0p = 0.0, g1p = 0.0;
    double ap = 0.0, bp = 0.0, c2p = 0.0, c4p = 0.0;
    bool fxp = true;
    for (int i = 0; i <= n; ++i) {
        const bool fx = (i == 0) || fixed(i);
        double a = 0.0, b = 0.0, c2 = 0.0, c4 = 0.0;     // element i (to the right of node i)
        if (i < n) {
            const double Ie = I(i);
            a = in.ka * Ie; b = in.kb * Ie; c2 = in.kc2 * Ie; c4 = in.kc4 * Ie;
        }
        // diagonal block D_i = A^R_{i-1} + A^L_i and load f_i
        double s00 = fx ? 1.0 : (ap + a);
        double s01 = fx ? 0.0 : (b - bp);
        double s11 = c4p + c4;
        double fu = nodal_point_load(in, i) + ((i == 0 || i == n) ? 0.5 * in.wl : in.wl);
        fu = fx ? 0.0 : fu;
        double fth = (i == 0) ? in.mfe : ((i == n) ? -in.mfe : 0.0);
        if (i > 0) {
            // coupling C_{i-1} (rows node i-1, cols node i) with constrained rows/cols zeroed
            const double c00 = (fxp || fx) ? 0.0 : -ap;
            const double c01 = fxp ? 0.0 : bp;
            const double c10 = fx ? 0.0 : -bp;
            const double c11 = c2p;
            // Y = L^-1 C (row 1 -= l * row 0), Z = D^-1 Y ;  S_i = D_i - Y^T Z ;  g_i = f_i - Z^T (L^-1 g)
            const double y10 = fma(-l, c00, c10);
            const double y11 = fma(-l, c01, c11);
            const double z00 = c00 * r0, z01 = c01 * r0;
            const double z10 = y10 * r1, z11 = y11 * r1;
            s00 -= fma(c00, z00, y10 * z10);
            s01 -= fma(c00, z01, y10 * z11);
            s11 -= fma(c01, z01, y11 * z11);
            const double h1 = fma(-l, g0p, g1p);
            fu -= fma(z00, g0p, z10 * h1);
            fth -= fma(z01, g0p, z11 * h1);
        }
        // scalar LDL^T of the 2x2 pivot block (what a half-bandwidth-3 band LDL^T does on these DOFs)
        r0 = 1.0 / s00;
        l = s01 * r0;
        const double d1 = fma(-l, s01, s11);
        r1 = 1.0 / d1;
        if (!(s00 > 0.0) || !(d1 > 0.0)) bad = 1;
        g0p = fu; g1p = fth;
        st.D(5 * i + 0) = l; st.D(5 * i + 1) = r0; st.D(5 * i + 2) = r1;
        st.D(5 * i + 3) = fu;  st.D(5 * i + 4) = fth;
        ap = a; bp = b; c2p = c2; c4p = c4; fxp = fx;
    }
    return bad;
}

// Back substitution right-to-left fused with ElasticBeam2d::getResistingForce.  emit(e, V, M)
// receives the FP64 end forces of element e; the solution overwrites the g slots.
template <int MAXF, class FixedFn, class InertiaFn, class EmitFn>
OPS_HD int solve_backward(const BeamConsts &k, const BeamInputs<MAXF> &in, const BeamStore &st,
                          FixedFn fixed, InertiaFn I, EmitFn emit)
{
    const int n = k.n;
    const double vfe = 0.5 * in.wl;        // w*Le/2
    double u1, t1;                          // solution at node i+1
    {
        const double l = st.D(5 * n + 0), r0 = st.D(5 * n + 1), r1 = st.D(5 * n + 2);
        const double g0 = st.D(5 * n + 3), g1 = st.D(5 * n + 4);
        t1 = fma(-l, g0, g1) * r1;
        u1 = fma(-l, t1, g0 * r0);
        st.D(5 * n + 3) = u1; st.D(5 * n + 4) = t1;
    }
    bool fx1 = fixed(n);
    for (int i = n - 1; i >= 0; --i) {
        const bool fx = (i == 0) || fixed(i);
        const double Ie = I(i);
        const double a = in.ka * Ie, b = in.kb * Ie, c2 = in.kc2 * Ie, c4 = in.kc4 * Ie;
        const double c00 = (fx || fx1) ? 0.0 : -a;
        const double c01 = fx ? 0.0 : b;
        const double c10 = fx1 ? 0.0 : -b;
        const double c11 = c2;
        const double r0 = st.D(5 * i + 3) - fma(c00, u1, c01 * t1);
        const double r1 = st.D(5 * i + 4) - fma(c10, u1, c11 * t1);
        const double l = st.D(5 * i + 0), p0 = st.D(5 * i + 1), p1 = st.D(5 * i + 2);
        const double t0 = fma(-l, r0, r1) * p1;
        const double u0 = fma(-l, t0, r0 * p0);
        st.D(5 * i + 3) = u0; st.D(5 * i + 4) = t0;
        // basic deformations and end forces (LinearCrdTransf2d / ElasticBeam2d)
        const double chord = (u0 - u1) * in.invLe;
        const double v1 = t0 + chord, v2 = t1 + chord;
        const double q1 = fma(c4, v1, fma(c2, v2, -in.mfe));
        const double q2 = fma(c2, v1, fma(c4, v2, in.mfe));
        const double V = fma(q1 + q2, in.invLe, -vfe);
        emit(i, V, q1);
        u1 = u0; t1 = t0; fx1 = fx;
    }
    return (u1 == u1 && t1 == t1) ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// FP32: torch.sum (ATen cpu/SumKernel.cpp cascade_sum: 8-lane vectors x 4 ILP rows, 4 cascade
// levels of 2^max(4, ceil_log2(steps)/4) steps).  get(e) returns element e.
// ---------------------------------------------------------------------------------------------
template <class Get>
OPS_HD float torch_sum_f32(int n, Get get)
{
    const int vec_size = n / 8;
    const int size_ilp = vec_size / 4;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    if (size_ilp < 16) {
        for (int i = 0; i < size_ilp; ++i) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += get(32 * i + j);
        }
    } else {
        int lg = 0;
        while ((1 << lg) < size_ilp) ++lg;
        int level_power = lg / 4;
        if (level_power < 4) level_power = 4;
        const int level_step = 1 << level_power;
        const int level_mask = level_step - 1;
        float lv[3][32];                         // levels 1..3 (level 0 is acc)
        for (int l = 0; l < 3; ++l)
            for (int j = 0; j < 32; ++j) lv[l][j] = 0.0f;
        int i = 0;
        for (; i + level_step <= size_ilp;) {
            for (int s = 0; s < level_step; ++s, ++i)
                for (int j = 0; j < 32; ++j) acc[j] += get(32 * i + j);
            for (int l = 1; l < 4; ++l) {
                for (int j = 0; j < 32; ++j) {
                    const float below = (l == 1) ? acc[j] : lv[l - 2][j];
                    lv[l - 1][j] += below;
                    if (l == 1) acc[j] = 0.0f; else lv[l - 2][j] = 0.0f;
                }
                const int mask = level_mask << (l * level_power);
                if ((i & mask) != 0) break;
            }
        }
        for (; i < size_ilp; ++i)
            for (int j = 0; j < 32; ++j) acc[j] += get(32 * i + j);
        for (int l = 0; l < 3; ++l)
            for (int j = 0; j < 32; ++j) acc[j] += lv[l][j];
    }
    for (int v = size_ilp * 4; v < vec_size; ++v) {
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[l] += get(8 * v + l);
    }
    float s = 0.0f;
    for (int e = vec_size * 8; e < n; ++e) s += get(e);
#pragma unroll
    for (int l = 0; l < 8; ++l) s += ((acc[l] + acc[8 + l]) + acc[16 + l]) + acc[24 + l];
    return s;
}

// Loss terms, autograd's gradient with M,V constant, and the single-tensor Adam update + clamp for
// one element.  c = sum over load cases of M^2, h = same for V.  Returns d, q through pointers.
OPS_HD void element_update_f32(const BeamConsts &k, float neg_step, float bc2_sqrt, float c, float h,
                               float &I, float &m, float &v, float &d_out, float &q_out)
{
    const float b = k.E2 * I + k.epsf;
    const float d = c / b;
    const float s = sqrtf(I);
    const float gg = k.Gf * (k.kf * s);
    const float q = h / gg;
    const float gb = ((-k.am) * (d / b)) * k.E2;
    const float gs = ((((-k.as_) * (q / gg)) * k.Gf) * k.kf) * (0.5f * (1.0f / s));
    const float g = (1.0f + gs) + gb;
    m = fmaf(k.w1, g - m, m);
    v = fmaf(k.omb2f * g, g, v * k.b2f);
    const float denom = sqrtf(v) / bc2_sqrt + k.adam_epsf;
    const float x = I + (neg_step * m) / denom;
    I = x < k.clampf ? k.clampf : x;
    d_out = d;
    q_out = q;
}

// One full iteration for a single-load-case beam.  Returns the fp32 total loss; *bad is set when
// the factorisation failed.  I, m, v live in st.F(e), st.F(n+e), st.F(2n+e).
template <int MAXF, class FixedFn>
OPS_HD float beam_iteration(const BeamConsts &k, const BeamInputs<MAXF> &in, const BeamStore &st,
                            FixedFn fixed, float neg_step, float bc2_sqrt, int *bad)
{
    const int n = k.n;
    auto inertia = [&](int e) { return (double)st.F(e); };
    int rc = factor_forward<MAXF>(k, in, st, fixed, inertia);
    rc |= solve_backward<MAXF>(k, in, st, fixed, inertia, [&](int e, double V, double M) {
        float *mv = st.pairf(5 * e);
        mv[0] = (float)M;
        mv[1] = (float)V;
    });
    *bad = rc;
    const float sI = torch_sum_f32(n, [&](int e) { return st.F(e); });
    for (int e = 0; e < n; ++e) {
        const float *mv = st.pairf(5 * e);
        const float M = mv[0], V = mv[1];
        float I = st.F(e), m = st.F(n + e), v = st.F(2 * n + e), d, q;
        element_update_f32(k, neg_step, bc2_sqrt, M * M, V * V, I, m, v, d, q);
        st.F(e) = I; st.F(n + e) = m; st.F(2 * n + e) = v;
        float *dq = st.pairf(5 * e + 1);
        dq[0] = d;
        dq[1] = q;
    }
    const float sd = torch_sum_f32(n, [&](int e) { return st.pairf(5 * e + 1)[0]; });
    const float sq = torch_sum_f32(n, [&](int e) { return st.pairf(5 * e + 1)[1]; });
    return (sI + k.am * sd) + k.as_ * sq;
}

}  // namespace ops
