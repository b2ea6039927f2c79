// Production iteration: three-moment solve with EIGHT LANES PER BEAM and the optimiser state in registers.
//
// Replaces, per epoch (reference file:line): setup_model + analyze + eleResponse
// (OpenPyStruct_BeamOpt_training_SingleCore.py:176-190), the loss (:195-199), backward / Adam /
// ExponentialLR / clamp (:202-208) and the early-stop test (:211-219).
//
// Mapping.  Lane l of a beam's 8-lane group owns the elements e = 8 k + l, k = 0..EPL-1: lane l is
// SIMD lane l of the 8-float vectors ATen's cascade_sum works on, so torch.sum's summation order
// (beamopt_core.cuh, torch_sum_f32) falls out of per-lane running sums plus one fixed-order combine.
// Per element the lane keeps I, Adam's m and v, the gradient and the element's index inside its span
// in REGISTERS; the I-independent data of the three-moment form sit in shared memory, one
// conflict-free [slot][thread] column per lane.
//
// Three-moment form with I-independent coefficients (beamopt_flex.cuh derives the equations).
// With r_e = 1 / I_e, ke = index of element e inside its span, d = 1 / (elements in the span):
//     R0 = sum r_e        R1 = sum r_e ke        R2 = sum r_e ke^2
//     G  = sum r_e (g1 + g2)                     Q  = sum r_e (g1 x1 + g2 x2)
//     c = d^2 (6 R2 + 6 R1 + 2 R0)   S = d (2 R1 + R0)   b = 3 S - c   a = 6 R0 - 6 S + c   p = G - Q   q = Q
// (g1 = 2 M0 + m2 - w Le^2/4, g2 = 2 m2 + M0 - w Le^2/4 from the simply supported moment diagram M0 of
// the span, x1 = ke d, x2 = x1 + d; the common factor Le / (6 E) of the flexibilities cancels in the
// support-moment system and is only applied for the displacements).  Each lane accumulates the five
// sums over its own elements (PASS 1), the group reduces them through shared memory in a fixed
// order, every lane solves the <= 4-unknown tridiagonal system redundantly, and PASS 2 evaluates
//     Mc_e = M0_e + MS_l + (MS_r - MS_l) d ke ,   V_e = Q0_e + (MS_r - MS_l) d / Le
// followed by the fp32 loss terms, the frozen-M,V gradient and the Adam update of its elements.
//
// All cross-lane traffic goes through shared memory + __syncwarp(group mask), so the phase functions
// below contain no CUDA intrinsics and tests/hostsim runs the very same code lane by lane on the host.
#pragma once

#include "beamopt_flex.cuh"
#include "fastmath.cuh"

namespace ops {
namespace lanes {

constexpr int LPB = 8;                          // lanes per beam
constexpr int NSPAN = FLEX_MAXS - 1;            // spans between supports
constexpr int DUMMY = NSPAN;                    // span id of overhang elements and padding slots
constexpr int NSUM = 5;                         // R0, R1, R2, G, Q
constexpr int SCR_STAGE = NSPAN * NSUM;         // 3 slots after the partial sums: {row sum, tail} of sum I, d, q
constexpr int SCR_SLOTS = SCR_STAGE + 3;
constexpr int TAB_SLOTS = (NSPAN + 1) * 3;      // per span: MS_l, (MS_r - MS_l) d, (MS_r - MS_l) d / Le
constexpr int GX_DOUBLES = 2;                   // Moh, Qoh
constexpr int GX_INTS = 4;                      // m, last, nloads, setup status
constexpr int GROUP_DOUBLES = FlexStore::NUM_DOUBLES + TAB_SLOTS + GX_DOUBLES;
constexpr int GROUP_INTS = FlexStore::NUM_INTS + GX_INTS;

OPS_HD constexpr int lane_doubles(int epl) { return 4 * epl + SCR_SLOTS; }

template <int EPL>
struct LaneRegs {
    float I[EPL], m[EPL], v[EPL], g[EPL], ke[EPL];
    unsigned long long spans;                   // 3 bits per slot k: span id of element 8 k + l
};

// per-lane shared columns, entry k at base[k * ls]
struct LaneStore {
    double *gc, *qc, *m0, *q0;                  // [EPL]: G and Q coefficients, M0 and Q0 of the element
    double *scr;                                // [SCR_SLOTS]
    long ls;
};

// per-group shared columns, entry k at base[k * gs]
struct GroupStore {
    FlexStore fs;                               // supports, loads, RA, DXI; slots A..Q hold R0, R1, R2, G, Q
    double *tab;                                // [TAB_SLOTS]
    double *gd;                                 // [GX_DOUBLES]
    int *gi;                                    // [GX_INTS]
    long gs;
};

// torch.sum layout of an n-vector over slots k (beamopt_core.cuh): k < blk -> 4 ILP rows,
// blk <= k < vec -> row 0, k == vec -> scalar tail (lanes l < ntail)
struct SumShape {
    int vec, blk, ntail;
};
OPS_HD SumShape sum_shape(int n)
{
    SumShape s;
    s.vec = n / 8;
    s.blk = (s.vec / 4) * 4;
    s.ntail = n - 8 * s.vec;
    return s;
}

// ---------------------------------------------------------------------------------------------
// once per beam
// ---------------------------------------------------------------------------------------------
OPS_HD void group_publish(const FlexBeam &fb, int rc, const GroupStore &gs)
{
    gs.gi[0] = fb.m; gs.gi[gs.gs] = fb.last; gs.gi[2 * gs.gs] = fb.nloads; gs.gi[3 * gs.gs] = rc;
    gs.gd[0] = fb.Moh; gs.gd[gs.gs] = fb.Qoh;
}

OPS_HD int group_fetch(const BeamConsts &k, double L, const GroupStore &gs, FlexBeam &fb)
{
    flex_geometry(k, L, fb);
    fb.m = gs.gi[0]; fb.last = gs.gi[gs.gs]; fb.nloads = gs.gi[2 * gs.gs];
    fb.Moh = gs.gd[0]; fb.Qoh = gs.gd[gs.gs];
    return gs.gi[3 * gs.gs];
}

template <int EPL>
OPS_HD void lane_init(const BeamConsts &k, int n, const FlexBeam &fb, const GroupStore &gs, const LaneStore &ls,
                      int l, LaneRegs<EPL> &rg)
{
    const int m = fb.m, last = fb.last, nl = fb.nloads;
    for (int s = 0; s < SCR_SLOTS; ++s) ls.scr[(long)s * ls.ls] = 0.0;
    unsigned long long spans = 0;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        int sp = DUMMY;
        double kef = 0.0, G = 0.0, Qc = 0.0, M0 = 0.0, Q0 = 0.0;
        float I0 = 1.0f;                        // padding slots: harmless inertia, never read back
        if (e < n) {
            I0 = k.I0f;
            if (e < last) {
                int j = 0;
                for (int s = 1; s < m; ++s) j = (gs.fs.sup(s) <= e) ? s : j;
                const int na = gs.fs.sup(j);
                const double ke = (double)(e - na);
                const double d = gs.fs.span(j + 1, FlexStore::DXI);
                const double ra = gs.fs.span(j + 1, FlexStore::RA);
                double Qs = fma(ke, fb.wl, ra);
                double Ms = fma(ra, ke * fb.Le, fb.wl2h * (ke * ke));
                for (int q = 0; q < nl; ++q) {
                    const int nd = gs.fs.lnode(q);
                    if (nd > na && nd <= e) {
                        const double P = gs.fs.lval(q);
                        Qs += P;
                        Ms = fma(P, (double)(e - nd) * fb.Le, Ms);
                    }
                }
                const double m2 = fma(Qs, fb.Le, Ms + fb.wl2h);
                const double g1 = fma(2.0, Ms, m2) - fb.corr, g2 = fma(2.0, m2, Ms) - fb.corr;
                const double x1 = ke * d, x2 = x1 + d;
                G = g1 + g2;
                Qc = fma(g1, x1, g2 * x2);
                sp = j; kef = ke; M0 = Ms; Q0 = Qs;
            } else {
                const double r = (double)(n - e);
                double Ms = fb.wl2h * (r * r), Qs = fb.wl * r;
                for (int q = 0; q < nl; ++q) {
                    const int nd = gs.fs.lnode(q);
                    if (nd > e) {
                        const double P = gs.fs.lval(q);
                        Ms = fma(P, (double)(nd - e) * fb.Le, Ms);
                        Qs += P;
                    }
                }
                M0 = Ms; Q0 = -Qs;
            }
        }
        spans |= (unsigned long long)sp << (3 * kk);
        rg.I[kk] = I0; rg.m[kk] = 0.0f; rg.v[kk] = 0.0f; rg.g[kk] = 0.0f; rg.ke[kk] = (float)kef;
        ls.gc[(long)kk * ls.ls] = G; ls.qc[(long)kk * ls.ls] = Qc;
        ls.m0[(long)kk * ls.ls] = M0; ls.q0[(long)kk * ls.ls] = Q0;
    }
    rg.spans = spans;
}

// a beam rejected at set-up (mechanism / unsupported support count) still emits I_0 in its record
template <int EPL>
OPS_HD void lane_reset(const BeamConsts &k, LaneRegs<EPL> &rg)
{
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        rg.I[kk] = k.I0f; rg.m[kk] = 0.0f; rg.v[kk] = 0.0f; rg.g[kk] = 0.0f; rg.ke[kk] = 0.0f;
    }
    rg.spans = 0;
}

// ---------------------------------------------------------------------------------------------
// per epoch
// ---------------------------------------------------------------------------------------------
OPS_HD void flush_span(const LaneStore &ls, int j, double R0, double R1, double R2, double G, double Q)
{
    double *s = ls.scr + (long)(j * NSUM) * ls.ls;
    s[0] = R0; s[ls.ls] = R1; s[2 * ls.ls] = R2; s[3 * ls.ls] = G; s[4 * ls.ls] = Q;
}

// PASS 1: flexibility sums of the lane's elements, one partial per span in the lane's scratch column
template <int EPL>
OPS_HD void lane_pass1(const LaneRegs<EPL> &rg, const LaneStore &ls)
{
    int jc = (int)(rg.spans & 7u);
    double R0 = 0.0, R1 = 0.0, R2 = 0.0, G = 0.0, Q = 0.0;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int j = (int)((rg.spans >> (3 * kk)) & 7u);
        if (j != jc) {
            if (jc != DUMMY) flush_span(ls, jc, R0, R1, R2, G, Q);
            R0 = R1 = R2 = G = Q = 0.0;
            jc = j;
        }
        const double r = fm::rcp64((double)rg.I[kk]);
        const double ke = (double)rg.ke[kk];
        const double t = r * ke;
        R0 += r;
        R1 += t;
        R2 = fma(t, ke, R2);
        G = fma(r, ls.gc[(long)kk * ls.ls], G);
        Q = fma(r, ls.qc[(long)kk * ls.ls], Q);
    }
    if (jc != DUMMY) flush_span(ls, jc, R0, R1, R2, G, Q);
}

// group reduction of the partials: lane l owns the sums p = l, l + 8, ... and adds the eight lanes'
// partials in the fixed order l, l+1, ... (mod 8) -- deterministic and bank-conflict free
OPS_HD void lane_reduce(int l, const LaneStore &ls, const GroupStore &gs)
{
    const double *scr0 = ls.scr - l;
#pragma unroll
    for (int i = 0; i < (NSPAN * NSUM + LPB - 1) / LPB; ++i) {
        const int p = l + LPB * i;
        if (p < NSPAN * NSUM) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < LPB; ++r) s += scr0[(long)p * ls.ls + ((l + r) & (LPB - 1))];
            gs.fs.span(p / NSUM + 1, FlexStore::A + p % NSUM) = s;
        }
    }
}

// flexibility coefficients of span j (0-based) from its reduced sums, without the Le/(6E) factor
OPS_HD void span_flex(const GroupStore &gs, int j, double &a, double &b, double &c, double &p, double &q, double &d)
{
    const double R0 = gs.fs.span(j + 1, FlexStore::A), R1 = gs.fs.span(j + 1, FlexStore::B);
    const double R2 = gs.fs.span(j + 1, FlexStore::C), G = gs.fs.span(j + 1, FlexStore::P);
    q = gs.fs.span(j + 1, FlexStore::Q);
    d = gs.fs.span(j + 1, FlexStore::DXI);
    c = (d * d) * fma(6.0, R2, fma(6.0, R1, 2.0 * R0));
    const double S = d * fma(2.0, R1, R0);
    b = fma(3.0, S, -c);
    a = fma(-6.0, S, fma(6.0, R0, c));
    p = G - q;
}

// three-moment system for the support moments (every lane, redundantly); lane 0 publishes the
// per-span table PASS 2 reads.  Returns 1 when a pivot is not positive.
OPS_HD int group_solve(const FlexBeam &fb, const GroupStore &gs, int l)
{
    const int m = fb.m;
    double a[NSPAN], b[NSPAN], c[NSPAN], p[NSPAN], q[NSPAN], dx[NSPAN];
#pragma unroll
    for (int j = 0; j < NSPAN; ++j) {
        a[j] = b[j] = c[j] = p[j] = q[j] = 0.0;
        dx[j] = 0.0;
        if (j < m) span_flex(gs, j, a[j], b[j], c[j], p[j], q[j], dx[j]);
    }
    double MS[NSPAN + 1];
    MS[0] = 0.0;
#pragma unroll
    for (int j = 1; j <= NSPAN; ++j) MS[j] = (j == m) ? fb.Moh : 0.0;
    int bad = 0;
    double inv[NSPAN], rr[NSPAN];
    double ip = 0.0, rp = 0.0;
#pragma unroll
    for (int kk = 1; kk < NSPAN; ++kk) {
        inv[kk] = 0.0; rr[kk] = 0.0;
        if (kk < m) {
            double dd = c[kk - 1] + a[kk];
            double r_ = -(q[kk - 1] + p[kk]);
            if (kk == m - 1) r_ = fma(-b[kk], fb.Moh, r_);
            if (kk > 1) {
                const double bk = b[kk - 1];
                const double w = bk * ip;
                dd = fma(-w, bk, dd);
                r_ = fma(-w, rp, r_);
            }
            if (!(dd > 0.0)) bad = 1;
            ip = fm::rcp64(dd);
            rp = r_;
            inv[kk] = ip; rr[kk] = r_;
        }
    }
    double mnext = fb.Moh;
#pragma unroll
    for (int kk = NSPAN - 1; kk >= 1; --kk) {
        if (kk < m) {
            double r_ = rr[kk];
            if (kk < m - 1) r_ = fma(-b[kk], mnext, r_);
            mnext = r_ * inv[kk];
            MS[kk] = mnext;
        }
    }
    if (l == 0) {
#pragma unroll
        for (int j = 0; j <= NSPAN; ++j) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            if (j < NSPAN && j < m) {
                const double dM = MS[j + 1] - MS[j];
                t0 = MS[j];
                t1 = dM * dx[j];
                t2 = t1 * fb.invLe;
            }
            gs.tab[(long)(3 * j) * gs.gs] = t0;
            gs.tab[(long)(3 * j + 1) * gs.gs] = t1;
            gs.tab[(long)(3 * j + 2) * gs.gs] = t2;
        }
    }
    return bad;
}

// bending moment (three-moment sign: sagging positive) and shear at the node-i end of slot kk
template <int EPL>
OPS_HD void element_forces(const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, int kk,
                           double &Mc, double &Qv)
{
    const int j = (int)((rg.spans >> (3 * kk)) & 7u);
    const double *t = gs.tab + (long)(3 * j) * gs.gs;
    const double T0 = t[0], T1 = t[gs.gs], T2 = t[2 * gs.gs];
    Mc = fma(T1, (double)rg.ke[kk], ls.m0[(long)kk * ls.ls] + T0);
    Qv = ls.q0[(long)kk * ls.ls] + T2;
}

// loss terms d, q and autograd's gradient with M, V constant (element_update_f32, first half), with
// the branch-free division / square-root sequences.  Ranges: I in [clamp_min, 1e20) (the clamp,
// SingleCore:208), c = M^2 and h = V^2 zero or >= 2^-100.
OPS_HD void element_grad(const BeamConsts &k, float I, float c, float h, float &d, float &q, float &g)
{
    const float b = k.E2 * I + k.epsf;
    const float rb = fm::rcp_r(b);
    d = fm::div_r(c, b, rb);
    const float db = fm::div_r(d, b, rb);
    const float s = fm::sqrt_f(I);
    const float gg = k.Gf * (k.kf * s);
    const float rgg = fm::rcp_r(gg);
    q = fm::div_r(h, gg, rgg);
    const float qg = fm::div_r(q, gg, rgg);
    const float is = fm::div_f(1.0f, s);
    const float gb = ((-k.am) * db) * k.E2;
    const float gs_ = ((((-k.as_) * qg) * k.Gf) * k.kf) * (0.5f * is);
    g = (1.0f + gs_) + gb;
}

// PASS 2: end forces, loss terms, gradient (kept in rg.g), torch.sum partials of sum I, sum d, sum q
template <int EPL>
OPS_HD void lane_forces(const BeamConsts &k, int n, LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, int l)
{
    const SumShape sh = sum_shape(n);
    float aI[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ad[4] = {0.0f, 0.0f, 0.0f, 0.0f}, aq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float tI = 0.0f, td = 0.0f, tq = 0.0f;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        double Mc, Qv;
        element_forces<EPL>(rg, ls, gs, kk, Mc, Qv);
        const float Mf = (float)Mc, Vf = (float)Qv;
        float d, q;
        element_grad(k, rg.I[kk], Mf * Mf, Vf * Vf, d, q, rg.g[kk]);
        if (kk < sh.blk) {
            aI[kk & 3] += rg.I[kk]; ad[kk & 3] += d; aq[kk & 3] += q;
        } else if (kk < sh.vec) {
            aI[0] += rg.I[kk]; ad[0] += d; aq[0] += q;
        } else if (kk == sh.vec && l < sh.ntail) {
            tI = rg.I[kk]; td = d; tq = q;
        }
    }
    float *st = reinterpret_cast<float *>(ls.scr + (long)SCR_STAGE * ls.ls);
    const long fs_ = 2 * ls.ls;                 // float stride between slots
    st[0] = ((aI[0] + aI[1]) + aI[2]) + aI[3]; st[1] = tI;
    st[fs_] = ((ad[0] + ad[1]) + ad[2]) + ad[3]; st[fs_ + 1] = td;
    st[2 * fs_] = ((aq[0] + aq[1]) + aq[2]) + aq[3]; st[2 * fs_ + 1] = tq;
}

// total loss in torch's order: scalar tail first, then the eight vector lanes (every lane, redundantly)
OPS_HD float group_loss(const BeamConsts &k, int n, const LaneStore &ls, int l)
{
    const SumShape sh = sum_shape(n);
    const float *st0 = reinterpret_cast<const float *>(ls.scr - l + (long)SCR_STAGE * ls.ls);
    const long fs_ = 2 * ls.ls;
    float s[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float acc = 0.0f;
        for (int t = 0; t < sh.ntail; ++t) acc += st0[i * fs_ + 2 * t + 1];
#pragma unroll
        for (int r = 0; r < LPB; ++r) acc += st0[i * fs_ + 2 * r];
        s[i] = acc;
    }
    return (s[0] + k.am * s[1]) + k.as_ * s[2];
}

// Adam step + clamp on the lane's elements (element_update_f32, second half).  The fast square root
// needs v >= 2^-101; v is an EMA of g^2, so anything smaller means g vanished on every epoch so far --
// tested once per lane and epoch, with the generic operators as the (cold) alternative.
template <int EPL>
OPS_HD void lane_adam(const BeamConsts &k, LaneRegs<EPL> &rg, float neg_step, float bc2_sqrt)
{
    bool rare = false;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const float g = rg.g[kk];
        rg.m[kk] = fmaf(k.w1, g - rg.m[kk], rg.m[kk]);
        rg.v[kk] = fmaf(k.omb2f * g, g, rg.v[kk] * k.b2f);
        rare = rare || !(rg.v[kk] >= fm::SQRT_F_MIN);
    }
    if (!rare) {
        const float rbc = fm::rcp_r(bc2_sqrt);
#pragma unroll
        for (int kk = 0; kk < EPL; ++kk) {
            const float denom = fm::div_r(fm::sqrt_f(rg.v[kk]), bc2_sqrt, rbc) + k.adam_epsf;
            const float x = rg.I[kk] + fm::div_f(neg_step * rg.m[kk], denom);
            rg.I[kk] = x < k.clampf ? k.clampf : x;
        }
    } else {
#pragma unroll 1
        for (int kk = 0; kk < EPL; ++kk) {
            const float denom = sqrtf(rg.v[kk]) / bc2_sqrt + k.adam_epsf;
            const float x = rg.I[kk] + (neg_step * rg.m[kk]) / denom;
            rg.I[kk] = x < k.clampf ? k.clampf : x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// once per beam: the record (SingleCore:221-249).  M, V, u, theta belong to the LAST ANALYSED
// inertias, i.e. the ones still in rg.I when the stop decision is taken (before lane_adam).
// ---------------------------------------------------------------------------------------------
template <int EPL>
OPS_HD void lane_emit_forces(int n, const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, int l,
                             bool fields, float *shear, float *moment)
{
    float *stage = reinterpret_cast<float *>(ls.scr);        // I of slot kk at float index (kk >> 1) * 2 ls + (kk & 1)
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        stage[(long)(kk >> 1) * 2 * ls.ls + (kk & 1)] = rg.I[kk];
        if (e < n) {
            double Mc = 0.0, Qv = 0.0;
            if (fields) element_forces<EPL>(rg, ls, gs, kk, Mc, Qv);
            shear[e] = fields ? (float)Qv : 0.0f;
            moment[e] = fields ? (float)(-Mc) : 0.0f;
        }
    }
}

// lane 0: displacements by integrating the curvature (flex_deflections_march) from the staged inertias
OPS_HD void group_emit_displacements(const BeamConsts &k, const FlexBeam &fb, const LaneStore &ls0,
                                     const GroupStore &gs, bool fields, double *defl, double *rot)
{
    const int nn = k.nn;
    if (!fields) {
        for (int i = 0; i < nn; ++i) { defl[i] = 0.0; rot[i] = 0.0; }
        return;
    }
    const int m = fb.m;
    for (int j = 0; j < m; ++j) {
        double a, b, c, p, q, d;
        span_flex(gs, j, a, b, c, p, q, d);
        gs.fs.span(j + 1, FlexStore::A) = fb.kc6 * a;
        gs.fs.span(j + 1, FlexStore::B) = fb.kc6 * b;
        gs.fs.span(j + 1, FlexStore::P) = fb.kc6 * p;
        gs.fs.ms(j) = gs.tab[(long)(3 * j) * gs.gs];
    }
    gs.fs.ms(m) = fb.Moh;
    const float *stage = reinterpret_cast<const float *>(ls0.scr);
    auto inertia = [&](int e) {
        const int kk = e >> 3, ln = e & (LPB - 1);
        return (double)stage[(long)(kk >> 1) * 2 * ls0.ls + 2 * ln + (kk & 1)];
    };
    flex_deflections_march(k, fb, gs.fs, inertia, [&](int i, double u, double th) {
        const bool z = k.zero_last_node && i == nn - 1;
        defl[i] = z ? 0.0 : u;
        rot[i] = z ? 0.0 : th;
    });
}

template <int EPL>
OPS_HD void lane_emit_inertias(int n, const LaneRegs<EPL> &rg, int l, float *I_out)
{
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        if (e < n) I_out[e] = rg.I[kk];
    }
}

}  // namespace lanes
}  // namespace ops
