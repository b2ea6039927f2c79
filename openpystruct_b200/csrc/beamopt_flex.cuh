// Three-moment (flexibility) form of the per-iteration beam solve -- the production iteration.
//
// What the reference does per epoch (OpenPyStruct_BeamOpt_training_SingleCore.py:176-190): build
// K(I), BandSPD-solve K u = f, read the element end forces.  For an Euler-Bernoulli beam with
// piecewise-constant EI, nodal loads and a uniform load, the cubic-Hermite stiffness model is nodally
// exact, so the same end forces follow from the exact Schur complement of K onto the bending
// moments over the supports (Clapeyron's three-moment equations with variable EI):
//
//   span j between supports s_{j-1}, s_j :  M(x) = M0_j(x) + M_{j-1} (1 - xi) + M_j xi
//   slope continuity at support k        :  b_k M_{k-1} + (c_k + a_{k+1}) M_k + b_{k+1} M_{k+1} = -(q_k + p_{k+1})
//   a_j = int (1-xi)^2/EI, b_j = int xi(1-xi)/EI, c_j = int xi^2/EI, p_j = int M0 (1-xi)/EI, q_j = int M0 xi/EI
//
// i.e. one streaming pass over the elements (5 accumulators per span, exact element integrals), a
// tridiagonal system with at most MAXS-2 unknowns, and one statics march that yields every element's
// end shear and moment.  Per beam the FP64 state is O(#supports) instead of the 5 doubles per node
// of the banded factor (beamopt_core.cuh), which is what lets hundreds of beams per SM stay on chip.
// Against an 80-bit solve of K u = f it is ~30x more accurate than FP64 banded Cholesky (errors
// ~5e-12 vs ~2e-10 on the default bridge), so parity with the reference's solver is
// bounded by the reference's own rounding.  Displacements (needed once per beam, for the emitted
// record) come from integrating the curvature M/EI span by span (deflections_march).
//
// FP32 loss / gradient / Adam: element_update_f32 from beamopt_core.cuh (torch CPU op order), with
// torch.sum's cascade reproduced by 32 running partial sums per reduction.
#pragma once

#include "beamopt_core.cuh"

namespace ops {

constexpr int FLEX_MAXS = 6;     // pin + up to 5 rollers (every reference configuration)
constexpr int FLEX_MAXF = 8;     // point loads per beam

// Per-beam small arrays ("span store"), element k at base[k * stride].
//   doubles: per span j = 1..m (slot j-1): RA, DXI, A, B, C, P, Q  ;  per support j = 0..m: MS
//   ints   : SUP[j] j = 0..m ; LNODE[i] sorted load nodes ; then LVAL (doubles) sorted load values
struct FlexStore {
    double *sd;
    int *si;
    long stride;
    enum { RA = 0, DXI = 1, A = 2, B = 3, C = 4, P = 5, Q = 6, PER_SPAN = 7 };
    enum { MS_BASE = PER_SPAN * (FLEX_MAXS - 1), LVAL_BASE = MS_BASE + FLEX_MAXS,
           NUM_DOUBLES = LVAL_BASE + FLEX_MAXF };
    enum { SUP_BASE = 0, LNODE_BASE = FLEX_MAXS, NUM_INTS = FLEX_MAXS + FLEX_MAXF + 1 };
    OPS_HD double &span(int j, int what) const { return sd[(long)((j - 1) * PER_SPAN + what) * stride]; }
    OPS_HD double &ms(int j) const { return sd[(long)(MS_BASE + j) * stride]; }
    OPS_HD double &lval(int i) const { return sd[(long)(LVAL_BASE + i) * stride]; }
    OPS_HD int &sup(int j) const { return si[(long)(SUP_BASE + j) * stride]; }
    OPS_HD int &lnode(int i) const { return si[(long)(LNODE_BASE + i) * stride]; }
};

// fp32 optimiser state, each array element e at base[e * stride] (arrays may live in different memories)
struct OptState {
    float *Icur, *Inext, *m, *v;
    long sI, sIn, sm, sv;
    float *accd, *accq;     // 2 x 64 running partial sums (level 0: [0,32), level 1: [32,64))
    long sacc;
};

struct FlexBeam {
    int m;                  // spans between supports (supports = m + 1); m = 0 -> mechanism
    int last;               // node of the last support
    int nloads;             // point loads on free nodes
    double Le, invLe, kc6, wl, wl2h, corr;   // Le, 1/Le, Le/(6E), w*Le, w*Le^2/2, w*Le^2/4
    double Moh, Qoh;        // overhang: moment over / shear just right of the last support
    double EIk;             // Le/E
};

// Element-length dependent constants of a beam (pure arithmetic on L).
OPS_HD void flex_geometry(const BeamConsts &k, double L, FlexBeam &fb)
{
    const double Le = L / (double)k.n;
    fb.Le = Le;
    fb.invLe = 1.0 / Le;
    fb.kc6 = Le / (6.0 * k.E);
    fb.EIk = Le / k.E;
    fb.wl = k.udl * Le;
    fb.wl2h = 0.5 * k.udl * Le * Le;
    fb.corr = 0.25 * k.udl * Le * Le;
}

// One-time (per beam) set-up of the I-independent data.  fixed(i): uy_i constrained (node 0 implied).
// Returns 0 ok, 1 mechanism (no roller), 3 more supports than FLEX_MAXS.
template <class FixedFn>
OPS_HD int flex_setup(const BeamConsts &k, double L, FixedFn fixed, int nforces, const int *fnode,
                      const double *fval, const FlexStore &fs, FlexBeam &fb)
{
    const int n = k.n;
    flex_geometry(k, L, fb);
    const double Le = fb.Le;
    // supports, ascending
    int m = 0;
    fs.sup(0) = 0;
    for (int i = 1; i <= n; ++i) {
        if (fixed(i)) {
            if (m + 1 >= FLEX_MAXS) return 3;
            ++m;
            fs.sup(m) = i;
        }
    }
    fb.m = m;
    if (m == 0) return 1;
    const int last = fs.sup(m);
    fb.last = last;
    // point loads: merge duplicates, drop loads on supports (they go straight into the reaction and
    // change no moment), insertion-sort by node
    int nl = 0;
    for (int i = 0; i < nforces; ++i) {
        const int nd = fnode[i];
        if (nd < 0 || nd > n) continue;
        bool on_support = (nd == 0);
        for (int j = 1; j <= m; ++j) on_support = on_support || (fs.sup(j) == nd);
        if (on_support) continue;
        int pos = -1;
        for (int q = 0; q < nl; ++q) if (fs.lnode(q) == nd) pos = q;
        if (pos >= 0) { fs.lval(pos) += fval[i]; continue; }
        int q = nl;
        while (q > 0 && fs.lnode(q - 1) > nd) {
            fs.lnode(q) = fs.lnode(q - 1);
            fs.lval(q) = fs.lval(q - 1);
            --q;
        }
        fs.lnode(q) = nd;
        fs.lval(q) = fval[i];
        ++nl;
    }
    fs.lnode(nl) = 0x7fffffff;     // sentinel
    fb.nloads = nl;
    // simply supported left reaction of every span and the overhang statics
    for (int j = 1; j <= m; ++j) {
        const int na = fs.sup(j - 1), nb = fs.sup(j);
        const double l = (double)(nb - na) * Le;
        double mom = 0.5 * k.udl * l * l;
        for (int q = 0; q < nl; ++q) {
            const int nd = fs.lnode(q);
            if (nd > na && nd < nb) mom += fs.lval(q) * ((double)(nb - nd) * Le);
        }
        fs.span(j, FlexStore::RA) = -mom / l;
        fs.span(j, FlexStore::DXI) = 1.0 / (double)(nb - na);
    }
    const double loh = (double)(n - last) * Le;
    double moh = 0.5 * k.udl * loh * loh, qoh = k.udl * loh;
    for (int q = 0; q < nl; ++q) {
        const int nd = fs.lnode(q);
        if (nd > last) { moh += fs.lval(q) * ((double)(nd - last) * Le); qoh += fs.lval(q); }
    }
    fb.Moh = moh;
    fb.Qoh = -qoh;
    return 0;
}

// Pass 1: flexibility integrals of every span for the current inertias, then the three-moment
// system for the support moments MS[0..m].  I(e): element inertia as double.
template <class InertiaFn>
OPS_HD int flex_support_moments(const BeamConsts &k, const FlexBeam &fb, const FlexStore &fs, InertiaFn I)
{
    const int m = fb.m, last = fb.last;
    int j = 0, next_sup = 0, li = 0, next_load = fs.lnode(0);
    double a = 0.0, b = 0.0, c = 0.0, p = 0.0, q = 0.0;
    double x1 = 0.0, f1 = 1.0, dxi = 0.0, Q0 = 0.0, M0 = 0.0;
    for (int e = 0; e < last; ++e) {
        if (e == next_sup) {
            if (j > 0) {
                fs.span(j, FlexStore::A) = a; fs.span(j, FlexStore::B) = b; fs.span(j, FlexStore::C) = c;
                fs.span(j, FlexStore::P) = p; fs.span(j, FlexStore::Q) = q;
            }
            ++j;
            next_sup = fs.sup(j);
            dxi = fs.span(j, FlexStore::DXI);
            Q0 = fs.span(j, FlexStore::RA);
            M0 = 0.0; x1 = 0.0; f1 = 1.0;
            a = b = c = p = q = 0.0;
        }
        const double c6 = fb.kc6 / I(e);
        const double x2 = x1 + dxi, f2 = 1.0 - x2;
        const double m2 = fma(Q0, fb.Le, M0 + fb.wl2h);
        // elastic weights of the element for the three moment diagrams (linear end values, the
        // simply supported one also carries the UDL parabola): W1 = c6 (2 v1 + v2), W2 = c6 (v1 + 2 v2)
        const double w10 = c6 * (fma(2.0, M0, m2) - fb.corr), w20 = c6 * (fma(2.0, m2, M0) - fb.corr);
        const double w1f = c6 * fma(2.0, f1, f2), w2f = c6 * fma(2.0, f2, f1);
        const double w1x = c6 * fma(2.0, x1, x2), w2x = c6 * fma(2.0, x2, x1);
        a = fma(w1f, f1, fma(w2f, f2, a));
        b = fma(w1f, x1, fma(w2f, x2, b));
        c = fma(w1x, x1, fma(w2x, x2, c));
        p = fma(w10, f1, fma(w20, f2, p));
        q = fma(w10, x1, fma(w20, x2, q));
        Q0 += fb.wl;
        if (e + 1 == next_load) { Q0 += fs.lval(li); ++li; next_load = fs.lnode(li); }
        M0 = m2; x1 = x2; f1 = f2;
    }
    fs.span(j, FlexStore::A) = a; fs.span(j, FlexStore::B) = b; fs.span(j, FlexStore::C) = c;
    fs.span(j, FlexStore::P) = p; fs.span(j, FlexStore::Q) = q;
    // three-moment system (Thomas); MS[0] = 0 (pinned end), MS[m] = overhang moment
    int bad = 0;
    fs.ms(0) = 0.0;
    fs.ms(m) = fb.Moh;
    if (m > 1) {
        // forward elimination; the pivots / right-hand sides overwrite C[k] / Q[k], k = 1..m-1
        double dprev = 0.0, rprev = 0.0;
        for (int kk = 1; kk < m; ++kk) {
            double dd = fs.span(kk, FlexStore::C) + fs.span(kk + 1, FlexStore::A);
            double rr = -(fs.span(kk, FlexStore::Q) + fs.span(kk + 1, FlexStore::P));
            if (kk == m - 1) rr -= fs.span(m, FlexStore::B) * fb.Moh;
            if (kk > 1) {
                const double bk = fs.span(kk, FlexStore::B);
                const double wgt = bk / dprev;
                dd = fma(-wgt, bk, dd);
                rr = fma(-wgt, rprev, rr);
            }
            if (!(dd > 0.0)) bad = 1;
            fs.span(kk, FlexStore::C) = dd;
            fs.span(kk, FlexStore::Q) = rr;
            dprev = dd; rprev = rr;
        }
        double mnext = fb.Moh;
        for (int kk = m - 1; kk >= 1; --kk) {
            double rr = fs.span(kk, FlexStore::Q);
            if (kk < m - 1) rr = fma(-fs.span(kk + 1, FlexStore::B), mnext, rr);
            mnext = rr / fs.span(kk, FlexStore::C);
            fs.ms(kk) = mnext;
        }
    }
    return bad;
}

// Pass 2: statics march.  emit(e, V, M) receives eleResponse(e,'forces')[1], [2]
// (global Fy and Mz at the node-i end: V = internal shear, M = -sagging moment).
template <class EmitFn>
OPS_HD void flex_forces_march(const BeamConsts &k, const FlexBeam &fb, const FlexStore &fs, EmitFn emit)
{
    const int n = k.n, m = fb.m, last = fb.last;
    int j = 0, next_sup = 0, li = 0, next_load = fs.lnode(0);
    double Q = 0.0, Mc = 0.0;
    for (int e = 0; e < n; ++e) {
        if (e == next_sup) {
            if (e < last) {
                ++j;
                next_sup = fs.sup(j);
                const double ml = fs.ms(j - 1), mr = fs.ms(j);
                const double invl = fs.span(j, FlexStore::DXI) / fb.Le;
                Q = fma(mr - ml, invl, fs.span(j, FlexStore::RA));
                Mc = ml;
            } else {
                next_sup = 0x7fffffff;
                Q = fb.Qoh;
                Mc = fb.Moh;
            }
        }
        emit(e, Q, -Mc);
        Mc = fma(Q, fb.Le, Mc + fb.wl2h);
        Q += fb.wl;
        if (e + 1 == next_load) { Q += fs.lval(li); ++li; next_load = fs.lnode(li); }
    }
    (void)m;
}

// Displacements of the analysed model: integrate the curvature M/EI span by span.  The slope at
// the left end of span j is -(p_j + M_{j-1} a_j + M_j b_j) (needs A, B, P of flex_support_moments,
// which the Thomas sweep leaves untouched); supports are set to exactly zero.  out(i, uy, theta).
template <class InertiaFn, class OutFn>
OPS_HD void flex_deflections_march(const BeamConsts &k, const FlexBeam &fb, const FlexStore &fs,
                                   InertiaFn I, OutFn out)
{
    const int n = k.n, last = fb.last;
    const double wl2_12 = fb.wl2h / 6.0, wl2_24 = fb.wl2h / 12.0;
    int j = 0, next_sup = 0, li = 0, next_load = fs.lnode(0);
    double Q = 0.0, Mc = 0.0, u = 0.0, t = 0.0;
    for (int e = 0; e < n; ++e) {
        if (e == next_sup) {
            u = 0.0;
            if (e < last) {
                ++j;
                next_sup = fs.sup(j);
                const double ml = fs.ms(j - 1), mr = fs.ms(j);
                const double invl = fs.span(j, FlexStore::DXI) / fb.Le;
                Q = fma(mr - ml, invl, fs.span(j, FlexStore::RA));
                Mc = ml;
                t = -(fs.span(j, FlexStore::P) + fma(ml, fs.span(j, FlexStore::A), mr * fs.span(j, FlexStore::B)));
            } else {
                next_sup = 0x7fffffff;
                Q = fb.Qoh;
                Mc = fb.Moh;            // theta continues from the last span
            }
        }
        out(e, u, t);
        const double m2 = fma(Q, fb.Le, Mc + fb.wl2h);
        const double ce = fb.EIk / I(e);                       // Le / (E I)
        u = u + fma(t, fb.Le, (ce * fb.Le) * (Mc / 3.0 + m2 / 6.0 - wl2_24));
        t = t + ce * (0.5 * (Mc + m2) - wl2_12);
        Mc = m2;
        Q += fb.wl;
        if (e + 1 == next_load) { Q += fs.lval(li); ++li; next_load = fs.lnode(li); }
    }
    out(n, (last == n) ? 0.0 : u, t);
}

// ---------------------------------------------------------------------------------------------
// running torch.sum partials: x_e is added into acc[e mod 32] in ascending e, which is the order
// of ATen's cascade_sum for fewer than 16 blocks of 32 (n < 512); longer vectors flush level 0
// into level 1 every 16 blocks (n < 8192).  See torch_sum_f32 in beamopt_core.cuh for the scheme.
// ---------------------------------------------------------------------------------------------
struct SumPlan {
    int vec_size, size_ilp, main_end, vec_end;   // [0, main_end) blocks of 32, [main_end, vec_end) whole vectors -> row 0
    bool cascade;
};

OPS_HD SumPlan make_sum_plan(int n)
{
    SumPlan s;
    s.vec_size = n / 8;
    s.size_ilp = s.vec_size / 4;
    s.main_end = s.size_ilp * 32;
    s.vec_end = s.vec_size * 8;
    s.cascade = s.size_ilp >= 16;
    return s;
}

OPS_HD float finish_running_sum(const float *acc, long stride, float tail)
{
    // rows ((r0 + r1) + r2) + r3 per lane, then the scalar tail first, then the 8 lanes in order
    float s = tail;
    for (int l = 0; l < 8; ++l) {
        const float r0 = acc[(long)l * stride], r1 = acc[(long)(8 + l) * stride];
        const float r2 = acc[(long)(16 + l) * stride], r3 = acc[(long)(24 + l) * stride];
        s += ((r0 + r1) + r2) + r3;
    }
    return s;
}

// level 1 += level 0 ; level 0 = 0   (after 16 complete blocks)
OPS_HD void running_sum_flush(float *acc, long stride)
{
    for (int i = 0; i < 32; ++i) {
        float *a0 = acc + (long)i * stride, *a1 = acc + (long)(32 + i) * stride;
        *a1 += *a0;
        *a0 = 0.0f;
    }
}

// end of the block region: level 0 += level 1 (ATen merges the levels BEFORE the left-over whole
// vectors are added to row 0)
OPS_HD void running_sum_merge(float *acc, long stride)
{
    for (int i = 0; i < 32; ++i) {
        float *a0 = acc + (long)i * stride, *a1 = acc + (long)(32 + i) * stride;
        *a0 += *a1;
        *a1 = 0.0f;
    }
}

// One full iteration (single load case).  Reads Icur, writes Inext (the caller swaps), updates m, v.
// Returns the fp32 total loss; *bad set on failure.
OPS_HD float flex_iteration(const BeamConsts &k, const FlexBeam &fb, const FlexStore &fs, const OptState &os,
                            float neg_step, float bc2_sqrt, int *bad)
{
    const int n = k.n;
    *bad = flex_support_moments(k, fb, fs, [&](int e) { return (double)os.Icur[(long)e * os.sI]; });
    const SumPlan sp = make_sum_plan(n);
    const int nacc = sp.cascade ? 64 : 32;
    for (int i = 0; i < nacc; ++i) { os.accd[(long)i * os.sacc] = 0.0f; os.accq[(long)i * os.sacc] = 0.0f; }
    const float sI = torch_sum_f32(n, [&](int e) { return os.Icur[(long)e * os.sI]; });
    float taild = 0.0f, tailq = 0.0f;
    flex_forces_march(k, fb, fs, [&](int e, double V, double M) {
        const float Mf = (float)M, Vf = (float)V;
        float I = os.Icur[(long)e * os.sI], mm = os.m[(long)e * os.sm], vv = os.v[(long)e * os.sv], d, q;
        element_update_f32(k, neg_step, bc2_sqrt, Mf * Mf, Vf * Vf, I, mm, vv, d, q);
        os.Inext[(long)e * os.sIn] = I;
        os.m[(long)e * os.sm] = mm;
        os.v[(long)e * os.sv] = vv;
        if (e < sp.vec_end) {
            const int slot = (e < sp.main_end) ? (e & 31) : ((e - sp.main_end) & 7);
            float *pd = os.accd + (long)slot * os.sacc, *pq = os.accq + (long)slot * os.sacc;
            *pd += d;
            *pq += q;
            if (sp.cascade && e < sp.main_end) {
                if ((e & 511) == 511) { running_sum_flush(os.accd, os.sacc); running_sum_flush(os.accq, os.sacc); }
                if (e == sp.main_end - 1) { running_sum_merge(os.accd, os.sacc); running_sum_merge(os.accq, os.sacc); }
            }
        } else {
            taild += d;
            tailq += q;
        }
    });
    const float sd = finish_running_sum(os.accd, os.sacc, taild);
    const float sq = finish_running_sum(os.accq, os.sacc, tailq);
    return (sI + k.am * sd) + k.as_ * sq;
}

}  // namespace ops
