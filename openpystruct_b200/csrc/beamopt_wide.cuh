// Shared-memory-state form of the production iteration: LPB lanes per beam (8 or 32), the fp32 optimiser
// state in shared memory instead of registers.  Two uses:
//   LPB = 32 (one warp per beam, rolled loop): fine discretisations, any n that fits shared memory;
//   LPB = 8  (four beams per warp, unrolled):  the reference's 100-element beams at twice the resident
//            beams per SM of beamopt_lanes.cuh, whose iteration is latency bound (profiles/).
//
// Same iteration as beamopt_lanes.cuh -- the three-moment solve of beamopt_flex.cuh with the fp32 loss /
// frozen-M,V gradient / Adam half in torch's CPU operation order -- replacing, per epoch (reference
// file:line): setup_model + analyze + eleResponse (OpenPyStruct_BeamOpt_training_SingleCore.py:176-190),
// the loss (:195-199), backward / Adam / ExponentialLR / clamp (:202-208), the early-stop test (:211-219).
//
// Mapping.  Lane l of the beam's LPB-lane group owns the elements e = LPB k + l ("slot" k).  Element e
// feeds accumulator e mod 32 of the 32 partial sums ATen's cascade_sum keeps (8-float vectors x 4 ILP
// rows): one running sum per lane for LPB = 32 (plus the level-1 flush every 16 slots for n >= 512),
// four row sums per lane for LPB = 8, so torch.sum's order falls out of lane-local additions and one
// fixed-order combine.  The fp32 state (I double-buffered, Adam's m and v) sits in shared memory in
// plain element order, which is conflict free for this mapping.
//
// I-independent statics as per-SEGMENT polynomials instead of per-element values.  Between two
// consecutive supports / load nodes the simply supported span statics of element e are
//     M0 = wl2h ke^2 + c1 ke + c0 ,   Q0 = wl ke + q0            (ke = e - first node of the span)
// and the three-moment result adds MS_l + (MS_r - MS_l) d ke to M0 and (MS_r - MS_l) d / Le to Q0, i.e.
// only changes c0, c1, q0.  After the solve of an epoch the <= 15 segment rows are rewritten
// ({C0, C1, Qc}) and every element evaluates  M = (wl2h ke + C1) ke + C0,  V = wl ke + Qc  -- three
// DFMAs and one row fetch.  The flexibility weights of PASS 1 (G = g1 + g2 and g2 of beamopt_lanes.cuh)
// are quadratics of ke with static per-segment coefficients in the same row.
//
// One sweep per epoch: forces -> loss terms -> gradient -> Adam -> 1 / I_new -> flexibility sums of the
// NEXT epoch, so I, m, v are read and written once.  The analysed inertias stay in the other I buffer
// for the record (SingleCore:221-249: fields of the last analysed model, I after the last step).
//
// Span sums: a lane keeps five running sums for the span its slots are in; when a span's last element
// falls into slot k (a group-uniform event), every lane contributes its sums of that span and the group
// adds them in the order of a fixed butterfly (deterministic).  No per-lane scratch.  Which spans close in a
// slot is one byte per slot (`cmask`, written at set-up): the 32-lane kernel tests a batch's four bytes at once
// and walks the set bits, so a sweep has no scan over the spans and no warp votes (beamopt_wide.cu).
//
// The functions here contain no CUDA intrinsics: the two cross-lane steps (butterfly sum, final loss
// combine) are done by the caller -- shuffles in beamopt_wide.cu, arrays in tests/hostsim.
#pragma once

#include "beamopt_flex.cuh"
#include "fastmath.cuh"
#include "beamopt_lanes.cuh"      // the packed fp32 chain (OPS_FP32_CHAIN, OPS_ADAM_STEP): one source for both kernel families

namespace ops {
namespace wide {

#ifndef OPS_WIDE_NB
#define OPS_WIDE_NB 4
#endif
constexpr int NB = OPS_WIDE_NB;                 // slots per batch
constexpr int NSPAN = FLEX_MAXS - 1;
constexpr int DUMMY = NSPAN;                    // span id of overhang elements and padding
constexpr int NSUM = 5;                         // R0, R1, R2, G, Q of beamopt_lanes.cuh
constexpr int MAXSEG = 16;                      // spans + loads + overhang + padding <= 5 + 8 + 1 + 1
// row = 16-byte pairs, fetched with 128-bit shared loads in the order the sweep needs them.  The statics
// {c0, c1, q0} a row is rewritten FROM every epoch: LPB = 8 keeps them behind the row (12 doubles); LPB = 32 --
// segment s is rewritten by lane s (wide_solve) -- keeps them in that lane's registers (SegStatics, rows of 8
// doubles): 512 B per beam less, which is the twelfth 1000-element beam of an SM.
template <int LPB>
OPS_HD constexpr int row_doubles() { return LPB == 32 ? 8 : 12; }
enum { R_NA = 0, R_C0 = 1,                      // first node of the span; this epoch's C0
       R_C1 = 2, R_QC = 3,                      // this epoch's C1, Qc
       R_G1 = 4, R_H1 = 5,                      // PASS 1 weights G = (G2 ke + G1) ke + G0, g2 = (H2 ke + H1) ke + H0
       R_G0 = 6, R_H0 = 7,
       R_B0 = 8, R_B1 = 9, R_BQ = 10 };         // LPB = 8: statics {c0, c1, q0}
struct alignas(16) Pair {                       // two doubles moved with one 128-bit shared access
    double x, y;
};
enum { GI_M = 0, GI_LAST = 1, GI_NLOADS = 2, GI_RC = 3, GI_NSEG = 4, GI_CLOSE = 5, GI_SEGSPAN = GI_CLOSE + NSPAN,
       GI_SEGSTART = GI_SEGSPAN + MAXSEG, GI_INTS = GI_SEGSTART + MAXSEG + 1 };

// shared memory of one beam
struct WideStore {
    float *I0, *m, *v;                          // [K * LPB], element order; I is double-buffered: I0 + (epoch & 1) * el
    int el;                                     // K * LPB
    unsigned char *seg;                         // [K * LPB]: segment id | span id << 4 | (a span closes in this slot) << 7
    unsigned char *cmask;                       // [K]: bit j = span j closes in this slot (the spans behind bit 7 of seg)
    double *rows;                               // [MAXSEG * row_doubles<LPB>()]
    double *tot;                                // [NSPAN * NSUM] span sums of the inertias about to be analysed
    double *gd;                                 // Moh, Qoh
    int *gi;                                    // [GI_INTS]
    FlexStore fs;                               // stride 1
};

// torch.sum layout of an n-vector over slots (beamopt_core.cuh, torch_sum_f32)
struct WideShape {
    int n, K;                                   // K slots of LPB elements
    int size_ilp;                               // whole 32-blocks
    int blk;                                    // slots of whole 32-blocks (4 ILP rows)
    int vec;                                    // LPB = 8: slots of whole 8-vectors; LPB = 32: = blk
    int casc_slots;                             // LPB = 32: slots covered by complete 16-block cascade chunks (0: none)
    int nlv, ntail;                             // left-over 8-vectors (0..3) and scalar tail (0..7) behind the blocks
};
// slots per lane: LPB = 8 runs compile-time slot counts (the instances of beamopt_wide.cu), padded with
// inert elements behind n; LPB = 32 runs any count
template <int LPB>
OPS_HD int wide_slots(int n)
{
    if (LPB == 8) return n <= 32 ? 4 : (n <= 64 ? 8 : (n <= 104 ? 13 : 21));
    return (n + LPB - 1) / LPB;
}

template <int LPB>
OPS_HD WideShape wide_shape(int n)
{
    WideShape s;
    s.n = n;
    s.K = wide_slots<LPB>(n);
    const int vec = n / 8;
    s.size_ilp = vec / 4;
    s.blk = s.size_ilp * (32 / LPB);
    s.vec = LPB == 8 ? vec : s.blk;
    s.casc_slots = s.size_ilp >= 16 ? (s.size_ilp / 16) * 16 : 0;
    s.nlv = vec - 4 * s.size_ilp;
    s.ntail = n - 8 * vec;
    return s;
}
// LPB = 8: up to 21 slots, sums without the cascade levels; LPB = 32 handles one cascade level (n < 8192)
template <int LPB>
OPS_HD bool wide_shape_ok(int n) { return n >= 1 && (LPB == 8 ? n <= 168 : n < 8192); }

template <int LPB>
OPS_HD size_t wide_beam_bytes(int n)
{
    const size_t el = (size_t)wide_slots<LPB>(n) * LPB;
    size_t b = (size_t)(MAXSEG * row_doubles<LPB>() + NSPAN * NSUM + 2 + FlexStore::NUM_DOUBLES) * 8;   // doubles first
    b += 4 * el * 4;
    b += (size_t)(GI_INTS + FlexStore::NUM_INTS) * 4;
    b += el + (size_t)wide_slots<LPB>(n);
    return (b + 15) / 16 * 16;
}
template <int LPB>
OPS_HD void wide_carve(unsigned char *base, int n, WideStore &ws)
{
    const size_t el = (size_t)wide_slots<LPB>(n) * LPB;
    double *d = reinterpret_cast<double *>(base);
    ws.rows = d; d += MAXSEG * row_doubles<LPB>();
    ws.tot = d; d += NSPAN * NSUM;
    ws.gd = d; d += 2;
    ws.fs.sd = d; d += FlexStore::NUM_DOUBLES;
    float *f = reinterpret_cast<float *>(d);
    ws.I0 = f; f += 2 * el;
    ws.el = (int)el;
    ws.m = f; f += el;
    ws.v = f; f += el;
    int *i = reinterpret_cast<int *>(f);
    ws.gi = i; i += GI_INTS;
    ws.fs.si = i; i += FlexStore::NUM_INTS;
    ws.fs.stride = 1;
    ws.seg = reinterpret_cast<unsigned char *>(i);
    ws.cmask = ws.seg + el;
}

// ---------------------------------------------------------------------------------------------
// once per beam, lane 0: supports, loads and span reactions (flex_setup), then the segment rows
// ---------------------------------------------------------------------------------------------
template <int LPB, class FixedFn>
OPS_HD int wide_setup(const BeamConsts &k, double L, FixedFn fixed, const int *fnode, const double *fval,
                      const WideStore &ws)
{
    FlexBeam fb;
    fb.m = 0; fb.last = 0; fb.nloads = 0; fb.Moh = 0.0; fb.Qoh = 0.0;
    const int rc = flex_setup(k, L, fixed, k.max_forces, fnode, fval, ws.fs, fb);
    ws.gi[GI_M] = rc ? 0 : fb.m; ws.gi[GI_LAST] = fb.last; ws.gi[GI_NLOADS] = fb.nloads; ws.gi[GI_RC] = rc;
    ws.gd[0] = fb.Moh; ws.gd[1] = fb.Qoh;
    for (int i = 0; i < NSPAN * NSUM; ++i) ws.tot[i] = 0.0;
    for (int j = 0; j < NSPAN; ++j) ws.gi[GI_CLOSE + j] = -1;
    for (int s_ = 0; s_ < wide_slots<LPB>(k.n); ++s_) ws.cmask[s_] = 0;
    int ns = 0;
    if (rc == 0) {
        const int n = k.n, m = fb.m, last = fb.last, nl = fb.nloads;
        const double Le = fb.Le;
        const double cG = fma(3.0, fb.wl2h, -2.0 * fb.corr), cH = fma(2.0, fb.wl2h, -fb.corr);
        const double k3Le = 3.0 * Le, k2Le = 2.0 * Le;
        auto emit = [&](int ea, int span, int na, double c0, double c1, double q0) {
            constexpr int ROW = row_doubles<LPB>();
            double *r = ws.rows + ns * ROW;
            const bool live = span != DUMMY;
            r[R_NA] = (double)na;
            r[R_C0] = c0; r[R_C1] = c1; r[R_QC] = q0;            // (LPB = 32: the statics, until wide_fetch_statics has taken them)
            if (ROW > 8) { r[R_B0] = c0; r[R_B1] = c1; r[R_BQ] = q0; r[11] = 0.0; }
            r[R_G1] = live ? fma(6.0, c1, k3Le * fb.wl) : 0.0;
            r[R_G0] = live ? fma(6.0, c0, fma(k3Le, q0, cG)) : 0.0;
            r[R_H1] = live ? fma(3.0, c1, k2Le * fb.wl) : 0.0;
            r[R_H0] = live ? fma(3.0, c0, fma(k2Le, q0, cH)) : 0.0;
            ws.gi[GI_SEGSTART + ns] = ea;
            ws.gi[GI_SEGSPAN + ns] = span;
            ++ns;
        };
        int li = 0;
        for (int j = 1; j <= m; ++j) {
            const int na = ws.fs.sup(j - 1), nb = ws.fs.sup(j);
            const double ra = ws.fs.span(j, FlexStore::RA);
            ws.gi[GI_CLOSE + j - 1] = (nb - 1) / LPB;
            ws.cmask[(nb - 1) / LPB] |= (unsigned char)(1u << (j - 1));
            double SP = 0.0, D = 0.0;                    // sum P, sum P (na - nd) over the loads left of the segment
            emit(na, j - 1, na, 0.0, Le * ra, ra);
            while (li < nl && ws.fs.lnode(li) < nb) {
                const int nd = ws.fs.lnode(li);
                const double P = ws.fs.lval(li);
                SP += P;
                D = fma(P, (double)(na - nd), D);
                emit(nd, j - 1, na, Le * D, Le * (ra + SP), ra + SP);
                ++li;
            }
        }
        if (last < n) {
            // overhang: M0 = wl2h r^2 + sum_{nd > e} P (nd - e) Le, Q0 = -(wl r + sum_{nd > e} P), r = n - e
            const double R = (double)(n - last);
            int ea = last;
            for (;;) {
                double SP = 0.0, D = 0.0;
                int next = n;
                for (int q = 0; q < nl; ++q) {
                    const int nd = ws.fs.lnode(q);
                    if (nd > ea) {
                        SP += ws.fs.lval(q);
                        D = fma(ws.fs.lval(q), (double)(nd - last), D);
                        if (nd < next) next = nd;
                    }
                }
                emit(ea, DUMMY, last, fma(fb.wl2h, R * R, Le * D), fma(-Le, SP, -2.0 * fb.wl2h * R), -fma(fb.wl, R, SP));
                if (next >= n) break;
                ea = next;
            }
        }
        emit(n, DUMMY, 0, 0.0, 0.0, 0.0);                // padding slots e >= n
    }
    ws.gi[GI_NSEG] = ns;
    return rc;
}

// every lane, after wide_setup is visible
OPS_HD int wide_fetch(const BeamConsts &k, double L, const WideStore &ws, FlexBeam &fb)
{
    flex_geometry(k, L, fb);
    fb.m = ws.gi[GI_M]; fb.last = ws.gi[GI_LAST]; fb.nloads = ws.gi[GI_NLOADS];
    fb.Moh = ws.gd[0]; fb.Qoh = ws.gd[1];
    return ws.gi[GI_RC];
}

// LPB = 32: statics {c0, c1, q0} of segment l, the one lane l rewrites -- registers on the device
struct SegStatics {
    double b[3];
};
// every lane, after wide_setup is visible and before the first wide_solve (the rows still hold the statics)
template <int LPB>
OPS_HD void wide_fetch_statics(const WideStore &ws, int l, SegStatics &st)
{
    static_assert(MAXSEG <= 32, "one segment per lane");
    st.b[0] = st.b[1] = st.b[2] = 0.0;
    if (LPB == 32 && l < ws.gi[GI_NSEG]) {
        const double *r = ws.rows + l * row_doubles<LPB>();
        st.b[0] = r[R_C0]; st.b[1] = r[R_C1]; st.b[2] = r[R_QC];
    }
}

template <int LPB>
OPS_HD void wide_lane_init(const BeamConsts &k, const WideShape &sh, const WideStore &ws, int l)
{
    const int n = sh.n, ns = ws.gi[GI_NSEG];
    for (int s = 0; s < sh.K; ++s) {
        const int e = LPB * s + l;
        int sid = 0;
        for (int q = 1; q < ns; ++q) sid = (ws.gi[GI_SEGSTART + q] <= e) ? q : sid;
        unsigned byte = (unsigned)sid | ((unsigned)ws.gi[GI_SEGSPAN + sid] << 4);
        for (int j = 0; j < NSPAN; ++j) byte |= (ws.gi[GI_CLOSE + j] == s) ? 0x80u : 0u;
        ws.seg[e] = (unsigned char)byte;
        const float I0 = e < n ? k.I0f : 1.0f;           // padding: harmless inertia, never read back
        ws.I0[e] = I0; ws.I0[ws.el + e] = I0; ws.m[e] = 0.0f; ws.v[e] = 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 half, batches of N elements written ONE OPERATION PER STAGE across the batch (see
// beamopt_lanes.cuh, lane_forces / lane_adam: same sequences, same operand ranges)
// ---------------------------------------------------------------------------------------------
#define OPS_W _Pragma("unroll") for (int i = 0; i < N; ++i)

// loss terms d = c / b, q = h / gg and autograd's gradient g with c = M^2, h = V^2 constant
template <int N>
OPS_HD void loss_grad_batch(const BeamConsts &k, const float (&I)[N], const float (&c)[N], const float (&h)[N],
                            float (&d)[N], float (&q)[N], float (&g)[N])
{
    float s[N], gg[N], rs[N], db[N], qg[N];
#if defined(__CUDA_ARCH__)
    float b[N], y[N], rb[N], rgg[N], t0[N], t1[N], t2[N];
    OPS_W { b[i] = k.E2 * I[i]; y[i] = fm::rsq_a(I[i]); }
    OPS_W b[i] = b[i] + k.epsf;
    OPS_W { rb[i] = fm::rcp_a(b[i]); t0[i] = I[i] * y[i]; y[i] = y[i] * 0.5f; }
    OPS_W { t1[i] = fmaf(-b[i], rb[i], 1.0f); t2[i] = fmaf(-t0[i], t0[i], I[i]); }
    OPS_W { rb[i] = fmaf(rb[i], t1[i], rb[i]); s[i] = fmaf(t2[i], y[i], t0[i]); }
    OPS_W { t0[i] = c[i] * rb[i]; gg[i] = k.kf * s[i]; rs[i] = fm::rcp_a(s[i]); }
    OPS_W { t1[i] = fmaf(-b[i], t0[i], c[i]); gg[i] = k.Gf * gg[i]; t2[i] = fmaf(-s[i], rs[i], 1.0f); }
    OPS_W { d[i] = fmaf(rb[i], t1[i], t0[i]); rgg[i] = fm::rcp_a(gg[i]); rs[i] = fmaf(rs[i], t2[i], rs[i]); }
    OPS_W { t0[i] = d[i] * rb[i]; t1[i] = fmaf(-gg[i], rgg[i], 1.0f); t2[i] = fmaf(-s[i], rs[i], 1.0f); }
    OPS_W { db[i] = fmaf(-b[i], t0[i], d[i]); rgg[i] = fmaf(rgg[i], t1[i], rgg[i]); rs[i] = fmaf(rs[i], t2[i], rs[i]); }
    OPS_W { db[i] = fmaf(rb[i], db[i], t0[i]); t1[i] = h[i] * rgg[i]; rs[i] = 0.5f * rs[i]; }
    OPS_W { db[i] = (-k.am) * db[i]; t2[i] = fmaf(-gg[i], t1[i], h[i]); }
    OPS_W { db[i] = db[i] * k.E2; q[i] = fmaf(rgg[i], t2[i], t1[i]); }
    OPS_W t0[i] = q[i] * rgg[i];
    OPS_W t1[i] = fmaf(-gg[i], t0[i], q[i]);
    OPS_W qg[i] = fmaf(rgg[i], t1[i], t0[i]);
#else
    OPS_W {
        const float b = k.E2 * I[i] + k.epsf;
        s[i] = sqrtf(I[i]);
        d[i] = c[i] / b;
        db[i] = ((-k.am) * (d[i] / b)) * k.E2;
        gg[i] = k.Gf * (k.kf * s[i]);
        q[i] = h[i] / gg[i];
        qg[i] = q[i] / gg[i];
        rs[i] = 0.5f * (1.0f / s[i]);
    }
#endif
    OPS_W qg[i] = (-k.as_) * qg[i];
    OPS_W qg[i] = qg[i] * k.Gf;
    OPS_W qg[i] = qg[i] * k.kf;
    OPS_W qg[i] = qg[i] * rs[i];
    OPS_W qg[i] = 1.0f + qg[i];
    OPS_W g[i] = qg[i] + db[i];
}

// torch's single-tensor Adam step + clamp (element_update_f32, second half); rbc = fm::rcp_r(bc2_sqrt)
template <int N>
OPS_HD void adam_batch(const BeamConsts &k, float neg_step, float bc2_sqrt, float rbc, float (&I)[N], float (&m)[N],
                       float (&v)[N], const float (&g)[N])
{
    {
        float t0[N], t1[N];
        OPS_W { t0[i] = g[i] - m[i]; t1[i] = k.omb2f * g[i]; v[i] = v[i] * k.b2f; }
        OPS_W { m[i] = fmaf(k.w1, t0[i], m[i]); v[i] = fmaf(t1[i], g[i], v[i]); }
    }
#if defined(__CUDA_ARCH__)
    bool rare = false;
    OPS_W rare = rare || !(v[i] >= fm::SQRT_F_MIN);
    if (!rare) {
        float y[N], den[N], rd[N], t0[N], t1[N], num[N];
        OPS_W y[i] = fm::rsq_a(v[i]);
        OPS_W { t0[i] = v[i] * y[i]; y[i] = y[i] * 0.5f; num[i] = neg_step * m[i]; }
        OPS_W t1[i] = fmaf(-t0[i], t0[i], v[i]);
        OPS_W t0[i] = fmaf(t1[i], y[i], t0[i]);
        OPS_W t1[i] = t0[i] * rbc;
        OPS_W den[i] = fmaf(-bc2_sqrt, t1[i], t0[i]);
        OPS_W den[i] = fmaf(rbc, den[i], t1[i]);
        OPS_W den[i] = den[i] + k.adam_epsf;
        OPS_W rd[i] = fm::rcp_a(den[i]);
        OPS_W t0[i] = fmaf(-den[i], rd[i], 1.0f);
        OPS_W rd[i] = fmaf(rd[i], t0[i], rd[i]);
        OPS_W t0[i] = num[i] * rd[i];
        OPS_W t1[i] = fmaf(-den[i], t0[i], num[i]);
        OPS_W t0[i] = fmaf(rd[i], t1[i], t0[i]);
        OPS_W t0[i] = I[i] + t0[i];
        OPS_W I[i] = t0[i] < k.clampf ? k.clampf : t0[i];
        return;
    }
#else
    (void)rbc;
#endif
    OPS_W {
        const float den = sqrtf(v[i]) / bc2_sqrt + k.adam_epsf;
        const float x = I[i] + (neg_step * m[i]) / den;
        I[i] = x < k.clampf ? k.clampf : x;
    }
}
#undef OPS_W

// The same two halves on PAIRS of slots (2 j, 2 j + 1) with sm_100a's packed fp32 instructions -- the stage-major chain of
// beamopt_lanes.cuh itself (FFMA2 / FMUL2 / FADD2, each half rounded like the scalar instruction; un-fused product-then-
// sum sites written with scalar adds, scripts/sass_packed_audit.py).  Used for batches of an even slot count.
template <int NPB>
OPS_HD void loss_grad_pairs(const BeamConsts &k, const fm::F2 (&I)[NPB], const fm::F2 (&c)[NPB], const fm::F2 (&h)[NPB],
                            fm::F2 (&d)[NPB], fm::F2 (&q)[NPB], fm::F2 (&g)[NPB])
{
    using fm::F2; using fm::splat; using fm::neg2; using fm::mul2; using fm::add2; using fm::fma2;
    constexpr int NB = NPB;
    const F2 one = splat(1.0f), half = splat(0.5f);
    F2 nb[NB], y[NB], rb[NB], s[NB], gg[NB], rgg[NB], rs[NB], db[NB], qg[NB], t0[NB], t1[NB], t2[NB];
#define OPS_P _Pragma("unroll") for (int i = 0; i < NB; ++i)
    OPS_FP32_CHAIN
#undef OPS_P
}

template <int NPB>
OPS_HD void adam_pairs(const BeamConsts &k, float neg_step, float bc2_sqrt, float rbc, fm::F2 (&I)[NPB], fm::F2 (&m_)[NPB],
                       fm::F2 (&v_)[NPB], const fm::F2 (&g)[NPB])
{
    using fm::F2; using fm::splat; using fm::neg2; using fm::mul2; using fm::add2; using fm::fma2;
    constexpr int NB = NPB, p0 = 0;
    const F2 one = splat(1.0f), half = splat(0.5f);
    F2 y[NB], t0[NB], t1[NB], t2[NB];
    struct { F2 I[NB]; } rg;
#define OPS_P _Pragma("unroll") for (int i = 0; i < NB; ++i)
    OPS_ADAM_STEP
    OPS_P I[i] = rg.I[p0 + i];
#undef OPS_P
    (void)one;
}

// ---------------------------------------------------------------------------------------------
// per epoch
// ---------------------------------------------------------------------------------------------
// registers of a lane that live across the slots of one sweep
template <int LPB>
struct LaneCtx {
    static constexpr int RPL = 32 / LPB;        // torch.sum accumulators per lane
    double a[NSUM];                             // running flexibility sums of span sa
    int sa;
    float acc[3][RPL];                          // partial sums of I, d, q (level 0)
    float lv[3];                                // LPB = 32: cascade level 1
    float left[3];                              // LPB = 32: the lane's element behind the blocks; LPB = 8: scalar tail
    double ed;                                  // element index of the lane in slot 0, as double
};
template <int LPB>
OPS_HD void ctx_reset(LaneCtx<LPB> &cx, int l)
{
#pragma unroll
    for (int i = 0; i < NSUM; ++i) cx.a[i] = 0.0;
    cx.sa = DUMMY;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int r = 0; r < LaneCtx<LPB>::RPL; ++r) cx.acc[i][r] = 0.0f;
        cx.lv[i] = 0.0f; cx.left[i] = 0.0f;
    }
    cx.ed = (double)l;
}

// what a batch hands to the span bookkeeping, per slot: 1 / I_new, ke, the weights G and g1 x1 + g2 x2
template <int N>
struct BatchOut {
    double r[N], ke[N], G[N], H[N];
    int sp[N];
    bool close[N];
};

struct SweepConsts {
    double G2, H2;                              // 6 wl2h, 3 wl2h
    float neg_step, bc2_sqrt, rbc;
};

// One batch of N slots kb .. kb + N - 1 of lane l (callers with LPB = 8 pass compile-time kb so that the
// accumulator rows are static).  run = false: only the flexibility terms of the inertias in Icur (the
// first pass of a beam; Inew, m, v are not touched).
template <int LPB, int N>
OPS_HD void sweep_batch(const BeamConsts &k, const FlexBeam &fb, const WideShape &sh, const WideStore &ws, int l, int kb,
                        bool run, const float *Icur, float *Inew, const SweepConsts &sc, LaneCtx<LPB> &cx, BatchOut<N> &bo)
{
    constexpr int RPL = LaneCtx<LPB>::RPL;
    float I[N];
    const Pair *row[N];
    unsigned sb[N];
    Pair nc[N];                                  // {NA, C0}
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int e = LPB * (kb + i) + l;
        sb[i] = ws.seg[e];
        I[i] = Icur[e];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        row[i] = reinterpret_cast<const Pair *>(ws.rows + (sb[i] & 0x0fu) * row_doubles<LPB>());
        nc[i] = row[i][0];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) bo.ke[i] = (cx.ed + (double)(LPB * (kb + i))) - nc[i].x;
    if (run) {
        float m[N], v[N], c[N], h[N], d[N], q[N], g[N];
        double Mc[N], Qv[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int e = LPB * (kb + i) + l;
            m[i] = ws.m[e]; v[i] = ws.v[e];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const Pair cq = row[i][1];           // {C1, Qc}
            Mc[i] = fma(fb.wl2h, bo.ke[i], cq.x); Qv[i] = fma(fb.wl, bo.ke[i], cq.y);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) Mc[i] = fma(Mc[i], bo.ke[i], nc[i].y);
#pragma unroll
        for (int i = 0; i < N; ++i) { c[i] = (float)Mc[i]; h[i] = (float)Qv[i]; }
#pragma unroll
        for (int i = 0; i < N; ++i) { c[i] = c[i] * c[i]; h[i] = h[i] * h[i]; }
        if constexpr (N % 2 == 0) {
            fm::F2 Ip[N / 2], cp[N / 2], hp[N / 2], dp[N / 2], qp[N / 2], gp[N / 2];
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                Ip[j] = fm::f2(I[2 * j], I[2 * j + 1]); cp[j] = fm::f2(c[2 * j], c[2 * j + 1]); hp[j] = fm::f2(h[2 * j], h[2 * j + 1]);
            }
            loss_grad_pairs<N / 2>(k, Ip, cp, hp, dp, qp, gp);
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                d[2 * j] = dp[j].x; d[2 * j + 1] = dp[j].y; q[2 * j] = qp[j].x; q[2 * j + 1] = qp[j].y;
                g[2 * j] = gp[j].x; g[2 * j + 1] = gp[j].y;
            }
        } else {
            loss_grad_batch<N>(k, I, c, h, d, q, g);
        }
        // torch.sum partials.  LPB = 32, the batch inside the whole 32-blocks and inside one 16-block cascade chunk (all
        // but the last batch of a sweep): one accumulator per quantity, one test per batch instead of three per slot
        const bool plain = LPB == 32 && kb + N <= sh.blk && (kb & 15) + N <= 16;
        if (plain) {
#pragma unroll
            for (int i = 0; i < N; ++i) { cx.acc[0][0] += I[i]; cx.acc[1][0] += d[i]; cx.acc[2][0] += q[i]; }
            if (kb + N <= sh.casc_slots && ((kb + N) & 15) == 0) {                   // a 16-block chunk is complete
#pragma unroll
                for (int w = 0; w < 3; ++w) { cx.lv[w] += cx.acc[w][0]; cx.acc[w][0] = 0.0f; }
            }
        } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int s = kb + i;
            if (s < sh.blk) {
                cx.acc[0][s & (RPL - 1)] += I[i]; cx.acc[1][s & (RPL - 1)] += d[i]; cx.acc[2][s & (RPL - 1)] += q[i];
            } else if (s < sh.vec) {                     // LPB = 8: whole vectors behind the blocks go to row 0
                cx.acc[0][0] += I[i]; cx.acc[1][0] += d[i]; cx.acc[2][0] += q[i];
            } else if (s == sh.vec) {
                const bool in = LPB * s + l < sh.n;
                cx.left[0] = in ? I[i] : 0.0f; cx.left[1] = in ? d[i] : 0.0f; cx.left[2] = in ? q[i] : 0.0f;
            }
            if (LPB == 32 && s + 1 <= sh.casc_slots && ((s + 1) & 15) == 0) {   // a 16-block chunk is complete
#pragma unroll
                for (int w = 0; w < 3; ++w) { cx.lv[w] += cx.acc[w][0]; cx.acc[w][0] = 0.0f; }
            }
        }
        }
        if constexpr (N % 2 == 0) {
            fm::F2 Ip[N / 2], mp[N / 2], vp[N / 2], gp[N / 2];
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                Ip[j] = fm::f2(I[2 * j], I[2 * j + 1]); mp[j] = fm::f2(m[2 * j], m[2 * j + 1]);
                vp[j] = fm::f2(v[2 * j], v[2 * j + 1]); gp[j] = fm::f2(g[2 * j], g[2 * j + 1]);
            }
            adam_pairs<N / 2>(k, sc.neg_step, sc.bc2_sqrt, sc.rbc, Ip, mp, vp, gp);
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                I[2 * j] = Ip[j].x; I[2 * j + 1] = Ip[j].y; m[2 * j] = mp[j].x; m[2 * j + 1] = mp[j].y;
                v[2 * j] = vp[j].x; v[2 * j + 1] = vp[j].y;
            }
        } else {
            adam_batch<N>(k, sc.neg_step, sc.bc2_sqrt, sc.rbc, I, m, v, g);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int e = LPB * (kb + i) + l;
            Inew[e] = I[i]; ws.m[e] = m[i]; ws.v[e] = v[i];
        }
    }
    {
        double Id[N];
        Pair g0[N];                              // {G0, H0}
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const Pair g1 = row[i][2];           // {G1, H1}
            g0[i] = row[i][3];
            Id[i] = (double)I[i];
            bo.G[i] = fma(sc.G2, bo.ke[i], g1.x);
            bo.H[i] = fma(sc.H2, bo.ke[i], g1.y);
        }
#if defined(__CUDA_ARCH__)
        double er[N];
#pragma unroll
        for (int i = 0; i < N; ++i) bo.r[i] = fm::rcp64_a(Id[i]);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            er[i] = fma(-Id[i], bo.r[i], 1.0);
            bo.G[i] = fma(bo.G[i], bo.ke[i], g0[i].x);
            bo.H[i] = fma(bo.H[i], bo.ke[i], g0[i].y);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) er[i] = fma(er[i], er[i], er[i]);
#pragma unroll
        for (int i = 0; i < N; ++i) bo.r[i] = fma(bo.r[i], er[i], bo.r[i]);
#else
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bo.r[i] = 1.0 / Id[i];
            bo.G[i] = fma(bo.G[i], bo.ke[i], g0[i].x);
            bo.H[i] = fma(bo.H[i], bo.ke[i], g0[i].y);
        }
#endif
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bo.H[i] = fma(bo.G[i], bo.ke[i], bo.H[i]);
            bo.sp[i] = (int)((sb[i] >> 4) & 7u);
            bo.close[i] = (sb[i] & 0x80u) != 0;
        }
    }
}

// flexibility contributions of slot i: r, r ke, r ke^2, r G, r (G ke + g2)
template <int N>
OPS_HD void slot_terms(const BatchOut<N> &bo, int i, double (&x)[NSUM])
{
    const double t = bo.r[i] * bo.ke[i];
    x[0] = bo.r[i]; x[1] = t; x[2] = t * bo.ke[i]; x[3] = bo.r[i] * bo.G[i]; x[4] = bo.r[i] * bo.H[i];
}

// slot i enters the lane's running sums (restart when the lane's slots change span);
// aold / sold = the sums before, needed where a span closes in this slot
template <int LPB>
OPS_HD void slot_accumulate(LaneCtx<LPB> &cx, const double (&x)[NSUM], int sp, double (&aold)[NSUM], int &sold)
{
    const double keep = (sp == cx.sa) ? 1.0 : 0.0;
    sold = cx.sa;
#pragma unroll
    for (int w = 0; w < NSUM; ++w) { aold[w] = cx.a[w]; cx.a[w] = fma(cx.a[w], keep, x[w]); }
    cx.sa = sp;
}

// the lane's share of span j when j closes in this slot: its sums of j so far and / or this slot's element
OPS_HD void close_value(const double (&aold)[NSUM], int sold, const double (&x)[NSUM], int sp, int j, double (&v)[NSUM])
{
    const double fo = (sold == j) ? 1.0 : 0.0, fx = (sp == j) ? 1.0 : 0.0;
#pragma unroll
    for (int w = 0; w < NSUM; ++w) v[w] = fma(x[w], fx, aold[w] * fo);
}

// end of a sweep: merge the cascade level (ATen: level 0 += level 1 before the left-over vectors)
template <int LPB>
OPS_HD void ctx_finish(const WideShape &sh, LaneCtx<LPB> &cx)
{
    if (LPB == 32 && sh.casc_slots) {
#pragma unroll
        for (int w = 0; w < 3; ++w) cx.acc[w][0] += cx.lv[w];
    }
}

// the lane's row sum ((r0 + r1) + r2) + r3 of ATen's final combine (LPB = 8)
template <int LPB>
OPS_HD float ctx_rowsum(const LaneCtx<LPB> &cx, int w)
{
    float s = cx.acc[w][0];
#pragma unroll
    for (int r = 1; r < LaneCtx<LPB>::RPL; ++r) s += cx.acc[w][r];
    return s;
}

// total loss from the lanes' partials in torch's order (host form; beamopt_wide.cu does the same
// additions with shuffles).  LPB = 32: acc[w][lane] = level-0 sums, left[w][lane] = slot behind the blocks;
// LPB = 8: acc[w][lane] = ctx_rowsum, left[w][lane] = scalar tail.
template <int LPB>
OPS_HD float wide_loss_arrays(const BeamConsts &k, const WideShape &sh, const float (*acc)[LPB], const float (*left)[LPB])
{
    float s[3];
    for (int w = 0; w < 3; ++w) {
        float t = 0.0f;
        if (LPB == 32) {
            float a8[8];
            for (int l = 0; l < 8; ++l) {
                a8[l] = acc[w][l];
                for (int v = 0; v < sh.nlv; ++v) a8[l] += left[w][8 * v + l];
            }
            for (int i = 0; i < sh.ntail; ++i) t += left[w][8 * sh.nlv + i];
            for (int l = 0; l < 8; ++l) t += ((a8[l] + acc[w][(8 + l) % LPB]) + acc[w][(16 + l) % LPB]) + acc[w][(24 + l) % LPB];
        } else {
            for (int i = 0; i < sh.ntail; ++i) t += left[w][i];
            for (int l = 0; l < 8; ++l) t += acc[w][l];
        }
        s[w] = t;
    }
    return (s[0] + k.am * s[1]) + k.as_ * s[2];
}

// Flexibility coefficients from the span sums, the three-moment system (every lane, redundantly; see
// lane_reduce / group_solve of beamopt_lanes.cuh), then the lanes rewrite {C0, C1, Qc} of the segment
// rows and lane 0 keeps a, b, p and the support moments for the displacement record.  Returns 1 on a
// bad pivot.
template <int LPB>
OPS_HD int wide_solve(const FlexBeam &fb, const WideStore &ws, int l, const SegStatics &st)
{
    const int m = fb.m;
    double a[NSPAN], b[NSPAN], c[NSPAN], p[NSPAN], q[NSPAN], dx[NSPAN];
#pragma unroll
    for (int j = 0; j < NSPAN; ++j) {
        const double R0 = ws.tot[j * NSUM], R1 = ws.tot[j * NSUM + 1], R2 = ws.tot[j * NSUM + 2];
        const double Gs = ws.tot[j * NSUM + 3], Qs = ws.tot[j * NSUM + 4];
        const double d = (j < m) ? ws.fs.span(j + 1, FlexStore::DXI) : 0.0;
        dx[j] = d;
        c[j] = (d * d) * fma(6.0, R2, fma(6.0, R1, 2.0 * R0));
        const double S = d * fma(2.0, R1, R0);
        a[j] = fma(-6.0, S, fma(6.0, R0, c[j]));
        b[j] = fma(3.0, S, -c[j]);
        q[j] = d * Qs;
        p[j] = Gs - q[j];
    }
    double MS[NSPAN + 1];
    MS[0] = 0.0;
#pragma unroll
    for (int j = 1; j <= NSPAN; ++j) MS[j] = (j == m) ? fb.Moh : 0.0;
    bool bad = false;
    double inv[NSPAN], rr[NSPAN];
    double ip = 0.0, rp = 0.0;
#pragma unroll
    for (int kk = 1; kk < NSPAN; ++kk) {
        double dd = c[kk - 1] + a[kk];
        double r_ = -(q[kk - 1] + p[kk]);
        if (kk > 1) {
            const double bk = b[kk - 1];
            const double w = bk * ip;
            dd = fma(-w, bk, dd);
            r_ = fma(-w, rp, r_);
        }
        const bool live = kk < m;
        bad = bad || (live && !(dd > 0.0));
        dd = live ? dd : 1.0;
        ip = fm::rcp64(dd);
        rp = r_;
        inv[kk] = ip; rr[kk] = r_;
    }
#pragma unroll
    for (int kk = NSPAN - 1; kk >= 1; --kk) {
        const double x = fma(-b[kk], MS[kk + 1], rr[kk]) * inv[kk];
        MS[kk] = (kk < m) ? x : MS[kk];
    }
    // per-span additions to the segment statics (none for the overhang / padding rows)
    const int ns = ws.gi[GI_NSEG];
    for (int s = l; s < ns; s += LPB) {
        const int sp = ws.gi[GI_SEGSPAN + s];
        double ml = 0.0, t1 = 0.0;
#pragma unroll
        for (int j = 0; j < NSPAN; ++j) {
            const bool hit = (j == sp) && (j < m);
            ml = hit ? MS[j] : ml;
            t1 = hit ? (MS[j + 1] - MS[j]) * dx[j] : t1;
        }
        double *r = ws.rows + s * row_doubles<LPB>();
        const bool in_regs = LPB == 32;                          // (then s = l: one trip)
        r[R_C0] = (in_regs ? st.b[0] : r[R_B0]) + ml;
        r[R_C1] = (in_regs ? st.b[1] : r[R_B1]) + t1;
        r[R_QC] = (in_regs ? st.b[2] : r[R_BQ]) + t1 * fb.invLe;
    }
    if (l == 0) {
#pragma unroll
        for (int j = 0; j < NSPAN; ++j) {
            ws.fs.span(j + 1, FlexStore::A) = a[j];
            ws.fs.span(j + 1, FlexStore::B) = b[j];
            ws.fs.span(j + 1, FlexStore::P) = p[j];
            ws.fs.ms(j) = MS[j];
        }
        ws.fs.ms(NSPAN) = MS[NSPAN];
    }
    return bad ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// once per beam: the record (SingleCore:221-249)
// ---------------------------------------------------------------------------------------------
// forces of the analysed model from the segment rows, inertias after the last step
template <int LPB>
OPS_HD void wide_emit_lane(const FlexBeam &fb, const WideShape &sh, const WideStore &ws, int l, bool fields,
                           const float *Ilast, float *shear, float *moment, float *I_out)
{
    for (int s = 0; s < sh.K; ++s) {
        const int e = LPB * s + l;
        if (e < sh.n) {
            float V = 0.0f, M = 0.0f;
            if (fields) {
                const double *r = ws.rows + (ws.seg[e] & 0x0fu) * row_doubles<LPB>();
                const double ke = (double)e - r[R_NA];
                const double Mc = fma(fma(fb.wl2h, ke, r[R_C1]), ke, r[R_C0]);
                const double Qv = fma(fb.wl, ke, r[R_QC]);
                V = (float)Qv; M = (float)(-Mc);
            }
            shear[e] = V; moment[e] = M;
            if (I_out) I_out[e] = Ilast[e];
        }
    }
}

// lane 0: displacements by integrating the curvature of the analysed inertias (flex_deflections_march)
OPS_HD void wide_emit_displacements(const BeamConsts &k, const FlexBeam &fb, const WideStore &ws, const float *Ianalysed,
                                    bool fields, double *defl, double *rot)
{
    const int nn = k.nn;
    if (!fields) {
        for (int i = 0; i < nn; ++i) { defl[i] = 0.0; rot[i] = 0.0; }
        return;
    }
    const int m = fb.m;
    for (int j = 0; j < m; ++j) {
        ws.fs.span(j + 1, FlexStore::A) *= fb.kc6;
        ws.fs.span(j + 1, FlexStore::B) *= fb.kc6;
        ws.fs.span(j + 1, FlexStore::P) *= fb.kc6;
    }
    ws.fs.ms(m) = fb.Moh;
    flex_deflections_march(k, fb, ws.fs, [&](int e) { return (double)Ianalysed[e]; }, [&](int i, double u, double th) {
        const bool z = k.zero_last_node && i == nn - 1;
        defl[i] = z ? 0.0 : u;
        rot[i] = z ? 0.0 : th;
    });
}

}  // namespace wide
}  // namespace ops
