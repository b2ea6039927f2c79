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
// in REGISTERS; the I-independent statics of the element ({M0, Q0} of the simply supported span) sit in
// shared memory, one conflict-free [slot][thread] column per lane, and the flexibility weights are
// derived from them on the fly (G = 6 M0 + 3 Le Q0 + const, g1 x1 + g2 x2 = d (G ke + g2)), which is what
// lets 40+ beams stay resident per SM.
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
// Load cases sharing one inertia vector (not in the reference): the NC cases of a beam
// run on NC adjacent groups (a "team").  Every group carries the same I, m, v and the coefficients of its
// own case; after PASS 2's forces the groups exchange M^2, V^2 through shared memory, add them in case
// order (c_0 + c_1 + ..., a fixed order) and then compute identical gradients, losses, Adam
// steps and stop decisions -- the team stays in lockstep without any further communication.
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
constexpr int TAB_SLOTS = 2 * (NSPAN + 1);      // span table, contiguous per group: Pair {MS_l, (MS_r - MS_l) d} [6], 16-byte rows (a
                                                // 128-bit access is served per quarter warp = per group: no bank conflicts between groups)
constexpr int GX_DOUBLES = 2;                   // Moh, Qoh
constexpr int GX_INTS = 6;                      // m, last, nloads, setup status, beam index of the team (lo, hi)
constexpr int GROUP_DOUBLES = FlexStore::NUM_DOUBLES + GX_DOUBLES;   // strided [slot][group] columns (+ TAB_SLOTS contiguous)
constexpr int GROUP_INTS = FlexStore::NUM_INTS + GX_INTS;

OPS_HD constexpr int lane_doubles(int epl, int nc) { return 2 * epl + SCR_SLOTS + (nc > 1 ? epl : 0); }

template <int EPL>
struct LaneRegs {
    float I[EPL], m[EPL], v[EPL], g[EPL], ke[EPL];
    unsigned long long spans;                   // 3 bits per slot k: span id of element 8 k + l
    unsigned int starts, ends;                  // bit k: slot k opens / closes a run of slots of one (real) span
};

struct alignas(16) Pair {                       // two doubles moved with one 128-bit shared access
    double x, y;
};

// Where a beam's record goes: the same row of up to MAX_DEST sets of dataset arrays (this GPU's and, for the
// in-kernel dataset gather, its peers' over NVLink).  Layouts: include/openpystruct_b200.h, ops_beamopt_launch.
constexpr int MAX_DEST = 8;
struct RecordDest {
    int nd;
    float *I[MAX_DEST];
    double *defl[MAX_DEST], *rot[MAX_DEST];
    float *shear[MAX_DEST], *moment[MAX_DEST];
    int *epochs[MAX_DEST];
    float *loss[MAX_DEST];
    int *status[MAX_DEST];
};

// per-lane shared columns, entry k at base[k * ls]
struct alignas(8) PairF {                       // {M^2, V^2} of one load case, fp32
    float c, h;
};

struct LaneStore {
    Pair *mq;                                   // [EPL]: {M0, Q0} of the element
    double *scr;                                // [SCR_SLOTS]
    PairF *xc;                                  // [EPL] (multi-case kernels only): squares exchanged between the case groups
    long ls;
};

// per-group shared data: strided columns (entry k at base[k * gs]) and the contiguous span table
struct GroupStore {
    FlexStore fs;                               // supports, loads, RA, DXI; slots A..Q hold a, b, c, p, q (no Le/6E)
    double *tab;                                // [TAB_SLOTS] contiguous, 16-byte aligned (TAB_ALIGN on the device)
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

// byte offset of the span-table row of slot kk (16 j, j = span id of the slot): one shift and one mask
OPS_HD unsigned int row_offset(unsigned long long spans, int kk)
{
    return (3 * kk >= 4) ? ((unsigned int)(spans >> (3 * kk - 4)) & 0x70u) : ((unsigned int)(spans << (4 - 3 * kk)) & 0x70u);
}
OPS_HD Pair table_row(const double *tab, unsigned int off)
{
    return *reinterpret_cast<const Pair *>(reinterpret_cast<const char *>(tab) + off);
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
    unsigned int starts = 0, ends = 0;
    int prev = -1;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        int sp = DUMMY;
        double kef = 0.0, M0 = 0.0, Q0 = 0.0;
        float I0 = 1.0f;                        // padding slots: harmless inertia, never read back
        if (e < n) {
            I0 = k.I0f;
            if (e < last) {
                int j = 0;
                for (int s = 1; s < m; ++s) j = (gs.fs.sup(s) <= e) ? s : j;
                const int na = gs.fs.sup(j);
                const double ke = (double)(e - na);
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
        if (sp != prev) {
            starts |= 1u << kk;
            if (kk > 0 && prev != DUMMY) ends |= 1u << (kk - 1);
        }
        prev = sp;
        rg.I[kk] = I0; rg.m[kk] = 0.0f; rg.v[kk] = 0.0f; rg.g[kk] = 0.0f; rg.ke[kk] = (float)kef;
        Pair mq; mq.x = M0; mq.y = Q0;
        ls.mq[(long)kk * ls.ls] = mq;
    }
    if (prev != DUMMY) ends |= 1u << (EPL - 1);
    rg.spans = spans; rg.starts = starts; rg.ends = ends;
}

// a beam rejected at set-up (mechanism / unsupported support count) still emits I_0 in its record
template <int EPL>
OPS_HD void lane_reset(const BeamConsts &k, LaneRegs<EPL> &rg)
{
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        rg.I[kk] = k.I0f; rg.m[kk] = 0.0f; rg.v[kk] = 0.0f; rg.g[kk] = 0.0f; rg.ke[kk] = 0.0f;
    }
    rg.spans = 0; rg.starts = 0; rg.ends = 0;
}

// ---------------------------------------------------------------------------------------------
// per epoch
// ---------------------------------------------------------------------------------------------
// PASS 1: flexibility sums of the lane's elements, one partial per span in the lane's scratch column.
// The five running sums restart where the lane's slots enter a new span (multiplication by an exact
// 0/1 factor instead of a branch, so the thirteen element bodies stay one basic block apart from the
// rare store of a finished partial).
struct SpanSums {
    double R0, R1, R2, G, Q;
};

// per-beam constants of the flexibility weights: with m2 = M0 + Le Q0 + w Le^2/2 (moment at the right node),
//   G  = g1 + g2 = 6 M0 + 3 Le Q0 + (3 w Le^2/2 - 2 corr)        g2 = 3 M0 + 2 Le Q0 + (w Le^2 - corr)
//   g1 x1 + g2 x2 = d (G ke + g2)                                  (x1 = ke d, x2 = x1 + d)
// so PASS 1 accumulates sum r G and sum r (G ke + g2); the factor d of the span is applied in the reduction.
struct Pass1Consts {
    double k3Le, k2Le, cG, cg2;
};
OPS_HD Pass1Consts pass1_consts(const FlexBeam &fb)
{
    Pass1Consts c;
    c.k3Le = 3.0 * fb.Le;
    c.k2Le = 2.0 * fb.Le;
    c.cG = fma(3.0, fb.wl2h, -2.0 * fb.corr);
    c.cg2 = fma(2.0, fb.wl2h, -fb.corr);
    return c;
}

template <int EPL>
OPS_HD void pass1_accumulate(const LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc, int kk, double r,
                             SpanSums &a)
{
    const double keep = ((rg.starts >> kk) & 1u) ? 0.0 : 1.0;
    const double ke = (double)rg.ke[kk];
    const Pair mq = ls.mq[(long)kk * ls.ls];
    const double G = fma(6.0, mq.x, fma(pc.k3Le, mq.y, pc.cG));
    const double g2 = fma(3.0, mq.x, fma(pc.k2Le, mq.y, pc.cg2));
    const double t = r * ke;
    a.R0 = fma(a.R0, keep, r);
    a.R1 = fma(a.R1, keep, t);
    a.R2 = fma(t, ke, a.R2 * keep);
    a.G = fma(r, G, a.G * keep);
    a.Q = fma(r, fma(G, ke, g2), a.Q * keep);
    if ((rg.ends >> kk) & 1u) {
        const int j = (int)((rg.spans >> (3 * kk)) & 7u);
        double *s = ls.scr + (long)(j * NSUM) * ls.ls;
        s[0] = a.R0; s[ls.ls] = a.R1; s[2 * ls.ls] = a.R2; s[3 * ls.ls] = a.G; s[4 * ls.ls] = a.Q;
    }
}

template <int EPL>
OPS_HD void lane_pass1(const LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc)
{
    SpanSums a = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) pass1_accumulate<EPL>(rg, ls, pc, kk, fm::rcp64((double)rg.I[kk]), a);
}

// Group reduction of the partials and the flexibility coefficients of the span: lane j < NSPAN adds
// the eight lanes' partials of span j in the fixed order j, j+1, ... (mod 8) -- deterministic and
// bank-conflict free -- and leaves a, b, c, p, q (without the Le/(6E) factor) in the span's slots
// (zeros for the unused spans j >= m, whose partials are never written).
OPS_HD void lane_reduce(int l, int m, const LaneStore &ls, const GroupStore &gs)
{
    if (l < NSPAN) {
        const double *scr0 = ls.scr - l + (long)(l * NSUM) * ls.ls;
        double R[NSUM];
#pragma unroll
        for (int i = 0; i < NSUM; ++i) {
            double s0 = scr0[(long)i * ls.ls + ((l + 0) & 7)] + scr0[(long)i * ls.ls + ((l + 1) & 7)];
            double s1 = scr0[(long)i * ls.ls + ((l + 2) & 7)] + scr0[(long)i * ls.ls + ((l + 3) & 7)];
            double s2 = scr0[(long)i * ls.ls + ((l + 4) & 7)] + scr0[(long)i * ls.ls + ((l + 5) & 7)];
            double s3 = scr0[(long)i * ls.ls + ((l + 6) & 7)] + scr0[(long)i * ls.ls + ((l + 7) & 7)];
            R[i] = (s0 + s1) + (s2 + s3);
        }
        const double d = (l < m) ? gs.fs.span(l + 1, FlexStore::DXI) : 0.0;
        const double c = (d * d) * fma(6.0, R[2], fma(6.0, R[1], 2.0 * R[0]));
        const double S = d * fma(2.0, R[1], R[0]);
        gs.fs.span(l + 1, FlexStore::A) = fma(-6.0, S, fma(6.0, R[0], c));
        gs.fs.span(l + 1, FlexStore::B) = fma(3.0, S, -c);
        gs.fs.span(l + 1, FlexStore::C) = c;
        const double q = d * R[4];
        gs.fs.span(l + 1, FlexStore::P) = R[3] - q;
        gs.fs.span(l + 1, FlexStore::Q) = q;
    }
}

// Three-moment system for the support moments (every lane, redundantly): unknowns MS[1..m-1],
// MS[0] = 0 (pinned end), MS[m] = overhang moment, Thomas elimination written without data-dependent
// control flow: rows kk >= m are computed on the zeros lane_reduce leaves for unused spans (they may
// come out Inf / NaN, feed only later unused rows in the forward sweep and are discarded by the one
// select of the back substitution); every array index is a compile-time constant, so the working set
// stays in registers.  Lane 0 publishes the per-span table PASS 2 reads (rows j >= m are never read:
// no element carries such a span id).  Returns 1 when a pivot of a real row is not positive.
OPS_HD int group_solve(const FlexBeam &fb, const GroupStore &gs, int l)
{
    const int m = fb.m;
    double a[NSPAN], b[NSPAN], c[NSPAN], p[NSPAN], q[NSPAN];
#pragma unroll
    for (int j = 0; j < NSPAN; ++j) {
        a[j] = gs.fs.span(j + 1, FlexStore::A); b[j] = gs.fs.span(j + 1, FlexStore::B);
        c[j] = gs.fs.span(j + 1, FlexStore::C); p[j] = gs.fs.span(j + 1, FlexStore::P);
        q[j] = gs.fs.span(j + 1, FlexStore::Q);
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
        bad = bad || (kk < m && !(dd > 0.0));
        ip = fm::rcp64(dd);
        rp = r_;
        inv[kk] = ip; rr[kk] = r_;
    }
    // back substitution; the known end moment MS[m] enters through the same b[kk] MS[kk + 1] term
#pragma unroll
    for (int kk = NSPAN - 1; kk >= 1; --kk) {
        const double x = fma(-b[kk], MS[kk + 1], rr[kk]) * inv[kk];
        MS[kk] = (kk < m) ? x : MS[kk];
    }
    if (l == 0) {
#pragma unroll
        for (int j = 0; j < NSPAN; ++j) {
            const double dxj = gs.fs.span(j + 1, FlexStore::DXI);
            Pair lo;
            lo.x = MS[j]; lo.y = (MS[j + 1] - MS[j]) * dxj;
            reinterpret_cast<Pair *>(gs.tab)[j] = lo;
        }
        // the DUMMY row (overhang, padding) stays zero: written once per beam by group_table_init
    }
    return bad ? 1 : 0;
}

OPS_HD void group_table_init(const GroupStore &gs)
{
    for (int i = 0; i < TAB_SLOTS; ++i) gs.tab[i] = 0.0;
}

// bending moment (three-moment sign: sagging positive) and shear at the node-i end of slot kk
template <int EPL>
OPS_HD void element_forces(const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe, int kk,
                           double &Mc, double &Qv)
{
    const Pair lo = table_row(gs.tab, row_offset(rg.spans, kk));
    const Pair mq = ls.mq[(long)kk * ls.ls];
    Mc = fma(lo.y, (double)rg.ke[kk], mq.x + lo.x);
    Qv = fma(lo.y, invLe, mq.y);
}

// Elements are processed in batches of NB slots, STAGE BY STAGE across the batch: the stages of one
// element are a ~150-cycle dependent chain (MUFU -> Newton -> quotient -> ...), and written element
// by element ptxas leaves the chains serial at this register budget; stage-major order puts NB
// independent instructions between dependent ones.
#ifndef OPS_LANES_NB
#define OPS_LANES_NB 5
#endif
constexpr int NB = OPS_LANES_NB;

// multi-case kernels, before PASS 2: M^2, V^2 of this group's load case into the exchange column
template <int EPL>
OPS_HD void lane_case_squares(const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe)
{
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        double Mc, Qv;
        element_forces<EPL>(rg, ls, gs, invLe, kk, Mc, Qv);
        const float Mf = (float)Mc, Vf = (float)Qv;
        PairF x;
        x.c = Mf * Mf; x.h = Vf * Vf;
        ls.xc[(long)kk * ls.ls] = x;
    }
}

// PASS 2: end forces, loss terms d, q and autograd's gradient with M, V constant (element_update_f32,
// first half; the gradient is kept in rg.g), torch.sum partials of sum I, sum d, sum q.
// NC > 1: the squares come from the exchange columns of the team (case_id = this group's case).
// fp32 ranges for the branch-free division / square root: I in [clamp_min, 1e20) (the clamp,
// SingleCore:208), c = M^2 and h = V^2 zero or >= 2^-100.
template <int EPL, int NC, int NBX = NB>
OPS_HD void lane_forces(const BeamConsts &k, int n, LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe,
                        int l, int case_id)
{
    constexpr int NB = NBX;                     // slots per batch of this instance (shadows the default)
    const SumShape sh = sum_shape(n);
    float aI[4] = {0.0f, 0.0f, 0.0f, 0.0f}, ad[4] = {0.0f, 0.0f, 0.0f, 0.0f}, aq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float tI = 0.0f, td = 0.0f, tq = 0.0f;
    const PairF *x0 = ls.xc - (long)case_id * LPB;             // this lane's column in the team's case-0 group
    // one operation per stage across the batch (OPS_B: "for every live slot i of the batch"); the fast-path
    // sequences of fastmath.cuh are written out step by step so that consecutive instructions are independent
#define OPS_B _Pragma("unroll") for (int i = 0; i < NB; ++i) if (k0 + i < EPL)
#pragma unroll
    for (int k0 = 0; k0 < EPL; k0 += NB) {
        float I[NB], c[NB], h[NB], b[NB], y[NB], rb[NB], s[NB], gg[NB], rgg[NB], rs[NB], d[NB], db[NB], q[NB], qg[NB];
        float t0[NB], t1[NB], t2[NB];
        OPS_B I[i] = rg.I[k0 + i];
        if (NC == 1) {
            Pair lo[NB], mq[NB];
            double ke[NB], Mc[NB], Qv[NB];
            OPS_B {
                lo[i] = table_row(gs.tab, row_offset(rg.spans, k0 + i));
                mq[i] = ls.mq[(long)(k0 + i) * ls.ls];
                ke[i] = (double)rg.ke[k0 + i];
            }
            OPS_B { Mc[i] = mq[i].x + lo[i].x; Qv[i] = fma(lo[i].y, invLe, mq[i].y); }
            OPS_B Mc[i] = fma(lo[i].y, ke[i], Mc[i]);
            OPS_B { c[i] = (float)Mc[i]; h[i] = (float)Qv[i]; }
            OPS_B { c[i] = c[i] * c[i]; h[i] = h[i] * h[i]; }
        } else {
            OPS_B {
                PairF x = x0[(long)(k0 + i) * ls.ls];
                c[i] = x.c; h[i] = x.h;
#pragma unroll
                for (int cc = 1; cc < NC; ++cc) {
                    x = x0[(long)(k0 + i) * ls.ls + cc * LPB];
                    c[i] += x.c; h[i] += x.h;
                }
            }
        }
        OPS_B { b[i] = k.E2 * I[i]; y[i] = fm::rsq_a(I[i]); }
        OPS_B b[i] = b[i] + k.epsf;
        // rb = refined 1 / b ; s = sqrt(I)
        OPS_B { rb[i] = fm::rcp_a(b[i]); t0[i] = I[i] * y[i]; y[i] = y[i] * 0.5f; }
#if defined(__CUDA_ARCH__)
        OPS_B { t1[i] = fmaf(-b[i], rb[i], 1.0f); t2[i] = fmaf(-t0[i], t0[i], I[i]); }
        OPS_B { rb[i] = fmaf(rb[i], t1[i], rb[i]); s[i] = fmaf(t2[i], y[i], t0[i]); }
        // d = c / b ; gg = Gf (kf s) ; rs = 1 / s
        OPS_B { t0[i] = c[i] * rb[i]; gg[i] = k.kf * s[i]; rs[i] = fm::rcp_a(s[i]); }
        OPS_B { t1[i] = fmaf(-b[i], t0[i], c[i]); gg[i] = k.Gf * gg[i]; t2[i] = fmaf(-s[i], rs[i], 1.0f); }
        OPS_B { d[i] = fmaf(rb[i], t1[i], t0[i]); rgg[i] = fm::rcp_a(gg[i]); rs[i] = fmaf(rs[i], t2[i], rs[i]); }
        // db = d / b ; rgg = refined 1 / gg ; rs = RN(1 / s)
        OPS_B { t0[i] = d[i] * rb[i]; t1[i] = fmaf(-gg[i], rgg[i], 1.0f); t2[i] = fmaf(-s[i], rs[i], 1.0f); }
        OPS_B { db[i] = fmaf(-b[i], t0[i], d[i]); rgg[i] = fmaf(rgg[i], t1[i], rgg[i]); rs[i] = fmaf(rs[i], t2[i], rs[i]); }
        OPS_B { db[i] = fmaf(rb[i], db[i], t0[i]); t1[i] = h[i] * rgg[i]; rs[i] = 0.5f * rs[i]; }
        // q = h / gg ; bending branch gb = ((-am) db) E2
        OPS_B { db[i] = (-k.am) * db[i]; t2[i] = fmaf(-gg[i], t1[i], h[i]); }
        OPS_B { db[i] = db[i] * k.E2; q[i] = fmaf(rgg[i], t2[i], t1[i]); }
        // qg = q / gg
        OPS_B t0[i] = q[i] * rgg[i];
        OPS_B t1[i] = fmaf(-gg[i], t0[i], q[i]);
        OPS_B qg[i] = fmaf(rgg[i], t1[i], t0[i]);
#else
        OPS_B {
            s[i] = sqrtf(I[i]);
            d[i] = c[i] / b[i];
            db[i] = ((-k.am) * (d[i] / b[i])) * k.E2;
            gg[i] = k.Gf * (k.kf * s[i]);
            q[i] = h[i] / gg[i];
            qg[i] = q[i] / gg[i];
            rs[i] = 0.5f * (1.0f / s[i]);
        }
        (void)rb; (void)rgg; (void)t0; (void)t1; (void)t2; (void)y;
#endif
        // shear branch gs = ((((-as) qg) Gf) kf) (0.5 / s) ; g = (1 + gs) + gb
        OPS_B qg[i] = (-k.as_) * qg[i];
        OPS_B qg[i] = qg[i] * k.Gf;
        OPS_B qg[i] = qg[i] * k.kf;
        OPS_B qg[i] = qg[i] * rs[i];
        OPS_B qg[i] = 1.0f + qg[i];
        OPS_B rg.g[k0 + i] = qg[i] + db[i];
        OPS_B {
            const int kk = k0 + i;
            if (kk < sh.blk) {
                aI[kk & 3] += I[i]; ad[kk & 3] += d[i]; aq[kk & 3] += q[i];
            } else if (kk < sh.vec) {
                aI[0] += I[i]; ad[0] += d[i]; aq[0] += q[i];
            } else if (kk == sh.vec && l < sh.ntail) {
                tI = I[i]; td = d[i]; tq = q[i];
            }
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

// Adam step + clamp on the lane's elements (element_update_f32, second half) and, fused behind it
// when PASS1 is set, PASS 1 of the NEXT epoch on the updated inertias (stage-major batches as above).
// The fast square root needs v >= 2^-101; v is an EMA of g^2, so anything smaller means g vanished on
// every epoch so far -- tested once per lane and epoch, with the generic operators as the (cold)
// alternative.
template <int EPL, bool PASS1, int NBX = NB>
OPS_HD void lane_adam(const BeamConsts &k, LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc,
                      float neg_step, float bc2_sqrt)
{
    constexpr int NB = NBX;
    bool rare = false;
    {
        float t0[EPL], t1[EPL];
#define OPS_A _Pragma("unroll") for (int kk = 0; kk < EPL; ++kk)
        OPS_A { t0[kk] = rg.g[kk] - rg.m[kk]; t1[kk] = k.omb2f * rg.g[kk]; rg.v[kk] = rg.v[kk] * k.b2f; }
        OPS_A { rg.m[kk] = fmaf(k.w1, t0[kk], rg.m[kk]); rg.v[kk] = fmaf(t1[kk], rg.g[kk], rg.v[kk]); }
        float vmin = rg.v[0];
        OPS_A vmin = fminf(vmin, rg.v[kk]);                         // (v is never NaN here: the loss was finite)
        rare = !(vmin >= fm::SQRT_F_MIN);
#undef OPS_A
    }
    if (!rare) {
        const float rbc = fm::rcp_r(bc2_sqrt);
        SpanSums a = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k0 = 0; k0 < EPL; k0 += NB) {
            float y[NB], den[NB], rd[NB], t0[NB], t1[NB], num[NB];
            double Id[NB], r[NB];
#if defined(__CUDA_ARCH__)
            // sqrt(v) / bc2_sqrt + eps
            OPS_B y[i] = fm::rsq_a(rg.v[k0 + i]);
            OPS_B { t0[i] = rg.v[k0 + i] * y[i]; y[i] = y[i] * 0.5f; num[i] = neg_step * rg.m[k0 + i]; }
            OPS_B t1[i] = fmaf(-t0[i], t0[i], rg.v[k0 + i]);
            OPS_B t0[i] = fmaf(t1[i], y[i], t0[i]);                  // sqrt(v)
            OPS_B t1[i] = t0[i] * rbc;
            OPS_B den[i] = fmaf(-bc2_sqrt, t1[i], t0[i]);
            OPS_B den[i] = fmaf(rbc, den[i], t1[i]);
            OPS_B den[i] = den[i] + k.adam_epsf;
            // I + (neg_step m) / denom, clamp
            OPS_B rd[i] = fm::rcp_a(den[i]);
            OPS_B t0[i] = fmaf(-den[i], rd[i], 1.0f);
            OPS_B rd[i] = fmaf(rd[i], t0[i], rd[i]);
            OPS_B t0[i] = num[i] * rd[i];
            OPS_B t1[i] = fmaf(-den[i], t0[i], num[i]);
            OPS_B t0[i] = fmaf(rd[i], t1[i], t0[i]);
            OPS_B t0[i] = rg.I[k0 + i] + t0[i];
            OPS_B rg.I[k0 + i] = fmaxf(t0[i], k.clampf);           // one FMNMX (t0 is finite: den > 0, v and m finite)
            if (PASS1) {
                double e[NB];
                OPS_B Id[i] = (double)rg.I[k0 + i];
                OPS_B r[i] = fm::rcp64_a(Id[i]);
                OPS_B e[i] = fma(-Id[i], r[i], 1.0);           // rcp64_n, stage by stage
                OPS_B e[i] = fma(e[i], e[i], e[i]);
                OPS_B r[i] = fma(r[i], e[i], r[i]);
            }
#else
            OPS_B {
                den[i] = sqrtf(rg.v[k0 + i]) / bc2_sqrt + k.adam_epsf;
                const float x = rg.I[k0 + i] + (neg_step * rg.m[k0 + i]) / den[i];
                rg.I[k0 + i] = x < k.clampf ? k.clampf : x;
                Id[i] = (double)rg.I[k0 + i];
                r[i] = 1.0 / Id[i];
            }
            (void)y; (void)rd; (void)t0; (void)t1; (void)num; (void)rbc;
#endif
            if (PASS1) {
                OPS_B pass1_accumulate<EPL>(rg, ls, pc, k0 + i, r[i], a);
            }
        }
    } else {
#pragma unroll                                  // (static indices: a rolled loop would push the state arrays to local memory)
        for (int kk = 0; kk < EPL; ++kk) {
            const float denom = sqrtf(rg.v[kk]) / bc2_sqrt + k.adam_epsf;
            const float x = rg.I[kk] + (neg_step * rg.m[kk]) / denom;
            rg.I[kk] = x < k.clampf ? k.clampf : x;
        }
        if (PASS1) lane_pass1<EPL>(rg, ls, pc);
    }
}
#undef OPS_B

// ---------------------------------------------------------------------------------------------
// once per beam: the record (SingleCore:221-249).  M, V, u, theta belong to the LAST ANALYSED
// inertias, i.e. the ones still in rg.I when the stop decision is taken (before lane_adam).
// ---------------------------------------------------------------------------------------------
template <int EPL>
OPS_HD void lane_emit_forces(int n, const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe, int l,
                             bool fields, float *shear, float *moment)
{
    float *stage = reinterpret_cast<float *>(ls.scr);        // I of slot kk at float index (kk >> 1) * 2 ls + (kk & 1)
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        stage[(long)(kk >> 1) * 2 * ls.ls + (kk & 1)] = rg.I[kk];
        if (e < n) {
            double Mc = 0.0, Qv = 0.0;
            if (fields) element_forces<EPL>(rg, ls, gs, invLe, kk, Mc, Qv);
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
        gs.fs.span(j + 1, FlexStore::A) *= fb.kc6;
        gs.fs.span(j + 1, FlexStore::B) *= fb.kc6;
        gs.fs.span(j + 1, FlexStore::P) *= fb.kc6;
        gs.fs.ms(j) = gs.tab[2 * j];
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

// In-kernel dataset gather: the group copies the record it has just written (destination 0 = this GPU's dataset
// arrays) into the same rows of the other destinations, the peers' arrays over NVLink, eight lanes wide: 16-byte
// units (128 contiguous bytes per group and step) where the row is 16-byte aligned, 8- or 4-byte units otherwise.
// Deflections and rotations were written by lane 0 and the other rows by all lanes, so the caller puts a group
// barrier in front.
struct alignas(16) Bytes16 {
    unsigned long long a, b;
};
OPS_HD void lane_copy_bytes(void *dst, const void *src, long bytes, int l)
{
    const bool al16 = ((((unsigned long long)dst) | ((unsigned long long)src)) & 15ull) == 0;
    const bool al8 = ((((unsigned long long)dst) | ((unsigned long long)src)) & 7ull) == 0;
    long done = 0;
    if (al16) {
        const long n16 = bytes >> 4;
        for (long i = l; i < n16; i += LPB) reinterpret_cast<Bytes16 *>(dst)[i] = reinterpret_cast<const Bytes16 *>(src)[i];
        done = n16 << 4;
    } else if (al8) {
        const long n8 = bytes >> 3;
        for (long i = l; i < n8; i += LPB)
            reinterpret_cast<unsigned long long *>(dst)[i] = reinterpret_cast<const unsigned long long *>(src)[i];
        done = n8 << 3;
    }
    const long n4 = (bytes - done) >> 2;                    // every row is a multiple of 4 bytes
    for (long i = l; i < n4; i += LPB)
        reinterpret_cast<unsigned int *>(static_cast<char *>(dst) + done)[i] =
            reinterpret_cast<const unsigned int *>(static_cast<const char *>(src) + done)[i];
}

OPS_HD void lane_copy_record(int n, int nn, int l, const RecordDest &dst, long long row, long long rowc, bool first_case)
{
    for (int r = 1; r < dst.nd; ++r) {
        lane_copy_bytes(dst.shear[r] + rowc * n, dst.shear[0] + rowc * n, 4L * n, l);
        lane_copy_bytes(dst.moment[r] + rowc * n, dst.moment[0] + rowc * n, 4L * n, l);
        lane_copy_bytes(dst.defl[r] + rowc * nn, dst.defl[0] + rowc * nn, 8L * nn, l);
        lane_copy_bytes(dst.rot[r] + rowc * nn, dst.rot[0] + rowc * nn, 8L * nn, l);
        if (first_case) {
            lane_copy_bytes(dst.I[r] + row * n, dst.I[0] + row * n, 4L * n, l);
            if (l == 0) {
                dst.epochs[r][row] = dst.epochs[0][row];
                dst.loss[r][row] = dst.loss[0][row];
                dst.status[r][row] = dst.status[0][row];
            }
        }
    }
}

}  // namespace lanes
}  // namespace ops
