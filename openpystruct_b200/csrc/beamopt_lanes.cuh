// Production iteration: three-moment solve with EIGHT LANES PER BEAM and the optimiser state in registers.
//
// Replaces, per epoch (reference file:line): setup_model + analyze + eleResponse
// (OpenPyStruct_BeamOpt_training_SingleCore.py:176-190), the loss (:195-199), backward / Adam /
// ExponentialLR / clamp (:202-208) and the early-stop test (:211-219).
//
// Mapping.  Lane l of a beam's 8-lane group owns the elements e = 8 k + l, k = 0..EPL-1: lane l is
// SIMD lane l of the 8-float vectors ATen's cascade_sum works on, so torch.sum's summation order
// (beamopt_core.cuh, torch_sum_f32) falls out of per-lane running sums plus one fixed-order combine.
// Per element the lane keeps I, Adam's m and v in REGISTERS -- as PAIRS of consecutive slots (2 j, 2 j + 1), the
// operands of sm_100a's packed fp32 instructions (fastmath.cuh) -- plus one byte for the element's index inside
// its span; the I-independent statics of the element ({M0, Q0} of the simply supported span) sit in
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
// sums over its own elements, the group reduces them through shared memory in a fixed
// order, every lane solves the <= 4-unknown tridiagonal system redundantly, and ONE PASS over the lane's
// elements (lane_pass) evaluates
//     Mc_e = M0_e + MS_l + (MS_r - MS_l) d ke ,   V_e = Q0_e + (MS_r - MS_l) d / Le
// followed by the fp32 loss terms, the frozen-M,V gradient, the Adam update of its elements and the five sums of
// the NEXT epoch on the updated inertias -- {M0, Q0} and ke are fetched once per element and epoch, and neither
// the gradient nor anything else per element survives the pass.
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
// Stride of the [slot][group] columns (register / shared-memory instances).  The five lanes that publish a, b, c, p, q of
// their span (lane_reduce) write the slots 7 j + X: with a stride of G = 40 or 48 groups they fall on two banks (6
// wavefronts per store instead of 2; ncu, profiles/r02_ncu_summary.md); a stride = 12 (mod 16) puts consecutive spans
// 8 banks apart (measured: +1.1 % on many-round batches and on the 8-case teams, profiles/r02_group_stride_ab.txt).
// Build knob OPS_LANES_NO_GSPAD (A/B).
OPS_HD constexpr int group_stride(int G)
{
#ifdef OPS_LANES_NO_GSPAD
    return G;
#else
    return G + (((12 - G) % 16) + 16) % 16;
#endif
}

// multi-case teams: pair j of the lane's slots is OWNED by group j % NC (its fp32 chain, Adam state and exchange rows)
OPS_HD constexpr int team_owned_pairs(int epl, int nc) { return ((epl + 1) / 2 + nc - 1) / nc; }
constexpr int XB_ROWS = 4;                      // exchange rows of an owned pair: {I_old, d, q, I_new}, one fp32 pair each
OPS_HD constexpr int lane_doubles(int epl, int nc)
{
    return 2 * epl + SCR_SLOTS + (nc > 1 ? epl + XB_ROWS * team_owned_pairs(epl, nc) : 0);
}
// shared memory of a CTA of G groups (lanes_plan and the kernel's carve-up agree on this)
OPS_HD constexpr size_t cta_smem_bytes(int G, int epl, int num_cases)
{
    return (size_t)G * ((size_t)LPB * 8 * lane_doubles(epl, num_cases) + (size_t)TAB_SLOTS * 8) +
           (size_t)group_stride(G) * ((size_t)GROUP_DOUBLES * 8 + (size_t)GROUP_INTS * 4);
}

template <int EPL>
struct LaneRegs {
    static constexpr int NP = (EPL + 1) / 2;    // slot pairs; an odd EPL leaves one padding slot (kk = EPL)
    fm::F2 I[NP], m[NP], v[NP];                 // pair j = slots 2 j, 2 j + 1
    unsigned int ke[(2 * NP + 3) / 4];          // one byte per slot: index of the element inside its span (<= 167)
    unsigned long long spans;                   // 3 bits per slot k < EPL: span id of element 8 k + l
    unsigned int starts, ends;                  // bit k: slot k opens / closes a run of slots of one (real) span
};

// element index inside its span of slot kk as a double (I2F.F64.U8 with a byte selector: one conversion)
// Conversions of the pass WITHOUT the XU pipe (I2F / F2F are 8 issue cycles per warp there, the busiest pipe of the
// kernel; build knob OPS_LANES_XUTRIM, exact either way): a small unsigned integer as 2^52 + k minus 2^52 (one DADD), a
// positive normal float widened by re-biasing the exponent with integer instructions.
OPS_HD double small_uint_to_double(unsigned int k)
{
#if defined(__CUDA_ARCH__) && defined(OPS_LANES_XUTRIM)
    return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0;
#else
    return (double)k;
#endif
}
OPS_HD double pos_normal_float_to_double(float f)
{
#if defined(__CUDA_ARCH__) && defined(OPS_LANES_XUTRIM)
    const unsigned int b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + (896u << 20)), (int)(b << 29));
#else
    return (double)f;
#endif
}

template <int EPL>
OPS_HD double slot_ke(const LaneRegs<EPL> &rg, int kk)
{
    return small_uint_to_double((rg.ke[kk >> 2] >> (8 * (kk & 3))) & 0xffu);
}
template <int EPL>
OPS_HD float &slot_ref(fm::F2 (&a)[LaneRegs<EPL>::NP], int kk) { return (kk & 1) ? a[kk >> 1].y : a[kk >> 1].x; }

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
    double *xb;                                 // [XB_ROWS * owned pairs] (multi-case): the owner's rows of its pairs
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

// ---------------------------------------------------------------------------------------------
// Where a lane keeps what only IT ever touches: the statics {M0, Q0} of its elements and Adam's m, v.
//
// HomeShared (every instance but the tensor-memory one): {M0, Q0} in the lane's shared-memory column, m and v in
// LaneRegs.  HomeTm (beamopt_lanes_tm.cu): both in TENSOR MEMORY -- the 256 KB per SM next to the tensor cores that
// this path otherwise leaves idle -- through tcgen05.st / tcgen05.ld (32x32b: thread t of a warp owns TMEM lane
// 32 (warp % 4) + t, the warps sharing a lane quarter own disjoint column blocks).  That takes 1.7 KB per beam out of
// shared memory and 28 registers out of every thread, which is what lets 17 warps = 68 beams stay resident per SM
// instead of 10.  Tensor-memory instructions are WARP-COLLECTIVE (.sync.aligned): every function that
// touches a HomeTm is called by all 32 lanes of a warp, and the lanes whose group has nothing to do compute on whatever
// their columns hold and keep the results to themselves (`active` arguments below).
// Column layout of a lane (32-bit words): [4 kk, 4 kk + 4) = {M0, Q0} of slot kk (kk < 2 NP), then
// [8 NP + 4 j, 8 NP + 4 j + 4) = {m.x, m.y, v.x, v.y} of pair j.
// Host build (tests/hostsim): the same layout in a plain array per lane.
// ---------------------------------------------------------------------------------------------
OPS_HD void words_of(double d, unsigned int &lo, unsigned int &hi)
{
#if defined(__CUDA_ARCH__)
    lo = (unsigned int)__double2loint(d); hi = (unsigned int)__double2hiint(d);
#else
    unsigned long long u; memcpy(&u, &d, 8); lo = (unsigned int)u; hi = (unsigned int)(u >> 32);
#endif
}
OPS_HD double double_of(unsigned int lo, unsigned int hi)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    const unsigned long long u = ((unsigned long long)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d;
#endif
}
OPS_HD unsigned int word_of(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned int u; memcpy(&u, &f, 4); return u;
#endif
}
OPS_HD float float_of(unsigned int u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

struct HomeShared {
    static constexpr bool TM = false;
    OPS_HD void wait_loads() const {}
    OPS_HD void wait_stores() const {}
    // {M0, Q0} of slots k0 .. k0 + N - 1 (those < EPL)
    template <int EPL, int N>
    OPS_HD void fetch_mq(const LaneStore &ls, int k0, Pair (&mq)[N]) const
    {
#pragma unroll
        for (int s = 0; s < N; ++s) if (k0 + s < EPL) mq[s] = ls.mq[(long)(k0 + s) * ls.ls];
    }
    template <int EPL, int N>
    OPS_HD void fetch_mv(const LaneRegs<EPL> &rg, int p0, fm::F2 (&m)[N], fm::F2 (&v)[N]) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) if (p0 + i < LaneRegs<EPL>::NP) { m[i] = rg.m[p0 + i]; v[i] = rg.v[p0 + i]; }
    }
    template <int EPL, int N>
    OPS_HD void put_mv(LaneRegs<EPL> &rg, int p0, const fm::F2 (&m)[N], const fm::F2 (&v)[N]) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) if (p0 + i < LaneRegs<EPL>::NP) { rg.m[p0 + i] = m[i]; rg.v[p0 + i] = v[i]; }
    }
    // lane_init: {M0, Q0} of slot kk
    OPS_HD void stage_mq(const LaneStore &ls, int kk, const Pair &mq) const { ls.mq[(long)kk * ls.ls] = mq; }
};

struct HomeTm {
    static constexpr bool TM = true;
    unsigned int base;                          // device: tensor-memory address of the lane's first column (lane quarter of the warp in bits 16+)
    unsigned int *w;                            // host simulator: the lane's words
    template <int N>
    OPS_HD void ld(int col, unsigned int (&r)[N]) const
    {
        static_assert(N == 4 || N == 8 || N == 16, "tcgen05.ld.32x32b.xN");
#if defined(__CUDA_ARCH__)
        const unsigned int a = base + (unsigned int)col;
        if constexpr (N == 4)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
        else if constexpr (N == 8)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
        else
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a));
#else
        for (int i = 0; i < N; ++i) r[i] = w[col + i];
#endif
    }
    template <int N>
    OPS_HD void st(int col, const unsigned int (&r)[N]) const
    {
        static_assert(N == 4 || N == 8, "tcgen05.st.32x32b.xN");
#if defined(__CUDA_ARCH__)
        const unsigned int a = base + (unsigned int)col;
        if constexpr (N == 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0, %1, %2, %3};"
                         ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(a) : "memory");
        else
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
                         ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(a) : "memory");
#else
        for (int i = 0; i < N; ++i) w[col + i] = r[i];
#endif
    }
    OPS_HD void wait_loads() const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
    }
    OPS_HD void wait_stores() const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
    }
    // words [c0, c0 + 4 NQ) as NQ quads, with the widest loads
    template <int NQ>
    OPS_HD void ld_quads(int c0, unsigned int (&q)[NQ][4]) const
    {
        static_assert(NQ >= 1 && NQ <= 6, "batch size");
        if constexpr (NQ >= 4) {
            unsigned int r[16];
            ld<16>(c0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) q[i >> 2][i & 3] = r[i];
            if constexpr (NQ == 5) ld<4>(c0 + 16, q[4]);
            if constexpr (NQ == 6) {
                unsigned int t[8];
                ld<8>(c0 + 16, t);
#pragma unroll
                for (int i = 0; i < 8; ++i) q[4 + (i >> 2)][i & 3] = t[i];
            }
        } else if constexpr (NQ >= 2) {
            unsigned int r[8];
            ld<8>(c0, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i >> 2][i & 3] = r[i];
            if constexpr (NQ == 3) ld<4>(c0 + 8, q[2]);
        } else {
            ld<4>(c0, q[0]);
        }
    }
    template <int EPL, int N>
    OPS_HD void fetch_mq(const LaneStore &, int k0, Pair (&mq)[N]) const
    {
        unsigned int q[N][4];
        ld_quads<N>(4 * k0, q);                 // (slots up to 2 NP - 1 have columns; the padding slot reads zeros)
        wait_loads();
#pragma unroll
        for (int s = 0; s < N; ++s) { mq[s].x = double_of(q[s][0], q[s][1]); mq[s].y = double_of(q[s][2], q[s][3]); }
    }
    template <int EPL, int N>
    OPS_HD void fetch_mv(const LaneRegs<EPL> &, int p0, fm::F2 (&m)[N], fm::F2 (&v)[N]) const
    {
        unsigned int q[N][4];
        ld_quads<N>(8 * LaneRegs<EPL>::NP + 4 * p0, q);
        wait_loads();
#pragma unroll
        for (int i = 0; i < N; ++i) { m[i] = fm::f2(float_of(q[i][0]), float_of(q[i][1])); v[i] = fm::f2(float_of(q[i][2]), float_of(q[i][3])); }
    }
    template <int EPL, int N>
    OPS_HD void put_mv(LaneRegs<EPL> &, int p0, const fm::F2 (&m)[N], const fm::F2 (&v)[N]) const
    {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (p0 + i < LaneRegs<EPL>::NP) {
                const unsigned int r[4] = {word_of(m[i].x), word_of(m[i].y), word_of(v[i].x), word_of(v[i].y)};
                st<4>(8 * LaneRegs<EPL>::NP + 4 * (p0 + i), r);
            }
        }
    }
    // lane_init stages {M0, Q0} in the lane's scratch column (slots 2 kk, 2 kk + 1); tm_commit moves them over
    OPS_HD void stage_mq(const LaneStore &ls, int kk, const Pair &mq) const
    {
        ls.scr[(long)(2 * kk) * ls.ls] = mq.x; ls.scr[(long)(2 * kk + 1) * ls.ls] = mq.y;
    }
    static OPS_HD constexpr int columns(int epl) { return 12 * ((epl + 1) / 2); }
};

// New beams of some groups of the warp (`fresh` lanes; called by ALL lanes): the staged {M0, Q0} and m = v = 0 go to
// tensor memory; a store writes all 32 lanes, so the other lanes write back what they hold (read-modify-write); the
// fresh lanes' scratch columns are cleared for the partial sums.
template <int EPL>
OPS_HD void tm_commit(const HomeTm &hm, const LaneStore &ls, bool fresh)
{
    constexpr int NP = LaneRegs<EPL>::NP;
    hm.wait_stores();
#pragma unroll
    for (int kk = 0; kk < 2 * NP; ++kk) {
        unsigned int r[4];
        hm.ld<4>(4 * kk, r);
        hm.wait_loads();
        if (fresh) {
            double M0 = 0.0, Q0 = 0.0;
            if (kk < EPL) { M0 = ls.scr[(long)(2 * kk) * ls.ls]; Q0 = ls.scr[(long)(2 * kk + 1) * ls.ls]; }
            words_of(M0, r[0], r[1]); words_of(Q0, r[2], r[3]);
        }
        hm.st<4>(4 * kk, r);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        unsigned int r[4];
        hm.ld<4>(8 * NP + 4 * j, r);
        hm.wait_loads();
        if (fresh) { r[0] = 0u; r[1] = 0u; r[2] = 0u; r[3] = 0u; }
        hm.st<4>(8 * NP + 4 * j, r);
    }
    hm.wait_stores();
    if (fresh) {
        for (int s = 0; s < SCR_SLOTS; ++s) ls.scr[(long)s * ls.ls] = 0.0;
    }
}

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

template <int EPL, class HM = HomeShared>
OPS_HD void lane_init(const BeamConsts &k, int n, const FlexBeam &fb, const GroupStore &gs, const LaneStore &ls,
                      int l, LaneRegs<EPL> &rg, const HM &hm = HM())
{
    constexpr int NP = LaneRegs<EPL>::NP;
    const int m = fb.m, last = fb.last, nl = fb.nloads;
    for (int s = 0; s < SCR_SLOTS; ++s) ls.scr[(long)s * ls.ls] = 0.0;
    unsigned long long spans = 0;
    unsigned int starts = 0, ends = 0;
    int prev = -1;
#pragma unroll
    for (int w = 0; w < (2 * NP + 3) / 4; ++w) rg.ke[w] = 0u;
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        const int e = LPB * kk + l;
        int sp = DUMMY;
        int kei = 0;
        double M0 = 0.0, Q0 = 0.0;
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
                sp = j; kei = e - na; M0 = Ms; Q0 = Qs;
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
        slot_ref<EPL>(rg.I, kk) = I0; slot_ref<EPL>(rg.m, kk) = 0.0f; slot_ref<EPL>(rg.v, kk) = 0.0f;
        rg.ke[kk >> 2] |= (unsigned int)kei << (8 * (kk & 3));
        Pair mq; mq.x = M0; mq.y = Q0;
        hm.stage_mq(ls, kk, mq);
    }
    if (EPL & 1) { rg.I[NP - 1].y = 1.0f; rg.m[NP - 1].y = 0.0f; rg.v[NP - 1].y = 0.0f; }     // the pair's padding half
    if (prev != DUMMY) ends |= 1u << (EPL - 1);
    rg.spans = spans; rg.starts = starts; rg.ends = ends;
}

// a beam rejected at set-up (mechanism / unsupported support count) still emits I_0 in its record
template <int EPL>
OPS_HD void lane_reset(const BeamConsts &k, LaneRegs<EPL> &rg)
{
    constexpr int NP = LaneRegs<EPL>::NP;
#pragma unroll
    for (int j = 0; j < NP; ++j) { rg.I[j] = fm::splat(k.I0f); rg.m[j] = fm::splat(0.0f); rg.v[j] = fm::splat(0.0f); }
#pragma unroll
    for (int w = 0; w < (2 * NP + 3) / 4; ++w) rg.ke[w] = 0u;
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

// one element's contribution to the running sums; `store`: finished partials go to the lane's scratch column
template <int EPL>
OPS_HD void pass1_accumulate(const LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc, int kk, double r,
                             double ke, const Pair &mq, bool store, SpanSums &a)
{
    const double keep = ((rg.starts >> kk) & 1u) ? 0.0 : 1.0;
    const double G = fma(6.0, mq.x, fma(pc.k3Le, mq.y, pc.cG));
    const double g2 = fma(3.0, mq.x, fma(pc.k2Le, mq.y, pc.cg2));
    const double t = r * ke;
    a.R0 = fma(a.R0, keep, r);
    a.R1 = fma(a.R1, keep, t);
    a.R2 = fma(t, ke, a.R2 * keep);
    a.G = fma(r, G, a.G * keep);
    a.Q = fma(r, fma(G, ke, g2), a.Q * keep);
    if (((rg.ends >> kk) & 1u) && store) {
        const int j = (int)((rg.spans >> (3 * kk)) & 7u);
        double *s = ls.scr + (long)(j * NSUM) * ls.ls;
        s[0] = a.R0; s[ls.ls] = a.R1; s[2 * ls.ls] = a.R2; s[3 * ls.ls] = a.G; s[4 * ls.ls] = a.Q;
    }
}

// the sums on their own: first epoch of a beam (parked = false), and after an epoch whose pass parked the inertias
// in the first slots of the scratch column instead (those slots are cleared first: a lane only ever writes the
// partials of the spans it touches, the others must read zero).  HomeTm: called by the whole warp, the lanes that
// are not `active` leave their scratch column alone.
template <int EPL, class HM = HomeShared>
OPS_HD void lane_pass1(const LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc, bool parked, const HM &hm = HM(),
                       bool active = true)
{
    if (parked && active) {
#pragma unroll
        for (int j = 0; j < LaneRegs<EPL>::NP; ++j) ls.scr[(long)j * ls.ls] = 0.0;
    }
    SpanSums a = {0.0, 0.0, 0.0, 0.0, 0.0};
    constexpr int SB = HM::TM ? 4 : 1;          // slots per fetch
#pragma unroll
    for (int k0 = 0; k0 < EPL; k0 += SB) {
        Pair mq[SB];
        hm.template fetch_mq<EPL, SB>(ls, k0, mq);
#pragma unroll
        for (int s_ = 0; s_ < SB; ++s_) {
            const int kk = k0 + s_;
            if (kk < EPL) {
                const float If = (kk & 1) ? rg.I[kk >> 1].y : rg.I[kk >> 1].x;
                pass1_accumulate<EPL>(rg, ls, pc, kk, fm::rcp64((double)If), slot_ke<EPL>(rg, kk), mq[s_], active, a);
            }
        }
    }
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
OPS_HD void element_forces_mq(const LaneRegs<EPL> &rg, const GroupStore &gs, double invLe, int kk, const Pair &mq,
                              double &Mc, double &Qv)
{
    const Pair lo = table_row(gs.tab, row_offset(rg.spans, kk));
    Mc = fma(lo.y, slot_ke<EPL>(rg, kk), mq.x + lo.x);
    Qv = fma(lo.y, invLe, mq.y);
}
template <int EPL>
OPS_HD void element_forces(const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe, int kk,
                           double &Mc, double &Qv)
{
    element_forces_mq<EPL>(rg, gs, invLe, kk, ls.mq[(long)kk * ls.ls], Mc, Qv);
}

// Slots are processed in batches of NBP PAIRS, STAGE BY STAGE across the batch: the stages of one element are a
// ~150-cycle dependent chain (MUFU -> Newton -> quotient -> ...), and written element by element ptxas leaves the
// chains serial at this register budget; stage-major order puts the batch's independent instructions between
// dependent ones.  A packed fp32 instruction occupies the FMA pipe for two cycles, so two pairs per stage already
// cover its four-cycle latency.
#ifndef OPS_LANES_NBP
#define OPS_LANES_NBP 3
#endif
constexpr int NBP = OPS_LANES_NBP;

// multi-case kernels, before the pass: M^2, V^2 of this group's load case into the exchange column.  mqk: {M0, Q0} of
// the lane's first TEAM_KEEP slots, kept in registers since team_pass1 of the same epoch fetched them (a team member
// carries one or two pairs of optimiser state, not seven; the team kernel issues 581 shared-memory wavefronts per
// warp-epoch against ~400 of the single-case kernel, profiles/r02_ncu_summary.md)
#ifndef OPS_TEAM_KEEP
#define OPS_TEAM_KEEP 8
#endif
// slots whose {M0, Q0} stay in registers, the rest is fetched again: all 13 spill; measured 0 / 8 / 13 kept:
// 343 k / 350 k / 345 k beams/s with 8 load cases (profiles/r02_team_keep_ab.txt)
constexpr int TEAM_KEEP = OPS_TEAM_KEEP;
template <int EPL>
OPS_HD void lane_case_squares(const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe,
                              const Pair (&mqk)[EPL])
{
#pragma unroll
    for (int kk = 0; kk < EPL; ++kk) {
        double Mc, Qv;
        element_forces_mq<EPL>(rg, gs, invLe, kk, kk < TEAM_KEEP ? mqk[kk] : ls.mq[(long)kk * ls.ls], Mc, Qv);
        const float Mf = (float)Mc, Vf = (float)Qv;
        PairF x;
        x.c = Mf * Mf; x.h = Vf * Vf;
        ls.xc[(long)kk * ls.ls] = x;
    }
}

// The fp32 half of one batch of slot pairs, stage by stage (used by the single-case pass and by the owners of a team).
// In scope: OPS_P ("for every live pair i of the batch"), k, one, half and the per-pair arrays
// I, c, h (inputs) -> d, q, g (outputs), scratch nb, y, rb, s, gg, rgg, rs, db, qg, t0, t1, t2.  Stages:
//   nb = -(2E I + eps) ; y ~ 1 / sqrt(I)            rb = refined 1 / b ; s = sqrt(I)
//   d = c / b ; gg = Gf (kf s) ; rs = 1 / s          db = d / b ; rgg = refined 1 / gg ; rs = RN(1 / s)
//   q = h / gg ; bending branch gb = ((-am) db) E2   qg = q / gg
//   shear branch gs = ((((-as) qg) Gf) kf) (0.5 / s) ; g = (1 + gs) + gb   (sums of products: scalar adds, fastmath.cuh)
#define OPS_FP32_CHAIN \
        OPS_P { nb[i] = fm::mul2_add(I[i], splat(-k.E2), splat(-k.epsf)); y[i] = fm::rsq2_a(I[i]); } \
        OPS_P { rb[i] = fm::rcp2_a(neg2(nb[i])); t0[i] = mul2(I[i], y[i]); y[i] = mul2(y[i], half); } \
        OPS_P { t1[i] = fma2(nb[i], rb[i], one); t2[i] = fma2(neg2(t0[i]), t0[i], I[i]); } \
        OPS_P { rb[i] = fma2(rb[i], t1[i], rb[i]); s[i] = fma2(t2[i], y[i], t0[i]); } \
        OPS_P { t0[i] = mul2(c[i], rb[i]); gg[i] = mul2(splat(k.kf), s[i]); rs[i] = fm::rcp2_a(s[i]); } \
        OPS_P { t1[i] = fma2(nb[i], t0[i], c[i]); gg[i] = mul2(splat(k.Gf), gg[i]); t2[i] = fma2(neg2(s[i]), rs[i], one); } \
        OPS_P { d[i] = fma2(rb[i], t1[i], t0[i]); rgg[i] = fm::rcp2_a(gg[i]); rs[i] = fma2(rs[i], t2[i], rs[i]); } \
        OPS_P { t0[i] = mul2(d[i], rb[i]); t1[i] = fma2(neg2(gg[i]), rgg[i], one); t2[i] = fma2(neg2(s[i]), rs[i], one); } \
        OPS_P { db[i] = fma2(nb[i], t0[i], d[i]); rgg[i] = fma2(rgg[i], t1[i], rgg[i]); rs[i] = fma2(rs[i], t2[i], rs[i]); } \
        OPS_P { db[i] = fma2(rb[i], db[i], t0[i]); t1[i] = mul2(h[i], rgg[i]); rs[i] = mul2(half, rs[i]); } \
        OPS_P { db[i] = mul2(splat(-k.am), db[i]); t2[i] = fma2(neg2(gg[i]), t1[i], h[i]); } \
        OPS_P { db[i] = mul2(db[i], splat(k.E2)); q[i] = fma2(rgg[i], t2[i], t1[i]); } \
        OPS_P t0[i] = mul2(q[i], rgg[i]); \
        OPS_P t1[i] = fma2(neg2(gg[i]), t0[i], q[i]); \
        OPS_P qg[i] = fma2(rgg[i], t1[i], t0[i]); \
        OPS_P qg[i] = mul2(splat(-k.as_), qg[i]); \
        OPS_P qg[i] = mul2(qg[i], splat(k.Gf)); \
        OPS_P qg[i] = mul2(qg[i], splat(k.kf)); \
        OPS_P qg[i] = fm::mul2_add(qg[i], rs[i], one); \
        OPS_P g[i] = fm::f2(qg[i].x + db[i].x, qg[i].y + db[i].y);

// Adam's m, v, the step and the clamp for the batch (element_update_f32, second half).  In scope besides the above:
// rg, p0, neg_step, bc2_sqrt, rbc and the batch's Adam state m_[i], v_[i] (updated in place); reads I[i], g[i];
// leaves the stepped inertias in rg.I[p0 + i].
//   sqrt(v) / bc2_sqrt + eps ; I + (neg_step m) / denom ; clamp   (t0 is finite: den > 0, v and m finite; v is never
//   NaN here: the loss was finite).  v below the fast square root's range: the generic operators (cold).
#define OPS_ADAM_STEP \
        F2 vmin = splat(3.0e38f); \
        OPS_P { t0[i] = add2(g[i], neg2(m_[i])); t1[i] = mul2(splat(k.omb2f), g[i]); t2[i] = mul2(v_[i], splat(k.b2f)); } \
        OPS_P { m_[i] = fma2(splat(k.w1), t0[i], m_[i]); v_[i] = fma2(t1[i], g[i], t2[i]); } \
        OPS_P { vmin.x = fminf(vmin.x, v_[i].x); vmin.y = fminf(vmin.y, v_[i].y); } \
        if (fminf(vmin.x, vmin.y) >= fm::SQRT_F_MIN) { \
            F2 den[NB], rd[NB], num[NB]; \
            OPS_P y[i] = fm::rsq2_a(v_[i]); \
            OPS_P { t0[i] = mul2(v_[i], y[i]); y[i] = mul2(y[i], half); num[i] = mul2(splat(neg_step), m_[i]); } \
            OPS_P t1[i] = fma2(neg2(t0[i]), t0[i], v_[i]); \
            OPS_P t0[i] = fma2(t1[i], y[i], t0[i]); \
            OPS_P t1[i] = mul2(t0[i], splat(rbc)); \
            OPS_P den[i] = fma2(splat(-bc2_sqrt), t1[i], t0[i]); \
            OPS_P den[i] = fma2(splat(rbc), den[i], t1[i]); \
            OPS_P den[i] = add2(den[i], splat(k.adam_epsf)); \
            OPS_P rd[i] = fm::rcp2_a(den[i]); \
            OPS_P t0[i] = fma2(neg2(den[i]), rd[i], one); \
            OPS_P rd[i] = fma2(rd[i], t0[i], rd[i]); \
            OPS_P t0[i] = mul2(num[i], rd[i]); \
            OPS_P t1[i] = fma2(neg2(den[i]), t0[i], num[i]); \
            OPS_P t0[i] = fma2(rd[i], t1[i], t0[i]); \
            OPS_P t0[i] = add2(I[i], t0[i]); \
            OPS_P rg.I[p0 + i] = fm::f2(fmaxf(t0[i].x, k.clampf), fmaxf(t0[i].y, k.clampf)); \
        } else { \
            OPS_P { \
                const float dx = sqrtf(v_[i].x) / bc2_sqrt + k.adam_epsf, dy = sqrtf(v_[i].y) / bc2_sqrt + k.adam_epsf; \
                const float xx = I[i].x + (neg_step * m_[i].x) / dx, xy = I[i].y + (neg_step * m_[i].y) / dy; \
                rg.I[p0 + i] = fm::f2(xx < k.clampf ? k.clampf : xx, xy < k.clampf ? k.clampf : xy); \
            } \
        }

// torch.sum partials of one quantity: the four ILP rows as two packed pairs {row 0, row 1}, {row 2, row 3}, and the
// scalar tail.  Pair j = slots 2 j, 2 j + 1 falls on rows (0, 1) or (2, 3) of the 4-row block, so inside the block
// one packed add per pair keeps every row's own summation order.
struct SumAcc {
    fm::F2 r01, r23;
    float tail;
};
OPS_HD void sum_slot(SumAcc &a, const SumShape &sh, int kk, int l, float x)
{
    if (kk < sh.blk) {
        const int r = kk & 3;
        if (r == 0) a.r01.x += x; else if (r == 1) a.r01.y += x; else if (r == 2) a.r23.x += x; else a.r23.y += x;
    } else if (kk < sh.vec) {
        a.r01.x += x;
    } else if (kk == sh.vec && l < sh.ntail) {
        a.tail = x;
    }
}
template <int EPL>
OPS_HD void sum_pair(SumAcc &a, const SumShape &sh, int j, int l, fm::F2 x)
{
    if (2 * j + 1 < sh.blk) {
        if (j & 1) a.r23 = fm::add2(a.r23, x); else a.r01 = fm::add2(a.r01, x);
    } else {
        sum_slot(a, sh, 2 * j, l, x.x);
        if (2 * j + 1 < EPL) sum_slot(a, sh, 2 * j + 1, l, x.y);
    }
}
OPS_HD float sum_rows(const SumAcc &a) { return ((a.r01.x + a.r01.y) + a.r23.x) + a.r23.y; }

// THE PASS over the lane's elements, once per epoch: end forces from the span table, loss terms d, q and autograd's
// gradient with M, V constant (element_update_f32, first half), torch.sum partials of sum I, sum d, sum q; then --
// the record of a beam needs nothing of the step but the new I, and the step is taken whether or not this epoch
// turns out to be the beam's last (SingleCore:202-208 come before the stop test :211-219) -- the Adam step + clamp
// on the lane's elements (second half) and the flexibility sums of the NEXT epoch on the updated inertias.
// NC > 1: the squares come from the exchange columns of the team (case_id = this group's case).
//
// stage_I: this epoch may be the beam's last (the caller knows: epoch count, patience counter), so the inertias the
// forces were computed FROM -- the record's displacements integrate them -- are parked in the scratch column, in
// place of the next epoch's partial sums; if the beam goes on after all, the caller redoes the sums (lane_pass1).
// fp32 ranges for the branch-free division / square root: I in [clamp_min, 1e20) (the clamp, SingleCore:208),
// c = M^2 and h = V^2 zero or >= 2^-100.  The fast square root of Adam's v needs v >= 2^-101; v is an EMA of g^2, so
// anything smaller means g vanished on every epoch so far -- tested per batch, with the generic operators as the
// (cold) alternative.
template <int EPL, int NC, int NBX = NBP, class HM = HomeShared>
OPS_HD void lane_pass(const BeamConsts &k, int n, LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs,
                      const Pass1Consts &pc, double invLe, int l, int case_id, float neg_step, float bc2_sqrt, bool stage_I,
                      const HM &hm = HM())
{
    using fm::F2; using fm::splat; using fm::neg2; using fm::mul2; using fm::add2; using fm::fma2;
    constexpr int NB = NBX;                     // pairs per batch of this instance
    constexpr int NP = LaneRegs<EPL>::NP;
    const SumShape sh = sum_shape(n);
    SumAcc aI = {splat(0.0f), splat(0.0f), 0.0f}, ad = aI, aq = aI;
    SpanSums acc = {0.0, 0.0, 0.0, 0.0, 0.0};
    const PairF *x0 = ls.xc - (long)case_id * LPB;             // this lane's column in the team's case-0 group
    const F2 one = splat(1.0f), half = splat(0.5f);
    const float rbc = fm::rcp_r(bc2_sqrt);
    // one operation per stage across the batch (OPS_P: "for every live pair i of the batch", OPS_S: "for every live
    // slot s of the batch", slot kk = 2 p0 + s); the fast-path sequences of fastmath.cuh are written out step by
    // step so that consecutive instructions are independent
#define OPS_P _Pragma("unroll") for (int i = 0; i < NB; ++i) if (p0 + i < NP)
#define OPS_S _Pragma("unroll") for (int s_ = 0; s_ < 2 * NB; ++s_) if (2 * p0 + s_ < EPL)
#pragma unroll
    for (int p0 = 0; p0 < NP; p0 += NB) {
        F2 I[NB], c[NB], h[NB], nb[NB], y[NB], rb[NB], s[NB], gg[NB], rgg[NB], rs[NB], d[NB], db[NB], q[NB], qg[NB], g[NB];
        F2 t0[NB], t1[NB], t2[NB];
        F2 m_[NB], v_[NB];
        Pair mq[2 * NB];
        double ke[2 * NB];
        OPS_P I[i] = rg.I[p0 + i];
        if (NC == 1) {
            Pair lo[2 * NB];
            double Mc[2 * NB], Qv[2 * NB];
#pragma unroll
            for (int s_ = 0; s_ < 2 * NB; ++s_) { Mc[s_] = 0.0; Qv[s_] = 0.0; }      // (the padding half of an odd EPL)
            if constexpr (HM::TM) hm.template fetch_mq<EPL, 2 * NB>(ls, 2 * p0, mq);
            OPS_S {
                lo[s_] = table_row(gs.tab, row_offset(rg.spans, 2 * p0 + s_));
                if constexpr (!HM::TM) mq[s_] = ls.mq[(long)(2 * p0 + s_) * ls.ls];
                ke[s_] = slot_ke<EPL>(rg, 2 * p0 + s_);
            }
            OPS_S { Mc[s_] = mq[s_].x + lo[s_].x; Qv[s_] = fma(lo[s_].y, invLe, mq[s_].y); }
            OPS_S Mc[s_] = fma(lo[s_].y, ke[s_], Mc[s_]);
            OPS_P { c[i] = fm::f2((float)Mc[2 * i], (float)Mc[2 * i + 1]); h[i] = fm::f2((float)Qv[2 * i], (float)Qv[2 * i + 1]); }
            OPS_P { c[i] = mul2(c[i], c[i]); h[i] = mul2(h[i], h[i]); }
        } else {
            OPS_P {
                c[i] = splat(0.0f); h[i] = splat(0.0f);
#pragma unroll
                for (int hs = 0; hs < 2; ++hs) {
                    const int kk = 2 * (p0 + i) + hs;
                    if (kk < EPL) {
                        PairF x = x0[(long)kk * ls.ls];
                        float cs = x.c, hq = x.h;
#pragma unroll
                        for (int cc = 1; cc < NC; ++cc) {
                            x = x0[(long)kk * ls.ls + cc * LPB];
                            cs += x.c; hq += x.h;
                        }
                        if (hs) { c[i].y = cs; h[i].y = hq; } else { c[i].x = cs; h[i].x = hq; }
                    }
                }
            }
        }
        OPS_FP32_CHAIN
        OPS_P {
            sum_pair<EPL>(aI, sh, p0 + i, l, I[i]); sum_pair<EPL>(ad, sh, p0 + i, l, d[i]); sum_pair<EPL>(aq, sh, p0 + i, l, q[i]);
        }
        if (stage_I) {
            OPS_P *reinterpret_cast<F2 *>(ls.scr + (long)(p0 + i) * ls.ls) = I[i];
        }
        hm.template fetch_mv<EPL, NB>(rg, p0, m_, v_);
        OPS_ADAM_STEP
        hm.template put_mv<EPL, NB>(rg, p0, m_, v_);
        // flexibility sums of the next epoch
        {
            double Id[2 * NB], r[2 * NB], e[2 * NB];
            if (NC > 1) {
                OPS_S { mq[s_] = ls.mq[(long)(2 * p0 + s_) * ls.ls]; ke[s_] = slot_ke<EPL>(rg, 2 * p0 + s_); }
            }
            OPS_S Id[s_] = pos_normal_float_to_double((s_ & 1) ? rg.I[p0 + (s_ >> 1)].y : rg.I[p0 + (s_ >> 1)].x);   // (>= clamp_min, finite)
            OPS_S r[s_] = fm::rcp64_a(Id[s_]);
            OPS_S e[s_] = fma(-Id[s_], r[s_], 1.0);           // rcp64_n, stage by stage
            OPS_S e[s_] = fma(e[s_], e[s_], e[s_]);
            OPS_S r[s_] = fma(r[s_], e[s_], r[s_]);
            OPS_S pass1_accumulate<EPL>(rg, ls, pc, 2 * p0 + s_, r[s_], ke[s_], mq[s_], !stage_I, acc);
        }
    }
#undef OPS_P
#undef OPS_S
    float *st = reinterpret_cast<float *>(ls.scr + (long)SCR_STAGE * ls.ls);
    const long fs_ = 2 * ls.ls;                 // float stride between slots
    st[0] = sum_rows(aI); st[1] = aI.tail;
    st[fs_] = sum_rows(ad); st[fs_ + 1] = ad.tail;
    st[2 * fs_] = sum_rows(aq); st[2 * fs_ + 1] = aq.tail;
}

// ---------------------------------------------------------------------------------------------
// Load cases sharing one inertia vector: the team's division of labour.
//
// The NC groups of a team (one per load case) used to carry identical I, m, v and to repeat the whole fp32 half.
// Now slot pair j of a lane belongs to ONE group, its owner j % NC: only the owner runs the pair's fp32 chain and
// Adam step and keeps its m, v; what the others need travels through the owner's exchange rows {I_old, d, q, I_new}
// (XB_ROWS fp32 pairs per owned pair, in the owner's lane column).  Per epoch, separated by team barriers:
//   P1 (every group, its own case): flexibility sums from I_new of ALL pairs (team_pass1), reduction, support moments,
//      M^2 and V^2 of its case for all slots into the squares exchange (lane_case_squares);
//   P2 (owners): squares summed over the cases in case order -> loss terms, gradient, Adam, clamp -> exchange rows;
//   P3 (every group, redundantly: the stop decision must be the same everywhere): torch.sum partials of I_old, d, q
//      over all pairs -> loss.
// Nothing else of the state lives in registers, and the record reads I_old / I_new from the exchange.
// ---------------------------------------------------------------------------------------------
template <int NC>
OPS_HD fm::F2 *team_row(const LaneStore &ls, int case_id, int j, int f)
{
    return reinterpret_cast<fm::F2 *>(ls.xb + ((long)(j % NC) - case_id) * LPB + (long)((j / NC) * XB_ROWS + f) * ls.ls);
}

// once per beam (after lane_init): the owned pairs start from I_0, their exchange rows likewise
template <int EPL, int NC>
OPS_HD void team_init(const BeamConsts &k, int n, const LaneStore &ls, int l, int case_id, LaneRegs<EPL> &rg)
{
    constexpr int NP = LaneRegs<EPL>::NP, NPO = team_owned_pairs(EPL, NC);
#pragma unroll
    for (int p_ = 0; p_ < NPO; ++p_) {
        const int j = case_id + p_ * NC;
        const int e0 = LPB * (2 * j) + l, e1 = LPB * (2 * j + 1) + l;
        const fm::F2 I0 = fm::f2((2 * j < EPL && e0 < n) ? k.I0f : 1.0f, (2 * j + 1 < EPL && e1 < n) ? k.I0f : 1.0f);
        rg.I[p_] = I0; rg.m[p_] = fm::splat(0.0f); rg.v[p_] = fm::splat(0.0f);
        if (j < NP) {
            fm::F2 *row = reinterpret_cast<fm::F2 *>(ls.xb + (long)(p_ * XB_ROWS) * ls.ls);
            row[0] = I0;
            *reinterpret_cast<fm::F2 *>(ls.xb + (long)(p_ * XB_ROWS + 1) * ls.ls) = fm::splat(0.0f);
            *reinterpret_cast<fm::F2 *>(ls.xb + (long)(p_ * XB_ROWS + 2) * ls.ls) = fm::splat(0.0f);
            *reinterpret_cast<fm::F2 *>(ls.xb + (long)(p_ * XB_ROWS + 3) * ls.ls) = I0;
        }
    }
}

// P1: the five sums of this group's case from the team's current inertias (stage-major batches of six slots);
// the {M0, Q0} it fetches stay in mqk for lane_case_squares
template <int EPL, int NC>
OPS_HD void team_pass1(const LaneRegs<EPL> &rg, const LaneStore &ls, const Pass1Consts &pc, int case_id, Pair (&mqk)[EPL])
{
    constexpr int SB = 6;
    SpanSums acc = {0.0, 0.0, 0.0, 0.0, 0.0};
#define OPS_T _Pragma("unroll") for (int s_ = 0; s_ < SB; ++s_) if (k0 + s_ < EPL)
#pragma unroll
    for (int k0 = 0; k0 < EPL; k0 += SB) {
        double Id[SB], r[SB], e[SB], ke[SB];
        Pair mq[SB];
        OPS_T {
            const int kk = k0 + s_;
            const fm::F2 In = *team_row<NC>(ls, case_id, kk >> 1, 3);
            Id[s_] = (double)((kk & 1) ? In.y : In.x);
            mq[s_] = ls.mq[(long)kk * ls.ls];
            if (kk < TEAM_KEEP) mqk[kk] = mq[s_];
            ke[s_] = slot_ke<EPL>(rg, kk);
        }
        OPS_T r[s_] = fm::rcp64_a(Id[s_]);
        OPS_T e[s_] = fma(-Id[s_], r[s_], 1.0);
        OPS_T e[s_] = fma(e[s_], e[s_], e[s_]);
        OPS_T r[s_] = fma(r[s_], e[s_], r[s_]);
        OPS_T pass1_accumulate<EPL>(rg, ls, pc, k0 + s_, r[s_], ke[s_], mq[s_], true, acc);
    }
#undef OPS_T
}

// P2: the owner's pairs -- squares of all cases, fp32 chain, Adam, exchange rows
template <int EPL, int NC>
OPS_HD void team_owner_update(const BeamConsts &k, LaneRegs<EPL> &rg, const LaneStore &ls, int case_id, float neg_step,
                              float bc2_sqrt)
{
    using fm::F2; using fm::splat; using fm::neg2; using fm::mul2; using fm::add2; using fm::fma2;
    constexpr int NP = LaneRegs<EPL>::NP, NB = team_owned_pairs(EPL, NC), p0 = 0;
    const PairF *x0 = ls.xc - (long)case_id * LPB;             // this lane's column in the team's case-0 group
    const F2 one = splat(1.0f), half = splat(0.5f);
    const float rbc = fm::rcp_r(bc2_sqrt);
    F2 I[NB], c[NB], h[NB], nb[NB], y[NB], rb[NB], s[NB], gg[NB], rgg[NB], rs[NB], d[NB], db[NB], q[NB], qg[NB], g[NB];
    F2 t0[NB], t1[NB], t2[NB];
#define OPS_P _Pragma("unroll") for (int i = 0; i < NB; ++i)
    OPS_P {
        const int j = case_id + i * NC;
        I[i] = rg.I[i];
        c[i] = splat(0.0f); h[i] = splat(0.0f);
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
            const int kk = 2 * j + hs;
            if (j < NP && kk < EPL) {
                PairF x = x0[(long)kk * ls.ls];
                float cs = x.c, hq = x.h;
#pragma unroll
                for (int cc = 1; cc < NC; ++cc) {
                    x = x0[(long)kk * ls.ls + cc * LPB];
                    cs += x.c; hq += x.h;
                }
                if (hs) { c[i].y = cs; h[i].y = hq; } else { c[i].x = cs; h[i].x = hq; }
            }
        }
    }
    OPS_FP32_CHAIN
    {
        F2 m_[NB], v_[NB];
        OPS_P { m_[i] = rg.m[i]; v_[i] = rg.v[i]; }
        { OPS_ADAM_STEP }
        OPS_P { rg.m[i] = m_[i]; rg.v[i] = v_[i]; }
    }
    OPS_P {
        const int j = case_id + i * NC;
        if (j < NP) {
            *reinterpret_cast<F2 *>(ls.xb + (long)(i * XB_ROWS + 0) * ls.ls) = I[i];
            *reinterpret_cast<F2 *>(ls.xb + (long)(i * XB_ROWS + 1) * ls.ls) = d[i];
            *reinterpret_cast<F2 *>(ls.xb + (long)(i * XB_ROWS + 2) * ls.ls) = q[i];
            *reinterpret_cast<F2 *>(ls.xb + (long)(i * XB_ROWS + 3) * ls.ls) = rg.I[i];
        }
    }
#undef OPS_P
}

// P3: torch.sum partials of sum I, sum d, sum q over ALL pairs of the lane, from the owners' rows
template <int EPL, int NC>
OPS_HD void team_loss_sums(int n, const LaneStore &ls, int l, int case_id)
{
    constexpr int NP = LaneRegs<EPL>::NP;
    const SumShape sh = sum_shape(n);
    SumAcc aI = {fm::splat(0.0f), fm::splat(0.0f), 0.0f}, ad = aI, aq = aI;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        sum_pair<EPL>(aI, sh, j, l, *team_row<NC>(ls, case_id, j, 0));
        sum_pair<EPL>(ad, sh, j, l, *team_row<NC>(ls, case_id, j, 1));
        sum_pair<EPL>(aq, sh, j, l, *team_row<NC>(ls, case_id, j, 2));
    }
    float *st = reinterpret_cast<float *>(ls.scr + (long)SCR_STAGE * ls.ls);
    const long fs_ = 2 * ls.ls;
    st[0] = sum_rows(aI); st[1] = aI.tail;
    st[fs_] = sum_rows(ad); st[fs_ + 1] = ad.tail;
    st[2 * fs_] = sum_rows(aq); st[2 * fs_ + 1] = aq.tail;
}

// inertia of element e of the beam from the exchange rows (f = 0: last analysed, f = 3: after the last step), read by
// ANY lane of the group (lane 0's displacement march, the emission of I_values): lane ln's column is ln further on
template <int NC>
OPS_HD float team_inertia(const LaneStore &ls, int lane_of_reader, int case_id, int e, int f)
{
    const int kk = e >> 3, ln = e & (LPB - 1), j = kk >> 1;
    const fm::F2 v = *reinterpret_cast<const fm::F2 *>(ls.xb + (ln - lane_of_reader) + ((long)(j % NC) - case_id) * LPB +
                                                        (long)((j / NC) * XB_ROWS + f) * ls.ls);
    return (kk & 1) ? v.y : v.x;
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

// ---------------------------------------------------------------------------------------------
// once per beam: the record (SingleCore:221-249).  M, V, u, theta belong to the LAST ANALYSED
// inertias, i.e. the ones the beam's last pass started from (parked by it, stage_I); rg.I holds the stepped ones.
// ---------------------------------------------------------------------------------------------
template <int EPL, class HM = HomeShared>
OPS_HD void lane_emit_forces(int n, const LaneRegs<EPL> &rg, const LaneStore &ls, const GroupStore &gs, double invLe, int l,
                             bool fields, float *shear, float *moment, const HM &hm = HM(), bool active = true)
{
    constexpr int SB = HM::TM ? 4 : 1;          // slots per fetch (HomeTm: the whole warp calls, `active` lanes write)
#pragma unroll
    for (int k0 = 0; k0 < EPL; k0 += SB) {
        Pair mq[SB];
        hm.template fetch_mq<EPL, SB>(ls, k0, mq);
#pragma unroll
        for (int s_ = 0; s_ < SB; ++s_) {
            const int kk = k0 + s_;
            const int e = LPB * kk + l;
            if (kk < EPL && e < n && active) {
                double Mc = 0.0, Qv = 0.0;
                if (fields) element_forces_mq<EPL>(rg, gs, invLe, kk, mq[s_], Mc, Qv);
                shear[e] = fields ? (float)Qv : 0.0f;
                moment[e] = fields ? (float)(-Mc) : 0.0f;
            }
        }
    }
}

// the inertias the beam's last pass parked in the scratch columns (pair j of a lane = float pair at slot j), seen from lane 0
struct ParkedInertia {
    const float *stage;
    long ls;
    OPS_HD double operator()(int e) const
    {
        const int kk = e >> 3, ln = e & (LPB - 1);
        return (double)stage[(long)(kk >> 1) * 2 * ls + 2 * ln + (kk & 1)];
    }
};

// lane 0: displacements by integrating the curvature (flex_deflections_march); `inertia(e)` = the last analysed I_e
template <class InertiaFn>
OPS_HD void group_emit_displacements(const BeamConsts &k, const FlexBeam &fb, const GroupStore &gs, bool fields, InertiaFn inertia,
                                     double *defl, double *rot)
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
        if (e < n) I_out[e] = (kk & 1) ? rg.I[kk >> 1].y : rg.I[kk >> 1].x;
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
