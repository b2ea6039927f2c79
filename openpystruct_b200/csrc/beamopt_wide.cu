// sm_100a kernels of the shared-memory-state iteration (beamopt_wide.cuh): LPB lanes per beam, persistent
// groups, one CTA per SM.  A group that finishes its beam (early stop, SingleCore:211-219, or max_e
// epochs) takes the CTA's next beam; nothing crosses groups, nothing but the beam's inputs (read once) and
// its record (written once) crosses HBM.  Cross-lane steps of a group: one butterfly sum per closing span
// and the final loss combine, both with width-LPB shuffles.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (one IEEE rounding per written op).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beamopt_internal.cuh"
#include "beamopt_wide.cuh"

namespace ops {

using namespace wide;

// All collectives below are executed by the whole (converged) warp with the full mask; a group is the
// width-LPB segment of its lanes.  The kernel keeps the groups of a warp in lockstep for that reason.
constexpr unsigned FULL = 0xffffffffu;

template <int LPB>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int step = 1; step < LPB; step <<= 1) v += __shfl_xor_sync(FULL, v, step, LPB);
    return v;
}

// The five span sums of a whole warp (LPB = 32) with the additions of the butterfly above -- level by level v_l + v_(l ^ step),
// so the same bits -- but each level done by HALF of the lanes per value (recursive halving): after level 1 the even lanes
// carry the sums {0, 1, 2} and the odd lanes {3, 4}, after level 3 every lane carries one, 16 shuffles and 8 additions
// instead of 50 and 25.  The totals end up in lanes 0, 4, 2, 1, 3 (sums 0..4), which store them.
__device__ __forceinline__ void warp_sum5_store(double (&v)[NSUM], int l, double *tot)
{
    static_assert(NSUM == 5, "five sums");
    const bool b0 = l & 1, b1 = l & 2, b2 = l & 4;
    // level 1: even lanes keep {0, 1, 2}, odd lanes {3, 4}
    const double r0 = __shfl_xor_sync(FULL, b0 ? v[0] : v[3], 1);
    const double r1 = __shfl_xor_sync(FULL, b0 ? v[1] : v[4], 1);
    const double r2 = __shfl_xor_sync(FULL, v[2], 1);
    const double a0 = (b0 ? v[3] : v[0]) + r0, a1 = (b0 ? v[4] : v[1]) + r1, a2 = v[2] + r2;      // (a2: even lanes only)
    // level 2: classes (b0, b1) = (0,0) keep {0, 1}, (0,1) keeps 2, (1,0) keeps 3, (1,1) keeps 4
    const double other = b0 ? a1 : a2;
    const double rx = __shfl_xor_sync(FULL, b1 ? a0 : other, 2);
    const double ry = __shfl_xor_sync(FULL, a1, 2);
    const double c0 = (b1 ? other : a0) + rx, c1 = a1 + ry;                                        // (c1: class (0,0) only)
    // level 3: class (0,0) splits {0, 1} over b2, the others go on with their one sum
    const bool z = !b0 && !b1;
    const double rz = __shfl_xor_sync(FULL, (z && !b2) ? c1 : c0, 4);
    double e = ((z && b2) ? c1 : c0) + rz;
    // levels 4, 5
    e += __shfl_xor_sync(FULL, e, 8);
    e += __shfl_xor_sync(FULL, e, 16);
    if (l < 5) tot[(0x14230u >> (4 * l)) & 7u] = e;               // lanes 0, 1, 2, 3, 4 hold the sums 0, 3, 2, 4, 1
}

// span j closes in this slot
template <int LPB>
__device__ __forceinline__ void dev_close_span(const WideStore &ws, int j, bool mine, const double (&aold)[NSUM], int sold,
                                               const double (&x)[NSUM], int sp, int l)
{
    double v[NSUM];
    close_value(aold, sold, x, sp, j, v);
    if constexpr (LPB == 32) {
        warp_sum5_store(v, l, ws.tot + j * NSUM);
    } else {
#pragma unroll
        for (int w = 0; w < NSUM; ++w) v[w] = group_sum<LPB>(v[w]);
        if (mine && l == 0) {
#pragma unroll
            for (int w = 0; w < NSUM; ++w) ws.tot[j * NSUM + w] = v[w];
        }
    }
}

// spans whose last element lies in slot s (for any group of the warp): every lane's share of the span's five sums,
// added across the group in the butterfly's order, published in the beam's `tot`
template <int LPB>
__device__ __forceinline__ void dev_close(const WideStore &ws, int s, bool cl, const double (&aold)[NSUM], int sold,
                                          const double (&x)[NSUM], int sp, int l)
{
    if constexpr (LPB == 32) {
        // one beam per warp: which spans close here is a warp-uniform byte of the beam (no scan over the spans, no votes)
        unsigned int m = ws.cmask[s];
#pragma unroll 1
        while (m) {
            const int j = __ffs((int)m) - 1;
            m &= m - 1;
            dev_close_span<LPB>(ws, j, true, aold, sold, x, sp, l);
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < NSPAN; ++j) {
            const bool mine = cl && ws.gi[GI_CLOSE + j] == s;
            if (!__any_sync(FULL, mine)) continue;
            dev_close_span<LPB>(ws, j, mine, aold, sold, x, sp, l);
        }
    }
}

template <int LPB, int N>
__device__ __forceinline__ void dev_batch(const BeamConsts &k, const FlexBeam &fb, const WideShape &sh, const WideStore &ws,
                                          int l, int kb, bool run, const float *Icur, float *Inew, const SweepConsts &sc,
                                          LaneCtx<LPB> &cx)
{
    BatchOut<N> bo;
    sweep_batch<LPB, N>(k, fb, sh, ws, l, kb, run, Icur, Inew, sc, cx, bo);
    if constexpr (LPB == 32 && N == 4) {
        // most batches close no span: one test of the batch's four bytes (kb is a multiple of four), then no branches
        if (*reinterpret_cast<const unsigned int *>(ws.cmask + kb) == 0u) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double x[NSUM], aold[NSUM];
                int sold;
                slot_terms<N>(bo, i, x);
                slot_accumulate<LPB>(cx, x, bo.sp[i], aold, sold);
            }
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x[NSUM], aold[NSUM];
        int sold;
        slot_terms<N>(bo, i, x);
        slot_accumulate<LPB>(cx, x, bo.sp[i], aold, sold);
        // (LPB = 32: the flag is the same on all lanes of the warp -- they belong to one beam)
        const bool closing = LPB == 32 ? bo.close[i] : (__any_sync(FULL, bo.close[i]) != 0);
        if (closing) dev_close<LPB>(ws, kb + i, bo.close[i], aold, sold, x, bo.sp[i], l);
    }
}

// compile-time slot count: balanced batches of at most NBX slots, everything static
template <int LPB, int KFIX, int NBX, int BI = 0>
__device__ __forceinline__ void unrolled_batches(const BeamConsts &k, const FlexBeam &fb, const WideShape &sh,
                                                 const WideStore &ws, int l, bool run, const float *Icur, float *Inew,
                                                 const SweepConsts &sc, LaneCtx<LPB> &cx)
{
    constexpr int nbatch = (KFIX + NBX - 1) / NBX;
    if constexpr (BI < nbatch) {
        constexpr int base = KFIX / nbatch, extra = KFIX % nbatch;
        constexpr int size = base + (BI < extra ? 1 : 0);
        constexpr int start = BI * base + (BI < extra ? BI : extra);
        dev_batch<LPB, size>(k, fb, sh, ws, l, start, run, Icur, Inew, sc, cx);
        unrolled_batches<LPB, KFIX, NBX, BI + 1>(k, fb, sh, ws, l, run, Icur, Inew, sc, cx);
    }
}

template <int LPB, int KFIX, int NBX>
__device__ __forceinline__ void dev_sweep(const BeamConsts &k, const FlexBeam &fb, const WideShape &sh, const WideStore &ws,
                                          int l, bool run, const float *Icur, float *Inew, const SweepConsts &sc,
                                          LaneCtx<LPB> &cx)
{
    ctx_reset<LPB>(cx, l);
    if constexpr (KFIX > 0) {
        unrolled_batches<LPB, KFIX, NBX>(k, fb, sh, ws, l, run, Icur, Inew, sc, cx);
    } else {
        int kb = 0;
#pragma unroll 1
        for (; kb + NBX <= sh.K; kb += NBX) dev_batch<LPB, NBX>(k, fb, sh, ws, l, kb, run, Icur, Inew, sc, cx);
#pragma unroll 1
        for (; kb < sh.K; ++kb) dev_batch<LPB, 1>(k, fb, sh, ws, l, kb, run, Icur, Inew, sc, cx);
    }
    ctx_finish<LPB>(sh, cx);
}

// total loss in torch's order (wide_loss_arrays with shuffles); every lane of the group obtains it
template <int LPB>
__device__ __forceinline__ float dev_loss(const BeamConsts &k, const WideShape &sh, const LaneCtx<LPB> &cx, int l)
{
    float s[3];
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        const float left = cx.left[w];
        float row, t = 0.0f;
        if (LPB == 32) {
            const float acc = cx.acc[w][0];
            float a8 = acc;
            for (int v = 0; v < sh.nlv; ++v) a8 += __shfl_sync(FULL, left, (l + 8 * v) & 31);
            const float v8 = __shfl_sync(FULL, acc, (l + 8) & 31), v16 = __shfl_sync(FULL, acc, (l + 16) & 31);
            const float v24 = __shfl_sync(FULL, acc, (l + 24) & 31);
            row = ((a8 + v8) + v16) + v24;
            for (int i = 0; i < sh.ntail; ++i) t += __shfl_sync(FULL, left, 8 * sh.nlv + i);
        } else {
            row = ctx_rowsum<LPB>(cx, w);
            for (int i = 0; i < sh.ntail; ++i) t += __shfl_sync(FULL, left, i, LPB);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) t += __shfl_sync(FULL, row, r, LPB);
        s[w] = t;
    }
    return (s[0] + k.am * s[1]) + k.as_ * s[2];
}

// record of a beam: fields of the last analysed inertias (the I buffer the last sweep read), inertias after
// the last Adam step (SingleCore:221-249); a beam rejected at set-up emits I_0
template <int LPB>
__device__ __forceinline__ void dev_record(const BeamConsts &k, const FlexBeam &fb, const WideShape &sh, const WideStore &ws,
                                           const OptPtrs &p, long long b, int l, unsigned gmask, int t, int bad, float lossf,
                                           bool setup_bad)
{
    const int n = k.n, nn = n + 1;
    const bool fields = (t > 0) && (bad == 0);
    wide_emit_lane<LPB>(fb, sh, ws, l, fields, ws.I0 + (t & 1) * ws.el, p.shear + b * n, p.moment + b * n,
                        setup_bad ? nullptr : p.I_values + b * n);
    if (setup_bad)
        for (int e = l; e < n; e += LPB) p.I_values[b * n + e] = k.I0f;
    if (l == 0) {
        wide_emit_displacements(k, fb, ws, ws.I0 + ((t + 1) & 1) * ws.el, fields, p.defl + b * nn, p.rot + b * nn);
        p.epochs[b] = t;
        p.loss[b] = lossf;
        p.status[b] = bad;
    }
    __syncwarp(gmask);
}

// Group states: the groups of a warp run the epoch phases in lockstep (the collectives are warp wide); a
// group without a running beam executes the sweep in its first-pass form on whatever its shared memory
// holds, which costs nothing but the tail of the launch.
enum { ST_NEED = 0, ST_FIRST = 1, ST_RUN = 2, ST_IDLE = 3 };

template <int LPB, int KFIX, int NBX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
beamopt_wide_kernel(const BeamConsts k, const long long B, const OptPtrs p, const int beam_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, l = tid & (LPB - 1), g = tid / LPB;
    const unsigned gmask = LPB == 32 ? FULL : (((1u << LPB) - 1u) << (tid & 31 & ~(LPB - 1)));
    const int n = k.n, nn = n + 1;
    const WideShape sh = wide_shape<LPB>(n);
    WideStore ws;
    wide_carve<LPB>(smem_raw + (size_t)g * beam_bytes, n, ws);
    for (int i = l; i < beam_bytes / 4; i += LPB) reinterpret_cast<int *>(smem_raw + (size_t)g * beam_bytes)[i] = 0;

    __shared__ unsigned int cta_next;
    if (tid == 0) cta_next = 0;
    __syncthreads();

    FlexBeam fb;
    flex_geometry(k, 1.0, fb);
    fb.m = 0; fb.last = 0; fb.nloads = 0; fb.Moh = 0.0; fb.Qoh = 0.0;
    SweepConsts sc;
    sc.G2 = 0.0; sc.H2 = 0.0; sc.neg_step = 0.0f; sc.bc2_sqrt = 1.0f; sc.rbc = 1.0f;
    LaneCtx<LPB> cx;
    SegStatics st = {{0.0, 0.0, 0.0}};
    long long b = -1;
    int state = ST_NEED, t = 0, counter = 0, bad = 0;
    double best = INFINITY;
    float lossf = NAN;

    while (true) {
        while (state == ST_NEED) {
            if (l == 0) b = (long long)blockIdx.x + (long long)gridDim.x * atomicAdd(&cta_next, 1u);
            b = __shfl_sync(gmask, b, 0, LPB);
            if (b >= B) { state = ST_IDLE; break; }
            if (l == 0) {
                int fnode[FLEX_MAXF];
                double fval[FLEX_MAXF];
                for (int j = 0; j < k.max_forces; ++j) {
                    fnode[j] = p.force_nodes[b * k.max_forces + j];
                    fval[j] = p.force_vals[b * k.max_forces + j];
                }
                const uint8_t *fx = p.fixed_uy + b * nn;
                wide_setup<LPB>(k, p.L[b], [&](int i) { return fx[i] != 0; }, fnode, fval, ws);
            }
            __syncwarp(gmask);
            bad = wide_fetch(k, p.L[b], ws, fb);
            wide_fetch_statics<LPB>(ws, l, st);
            sc.G2 = 6.0 * fb.wl2h; sc.H2 = 3.0 * fb.wl2h;
            t = 0; counter = 0; best = INFINITY; lossf = NAN;
            if (bad != 0 || k.max_epochs <= 0) {
                if (bad == 0) wide_lane_init<LPB>(k, sh, ws, l);
                __syncwarp(gmask);
                dev_record<LPB>(k, fb, sh, ws, p, b, l, gmask, 0, bad, lossf, bad != 0);
                continue;
            }
            wide_lane_init<LPB>(k, sh, ws, l);
            state = ST_FIRST;
        }
        if (!__any_sync(FULL, state != ST_IDLE)) break;
        // ---- the warp is converged from here to the end of the iteration ----
        const bool run = state == ST_RUN;
        int rc = 0;
        if (run) {
            sc.neg_step = __ldg(p.sched + 2 * t);
            sc.bc2_sqrt = __ldg(p.sched + 2 * t + 1);
            sc.rbc = fm::rcp_r(sc.bc2_sqrt);
            rc = wide_solve<LPB>(fb, ws, l, st);
        }
        __syncwarp();
        const int cur = run ? (t & 1) : 0;
        dev_sweep<LPB, KFIX, NBX>(k, fb, sh, ws, l, run, ws.I0 + cur * ws.el, ws.I0 + (cur ^ 1) * ws.el, sc, cx);
        const float lv_ = dev_loss<LPB>(k, sh, cx, l);
        __syncwarp();
        if (run) {
            lossf = lv_;
            ++t;
            bool done = false;
            if (rc || !(lossf - lossf == 0.0f)) { bad = 1; done = true; }
            if (k.early_stop) {
                const double lv = (double)lossf;
                if (lv < best - k.tol) { best = lv; counter = 0; } else { ++counter; }
                if (counter >= k.patience) done = true;
            }
            if (t >= k.max_epochs) done = true;
            if (done) {
                dev_record<LPB>(k, fb, sh, ws, p, b, l, gmask, t, bad, lossf, false);
                state = ST_NEED;
            }
        } else if (state == ST_FIRST) {
            state = ST_RUN;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#ifndef OPS_WIDE32_MAXT
#define OPS_WIDE32_MAXT 384     // 168-register cap (3 warps per SM sub-partition): +7 % on 1000-element beams over 512 / 128
#endif
constexpr int WIDE8_MAXT = 512, WIDE32_MAXT = OPS_WIDE32_MAXT;

bool wide_supported(const BeamConsts &k, int num_cases, int lpb, int smem_optin)
{
    if (num_cases != 1 || k.max_forces > FLEX_MAXF) return false;
    if (lpb == 8) return wide_shape_ok<8>(k.n) && wide_beam_bytes<8>(k.n) * 4 + 64 <= (size_t)smem_optin;
    if (lpb == 32) return wide_shape_ok<32>(k.n) && wide_beam_bytes<32>(k.n) + 64 <= (size_t)smem_optin;
    return false;
}

int wide_plan(const BeamConsts &k, int lpb, int64_t B, int sms, int smem_optin, WidePlan *pl)
{
    pl->lpb = lpb;
    pl->beam_bytes = (int)(lpb == 8 ? wide_beam_bytes<8>(k.n) : wide_beam_bytes<32>(k.n));
    const int maxt = lpb == 8 ? WIDE8_MAXT : WIDE32_MAXT;
    int groups = (int)(((size_t)smem_optin - 64) / (size_t)pl->beam_bytes);
    int T = groups * lpb / 32 * 32;
    if (T > maxt) T = maxt;
    const char *thr_env = getenv("OPS_WIDE_THREADS");             // profiling knob
    if (thr_env && atoi(thr_env) >= 32 && atoi(thr_env) <= T) T = atoi(thr_env) / 32 * 32;
    if (T < 32) return -2;
    pl->threads = T;
    pl->smem_bytes = (size_t)pl->beam_bytes * (T / lpb);
    const long per_cta = T / lpb;
    long want = (long)((B + per_cta - 1) / per_cta);
    pl->blocks = (int)(want < sms ? want : sms);
    if (pl->blocks < 1) pl->blocks = 1;
    if (pl->blocks < sms && B > pl->blocks) pl->blocks = (int)(B < sms ? B : sms);   // spread a small batch over all SMs
    return 0;
}

template <int LPB, int KFIX, int NBX, int MAXT>
static cudaError_t launch_instance(const BeamConsts &k, long long B, const OptPtrs &p, const WidePlan &pl, cudaStream_t stream)
{
    auto kern = beamopt_wide_kernel<LPB, KFIX, NBX, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<pl.blocks, pl.threads, pl.smem_bytes, stream>>>(k, B, p, pl.beam_bytes);
    return cudaGetLastError();
}

cudaError_t wide_launch(const BeamConsts &k, long long B, const OptPtrs &p, const WidePlan &pl, cudaStream_t stream)
{
    if (pl.lpb == 8) {
        switch (wide_slots<8>(k.n)) {
        case 4: return launch_instance<8, 4, 4, WIDE8_MAXT>(k, B, p, pl, stream);
        case 8: return launch_instance<8, 8, 4, WIDE8_MAXT>(k, B, p, pl, stream);
        case 13: return launch_instance<8, 13, 5, WIDE8_MAXT>(k, B, p, pl, stream);       // the reference's 100 elements
        default: return launch_instance<8, 21, 5, WIDE8_MAXT>(k, B, p, pl, stream);
        }
    }
    return launch_instance<32, 0, NB, WIDE32_MAXT>(k, B, p, pl, stream);
}

}  // namespace ops
