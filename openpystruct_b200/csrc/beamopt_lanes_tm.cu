// sm_100a kernel of the production iteration with the lane-private data in TENSOR MEMORY: eight lanes per beam,
// 16 warps = 64 beams resident per SM.
//
// Same arithmetic, same phase functions (beamopt_lanes.cuh) and same results, bit for bit, as beamopt_lanes_kernel
// (beamopt_lanes.cu).  What differs is where a lane keeps what only it ever touches (HomeTm in beamopt_lanes.cuh):
// the statics {M0, Q0} of its 13 elements (208 B) and Adam's m, v (104 B) live in the SM's 256 KB of tensor memory --
// idle on this path, which has no contraction for the tensor cores -- and travel through tcgen05.ld / tcgen05.st
// (SASS LDTM / STTM; measured: 25 cycles for a load + wait, 740 B per cycle and SM with 17 warps,
// scripts/ubench/tmem_ubench.cu).  That takes 1.7 KB per beam out of shared memory and 28 registers out of every
// thread: 512 threads of 128 registers and 64 beams x 2.4 KB fit one SM, where the register / shared-memory home
// stops at 10-12 warps of 168 registers.
//
// What it buys (profiles/r02_tensor_memory_instance.md): 64 beams per SM in ONE round of 6.6 us per epoch where the
// register instance needs two (40 + 24: 7.9 us) -- +19 % at 9 472 beams -- but no more than that: the epoch of a warp
// is longer at 128 registers, and at 16 warps the schedulers still issue on only 61 % of the cycles (register instance,
// 10 warps: 51 %), so per beam and epoch the two instances are within 8 % of each other and the register instance wins
// every batch that is not a single round of 52..64 beams per SM (lanes_plan picks by that rule).  A fifth warp per
// scheduler (17 / 20 warps, <= 96 registers) spills and is slower still.  More resident warps are therefore NOT what
// the iteration lacks; the measurement retires that hypothesis of round 1 / DESIGN 3.4.
//
// Tensor-memory instructions are warp-collective, a group's beams end at different epochs.  So every phase that touches
// tensor memory is entered by the whole warp on a warp-uniform condition and the lanes without work compute on whatever
// their columns hold, keeping the results to themselves:
//   * new beams: lane_init stages {M0, Q0} in the lane's scratch column, tm_commit moves them over for the `fresh` lanes
//     of the warp (read-modify-write for the others);
//   * the sums on their own (first epoch / after a parked epoch): lane_pass1 with an `active` predicate;
//   * THE PASS: all lanes; a group without a running beam touches only its own columns and registers;
//   * the record: lane_emit_forces with an `active` predicate, the rest of the record path is per group as before.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beamopt_internal.cuh"
#include "beamopt_lanes.cuh"

namespace ops {

using namespace lanes;

#ifndef OPS_TM_THREADS
#define OPS_TM_THREADS 512
#endif
constexpr int TM_THREADS = OPS_TM_THREADS;
#ifndef OPS_TM_NBP
#define OPS_TM_NBP 2
#endif
constexpr int TM_NBP = OPS_TM_NBP;              // slot pairs per stage-major batch (the register budget is 120)
constexpr int TM_COLUMNS = 512;                 // the whole tensor memory of the SM (one CTA per SM)

static size_t tm_group_bytes()
{
    return (size_t)LPB * 8 * SCR_SLOTS + (size_t)(TAB_SLOTS + GROUP_DOUBLES) * 8 + (size_t)GROUP_INTS * 4;
}

template <int EPL, int NFIX, int TFIX, bool SC>
__global__ void __launch_bounds__(TFIX, 1)
beamopt_lanes_tm_kernel(const BeamConsts k, const long long B, const OptPtrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T = TFIX, G = T / LPB, WARPS = T / 32;
    constexpr int COLS = HomeTm::columns(EPL);                      // tensor-memory columns of a warp
    static_assert(((WARPS + 3) / 4) * COLS <= TM_COLUMNS, "tensor memory: the warps of a lane quarter do not fit");
    const int tid = threadIdx.x, l = tid & (LPB - 1), g = tid / LPB, w = tid >> 5;
    const unsigned gmask = 0xffu << (tid & 24);
    const int n = NFIX ? NFIX : k.n;
    const int nn = n + 1;

    double *lane_d = reinterpret_cast<double *>(smem_raw);
    double *tab_d = lane_d + (size_t)SCR_SLOTS * T;
    double *grp_d = tab_d + (size_t)TAB_SLOTS * G;
    int *grp_i = reinterpret_cast<int *>(grp_d + (size_t)GROUP_DOUBLES * G);
    LaneStore ls;
    ls.ls = T;
    ls.mq = nullptr;
    ls.scr = lane_d + tid;
    ls.xc = nullptr;
    ls.xb = nullptr;
    GroupStore gs;
    gs.gs = G;
    gs.tab = tab_d + (size_t)TAB_SLOTS * g;
    gs.fs.sd = grp_d + g;
    gs.fs.stride = G;
    gs.gd = gs.fs.sd + (size_t)FlexStore::NUM_DOUBLES * G;
    gs.fs.si = grp_i + g;
    gs.gi = gs.fs.si + (size_t)FlexStore::NUM_INTS * G;

    // tensor memory: warp 0 allocates all columns; thread t of warp w owns TMEM lane 32 (w % 4) + t, columns
    // [(w / 4) COLS, (w / 4 + 1) COLS)
    __shared__ unsigned int cta_next;
    __shared__ unsigned int tm_base;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((unsigned int)__cvta_generic_to_shared(&tm_base)), "n"(TM_COLUMNS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) cta_next = (unsigned int)G;                       // (static first round, beamopt_lanes.cu)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    HomeTm hm;
    hm.base = tm_base + ((unsigned int)(32 * (w & 3)) << 16) + (unsigned int)((w >> 2) * COLS);

    LaneRegs<EPL> rg;
    lane_reset<EPL>(k, rg);
    FlexBeam fb;
    memset(&fb, 0, sizeof fb);
    Pass1Consts pc = {0.0, 0.0, 0.0, 0.0};
    long long b = -1;
    unsigned int my_next = (unsigned int)g;
    bool have = false, exhausted = false, resume = false;
    int t = 0, counter = 0, bad = 0;
    double best = INFINITY;
    float lossf = NAN;

    while (true) {
        bool fresh = false;
        if (!have && !exhausted) {
            long long nb = 0;
            if (b < 0) {                                            // static first turn (beamopt_lanes.cu)
                nb = (long long)blockIdx.x + (long long)gridDim.x * my_next;
                my_next += (unsigned int)G;
            } else {
                if (l == 0) nb = (long long)blockIdx.x + (long long)gridDim.x * atomicAdd(&cta_next, 1u);
                nb = __shfl_sync(gmask, nb, 0, LPB);
            }
            if (nb < B) {
                b = nb;
                have = true;
                t = 0; counter = 0; best = INFINITY; lossf = NAN;
                if (l == 0) {
                    int fnode[FLEX_MAXF];
                    double fval[FLEX_MAXF];
                    for (int j = 0; j < k.max_forces; ++j) {
                        fnode[j] = p.force_nodes[b * k.max_forces + j];
                        fval[j] = p.force_vals[b * k.max_forces + j];
                    }
                    const uint8_t *fx = p.fixed_uy + b * nn;
                    FlexBeam f0;
                    const int rc = flex_setup(k, p.L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode,
                                              fval, gs.fs, f0);
                    group_publish(f0, rc, gs);
                    group_table_init(gs);
                }
                __syncwarp(gmask);
                bad = group_fetch(k, p.L[b], gs, fb);
                pc = pass1_consts(fb);
                if (!bad) {
                    lane_init<EPL, HomeTm>(k, n, fb, gs, ls, l, rg, hm);    // {M0, Q0} staged in the scratch column
                    fresh = true;
                } else {
                    lane_reset<EPL>(k, rg);
                }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // warp-collective: new beams' statics into tensor memory, and the flexibility sums no pass has left behind
        if (__any_sync(0xffffffffu, fresh)) tm_commit<EPL>(hm, ls, fresh);
        if (__any_sync(0xffffffffu, fresh || resume)) {
            hm.wait_stores();
            lane_pass1<EPL, HomeTm>(rg, ls, pc, resume, hm, fresh || resume);
        }
        resume = false;

        const bool run = have && k.max_epochs > 0 && bad == 0;
        bool done = have && !run;
        float neg_step = 0.0f, bc2_sqrt = 1.0f;
        int rc = 0;
        if (run) {
            neg_step = __ldg(p.sched + 2 * t);
            bc2_sqrt = __ldg(p.sched + 2 * t + 1);
        }
        __syncwarp();
        if (run) lane_reduce(l, fb.m, ls, gs);
        __syncwarp();
        if (run) rc = group_solve(fb, gs, l);
        __syncwarp();
        // this epoch may be the beam's last: the pass parks the inertias it starts from for the record
        const bool stage_I = run && ((t + 1 >= k.max_epochs) || (k.early_stop && counter + 1 >= k.patience));
        if (__any_sync(0xffffffffu, run)) {
            hm.wait_stores();                                       // (last epoch's m, v)
            lane_pass<EPL, 1, TM_NBP, HomeTm>(k, n, rg, ls, gs, pc, fb.invLe, l, 0, neg_step, bc2_sqrt, stage_I, hm);
            if (have && !run) lane_reset<EPL>(k, rg);               // a rejected beam's record emits I_0
        }
        __syncwarp();
        if (run) {
            lossf = group_loss(k, n, ls, l);
            ++t;
            if (rc || !(lossf - lossf == 0.0f)) { bad = 1; done = true; }
            if (k.early_stop) {
                const double lv = (double)lossf;
                if (lv < best - k.tol) { best = lv; counter = 0; } else { ++counter; }
                if (counter >= k.patience) done = true;
            }
            if (t >= k.max_epochs) done = true;
        }
        // Record of a beam (once per beam): fields of the last analysed inertias (parked by the beam's last pass), the
        // inertias after the last Adam step; SC instances then copy the rows to the peers.
        const bool rec = have && done;
        const bool fields = (t > 0) && (bad == 0);
        const long long row = p.row0 + b;
        if (__any_sync(0xffffffffu, rec))
            lane_emit_forces<EPL, HomeTm>(n, rg, ls, gs, fb.invLe, l, fields, p.shear + row * n, p.moment + row * n, hm, rec);
        if (rec) {
            if (l == 0) {
                const ParkedInertia parked = {reinterpret_cast<const float *>(ls.scr), ls.ls};
                group_emit_displacements(k, fb, gs, fields, parked, p.defl + row * nn, p.rot + row * nn);
                p.epochs[row] = t;
                p.loss[row] = lossf;
                p.status[row] = bad;
            }
            lane_emit_inertias<EPL>(n, rg, l, p.I_values + row * n);
            if (SC && p.dest.nd > 1) {                              // dataset gather: every lane re-reads rows other lanes wrote
                __syncwarp(gmask);
                lane_copy_record(n, nn, l, p.dest, row, row, true);
            }
            __syncwarp(gmask);
            have = false;
        } else if (stage_I) {
            resume = true;                                          // the beam goes on: the sums the parked inertias displaced
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "n"(TM_COLUMNS) : "memory");
    }
}

// The tensor-memory instances exist for the 13-slot discretisations (n <= 104, the reference's n = 100 with compile-time
// element count); they pay once a batch needs more beams per SM than the register / shared-memory instance holds.
bool lanes_tm_supported(int epl, int num_cases) { return num_cases == 1 && epl == 13; }

int lanes_tm_plan(int64_t B, int sms, int smem_optin, LanesPlan *pl)
{
    const size_t per_group = tm_group_bytes();
    const int T = TM_THREADS;
    if (per_group * (T / LPB) > (size_t)smem_optin) return -2;
    pl->tm = 1;
    pl->threads = T;
    pl->smem_bytes = per_group * (T / LPB);
    const long per_cta = T / LPB;
    long want = (long)((B + per_cta - 1) / per_cta);
    pl->blocks = (int)(want < sms ? want : sms);
    if (pl->blocks < 1) pl->blocks = 1;
    if (pl->blocks < sms && B > pl->blocks) pl->blocks = (int)(B < sms ? B : sms);
    return 0;
}

template <int EPL, int NFIX, bool SC>
static cudaError_t launch_tm(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream)
{
    auto kern = beamopt_lanes_tm_kernel<EPL, NFIX, TM_THREADS, SC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<pl.blocks, pl.threads, pl.smem_bytes, stream>>>(k, B, p);
    return cudaGetLastError();
}

cudaError_t lanes_tm_launch(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream)
{
    const bool sc = p.dest.nd > 1 || getenv("OPS_FORCE_SC") != nullptr;
    if (pl.epl != 13 || pl.num_cases != 1 || pl.threads != TM_THREADS) return cudaErrorInvalidValue;
    if (pl.nfix == 100) return sc ? launch_tm<13, 100, true>(k, B, p, pl, stream) : launch_tm<13, 100, false>(k, B, p, pl, stream);
#ifdef OPS_LANES_DEV
    return cudaErrorInvalidValue;
#else
    return sc ? launch_tm<13, 0, true>(k, B, p, pl, stream) : launch_tm<13, 0, false>(k, B, p, pl, stream);
#endif
}

}  // namespace ops
