// sm_100a kernel of the production iteration: eight lanes per beam, persistent groups.
//
// One CTA per SM; a CTA is T/8 independent 8-lane groups, four per warp.  A group that finishes its
// beam (early stop, SingleCore:211-219, or max_e epochs) pulls the next beam index from a global
// counter, so ragged stopping never idles a group.  Phases of a group are separated by
// __syncwarp(group mask); nothing crosses warps, nothing but the beam's inputs (read once) and its
// record (written once) crosses HBM.  Arithmetic: beamopt_lanes.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (one IEEE rounding per written op).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "beamopt_internal.cuh"
#include "beamopt_lanes.cuh"

namespace ops {

using namespace lanes;

#ifndef OPS_LANES_MAXT
#define OPS_LANES_MAXT 320
#endif
constexpr int LANES_MAX_THREADS = OPS_LANES_MAXT;
// Many-round batches of the reference's discretisation: 12 warps per SM.  An iteration of a CTA costs a fixed
// latency plus a part that grows with the resident warps (profiles/): 48 beams per SM and round beat 40 once a
// batch runs several full rounds (+5 % at 6 rounds), while the 10 000-beam batch (1.7 rounds of 40) is faster at 320.
#ifndef OPS_LANES_BIGT
#define OPS_LANES_BIGT 384
#endif
constexpr int LANES_BIG_THREADS = OPS_LANES_BIGT;
constexpr int LANES_BIG_MIN_ROUNDS = 3;

// synchronisation of the NC groups of a team: one warp (NC <= 4) or 8 NC / 32 whole warps (named barrier)
template <int NC>
__device__ __forceinline__ void team_sync(unsigned team_mask, int barrier_id)
{
    if (NC == 1) return;
    if (NC * LPB <= 32) __syncwarp(team_mask);
    else asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "n"(NC * LPB) : "memory");
}

// Record of a beam (once per beam): fields of the last analysed inertias (parked by the beam's last pass), the
// inertias after the last Adam step; SC instances then copy the rows to the peers.
#define OPS_RECORD_PATH                                                                                              \
    {                                                                                                                \
        const bool fields = (t > 0) && (bad == 0);                                                                   \
        const long long row = p.row0 + b, rowc = row * NC + case_id; /* dataset rows of the beam / its load case */  \
        lane_emit_forces<EPL>(n, rg, ls, gs, fb.invLe, l, fields, p.shear + rowc * n, p.moment + rowc * n);          \
        if (l == 0) {                                                                                                \
            if constexpr (NC == 1) {                                                                                 \
                const ParkedInertia parked = {reinterpret_cast<const float *>(ls.scr), ls.ls};                       \
                group_emit_displacements(k, fb, gs, fields, parked, p.defl + rowc * nn, p.rot + rowc * nn);          \
            } else {                                                                                                 \
                group_emit_displacements(k, fb, gs, fields,                                                          \
                                         [&](int e) { return (double)team_inertia<NC>(ls, 0, case_id, e, 0); },      \
                                         p.defl + rowc * nn, p.rot + rowc * nn);                                     \
            }                                                                                                        \
            if (case_id == 0) {                                                                                      \
                p.epochs[row] = t;                                                                                   \
                p.loss[row] = lossf;                                                                                 \
                p.status[row] = bad;                                                                                 \
            }                                                                                                        \
        }                                                                                                            \
        if constexpr (NC == 1) lane_emit_inertias<EPL>(n, rg, l, p.I_values + row * n);                              \
        else if (case_id == 0) {                                                                                     \
            _Pragma("unroll") for (int kk = 0; kk < EPL; ++kk)                                                       \
                if (LPB * kk + l < n) p.I_values[row * n + LPB * kk + l] = team_inertia<NC>(ls, l, 0, LPB * kk + l, t > 0 ? 3 : 0); \
        }                                                                                                            \
        if (SC && p.dest.nd > 1) { /* dataset gather: every lane re-reads rows other lanes wrote */                  \
            __syncwarp(gmask);                                                                                       \
            lane_copy_record(n, nn, l, p.dest, row, rowc, case_id == 0);                                             \
        }                                                                                                            \
        __syncwarp(gmask);                                                                                           \
        have = false;                                                                                                \
    }
// SC: instance with the in-kernel dataset gather (the peers' copies of a finished beam's rows).  A separate
// instance because the mere presence of that cold code -- wherever it is placed -- costs the 320-thread instance
// 4-7 % on 10 000 beams: same epoch-loop instruction count, but ptxas orders the loop differently and ncu shows
// 10 % more fixed-latency `wait` stalls (profiles/r01_v6_ab_scatter.txt, r01_v6_scatter_instance_ncu.md).
// TFIX: compile-time CTA size (0 = blockDim.x).  With it every shared-memory column stride is an immediate
// and the [slot][thread] addressing costs no integer instructions.
#ifdef OPS_LANES_MAXNREG
#define OPS_LANES_BOUNDS __maxnreg__(OPS_LANES_MAXNREG)      /* A/B knob: register cap instead of the thread bound */
#else
#define OPS_LANES_BOUNDS __launch_bounds__(TFIX ? TFIX : LANES_MAX_THREADS, 1)
#endif
template <int EPL, int NFIX, int NC, int TFIX, bool SC>
__global__ void OPS_LANES_BOUNDS
beamopt_lanes_kernel(const BeamConsts k, const long long B, const OptPtrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NBK = lanes::NBP;                                // slot pairs per stage-major batch
    const int T = TFIX ? TFIX : (int)blockDim.x, G = T / LPB;
    const int tid = threadIdx.x, l = tid & (LPB - 1), g = tid / LPB;
    const unsigned gmask = 0xffu << (tid & 24);
    const int case_id = g % NC;                                    // load case of this group
    const unsigned team_mask = (NC * LPB >= 32) ? 0xffffffffu
                                                : (((1u << (NC * LPB)) - 1u) << ((tid & 31) / (NC * LPB) * (NC * LPB)));
    const int barrier_id = 1 + g / NC;                             // named barrier of the team (NC = 8)
    const int n = NFIX ? NFIX : k.n;
    const int nn = n + 1;

    double *lane_d = reinterpret_cast<double *>(smem_raw);
    double *tab_d = lane_d + (size_t)lane_doubles(EPL, NC) * T;
    const int GS = group_stride(G);
    double *grp_d = tab_d + (size_t)TAB_SLOTS * G;
    int *grp_i = reinterpret_cast<int *>(grp_d + (size_t)GROUP_DOUBLES * GS);
    LaneStore ls;
    ls.ls = T;
    ls.mq = reinterpret_cast<Pair *>(lane_d) + tid;
    ls.scr = lane_d + (size_t)2 * EPL * T + tid;
    ls.xc = reinterpret_cast<PairF *>(lane_d + (size_t)(2 * EPL + SCR_SLOTS) * T) + tid;
    ls.xb = lane_d + (size_t)(2 * EPL + SCR_SLOTS + EPL) * T + tid;
    GroupStore gs;
    gs.gs = GS;
    gs.tab = tab_d + (size_t)TAB_SLOTS * g;
    gs.fs.sd = grp_d + g;
    gs.fs.stride = GS;
    gs.gd = gs.fs.sd + (size_t)FlexStore::NUM_DOUBLES * GS;
    gs.fs.si = grp_i + g;
    gs.gi = gs.fs.si + (size_t)FlexStore::NUM_INTS * GS;
    int *team_gi = gs.gi - case_id;                                // the case-0 group's ints

    // Work distribution: beam b belongs to CTA b mod gridDim.x, and the groups of a CTA take the CTA's
    // beams in order from a counter in shared memory.  Every SM gets the same number of beams (+-1), so
    // the last, partially filled round runs with few groups on EVERY SM (short iterations) instead of
    // full on some SMs and empty on others; ragged stopping is still absorbed inside the CTA.
    // Which group takes which beam.  The FIRST beam of a group is static (group g / team g takes the CTA's g-th beam),
    // the others come from a counter that starts behind that first round.  A launch that does not fill every group --
    // the last chunk of a pipelined run, a small batch -- then runs on the LOWEST warps, i.e. evenly spread over the four
    // schedulers; a race for the counter hands the beams to whichever warps arrive first, often several of one
    // scheduler, and the partial round takes as long as a full one (16 beams per SM: 4.4 us per epoch against 3.0 us,
    // 27.6: 4.5 against 3.7; profiles/r02_static_first_round.txt).  Later rounds of the same launch stay on the counter:
    // measured faster than a static order there (10 000 beams: 4.78 ms against 4.87 ms).
    __shared__ unsigned int cta_next;
    if (tid == 0) cta_next = (unsigned int)(G / NC);
    __syncthreads();
    unsigned int my_next = (unsigned int)(g / NC);                 // CTA-local index of the group's next static beam
#ifdef OPS_LANES_STAGGER_NS
    // A/B knob: warps start their first epoch OPS_LANES_STAGGER_NS apart (per scheduler slot w / 4, plus a quarter of
    // that per scheduler), so that the resident warps are not all in the same phase of the epoch at the same time
    {
        const unsigned w = tid >> 5;
        unsigned ns = (w >> 2) * OPS_LANES_STAGGER_NS + (w & 3) * (OPS_LANES_STAGGER_NS / 4);
        while (ns > 0) { const unsigned step = ns > 1000000u ? 1000000u : ns; __nanosleep(step); ns -= step; }
    }
#endif

    LaneRegs<EPL> rg;
    FlexBeam fb;
    Pass1Consts pc = {0.0, 0.0, 0.0, 0.0};
    long long b = -1;
    bool have = false, exhausted = false;
    int t = 0, counter = 0, bad = 0;
    double best = INFINITY;
    float lossf = NAN;

    while (true) {
        if (!have && !exhausted) {
            long long nb = 0;
            const bool fixed_turn = b < 0;                         // the group's first beam
            if (fixed_turn) {
                nb = (long long)blockIdx.x + (long long)gridDim.x * my_next;
                my_next += (unsigned int)(G / NC);
            } else if (NC == 1) {
                if (l == 0) nb = (long long)blockIdx.x + (long long)gridDim.x * atomicAdd(&cta_next, 1u);
                nb = __shfl_sync(gmask, nb, 0, LPB);
            } else {
                if (case_id == 0 && l == 0) {
                    nb = (long long)blockIdx.x + (long long)gridDim.x * atomicAdd(&cta_next, 1u);
                    team_gi[4 * GS] = (int)(nb & 0xffffffffLL);
                    team_gi[5 * GS] = (int)(nb >> 32);
                }
                team_sync<NC>(team_mask, barrier_id);
                nb = ((long long)team_gi[5 * GS] << 32) | (unsigned int)team_gi[4 * GS];
                team_sync<NC>(team_mask, barrier_id);
            }
            if (nb < B) {
                b = nb;
                have = true;
                t = 0; counter = 0; best = INFINITY; lossf = NAN;
                const long long bc = b * NC + case_id;             // row of this group's load case
                if (l == 0) {
                    int fnode[FLEX_MAXF];
                    double fval[FLEX_MAXF];
                    for (int j = 0; j < k.max_forces; ++j) {
                        fnode[j] = p.force_nodes[bc * k.max_forces + j];
                        fval[j] = p.force_vals[bc * k.max_forces + j];
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
                    lane_init<EPL>(k, n, fb, gs, ls, l, rg);
                    if (NC == 1) lane_pass1<EPL>(rg, ls, pc, false);
                } else {
                    lane_reset<EPL>(k, rg);
                }
                if constexpr (NC > 1) {
                    team_init<EPL, NC>(k, n, ls, l, case_id, rg);   // (bad beams too: the record emits I_0)
                    team_sync<NC>(team_mask, barrier_id);           // the owners' rows are read by every group of the team
                }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // One epoch.  The phases are separated by FULL-warp barriers that every lane reaches (a group without
        // a running beam only keeps the appointment): a barrier on the group's own 8-lane mask is a
        // non-uniform mask, which costs a MATCH / REDUX / VOTE sequence per barrier instead of nothing.
        const bool run = have && k.max_epochs > 0 && bad == 0;
        bool done = have && !run;
        float neg_step = 0.0f, bc2_sqrt = 1.0f;
        int rc = 0;
        if (run) {
            neg_step = __ldg(p.sched + 2 * t);
            bc2_sqrt = __ldg(p.sched + 2 * t + 1);
        }
        Pair mqk[NC > 1 ? EPL : 1];                                 // teams: {M0, Q0} from the sums to the forces, in registers
        if constexpr (NC > 1) {
            if (run) team_pass1<EPL, NC>(rg, ls, pc, case_id, mqk); // P1: this case's sums from the team's inertias
        }
        __syncwarp();
        if (run) lane_reduce(l, fb.m, ls, gs);
        __syncwarp();
        if (run) rc = group_solve(fb, gs, l);
        __syncwarp();
        // this epoch may be the beam's last: the pass parks the inertias it starts from for the record
        const bool stage_I = run && ((t + 1 >= k.max_epochs) || (k.early_stop && counter + 1 >= k.patience));
        if constexpr (NC == 1) {
            if (run) lane_pass<EPL, NC, NBK>(k, n, rg, ls, gs, pc, fb.invLe, l, case_id, neg_step, bc2_sqrt, stage_I);
            __syncwarp();
        } else {
            if (run) lane_case_squares<EPL>(rg, ls, gs, fb.invLe, mqk);
            if (NC * LPB <= 32) __syncwarp();
            else if (run) team_sync<NC>(team_mask, barrier_id);     // whole warps belong to one team: uniform
            if (run) team_owner_update<EPL, NC>(k, rg, ls, case_id, neg_step, bc2_sqrt);     // P2
            if (NC * LPB <= 32) __syncwarp();
            else if (run) team_sync<NC>(team_mask, barrier_id);
            if (run) team_loss_sums<EPL, NC>(n, ls, l, case_id);                              // P3
            __syncwarp();
        }
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
        if (have && done) {
            OPS_RECORD_PATH
        } else if (NC == 1 && stage_I) {
            lane_pass1<EPL>(rg, ls, pc, true);                      // the beam goes on: the sums the parked inertias displaced
        }
    }
}
#undef OPS_RECORD_PATH

bool lanes_supported(const BeamConsts &k, int num_cases)
{
    const bool cases_ok = num_cases == 1 || num_cases == 2 || num_cases == 4 || num_cases == 8;
    return cases_ok && k.n >= 1 && k.n <= 8 * 21 && k.max_forces <= FLEX_MAXF && (num_cases == 1 || k.n <= 104);
}

static int pick_epl(int n)
{
    if (n <= 32) return 4;
    if (n <= 64) return 8;
    if (n <= 104) return 13;
    return 21;
}

int lanes_plan(const BeamConsts &k, int num_cases, int64_t B, int sms, int smem_optin, LanesPlan *pl)
{
    pl->epl = pick_epl(k.n);
    pl->nfix = (k.n == 100) ? 100 : 0;
    pl->num_cases = num_cases;
    pl->tm = 0;
    // Tensor-memory instance (beamopt_lanes_tm.cu): 64 beams per SM and round instead of 40.  Measured (profiles/
    // r02_tensor_memory_instance.md): its epoch is slower per warp (128 registers), so it wins exactly where it turns two
    // rounds of this instance into one -- fixed epoch counts with 52..64 beams per SM (+6 % .. +19 %).
    // OPS_LANES_TM = 0 / 1 forces the choice (profiling knob; same results either way).
    {
        const char *tm_env = getenv("OPS_LANES_TM");
        const int64_t per_sm = (B + sms - 1) / sms;
        const bool want_tm = tm_env ? atoi(tm_env) != 0 : (!k.early_stop && per_sm >= 52 && per_sm <= 64);
        if (want_tm && lanes_tm_supported(pl->epl, num_cases) && lanes_tm_plan(B, sms, smem_optin, pl) == 0) return 0;
    }
    const size_t per_group = (size_t)LPB * 8 * lane_doubles(pl->epl, num_cases) + (size_t)(TAB_SLOTS + GROUP_DOUBLES) * 8 +
                             (size_t)GROUP_INTS * 4;
    int groups = (int)((size_t)smem_optin / per_group);
    int T = groups * LPB / 32 * 32;
    int cap = LANES_MAX_THREADS;
    if (num_cases == 1 && pl->nfix == 100 && LANES_BIG_THREADS > cap &&
        B >= (int64_t)sms * (LANES_BIG_THREADS / LPB) * LANES_BIG_MIN_ROUNDS)
        cap = LANES_BIG_THREADS;
    if (T > cap) T = cap;
    const char *thr_env = getenv("OPS_LANES_THREADS");            // profiling knob
    if (thr_env && atoi(thr_env) >= 32 && atoi(thr_env) <= T) T = atoi(thr_env) / 32 * 32;
    const int team_threads = num_cases * LPB;                     // whole teams per CTA (and whole warps per team)
    const int quantum = team_threads > 32 ? team_threads : 32;
    T = T / quantum * quantum;
    while (T >= quantum && cta_smem_bytes(T / LPB, pl->epl, num_cases) > (size_t)smem_optin) T -= quantum;   // (the padded stride)
    if (T < quantum) return -2;
    pl->threads = T;
    pl->smem_bytes = cta_smem_bytes(T / LPB, pl->epl, num_cases);
    const long per_cta = T / team_threads;
    long want = (long)((B + per_cta - 1) / per_cta);
    pl->blocks = (int)(want < sms ? want : sms);
    if (pl->blocks < 1) pl->blocks = 1;
    // spread a batch that does not fill every CTA evenly over all SMs
    if (pl->blocks < sms && B > pl->blocks) {
        pl->blocks = (int)(B < sms ? B : sms);
    }
    return 0;
}

template <int EPL, int NFIX, int NC, int TFIX = 0, bool SC = false>
static cudaError_t launch_instance(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl,
                                   cudaStream_t stream)
{
    auto kern = beamopt_lanes_kernel<EPL, NFIX, NC, TFIX, SC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<pl.blocks, pl.threads, pl.smem_bytes, stream>>>(k, B, p);
    return cudaGetLastError();
}

// instances with the in-kernel dataset gather: every single-case instance, and the 13-slot multi-case ones
bool lanes_scatter_supported(const LanesPlan &pl) { return pl.num_cases == 1 || pl.epl == 13; }

#ifdef OPS_LANES_DEV
// development build (kernel iteration on the reference's discretisation only: compiles in a fraction of the time)
cudaError_t lanes_launch(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream)
{
    if (pl.tm) return lanes_tm_launch(k, B, p, pl, stream);
    if (pl.num_cases != 1 || pl.nfix != 100 || p.dest.nd > 1) return cudaErrorInvalidValue;
    if (pl.threads == LANES_MAX_THREADS) return launch_instance<13, 100, 1, LANES_MAX_THREADS, false>(k, B, p, pl, stream);
    if (pl.threads == LANES_BIG_THREADS) return launch_instance<13, 100, 1, LANES_BIG_THREADS, false>(k, B, p, pl, stream);
    return launch_instance<13, 100, 1, 0, false>(k, B, p, pl, stream);
}
#else
template <bool SC>
static cudaError_t launch_single_case(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl,
                                      cudaStream_t stream)
{
    if (pl.nfix == 100 && pl.threads == LANES_MAX_THREADS)
        return launch_instance<13, 100, 1, LANES_MAX_THREADS, SC>(k, B, p, pl, stream);  // the reference's discretisation
    if (pl.nfix == 100 && pl.threads == LANES_BIG_THREADS)
        return launch_instance<13, 100, 1, LANES_BIG_THREADS, SC>(k, B, p, pl, stream);
    if (pl.nfix == 100) return launch_instance<13, 100, 1, 0, SC>(k, B, p, pl, stream);
    switch (pl.epl) {
    case 4: return launch_instance<4, 0, 1, 0, SC>(k, B, p, pl, stream);
    case 8: return launch_instance<8, 0, 1, 0, SC>(k, B, p, pl, stream);
    case 13: return launch_instance<13, 0, 1, 0, SC>(k, B, p, pl, stream);
    default: return launch_instance<21, 0, 1, 0, SC>(k, B, p, pl, stream);
    }
}

cudaError_t lanes_launch(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream)
{
    if (pl.tm) return lanes_tm_launch(k, B, p, pl, stream);
    const bool sc = p.dest.nd > 1 || getenv("OPS_FORCE_SC") != nullptr;      // (the knob: profiling of the scatter instances on one GPU)
    if (sc && !lanes_scatter_supported(pl)) return cudaErrorInvalidValue;
    if (pl.num_cases > 1) {
        // load cases sharing one inertia vector: the reference's discretisation with compile-time element count and CTA
        // size (BASELINE config 4), and the generic <= 104-element instances
        if (pl.nfix == 100 && pl.threads == LANES_MAX_THREADS) {
            switch (pl.num_cases) {
            case 2: return sc ? launch_instance<13, 100, 2, LANES_MAX_THREADS, true>(k, B, p, pl, stream)
                              : launch_instance<13, 100, 2, LANES_MAX_THREADS, false>(k, B, p, pl, stream);
            case 4: return sc ? launch_instance<13, 100, 4, LANES_MAX_THREADS, true>(k, B, p, pl, stream)
                              : launch_instance<13, 100, 4, LANES_MAX_THREADS, false>(k, B, p, pl, stream);
            case 8: return sc ? launch_instance<13, 100, 8, LANES_MAX_THREADS, true>(k, B, p, pl, stream)
                              : launch_instance<13, 100, 8, LANES_MAX_THREADS, false>(k, B, p, pl, stream);
            default: return cudaErrorInvalidValue;
            }
        }
        switch (pl.num_cases * 100 + pl.epl) {
        case 204: return launch_instance<4, 0, 2>(k, B, p, pl, stream);
        case 208: return launch_instance<8, 0, 2>(k, B, p, pl, stream);
        case 213: return sc ? launch_instance<13, 0, 2, 0, true>(k, B, p, pl, stream) : launch_instance<13, 0, 2>(k, B, p, pl, stream);
        case 404: return launch_instance<4, 0, 4>(k, B, p, pl, stream);
        case 408: return launch_instance<8, 0, 4>(k, B, p, pl, stream);
        case 413: return sc ? launch_instance<13, 0, 4, 0, true>(k, B, p, pl, stream) : launch_instance<13, 0, 4>(k, B, p, pl, stream);
        case 804: return launch_instance<4, 0, 8>(k, B, p, pl, stream);
        case 808: return launch_instance<8, 0, 8>(k, B, p, pl, stream);
        case 813: return sc ? launch_instance<13, 0, 8, 0, true>(k, B, p, pl, stream) : launch_instance<13, 0, 8>(k, B, p, pl, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    return sc ? launch_single_case<true>(k, B, p, pl, stream) : launch_single_case<false>(k, B, p, pl, stream);
}
#endif

// ---------------------------------------------------------------------------------------------
// self test of fastmath.cuh against the compiler's IEEE operators (ops_fastmath_selftest)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int lcg(unsigned long long &s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (unsigned int)(s >> 32);
}

// random float with a biased exponent in [elo, ehi] and a random mantissa; sign positive
__device__ __forceinline__ float rand_float(unsigned long long &s, int elo, int ehi)
{
    const unsigned int m = lcg(s) & 0x7fffffu;
    const unsigned int e = (unsigned int)elo + lcg(s) % (unsigned int)(ehi - elo + 1);
    return __uint_as_float((e << 23) | m);
}

__global__ void fastmath_selftest_kernel(int per_thread, unsigned long long seed, unsigned long long *out)
{
    unsigned long long s = seed + 0x9E3779B97F4A7C15ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad_div = 0, bad_sqrt = 0, bad_rcp = 0, bad_rsq = 0;
    double worst = 0.0, worst_seed = 0.0;
    for (int i = 0; i < per_thread; ++i) {
        // division: b in [2^-60, 2^60], a zero or in [2^-60, 2^60]  (quotient within [2^-120, 2^120])
        const float b = rand_float(s, 67, 187);
        float a = rand_float(s, 67, 187);
        if ((lcg(s) & 1023u) == 0) a = 0.0f;
        if (lcg(s) & 1u) a = -a;
        const float r = fm::rcp_r(b);
        // (a signed zero numerator keeps the magnitude but not the sign: every numerator of the kernel is
        //  either non-negative or only ever added to a non-zero value)
        const float qf = fm::div_r(a, b, r), qi = a / b;
        if (__float_as_uint(qf) != __float_as_uint(qi) && !(qf == 0.0f && qi == 0.0f)) ++bad_div;
        if (__float_as_uint(fm::rcp_f(b)) != __float_as_uint(1.0f / b)) ++bad_rcp;
        // square root: x in [2^-101, 2^127]
        const float x = rand_float(s, 26, 254);
        if (__float_as_uint(fm::sqrt_f(x)) != __float_as_uint(sqrtf(x))) ++bad_sqrt;
        // fp64 reciprocal of an fp32-valued argument in [2^-40, 2^40]
        const double xd = (double)rand_float(s, 87, 167);
        const double e = fabs(fm::rcp64(xd) * xd - 1.0);
        worst = e > worst ? e : worst;
        const double e0 = fabs(fm::rcp64_a(xd) * xd - 1.0);          // the bare MUFU.RCP64H seed
        worst_seed = e0 > worst_seed ? e0 : worst_seed;
        // packed fp32 pairs (FFMA2 / FMUL2 / FADD2, fastmath.cuh): each half must carry the bits of the scalar operation,
        // and the un-fused product-then-sum must not have been contracted
        {
            const fm::F2 pa = fm::f2(a, -x), pb = fm::f2(b, qi), pc = fm::f2(x, a);
            const fm::F2 f = fm::fma2(pa, pb, pc), m2 = fm::mul2(pa, pb), s2 = fm::add2(pa, pc), ma = fm::mul2_add(pa, pb, pc);
            const fm::F2 fn = fm::fma2(fm::neg2(pa), pb, fm::splat(1.0f));
            bad_rsq += __float_as_uint(f.x) != __float_as_uint(fmaf(pa.x, pb.x, pc.x));
            bad_rsq += __float_as_uint(f.y) != __float_as_uint(fmaf(pa.y, pb.y, pc.y));
            bad_rsq += __float_as_uint(m2.x) != __float_as_uint(pa.x * pb.x);
            bad_rsq += __float_as_uint(m2.y) != __float_as_uint(pa.y * pb.y);
            bad_rsq += __float_as_uint(s2.x) != __float_as_uint(pa.x + pc.x);
            bad_rsq += __float_as_uint(s2.y) != __float_as_uint(pa.y + pc.y);
            bad_rsq += __float_as_uint(ma.x) != __float_as_uint(__fadd_rn(__fmul_rn(pa.x, pb.x), pc.x));
            bad_rsq += __float_as_uint(ma.y) != __float_as_uint(__fadd_rn(__fmul_rn(pa.y, pb.y), pc.y));
            bad_rsq += __float_as_uint(fn.x) != __float_as_uint(fmaf(-pa.x, pb.x, 1.0f));
            bad_rsq += __float_as_uint(fn.y) != __float_as_uint(fmaf(-pa.y, pb.y, 1.0f));
        }
    }
    atomicAdd(out + 0, bad_div);
    atomicAdd(out + 1, bad_sqrt);
    atomicAdd(out + 2, bad_rcp);
    atomicMax(out + 3, (unsigned long long)__double_as_longlong(worst));
    atomicAdd(out + 4, bad_rsq);
    atomicMax(out + 5, (unsigned long long)__double_as_longlong(worst_seed));
}

cudaError_t fastmath_selftest(long long samples, unsigned long long *host_out4, double *worst_rcp64, cudaStream_t stream)
{
    unsigned long long *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, 64);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d, 0, 64, stream);
    const int threads = 256, blocks = 592;
    int per_thread = (int)(samples / ((long long)threads * blocks)) + 1;
    fastmath_selftest_kernel<<<blocks, threads, 0, stream>>>(per_thread, 0x1234567ULL, d);
    e = cudaGetLastError();
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 64, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    double w, w0;
    memcpy(&w, &h[3], 8);
    memcpy(&w0, &h[5], 8);
    *worst_rcp64 = w;
    if (getenv("OPS_SELFTEST_VERBOSE")) printf("rcp64 seed max rel err %.3e, refined %.3e\n", w0, w);
    host_out4[0] = h[0]; host_out4[1] = h[1]; host_out4[2] = h[2] + h[4];     // reciprocal mismatches incl. 1/sqrt
    host_out4[3] = (unsigned long long)per_thread * threads * blocks;
    return cudaSuccess;
}

}  // namespace ops
