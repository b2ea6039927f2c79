// sm_100a kernels + C ABI (include/openpystruct_b200.h) of the beam moment-of-inertia optimiser.
//
// Execution model: persistent grid, ONE THREAD PER BEAM.  A lane that finishes its beam (early stop,
// SingleCore:211-219) pulls the next beam index from a global counter, so ragged stopping does not
// idle the warp.  The per-beam factor (5 doubles per node) and the fp32 optimiser state (I, m, v)
// live in shared memory laid out [slot][thread]; nothing but the beam's inputs (read once) and its
// results (written once) crosses HBM.  For discretisations whose state does not fit in shared
// memory the same code runs on a global-memory scratch (`workspace`) laid out the same way.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see beamopt_core.cuh).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/openpystruct_b200.h"
#include "beamopt_core.cuh"
#include "beamopt_flex.cuh"
#include "beamopt_internal.cuh"

#define OPS_VERSION "openpystruct_b200 0.1.0 sm_100a"

namespace ops {

// fixed-uy bitmask words per beam, laid out [word][thread]
struct MaskStore {
    uint32_t *w;
    long stride;
    OPS_HD bool operator()(int i) const
    {
        return (w[(long)(i >> 5) * stride] >> (i & 31)) & 1u;
    }
};

template <int MAXF, bool SMEM>
__global__ void __launch_bounds__(128) beamopt_kernel(const BeamConsts k, const long long B, const OptPtrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int n = k.n, nn = k.nn;
    const int mask_words = (nn + 31) >> 5;

    BeamStore st;
    MaskStore fixed;
    if (SMEM) {
        double *sd = reinterpret_cast<double *>(smem_raw);
        float *sf = reinterpret_cast<float *>(sd + (size_t)5 * nn * T);
        uint32_t *sm = reinterpret_cast<uint32_t *>(sf + (size_t)3 * n * T);
        st.d = sd + threadIdx.x;
        st.f = sf + threadIdx.x;
        st.stride = T;
        fixed.w = sm + threadIdx.x;
        fixed.stride = T;
    } else {
        const long total = (long)gridDim.x * T;
        const long gtid = (long)blockIdx.x * T + threadIdx.x;
        st.d = p.ws_d + gtid;
        st.f = p.ws_f + gtid;
        st.stride = total;
        fixed.w = p.ws_mask + gtid;
        fixed.stride = total;
    }

    BeamInputs<MAXF> in;
    long long b = -1;
    bool have = false, exhausted = false;
    int t = 0, counter = 0, bad = 0;
    double best = INFINITY;
    float lossf = NAN;

    while (true) {
        if (!have && !exhausted) {
            b = (long long)atomicAdd(p.counter, 1ULL);
            if (b < B) {
                have = true;
                t = 0; counter = 0; bad = 0; best = INFINITY; lossf = NAN;
                beam_geometry<MAXF>(k, p.L[b], in);
#pragma unroll
                for (int j = 0; j < MAXF; ++j) {
                    const bool on = j < k.max_forces;
                    in.fnode[j] = on ? p.force_nodes[b * k.max_forces + j] : -1;
                    in.fval[j] = on ? p.force_vals[b * k.max_forces + j] : 0.0;
                }
                for (int w = 0; w < mask_words; ++w) {
                    uint32_t bits = 0;
                    for (int i = 0; i < 32 && 32 * w + i < nn; ++i)
                        bits |= (p.fixed_uy[b * nn + 32 * w + i] ? 1u : 0u) << i;
                    fixed.w[(long)w * fixed.stride] = bits;
                }
                for (int e = 0; e < n; ++e) {
                    st.F(e) = k.I0f;
                    st.F(n + e) = 0.0f;
                    st.F(2 * n + e) = 0.0f;
                }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
            int rc = 0;
            bool done = (k.max_epochs <= 0);
            if (!done) {
                const float neg_step = __ldg(p.sched + 2 * t);
                const float bc2_sqrt = __ldg(p.sched + 2 * t + 1);
                lossf = beam_iteration<MAXF>(k, in, st, fixed, neg_step, bc2_sqrt, &rc);
                ++t;
                if (rc) { bad = 1; done = true; }
                if (k.early_stop) {
                    const double l = (double)lossf;
                    if (l < best - k.tol) { best = l; counter = 0; } else { ++counter; }
                    if (counter >= k.patience) done = true;
                }
                if (t >= k.max_epochs) done = true;
            }
            if (done) {
                for (int e = 0; e < n; ++e) {
                    p.I_values[b * n + e] = st.F(e);
                    const float *mv = st.pairf(5 * e);
                    p.moment[b * n + e] = t > 0 ? mv[0] : 0.0f;
                    p.shear[b * n + e] = t > 0 ? mv[1] : 0.0f;
                }
                for (int i = 0; i < nn; ++i) {
                    const bool z = (t == 0) || (k.zero_last_node && i == nn - 1);
                    p.defl[b * nn + i] = z ? 0.0 : st.D(5 * i + 3);
                    p.rot[b * nn + i] = z ? 0.0 : st.D(5 * i + 4);
                }
                p.epochs[b] = t;
                p.loss[b] = lossf;
                p.status[b] = bad;
                have = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Production kernel: three-moment iteration (beamopt_flex.cuh), one thread per beam, persistent
// lanes.  Per-beam state: fp32 I (double-buffered: the analysed and the updated vector), Adam m, v,
// the running torch.sum partials and the O(#supports) span store.  `lay` says which of the fp32
// arrays live in shared memory and which in an L2-resident global scratch laid out [e][thread].
// ---------------------------------------------------------------------------------------------
struct FlexLayout {
    int I_global, m_global, v_global;   // 0 = shared memory, 1 = global scratch
    int nacc;                           // running-sum slots per reduction (32, or 64 with the level cascade)
};

__global__ void __launch_bounds__(256) beamopt_flex_kernel(const BeamConsts k, const long long B, const OptPtrs p,
                                                           const FlexLayout lay)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int n = k.n, nn = k.nn;
    const long total = (long)gridDim.x * T;
    const long gtid = (long)blockIdx.x * T + threadIdx.x;

    // carve shared memory: doubles, then floats, then ints
    double *sd = reinterpret_cast<double *>(smem_raw);
    float *sf = reinterpret_cast<float *>(sd + (size_t)FlexStore::NUM_DOUBLES * T);
    FlexStore fs;
    fs.sd = sd + threadIdx.x;
    fs.stride = T;
    OptState os;
    float *gf = p.ws_f + gtid;          // global scratch arrays: which * n * total
    int gslot = 0;
    auto place = [&](int global, int count, float *&ptr, long &stride) {
        if (global) { ptr = gf + (long)gslot * n * total; stride = total; gslot += 1; (void)count; }
        else { ptr = sf + threadIdx.x; stride = T; sf += (size_t)count * T; }
    };
    place(lay.I_global, n, os.Icur, os.sI);
    place(lay.I_global, n, os.Inext, os.sIn);
    place(lay.m_global, n, os.m, os.sm);
    place(lay.v_global, n, os.v, os.sv);
    os.accd = sf + threadIdx.x; sf += (size_t)lay.nacc * T;
    os.accq = sf + threadIdx.x; sf += (size_t)lay.nacc * T;
    os.sacc = T;
    fs.si = reinterpret_cast<int *>(sf) + threadIdx.x;

    FlexBeam fb;
    long long b = -1;
    bool have = false, exhausted = false;
    int t = 0, counter = 0, bad = 0;
    double best = INFINITY;
    float lossf = NAN;

    while (true) {
        if (!have && !exhausted) {
            b = (long long)atomicAdd(p.counter, 1ULL);
            if (b < B) {
                have = true;
                t = 0; counter = 0; best = INFINITY; lossf = NAN;
                int fnode[FLEX_MAXF];
                double fval[FLEX_MAXF];
                for (int j = 0; j < k.max_forces; ++j) {
                    fnode[j] = p.force_nodes[b * k.max_forces + j];
                    fval[j] = p.force_vals[b * k.max_forces + j];
                }
                const uint8_t *fx = p.fixed_uy + b * nn;
                bad = flex_setup(k, p.L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, fs, fb);
                for (int e = 0; e < n; ++e) {
                    os.Icur[(long)e * os.sI] = k.I0f;
                    os.Inext[(long)e * os.sIn] = k.I0f;
                    os.m[(long)e * os.sm] = 0.0f;
                    os.v[(long)e * os.sv] = 0.0f;
                }
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
            bool done = (k.max_epochs <= 0) || (bad != 0);
            if (!done) {
                int rc = 0;
                const float neg_step = __ldg(p.sched + 2 * t);
                const float bc2_sqrt = __ldg(p.sched + 2 * t + 1);
                lossf = flex_iteration(k, fb, fs, os, neg_step, bc2_sqrt, &rc);
                ++t;
                // ping-pong: Icur <- updated inertias, Inext <- the vector that was just analysed
                { float *tp = os.Icur; os.Icur = os.Inext; os.Inext = tp; long ts = os.sI; os.sI = os.sIn; os.sIn = ts; }
                if (rc || !(lossf - lossf == 0.0f)) { bad = 1; done = true; }
                if (k.early_stop) {
                    const double l = (double)lossf;
                    if (l < best - k.tol) { best = l; counter = 0; } else { ++counter; }
                    if (counter >= k.patience) done = true;
                }
                if (t >= k.max_epochs) done = true;
            }
            if (done) {
                const bool fields = (t > 0) && (bad == 0);
                for (int e = 0; e < n; ++e) p.I_values[b * n + e] = os.Icur[(long)e * os.sI];
                if (fields) {
                    flex_forces_march(k, fb, fs, [&](int e, double V, double M) {
                        p.shear[b * n + e] = (float)V;
                        p.moment[b * n + e] = (float)M;
                    });
                    flex_deflections_march(k, fb, fs, [&](int e) { return (double)os.Inext[(long)e * os.sIn]; },
                                           [&](int i, double u, double th) {
                                               const bool z = k.zero_last_node && i == nn - 1;
                                               p.defl[b * nn + i] = z ? 0.0 : u;
                                               p.rot[b * nn + i] = z ? 0.0 : th;
                                           });
                } else {
                    for (int e = 0; e < n; ++e) { p.shear[b * n + e] = 0.0f; p.moment[b * n + e] = 0.0f; }
                    for (int i = 0; i < nn; ++i) { p.defl[b * nn + i] = 0.0; p.rot[b * nn + i] = 0.0; }
                }
                p.epochs[b] = t;
                p.loss[b] = lossf;
                p.status[b] = bad;
                have = false;
            }
        }
    }
}

struct SolvePtrs {
    const uint8_t *fixed_uy;
    const int32_t *force_nodes;
    const double *force_vals;
    const double *L;
    const double *I;
    double *defl, *rot, *shear, *moment;
    int32_t *status;
};

// One solve per beam, thread per beam, factor in shared memory [slot][thread].
template <int MAXF>
__global__ void __launch_bounds__(128) beamsolve_kernel(const BeamConsts k, const long long B, const SolvePtrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int n = k.n, nn = k.nn;
    const long long b = (long long)blockIdx.x * T + threadIdx.x;
    if (b >= B) return;
    BeamStore st;
    st.d = reinterpret_cast<double *>(smem_raw) + threadIdx.x;
    st.f = nullptr;
    st.stride = T;
    BeamInputs<MAXF> in;
    beam_geometry<MAXF>(k, p.L[b], in);
#pragma unroll
    for (int j = 0; j < MAXF; ++j) {
        const bool on = j < k.max_forces;
        in.fnode[j] = on ? p.force_nodes[b * k.max_forces + j] : -1;
        in.fval[j] = on ? p.force_vals[b * k.max_forces + j] : 0.0;
    }
    const uint8_t *fx = p.fixed_uy + b * nn;
    const double *Ib = p.I + b * n;
    auto fixed = [&](int i) { return fx[i] != 0; };
    auto inertia = [&](int e) { return Ib[e]; };
    int rc = factor_forward<MAXF>(k, in, st, fixed, inertia);
    rc |= solve_backward<MAXF>(k, in, st, fixed, inertia, [&](int e, double V, double M) {
        p.shear[b * n + e] = V;
        p.moment[b * n + e] = M;
    });
    for (int i = 0; i < nn; ++i) {
        p.defl[b * nn + i] = st.D(5 * i + 3);
        p.rot[b * nn + i] = st.D(5 * i + 4);
    }
    p.status[b] = rc;
}

// One three-moment solve per beam (FP64 outputs); span store in shared memory.
__global__ void __launch_bounds__(128) beamsolve_flex_kernel(const BeamConsts k, const long long B, const SolvePtrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int n = k.n, nn = k.nn;
    const long long b = (long long)blockIdx.x * T + threadIdx.x;
    if (b >= B) return;
    FlexStore fs;
    fs.sd = reinterpret_cast<double *>(smem_raw) + threadIdx.x;
    fs.si = reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + (size_t)FlexStore::NUM_DOUBLES * T) + threadIdx.x;
    fs.stride = T;
    FlexBeam fb;
    int fnode[FLEX_MAXF];
    double fval[FLEX_MAXF];
    for (int j = 0; j < k.max_forces; ++j) {
        fnode[j] = p.force_nodes[b * k.max_forces + j];
        fval[j] = p.force_vals[b * k.max_forces + j];
    }
    const uint8_t *fx = p.fixed_uy + b * nn;
    const double *Ib = p.I + b * n;
    int bad = flex_setup(k, p.L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, fs, fb);
    if (!bad) {
        bad = flex_support_moments(k, fb, fs, [&](int e) { return Ib[e]; });
        flex_forces_march(k, fb, fs, [&](int e, double V, double M) {
            p.shear[b * n + e] = V;
            p.moment[b * n + e] = M;
        });
        flex_deflections_march(k, fb, fs, [&](int e) { return Ib[e]; }, [&](int i, double u, double th) {
            p.defl[b * nn + i] = u;
            p.rot[b * nn + i] = th;
        });
    }
    p.status[b] = bad;
}

static int make_consts(const OpsBeamOptParams *p, BeamConsts *k)
{
    if (!p || p->struct_size != (int32_t)sizeof(OpsBeamOptParams)) return OPS_E_BADARG;
    if (p->num_nodes < 2 || p->num_cases < 1 || p->max_forces < 0 || p->max_epochs < 0) return OPS_E_BADARG;
    if (p->num_cases != 1 && !(p->solver == OPS_SOLVER_THREE_MOMENT && p->num_nodes <= 105 &&
                                (p->num_cases == 2 || p->num_cases == 4 || p->num_cases == 8)))
        return OPS_E_UNSUPP;       // shared-I load cases: 2, 4 or 8 per beam, lanes kernel only
    if (p->max_forces > 8) return OPS_E_UNSUPP;
    if (p->solver != OPS_SOLVER_THREE_MOMENT && p->solver != OPS_SOLVER_BAND_LDLT &&
        p->solver != OPS_SOLVER_THREE_MOMENT_THREAD && p->solver != OPS_SOLVER_THREE_MOMENT_SMEM8 &&
        p->solver != OPS_SOLVER_THREE_MOMENT_SMEM32)
        return OPS_E_BADARG;
    if (p->reserved != 0) return OPS_E_BADARG;
    k->nn = p->num_nodes;
    k->n = p->num_nodes - 1;
    k->max_forces = p->max_forces;
    k->max_epochs = p->max_epochs;
    k->patience = p->patience;
    k->early_stop = p->early_stop;
    k->zero_last_node = p->zero_last_node;
    k->E = p->E;
    k->udl = p->udl;
    k->tol = p->tolerance;
    k->I0f = (float)p->I0;
    k->E2 = (float)(2.0 * p->E);
    k->Gf = (float)p->G;
    k->kf = (float)p->shear_k;
    k->am = (float)p->alpha_moment;
    k->as_ = (float)p->alpha_shear;
    k->epsf = (float)p->bending_eps;
    k->clampf = (float)p->clamp_min;
    k->w1 = (float)(1.0 - p->beta1);
    k->b2f = (float)p->beta2;
    k->omb2f = (float)(1.0 - p->beta2);
    k->adam_epsf = (float)p->adam_eps;
    return 0;
}

struct LaunchPlan {
    bool lanes;          // eight-lanes-per-beam three-moment kernel (beamopt_lanes.cu)
    LanesPlan lp;
    bool wide;           // shared-memory-state three-moment kernels (beamopt_wide.cu)
    WidePlan wp;
    bool flex;           // thread-per-beam three-moment kernel
    FlexLayout lay;
    bool smem;
    int threads;         // per CTA
    int blocks;
    size_t smem_bytes;
    size_t ws_d_bytes, ws_f_bytes, ws_mask_bytes;   // global scratch (0 when smem)
};

static size_t per_beam_bytes(int nn)
{
    const int n = nn - 1;
    return (size_t)5 * nn * 8 + (size_t)3 * n * 4 + (size_t)((nn + 31) / 32) * 4;
}

static size_t flex_smem_per_thread(int n, const FlexLayout &lay)
{
    const int smem_arrays = (lay.I_global ? 0 : 2) + (lay.m_global ? 0 : 1) + (lay.v_global ? 0 : 1);
    return (size_t)FlexStore::NUM_DOUBLES * 8 + ((size_t)smem_arrays * n + 2 * (size_t)lay.nacc) * 4 +
           (size_t)FlexStore::NUM_INTS * 4;
}

static int plan_flex(const BeamConsts &k, int64_t B, int sms, int smem_optin, LaunchPlan *pl)
{
    memset(pl, 0, sizeof *pl);
    pl->flex = true;
    pl->lay.nacc = (k.n / 32) >= 16 ? 64 : 32;
    // Layout: more resident warps beat keeping everything in shared memory (profiles/): by default
    // Adam's m, v go to the L2-resident global scratch, I stays in shared memory; when even that
    // leaves fewer than 64 threads per SM (fine discretisations) I moves out as well.
    // OPS_FLEX_LAYOUT = smem | mv | all and OPS_FLEX_THREADS override (profiling knobs).
    const char *lay_env = getenv("OPS_FLEX_LAYOUT");
    const char *thr_env = getenv("OPS_FLEX_THREADS");
    int mode = 1;
    if (lay_env) mode = !strcmp(lay_env, "smem") ? 0 : (!strcmp(lay_env, "all") ? 2 : 1);
    for (;; ++mode) {
        pl->lay.I_global = mode >= 2;
        pl->lay.m_global = pl->lay.v_global = mode >= 1;
        const size_t per = flex_smem_per_thread(k.n, pl->lay);
        int T = (int)((size_t)smem_optin / per) / 32 * 32;
        if (T > 256) T = 256;
        if (thr_env && atoi(thr_env) >= 32 && atoi(thr_env) <= T) T = atoi(thr_env) / 32 * 32;
        if (T >= 64 || mode >= 2) {
            if (T < 32) return OPS_E_UNSUPP;
            pl->threads = T;
            pl->smem_bytes = per * T;
            break;
        }
    }
    const long want = (long)((B + pl->threads - 1) / pl->threads);
    pl->blocks = (int)(want < sms ? want : sms);
    if (pl->blocks < 1) pl->blocks = 1;
    const size_t total = (size_t)pl->blocks * pl->threads;
    const int garrays = (pl->lay.I_global ? 2 : 0) + (pl->lay.m_global ? 1 : 0) + (pl->lay.v_global ? 1 : 0);
    pl->ws_f_bytes = total * garrays * k.n * 4;
    return 0;
}

static int plan_launch_on(const BeamConsts &k, int num_cases, int64_t B, int solver, int sms, int smem_optin, LaunchPlan *pl);

static int plan_launch(const BeamConsts &k, int num_cases, int64_t B, int solver, LaunchPlan *pl)
{
    int dev = 0, sms = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return (int)e;
    return plan_launch_on(k, num_cases, B, solver, sms, smem_optin, pl);
}

// the choice of kernel family and launch geometry: host arithmetic on (parameters, batch size, SM count, shared memory)
static int plan_launch_on(const BeamConsts &k, int num_cases, int64_t B, int solver, int sms, int smem_optin, LaunchPlan *pl)
{
    if (solver == OPS_SOLVER_THREE_MOMENT && lanes_supported(k, num_cases)) {
        memset(pl, 0, sizeof *pl);
        pl->lanes = true;
        return lanes_plan(k, num_cases, B, sms, smem_optin, &pl->lp);
    }
    if (num_cases != 1) return OPS_E_UNSUPP;
    {
        // fine discretisations run one warp per beam; the explicit selectors pick a lane count
        const int lpb = solver == OPS_SOLVER_THREE_MOMENT_SMEM8 ? 8 :
                        ((solver == OPS_SOLVER_THREE_MOMENT_SMEM32 || solver == OPS_SOLVER_THREE_MOMENT) ? 32 : 0);
        if (lpb && wide_supported(k, num_cases, lpb, smem_optin)) {
            memset(pl, 0, sizeof *pl);
            pl->wide = true;
            return wide_plan(k, lpb, B, sms, smem_optin, &pl->wp);
        }
        if (solver == OPS_SOLVER_THREE_MOMENT_SMEM8 || solver == OPS_SOLVER_THREE_MOMENT_SMEM32) return OPS_E_UNSUPP;
    }
    if (solver == OPS_SOLVER_THREE_MOMENT || solver == OPS_SOLVER_THREE_MOMENT_THREAD)
        return plan_flex(k, B, sms, smem_optin, pl);
    const size_t pb = per_beam_bytes(k.nn);
    memset(pl, 0, sizeof *pl);
    // experiment knobs (profiling only): OPS_BEAMOPT_GLOBAL=1 forces the global-scratch variant,
    // OPS_BEAMOPT_CTAS_PER_SM sets its residency.
    const char *force_global = getenv("OPS_BEAMOPT_GLOBAL");
    const char *ctas_env = getenv("OPS_BEAMOPT_CTAS_PER_SM");
    const int ctas_per_sm = ctas_env ? atoi(ctas_env) : 4;
    if (32 * pb <= (size_t)smem_optin && !(force_global && force_global[0] == '1')) {
        pl->smem = true;
        pl->threads = 32;
        pl->smem_bytes = 32 * pb;
        int per_sm = (int)((size_t)smem_optin / pl->smem_bytes);
        if (per_sm < 1) per_sm = 1;
        long want = (long)((B + 31) / 32);
        long cap = (long)sms * per_sm;
        pl->blocks = (int)(want < cap ? want : cap);
    } else {
        pl->smem = false;
        pl->threads = 64;
        long want = (long)((B + 63) / 64);
        long cap = (long)sms * (ctas_per_sm > 0 ? ctas_per_sm : 4);
        pl->blocks = (int)(want < cap ? want : cap);
        const size_t total = (size_t)pl->blocks * pl->threads;
        pl->ws_d_bytes = total * 5 * k.nn * 8;
        pl->ws_f_bytes = total * 3 * k.n * 4;
        pl->ws_mask_bytes = total * ((k.nn + 31) / 32) * 4;
    }
    if (pl->blocks < 1) pl->blocks = 1;
    return 0;
}

// DFMA throughput probe: 8 independent chains per thread keep the FP64 pipe full.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;     // never true; keeps the chains alive
}

// Issue-rate probe of one instruction class: 8 independent chains per thread, 1024 threads per SM.
//   0 DFMA   1 FFMA (3 register operands)   2 FMUL   3 FADD   4 MUFU.RCP   5 F2F.F64.F32 + F2F.F32.F64
//   6 IMAD   7 LOP3   8 FFMA with an immediate operand
template <int OP>
__global__ void __launch_bounds__(256) pipe_probe_kernel(int iters, float seed, float *sink)
{
    float a[8];
    double d[8];
    int n[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = seed + threadIdx.x + j; d[j] = a[j]; n[j] = (int)a[j]; }
    const float m = 0.9999f + seed * 1e-9f, c = 1e-7f + seed * 1e-12f;
    const double md = m, cd = c;
    const int mi = 3 + (int)seed;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) d[j] = fma(d[j], md, cd);
            if (OP == 1) a[j] = fmaf(a[j], m, c);
            if (OP == 2) a[j] = a[j] * m;
            if (OP == 3) a[j] = a[j] + c;
            if (OP == 4) { float t_; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t_) : "f"(a[j])); a[j] = t_; }
            if (OP == 5) { d[j] = (double)a[j]; asm volatile("" : "+d"(d[j])); a[j] = (float)d[j]; asm volatile("" : "+f"(a[j])); }
            if (OP == 6) n[j] = n[j] * mi + i;
            if (OP == 7) n[j] = (n[j] ^ mi) & (i | 0x55);
            if (OP == 8) a[j] = fmaf(a[j], 0.9999f, 1e-7f);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + (float)d[j] + (float)n[j];
    if (s == 12345.678f) sink[0] = s;
}

}  // namespace ops

using namespace ops;

extern "C" {

const char *ops_beamopt_version(void) { return OPS_VERSION; }

int ops_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ops_set_device(int device)
{
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

int ops_beamopt_fill_schedule(const OpsBeamOptParams *p, float *host_table)
{
    if (!p || !host_table || p->struct_size != (int32_t)sizeof(OpsBeamOptParams)) return OPS_E_BADARG;
    // torch/optim/adam.py (_single_tensor_adam) + lr_scheduler.ExponentialLR, all in double:
    //   bias_correction1 = 1 - beta1**step ; bias_correction2 = 1 - beta2**step
    //   step_size = lr / bias_correction1 ; bias_correction2_sqrt = bias_correction2**0.5
    //   lr <- lr * gamma after every optimizer step
    double lr = p->lr;
    for (int t = 1; t <= p->max_epochs; ++t) {
        const double bc1 = 1.0 - pow(p->beta1, (double)t);
        const double bc2 = 1.0 - pow(p->beta2, (double)t);
        host_table[2 * (t - 1)] = (float)(-(lr / bc1));
        host_table[2 * (t - 1) + 1] = (float)pow(bc2, 0.5);
        lr = lr * p->gamma;
    }
    return 0;
}

int ops_beamopt_plan(const OpsBeamOptParams *p, int64_t B, int32_t sms, int32_t smem_optin, OpsLaunchPlanInfo *out)
{
    BeamConsts k;
    if (!p || !out || B < 0) return OPS_E_BADARG;
    int rc = make_consts(p, &k);
    if (rc != 0) return rc;
    LaunchPlan pl;
    rc = (sms > 0 && smem_optin > 0) ? plan_launch_on(k, p->num_cases, B, p->solver, sms, smem_optin, &pl)
                                     : plan_launch(k, p->num_cases, B, p->solver, &pl);
    if (rc != 0) return rc;
    memset(out, 0, sizeof *out);
    if (pl.lanes) {
        out->family = pl.lp.tm ? OPS_PLAN_LANES_TM : OPS_PLAN_LANES;
        out->threads = pl.lp.threads; out->blocks = pl.lp.blocks; out->smem_bytes = (int64_t)pl.lp.smem_bytes;
        out->lanes_per_beam = 8 * pl.lp.num_cases;
        out->beams_per_cta = pl.lp.threads / out->lanes_per_beam;
        out->scatter = lanes_scatter_supported(pl.lp) ? 1 : 0;
    } else if (pl.wide) {
        out->family = OPS_PLAN_WIDE;
        out->threads = pl.wp.threads; out->blocks = pl.wp.blocks; out->smem_bytes = (int64_t)pl.wp.smem_bytes;
        out->lanes_per_beam = pl.wp.lpb;
        out->beams_per_cta = pl.wp.threads / pl.wp.lpb;
    } else {
        out->family = pl.flex ? OPS_PLAN_THREAD_THREE_MOMENT : OPS_PLAN_THREAD_LDLT;
        out->threads = pl.threads; out->blocks = pl.blocks; out->smem_bytes = (int64_t)pl.smem_bytes;
        out->lanes_per_beam = 1;
        out->beams_per_cta = pl.threads;
    }
    out->workspace_bytes = (int64_t)(256 + pl.ws_d_bytes + pl.ws_f_bytes + pl.ws_mask_bytes + 512);
    return 0;
}

size_t ops_beamopt_workspace_bytes(const OpsBeamOptParams *p, int64_t B)
{
    BeamConsts k;
    if (make_consts(p, &k) != 0 || B < 0) return 0;
    LaunchPlan pl;
    if (plan_launch(k, p->num_cases, B, p->solver, &pl) != 0) return 0;
    return 256 + pl.ws_d_bytes + pl.ws_f_bytes + pl.ws_mask_bytes + 512;
}

// One launch; n_dest == 0: the record arrays given here (rows 0..B-1); n_dest >= 1: `dests` holds n_dest sets of the
// eight record arrays (order of OpsBeamOptRecordArrays) and beam b is written to row row0 + b of every set.
static int launch_impl(const OpsBeamOptParams *p, int64_t B,
                       const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                       const double *L, const float *d_schedule,
                       float *I_values, double *deflections, double *rotations, float *shear,
                       float *moment, int32_t *epochs, float *loss, int32_t *status,
                       int n_dest, const OpsBeamOptRecordArrays *dests, int64_t row0,
                       void *d_workspace, size_t workspace_bytes, void *cuda_stream)
{
    BeamConsts k;
    int rc = make_consts(p, &k);
    if (rc) return rc;
    if (B < 0) return OPS_E_BADARG;
    if (B == 0) return 0;
    if (n_dest == 0 && (!I_values || !deflections || !rotations || !shear || !moment || !epochs || !loss || !status))
        return OPS_E_BADARG;
    if (!fixed_uy || !L || !d_workspace || (p->max_forces > 0 && (!force_nodes || !force_vals)) ||
        (p->max_epochs > 0 && !d_schedule))
        return OPS_E_BADARG;
    if (n_dest < 0 || n_dest > lanes::MAX_DEST || (n_dest > 0 && !dests) || row0 < 0) return OPS_E_BADARG;
    LaunchPlan pl;
    rc = plan_launch(k, p->num_cases, B, p->solver, &pl);
    if (rc) return rc;
    if (n_dest > 1 && !(pl.lanes && lanes_scatter_supported(pl.lp))) return OPS_E_UNSUPP;   // the lanes kernel's scatter instances
    const size_t need = 256 + pl.ws_d_bytes + pl.ws_f_bytes + pl.ws_mask_bytes + 512;
    if (workspace_bytes < need) return OPS_E_WORKSPACE;
    cudaStream_t stream = (cudaStream_t)cuda_stream;

    unsigned char *ws = (unsigned char *)d_workspace;
    OptPtrs q;
    q.fixed_uy = fixed_uy; q.force_nodes = force_nodes; q.force_vals = force_vals; q.L = L;
    q.sched = d_schedule; q.I_values = I_values; q.defl = deflections; q.rot = rotations;
    q.shear = shear; q.moment = moment; q.epochs = epochs; q.loss = loss; q.status = status;
    memset(&q.dest, 0, sizeof(q.dest));
    q.row0 = n_dest > 0 ? (long long)row0 : 0;
    q.dest.nd = n_dest > 0 ? n_dest : 1;
    for (int r = 0; r < q.dest.nd; ++r) {
        OpsBeamOptRecordArrays a;
        if (n_dest > 0) a = dests[r];
        else { a.I_values = I_values; a.deflections = deflections; a.rotations = rotations; a.shear = shear;
               a.moment = moment; a.epochs = epochs; a.loss = loss; a.status = status; }
        if (!a.I_values || !a.deflections || !a.rotations || !a.shear || !a.moment || !a.epochs || !a.loss || !a.status)
            return OPS_E_BADARG;
        q.dest.I[r] = a.I_values; q.dest.defl[r] = a.deflections; q.dest.rot[r] = a.rotations;
        q.dest.shear[r] = a.shear; q.dest.moment[r] = a.moment; q.dest.epochs[r] = a.epochs;
        q.dest.loss[r] = a.loss; q.dest.status[r] = a.status;
    }
    // destination 0 is where the kernel writes the record; it copies the rows to the others
    q.I_values = q.dest.I[0]; q.defl = q.dest.defl[0]; q.rot = q.dest.rot[0]; q.shear = q.dest.shear[0];
    q.moment = q.dest.moment[0]; q.epochs = q.dest.epochs[0]; q.loss = q.dest.loss[0]; q.status = q.dest.status[0];
    q.counter = (unsigned long long *)ws;
    size_t off = 256;
    q.ws_d = (double *)(ws + off); off += (pl.ws_d_bytes + 255) / 256 * 256;
    q.ws_f = (float *)(ws + off);  off += (pl.ws_f_bytes + 255) / 256 * 256;
    q.ws_mask = (uint32_t *)(ws + off);
    cudaError_t e = cudaMemsetAsync(q.counter, 0, 256, stream);
    if (e != cudaSuccess) return (int)e;

    if (pl.lanes) {
        e = lanes_launch(k, (long long)B, q, pl.lp, stream);
        return e == cudaSuccess ? 0 : (int)e;
    }
    if (pl.wide) {
        e = wide_launch(k, (long long)B, q, pl.wp, stream);
        return e == cudaSuccess ? 0 : (int)e;
    }
    if (pl.flex) {
        e = cudaFuncSetAttribute(beamopt_flex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pl.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        beamopt_flex_kernel<<<pl.blocks, pl.threads, pl.smem_bytes, stream>>>(k, (long long)B, q, pl.lay);
        e = cudaGetLastError();
        return e == cudaSuccess ? 0 : (int)e;
    }
    const int mf = p->max_forces <= 4 ? 4 : 8;
#define OPS_LAUNCH(MAXF, SMEM)                                                                      \
    do {                                                                                            \
        auto kern = beamopt_kernel<MAXF, SMEM>;                                                     \
        if (SMEM) {                                                                                 \
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                     (int)pl.smem_bytes);                                           \
            if (e != cudaSuccess) return (int)e;                                                    \
        }                                                                                           \
        kern<<<pl.blocks, pl.threads, SMEM ? pl.smem_bytes : 0, stream>>>(k, (long long)B, q);      \
    } while (0)
    if (pl.smem) { if (mf == 4) OPS_LAUNCH(4, true); else OPS_LAUNCH(8, true); }
    else         { if (mf == 4) OPS_LAUNCH(4, false); else OPS_LAUNCH(8, false); }
#undef OPS_LAUNCH
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

int ops_beamopt_launch(const OpsBeamOptParams *p, int64_t B,
                       const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                       const double *L, const float *d_schedule,
                       float *I_values, double *deflections, double *rotations, float *shear,
                       float *moment, int32_t *epochs, float *loss, int32_t *status,
                       void *d_workspace, size_t workspace_bytes, void *cuda_stream)
{
    return launch_impl(p, B, fixed_uy, force_nodes, force_vals, L, d_schedule, I_values, deflections, rotations, shear,
                       moment, epochs, loss, status, 0, nullptr, 0, d_workspace, workspace_bytes, cuda_stream);
}

// ---------------------------------------------------------------------------------------------
// Dataset gather with a FIXED epoch count: all beams of a round finish in the same epoch, so the in-kernel copy of a
// round's records to the peers (lane_copy_record) is a burst at the end of the round that no iteration is left to hide
// (8 GPUs: 0.85 ms of a 5.6 ms step, profiles/r02_scatter_ab_n8.txt).  Such batches run as a PIPELINE instead: the
// plain kernel instance, launched in chunks of whole rounds into this GPU's arrays, and after every chunk a copy kernel
// on a side stream that stores the chunk's rows -- contiguous blocks of the eight arrays -- into the peers' arrays over
// NVLink while the next chunk iterates (it fits beside the compute CTA: no shared memory, 128 threads).  Only the last
// chunk's copy is exposed.  Early-stopped batches keep the in-kernel copy: there the beams end at different epochs and
// the copies hide behind the other groups' iterations.
// ---------------------------------------------------------------------------------------------
#define OPS_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
struct PeerCopyJob {
    int nd, arrays;
    char *base[lanes::MAX_DEST][8];             // [destination][array]
    long long first[8], bytes[8];               // byte range of the chunk's rows inside each array
};

struct alignas(16) CopyUnit16 { unsigned long long a, b; };

// every 16-byte unit is read once and stored to all peers; ragged 4-byte heads / tails of a range (rows are multiples
// of 4 bytes, the array bases 256-byte aligned) go one word at a time
__global__ void __launch_bounds__(128) peer_copy_kernel(const PeerCopyJob job)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
    for (int a = 0; a < job.arrays; ++a) {
        const long long lo = job.first[a], hi = lo + job.bytes[a];
        long long body_lo = (lo + 15) & ~15LL, body_hi = hi & ~15LL;
        if (body_lo > body_hi) { body_lo = hi; body_hi = hi; }
        const char *src = job.base[0][a];
        for (long long u = tid; u < ((body_hi - body_lo) >> 4); u += nthreads) {
            const CopyUnit16 v = *reinterpret_cast<const CopyUnit16 *>(src + body_lo + (u << 4));
            for (int r = 1; r < job.nd; ++r) *reinterpret_cast<CopyUnit16 *>(job.base[r][a] + body_lo + (u << 4)) = v;
        }
        const long long head = (body_lo - lo) >> 2, tail = (hi - body_hi) >> 2;
        for (long long u = tid; u < head + tail; u += nthreads) {
            const long long off = u < head ? lo + (u << 2) : body_hi + ((u - head) << 2);
            const unsigned int v = *reinterpret_cast<const unsigned int *>(src + off);
            for (int r = 1; r < job.nd; ++r) *reinterpret_cast<unsigned int *>(job.base[r][a] + off) = v;
        }
    }
}

// side stream + fork / join events of the pipeline, one set per device, created on first use
struct ScatterPipe {
    cudaStream_t side;
    cudaEvent_t fork, join;
    bool ready;
};
static ScatterPipe g_pipe[64];
static std::mutex g_pipe_mutex;

static int scatter_pipe(ScatterPipe **out)
{
    int dev = 0;
    OPS_CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return OPS_E_BADARG;
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    ScatterPipe &sp = g_pipe[dev];
    if (!sp.ready) {
        OPS_CU(cudaStreamCreateWithFlags(&sp.side, cudaStreamNonBlocking));
        OPS_CU(cudaEventCreateWithFlags(&sp.fork, cudaEventDisableTiming));
        OPS_CU(cudaEventCreateWithFlags(&sp.join, cudaEventDisableTiming));
        sp.ready = true;
    }
    *out = &sp;
    return 0;
}

static int launch_scatter_pipelined(const OpsBeamOptParams *p, int64_t B, int64_t chunk,
                                    const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                                    const double *L, const float *d_schedule, int n_dest,
                                    const OpsBeamOptRecordArrays *dests, int64_t row0, void *d_workspace,
                                    size_t workspace_bytes, cudaStream_t stream)
{
    ScatterPipe *sp = nullptr;
    int rc = scatter_pipe(&sp);
    if (rc) return rc;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const size_t nn = (size_t)p->num_nodes, n = nn - 1, C = (size_t)p->num_cases, F = (size_t)p->max_forces * C;
    const size_t row_bytes[8] = {n * 4, C * nn * 8, C * nn * 8, C * n * 4, C * n * 4, 4, 4, 4};
    for (int64_t r0 = 0; r0 < B; r0 += chunk) {
        const int64_t m = (B - r0) < chunk ? (B - r0) : chunk;
        rc = launch_impl(p, m, fixed_uy + (size_t)r0 * nn, force_nodes ? force_nodes + (size_t)r0 * F : nullptr,
                         force_vals ? force_vals + (size_t)r0 * F : nullptr, L + r0, d_schedule, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, nullptr, nullptr, 1, dests, row0 + r0, d_workspace, workspace_bytes,
                         (void *)stream);
        if (rc) return rc;
        PeerCopyJob job;
        memset(&job, 0, sizeof job);
        job.nd = n_dest; job.arrays = 8;
        for (int r = 0; r < n_dest; ++r) {
            const OpsBeamOptRecordArrays &a = dests[r];
            char *b8[8] = {(char *)a.I_values, (char *)a.deflections, (char *)a.rotations, (char *)a.shear, (char *)a.moment,
                           (char *)a.epochs, (char *)a.loss, (char *)a.status};
            for (int i = 0; i < 8; ++i) { if (!b8[i]) return OPS_E_BADARG; job.base[r][i] = b8[i]; }
        }
        for (int i = 0; i < 8; ++i) {
            job.first[i] = (long long)((size_t)(row0 + r0) * row_bytes[i]);
            job.bytes[i] = (long long)((size_t)m * row_bytes[i]);
        }
        OPS_CU(cudaEventRecord(sp->fork, stream));
        OPS_CU(cudaStreamWaitEvent(sp->side, sp->fork, 0));
        peer_copy_kernel<<<sms * 4, 128, 0, sp->side>>>(job);
        OPS_CU(cudaGetLastError());
    }
    OPS_CU(cudaEventRecord(sp->join, sp->side));
    OPS_CU(cudaStreamWaitEvent(stream, sp->join, 0));
    return 0;
}

int ops_beamopt_launch_scatter(const OpsBeamOptParams *p, int64_t B,
                               const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                               const double *L, const float *d_schedule,
                               int n_dest, const OpsBeamOptRecordArrays *dests, int64_t row0,
                               void *d_workspace, size_t workspace_bytes, void *cuda_stream)
{
    if (n_dest < 1) return OPS_E_BADARG;
    // Fixed epoch count, more than one peer, a batch of a few rounds: chunks of whole rounds + the copy kernel.  With a
    // single peer the copy inside the kernel is as good as free (2 GPUs: 4.83 ms against 4.81 ms on one), with seven it
    // is 0.8 ms of a 1.7-round step.  Over many rounds the in-kernel copies of all rounds but the last hide behind the
    // following rounds, and the pipeline's chunk granularity costs more than that last burst (1 M beams on 8 GPUs:
    // 53.0 ms in-kernel, 55.9 ms pipelined), so batches beyond eight rounds keep the in-kernel copy.
    // OPS_SCATTER_IN_KERNEL / OPS_SCATTER_PIPELINED force either form (A/B runs; same results).
    bool pipelined = n_dest > 2;
    if (pipelined) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t per_round = (int64_t)sms * 40 / (p && p->num_cases > 0 ? p->num_cases : 1);
        pipelined = B < 8 * per_round;
    }
    if (getenv("OPS_SCATTER_PIPELINED")) pipelined = true;
    else if (getenv("OPS_SCATTER_IN_KERNEL")) pipelined = false;
    if (n_dest > 1 && n_dest <= lanes::MAX_DEST && dests && B > 0 && p && !p->early_stop && pipelined) {
        BeamConsts k;
        int rc = make_consts(p, &k);
        if (rc) return rc;
        LaunchPlan pl;
        rc = plan_launch(k, p->num_cases, B, p->solver, &pl);
        if (rc) return rc;
        if (!(pl.lanes && lanes_scatter_supported(pl.lp))) return OPS_E_UNSUPP;
        int64_t chunk = (int64_t)pl.lp.blocks * (pl.lp.threads / (lanes::LPB * p->num_cases));
        if (!pl.lp.tm && pl.lp.threads > 320) chunk *= 3;            // the 384-thread instance is chosen for >= 3 rounds
        if (chunk <= 0 || B < chunk + chunk / 4) chunk = B;
        return launch_scatter_pipelined(p, B, chunk, fixed_uy, force_nodes, force_vals, L, d_schedule, n_dest, dests, row0,
                                        d_workspace, workspace_bytes, (cudaStream_t)cuda_stream);
    }
    return launch_impl(p, B, fixed_uy, force_nodes, force_vals, L, d_schedule, nullptr, nullptr, nullptr, nullptr,
                       nullptr, nullptr, nullptr, nullptr, n_dest, dests, row0, d_workspace, workspace_bytes, cuda_stream);
}

// Whether ops_beamopt_launch_scatter serves this configuration (a function of the parameter block alone, so every rank
// of a job gets the same answer without talking to the others -- ranks with an empty shard included).
int ops_beamopt_scatter_supported(const OpsBeamOptParams *p)
{
    BeamConsts k;
    if (make_consts(p, &k)) return 0;
    LaunchPlan pl;
    if (plan_launch(k, p->num_cases, 1, p->solver, &pl)) return 0;
    return (pl.lanes && lanes_scatter_supported(pl.lp)) ? 1 : 0;
}

// --- peer-visible device buffers (CUDA IPC): the dataset arrays the in-kernel scatter writes over NVLink ---
int ops_peer_alloc(size_t bytes, void **dptr, unsigned char *handle64)
{
    if (!dptr || !handle64 || bytes == 0) return OPS_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, d);
    if (e != cudaSuccess) { cudaFree(d); return (int)e; }
    memcpy(handle64, &h, 64);
    *dptr = d;
    return 0;
}

int ops_peer_open(const unsigned char *handle64, void **dptr)
{
    if (!dptr || !handle64) return OPS_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *d = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *dptr = d;
    return 0;
}

int ops_peer_close(void *dptr)
{
    cudaError_t e = cudaIpcCloseMemHandle(dptr);
    return e == cudaSuccess ? 0 : (int)e;
}

int ops_peer_free(void *dptr)
{
    cudaError_t e = cudaFree(dptr);
    return e == cudaSuccess ? 0 : (int)e;
}

int ops_beamsolve_launch(const OpsBeamOptParams *p, int64_t B,
                         const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                         const double *L, const double *I_f64,
                         double *deflections, double *rotations, double *shear, double *moment,
                         int32_t *status, void *cuda_stream)
{
    BeamConsts k;
    int rc = make_consts(p, &k);
    if (rc) return rc;
    if (B < 0) return OPS_E_BADARG;
    if (B == 0) return 0;
    if (!fixed_uy || !L || !I_f64 || !deflections || !rotations || !shear || !moment || !status ||
        (p->max_forces > 0 && (!force_nodes || !force_vals)))
        return OPS_E_BADARG;
    int dev = 0, smem_optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return (int)e;
    SolvePtrs q;
    q.fixed_uy = fixed_uy; q.force_nodes = force_nodes; q.force_vals = force_vals; q.L = L; q.I = I_f64;
    q.defl = deflections; q.rot = rotations; q.shear = shear; q.moment = moment; q.status = status;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    if (p->solver != OPS_SOLVER_BAND_LDLT) {
        const int threads = 64;
        const size_t smem = (size_t)threads * (FlexStore::NUM_DOUBLES * 8 + FlexStore::NUM_INTS * 4);
        e = cudaFuncSetAttribute(beamsolve_flex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        beamsolve_flex_kernel<<<(int)((B + threads - 1) / threads), threads, smem, stream>>>(k, (long long)B, q);
        e = cudaGetLastError();
        return e == cudaSuccess ? 0 : (int)e;
    }
    const size_t pb = (size_t)5 * k.nn * 8;
    int threads = 32;
    while (threads > 1 && threads * pb > (size_t)smem_optin) threads >>= 1;
    if (threads * pb > (size_t)smem_optin) return OPS_E_UNSUPP;
    const size_t smem = threads * pb;
    const int blocks = (int)((B + threads - 1) / threads);
    const int mf = p->max_forces <= 4 ? 4 : 8;
    if (mf == 4) {
        auto kern = beamsolve_kernel<4>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<blocks, threads, smem, stream>>>(k, (long long)B, q);
    } else {
        auto kern = beamsolve_kernel<8>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<blocks, threads, smem, stream>>>(k, (long long)B, q);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

int ops_fp64_peak_probe(int iters, double *tflops, float *elapsed_ms, void *cuda_stream)
{
    if (iters <= 0 || !tflops) return OPS_E_BADARG;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    double *sink = nullptr;
    cudaEvent_t ev0, ev1;
    e = cudaMalloc((void **)&sink, 8);
    if (e != cudaSuccess) return (int)e;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    const int blocks = sms * 8, threads = 256;
    fp64_probe_kernel<<<blocks, threads, 0, stream>>>(iters / 8 + 1, 1.0, sink);   // warm-up
    cudaEventRecord(ev0, stream);
    fp64_probe_kernel<<<blocks, threads, 0, stream>>>(iters, 1.0, sink);
    cudaEventRecord(ev1, stream);
    e = cudaStreamSynchronize(stream);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaFree(sink);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / ((double)ms * 1e-3) / 1e12;
    if (elapsed_ms) *elapsed_ms = ms;
    return 0;
}

int ops_fastmath_selftest(int64_t samples, int64_t *mismatches3, int64_t *samples_run, double *rcp64_max_rel_err,
                          void *cuda_stream)
{
    if (samples <= 0 || !mismatches3 || !samples_run || !rcp64_max_rel_err) return OPS_E_BADARG;
    unsigned long long out[4] = {0, 0, 0, 0};
    cudaError_t e = fastmath_selftest((long long)samples, out, rcp64_max_rel_err, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    mismatches3[0] = (int64_t)out[0]; mismatches3[1] = (int64_t)out[1]; mismatches3[2] = (int64_t)out[2];
    *samples_run = (int64_t)out[3];
    return 0;
}

int ops_pipe_probe(int op, int iters, double *warp_inst_per_clk_per_sm, void *cuda_stream)
{
    if (iters <= 0 || !warp_inst_per_clk_per_sm || op < 0 || op > 8) return OPS_E_BADARG;
    int dev = 0, sms = 0, khz = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    float *sink = nullptr;
    e = cudaMalloc((void **)&sink, 4);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    const int blocks = sms * 4, threads = 256;
    void (*kerns[9])(int, float, float *) = {pipe_probe_kernel<0>, pipe_probe_kernel<1>, pipe_probe_kernel<2>,
                                             pipe_probe_kernel<3>, pipe_probe_kernel<4>, pipe_probe_kernel<5>,
                                             pipe_probe_kernel<6>, pipe_probe_kernel<7>, pipe_probe_kernel<8>};
    kerns[op]<<<blocks, threads, 0, stream>>>(iters / 8 + 1, 1.0f, sink);
    cudaEventRecord(ev0, stream);
    kerns[op]<<<blocks, threads, 0, stream>>>(iters, 1.0f, sink);
    cudaEventRecord(ev1, stream);
    e = cudaStreamSynchronize(stream);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaFree(sink);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const double per_op = (op == 5) ? 2.0 : 1.0;
    const double warp_inst = per_op * 8.0 * (double)iters * blocks * threads / 32.0;
    *warp_inst_per_clk_per_sm = warp_inst / ((double)ms * 1e-3 * (double)khz * 1e3) / sms;
    return 0;
}

#define OPS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)

int ops_beamopt_run_host(const OpsBeamOptParams *p, int64_t B,
                         const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                         const double *L,
                         float *I_values, double *deflections, double *rotations, float *shear,
                         float *moment, int32_t *epochs, float *loss, int32_t *status,
                         int device, float *elapsed_ms);

struct OpsBeamOptSession {
    OpsBeamOptParams p;
    int64_t max_beams;
    int device;
    cudaStream_t stream, copy;           // launches / device->host copies of finished chunks
    cudaEvent_t ev0, ev1, evc;
    unsigned char *dbuf, *hbuf;          // one device and one pinned host allocation, same offsets
    size_t in_bytes, out_bytes, ws_bytes;
    size_t o_fixed, o_fn, o_fv, o_L, o_I, o_defl, o_rot, o_sh, o_mo, o_ep, o_loss, o_st, o_sched, o_ws;
};

int ops_beamopt_session_create(const OpsBeamOptParams *p, int64_t max_beams, int device, OpsBeamOptSession **out)
{
    BeamConsts k;
    int rc = make_consts(p, &k);
    if (rc) return rc;
    if (!out || max_beams <= 0) return OPS_E_BADARG;
    OpsBeamOptSession *s = (OpsBeamOptSession *)calloc(1, sizeof *s);
    if (!s) return OPS_E_BADARG;
    s->p = *p; s->max_beams = max_beams; s->device = device;
    const size_t B = (size_t)max_beams, nn = k.nn, n = k.n, C = (size_t)p->num_cases, F = (size_t)p->max_forces;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    // inputs first (one contiguous H2D range), then outputs (one contiguous D2H range), then device-only
    s->o_fixed = take(B * nn); s->o_fn = take(B * C * F * 4); s->o_fv = take(B * C * F * 8); s->o_L = take(B * 8);
    s->in_bytes = off;
    s->o_I = take(B * n * 4); s->o_defl = take(B * C * nn * 8); s->o_rot = take(B * C * nn * 8);
    s->o_sh = take(B * C * n * 4); s->o_mo = take(B * C * n * 4);
    s->o_ep = take(B * 4); s->o_loss = take(B * 4); s->o_st = take(B * 4);
    s->out_bytes = off - s->in_bytes;
    const size_t host_bytes = off;
    s->o_sched = take((size_t)(p->max_epochs > 0 ? p->max_epochs : 1) * 8);
    float *sched_h = nullptr;
    OPS_CUDA(cudaSetDevice(device));
    s->ws_bytes = ops_beamopt_workspace_bytes(p, max_beams);
    if (s->ws_bytes == 0) { rc = OPS_E_BADARG; goto done; }
    s->o_ws = take(s->ws_bytes);
    OPS_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    OPS_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    OPS_CUDA(cudaEventCreate(&s->ev0));
    OPS_CUDA(cudaEventCreate(&s->ev1));
    OPS_CUDA(cudaEventCreateWithFlags(&s->evc, cudaEventDisableTiming));
    OPS_CUDA(cudaMalloc((void **)&s->dbuf, off));
    OPS_CUDA(cudaHostAlloc((void **)&s->hbuf, host_bytes, cudaHostAllocDefault));
    memset(s->hbuf, 0, host_bytes);
    sched_h = (float *)malloc((size_t)(p->max_epochs > 0 ? p->max_epochs : 1) * 8);
    if (!sched_h) { rc = OPS_E_BADARG; goto done; }
    ops_beamopt_fill_schedule(p, sched_h);
    OPS_CUDA(cudaMemcpyAsync(s->dbuf + s->o_sched, sched_h, (size_t)(p->max_epochs > 0 ? p->max_epochs : 1) * 8,
                             cudaMemcpyHostToDevice, s->stream));
    OPS_CUDA(cudaStreamSynchronize(s->stream));
done:
    free(sched_h);
    if (rc) { ops_beamopt_session_destroy(s); if (rc > 0) cudaGetLastError(); return rc; }
    *out = s;
    return 0;
}

int ops_beamopt_session_arrays(OpsBeamOptSession *s, OpsBeamOptHostArrays *a)
{
    if (!s || !a) return OPS_E_BADARG;
    unsigned char *h = s->hbuf;
    a->fixed_uy = h + s->o_fixed; a->force_nodes = (int32_t *)(h + s->o_fn); a->force_vals = (double *)(h + s->o_fv);
    a->L = (double *)(h + s->o_L); a->I_values = (float *)(h + s->o_I); a->deflections = (double *)(h + s->o_defl);
    a->rotations = (double *)(h + s->o_rot); a->shear = (float *)(h + s->o_sh); a->moment = (float *)(h + s->o_mo);
    a->epochs = (int32_t *)(h + s->o_ep); a->loss = (float *)(h + s->o_loss); a->status = (int32_t *)(h + s->o_st);
    return 0;
}

// Beams per launch of a session run.  A run is a pipeline: the whole batch's inputs go up, then the batch is optimised
// in chunks of whole rounds (a round = the beams resident on the GPU at once) and every chunk's record comes down on a
// second stream while the next chunk iterates, so only the last chunk's device->host copy is exposed.  With fixed
// epochs a chunk is one round (three of the 384-thread instance); with early stopping it is at least six, because a
// launch ends with its slowest beam and refills its groups only from its own chunk.  B itself = no pipelining.
static int64_t session_chunk_beams(const OpsBeamOptSession *s, int64_t B)
{
    BeamConsts k;
    if (make_consts(&s->p, &k) != 0) return B;
    LaunchPlan pl;
    if (plan_launch(k, s->p.num_cases, B, s->p.solver, &pl) != 0) return B;
    int64_t per_round = 0;
    int rounds = 1;
    if (pl.lanes) {
        per_round = (int64_t)pl.lp.blocks * (pl.lp.threads / (lanes::LPB * s->p.num_cases));
        if (pl.lp.threads > 320) rounds = 3;
    } else if (pl.wide) {
        per_round = (int64_t)pl.wp.blocks * (pl.wp.threads / pl.wp.lpb);
    } else {
        return B;
    }
    if (s->p.early_stop && rounds < 6) rounds = 6;
    const int64_t chunk = per_round * rounds;
    if (chunk <= 0 || B < chunk + chunk / 4) return B;
    return chunk;
}

int ops_beamopt_session_run(OpsBeamOptSession *s, int64_t B, float *elapsed_ms)
{
    if (!s || B < 0 || B > s->max_beams) return OPS_E_BADARG;
    if (B == 0) return 0;
    int rc = 0;
    const size_t nn = (size_t)s->p.num_nodes, n = nn - 1, C = (size_t)s->p.num_cases;
    const size_t F = (size_t)s->p.max_forces * C, b = (size_t)B;
    unsigned char *d = s->dbuf, *h = s->hbuf;
    auto h2d = [&](size_t o, size_t bytes) { return cudaMemcpyAsync(d + o, h + o, bytes, cudaMemcpyHostToDevice, s->stream); };
    auto d2h = [&](size_t o, size_t first, size_t rows, size_t row_bytes) {       // rows [first, first + rows) of an output
        return cudaMemcpyAsync(h + o + first * row_bytes, d + o + first * row_bytes, rows * row_bytes,
                               cudaMemcpyDeviceToHost, s->copy);
    };
    const int64_t chunk = session_chunk_beams(s, B);
    OPS_CUDA(cudaSetDevice(s->device));
    if (B == s->max_beams) {
        OPS_CUDA(h2d(0, s->in_bytes));
    } else {
        OPS_CUDA(h2d(s->o_fixed, b * nn));
        if (F > 0) { OPS_CUDA(h2d(s->o_fn, b * F * 4)); OPS_CUDA(h2d(s->o_fv, b * F * 8)); }
        OPS_CUDA(h2d(s->o_L, b * 8));
    }
    OPS_CUDA(cudaEventRecord(s->ev0, s->stream));
    for (int64_t r0 = 0; r0 < B; r0 += chunk) {
        const size_t f = (size_t)r0, m = (size_t)((B - r0) < chunk ? (B - r0) : chunk);
        rc = ops_beamopt_launch(&s->p, (int64_t)m, d + s->o_fixed + f * nn, (const int32_t *)(d + s->o_fn) + f * F,
                                (const double *)(d + s->o_fv) + f * F, (const double *)(d + s->o_L) + f,
                                (const float *)(d + s->o_sched), (float *)(d + s->o_I) + f * n,
                                (double *)(d + s->o_defl) + f * C * nn, (double *)(d + s->o_rot) + f * C * nn,
                                (float *)(d + s->o_sh) + f * C * n, (float *)(d + s->o_mo) + f * C * n,
                                (int32_t *)(d + s->o_ep) + f, (float *)(d + s->o_loss) + f, (int32_t *)(d + s->o_st) + f,
                                d + s->o_ws, s->ws_bytes, s->stream);
        if (rc) goto done;
        // this chunk's record goes down on the copy stream as soon as its launch has finished
        OPS_CUDA(cudaEventRecord(s->evc, s->stream));
        OPS_CUDA(cudaStreamWaitEvent(s->copy, s->evc, 0));
        OPS_CUDA(d2h(s->o_I, f, m, n * 4)); OPS_CUDA(d2h(s->o_defl, f, m, C * nn * 8)); OPS_CUDA(d2h(s->o_rot, f, m, C * nn * 8));
        OPS_CUDA(d2h(s->o_sh, f, m, C * n * 4)); OPS_CUDA(d2h(s->o_mo, f, m, C * n * 4));
        OPS_CUDA(d2h(s->o_ep, f, m, 4)); OPS_CUDA(d2h(s->o_loss, f, m, 4)); OPS_CUDA(d2h(s->o_st, f, m, 4));
    }
    OPS_CUDA(cudaEventRecord(s->ev1, s->stream));
    OPS_CUDA(cudaStreamSynchronize(s->stream));
    OPS_CUDA(cudaStreamSynchronize(s->copy));
    if (elapsed_ms) OPS_CUDA(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
done:
    if (rc > 0) cudaGetLastError();
    if (rc) { cudaStreamSynchronize(s->stream); cudaStreamSynchronize(s->copy); }
    return rc;
}

void ops_beamopt_session_destroy(OpsBeamOptSession *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->dbuf) cudaFree(s->dbuf);
    if (s->hbuf) cudaFreeHost(s->hbuf);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evc) cudaEventDestroy(s->evc);
    if (s->copy) { cudaStreamSynchronize(s->copy); cudaStreamDestroy(s->copy); }
    if (s->stream) cudaStreamDestroy(s->stream);
    free(s);
}

// one-shot convenience: a session for exactly this batch (allocation + pinned staging inside)
int ops_beamopt_run_host(const OpsBeamOptParams *p, int64_t B,
                         const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                         const double *L,
                         float *I_values, double *deflections, double *rotations, float *shear,
                         float *moment, int32_t *epochs, float *loss, int32_t *status,
                         int device, float *elapsed_ms)
{
    BeamConsts k;
    int rc = make_consts(p, &k);
    if (rc) return rc;
    if (B < 0) return OPS_E_BADARG;
    if (B == 0) return 0;
    if (!fixed_uy || !L || !I_values || !deflections || !rotations || !shear || !moment || !epochs ||
        !loss || !status || (p->max_forces > 0 && (!force_nodes || !force_vals)))
        return OPS_E_BADARG;
    OpsBeamOptSession *s = nullptr;
    rc = ops_beamopt_session_create(p, B, device, &s);
    if (rc) return rc;
    OpsBeamOptHostArrays a;
    ops_beamopt_session_arrays(s, &a);
    const size_t b = (size_t)B, nn = (size_t)k.nn, n = (size_t)k.n, C = (size_t)p->num_cases;
    const size_t F = (size_t)p->max_forces * C;
    memcpy(a.fixed_uy, fixed_uy, b * nn);
    if (F > 0) { memcpy(a.force_nodes, force_nodes, b * F * 4); memcpy(a.force_vals, force_vals, b * F * 8); }
    memcpy(a.L, L, b * 8);
    rc = ops_beamopt_session_run(s, B, elapsed_ms);
    if (rc == 0) {
        memcpy(I_values, a.I_values, b * n * 4);
        memcpy(deflections, a.deflections, b * C * nn * 8);
        memcpy(rotations, a.rotations, b * C * nn * 8);
        memcpy(shear, a.shear, b * C * n * 4);
        memcpy(moment, a.moment, b * C * n * 4);
        memcpy(epochs, a.epochs, b * 4);
        memcpy(loss, a.loss, b * 4);
        memcpy(status, a.status, b * 4);
    }
    ops_beamopt_session_destroy(s);
    return rc;
}

}  // extern "C"
