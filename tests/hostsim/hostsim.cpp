// DEBUG AID FOR TESTS ONLY -- not part of the product, never loaded by openpystruct_b200.
//
// Compiles the kernel's per-beam arithmetic (openpystruct_b200/csrc/beamopt_core.cuh, which has no
// CUDA intrinsics) with g++ so that the `-m "not gpu"` suite can check the exact operation order the
// GPU executes against the CPU oracle in the build container, where no GPU exists.  It mirrors the
// control flow of beamopt_kernel for one beam at a time (stride 1 storage).  The product library has
// no host compute path; this file is linked into tests/hostsim/_build/libhostsim.so only.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/openpystruct_b200.h"
#include "../../openpystruct_b200/csrc/beamopt_core.cuh"
#include "../../openpystruct_b200/csrc/beamopt_flex.cuh"

using namespace ops;

static void consts_from(const OpsBeamOptParams *p, BeamConsts *k)
{
    k->nn = p->num_nodes; k->n = p->num_nodes - 1; k->max_forces = p->max_forces;
    k->max_epochs = p->max_epochs; k->patience = p->patience; k->early_stop = p->early_stop;
    k->zero_last_node = p->zero_last_node; k->E = p->E; k->udl = p->udl; k->tol = p->tolerance;
    k->I0f = (float)p->I0; k->E2 = (float)(2.0 * p->E); k->Gf = (float)p->G; k->kf = (float)p->shear_k;
    k->am = (float)p->alpha_moment; k->as_ = (float)p->alpha_shear; k->epsf = (float)p->bending_eps;
    k->clampf = (float)p->clamp_min; k->w1 = (float)(1.0 - p->beta1); k->b2f = (float)p->beta2;
    k->omb2f = (float)(1.0 - p->beta2); k->adam_epsf = (float)p->adam_eps;
}

extern "C" int hostsim_beamopt(const OpsBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                               const int32_t *force_nodes, const double *force_vals, const double *L,
                               const float *sched, float *I_values, double *defl, double *rot,
                               float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    constexpr int MAXF = 8;
    if (p->num_cases != 1 || p->max_forces > MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    const int n = k.n, nn = k.nn;
    std::vector<double> d((size_t)5 * nn);
    std::vector<float> f((size_t)3 * n);
    BeamStore st{d.data(), f.data(), 1};
    for (int64_t b = 0; b < B; ++b) {
        BeamInputs<MAXF> in;
        beam_geometry<MAXF>(k, L[b], in);
        for (int j = 0; j < MAXF; ++j) {
            const bool on = j < k.max_forces;
            in.fnode[j] = on ? force_nodes[b * k.max_forces + j] : -1;
            in.fval[j] = on ? force_vals[b * k.max_forces + j] : 0.0;
        }
        const uint8_t *fx = fixed_uy + b * nn;
        auto fixed = [&](int i) { return fx[i] != 0; };
        for (int e = 0; e < n; ++e) { st.F(e) = k.I0f; st.F(n + e) = 0.0f; st.F(2 * n + e) = 0.0f; }
        int t = 0, counter = 0, bad = 0;
        double best = INFINITY;
        float lossf = NAN;
        while (t < k.max_epochs) {
            int rc = 0;
            lossf = beam_iteration<MAXF>(k, in, st, fixed, sched[2 * t], sched[2 * t + 1], &rc);
            ++t;
            if (rc) { bad = 1; break; }
            if (k.early_stop) {
                const double l = (double)lossf;
                if (l < best - k.tol) { best = l; counter = 0; } else { ++counter; }
                if (counter >= k.patience) break;
            }
        }
        for (int e = 0; e < n; ++e) {
            I_values[b * n + e] = st.F(e);
            moment[b * n + e] = t > 0 ? st.pairf(5 * e)[0] : 0.0f;
            shear[b * n + e] = t > 0 ? st.pairf(5 * e)[1] : 0.0f;
        }
        for (int i = 0; i < nn; ++i) {
            const bool z = (t == 0) || (k.zero_last_node && i == nn - 1);
            defl[b * nn + i] = z ? 0.0 : st.D(5 * i + 3);
            rot[b * nn + i] = z ? 0.0 : st.D(5 * i + 4);
        }
        epochs[b] = t; loss[b] = lossf; status[b] = bad;
    }
    return 0;
}

extern "C" int hostsim_beamsolve(const OpsBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                                 const int32_t *force_nodes, const double *force_vals, const double *L,
                                 const double *I, double *defl, double *rot, double *shear, double *moment,
                                 int32_t *status)
{
    constexpr int MAXF = 8;
    if (p->max_forces > MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    const int n = k.n, nn = k.nn;
    std::vector<double> d((size_t)5 * nn);
    BeamStore st{d.data(), nullptr, 1};
    for (int64_t b = 0; b < B; ++b) {
        BeamInputs<MAXF> in;
        beam_geometry<MAXF>(k, L[b], in);
        for (int j = 0; j < MAXF; ++j) {
            const bool on = j < k.max_forces;
            in.fnode[j] = on ? force_nodes[b * k.max_forces + j] : -1;
            in.fval[j] = on ? force_vals[b * k.max_forces + j] : 0.0;
        }
        const uint8_t *fx = fixed_uy + b * nn;
        const double *Ib = I + b * n;
        auto fixed = [&](int i) { return fx[i] != 0; };
        auto inertia = [&](int e) { return Ib[e]; };
        int rc = factor_forward<MAXF>(k, in, st, fixed, inertia);
        rc |= solve_backward<MAXF>(k, in, st, fixed, inertia, [&](int e, double V, double M) {
            shear[b * n + e] = V; moment[b * n + e] = M; });
        for (int i = 0; i < nn; ++i) { defl[b * nn + i] = st.D(5 * i + 3); rot[b * nn + i] = st.D(5 * i + 4); }
        status[b] = rc;
    }
    return 0;
}

// Three-moment production iteration (beamopt_flex.cuh) with the control flow of beamopt_flex_kernel.
extern "C" int hostsim_beamopt_flex(const OpsBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                                    const int32_t *force_nodes, const double *force_vals, const double *L,
                                    const float *sched, float *I_values, double *defl, double *rot,
                                    float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    if (p->num_cases != 1 || p->max_forces > FLEX_MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    const int n = k.n, nn = k.nn;
    std::vector<double> sd(FlexStore::NUM_DOUBLES);
    std::vector<int> si(FlexStore::NUM_INTS);
    std::vector<float> f((size_t)4 * n + 128);
    FlexStore fs{sd.data(), si.data(), 1};
    for (int64_t b = 0; b < B; ++b) {
        OptState os;
        os.Icur = f.data(); os.Inext = f.data() + n; os.m = f.data() + 2 * n; os.v = f.data() + 3 * n;
        os.sI = os.sIn = os.sm = os.sv = 1;
        os.accd = f.data() + 4 * n; os.accq = os.accd + 64; os.sacc = 1;
        FlexBeam fb;
        int fnode[FLEX_MAXF];
        double fval[FLEX_MAXF];
        for (int j = 0; j < k.max_forces; ++j) {
            fnode[j] = force_nodes[b * k.max_forces + j];
            fval[j] = force_vals[b * k.max_forces + j];
        }
        const uint8_t *fx = fixed_uy + b * nn;
        int bad = flex_setup(k, L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, fs, fb);
        for (int e = 0; e < n; ++e) { os.Icur[e] = k.I0f; os.Inext[e] = k.I0f; os.m[e] = 0.0f; os.v[e] = 0.0f; }
        int t = 0, counter = 0;
        double best = INFINITY;
        float lossf = NAN;
        while (t < k.max_epochs && !bad) {
            int rc = 0;
            lossf = flex_iteration(k, fb, fs, os, sched[2 * t], sched[2 * t + 1], &rc);
            ++t;
            std::swap(os.Icur, os.Inext);
            if (rc || !(lossf - lossf == 0.0f)) { bad = 1; break; }
            if (k.early_stop) {
                const double l = (double)lossf;
                if (l < best - k.tol) { best = l; counter = 0; } else { ++counter; }
                if (counter >= k.patience) break;
            }
        }
        const bool fields = (t > 0) && (bad == 0);
        for (int e = 0; e < n; ++e) I_values[b * n + e] = os.Icur[e];
        if (fields) {
            flex_forces_march(k, fb, fs, [&](int e, double V, double M) {
                shear[b * n + e] = (float)V; moment[b * n + e] = (float)M; });
            flex_deflections_march(k, fb, fs, [&](int e) { return (double)os.Inext[e]; },
                                   [&](int i, double u, double th) {
                                       const bool z = k.zero_last_node && i == nn - 1;
                                       defl[b * nn + i] = z ? 0.0 : u; rot[b * nn + i] = z ? 0.0 : th; });
        } else {
            for (int e = 0; e < n; ++e) { shear[b * n + e] = 0.0f; moment[b * n + e] = 0.0f; }
            for (int i = 0; i < nn; ++i) { defl[b * nn + i] = 0.0; rot[b * nn + i] = 0.0; }
        }
        epochs[b] = t; loss[b] = lossf; status[b] = bad;
    }
    return 0;
}

// One three-moment solve for given inertias, FP64 outputs.
extern "C" int hostsim_beamsolve_flex(const OpsBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                                      const int32_t *force_nodes, const double *force_vals, const double *L,
                                      const double *I, double *defl, double *rot, double *shear, double *moment,
                                      int32_t *status)
{
    if (p->max_forces > FLEX_MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    const int n = k.n, nn = k.nn;
    std::vector<double> sd(FlexStore::NUM_DOUBLES);
    std::vector<int> si(FlexStore::NUM_INTS);
    FlexStore fs{sd.data(), si.data(), 1};
    for (int64_t b = 0; b < B; ++b) {
        FlexBeam fb;
        int fnode[FLEX_MAXF];
        double fval[FLEX_MAXF];
        for (int j = 0; j < k.max_forces; ++j) {
            fnode[j] = force_nodes[b * k.max_forces + j];
            fval[j] = force_vals[b * k.max_forces + j];
        }
        const uint8_t *fx = fixed_uy + b * nn;
        const double *Ib = I + b * n;
        int bad = flex_setup(k, L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, fs, fb);
        if (!bad) {
            bad = flex_support_moments(k, fb, fs, [&](int e) { return Ib[e]; });
            flex_forces_march(k, fb, fs, [&](int e, double V, double M) { shear[b * n + e] = V; moment[b * n + e] = M; });
            flex_deflections_march(k, fb, fs, [&](int e) { return Ib[e]; },
                                   [&](int i, double u, double th) { defl[b * nn + i] = u; rot[b * nn + i] = th; });
        }
        status[b] = bad;
    }
    return 0;
}
