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

// Eight-lanes-per-beam production iteration (beamopt_lanes.cuh) with the control flow of
// beamopt_lanes_kernel: the phase functions are the device code itself, run lane by lane; the
// __syncwarp points of the kernel are the boundaries between the lane loops.
#include "../../openpystruct_b200/csrc/beamopt_lanes.cuh"

template <int EPL, int NC>
static int lanes_run(const BeamConsts &k, int64_t B, const uint8_t *fixed_uy, const int32_t *force_nodes,
                     const double *force_vals, const double *L, const float *sched, float *I_values,
                     double *defl, double *rot, float *shear, float *moment, int32_t *epochs, float *loss,
                     int32_t *status)
{
    using namespace ops::lanes;
    const int n = k.n, nn = k.nn;
    constexpr int TL = NC * LPB;                       // lanes of a team; column index = case * 8 + lane
    std::vector<Pair> lane_p((size_t)EPL * TL);
    std::vector<double> lane_s((size_t)SCR_SLOTS * TL), grp_d((size_t)GROUP_DOUBLES * NC);
    std::vector<PairF> lane_x((size_t)EPL * TL);
    std::vector<double> lane_b((size_t)XB_ROWS * team_owned_pairs(EPL, NC) * TL);
    std::vector<Pair> tab_p((size_t)TAB_SLOTS / 2 * NC);
    std::vector<int> grp_i((size_t)GROUP_INTS * NC);
    LaneStore ls[NC][LPB];
    GroupStore gs[NC];
    for (int c = 0; c < NC; ++c) {
        for (int l = 0; l < LPB; ++l) {
            const int col = c * LPB + l;
            ls[c][l].ls = TL;
            ls[c][l].mq = lane_p.data() + col;
            ls[c][l].scr = lane_s.data() + col;
            ls[c][l].xc = lane_x.data() + col;
            ls[c][l].xb = lane_b.data() + col;
        }
        gs[c].gs = NC;
        gs[c].tab = reinterpret_cast<double *>(tab_p.data()) + (size_t)TAB_SLOTS * c;
        gs[c].fs.sd = grp_d.data() + c; gs[c].fs.stride = NC;
        gs[c].gd = gs[c].fs.sd + (size_t)FlexStore::NUM_DOUBLES * NC;
        gs[c].fs.si = grp_i.data() + c;
        gs[c].gi = gs[c].fs.si + (size_t)FlexStore::NUM_INTS * NC;
    }
    for (int64_t b = 0; b < B; ++b) {
        static LaneRegs<EPL> rg[NC][LPB];
        FlexBeam fb[NC];
        int bad = 0;
        const uint8_t *fx = fixed_uy + b * nn;
        for (int c = 0; c < NC; ++c) {
            int fnode[FLEX_MAXF];
            double fval[FLEX_MAXF];
            for (int j = 0; j < k.max_forces; ++j) {
                fnode[j] = force_nodes[(b * NC + c) * k.max_forces + j];
                fval[j] = force_vals[(b * NC + c) * k.max_forces + j];
            }
            FlexBeam f0;
            const int rc = flex_setup(k, L[b], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, gs[c].fs, f0);
            group_publish(f0, rc, gs[c]);
            group_table_init(gs[c]);
            bad = group_fetch(k, L[b], gs[c], fb[c]);
            for (int l = 0; l < LPB; ++l) {
                if (!bad) {
                    lane_init<EPL>(k, n, fb[c], gs[c], ls[c][l], l, rg[c][l]);
                    if (NC == 1) lane_pass1<EPL>(rg[c][l], ls[c][l], pass1_consts(fb[c]), false);
                } else lane_reset<EPL>(k, rg[c][l]);
                if constexpr (NC > 1) team_init<EPL, NC>(k, n, ls[c][l], l, c, rg[c][l]);
            }
        }
        int t = 0, counter = 0;
        double best = INFINITY;
        float lossf = NAN, neg_step = 0.0f, bc2_sqrt = 1.0f;
        bool done = (k.max_epochs <= 0) || bad;
        while (!done) {
            neg_step = sched[2 * t]; bc2_sqrt = sched[2 * t + 1];
            int rc = 0;
            static Pair mqk[NC][LPB][EPL];                         // (registers of the lane on the device)
            if constexpr (NC > 1) {
                for (int c = 0; c < NC; ++c)
                    for (int l = 0; l < LPB; ++l) team_pass1<EPL, NC>(rg[c][l], ls[c][l], pass1_consts(fb[c]), c, mqk[c][l]);
            }
            for (int c = 0; c < NC; ++c) {
                for (int l = 0; l < LPB; ++l) lane_reduce(l, fb[c].m, ls[c][l], gs[c]);
                for (int l = LPB - 1; l >= 0; --l) rc |= group_solve(fb[c], gs[c], l);
                if (NC > 1) for (int l = 0; l < LPB; ++l) lane_case_squares<EPL>(rg[c][l], ls[c][l], gs[c], fb[c].invLe, mqk[c][l]);
            }
            float lv[NC][LPB];
            const bool stage_I = (t + 1 >= k.max_epochs) || (k.early_stop && counter + 1 >= k.patience);
            if constexpr (NC == 1) {
                for (int l = 0; l < LPB; ++l)
                    lane_pass<EPL, NC>(k, n, rg[0][l], ls[0][l], gs[0], pass1_consts(fb[0]), fb[0].invLe, l, 0, neg_step, bc2_sqrt,
                                       stage_I);
            } else {
                for (int c = 0; c < NC; ++c)
                    for (int l = 0; l < LPB; ++l) team_owner_update<EPL, NC>(k, rg[c][l], ls[c][l], c, neg_step, bc2_sqrt);
                for (int c = 0; c < NC; ++c)
                    for (int l = 0; l < LPB; ++l) team_loss_sums<EPL, NC>(n, ls[c][l], l, c);
            }
            for (int c = 0; c < NC; ++c)
                for (int l = 0; l < LPB; ++l) lv[c][l] = group_loss(k, n, ls[c][l], l);
            for (int c = 0; c < NC; ++c)
                for (int l = 0; l < LPB; ++l) if (memcmp(&lv[c][l], &lv[0][0], 4) != 0) return -100;   // the team must agree
            lossf = lv[0][0];
            ++t;
            if (rc || !(lossf - lossf == 0.0f)) { bad = 1; done = true; }
            if (k.early_stop) {
                const double l_ = (double)lossf;
                if (l_ < best - k.tol) { best = l_; counter = 0; } else { ++counter; }
                if (counter >= k.patience) done = true;
            }
            if (t >= k.max_epochs) done = true;
            if (NC == 1 && !done && stage_I)
                for (int l = 0; l < LPB; ++l) lane_pass1<EPL>(rg[0][l], ls[0][l], pass1_consts(fb[0]), true);
        }
        const bool fields = (t > 0) && (bad == 0);
        for (int c = 0; c < NC; ++c) {
            const int64_t bc = b * NC + c;
            for (int l = 0; l < LPB; ++l)
                lane_emit_forces<EPL>(n, rg[c][l], ls[c][l], gs[c], fb[c].invLe, l, fields, shear + bc * n, moment + bc * n);
            if constexpr (NC == 1) {
                const ParkedInertia parked = {reinterpret_cast<const float *>(ls[c][0].scr), ls[c][0].ls};
                group_emit_displacements(k, fb[c], gs[c], fields, parked, defl + bc * nn, rot + bc * nn);
                for (int l = 0; l < LPB; ++l) lane_emit_inertias<EPL>(n, rg[c][l], l, I_values + b * n);
            } else {
                group_emit_displacements(k, fb[c], gs[c], fields,
                                         [&](int e) { return (double)team_inertia<NC>(ls[c][0], 0, c, e, 0); },
                                         defl + bc * nn, rot + bc * nn);
                if (c == 0)
                    for (int e = 0; e < n; ++e) I_values[b * n + e] = team_inertia<NC>(ls[0][0], 0, 0, e, t > 0 ? 3 : 0);
            }
        }
        epochs[b] = t; loss[b] = lossf; status[b] = bad;
    }
    return 0;
}

// The peers' copies of one beam's rows (lane_copy_record, the in-kernel dataset gather) run lane by lane on two
// host "GPUs": dest[0] holds the record, dest[1] receives it.  Arrays are caller-owned, whole-dataset sized.
extern "C" int hostsim_copy_record(int n, int nn, int64_t row, int64_t rowc, int first_case,
                                   float *I0, double *defl0, double *rot0, float *shear0, float *moment0,
                                   int32_t *epochs0, float *loss0, int32_t *status0,
                                   float *I1, double *defl1, double *rot1, float *shear1, float *moment1,
                                   int32_t *epochs1, float *loss1, int32_t *status1)
{
    using namespace ops::lanes;
    RecordDest d;
    memset(&d, 0, sizeof(d));
    d.nd = 2;
    d.I[0] = I0; d.defl[0] = defl0; d.rot[0] = rot0; d.shear[0] = shear0; d.moment[0] = moment0;
    d.epochs[0] = epochs0; d.loss[0] = loss0; d.status[0] = status0;
    d.I[1] = I1; d.defl[1] = defl1; d.rot[1] = rot1; d.shear[1] = shear1; d.moment[1] = moment1;
    d.epochs[1] = epochs1; d.loss[1] = loss1; d.status[1] = status1;
    for (int l = 0; l < LPB; ++l) lane_copy_record(n, nn, l, d, row, rowc, first_case != 0);
    return 0;
}

// Tensor-memory instance (beamopt_lanes_tm.cu): ONE WARP of the kernel -- four groups taking beams from a counter --
// with the kernel's loop body: every phase that touches tensor memory (HomeTm; here a plain array of words per lane) is
// entered by all 32 lanes on a warp-uniform condition, the lanes without work compute on whatever they hold.
template <int EPL>
static int lanes_run_tm(const BeamConsts &k, int64_t B, const uint8_t *fixed_uy, const int32_t *force_nodes,
                        const double *force_vals, const double *L, const float *sched, float *I_values,
                        double *defl, double *rot, float *shear, float *moment, int32_t *epochs, float *loss,
                        int32_t *status, int nbp)
{
    using namespace ops::lanes;
    const int n = k.n, nn = k.nn;
    constexpr int NG = 4, TL = NG * LPB, COLS = HomeTm::columns(EPL);
    std::vector<double> lane_s((size_t)SCR_SLOTS * TL), grp_d((size_t)GROUP_DOUBLES * NG);
    std::vector<Pair> tab_p((size_t)TAB_SLOTS / 2 * NG);
    std::vector<int> grp_i((size_t)GROUP_INTS * NG);
    std::vector<unsigned int> words((size_t)COLS * TL, 0xdeadbeefu);
    LaneStore ls[NG][LPB];
    GroupStore gs[NG];
    HomeTm hm[NG][LPB];
    static LaneRegs<EPL> rg[NG][LPB];
    for (int g = 0; g < NG; ++g) {
        for (int l = 0; l < LPB; ++l) {
            const int col = g * LPB + l;
            ls[g][l].ls = TL; ls[g][l].mq = nullptr; ls[g][l].scr = lane_s.data() + col; ls[g][l].xc = nullptr; ls[g][l].xb = nullptr;
            hm[g][l].base = 0; hm[g][l].w = words.data() + (size_t)COLS * col;
            lane_reset<EPL>(k, rg[g][l]);
        }
        gs[g].gs = NG;
        gs[g].tab = reinterpret_cast<double *>(tab_p.data()) + (size_t)TAB_SLOTS * g;
        gs[g].fs.sd = grp_d.data() + g; gs[g].fs.stride = NG;
        gs[g].gd = gs[g].fs.sd + (size_t)FlexStore::NUM_DOUBLES * NG;
        gs[g].fs.si = grp_i.data() + g;
        gs[g].gi = gs[g].fs.si + (size_t)FlexStore::NUM_INTS * NG;
    }
    struct GroupState {
        FlexBeam fb; Pass1Consts pc; int64_t b; bool have, exhausted, resume, fresh; int t, counter, bad; double best; float lossf;
    } st[NG];
    for (int g = 0; g < NG; ++g) {
        memset(&st[g], 0, sizeof st[g]);
        st[g].b = -1; st[g].best = INFINITY; st[g].lossf = NAN;
    }
    int64_t next = 0;
    while (true) {
        bool any_have = false, any_fresh = false, any_sums = false;
        for (int g = 0; g < NG; ++g) {
            GroupState &s = st[g];
            s.fresh = false;
            if (!s.have && !s.exhausted) {
                const int64_t nb = next++;
                if (nb < B) {
                    s.b = nb; s.have = true; s.t = 0; s.counter = 0; s.best = INFINITY; s.lossf = NAN;
                    int fnode[FLEX_MAXF];
                    double fval[FLEX_MAXF];
                    for (int j = 0; j < k.max_forces; ++j) { fnode[j] = force_nodes[nb * k.max_forces + j]; fval[j] = force_vals[nb * k.max_forces + j]; }
                    const uint8_t *fx = fixed_uy + nb * nn;
                    FlexBeam f0;
                    const int rc = flex_setup(k, L[nb], [&](int i) { return fx[i] != 0; }, k.max_forces, fnode, fval, gs[g].fs, f0);
                    group_publish(f0, rc, gs[g]);
                    group_table_init(gs[g]);
                    s.bad = group_fetch(k, L[nb], gs[g], s.fb);
                    s.pc = pass1_consts(s.fb);
                    for (int l = 0; l < LPB; ++l) {
                        if (!s.bad) lane_init<EPL, HomeTm>(k, n, s.fb, gs[g], ls[g][l], l, rg[g][l], hm[g][l]);
                        else lane_reset<EPL>(k, rg[g][l]);
                    }
                    s.fresh = !s.bad;
                } else s.exhausted = true;
            }
            any_have = any_have || s.have; any_fresh = any_fresh || s.fresh; any_sums = any_sums || s.fresh || s.resume;
        }
        if (!any_have) break;
        if (any_fresh)
            for (int g = 0; g < NG; ++g) for (int l = 0; l < LPB; ++l) tm_commit<EPL>(hm[g][l], ls[g][l], st[g].fresh);
        if (any_sums)
            for (int g = 0; g < NG; ++g) for (int l = 0; l < LPB; ++l)
                lane_pass1<EPL, HomeTm>(rg[g][l], ls[g][l], st[g].pc, st[g].resume, hm[g][l], st[g].fresh || st[g].resume);
        bool run[NG], done[NG], stage_I[NG], any_run = false, any_rec = false;
        int rc[NG];
        for (int g = 0; g < NG; ++g) {
            GroupState &s = st[g];
            s.resume = false;
            run[g] = s.have && k.max_epochs > 0 && s.bad == 0;
            done[g] = s.have && !run[g];
            rc[g] = 0;
            any_run = any_run || run[g];
            if (run[g]) {
                for (int l = 0; l < LPB; ++l) lane_reduce(l, s.fb.m, ls[g][l], gs[g]);
                for (int l = LPB - 1; l >= 0; --l) rc[g] |= group_solve(s.fb, gs[g], l);
            }
            stage_I[g] = run[g] && ((s.t + 1 >= k.max_epochs) || (k.early_stop && s.counter + 1 >= k.patience));
        }
        if (any_run) {
            for (int g = 0; g < NG; ++g) {
                GroupState &s = st[g];
                const float neg_step = run[g] ? sched[2 * s.t] : 0.0f, bc2_sqrt = run[g] ? sched[2 * s.t + 1] : 1.0f;
                for (int l = 0; l < LPB; ++l) {
                    if (nbp == 1) lane_pass<EPL, 1, 1, HomeTm>(k, n, rg[g][l], ls[g][l], gs[g], s.pc, s.fb.invLe, l, 0, neg_step, bc2_sqrt, stage_I[g], hm[g][l]);
                    else if (nbp == 3) lane_pass<EPL, 1, 3, HomeTm>(k, n, rg[g][l], ls[g][l], gs[g], s.pc, s.fb.invLe, l, 0, neg_step, bc2_sqrt, stage_I[g], hm[g][l]);
                    else lane_pass<EPL, 1, 2, HomeTm>(k, n, rg[g][l], ls[g][l], gs[g], s.pc, s.fb.invLe, l, 0, neg_step, bc2_sqrt, stage_I[g], hm[g][l]);
                    if (s.have && !run[g]) lane_reset<EPL>(k, rg[g][l]);
                }
            }
        }
        for (int g = 0; g < NG; ++g) {
            GroupState &s = st[g];
            if (run[g]) {
                float lv[LPB];
                for (int l = 0; l < LPB; ++l) lv[l] = group_loss(k, n, ls[g][l], l);
                for (int l = 1; l < LPB; ++l) if (memcmp(&lv[l], &lv[0], 4) != 0) return -100;
                s.lossf = lv[0];
                ++s.t;
                if (rc[g] || !(s.lossf - s.lossf == 0.0f)) { s.bad = 1; done[g] = true; }
                if (k.early_stop) {
                    const double l_ = (double)s.lossf;
                    if (l_ < s.best - k.tol) { s.best = l_; s.counter = 0; } else { ++s.counter; }
                    if (s.counter >= k.patience) done[g] = true;
                }
                if (s.t >= k.max_epochs) done[g] = true;
            }
            any_rec = any_rec || (s.have && done[g]);
        }
        for (int g = 0; g < NG; ++g) {
            GroupState &s = st[g];
            const bool rec = s.have && done[g];
            const bool fields = (s.t > 0) && (s.bad == 0);
            const int64_t row = s.b < 0 ? 0 : s.b;
            if (any_rec)
                for (int l = 0; l < LPB; ++l)
                    lane_emit_forces<EPL, HomeTm>(n, rg[g][l], ls[g][l], gs[g], s.fb.invLe, l, fields, shear + row * n, moment + row * n,
                                                  hm[g][l], rec);
            if (rec) {
                const ParkedInertia parked = {reinterpret_cast<const float *>(ls[g][0].scr), ls[g][0].ls};
                group_emit_displacements(k, s.fb, gs[g], fields, parked, defl + row * nn, rot + row * nn);
                for (int l = 0; l < LPB; ++l) lane_emit_inertias<EPL>(n, rg[g][l], l, I_values + row * n);
                epochs[row] = s.t; loss[row] = s.lossf; status[row] = s.bad;
                s.have = false;
            } else if (stage_I[g]) {
                s.resume = true;
            }
        }
    }
    return 0;
}

extern "C" int hostsim_beamopt_lanes_tm(const OpsBeamOptParams *p, int nbp, int64_t B, const uint8_t *fixed_uy,
                                        const int32_t *force_nodes, const double *force_vals, const double *L,
                                        const float *sched, float *I_values, double *defl, double *rot,
                                        float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    if (p->max_forces > FLEX_MAXF || p->num_cases != 1) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    if (k.n <= 64 || k.n > 104) return OPS_E_UNSUPP;
    return lanes_run_tm<13>(k, B, fixed_uy, force_nodes, force_vals, L, sched, I_values, defl, rot, shear, moment, epochs,
                            loss, status, nbp);
}

extern "C" int hostsim_beamopt_lanes(const OpsBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                                     const int32_t *force_nodes, const double *force_vals, const double *L,
                                     const float *sched, float *I_values, double *defl, double *rot,
                                     float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    if (p->max_forces > FLEX_MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
#define RUN(EPL, NC) lanes_run<EPL, NC>(k, B, fixed_uy, force_nodes, force_vals, L, sched, I_values, defl, rot, shear, \
                                        moment, epochs, loss, status)
    if (p->num_cases == 1) {
        if (k.n <= 32) return RUN(4, 1);
        if (k.n <= 64) return RUN(8, 1);
        if (k.n <= 104) return RUN(13, 1);
        if (k.n <= 168) return RUN(21, 1);
    } else if (k.n <= 104) {
        if (p->num_cases == 2) return k.n <= 32 ? RUN(4, 2) : (k.n <= 64 ? RUN(8, 2) : RUN(13, 2));
        if (p->num_cases == 4) return k.n <= 32 ? RUN(4, 4) : (k.n <= 64 ? RUN(8, 4) : RUN(13, 4));
        if (p->num_cases == 8) return k.n <= 32 ? RUN(4, 8) : (k.n <= 64 ? RUN(8, 8) : RUN(13, 8));
    }
#undef RUN
    return OPS_E_UNSUPP;
}

// Shared-memory-state iteration (beamopt_wide.cuh) with the control flow of beamopt_wide_kernel: the
// per-lane functions are the device code itself, run lane by lane; the two cross-lane steps (butterfly
// sum of a closing span, final loss combine) are done on arrays here and with shuffles on the device.
#include "../../openpystruct_b200/csrc/beamopt_wide.cuh"

template <int LPB, int N>
static void wide_host_batch(const BeamConsts &k, const FlexBeam &fb, const ops::wide::WideShape &sh,
                            const ops::wide::WideStore &ws, int kb, bool run, const float *Icur, float *Inew,
                            const ops::wide::SweepConsts &sc, ops::wide::LaneCtx<LPB> *cx)
{
    using namespace ops::wide;
    static BatchOut<N> bo[LPB];
    for (int l = 0; l < LPB; ++l) sweep_batch<LPB, N>(k, fb, sh, ws, l, kb, run, Icur, Inew, sc, cx[l], bo[l]);
    for (int i = 0; i < N; ++i) {
        double x[LPB][NSUM], aold[LPB][NSUM];
        int sold[LPB];
        for (int l = 0; l < LPB; ++l) {
            slot_terms<N>(bo[l], i, x[l]);
            slot_accumulate<LPB>(cx[l], x[l], bo[l].sp[i], aold[l], sold[l]);
        }
        if (!bo[0].close[i]) continue;
        for (int j = 0; j < NSPAN; ++j) {
            if (ws.gi[GI_CLOSE + j] != kb + i) continue;
            double v[LPB][NSUM];
            for (int l = 0; l < LPB; ++l) close_value(aold[l], sold[l], x[l], bo[l].sp[i], j, v[l]);
            for (int step = 1; step < LPB; step <<= 1) {              // butterfly: v_l += v_{l ^ step}
                double nv[LPB][NSUM];
                for (int l = 0; l < LPB; ++l)
                    for (int w = 0; w < NSUM; ++w) nv[l][w] = v[l][w] + v[l ^ step][w];
                memcpy(v, nv, sizeof v);
            }
            for (int w = 0; w < NSUM; ++w) ws.tot[j * NSUM + w] = v[0][w];
        }
    }
}

template <int LPB>
static void wide_host_sweep(const BeamConsts &k, const FlexBeam &fb, const ops::wide::WideShape &sh,
                            const ops::wide::WideStore &ws, bool run, const float *Icur, float *Inew,
                            const ops::wide::SweepConsts &sc, ops::wide::LaneCtx<LPB> *cx)
{
    using namespace ops::wide;
    for (int l = 0; l < LPB; ++l) ctx_reset<LPB>(cx[l], l);
    int kb = 0;
    for (; kb + NB <= sh.K; kb += NB) wide_host_batch<LPB, NB>(k, fb, sh, ws, kb, run, Icur, Inew, sc, cx);
    for (; kb < sh.K; ++kb) wide_host_batch<LPB, 1>(k, fb, sh, ws, kb, run, Icur, Inew, sc, cx);
    for (int l = 0; l < LPB; ++l) ctx_finish<LPB>(sh, cx[l]);
}

template <int LPB>
static int wide_run(const BeamConsts &k, int64_t B, const uint8_t *fixed_uy, const int32_t *force_nodes,
                    const double *force_vals, const double *L, const float *sched, float *I_values, double *defl,
                    double *rot, float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    using namespace ops::wide;
    const int n = k.n, nn = k.nn;
    if (!wide_shape_ok<LPB>(n)) return OPS_E_UNSUPP;
    const WideShape sh = wide_shape<LPB>(n);
    std::vector<double> raw(wide_beam_bytes<LPB>(n) / 8 + 2);
    WideStore ws;
    wide_carve<LPB>(reinterpret_cast<unsigned char *>(raw.data()), n, ws);
    static LaneCtx<LPB> cx[LPB];
    static SegStatics st[LPB];
    for (int64_t b = 0; b < B; ++b) {
        int fnode[FLEX_MAXF];
        double fval[FLEX_MAXF];
        for (int j = 0; j < k.max_forces; ++j) {
            fnode[j] = force_nodes[b * k.max_forces + j];
            fval[j] = force_vals[b * k.max_forces + j];
        }
        const uint8_t *fx = fixed_uy + b * nn;
        wide_setup<LPB>(k, L[b], [&](int i) { return fx[i] != 0; }, fnode, fval, ws);
        FlexBeam fb;
        int bad = wide_fetch(k, L[b], ws, fb);
        for (int l = 0; l < LPB; ++l) wide_fetch_statics<LPB>(ws, l, st[l]);
        const bool setup_bad = bad != 0;
        SweepConsts sc;
        sc.G2 = 6.0 * fb.wl2h; sc.H2 = 3.0 * fb.wl2h; sc.neg_step = 0.0f; sc.bc2_sqrt = 1.0f; sc.rbc = 1.0f;
        if (!bad) {
            for (int l = 0; l < LPB; ++l) wide_lane_init<LPB>(k, sh, ws, l);
            wide_host_sweep<LPB>(k, fb, sh, ws, false, ws.I0, ws.I0 + ws.el, sc, cx);
        }
        int t = 0, counter = 0;
        double best = INFINITY;
        float lossf = NAN;
        bool done = (k.max_epochs <= 0) || bad;
        while (!done) {
            sc.neg_step = sched[2 * t]; sc.bc2_sqrt = sched[2 * t + 1]; sc.rbc = 1.0f / sc.bc2_sqrt;
            int rc = 0;
            for (int l = LPB - 1; l >= 0; --l) rc |= wide_solve<LPB>(fb, ws, l, st[l]);
            wide_host_sweep<LPB>(k, fb, sh, ws, true, ws.I0 + (t & 1) * ws.el, ws.I0 + ((t + 1) & 1) * ws.el, sc, cx);
            float acc[3][LPB], left[3][LPB];
            for (int w = 0; w < 3; ++w)
                for (int l = 0; l < LPB; ++l) {
                    acc[w][l] = LPB == 32 ? cx[l].acc[w][0] : ctx_rowsum<LPB>(cx[l], w);
                    left[w][l] = cx[l].left[w];
                }
            lossf = wide_loss_arrays<LPB>(k, sh, acc, left);
            ++t;
            if (rc || !(lossf - lossf == 0.0f)) { bad = 1; done = true; }
            if (k.early_stop) {
                const double l_ = (double)lossf;
                if (l_ < best - k.tol) { best = l_; counter = 0; } else { ++counter; }
                if (counter >= k.patience) done = true;
            }
            if (t >= k.max_epochs) done = true;
        }
        const bool fields = (t > 0) && (bad == 0);
        for (int l = 0; l < LPB; ++l)
            wide_emit_lane<LPB>(fb, sh, ws, l, fields, ws.I0 + (t & 1) * ws.el, shear + b * n, moment + b * n,
                                setup_bad ? nullptr : I_values + b * n);
        if (setup_bad) for (int e = 0; e < n; ++e) I_values[b * n + e] = k.I0f;
        wide_emit_displacements(k, fb, ws, ws.I0 + ((t + 1) & 1) * ws.el, fields, defl + b * nn, rot + b * nn);
        epochs[b] = t; loss[b] = lossf; status[b] = bad;
    }
    return 0;
}

extern "C" int hostsim_beamopt_wide(const OpsBeamOptParams *p, int lanes_per_beam, int64_t B, const uint8_t *fixed_uy,
                                    const int32_t *force_nodes, const double *force_vals, const double *L,
                                    const float *sched, float *I_values, double *defl, double *rot,
                                    float *shear, float *moment, int32_t *epochs, float *loss, int32_t *status)
{
    if (p->num_cases != 1 || p->max_forces > FLEX_MAXF) return OPS_E_UNSUPP;
    BeamConsts k;
    consts_from(p, &k);
    if (lanes_per_beam == 8)
        return wide_run<8>(k, B, fixed_uy, force_nodes, force_vals, L, sched, I_values, defl, rot, shear, moment, epochs,
                           loss, status);
    if (lanes_per_beam == 32)
        return wide_run<32>(k, B, fixed_uy, force_nodes, force_vals, L, sched, I_values, defl, rot, shear, moment, epochs,
                            loss, status);
    return OPS_E_UNSUPP;
}
