/*
 * TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the reference's per-sample
 * moment-of-inertia optimisation loop.  Never linked into, loaded by or called from the product
 * library (openpystruct_b200/csrc); only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may use it.
 *
 * PARITY UNPINNED: the reference (dsmyl6/OpenPyStruct) ships no tests/golden vectors and its FE
 * engine is the third-party, un-pinned `openseespy` wheel (environment.yml:13-14), absent from
 * /root/reference and not installable here.  Its published algorithm is restated:
 *   FP64 half  = OpenSees ElasticBeam2d + LinearCrdTransf2d + Plain constraints + BandSPD (LAPACK
 *                dpbsv: dpbtf2('U') + 2 x dtbsv), call sites SingleCore:89-124,176-190,221-232.
 *   FP32 half  = torch CPU semantics of SingleCore:163-219: torch.sum (ATen SumKernel cascade_sum,
 *                8-lane vectors x 4 ILP rows), autograd partials with M,V frozen, single-tensor Adam
 *                (lerp/addcmul fused, addcdiv not), ExponentialLR, clamp_, early stop.
 * It is pinned against (a) closed-form Euler-Bernoulli answers, (b) the reference's own source
 * executed on oracle/opensees_shim.py (tests/golden/), (c) the real torch ops (oracle/beamopt_port.py).
 * Only deviation from torch: IEEE sqrtf where torch's CPU sqrt (MKL VML) is 1 ulp off on ~0.7 % of
 * inputs.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every fp32/fp64 op below is one IEEE rounding;
 * fused operations are explicit fma()/fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "beamopt_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* FP64 half                                                                                   */
/* ------------------------------------------------------------------------------------------ */

#define REAL double
#define SUFFIX(x) x##_f64
#define RSQRT(x) sqrt(x)
#include "beam_fe.inc"
#undef REAL
#undef SUFFIX
#undef RSQRT

#define REAL long double
#define SUFFIX(x) x##_f80
#define RSQRT(x) sqrtl(x)
#include "beam_fe.inc"
#undef REAL
#undef SUFFIX
#undef RSQRT

/* ------------------------------------------------------------------------------------------ */
/* FP32 half                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* ATen/native/cpu/SumKernel.cpp: multi_row_sum<float[8-lane vector], nrows = 4>. */
static void multi_row_sum_v8x4(const float *x, int64_t size_ilp, float out[4][8])
{
    enum { NUM_LEVELS = 4 };
    int64_t lg = 0;
    while (((int64_t)1 << lg) < size_ilp) ++lg;           /* CeilLog2(size) */
    int64_t level_power = lg / NUM_LEVELS;
    if (level_power < 4) level_power = 4;
    const int64_t level_step = (int64_t)1 << level_power;
    const int64_t level_mask = level_step - 1;
    float acc[NUM_LEVELS][4][8];
    memset(acc, 0, sizeof acc);
    int64_t i = 0;
    for (; i + level_step <= size_ilp;) {
        for (int64_t j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) acc[0][k][l] += x[(4 * i + k) * 8 + l];
        for (int j = 1; j < NUM_LEVELS; ++j) {
            for (int k = 0; k < 4; ++k)
                for (int l = 0; l < 8; ++l) {
                    acc[j][k][l] += acc[j - 1][k][l];
                    acc[j - 1][k][l] = 0.0f;
                }
            const int64_t mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size_ilp; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[0][k][l] += x[(4 * i + k) * 8 + l];
    for (int j = 1; j < NUM_LEVELS; ++j)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[0][k][l] += acc[j][k][l];
    memcpy(out, acc[0], sizeof acc[0]);
}

/* torch.sum of a contiguous fp32 vector on CPU (vectorized_inner_sum -> row_sum -> multi_row_sum). */
float oracle_torch_sum_f32(const float *x, int64_t n)
{
    const int64_t vec_size = n / 8;
    const int64_t size_ilp = vec_size / 4;
    float part[4][8];
    multi_row_sum_v8x4(x, size_ilp, part);
    for (int64_t v = size_ilp * 4; v < vec_size; ++v)
        for (int l = 0; l < 8; ++l) part[0][l] += x[v * 8 + l];
    for (int k = 1; k < 4; ++k)
        for (int l = 0; l < 8; ++l) part[0][l] += part[k][l];
    float acc = 0.0f;
    for (int64_t k = vec_size * 8; k < n; ++k) acc += x[k];
    for (int l = 0; l < 8; ++l) acc += part[0][l];
    return acc;
}

typedef struct {
    float E2, Gf, kf, am, as, epsf, clampf, w1, b2f, omb2f, adam_epsf;
} fp32_consts;

static fp32_consts make_consts(const OracleBeamOptParams *p)
{
    fp32_consts c;
    c.E2 = (float)(2.0 * p->E);               /* 2 * E  is a Python double, wrapped into fp32 when applied */
    c.Gf = (float)p->G;
    c.kf = (float)p->shear_k;
    c.am = (float)p->alpha_moment;
    c.as = (float)p->alpha_shear;
    c.epsf = (float)p->bending_eps;
    c.clampf = (float)p->clamp_min;
    c.w1 = (float)(1.0 - p->beta1);           /* lerp weight */
    c.b2f = (float)p->beta2;
    c.omb2f = (float)(1.0 - p->beta2);
    c.adam_epsf = (float)p->adam_eps;
    return c;
}

/*
 * Loss (SingleCore:195-199) and the gradient autograd produces with M, V constant (SingleCore:202).
 * csq[e] = sum over load cases of M^2, hsq[e] likewise for V (one case in the reference).
 * scratch: 2*n floats.
 */
float oracle_loss_grad_f32(const OracleBeamOptParams *p, int64_t n, const float *I, const float *csq,
                           const float *hsq, float *grad, float *scratch)
{
    const fp32_consts k = make_consts(p);
    float *d = scratch, *q = scratch + n;
    for (int64_t e = 0; e < n; ++e) {
        const float b = k.E2 * I[e] + k.epsf;
        const float c = csq[e];
        d[e] = c / b;
        const float s = sqrtf(I[e]);
        const float gg = k.Gf * (k.kf * s);
        const float h = hsq[e];
        q[e] = h / gg;
        /* div backward: -g * ((a/b)/b); mul-by-scalar backward; pow(0.5) backward: 0.5 * I^-0.5 = 0.5*(1/sqrt) */
        const float gb = ((-k.am) * (d[e] / b)) * k.E2;
        const float gs = ((((-k.as) * (q[e] / gg)) * k.Gf) * k.kf) * (0.5f * (1.0f / sqrtf(I[e])));
        grad[e] = (1.0f + gs) + gb;
    }
    const float sI = oracle_torch_sum_f32(I, n);
    const float sd = oracle_torch_sum_f32(d, n);
    const float sq = oracle_torch_sum_f32(q, n);
    return (sI + k.am * sd) + k.as * sq;
}

/* Per-epoch Adam scalars, computed in double exactly like torch/optim/adam.py (_single_tensor_adam)
 * and torch/optim/lr_scheduler.py (ExponentialLR, chained form), then rounded to fp32 where torch
 * hands them to an fp32 kernel.  table[2*t] = -(lr_t / bias_correction1), table[2*t+1] = sqrt(bias_correction2). */
void oracle_adam_schedule(const OracleBeamOptParams *p, float *table)
{
    double lr = p->lr;
    for (int t = 1; t <= p->max_epochs; ++t) {
        const double bc1 = 1.0 - pow(p->beta1, (double)t);
        const double bc2 = 1.0 - pow(p->beta2, (double)t);
        const double step_size = lr / bc1;
        const double bc2_sqrt = pow(bc2, 0.5);
        table[2 * (t - 1)] = (float)(-step_size);
        table[2 * (t - 1) + 1] = (float)bc2_sqrt;
        lr = lr * p->gamma;
    }
}

void oracle_adam_step_f32(const OracleBeamOptParams *p, int64_t n, float neg_step, float bc2_sqrt,
                          const float *grad, float *I, float *m, float *v)
{
    const fp32_consts k = make_consts(p);
    for (int64_t e = 0; e < n; ++e) {
        const float g = grad[e];
        m[e] = fmaf(k.w1, g - m[e], m[e]);                       /* exp_avg.lerp_(grad, 1-beta1) */
        v[e] = fmaf(k.omb2f * g, g, v[e] * k.b2f);               /* mul_(beta2).addcmul_(g, g, 1-beta2) */
        const float denom = sqrtf(v[e]) / bc2_sqrt + k.adam_epsf;
        float x = I[e] + (neg_step * m[e]) / denom;              /* addcdiv_ */
        I[e] = x < k.clampf ? k.clampf : x;                      /* clamp_(min) ; NaN propagates like torch */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* the loop                                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* fe_precision: 0 = FP64 dpbsv restatement (what the reference's OpenSees BandSPD computes),
 * 1 = the same FE solve in x87 80-bit arithmetic ("truth" for arbitrating ill-conditioned beams). */
int oracle_beamopt_prec(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                        const int32_t *force_nodes, const double *force_vals, const double *L,
                        float *I_out, double *defl, double *rot, float *shear, float *moment,
                        int32_t *epochs, float *loss, int32_t *status, int fe_precision);
int oracle_beamopt_margin(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                          const int32_t *force_nodes, const double *force_vals, const double *L,
                          float *I_out, double *defl, double *rot, float *shear, float *moment,
                          int32_t *epochs, float *loss, int32_t *status, int fe_precision, double *min_margin_ulps);

int oracle_beamopt(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                   const int32_t *force_nodes, const double *force_vals, const double *L,
                   float *I_out, double *defl, double *rot, float *shear, float *moment,
                   int32_t *epochs, float *loss, int32_t *status)
{
    return oracle_beamopt_prec(p, B, fixed_uy, force_nodes, force_vals, L, I_out, defl, rot, shear, moment,
                               epochs, loss, status, 0);
}

int oracle_beamopt_prec(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                        const int32_t *force_nodes, const double *force_vals, const double *L,
                        float *I_out, double *defl, double *rot, float *shear, float *moment,
                        int32_t *epochs, float *loss, int32_t *status, int fe_precision)
{
    return oracle_beamopt_margin(p, B, fixed_uy, force_nodes, force_vals, L, I_out, defl, rot, shear, moment,
                                 epochs, loss, status, fe_precision, (double *)0);
}

/* min_margin_ulps (optional): per beam, the smallest |loss_t - (best - tolerance)| over the run's early-stop tests
 * (SingleCore:211), in units of the fp32 spacing of loss_t -- how many ulps of the loss the closest stop / "new best"
 * decision of the run was away from going the other way (INFINITY when no test had a finite best). */
int oracle_beamopt_margin(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                          const int32_t *force_nodes, const double *force_vals, const double *L,
                          float *I_out, double *defl, double *rot, float *shear, float *moment,
                          int32_t *epochs, float *loss, int32_t *status, int fe_precision, double *min_margin_ulps)
{
    const int nn = p->num_nodes, n = nn - 1, C = p->num_cases, F = p->max_forces;
    if (nn < 2 || C < 1 || F < 0) return -1;
    float *table = malloc(sizeof(float) * 2 * (size_t)(p->max_epochs > 0 ? p->max_epochs : 1));
    float *I = malloc(sizeof(float) * 8 * (size_t)n);
    double *I64 = malloc(sizeof(double) * (size_t)n);
    double *f_uy = malloc(sizeof(double) * (size_t)nn);
    double *V = malloc(sizeof(double) * (size_t)n), *M = malloc(sizeof(double) * (size_t)n);
    void *work = fe_precision ? malloc(sizeof(long double) * fe_work_doubles_f80(nn))
                              : malloc(sizeof(double) * fe_work_doubles_f64(nn));
    float *m = I + n, *v = I + 2 * n, *grad = I + 3 * n, *csq = I + 4 * n, *hsq = I + 5 * n, *scr = I + 6 * n;
    oracle_adam_schedule(p, table);
    for (int64_t b = 0; b < B; ++b) {
        const uint8_t *fx = fixed_uy + b * nn;
        for (int e = 0; e < n; ++e) { I[e] = (float)p->I0; m[e] = 0.0f; v[e] = 0.0f; }
        double best = INFINITY, margin = INFINITY;
        int counter = 0, ep = 0, st = 0;
        float lossf = NAN;
        for (int t = 0; t < p->max_epochs; ++t) {
            for (int e = 0; e < n; ++e) { I64[e] = (double)I[e]; csq[e] = 0.0f; hsq[e] = 0.0f; }
            for (int c = 0; c < C; ++c) {
                memset(f_uy, 0, sizeof(double) * (size_t)nn);
                for (int j = 0; j < F; ++j) {
                    const int32_t nd = force_nodes[(b * C + c) * F + j];
                    if (nd >= 0 && nd < nn) f_uy[nd] += force_vals[(b * C + c) * F + j];
                }
                double *u_c = defl + (b * C + c) * nn, *r_c = rot + (b * C + c) * nn;
                const int frc = fe_precision
                    ? beam_fe_solve_f80(nn, I64, L[b], fx, f_uy, p->udl, p->E, u_c, r_c, V, M, work)
                    : beam_fe_solve_f64(nn, I64, L[b], fx, f_uy, p->udl, p->E, u_c, r_c, V, M, work);
                if (frc) { st = 1; break; }
                float *Vc = shear + (b * C + c) * n, *Mc = moment + (b * C + c) * n;
                for (int e = 0; e < n; ++e) {
                    Vc[e] = (float)V[e];
                    Mc[e] = (float)M[e];
                    const float m2 = Mc[e] * Mc[e], v2 = Vc[e] * Vc[e];
                    csq[e] = c == 0 ? m2 : csq[e] + m2;
                    hsq[e] = c == 0 ? v2 : hsq[e] + v2;
                }
            }
            if (st) break;
            lossf = oracle_loss_grad_f32(p, n, I, csq, hsq, grad, scr);
            oracle_adam_step_f32(p, n, table[2 * t], table[2 * t + 1], grad, I, m, v);
            ++ep;
            if (p->early_stop) {
                const double l = (double)lossf;
                if (isfinite(best) && isfinite(l)) {
                    const double ulp = (double)nextafterf(fabsf(lossf), INFINITY) - (double)fabsf(lossf);
                    const double mg = fabs(l - (best - p->tolerance)) / ulp;
                    if (mg < margin) margin = mg;
                }
                if (l < best - p->tolerance) { best = l; counter = 0; } else { ++counter; }
                if (counter >= p->patience) break;
            }
        }
        if (min_margin_ulps) min_margin_ulps[b] = margin;
        if (p->zero_last_node)
            for (int c = 0; c < C; ++c) { defl[(b * C + c) * nn + nn - 1] = 0.0; rot[(b * C + c) * nn + nn - 1] = 0.0; }
        memcpy(I_out + b * n, I, sizeof(float) * (size_t)n);
        epochs[b] = ep;
        loss[b] = lossf;
        status[b] = st;
    }
    free(table); free(I); free(I64); free(f_uy); free(V); free(M); free(work);
    return 0;
}

/* One solve, no optimiser.  precision: 0 = FP64 (dpbsv restatement), 1 = x87 extended (80-bit) truth. */
int oracle_beam_solve(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                      const int32_t *force_nodes, const double *force_vals, const double *L,
                      const double *I, double *defl, double *rot, double *shear, double *moment,
                      int precision)
{
    const int nn = p->num_nodes, n = nn - 1, F = p->max_forces;
    double *f_uy = malloc(sizeof(double) * (size_t)nn);
    void *work = malloc(sizeof(long double) * fe_work_doubles_f80(nn));
    int rc = 0;
    for (int64_t b = 0; b < B; ++b) {
        memset(f_uy, 0, sizeof(double) * (size_t)nn);
        for (int j = 0; j < F; ++j) {
            const int32_t nd = force_nodes[b * F + j];
            if (nd >= 0 && nd < nn) f_uy[nd] += force_vals[b * F + j];
        }
        int r;
        if (precision == 0)
            r = beam_fe_solve_f64(nn, I + b * n, L[b], fixed_uy + b * nn, f_uy, p->udl, p->E,
                                  defl + b * nn, rot + b * nn, shear + b * n, moment + b * n, work);
        else
            r = beam_fe_solve_f80(nn, I + b * n, L[b], fixed_uy + b * nn, f_uy, p->udl, p->E,
                                  defl + b * nn, rot + b * nn, shear + b * n, moment + b * n, work);
        if (r) rc = 1;
    }
    free(f_uy); free(work);
    return rc;
}
