/* TEST INFRASTRUCTURE ONLY -- C interface of the CPU oracle (see beamopt_oracle.c). */
#ifndef BEAMOPT_ORACLE_H
#define BEAMOPT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same fields, order and meaning as OpsBeamOptParams in include/openpystruct_b200.h so that the
 * tests can hand one ctypes structure to both sides. */
typedef struct OracleBeamOptParams {
    int32_t struct_size;
    int32_t num_nodes;       /* nn; elements n = nn - 1 */
    int32_t num_cases;       /* C load cases sharing one I vector (reference: 1) */
    int32_t max_forces;      /* width of the force_nodes / force_vals rows */
    int32_t max_epochs;      /* max_e */
    int32_t patience;
    int32_t early_stop;      /* 0 = run exactly max_epochs */
    int32_t zero_last_node;  /* MultiCore:222-223 */
    int32_t solver;          /* ignored by the oracle (product-side selection) */
    int32_t reserved;
    double E, G, udl, I0, lr, gamma, alpha_moment, alpha_shear, tolerance;
    double shear_k, bending_eps, clamp_min, beta1, beta2, adam_eps;
} OracleBeamOptParams;

float oracle_torch_sum_f32(const float *x, int64_t n);
float oracle_loss_grad_f32(const OracleBeamOptParams *p, int64_t n, const float *I, const float *csq,
                           const float *hsq, float *grad, float *scratch);
void oracle_adam_schedule(const OracleBeamOptParams *p, float *table);
void oracle_adam_step_f32(const OracleBeamOptParams *p, int64_t n, float neg_step, float bc2_sqrt,
                          const float *grad, float *I, float *m, float *v);
int oracle_beamopt(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                   const int32_t *force_nodes, const double *force_vals, const double *L,
                   float *I_out, double *defl, double *rot, float *shear, float *moment,
                   int32_t *epochs, float *loss, int32_t *status);
int oracle_beamopt_prec(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                        const int32_t *force_nodes, const double *force_vals, const double *L,
                        float *I_out, double *defl, double *rot, float *shear, float *moment,
                        int32_t *epochs, float *loss, int32_t *status, int fe_precision);
/* oracle_beamopt_prec + per beam the smallest early-stop decision margin of the run, in fp32 ulps of the loss */
int oracle_beamopt_margin(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                          const int32_t *force_nodes, const double *force_vals, const double *L,
                          float *I_out, double *defl, double *rot, float *shear, float *moment,
                          int32_t *epochs, float *loss, int32_t *status, int fe_precision, double *min_margin_ulps);
int oracle_beam_solve(const OracleBeamOptParams *p, int64_t B, const uint8_t *fixed_uy,
                      const int32_t *force_nodes, const double *force_vals, const double *L,
                      const double *I, double *defl, double *rot, double *shear, double *moment,
                      int precision);

#ifdef __cplusplus
}
#endif
#endif
