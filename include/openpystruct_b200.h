/*
 * openpystruct_b200 -- C ABI of the B200-native beam moment-of-inertia optimiser.
 *
 * This is the drop-in boundary for ONE hot path of dsmyl6/OpenPyStruct: the per-sample optimisation
 * loop of the training-data generators.  Each entry point states the reference interface it replaces
 * (file:line into the reference repository):
 *
 *   generate_sample()'s epoch loop      OpenPyStruct_BeamOpt_training_SingleCore.py:163-232
 *                                       OpenPyStruct_BeamOpt_training_MultiCore.py:165-223
 *                                       OpenPyStruct_BeamOpt_training_GPU.py:173-251
 *                                       OpenPyStruct_BeamOpt.py:180-237
 *   which today crosses into native code through ~520 OpenSeesPy calls per epoch
 *   (setup_model SingleCore:89-124, ops.analyze :182, ops.eleResponse :189-190, ops.nodeDisp :224-232)
 *   and ~25 eager torch ops (loss :195-199, backward :202, Adam :203, ExponentialLR :204, clamp :208).
 *
 * Conventions
 *   - plain C, no C++/torch types; all arrays contiguous, row-major, BEAM-MAJOR;
 *   - every buffer is caller-owned; the library never allocates device memory in the *_launch calls;
 *   - *_launch calls are stream-ordered, re-entrant and never synchronise the device;
 *   - return value: 0 ok, <0 invalid argument (OPS_E_*), >0 a cudaError_t value;
 *   - per-beam numerical failure (non-SPD pivot / non-finite result; the reference's
 *     `analyze() != 0 -> return None`, MultiCore:184-186) is reported in status[b], not as an error;
 *   - there is NO CPU implementation behind this ABI: without a CUDA device every compute entry fails.
 */
#ifndef OPENPYSTRUCT_B200_H
#define OPENPYSTRUCT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPS_E_BADARG   (-1)   /* null pointer, non-positive size, struct_size mismatch */
#define OPS_E_UNSUPP   (-2)   /* configuration outside what the kernels implement */
#define OPS_E_WORKSPACE (-3)  /* workspace too small */

/* per-beam status[] values */
#define OPS_STATUS_OK          0
#define OPS_STATUS_SINGULAR    1   /* mechanism / non-SPD pivot / non-finite result: drop the sample */
#define OPS_STATUS_UNSUPPORTED 3   /* OPS_SOLVER_THREE_MOMENT* only: more than 5 rollers; use OPS_SOLVER_BAND_LDLT */

/* solver selection */
#define OPS_SOLVER_THREE_MOMENT 0  /* exact Schur complement of K onto the support moments (default, fastest):
                                      eight lanes per beam with the optimiser state in registers for
                                      num_nodes <= 169, else one warp per beam with the state in shared memory */
#define OPS_SOLVER_BAND_LDLT    1  /* in-place banded (block) LDL^T of K, factor kept in shared memory */
#define OPS_SOLVER_THREE_MOMENT_THREAD 2  /* three-moment, one thread per beam (any num_nodes) */
#define OPS_SOLVER_THREE_MOMENT_SMEM8  3  /* three-moment, 8 lanes per beam, optimiser state in shared memory (num_nodes <= 169) */
#define OPS_SOLVER_THREE_MOMENT_SMEM32 4  /* three-moment, one warp per beam, optimiser state in shared memory
                                              (what OPS_SOLVER_THREE_MOMENT runs beyond 169 nodes) */

/* Module-level constants of the reference generators (SingleCore:20-49, MultiCore:20-52, GPU:21-56,
 * BeamOpt:24-48) plus the literals of the loss (SingleCore:195-196) and torch's Adam defaults. */
typedef struct OpsBeamOptParams {
    int32_t struct_size;     /* sizeof(OpsBeamOptParams) -- ABI guard */
    int32_t num_nodes;       /* num_nodes (101); elements n = num_nodes - 1 */
    int32_t num_cases;       /* load cases sharing one I vector; the reference has 1 */
    int32_t max_forces;      /* row width of force_nodes / force_vals (M_forces_max = 4; BeamOpt uses 5) */
    int32_t max_epochs;      /* max_e (600; BeamOpt 1000) */
    int32_t patience;        /* SC 5, MC 10 (def default shadows the constant), GPU 100, BeamOpt 10 */
    int32_t early_stop;      /* 1 = reference behaviour; 0 = run exactly max_epochs (benchmark mode) */
    int32_t zero_last_node;  /* 1 = MultiCore:222-223 emits 0.0 for the last node's uy / theta */
    int32_t solver;          /* OPS_SOLVER_*: how K(I) u = f is solved each epoch (results agree to ~1e-10) */
    int32_t reserved;        /* must be 0 */
    double E;                /* 200e9 */
    double G;                /* E / (2 (1 + nu)) */
    double udl;              /* uniform_udl (-1000; BeamOpt -5000), applied to every element */
    double I0;               /* I_0 = 0.5 */
    double lr;               /* 0.01 */
    double gamma;            /* 0.98 (ExponentialLR) */
    double alpha_moment;     /* 1e-2 */
    double alpha_shear;      /* 1e-2 */
    double tolerance;        /* 5e-3 (GPU / BeamOpt scripts: 1e-2) */
    double shear_k;          /* 0.03  : A_approx = 0.03 * I**0.5      (SingleCore:196) */
    double bending_eps;      /* 1e-6  : 2 E I + 1e-6                  (SingleCore:195) */
    double clamp_min;        /* 1e-8  : I_tensor.clamp_(min=1e-8)     (SingleCore:208) */
    double beta1, beta2, adam_eps;   /* torch.optim.Adam defaults 0.9, 0.999, 1e-8 (SingleCore:166) */
} OpsBeamOptParams;

/* Library identification: "openpystruct_b200 <semver> sm_100a". */
const char *ops_beamopt_version(void);

/* Number of CUDA devices visible to the library (0 when there is none; compute entries then fail). */
int ops_device_count(void);

/* Makes `device` current for the calling thread (cudaSetDevice); the *_launch entries run on the
 * current device, which must be the one that owns the buffers and the stream. */
int ops_set_device(int device);

/*
 * Per-epoch Adam scalars.  torch computes bias corrections, the decayed learning rate
 * (ExponentialLR, lr_t = lr_{t-1} * gamma) and step_size in Python doubles and rounds them to fp32
 * when they meet the fp32 parameter tensor (torch/optim/adam.py:_single_tensor_adam); the kernel
 * must see bit-identical values, so they are produced on the host in double.
 * host_table receives 2*max_epochs floats: [-(lr_t/bc1_t), sqrt(bc2_t)] for t = 1..max_epochs.
 * Replaces: optimizer/scheduler construction SingleCore:166-167 and scheduler.step() :204.
 */
int ops_beamopt_fill_schedule(const OpsBeamOptParams *p, float *host_table);

/* Bytes of device scratch ops_beamopt_launch needs for B beams on the current device. */
size_t ops_beamopt_workspace_bytes(const OpsBeamOptParams *p, int64_t B);

/*
 * The fused optimisation loop for B independent beams (device pointers).
 *
 * Inputs
 *   fixed_uy     u8 [B][num_nodes]                1 = uy constrained (pin at node 0 is implied and
 *                                                 forced; rollers: ops.fix(node,0,1,0) SingleCore:101-102)
 *   force_nodes  i32[B][num_cases][max_forces]    0-based node index of each point load, <0 = unused
 *   force_vals   f64[B][num_cases][max_forces]    ops.load(node, 0.0, F, 0.0) SingleCore:112-113
 *   L            f64[B]                           beam length; node_positions = linspace(0, L, num_nodes)
 *   d_schedule   f32[2*max_epochs]                ops_beamopt_fill_schedule() copied to the device
 * Outputs (the tensors behind generate_sample's record, SingleCore:235-249)
 *   I_values     f32[B][n]                        optimised inertias AFTER the last Adam step + clamp
 *   deflections  f64[B][num_cases][num_nodes]     nodeDisp(i,2) of the LAST ANALYSED model
 *   rotations    f64[B][num_cases][num_nodes]     nodeDisp(i,3)        "
 *   shear        f32[B][num_cases][n]             eleResponse(e,'forces')[1] cast to fp32   "
 *   moment       f32[B][num_cases][n]             eleResponse(e,'forces')[2] cast to fp32   "
 *   epochs       i32[B]                           iterations executed (early stop SingleCore:211-219)
 *   loss         f32[B]                           total_loss of the last iteration
 *   status       i32[B]                           0 ok, 1 = factorisation/solve failed (sample to be dropped)
 */
int ops_beamopt_launch(const OpsBeamOptParams *p, int64_t B,
                       const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                       const double *L, const float *d_schedule,
                       float *I_values, double *deflections, double *rotations, float *shear,
                       float *moment, int32_t *epochs, float *loss, int32_t *status,
                       void *d_workspace, size_t workspace_bytes, void *cuda_stream);

/*
 * The same launch with the DATASET GATHER FUSED INTO THE KERNEL'S RECORD WRITE (beams shard over the GPUs
 * of a box with no per-iteration communication; the only exchange is the final gather of the dataset, which the
 * reference does by appending worker results in the parent process, MultiCore:258-270).  `dests` holds n_dest
 * (1..8) sets of the eight record arrays, each sized for the WHOLE dataset: this GPU's own FIRST, then its peers',
 * mapped into this process with ops_peer_open.  Beam b of this launch is written to row row0 + b of set 0 and
 * copied by the same thread group to that row of every other set (stores over NVLink) while the other beams keep
 * iterating, so when all ranks' launches have completed every GPU holds the complete dataset; the caller orders that with a barrier on the stream (e.g. a one-element NCCL all_reduce).
 * With a FIXED epoch count (early_stop = 0) and more than one peer the call runs as a pipeline instead -- the plain
 * kernel in chunks of whole rounds, and after every chunk a copy kernel on a library-owned side stream that stores the
 * chunk's rows into the peers' sets while the next chunk iterates (all beams of a round finish together there, so the
 * in-kernel copy would be an exposed burst); the side stream is joined back into `cuda_stream` before the call
 * returns, so the stream-ordering contract is the same.
 * Production (lanes) kernel only: OPS_E_UNSUPP for configurations that run another kernel.
 */
typedef struct OpsBeamOptRecordArrays {
    float *I_values;
    double *deflections, *rotations;
    float *shear, *moment;
    int32_t *epochs;
    float *loss;
    int32_t *status;
} OpsBeamOptRecordArrays;

int ops_beamopt_launch_scatter(const OpsBeamOptParams *p, int64_t B,
                               const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                               const double *L, const float *d_schedule,
                               int n_dest, const OpsBeamOptRecordArrays *dests, int64_t row0,
                               void *d_workspace, size_t workspace_bytes, void *cuda_stream);

/* 1 when ops_beamopt_launch_scatter serves this parameter block (production lanes kernel: default solver, num_nodes <= 169,
 * multi-case only at <= 105 nodes), else 0.  Depends on `p` alone: every rank of a job decides alike before any collective. */
int ops_beamopt_scatter_supported(const OpsBeamOptParams *p);

/*
 * Peer-visible device buffers for the scatter above (CUDA IPC; one process per GPU).  ops_peer_alloc: cudaMalloc on
 * the current device + a 64-byte handle to send to the other ranks; ops_peer_open: map a peer's buffer into this
 * process (peer access is enabled lazily); ops_peer_close / ops_peer_free undo them.  0 / OPS_E_* / cudaError_t.
 */
int ops_peer_alloc(size_t bytes, void **dptr, unsigned char *handle64);
int ops_peer_open(const unsigned char *handle64, void **dptr);
int ops_peer_close(void *dptr);
int ops_peer_free(void *dptr);

/*
 * One static solve per beam for given inertias, everything in FP64 (no optimiser): the
 * setup_model + analyze + eleResponse + nodeDisp sequence (SingleCore:176-190, 224-232) in isolation.
 * I_f64[B][n]; single load case (force arrays [B][max_forces]); outputs f64 [B][num_nodes] / [B][n].
 */
int ops_beamsolve_launch(const OpsBeamOptParams *p, int64_t B,
                         const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                         const double *L, const double *I_f64,
                         double *deflections, double *rotations, double *shear, double *moment,
                         int32_t *status, void *cuda_stream);

/*
 * Host-buffer convenience around ops_beamopt_launch for callers that are not torch (ctypes / cffi
 * from the reference's scripts): allocates device buffers, copies inputs H2D, runs, copies every
 * output D2H, synchronises, frees.  Same arrays as above but HOST pointers; `device` is the CUDA
 * ordinal.  elapsed_ms (optional) receives the device time of the launch alone.
 */
int ops_beamopt_run_host(const OpsBeamOptParams *p, int64_t B,
                         const uint8_t *fixed_uy, const int32_t *force_nodes, const double *force_vals,
                         const double *L,
                         float *I_values, double *deflections, double *rotations, float *shear,
                         float *moment, int32_t *epochs, float *loss, int32_t *status,
                         int device, float *elapsed_ms);

/*
 * Host-buffer SESSION: the same work as ops_beamopt_run_host for callers that generate many batches
 * (the reference's main() loop over joblib batches, MultiCore:246-262).  The session owns the device
 * buffers, a stream and PINNED host buffers for one device and up to max_beams beams per run, so a
 * run costs one H2D copy of the inputs, one launch and one D2H copy of the record arrays -- no
 * allocation, no page faults on fresh host pages.
 *
 *   ops_beamopt_session_create   allocates everything (returns 0 / OPS_E_* / cudaError_t)
 *   ops_beamopt_session_arrays   the pinned host arrays (layouts of ops_beamopt_launch, max_beams rows):
 *                                the caller fills the four input arrays in place before a run and
 *                                reads the eight output arrays after it (valid until the next run)
 *   ops_beamopt_session_run      H2D, launch, D2H, synchronise for the first B beams; elapsed_ms
 *                                (optional) = device time of the launch alone
 *   ops_beamopt_session_destroy  frees everything
 */
typedef struct OpsBeamOptSession OpsBeamOptSession;

typedef struct OpsBeamOptHostArrays {
    uint8_t *fixed_uy;
    int32_t *force_nodes;
    double *force_vals;
    double *L;
    float *I_values;
    double *deflections, *rotations;
    float *shear, *moment;
    int32_t *epochs;
    float *loss;
    int32_t *status;
} OpsBeamOptHostArrays;

int ops_beamopt_session_create(const OpsBeamOptParams *p, int64_t max_beams, int device, OpsBeamOptSession **out);
int ops_beamopt_session_arrays(OpsBeamOptSession *s, OpsBeamOptHostArrays *arrays);
int ops_beamopt_session_run(OpsBeamOptSession *s, int64_t B, float *elapsed_ms);
void ops_beamopt_session_destroy(OpsBeamOptSession *s);

/*
 * Diagnostic for the roofline denominator (no reference counterpart): runs `chains` independent
 * DFMA chains of length `iters` per thread on a full grid of the current device and reports the
 * sustained FP64 rate in TFLOP/s (FMA = 2 flop), timed with CUDA events on `cuda_stream`.
 * Synchronises the stream.
 */
int ops_fp64_peak_probe(int iters, double *tflops, float *elapsed_ms, void *cuda_stream);

/*
 * Diagnostic (no reference counterpart -- the reference has no launch to plan): the kernel family and launch
 * geometry ops_beamopt_launch would use for a batch of B beams.  Host arithmetic only: with sms > 0 and
 * smem_optin > 0 (a B200: 148 SMs, 232 448 bytes of opt-in shared memory per CTA) no device is touched, with
 * 0 / 0 the current device is queried.  Returns 0, OPS_E_BADARG or OPS_E_UNSUPP like the launch would.
 */
enum { OPS_PLAN_LANES = 0,               /* eight lanes per beam (x load cases), state in registers */
       OPS_PLAN_LANES_TM = 1,            /* the same with the lane-private data in tensor memory */
       OPS_PLAN_WIDE = 2,                /* lanes_per_beam = 8 or 32, state in shared memory (fine discretisations) */
       OPS_PLAN_THREAD_THREE_MOMENT = 3, /* thread per beam, three-moment solve */
       OPS_PLAN_THREAD_LDLT = 4 };       /* thread per beam, banded LDL^T */
typedef struct {
    int32_t family;
    int32_t threads, blocks;             /* per CTA; CTAs (persistent: at most one per SM for families 0..2) */
    int32_t lanes_per_beam;              /* threads working on one beam (all its load cases) */
    int32_t beams_per_cta;               /* beams resident per CTA and round */
    int32_t scatter;                     /* 1: ops_beamopt_launch_scatter serves this configuration */
    int64_t smem_bytes;                  /* dynamic shared memory per CTA */
    int64_t workspace_bytes;             /* = ops_beamopt_workspace_bytes */
} OpsLaunchPlanInfo;
int ops_beamopt_plan(const OpsBeamOptParams *p, int64_t B, int32_t sms, int32_t smem_optin, OpsLaunchPlanInfo *out);

/*
 * Diagnostic (no reference counterpart): sustained issue rate of one instruction class, in warp
 * instructions per clock per SM at the device's maximum clock (8 independent chains per thread,
 * 1024 threads per SM).  op: 0 DFMA, 1 FFMA, 2 FMUL, 3 FADD, 4 MUFU.RCP, 5 F2F (f32<->f64), 6 IMAD,
 * 7 LOP3, 8 FFMA with immediate operands.  Used by profiles/ to put the kernel's instruction mix
 * against the pipes that execute it.
 */
int ops_pipe_probe(int op, int iters, double *warp_inst_per_clk_per_sm, void *cuda_stream);

/*
 * Diagnostic (no reference counterpart): checks the branch-free fp32 division / square-root
 * sequences of the production kernel (csrc/fastmath.cuh) against the compiler's IEEE `/` and sqrtf
 * on `samples` random operands inside the ranges the kernel guarantees.  mismatches3 receives the
 * number of results that are not bit-identical {a/b, sqrt(x), 1/b and the packed fp32 pair operations (FFMA2 / FMUL2 /
 * FADD2 against the scalar instructions, incl. the product-then-sum that must stay un-fused)}; rcp64_max_rel_err the largest
 * |x * rcp64(x) - 1| of the FP64 reciprocal.  Synchronises the stream.
 */
int ops_fastmath_selftest(int64_t samples, int64_t *mismatches3, int64_t *samples_run, double *rcp64_max_rel_err,
                          void *cuda_stream);

/*
 * Host side of the seam: the reference's support / load sampling (SingleCore:133-160; it "stays on the host, unchanged",
 * SURVEY 8a row 3) drawn natively from a bit-exact replica of CPython's `random` -- the same stream random.seed(seed)
 * gives the reference's own statements -- straight into the arrays of ops_beamopt_launch.  No device involved.
 *   ops_sampler_create      random.Random(seed), 0 <= seed < 2^64
 *   ops_sampler_random / ops_sampler_randint   one random.random() / random.randint(a, b) of the stream (tests, frames)
 *   ops_sampler_draw_cases  `count` consecutive generate_sample() draws (count = beams * num_cases; consecutive cases
 *                           share a beam and the supports / length of its first case).  flag = 0: the fixed
 *                           roller_nodes / available_nodes lists (1-based tags, SingleCore:58-66); flag = 1: per sample
 *                           L = L_min + uniform(0, L_max) and 1..N_rollers_max rollers by choice + remove (:133-151).
 *                           Then always k = randint(1, M_forces_max), sample(available, k), k x uniform(min_force, max_force).
 *     fixed_uy u8[beams][num_nodes], force_nodes i32[count][max_forces] (0-based, -1 unused), force_vals f64[count][max_forces],
 *     L_out f64[beams]: the ABI arrays;  roller_tags i32[count][roller_width], force_tags i32[count][max_forces] (1-based tags,
 *     0 unused) and case_L f64[count]: what the record needs besides (roller_nodes, force_nodes, L of the 13-key dict).
 */
typedef struct OpsSampler OpsSampler;
int ops_sampler_create(uint64_t seed, OpsSampler **out);
void ops_sampler_destroy(OpsSampler *s);
double ops_sampler_random(OpsSampler *s);
int ops_sampler_randint(OpsSampler *s, int a, int b);
int ops_sampler_draw_cases(OpsSampler *s, int64_t count, int32_t num_nodes, int32_t flag, double L,
                           const int32_t *roller_nodes, int32_t n_rollers, const int32_t *available_nodes,
                           int32_t n_available, double L_max, double L_min, int32_t N_rollers_max, int32_t M_forces_max,
                           double max_force, double min_force, int32_t num_cases, int32_t max_forces,
                           uint8_t *fixed_uy, int32_t *force_nodes, double *force_vals, double *L_out,
                           int32_t *roller_tags, int32_t roller_width, int32_t *force_tags, double *case_L);

/*
 * ---------------------------------------------------------------------------------------------------------------------
 * Frame optimiser: the optimisation loop of OpenPyStruct_FrameOpt_Discrete_Beta.py:179-206 for a
 * batch of rectangular frames, one CTA per frame.  Per epoch it replaces setup_frame_model (:75-139) + ops.analyze(1)
 * (:183) + compute_combined_loss (:141-160) + backward / Adam / clamp (:184-189) + the early-stop test (:194-205).
 * Constants: the script's module-level block (:14-44) plus the literals of the loss and torch's Adam defaults.
 */
typedef struct OpsFrameOptParams {
    int32_t struct_size;     /* sizeof(OpsFrameOptParams) -- ABI guard */
    int32_t max_bays;        /* max_bays (10): upper bound of num_bays[] in a launch, sizes the shared memory (<= 16) */
    int32_t max_stories;     /* max_stories (10) */
    int32_t max_epochs;      /* num_epochs (5000) */
    int32_t patience;        /* 10 */
    int32_t early_stop;      /* 1 = reference behaviour; 0 = run exactly max_epochs */
    double E;                /* 200e9 */
    double G;                /* E / (2 (1 + nu)) */
    double A;                /* 0.02: cross-section area of every member */
    double I0;               /* 5e-4 */
    double alpha_moment;     /* 1e-2 */
    double alpha_shear;      /* 1e-2 */
    double shear_k;          /* 0.03: A_local = k * sqrt(I) (:156) */
    double bending_eps;      /* 1e-8: 2 E I + 1e-8 (:155) */
    double lateral_load;     /* 1e4: nodal load in x on the left-hand nodes above ground (:129-131) */
    double vertical_load;    /* -1e4: eleLoad -beamUniform w w on every beam (:135-138) */
    double lr;               /* 0.005, constant (no scheduler) */
    double tolerance;        /* 1e-3 */
    double bay_width;        /* 6.0 */
    double story_height;     /* 3.0 */
    double clamp_min;        /* 1e-8 (:188-189) */
    double beta1, beta2, adam_eps;   /* torch.optim.Adam defaults */
} OpsFrameOptParams;

/* columns + beams of the largest frame a launch admits = row width of I_values / moment / shear */
int ops_frameopt_max_elements(const OpsFrameOptParams *p);

/* host_table receives 2*max_epochs floats [-(lr/bc1_t), sqrt(bc2_t)], t = 1..max_epochs (see ops_beamopt_fill_schedule) */
int ops_frameopt_fill_schedule(const OpsFrameOptParams *p, float *host_table);

/*
 * Device pointers, stream-ordered, caller-owned buffers (conventions of ops_beamopt_launch).  Members are numbered like
 * the reference: columns story by story (:104-111), then the beams of every elevated story (:114-121).
 *   num_bays, num_stories  i32[B]                  the frame (the script draws both with random.randint, :50-51)
 *   d_schedule             f32[2*max_epochs]
 *   I_values               f32[B][max_elements]    opt_I: inertias after the last Adam step + clamp (unused tail 0)
 *   loss_history           f32[B][max_epochs]      total_loss.item() per epoch (:191-192), NaN beyond `epochs`
 *   moment, shear          f64[B][max_elements]    eleResponse(e,'forces')[2|1] of the LAST analysed inertias
 *   best_loss              f64[B]                  best_loss of the stop test (:194-196)
 *   epochs                 i32[B]                  iterations executed
 *   status                 i32[B]                  0 ok, 1 non-SPD pivot / non-finite loss, 2 frame outside max_bays / max_stories
 */
int ops_frameopt_launch(const OpsFrameOptParams *p, int64_t B, const int32_t *num_bays, const int32_t *num_stories,
                        const float *d_schedule, float *I_values, float *loss_history, double *moment, double *shear,
                        double *best_loss, int32_t *epochs, int32_t *status, void *cuda_stream);

/* host-buffer convenience around ops_frameopt_launch (allocates, copies, runs, copies back, frees) */
int ops_frameopt_run_host(const OpsFrameOptParams *p, int64_t B, const int32_t *num_bays, const int32_t *num_stories,
                          float *I_values, float *loss_history, double *moment, double *shear, double *best_loss,
                          int32_t *epochs, int32_t *status, int device, float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* OPENPYSTRUCT_B200_H */
