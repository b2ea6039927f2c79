// Declarations shared by the translation units of libopenpystruct_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "beamopt_core.cuh"
#include "beamopt_lanes.cuh"

namespace ops {

// device pointers of one ops_beamopt_launch call (include/openpystruct_b200.h documents the arrays)
struct OptPtrs {
    const uint8_t *fixed_uy;
    const int32_t *force_nodes;
    const double *force_vals;
    const double *L;
    const float *sched;
    float *I_values;
    double *defl, *rot;
    float *shear, *moment;
    int32_t *epochs;
    float *loss;
    int32_t *status;
    unsigned long long *counter;
    double *ws_d;      // global scratch (thread-per-beam kernels only)
    float *ws_f;
    uint32_t *ws_mask;
    // Record destinations of the lanes kernel: beam b of this launch is row row0 + b of EVERY destination's arrays.
    // One destination (row0 = 0) = the arrays above; several = the same dataset arrays of the peer GPUs, mapped
    // into this process (ops_peer_*): the kernel's record write is the dataset gather (ops_beamopt_launch_scatter).
    lanes::RecordDest dest;
    long long row0;
};

// eight-lanes-per-beam three-moment kernel (beamopt_lanes.cu)
struct LanesPlan {
    int epl;             // element slots per lane (template instance)
    int nfix;            // compile-time element count of the instance (0 = run-time n)
    int num_cases;       // load cases per beam = groups per team
    int threads, blocks;
    size_t smem_bytes;
    int tm;              // tensor-memory instance (beamopt_lanes_tm.cu): {M0, Q0}, m, v of a lane in TMEM, 17 warps per SM
};
bool lanes_supported(const BeamConsts &k, int num_cases);
bool lanes_scatter_supported(const LanesPlan &pl);
int lanes_plan(const BeamConsts &k, int num_cases, int64_t B, int sms, int smem_optin, LanesPlan *pl);
cudaError_t lanes_launch(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream);
bool lanes_tm_supported(int epl, int num_cases);
int lanes_tm_plan(int64_t B, int sms, int smem_optin, LanesPlan *pl);
cudaError_t lanes_tm_launch(const BeamConsts &k, long long B, const OptPtrs &p, const LanesPlan &pl, cudaStream_t stream);

// shared-memory-state three-moment kernels, 8 or 32 lanes per beam (beamopt_wide.cu)
struct WidePlan {
    int lpb;             // lanes per beam
    int beam_bytes;      // shared memory of one beam
    int threads, blocks;
    size_t smem_bytes;
};
bool wide_supported(const BeamConsts &k, int num_cases, int lpb, int smem_optin);
int wide_plan(const BeamConsts &k, int lpb, int64_t B, int sms, int smem_optin, WidePlan *pl);
cudaError_t wide_launch(const BeamConsts &k, long long B, const OptPtrs &p, const WidePlan &pl, cudaStream_t stream);

// fastmath.cuh against the compiler's IEEE operators; out4 = {div mismatches, sqrt mismatches,
// reciprocal mismatches, samples}
cudaError_t fastmath_selftest(long long samples, unsigned long long *host_out4, double *worst_rcp64,
                              cudaStream_t stream);

}  // namespace ops
