// sm_100a kernel of the frame optimiser: one CTA per frame, the whole optimisation loop fused.
//
// Replaces, per epoch (reference file:line, OpenPyStruct_FrameOpt_Discrete_Beta.py): setup_frame_model (:75-139: nodes,
// clamped ground nodes, elasticBeamColumn columns then beams, lateral nodal loads, beamUniform on the beams, BandGeneral),
// ops.analyze(1) (:183), compute_combined_loss (:141-160: eleResponse(e,'forces')[1|2] per element, fp32 loss accumulated
// element by element), total_loss.backward() (:184), Adam without learning-rate decay (:174, 185), clamp_(1e-8)
// (:188-189) and the early-stop test (:194-205).
//
// FE half (FP64).  Degrees of freedom only at the elevated nodes, (ux, uy, rz) per node in tag order: K(I) is SPD and
// banded with half bandwidth 3 (bays + 1) + 2.  Its lower band lives in shared memory, assembled by GATHER (one thread
// per node sums the <= 4 members meeting there in a fixed order: deterministic, no atomics) from closed-form global
// member matrices of the two orientations (no rotation at run time), factored by a right-looking band Cholesky with
// the forward substitution riding on the same column sweep (two CTA barriers per column), back substitution by one
// warp (shuffle-reduced dot products in a fixed order).  Global end forces [Fy_i, Mz_i] per member incl. the fixed-end
// terms of `eleLoad -beamUniform Wy Wx` (the reference passes the SAME value as transverse and axial load, :138).
//
// fp32 half (torch's CPU operation order, every line one rounded operation; checked bit for bit against autograd on
// random inputs, tests/test_frames.py).  Per member, with M, V the Python doubles eleResponse returned:
//     b = fp32(2E) I + 1e-8f ;  rb = 1 / b ;  d = rb * fp32(M M)          (python_float / tensor is reciprocal() * float)
//     s = sqrt(I) ;  gg = fp32(G) (fp32(k) s) ;  rg = 1 / gg ;  q = rg * fp32(V V)
//     be = (((0 + d_0) + d_1) + ...) ,  se likewise                      (Python `+=` in member order)
//     total = (SUM(I) + a_m be) + a_s se                                   (SUM = torch.sum's cascade)
//     gb = ((-(a_m fp32(M M))) (rb rb)) fp32(2E)                           (mul, reciprocal, add, mul backward)
//     gs = ((((-(a_s fp32(V V))) (rg rg)) fp32(G)) fp32(k)) (0.5f (1 / sqrt(I)))
//     g  = 1 + (gs + gb)                                                   (both uses of I[e] meet at the select node)
// followed by torch's single-tensor Adam (lerp as FMA, addcmul as FMA, addcdiv un-fused) with per-epoch scalars from a
// host table, and clamp.  Divisions and square roots are the compiler's IEEE operators (-fmad=false build).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/openpystruct_b200.h"
#include "beamopt_core.cuh"

namespace ops {
namespace frame {

constexpr int THREADS = 128;
constexpr int WSLOTS = 8;                       // window positions of the trailing update a thread keeps in registers
constexpr int MAX_DIM = 16;                     // bays, stories <= 16 (the reference draws 1..10)

struct Consts {
    int max_epochs, patience, early_stop, max_bays, max_stories;
    double E, A, tol, lateral, vertical, bay_width, story_height;
    float I0f, E2, Gf, kf, am, as_, epsf, clampf, w1, b2f, omb2f, adam_epsf;
};

struct Ptrs {
    const int32_t *bays, *stories;
    const float *sched;
    float *I_values;            // [B][max_elems]
    float *loss_hist;           // [B][max_epochs]
    double *moment, *shear;     // [B][max_elems]: of the last analysed inertias
    double *best;               // [B]
    int32_t *epochs, *status;   // [B]
    int max_elems;
};

__host__ __device__ inline int num_elems(int bays, int stories) { return stories * (bays + 1) + stories * bays; }
__host__ __device__ inline int num_dofs(int bays, int stories) { return 3 * stories * (bays + 1); }
__host__ __device__ inline int half_bw(int bays) { return 3 * (bays + 1) + 2; }

// shared memory of one CTA, sized on the host for the largest frame the launch admits
__host__ __device__ inline size_t smem_bytes(int max_bays, int max_stories)
{
    const size_t n = (size_t)num_dofs(max_bays, max_stories), hb = (size_t)half_bw(max_bays), ne = (size_t)num_elems(max_bays, max_stories);
    size_t doubles = (hb + 1) * n      // lower band of K, then of its Cholesky factor
                     + 3 * n           // load vector, right-hand side / solution, reciprocal pivots
                     + 2 * (hb + 1)    // scaled column of the current elimination step (double-buffered)
                     + 2 * ne;         // M, V
    size_t floats = 7 * ne;            // I, m, v, d, q, g, (spare)
    size_t bytes = doubles * 8 + floats * 4 + (hb + 1) * (hb + 2);   // + (r, c) pair table of the trailing update
    return (bytes + 15) / 16 * 16;
}

struct Member {
    double a, b, c, d, t;       // EA/L, 12EI/L^3, 6EI/L^2, 4EI/L, 2EI/L
};
__device__ __forceinline__ Member member(const Consts &k, double I, double L)
{
    Member m;
    const double EI = k.E * I;
    m.a = k.E * k.A / L;
    m.b = 12.0 * EI / (L * L * L);
    m.c = 6.0 * EI / (L * L);
    m.d = 4.0 * EI / L;
    m.t = 2.0 * EI / L;
    return m;
}

__global__ void __launch_bounds__(THREADS) frameopt_kernel(const Consts k, const long long B, const Ptrs p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    for (long long fr = blockIdx.x; fr < B; fr += gridDim.x) {
        const int bays = p.bays[fr], stories = p.stories[fr];
        const bool valid = bays >= 1 && stories >= 1 && bays <= k.max_bays && stories <= k.max_stories;
        const int nb1 = bays + 1, n = valid ? num_dofs(bays, stories) : 0, hb = valid ? half_bw(bays) : 0;
        const int nnodes = valid ? stories * nb1 : 0;                // elevated nodes
        const int n_col = stories * nb1, ne = valid ? num_elems(bays, stories) : 0;
        double *ab = reinterpret_cast<double *>(smem_raw);       // ab[d * n + j] = A[j + d][j]
        double *f0 = ab + (size_t)(hb + 1) * n;
        double *u = f0 + n;
        double *rinv = u + n;
        double *lcol = rinv + n;
        double *lcol2 = lcol + (hb + 1);
        double *Mv = lcol2 + (hb + 1);
        double *Vv = Mv + ne;
        float *If = reinterpret_cast<float *>(Vv + ne);
        float *mf = If + ne, *vf = mf + ne, *df = vf + ne, *qf = df + ne, *gf = qf + ne;
        unsigned char *pairs = reinterpret_cast<unsigned char *>(gf + ne);
        __shared__ int s_flag[4];                                // bad pivot, done, epochs, counter
        __shared__ double s_best;

        // ---- once per frame: load vector, pair table, state
        const int npairs = hb * (hb + 1) / 2;
        for (int i = tid; i < n; i += THREADS) f0[i] = 0.0;
        for (int i = tid; i < ne; i += THREADS) { If[i] = k.I0f; mf[i] = 0.0f; vf[i] = 0.0f; Mv[i] = 0.0; Vv[i] = 0.0; }
        for (int i = tid; i < npairs; i += THREADS) {            // idx -> (r, c), 1 <= c <= r <= hb
            int r = 1, left = i;
            while (left >= r) { left -= r; ++r; }
            pairs[2 * i] = (unsigned char)r; pairs[2 * i + 1] = (unsigned char)(left + 1);
        }
        if (tid == 0) { s_flag[0] = valid ? 0 : 1; s_flag[1] = 0; s_flag[2] = 0; s_flag[3] = 0; s_best = INFINITY; }
        __syncthreads();
        // the trailing update's fixed window positions of this thread (warps 1..3): slot s_ = entry s_ * 96 + (tid - 32) of
        // the pair table shifted by one, i.e. (r, c) with 2 <= c <= r; r = 0 marks an unused slot
        const int warp = tid >> 5, lane = tid & 31;
        unsigned int wr[WSLOTS / 4], wc[WSLOTS / 4];
        int woff[WSLOTS];
#pragma unroll
        for (int w_ = 0; w_ < WSLOTS / 4; ++w_) { wr[w_] = 0u; wc[w_] = 0u; }
#pragma unroll
        for (int s_ = 0; s_ < WSLOTS; ++s_) {
            const int i = s_ * (THREADS - 32) + (tid - 32);
            int r = 0, c = 0;
            if (warp > 0 && i < (hb - 1) * hb / 2) { r = pairs[2 * i] + 1; c = pairs[2 * i + 1] + 1; }
            wr[s_ >> 2] |= (unsigned int)r << (8 * (s_ & 3));
            wc[s_ >> 2] |= (unsigned int)c << (8 * (s_ & 3));
            woff[s_] = (r - c) * n + c;
        }
        {
            const double w = k.vertical, L = k.bay_width;
            for (int nd = tid; nd < nnodes; nd += THREADS) {      // elevated node (s, b), s = 1 .. stories
                const int b = nd % nb1;
                double fx = (b == 0) ? k.lateral : 0.0, fy = 0.0, mz = 0.0;
                if (b < bays) { fx += w * L / 2; fy += w * L / 2; mz += w * L * L / 12; }     // node i of the beam to the right
                if (b > 0) { fx += w * L / 2; fy += w * L / 2; mz -= w * L * L / 12; }        // node j of the beam to the left
                f0[3 * nd] = fx; f0[3 * nd + 1] = fy; f0[3 * nd + 2] = mz;
            }
        }
        __syncthreads();

        int t = 0;
        while (valid && k.max_epochs > 0) {
            // ---- assembly by gather: thread = elevated node
            for (int i = tid; i < (hb + 1) * n; i += THREADS) ab[i] = 0.0;
            for (int i = tid; i < n; i += THREADS) u[i] = f0[i];
            __syncthreads();
            for (int nd = tid; nd < nnodes; nd += THREADS) {
                const int s = nd / nb1 + 1, b = nd % nb1, j0 = 3 * nd;
                // columns: below (story s-1 -> s), above (s -> s+1); beams: left (b-1 -> b), right (b -> b+1)
                const Member cb = member(k, (double)If[(s - 1) * nb1 + b], k.story_height);
                double d00 = cb.b, d11 = cb.a, d22 = cb.d, d20 = cb.c, d21 = 0.0;          // K_jj of a column
                if (s < stories) {
                    const Member ca = member(k, (double)If[s * nb1 + b], k.story_height);
                    d00 += ca.b; d11 += ca.a; d22 += ca.d; d20 += -ca.c;                    // K_ii of a column
                    // coupling to node (s+1, b): block (q, p) = K_ij^T, K_ij = [[-b,0,-c],[0,-a,0],[c,0,t]]
                    const int dq = 3 * nb1;
                    ab[(size_t)(dq + 0 - 0) * n + j0 + 0] = -ca.b;     // (r0, c0) = K_ij[0][0]
                    ab[(size_t)(dq + 2 - 0) * n + j0 + 0] = -ca.c;     // (r2, c0) = K_ij[0][2]
                    ab[(size_t)(dq + 1 - 1) * n + j0 + 1] = -ca.a;     // (r1, c1) = K_ij[1][1]
                    ab[(size_t)(dq + 0 - 2) * n + j0 + 2] = ca.c;      // (r0, c2) = K_ij[2][0]
                    ab[(size_t)(dq + 2 - 2) * n + j0 + 2] = ca.t;      // (r2, c2) = K_ij[2][2]
                }
                if (b > 0) {
                    const Member bl = member(k, (double)If[n_col + (s - 1) * bays + b - 1], k.bay_width);
                    d00 += bl.a; d11 += bl.b; d22 += bl.d; d21 += -bl.c;                    // K_jj of a beam
                }
                if (b < bays) {
                    const Member br = member(k, (double)If[n_col + (s - 1) * bays + b], k.bay_width);
                    d00 += br.a; d11 += br.b; d22 += br.d; d21 += br.c;                     // K_ii of a beam
                    // coupling to node (s, b+1): K_ij = [[-a,0,0],[0,-b,c],[0,-c,t]], block (q, p) = K_ij^T
                    ab[(size_t)(3 + 0 - 0) * n + j0 + 0] = -br.a;      // (r0, c0)
                    ab[(size_t)(3 + 1 - 1) * n + j0 + 1] = -br.b;      // (r1, c1) = K_ij[1][1]
                    ab[(size_t)(3 + 2 - 1) * n + j0 + 1] = br.c;       // (r2, c1) = K_ij[1][2]
                    ab[(size_t)(3 + 1 - 2) * n + j0 + 2] = -br.c;      // (r1, c2) = K_ij[2][1]
                    ab[(size_t)(3 + 2 - 2) * n + j0 + 2] = br.t;       // (r2, c2) = K_ij[2][2]
                }
                ab[j0] = d00; ab[j0 + 1] = d11; ab[j0 + 2] = d22;                            // diagonal (d = 0)
                ab[(size_t)2 * n + j0] = d20;                                                // A[j0+2][j0]
                ab[(size_t)1 * n + j0 + 1] = d21;                                            // A[j0+2][j0+1]
            }
            __syncthreads();
            // ---- band Cholesky + forward substitution, column by column, ONE CTA barrier per column.
            // Warp 0 owns the critical path: it scales column j (pivot, reciprocal square root, y_j), and in the trailing
            // update it takes exactly the entries it needs next -- column j + 1 and the right-hand side -- so that it can
            // scale column j + 1 without waiting for anybody; warps 1..3 apply column j to the rest of the window
            // (columns j + 2 ...) meanwhile, each thread on a FIXED set of window positions (r, c) it worked out once per
            // frame (registers; eight per thread cover half bandwidths up to 38, the pair table serves the rest), so
            // that its loads are issued back to back instead of one dependent chain per entry.  The scaled column
            // travels through a double-buffered `lcol`; the barrier of step j + 1 is what orders the other warps'
            // updates of step j before anything of step j + 1 touches them.  Every entry receives the same rank-1
            // updates in the same (column) order as a plain right-looking sweep: same bits.
            {
                for (int j = 0; j < n; ++j) {
                    const int lim = (n - 1 - j < hb) ? n - 1 - j : hb;
                    double *lc = (j & 1) ? lcol2 : lcol;
                    double yj = 0.0;
                    if (warp == 0) {
                        const double piv = ab[j];
                        const double ri = rsqrt(piv);
                        yj = u[j] * ri;
                        __syncwarp();                                    // (every lane has read u[j])
                        if (lane == 0) {
                            if (!(piv > 0.0)) s_flag[0] = 1;
                            rinv[j] = ri;
                            u[j] = yj;
                        }
                        for (int t = lane + 1; t <= lim; t += 32) {
                            const double l = ab[(size_t)t * n + j] * ri;
                            ab[(size_t)t * n + j] = l;
                            lc[t] = l;
                        }
                    }
                    __syncthreads();
                    if (warp == 0) {
                        // column j + 1 of the window: A[j + r][j + 1] -= l_r l_1 ; right-hand side: u[j + t] -= l_t y_j
                        if (lim >= 1) {
                            const double l1 = lc[1];
                            for (int r = lane + 1; r <= lim; r += 32) {
                                const double lr = lc[r];
                                ab[(size_t)(r - 1) * n + j + 1] = fma(-lr, l1, ab[(size_t)(r - 1) * n + j + 1]);
                                u[j + r] = fma(-lr, yj, u[j + r]);
                            }
                        }
                        __syncwarp();
                    } else {
                        // window positions (r, c), 2 <= c <= r <= lim: this thread's fixed slots first
                        double a_[WSLOTS], lr_[WSLOTS], lc_[WSLOTS];
#pragma unroll
                        for (int s_ = 0; s_ < WSLOTS; ++s_) {
                            const int r = (int)((wr[s_ >> 2] >> (8 * (s_ & 3))) & 0xffu);
                            const bool on = r >= 2 && r <= lim;
                            const int c = (int)((wc[s_ >> 2] >> (8 * (s_ & 3))) & 0xffu);
                            lr_[s_] = on ? lc[r] : 0.0;
                            lc_[s_] = on ? lc[c] : 0.0;
                            a_[s_] = on ? ab[(size_t)woff[s_] + j] : 0.0;
                        }
#pragma unroll
                        for (int s_ = 0; s_ < WSLOTS; ++s_) {
                            const int r = (int)((wr[s_ >> 2] >> (8 * (s_ & 3))) & 0xffu);
                            if (r >= 2 && r <= lim) ab[(size_t)woff[s_] + j] = fma(-lr_[s_], lc_[s_], a_[s_]);
                        }
                        const int np = (lim - 1) * lim / 2;
                        for (int i = tid - 32 + WSLOTS * (THREADS - 32); i < np; i += THREADS - 32) {
                            const int r = pairs[2 * i] + 1, c = pairs[2 * i + 1] + 1;
                            ab[(size_t)(r - c) * n + j + c] = fma(-lc[r], lc[c], ab[(size_t)(r - c) * n + j + c]);
                        }
                    }
                }
                __syncthreads();
            }
            // ---- back substitution L^T x = y by columns (the order of LAPACK's dtbsv): x_j = y_j / l_jj, then
            // y_{j-d} -= l_{j, j-d} x_j for the <= hb entries of row j of the factor -- one warp, no reduction
            if (warp == 0) {
                for (int j = n - 1; j >= 0; --j) {
                    const double xj = u[j] * rinv[j];
                    const int lim = j < hb ? j : hb;
                    __syncwarp();                                        // (every lane has read u[j])
                    if (lane == 0) u[j] = xj;
                    for (int d = lane + 1; d <= lim; d += 32)
                        u[j - d] = fma(-ab[(size_t)d * n + (j - d)], xj, u[j - d]);
                    __syncwarp();
                }
            }
            __syncthreads();
            // ---- member end forces (global, node-i end), loss terms, gradient
            for (int e = tid; e < ne; e += THREADS) {
                const double I = (double)If[e];
                double V, M;
                if (e < n_col) {
                    const int s0 = e / nb1, b = e % nb1;                  // from node (s0, b) up to (s0 + 1, b)
                    const Member m = member(k, I, k.story_height);
                    const int nj = 3 * (s0 * nb1 + b);                    // elevated node index (s0 + 1 - 1) * nb1 + b
                    double uxi = 0.0, uyi = 0.0, rzi = 0.0;
                    if (s0 >= 1) { const int ni = 3 * ((s0 - 1) * nb1 + b); uxi = u[ni]; uyi = u[ni + 1]; rzi = u[ni + 2]; }
                    const double uxj = u[nj], uyj = u[nj + 1], rzj = u[nj + 2];
                    V = m.a * (uyi - uyj);
                    M = fma(m.c, uxj - uxi, fma(m.d, rzi, m.t * rzj));
                } else {
                    const int q = e - n_col, s = q / bays + 1, b = q % bays;
                    const Member m = member(k, I, k.bay_width);
                    const int ni = 3 * ((s - 1) * nb1 + b), nj = ni + 3;
                    const double uyi = u[ni + 1], rzi = u[ni + 2], uyj = u[nj + 1], rzj = u[nj + 2];
                    const double w = k.vertical, L = k.bay_width;
                    V = fma(m.b, uyi - uyj, m.c * (rzi + rzj)) - w * L / 2;
                    M = fma(m.c, uyi - uyj, fma(m.d, rzi, m.t * rzj)) - w * L * L / 12;
                }
                Mv[e] = M; Vv[e] = V;
                const float If_ = If[e];
                const float c = (float)(M * M), h = (float)(V * V);
                const float bb = k.E2 * If_ + k.epsf;
                const float rb = 1.0f / bb;
                const float s_ = sqrtf(If_);
                const float gg = k.Gf * (k.kf * s_);
                const float rg = 1.0f / gg;
                df[e] = rb * c;
                qf[e] = rg * h;
                const float gb = ((-(k.am * c)) * (rb * rb)) * k.E2;
                const float gs = ((((-(k.as_ * h)) * (rg * rg)) * k.Gf) * k.kf) * (0.5f * (1.0f / s_));
                gf[e] = 1.0f + (gs + gb);
            }
            __syncthreads();
            // ---- loss (member order), stop test
            if (tid == 0) {
                float be = 0.0f, se = 0.0f;
                for (int e = 0; e < ne; ++e) { be += df[e]; se += qf[e]; }
                const float sI = torch_sum_f32(ne, [&](int e) { return If[e]; });
                const float total = (sI + k.am * be) + k.as_ * se;
                p.loss_hist[fr * k.max_epochs + t] = total;
                int done = 0;
                if (s_flag[0] || !(total - total == 0.0f)) { s_flag[0] = 1; done = 1; }
                if (k.early_stop) {
                    const double lv = (double)total;
                    if (lv < s_best - k.tol) { s_best = lv; s_flag[3] = 0; } else { ++s_flag[3]; }
                    if (s_flag[3] >= k.patience) done = 1;
                } else if ((double)total < s_best) {
                    s_best = (double)total;
                }
                if (t + 1 >= k.max_epochs) done = 1;
                s_flag[1] = done; s_flag[2] = t + 1;
            }
            // ---- Adam + clamp (the step is taken before the stop test in the reference, :185-205)
            {
                const float neg_step = __ldg(p.sched + 2 * t), bc2_sqrt = __ldg(p.sched + 2 * t + 1);
                for (int e = tid; e < ne; e += THREADS) {
                    const float g = gf[e];
                    const float m = fmaf(k.w1, g - mf[e], mf[e]);
                    const float v = fmaf(k.omb2f * g, g, vf[e] * k.b2f);
                    mf[e] = m; vf[e] = v;
                    const float denom = sqrtf(v) / bc2_sqrt + k.adam_epsf;
                    const float x = If[e] + (neg_step * m) / denom;
                    gf[e] = x < k.clampf ? k.clampf : x;               // (If is still read by thread 0's torch.sum)
                }
            }
            __syncthreads();
            for (int e = tid; e < ne; e += THREADS) If[e] = gf[e];
            ++t;
            const int done = s_flag[1];
            __syncthreads();
            if (done) break;
        }
        // ---- record
        for (int e = tid; e < p.max_elems; e += THREADS) {
            p.I_values[fr * p.max_elems + e] = e < ne ? If[e] : 0.0f;
            p.moment[fr * p.max_elems + e] = e < ne ? Mv[e] : 0.0;
            p.shear[fr * p.max_elems + e] = e < ne ? Vv[e] : 0.0;
        }
        for (int e = t + tid; e < k.max_epochs; e += THREADS) p.loss_hist[fr * k.max_epochs + e] = NAN;
        if (tid == 0) {
            p.epochs[fr] = t;
            p.status[fr] = valid ? s_flag[0] : 2;
            p.best[fr] = s_best;
        }
        __syncthreads();
    }
}

static int make_consts(const OpsFrameOptParams *p, Consts *k)
{
    if (!p || p->struct_size != (int32_t)sizeof(OpsFrameOptParams)) return OPS_E_BADARG;
    if (p->max_epochs < 0 || p->patience < 0 || p->max_bays < 1 || p->max_stories < 1) return OPS_E_BADARG;
    if (p->max_bays > MAX_DIM || p->max_stories > MAX_DIM) return OPS_E_UNSUPP;
    k->max_epochs = p->max_epochs; k->patience = p->patience; k->early_stop = p->early_stop;
    k->max_bays = p->max_bays; k->max_stories = p->max_stories;
    k->E = p->E; k->A = p->A; k->tol = p->tolerance; k->lateral = p->lateral_load; k->vertical = p->vertical_load;
    k->bay_width = p->bay_width; k->story_height = p->story_height;
    k->I0f = (float)p->I0; k->E2 = (float)(2.0 * p->E); k->Gf = (float)p->G; k->kf = (float)p->shear_k;
    k->am = (float)p->alpha_moment; k->as_ = (float)p->alpha_shear; k->epsf = (float)p->bending_eps;
    k->clampf = (float)p->clamp_min; k->w1 = (float)(1.0 - p->beta1); k->b2f = (float)p->beta2;
    k->omb2f = (float)(1.0 - p->beta2); k->adam_epsf = (float)p->adam_eps;
    return 0;
}

}  // namespace frame
}  // namespace ops

using namespace ops;

extern "C" {

int ops_frameopt_max_elements(const OpsFrameOptParams *p)
{
    if (!p || p->max_bays < 1 || p->max_stories < 1) return OPS_E_BADARG;
    return frame::num_elems(p->max_bays, p->max_stories);
}

int ops_frameopt_fill_schedule(const OpsFrameOptParams *p, float *host_table)
{
    if (!p || !host_table || p->struct_size != (int32_t)sizeof(OpsFrameOptParams)) return OPS_E_BADARG;
    // torch.optim.Adam (_single_tensor_adam), constant learning rate (no scheduler in the frame script, :174)
    for (int t = 1; t <= p->max_epochs; ++t) {
        const double bc1 = 1.0 - pow(p->beta1, (double)t);
        const double bc2 = 1.0 - pow(p->beta2, (double)t);
        host_table[2 * (t - 1)] = (float)(-(p->lr / bc1));
        host_table[2 * (t - 1) + 1] = (float)pow(bc2, 0.5);
    }
    return 0;
}

int ops_frameopt_launch(const OpsFrameOptParams *p, int64_t B, const int32_t *num_bays, const int32_t *num_stories,
                        const float *d_schedule, float *I_values, float *loss_history, double *moment, double *shear,
                        double *best_loss, int32_t *epochs, int32_t *status, void *cuda_stream)
{
    frame::Consts k;
    const int rc = frame::make_consts(p, &k);
    if (rc) return rc;
    if (B < 0) return OPS_E_BADARG;
    if (B == 0) return 0;
    if (!num_bays || !num_stories || !d_schedule || !I_values || !loss_history || !moment || !shear || !best_loss ||
        !epochs || !status)
        return OPS_E_BADARG;
    int dev = 0, sms = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t smem = frame::smem_bytes(k.max_bays, k.max_stories);
    if (smem > (size_t)optin) return OPS_E_UNSUPP;
    e = cudaFuncSetAttribute(frame::frameopt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame::frameopt_kernel, frame::THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)sms * per_sm;
    if (blocks > B) blocks = B;
    frame::Ptrs q;
    q.bays = num_bays; q.stories = num_stories; q.sched = d_schedule; q.I_values = I_values; q.loss_hist = loss_history;
    q.moment = moment; q.shear = shear; q.best = best_loss; q.epochs = epochs; q.status = status;
    q.max_elems = frame::num_elems(k.max_bays, k.max_stories);
    frame::frameopt_kernel<<<(int)blocks, frame::THREADS, smem, (cudaStream_t)cuda_stream>>>(k, (long long)B, q);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

#define OPS_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)

int ops_frameopt_run_host(const OpsFrameOptParams *p, int64_t B, const int32_t *num_bays, const int32_t *num_stories,
                          float *I_values, float *loss_history, double *moment, double *shear, double *best_loss,
                          int32_t *epochs, int32_t *status, int device, float *elapsed_ms)
{
    if (!p || p->struct_size != (int32_t)sizeof(OpsFrameOptParams) || B < 0) return OPS_E_BADARG;
    if (B == 0) return 0;
    int rc = 0;
    const int me = frame::num_elems(p->max_bays, p->max_stories);
    const size_t nE = (size_t)B * me, nH = (size_t)B * (p->max_epochs > 0 ? p->max_epochs : 1);
    int32_t *d_b = nullptr, *d_s = nullptr, *d_ep = nullptr, *d_st = nullptr;
    float *d_sched = nullptr, *d_I = nullptr, *d_h = nullptr;
    double *d_M = nullptr, *d_V = nullptr, *d_best = nullptr;
    float *h_sched = (float *)malloc(sizeof(float) * 2 * (size_t)(p->max_epochs > 0 ? p->max_epochs : 1));
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!h_sched) return OPS_E_BADARG;
    OPS_CUDA(cudaSetDevice(device));
    rc = ops_frameopt_fill_schedule(p, h_sched);
    if (rc) goto done;
    OPS_CUDA(cudaStreamCreate(&st));
    OPS_CUDA(cudaEventCreate(&e0)); OPS_CUDA(cudaEventCreate(&e1));
    OPS_CUDA(cudaMalloc((void **)&d_b, 4 * (size_t)B)); OPS_CUDA(cudaMalloc((void **)&d_s, 4 * (size_t)B));
    OPS_CUDA(cudaMalloc((void **)&d_ep, 4 * (size_t)B)); OPS_CUDA(cudaMalloc((void **)&d_st, 4 * (size_t)B));
    OPS_CUDA(cudaMalloc((void **)&d_sched, 8 * (size_t)(p->max_epochs > 0 ? p->max_epochs : 1)));
    OPS_CUDA(cudaMalloc((void **)&d_I, 4 * nE)); OPS_CUDA(cudaMalloc((void **)&d_h, 4 * nH));
    OPS_CUDA(cudaMalloc((void **)&d_M, 8 * nE)); OPS_CUDA(cudaMalloc((void **)&d_V, 8 * nE));
    OPS_CUDA(cudaMalloc((void **)&d_best, 8 * (size_t)B));
    OPS_CUDA(cudaMemcpyAsync(d_b, num_bays, 4 * (size_t)B, cudaMemcpyHostToDevice, st));
    OPS_CUDA(cudaMemcpyAsync(d_s, num_stories, 4 * (size_t)B, cudaMemcpyHostToDevice, st));
    OPS_CUDA(cudaMemcpyAsync(d_sched, h_sched, 8 * (size_t)(p->max_epochs > 0 ? p->max_epochs : 1), cudaMemcpyHostToDevice, st));
    OPS_CUDA(cudaEventRecord(e0, st));
    rc = ops_frameopt_launch(p, B, d_b, d_s, d_sched, d_I, d_h, d_M, d_V, d_best, d_ep, d_st, st);
    if (rc) goto done;
    OPS_CUDA(cudaEventRecord(e1, st));
    OPS_CUDA(cudaMemcpyAsync(I_values, d_I, 4 * nE, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(loss_history, d_h, 4 * nH, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(moment, d_M, 8 * nE, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(shear, d_V, 8 * nE, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(best_loss, d_best, 8 * (size_t)B, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(epochs, d_ep, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaMemcpyAsync(status, d_st, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
    OPS_CUDA(cudaStreamSynchronize(st));
    if (elapsed_ms) OPS_CUDA(cudaEventElapsedTime(elapsed_ms, e0, e1));
done:
    free(h_sched);
    cudaFree(d_b); cudaFree(d_s); cudaFree(d_ep); cudaFree(d_st); cudaFree(d_sched); cudaFree(d_I); cudaFree(d_h);
    cudaFree(d_M); cudaFree(d_V); cudaFree(d_best);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (rc > 0) cudaGetLastError();
    return rc;
}

}  // extern "C"
