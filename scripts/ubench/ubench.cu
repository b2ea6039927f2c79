// Micro-benchmarks of the sm_100a issue model behind the lanes kernel's instruction mix (scripts/ubench/README in
// profiles/): per-SM-sub-partition cycles per instruction for single classes and for 1:1 / n:1 mixes, at 1..4 warps
// per scheduler, plus dependent-chain latencies.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

enum { FFMA, FFMA2, FMUL2, FADD2, DFMA, MUFU, MUFU64, F2F_DS, F2F_SD, IMAD, LOP3, ISETP_SEL, LDS128, NOPS };

template <int OP>
__device__ __forceinline__ void op1(float &a, float2 &p, double &d, int &n, float m, float c, double md, double cd, const double2 *sm)
{
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(m), "f"(c));
    if (OP == FFMA2) {
        unsigned long long r, mm, cc;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(p.x), "f"(p.y));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r) : "l"(mm), "l"(cc));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(p.x), "=f"(p.y) : "l"(r));
    }
    if (OP == FMUL2) {
        unsigned long long r, mm;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(p.x), "f"(p.y));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r) : "l"(mm));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(p.x), "=f"(p.y) : "l"(r));
    }
    if (OP == FADD2) {
        unsigned long long r, cc;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(p.x), "f"(p.y));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r) : "l"(cc));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(p.x), "=f"(p.y) : "l"(r));
    }
    if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(md), "d"(cd));
    if (OP == MUFU) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a));
    if (OP == MUFU64) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d));
    if (OP == F2F_DS) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a));
    if (OP == F2F_SD) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a) : "d"(d));
    if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n) : "r"(n | 3), "r"(n));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n) : "r"((int)__float_as_int(m)), "r"((int)__float_as_int(c)));
    if (OP == ISETP_SEL) { int t; asm volatile("{ .reg .pred q; setp.gt.s32 q, %1, %2; selp.b32 %0, %1, %2, q; }" : "=r"(t) : "r"(n), "r"((int)__float_as_int(m))); n = t; }
    if (OP == LDS128) { double2 v = sm[(n & 31)]; d += v.x; }
}

// NA ops of class A then NB ops of class B per chain step, 8 independent chains per thread
template <int A, int NA, int B, int NBB>
__global__ void mix_kernel(int iters, float seed, float *sink, long long *cycles)
{
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(seed, seed);
    __syncthreads();
    float a[8]; float2 p[8]; double d[8]; int n[8];
    float a2[8]; float2 p2[8]; double d2[8]; int n2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = seed + threadIdx.x + j; p[j] = make_float2(a[j], a[j] + 1); d[j] = a[j]; n[j] = (int)a[j];
        a2[j] = a[j] + 3; p2[j] = make_float2(a2[j], a2[j] + 1); d2[j] = a2[j]; n2[j] = n[j] + 3;
    }
    const float m = 0.9999f + seed * 1e-9f, c = 1e-7f + seed * 1e-12f;
    const double md = m, cd = c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int r = 0; r < NA; ++r) op1<A>(a[j], p[j], d[j], n[j], m, c, md, cd, sm);
#pragma unroll
            for (int r = 0; r < NBB; ++r) op1<B>(a2[j], p2[j], d2[j], n2[j], m, c, md, cd, sm);
        }
    }
    long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + p[j].x + p[j].y + (float)d[j] + (float)n[j] + a2[j] + p2[j].x + p2[j].y + (float)d2[j] + (float)n2[j];
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// latency: one dependent chain, one warp
template <int A>
__global__ void lat_kernel(int iters, float seed, float *sink, long long *cycles)
{
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(seed, seed);
    __syncthreads();
    float a = seed + threadIdx.x; float2 p = make_float2(a, a + 1); double d = a; int n = (int)a;
    const float m = 0.9999f + seed * 1e-9f, c = 1e-7f + seed * 1e-12f;
    const double md = m, cd = c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (A == F2F_DS || A == F2F_SD) { op1<F2F_DS>(a, p, d, n, m, c, md, cd, sm); op1<F2F_SD>(a, p, d, n, m, c, md, cd, sm); }
            else op1<A>(a, p, d, n, m, c, md, cd, sm);
        }
    }
    long long t1 = clock64();
    float s = a + p.x + p.y + (float)d + (float)n;
    if (s == 12345.678f) sink[0] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

static float *sink; static long long *cyc;

template <int A, int NA, int B, int NBB>
void run_mix(const char *name)
{
    printf("%-34s", name);
    for (int warps = 1; warps <= 4; ++warps) {           // warps per scheduler; one CTA per SM
        const int threads = warps * 128, iters = 2000;
        mix_kernel<A, NA, B, NBB><<<148, threads>>>(iters, 1.0f, sink, cyc);
        CHK(cudaDeviceSynchronize());
        long long h; CHK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double per_warp_inst = (double)h / ((double)iters * 8 * (NA + NBB) * warps);
        printf("  w%d %6.3f", warps, per_warp_inst);   // scheduler cycles per warp instruction
    }
    printf("   (cycles per warp-instruction per scheduler)\n");
}

template <int A>
void run_lat(const char *name)
{
    lat_kernel<A><<<1, 32>>>(2000, 1.0f, sink, cyc);
    CHK(cudaDeviceSynchronize());
    long long h; CHK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double per = (double)h / (2000.0 * 16);
    printf("latency %-26s %7.2f cycles%s\n", name, per, (A == F2F_DS || A == F2F_SD) ? " (f32->f64 + f64->f32 pair)" : "");
}

int main()
{
    CHK(cudaMalloc(&sink, 4)); CHK(cudaMalloc(&cyc, 8));
    run_mix<FFMA, 1, NOPS, 0>("FFMA");
    run_mix<FFMA2, 1, NOPS, 0>("FFMA2");
    run_mix<FMUL2, 1, NOPS, 0>("FMUL2");
    run_mix<FADD2, 1, NOPS, 0>("FADD2");
    run_mix<DFMA, 1, NOPS, 0>("DFMA");
    run_mix<MUFU, 1, NOPS, 0>("MUFU.RCP");
    run_mix<MUFU64, 1, NOPS, 0>("MUFU.RCP64H");
    run_mix<F2F_DS, 1, F2F_SD, 1>("F2F pair (f32->f64->f32)");
    run_mix<IMAD, 1, NOPS, 0>("IMAD");
    run_mix<LOP3, 1, NOPS, 0>("LOP3");
    run_mix<ISETP_SEL, 1, NOPS, 0>("ISETP+SEL");
    run_mix<LDS128, 1, NOPS, 0>("LDS.128 + DADD");
    run_mix<FFMA, 1, DFMA, 1>("FFMA : DFMA 1:1");
    run_mix<FFMA, 2, DFMA, 1>("FFMA : DFMA 2:1");
    run_mix<FFMA, 3, DFMA, 1>("FFMA : DFMA 3:1");
    run_mix<FFMA2, 1, DFMA, 1>("FFMA2 : DFMA 1:1");
    run_mix<FFMA2, 2, DFMA, 1>("FFMA2 : DFMA 2:1");
    run_mix<FFMA, 1, FFMA2, 1>("FFMA : FFMA2 1:1");
    run_mix<FFMA, 1, LOP3, 1>("FFMA : LOP3 1:1");
    run_mix<FFMA, 1, IMAD, 1>("FFMA : IMAD 1:1");
    run_mix<DFMA, 1, LOP3, 1>("DFMA : LOP3 1:1");
    run_mix<DFMA, 1, IMAD, 1>("DFMA : IMAD 1:1");
    run_mix<FFMA, 4, MUFU, 1>("FFMA : MUFU 4:1");
    run_mix<FFMA, 8, MUFU, 1>("FFMA : MUFU 8:1");
    run_mix<DFMA, 4, MUFU, 1>("DFMA : MUFU 4:1");
    run_mix<FFMA, 4, F2F_DS, 1>("FFMA : F2F(f32->f64) 4:1");
    run_mix<FFMA, 4, F2F_SD, 1>("FFMA : F2F(f64->f32) 4:1");
    run_mix<MUFU, 1, F2F_DS, 1>("MUFU : F2F(f32->f64) 1:1");
    run_mix<FFMA, 4, LDS128, 1>("FFMA : LDS.128+DADD 4:1");
    run_lat<FFMA>("FFMA"); run_lat<FFMA2>("FFMA2"); run_lat<FMUL2>("FMUL2"); run_lat<FADD2>("FADD2");
    run_lat<DFMA>("DFMA"); run_lat<MUFU>("MUFU.RCP"); run_lat<MUFU64>("MUFU.RCP64H");
    run_lat<F2F_DS>("F2F"); run_lat<IMAD>("IMAD"); run_lat<LOP3>("LOP3"); run_lat<ISETP_SEL>("ISETP+SEL"); run_lat<LDS128>("LDS.128+DADD");
    return 0;
}
