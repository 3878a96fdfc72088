// ubench_fp32x2.cu -- what does sm_100a's packed FP32 (FFMA2 / FMUL2 / FADD2) buy?
//
// Dependent-free instruction streams on every SM; prints warp-instructions per clock per SM and
// the FP32 rate for scalar FFMA, packed FFMA2/FMUL2/FADD2, and mixes with MUFU / ALU-pipe work.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32x2 scripts/ubench_fp32x2.cu
// The numbers decide how K1 / E3 pair their samples (kernels section).
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kThreads = 256;
constexpr int kIters = 2048;

template <int kMode>
__global__ void __launch_bounds__(kThreads) stream_kernel(float* out, float m_, float c_) {
    const float t = threadIdx.x * 1e-3f;
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(t + i, t + i + 0.5f);
    const float2 m = make_float2(m_, m_), c = make_float2(c_, c_);
    float x0 = t + 1.0f, x1 = t + 2.0f;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (kMode == 0) {           // scalar FFMA x2
                    a[i].x = fmaf(a[i].x, m_, c_);
                    a[i].y = fmaf(a[i].y, m_, c_);
                } else if (kMode == 1) {    // FFMA2
                    a[i] = __ffma2_rn(a[i], m, c);
                } else if (kMode == 2) {    // FMUL2
                    a[i] = __fmul2_rn(a[i], m);
                } else if (kMode == 3) {    // FADD2
                    a[i] = __fadd2_rn(a[i], c);
                } else if (kMode == 4) {    // 8 FFMA2 : 1 MUFU
                    a[i] = __ffma2_rn(a[i], m, c);
                } else if (kMode == 5) {    // 8 FFMA2 : 2 MUFU : 2 FMNMX (K1-like mix)
                    a[i] = __ffma2_rn(a[i], m, c);
                } else if (kMode == 6) {    // scalar FMUL (no immediate) x2
                    a[i].x = a[i].x * m_;
                    a[i].y = a[i].y * m_;
                } else if (kMode == 7) {    // FFMA2, three distinct register operands
                    a[i] = __ffma2_rn(a[i], a[(i + 1) & 7], a[(i + 3) & 7]);
                } else if (kMode == 8) {    // scalar FFMA x2, three distinct register operands
                    a[i].x = fmaf(a[i].x, a[(i + 1) & 7].x, a[(i + 3) & 7].x);
                    a[i].y = fmaf(a[i].y, a[(i + 1) & 7].y, a[(i + 3) & 7].y);
                } else if (kMode == 9) {    // FMUL2 reg x reg distinct
                    a[i] = __fmul2_rn(a[i], a[(i + 1) & 7]);
                } else if (kMode == 10) {   // FFMA2 with a dependent chain of length 2 (latency probe): a = fma(fma(a,m,c),m,c)
                    a[i] = __ffma2_rn(__ffma2_rn(a[i], m, c), m, c);
                }
            }
            if (kMode == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x0)); }
            if (kMode == 5) {
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x0));
                asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x1));
                x0 = fmaxf(x0, 1e-30f); x1 = fmaxf(x1, 1e-30f);
            }
        }
    }
    float s = x0 + x1;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * kThreads + threadIdx.x] = s;
}

template <int kMode>
void run(const char* name, double flops_per_inner, double insts_per_inner, float* d_out, int sms, double ghz) {
    const int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<kMode><<<blocks, kThreads>>>(d_out, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 1 && ms < best) best = ms;
    }
    const double inner = (double)kIters * 4 * blocks * kThreads;           // executions of the u-loop body per thread-set
    const double tflops = flops_per_inner * inner / (best * 1e-3) / 1e12;
    const double winst_per_clk_sm = insts_per_inner * inner / 32.0 / (best * 1e-3) / (ghz * 1e9) / sms;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s  %5.2f warp-inst/clk/SM (at %.3f GHz)\n", name, best, tflops, winst_per_clk_sm, ghz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float* d_out; cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 8 * kThreads * 4);
    printf("%s, %d SMs, clock attr %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<0>("scalar FFMA (16 per body)", 32, 16, d_out, p.multiProcessorCount, ghz);
    run<6>("scalar FMUL reg*reg (16 per body)", 16, 16, d_out, p.multiProcessorCount, ghz);
    run<1>("FFMA2 (8 per body)", 32, 8, d_out, p.multiProcessorCount, ghz);
    run<2>("FMUL2 (8 per body)", 16, 8, d_out, p.multiProcessorCount, ghz);
    run<3>("FADD2 (8 per body)", 16, 8, d_out, p.multiProcessorCount, ghz);
    run<4>("8 FFMA2 + 1 MUFU", 32, 9, d_out, p.multiProcessorCount, ghz);
    run<5>("8 FFMA2 + 2 MUFU + 2 FMNMX", 32, 12, d_out, p.multiProcessorCount, ghz);
    run<7>("FFMA2 reg,reg,reg distinct (8 per body)", 32, 8, d_out, p.multiProcessorCount, ghz);
    run<8>("scalar FFMA reg,reg,reg distinct (16)", 32, 16, d_out, p.multiProcessorCount, ghz);
    run<9>("FMUL2 reg,reg distinct (8 per body)", 16, 8, d_out, p.multiProcessorCount, ghz);
    run<10>("FFMA2 dependent pairs (16 per body)", 64, 16, d_out, p.multiProcessorCount, ghz);
    return 0;
}
