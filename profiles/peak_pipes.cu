// peak_pipes.cu — measured issue-rate ceilings of the CUDA-core pipes the non-GEMM kernels are bound by (SURVEY.md §8(d):
// "measure the FP32/MUFU peak with a micro-benchmark; MEASURED_PEAKS.json has no fp32 entry").
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/peak_pipes profiles/peak_pipes.cu && profiles/peak_pipes
//
// Every test keeps 8 independent dependency chains per thread, 1024 threads per SM x 2 resident blocks, and reports
// lane-operations per second over the whole GPU (CUDA events, best of 5).  One JSON line on stdout.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int OP>
__global__ void __launch_bounds__(512) pipe_kernel(float* out, float a, float b) {
    float x[CHAINS];
    double y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = a + (float)(threadIdx.x + c); y[c] = (double)x[c]; }
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) x[c] = __fmaf_rn(x[c], a, b);                                       // FFMA
            else if (OP == 1) x[c] = __fadd_rn(x[c], b);                                     // FADD (separately rounded sums)
            else if (OP == 2) x[c] = __fmul_rn(x[c], a);                                     // FMUL
            else if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));       // MUFU.RCP
            else if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));       // MUFU.LG2
            else if (OP == 5) x[c] = __fdiv_rn(x[c], a);                                     // IEEE division (the chi-square term)
            else if (OP == 6) y[c] = __fma_rn(y[c], (double)a, (double)b);                   // DFMA
            else if (OP == 7) y[c] = __dadd_rn(y[c], (double)b);                             // DADD
            else if (OP == 8) { float t = __fsub_rn(x[c], a); x[c] = __fadd_rn(x[c], __fmul_rn(t, t)); }   // the L2 term: 3 rounded ops
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + (float)y[c];
    if (s == 12345.678f) out[0] = s;
}

template <int OP>
static double run(int n_sm, float* out, double ops_per_iter) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = n_sm * 4;
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(a);
        pipe_kernel<OP><<<blocks, 512>>>(out, 1.0000001f, 1e-7f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    return (double)blocks * 512 * ITERS * CHAINS * ops_per_iter / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    float* out; cudaMalloc(&out, 4);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_khz_max\": %d, \"unit\": \"lane-ops/s\", ", p.name, p.multiProcessorCount, clk);
    printf("\"ffma\": %.4e, ", run<0>(p.multiProcessorCount, out, 1));
    printf("\"fadd\": %.4e, ", run<1>(p.multiProcessorCount, out, 1));
    printf("\"fmul\": %.4e, ", run<2>(p.multiProcessorCount, out, 1));
    printf("\"mufu_rcp\": %.4e, ", run<3>(p.multiProcessorCount, out, 1));
    printf("\"mufu_lg2\": %.4e, ", run<4>(p.multiProcessorCount, out, 1));
    printf("\"fdiv_ieee\": %.4e, ", run<5>(p.multiProcessorCount, out, 1));
    printf("\"dfma\": %.4e, ", run<6>(p.multiProcessorCount, out, 1));
    printf("\"dadd\": %.4e, ", run<7>(p.multiProcessorCount, out, 1));
    printf("\"l2_term_sub_mul_add\": %.4e}\n", run<8>(p.multiProcessorCount, out, 1));
    return 0;
}
