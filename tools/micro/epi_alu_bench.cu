// Microbenchmark: ALU throughput of the FC1 / FC2-dgrad epilogue arithmetic (packed fp32x2 GELU) as a function of the number of
// resident warps per SM, to decide how many epilogue warps the GEMM kernel needs. Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -I once-for-both_b200/csrc tools/micro/epi_alu_bench.cu -o gpurun_out/epi_alu_bench
#include "ptx.cuh"
#include <cstdio>
using namespace ofb;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, float seed) {
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(seed + threadIdx.x * 1e-3f + i, seed - i * 0.37f);
    float2 acc = make_float2(0.f, 0.f);
    const float2 b2 = splat2(0.01f), g2 = splat2(0.9f), r2 = splat2(1.1f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {            // FC1 epilogue arithmetic
                const float2 u = add2(v[i], b2);
                const float2 z = mul2(u, g2);
                float2 h = mul2(z, gelu_cdf2(z));
                h = mul2(h, r2);
                acc = add2(acc, make_float2(__uint_as_float(pack_bf16x2(u.x, u.y)), __uint_as_float(pack_bf16x2(h.x, h.y))));
            } else if (MODE == 1) {     // FC2-dgrad epilogue arithmetic
                float2 Phi, dg;
                gelu_terms2(mul2(v[i], g2), Phi, dg);
                const float2 w = mul2(b2, dg);
                acc = fma2(w, v[i], acc);
                acc = add2(acc, mul2(w, r2));
            } else if (MODE == 2) {     // pure FFMA2 chain, 16 independent accumulators
                v[i] = fma2(v[i], g2, b2);
            } else if (MODE == 3) {     // scalar FFMA, immediate operand
                v[i].x = fmaf(v[i].x, 0.9f, 0.01f); v[i].y = fmaf(v[i].y, 0.9f, 0.01f);
            } else {                    // scalar FFMA, three registers
                v[i].x = fmaf(v[i].x, g2.x, b2.x); v[i].y = fmaf(v[i].y, g2.y, b2.y);
            }
        }
        if (MODE <= 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = add2(v[i], splat2(1e-3f));
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = add2(acc, v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

template <int MODE>
void run(const char* name, float* out) {
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 4000;
        k<MODE><<<148, warps * 32>>>(out, 10, 0.5f);
        cudaEventRecord(e0);
        k<MODE><<<148, warps * 32>>>(out, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double pairs = double(iters) * 16 * warps * 32;        // per SM
        printf("%-28s warps/SM %2d: %.3f ms  -> %.2f element-pairs / ns / SM (%.1f warp-pairs per us per SMSP)\n", name, warps, ms,
               pairs / (ms * 1e6), pairs / 32 / 4 / (ms * 1e3));
    }
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    run<0>("FC1 epilogue math", out);
    run<1>("FC2-dgrad epilogue math", out);
    run<2>("FFMA2 (16 independent)", out);
    run<3>("FFMA imm x2", out);
    run<4>("FFMA reg x2", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
