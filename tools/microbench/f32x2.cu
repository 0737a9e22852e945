// Microbenchmark: does the packed fp32x2 arithmetic of sm_100 (FFMA2 / FMUL2 / FADD2, `__ffma2_rn` ...) retire two IEEE
// binary32 operations per lane per issue slot? Every lane runs 8 independent dependent chains; time per instruction is compared
// between the scalar and the packed form. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false f32x2.cu -o f32x2
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void scalarFma(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packedFma(float* out, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], A, B);
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void scalarMulAdd(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fadd_rn(__fmul_rn(x[i], a), b);
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packedMulAdd(float* out, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fadd2_rn(__fmul2_rn(x[i], A), B);
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: packed FP + integer work in the freed issue slots
__global__ void mixedScalar(float* out, float a, float b, unsigned k) {
    float x[8]; unsigned u[4];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 4; i++) u[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
#pragma unroll
        for (int i = 0; i < 4; i++) { u[i] = (u[i] ^ k) + (u[i] >> 3); u[i] = (u[i] & 0xffffffu) * 3u + i; }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
}
__global__ void mixedPacked(float* out, float a, float b, unsigned k) {
    float2 x[4]; unsigned u[4];
    for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    for (int i = 0; i < 4; i++) u[i] = threadIdx.x + i;
    const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = __ffma2_rn(x[i], A, B);
#pragma unroll
        for (int i = 0; i < 4; i++) { u[i] = (u[i] ^ k) + (u[i] >> 3); u[i] = (u[i] & 0xffffffu) * 3u + i; }
    }
    float s = 0;
    for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] ^ u[1] ^ u[2] ^ u[3]);
}
template <typename F> float timeIt(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int grid = 148 * 8, block = 1024;
    const double lanes = (double)grid * block, flopsPerLaneScalar = 8.0 * ITER;
    float t;
    t = timeIt([&] { scalarFma<<<grid, block>>>(out, 0.999f, 0.001f); });    printf("scalar FFMA      : %.3f ms, %.1f G lane-op/s\n", t, lanes * flopsPerLaneScalar / t / 1e6);
    t = timeIt([&] { packedFma<<<grid, block>>>(out, 0.999f, 0.001f); });    printf("packed FFMA2     : %.3f ms, %.1f G lane-op/s (16 binary32 fma per iteration)\n", t, lanes * 2 * flopsPerLaneScalar / t / 1e6);
    t = timeIt([&] { scalarMulAdd<<<grid, block>>>(out, 0.999f, 0.001f); }); printf("scalar FMUL+FADD : %.3f ms, %.1f G lane-op/s\n", t, lanes * 2 * flopsPerLaneScalar / t / 1e6);
    t = timeIt([&] { packedMulAdd<<<grid, block>>>(out, 0.999f, 0.001f); }); printf("packed FMUL2+FADD2: %.3f ms, %.1f G lane-op/s\n", t, lanes * 4 * flopsPerLaneScalar / t / 1e6);
    t = timeIt([&] { mixedScalar<<<grid, block>>>(out, 0.999f, 0.001f, 12345u); }); printf("mixed scalar (8 fma + int) : %.3f ms\n", t);
    t = timeIt([&] { mixedPacked<<<grid, block>>>(out, 0.999f, 0.001f, 12345u); }); printf("mixed packed (4 fma2 + int): %.3f ms\n", t);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
