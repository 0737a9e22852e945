// Are FADD2 / FMUL2 / FFMA2 bit-identical to the scalar instructions (built with -fmad=false, no -ftz)? Random and special operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false f32x2_bits.cu -o f32x2_bits
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ uint32_t rnd(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__device__ float pick(uint32_t& s) {
    const uint32_t r = rnd(s), k = rnd(s) % 16;
    uint32_t b = r;
    if (k == 0) b = r & 0x807fffffu;                      // denormal / zero
    else if (k == 1) b = (r & 0x80000000u) | 0x7f800000u;  // inf
    else if (k == 2) b = r | 0x7f800000u;                  // NaN (any payload) or inf
    else if (k == 3) b = (r & 0x80000000u);                // +-0
    else if (k == 4) b = (r & 0x80ffffffu) | 0x00800000u;  // tiny normal
    else if (k == 5) b = (r & 0x807fffffu) | 0x7f000000u;  // huge
    return __uint_as_float(b);
}
__global__ void check(unsigned long long* bad, uint32_t* example, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        const float a0 = pick(s), a1 = pick(s), b0 = pick(s), b1 = pick(s), c0 = pick(s), c1 = pick(s);
        const float2 A = make_float2(a0, a1), B = make_float2(b0, b1), Cc = make_float2(c0, c1);
        const float2 add = __fadd2_rn(A, B), mul = __fmul2_rn(A, B), fma = __ffma2_rn(A, B, Cc), sub = __ffma2_rn(B, make_float2(-1.f, -1.f), A);
        const float sa0 = __fadd_rn(a0, b0), sa1 = __fadd_rn(a1, b1), sm0 = __fmul_rn(a0, b0), sm1 = __fmul_rn(a1, b1);
        const float sf0 = __fmaf_rn(a0, b0, c0), sf1 = __fmaf_rn(a1, b1, c1), ss0 = __fsub_rn(a0, b0), ss1 = __fsub_rn(a1, b1);
        const uint32_t p[8] = {__float_as_uint(add.x), __float_as_uint(add.y), __float_as_uint(mul.x), __float_as_uint(mul.y), __float_as_uint(fma.x), __float_as_uint(fma.y), __float_as_uint(sub.x), __float_as_uint(sub.y)};
        const uint32_t q[8] = {__float_as_uint(sa0), __float_as_uint(sa1), __float_as_uint(sm0), __float_as_uint(sm1), __float_as_uint(sf0), __float_as_uint(sf1), __float_as_uint(ss0), __float_as_uint(ss1)};
        for (int k = 0; k < 8; k++)
            if (p[k] != q[k]) {
                const unsigned long long n = atomicAdd(&bad[k / 2], 1ull);
                if (n < 4) { uint32_t* e = example + ((k / 2) * 4 + n) * 5; e[0] = __float_as_uint(k & 1 ? a1 : a0); e[1] = __float_as_uint(k & 1 ? b1 : b0); e[2] = __float_as_uint(k & 1 ? c1 : c0); e[3] = p[k]; e[4] = q[k]; }
            }
    }
}
int main() {
    unsigned long long* bad; uint32_t* ex;
    cudaMallocManaged(&bad, 4 * 8); cudaMallocManaged(&ex, 4 * 4 * 5 * 4);
    for (int i = 0; i < 4; i++) bad[i] = 0;
    for (int i = 0; i < 80; i++) ex[i] = 0;
    check<<<148 * 4, 256>>>(bad, ex, 20000);
    cudaDeviceSynchronize();
    const char* names[4] = {"FADD2 vs FADD", "FMUL2 vs FMUL", "FFMA2 vs FFMA", "FFMA2(b,-1,a) vs FSUB"};
    for (int k = 0; k < 4; k++) {
        printf("%s: %llu mismatches of %llu\n", names[k], bad[k], 2ull * 148 * 4 * 256 * 20000);
        for (int n = 0; n < 4 && n < (int)bad[k]; n++) { const uint32_t* e = ex + (k * 4 + n) * 5; printf("   a=%08x b=%08x c=%08x packed=%08x scalar=%08x\n", e[0], e[1], e[2], e[3], e[4]); }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
