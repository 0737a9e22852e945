// selftest.cu - exhaustive on-device checks of the instruction-lean sequences against the functions of the numeric contract
// they replace (include/plain_b200.h plain_device_selftest). Every kernel counts mismatching BIT PATTERNS (two NaNs agree).
#include "pass_common.cuh"

namespace pb {

__device__ __forceinline__ bool sameBits(float a, float b) { return (a != a && b != b) || dm::f2u(a) == dm::f2u(b); }

// [0] rcpf_nz, [1] rcpf_normal against __frcp_rn; [2] encoders; [4] floor2i
__global__ void selftestAllFloatsKernel(unsigned long long* out) {
    unsigned long long bad0 = 0, bad1 = 0, bad2 = 0, bad4 = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t k = 0; k < (1u << 20); k++, u += stride) {  // grid of 2^12 threads x 2^20 iterations = 2^32 values
        const float x = dm::u2f(u);
        const float ax = fabsf(x);
        const bool zeroOrDenormal = ax < 1.17549435e-38f;  // false for NaN
        if (!zeroOrDenormal && !sameBits(rcpf_nz(x), __frcp_rn(x))) bad0++;
        if (ax >= 1.17549435e-38f && ax < 8.507059173e37f && !sameBits(rcpf_normal(x), __frcp_rn(x))) bad1++;
        if (encodeSmallFloatFast(x, 6) != encodeSmallFloat(x, 6)) bad2++;
        if (encodeSmallFloatFast(x, 5) != encodeSmallFloat(x, 5)) bad2++;
        if (ax < 16777216.f) {
            const int i = floor2i(x);
            if (i != f2i(floorf_(x)) || !sameBits((float)i, floorf_(x))) bad4++;
        }
    }
    if (bad0) atomicAdd(out + 0, bad0);
    if (bad1) atomicAdd(out + 1, bad1);
    if (bad2) atomicAdd(out + 2, bad2);
    if (bad4) atomicAdd(out + 4, bad4);
}
// [6] sqrtf_normal / sqrt2_normal against sqrtf over every x with 2^-60 <= x <= 2^60; [7] divf_normal / div2_normal against IEEE division: every
// such divisor under the numerators the frame uses (0.25, near * far = 30, and four others), and pseudo-random moderate operand pairs
__global__ void selftestNormalRangeKernel(unsigned long long* out) {
#if defined(__CUDA_ARCH__)  // the sequences under test are device-only (inline PTX)
    unsigned long long bad6 = 0, bad7 = 0;
    const uint32_t lo = 0x21800000u, hi = 0x5d800000u;  // 2^-60 .. 2^60
    const uint32_t stride = gridDim.x * blockDim.x;
    const float numerators[6] = {0.25f, 30.f, 1.f, 3.1415927f, 1.0e-12f, 7.7e11f};
    for (uint32_t u = lo + blockIdx.x * blockDim.x + threadIdx.x; u <= hi; u += stride) {
        const float x = dm::u2f(u), nx = -x;
        if (!sameBits(sqrtf_normal(x), sqrtf(x))) bad6++;
        const float2 s2 = sqrt2_normal(make_float2(x, x * 1.5f));  // (x * 1.5 <= 2^60.6: inside the sequence's own range)
        if (!sameBits(s2.x, sqrtf(x)) || !sameBits(s2.y, sqrtf(x * 1.5f))) bad6++;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const float a = numerators[k];
            if (!sameBits(divf_normal(a, x), a / x) || !sameBits(divf_normal(a, nx), a / nx)) bad7++;
        }
        const float2 d2 = div2_normal(make_float2(0.25f, 30.f), make_float2(x, nx));
        if (!sameBits(d2.x, 0.25f / x) || !sameBits(d2.y, 30.f / nx)) bad7++;
        const float2 r2 = rcp2_normal(make_float2(x, nx));
        if (!sameBits(r2.x, __frcp_rn(x)) || !sameBits(r2.y, __frcp_rn(nx))) bad7++;
    }
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x + 1u) * 2654435761u;
    for (int k = 0; k < (1 << 16); k++) {  // 2^16 threads x 2^16 pairs
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        const float a = dm::u2f((lo + s % (hi - lo + 1u)) | (s & 0x80000000u));
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        const float b = dm::u2f((lo + s % (hi - lo + 1u)) | (s & 0x80000000u));
        if (!sameBits(divf_normal(a, b), a / b)) bad7++;
    }
    if (bad6) atomicAdd(out + 6, bad6);
    if (bad7) atomicAdd(out + 7, bad7);
#endif
}
// [3] decoders over every code, [5] FMNMX against the pinned min / max over pseudo-random operand pairs without -0
__global__ void selftestCodesKernel(unsigned long long* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad3 = 0, bad5 = 0;
    if (t < 2048u) {
        if (!sameBits(decodeSmallFloatFast(t, 6), decodeSmallFloat(t, 6))) bad3++;
        if (t < 1024u && !sameBits(decodeSmallFloatFast(t, 5), decodeSmallFloat(t, 5))) bad3++;
    }
    uint32_t s = (t + 1u) * 2654435761u;  // never zero for t < 2^16: xorshift stays non-zero
    for (int k = 0; k < 1024; k++) {
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        uint32_t a = s;
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        uint32_t b = (k & 7) == 0 ? a : s;  // equal operands every eighth pair
        if (a == 0x80000000u) a = 0u;
        if (b == 0x80000000u) b = 0u;
        const float x = dm::u2f(a), y = dm::u2f(b);
        if (!sameBits(fmin_nn(x, y), fminp(x, y)) || !sameBits(fmax_nn(x, y), fmaxp(x, y))) bad5++;
    }
    if (bad3) atomicAdd(out + 3, bad3);
    if (bad5) atomicAdd(out + 5, bad5);
}

bool runDeviceSelftest(cudaStream_t stream, unsigned long long* hostOut8, std::string& error) {
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 8 * sizeof(unsigned long long)) != cudaSuccess) { error = "device_selftest: cudaMalloc failed"; return false; }
    cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), stream);
    selftestAllFloatsKernel<<<16, 256, 0, stream>>>(d);
    selftestCodesKernel<<<256, 256, 0, stream>>>(d);
    selftestNormalRangeKernel<<<256, 256, 0, stream>>>(d);
    cudaError_t e = cudaMemcpyAsync(hostOut8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) { error = std::string("device_selftest: ") + cudaGetErrorString(e); return false; }
    return true;
}

}  // namespace pb
