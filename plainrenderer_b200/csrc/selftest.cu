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
    cudaError_t e = cudaMemcpyAsync(hostOut8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) { error = std::string("device_selftest: ") + cudaGetErrorString(e); return false; }
    return true;
}

}  // namespace pb
