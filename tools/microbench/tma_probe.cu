// Stand-alone probe of the TMA tile load the bloom A/B kernel uses: one 72x24 box of 32-bit texels from a w x h image, coordinates
// partly outside the image (zero fill). Prints whether the tile matches. nvcc -gencode arch=compute_100a,code=sm_100a tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#define TW 72
#define TH 24
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, uint32_t* out, int x0, int y0) {
    __shared__ __align__(128) uint32_t sRaw[TW * TH];
    __shared__ __align__(8) unsigned long long sBar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(&sBar)), "r"((uint32_t)(TW * TH * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smemAddr(sRaw)), "l"(&map), "r"(x0), "r"(y0), "r"(smemAddr(&sBar)) : "memory");
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        if (++spins > (1u << 20)) { if (threadIdx.x == 0) out[TW * TH] = 0xdead; return; }
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smemAddr(&sBar)) : "memory");
    }
    for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) out[i] = sRaw[i];
    if (threadIdx.x == 0) out[TW * TH] = spins;
}
int main() {
    const int w = 240, h = 135;
    std::vector<uint32_t> img(w * h);
    for (int i = 0; i < w * h; i++) img[i] = 0x10000000u + i;
    uint32_t *dImg, *dOut;
    cudaMalloc(&dImg, w * h * 4); cudaMalloc(&dOut, (TW * TH + 1) * 4);
    cudaMemcpy(dImg, img.data(), w * h * 4, cudaMemcpyHostToDevice);
    cudaDriverEntryPointQueryResult q; void* fn = nullptr;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(ge), (int)q, fn);
    auto encode = (PFN_cuTensorMapEncodeTiled)fn;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}; const cuuint64_t strides[1] = {(cuuint64_t)w * 4};
    const cuuint32_t box[2] = {TW, TH}, es[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dImg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    for (int t = 0; t < 3; t++) {
        const int x0 = t == 0 ? 124 : (t == 1 ? -4 : 200), y0 = t == 0 ? 13 : (t == 1 ? -3 : 120);
        cudaMemset(dOut, 0xff, (TW * TH + 1) * 4);
        probe<<<1, 256>>>(map, dOut, x0, y0);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint32_t> out(TW * TH + 1);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < TH; y++) for (int x = 0; x < TW; x++) {
            const int ax = x0 + x, ay = y0 + y;
            const uint32_t want = (ax >= 0 && ay >= 0 && ax < w && ay < h) ? img[ay * w + ax] : 0u;
            if (out[y * TW + x] != want) bad++;
        }
        printf("tile at (%d, %d): %s, spins/marker %u, %d wrong texels\n", x0, y0, cudaGetErrorString(e), out[TW * TH], bad);
    }
    return 0;
}
