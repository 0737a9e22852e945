// CPU check of the fused froxel column path (test infrastructure; built by tests/test_froxel_fusion_cpu.py with nvcc as HOST code).
// The phase functions of plainrenderer_b200/csrc/froxel_inc.cuh are host+device: the kernel (passes_volumetrics.cu froxelColumnKernel)
// calls each once per thread with a barrier in between, this file loops over the thread indices of a block per phase - the same
// statements, the same tables, the same order of the running sums. The test holds the four volumes it returns against the CPU
// oracle's four separate passes, bit for bit (Volumetrics.cpp:151-246 and the four shaders cited in froxel_inc.cuh).
#include <type_traits>
#include <vector>
#include "froxel_inc.cuh"

using namespace pb;

extern "C" __attribute__((visibility("default"))) int froxel_fused_host(
    int w, int h, int d, const unsigned char* noise, int noiseSize, const unsigned short* shadow, int shadowSize, const plain_shadow_cascade_info* cascades,
    const plain_light_buffer* light, const plain_volumetric_lighting_settings* settings, const unsigned short* history, const plain_global_shader_info* g,
    int yBegin, int yEnd, int zLanes, unsigned short* material, unsigned short* scattering, unsigned short* target, unsigned short* integrated) {
    if (d > FROXEL_MAX_DEPTH) return 1;
    FroxelFusedParams p;
    auto view = [&](const void* ptr, int vw, int vh, int vd) { ImgView v; v.ptr = (unsigned char*)ptr; v.w = vw; v.h = vh; v.d = vd; return v; };
    p.historyTarget = view(target, w, h, d);
    p.integrationVolume = view(integrated, w, h, d);
    p.materialVolume = view(material, w, h, d);
    p.scatteringVolume = view(scattering, w, h, d);
    p.in.noiseTexture = view(noise, noiseSize, noiseSize, noiseSize);
    p.in.sunShadowMap = view(shadow, shadowSize, shadowSize, 1);
    p.in.historyVolume = view(history, w, h, d);
    p.in.cascades = cascades;
    p.in.light = light;
    p.in.g = g;
    p.settings = settings;
    p.yBegin = yBegin;
    std::vector<FroxelBlockShared> shared(1);
    FroxelBlockShared& sh = shared[0];
    const int blocksX = (w + FROXEL_COLS - 1) / FROXEL_COLS;
    auto run = [&](auto lanes) {
        constexpr int ZLANES = decltype(lanes)::value;
        for (int by = 0; by < yEnd - yBegin; by++)              // blockIdx.y
            for (int blockX = 0; blockX < blocksX; blockX++) {  // blockIdx.x
                const int y = p.yBegin + by;
                const plain_volumetric_lighting_settings s = *p.settings;
                const Globals G = loadGlobals(p.in.g);
                for (int tid = 0; tid < FROXEL_COLS * ZLANES; tid++) froxelBlockPrologue<ZLANES>(sh, p, G, s, tid, blockX, y);
                for (int tid = 0; tid < FROXEL_COLS * ZLANES; tid++) froxelBlockPhase1<ZLANES, true>(sh, p, G, s, tid, blockX, y);
                for (int tid = 0; tid < FROXEL_COLS * ZLANES; tid++) froxelBlockPhase2(sh, p, tid, blockX);
                for (int tid = 0; tid < FROXEL_COLS * ZLANES; tid++) froxelBlockPhase3<ZLANES>(sh, p, tid, blockX, y);
            }
    };
    if (zLanes == 16) run(std::integral_constant<int, 16>());
    else if (zLanes == 32) run(std::integral_constant<int, 32>());
    else if (zLanes == 64) run(std::integral_constant<int, 64>());
    else return 2;
    return 0;
}
