// pass_common.cuh - what every CUDA pass launcher sees: the execution record of one compute pass resolved to device
// views (the counterpart of the descriptor sets the reference binds, RenderBackend.cpp:845-890), the launch helpers
// and the registry that maps the reference's shader file name to a launcher (ResourceDescriptions.h:112-120).
//
// Memory model (DESIGN.md "Data layout"): every image is one HBM allocation, mip levels concatenated at 256-byte
// aligned offsets, row-major, tightly packed. Uniform and storage buffers are plain device allocations; kernels read
// them through pointers, never by value, so an unchanged pass list can be replayed as a CUDA graph while the host
// only rewrites buffer contents (the global UBO, resolve weights, ... land through one staged copy per frame).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include "image_view.h"
#include "plain_b200.h"
#include "plain_frame_types.h"

namespace pb {
using namespace pv;

struct BindlessEntry {  // set 2: global texture array (RenderBackend.cpp:45); index == image handle index
    ImgView view;
    uint32_t format;
    uint32_t pad;
    // R16F 3-D images (SDF bricks) only: the corner-replicated copy, (w + 1) x (h + 1) x (d + 1) entries of eight halves. Entry
    // (i, j, k) holds the 2x2x2 clamp-to-edge footprint of a trilinear tap whose lower texel is (i - 1, j - 1, k - 1), in the
    // sampler's blend order, so a tap is ONE 128-bit load. Built by the backend whenever the image is written (backend.cu).
    const uint4* corners;
};

struct MipInfo { int w, h, d; size_t offset, bytes; };
struct DeviceImage {
    plain_image_desc desc{};
    unsigned char* ptr = nullptr;
    size_t bytes = 0;
    std::vector<MipInfo> mips;
    bool inUse = false;
    bool transparentTexels = false;  // RGBA8 created with initial data containing a texel of alpha < 255 (alpha test of the raster passes)
    long long lastUsedSubmission = -1;  // index of the last submission (render_frame) whose passes referenced the image
    uint4* corners = nullptr;           // R16F 3-D images: corner-replicated copy for the sphere tracer (BindlessEntry::corners)
    bool cornersStale = false;          // the image was written since the copy was built
    unsigned char* peerPtr[PLAIN_MAX_PEERS] = {};  // the other ranks' copies of this image (CUDA IPC mappings), row sharding
    cudaEvent_t downloadDone = nullptr; // recorded after the last asynchronous read-back of the image (created on first use)
    bool downloadPending = false;       // that read-back has not been ordered before a later writer yet
    bool deferredExchange = false;      // peers push rows of this image behind the frame (peer_push_rows_deferred): a read-back waits for them
};
struct DeviceBuffer {
    unsigned char* ptr = nullptr;
    size_t size = 0;
};

// ---- graphic passes: meshes, draws and the rasteriser's scratch ----
struct DeviceMesh {  // MeshBinary (MeshData.h:27-35) in HBM
    unsigned char* indices = nullptr;   // u16 when indexCount < 65535, else u32 (RenderBackend.cpp:483-488)
    unsigned char* vertices = nullptr;  // 28 bytes per vertex
    uint32_t indexCount = 0, vertexCount = 0, index32 = 0;
};
struct RasterDraw {  // one entry of the device draw table of a graphic pass execution
    const unsigned char* indices;
    const unsigned char* vertices;
    uint32_t firstPrimitive, triCount, index32, firstVertex, vertexCount, alphaTest;  // alphaTest: the albedo texture (push[0]) has transparent texels
    uint32_t push[4];
};
#define PLAIN_RASTER_VERTEX_FLOATS 16  // one entry of a pass's post-transform vertex cache (64 bytes)
struct DrawRecord { uint32_t mesh; uint8_t push[16]; };

struct ExecRecord {
    uint32_t pass;
    std::vector<plain_render_target> targets;  // graphic passes: attachments in order
    std::vector<DrawRecord> draws;
    // resolved by render_frame before the passes run (allocations are not allowed while a graph is being captured)
    const RasterDraw* rasterDraws = nullptr;
    uint32_t* rasterTriInfo = nullptr;          // per primitive: covered rows (y0 | y1 << 16) or ~0u; [totalTris] = big-triangle count; then the big-triangle list
    unsigned long long* rasterVis = nullptr;    // per pixel of the depth target: depth bits << 32 | primitive + 1
    float* rasterVertexCache = nullptr;         // per (draw, vertex): the vertex stage's outputs, PLAIN_RASTER_VERTEX_FLOATS floats each
    uint32_t rasterTotalTris = 0, rasterTotalVertices = 0;
    bool rasterAnyAlphaTest = false;            // a draw's albedo texture has transparent texels: the alpha-testing kernel variants run
    std::vector<plain_storage_buffer_resource> storageBuffers;
    std::vector<plain_uniform_buffer_resource> uniformBuffers;
    std::vector<plain_image_resource> sampledImages;
    std::vector<plain_image_resource> storageImages;
    std::vector<uint8_t> pushConstants;
    uint32_t dispatch[3];
    uint32_t rowBegin = 0, rowEnd = 0, shardPhase = 0;  // row sharding (plain_compute_pass_execution), 0/0/0 = whole pass
    // pass fusion (backend.cu planFusions): a producer whose only consumer in the submission computes its texels inline
    int fusedProducer = -1;   // consumer: index of the execution it absorbed
    bool fusedAway = false;   // producer: no launch (its place in the dependency order is kept)
    std::vector<int> fusedRun; // the execution that launches a run of dependent passes as ONE kernel, the run's executions in order: bloom mips >= 2 (one persistent
                               // launch by the run's FIRST execution), the froxel chain (one launch over froxel columns by the run's LAST execution)
};

struct Backend;
struct LaunchCtx;
typedef void (*LaunchFn)(LaunchCtx&);

struct PassRecord {
    std::string shader, name;
    std::map<uint32_t, std::vector<uint8_t>> spec;
    LaunchFn fn = nullptr;
    // graphic passes (GraphicPassDescription, ResourceDescriptions.h:129-143); spec holds the vertex stage's constants
    bool graphic = false;
    uint32_t cullMode = 0, clampDepth = 0, depthFunction = 0, depthWrite = 0, pushSize = 0, depthAttachment = 0;
    std::vector<plain_attachment> attachments;
};

struct LaunchCtx {
    Backend* be;
    const PassRecord* pass;
    const ExecRecord* exec;
    cudaStream_t stream;
    const plain_global_shader_info* g;  // device pointer (set 0 binding 0)
    const BindlessEntry* bindless;      // device pointer
    const float* tables;                // device pointer: exact lookup tables over 8-bit domains, see ShadingTables
    int smCount;
    bool failed = false;
    std::string error;

    ImgView target(uint32_t attachment, int expectFormat = -1);  // graphic passes
    ImgView sampled(uint32_t binding, int expectFormat = -1);
    ImgView storage(uint32_t binding, int expectFormat = -1);
    int sampledFormat(uint32_t binding);  // plain_image_format of the image bound at a sampled binding, -1 if none
    template <typename T> T* sbuf(uint32_t binding, size_t* size = nullptr) { return (T*)sbufRaw(binding, size); }
    template <typename T> const T* ubuf(uint32_t binding) { return (const T*)ubufRaw(binding); }
    void* sbufRaw(uint32_t binding, size_t* size);
    const void* ubufRaw(uint32_t binding);
    template <typename T> T spec(uint32_t location, T def) const {
        auto it = pass->spec.find(location);
        if (it == pass->spec.end() || it->second.size() < sizeof(T)) return def;
        T v;
        memcpy(&v, it->second.data(), sizeof(T));
        return v;
    }
    bool specBool(uint32_t location, bool def) const {  // VkBool32 or 1-byte bool
        auto it = pass->spec.find(location);
        if (it == pass->spec.end() || it->second.empty()) return def;
        for (uint8_t b : it->second) if (b) return true;
        return false;
    }
    template <typename T> T push(size_t offset = 0) const {
        T v{};
        if (exec->pushConstants.size() >= offset + sizeof(T)) memcpy(&v, exec->pushConstants.data() + offset, sizeof(T));
        return v;
    }
    void fail(const std::string& msg) { if (!failed) { failed = true; error = msg; } }
    // rows [y0, y1) of an output with `rows` rows that this execution has to produce (row sharding; whole range by default)
    void window(int rows, int& y0, int& y1) const {
        if (exec->rowBegin == 0 && exec->rowEnd == 0) { y0 = 0; y1 = rows; return; }
        y0 = std::min((int)exec->rowBegin, rows);
        y1 = std::min((int)exec->rowEnd, rows);
        if (y1 < y0) y1 = y0;
    }
    void countLaunch(int n = 1);
    const ExecRecord* be_exec(int index) const;  // another execution of the same submission (pass fusion)
    unsigned int* be_fusionCounters() const;      // 64 u32 of device scratch for the grid barriers of a fused run
    const PassRecord* be_pass(uint32_t pass) const;
    size_t be_imageCount() const;               // image table of the backend (bindless slot == image handle index)
    int be_imageFormat(uint32_t index) const;
};

// Exact tables over 8-bit input domains (built once per context by buildShadingTables, passes_shading.cu): tabulating a
// function on its whole domain returns the same bits as evaluating it.
//   unorm8[b]       = float(b) / 255                              (UNORM8 decode)
//   srgbToLinear[b] = sRGBToLinear(unorm8[b])                     (colorConversion.inc:15-23 on an 8-bit albedo channel)
//   pcf[b * 12 + i] = {cos(angle), sin(angle), sqrt(d), 0} of tap i of calcShadow for blue-noise byte b (triangle.frag:107-116)
//   disc[s * 32 + i] = {sqrt(rand), cos(angle), sin(angle), 0} of sample i of the spatial GI filter's xorshift disc for seed
//                     wang_hash(s), s = frameIndexMod4 + filterIndex in 0..7 (filterIndirectDiffuseSpatial.comp:60-70): the sequence is
//                     the same for every pixel of a dispatch, so it is tabulated once per context instead of once per block
#define PLAIN_DISC_SEEDS 8
struct ShadingTables {
    float unorm8[256];
    float srgbToLinear[256];
    float4 pcf[256 * 12];
    float4 disc[PLAIN_DISC_SEEDS * 32];
};
void buildShadingTables(ShadingTables* deviceTables, cudaStream_t stream);
bool runDeviceSelftest(cudaStream_t stream, unsigned long long* hostOut8, std::string& error);  // selftest.cu
void buildCornerBrick(uint4* corners, const unsigned char* texels, int w, int h, int d, cudaStream_t stream);  // passes_gi.cu

struct PassRegistration { PassRegistration(const char* shader, LaunchFn fn); };
#define PLAIN_PASS(fnname, shader)                       \
    static void fnname(pb::LaunchCtx& c);                \
    static pb::PassRegistration reg_##fnname(shader, fnname); \
    static void fnname(pb::LaunchCtx& c)

inline unsigned ceilDiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

// launch helper: counts the launch (bench "gpu_launches") and records launch errors into the context
#define PLAIN_LAUNCH(c, kernel, grid, block, smem, ...)                                                   \
    do {                                                                                                  \
        kernel<<<(grid), (block), (smem), (c).stream>>>(__VA_ARGS__);                                     \
        cudaError_t e__ = cudaPeekAtLastError();                                                          \
        if (e__ != cudaSuccess) (c).fail(std::string(#kernel) + ": " + cudaGetErrorString(e__));          \
        (c).countLaunch();                                                                                \
    } while (0)

// ---- device helpers shared by the pass kernels ----
struct Globals {  // the fields of the global UBO (global.inc:4-33) most kernels need, loaded once per thread
    vec3 camPos, fwd, up, right;
    float tanFovHalf, aspect, nearPlane, farPlane;
};
PV_HD Globals loadGlobals(const plain_global_shader_info* __restrict__ g) {
    Globals r;
    r.camPos = v3(g->cameraPosition[0], g->cameraPosition[1], g->cameraPosition[2]);
    r.fwd = v3(g->cameraForward[0], g->cameraForward[1], g->cameraForward[2]);
    r.up = v3(g->cameraUp[0], g->cameraUp[1], g->cameraUp[2]);
    r.right = v3(g->cameraRight[0], g->cameraRight[1], g->cameraRight[2]);
    r.tanFovHalf = g->cameraTanFovHalf;
    r.aspect = g->cameraAspectRatio;
    r.nearPlane = g->nearPlane;
    r.farPlane = g->farPlane;
    return r;
}

}  // namespace pb
