// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// backend.h - CPU implementation of include/plain_b200.h (symbols oracle_*): handle tables, deferred buffer
// fills, in-order replay of the recorded compute-pass executions (RenderBackend.cpp:259-265, 769-786,
// 896-911). Each pass is a scalar C++ restatement of one reference shader, looked up by shader file name.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "image.h"
#include "../include/plain_frame_types.h"

namespace orc {

struct Buffer {
    std::vector<uint8_t> data;
};

struct Mesh {  // MeshBinary (MeshData.h:27-35) as uploaded by create_meshes
    uint32_t indexCount = 0, vertexCount = 0;
    bool index32 = false;
    std::vector<uint8_t> indices, vertices;
};
struct DrawRecord {  // one mesh of a drawMeshes call with its push-constant block
    uint32_t mesh;
    uint8_t push[16];
};

struct ExecRecord {
    uint32_t pass;
    std::vector<plain_render_target> targets;  // graphic passes: attachments in order
    std::vector<DrawRecord> draws;
    std::vector<plain_storage_buffer_resource> storageBuffers;
    std::vector<plain_uniform_buffer_resource> uniformBuffers;
    std::vector<plain_image_resource> sampledImages;
    std::vector<plain_image_resource> storageImages;
    std::vector<uint8_t> pushConstants;
    uint32_t dispatch[3];
};

struct PassCtx;
typedef void (*PassFn)(PassCtx&);

struct PassRecord {
    std::string shader;
    std::string name;
    std::map<uint32_t, std::vector<uint8_t>> spec;
    PassFn fn = nullptr;
    // graphic passes (GraphicPassDescription, ResourceDescriptions.h:129-143); spec holds the vertex stage's constants
    bool graphic = false;
    uint32_t cullMode = 0, clampDepth = 0, depthFunction = 0, depthWrite = 0, pushSize = 0;
    std::vector<plain_attachment> attachments;
};

struct FillOrder {
    bool uniform;
    uint32_t buffer;
    std::vector<uint8_t> data;
};

struct Ctx {
    std::vector<Image> images;
    std::vector<Image> transientImages;
    Image swapchain;
    std::vector<Buffer> uniformBuffers;
    std::vector<Buffer> storageBuffers;
    std::vector<plain_sampler_desc> samplers;
    std::vector<PassRecord> passes;
    std::vector<Mesh> meshes;
    std::map<uint32_t, std::vector<uint64_t>> visibility;  // per depth image (handle index): depth bits << 32 | primitive + 1 of the last raster pass
    std::vector<ExecRecord> execs;
    std::vector<FillOrder> fills;
    uint32_t globalUniformBuffer = PLAIN_INVALID_INDEX;
    std::string lastError;
    std::vector<plain_pass_time> timings;
    bool timingEnabled = false;
    int threads = 1;

    Image* resolve(plain_image_handle h);
};

void parallelFor(int threads, int n, const std::function<void(int)>& fn);

struct PassCtx {
    Ctx* ctx;
    const PassRecord* pass;
    const ExecRecord* exec;
    plain_global_shader_info g;

    View target(uint32_t attachment) const;  // graphic passes
    View sampled(uint32_t binding) const;
    View storage(uint32_t binding) const;
    View bindless(uint32_t index) const;  // set 2: global texture array, index == image handle index
    uint8_t* sbuf(uint32_t binding, size_t* size = nullptr) const;
    const uint8_t* ubuf(uint32_t binding, size_t* size = nullptr) const;
    template <typename T> T spec(uint32_t location, T def) const {
        auto it = pass->spec.find(location);
        if (it == pass->spec.end() || it->second.size() < sizeof(T)) return def;
        T v; memcpy(&v, it->second.data(), sizeof(T)); return v;
    }
    bool specBool(uint32_t location, bool def) const {  // VkBool32 or 1-byte C++ bool
        auto it = pass->spec.find(location);
        if (it == pass->spec.end() || it->second.empty()) return def;
        for (uint8_t b : it->second) if (b) return true;
        return false;
    }
    template <typename T> T push(size_t offset = 0) const {
        T v{}; if (exec->pushConstants.size() >= offset + sizeof(T)) memcpy(&v, exec->pushConstants.data() + offset, sizeof(T)); return v;
    }
    // run fn(groupX, groupY, groupZ) for every workgroup of the dispatch, rows of groups spread over threads
    void forEachGroup(const std::function<void(int, int, int)>& fn) const;
    // run fn(x, y, z) for every invocation (gl_GlobalInvocationID) with the given local size
    void forEachInvocation(int lx, int ly, int lz, const std::function<void(int, int, int)>& fn) const;

    vec3 gv3(const float* p) const { return vec3(p[0], p[1], p[2]); }
    mat4 gm4(const float* p) const { mat4 m; for (int c = 0; c < 4; c++) m.c[c] = vec4(p[c * 4], p[c * 4 + 1], p[c * 4 + 2], p[c * 4 + 3]); return m; }
};

PassFn findPass(const std::string& shader);
struct PassRegistration { PassRegistration(const char* shader, PassFn fn); };
// a pass that takes precedence over the registered one of the same shader name: oracle/_ref/liboracle_refmain.so (oracle/ref/ref_shader_passes.cpp)
// runs the reference's own main() for the shaders that compile as C++; liboracle.so itself registers no override
struct PassOverride { PassOverride(const char* shader, PassFn fn); };
#define ORACLE_PASS(fnname, shader) static void fnname(PassCtx& c); static PassRegistration reg_##fnname(shader, fnname); static void fnname(PassCtx& c)

}  // namespace orc
