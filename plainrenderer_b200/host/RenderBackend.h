// RenderBackend.h - host-side C++ view of the C-ABI in include/plain_b200.h, with the method names and argument
// meaning of the reference's `class RenderBackend` (Plain/src/Runtime/Rendering/Backend/RenderBackend.h:33-110), so
// the frontend/technique mirrors below read like the reference callers. Every method is one C-ABI call; failures
// throw std::runtime_error carrying plain_last_error() (the reference prints and throws, RenderBackend.cpp:442-445).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "plain_b200.h"
#include "plain_frontend.h"

typedef plain_image_handle ImageHandle;
struct RenderPassHandle { uint32_t index = PLAIN_INVALID_INDEX; };
struct UniformBufferHandle { uint32_t index = PLAIN_INVALID_INDEX; };
struct StorageBufferHandle { uint32_t index = PLAIN_INVALID_INDEX; };
struct SamplerHandle { uint32_t index = PLAIN_INVALID_INDEX; };
typedef plain_image_desc ImageDescription;

// ResourceDescriptions.h:9-53
struct StorageBufferResource {
    StorageBufferResource(StorageBufferHandle b, bool ro, uint32_t bind) : buffer(b), readOnly(ro), binding(bind) {}
    StorageBufferHandle buffer; bool readOnly; uint32_t binding;
};
struct UniformBufferResource {
    UniformBufferResource(UniformBufferHandle b, uint32_t bind) : buffer(b), binding(bind) {}
    UniformBufferHandle buffer; uint32_t binding;
};
struct ImageResource {
    ImageResource(ImageHandle i, uint32_t mip, uint32_t bind) : image(i), mipLevel(mip), binding(bind) {}
    ImageHandle image; uint32_t mipLevel; uint32_t binding;
};
struct SamplerResource {
    SamplerResource(SamplerHandle s, uint32_t bind) : sampler(s), binding(bind) {}
    SamplerHandle sampler; uint32_t binding;
};
struct RenderPassResources {
    std::vector<SamplerResource> samplers;
    std::vector<StorageBufferResource> storageBuffers;
    std::vector<UniformBufferResource> uniformBuffers;
    std::vector<ImageResource> sampledImages;
    std::vector<ImageResource> storageImages;
};
struct RenderPassExecution { RenderPassHandle handle; RenderPassResources resources; };
struct ComputePassExecution {
    RenderPassExecution genericInfo;
    std::vector<char> pushConstants;
    uint32_t dispatchCount[3] = {1, 1, 1};
    uint32_t rowBegin = 0, rowEnd = 0, shardPhase = 0;  // row sharding (plain_compute_pass_execution), 0/0/0 = whole pass
};
// graphic passes (ResourceDescriptions.h:55-71, 80-143)
struct MeshHandle { uint32_t index = PLAIN_INVALID_INDEX; };
struct RenderTarget { ImageHandle image; uint32_t mipLevel; };
struct GraphicPassExecution { RenderPassExecution genericInfo; std::vector<RenderTarget> targets; uint32_t rowBegin = 0, rowEnd = 0; /* row sharding, 0/0 = all rows */ };
struct SpecialisationConstant { uint32_t location; std::vector<char> data; };
struct ShaderDescription { std::string srcPathRelative; std::vector<SpecialisationConstant> specialisationConstants; };
struct ComputePassDescription { ShaderDescription shaderDescription; std::string name; };
struct GraphicPassShaderDescriptions { ShaderDescription vertex, fragment; };
struct Attachment { plain_image_format format; plain_attachment_load_op loadOp; };
struct RasterizationConfig { plain_cull_mode cullMode = PLAIN_CULL_NONE; bool clampDepth = false; };
struct DepthTest { plain_depth_function function = PLAIN_DEPTH_ALWAYS; bool write = false; };
struct GraphicPassDescription {
    GraphicPassShaderDescriptions shaderDescriptions;
    std::vector<Attachment> attachments;
    RasterizationConfig rasterization;
    DepthTest depthTest;
    std::string name;
};
struct MeshBinary {  // MeshData.h:27-35 (geometry part)
    uint32_t indexCount = 0, vertexCount = 0;
    std::vector<uint16_t> indexBuffer;  // 16 or 32 bit indices, as stored
    std::vector<uint8_t> vertexBuffer;
};

inline std::vector<char> dataToCharArray(const void* data, size_t size) {  // GeneralUtils.cpp:8-12
    const char* p = (const char*)data;
    return std::vector<char>(p, p + size);
}

// what a row-sharded frame is made of: compute pass executions and the exchanges between them
struct ExchangeRequest {
    uint32_t kind = PLAIN_EXCHANGE_NONE;
    std::vector<ImageHandle> images;
    std::vector<uint32_t> mips, divisors;
    uint32_t buffer = PLAIN_INVALID_INDEX, elementCount = 0, haloRows = 0;
    bool deferred = false;  // only the NEXT frame reads the rows: the peer exchange may run behind the frame (peer_push_rows_deferred)
    std::string name;
};
struct ShardInfo {
    uint32_t rank = 0, count = 1, fullHeight = 0, y0 = 0, y1 = 0;  // [y0, y1): full-resolution rows of this rank
    bool active() const { return count > 1; }
};
#define PLAIN_SHARD_ROW_UNIT 32u
inline void shardBandRows(uint32_t fullHeight, uint32_t count, uint32_t rank, uint32_t* y0, uint32_t* y1) {
    // 32 rows: a 32x32 histogram tile, one 16-row block of the half-res trace, 4 fused HiZ levels (16 / 8 / 4 / 2 rows of the
    // half-res pyramid), 4 froxel rows. (The 32x32 culling tiles of the half-res trace span 64 rows: their lists are
    // computed by every rank.) Finer units balance the bands better: 2160 rows on 8 ranks = 256..288 rows instead of 256..320
    const uint32_t unit = PLAIN_SHARD_ROW_UNIT, units = (fullHeight + unit - 1) / unit;
    const uint32_t u0 = (uint32_t)((uint64_t)rank * units / count), u1 = (uint32_t)((uint64_t)(rank + 1) * units / count);
    *y0 = u0 * unit < fullHeight ? u0 * unit : fullHeight;
    *y1 = u1 * unit < fullHeight ? u1 * unit : fullHeight;
}

class RenderBackend {
public:
    struct Recorded {  // one entry of a recorded frame: a compute execution, an exchange or a graphic execution with its draws
        bool isExchange; ComputePassExecution exec; ExchangeRequest exchange;
        bool isGraphic = false; GraphicPassExecution graphic; std::vector<plain_handle> drawMeshes; std::vector<char> drawPush;
    };
    // ---- row sharding helpers (no counterpart in the reference) ----
    ShardInfo shard;
    // rows of an image whose height is fullHeight / divisor (or tiles of `divisor` rows) that belong to this rank, optionally
    // extended by `extend` rows on both sides (overlapped computation instead of a halo exchange); {0, 0} when not sharded
    void band(uint32_t divisor, uint32_t rows, uint32_t extend, uint32_t* a, uint32_t* b) const {
        if (!shard.active()) { *a = 0; *b = 0; return; }
        uint32_t lo = shard.y0 / divisor, hi = (shard.y1 + divisor - 1) / divisor;
        lo = lo > extend ? lo - extend : 0;
        hi = hi + extend < rows ? hi + extend : rows;
        if (lo > rows) lo = rows;
        *a = lo; *b = hi > lo ? hi : lo;
        if (*a == 0 && *b == 0) *b = 0;
    }

    void setup(int device, uint32_t width, uint32_t height) { if (PLAIN_FN(backend_create)(device, width, height, &m_ctx)) throw std::runtime_error("backend_create failed"); }
    void shutdown() { if (m_ctx) PLAIN_FN(backend_destroy)(m_ctx); m_ctx = nullptr; }
    plain_ctx* context() const { return m_ctx; }

    void resizeImages(const std::vector<ImageHandle>& images, uint32_t w, uint32_t h) { check(PLAIN_FN(resize_images)(m_ctx, images.data(), (uint32_t)images.size(), w, h)); }
    void newFrame() { check(PLAIN_FN(new_frame)(m_ctx)); }
    void setComputePassExecution(const ComputePassExecution& e) {
        if (m_recording) { m_recorded.push_back(Recorded{false, e, ExchangeRequest()}); return; }
        sendExecution(e);
    }
    // graphic passes: the execution is recorded where the frontend sets it; drawMeshes (called later, from renderScene) attaches
    // the draws to it; both reach the C-ABI when the segment is sent
    void setGraphicPassExecution(const GraphicPassExecution& e) {
        Recorded r{false, ComputePassExecution(), ExchangeRequest()};
        r.isGraphic = true;
        r.graphic = e;
        if (m_recording) { m_recorded.push_back(r); return; }
        m_immediateGraphic.push_back(r);
    }
    void drawMeshes(const std::vector<MeshHandle>& meshHandles, const char* pushConstantData, RenderPassHandle pass, int workerIndex) {
        (void)workerIndex;
        const uint32_t ps = pass.index < m_graphicPushSize.size() ? m_graphicPushSize[pass.index] : 0;
        if (!ps) throw std::runtime_error("drawMeshes: not a graphic pass");
        Recorded* rec = nullptr;
        for (auto& r : m_recorded) if (r.isGraphic && r.graphic.genericInfo.handle.index == pass.index) rec = &r;
        for (auto& r : m_immediateGraphic) if (r.graphic.genericInfo.handle.index == pass.index) rec = &r;
        if (!rec) throw std::runtime_error("drawMeshes: the pass has no execution this frame");
        for (auto& m : meshHandles) rec->drawMeshes.push_back(m.index);
        rec->drawPush.insert(rec->drawPush.end(), pushConstantData, pushConstantData + (size_t)ps * meshHandles.size());
    }
    std::vector<MeshHandle> createMeshes(const std::vector<MeshBinary>& meshes) {
        std::vector<plain_mesh_binary> in;
        for (auto& m : meshes) in.push_back(plain_mesh_binary{m.indexCount, m.vertexCount, m.indexBuffer.data(), m.vertexBuffer.data()});
        std::vector<plain_handle> out(meshes.size());
        check(PLAIN_FN(create_meshes)(m_ctx, in.data(), (uint32_t)in.size(), out.data()));
        std::vector<MeshHandle> handles(meshes.size());
        for (size_t i = 0; i < out.size(); i++) handles[i].index = out[i];
        return handles;
    }
    RenderPassHandle createGraphicPass(const GraphicPassDescription& d) {
        std::vector<plain_spec_const> vs = specs(d.shaderDescriptions.vertex), fs = specs(d.shaderDescriptions.fragment);
        std::vector<plain_attachment> at;
        for (auto& a : d.attachments) at.push_back(plain_attachment{(uint32_t)a.format, (uint32_t)a.loadOp});
        plain_graphic_pass_desc x{};
        x.vertex_shader = d.shaderDescriptions.vertex.srcPathRelative.c_str(); x.vertex_consts = vs.data(); x.n_vertex_consts = (uint32_t)vs.size();
        x.fragment_shader = d.shaderDescriptions.fragment.srcPathRelative.c_str(); x.fragment_consts = fs.data(); x.n_fragment_consts = (uint32_t)fs.size();
        x.attachments = at.data(); x.n_attachments = (uint32_t)at.size();
        x.cull_mode = d.rasterization.cullMode; x.clamp_depth = d.rasterization.clampDepth ? 1u : 0u;
        x.depth_function = d.depthTest.function; x.depth_write = d.depthTest.write ? 1u : 0u;
        x.debug_name = d.name.c_str();
        RenderPassHandle h;
        check(PLAIN_FN(create_graphic_pass)(m_ctx, &x, &h.index));
        if (m_graphicPushSize.size() <= h.index) m_graphicPushSize.resize(h.index + 1, 0);
        m_graphicPushSize[h.index] = d.shaderDescriptions.vertex.srcPathRelative == "sunShadow.vert" ? 8u : 16u;
        return h;
    }
    void addExchange(const ExchangeRequest& x) { if (m_recording && shard.active()) m_recorded.push_back(Recorded{true, ComputePassExecution(), x}); }
    // deferred frames: executions are kept on the host and sent segment by segment (run until the next exchange)
    void beginRecording() { m_recording = true; m_recorded.clear(); m_pendingDeferred.clear(); m_cursor = 0; }
    void endRecording() { m_recording = false; }
    bool hasRecorded() const { return m_cursor < m_recorded.size(); }
    // sends executions up to the next exchange and submits them; returns true and fills `out` when an exchange is pending
    bool runSegment(plain_exchange* out) {
        while (m_cursor < m_recorded.size()) {
            const Recorded& r = m_recorded[m_cursor++];
            if (r.isGraphic) { sendGraphic(r); continue; }
            if (!r.isExchange) { sendExecution(r.exec); continue; }
            // a deferred exchange over peer memory does not end the submission: its pushes are issued behind the submission it falls
            // into (they only have to land before the next frame), so the passes around it stay in one graph with its parallel branches
            if (r.exchange.deferred && deferredOnDevice(r.exchange)) { m_pendingDeferred.push_back(&r.exchange); continue; }
            check(PLAIN_FN(submit_recorded_passes)(m_ctx));
            issuePendingDeferred();
            // two exchanges back to back share the second one's barrier (pushes enqueued before a barrier on any rank are visible after it)
            bool nextHasBarrier = false;
            if (m_cursor < m_recorded.size() && m_recorded[m_cursor].isExchange && !m_recorded[m_cursor].exchange.deferred) {
                const ExchangeRequest& nx = m_recorded[m_cursor].exchange;
                nextHasBarrier = m_peerExchange && shard.active() && (nx.kind == PLAIN_EXCHANGE_ALLREDUCE_SUM_U32 || imagesMapped(nx));
            }
            if (peerExchange(r.exchange, nextHasBarrier)) continue;  // done on the device: rows pushed into the peers' images + flag barrier
            std::memset(out, 0, sizeof(*out));
            out->kind = r.exchange.kind;
            out->halo_rows = r.exchange.haloRows;
            out->element_count = r.exchange.elementCount;
            out->deferred = r.exchange.deferred ? 1u : 0u;
            std::snprintf(out->name, sizeof(out->name), "%s", r.exchange.name.c_str());
            if (r.exchange.kind == PLAIN_EXCHANGE_ALLREDUCE_SUM_U32) {
                size_t size = 0;
                check(PLAIN_FN(get_storage_buffer_device_pointer)(m_ctx, r.exchange.buffer, &out->device_ptr[0], &size));
                out->n_images = 1;
                out->buffer = r.exchange.buffer;
            } else {
                out->n_images = (uint32_t)r.exchange.images.size();
                for (uint32_t i = 0; i < out->n_images && i < 4; i++) {
                    size_t size = 0;
                    check(PLAIN_FN(get_image_device_pointer)(m_ctx, r.exchange.images[i], r.exchange.mips[i], &out->device_ptr[i], &size));
                    ImageDescription d = getImageDescription(r.exchange.images[i]);
                    uint32_t rows = d.height >> r.exchange.mips[i];
                    if (rows < 1) rows = 1;
                    uint32_t slices = d.depth >> r.exchange.mips[i];
                    if (slices < 1) slices = 1;
                    out->rows[i] = rows;
                    out->depth[i] = slices;
                    out->row_pitch_bytes[i] = (uint32_t)(size / ((size_t)rows * slices));
                    out->row_divisor[i] = r.exchange.divisors[i];
                    out->image[i].index = r.exchange.images[i].index;
                    out->image[i].type = (uint32_t)r.exchange.images[i].type;
                    out->mip_level[i] = r.exchange.mips[i];
                }
            }
            return true;
        }
        check(PLAIN_FN(render_frame)(m_ctx, 1));
        issuePendingDeferred();
        if (m_deferredIssued) { check(PLAIN_FN(peer_flush_deferred)(m_ctx)); m_deferredIssued = false; }  // one barrier for the frame's deferred pushes
        return false;
    }
    // ---- peer exchange over NVLink (include/plain_b200.h): the sends of sharding.plan_row_exchange as row pushes ----
    bool m_peerExchange = false;
    bool m_peerDeferred = !(std::getenv("PLAIN_PEER_DEFERRED") && std::atoi(std::getenv("PLAIN_PEER_DEFERRED")) == 0);  // A / B switch
    bool m_deferredIssued = false;
    void bandOf(uint32_t rank, uint32_t divisor, uint32_t rows, uint32_t* a, uint32_t* b) const {
        uint32_t y0 = 0, y1 = shard.fullHeight;
        shardBandRows(shard.fullHeight, shard.count, rank, &y0, &y1);
        if (divisor < 1) divisor = 1;
        uint32_t lo = y0 / divisor, hi = (y1 + divisor - 1) / divisor;
        if (hi > rows) hi = rows;
        if (lo > hi) lo = hi;
        *a = lo; *b = hi;
    }
    std::vector<const ExchangeRequest*> m_pendingDeferred;
    bool imagesMapped(const ExchangeRequest& x) {
        for (size_t i = 0; i < x.images.size(); i++) {
            plain_image_handle h;
            h.type = (uint32_t)x.images[i].type;
            h.index = x.images[i].index;
            if (!PLAIN_FN(peer_image_ready)(m_ctx, h)) return false;
        }
        return true;
    }
    bool deferredOnDevice(const ExchangeRequest& x) {
        return m_peerExchange && m_peerDeferred && shard.active() && x.kind == PLAIN_EXCHANGE_ALLGATHER_ROWS && imagesMapped(x);
    }
    void issuePendingDeferred() {
        for (const ExchangeRequest* x : m_pendingDeferred)
            if (!peerExchange(*x)) throw std::runtime_error("deferred peer exchange: image not mapped");
        m_pendingDeferred.clear();
    }
    bool peerExchange(const ExchangeRequest& x, bool skipBarrier = false) {
        if (!m_peerExchange || !shard.active()) return false;
        if (x.kind == PLAIN_EXCHANGE_ALLREDUCE_SUM_U32) {
            check(PLAIN_FN(peer_allreduce_sum_u32)(m_ctx, x.buffer, x.elementCount));
            return true;
        }
        std::vector<plain_peer_push> pushes;
        for (size_t i = 0; i < x.images.size(); i++) {
            plain_image_handle h;
            h.type = (uint32_t)x.images[i].type;
            h.index = x.images[i].index;
            if (!PLAIN_FN(peer_image_ready)(m_ctx, h)) return false;
            const ImageDescription d = getImageDescription(x.images[i]);
            uint32_t rows = d.height >> x.mips[i];
            if (rows < 1) rows = 1;
            uint32_t a, b;
            bandOf(shard.rank, x.divisors[i], rows, &a, &b);
            auto push = [&](uint32_t peer, uint32_t r0, uint32_t r1) {
                if (r1 > r0) pushes.push_back(plain_peer_push{h, x.mips[i], r0, r1, peer});
            };
            if (x.kind == PLAIN_EXCHANGE_ALLGATHER_ROWS) {
                for (uint32_t p = 0; p < shard.count; p++)
                    if (p != shard.rank) push(p, a, b);
            } else if (x.kind == PLAIN_EXCHANGE_HALO_ROWS) {
                const uint32_t halo = x.haloRows;
                if (shard.rank > 0) push(shard.rank - 1, a, a + halo < b ? a + halo : b);                // the neighbour above needs my first rows
                if (shard.rank + 1 < shard.count) push(shard.rank + 1, b > a + halo ? b - halo : a, b);  // the neighbour below my last rows
            }
        }
        if (x.deferred && m_peerDeferred && x.kind == PLAIN_EXCHANGE_ALLGATHER_ROWS) {
            check(PLAIN_FN(peer_push_rows_deferred)(m_ctx, (uint32_t)pushes.size(), pushes.data()));
            m_deferredIssued = true;
            return true;
        }
        check(PLAIN_FN(peer_push_rows)(m_ctx, (uint32_t)pushes.size(), pushes.data()));
        if (!skipBarrier) check(PLAIN_FN(peer_barrier)(m_ctx));
        return true;
    }
    void sendGraphic(const Recorded& r);
    void sendExecution(const ComputePassExecution& e) {
        std::vector<plain_sampler_resource> sm;
        std::vector<plain_storage_buffer_resource> sb;
        std::vector<plain_uniform_buffer_resource> ub;
        std::vector<plain_image_resource> si, st;
        plain_compute_pass_execution x{};
        fill(e.genericInfo.resources, x.resources, sm, sb, ub, si, st);
        x.pass = e.genericInfo.handle.index;
        x.push_constants = e.pushConstants.empty() ? nullptr : e.pushConstants.data();
        x.push_constant_size = (uint32_t)e.pushConstants.size();
        for (int i = 0; i < 3; i++) x.dispatch_count[i] = e.dispatchCount[i];
        x.row_begin = e.rowBegin; x.row_end = e.rowEnd; x.shard_phase = e.shardPhase;
        check(PLAIN_FN(set_compute_pass_execution)(m_ctx, &x));
    }
    void prepareForDrawcallRecording() { check(PLAIN_FN(prepare_for_drawcall_recording)(m_ctx)); }
    void setUniformBufferData(UniformBufferHandle b, const void* data, size_t size) { check(PLAIN_FN(set_uniform_buffer_data)(m_ctx, b.index, data, size)); }
    void setStorageBufferData(StorageBufferHandle b, const void* data, size_t size) { check(PLAIN_FN(set_storage_buffer_data)(m_ctx, b.index, data, size)); }
    void setGlobalDescriptorSetResources(const RenderPassResources& r) {
        std::vector<plain_sampler_resource> sm;
        std::vector<plain_storage_buffer_resource> sb;
        std::vector<plain_uniform_buffer_resource> ub;
        std::vector<plain_image_resource> si, st;
        plain_pass_resources x{};
        fill(r, x, sm, sb, ub, si, st);
        check(PLAIN_FN(set_global_descriptor_set_resources)(m_ctx, &x));
    }
    void updateComputePassShaderDescription(RenderPassHandle pass, const ShaderDescription& d) {
        std::vector<plain_spec_const> sc = specs(d);
        check(PLAIN_FN(update_compute_pass_shader_description)(m_ctx, pass.index, d.srcPathRelative.c_str(), sc.data(), (uint32_t)sc.size()));
    }
    void renderFrame(bool present) {
        for (auto& r : m_immediateGraphic) sendGraphic(r);
        m_immediateGraphic.clear();
        check(PLAIN_FN(render_frame)(m_ctx, present ? 1 : 0));
    }
    uint32_t getImageGlobalTextureArrayIndex(ImageHandle image) { uint32_t i = 0; check(PLAIN_FN(get_image_global_texture_array_index)(m_ctx, image, &i)); return i; }
    RenderPassHandle createComputePass(const ComputePassDescription& d) {
        std::vector<plain_spec_const> sc = specs(d.shaderDescription);
        RenderPassHandle h;
        check(PLAIN_FN(create_compute_pass)(m_ctx, d.shaderDescription.srcPathRelative.c_str(), sc.data(), (uint32_t)sc.size(), d.name.c_str(), &h.index));
        return h;
    }
    ImageHandle createImage(const ImageDescription& d, const void* data, size_t size) { ImageHandle h; check(PLAIN_FN(create_image)(m_ctx, &d, data, size, &h)); return h; }
    UniformBufferHandle createUniformBuffer(size_t size, const void* initial = nullptr) { UniformBufferHandle h; check(PLAIN_FN(create_uniform_buffer)(m_ctx, size, initial, &h.index)); return h; }
    StorageBufferHandle createStorageBuffer(size_t size, const void* initial = nullptr) { StorageBufferHandle h; check(PLAIN_FN(create_storage_buffer)(m_ctx, size, initial, &h.index)); return h; }
    SamplerHandle createSampler(const plain_sampler_desc& d) { SamplerHandle h; check(PLAIN_FN(create_sampler)(m_ctx, &d, &h.index)); return h; }
    ImageHandle createTemporaryImage(const ImageDescription& d) { ImageHandle h; check(PLAIN_FN(create_temporary_image)(m_ctx, &d, &h)); return h; }
    ImageHandle getSwapchainInputImage() { ImageHandle h; check(PLAIN_FN(get_swapchain_input_image)(m_ctx, &h)); return h; }
    ImageDescription getImageDescription(ImageHandle h) { ImageDescription d; check(PLAIN_FN(get_image_description)(m_ctx, h, &d)); return d; }
    // additions of this build (inputs the out-of-scope raster passes would have written, read-back)
    void writeImage(ImageHandle h, uint32_t mip, const void* data, size_t size) { check(PLAIN_FN(write_image)(m_ctx, h, mip, data, size)); }
    void writeImageAsync(ImageHandle h, uint32_t mip, const void* data, size_t size) { check(PLAIN_FN(write_image_async)(m_ctx, h, mip, data, size)); }
    void writeImageRowsAsync(ImageHandle h, uint32_t mip, uint32_t r0, uint32_t r1, const void* data, size_t size) { check(PLAIN_FN(write_image_rows_async)(m_ctx, h, mip, r0, r1, data, size)); }
    void readImageRowsAsync(ImageHandle h, uint32_t mip, uint32_t r0, uint32_t r1, void* out, size_t size) { check(PLAIN_FN(read_image_rows_async)(m_ctx, h, mip, r0, r1, out, size)); }
    void readImage(ImageHandle h, uint32_t mip, void* out, size_t size) { check(PLAIN_FN(read_image)(m_ctx, h, mip, out, size)); }
    void readImageAsync(ImageHandle h, uint32_t mip, void* out, size_t size) { check(PLAIN_FN(read_image_async)(m_ctx, h, mip, out, size)); }
    void waitForGPUIdle() { check(PLAIN_FN(wait_for_gpu_idle)(m_ctx)); }

private:
    void check(int rc) { if (rc) throw std::runtime_error(std::string("RenderBackend: ") + PLAIN_FN(last_error)(m_ctx)); }
    static std::vector<plain_spec_const> specs(const ShaderDescription& d) {
        std::vector<plain_spec_const> sc;
        for (auto& c : d.specialisationConstants) sc.push_back({c.location, c.data.data(), (uint32_t)c.data.size()});
        return sc;
    }
    static void fill(const RenderPassResources& r, plain_pass_resources& x, std::vector<plain_sampler_resource>& sm, std::vector<plain_storage_buffer_resource>& sb,
                     std::vector<plain_uniform_buffer_resource>& ub, std::vector<plain_image_resource>& si, std::vector<plain_image_resource>& st) {
        for (auto& s : r.samplers) sm.push_back({s.sampler.index, s.binding});
        for (auto& s : r.storageBuffers) sb.push_back({s.buffer.index, s.readOnly ? 1u : 0u, s.binding});
        for (auto& s : r.uniformBuffers) ub.push_back({s.buffer.index, s.binding});
        for (auto& s : r.sampledImages) si.push_back({s.image, s.mipLevel, s.binding});
        for (auto& s : r.storageImages) st.push_back({s.image, s.mipLevel, s.binding});
        x.samplers = sm.data(); x.n_samplers = (uint32_t)sm.size();
        x.storage_buffers = sb.data(); x.n_storage_buffers = (uint32_t)sb.size();
        x.uniform_buffers = ub.data(); x.n_uniform_buffers = (uint32_t)ub.size();
        x.sampled_images = si.data(); x.n_sampled_images = (uint32_t)si.size();
        x.storage_images = st.data(); x.n_storage_images = (uint32_t)st.size();
    }
    plain_ctx* m_ctx = nullptr;
    std::vector<Recorded> m_recorded, m_immediateGraphic;
    std::vector<uint32_t> m_graphicPushSize;  // by pass handle; 0 = not a graphic pass
    size_t m_cursor = 0;
    bool m_recording = false;
};

inline void RenderBackend::sendGraphic(const Recorded& r) {
    std::vector<plain_sampler_resource> sm;
    std::vector<plain_storage_buffer_resource> sb;
    std::vector<plain_uniform_buffer_resource> ub;
    std::vector<plain_image_resource> si, st;
    plain_graphic_pass_execution x{};
    fill(r.graphic.genericInfo.resources, x.resources, sm, sb, ub, si, st);
    x.pass = r.graphic.genericInfo.handle.index;
    std::vector<plain_render_target> tg;
    for (auto& t : r.graphic.targets) tg.push_back(plain_render_target{t.image, t.mipLevel});
    x.targets = tg.data(); x.n_targets = (uint32_t)tg.size();
    x.row_begin = r.graphic.rowBegin; x.row_end = r.graphic.rowEnd;
    check(PLAIN_FN(set_graphic_pass_execution)(m_ctx, &x));
    if (!r.drawMeshes.empty()) check(PLAIN_FN(draw_meshes)(m_ctx, r.drawMeshes.data(), (uint32_t)r.drawMeshes.size(), r.drawPush.data(), x.pass, 0));
}

inline ImageDescription imageDesc2D(uint32_t w, uint32_t h, plain_image_format f, uint32_t usage, plain_mip_count mips = PLAIN_MIPS_ONE, uint32_t manualMips = 1) {
    ImageDescription d{};
    d.width = w; d.height = h; d.depth = 1;
    d.type = PLAIN_IMAGE_TYPE_2D; d.format = f; d.usage_flags = usage; d.mip_count = mips; d.manual_mip_count = manualMips; d.auto_create_mips = 0;
    return d;
}
inline ImageDescription imageDesc3D(uint32_t w, uint32_t h, uint32_t dep, plain_image_format f, uint32_t usage) {
    ImageDescription d = imageDesc2D(w, h, f, usage);
    d.depth = dep; d.type = PLAIN_IMAGE_TYPE_3D;
    return d;
}
