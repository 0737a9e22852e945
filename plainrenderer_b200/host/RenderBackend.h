// RenderBackend.h - host-side C++ view of the C-ABI in include/plain_b200.h, with the method names and argument
// meaning of the reference's `class RenderBackend` (Plain/src/Runtime/Rendering/Backend/RenderBackend.h:33-110), so
// the frontend/technique mirrors below read like the reference callers. Every method is one C-ABI call; failures
// throw std::runtime_error carrying plain_last_error() (the reference prints and throws, RenderBackend.cpp:442-445).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "plain_b200.h"

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
};
struct SpecialisationConstant { uint32_t location; std::vector<char> data; };
struct ShaderDescription { std::string srcPathRelative; std::vector<SpecialisationConstant> specialisationConstants; };
struct ComputePassDescription { ShaderDescription shaderDescription; std::string name; };

inline std::vector<char> dataToCharArray(const void* data, size_t size) {  // GeneralUtils.cpp:8-12
    const char* p = (const char*)data;
    return std::vector<char>(p, p + size);
}

class RenderBackend {
public:
    void setup(int device, uint32_t width, uint32_t height) { if (PLAIN_FN(backend_create)(device, width, height, &m_ctx)) throw std::runtime_error("backend_create failed"); }
    void shutdown() { if (m_ctx) PLAIN_FN(backend_destroy)(m_ctx); m_ctx = nullptr; }
    plain_ctx* context() const { return m_ctx; }

    void resizeImages(const std::vector<ImageHandle>& images, uint32_t w, uint32_t h) { check(PLAIN_FN(resize_images)(m_ctx, images.data(), (uint32_t)images.size(), w, h)); }
    void newFrame() { check(PLAIN_FN(new_frame)(m_ctx)); }
    void setComputePassExecution(const ComputePassExecution& e) {
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
    void renderFrame(bool present) { check(PLAIN_FN(render_frame)(m_ctx, present ? 1 : 0)); }
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
};

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
