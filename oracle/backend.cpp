// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// backend.cpp - resource tables and frame replay of the CPU oracle backend (include/plain_b200.h, oracle_* symbols).
#include "backend.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

struct plain_ctx {
    orc::Ctx c;
};

namespace orc {

static std::map<std::string, PassFn>& registry() {
    static std::map<std::string, PassFn> r;
    return r;
}
static std::map<std::string, PassFn>& overrides() {
    static std::map<std::string, PassFn> r;
    return r;
}
PassRegistration::PassRegistration(const char* shader, PassFn fn) { registry()[shader] = fn; }
PassOverride::PassOverride(const char* shader, PassFn fn) { overrides()[shader] = fn; }
PassFn findPass(const std::string& shader) {
    auto ov = overrides().find(shader);
    if (ov != overrides().end()) return ov->second;
    auto it = registry().find(shader);
    return it == registry().end() ? nullptr : it->second;
}

void parallelFor(int threads, int n, const std::function<void(int)>& fn) {
    if (threads <= 1 || n <= 1) {
        for (int i = 0; i < n; i++) fn(i);
        return;
    }
    int t = threads < n ? threads : n;
    std::vector<std::thread> pool;
    pool.reserve(t);
    for (int k = 0; k < t; k++) {
        pool.emplace_back([=, &fn]() {
            for (int i = k; i < n; i += t) fn(i);
        });
    }
    for (auto& th : pool) th.join();
}

Image* Ctx::resolve(plain_image_handle h) {
    if (h.type == PLAIN_IMAGE_HANDLE_SWAPCHAIN) return &swapchain;
    if (h.type == PLAIN_IMAGE_HANDLE_TRANSIENT) return h.index < transientImages.size() ? &transientImages[h.index] : nullptr;
    return h.index < images.size() ? &images[h.index] : nullptr;
}

static View makeView(Ctx* ctx, const plain_image_resource& r) {
    View v;
    Image* img = ctx->resolve(r.image);
    if (!img || r.mip_level >= img->mips.size()) return v;
    v.img = img;
    v.mip = (int)r.mip_level;
    return v;
}
View PassCtx::target(uint32_t attachment) const {
    if (attachment >= exec->targets.size()) return View();
    plain_image_resource r{exec->targets[attachment].image, exec->targets[attachment].mip_level, 0};
    return makeView(ctx, r);
}
View PassCtx::sampled(uint32_t binding) const {
    for (auto& r : exec->sampledImages) if (r.binding == binding) return makeView(ctx, r);
    return View();
}
View PassCtx::storage(uint32_t binding) const {
    for (auto& r : exec->storageImages) if (r.binding == binding) return makeView(ctx, r);
    return View();
}
View PassCtx::bindless(uint32_t index) const {
    View v;
    if (index < ctx->images.size()) { v.img = &ctx->images[index]; v.mip = 0; }
    return v;
}
uint8_t* PassCtx::sbuf(uint32_t binding, size_t* size) const {
    for (auto& r : exec->storageBuffers)
        if (r.binding == binding && r.buffer < ctx->storageBuffers.size()) {
            if (size) *size = ctx->storageBuffers[r.buffer].data.size();
            return ctx->storageBuffers[r.buffer].data.data();
        }
    return nullptr;
}
const uint8_t* PassCtx::ubuf(uint32_t binding, size_t* size) const {
    for (auto& r : exec->uniformBuffers)
        if (r.binding == binding && r.buffer < ctx->uniformBuffers.size()) {
            if (size) *size = ctx->uniformBuffers[r.buffer].data.size();
            return ctx->uniformBuffers[r.buffer].data.data();
        }
    return nullptr;
}
void PassCtx::forEachGroup(const std::function<void(int, int, int)>& fn) const {
    int gx = (int)exec->dispatch[0], gy = (int)exec->dispatch[1], gz = (int)exec->dispatch[2];
    parallelFor(ctx->threads, gy * gz, [&](int i) {
        int y = i % gy, z = i / gy;
        for (int x = 0; x < gx; x++) fn(x, y, z);
    });
}
void PassCtx::forEachInvocation(int lx, int ly, int lz, const std::function<void(int, int, int)>& fn) const {
    forEachGroup([&](int gx, int gy, int gz) {
        for (int z = 0; z < lz; z++)
            for (int y = 0; y < ly; y++)
                for (int x = 0; x < lx; x++) fn(gx * lx + x, gy * ly + y, gz * lz + z);
    });
}

}  // namespace orc

using namespace orc;

static int fail(plain_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->c.lastError = msg;
    return 1;
}

extern "C" {

int PLAIN_FN(backend_create)(int device, uint32_t width, uint32_t height, plain_ctx** out_ctx) {
    (void)device;
    if (!out_ctx) return 1;
    plain_ctx* ctx = new plain_ctx();
    plain_image_desc d{};
    d.width = width; d.height = height; d.depth = 1;
    d.type = PLAIN_IMAGE_TYPE_2D; d.format = PLAIN_FORMAT_BGRA8_UNORM;  // VulkanSurface.cpp:41-46
    d.usage_flags = PLAIN_USAGE_STORAGE; d.mip_count = PLAIN_MIPS_ONE;
    ctx->c.swapchain.allocate(d);
    const char* t = getenv("ORACLE_THREADS");
    int n = t ? atoi(t) : (int)std::thread::hardware_concurrency();
    ctx->c.threads = n > 0 ? n : 1;
    *out_ctx = ctx;
    return 0;
}
void PLAIN_FN(backend_destroy)(plain_ctx* ctx) { delete ctx; }
const char* PLAIN_FN(last_error)(plain_ctx* ctx) { return ctx ? ctx->c.lastError.c_str() : "null context"; }
int PLAIN_FN(recreate_swapchain)(plain_ctx* ctx, uint32_t width, uint32_t height) {
    plain_image_desc d = ctx->c.swapchain.desc;
    d.width = width; d.height = height;
    ctx->c.swapchain.allocate(d);
    return 0;
}

int PLAIN_FN(create_image)(plain_ctx* ctx, const plain_image_desc* desc, const void* initial_data, size_t initial_data_size, plain_image_handle* out) {
    if (!desc || !out) return fail(ctx, "create_image: null argument");
    if (formatBytesPerTexel(desc->format) == 0) return fail(ctx, "create_image: format not supported on the frame path");
    Image img;
    img.allocate(*desc);
    if (initial_data) {
        size_t off = 0;
        const uint8_t* src = (const uint8_t*)initial_data;
        size_t levels = desc->mip_count == PLAIN_MIPS_FULL_CHAIN_ALREADY_IN_DATA ? img.mips.size() : 1;
        for (size_t i = 0; i < levels; i++) {
            size_t n = img.mips[i].data.size();
            if (off + n > initial_data_size) return fail(ctx, "create_image: initial data too small");
            memcpy(img.mips[i].data.data(), src + off, n);
            off += n;
        }
        if (desc->format == PLAIN_FORMAT_RGBA8)
            for (size_t t = 3; t < img.mips[0].data.size(); t += 4) if (img.mips[0].data[t] != 255) { img.transparentTexels = true; break; }
    }
    ctx->c.images.push_back(std::move(img));
    out->type = PLAIN_IMAGE_HANDLE_DEFAULT;
    out->index = (uint32_t)ctx->c.images.size() - 1;
    return 0;
}
int PLAIN_FN(create_temporary_image)(plain_ctx* ctx, const plain_image_desc* desc, plain_image_handle* out) {
    // valid for one frame; reused across frames when the description matches (RenderBackend.cpp:1026-1123)
    for (size_t i = 0; i < ctx->c.transientImages.size(); i++) {
        Image& t = ctx->c.transientImages[i];
        if (!t.inUse && memcmp(&t.desc, desc, sizeof(*desc)) == 0) {
            t.inUse = true;
            out->type = PLAIN_IMAGE_HANDLE_TRANSIENT; out->index = (uint32_t)i;
            return 0;
        }
    }
    Image img;
    img.allocate(*desc);
    img.inUse = true;
    ctx->c.transientImages.push_back(std::move(img));
    out->type = PLAIN_IMAGE_HANDLE_TRANSIENT;
    out->index = (uint32_t)ctx->c.transientImages.size() - 1;
    return 0;
}
int PLAIN_FN(resize_images)(plain_ctx* ctx, const plain_image_handle* images, uint32_t n, uint32_t width, uint32_t height) {
    for (uint32_t i = 0; i < n; i++) {
        Image* img = ctx->c.resolve(images[i]);
        if (!img) return fail(ctx, "resize_images: invalid handle");
        plain_image_desc d = img->desc;
        d.width = width; d.height = height;
        img->allocate(d);
    }
    return 0;
}
int PLAIN_FN(get_image_description)(plain_ctx* ctx, plain_image_handle image, plain_image_desc* out) {
    Image* img = ctx->c.resolve(image);
    if (!img) return fail(ctx, "get_image_description: invalid handle");
    *out = img->desc;
    return 0;
}
int PLAIN_FN(get_image_global_texture_array_index)(plain_ctx* ctx, plain_image_handle image, uint32_t* out) {
    if (image.type != PLAIN_IMAGE_HANDLE_DEFAULT || image.index >= ctx->c.images.size()) return fail(ctx, "global texture index: invalid handle");
    *out = image.index;
    return 0;
}
int PLAIN_FN(create_uniform_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out) {
    Buffer b;
    b.data.assign(size, 0);
    if (initial_data) memcpy(b.data.data(), initial_data, size);
    ctx->c.uniformBuffers.push_back(std::move(b));
    *out = (uint32_t)ctx->c.uniformBuffers.size() - 1;
    return 0;
}
int PLAIN_FN(create_storage_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out) {
    Buffer b;
    b.data.assign(size, 0);
    if (initial_data) memcpy(b.data.data(), initial_data, size);
    ctx->c.storageBuffers.push_back(std::move(b));
    *out = (uint32_t)ctx->c.storageBuffers.size() - 1;
    return 0;
}
int PLAIN_FN(create_sampler)(plain_ctx* ctx, const plain_sampler_desc* desc, plain_handle* out) {
    ctx->c.samplers.push_back(*desc);
    *out = (uint32_t)ctx->c.samplers.size() - 1;
    return 0;
}
int PLAIN_FN(get_swapchain_input_image)(plain_ctx* ctx, plain_image_handle* out) {
    (void)ctx;
    out->type = PLAIN_IMAGE_HANDLE_SWAPCHAIN;
    out->index = 0;
    return 0;
}

static int fillPass(plain_ctx* ctx, PassRecord& p, const char* shader, const plain_spec_const* consts, uint32_t n) {
    p.shader = shader;
    p.spec.clear();
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t* d = (const uint8_t*)consts[i].data;
        p.spec[consts[i].location] = std::vector<uint8_t>(d, d + consts[i].size);
    }
    p.fn = findPass(p.shader);
    if (!p.fn) return fail(ctx, std::string("no oracle pass for shader '") + shader + "'");
    return 0;
}
int PLAIN_FN(create_compute_pass)(plain_ctx* ctx, const char* shader, const plain_spec_const* consts, uint32_t n_consts, const char* debug_name, plain_handle* out) {
    PassRecord p;
    if (fillPass(ctx, p, shader, consts, n_consts)) return 1;
    p.name = debug_name ? debug_name : shader;
    ctx->c.passes.push_back(std::move(p));
    *out = (uint32_t)ctx->c.passes.size() - 1;
    return 0;
}
int PLAIN_FN(update_compute_pass_shader_description)(plain_ctx* ctx, plain_handle pass, const char* shader, const plain_spec_const* consts, uint32_t n_consts) {
    if (pass >= ctx->c.passes.size()) return fail(ctx, "update pass: invalid handle");
    return fillPass(ctx, ctx->c.passes[pass], shader, consts, n_consts);
}
int PLAIN_FN(set_global_descriptor_set_resources)(plain_ctx* ctx, const plain_pass_resources* r) {
    for (uint32_t i = 0; i < r->n_uniform_buffers; i++)
        if (r->uniform_buffers[i].binding == 0) ctx->c.globalUniformBuffer = r->uniform_buffers[i].buffer;
    return 0;
}

int PLAIN_FN(new_frame)(plain_ctx* ctx) {
    ctx->c.execs.clear();
    for (auto& t : ctx->c.transientImages) t.inUse = false;
    return 0;
}
int PLAIN_FN(set_compute_pass_execution)(plain_ctx* ctx, const plain_compute_pass_execution* e) {
    if (e->pass >= ctx->c.passes.size()) return fail(ctx, "set_compute_pass_execution: invalid pass handle");
    ExecRecord r;
    r.pass = e->pass;
    const plain_pass_resources& s = e->resources;
    r.storageBuffers.assign(s.storage_buffers, s.storage_buffers + s.n_storage_buffers);
    r.uniformBuffers.assign(s.uniform_buffers, s.uniform_buffers + s.n_uniform_buffers);
    r.sampledImages.assign(s.sampled_images, s.sampled_images + s.n_sampled_images);
    r.storageImages.assign(s.storage_images, s.storage_images + s.n_storage_images);
    const uint8_t* pc = (const uint8_t*)e->push_constants;
    if (pc) r.pushConstants.assign(pc, pc + e->push_constant_size);
    for (int i = 0; i < 3; i++) r.dispatch[i] = e->dispatch_count[i];
    ctx->c.execs.push_back(std::move(r));
    return 0;
}
int PLAIN_FN(prepare_for_drawcall_recording)(plain_ctx* ctx) { (void)ctx; return 0; }

// ---- meshes and graphic passes (RenderBackend.h:57-96) ----
int PLAIN_FN(create_meshes)(plain_ctx* ctx, const plain_mesh_binary* meshes, uint32_t n, plain_handle* out) {
    for (uint32_t i = 0; i < n; i++) {
        const plain_mesh_binary& m = meshes[i];
        if (m.index_count % 3 != 0 || !m.index_buffer || !m.vertex_buffer) return fail(ctx, "create_meshes: triangle list with index and vertex data expected");
        Mesh r;
        r.indexCount = m.index_count; r.vertexCount = m.vertex_count;
        r.index32 = !(m.index_count < 65535u);  // RenderBackend.cpp:483-488
        const size_t ib = (size_t)m.index_count * (r.index32 ? 4 : 2), vb = (size_t)m.vertex_count * 28;
        r.indices.assign((const uint8_t*)m.index_buffer, (const uint8_t*)m.index_buffer + ib);
        r.vertices.assign((const uint8_t*)m.vertex_buffer, (const uint8_t*)m.vertex_buffer + vb);
        for (uint32_t k = 0; k < m.index_count; k++) {
            uint32_t idx = r.index32 ? ((const uint32_t*)r.indices.data())[k] : ((const uint16_t*)r.indices.data())[k];
            if (idx >= m.vertex_count) return fail(ctx, "create_meshes: index out of range");
        }
        ctx->c.meshes.push_back(std::move(r));
        out[i] = (uint32_t)ctx->c.meshes.size() - 1;
    }
    return 0;
}
int PLAIN_FN(create_graphic_pass)(plain_ctx* ctx, const plain_graphic_pass_desc* d, plain_handle* out) {
    if (!d || !d->vertex_shader || !d->fragment_shader) return fail(ctx, "create_graphic_pass: vertex and fragment shader expected");
    PassRecord p;
    p.graphic = true;
    p.shader = std::string(d->vertex_shader) + "+" + d->fragment_shader;
    p.name = d->debug_name ? d->debug_name : p.shader;
    for (uint32_t i = 0; i < d->n_vertex_consts; i++) {
        const uint8_t* s = (const uint8_t*)d->vertex_consts[i].data;
        p.spec[d->vertex_consts[i].location] = std::vector<uint8_t>(s, s + d->vertex_consts[i].size);
    }
    p.cullMode = d->cull_mode; p.clampDepth = d->clamp_depth; p.depthFunction = d->depth_function; p.depthWrite = d->depth_write;
    p.attachments.assign(d->attachments, d->attachments + d->n_attachments);
    p.pushSize = p.shader.rfind("sunShadow.vert", 0) == 0 ? 8u : 16u;
    p.fn = findPass(p.shader);
    if (!p.fn) return fail(ctx, std::string("no oracle rasteriser program for the shader pair '") + p.shader + "'");
    ctx->c.passes.push_back(std::move(p));
    *out = (uint32_t)ctx->c.passes.size() - 1;
    return 0;
}
int PLAIN_FN(set_graphic_pass_execution)(plain_ctx* ctx, const plain_graphic_pass_execution* e) {
    if (e->pass >= ctx->c.passes.size() || !ctx->c.passes[e->pass].graphic) return fail(ctx, "set_graphic_pass_execution: not a graphic pass");
    if (e->n_targets != ctx->c.passes[e->pass].attachments.size()) return fail(ctx, "set_graphic_pass_execution: one target per attachment expected");
    ExecRecord r;
    r.pass = e->pass;
    const plain_pass_resources& s = e->resources;
    r.storageBuffers.assign(s.storage_buffers, s.storage_buffers + s.n_storage_buffers);
    r.uniformBuffers.assign(s.uniform_buffers, s.uniform_buffers + s.n_uniform_buffers);
    r.sampledImages.assign(s.sampled_images, s.sampled_images + s.n_sampled_images);
    r.storageImages.assign(s.storage_images, s.storage_images + s.n_storage_images);
    r.targets.assign(e->targets, e->targets + e->n_targets);
    r.dispatch[0] = r.dispatch[1] = r.dispatch[2] = 0;
    ctx->c.execs.push_back(std::move(r));
    return 0;
}
int PLAIN_FN(draw_meshes)(plain_ctx* ctx, const plain_handle* meshes, uint32_t n, const void* push_constants, plain_handle pass, int32_t worker_index) {
    (void)worker_index;
    if (pass >= ctx->c.passes.size() || !ctx->c.passes[pass].graphic) return fail(ctx, "draw_meshes: not a graphic pass");
    ExecRecord* rec = nullptr;
    for (auto& e : ctx->c.execs) if (e.pass == pass) rec = &e;
    if (!rec) return fail(ctx, "draw_meshes: the pass has no execution this frame (set_graphic_pass_execution)");
    const uint32_t ps = ctx->c.passes[pass].pushSize;
    for (uint32_t i = 0; i < n; i++) {
        if (meshes[i] >= ctx->c.meshes.size()) return fail(ctx, "draw_meshes: invalid mesh handle");
        DrawRecord d;
        d.mesh = meshes[i];
        memset(d.push, 0, sizeof(d.push));
        memcpy(d.push, (const uint8_t*)push_constants + (size_t)i * ps, ps);
        rec->draws.push_back(d);
    }
    return 0;
}
int PLAIN_FN(set_uniform_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size) {
    if (buffer >= ctx->c.uniformBuffers.size() || size > ctx->c.uniformBuffers[buffer].data.size()) return fail(ctx, "set_uniform_buffer_data: invalid buffer/size");
    FillOrder f{true, buffer, std::vector<uint8_t>((const uint8_t*)data, (const uint8_t*)data + size)};
    ctx->c.fills.push_back(std::move(f));
    return 0;
}
int PLAIN_FN(set_storage_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size) {
    if (buffer >= ctx->c.storageBuffers.size() || size > ctx->c.storageBuffers[buffer].data.size()) return fail(ctx, "set_storage_buffer_data: invalid buffer/size");
    FillOrder f{false, buffer, std::vector<uint8_t>((const uint8_t*)data, (const uint8_t*)data + size)};
    ctx->c.fills.push_back(std::move(f));
    return 0;
}
int PLAIN_FN(render_frame)(plain_ctx* ctx, int present) {
    (void)present;
    Ctx& c = ctx->c;
    // all fills of the frame land before any pass (RenderBackend.cpp:896-911)
    for (auto& f : c.fills) {
        Buffer& b = f.uniform ? c.uniformBuffers[f.buffer] : c.storageBuffers[f.buffer];
        memcpy(b.data.data(), f.data.data(), f.data.size());
    }
    c.fills.clear();
    c.timings.clear();
    for (auto& e : c.execs) {
        PassCtx pc;
        pc.ctx = &c;
        pc.pass = &c.passes[e.pass];
        pc.exec = &e;
        memset(&pc.g, 0, sizeof(pc.g));
        if (c.globalUniformBuffer < c.uniformBuffers.size()) {
            auto& gb = c.uniformBuffers[c.globalUniformBuffer].data;
            memcpy(&pc.g, gb.data(), gb.size() < sizeof(pc.g) ? gb.size() : sizeof(pc.g));
        }
        auto t0 = std::chrono::steady_clock::now();
        pc.pass->fn(pc);
        auto t1 = std::chrono::steady_clock::now();
        if (c.timingEnabled) {
            plain_pass_time pt;
            snprintf(pt.name, sizeof(pt.name), "%s", pc.pass->name.c_str());
            pt.time_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
            c.timings.push_back(pt);
        }
    }
    return 0;
}
int PLAIN_FN(submit_recorded_passes)(plain_ctx* ctx) {
    const int rc = PLAIN_FN(render_frame)(ctx, 0);
    ctx->c.execs.clear();
    return rc;
}
int PLAIN_FN(wait_for_gpu_idle)(plain_ctx* ctx) { (void)ctx; return 0; }
int PLAIN_FN(get_renderpass_timings)(plain_ctx* ctx, plain_pass_time* out, uint32_t capacity, uint32_t* out_count) {
    uint32_t n = (uint32_t)ctx->c.timings.size();
    if (out_count) *out_count = n;
    for (uint32_t i = 0; i < n && i < capacity; i++) out[i] = ctx->c.timings[i];
    return 0;
}
int PLAIN_FN(set_timing_enabled)(plain_ctx* ctx, int enabled) { ctx->c.timingEnabled = enabled != 0; return 0; }

int PLAIN_FN(write_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, const void* data, size_t size) {
    Image* img = ctx->c.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, "write_image: invalid handle/mip");
    if (size != img->mips[mip].data.size()) return fail(ctx, "write_image: size mismatch");
    memcpy(img->mips[mip].data.data(), data, size);
    return 0;
}
int PLAIN_FN(read_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void* out, size_t size) {
    Image* img = ctx->c.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, "read_image: invalid handle/mip");
    if (size != img->mips[mip].data.size()) return fail(ctx, "read_image: size mismatch");
    memcpy(out, img->mips[mip].data.data(), size);
    return 0;
}
int PLAIN_FN(read_storage_buffer)(plain_ctx* ctx, plain_handle buffer, void* out, size_t size) {
    if (buffer >= ctx->c.storageBuffers.size() || size > ctx->c.storageBuffers[buffer].data.size()) return fail(ctx, "read_storage_buffer: invalid buffer/size");
    memcpy(out, ctx->c.storageBuffers[buffer].data.data(), size);
    return 0;
}
int PLAIN_FN(write_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, const void* data, size_t size) { return PLAIN_FN(write_image)(ctx, image, mip, data, size); }
int PLAIN_FN(read_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void* out, size_t size) { return PLAIN_FN(read_image)(ctx, image, mip, out, size); }
static int imageRows(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, void* host, size_t size, bool toImage) {
    Image* img = ctx->c.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, "image rows: invalid handle/mip");
    MipLevel& m = img->mips[mip];
    if (m.d != 1 || rowBegin > rowEnd || rowEnd > (uint32_t)m.h) return fail(ctx, "image rows: invalid row range");
    const size_t pitch = m.data.size() / (size_t)m.h;
    if (size != pitch * (rowEnd - rowBegin)) return fail(ctx, "image rows: size mismatch");
    if (toImage) memcpy(m.data.data() + pitch * rowBegin, host, size);
    else memcpy(host, m.data.data() + pitch * rowBegin, size);
    return 0;
}
int PLAIN_FN(write_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, const void* data, size_t size) { return imageRows(ctx, image, mip, rowBegin, rowEnd, (void*)data, size, true); }
int PLAIN_FN(read_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, void* out, size_t size) { return imageRows(ctx, image, mip, rowBegin, rowEnd, out, size, false); }
int PLAIN_FN(get_image_device_pointer)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void** out_ptr, size_t* out_size) {
    Image* img = ctx->c.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, "get_image_device_pointer: invalid handle/mip");
    *out_ptr = img->mips[mip].data.data();
    if (out_size) *out_size = img->mips[mip].data.size();
    return 0;
}
int PLAIN_FN(get_storage_buffer_device_pointer)(plain_ctx* ctx, plain_handle buffer, void** out_ptr, size_t* out_size) {
    if (buffer >= ctx->c.storageBuffers.size()) return fail(ctx, "get_storage_buffer_device_pointer: invalid buffer");
    *out_ptr = ctx->c.storageBuffers[buffer].data.data();
    if (out_size) *out_size = ctx->c.storageBuffers[buffer].data.size();
    return 0;
}
int PLAIN_FN(get_last_frame_launch_count)(plain_ctx* ctx, uint32_t* out) { (void)ctx; *out = 0; return 0; }
int PLAIN_FN(set_graph_replay_enabled)(plain_ctx* ctx, int enabled) { (void)ctx; (void)enabled; return 0; }
int PLAIN_FN(set_pass_fusion_enabled)(plain_ctx*, int) { return 0; }  // the oracle runs every pass as recorded
// peer exchange over NVLink: CUDA backend only
static int peerUnsupported(plain_ctx* ctx) { ctx->c.lastError = "peer exchange is not available in the CPU oracle"; return 1; }
int PLAIN_FN(peer_init)(plain_ctx* ctx, uint32_t, uint32_t) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_get_sync_handle)(plain_ctx* ctx, void*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_open_sync)(plain_ctx* ctx, uint32_t, const void*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_get_image_handle)(plain_ctx* ctx, plain_image_handle, void*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_open_image)(plain_ctx* ctx, plain_image_handle, uint32_t, const void*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_image_ready)(plain_ctx*, plain_image_handle) { return 0; }
int PLAIN_FN(peer_push_rows)(plain_ctx* ctx, uint32_t, const plain_peer_push*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_barrier)(plain_ctx* ctx) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_push_rows_deferred)(plain_ctx* ctx, uint32_t, const plain_peer_push*) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_flush_deferred)(plain_ctx* ctx) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_allreduce_sum_u32)(plain_ctx* ctx, plain_handle, uint32_t) { return peerUnsupported(ctx); }
int PLAIN_FN(peer_error)(plain_ctx*, uint32_t* out_error) { *out_error = 0; return 0; }
int PLAIN_FN(peer_error_poll)(plain_ctx*, uint32_t* out_error) { *out_error = 0; return 0; }
int PLAIN_FN(set_concurrent_passes_enabled)(plain_ctx* ctx, int) { (void)ctx; return 0; }
int PLAIN_FN(join_transfers)(plain_ctx* ctx) { (void)ctx; return 0; }
int PLAIN_FN(get_stream)(plain_ctx* ctx, void** out_stream) { (void)ctx; *out_stream = nullptr; return 0; }
int PLAIN_FN(device_selftest)(plain_ctx* ctx, uint64_t* out_mismatches) { (void)ctx; for (int i = 0; i < 8; i++) out_mismatches[i] = 0; return 0; }

}  // extern "C"
