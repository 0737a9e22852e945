// backend.cu - CUDA implementation of include/plain_b200.h (symbols plain_*): the drop-in for the compute-pass
// subset of the reference's RenderBackend (Plain/src/Runtime/Rendering/Backend/RenderBackend.h:33-110).
//
//   * handles are indices into tables owned by the context (RenderHandles.h:4-33); no public destroy
//   * set_*_buffer_data copies the bytes before returning and applies them at render_frame, before any pass of the
//     frame (RenderBackend.cpp:315-321, 896-911): the fills of a frame are packed into one pinned staging block,
//     moved with ONE host->device copy and scattered to their buffers by one kernel
//   * executions are replayed in submission order on one stream (RenderBackend.cpp:769-786); the in-order stream
//     replaces the reference's barriers
//   * an unchanged pass list (same passes, resources, push constants) is replayed as an instantiated CUDA graph:
//     kernels read uniform/storage buffers through pointers, so only buffer contents differ between frames
//   * there is no CPU fallback: every pass is a CUDA kernel looked up by the reference's shader file name; an unknown
//     shader name or a missing device is an error
#include <cstdlib>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include "pass_common.cuh"

namespace pb {

static std::map<std::string, LaunchFn>& registry() {
    static std::map<std::string, LaunchFn> r;
    return r;
}
PassRegistration::PassRegistration(const char* shader, LaunchFn fn) { registry()[shader] = fn; }

struct FillOrder {
    bool uniform;
    uint32_t buffer;
    size_t offsetInStaging, size;
};
struct FillSegment {  // one entry of the scatter table that travels at the head of the staging block
    unsigned long long dst;
    uint32_t srcOffset, size;
};
static const size_t kMaxFillSegments = 64;
static const size_t kStagingHeader = kMaxFillSegments * sizeof(FillSegment);
static const size_t kStagingBytes = 1u << 20;

struct CachedGraph {
    cudaGraphExec_t exec = nullptr;
    uint32_t launches = 0;
    uint64_t lastUse = 0;  // submission counter of the last replay (least recently used entries are evicted, kMaxCachedGraphs)
};

struct Backend {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    // Asynchronous image uploads / read-backs run on their own streams (copy engines) so that the raster-pass outputs of
    // frame N+1 travel while the passes of frame N execute. Ordering: an upload into an image waits for the last
    // submission that referenced it (submissionDone ring) and for a pending read-back of it; a submission waits for
    // every upload issued before it and for the read-backs still in flight; a read-back waits for the passes
    // submitted so far.
    cudaStream_t uploadStream = nullptr, downloadStream = nullptr;
    static const int kSubmissionRing = 64;  // a row-sharded frame is ~11 submissions; an image is reused after 2 frames
    cudaEvent_t submissionDone[kSubmissionRing] = {};
    cudaEvent_t uploadsDone = nullptr, computeMark = nullptr;
    long long submissionCounter = 0;
    bool uploadsPending = false;
    // Passes of one submission are scheduled onto the pass stream + side streams from the hazards between their declared
    // resources (what the reference derives its barriers from, RenderBackend.cpp:632-767): a pass waits for the passes it
    // conflicts with and nothing else, so the small low-occupancy passes (sky LUTs, histogram, HiZ tail, culling) run next to
    // each other and next to the froxel / GI chains instead of serially. Captured into the frame's CUDA graph as parallel branches.
    static const int kSideStreams = 2;
    cudaStream_t sideStreams[kSideStreams] = {};
    std::vector<cudaEvent_t> passEvents;
    cudaEvent_t forkEvent = nullptr;
    bool concurrentPasses = true;
    // peer exchange over NVLink (row-sharded frames): a sync block per rank, mapped by every other rank
    //   u32 flags[PLAIN_MAX_PEERS]   flags[p] = epoch of the last barrier rank p has signalled (written by rank p)
    //   u32 error                    set when a barrier gave up waiting
    //   u32 scratch[2][PLAIN_MAX_PEERS][kPeerReduceMax]   all-reduce partials, double-buffered by all-reduce parity
    static const uint32_t kPeerReduceMax = 256;
    //   u32 flagsDeferred[PLAIN_MAX_PEERS]   the same for the barrier of the deferred (next-frame) exchanges, which runs on its own stream
    static const size_t kPeerFlagsOffset = 0, kPeerErrorOffset = PLAIN_MAX_PEERS, kPeerScratchOffset = 2 * PLAIN_MAX_PEERS;
    static const size_t kPeerFlagsDeferredOffset = 2 * PLAIN_MAX_PEERS + 2 * PLAIN_MAX_PEERS * kPeerReduceMax;
    static const size_t kPeerSyncWords = kPeerFlagsDeferredOffset + PLAIN_MAX_PEERS;
    uint32_t peerRank = 0, peerCount = 0;
    uint32_t* peerSync[PLAIN_MAX_PEERS] = {};  // [peerRank] = own block
    uint32_t peerEpoch = 0, peerReduceCount = 0;
    // Deferred exchanges: rows that only the NEXT frame reads (GI history, froxel history, TAA history) are pushed on peerStream,
    // behind an event of the pass stream, while the rest of the frame runs; one barrier on the deferred flag set closes them
    // (peer_flush_deferred) and the first submission of the next frame waits for it.
    cudaStream_t peerStream = nullptr;
    cudaEvent_t peerFork = nullptr, peerDeferredDone = nullptr;
    uint32_t peerEpochDeferred = 0;
    bool peerDeferredDirty = false, peerDeferredPending = false, frameFresh = false;
    uint32_t peerBarriersThisFrame = 0;  // synchronous barriers since new_frame: a deferred push needs at least one before it (write-after-read across frames)
    std::vector<DeviceImage> images, transientImages;
    // two presentable images, flipped by new_frame (a swapchain hands out a different image every frame): the read-back of
    // frame N does not hold up the tonemapping pass of frame N+1
    DeviceImage swapchainImages[2];
    int swapchainCurrent = 0;
    std::vector<DeviceBuffer> uniformBuffers, storageBuffers;
    std::vector<plain_sampler_desc> samplers;
    std::vector<PassRecord> passes;
    std::vector<DeviceMesh> meshes;
    struct RasterScratch { RasterDraw* draws = nullptr; size_t drawCapacity = 0; uint32_t* triInfo = nullptr; size_t triCapacity = 0; float* vertexCache = nullptr; size_t vertexCapacity = 0; };
    std::unordered_map<uint32_t, RasterScratch> rasterScratch;  // by pass handle
    struct VisBuffer { unsigned long long* ptr = nullptr; size_t texels = 0; };
    std::unordered_map<uint64_t, VisBuffer> visBuffers;         // by depth target (image handle + mip): shared by the prepass and the G-buffer fill
    std::vector<RasterDraw> rasterDrawStaging;
    std::vector<ExecRecord> execs;
    std::vector<FillOrder> fills;
    uint32_t globalUniformBuffer = PLAIN_INVALID_INDEX;
    std::string lastError;

    unsigned char* stagingHost = nullptr;  // pinned
    unsigned char* stagingDevice = nullptr;
    size_t stagingUsed = kStagingHeader;
    cudaEvent_t stagingConsumed = nullptr;
    bool stagingInFlight = false;

    BindlessEntry* bindlessDevice = nullptr;
    uint32_t* peerErrorHost = nullptr;  // pinned mirror of the peer error word (peer_error_poll)
    ShadingTables* tablesDevice = nullptr;
    static const uint32_t kMaxBindless = 4096;

    bool timingEnabled = false;
    std::vector<plain_pass_time> timings;
    std::vector<cudaEvent_t> timingEvents;

    bool graphEnabled = false;
    unsigned int* fusionCounters = nullptr;  // device scratch of fused runs (grid-barrier counters + error word)
    bool fusionEnabled = true;  // planFusions: indirectLightUpscale.comp folded into gbufferShading.comp (set_pass_fusion_enabled)
    std::unordered_map<uint64_t, CachedGraph> graphs;
    static const size_t kMaxCachedGraphs = 32;
    uint64_t graphUseCounter = 0;
    uint32_t passEpoch = 0;  // bumped when a pass description or an image allocation changes: invalidates cached graphs
    uint32_t lastFrameLaunches = 0;
    uint32_t launchCounter = 0;

    DeviceImage* resolve(plain_image_handle h) {
        if (h.type == PLAIN_IMAGE_HANDLE_SWAPCHAIN) return &swapchainImages[swapchainCurrent];
        if (h.type == PLAIN_IMAGE_HANDLE_TRANSIENT) return h.index < transientImages.size() ? &transientImages[h.index] : nullptr;
        return h.index < images.size() ? &images[h.index] : nullptr;
    }
};

static int formatBytesPerTexel(uint32_t fmt) {
    switch (fmt) {
        case PLAIN_FORMAT_R8: return 1;
        case PLAIN_FORMAT_RG8: return 2;
        case PLAIN_FORMAT_RGBA8: return 4;
        case PLAIN_FORMAT_R16_SFLOAT: return 2;
        case PLAIN_FORMAT_RG16_SFLOAT: return 4;
        case PLAIN_FORMAT_RG32_SFLOAT: return 8;
        case PLAIN_FORMAT_RG16_SNORM: return 4;
        case PLAIN_FORMAT_RGBA16_SFLOAT: return 8;
        case PLAIN_FORMAT_RGBA16_SNORM: return 8;
        case PLAIN_FORMAT_RGBA32_SFLOAT: return 16;
        case PLAIN_FORMAT_R11G11B10_UFLOAT: return 4;
        case PLAIN_FORMAT_DEPTH16: return 2;
        case PLAIN_FORMAT_DEPTH32: return 4;
        case PLAIN_FORMAT_BGRA8_UNORM: return 4;
        case PLAIN_FORMAT_RGBA32_UINT: return 16;
        default: return 0;  // BCn: material textures are baked into the G-buffer, not on the frame path
    }
}

static int computeMipCount(const plain_image_desc& d) {
    if (d.mip_count == PLAIN_MIPS_ONE) return 1;
    if (d.mip_count == PLAIN_MIPS_MANUAL) return (int)d.manual_mip_count;
    uint32_t m = d.width > d.height ? d.width : d.height;
    if (d.depth > m) m = d.depth;
    int n = 1;
    while (m > 1) { m >>= 1; n++; }  // 1 + floor(log2(max)), MathUtils.cpp:17-19
    return n;
}

static bool hasCornerBrick(const plain_image_desc& d) { return d.type == PLAIN_IMAGE_TYPE_3D && d.format == PLAIN_FORMAT_R16_SFLOAT; }
static bool allocateImage(Backend& b, DeviceImage& img, const plain_image_desc& d) {
    if (img.ptr) { cudaStreamSynchronize(b.uploadStream); cudaStreamSynchronize(b.downloadStream); cudaStreamSynchronize(b.stream); if (b.peerStream) cudaStreamSynchronize(b.peerStream); cudaFree(img.ptr); img.ptr = nullptr; }
    // the peers' copies were mapped for the old allocation (and the peers hold mappings of it): drop this rank's mappings so that
    // peer_image_ready turns false and the next exchange of the image goes back through the caller, which re-maps it on every rank
    for (uint32_t pr = 0; pr < PLAIN_MAX_PEERS; pr++) if (img.peerPtr[pr]) { cudaIpcCloseMemHandle(img.peerPtr[pr]); img.peerPtr[pr] = nullptr; }
    img.deferredExchange = false;
    if (img.corners) { cudaFree(img.corners); img.corners = nullptr; }
    img.desc = d;
    const int n = computeMipCount(d), bpt = formatBytesPerTexel(d.format);
    img.mips.assign(n, MipInfo());
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        MipInfo& m = img.mips[i];
        m.w = std::max((int)d.width >> i, 1);
        m.h = std::max((int)d.height >> i, 1);
        m.d = std::max((int)(d.depth ? d.depth : 1) >> i, 1);
        m.offset = off;
        m.bytes = (size_t)m.w * m.h * m.d * bpt;
        off = (off + m.bytes + 255) & ~(size_t)255;
    }
    img.bytes = off;
    if (cudaMalloc(&img.ptr, img.bytes ? img.bytes : 256) != cudaSuccess) return false;
    cudaMemsetAsync(img.ptr, 0, img.bytes ? img.bytes : 256, b.stream);
    if (hasCornerBrick(d)) {  // an all-zero brick has an all-zero copy
        const size_t entries = (size_t)(img.mips[0].w + 1) * (img.mips[0].h + 1) * (img.mips[0].d + 1);
        if (cudaMalloc(&img.corners, entries * sizeof(uint4)) != cudaSuccess) { cudaFree(img.ptr); img.ptr = nullptr; return false; }
        cudaMemsetAsync(img.corners, 0, entries * sizeof(uint4), b.stream);
        img.cornersStale = false;
    }
    b.passEpoch++;
    return true;
}

static ImgView makeView(LaunchCtx& c, const plain_image_resource& r, int expectFormat, const char* what, uint32_t binding) {
    ImgView v{nullptr, 0, 0, 0};
    DeviceImage* img = c.be->resolve(r.image);
    if (!img || r.mip_level >= img->mips.size()) {
        c.fail(c.pass->shader + ": " + what + " binding " + std::to_string(binding) + " has an invalid image handle or mip level");
        return v;
    }
    if (expectFormat >= 0 && (int)img->desc.format != expectFormat) {
        c.fail(c.pass->shader + ": " + what + " binding " + std::to_string(binding) + " has format " + std::to_string(img->desc.format) + ", expected " + std::to_string(expectFormat));
        return v;
    }
    const MipInfo& m = img->mips[r.mip_level];
    v.ptr = img->ptr + m.offset;
    v.w = m.w; v.h = m.h; v.d = m.d;
    return v;
}
ImgView LaunchCtx::target(uint32_t attachment, int expectFormat) {
    if (attachment >= exec->targets.size()) { fail(pass->shader + ": no render target for attachment " + std::to_string(attachment)); return ImgView{nullptr, 0, 0, 0}; }
    const plain_image_resource r{exec->targets[attachment].image, exec->targets[attachment].mip_level, attachment};
    return makeView(*this, r, expectFormat, "render target", attachment);
}
ImgView LaunchCtx::sampled(uint32_t binding, int expectFormat) {
    for (auto& r : exec->sampledImages) if (r.binding == binding) return makeView(*this, r, expectFormat, "sampled image", binding);
    fail(pass->shader + ": no sampled image at binding " + std::to_string(binding));
    return ImgView{nullptr, 0, 0, 0};
}
ImgView LaunchCtx::storage(uint32_t binding, int expectFormat) {
    for (auto& r : exec->storageImages) if (r.binding == binding) return makeView(*this, r, expectFormat, "storage image", binding);
    fail(pass->shader + ": no storage image at binding " + std::to_string(binding));
    return ImgView{nullptr, 0, 0, 0};
}
int LaunchCtx::sampledFormat(uint32_t binding) {
    for (auto& r : exec->sampledImages)
        if (r.binding == binding) { DeviceImage* img = be->resolve(r.image); return img ? (int)img->desc.format : -1; }
    return -1;
}
void* LaunchCtx::sbufRaw(uint32_t binding, size_t* size) {
    for (auto& r : exec->storageBuffers)
        if (r.binding == binding && r.buffer < be->storageBuffers.size()) {
            if (size) *size = be->storageBuffers[r.buffer].size;
            return be->storageBuffers[r.buffer].ptr;
        }
    fail(pass->shader + ": no storage buffer at binding " + std::to_string(binding));
    return nullptr;
}
const void* LaunchCtx::ubufRaw(uint32_t binding) {
    for (auto& r : exec->uniformBuffers)
        if (r.binding == binding && r.buffer < be->uniformBuffers.size()) return be->uniformBuffers[r.buffer].ptr;
    fail(pass->shader + ": no uniform buffer at binding " + std::to_string(binding));
    return nullptr;
}
void LaunchCtx::countLaunch(int n) { be->launchCounter += (uint32_t)n; }
const ExecRecord* LaunchCtx::be_exec(int index) const { return &be->execs[(size_t)index]; }
unsigned int* LaunchCtx::be_fusionCounters() const { return be->fusionCounters; }
const PassRecord* LaunchCtx::be_pass(uint32_t p) const { return &be->passes[p]; }
size_t LaunchCtx::be_imageCount() const { return be->images.size(); }
int LaunchCtx::be_imageFormat(uint32_t index) const { return index < be->images.size() ? (int)be->images[index].desc.format : -1; }

// scatter the staged fills to their buffers: one block per segment, 4-byte words (+ byte tail)
__global__ void scatterFillsKernel(const unsigned char* __restrict__ staging, int nSegments) {
    const FillSegment seg = ((const FillSegment*)staging)[blockIdx.x];
    if ((int)blockIdx.x >= nSegments) return;
    unsigned char* dst = (unsigned char*)seg.dst;
    const unsigned char* src = staging + seg.srcOffset;
    if ((((unsigned long long)dst | seg.srcOffset) & 3u) == 0) {
        const uint32_t words = seg.size >> 2;
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
        for (uint32_t i = (words << 2) + threadIdx.x; i < seg.size; i += blockDim.x) dst[i] = src[i];
    } else {
        for (uint32_t i = threadIdx.x; i < seg.size; i += blockDim.x) dst[i] = src[i];
    }
}

static uint64_t fnv(uint64_t h, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t hashExecs(const Backend& b) {
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, &b.passEpoch, sizeof(b.passEpoch));
    h = fnv(h, &b.globalUniformBuffer, sizeof(uint32_t));
    h = fnv(h, &b.swapchainCurrent, sizeof(int));  // the swapchain handle resolves to a different allocation every other frame
    for (auto& e : b.execs) {
        h = fnv(h, &e.pass, 4);
        uint32_t n;
        n = (uint32_t)e.storageBuffers.size(); h = fnv(h, &n, 4); h = fnv(h, e.storageBuffers.data(), n * sizeof(e.storageBuffers[0]));
        n = (uint32_t)e.uniformBuffers.size(); h = fnv(h, &n, 4); h = fnv(h, e.uniformBuffers.data(), n * sizeof(e.uniformBuffers[0]));
        n = (uint32_t)e.sampledImages.size(); h = fnv(h, &n, 4); h = fnv(h, e.sampledImages.data(), n * sizeof(e.sampledImages[0]));
        n = (uint32_t)e.storageImages.size(); h = fnv(h, &n, 4); h = fnv(h, e.storageImages.data(), n * sizeof(e.storageImages[0]));
        n = (uint32_t)e.pushConstants.size(); h = fnv(h, &n, 4); h = fnv(h, e.pushConstants.data(), n);
        h = fnv(h, e.dispatch, sizeof(e.dispatch));
        h = fnv(h, &e.rowBegin, 12);
        n = (uint32_t)e.targets.size(); h = fnv(h, &n, 4); h = fnv(h, e.targets.data(), n * sizeof(e.targets[0]));
        n = (uint32_t)e.draws.size(); h = fnv(h, &n, 4); h = fnv(h, e.draws.data(), n * sizeof(e.draws[0]));  // (the alpha-test variant follows from the draws' textures, fixed at create_image)
    }
    return h;
}

// the compute stream waits for the uploads and read-backs issued so far
static void joinTransfers(Backend& b) {
    if (b.uploadsPending) {
        cudaEventRecord(b.uploadsDone, b.uploadStream);
        cudaStreamWaitEvent(b.stream, b.uploadsDone, 0);
        b.uploadsPending = false;
    }
}
// a submission that references an image with a read-back in flight waits for that read-back (write-after-read)
static void orderAfterDownload(Backend& b, DeviceImage* img) {
    if (img && img->downloadPending) {
        cudaStreamWaitEvent(b.stream, img->downloadDone, 0);
        img->downloadPending = false;
    }
}
// stream for an asynchronous copy of `img`, ordered against the passes and the other copy direction
static cudaStream_t transferStream(Backend& b, DeviceImage& img, bool toDevice) {
    if (toDevice) {
        if (img.lastUsedSubmission >= 0) {
            // the ring slot holds this submission's event or, once overwritten, a later one: either completes after it
            cudaStreamWaitEvent(b.uploadStream, b.submissionDone[img.lastUsedSubmission % Backend::kSubmissionRing], 0);
        }
        if (img.downloadPending) cudaStreamWaitEvent(b.uploadStream, img.downloadDone, 0);
        b.uploadsPending = true;
        return b.uploadStream;
    }
    cudaEventRecord(b.computeMark, b.stream);
    cudaStreamWaitEvent(b.downloadStream, b.computeMark, 0);
    if (img.deferredExchange && b.peerDeferredPending) cudaStreamWaitEvent(b.downloadStream, b.peerDeferredDone, 0);  // rows the peers pushed behind the frame
    if (b.uploadsPending) {  // a read-back of an image that is being uploaded sees the upload
        cudaEventRecord(b.uploadsDone, b.uploadStream);
        cudaStreamWaitEvent(b.downloadStream, b.uploadsDone, 0);
    }
    return b.downloadStream;
}
static void markDownloaded(Backend& b, DeviceImage& img) {
    if (!img.downloadDone) cudaEventCreateWithFlags(&img.downloadDone, cudaEventDisableTiming);
    cudaEventRecord(img.downloadDone, b.downloadStream);
    img.downloadPending = true;
}
// blocking operations on the compute stream first drain the transfer streams
static void drainTransfers(Backend& b) {
    cudaStreamSynchronize(b.uploadStream);
    cudaStreamSynchronize(b.downloadStream);
    joinTransfers(b);
}

// hazards between the passes of a submission, from their declared resources: storage images and writable storage buffers are
// read-write, sampled images and read-only storage buffers are reads (uniform buffers change only through the fills, which
// land before the first pass). deps[i] = earlier passes that pass i must wait for (RAW, WAR, WAW).
static void passDependencies(Backend& b, std::vector<std::vector<int>>& deps) {
    struct Use { int lastWriter = -1; std::vector<int> readers; };
    std::unordered_map<uint64_t, Use> uses;
    auto imageKey = [](const plain_image_resource& r) { return ((uint64_t)(r.image.type & 3u) << 60) | ((uint64_t)r.image.index << 8) | (uint64_t)(r.mip_level & 0xffu); };
    deps.assign(b.execs.size(), {});
    for (size_t i = 0; i < b.execs.size(); i++) {
        const ExecRecord& e = b.execs[i];
        std::vector<int>& d = deps[i];
        auto read = [&](uint64_t key) {
            Use& u = uses[key];
            if (u.lastWriter >= 0 && u.lastWriter != (int)i) d.push_back(u.lastWriter);
            u.readers.push_back((int)i);
        };
        auto write = [&](uint64_t key) {
            Use& u = uses[key];
            if (u.lastWriter >= 0 && u.lastWriter != (int)i) d.push_back(u.lastWriter);  // a pass may bind one resource twice
            for (int r : u.readers) if (r != (int)i) d.push_back(r);
            u.readers.clear();
            u.lastWriter = (int)i;
        };
        for (auto& r : e.sampledImages) read(imageKey(r));
        for (auto& r : e.storageBuffers) if (r.read_only) read((1ull << 63) | r.buffer);
        for (auto& r : e.storageImages) write(imageKey(r));
        for (auto& t : e.targets) write(imageKey(plain_image_resource{t.image, t.mip_level, 0}));  // attachments (a loaded depth attachment is read and written)
        for (auto& r : e.storageBuffers) if (!r.read_only) write((1ull << 63) | r.buffer);
        std::sort(d.begin(), d.end());
        d.erase(std::unique(d.begin(), d.end()), d.end());
    }
}

// Pass fusion. indirectLightUpscale.comp writes two full-resolution images (12 bytes per pixel) that gbufferShading.comp reads back at
// the pixel's own texel (triangle.frag:294-320): when both are in one submission, nothing else in it touches the two images and they
// have the colour target's extent, the shading kernel computes the upscaled texel itself (giUpscalePixel, rounded through binary16 like
// the store) and the upscale is not launched. The images then keep their previous contents: callers that read them back (tests,
// debugging) switch fusion off with set_pass_fusion_enabled(ctx, 0). The producer keeps its place in the dependency order, so the
// consumer still waits for everything the producer's inputs depend on, and for the same row window.
static void planFusions(Backend& b) {
    for (auto& e : b.execs) { e.fusedProducer = -1; e.fusedAway = false; e.fusedRun.clear(); }
    if (!b.fusionEnabled) return;
    // runs of consecutive bloomDownsample / bloomUpsample executions over WHOLE small levels (<= 600 K texels: mips >= 2 at 3840x2160;
    // Bloom.cpp:65-122 emits them back to back and each reads what the previous ones wrote): one persistent launch (passes_post.cu bloomTailKernel)
    for (size_t i = 0; i < b.execs.size();) {
        auto smallBloomLevel = [&](const ExecRecord& e) {
            if (b.passes[e.pass].graphic || (b.passes[e.pass].shader != "bloomDownsample.comp" && b.passes[e.pass].shader != "bloomUpsample.comp")) return false;
            if (e.rowBegin != 0 || e.rowEnd != 0 || e.storageImages.empty()) return false;
            DeviceImage* img = b.resolve(e.storageImages[0].image);
            if (!img || e.storageImages[0].mip_level >= img->mips.size()) return false;
            const MipInfo& m = img->mips[e.storageImages[0].mip_level];
            return (size_t)m.w * m.h <= 600000u && m.d == 1;
        };
        size_t j = i;
        while (j < b.execs.size() && smallBloomLevel(b.execs[j]) && j - i < 10) j++;
        if (j - i >= 2) {
            for (size_t k = i; k < j; k++) { b.execs[i].fusedRun.push_back((int)k); if (k > i) b.execs[k].fusedAway = true; }
            i = j;
        } else {
            i = j > i ? j : i + 1;
        }
    }
    auto same = [](const plain_image_resource& a, const plain_image_resource& c) { return a.image.type == c.image.type && a.image.index == c.image.index && a.mip_level == c.mip_level; };
    auto find = [](const std::vector<plain_image_resource>& v, uint32_t binding) -> const plain_image_resource* { for (auto& r : v) if (r.binding == binding) return &r; return nullptr; };
    for (size_t j = 0; j < b.execs.size(); j++) {
        ExecRecord& consumer = b.execs[j];
        if (b.passes[consumer.pass].graphic || b.passes[consumer.pass].shader != "gbufferShading.comp") continue;
        const plain_image_resource* y = find(consumer.sampledImages, 15), *cg = find(consumer.sampledImages, 16);
        if (!y || !cg) continue;
        int producer = -1;
        for (size_t i = 0; i < j; i++) {
            const ExecRecord& e = b.execs[i];
            if (b.passes[e.pass].graphic || b.passes[e.pass].shader != "indirectLightUpscale.comp") continue;
            const plain_image_resource* oy = find(e.storageImages, 0), *oc = find(e.storageImages, 1);
            if (oy && oc && same(*oy, *y) && same(*oc, *cg)) producer = (int)i;
        }
        if (producer < 0) continue;
        const ExecRecord& pe = b.execs[(size_t)producer];
        if (pe.rowBegin != consumer.rowBegin || pe.rowEnd != consumer.rowEnd) continue;  // both cover the same rows (row sharding: band + 8)
        bool othersTouch = false;
        for (size_t k = 0; k < b.execs.size() && !othersTouch; k++) {
            if ((int)k == producer || k == j) continue;
            for (auto& r : b.execs[k].sampledImages) othersTouch = othersTouch || same(r, *y) || same(r, *cg);
            for (auto& r : b.execs[k].storageImages) othersTouch = othersTouch || same(r, *y) || same(r, *cg);
        }
        if (othersTouch) continue;
        consumer.fusedProducer = producer;
        b.execs[(size_t)producer].fusedAway = true;
    }
    // The froxel chain (Volumetrics.cpp:151-246 emits froxelVolumeMaterial -> froxelLightScattering -> volumeLightingReprojection ->
    // volumetricLightingIntegration back to back, each reading its predecessor's volume at its own froxel only): ONE launch over froxel
    // columns (passes_volumetrics.cu froxelColumnKernel), issued in the place of the LAST execution - by then the stream has waited for
    // everything the four depend on. The material and scattering volumes are then not written (nothing else in the submission may touch them).
    static const bool fuseFroxels = !getenv("PLAIN_FROXEL_FUSION") || atoi(getenv("PLAIN_FROXEL_FUSION")) != 0;  // A / B switch: 0 = the four kernels
    for (size_t i = 0; fuseFroxels && i + 3 < b.execs.size(); i++) {
        static const char* const chain[4] = {"froxelVolumeMaterial.comp", "froxelLightScattering.comp", "volumeLightingReprojection.comp", "volumetricLightingIntegration.comp"};
        bool ok = true;
        for (size_t k = 0; k < 4 && ok; k++) {
            const ExecRecord& e = b.execs[i + k];
            ok = !b.passes[e.pass].graphic && b.passes[e.pass].shader == chain[k] && !e.fusedAway && e.fusedRun.empty() && e.fusedProducer < 0 &&
                 e.rowBegin == b.execs[i].rowBegin && e.rowEnd == b.execs[i].rowEnd;
        }
        if (!ok) continue;
        const ExecRecord &mat = b.execs[i], &sca = b.execs[i + 1], &rep = b.execs[i + 2], &itg = b.execs[i + 3];
        const plain_image_resource *matOut = find(mat.storageImages, 0), *scaOut = find(sca.storageImages, 0), *scaIn = find(sca.sampledImages, 2);
        const plain_image_resource *repOut = find(rep.storageImages, 0), *repIn = find(rep.sampledImages, 1), *repHistory = find(rep.sampledImages, 2);
        const plain_image_resource *itgOut = find(itg.storageImages, 0), *itgIn = find(itg.sampledImages, 1);
        if (!matOut || !scaOut || !scaIn || !repOut || !repIn || !repHistory || !itgOut || !itgIn) continue;
        if (!same(*matOut, *scaIn) || !same(*scaOut, *repIn) || !same(*repOut, *itgIn)) continue;
        const plain_image_resource* written[4] = {matOut, scaOut, repOut, itgOut};
        bool distinct = true;
        for (int a = 0; a < 4; a++) {
            distinct = distinct && !same(*written[a], *repHistory);  // the history is sampled at reprojected positions: it must not be a volume the launch writes
            for (int c = a + 1; c < 4; c++) distinct = distinct && !same(*written[a], *written[c]);
        }
        if (!distinct) continue;
        auto level = [&](const plain_image_resource& r) -> const MipInfo* {
            DeviceImage* img = b.resolve(r.image);
            return img && r.mip_level < img->mips.size() ? &img->mips[r.mip_level] : nullptr;
        };
        const MipInfo* ref = level(*repOut);
        if (!ref || ref->d > 128) continue;
        bool sameExtent = true;
        for (int a = 0; a < 4; a++) { const MipInfo* m = level(*written[a]); sameExtent = sameExtent && m && m->w == ref->w && m->h == ref->h && m->d == ref->d; }
        if (!sameExtent) continue;
        bool covered = (int)itg.dispatch[0] * 8 >= ref->w && (int)itg.dispatch[1] * 8 >= ref->h;
        for (size_t k = 0; k < 3; k++) covered = covered && (int)b.execs[i + k].dispatch[0] * 4 >= ref->w && (int)b.execs[i + k].dispatch[1] * 4 >= ref->h && (int)b.execs[i + k].dispatch[2] * 4 >= ref->d;
        if (!covered) continue;
        auto settingsBuffer = [](const ExecRecord& e, uint32_t binding) -> long long { for (auto& r : e.uniformBuffers) if (r.binding == binding) return (long long)r.buffer; return -1 - (long long)binding; };
        const long long settings = settingsBuffer(mat, 2);
        if (settings < 0 || settingsBuffer(sca, 5) != settings || settingsBuffer(rep, 3) != settings || settingsBuffer(itg, 2) != settings) continue;
        bool othersTouch = false;
        for (size_t k = 0; k < b.execs.size() && !othersTouch; k++) {
            if (k >= i && k < i + 4) continue;
            for (auto& r : b.execs[k].sampledImages) othersTouch = othersTouch || same(r, *matOut) || same(r, *scaOut);
            for (auto& r : b.execs[k].storageImages) othersTouch = othersTouch || same(r, *matOut) || same(r, *scaOut);
        }
        if (othersTouch) continue;
        for (size_t k = 0; k < 4; k++) b.execs[i + 3].fusedRun.push_back((int)(i + k));
        for (size_t k = 0; k < 3; k++) b.execs[i + k].fusedAway = true;
        i += 3;
    }
}

#define SCHED_CHECK(call, what)                                                                                         \
    do {                                                                                                                \
        cudaError_t e__ = (call);                                                                                       \
        if (e__ != cudaSuccess) { b.lastError = std::string("pass scheduling: ") + what + ": " + cudaGetErrorString(e__); return false; } \
    } while (0)
static bool runPasses(Backend& b, bool withTiming) {
    const bool concurrent = b.concurrentPasses && !withTiming && b.execs.size() > 1;
    std::vector<std::vector<int>> deps;
    std::vector<int> streamOf;
    cudaStream_t streams[1 + Backend::kSideStreams];
    int tail[1 + Backend::kSideStreams];
    bool forked[1 + Backend::kSideStreams];
    int nextSide = 0;
    if (concurrent) {
        passDependencies(b, deps);
        streamOf.assign(b.execs.size(), 0);
        streams[0] = b.stream;
        for (int k = 0; k < Backend::kSideStreams; k++) streams[1 + k] = b.sideStreams[k];
        for (int k = 0; k <= Backend::kSideStreams; k++) { tail[k] = -1; forked[k] = k == 0; }
        SCHED_CHECK(cudaEventRecord(b.forkEvent, b.stream), "fork record");  // everything enqueued before this submission (fills, uploads, exchanges)
    }
    size_t ev = 0;
    for (size_t i = 0; i < b.execs.size(); i++) {
        ExecRecord& e = b.execs[i];
        LaunchCtx c;
        c.be = &b;
        c.pass = &b.passes[e.pass];
        c.exec = &e;
        c.stream = b.stream;
        c.smCount = b.smCount;
        c.g = b.globalUniformBuffer < b.uniformBuffers.size() ? (const plain_global_shader_info*)b.uniformBuffers[b.globalUniformBuffer].ptr : nullptr;
        c.bindless = b.bindlessDevice;
        c.tables = (const float*)b.tablesDevice;
        if (!c.g) { b.lastError = "render_frame: no global uniform buffer bound (set_global_descriptor_set_resources)"; return false; }
        int s = 0;
        if (concurrent) {
            // the pass follows a pass it depends on when that pass is the tail of a stream (the pass stream first); a pass that
            // depends on no tail starts a parallel branch on a side stream
            const std::vector<int>& d = deps[i];
            auto dependsOn = [&](int p) { return p >= 0 && std::binary_search(d.begin(), d.end(), p); };
            if (i == 0 || dependsOn(tail[0])) s = 0;
            else {
                s = -1;
                for (int k = 1; k <= Backend::kSideStreams && s < 0; k++) if (dependsOn(tail[k])) s = k;
                if (s < 0) { s = 1 + nextSide; nextSide = (nextSide + 1) % Backend::kSideStreams; }
            }
            if (!forked[s]) { SCHED_CHECK(cudaStreamWaitEvent(streams[s], b.forkEvent, 0), "fork wait"); forked[s] = true; }
            for (int p : d) if (streamOf[(size_t)p] != s) SCHED_CHECK(cudaStreamWaitEvent(streams[s], b.passEvents[(size_t)p], 0), "dependency wait");  // same stream: ordered already
            streamOf[i] = s;
            tail[s] = (int)i;
            c.stream = streams[s];
        }
        if (withTiming) {
            while (b.timingEvents.size() < ev + 2) { cudaEvent_t x; cudaEventCreate(&x); b.timingEvents.push_back(x); }
            cudaEventRecord(b.timingEvents[ev], b.stream);
        }
        if (!e.fusedAway) c.pass->fn(c);  // a fused-away producer keeps its events: its consumer inherits its dependencies through them
        if (withTiming) { cudaEventRecord(b.timingEvents[ev + 1], b.stream); ev += 2; }
        if (concurrent) SCHED_CHECK(cudaEventRecord(b.passEvents[i], c.stream), "pass event record");
        if (c.failed) { b.lastError = c.error; if (concurrent) for (int k = 1; k <= Backend::kSideStreams; k++) if (forked[k] && tail[k] >= 0) cudaStreamWaitEvent(b.stream, b.passEvents[(size_t)tail[k]], 0); return false; }
    }
    if (concurrent)  // join: whatever follows on the pass stream (next submission, exchanges, read-backs) sees every pass
        for (int k = 1; k <= Backend::kSideStreams; k++) if (tail[k] >= 0) SCHED_CHECK(cudaStreamWaitEvent(b.stream, b.passEvents[(size_t)tail[k]], 0), "join");
    return true;
}

// ---------------- peer exchange kernels ----------------
struct PushSegment { const unsigned char* src; unsigned char* dst; unsigned long long bytes, sliceStride; unsigned int slices; };  // `bytes` contiguous bytes in each of `slices` slices
static const int kMaxPushSegments = 28;
struct PushArgs { PushSegment seg[kMaxPushSegments]; };
// blockIdx.y = segment; the blocks of a segment stride over it. Rows of an image level are contiguous, and all levels on
// the frame path have 16-byte aligned rows, so the copy is 128-bit loads from local HBM and 128-bit stores over NVLink.
__global__ void __launch_bounds__(256) peerPushKernel(const __grid_constant__ PushArgs a) {
    const PushSegment sg = a.seg[blockIdx.y];
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if ((((size_t)sg.src | (size_t)sg.dst | (size_t)sg.bytes | (size_t)sg.sliceStride) & 15) == 0) {
        const size_t perSlice = sg.bytes / 16, n = perSlice * sg.slices, strideVec = sg.sliceStride / 16;
        const uint4* s = (const uint4*)sg.src;
        uint4* d = (uint4*)sg.dst;
        for (size_t i = tid; i < n; i += stride) {
            const size_t slice = i / perSlice, off = slice * strideVec + (i - slice * perSlice);
            d[off] = s[off];
        }
    } else {
        const size_t n = (size_t)sg.bytes * sg.slices;
        for (size_t i = tid; i < n; i += stride) {
            const size_t slice = i / sg.bytes, off = slice * sg.sliceStride + (i - slice * sg.bytes);
            sg.dst[off] = sg.src[off];
        }
    }
}
struct BarrierArgs {
    uint32_t* peerFlags[PLAIN_MAX_PEERS];  // flags array of every rank's sync block
    uint32_t* localError;
    uint32_t rank, count, epoch;
    long long timeoutCycles;
};
// one warp: lane p signals rank p (release: the pushes enqueued before this kernel have completed, the fence orders them
// before the flag at system scope), then waits until rank p has signalled this rank
__global__ void peerBarrierKernel(const __grid_constant__ BarrierArgs a) {
    const uint32_t p = threadIdx.x;
    if (p < a.count && p != a.rank) {
        __threadfence_system();
        *(volatile uint32_t*)(a.peerFlags[p] + a.rank) = a.epoch;
        const volatile uint32_t* mine = a.peerFlags[a.rank] + p;
        const long long start = clock64();
        while ((int)(*mine - a.epoch) < 0) {
            if (clock64() - start > a.timeoutCycles) { atomicExch(a.localError, 1u); break; }
            __nanosleep(64);
        }
        __threadfence_system();
    }
}
struct ReduceArgs {
    uint32_t* peerScratch[PLAIN_MAX_PEERS];  // scratch[parity] of every rank's sync block
    uint32_t* buffer;
    uint32_t rank, count, n;
};
__global__ void peerReducePushKernel(const __grid_constant__ ReduceArgs a) {  // block p writes this rank's partials into rank p's scratch
    for (uint32_t i = threadIdx.x; i < a.n; i += blockDim.x) a.peerScratch[blockIdx.x][(size_t)a.rank * Backend::kPeerReduceMax + i] = a.buffer[i];
}
__global__ void peerReduceSumKernel(const __grid_constant__ ReduceArgs a) {  // ascending rank order: integer sums, deterministic
    for (uint32_t i = threadIdx.x; i < a.n; i += blockDim.x) {
        uint32_t sum = 0;
        for (uint32_t r = 0; r < a.count; r++) sum += a.peerScratch[a.rank][(size_t)r * Backend::kPeerReduceMax + i];
        a.buffer[i] = sum;
    }
}

}  // namespace pb

using namespace pb;

struct plain_ctx {
    Backend b;
};

static int fail(plain_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->b.lastError = msg;
    return 1;
}
#define CU_CHECK(ctx, call)                                                                            \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) return fail(ctx, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// the corner-replicated copies of the SDF bricks written since they were built (enqueued on the pass stream; no allocation: capturable)
static void refreshCornerBricks(Backend& b) {
    for (DeviceImage& img : b.images)
        if (img.corners && img.cornersStale) {
            buildCornerBrick(img.corners, img.ptr, img.mips[0].w, img.mips[0].h, img.mips[0].d, b.stream);
            img.cornersStale = false;
        }
}
static void updateBindless(Backend& b, uint32_t index) {
    if (index >= Backend::kMaxBindless) return;
    const DeviceImage& img = b.images[index];
    BindlessEntry e;
    e.view.ptr = img.ptr; e.view.w = img.mips[0].w; e.view.h = img.mips[0].h; e.view.d = img.mips[0].d;
    e.format = img.desc.format; e.pad = 0;
    e.corners = img.corners;
    cudaMemcpyAsync(b.bindlessDevice + index, &e, sizeof(e), cudaMemcpyHostToDevice, b.stream);
    cudaStreamSynchronize(b.stream);
}

extern "C" {
static int joinDeferredExchanges(plain_ctx* ctx, cudaStream_t stream);

int PLAIN_FN(backend_create)(int device, uint32_t width, uint32_t height, plain_ctx** out_ctx) {
    if (!out_ctx) return 1;
    *out_ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        fprintf(stderr, "plain_backend_create: CUDA device %d not available (%d devices) - this backend has no CPU path\n", device, count);
        return 1;
    }
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    plain_ctx* ctx = new plain_ctx();
    Backend& b = ctx->b;
    b.device = device;
    cudaDeviceGetAttribute(&b.smCount, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&b.uploadStream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&b.downloadStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return 1; }
    cudaEventCreateWithFlags(&b.stagingConsumed, cudaEventDisableTiming);
    for (auto& e : b.submissionDone) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& st : b.sideStreams) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&b.forkEvent, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&b.uploadsDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&b.computeMark, cudaEventDisableTiming);
    if (cudaMallocHost(&b.stagingHost, kStagingBytes) != cudaSuccess || cudaMalloc(&b.stagingDevice, kStagingBytes) != cudaSuccess ||
        cudaMalloc(&b.bindlessDevice, sizeof(BindlessEntry) * Backend::kMaxBindless) != cudaSuccess || cudaMalloc(&b.tablesDevice, sizeof(ShadingTables)) != cudaSuccess) {
        fprintf(stderr, "plain_backend_create: allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return 1;
    }
    cudaMemsetAsync(b.bindlessDevice, 0, sizeof(BindlessEntry) * Backend::kMaxBindless, b.stream);
    buildShadingTables(b.tablesDevice, b.stream);
    if (cudaMalloc(&b.fusionCounters, 64 * sizeof(unsigned int)) != cudaSuccess) { delete ctx; return 1; }
    cudaMemsetAsync(b.fusionCounters, 0, 64 * sizeof(unsigned int), b.stream);
    plain_image_desc d{};
    d.width = width; d.height = height; d.depth = 1;
    d.type = PLAIN_IMAGE_TYPE_2D; d.format = PLAIN_FORMAT_BGRA8_UNORM;  // VulkanSurface.cpp:41-46
    d.usage_flags = PLAIN_USAGE_STORAGE; d.mip_count = PLAIN_MIPS_ONE;
    for (auto& sc : b.swapchainImages) if (!allocateImage(b, sc, d)) { delete ctx; return 1; }
    *out_ctx = ctx;
    return 0;
}
void PLAIN_FN(backend_destroy)(plain_ctx* ctx) {
    if (!ctx) return;
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    cudaStreamSynchronize(b.uploadStream);
    cudaStreamSynchronize(b.downloadStream);
    cudaStreamSynchronize(b.stream);
    if (b.peerStream) cudaStreamSynchronize(b.peerStream);
    for (auto& g : b.graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    for (uint32_t p = 0; p < b.peerCount; p++) {
        for (auto& i : b.images) if (i.peerPtr[p]) cudaIpcCloseMemHandle(i.peerPtr[p]);
        for (auto& i : b.transientImages) if (i.peerPtr[p]) cudaIpcCloseMemHandle(i.peerPtr[p]);
        if (p != b.peerRank && b.peerSync[p]) cudaIpcCloseMemHandle(b.peerSync[p]);
    }
    if (b.peerCount) cudaFree(b.peerSync[b.peerRank]);
    for (auto& i : b.images) { cudaFree(i.ptr); cudaFree(i.corners); }
    for (auto& i : b.transientImages) cudaFree(i.ptr);
    for (auto& sc : b.swapchainImages) { cudaFree(sc.ptr); if (sc.downloadDone) cudaEventDestroy(sc.downloadDone); }
    for (auto& i : b.images) if (i.downloadDone) cudaEventDestroy(i.downloadDone);
    for (auto& i : b.transientImages) if (i.downloadDone) cudaEventDestroy(i.downloadDone);
    for (auto& m : b.meshes) { cudaFree(m.indices); cudaFree(m.vertices); }
    for (auto& kv : b.rasterScratch) { cudaFree(kv.second.draws); cudaFree(kv.second.triInfo); cudaFree(kv.second.vertexCache); }
    for (auto& kv : b.visBuffers) cudaFree(kv.second.ptr);
    for (auto& u : b.uniformBuffers) cudaFree(u.ptr);
    for (auto& s : b.storageBuffers) cudaFree(s.ptr);
    for (auto& e : b.timingEvents) cudaEventDestroy(e);
    cudaFree(b.stagingDevice);
    cudaFreeHost(b.stagingHost);
    cudaFree(b.bindlessDevice);
    if (b.peerErrorHost) cudaFreeHost(b.peerErrorHost);
    if (b.peerStream) { cudaStreamDestroy(b.peerStream); cudaEventDestroy(b.peerFork); cudaEventDestroy(b.peerDeferredDone); }
    cudaFree(b.tablesDevice);
    cudaFree(b.fusionCounters);
    cudaEventDestroy(b.stagingConsumed);
    for (auto& e : b.submissionDone) cudaEventDestroy(e);
    for (auto& st : b.sideStreams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto& e : b.passEvents) cudaEventDestroy(e);
    cudaEventDestroy(b.forkEvent);
    cudaEventDestroy(b.uploadsDone);
    cudaEventDestroy(b.computeMark);
    cudaStreamDestroy(b.uploadStream);
    cudaStreamDestroy(b.downloadStream);
    cudaStreamDestroy(b.stream);
    delete ctx;
}
const char* PLAIN_FN(last_error)(plain_ctx* ctx) { return ctx ? ctx->b.lastError.c_str() : "null context"; }
int PLAIN_FN(recreate_swapchain)(plain_ctx* ctx, uint32_t width, uint32_t height) {
    plain_image_desc d = ctx->b.swapchainImages[0].desc;
    d.width = width; d.height = height;
    for (auto& sc : ctx->b.swapchainImages) if (!allocateImage(ctx->b, sc, d)) return fail(ctx, "recreate_swapchain: allocation failed");
    return 0;
}

int PLAIN_FN(create_image)(plain_ctx* ctx, const plain_image_desc* desc, const void* initial_data, size_t initial_data_size, plain_image_handle* out) {
    if (!desc || !out) return fail(ctx, "create_image: null argument");
    if (formatBytesPerTexel(desc->format) == 0) return fail(ctx, "create_image: format not supported on the frame path");
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceImage img;
    if (!allocateImage(b, img, *desc)) return fail(ctx, std::string("create_image: cudaMalloc failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (initial_data) {
        size_t off = 0;
        const size_t levels = desc->mip_count == PLAIN_MIPS_FULL_CHAIN_ALREADY_IN_DATA ? img.mips.size() : 1;
        for (size_t i = 0; i < levels; i++) {
            if (off + img.mips[i].bytes > initial_data_size) { cudaFree(img.ptr); return fail(ctx, "create_image: initial data too small"); }
            CU_CHECK(ctx, cudaMemcpyAsync(img.ptr + img.mips[i].offset, (const uint8_t*)initial_data + off, img.mips[i].bytes, cudaMemcpyHostToDevice, b.stream));
            off += img.mips[i].bytes;
        }
        CU_CHECK(ctx, cudaStreamSynchronize(b.stream));  // the caller's memory may go away after the call (RenderBackend.cpp:315-321)
        img.cornersStale = img.corners != nullptr;
        if (desc->format == PLAIN_FORMAT_RGBA8)
            for (size_t t = 3; t < img.mips[0].bytes; t += 4) if (((const uint8_t*)initial_data)[t] != 255) { img.transparentTexels = true; break; }
    }
    b.images.push_back(std::move(img));
    out->type = PLAIN_IMAGE_HANDLE_DEFAULT;
    out->index = (uint32_t)b.images.size() - 1;
    updateBindless(b, out->index);
    return 0;
}
int PLAIN_FN(create_temporary_image)(plain_ctx* ctx, const plain_image_desc* desc, plain_image_handle* out) {
    // valid for one frame; the allocation is reused across frames when the description matches (RenderBackend.cpp:1026-1123)
    Backend& b = ctx->b;
    for (size_t i = 0; i < b.transientImages.size(); i++) {
        DeviceImage& t = b.transientImages[i];
        if (!t.inUse && memcmp(&t.desc, desc, sizeof(*desc)) == 0) {
            t.inUse = true;
            out->type = PLAIN_IMAGE_HANDLE_TRANSIENT; out->index = (uint32_t)i;
            return 0;
        }
    }
    if (formatBytesPerTexel(desc->format) == 0) return fail(ctx, "create_temporary_image: format not supported");
    cudaSetDevice(b.device);
    DeviceImage img;
    if (!allocateImage(b, img, *desc)) return fail(ctx, "create_temporary_image: cudaMalloc failed");
    img.inUse = true;
    b.transientImages.push_back(std::move(img));
    out->type = PLAIN_IMAGE_HANDLE_TRANSIENT;
    out->index = (uint32_t)b.transientImages.size() - 1;
    return 0;
}
int PLAIN_FN(resize_images)(plain_ctx* ctx, const plain_image_handle* images, uint32_t n, uint32_t width, uint32_t height) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    for (uint32_t i = 0; i < n; i++) {
        DeviceImage* img = b.resolve(images[i]);
        if (!img) return fail(ctx, "resize_images: invalid handle");
        plain_image_desc d = img->desc;
        d.width = width; d.height = height;
        if (!allocateImage(b, *img, d)) return fail(ctx, "resize_images: cudaMalloc failed");
        if (images[i].type == PLAIN_IMAGE_HANDLE_DEFAULT) updateBindless(b, images[i].index);
    }
    return 0;
}
int PLAIN_FN(get_image_description)(plain_ctx* ctx, plain_image_handle image, plain_image_desc* out) {
    DeviceImage* img = ctx->b.resolve(image);
    if (!img) return fail(ctx, "get_image_description: invalid handle");
    *out = img->desc;
    return 0;
}
int PLAIN_FN(get_image_global_texture_array_index)(plain_ctx* ctx, plain_image_handle image, uint32_t* out) {
    if (image.type != PLAIN_IMAGE_HANDLE_DEFAULT || image.index >= ctx->b.images.size() || image.index >= Backend::kMaxBindless) return fail(ctx, "global texture index: invalid handle");
    *out = image.index;
    return 0;
}
static int createBuffer(plain_ctx* ctx, std::vector<DeviceBuffer>& table, size_t size, const void* initial_data, plain_handle* out) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceBuffer buf;
    buf.size = size;
    CU_CHECK(ctx, cudaMalloc(&buf.ptr, size ? size : 4));
    CU_CHECK(ctx, cudaMemsetAsync(buf.ptr, 0, size ? size : 4, b.stream));
    if (initial_data) {
        CU_CHECK(ctx, cudaMemcpyAsync(buf.ptr, initial_data, size, cudaMemcpyHostToDevice, b.stream));
        CU_CHECK(ctx, cudaStreamSynchronize(b.stream));
    }
    table.push_back(buf);
    *out = (uint32_t)table.size() - 1;
    b.passEpoch++;
    return 0;
}
int PLAIN_FN(create_uniform_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out) { return createBuffer(ctx, ctx->b.uniformBuffers, size, initial_data, out); }
int PLAIN_FN(create_storage_buffer)(plain_ctx* ctx, size_t size, const void* initial_data, plain_handle* out) { return createBuffer(ctx, ctx->b.storageBuffers, size, initial_data, out); }
int PLAIN_FN(create_sampler)(plain_ctx* ctx, const plain_sampler_desc* desc, plain_handle* out) {
    // samplers are compile-time policy of each kernel (the 8 immutable samplers of global.inc:35-42); the table only keeps handles valid
    ctx->b.samplers.push_back(*desc);
    *out = (uint32_t)ctx->b.samplers.size() - 1;
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
    auto it = registry().find(p.shader);
    if (it == registry().end()) return fail(ctx, std::string("no CUDA kernel registered for shader '") + shader + "'");
    p.fn = it->second;
    ctx->b.passEpoch++;
    return 0;
}
int PLAIN_FN(create_compute_pass)(plain_ctx* ctx, const char* shader, const plain_spec_const* consts, uint32_t n_consts, const char* debug_name, plain_handle* out) {
    PassRecord p;
    if (fillPass(ctx, p, shader, consts, n_consts)) return 1;
    p.name = debug_name ? debug_name : shader;
    ctx->b.passes.push_back(std::move(p));
    *out = (uint32_t)ctx->b.passes.size() - 1;
    return 0;
}
int PLAIN_FN(update_compute_pass_shader_description)(plain_ctx* ctx, plain_handle pass, const char* shader, const plain_spec_const* consts, uint32_t n_consts) {
    if (pass >= ctx->b.passes.size()) return fail(ctx, "update pass: invalid handle");
    return fillPass(ctx, ctx->b.passes[pass], shader, consts, n_consts);
}
int PLAIN_FN(set_global_descriptor_set_resources)(plain_ctx* ctx, const plain_pass_resources* r) {
    for (uint32_t i = 0; i < r->n_uniform_buffers; i++)
        if (r->uniform_buffers[i].binding == 0) ctx->b.globalUniformBuffer = r->uniform_buffers[i].buffer;
    return 0;
}

int PLAIN_FN(new_frame)(plain_ctx* ctx) {
    ctx->b.execs.clear();
    ctx->b.launchCounter = 0;  // kernels of the frame: all submissions + the peer exchange kernels between them
    ctx->b.timings.clear();  // pass timings accumulate over the submissions of a frame (a row-sharded frame has one per segment)
    ctx->b.swapchainCurrent ^= 1;  // the next presentable image (RenderBackend.cpp:608-612 getSwapchainInputImage)
    ctx->b.frameFresh = true;
    ctx->b.peerBarriersThisFrame = 0;
    for (auto& t : ctx->b.transientImages) t.inUse = false;
    return 0;
}
int PLAIN_FN(set_compute_pass_execution)(plain_ctx* ctx, const plain_compute_pass_execution* e) {
    if (e->pass >= ctx->b.passes.size()) return fail(ctx, "set_compute_pass_execution: invalid pass handle");
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
    r.rowBegin = e->row_begin; r.rowEnd = e->row_end; r.shardPhase = e->shard_phase;
    ctx->b.execs.push_back(std::move(r));
    return 0;
}
int PLAIN_FN(prepare_for_drawcall_recording)(plain_ctx* ctx) { (void)ctx; return 0; }

// ---- meshes and graphic passes (RenderBackend.h:57-96) ----
int PLAIN_FN(create_meshes)(plain_ctx* ctx, const plain_mesh_binary* meshes, uint32_t n, plain_handle* out) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    for (uint32_t i = 0; i < n; i++) {
        const plain_mesh_binary& m = meshes[i];
        if (m.index_count % 3 != 0 || !m.index_buffer || !m.vertex_buffer) return fail(ctx, "create_meshes: triangle list with index and vertex data expected");
        DeviceMesh d;
        d.indexCount = m.index_count; d.vertexCount = m.vertex_count;
        d.index32 = m.index_count < 65535u ? 0u : 1u;
        const size_t ib = (size_t)m.index_count * (d.index32 ? 4 : 2), vb = (size_t)m.vertex_count * 28;
        for (uint32_t k = 0; k < m.index_count; k++) {
            const uint32_t idx = d.index32 ? ((const uint32_t*)m.index_buffer)[k] : ((const uint16_t*)m.index_buffer)[k];
            if (idx >= m.vertex_count) return fail(ctx, "create_meshes: index out of range");
        }
        CU_CHECK(ctx, cudaMalloc(&d.indices, ib ? ib : 4));
        CU_CHECK(ctx, cudaMalloc(&d.vertices, vb ? vb : 4));
        CU_CHECK(ctx, cudaMemcpy(d.indices, m.index_buffer, ib, cudaMemcpyHostToDevice));
        CU_CHECK(ctx, cudaMemcpy(d.vertices, m.vertex_buffer, vb, cudaMemcpyHostToDevice));
        b.meshes.push_back(d);
        out[i] = (uint32_t)b.meshes.size() - 1;
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
    p.depthAttachment = PLAIN_INVALID_INDEX;
    for (uint32_t i = 0; i < d->n_attachments; i++)
        if (d->attachments[i].format == PLAIN_FORMAT_DEPTH32 || d->attachments[i].format == PLAIN_FORMAT_DEPTH16) p.depthAttachment = i;
    if (p.depthAttachment == PLAIN_INVALID_INDEX) return fail(ctx, "create_graphic_pass: a depth attachment is required (the rasteriser resolves visibility through it)");
    auto it = registry().find(p.shader);
    if (it == registry().end()) return fail(ctx, std::string("no CUDA rasteriser program registered for the shader pair '") + p.shader + "'");
    p.fn = it->second;
    ctx->b.passes.push_back(std::move(p));
    ctx->b.passEpoch++;
    *out = (uint32_t)ctx->b.passes.size() - 1;
    return 0;
}
int PLAIN_FN(set_graphic_pass_execution)(plain_ctx* ctx, const plain_graphic_pass_execution* e) {
    Backend& b = ctx->b;
    if (e->pass >= b.passes.size() || !b.passes[e->pass].graphic) return fail(ctx, "set_graphic_pass_execution: not a graphic pass");
    if (e->n_targets != b.passes[e->pass].attachments.size()) return fail(ctx, "set_graphic_pass_execution: one target per attachment expected");
    ExecRecord r;
    r.pass = e->pass;
    const plain_pass_resources& s = e->resources;
    r.storageBuffers.assign(s.storage_buffers, s.storage_buffers + s.n_storage_buffers);
    r.uniformBuffers.assign(s.uniform_buffers, s.uniform_buffers + s.n_uniform_buffers);
    r.sampledImages.assign(s.sampled_images, s.sampled_images + s.n_sampled_images);
    r.storageImages.assign(s.storage_images, s.storage_images + s.n_storage_images);
    r.targets.assign(e->targets, e->targets + e->n_targets);
    r.dispatch[0] = r.dispatch[1] = r.dispatch[2] = 0;
    r.rowBegin = e->row_begin; r.rowEnd = e->row_end;
    b.execs.push_back(std::move(r));
    return 0;
}
int PLAIN_FN(draw_meshes)(plain_ctx* ctx, const plain_handle* meshes, uint32_t n, const void* push_constants, plain_handle pass, int32_t worker_index) {
    (void)worker_index;
    Backend& b = ctx->b;
    if (pass >= b.passes.size() || !b.passes[pass].graphic) return fail(ctx, "draw_meshes: not a graphic pass");
    ExecRecord* rec = nullptr;
    for (auto& e : b.execs) if (e.pass == pass) rec = &e;
    if (!rec) return fail(ctx, "draw_meshes: the pass has no execution this frame (set_graphic_pass_execution)");
    const uint32_t ps = b.passes[pass].pushSize;
    for (uint32_t i = 0; i < n; i++) {
        if (meshes[i] >= b.meshes.size()) return fail(ctx, "draw_meshes: invalid mesh handle");
        DrawRecord d;
        d.mesh = meshes[i];
        memset(d.push, 0, sizeof(d.push));
        memcpy(d.push, (const uint8_t*)push_constants + (size_t)i * ps, ps);
        rec->draws.push_back(d);
    }
    return 0;
}
// draw tables, per-primitive scratch and visibility buffers of the graphic passes of this submission: allocated and uploaded
// before the passes run (nothing may allocate while the pass list is captured into a graph)
static int prepareRaster(plain_ctx* ctx) {
    Backend& b = ctx->b;
    for (auto& e : b.execs) {
        const PassRecord& p = b.passes[e.pass];
        if (!p.graphic) continue;
        Backend::RasterScratch& sc = b.rasterScratch[e.pass];
        b.rasterDrawStaging.clear();
        uint32_t first = 0, firstVertex = 0;
        bool anyAlpha = false;
        for (auto& d : e.draws) {
            const DeviceMesh& m = b.meshes[d.mesh];
            RasterDraw rd;
            rd.indices = m.indices; rd.vertices = m.vertices;
            rd.firstPrimitive = first; rd.triCount = m.indexCount / 3; rd.index32 = m.index32;
            rd.firstVertex = firstVertex; rd.vertexCount = m.vertexCount;
            memcpy(rd.push, d.push, 16);
            rd.alphaTest = (rd.push[0] < b.images.size() && b.images[rd.push[0]].desc.format == PLAIN_FORMAT_RGBA8 && b.images[rd.push[0]].transparentTexels) ? 1u : 0u;
            anyAlpha = anyAlpha || rd.alphaTest != 0u;
            b.rasterDrawStaging.push_back(rd);
            first += rd.triCount;
            firstVertex += rd.vertexCount;
        }
        if ((size_t)firstVertex > sc.vertexCapacity) {
            if (sc.vertexCache) cudaFree(sc.vertexCache);
            sc.vertexCapacity = (size_t)firstVertex * 2 + 1024;
            CU_CHECK(ctx, cudaMalloc(&sc.vertexCache, sc.vertexCapacity * PLAIN_RASTER_VERTEX_FLOATS * sizeof(float)));
            b.passEpoch++;
        }
        if (b.rasterDrawStaging.size() > sc.drawCapacity) {
            if (sc.draws) cudaFree(sc.draws);
            sc.drawCapacity = b.rasterDrawStaging.size() * 2 + 16;
            CU_CHECK(ctx, cudaMalloc(&sc.draws, sc.drawCapacity * sizeof(RasterDraw)));
            b.passEpoch++;
        }
        const size_t triWords = (size_t)first * 2 + 2;  // rows per primitive, the big-triangle counter, the big-triangle list
        if (triWords > sc.triCapacity) {
            if (sc.triInfo) cudaFree(sc.triInfo);
            sc.triCapacity = triWords * 2 + 1024;
            CU_CHECK(ctx, cudaMalloc(&sc.triInfo, sc.triCapacity * sizeof(uint32_t)));
            b.passEpoch++;
        }
        if (!b.rasterDrawStaging.empty())  // pageable source: the copy has left the host vector when the call returns
            CU_CHECK(ctx, cudaMemcpyAsync(sc.draws, b.rasterDrawStaging.data(), b.rasterDrawStaging.size() * sizeof(RasterDraw), cudaMemcpyHostToDevice, b.stream));
        const plain_render_target& dt = e.targets[p.depthAttachment];
        DeviceImage* img = b.resolve(dt.image);
        if (!img || dt.mip_level >= img->mips.size()) return fail(ctx, p.shader + ": invalid depth target");
        const size_t texels = (size_t)img->mips[dt.mip_level].w * img->mips[dt.mip_level].h;
        Backend::VisBuffer& vb = b.visBuffers[((uint64_t)(dt.image.type & 3u) << 60) | ((uint64_t)dt.image.index << 8) | (uint64_t)(dt.mip_level & 0xffu)];
        if (vb.texels != texels) {
            if (vb.ptr) cudaFree(vb.ptr);
            CU_CHECK(ctx, cudaMalloc(&vb.ptr, (texels ? texels : 1) * sizeof(unsigned long long)));
            CU_CHECK(ctx, cudaMemsetAsync(vb.ptr, 0, (texels ? texels : 1) * sizeof(unsigned long long), b.stream));
            vb.texels = texels;
            b.passEpoch++;
        }
        e.rasterDraws = sc.draws;
        e.rasterTriInfo = sc.triInfo;
        e.rasterVis = vb.ptr;
        e.rasterTotalTris = first;
        e.rasterVertexCache = sc.vertexCache;
        e.rasterTotalVertices = firstVertex;
        e.rasterAnyAlphaTest = anyAlpha;
    }
    return 0;
}

static int stageFill(plain_ctx* ctx, bool uniform, plain_handle buffer, const void* data, size_t size) {
    Backend& b = ctx->b;
    std::vector<DeviceBuffer>& table = uniform ? b.uniformBuffers : b.storageBuffers;
    if (buffer >= table.size() || size > table[buffer].size) return fail(ctx, "set_buffer_data: invalid buffer/size");
    if (b.stagingInFlight) {  // the previous frame's staged copy must have left the pinned block before it is rewritten
        cudaEventSynchronize(b.stagingConsumed);
        b.stagingInFlight = false;
    }
    const size_t aligned = (size + 15) & ~(size_t)15;
    if (b.fills.size() >= kMaxFillSegments || b.stagingUsed + aligned > kStagingBytes) return fail(ctx, "set_buffer_data: staging block full (too many / too large fills in one frame)");
    memcpy(b.stagingHost + b.stagingUsed, data, size);
    b.fills.push_back(FillOrder{uniform, buffer, b.stagingUsed, size});
    b.stagingUsed += aligned;
    return 0;
}
int PLAIN_FN(set_uniform_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size) { return stageFill(ctx, true, buffer, data, size); }
int PLAIN_FN(set_storage_buffer_data)(plain_ctx* ctx, plain_handle buffer, const void* data, size_t size) { return stageFill(ctx, false, buffer, data, size); }

int PLAIN_FN(render_frame)(plain_ctx* ctx, int present) {
    (void)present;
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    joinTransfers(b);  // uploads issued before this submission are visible to its passes; read-backs in flight keep their source
    if (b.frameFresh) {  // the previous frame's deferred exchanges (next-frame data) have landed before this frame's first pass
        b.frameFresh = false;
        if (joinDeferredExchanges(ctx, b.stream)) return 1;
        b.peerDeferredPending = false;
    }
    refreshCornerBricks(b);  // SDF bricks written since the last submission (direct launches, outside the captured graph)
    while (b.passEvents.size() < b.execs.size()) { cudaEvent_t x; cudaEventCreateWithFlags(&x, cudaEventDisableTiming); b.passEvents.push_back(x); }
    if (prepareRaster(ctx)) return 1;
    for (auto& e : b.execs) {
        for (auto& t : e.targets) if (DeviceImage* img = b.resolve(t.image)) { img->lastUsedSubmission = b.submissionCounter; orderAfterDownload(b, img); }
        for (auto& r : e.sampledImages) if (DeviceImage* img = b.resolve(r.image)) img->lastUsedSubmission = b.submissionCounter;
        for (auto& r : e.storageImages) if (DeviceImage* img = b.resolve(r.image)) {
            img->lastUsedSubmission = b.submissionCounter; orderAfterDownload(b, img);
            if (img->corners) img->cornersStale = true;  // a pass writes an SDF brick: its corner copy is rebuilt before the NEXT submission
        }
    }
    // all fills of the frame land before any pass (RenderBackend.cpp:896-911): one H2D copy + one scatter kernel
    if (!b.fills.empty()) {
        FillSegment* table = (FillSegment*)b.stagingHost;
        for (size_t i = 0; i < b.fills.size(); i++) {
            const FillOrder& f = b.fills[i];
            DeviceBuffer& dst = f.uniform ? b.uniformBuffers[f.buffer] : b.storageBuffers[f.buffer];
            table[i].dst = (unsigned long long)dst.ptr;
            table[i].srcOffset = (uint32_t)f.offsetInStaging;
            table[i].size = (uint32_t)f.size;
        }
        CU_CHECK(ctx, cudaMemcpyAsync(b.stagingDevice, b.stagingHost, b.stagingUsed, cudaMemcpyHostToDevice, b.stream));
        cudaEventRecord(b.stagingConsumed, b.stream);
        b.stagingInFlight = true;
        scatterFillsKernel<<<(unsigned)b.fills.size(), 128, 0, b.stream>>>(b.stagingDevice, (int)b.fills.size());
        b.launchCounter++;
        b.fills.clear();
        b.stagingUsed = kStagingHeader;
    }
    const uint32_t fillLaunches = b.launchCounter;
    planFusions(b);
    if (b.timingEnabled) {
        if (!runPasses(b, true)) return 1;
        CU_CHECK(ctx, cudaStreamSynchronize(b.stream));
        for (size_t i = 0; i < b.execs.size(); i++) {
            plain_pass_time pt;
            snprintf(pt.name, sizeof(pt.name), "%s", b.passes[b.execs[i].pass].name.c_str());
            cudaEventElapsedTime(&pt.time_ms, b.timingEvents[2 * i], b.timingEvents[2 * i + 1]);
            b.timings.push_back(pt);
        }
    } else if (b.graphEnabled) {
        const uint64_t key = hashExecs(b);
        auto it = b.graphs.find(key);
        if (it == b.graphs.end()) {
            cudaGraph_t graph = nullptr;
            CU_CHECK(ctx, cudaStreamBeginCapture(b.stream, cudaStreamCaptureModeThreadLocal));
            const bool ok = runPasses(b, false);
            cudaError_t ce = cudaStreamEndCapture(b.stream, &graph);
            if (!ok) { if (graph) cudaGraphDestroy(graph); return 1; }
            if (ce != cudaSuccess) return fail(ctx, std::string("graph capture failed: ") + cudaGetErrorString(ce));
            CachedGraph cg;
            cg.launches = b.launchCounter - fillLaunches;
            ce = cudaGraphInstantiate(&cg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(ctx, std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ce));
            // a frame-varying push constant or draw list makes a new key every frame: keep the cache bounded (evict the least recently used)
            if (b.graphs.size() >= Backend::kMaxCachedGraphs) {
                auto victim = b.graphs.begin();
                for (auto g = b.graphs.begin(); g != b.graphs.end(); ++g) if (g->second.lastUse < victim->second.lastUse) victim = g;
                if (victim->second.exec) cudaGraphExecDestroy(victim->second.exec);
                b.graphs.erase(victim);
            }
            it = b.graphs.emplace(key, cg).first;
        }
        it->second.lastUse = ++b.graphUseCounter;
        CU_CHECK(ctx, cudaGraphLaunch(it->second.exec, b.stream));
        b.launchCounter = fillLaunches + it->second.launches;
    } else {
        if (!runPasses(b, false)) return 1;
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(ctx, std::string("render_frame: ") + cudaGetErrorString(e));
    b.lastFrameLaunches = b.launchCounter;
    cudaEventRecord(b.submissionDone[b.submissionCounter % Backend::kSubmissionRing], b.stream);
    b.submissionCounter++;
    return 0;
}
int PLAIN_FN(submit_recorded_passes)(plain_ctx* ctx) {
    const int rc = PLAIN_FN(render_frame)(ctx, 0);
    ctx->b.execs.clear();
    return rc;
}
int PLAIN_FN(wait_for_gpu_idle)(plain_ctx* ctx) {
    cudaSetDevice(ctx->b.device);
    if (joinDeferredExchanges(ctx, ctx->b.stream)) return 1;  // a read of an exchanged image after this call sees the peers' rows
    ctx->b.peerDeferredPending = false;
    drainTransfers(ctx->b);
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->b.stream));
    if (ctx->b.fusionCounters) {  // the grid barrier of a fused run gave up (its blocks were not co-resident for 50 ms): frames since the last check are invalid
        unsigned int fusionError = 0;
        CU_CHECK(ctx, cudaMemcpy(&fusionError, ctx->b.fusionCounters + 63, sizeof(fusionError), cudaMemcpyDeviceToHost));
        if (fusionError) {
            cudaMemset(ctx->b.fusionCounters + 63, 0, sizeof(fusionError));
            return fail(ctx, "wait_for_gpu_idle: the grid barrier of a fused pass run (bloom mips >= 2) timed out; disable it with set_pass_fusion_enabled(ctx, 0)");
        }
    }
    return 0;
}
int PLAIN_FN(join_transfers)(plain_ctx* ctx) {  // the compute stream waits for every asynchronous upload and read-back issued so far
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    joinTransfers(b);
    cudaEventRecord(b.computeMark, b.downloadStream);
    cudaStreamWaitEvent(b.stream, b.computeMark, 0);
    return 0;
}
int PLAIN_FN(get_renderpass_timings)(plain_ctx* ctx, plain_pass_time* out, uint32_t capacity, uint32_t* out_count) {
    const uint32_t n = (uint32_t)ctx->b.timings.size();
    if (out_count) *out_count = n;
    for (uint32_t i = 0; i < n && i < capacity; i++) out[i] = ctx->b.timings[i];
    return 0;
}
int PLAIN_FN(set_timing_enabled)(plain_ctx* ctx, int enabled) { ctx->b.timingEnabled = enabled != 0; return 0; }

static int imageCopy(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void* host, size_t size, bool toDevice, bool sync, const char* what) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceImage* img = b.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, std::string(what) + ": invalid handle/mip");
    if (size != img->mips[mip].bytes) return fail(ctx, std::string(what) + ": size mismatch");
    unsigned char* dev = img->ptr + img->mips[mip].offset;
    if (sync) drainTransfers(b);
    if (sync && img->deferredExchange && joinDeferredExchanges(ctx, b.stream)) return 1;
    cudaStream_t st = sync ? b.stream : transferStream(b, *img, toDevice);
    if (toDevice) CU_CHECK(ctx, cudaMemcpyAsync(dev, host, size, cudaMemcpyHostToDevice, st));
    else CU_CHECK(ctx, cudaMemcpyAsync(host, dev, size, cudaMemcpyDeviceToHost, st));
    if (toDevice && img->corners) img->cornersStale = true;
    if (!sync && !toDevice) markDownloaded(b, *img);
    if (sync) CU_CHECK(ctx, cudaStreamSynchronize(b.stream));
    return 0;
}
int PLAIN_FN(write_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, const void* data, size_t size) { return imageCopy(ctx, image, mip, (void*)data, size, true, true, "write_image"); }
int PLAIN_FN(read_image)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void* out, size_t size) { return imageCopy(ctx, image, mip, out, size, false, true, "read_image"); }
int PLAIN_FN(write_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, const void* data, size_t size) { return imageCopy(ctx, image, mip, (void*)data, size, true, false, "write_image_async"); }
int PLAIN_FN(read_image_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void* out, size_t size) { return imageCopy(ctx, image, mip, out, size, false, false, "read_image_async"); }
static int imageRowsCopy(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, void* host, size_t size, bool toDevice, const char* what) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceImage* img = b.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, std::string(what) + ": invalid handle/mip");
    const MipInfo& m = img->mips[mip];
    if (m.d != 1 || rowBegin > rowEnd || rowEnd > (uint32_t)m.h) return fail(ctx, std::string(what) + ": invalid row range");
    const size_t pitch = m.bytes / (size_t)m.h;
    if (size != pitch * (rowEnd - rowBegin)) return fail(ctx, std::string(what) + ": size mismatch");
    if (size == 0) return 0;
    unsigned char* dev = img->ptr + m.offset + pitch * rowBegin;
    cudaStream_t st = transferStream(b, *img, toDevice);
    if (toDevice) CU_CHECK(ctx, cudaMemcpyAsync(dev, host, size, cudaMemcpyHostToDevice, st));
    else CU_CHECK(ctx, cudaMemcpyAsync(host, dev, size, cudaMemcpyDeviceToHost, st));
    if (toDevice && img->corners) img->cornersStale = true;
    if (!toDevice) markDownloaded(b, *img);
    return 0;
}
int PLAIN_FN(write_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, const void* data, size_t size) { return imageRowsCopy(ctx, image, mip, rowBegin, rowEnd, (void*)data, size, true, "write_image_rows_async"); }
int PLAIN_FN(read_image_rows_async)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, uint32_t rowBegin, uint32_t rowEnd, void* out, size_t size) { return imageRowsCopy(ctx, image, mip, rowBegin, rowEnd, out, size, false, "read_image_rows_async"); }
int PLAIN_FN(read_storage_buffer)(plain_ctx* ctx, plain_handle buffer, void* out, size_t size) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (buffer >= b.storageBuffers.size() || size > b.storageBuffers[buffer].size) return fail(ctx, "read_storage_buffer: invalid buffer/size");
    drainTransfers(b);
    CU_CHECK(ctx, cudaMemcpyAsync(out, b.storageBuffers[buffer].ptr, size, cudaMemcpyDeviceToHost, b.stream));
    CU_CHECK(ctx, cudaStreamSynchronize(b.stream));
    return 0;
}
int PLAIN_FN(get_image_device_pointer)(plain_ctx* ctx, plain_image_handle image, uint32_t mip, void** out_ptr, size_t* out_size) {
    DeviceImage* img = ctx->b.resolve(image);
    if (!img || mip >= img->mips.size()) return fail(ctx, "get_image_device_pointer: invalid handle/mip");
    *out_ptr = img->ptr + img->mips[mip].offset;
    if (out_size) *out_size = img->mips[mip].bytes;
    return 0;
}
int PLAIN_FN(get_storage_buffer_device_pointer)(plain_ctx* ctx, plain_handle buffer, void** out_ptr, size_t* out_size) {
    if (buffer >= ctx->b.storageBuffers.size()) return fail(ctx, "get_storage_buffer_device_pointer: invalid buffer");
    *out_ptr = ctx->b.storageBuffers[buffer].ptr;
    if (out_size) *out_size = ctx->b.storageBuffers[buffer].size;
    return 0;
}
int PLAIN_FN(get_last_frame_launch_count)(plain_ctx* ctx, uint32_t* out) { *out = ctx->b.lastFrameLaunches; return 0; }
int PLAIN_FN(set_graph_replay_enabled)(plain_ctx* ctx, int enabled) { ctx->b.graphEnabled = enabled != 0; return 0; }
int PLAIN_FN(set_pass_fusion_enabled)(plain_ctx* ctx, int enabled) { ctx->b.fusionEnabled = enabled != 0; ctx->b.passEpoch++; return 0; }
int PLAIN_FN(set_concurrent_passes_enabled)(plain_ctx* ctx, int enabled) {
    if ((enabled != 0) != ctx->b.concurrentPasses) ctx->b.passEpoch++;  // cached graphs were captured with the other schedule
    ctx->b.concurrentPasses = enabled != 0;
    return 0;
}

// ---------------- peer exchange over NVLink ----------------
int PLAIN_FN(peer_init)(plain_ctx* ctx, uint32_t rank, uint32_t count) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (count < 1 || count > PLAIN_MAX_PEERS || rank >= count) return fail(ctx, "peer_init: invalid rank/count");
    if (b.peerSync[b.peerRank]) return fail(ctx, "peer_init: already initialised");
    uint32_t* block = nullptr;
    CU_CHECK(ctx, cudaMalloc(&block, Backend::kPeerSyncWords * 4));
    CU_CHECK(ctx, cudaMemset(block, 0, Backend::kPeerSyncWords * 4));
    b.peerRank = rank; b.peerCount = count;
    b.peerSync[rank] = block;
    return 0;
}
int PLAIN_FN(peer_get_sync_handle)(plain_ctx* ctx, void* out_handle) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (!b.peerCount) return fail(ctx, "peer_get_sync_handle: peer_init first");
    static_assert(sizeof(cudaIpcMemHandle_t) == PLAIN_IPC_HANDLE_BYTES, "IPC handle size");
    CU_CHECK(ctx, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out_handle, b.peerSync[b.peerRank]));
    return 0;
}
int PLAIN_FN(peer_open_sync)(plain_ctx* ctx, uint32_t peer, const void* handle) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (peer >= b.peerCount || peer == b.peerRank) return fail(ctx, "peer_open_sync: invalid peer");
    if (b.peerSync[peer]) return 0;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CU_CHECK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    b.peerSync[peer] = (uint32_t*)p;
    return 0;
}
int PLAIN_FN(peer_get_image_handle)(plain_ctx* ctx, plain_image_handle image, void* out_handle) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceImage* img = b.resolve(image);
    if (!img || !img->ptr || image.type == PLAIN_IMAGE_HANDLE_SWAPCHAIN) return fail(ctx, "peer_get_image_handle: invalid image");
    CU_CHECK(ctx, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out_handle, img->ptr));
    return 0;
}
int PLAIN_FN(peer_open_image)(plain_ctx* ctx, plain_image_handle image, uint32_t peer, const void* handle) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    DeviceImage* img = b.resolve(image);
    if (!img || image.type == PLAIN_IMAGE_HANDLE_SWAPCHAIN) return fail(ctx, "peer_open_image: invalid image");
    if (peer >= b.peerCount || peer == b.peerRank) return fail(ctx, "peer_open_image: invalid peer");
    if (img->peerPtr[peer]) return 0;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CU_CHECK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    img->peerPtr[peer] = (unsigned char*)p;
    return 0;
}
int PLAIN_FN(peer_image_ready)(plain_ctx* ctx, plain_image_handle image) {
    Backend& b = ctx->b;
    DeviceImage* img = b.resolve(image);
    if (!img || b.peerCount < 2) return 0;
    for (uint32_t p = 0; p < b.peerCount; p++)
        if (p != b.peerRank && !img->peerPtr[p]) return 0;
    return 1;
}
static int peerPushRows(plain_ctx* ctx, uint32_t n, const plain_peer_push* pushes, cudaStream_t stream, int blocksPerSm) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (b.peerCount < 2) return fail(ctx, "peer_push_rows: peer_init first");
    PushArgs args;
    int used = 0;
    auto flush = [&]() {
        if (!used) return;
        const unsigned blocksPerSegment = (unsigned)std::max(1, blocksPerSm * b.smCount / used);
        peerPushKernel<<<dim3(blocksPerSegment, (unsigned)used), 256, 0, stream>>>(args);
        b.launchCounter++;
        used = 0;
    };
    for (uint32_t i = 0; i < n; i++) {
        const plain_peer_push& q = pushes[i];
        DeviceImage* img = b.resolve(q.image);
        if (!img || q.mip_level >= img->mips.size()) return fail(ctx, "peer_push_rows: invalid image/mip");
        if (q.peer >= b.peerCount || q.peer == b.peerRank || !img->peerPtr[q.peer]) return fail(ctx, "peer_push_rows: peer image not mapped (peer_open_image)");
        const MipInfo& m = img->mips[q.mip_level];
        if (q.row_begin > q.row_end || q.row_end > (uint32_t)m.h) return fail(ctx, "peer_push_rows: invalid row range");
        if (q.row_end == q.row_begin) continue;
        const size_t pitch = m.bytes / ((size_t)m.h * m.d), off = m.offset + pitch * q.row_begin;  // 3-D levels: the rows in every slice
        args.seg[used].src = img->ptr + off;
        args.seg[used].dst = img->peerPtr[q.peer] + off;
        args.seg[used].bytes = (unsigned long long)(pitch * (q.row_end - q.row_begin));
        args.seg[used].sliceStride = (unsigned long long)(pitch * m.h);
        args.seg[used].slices = (unsigned int)m.d;
        if (++used == kMaxPushSegments) flush();
    }
    flush();
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(ctx, std::string("peer_push_rows: ") + cudaGetErrorString(e));
    return 0;
}
int PLAIN_FN(peer_push_rows)(plain_ctx* ctx, uint32_t n, const plain_peer_push* pushes) { return peerPushRows(ctx, n, pushes, ctx->b.stream, 4); }
static int peerBarrier(plain_ctx* ctx, bool deferred = false) {
    Backend& b = ctx->b;
    BarrierArgs a{};
    for (uint32_t p = 0; p < b.peerCount; p++) {
        if (!b.peerSync[p]) return fail(ctx, "peer_barrier: sync block of a peer not mapped (peer_open_sync)");
        a.peerFlags[p] = b.peerSync[p] + (deferred ? Backend::kPeerFlagsDeferredOffset : Backend::kPeerFlagsOffset);
    }
    a.localError = b.peerSync[b.peerRank] + Backend::kPeerErrorOffset;
    a.rank = b.peerRank; a.count = b.peerCount; a.epoch = deferred ? ++b.peerEpochDeferred : ++b.peerEpoch;
    if (!deferred) b.peerBarriersThisFrame++;
    // a rank that is late on the host (first graph instantiation, lazy IPC mapping, CPU contention) must not trip the others: 20 s by
    // default, PLAIN_PEER_TIMEOUT_MS to change it. A timeout sets the sticky error word; callers poll it every frame (peer_error_poll)
    static const long long timeoutMs = getenv("PLAIN_PEER_TIMEOUT_MS") ? atoll(getenv("PLAIN_PEER_TIMEOUT_MS")) : 20000ll;
    a.timeoutCycles = timeoutMs * 1900000ll;  // SM clock ~1.9 GHz
    peerBarrierKernel<<<1, 32, 0, deferred ? b.peerStream : b.stream>>>(a);
    b.launchCounter++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(ctx, std::string("peer_barrier: ") + cudaGetErrorString(e));
    return 0;
}
// Rows that only the next frame reads: pushed on peerStream behind everything the pass stream holds so far, next to the passes that
// follow. Fewer blocks than a synchronous push: the copy shares the SMs with the frame's kernels and has most of a frame to finish.
int PLAIN_FN(peer_push_rows_deferred)(plain_ctx* ctx, uint32_t n, const plain_peer_push* pushes) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (b.peerCount < 2) return fail(ctx, "peer_push_rows_deferred: peer_init first");
    if (b.peerBarriersThisFrame == 0) return fail(ctx, "peer_push_rows_deferred: no synchronous peer barrier in this frame yet - the peers may still be reading the rows' previous contents");
    if (!b.peerStream) {
        CU_CHECK(ctx, cudaStreamCreateWithFlags(&b.peerStream, cudaStreamNonBlocking));
        CU_CHECK(ctx, cudaEventCreateWithFlags(&b.peerFork, cudaEventDisableTiming));
        CU_CHECK(ctx, cudaEventCreateWithFlags(&b.peerDeferredDone, cudaEventDisableTiming));
    }
    CU_CHECK(ctx, cudaEventRecord(b.peerFork, b.stream));
    CU_CHECK(ctx, cudaStreamWaitEvent(b.peerStream, b.peerFork, 0));
    if (peerPushRows(ctx, n, pushes, b.peerStream, 1)) return 1;
    for (uint32_t i = 0; i < n; i++) if (DeviceImage* img = b.resolve(pushes[i].image)) img->deferredExchange = true;
    b.peerDeferredDirty = true;
    return 0;
}
// closes the deferred pushes of the frame: one barrier on the deferred flag set, on peerStream. The next frame's first submission
// (and any read-back) waits for it - by then every peer's rows have landed here and every peer has read what this rank will overwrite.
int PLAIN_FN(peer_flush_deferred)(plain_ctx* ctx) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (!b.peerDeferredDirty) return 0;
    // the barrier also orders this rank's consumers of the images (this frame's passes) before the peers' next pushes into them
    CU_CHECK(ctx, cudaEventRecord(b.peerFork, b.stream));
    CU_CHECK(ctx, cudaStreamWaitEvent(b.peerStream, b.peerFork, 0));
    if (peerBarrier(ctx, true)) return 1;
    CU_CHECK(ctx, cudaEventRecord(b.peerDeferredDone, b.peerStream));
    b.peerDeferredDirty = false;
    b.peerDeferredPending = true;
    return 0;
}
static int joinDeferredExchanges(plain_ctx* ctx, cudaStream_t stream) {
    Backend& b = ctx->b;
    if (b.peerDeferredDirty && PLAIN_FN(peer_flush_deferred)(ctx)) return 1;
    if (b.peerDeferredPending) CU_CHECK(ctx, cudaStreamWaitEvent(stream, b.peerDeferredDone, 0));
    return 0;
}
int PLAIN_FN(peer_barrier)(plain_ctx* ctx) {
    cudaSetDevice(ctx->b.device);
    if (ctx->b.peerCount < 2) return fail(ctx, "peer_barrier: peer_init first");
    return peerBarrier(ctx);
}
int PLAIN_FN(peer_allreduce_sum_u32)(plain_ctx* ctx, plain_handle storage_buffer, uint32_t count) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    if (b.peerCount < 2) return fail(ctx, "peer_allreduce_sum_u32: peer_init first");
    if (storage_buffer >= b.storageBuffers.size() || count > Backend::kPeerReduceMax || (size_t)count * 4 > b.storageBuffers[storage_buffer].size) return fail(ctx, "peer_allreduce_sum_u32: invalid buffer/count");
    ReduceArgs a{};
    const size_t parity = (b.peerReduceCount++ & 1u) * (size_t)PLAIN_MAX_PEERS * Backend::kPeerReduceMax;
    for (uint32_t p = 0; p < b.peerCount; p++) {
        if (!b.peerSync[p]) return fail(ctx, "peer_allreduce_sum_u32: sync block of a peer not mapped (peer_open_sync)");
        a.peerScratch[p] = b.peerSync[p] + Backend::kPeerScratchOffset + parity;
    }
    a.buffer = (uint32_t*)b.storageBuffers[storage_buffer].ptr;
    a.rank = b.peerRank; a.count = b.peerCount; a.n = count;
    peerReducePushKernel<<<b.peerCount, 256, 0, b.stream>>>(a);
    b.launchCounter++;
    if (peerBarrier(ctx)) return 1;
    peerReduceSumKernel<<<1, 256, 0, b.stream>>>(a);
    b.launchCounter++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(ctx, std::string("peer_allreduce_sum_u32: ") + cudaGetErrorString(e));
    return 0;
}
int PLAIN_FN(peer_error_poll)(plain_ctx* ctx, uint32_t* out_error) {
    // non-blocking: returns the value the PREVIOUS poll fetched (the word is sticky, so a time-out shows up one poll later at the latest)
    // and enqueues the next 4-byte read-back behind the work submitted so far
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    *out_error = 0;
    if (!b.peerCount) return 0;
    if (!b.peerErrorHost) {
        if (cudaHostAlloc((void**)&b.peerErrorHost, 4, cudaHostAllocDefault) != cudaSuccess) return fail(ctx, "peer_error_poll: cudaHostAlloc failed");
        *b.peerErrorHost = 0;
    }
    *out_error = *(volatile uint32_t*)b.peerErrorHost;
    CU_CHECK(ctx, cudaMemcpyAsync(b.peerErrorHost, b.peerSync[b.peerRank] + Backend::kPeerErrorOffset, 4, cudaMemcpyDeviceToHost, b.stream));
    return 0;
}
int PLAIN_FN(peer_error)(plain_ctx* ctx, uint32_t* out_error) {
    Backend& b = ctx->b;
    cudaSetDevice(b.device);
    *out_error = 0;
    if (!b.peerCount) return 0;
    drainTransfers(b);
    CU_CHECK(ctx, cudaMemcpyAsync(out_error, b.peerSync[b.peerRank] + Backend::kPeerErrorOffset, 4, cudaMemcpyDeviceToHost, b.stream));
    CU_CHECK(ctx, cudaStreamSynchronize(b.stream));
    if (*out_error) {  // reported: clear the sticky word so that a caller that recovers (re-handshake, retry) starts clean
        CU_CHECK(ctx, cudaMemsetAsync(b.peerSync[b.peerRank] + Backend::kPeerErrorOffset, 0, 4, b.stream));
        if (b.peerErrorHost) *b.peerErrorHost = 0;
    }
    return 0;
}
int PLAIN_FN(get_stream)(plain_ctx* ctx, void** out_stream) { *out_stream = (void*)ctx->b.stream; return 0; }
int PLAIN_FN(device_selftest)(plain_ctx* ctx, uint64_t* out_mismatches) {
    if (!ctx || !out_mismatches) return 1;
    cudaSetDevice(ctx->b.device);
    unsigned long long counts[8] = {};
    if (!pb::runDeviceSelftest(ctx->b.stream, counts, ctx->b.lastError)) return 1;
    for (int i = 0; i < 8; i++) out_mismatches[i] = counts[i];
    return 0;
}

}  // extern "C"
