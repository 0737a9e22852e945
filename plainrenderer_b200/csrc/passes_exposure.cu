// passes_exposure.cu - luminance histogram, auto exposure, Hillaire sky LUTs.
//   histogramPerTile.comp:32-65, histogramReset.comp:11-16, histogramCombineTiles.comp:27-34, preExposeLights.comp:28-88,
//   skyTransmissionLut.comp:16-48, skyMultiscatterLut.comp:19-124, skyLut.comp:24-97
#include "shader_inc.cuh"

namespace pb {

// ---------------- histogramPerTile.comp ----------------
// One 256-thread block per 32x32 tile, four pixels per thread (one 128-bit load of packed R11G11B10 texels when the
// row is 16-byte aligned). Bins are counted in shared memory with match-any aggregation per warp: neighbouring
// pixels mostly share a bin, so a warp issues a handful of shared atomics instead of 32.
template <int NBINS_MAX>
__global__ void __launch_bounds__(256) histogramPerTileKernel(ImgView src, uint32_t* __restrict__ perTile, size_t perTileCount, const plain_light_buffer* __restrict__ light,
                                                               uint32_t nBins, float minLuminance, float maxLuminance, int tileRowOffset) {
    __shared__ uint32_t hist[NBINS_MAX];
    for (uint32_t i = threadIdx.x; i < nBins; i += 256) hist[i] = 0;
    __syncthreads();
    const float exposure = light->previousFrameExposure;
    const float minLuminanceLog = dm::log(minLuminance), maxLuminanceLog = dm::log(maxLuminance);
    const float range = maxLuminanceLog - minLuminanceLog;
    const uint32_t maxIndex = nBins - 1;
    const int tileRow = (int)blockIdx.y + tileRowOffset;  // row sharding: this launch covers tile rows [tileRowOffset, tileRowOffset + gridDim.y)
    const int tx = blockIdx.x * 32, ty = tileRow * 32;
    const int lx = (threadIdx.x & 7) * 4, ly = threadIdx.x >> 3;  // 8 threads x 4 px per row, 32 rows
    const int x0 = tx + lx, y = ty + ly;
    uint32_t texels[4];
    int valid = 0;
    if (y < src.h && x0 < src.w) {
        const uint32_t* row = (const uint32_t*)src.ptr + (size_t)y * src.w;
        if (x0 + 3 < src.w && ((src.w & 3) == 0)) {
            const uint4 t = __ldg((const uint4*)(row + x0));
            texels[0] = t.x; texels[1] = t.y; texels[2] = t.z; texels[3] = t.w;
            valid = 4;
        } else {
            for (int i = 0; i < 4 && x0 + i < src.w; i++) { texels[i] = __ldg(row + x0 + i); valid = i + 1; }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const bool active = i < valid;
        uint32_t bin = 0xffffffffu;
        if (active) {
            const vec3 color = unpackR11G11B10(texels[i]);
            const float luminance = dot(color, v3(0.2126f, 0.7152f, 0.0722f)) / exposure;
            const float luminanceLog = dm::log(luminance);
            bin = f2u((float)maxIndex * clampf((luminanceLog - minLuminanceLog) / range, 0.f, 1.f));
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (active && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
    }
    __syncthreads();
    // invocation (b % 32, b / 32) of the reference's 32x32 group writes bin b unless it left the image (:37-39, :58-64)
    const uint32_t tileIndex = blockIdx.x + (uint32_t)tileRow * ((src.w + 31) / 32);
    for (uint32_t b = threadIdx.x; b < nBins; b += 256) {
        const int px = tx + (int)(b % 32), py = ty + (int)(b / 32);
        if (px >= src.w || py >= src.h) continue;
        const size_t idx = (size_t)tileIndex * nBins + b;
        if (idx < perTileCount) perTile[idx] = hist[b];
    }
}

PLAIN_PASS(launch_histogramPerTile, "histogramPerTile.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    const float minLum = c.spec<float>(1, 1.f), maxLum = c.spec<float>(2, 100.f);
    size_t perTileSize = 0;
    uint32_t* perTile = c.sbuf<uint32_t>(0, &perTileSize);
    const ImgView src = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const plain_light_buffer* light = c.sbuf<plain_light_buffer>(3);
    if (c.failed) return;
    if (nBins > 256) { c.fail("histogramPerTile.comp: at most 256 bins"); return; }
    int ty0, ty1;
    c.window((int)c.exec->dispatch[1], ty0, ty1);  // row sharding unit: rows of 32x32 tiles
    if (ty1 <= ty0 || c.exec->dispatch[0] == 0) return;
    dim3 grid(c.exec->dispatch[0], ty1 - ty0);
    PLAIN_LAUNCH(c, histogramPerTileKernel<256>, grid, 256, 0, src, perTile, perTileSize / 4, light, nBins, minLum, maxLum, ty0);
}

__global__ void histogramResetKernel(uint32_t* histogram, uint32_t nBins, uint32_t invocations) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < invocations && i < nBins) histogram[i] = 0;
}
PLAIN_PASS(launch_histogramReset, "histogramReset.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    uint32_t* histogram = c.sbuf<uint32_t>(1);
    if (c.failed) return;
    const uint32_t inv = c.exec->dispatch[0] * 64;
    PLAIN_LAUNCH(c, histogramResetKernel, ceilDiv(inv, 128), 128, 0, histogram, nBins, inv);
}

// histogramCombineTiles.comp: one thread per bin and a slab of tiles per block; coalesced over bins, one global
// atomic per (block, bin). Integer sums are order independent, so the result equals the reference's atomicAdd chain.
__global__ void __launch_bounds__(256) histogramCombineKernel(const uint32_t* __restrict__ perTile, size_t perTileCount, uint32_t* histogram, size_t histCount,
                                                               uint32_t nBins, uint32_t tileBegin, uint32_t tileCount, uint32_t binInvocations, uint32_t tilesPerBlock) {
    const uint32_t binsPerRow = min(nBins + 1, binInvocations);  // bin == nBins passes the reference's '>' test (:29) and falls outside both buffers
    const uint32_t rowsPerIter = 256 / binsPerRow > 0 ? 256 / binsPerRow : 1;
    const uint32_t bin = threadIdx.x % binsPerRow, sub = threadIdx.x / binsPerRow;
    if (binsPerRow > 256 || sub >= rowsPerIter) return;
    const uint32_t t0 = tileBegin + blockIdx.x * tilesPerBlock, t1 = min(t0 + tilesPerBlock, tileCount);  // tiles [tileBegin, tileCount)
    uint32_t sum = 0;
    // eight independent loads in flight per thread (the loop was one load per ~500-cycle round trip: 16 us of pure latency, ncu long_scoreboard 53)
    uint32_t t = t0 + sub;
    for (; t + 7 * rowsPerIter < t1; t += 8 * rowsPerIter) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const size_t idx = (size_t)(t + k * rowsPerIter) * nBins + bin;
            v[k] = idx < perTileCount ? __ldg(perTile + idx) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) sum += v[k];
    }
    for (; t < t1; t += rowsPerIter) {
        const size_t idx = (size_t)t * nBins + bin;
        if (idx < perTileCount) sum += __ldg(perTile + idx);
    }
    if (bin < histCount && sum) atomicAdd(histogram + bin, sum);
}
PLAIN_PASS(launch_histogramCombine, "histogramCombineTiles.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    size_t perTileSize = 0, histSize = 0;
    const uint32_t* perTile = c.sbuf<uint32_t>(0, &perTileSize);
    uint32_t* histogram = c.sbuf<uint32_t>(1, &histSize);
    if (c.failed) return;
    const uint32_t tiles = c.exec->dispatch[0], binInv = c.exec->dispatch[1] * 64;
    if (nBins + 1 > 256) { c.fail("histogramCombineTiles.comp: at most 255 bins"); return; }
    const uint32_t tilesPerBlock = 32;
    int t0, t1;
    c.window((int)tiles, t0, t1);  // row sharding unit: tile indices (a rank sums the tiles of its own tile rows; the partial histograms are all-reduced)
    if (t1 <= t0) return;
    PLAIN_LAUNCH(c, histogramCombineKernel, ceilDiv((unsigned)(t1 - t0), tilesPerBlock), 256, 0, perTile, perTileSize / 4, histogram, histSize / 4, nBins, (uint32_t)t0, (uint32_t)t1, binInv, tilesPerBlock);
}

// ---------------- preExposeLights.comp:28-88 (single thread in the reference) ----------------
__device__ float offsetFromSceneEV(float sceneEV100) {  // :28-38
    const float darkExp = 2.84f, lightExp = 12.81f, lightOffset = 1.47f, darkOffset = -3.17f;
    const float t = clampf((sceneEV100 - darkExp) / (lightExp - darkOffset), 0.f, 1.f);
    return mixf(darkOffset, lightOffset, t);
}
__global__ void preExposeLightsKernel(plain_light_buffer* lightBuffer, const uint32_t* __restrict__ histogram, ImgView transmissionLut,
                                      const plain_global_shader_info* __restrict__ g, int nBins, float minLuminance, float maxLuminance) {
    // one warp (round 2): the per-bin work of :28-38 - the running pixel count (an exact integer prefix sum), the percentage test, the bin's
    // luminance exp(...) and the product count * luminance - is done by the lanes, four bins each; lane 0 then adds the products up in bin
    // order, which is the only serial part (the sum's rounding depends on the order). A single thread took 25 us on the critical path of a
    // row-sharded frame. More than 128 bins (never: nHistogramBins = 128) fall back to the serial loop.
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const float minLuminanceLog = dm::log(minLuminance), maxLuminanceLog = dm::log(maxLuminance);
    const uint32_t pixelCount = (uint32_t)(g->screenResolution[0] * g->screenResolution[1]);
    float mean = 0.f;
    uint32_t countedPixels = 0;
    __shared__ float sTerm[128];
    __shared__ uint32_t sCounted[128];
    if (nBins <= 128) {
        const int lane = threadIdx.x, base = lane * 4;
        uint32_t h[4], mine = 0;
        for (int k = 0; k < 4; k++) { h[k] = base + k < nBins ? histogram[base + k] : 0u; mine += h[k]; }
        uint32_t incl = mine;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        uint32_t currentPixelCount = incl - mine;
        for (int k = 0; k < 4; k++) {
            const int i = base + k;
            currentPixelCount += h[k];
            const float percentage = (float)currentPixelCount / (float)pixelCount;
            float term = 0.f;
            uint32_t counted = 0xffffffffu;  // marker: the bin is outside the window
            if (i < nBins && percentage < 0.95f && percentage >= 0.5f) {
                const float binValueLog = minLuminanceLog + (maxLuminanceLog - minLuminanceLog) * (float)i / ((float)nBins - 1.f);
                const float binValueLinear = dm::exp(binValueLog);
                term = (float)h[k] * binValueLinear;
                counted = h[k];
            }
            sTerm[i] = term; sCounted[i] = counted;
        }
        __syncwarp();
        if (lane != 0) return;
        for (int i = 0; i < nBins; i++)
            if (sCounted[i] != 0xffffffffu) { mean += sTerm[i]; countedPixels += sCounted[i]; }
    } else {
        if (threadIdx.x != 0) return;
        uint32_t currentPixelCount = 0;
        for (int i = 0; i < nBins; i++) {
            const uint32_t h = histogram[i];
            currentPixelCount += h;
            const float percentage = (float)currentPixelCount / (float)pixelCount;
            if (percentage < 0.95f && percentage >= 0.5f) {
                const float binValueLog = minLuminanceLog + (maxLuminanceLog - minLuminanceLog) * (float)i / ((float)nBins - 1.f);
                const float binValueLinear = dm::exp(binValueLog);
                mean += (float)h * binValueLinear;
                countedPixels += h;
            }
        }
    }
    mean /= (float)countedPixels;
    const float sceneEV100 = dm::log2(mean * 100.f / 12.5f);
    float exposureOffset = offsetFromSceneEV(sceneEV100);
    exposureOffset += g->exposureOffset;
    float targetEV100 = sceneEV100 - exposureOffset;
    targetEV100 = fmaxp(targetEV100, 10.f);
    const float previousEV100 = dm::log2(1.f / (fmaxp(lightBuffer->previousFrameExposure, 0.000001f) * 1.2f));
    const float evDelta = targetEV100 - previousEV100;
    const float evMaxChange = g->exposureAdaptionSpeedEvPerSec * g->deltaTime;
    const float evChange = signf(evDelta) * fminp(absf(evDelta), absf(evMaxChange));
    const float currentEV100 = previousEV100 + evChange;
    const float exposure = 1.f / (dm::pow(2.f, currentEV100) * 1.2f);
    lightBuffer->sunStrengthExposed = g->sunStrength * exposure;
    lightBuffer->previousFrameExposure = exposure;
    const vec2 lutUV = v2(0.f, -g->sunDirection[1] * 0.5f + 0.5f);
    const vec3 sunColor = sampleR11LinearClamp(transmissionLut, lutUV);
    lightBuffer->sunColor[0] = sunColor.x; lightBuffer->sunColor[1] = sunColor.y; lightBuffer->sunColor[2] = sunColor.z;
}
PLAIN_PASS(launch_preExposeLights, "preExposeLights.comp") {
    const int nBins = c.spec<int>(0, 64);
    const float minLum = c.spec<float>(1, 1.f), maxLum = c.spec<float>(2, 100.f);
    plain_light_buffer* light = c.sbuf<plain_light_buffer>(0);
    const uint32_t* histogram = c.sbuf<uint32_t>(1);
    const ImgView lut = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    if (c.failed) return;
    PLAIN_LAUNCH(c, preExposeLightsKernel, 1, 32, 0, light, histogram, lut, c.g, nBins, minLum, maxLum);
}

// ---------------- skyTransmissionLut.comp:16-48 ----------------
// eight lanes per texel like skyLutKernel below: lane l evaluates the decay factors of steps [5l, 5l + 5), lane 0 multiplies them up in step order
#define SKYT_LANES 8
#define SKYT_TEXELS_PER_BLOCK 16
__global__ void __launch_bounds__(SKYT_LANES * SKYT_TEXELS_PER_BLOCK) skyTransmissionLutKernel(ImgView lut, const plain_atmosphere_settings* __restrict__ ap, int limitX, int limitY) {
    constexpr int sampleCount = 40;
    __shared__ float sDecay[SKYT_TEXELS_PER_BLOCK][sampleCount][3];
    const int local = threadIdx.x / SKYT_LANES, l = threadIdx.x % SKYT_LANES;
    const int wLim = min(limitX, lut.w), hLim = min(limitY, lut.h);
    const int texel = blockIdx.x * SKYT_TEXELS_PER_BLOCK + local;
    const bool active = texel < wLim * hLim;
    const int ux = active ? texel % wLim : 0, uy = active ? texel / wLim : 0;
    const plain_atmosphere_settings a = *ap;
    const float x = (float)ux / (float)(lut.w - 1);
    const float y = (float)uy / (float)(lut.h - 1);
    const float height = mixf(0.f, a.atmosphereHeight, x);
    float upDot = y * 2.f - 1.f;
    upDot = fmaxp(upDot, -0.999f);
    const vec3 V = v3(0.f, -upDot, sqrtf_(1.f - (upDot * upDot)));
    const vec3 P = v3(0.f, -height - a.earthRadius, 0.f);
    const vec3 earthCenter = v3(0.f);
    const Intersection intersection = rayEarthIntersection(P - 0.01f, V, earthCenter, a.earthRadius, a.atmosphereHeight);
    const float pathLength = fmaxp(length(intersection.pos - P), 0.01f);
    const float stepLength = pathLength / (float)sampleCount;
    vec3 currentPos = intersection.pos;
    const vec3 step = V * stepLength;
    const int first = 5 * l, last = first + 5;
    for (int i = 0; i < first; i++) currentPos = currentPos - step;  // where the serial loop stands before step `first`
    if (active) {
        for (int i = first; i < last; i++) {
            currentPos = currentPos - step;
            const float currentHeight = fmaxp(length(earthCenter - currentPos) - a.earthRadius, 0.f);
            const AtmosphereCoefficients co = calculateCoefficients(currentHeight, a);
            const vec3 decay = vexp(-co.extinction * stepLength);
            sDecay[local][i][0] = decay.x; sDecay[local][i][1] = decay.y; sDecay[local][i][2] = decay.z;
        }
    }
    __syncthreads();
    if (!active || l != 0) return;
    vec3 absorption = v3(1.f);
    for (int i = 0; i < sampleCount; i++) absorption = absorption * v3(sDecay[local][i][0], sDecay[local][i][1], sDecay[local][i][2]);
    absorption = intersection.hitEarth ? v3(0.f) : absorption;
    storeR11(lut, ux, uy, absorption);
}
PLAIN_PASS(launch_skyTransmissionLut, "skyTransmissionLut.comp") {
    const ImgView lut = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const plain_atmosphere_settings* a = c.ubuf<plain_atmosphere_settings>(1);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    const int texels = std::min(limX, lut.w) * std::min(limY, lut.h);
    if (texels <= 0) return;
    PLAIN_LAUNCH(c, skyTransmissionLutKernel, dim3(ceilDiv((unsigned)texels, SKYT_TEXELS_PER_BLOCK)), SKYT_LANES * SKYT_TEXELS_PER_BLOCK, 0, lut, a, limX, limY);
}

// ---------------- skyMultiscatterLut.comp:19-124 ('approximation' path) ----------------
// One block of 64 threads per LUT texel: the 64 (i, j) directions are spread over the threads (one each) and the two partial sums are
// combined in the reference's (i, j) order by a serial pass of lane 0 over the per-direction terms, so the result
// does not depend on the lane mapping.
__global__ void __launch_bounds__(64) skyMultiscatterLutKernel(ImgView multiscatterLut, ImgView transmissionLut, const plain_atmosphere_settings* __restrict__ ap, int limitX, int limitY) {
    __shared__ float terms[64][6];
    const int ux = blockIdx.x, uy = blockIdx.y;
    if (ux >= limitX || uy >= limitY || ux >= multiscatterLut.w || uy >= multiscatterLut.h) return;
    const plain_atmosphere_settings a = *ap;
    const float x = (float)ux / (float)multiscatterLut.w;
    const float y = (float)uy / (float)multiscatterLut.h;
    const float height = mixf(0.f, a.atmosphereHeight, x);
    const vec3 P = v3(0.f, -height - a.earthRadius, 0.f);
    const vec3 earthCenter = v3(0.f);
    const float upDot = y * 2.f - 1.f;
    const vec3 L = v3(0.f, -upDot, sqrtf_(1.f - (upDot * upDot)));
    const float isotropicPhase = 1.f / (4.f * PV_PI);
    const int sampleCountSqrt = 8;
    const float sampleCountSqrtRcp = 1.f / (float)sampleCountSqrt;
    for (int dirIdx = threadIdx.x; dirIdx < 64; dirIdx += 64) {  // one direction per thread (round 2: two per lane of one warp took twice as long on the critical path of a sharded frame)
        const int i = dirIdx / 8;
        const float theta = PV_PI * (float)i * sampleCountSqrtRcp;
        const float sinTheta = dm::sin(theta), cosTheta = dm::cos(theta);
        vec3 V = v3(sinTheta * cosTheta, -cosTheta, sinTheta * sinTheta);
        const int innerSampleCount = 20;
        vec3 inscattered = v3(0.f);
        const Intersection intersection = rayEarthIntersection(P, V, earthCenter, a.earthRadius, a.atmosphereHeight);
        vec3 currentPosition = P;
        const float stepSize = intersection.distance / (float)innerSampleCount;
        V = V * stepSize;
        vec3 L_f = v3(0.f);
        const vec3 earthAlbedo = v3(0.3f);
        const vec3 earthHitNormal = normalize(intersection.pos - earthCenter);
        const float earthNoL = clampf(dot(earthHitNormal, L), 0.f, 1.f);
        const vec3 up = normalize(currentPosition - earthCenter);
        const vec2 lutUV = computeLutUV(0.f, a.atmosphereHeight, up, L);
        const vec3 transmissionToIntersection = sampleR11LinearClamp(transmissionLut, lutUV);
        const vec3 earthLit = earthAlbedo / PV_PI * transmissionToIntersection * earthNoL;
        vec3 direct = intersection.hitEarth ? earthLit : v3(0.f);
        vec3 transmission = v3(1.f);
        const float currentHeight = -currentPosition.y - a.earthRadius;
        for (int k = 0; k < innerSampleCount; k++) {
            currentPosition = currentPosition + V;
            const vec3 upc = v3(0.f, -1.f, 0.f);
            const AtmosphereCoefficients co = calculateCoefficients(height, a);
            const vec3 scatteringCo = co.scatterRayleigh + co.scatterMie;
            const vec2 lutUV2 = computeLutUV(currentHeight, a.atmosphereHeight, upc, L);
            const vec3 transmissionSun = sampleR11LinearClamp(transmissionLut, lutUV2);
            const vec3 coefficientIntegral = integrateInscattering(scatteringCo, co.extinction, stepSize);
            L_f = L_f + coefficientIntegral * transmission;
            const vec3 scatterIntegral = coefficientIntegral * transmissionSun * isotropicPhase;
            inscattered = inscattered + scatterIntegral * transmission;
            transmission = transmission * vexp(-co.extinction * stepSize);
        }
        direct = direct * transmission;
        const vec3 fTerm = L_f * sinTheta;
        const vec3 lTerm = (direct * transmission + inscattered) * sinTheta;
        terms[dirIdx][0] = fTerm.x; terms[dirIdx][1] = fTerm.y; terms[dirIdx][2] = fTerm.z;
        terms[dirIdx][3] = lTerm.x; terms[dirIdx][4] = lTerm.y; terms[dirIdx][5] = lTerm.z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        vec3 f_ms = v3(0.f), L_2nd = v3(0.f);
        for (int d = 0; d < 64; d++) {
            f_ms = f_ms + v3(terms[d][0], terms[d][1], terms[d][2]);
            L_2nd = L_2nd + v3(terms[d][3], terms[d][4], terms[d][5]);
        }
        const float sampleCountInverse = 1.f / (float)(sampleCountSqrt * sampleCountSqrt);
        f_ms = f_ms * sampleCountInverse;
        L_2nd = L_2nd * sampleCountInverse;
        const vec3 F_ms = v3(1.f) / (1.f - f_ms);
        storeR11(multiscatterLut, ux, uy, L_2nd * F_ms);
    }
}
PLAIN_PASS(launch_skyMultiscatterLut, "skyMultiscatterLut.comp") {
    const ImgView lut = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const ImgView transmission = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const plain_atmosphere_settings* a = c.ubuf<plain_atmosphere_settings>(3);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    dim3 grid(std::min(limX, lut.w), std::min(limY, lut.h));
    PLAIN_LAUNCH(c, skyMultiscatterLutKernel, grid, 64, 0, lut, transmission, a, limX, limY);
}

// ---------------- skyLut.comp:24-97 ----------------
__device__ float skyShadowRay(vec3 P, vec3 D, vec3 C, float earthRadius) {  // :24-34
    const vec3 L = C - P;
    const float t_ca = dot(L, D);
    const float d = sqrtf_(dot(L, L) - t_ca * t_ca);
    const float t_hc_earth = sqrtf_(earthRadius * earthRadius - d * d);
    const float t_earth = t_ca - t_hc_earth;
    return t_earth > 0.f ? 0.f : 1.f;
}
// Eight lanes per LUT texel (round 2): the 30 march steps of a texel (:52-93) are a serial chain only through `color` and `absorption`; what a
// step costs - the position's height, two LUT lookups, the shadow ray, the coefficients, three exponentials - depends on the step's position
// alone. Lane l evaluates steps [4l, 4l + 4) (lanes 6, 7: three each) from the position the serial loop would have reached (the same repeated
// additions), leaves each step's three vectors in shared memory, and lane 0 then runs the serial accumulation in step order. Same operations,
// same order per accumulator, a sixth of the latency: the kernel is 300 threads of pure dependency chain otherwise (ncu: 6 % occupancy, 55 us) -
// hidden beside the trace on one GPU, on the critical path of a row-sharded frame, where every rank computes the whole LUT.
#define SKY_LANES 8
#define SKY_TEXELS_PER_BLOCK 16
__global__ void __launch_bounds__(SKY_LANES * SKY_TEXELS_PER_BLOCK) skyLutKernel(ImgView skyLut, ImgView transmissionLut, ImgView multiscatterLut, const plain_atmosphere_settings* __restrict__ ap,
                                                    const plain_light_buffer* __restrict__ light, const plain_global_shader_info* __restrict__ g, int limitX, int limitY) {
    constexpr int sampleCount = 30;
    __shared__ float sStep[SKY_TEXELS_PER_BLOCK][sampleCount][9];  // per step: scatterIntegral, exp(-extinction * stepSize), the multiscattering term
    const int local = threadIdx.x / SKY_LANES, l = threadIdx.x % SKY_LANES;
    const int wLim = min(limitX, skyLut.w), hLim = min(limitY, skyLut.h);  // the dispatch truncates: rows 96-99 stay unwritten (Sky.cpp:311-312)
    const int texel = blockIdx.x * SKY_TEXELS_PER_BLOCK + local;
    const bool active = texel < wLim * hLim;
    const int ux = active ? texel % wLim : 0, uy = active ? texel / wLim : 0;
    const plain_atmosphere_settings a = *ap;
    const float sunStrengthExposed = light->sunStrengthExposed;
    const float x = (float)ux / (float)skyLut.w;
    const float y = (float)uy / (float)skyLut.h;
    const vec3 V = fromSkyLut(v2(x, y));
    const vec3 earthCenter = v3(0.f);
    const float bias = 0.002f;
    const vec3 P = v3(0.f, -a.earthRadius - bias, 0.f);
    const Intersection intersection = rayEarthIntersection(P, V, earthCenter, a.earthRadius, a.atmosphereHeight);
    const float stepSize = intersection.distance / (float)sampleCount;
    const vec3 L = v3(g->sunDirection[0], g->sunDirection[1], g->sunDirection[2]);
    const float VoL = dot(V, L);
    const float phaseR = phaseRayleigh(VoL);
    const float phaseMie = cornetteShanksPhase(VoL, a.mieScatteringExponent);
    const vec3 step = V * stepSize;
    const int first = l < 6 ? 4 * l : 24 + 3 * (l - 6), last = l < 6 ? first + 4 : first + 3;
    vec3 currentPosition = P;
    for (int i = 0; i < first; i++) currentPosition = currentPosition + step;  // where the serial loop stands before step `first`
    if (active) {
        for (int i = first; i < last; i++) {
            currentPosition = currentPosition + step;
            vec3 up = currentPosition - earthCenter;
            const float upLength = length(up);
            const float currentHeight = upLength - a.earthRadius;
            up = up / upLength;
            const vec2 lutUV = computeLutUV(currentHeight, a.atmosphereHeight, up, L);
            const vec3 transmission = sampleR11LinearClamp(transmissionLut, lutUV);
            vec3 incomingLight = sunStrengthExposed * transmission;
            incomingLight = incomingLight * skyShadowRay(currentPosition, L, earthCenter, a.earthRadius);
            const AtmosphereCoefficients co = calculateCoefficients(currentHeight, a);
            const vec3 inscatteringRayleight = co.scatterRayleigh * incomingLight * phaseR;
            const vec3 inscatteringMie = co.scatterMie * incomingLight * phaseMie;
            const vec3 inscattering = inscatteringRayleight + inscatteringMie;
            const vec3 scatterIntegral = integrateInscattering(inscattering, co.extinction, stepSize);
            const vec3 decay = vexp(-co.extinction * stepSize);
            const vec3 multiscattering = sampleR11LinearClamp(multiscatterLut, lutUV);
            const vec3 ms = multiscattering * incomingLight * (co.scatterRayleigh + co.scatterMie) * stepSize * transmission;
            float* o = sStep[local][i];
            o[0] = scatterIntegral.x; o[1] = scatterIntegral.y; o[2] = scatterIntegral.z;
            o[3] = decay.x; o[4] = decay.y; o[5] = decay.z;
            o[6] = ms.x; o[7] = ms.y; o[8] = ms.z;
        }
    }
    __syncthreads();
    if (!active || l != 0) return;
    vec3 absorption = v3(1.f);
    vec3 color = v3(0.f);
    for (int i = 0; i < sampleCount; i++) {  // :84-92 in step order
        const float* o = sStep[local][i];
        color = color + v3(o[0], o[1], o[2]) * absorption;
        absorption = absorption * v3(o[3], o[4], o[5]);
        color = color + v3(o[6], o[7], o[8]);
    }
    storeR11(skyLut, ux, uy, color);
}
PLAIN_PASS(launch_skyLut, "skyLut.comp") {
    const ImgView lut = c.storage(0, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const ImgView transmission = c.sampled(1, PLAIN_FORMAT_R11G11B10_UFLOAT), multiscatter = c.sampled(2, PLAIN_FORMAT_R11G11B10_UFLOAT);
    const plain_atmosphere_settings* a = c.ubuf<plain_atmosphere_settings>(4);
    const plain_light_buffer* light = c.sbuf<plain_light_buffer>(5);
    if (c.failed) return;
    const int limX = (int)c.exec->dispatch[0] * 8, limY = (int)c.exec->dispatch[1] * 8;
    const int texels = std::min(limX, lut.w) * std::min(limY, lut.h);
    if (texels <= 0) return;
    PLAIN_LAUNCH(c, skyLutKernel, dim3(ceilDiv((unsigned)texels, SKY_TEXELS_PER_BLOCK)), SKY_LANES * SKY_TEXELS_PER_BLOCK, 0, lut, transmission, multiscatter, a, light, c.g, limX, limY);
}

}  // namespace pb
