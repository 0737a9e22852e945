// ORACLE - test infrastructure only (see oracle/README.md). Never linked into the product library.
//
// passes_exposure.cpp - luminance histogram, auto exposure and the Hillaire sky LUTs.
#include "backend.h"
#include "shader_inc.h"

namespace orc {

// ---------------- histogramPerTile.comp:32-65 ----------------
ORACLE_PASS(pass_histogramPerTile, "histogramPerTile.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    const float minLuminance = c.spec<float>(1, 1.f);
    const float maxLuminance = c.spec<float>(2, 100.f);
    size_t perTileSize = 0;
    uint32_t* histogramPerTile = (uint32_t*)c.sbuf(0, &perTileSize);
    View src = c.sampled(2);
    plain_light_buffer light;
    memcpy(&light, c.sbuf(3), sizeof(light));
    const int tilesX = (int)dm::floor_((float)src.w() / 32.f + 0.999999f) ;  // placeholder, recomputed below
    (void)tilesX;
    c.forEachGroup([&](int gx, int gy, int) {
        std::vector<uint32_t> localHistogram(nBins, 0);
        bool any = false;
        const float minLuminanceLog = log(minLuminance);
        const float maxLuminanceLog = log(maxLuminance);
        for (int ly = 0; ly < 32; ly++)
            for (int lx = 0; lx < 32; lx++) {
                int x = gx * 32 + lx, y = gy * 32 + ly;
                if (x >= src.w() || y >= src.h()) continue;  // early return, histogramPerTile.comp:37-39
                any = true;
                vec3 color = src.fetch(x, y).xyz();
                float luminance = dot(color, vec3(0.2126f, 0.7152f, 0.0722f)) / light.previousFrameExposure;
                float luminanceLog = log(luminance);
                const uint32_t maxIndex = nBins - 1;
                uint32_t bin = f2uint((float)maxIndex * clamp((luminanceLog - minLuminanceLog) / (maxLuminanceLog - minLuminanceLog), 0.f, 1.f));
                localHistogram[bin] += 1;
            }
        // tileIndex = gl_WorkGroupID.x + gl_WorkGroupID.y * ceil(inputSize.x / 32)
        uint32_t tileIndex = (uint32_t)gx + (uint32_t)gy * (uint32_t)((src.w() + 31) / 32);
        // threads with localIndexFlat < nBins write; they exist iff the group's first row/columns are inside the image
        (void)any;
        for (uint32_t b = 0; b < nBins; b++) {
            // invocation (b % 32, b / 32) of this group writes bin b unless it returned early
            int x = gx * 32 + (int)(b % 32), y = gy * 32 + (int)(b / 32);
            if (x >= src.w() || y >= src.h()) continue;
            size_t idx = (size_t)tileIndex * nBins + b;
            if ((idx + 1) * 4 <= perTileSize) histogramPerTile[idx] = localHistogram[b];
        }
    });
}

// ---------------- histogramReset.comp:11-16 ----------------
ORACLE_PASS(pass_histogramReset, "histogramReset.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    uint32_t* histogram = (uint32_t*)c.sbuf(1);
    for (uint32_t i = 0; i < c.exec->dispatch[0] * 64; i++)
        if (i < nBins) histogram[i] = 0;
}

// ---------------- histogramCombineTiles.comp:27-34 ----------------
ORACLE_PASS(pass_histogramCombineTiles, "histogramCombineTiles.comp") {
    const uint32_t nBins = c.spec<uint32_t>(0, 64);
    size_t perTileSize = 0, histSize = 0;
    const uint32_t* histogramPerTile = (const uint32_t*)c.sbuf(0, &perTileSize);
    uint32_t* histogram = (uint32_t*)c.sbuf(1, &histSize);
    for (uint32_t gy = 0; gy < c.exec->dispatch[1]; gy++)
        for (uint32_t tile = 0; tile < c.exec->dispatch[0]; tile++)
            for (uint32_t lx = 0; lx < 64; lx++) {
                uint32_t bin = lx + 64 * gy;
                if (bin > nBins) continue;  // '>' as in the reference (:29); bin == nBins falls outside both buffers
                size_t src = (size_t)tile * nBins + bin;
                if ((size_t)(bin + 1) * 4 > histSize || (src + 1) * 4 > perTileSize) continue;  // out-of-bounds access is dropped
                histogram[bin] += histogramPerTile[src];
            }
}

// ---------------- preExposeLights.comp:28-88 ----------------
static float offsetFromSceneEV(float sceneEV100) {
    float darkExp = 2.84f;
    float lightExp = 12.81f;
    float lightOffset = 1.47f;
    float darkOffset = -3.17f;
    float t = clamp((sceneEV100 - darkExp) / (lightExp - darkOffset), 0.f, 1.f);
    return mix(darkOffset, lightOffset, t);
}
ORACLE_PASS(pass_preExposeLights, "preExposeLights.comp") {
    const int nBins = c.spec<int>(0, 64);
    const float minLuminance = c.spec<float>(1, 1.f);
    const float maxLuminance = c.spec<float>(2, 100.f);
    plain_light_buffer* lightBuffer = (plain_light_buffer*)c.sbuf(0);
    const uint32_t* histogram = (const uint32_t*)c.sbuf(1);
    View transmissionLut = c.sampled(2);
    const plain_global_shader_info& g = c.g;

    const float minLuminanceLog = log(minLuminance);
    const float maxLuminanceLog = log(maxLuminance);
    uint32_t pixelCount = (uint32_t)(g.screenResolution[0] * g.screenResolution[1]);
    float mean = 0.f;
    uint32_t countedPixels = 0;
    uint32_t currentPixelCount = 0;
    for (int i = 0; i < nBins; i++) {
        currentPixelCount += histogram[i];
        float percentage = (float)currentPixelCount / (float)pixelCount;
        if (percentage < 0.95f && percentage >= 0.5f) {
            float binValueLog = minLuminanceLog + (maxLuminanceLog - minLuminanceLog) * (float)i / ((float)nBins - 1.f);
            float binValueLinear = exp(binValueLog);
            mean += (float)histogram[i] * binValueLinear;
            countedPixels += histogram[i];
        }
    }
    mean /= (float)countedPixels;
    float sceneEV100 = log2(mean * 100.f / 12.5f);
    float exposureOffset = offsetFromSceneEV(sceneEV100);
    exposureOffset += g.exposureOffset;
    float targetEV100 = sceneEV100 - exposureOffset;
    targetEV100 = max(targetEV100, 10.f);
    float previousEV100 = log2(1.f / (max(lightBuffer->previousFrameExposure, 0.000001f) * 1.2f));
    float evDelta = targetEV100 - previousEV100;
    float evMaxChange = g.exposureAdaptionSpeedEvPerSec * g.deltaTime;
    float evChange = sign(evDelta) * min(abs(evDelta), abs(evMaxChange));
    float currentEV100 = previousEV100 + evChange;
    float exposure = 1.f / (pow(2.f, currentEV100) * 1.2f);
    lightBuffer->sunStrengthExposed = g.sunStrength * exposure;
    lightBuffer->previousFrameExposure = exposure;
    vec2 lutUV = vec2(0.f, -g.sunDirection[1] * 0.5f + 0.5f);
    vec3 sunColor = texture(transmissionLut, s_linearClamp, lutUV).xyz();
    lightBuffer->sunColor[0] = sunColor.x; lightBuffer->sunColor[1] = sunColor.y; lightBuffer->sunColor[2] = sunColor.z;
}

// ---------------- skyTransmissionLut.comp:16-48 ----------------
ORACLE_PASS(pass_skyTransmissionLut, "skyTransmissionLut.comp") {
    View lut = c.storage(0);
    plain_atmosphere_settings a;
    memcpy(&a, c.ubuf(1), sizeof(a));
    c.forEachInvocation(8, 8, 1, [&](int ux, int uy, int) {
        float x = (float)ux / (float)(lut.w() - 1);
        float y = (float)uy / (float)(lut.h() - 1);
        float height = mix(0.f, a.atmosphereHeight, x);
        float upDot = y * 2.f - 1.f;
        upDot = max(upDot, -0.999f);
        vec3 V = vec3(0.f, -upDot, sqrt(1.f - (upDot * upDot)));
        vec3 P = vec3(0.f, -height - a.earthRadius, 0.f);
        vec3 earthCenter = vec3(0.f);
        Intersection intersection = rayEarthIntersection(P - 0.01f, V, earthCenter, a.earthRadius, a.atmosphereHeight);
        float pathLength = max(distance(intersection.pos, P), 0.01f);
        const int sampleCount = 40;
        float stepLength = pathLength / (float)sampleCount;
        vec3 currentPos = intersection.pos;
        vec3 absorption = vec3(1.f);
        vec3 step = V * stepLength;
        for (int i = 0; i < sampleCount; i++) {
            currentPos -= step;
            float currentHeight = max(distance(earthCenter, currentPos) - a.earthRadius, 0.f);
            AtmosphereCoefficients co = calculateCoefficients(currentHeight, a);
            absorption *= exp(-co.extinction * stepLength);
        }
        absorption = intersection.hitEarth ? vec3(0.f) : absorption;
        lut.store(ux, uy, 0, vec4(absorption, 0.f));
    });
}

// ---------------- skyMultiscatterLut.comp:19-124 ('approximation' path) ----------------
ORACLE_PASS(pass_skyMultiscatterLut, "skyMultiscatterLut.comp") {
    View multiscatterLut = c.storage(0);
    View transmissionLut = c.sampled(1);
    plain_atmosphere_settings a;
    memcpy(&a, c.ubuf(3), sizeof(a));
    c.forEachInvocation(8, 8, 1, [&](int ux, int uy, int) {
        float x = (float)ux / (float)multiscatterLut.w();
        float y = (float)uy / (float)multiscatterLut.h();
        float height = mix(0.f, a.atmosphereHeight, x);
        vec3 P = vec3(0.f, -height - a.earthRadius, 0.f);
        vec3 earthCenter = vec3(0.f);
        float upDot = y * 2.f - 1.f;
        vec3 L = vec3(0.f, -upDot, sqrt(1.f - (upDot * upDot)));
        vec3 L_2nd = vec3(0.f);
        vec3 f_ms = vec3(0.f);
        float isotropicPhase = 1.f / (4.f * pi);
        int sampleCountSqrt = 8;
        float sampleCountSqrtRcp = 1.f / (float)sampleCountSqrt;
        for (int i = 0; i < sampleCountSqrt; i++) {
            for (int j = 0; j < sampleCountSqrt; j++) {
                float theta = pi * (float)i * sampleCountSqrtRcp;
                float phi = 2.f * pi * (float)j * sampleCountSqrtRcp;
                (void)phi;
                float sinTheta = sin(theta);
                float cosTheta = cos(theta);
                vec3 V = vec3(sinTheta * cosTheta, -cosTheta, sinTheta * sinTheta);
                int innerSampleCount = 20;
                vec3 inscattered = vec3(0.f);
                Intersection intersection = rayEarthIntersection(P, V, earthCenter, a.earthRadius, a.atmosphereHeight);
                vec3 currentPosition = P;
                float stepSize = intersection.distance / (float)innerSampleCount;
                V *= stepSize;
                vec3 L_f = vec3(0.f);
                vec3 earthAlbedo = vec3(0.3f);
                vec3 earthHitNormal = normalize(intersection.pos - earthCenter);
                float earthNoL = clamp(dot(earthHitNormal, L), 0.f, 1.f);
                vec3 up = normalize(currentPosition - earthCenter);
                vec2 lutUV = computeLutUV(0.f, a.atmosphereHeight, up, L);
                vec3 transmissionToIntersection = texture(transmissionLut, s_linearClamp, lutUV).xyz();
                vec3 incomingLight = transmissionToIntersection;
                vec3 earthLit = earthAlbedo / pi * incomingLight * earthNoL;
                vec3 direct = intersection.hitEarth ? earthLit : vec3(0.f);
                vec3 transmission = vec3(1.f);
                float currentHeight = -currentPosition.y - a.earthRadius;
                for (int k = 0; k < innerSampleCount; k++) {
                    currentPosition += V;
                    vec3 upc = vec3(0.f, -1.f, 0.f);
                    AtmosphereCoefficients co = calculateCoefficients(height, a);
                    vec3 scatteringCo = co.scatterRayleigh + co.scatterMie;
                    vec2 lutUV2 = computeLutUV(currentHeight, a.atmosphereHeight, upc, L);
                    vec3 transmissionSun = texture(transmissionLut, s_linearClamp, lutUV2).xyz();
                    vec3 coefficientIntegral = integrateInscattering(scatteringCo, co.extinction, stepSize);
                    L_f += coefficientIntegral * transmission;
                    vec3 scatterIntegral = coefficientIntegral * transmissionSun * isotropicPhase;
                    inscattered = inscattered + scatterIntegral * transmission;
                    transmission *= exp(-co.extinction * stepSize);
                }
                direct *= transmission;
                f_ms += L_f * sinTheta;
                L_2nd += (direct * transmission + inscattered) * sinTheta;
            }
        }
        float sampleCountInverse = 1.f / (float)(sampleCountSqrt * sampleCountSqrt);
        f_ms *= sampleCountInverse;
        L_2nd *= sampleCountInverse;
        vec3 F_ms = vec3(1.f) / (1.f - f_ms);
        vec3 multiscatter = L_2nd * F_ms;
        multiscatterLut.store(ux, uy, 0, vec4(multiscatter, 0.f));
    });
}

// ---------------- skyLut.comp:24-97 ----------------
static float shadowRay(vec3 P, vec3 D, vec3 C, float earthRadius) {
    vec3 L = C - P;
    float t_ca = dot(L, D);
    float d = sqrt(dot(L, L) - t_ca * t_ca);
    float t_hc_earth = sqrt(earthRadius * earthRadius - d * d);
    float t_earth = t_ca - t_hc_earth;
    return t_earth > 0.f ? 0.f : 1.f;
}
ORACLE_PASS(pass_skyLut, "skyLut.comp") {
    View skyLut = c.storage(0);
    View transmissionLut = c.sampled(1);
    View multiscatterLut = c.sampled(2);
    plain_atmosphere_settings a;
    memcpy(&a, c.ubuf(4), sizeof(a));
    plain_light_buffer light;
    memcpy(&light, c.sbuf(5), sizeof(light));
    const vec3 sunDir = c.gv3(c.g.sunDirection);
    // dispatch is (200/8, 100/8) = 25 x 12 groups: rows 96-99 are never written (Sky.cpp:311-312)
    c.forEachInvocation(8, 8, 1, [&](int ux, int uy, int) {
        float x = (float)ux / (float)skyLut.w();
        float y = (float)uy / (float)skyLut.h();
        vec3 V = fromSkyLut(vec2(x, y));
        vec3 earthCenter = vec3(0.f);
        float bias = 0.002f;
        vec3 P = vec3(0.f, -a.earthRadius - bias, 0.f);
        Intersection intersection = rayEarthIntersection(P, V, earthCenter, a.earthRadius, a.atmosphereHeight);
        const int sampleCount = 30;
        float stepSize = intersection.distance / (float)sampleCount;
        vec3 L = sunDir;
        float VoL = dot(V, L);
        float phaseR = phaseRayleigh(VoL);
        float phaseMie = cornetteShanksPhase(VoL, a.mieScatteringExponent);
        vec3 currentPosition = P;
        vec3 absorption = vec3(1.f);
        vec3 color = vec3(0.f);
        vec3 step = V * stepSize;
        for (int i = 0; i < sampleCount; i++) {
            currentPosition += step;
            vec3 up = currentPosition - earthCenter;
            float upLength = length(up);
            float currentHeight = upLength - a.earthRadius;
            up /= upLength;
            vec2 lutUV = computeLutUV(currentHeight, a.atmosphereHeight, up, L);
            vec3 transmission = texture(transmissionLut, s_linearClamp, lutUV).xyz();
            vec3 incomingLight = light.sunStrengthExposed * transmission;
            incomingLight *= shadowRay(currentPosition, L, earthCenter, a.earthRadius);
            AtmosphereCoefficients co = calculateCoefficients(currentHeight, a);
            vec3 inscatteringRayleight = co.scatterRayleigh * incomingLight * phaseR;
            vec3 inscatteringMie = co.scatterMie * incomingLight * phaseMie;
            vec3 inscattering = inscatteringRayleight + inscatteringMie;
            vec3 scatterIntegral = integrateInscattering(inscattering, co.extinction, stepSize);
            color = color + scatterIntegral * absorption;
            absorption *= exp(-co.extinction * stepSize);
            vec3 multiscattering = texture(multiscatterLut, s_linearClamp, lutUV).xyz();
            color += multiscattering * incomingLight * (co.scatterRayleigh + co.scatterMie) * stepSize * transmission;
        }
        skyLut.store(ux, uy, 0, vec4(color, 0.f));
    });
}

}  // namespace orc
