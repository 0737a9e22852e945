// ORACLE - test infrastructure only. The one pass of the frame path whose shader is this build's own recast: gbufferShading.comp runs
// triangle.frag:177-341 over the packed G-buffer (include/plain_frame_types.h) instead of over rasterised fragments. To hold the oracle's
// restatement of those lines against the reference's own text, the geometry branch of the pass can be handed to a hook: oracle/_ref/
// liboracle_refmain.so (oracle/ref/ref_shader_passes.cpp) installs one that runs the reference's triangle.frag main() for the pixel, fed
// with what the G-buffer texel defines as the fragment's inputs. liboracle.so installs none.
#pragma once
#include "backend.h"

namespace orc {

struct FragmentInputs {
    int x, y;              // gl_FragCoord = (x + 0.5, y + 0.5)
    vec3 passPos;          // world position reconstructed from the depth texel (passes_shading.cpp shadeGeometry)
    vec3 N;                // the G-buffer's shading normal = the fragment's normal after normal mapping
    vec3 albedoTexel;      // what the albedo fetch returned (sRGB 8-bit)
    float specG, specB;    // roughness / metalness channels of the specular fetch
    vec3 dxN[2], dyN[2];   // the normals of the quad's pixels in this pixel's row (left, right) and column (top, bottom): dFdxFine / dFdyFine operands
};
struct ShadeGeometryHook {
    void (*beginPass)(PassCtx& c);               // once per execution of gbufferShading.comp, before any pixel
    vec3 (*shade)(const FragmentInputs& in);     // any thread
};
extern ShadeGeometryHook g_shadeGeometryHook;

// depthPrepass.frag behind the resolve step of the oracle's depth prepass (passes_raster.cpp): the fragment that won the pixel, with the varyings
// the rasteriser interpolated (clip-space position this frame / last frame, geometric normal = passTBN[2]) -> motion vector and encoded normal
struct PrepassFragmentHook {
    void (*beginPass)(PassCtx& c);
    void (*shade)(vec4 passPos, vec4 passPosPrevious, vec3 passNormal, vec2& motion, vec3& normal);  // any thread
};
extern PrepassFragmentHook g_prepassFragmentHook;

}  // namespace orc
