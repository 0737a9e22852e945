// ORACLE - test infrastructure only. Whole compute shaders of the reference (resources/shaders/*.comp with their include files), compiled
// as C++ from where they lie under /root/reference (oracle/ref/glsl_shader_to_cpp.py adapts spelling and turns the interface declarations
// into variables; oracle/build_ref.sh builds this file + the oracle's own objects into oracle/_ref/liboracle_refmain.so). Each shader is
// registered as an OVERRIDE of the oracle's restatement of that pass: the library is the oracle with these passes executed by the
// reference's own main(). tests/test_oracle_vs_reference_shaders.py renders the same frames through liboracle.so and through this library
// and compares every image and buffer bit for bit - which pins the oracle's main() bodies, not only its include functions.
#include "glsl_shader.h"

#include "shading_hook.h"

namespace refglsl {
#include "shaders_generated.h"
#include "shader_triangle.h"  // triangle.frag: not a pass of its own here, see the hook at the end of this file
#include "shader_depthPrepass.h"  // depthPrepass.frag: likewise
}  // namespace refglsl

#include <map>
#include <string>
#include <tuple>
#include <vector>
namespace {
int g_runs = 0;  // executions that went through a reference main()
std::map<std::string, int>& runsOf() { static std::map<std::string, int> m; return m; }
}
// brdfLut.comp is a pure function of its specialisation constant and extent and costs 2.7e8 samples: like the oracle's own pass (passes_shading.cpp),
// the first evaluation in a process is the reference's main(), later ones copy its texels
static bool lutCached(orc::PassCtx& c, const char* file, bool store) {
    static std::map<std::tuple<int, int, int>, std::vector<uint8_t>> cache;
    if (std::string(file) != "brdfLut.comp") return false;
    orc::View lut = c.storage(0);
    std::vector<uint8_t>& texels = lut.img->mips[(size_t)lut.mip].data;
    const auto key = std::make_tuple(c.spec<int>(0, 0), lut.w(), lut.h());
    if (store) { cache[key] = texels; return true; }
    auto it = cache.find(key);
    if (it == cache.end() || it->second.size() != texels.size()) return false;
    texels = it->second;
    return true;
}
#define REF_SHADER(name, file)                                                                             \
    static void run_##name(orc::PassCtx& c) {                                                              \
        if (lutCached(c, file, false)) return;                                                             \
        refglsl::ref_##name::bind(c);                                                                      \
        refglsl::dispatch(c, refglsl::ref_##name::local_size, refglsl::ref_##name::serial, refglsl::ref_##name::fibers, refglsl::ref_##name::shader_main); \
        refglsl::ref_##name::unbind(c);                                                                    \
        g_runs++;                                                                                          \
        runsOf()[file]++;                                                                                  \
        lutCached(c, file, true);                                                                          \
    }                                                                                                      \
    static orc::PassOverride override_##name(file, run_##name);
#include "shaders_registered.h"

extern "C" __attribute__((visibility("default"))) int oracle_refmain_runs() { return g_runs; }
extern "C" __attribute__((visibility("default"))) int oracle_refmain_runs_of(const char* shader) { return runsOf()[shader]; }
extern "C" __attribute__((visibility("default"))) const char* oracle_refmain_shaders() { return REF_SHADER_LIST; }

// ---- triangle.frag: the reference's fragment shader behind the geometry branch of gbufferShading.comp (oracle/shading_hook.h) ----
// The pass binds what triangle.frag binds, under the same numbers (RenderFrontend.cpp shading execution); the fragment's varyings come from the
// G-buffer texel: passPos = the reconstructed world position, the normal map fetch returns NaN so that main() takes its own fallback
// `N = passTBN[2]` (triangle.frag:195-197) with passTBN[2] = the G-buffer's normal, the albedo / specular fetches return the stored texels.
namespace {
int g_triangleRuns = 0;
void beginTriangle(orc::PassCtx& c) {
    using namespace refglsl::ref_triangle;
    bind(c);
    albedoTextureIndex = -1;
    specularTextureIndex = -2;
    normalTextureIndex = -3;
    runsOf()["triangle.frag"]++;
    g_triangleRuns++;
    g_runs++;
}
gl::vec3 shadeTriangle(const orc::FragmentInputs& in) {
    using namespace refglsl::ref_triangle;
    const float nan = dm::u2f(0x7fc00000u);
    refglsl::gl_FragCoord = refglsl::vec4((float)in.x + 0.5f, (float)in.y + 0.5f, 0.f, 1.f);
    refglsl::g_materialTexel[0] = refglsl::vec4(in.albedoTexel.x, in.albedoTexel.y, in.albedoTexel.z, 1.f);
    refglsl::g_materialTexel[1] = refglsl::vec4(0.f, in.specG, in.specB, 1.f);
    refglsl::g_materialTexel[2] = refglsl::vec4(nan, nan, 0.f, 1.f);
    for (int k = 0; k < 2; k++) { refglsl::g_dxN[k] = in.dxN[k]; refglsl::g_dyN[k] = in.dyN[k]; }
    passUV = refglsl::vec2(0.f, 0.f);
    passPos = in.passPos;
    passTBN[0] = gl::vec3(0.f, 0.f, 0.f);
    passTBN[1] = gl::vec3(0.f, 0.f, 0.f);
    passTBN[2] = in.N;
    shader_main();
    return color;
}
struct InstallTriangle { InstallTriangle() { orc::g_shadeGeometryHook.beginPass = beginTriangle; orc::g_shadeGeometryHook.shade = shadeTriangle; } } g_installTriangle;
}  // namespace

// ---- depthPrepass.frag: the fragment stage behind the resolve step of the oracle's depth prepass (oracle/shading_hook.h) ----
// The fragment reached the resolve step, so its alpha test has passed: the albedo fetch returns alpha 1; the normal-map fetch is irrelevant
// (depthPrepass.frag:48 overwrites the normal-mapped normal with the geometric one).
namespace {
void beginPrepass(orc::PassCtx& c) {
    using namespace refglsl::ref_depthPrepass;
    bind(c);
    albedoTextureIndex = -1;
    normalTextureIndex = -3;
    runsOf()["depthPrepass.frag"]++;
    g_runs++;
}
void shadePrepass(gl::vec4 pos, gl::vec4 posPrevious, gl::vec3 n, gl::vec2& motionOut, gl::vec3& normalOut) {
    using namespace refglsl::ref_depthPrepass;
    refglsl::g_materialTexel[0] = refglsl::vec4(0.f, 0.f, 0.f, 1.f);
    refglsl::g_materialTexel[2] = refglsl::vec4(0.5f, 0.5f, 0.f, 1.f);
    refglsl::g_discarded = false;
    passUV = refglsl::vec2(0.f, 0.f);
    passPos = pos;
    passPosPrevious = posPrevious;
    passTBN[0] = gl::vec3(1.f, 0.f, 0.f);
    passTBN[1] = gl::vec3(0.f, 1.f, 0.f);
    passTBN[2] = n;
    shader_main();
    motionOut = motion;
    normalOut = normal;
}
struct InstallPrepass { InstallPrepass() { orc::g_prepassFragmentHook.beginPass = beginPrepass; orc::g_prepassFragmentHook.shade = shadePrepass; } } g_installPrepass;
}  // namespace
