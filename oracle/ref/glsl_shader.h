// ORACLE - test infrastructure only. What whole compute shaders of the reference need beyond oracle/ref/glsl_ref.h to compile as C++
// (oracle/ref/ref_shader_passes.cpp): storage images, the compute built-in variables, unsigned-vector spelling, swizzle stores.
// Every arithmetic operation is oracle/glsl.h's (the numeric contract) and every texel access oracle/image.h's (the sampler the
// oracle's own passes use): this header only adds spelling.
#pragma once
#include <ucontext.h>
#include <vector>
#include "glsl_ref.h"
#include "backend.h"

namespace refglsl {

// ---- compute built-ins, set by dispatch() below ----
static thread_local uvec3 gl_GlobalInvocationID, gl_WorkGroupID, gl_LocalInvocationID, gl_NumWorkGroups;
static thread_local uint gl_LocalInvocationIndex;

// ---- fragment-shader inputs (triangle.frag through oracle/shading_hook.h): set per pixel by the hook ----
static thread_local vec4 gl_FragCoord;
static thread_local gl::vec3 g_dxN[2], g_dyN[2];                 // the quad's values of the one quantity the shader differentiates (the normal)
inline gl::vec3 dFdxFine(gl::vec3) { return g_dxN[1] - g_dxN[0]; }  // right - left in the pixel's row
inline gl::vec3 dFdyFine(gl::vec3) { return g_dyN[1] - g_dyN[0]; }  // bottom - top in its column
static thread_local bool g_discarded;                             // `discard;` (the converter spells it { g_discarded = true; return; })
struct textureCube {};                                            // declared by triangle.frag, not read by its main()
// the three material fetches of triangle.frag:178-180 return what the G-buffer texel says they returned: the bindless table hands out these
// slots for the (negative) material indices the hook sets, and texture(sampler, uv, bias) recognises them
static View g_materialSlot[3];
static thread_local vec4 g_materialTexel[3];

// ---- storage images ----
struct image2D { View v; image2D() {} explicit image2D(View view) : v(view) {} };
struct image3D { View v; image3D() {} explicit image3D(View view) : v(view) {} };
inline ivec2 imageSize(const image2D& i) { return ivec2(i.v.w(), i.v.h()); }
inline ivec3 imageSize(const image3D& i) { return ivec3(i.v.w(), i.v.h(), i.v.d()); }
inline vec4 imageLoad(const image2D& i, ivec2 p) { return i.v.fetch(p.x, p.y, 0); }
inline vec4 imageLoad(const image3D& i, ivec3 p) { return i.v.fetch(p.x, p.y, p.z); }
inline void imageStore(const image2D& i, ivec2 p, vec4 c) { i.v.store(p.x, p.y, 0, c); }
inline void imageStore(const image3D& i, ivec3 p, vec4 c) { i.v.store(p.x, p.y, p.z, c); }
inline vec4 texelFetch(sampler3D s, ivec3 p, int) { return s.t->fetch(p.x, p.y, p.z); }
inline ivec2 textureSize(sampler2D s, int) { return ivec2(s.t->w(), s.t->h()); }
inline vec4 textureLod(sampler2D s, gl::vec2 uv, float lod) {  // an explicit level, counted from the bound view's level (depthHiZPyramid.comp reads the pyramid's earlier levels)
    View v = *s.t;
    v.mip += (int)lod;
    return orc::texture(v, *s.s, uv);
}
inline vec4 texture(sampler2D s, gl::vec2 uv, float /*bias*/) {  // views are single mip levels; material slots: see above
    if (s.t >= g_materialSlot && s.t < g_materialSlot + 3) return g_materialTexel[s.t - g_materialSlot];
    return orc::texture(*s.t, *s.s, uv);
}
inline vec4 textureGather(sampler2D s, gl::vec2 uv, int = 0) { return orc::textureGather(*s.t, *s.s, uv); }

// ---- unsigned / signed vector spelling (GLSL converts implicitly; integer -> float conversions are exact below 2^24) ----
inline gl::vec3 operator+(uvec3 a, float b) { return gl::vec3((float)a.x + b, (float)a.y + b, (float)a.z + b); }
inline gl::vec2 operator+(uvec2 a, float b) { return gl::vec2((float)a.x + b, (float)a.y + b); }
inline gl::vec3 operator/(gl::vec3 a, uvec3 b) { return a / gl::vec3((float)b.x, (float)b.y, (float)b.z); }
inline gl::vec2 operator/(gl::vec2 a, uvec2 b) { return a / gl::vec2((float)b.x, (float)b.y); }
inline gl::vec2 operator/(gl::vec2 a, gl::ivec2 b) { return a / gl::vec2((float)b.x, (float)b.y); }
inline gl::vec2 operator+(gl::ivec2 a, float b) { return gl::vec2((float)a.x + b, (float)a.y + b); }
inline gl::vec2 operator-(gl::ivec2 a, float b) { return gl::vec2((float)a.x - b, (float)a.y - b); }
inline gl::vec2 operator/(float a, gl::ivec2 b) { return a / gl::vec2((float)b.x, (float)b.y); }
inline gl::vec2 operator*(gl::vec2 a, gl::ivec2 b) { return a * gl::vec2((float)b.x, (float)b.y); }
inline gl::vec2 operator/(gl::ivec2 a, gl::vec2 b) { return gl::vec2((float)a.x, (float)a.y) / b; }
inline float ceil(float x) { const float f = gl::floor(x); return f < x ? f + 1.f : f; }  // exact either way
struct bvec2 { bool x, y; };
inline bvec2 greaterThanEqual(gl::ivec2 a, uvec2 b) { return bvec2{(uint)a.x >= b.x, (uint)a.y >= b.y}; }  /* GLSL converts int -> uint implicitly */
inline bvec2 greaterThan(gl::ivec2 a, gl::ivec2 b) { return bvec2{a.x > b.x, a.y > b.y}; }
inline bvec2 lessThan(gl::ivec2 a, gl::ivec2 b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bvec3 greaterThanEqual(uvec3 a, uvec3 b) { return bvec3{a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bvec2 greaterThanEqual(uvec2 a, uvec2 b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bvec2 greaterThanEqual(gl::ivec2 a, gl::ivec2 b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bool any(bvec2 v) { return v.x || v.y; }
inline bvec2 greaterThan(gl::vec2 a, gl::vec2 b) { return bvec2{a.x > b.x, a.y > b.y}; }
inline bvec2 lessThan(gl::vec2 a, gl::vec2 b) { return bvec2{a.x < b.x, a.y < b.y}; }
struct bvec4 { bool x, y, z, w; };
inline bool any(bvec4 v) { return v.x || v.y || v.z || v.w; }
using gl::isnan;
inline bvec2 isnan(gl::vec2 a) { return bvec2{gl::isnan(a.x), gl::isnan(a.y)}; }
inline bvec3 isnan(gl::vec3 a) { return bvec3{gl::isnan(a.x), gl::isnan(a.y), gl::isnan(a.z)}; }
inline bvec4 isnan(gl::vec4 a) { return bvec4{gl::isnan(a.x), gl::isnan(a.y), gl::isnan(a.z), gl::isnan(a.w)}; }

// ---- atomics (the dispatch runs on one thread, in invocation order) ----
inline uint atomicAdd(uint& mem, uint v) { const uint old = mem; mem += v; return old; }
inline uint atomicMax(uint& mem, uint v) { const uint old = mem; if (v > mem) mem = v; return old; }
inline uint atomicMin(uint& mem, uint v) { const uint old = mem; if (v < mem) mem = v; return old; }

// ---- swizzle stores: X.xy = E; X.xyz = E; ----
template <typename V> inline void assign_xy(V& v, gl::vec2 e) { v.x = e.x; v.y = e.y; }
template <typename V> inline void assign_xyz(V& v, gl::vec3 e) { v.x = e.x; v.y = e.y; v.z = e.z; }

// ---- workgroup barriers: the invocations of a workgroup as fibers on one OS thread ----
// barrier() hands control back to the workgroup's scheduler, which resumes the next invocation; when every live invocation has reached the
// barrier (or returned) the round starts over. Deterministic (invocation order within every phase), and `shared` variables are plain
// thread_local statics because a workgroup never leaves its OS thread.
struct Fibers {
    static const size_t kStack = 128 * 1024;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stack;
    std::vector<char> done;
    ucontext_t scheduler;
    int current = -1;
    void (*body)() = nullptr;
};
static thread_local Fibers g_fibers;
inline void fiberEntry() { g_fibers.body(); g_fibers.done[(size_t)g_fibers.current] = 1; }  // uc_link returns to the scheduler
inline void barrier() { Fibers& f = g_fibers; swapcontext(&f.ctx[(size_t)f.current], &f.scheduler); }
inline void memoryBarrier() {}
inline void memoryBarrierShared() {}
inline void memoryBarrierBuffer() {}
inline void groupMemoryBarrier() {}
template <typename T> inline T subgroupBroadcastFirst(T v) { return v; }  // the value is uniform over the subgroup where the shaders use it

// run body() for every invocation of the execution's dispatch. serial (shaders with atomics on buffers: the append order is the invocation order,
// as in the oracle's restatement): one thread, workgroups in order. Otherwise the workgroups are spread over the oracle's threads
// (PassCtx::forEachGroup) - the built-in variables above are thread_local, everything else a shader reads is bound before the dispatch and not
// written during it. fibers: the shader calls barrier().
inline void setInvocation(const int local[3], int gx, int gy, int gz, int i) {
    const int lx = i % local[0], ly = (i / local[0]) % local[1], lz = i / (local[0] * local[1]);
    gl_LocalInvocationID = uvec3((uint)lx, (uint)ly, (uint)lz);
    gl_LocalInvocationIndex = (uint)i;
    gl_GlobalInvocationID = uvec3((uint)(gx * local[0] + lx), (uint)(gy * local[1] + ly), (uint)(gz * local[2] + lz));
}
inline void dispatch(const orc::PassCtx& c, const int local[3], bool serial, bool fibers, void (*body)()) {
    const uvec3 groups(c.exec->dispatch[0], c.exec->dispatch[1], c.exec->dispatch[2]);
    const int n = local[0] * local[1] * local[2];
    auto group = [&](int gx, int gy, int gz) {
        gl_NumWorkGroups = groups;
        gl_WorkGroupID = uvec3((uint)gx, (uint)gy, (uint)gz);
        if (!fibers) {
            for (int i = 0; i < n; i++) { setInvocation(local, gx, gy, gz, i); body(); }
            return;
        }
        Fibers& f = g_fibers;
        if ((int)f.ctx.size() < n) { f.ctx.resize((size_t)n); f.stack.resize((size_t)n); for (auto& st : f.stack) if (st.empty()) st.resize(Fibers::kStack); }
        f.done.assign((size_t)n, 0);
        f.body = body;
        for (int i = 0; i < n; i++) {
            getcontext(&f.ctx[(size_t)i]);
            f.ctx[(size_t)i].uc_stack.ss_sp = f.stack[(size_t)i].data();
            f.ctx[(size_t)i].uc_stack.ss_size = Fibers::kStack;
            f.ctx[(size_t)i].uc_link = &f.scheduler;
            makecontext(&f.ctx[(size_t)i], fiberEntry, 0);
        }
        for (int live = n; live > 0;) {
            live = 0;
            for (int i = 0; i < n; i++) {
                if (f.done[(size_t)i]) continue;
                f.current = i;
                setInvocation(local, gx, gy, gz, i);
                swapcontext(&f.scheduler, &f.ctx[(size_t)i]);  // until its next barrier() or its return
                if (!f.done[(size_t)i]) live++;
            }
        }
    };
    if (!serial) { c.forEachGroup(group); return; }
    for (uint gz = 0; gz < groups.z; gz++)
        for (uint gy = 0; gy < groups.y; gy++)
            for (uint gx = 0; gx < groups.x; gx++) group((int)gx, (int)gy, (int)gz);
}

// ---- set 2: the global texture array (RenderBackend.cpp:45), index == image handle index ----
struct BindlessTextures {
    std::vector<View> views;
    void bind(const orc::PassCtx& c) { views.resize(c.ctx->images.size()); for (size_t i = 0; i < views.size(); i++) views[i] = c.bindless((uint32_t)i); }
    const View* operator[](int i) const { return i < 0 ? &g_materialSlot[-i - 1] : &views[(size_t)i]; }
};
inline bool all(bvec2 v) { return v.x && v.y; }
inline bool all(bvec3 v) { return v.x && v.y && v.z; }

}  // namespace refglsl
