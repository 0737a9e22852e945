#!/usr/bin/env python
"""ORACLE - test infrastructure only. Turns whole compute shaders of the reference (resources/shaders/*.comp, read where they lie under
/root/reference, with the include files they name) into text a C++ compiler accepts against oracle/ref/glsl_shader.h, WITHOUT restating
them: every statement of `main()` and of the included functions stays the reference's own; only what GLSL and C++ spell differently is
touched (the rewrites of glsl_to_cpp.py plus the ones listed below), and the interface declarations (`layout(...) uniform ...`) become
plain variables that generated `bind()` / `unbind()` functions fill from / write back to the oracle backend's resources.

    glsl_shader_to_cpp.py <reference shader dir> <output dir> <shader.comp> [...]

One header per shader: oracle/_ref/glsl/shader_<name>.h (build output, git-ignored: no reference source enters the repository), holding
`namespace ref_<name> { includes; resources; spec constants; shader_main(); bind(); unbind(); local_size[3]; }`.
oracle/ref/ref_shader_passes.cpp includes them and registers each as an override of the oracle's own pass for that shader name, in a
second library (oracle/_ref/liboracle_refmain.so); tests/test_oracle_vs_reference_shaders.py renders the same frames through both
libraries and compares every image and buffer bit for bit.

Rewrites beyond glsl_to_cpp.py (each mechanical):
  * `layout(local_size_x = ..) in;` -> local_size[]; `layout(set, binding[, format]) uniform image2D/image3D/texture2D/texture3D/sampler NAME;`
    -> `static TYPE NAME;`; uniform / buffer blocks -> their members as static variables (a block with an instance name: a struct + instance);
    `layout(constant_id = N) const T NAME = V;` -> `static T NAME = V;` (bound from the pass's specialisation constants);
    `layout(push_constant) uniform B { .. };` -> static members (bound from the execution's push-constant bytes, std430 offsets)
  * block members are filled leaf by leaf at their std140 / std430 offsets (computed here), never by copying a C++ struct
  * `void main()` -> `static void shader_main()`; `X.xy = E;` / `X.xyz = E;` / `X.rgb = E;` (swizzle stores) -> `assign_xy(X, E);` ...;
    `X.xy op= E;` likewise; `T name[N] = { .. };` initialiser lists stay (C++ aggregate initialisation)
  * include guards (#ifndef G / #define G / #endif) go - files are spliced once; every other preprocessor line stays (skyMultiscatterLut.comp
    switches code with #define / #ifdef) and the macros are #undef'd at the end of the generated header
  * arrays sized by a specialisation constant or unsized (`uint histogram[constNBins]`, `BoundingBox instanceBBs[]`) are pointers into the bound
    buffer (struct elements only when their std430 layout equals the C++ struct's, checked here)
  * `layout(set = 2, ..) uniform texture2D[] textures;` -> one table of views over the backend's images (glsl_shader.h BindlessTextures)
  * `shared T x[..];` -> `static thread_local` (a workgroup runs on one OS thread); `barrier()` yields to the workgroup's fiber scheduler
  * GLSL scoping: a local is not in scope in its own initialiser, C++'s is - `float depth = texture(sampler2D(depth, s), uv).r;` and
    `float phase = phase(x);` get the texture / the local renamed (suffix _tex / _v)
  * `inout vec3[8] p` -> `vec3 (&p)[8]`; `vec3 n[3][3]` parameters -> the Nb33 type of glsl_to_cpp.py; `cullingTileSize.x` (a swizzled scalar) -> the scalar
  * fragment shaders (triangle.frag, depthPrepass.frag, run behind oracle/shading_hook.h): `layout(location = N) in / out T x;` -> `static thread_local T x;`,
    `discard;` -> `{ g_discarded = true; return; }`, textureCube resources are declared and left unbound
"""
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from glsl_to_cpp import convert as convert_spelling, strip_comments  # noqa: E402

SCALARS = {"float": "float", "int": "int", "uint": "uint", "bool": "bool"}
VECS = {"vec2": ("float", 2), "vec3": ("float", 3), "vec4": ("float", 4), "ivec2": ("int", 2), "ivec3": ("int", 3), "ivec4": ("int", 4),
        "uvec2": ("uint", 2), "uvec3": ("uint", 3), "uvec4": ("uint", 4)}
COMP = "xyzw"
RESOURCE_TYPES = ("image2D", "image3D", "texture2D", "texture3D", "textureCube", "sampler")


def round_up(x, a):
    return (x + a - 1) // a * a


class Layout:
    """std140 / std430 offsets of a block's members (OpenGL 4.6 spec 7.6.2.2)."""

    def __init__(self, structs, consts, std):
        self.structs, self.consts, self.std = structs, consts, std

    def array_len(self, text):
        text = text.strip()
        if text == "":
            return None  # unsized
        if text in self.consts:
            return int(self.consts[text])
        try:
            return int(text)
        except ValueError:
            return None  # sized by specialisation constants (histogramCombineTiles.comp:15): bound as a pointer into the buffer

    def type_info(self, t):
        """(align, size, leaves) with leaves = [(suffix, scalar, offset)]"""
        if t in SCALARS:
            return 4, 4, [("", t, 0)]
        if t in VECS:
            s, n = VECS[t]
            return (8 if n == 2 else 16), 4 * n, [("." + COMP[i], s, 4 * i) for i in range(n)]
        if t in ("mat4", "mat4x4"):
            return 16, 64, [(".c[%d].%s" % (c, COMP[r]), "float", 16 * c + 4 * r) for c in range(4) for r in range(4)]
        if t in self.structs:
            align, off, leaves = 0, 0, []
            for mt, mn, ml in self.structs[t]:
                a, s, lv = self.member_info(mt, ml)
                off = round_up(off, a)
                leaves += [("." + mn + sfx, sc, off + o) for sfx, sc, o in lv]
                off += s
                align = max(align, a)
            if self.std == "std140":
                align = round_up(align, 16)
            return align, round_up(off, align), leaves
        raise ValueError("unknown GLSL type in a block: " + t)

    def natural(self, t):
        """(size, leaves) of the C++ struct the converted GLSL text declares: members packed at 4-byte alignment"""
        if t in SCALARS:
            return 4, [("", t, 0)]
        if t in VECS:
            sc, n = VECS[t]
            return 4 * n, [("." + COMP[i], sc, 4 * i) for i in range(n)]
        if t in ("mat4", "mat4x4"):
            return 64, [(".c[%d].%s" % (c, COMP[r]), "float", 16 * c + 4 * r) for c in range(4) for r in range(4)]
        off, leaves = 0, []
        for mt, mn, ml in self.structs[t]:
            s, lv = self.natural(mt)
            for i in range(1 if ml is None else self.array_len(ml)):
                sfx = "" if ml is None else "[%d]" % i
                leaves += [("." + mn + sfx + x, sc, off + o) for x, sc, o in lv]
                off += s
        return off, leaves

    def same_as_natural(self, t):
        a, s, lv = self.type_info(t)
        ns, nlv = self.natural(t)
        return round_up(s, a) == ns and lv == nlv

    def member_info(self, t, arr):
        a, s, lv = self.type_info(t)
        if arr is None:
            return a, s, lv
        n = self.array_len(arr)
        if self.std == "std140":
            a = round_up(a, 16)
        stride = round_up(s, a)
        if n is None:
            return a, 0, [("[]", t, stride)]  # unsized: pointer + stride, bound directly
        leaves = []
        for i in range(n):
            leaves += [("[%d]%s" % (i, sfx), sc, i * stride + o) for sfx, sc, o in lv]
        return a, stride * n, leaves


MEMBER = re.compile(r"\s*(\w+)\s+(\w+)\s*(?:\[\s*([^\]]*?)\s*\])?\s*;")


def parse_members(body):
    out = []
    for m in MEMBER.finditer(body):
        out.append((m.group(1), m.group(2), m.group(3)))
    return out


def gather(shader_dir, name, seen):
    """the shader text with its includes spliced in (each file once, in first-use order)"""
    text = strip_comments((shader_dir / name).read_text())
    text = re.sub(r"^\s*#\s*(version|extension)\b[^\n]*$", "", text, flags=re.M)
    # the include guard (#ifndef G / #define G ... #endif) goes, files are spliced once; every other preprocessor line stays - the C++ preprocessor
    # reads #define / #ifdef / #else / #endif the same way (the macros are #undef'd at the end of the generated header)
    g = re.match(r"\s*#\s*ifndef\s+(\w+)\s*\n\s*#\s*define\s+(\w+)[^\n]*\n", text)
    if g and g.group(1) == g.group(2):
        text = text[g.end():]
        k = text.rfind("#endif")
        text = text[:k] + re.sub(r"^#endif[^\n]*", "", text[k:])

    def splice(m):
        inc = m.group(1)
        if inc in seen:
            return ""
        seen.add(inc)
        return "\n" + gather(shader_dir, inc, seen) + "\n"
    return re.sub(r'^\s*#include\s+"([^"]+)"\s*$', splice, text, flags=re.M)


def convert_shader(shader_dir, shader):
    ns = "ref_" + re.sub(r"\W", "_", Path(shader).stem)
    text = gather(shader_dir, shader, set())
    # compile-time constants that may size an array; specialisation constants are NOT among them (their value comes with the pass), so an
    # array they size is bound as a pointer into the buffer
    consts = dict(re.findall(r"^\s*const\s+(?:int|uint)\s+(\w+)\s*=\s*(\d+)\s*;", text, flags=re.M))
    structs = {}
    for m in re.finditer(r"\bstruct\s+(\w+)\s*\{(.*?)\}\s*;", text, flags=re.S):
        structs[m.group(1)] = parse_members(m.group(2))
    bind, unbind, decls = [], [], []
    local = [1, 1, 1]

    m = re.search(r"layout\s*\(([^)]*local_size[^)]*)\)\s*in\s*;", text)
    if m:
        for k, axis in enumerate("xyz"):
            mm = re.search(r"local_size_%s\s*=\s*(\d+)" % axis, m.group(1))
            if mm:
                local[k] = int(mm.group(1))
        text = text.replace(m.group(0), "")

    def leaf_code(var, leaves, base, writable):
        for sfx, sc, off in leaves:
            lv = var + sfx
            if sc == "bool":
                bind.append("    { uint32_t t; memcpy(&t, %s + %d, 4); %s = t != 0; }" % (base, off, lv))
                if writable:
                    unbind.append("    { uint32_t t = %s ? 1u : 0u; memcpy(%s + %d, &t, 4); }" % (lv, base, off))
            else:
                bind.append("    memcpy(&%s, %s + %d, 4);" % (lv, base, off))
                if writable:
                    unbind.append("    memcpy(%s + %d, &%s, 4);" % (base, off, lv))

    # set 2: the global texture array, `uniform texture3D[] textures;` -> one table of views over every image of the backend
    def bindless(m):
        bind.append("    %s.bind(c);" % m.group(2))
        return "static BindlessTextures %s;" % m.group(2)
    text = re.sub(r"layout\s*\(\s*set\s*=\s*2[^)]*\)\s*uniform\s+(texture[23]D)\s*\[\s*\]\s*(\w+)\s*;", bindless, text)
    # workgroup-shared variables: a workgroup runs on one OS thread (glsl_shader.h Fibers)
    text = re.sub(r"^\s*shared\s+(\w+)\s*((?:\[\d+\])+)\s*(\w+)\s*;", r"static thread_local \1 \3\2;", text, flags=re.M)
    text = re.sub(r"^\s*shared\s+(\w+)\s+(\w+)\s*\[\s*([A-Za-z_]\w*)\s*\]\s*;", r"static thread_local \1 \2[4096];  // sized by the specialisation constant \3", text, flags=re.M)
    text = re.sub(r"^\s*shared\s+(\w+)\s+(\w+)\s*((?:\[\w+\])*)\s*;", r"static thread_local \1 \2\3;", text, flags=re.M)
    # resources: images / textures / samplers
    def resource(m):
        quals, typ, name, arr = m.group(1), m.group(2), m.group(3), m.group(4)
        set_ = int(re.search(r"set\s*=\s*(\d+)", quals).group(1)) if re.search(r"set\s*=", quals) else 0
        binding = int(re.search(r"binding\s*=\s*(\d+)", quals).group(1))
        if arr is not None:
            raise ValueError("resource arrays (bindless textures) are not supported: " + name)
        if typ == "sampler":
            return "static const sampler %s = &orc::s_%s;" % (name, name.replace("g_sampler_", ""))
        if typ == "textureCube":
            return "static textureCube %s;" % name
        decls.append((typ, name))
        if typ.startswith("image"):
            bind.append("    %s = %s(c.storage(%d));" % (name, typ, binding))
        else:
            bind.append("    %s_view = c.sampled(%d); %s = &%s_view;" % (name, binding, name, name))
            return "static orc::View %s_view; static %s %s;" % (name, typ, name)
        assert set_ == 1, "images / textures of set %d" % set_
        return "static %s %s;" % (typ, name)
    text = re.sub(r"layout\s*\(([^)]*)\)\s*uniform\s+(%s)\s+(\w+)\s*(\[\s*\w*\s*\])?\s*;" % "|".join(RESOURCE_TYPES), resource, text)

    # fragment-shader varyings and outputs (triangle.frag): per-pixel variables the hook of oracle/shading_hook.h sets / reads
    text = re.sub(r"layout\s*\(\s*location\s*=\s*\d+\s*\)\s*(?:in|out)\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text)
    # specialisation constants
    def spec(m):
        cid, typ, name, val = int(m.group(1)), m.group(2), m.group(3), m.group(4).strip()
        if typ == "bool":
            bind.append("    %s = c.specBool(%d, %s);" % (name, cid, val))
        else:
            bind.append("    %s = c.spec<%s>(%d, %s);" % (name, typ, cid, val))
        return "static %s %s = %s;" % (typ, name, val)
    text = re.sub(r"layout\s*\(\s*constant_id\s*=\s*(\d+)\s*\)\s*const\s+(\w+)\s+(\w+)\s*=\s*([^;]+);", spec, text)

    # uniform / buffer / push-constant blocks
    def block(m):
        quals, kind, bname, body, inst = m.group(1), m.group(2), m.group(3), m.group(4), m.group(5)
        push = "push_constant" in quals
        std = "std430" if (push or "std430" in quals or kind == "buffer") and "std140" not in quals else "std140"
        members = parse_members(body)
        lay = Layout(structs, consts, std)
        if push:
            base, writable = "push", False
            pre = "    uint8_t push[256] = {0}; memcpy(push, c.exec->pushConstants.data(), c.exec->pushConstants.size() < 256 ? c.exec->pushConstants.size() : 256);"
        else:
            set_ = int(re.search(r"set\s*=\s*(\d+)", quals).group(1)) if re.search(r"set\s*=", quals) else 0
            binding = int(re.search(r"binding\s*=\s*(\d+)", quals).group(1))
            base = "b_%s" % bname
            writable = kind == "buffer"
            if set_ == 0:
                pre = "    const uint8_t* %s = (const uint8_t*)&c.g;" % base
                writable = False
            elif kind == "buffer":
                pre = "    uint8_t* %s = c.sbuf(%d);" % (base, binding)
            else:
                pre = "    const uint8_t* %s = c.ubuf(%d);" % (base, binding)
        bind.append(pre)
        if writable:
            unbind.append(pre)
        out, off = [], 0
        prefix = (inst + ".") if inst else ""
        for mt, mn, ml in members:
            a, s, lv = lay.member_info(mt, ml)
            off = round_up(off, a)
            if lv and lv[0][0] == "[]":  # unsized array / array sized by a specialisation constant: a pointer into the buffer
                elem = lv[0][1]
                if (mt, mn, ml) != members[-1]:
                    raise ValueError("%s: the array %s of unknown size is not the last member of its block" % (shader, mn))
                if elem not in SCALARS and not (elem in structs and lay.same_as_natural(elem)):
                    raise ValueError("%s: the array %s of unknown size has elements of type %s whose block layout differs from the C++ struct's" % (shader, mn, elem))
                out.append("static %s* %s;" % (elem, mn) if not inst else "%s* %s;" % (elem, mn))
                bind.append("    %s%s = (%s*)(%s + %d);" % (prefix, mn, elem, base, off))
                continue
            cpp_t = "mat4" if mt == "mat4x4" else mt
            decl = "%s %s%s;" % (cpp_t, mn, "[%d]" % lay.array_len(ml) if ml is not None else "")
            out.append(decl if inst else "static " + decl)
            leaf_code(prefix + mn, [(sfx, sc, off + o) for sfx, sc, o in lv], base, writable)
            off += s
        if inst:
            return "struct %s_t { %s }; static %s_t %s;" % (bname, " ".join(out), bname, inst)
        return "\n".join(out)
    text = re.sub(r"layout\s*\(([^)]*)\)\s*(?:(?:readonly|writeonly|coherent|restrict|volatile)\s+)*(uniform|buffer)\s+(\w+)\s*\{(.*?)\}\s*(\w*)\s*;", block, text, flags=re.S)

    # GLSL: a local variable is not in scope in its own initialiser, so `float depth = texture(sampler2D(depth, s), uv).r;` reads the texture
    # of the same name; C++ would read the local. Such textures get a suffix in their declaration and wherever a sampler is built from them.
    for typ, name in decls:
        if not typ.startswith("image") and re.search(r"\b(?:float|vec[234]|int|uint)\s+%s\s*=" % name, text):
            text = re.sub(r"\b(sampler[23]D\s*\(\s*)%s\b" % name, r"\1%s_tex" % name, text)
            text = text.replace("static %s %s;" % (typ, name), "static %s %s_tex;" % (typ, name))
            for k, line in enumerate(bind):
                bind[k] = line.replace("%s = &%s_view;" % (name, name), "%s_tex = &%s_view;" % (name, name))
    if re.search(r"\blayout\s*\(", text):
        raise ValueError("%s: an interface declaration was not understood: %s" % (shader, re.search(r"\blayout\s*\([^\n]*", text).group(0)))
    if re.search(r"^\s*shared\b", text, flags=re.M):
        raise ValueError("%s: a shared declaration was not understood: %s" % (shader, re.search(r"^\s*shared\b[^\n]*", text, flags=re.M).group(0)))

    # swizzle stores, before the spelling pass turns swizzles into calls
    sw = r"(xy|xyz|rg|rgb)"
    text = re.sub(r"^(\s*)([\w.\[\]]+)\.%s\s*=(?!=)\s*([^;]+);" % sw, lambda m: "%sassign_%s(%s, %s);" % (m.group(1), {"rg": "xy", "rgb": "xyz"}.get(m.group(3), m.group(3)), m.group(2), m.group(4)), text, flags=re.M)
    text = re.sub(r"^(\s*)([\w.\[\]]+)\.%s\s*([-+*/])=\s*([^;]+);" % sw,
                  lambda m: "%sassign_%s(%s, %s.%s %s (%s));" % (m.group(1), {"rg": "xy", "rgb": "xyz"}.get(m.group(3), m.group(3)), m.group(2), m.group(2), m.group(3), m.group(4), m.group(5)), text, flags=re.M)
    # GLSL: a local variable is not in scope in its own initialiser - `float phase = phase(x);` calls the function. C++ would call the float:
    # the local gets a suffix from its declaration to the end of the enclosing block
    while True:
        m = re.search(r"\b(float|int|uint|vec[234])\s+(\w+)\s*=\s*\2\s*\(", text)
        if not m:
            break
        name, depth, end = m.group(2), 0, len(text)
        for k in range(m.start(), len(text)):
            if text[k] == "{":
                depth += 1
            elif text[k] == "}":
                depth -= 1
                if depth < 0:
                    end = k
                    break
        block = text[m.start():end]
        block = block.replace(m.group(0), "%s %s_v = %s(" % (m.group(1), name, name), 1)
        block = re.sub(r"\b%s\b(?!\s*\(|_v)" % name, name + "_v", block)
        text = text[:m.start()] + block + text[end:]
    text = re.sub(r"\b(cullingTileSize)\.x\b", r"\1", text)  # GLSL lets a scalar be swizzled: s.x is s
    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s*\[(\d+)\]\s+(\w+)", r"\1 (&\3)[\2]", text)   # `inout vec3[8] p` -> a reference to an array
    text = re.sub(r"\bvec3\s+(\w+)\s*\[3\]\s*\[3\]", r"vec3[3][3] \1", text)  # C-style array declarator -> the type spelling glsl_to_cpp.py maps to Nb33
    text = re.sub(r"\bdiscard\s*;", "{ g_discarded = true; return; }", text)  # fragment shaders
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "static void shader_main()", text)
    text = convert_spelling(text)
    macros = sorted(set(re.findall(r"^\s*#\s*define\s+(\w+)", text, flags=re.M)))
    text += "\n" + "".join("#undef %s\n" % m for m in macros)
    head = "// GENERATED by oracle/ref/glsl_shader_to_cpp.py from %s - build output, not source. Do not commit.\n" % (shader_dir / shader)
    serial = "true" if re.search(r"\batomic\w+\s*\(", text) else "false"  # appends keep the invocation order
    fibers = "true" if re.search(r"\bbarrier\s*\(", text) else "false"
    text = "static const uvec3 gl_WorkGroupSize(%d, %d, %d);\n" % tuple(local) + text
    body = "namespace %s {\n%s\nstatic const int local_size[3] = {%d, %d, %d};\nstatic const bool serial = %s, fibers = %s;\nstatic void bind(orc::PassCtx& c) {\n%s\n}\nstatic void unbind(orc::PassCtx& c) {\n%s\n}\n}  // namespace %s\n" % (
        ns, text, local[0], local[1], local[2], serial, fibers, "\n".join(bind), "\n".join(unbind) if unbind else "    (void)c;", ns)
    return head + body


def main():
    src, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    for shader in sys.argv[3:]:
        (out / ("shader_%s.h" % Path(shader).stem)).write_text(convert_shader(src, shader))


if __name__ == "__main__":
    main()
