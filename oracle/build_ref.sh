#!/bin/bash
# ORACLE - test infrastructure only. Compiles the one part of the reference that builds here - the asset pipeline
# (PlainAssetPipeline: glTF import + CPU SDF bake, BASELINE configs[0]) - from its sources where they lie under
# /root/reference, into oracle/_ref/ (git-ignored, travels to the GPU box). No reference source is copied into the repo.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/Plain/src/AssetPipeline" ] || { echo "reference not present at $REF"; exit 0; }
mkdir -p "$OUT"
# the reference's own host-side functions that feed the frame path (camera matrices, view frustum, culling, Hammersley jitter, SDF
# bounding-box padding, vertex compression) behind the C entry points of oracle/ref/ref_host_shim.cpp: tests/test_host_vs_reference.py
if [ ! -f "$OUT/libref_host.so" ] || [ "$HERE/ref/ref_host_shim.cpp" -nt "$OUT/libref_host.so" ] || [ "$HERE/build_ref.sh" -nt "$OUT/libref_host.so" ]; then
    g++ -std=c++17 -O2 -w -fpermissive -shared -fPIC -fvisibility=hidden -include cassert -include cstring -include stdexcept -include math.h -include stdlib.h \
        -I"$HERE/shim" -I"$REF/Plain/src" -I"$REF/Plain/src/Common" -I"$REF/Plain/src/Runtime/Rendering" -I"$REF/Plain/vendor" -I"$REF/Plain/vendor/glm" \
        "$HERE/ref/ref_host_shim.cpp" "$REF"/Plain/src/Runtime/Rendering/Camera.cpp "$REF"/Plain/src/Runtime/Rendering/ViewFrustum.cpp "$REF"/Plain/src/Runtime/Rendering/Culling.cpp \
        "$REF"/Plain/src/Common/Utilities/MathUtils.cpp "$REF"/Plain/src/Common/sdfUtilities.cpp "$REF"/Plain/src/Common/CompressedTypes.cpp "$REF"/Plain/src/Common/AABB.cpp \
        -o "$OUT/libref_host.so"
    echo "built $OUT/libref_host.so"
fi
# the reference's pure GLSL include files compiled as C++ from where they lie (oracle/ref/glsl_to_cpp.py adapts spelling only; its output
# is a build product under _ref/glsl/, never committed) against the oracle's built-ins and image sampler: tests/test_oracle_vs_reference_glsl.py
if [ -d "$REF/resources/shaders" ]; then
    if [ ! -f "$OUT/libref_glsl.so" ] || [ "$HERE/ref/ref_glsl_shim.cpp" -nt "$OUT/libref_glsl.so" ] || [ "$HERE/ref/glsl_ref.h" -nt "$OUT/libref_glsl.so" ] || \
       [ "$HERE/ref/glsl_to_cpp.py" -nt "$OUT/libref_glsl.so" ] || [ "$HERE/inc_eval.h" -nt "$OUT/libref_glsl.so" ] || [ "$HERE/glsl.h" -nt "$OUT/libref_glsl.so" ] || [ "$HERE/image.h" -nt "$OUT/libref_glsl.so" ]; then
        python3 "$HERE/ref/glsl_to_cpp.py" "$REF/resources/shaders" "$OUT/glsl/reference_inc.h"
        FMA=""; [ "$(uname -m)" = "x86_64" ] && FMA="-mfma"
        g++ -std=c++17 -O2 -ffp-contract=off $FMA -fno-fast-math -shared -fPIC -fvisibility=hidden -w -I"$HERE" -I"$HERE/ref" -I"$OUT/glsl" -I"$HERE/../include" -I"$HERE/../plainrenderer_b200/csrc" \
            "$HERE/ref/ref_glsl_shim.cpp" -o "$OUT/libref_glsl.so"
        echo "built $OUT/libref_glsl.so"
    fi
fi
# whole compute shaders of the reference (main() included) compiled as C++ from where they lie and registered as overrides of the oracle's own
# passes: oracle/_ref/liboracle_refmain.so = the oracle's objects + oracle/ref/ref_shader_passes.cpp (tests/test_oracle_vs_reference_shaders.py).
# The generated headers are build products under _ref/glsl/, never committed.
REF_SHADERS="$(cat "$HERE/ref/ref_shaders.txt" | grep -v '^#' | tr '\n' ' ')"
if [ -d "$REF/resources/shaders" ] && [ -n "$REF_SHADERS" ]; then
    make -C "$HERE" >/dev/null
    NEWEST_OBJ=$(ls -t "$HERE"/_build/o_*.o "$HERE"/_build/h_*.o | head -1)
    if [ ! -f "$OUT/liboracle_refmain.so" ] || [ "$HERE/ref/ref_shader_passes.cpp" -nt "$OUT/liboracle_refmain.so" ] || [ "$HERE/ref/glsl_shader.h" -nt "$OUT/liboracle_refmain.so" ] || \
       [ "$HERE/ref/glsl_ref.h" -nt "$OUT/liboracle_refmain.so" ] || [ "$HERE/ref/glsl_shader_to_cpp.py" -nt "$OUT/liboracle_refmain.so" ] || [ "$HERE/ref/glsl_to_cpp.py" -nt "$OUT/liboracle_refmain.so" ] || \
       [ "$HERE/ref/ref_shaders.txt" -nt "$OUT/liboracle_refmain.so" ] || [ "$NEWEST_OBJ" -nt "$OUT/liboracle_refmain.so" ]; then
        mkdir -p "$OUT/glsl"
        python3 "$HERE/ref/glsl_shader_to_cpp.py" "$REF/resources/shaders" "$OUT/glsl" $REF_SHADERS triangle.frag depthPrepass.frag  # the two fragment shaders: behind oracle/shading_hook.h
        : > "$OUT/glsl/shaders_generated.h"; : > "$OUT/glsl/shaders_registered.h"
        for s in $REF_SHADERS; do
            n="${s%.comp}"
            echo "#include \"shader_$n.h\"" >> "$OUT/glsl/shaders_generated.h"
            echo "REF_SHADER($n, \"$s\")" >> "$OUT/glsl/shaders_registered.h"
        done
        echo "#define REF_SHADER_LIST \"$REF_SHADERS\"" >> "$OUT/glsl/shaders_registered.h"
        FMA=""; [ "$(uname -m)" = "x86_64" ] && FMA="-mfma"
        g++ -std=c++17 -O2 -ffp-contract=off $FMA -fno-fast-math -fPIC -fvisibility=hidden -w -DPLAIN_FN_PREFIX=oracle_ -DPLAIN_FRONTEND_PREFIX=oracle_frontend_ -DPLAIN_ASSET_PREFIX=oracle_asset_ \
            -I"$HERE" -I"$HERE/ref" -I"$OUT/glsl" -I"$HERE/../include" -I"$HERE/../plainrenderer_b200/csrc" -c "$HERE/ref/ref_shader_passes.cpp" -o "$OUT/glsl/ref_shader_passes.o"
        g++ -shared -o "$OUT/liboracle_refmain.so" "$HERE"/_build/o_*.o "$HERE"/_build/h_*.o "$OUT/glsl/ref_shader_passes.o" -lpthread
        echo "built $OUT/liboracle_refmain.so ($REF_SHADERS)"
    fi
fi
[ -x "$OUT/PlainAssetPipeline" ] && [ "$OUT/PlainAssetPipeline" -nt "$HERE/build_ref.sh" ] && { echo "up to date: $OUT/PlainAssetPipeline"; exit 0; }
# -include math.h / stdlib.h: libstdc++'s C++ wrappers pull the float overloads of abs/sin/cos/sqrt into the global namespace, as
# MSVC's headers do for the reference's own build. Without them the unqualified abs(float) calls of SceneSDF.cpp resolve to
# int abs(int) under g++ (every ray is rejected as "parallel", distances come out as square roots of integers).
g++ -std=c++17 -O2 -w -fpermissive -include cassert -include cstring -include condition_variable -include stdexcept -include math.h -include stdlib.h \
    -I"$HERE/shim" -I"$REF/Plain/src" -I"$REF/Plain/src/Common" -I"$REF/Plain/vendor" -I"$REF/Plain/vendor/glm" -I"$REF/Plain/vendor/tinygltf" \
    "$REF"/Plain/src/AssetPipeline/*.cpp "$REF"/Plain/src/Common/*.cpp "$REF"/Plain/src/Common/Utilities/*.cpp \
    -lpthread -o "$OUT/PlainAssetPipeline"
echo "built $OUT/PlainAssetPipeline"
