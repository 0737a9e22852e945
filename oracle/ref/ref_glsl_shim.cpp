// ORACLE - test infrastructure only. The reference's own GLSL include files (resources/shaders/*.inc), compiled as C++ from where
// they lie under /root/reference (oracle/ref/glsl_to_cpp.py only adapts spelling; oracle/build_ref.sh builds this file into
// oracle/_ref/libref_glsl.so), behind the evaluation entry points of oracle/inc_eval.h. tests/test_oracle_vs_reference_glsl.py
// holds the oracle's restatement of every function (oracle/shader_inc.h, passes_gi.cpp, passes_post.cpp) against it, bit for bit.
#include "glsl_ref.h"

namespace refglsl {
static float g_time = 0.f;  // global.inc: the one global uniform these files read (dither.inc)
#include "reference_inc.h"
}  // namespace refglsl

#define INC_NS refglsl
#define INC_PREFIX refglsl_
#define INC_IS_REFERENCE 1
#define INC_PART_PURE 1
#define INC_PART_SDF 1
#define INC_PART_TAA 1
#include "../inc_eval.h"
