// ORACLE - test infrastructure only. Batch evaluation of the oracle's restatement of the reference's pure GLSL include functions
// (oracle/shader_inc.h) behind oracle_inc_eval_pure: the counterpart of oracle/_ref/libref_glsl.so's refglsl_eval_pure
// (tests/test_oracle_vs_reference_glsl.py compares the two bit for bit).
#include "backend.h"
#include "shader_inc.h"
#define INC_NS orc
#define INC_PREFIX oracle_inc_
#define INC_IS_REFERENCE 0
#define INC_PART_PURE 1
#include "inc_eval.h"
