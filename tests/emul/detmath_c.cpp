// Test-only C wrapper: exposes detmath.h on arrays so tests/test_detmath.py can compare it with float64 numpy.
#include "detmath.h"
#define W1(name) extern "C" void dmw_##name(const float* x, float* o, int n) { for (int i = 0; i < n; i++) o[i] = dm::name(x[i]); }
#define W2(name) extern "C" void dmw_##name(const float* x, const float* y, float* o, int n) { for (int i = 0; i < n; i++) o[i] = dm::name(x[i], y[i]); }
W1(exp) W1(exp2) W1(log) W1(log2) W1(sin) W1(cos) W1(tan) W1(asin) W1(acos) W1(atan)
W2(pow) W2(atan2)
