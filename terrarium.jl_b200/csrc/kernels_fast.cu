// Fast math: FMA contraction and algebraic shortcuts (cbrt/sqrt for van Genuchten n = 2, reciprocal multiplies).
#define TRM_FAST 1
#include "kernels.inl"
