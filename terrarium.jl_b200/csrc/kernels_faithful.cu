// Faithful math: reference operation order, no FMA contraction (this file is built with -fmad=false).
#define TRM_FAST 0
#include "kernels.inl"
