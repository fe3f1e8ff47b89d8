// kernel_set.h -- launch entry points of one math-mode translation unit (faithful or fast).
#pragma once
#include "stage_kernel.cuh"

namespace trm {

// which stage_kernel instantiation to launch
enum Variant { VAR_EULER_RECOMPUTE = 0, VAR_EULER_LOAD = 1, VAR_GENERIC = 2 };

struct KernelSet {
    cudaError_t (*stage_f32)(int phys, int variant, const StageArgs<float>& a, int block, cudaStream_t st);
    cudaError_t (*stage_f64)(int phys, int variant, const StageArgs<double>& a, int block, cudaStream_t st);
    cudaError_t (*init_f32)(int64_t ncol, int64_t ld, int nz, int richards, const float* metrics, const DevParams<float>& p,
                            float* U, float* S, float* T, float* L, float* P, float* Wt, float* Sx, cudaStream_t st);
    cudaError_t (*init_f64)(int64_t ncol, int64_t ld, int nz, int richards, const double* metrics, const DevParams<double>& p,
                            double* U, double* S, double* T, double* L, double* P, double* Wt, double* Sx, cudaStream_t st);
    // ForwardEuler / Heun stage (mode = MODE_EULER | MODE_HEUN1 | MODE_HEUN2) with the pipeline state in shared
    // memory (euler_kernel.cuh); cudaErrorInvalidConfiguration = not applicable, use the generic kernel
    cudaError_t (*euler_f32)(int phys, int mode, int load_aux, const StageArgs<float>& a, cudaStream_t st);
    cudaError_t (*euler_f64)(int phys, int mode, int load_aux, const StageArgs<double>& a, cudaStream_t st);
    // LandModel surface block as its own launch (what = 0) / soil moisture limiting factor from stored fields (what = 1)
    cudaError_t (*surface_f32)(int what, const StageArgs<float>& a, cudaStream_t st);
    cudaError_t (*surface_f64)(int what, const StageArgs<double>& a, cudaStream_t st);
    // small domains: one warp per column, `nsteps` ForwardEuler (heun = 0) or Heun steps inside one launch
    // (warp_kernel.cuh); cudaErrorInvalidConfiguration = not applicable (LandModel, nz > 31)
    cudaError_t (*warp_f32)(int phys, int heun, int nsteps, const StageArgs<float>& a, cudaStream_t st);
    cudaError_t (*warp_f64)(int phys, int heun, int nsteps, const StageArgs<double>& a, cudaStream_t st);
};

const KernelSet& kernels_faithful();   // compiled with -fmad=false, reference operation order
const KernelSet& kernels_fast();       // FMA contraction + algebraic shortcuts

}  // namespace trm
