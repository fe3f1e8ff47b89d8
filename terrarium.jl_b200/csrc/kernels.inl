// kernels.inl -- instantiates the stage kernels for one math mode.  Included by kernels_faithful.cu
// (TRM_FAST = 0, built with -fmad=false) and kernels_fast.cu (TRM_FAST = 1).
#include "kernel_set.h"

namespace trm {
namespace {

constexpr bool kFast = TRM_FAST != 0;

template <class NF, int PHYS>
cudaError_t launch_phys(int variant, const StageArgs<NF>& a, int block, cudaStream_t st) {
    const int64_t nblk = (a.ncol + block - 1) / block;
    dim3 grid((unsigned)nblk), blk((unsigned)block);
    size_t smem = sizeof(NF) * 6 * (size_t)(a.nz + 3);
    switch (variant) {
        case VAR_EULER_RECOMPUTE: stage_kernel<NF, PHYS, MODE_EULER, 0, kFast><<<grid, blk, smem, st>>>(a); break;
        case VAR_EULER_LOAD:      stage_kernel<NF, PHYS, MODE_EULER, 1, kFast><<<grid, blk, smem, st>>>(a); break;
        default:                  stage_kernel<NF, PHYS, -1, -1, kFast><<<grid, blk, smem, st>>>(a); break;
    }
    return cudaGetLastError();
}

template <class NF>
cudaError_t launch_stage(int phys, int variant, const StageArgs<NF>& a, int block, cudaStream_t st) {
    switch (phys) {
        case PHYS_NOFLOW:   return launch_phys<NF, PHYS_NOFLOW>(variant, a, block, st);
        case PHYS_RICHARDS: return launch_phys<NF, PHYS_RICHARDS>(variant, a, block, st);
        default:            return launch_phys<NF, PHYS_LAND>(variant, a, block, st);
    }
}

template <class NF>
cudaError_t launch_init(int64_t ncol, int64_t ld, int nz, int richards, const NF* metrics, const DevParams<NF>& p,
                        NF* U, NF* S, NF* T, NF* L, NF* P, NF* Wt, NF* Sx, cudaStream_t st) {
    const int block = 128;
    init_kernel<NF, kFast><<<(unsigned)((ncol + block - 1) / block), block, 0, st>>>(ncol, ld, nz, richards, metrics, p, U, S, T, L, P, Wt, Sx);
    return cudaGetLastError();
}

}  // namespace

#if TRM_FAST
const KernelSet& kernels_fast() {
#else
const KernelSet& kernels_faithful() {
#endif
    static const KernelSet ks = {&launch_stage<float>, &launch_stage<double>, &launch_init<float>, &launch_init<double>};
    return ks;
}

}  // namespace trm
