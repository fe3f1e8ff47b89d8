// kernels.inl -- instantiates the stage kernels for one math mode.  Included by kernels_faithful.cu
// (TRM_FAST = 0, built with -fmad=false) and kernels_fast.cu (TRM_FAST = 1).
#include "kernel_set.h"
#include <cstdlib>
#include "euler_kernel.cuh"
#include "euler2_kernel.cuh"
#include "warp_kernel.cuh"

namespace trm {
namespace {

constexpr bool kFast = TRM_FAST != 0;

template <class NF, int PHYS>
cudaError_t launch_phys(int variant, const StageArgs<NF>& a, int block, cudaStream_t st) {
    const int64_t nblk = (a.ncol + block - 1) / block;
    dim3 grid((unsigned)nblk), blk((unsigned)block);
    size_t smem = sizeof(NF) * MET_COUNT * MET_STRIDE;
    switch (variant) {
        case VAR_EULER_RECOMPUTE: stage_kernel<NF, PHYS, MODE_EULER, 0, kFast><<<grid, blk, smem, st>>>(a); break;
        case VAR_EULER_LOAD:      stage_kernel<NF, PHYS, MODE_EULER, 1, kFast><<<grid, blk, smem, st>>>(a); break;
        default:                  stage_kernel<NF, PHYS, -1, -1, kFast><<<grid, blk, smem, st>>>(a); break;
    }
    return cudaGetLastError();
}

// ForwardEuler / Heun stages, streaming kernel with the pipeline state in shared memory (euler_kernel.cuh)
template <class NF, int PHYS, int LOAD, int MS, int MODE, bool VG2 = false>
cudaError_t launch_euler_variant(const StageArgs<NF>& a, cudaStream_t st) {
    constexpr size_t smem = EulerSmem<NF, LOAD, MS, MODE>::BYTES;
    // the opt-in to more than 48 KB of dynamic shared memory is a per-device attribute of the function: one flag per
    // device ordinal (handles on several GPUs may live in one process, e.g. two Julia integrators)
    static bool configured[64] = {false};
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(euler_kernel<NF, PHYS, LOAD, kFast, MS, MODE, VG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int64_t nblk = (a.ncol + TRM_EULER_BLOCK - 1) / TRM_EULER_BLOCK;
    euler_kernel<NF, PHYS, LOAD, kFast, MS, MODE, VG2><<<(unsigned)nblk, TRM_EULER_BLOCK, smem, st>>>(a);
    return cudaGetLastError();
}
template <class NF, int PHYS, int LOAD>
cudaError_t launch_euler_mode(int mode, const StageArgs<NF>& a, cudaStream_t st) {
    // fast math, van Genuchten n = 2 soil, closure fields recomputed: the instantiations without the run-time tests for
    // the general retention / conductivity formulas (the hot configuration of the benchmark)
    constexpr bool CAN_VG2 = kFast && LOAD == 0 && phys_richards(PHYS);
    const bool vg2 = CAN_VG2 && a.p.vg_n_is_2 && a.p.swrc == TRM_SWRC_VANGENUCHTEN && a.p.unsat_k == TRM_UNSATK_VANGENUCHTEN;
    const bool compact = a.nz + 3 <= EULER_MS_SMALL;   // compact metric rows keep one more block resident
    if (mode == MODE_HEUN1) {
        if (vg2) return compact ? launch_euler_variant<NF, PHYS, LOAD, EULER_MS_SMALL, MODE_HEUN1, CAN_VG2>(a, st)
                                : launch_euler_variant<NF, PHYS, LOAD, MET_STRIDE, MODE_HEUN1, CAN_VG2>(a, st);
        return compact ? launch_euler_variant<NF, PHYS, LOAD, EULER_MS_SMALL, MODE_HEUN1>(a, st)
                       : launch_euler_variant<NF, PHYS, LOAD, MET_STRIDE, MODE_HEUN1>(a, st);
    }
    if (mode == MODE_HEUN2) {   // stage state: always recomputed
        constexpr bool H2_VG2 = kFast && phys_richards(PHYS);
        if (H2_VG2 && a.p.vg_n_is_2 && a.p.swrc == TRM_SWRC_VANGENUCHTEN && a.p.unsat_k == TRM_UNSATK_VANGENUCHTEN)
            return compact ? launch_euler_variant<NF, PHYS, 0, EULER_MS_SMALL, MODE_HEUN2, H2_VG2>(a, st)
                           : launch_euler_variant<NF, PHYS, 0, MET_STRIDE, MODE_HEUN2, H2_VG2>(a, st);
        return launch_euler_variant<NF, PHYS, 0, MET_STRIDE, MODE_HEUN2>(a, st);
    }
    if (vg2) return compact ? launch_euler_variant<NF, PHYS, LOAD, EULER_MS_SMALL, MODE_EULER, CAN_VG2>(a, st)
                            : launch_euler_variant<NF, PHYS, LOAD, MET_STRIDE, MODE_EULER, CAN_VG2>(a, st);
    return compact ? launch_euler_variant<NF, PHYS, LOAD, EULER_MS_SMALL, MODE_EULER>(a, st)
                   : launch_euler_variant<NF, PHYS, LOAD, MET_STRIDE, MODE_EULER>(a, st);
}
#if TRM_FAST
// fast math: two columns per thread (euler2_kernel.cuh) -- Float32 with packed f32x2 arithmetic, every timestepper stage ;
// Float64 with 16-byte accesses, ForwardEuler stage (the Heun stages of Float64 run the recompute protocol of euler_kernel)
template <class T, int PHYS, int MS, int MODE, int SOIL>
cudaError_t launch_euler2_variant(const StageArgs<T>& a, cudaStream_t st) {
    constexpr size_t smem = Euler2Smem<T, MS, MODE, phys_land(PHYS)>::BYTES;
    static bool configured[64] = {false};
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(euler2_kernel<T, PHYS, MS, MODE, SOIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int64_t npairs = (a.ncol + 1) / 2;
    const int64_t nblk = (npairs + euler2_block<T>() - 1) / euler2_block<T>();
    euler2_kernel<T, PHYS, MS, MODE, SOIL><<<(unsigned)nblk, euler2_block<T>(), smem, st>>>(a);
    return cudaGetLastError();
}
template <class T, int PHYS, int SOIL>
cudaError_t launch_euler2_mode(int mode, const StageArgs<T>& a, cudaStream_t st) {
    const bool compact = a.nz + 3 <= EULER_MS_SMALL;
    if constexpr (sizeof(T) == 4) {
        if (mode == MODE_HEUN1) return compact ? launch_euler2_variant<T, PHYS, EULER_MS_SMALL, MODE_HEUN1, SOIL>(a, st) : launch_euler2_variant<T, PHYS, MET_STRIDE, MODE_HEUN1, SOIL>(a, st);
        if (mode == MODE_HEUN2) return compact ? launch_euler2_variant<T, PHYS, EULER_MS_SMALL, MODE_HEUN2, SOIL>(a, st) : launch_euler2_variant<T, PHYS, MET_STRIDE, MODE_HEUN2, SOIL>(a, st);
    }
    return compact ? launch_euler2_variant<T, PHYS, EULER_MS_SMALL, MODE_EULER, SOIL>(a, st) : launch_euler2_variant<T, PHYS, MET_STRIDE, MODE_EULER, SOIL>(a, st);
}
// Which pair instantiation covers this launch, -1 for none (closure fields recomputed; Richards soils: van Genuchten n = 2
// for retention curve and conductivity, or Brooks-Corey with an integer 1 / lambda + linear conductivity -- the reference's
// default hydraulics). TRM_F32X2=0 switches the Float32 pair kernels off, TRM_F64X2=1 the Float64 pair kernel on.
template <class T>
inline int euler2_soil(int phys, int mode, int load_aux, const StageArgs<T>& a) {
    const char* e = std::getenv(sizeof(T) == 4 ? "TRM_F32X2" : "TRM_F64X2");   // (read per launch: tests compare both kernels within one process)
    if ((e && e[0] == '0') || load_aux) return -1;
    if (sizeof(T) == 8 && mode != MODE_EULER) return -1;   // (the Float64 Heun stages run the recompute protocol of euler_kernel)
    int soil = -1;
    if (!phys_richards(phys)) soil = SOIL2_VG2;
    else if (a.p.vg_n_is_2 && a.p.swrc == TRM_SWRC_VANGENUCHTEN && a.p.unsat_k == TRM_UNSATK_VANGENUCHTEN) soil = SOIL2_VG2;
    else if (a.p.swrc == TRM_SWRC_BROOKSCOREY && a.p.bc_k > 0 && a.p.unsat_k == TRM_UNSATK_LINEAR) soil = SOIL2_BC_LINEAR;
    // Float64, van Genuchten n = 2 / immobile water: the one-column kernel has its own compile-time instantiation and is the
    // faster one on the 10 M-column step (3.17 ms against 3.24 - 3.34 ms, profiles/r02_sweep_compact_metrics.txt): the pair
    // kernel runs on request only (TRM_F64X2=1). The default Brooks-Corey + linear soil has a compile-time instantiation
    // only here (the one-column kernel takes the general out-of-line formulas: 4.9 against 3.1 ms per 60 steps on 131 072
    // columns of the reference's benchmark design), so it stays on the pair kernel.
    if (sizeof(T) == 8 && soil == SOIL2_VG2 && !(e && e[0] == '1')) return -1;
    return soil;
}
template <class T>
inline cudaError_t launch_euler2(int phys, int mode, int soil, const StageArgs<T>& a, cudaStream_t st) {
#ifdef TRM_DEV_MIN   // (kernel tuning builds, profiles/build_variant.sh: only the instantiations of the soil benchmark)
    if (phys != PHYS_RICHARDS || soil != SOIL2_VG2) return cudaErrorInvalidConfiguration;
    return launch_euler2_mode<T, PHYS_RICHARDS, SOIL2_VG2>(mode, a, st);
#else
    switch (phys) {
        case PHYS_NOFLOW:   return launch_euler2_mode<T, PHYS_NOFLOW, SOIL2_VG2>(mode, a, st);
        case PHYS_RICHARDS: return soil == SOIL2_VG2 ? launch_euler2_mode<T, PHYS_RICHARDS, SOIL2_VG2>(mode, a, st) : launch_euler2_mode<T, PHYS_RICHARDS, SOIL2_BC_LINEAR>(mode, a, st);
        case PHYS_LAND:     return soil == SOIL2_VG2 ? launch_euler2_mode<T, PHYS_LAND, SOIL2_VG2>(mode, a, st) : launch_euler2_mode<T, PHYS_LAND, SOIL2_BC_LINEAR>(mode, a, st);
        default:            return launch_euler2_mode<T, PHYS_LAND_NOFLOW, SOIL2_VG2>(mode, a, st);
    }
#endif
}
#endif

template <class NF>
cudaError_t launch_euler(int phys, int mode, int load_aux, const StageArgs<NF>& a, cudaStream_t st) {
    // the kernel addresses layers with 32-bit element offsets; larger fields run the generic streaming kernel
    if ((uint64_t)a.nz * (uint64_t)a.ld >= (1ull << 32)) return cudaErrorInvalidConfiguration;
#if TRM_FAST
    {
        const int soil = euler2_soil<NF>(phys, mode, load_aux, a);
        if (soil >= 0) return launch_euler2<NF>(phys, mode, soil, a, st);
    }
#endif
#ifdef TRM_DEV_MIN
    if (phys != PHYS_RICHARDS) return cudaErrorInvalidConfiguration;
    return load_aux ? launch_euler_mode<NF, PHYS_RICHARDS, 1>(mode, a, st) : launch_euler_mode<NF, PHYS_RICHARDS, 0>(mode, a, st);
#endif
    switch (phys * 2 + (load_aux ? 1 : 0)) {
        case 0: return launch_euler_mode<NF, PHYS_NOFLOW, 0>(mode, a, st);
        case 1: return launch_euler_mode<NF, PHYS_NOFLOW, 1>(mode, a, st);
        case 2: return launch_euler_mode<NF, PHYS_RICHARDS, 0>(mode, a, st);
        case 3: return launch_euler_mode<NF, PHYS_RICHARDS, 1>(mode, a, st);
        case 4: return launch_euler_mode<NF, PHYS_LAND, 0>(mode, a, st);
        case 5: return launch_euler_mode<NF, PHYS_LAND, 1>(mode, a, st);
        case 6: return launch_euler_mode<NF, PHYS_LAND_NOFLOW, 0>(mode, a, st);
        default: return launch_euler_mode<NF, PHYS_LAND_NOFLOW, 1>(mode, a, st);
    }
}

// one warp per column, nsteps steps per launch (warp_kernel.cuh): nz <= 31 ; LandModel: one step per launch, after surface_kernel
template <class NF, bool LAND>
cudaError_t launch_warp_land(bool rich, int heun, int nsteps, const StageArgs<NF>& a, cudaStream_t st) {
    constexpr int cols = TRM_WARP_BLOCK / 32;
    const unsigned nblk = (unsigned)((a.ncol + cols - 1) / cols);
    if (!rich) { column_warp_kernel<NF, false, kFast, WSOIL_GENERIC, LAND><<<nblk, TRM_WARP_BLOCK, 0, st>>>(a, nsteps, heun); return cudaGetLastError(); }
#if TRM_FAST
    if (a.p.vg_n_is_2 && a.p.swrc == TRM_SWRC_VANGENUCHTEN && a.p.unsat_k == TRM_UNSATK_VANGENUCHTEN)
        column_warp_kernel<NF, true, true, WSOIL_VG2, LAND><<<nblk, TRM_WARP_BLOCK, 0, st>>>(a, nsteps, heun);
    else if (a.p.swrc == TRM_SWRC_BROOKSCOREY && a.p.bc_k > 0 && a.p.unsat_k == TRM_UNSATK_LINEAR)
        column_warp_kernel<NF, true, true, WSOIL_BC_LINEAR, LAND><<<nblk, TRM_WARP_BLOCK, 0, st>>>(a, nsteps, heun);
    else
#endif
        column_warp_kernel<NF, true, kFast, WSOIL_GENERIC, LAND><<<nblk, TRM_WARP_BLOCK, 0, st>>>(a, nsteps, heun);
    return cudaGetLastError();
}
template <class NF>
cudaError_t launch_warp(int phys, int heun, int nsteps, const StageArgs<NF>& a, cudaStream_t st) {
    if (a.nz > WARP_MAX_NZ || a.nz < 1 || (phys_land(phys) && nsteps != 1)) return cudaErrorInvalidConfiguration;
    return phys_land(phys) ? launch_warp_land<NF, true>(phys_richards(phys), heun, nsteps, a, st)
                           : launch_warp_land<NF, false>(phys_richards(phys), heun, nsteps, a, st);
}

template <class NF>
cudaError_t launch_stage(int phys, int variant, const StageArgs<NF>& a, int block, cudaStream_t st) {
    switch (phys) {
        case PHYS_NOFLOW:   return launch_phys<NF, PHYS_NOFLOW>(variant, a, block, st);
        case PHYS_RICHARDS: return launch_phys<NF, PHYS_RICHARDS>(variant, a, block, st);
        case PHYS_LAND:     return launch_phys<NF, PHYS_LAND>(variant, a, block, st);
        default:            return launch_phys<NF, PHYS_LAND_NOFLOW>(variant, a, block, st);
    }
}

template <class NF>
cudaError_t launch_surface(int what, const StageArgs<NF>& a, cudaStream_t st) {
    const unsigned nblk = (unsigned)((a.ncol + 127) / 128);
    if (what == 0) surface_kernel<NF, kFast><<<nblk, 128, 0, st>>>(a);
    else beta_kernel<NF, kFast><<<nblk, 128, 0, st>>>(a);
    return cudaGetLastError();
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
    static const KernelSet ks = {&launch_stage<float>, &launch_stage<double>, &launch_init<float>, &launch_init<double>,
                                 &launch_euler<float>, &launch_euler<double>, &launch_surface<float>, &launch_surface<double>,
                                 &launch_warp<float>, &launch_warp<double>};
    return ks;
}

}  // namespace trm
