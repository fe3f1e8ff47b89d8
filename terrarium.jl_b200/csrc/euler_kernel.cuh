// euler_kernel.cuh -- ForwardEuler / Heun stages, streaming kernel with the pipeline state in shared memory.
//
// Same algorithm and the same per-cell arithmetic as stage_kernel (one thread = one column, one sweep
// bottom -> top, see the header of stage_kernel.cuh for the reference functions), but
//   * the values that travel between pipeline iterations (closure fields, conductivities, fluxes) live in a
//     per-thread strip of shared memory `[field][thread]` -- one slot per field, read by the next iteration before it
//     is overwritten; only Kf, which the iteration after next still needs, alternates between two slots -- instead of
//     registers: no register rotation, one copy of the loop body, 80 registers -> 6 resident blocks (24 warps) per SM;
//   * the raw U / sat loads are prefetched by cp.async (LDGSTS) four layers ahead into an 8-deep shared-memory ring in
//     which a layer stays from its prefetch until its update (never copied); T / liq / psi of the LOAD variant and the
//     k1 / base state of Heun stage 2 come through rings of their own;
//   * inner iterations (no halo, boundary face, Flux BC) run a second instantiation of the loop body without the
//     layer-index tests; a third axis (VG2) removes the run-time tests for the general retention / conductivity formulas
//     when the host has checked that the soil is van Genuchten with n = 2;
//   * output stores are evict-first (st.global.cs);
//   * the LandModel variants do not evaluate the surface block: surface_kernel (stage_kernel.cuh) runs before the stage
//     and leaves the ground heat flux / infiltration in their 2-D fields; for the vegetated model the stage accumulates
//     the soil moisture limiting factor of the state it WRITES for the next surface launch.
// Shared memory is addressed through explicit 32-bit shared addresses (`thread register + immediate`).
#pragma once

#include <type_traits>

#include "stage_kernel.cuh"

namespace trm {

#ifndef TRM_EULER_BLOCK
#define TRM_EULER_BLOCK 128     // threads per block (compile time: it is the stride of the smem strips)
#endif
#ifndef TRM_EULER_MIN_BLOCKS
#define TRM_EULER_MIN_BLOCKS 6   // <= 80 registers per thread, 24 resident warps per SM (measured: 4 blocks 5.08 ms, 5: 4.52, 6: 4.16 per 10 M-column step)
#endif
#ifndef TRM_EULER_F32_BLOCKS
#define TRM_EULER_F32_BLOCKS 8
#endif
#ifndef TRM_EULER_H2_BLOCKS
#define TRM_EULER_H2_BLOCKS 4    // Heun stage 2 (second ring: 44 - 52 KB of shared memory per block)
#endif
#ifndef TRM_EULER_LAND_BLOCKS
#define TRM_EULER_LAND_BLOCKS 6   // the LandModel variants no longer contain the surface block (surface_kernel): same budget as the soil kernel
#endif
// Resident blocks per SM the register allocator must allow. Register spills are ruinous here (shared memory
// leaves little L1 for local memory), so the faithful math mode with its inlined pow / IEEE division sequences
// gets a looser bound.
template <class NF, int PHYS, bool FAST, int MODE = MODE_EULER>
constexpr int euler_min_blocks() {
    return !FAST ? (sizeof(NF) == 8 ? 3 : 4) : (MODE == MODE_HEUN2 ? TRM_EULER_H2_BLOCKS : (phys_land(PHYS) ? TRM_EULER_LAND_BLOCKS : (sizeof(NF) == 4 ? TRM_EULER_F32_BLOCKS : TRM_EULER_MIN_BLOCKS)));
}

// volatile without a "memory" clobber: the shared-memory accesses of a thread keep their program order among
// themselves (every strip / ring location is private to one thread), while ordinary loads, stores and arithmetic
// may be scheduled across them
__device__ __forceinline__ void sts(uint32_t a, float v)  { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v)); }
__device__ __forceinline__ void sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v)); }
__device__ __forceinline__ float  ldsv(uint32_t a, float*)  { float v;  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double ldsv(uint32_t a, double*) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(dst), "l"(src), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }
// output stores: nothing written by a stage is read again before the next launch has streamed through gigabytes of
// other data, so the lines are marked evict-first (st.global.cs; measured 3.54 vs 3.56 ms, TRM_NO_STCS switches it off)
template <class NF> __device__ __forceinline__ void stg(NF* p, NF v) {
#ifdef TRM_NO_STCS
    *p = v;
#else
    __stcs(p, v);
#endif
}
// integer views used by the fast-math control flow: sign word of a value ; v < 1 for v >= 0 or v < 0 (not NaN)
__device__ __forceinline__ int sign_word(double v) { return __double2hiint(v); }
__device__ __forceinline__ int sign_word(float v)  { return __float_as_int(v); }
__device__ __forceinline__ bool below_one(double v) { return __double2hiint(v) < 0x3FF00000; }
__device__ __forceinline__ bool below_one(float v)  { return __float_as_int(v) < 0x3F800000; }

// fields of the per-thread pipeline strip. One slot each (written at the end of iteration m, read by iteration
// m+1 before it is overwritten), except Kf which iteration m+2 still needs (two alternating slots, EF_KF and
// EF_KF + 1) and, for the vegetated LandModel, the running soil moisture limiting factor of the state being written
// (EF_BETA). The LandModel variants do not evaluate the surface block: surface_kernel (stage_kernel.cuh) has run before
// and left the ground heat flux and the infiltration in their 2-D fields.
enum EulerField { EF_KF = 0, EF_T = 2, EF_P, EF_KAP, EF_KC, EF_QH, EF_G, EF_DQH, EF_QD, EF_BETA, EF_BCT /* prefetched TEMPERATURE_TOP boundary value */, EF_COUNT };
#ifndef TRM_EULER_DIST
#define TRM_EULER_DIST 4   // measured on the 10 M-column step: 3 layers ahead 3.675 ms, 4: 3.638, 5: 3.646
#endif
// layers the cp.async prefetch runs ahead of the layer entering the pipeline. Heun stage 2 keeps 3: its second ring
// (k1 and the base state) would need 8 instead of 4 slots per field and cost a resident block.
__host__ __device__ constexpr int euler_dist(int mode) { return mode == MODE_HEUN2 ? 3 : TRM_EULER_DIST; }
// depth of the prefetch rings that are consumed when a layer enters (T, liq, psi of the LOAD variant; Heun stage 2 ring)
__host__ __device__ constexpr int euler_pf(int mode) { return euler_dist(mode) < 4 ? 4 : 8; }
static_assert(TRM_EULER_DIST >= 1 && TRM_EULER_DIST <= 5, "the U / sat ring holds layers m-2 .. m+DIST in 8 slots");
constexpr int EULER_RD = 8;   // depth of the U / sat ring: layer k stays in slot (k & 7) from its prefetch (iteration
                              // k-3) until it is updated (iteration k+2), so the raw values are never copied

// position of a pipeline iteration in the column (see `iterate`): IP_GEN = layer-index tests at run time ; the others make
// them compile-time constants: m = 1, 2, 3 ; 4 <= m <= nz-DIST (the layer DIST ahead exists: prefetch) ; nz-DIST < m < nz ;
// m = nz (top layer enters) ; m = nz+1 (halo above the surface) ; m = nz+2 (top layer is updated). Needs nz >= 4.
enum IterPos { IP_GEN = 0, IP_M1, IP_M2, IP_M3, IP_INNER, IP_INNER_NP, IP_NZ, IP_HALO, IP_LAST };
template <int P> using IterTag = std::integral_constant<int, P>;
#ifndef TRM_EULER_SPEC
#define TRM_EULER_SPEC 1
#endif

constexpr int EULER_MS_SMALL = 40;   // compact metric rows for nz <= 37: read from the kernel parameters, no shared-memory copy
static_assert(EULER_MS_SMALL == CMET_STRIDE, "the compact variants read StageArgs::cmet");

// MODE: which timestepper stage the launch is (subset of StageMode): MODE_EULER, MODE_HEUN1 (stage state and k1 out,
// no closure fields), MODE_HEUN2 (tendencies evaluated on the stage state, averaged with k1, applied to the base state:
// k1 and the base U / sat of a layer are prefetched into a second cp.async ring two iterations before its update).
template <class NF, int LOAD, int MS, int MODE = MODE_EULER>
struct EulerSmem {
    static constexpr int PF_ = euler_pf(MODE);
    static constexpr bool CMET = MS == EULER_MS_SMALL;                                // compact rows: read from the kernel parameters (MetricsC)
    static constexpr int METRICS = CMET ? 0 : MET_COUNT * MS;                         // elements (the root fraction row is only filled by LandModel kernels)
    static constexpr int STRIP = EF_COUNT * TRM_EULER_BLOCK;
    static constexpr int RING = (2 * EULER_RD + (LOAD ? 3 : 0) * PF_) * TRM_EULER_BLOCK;   // U, sat (, T, liq, psi)
    // Heun stage 2: k1U, k1S, bU, bS of the layer to be updated (stored protocol) / k1U, k1S next to U, sat (recompute protocol)
    static constexpr int XRING = (MODE != MODE_HEUN2 ? 0 : (heun_recompute<NF>() ? 2 * EULER_RD : 4 * PF_)) * TRM_EULER_BLOCK;
    static constexpr size_t BYTES = sizeof(NF) * (size_t)(METRICS + STRIP + RING + XRING);
};

// Heun stage 1, recompute protocol, slow path of one column (Richards): a layer of the stage state went negative, so the
// stage copy needs the downward sweep of adjust_saturation_profile! (soil_hydrology.jl:201-216), which stage 2 cannot rebuild
// layer by layer. The column's stage state is formed here in full from the base state and the k1 this thread has just stored
// (explicit_step! with the time-n Flux BCs, upward sweep, downward sweep), written to yU / yS, and the column is flagged.
template <class NF, bool FAST, class Met, bool LAND>
__device__ __noinline__ void heun1_slow_column(const StageArgs<NF>& A, Met met, int64_t c) {
    const DevParams<NF>& p = A.p;
    const int nz = A.nz;
    const int64_t ld = A.ld;
    const NF dt = A.dt;
    auto bc_input = [&](int slot) -> NF { return eval_input(A.in[A.bc[slot].input], c, A.t_b, 1); };   // Flux BCs: time of the base state
    NF carry = NF(0);
#pragma unroll 1
    for (int k = 1; k <= nz; ++k) {
        const int64_t o = (int64_t)(k - 1) * ld + c;
        NF tU = A.oTU[o], tS = A.oTS[o];
        if (k == nz) {
            if (LAND) { tU -= A.G[c] / met.dzc(nz); tS -= (-A.infil[c]) / met.dzc(nz); }
            else {
                if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) tU -= bc_input(TRM_BC_ENERGY_TOP) / met.dzc(nz);
                if (A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) tS -= bc_input(TRM_BC_SATURATION_TOP) / met.dzc(nz);
            }
        }
        if (k == 1) {
            if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) tU += bc_input(TRM_BC_ENERGY_BOTTOM) / met.dzc(1);
            if (A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) tS += bc_input(TRM_BC_SATURATION_BOTTOM) / met.dzc(1);
        }
        A.yU[o] = A.bU[o] + tU * dt;
        NF s = A.bS[o] + tS * dt;
        s = s + carry;
        if (k < nz) {
            const NF e = M<NF, FAST>::pos(s - 1);
            s -= e;
            carry = FAST ? e * met.dzc(k) * met.rdzc(k + 1) : e * met.dzc(k) / met.dzc(k + 1);
        }
        A.yS[o] = s;
    }
    NF carry_dn = NF(0);
#pragma unroll 1
    for (int k = nz; k >= 1; --k) {
        const int64_t o = (int64_t)(k - 1) * ld + c;
        NF s = A.yS[o];
        if (k < nz) s -= carry_dn;
        if (k >= 2) {
            const NF d = jmax(-s, NF(0));
            s += d;
            carry_dn = d * met.dzc(k) / met.dzc(k - 1);
        }
        if (k == nz) s -= jmax(s - 1, NF(0));   // (the surface excess water of the stage copy is not used)
        if (k == 1) s = jmax(s, NF(0));
        A.yS[o] = s;
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 1; k <= nz; ++k) if (idx == 0 && A.yS[(int64_t)(k - 1) * ld + c] < 1) idx = k;
    if (idx == 0) idx = nz + 1;
    A.yWt[c] = met.zF(idx);
    A.hflag_out[c] = NF(1);
    if (LAND && has_veg(A)) {   // soil moisture limiting factor of the stage state, for the vegetation block of stage 2
        NF beta = NF(0);
#pragma unroll 1
        for (int k = 1; k <= nz; ++k) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            NF Tc, lc;
            energy_to_temperature<NF, FAST>(p, A.yU[o], A.yS[o], Tc, lc);
            beta += FAST ? plant_available_water_fast(A.vp, p, A.yS[o], lc) * met.root(k) : plant_available_water(A.vp, p, A.yS[o], lc) * met.root(k) / met.dzc(k) * met.dzc(k);
        }
        A.ybeta[c] = beta;
    }
}

// VG2: van Genuchten n = 2 for retention curve and conductivity, checked on the host (see cell_conductivity)
template <class NF, int PHYS, int LOAD_CT, bool FAST, int MS, int MODE = MODE_EULER, bool VG2 = false>
__global__ void __launch_bounds__(TRM_EULER_BLOCK, (euler_min_blocks<NF, PHYS, FAST, MODE>())) euler_kernel(const __grid_constant__ StageArgs<NF> A) {
    constexpr bool RICH = phys_richards(PHYS);
    constexpr bool LAND = phys_land(PHYS);
    constexpr bool LOAD = LOAD_CT != 0;
    constexpr int B = TRM_EULER_BLOCK;
    constexpr int ES = (int)sizeof(NF);
    using Mx = M<NF, FAST>;
    using SM = EulerSmem<NF, LOAD_CT, MS, MODE>;
    constexpr bool H1 = MODE == MODE_HEUN1, H2 = MODE == MODE_HEUN2;
    constexpr int DIST_ = euler_dist(MODE), PF_ = euler_pf(MODE);
    constexpr bool CLOSE = !H1;    // Heun stage 1 leaves the closure fields of the stage state to stage 2 (recomputed there)
    constexpr bool RC = heun_recompute<NF>() && (H1 || H2);   // Heun "recompute" protocol (see heun_recompute)
    // one instantiation of the loop body per position in the column (hot configurations only: code size / build time)
    constexpr bool SPEC = TRM_EULER_SPEC && FAST && !LOAD && (MS == EULER_MS_SMALL || H2);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nz = A.nz;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    using Met = std::conditional_t<SM::CMET, MetricsC<NF>, Metrics<NF, MS>>;
    Met met;
    if constexpr (SM::CMET) met.A = &A;
    else {
        NF* sm = reinterpret_cast<NF*>(smem_raw);
        for (int q = 0; q < met_rows(LAND); ++q)
            for (int i = threadIdx.x; i < nz + 3; i += B) sm[q * MS + i] = A.metrics[q * MET_STRIDE + i];
        __syncthreads();
        met.base = smem_base;
    }

    const int64_t c = (int64_t)blockIdx.x * B + threadIdx.x;
    if (c >= A.ncol) return;
    const int64_t ld = A.ld;
    const DevParams<NF>& p = A.p;
    const NF dt = A.dt;

    // shared addresses of this thread's strip: field f is  strip0 + f * B * ES ; the two Kf slots alternate
    const uint32_t strip0 = smem_base + (uint32_t)((SM::METRICS + threadIdx.x) * ES);
    uint32_t kf_cur = strip0 + B * ES, kf_prv = strip0;   // Kf[m] is written to kf_cur, which still holds Kf[m-2] ; kf_prv holds Kf[m-1]
    const uint32_t ring0 = smem_base + (uint32_t)((SM::METRICS + SM::STRIP + threadIdx.x) * ES);
    auto rd = [&](int f) { return ldsv(strip0 + (uint32_t)(f * B * ES), (NF*)nullptr); };
    auto wr = [&](int f, NF v) { sts(strip0 + (uint32_t)(f * B * ES), v); };
    // U and sat of layer k while it is in flight (iterations k-3 .. k+2)
    auto ringU = [&](int k) { return ring0 + (uint32_t)((k & (EULER_RD - 1)) * B * ES); };
    auto ringS = [&](int k) { return ring0 + (uint32_t)(((k & (EULER_RD - 1)) + EULER_RD) * B * ES); };
    // the only strip values read before the pipeline has written them: Kf[0] = 0 (never written by the reference)
    // and the (unused) differences formed against the not yet existing layer 0
    sts(kf_cur, NF(0)); sts(kf_prv, NF(0)); wr(EF_QH, NF(0)); wr(EF_G, NF(0)); wr(EF_KC, NF(0));

    auto bc_input = [&](int slot) -> NF {
        const int kind = A.bc[slot].kind;
        if (kind == TRM_BC_DEFAULT) return NF(0);
        return eval_input(A.in[A.bc[slot].input], c, kind == TRM_BC_FLUX ? A.t_b : A.t_x, kind == TRM_BC_FLUX ? 1 : 0);
    };
    const NF wtx = (RICH && !LOAD) ? A.xWt[c] : NF(0);
    // per-column surface temperature vector (device memory, or mapped host memory bound with trm_bind_host_io): fetched
    // now, together with the first layer, and consumed ~nz iterations later when the halo above the surface is formed
    const bool bct_pre = A.bct_pre != 0;
    if (bct_pre) cp_async<ES>(strip0 + (uint32_t)(EF_BCT * B * ES), A.in[A.bc[TRM_BC_TEMPERATURE_TOP].input].a + c);

    // ---- raw prefetch ring: layer k lives in ring slot (k & 3); one cp.async group per layer ----
    // element offsets fit 32 bits (the launcher checks nz * ld < 2^32): one IMAD.WIDE.U32 per address
    uint32_t oin = (uint32_t)c;
    const uint32_t xring0 = ring0 + (uint32_t)(SM::RING * ES);
    uint32_t oext = (uint32_t)c;   // (Heun stage 2) element offset of the next layer of the extra ring
    // (recompute protocol, stage 2) k1 of layer k lives next to U / sat of layer k, from its prefetch until its update
    auto k1U_slot = [&](int k) { return xring0 + (uint32_t)((k & (EULER_RD - 1)) * B * ES); };
    auto k1S_slot = [&](int k) { return xring0 + (uint32_t)(((k & (EULER_RD - 1)) + EULER_RD) * B * ES); };
    auto prefetch = [&](int k, bool always = false) {
        if (always || k <= nz) {
            cp_async<ES>(ringU(k), A.xU + oin);
            cp_async<ES>(ringS(k), A.xS + oin);
            if (LOAD) {
                const uint32_t dst = ring0 + (uint32_t)((2 * EULER_RD + (k & (PF_ - 1))) * B * ES);
                cp_async<ES>(dst, A.xT + oin);
                cp_async<ES>(dst + PF_ * B * ES, A.xL + oin);
                if (RICH) cp_async<ES>(dst + 2 * PF_ * B * ES, A.xP + oin);
            }
            if (H2 && RC) { cp_async<ES>(k1U_slot(k), A.k1U + oin); if (RICH) cp_async<ES>(k1S_slot(k), A.k1S + oin); }
            oin += (uint32_t)ld;
        }
        if (H2 && !RC) {
            // Heun stage 2: k1 and the base state of layer k-2, needed when that layer is updated (iteration k)
            const int kk = k - 2;
            if (kk >= 1 && kk <= nz) {
                const uint32_t dst = xring0 + (uint32_t)((kk & (PF_ - 1)) * B * ES);
                cp_async<ES>(dst, A.k1U + oext);
                cp_async<ES>(dst + 2 * PF_ * B * ES, A.bU + oext);
                if (RICH) { cp_async<ES>(dst + PF_ * B * ES, A.k1S + oext); cp_async<ES>(dst + 3 * PF_ * B * ES, A.bS + oext); }
                oext += (uint32_t)ld;
            }
        }
        cp_async_commit();   // (an empty group when nothing is left keeps the group count in step with the iteration count)
    };
#pragma unroll
    for (int k = 1; k <= DIST_; ++k) prefetch(k);

    NF carry = NF(0);          // over-saturation handed to the layer above (upward sweep of adjust_saturation_profile!)
    // a negative saturation needs the downward sweep -> slow path. Fast math ORs the sign words of the updated
    // saturations (one integer instruction per layer; -0 takes the slow path too, which gives the same result)
    int neg_acc = 0;
    auto any_neg = [&]() { return neg_acc < 0; };
    int idx = 0;               // lowest unsaturated layer (compute_water_table!), 0 = not found yet
    NF wt_new = NF(0);
    NF Sx_new = NF(0);
    if (RICH && !H1) Sx_new = A.bSx[c] + NF(0) * dt;   // surface_excess_water tendency is zero (soil_hydrology.jl:260-267)
    uint32_t oout = (uint32_t)c;   // element offset of layer m-2
    if (LAND && has_veg(A)) wr(EF_BETA, NF(0));   // vegetated LandModel: soil moisture limiting factor (plant_available_water.jl:31-35)
    // recompute protocol, stage 2: upward-sweep carry of the stage state being rebuilt ; columns whose stage state was stored
    NF carry1 = NF(0);
    const bool flagged = (H2 && RC && RICH) ? (A.hflag_in[c] != NF(0)) : false;
    uint32_t oent = (uint32_t)c;   // element offset of the entering layer (stored stage state of a flagged column)
    // Flux boundary conditions (compute_z_bcs!, abstract_timestepper.jl:69 ): the fluxes of the time-n state.
    // LandModel: ground heat flux and infiltration left by surface_kernel (land_model.jl:56-62)
    auto apply_top_flux = [&](NF& tU, NF& tS) {
        if (LAND) { const NF G_top = A.G[c]; tU -= G_top / met.dzc(nz); if (RICH) { const NF infil_top = A.infil[c]; tS -= (-infil_top) / met.dzc(nz); } }
        else {
            if (A.bc[TRM_BC_ENERGY_TOP].kind == TRM_BC_FLUX) tU -= bc_input(TRM_BC_ENERGY_TOP) / met.dzc(nz);
            if (RICH && A.bc[TRM_BC_SATURATION_TOP].kind == TRM_BC_FLUX) tS -= bc_input(TRM_BC_SATURATION_TOP) / met.dzc(nz);
        }
    };
    auto apply_bottom_flux = [&](NF& tU, NF& tS) {
        if (A.bc[TRM_BC_ENERGY_BOTTOM].kind == TRM_BC_FLUX) tU += bc_input(TRM_BC_ENERGY_BOTTOM) / met.dzc(1);
        if (RICH && A.bc[TRM_BC_SATURATION_BOTTOM].kind == TRM_BC_FLUX) tS += bc_input(TRM_BC_SATURATION_BOTTOM) / met.dzc(1);
    };

    // One pipeline iteration. The position tag (compile time, enum IterPos) says where in the column the iteration sits,
    // so that every layer-index special case below (halo, boundary face, Flux BC, does the prefetched layer exist) is
    // statically false / true; IP_GEN keeps them as run-time tests on m (columns of fewer than four layers, faithful
    // math). All instantiations share this source.
    auto iterate = [&](const int m, auto pos_tag) {
        constexpr int POS = decltype(pos_tag)::value;
        constexpr bool GEN = POS == IP_GEN;
        const bool m_is_1 = GEN ? m == 1 : POS == IP_M1;
        const bool m_is_nz = GEN ? m == nz : POS == IP_NZ;
        const bool enters = GEN ? m <= nz : POS <= IP_NZ;             // layer m exists and enters the pipeline
        const bool is_halo = GEN ? m == nz + 1 : POS == IP_HALO;      // the halo cell above the surface enters
        const bool has_face = GEN ? m <= nz + 1 : POS != IP_LAST;     // face m carries a heat flux / head gradient
        const bool has_darcy = GEN ? m >= 2 : POS != IP_M1;           // face m-1 carries a Darcy flux
        const bool updates = GEN ? m >= 3 : POS >= IP_M3;             // layer j = m-2 is updated
        const bool j_is_1 = GEN ? m == 3 : POS == IP_M3;
        const bool j_is_nz = GEN ? m == nz + 2 : POS == IP_LAST;
        const bool j_ge_2 = GEN ? m >= 4 : POS > IP_M3;
        if (POS == IP_INNER) prefetch(m + DIST_, true);
        else if (POS >= IP_INNER_NP && !(H2 && !RC)) cp_async_commit();   // nothing left to prefetch: the (empty) group keeps the count in step
        else prefetch(m + DIST_, false);
        // ---- layer m (or the halo above the surface) enters the pipeline ----
        NF Tn, Pn = NF(0), kapn, Kfn = NF(0);     // T, psi, kappa of layer m ; Kf[m]
        const NF Kf1 = RICH ? ldsv(kf_prv, (NF*)nullptr) : NF(0);   // Kf[m-1]
        if (enters) {
            cp_async_wait<DIST_>();   // all but the DIST most recent groups have landed: layer m is in the ring
            NF Ur = ldsv(ringU(m), (NF*)nullptr);
            NF sr = ldsv(ringS(m), (NF*)nullptr);
            if (H2 && RC) {
                // stage state of layer m: what stage 1 computed from the same base state and k1, and did not store
                // (explicit_step! + upward sweep of adjust_saturation_profile! of the stage copy, heun.jl:45-49)
                if (flagged) { Ur = A.sU[oent]; sr = A.sS[oent]; }
                else {
                    NF t1U = ldsv(k1U_slot(m), (NF*)nullptr), t1S = RICH ? ldsv(k1S_slot(m), (NF*)nullptr) : NF(0);
                    if (m_is_nz) apply_top_flux(t1U, t1S);
                    if (m_is_1) apply_bottom_flux(t1U, t1S);
                    Ur = Ur + t1U * dt;
                    if (RICH) {
                        sr = sr + t1S * dt;
                        sr = sr + carry1;
                        if (!m_is_nz) {
                            const NF e = Mx::pos(sr - 1);
                            sr -= e;
                            carry1 = FAST ? e * met.dzc(m) * met.rdzc(m + 1) : e * met.dzc(m) / met.dzc(m + 1);
                        }
                        if (!FAST && !m_is_1) sr = sr + jmax(-sr, NF(0));
                        if (m_is_nz) sr -= Mx::pos(sr - 1);       // top excess of the stage copy (its surface excess water is not used)
                        if (!FAST && m_is_1) sr = jmax(sr, NF(0));
                    }
                }
                oent += (uint32_t)ld;
            }
            NF ln;
            if (LOAD) {
                const uint32_t src = ring0 + (uint32_t)((2 * EULER_RD + (m & (PF_ - 1))) * B * ES);
                Tn = ldsv(src, (NF*)nullptr);
                ln = ldsv(src + PF_ * B * ES, (NF*)nullptr);
                if (RICH) Pn = ldsv(src + 2 * PF_ * B * ES, (NF*)nullptr);
            } else {
                energy_to_temperature<NF, FAST>(p, Ur, sr, Tn, ln);
                if (RICH) Pn = pressure_head<NF, FAST, VG2>(p, sr, wtx, met.zC(m), met.psiz(m));
            }
            kapn = FAST ? thermal_conductivity_fast(p, sr, ln) : thermal_conductivity(p, sr, ln);
            if (RICH) {
                // cell conductivity and face conductivity Kf[m], soil_hydrology.jl:249-276
                const NF Kcn = cell_conductivity<NF, FAST, VG2>(p, sr, ln);
                Kfn = (m_is_1 || m_is_nz) ? Kcn : Mx::mn(Kcn, rd(EF_KC));   // Kf[1] = Kc[1], Kf[Nz] = Kc[Nz]
                wr(EF_KC, Kcn);
            }
        } else if (is_halo) {   // halo above the surface, built from layer nz (prv)
            Tn = halo_value(A.bc[TRM_BC_TEMPERATURE_TOP].kind, rd(EF_T), bct_pre ? rd(EF_BCT) : bc_input(TRM_BC_TEMPERATURE_TOP), met.dzf(nz + 1), true);
            // conductivity of the halo cell: same (sat, liq) as layer nz when the saturation halo is a copy, else
            // sat = 0, for which the liquid fraction drops out of the constituent sum
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapn = copy ? rd(EF_KAP) : (FAST ? thermal_conductivity_fast(p, NF(0), NF(1)) : thermal_conductivity(p, NF(0), NF(1)));
            if (RICH) Pn = halo_value(A.bc[TRM_BC_PRESSURE_TOP].kind, rd(EF_P), bc_input(TRM_BC_PRESSURE_TOP), met.dzf(nz + 1), true);
            Kfn = Kf1;              // Kf[Nz+1] = Kf[Nz] ; Kf[Nz+2] is a halo face (0)
        } else {
            Tn = NF(0); kapn = NF(0);
        }
        // ---- lower neighbour of layer m: layer m-1, or the halo below the bottom layer for m = 1 ----
        NF Tp, kapp, Pp = NF(0);
        if (m_is_1) {
            Tp = halo_value(A.bc[TRM_BC_TEMPERATURE_BOTTOM].kind, Tn, bc_input(TRM_BC_TEMPERATURE_BOTTOM), met.dzf(1), false);
            const bool copy = RICH || p.sat_halo == TRM_HALO_COPY;
            kapp = copy ? kapn : (FAST ? thermal_conductivity_fast(p, NF(0), NF(1)) : thermal_conductivity(p, NF(0), NF(1)));
            if (RICH) Pp = halo_value(A.bc[TRM_BC_PRESSURE_BOTTOM].kind, Pn, bc_input(TRM_BC_PRESSURE_BOTTOM), met.dzf(1), false);
        } else {
            Tp = rd(EF_T); kapp = rd(EF_KAP);
            if (RICH) Pp = rd(EF_P);
        }
        // ---- heat flux and head gradient at face m (diffusive_heat_flux, soil_energy.jl:134-149) ----
        NF qhn = NF(0), gn = NF(0);
        if (has_face) {
            qhn = -((kapn + kapp) / 2) * ((Tn - Tp) * met.rdzf(m));
            if (RICH) gn = (Pn - Pp) * met.rdzf(m);
        }
        const NF dqhn = qhn - rd(EF_QH);
        // ---- Darcy flux at face m-1 (darcy_flux, soil_hydrology_rre.jl:119-131) ----
        NF qdn = NF(0);
        if (RICH && has_darcy) {
            const NF g = rd(EF_G);
            const NF Kf2 = ldsv(kf_cur, (NF*)nullptr);   // Kf[m-2] (0 for m = 2: Kf[0] is never written by the reference)
            NF Kk;
            if (FAST) Kk = Mx::mn(Kf1, g < 0 ? Kf2 : Kfn);
            else Kk = (g < 0 ? jmin(Kf2, Kf1) : NF(0)) + (g >= 0 ? jmin(Kf1, Kfn) : NF(0));
            qdn = -Kk * g;
        }

        // ---- LandModel surface processes, once the top layer is the one about to be updated ----
        // (LandModel: the fluxes coupling the surface to the top soil layer were evaluated by surface_kernel on the time-n
        //  state; Heun stage 2 applies the same time-n fluxes, heun.jl:63-66 -- see apply_top_flux)

        if (updates) {
            // ---- tendencies of layer j = m-2 ----
            const int j = m - 2;
            const uint32_t o = oout;
            oout += (uint32_t)ld;
            NF tU = -(rd(EF_DQH) * met.rdzc(j));                          // soil_energy.jl:112-131
            NF tS = NF(0);
            if (RICH) {
                const NF dth = -((qdn - rd(EF_QD)) * met.rdzc(j)) + NF(0) + p.vwcf;   // soil_hydrology_rre.jl:95-117
                tS = FAST ? dth * p.rpor : dth / p.por;                          // soil_hydrology.jl:222-237
            }
            NF Ub, sb;   // base state the update is applied to
            if (H2 && RC) {   // average_tendencies! (heun.jl:27-35) ; k1 and the base state sit in the 8-deep rings
                tU = (ldsv(k1U_slot(j), (NF*)nullptr) + tU) / 2;
                Ub = ldsv(ringU(j), (NF*)nullptr); sb = ldsv(ringS(j), (NF*)nullptr);
                if (RICH) tS = (ldsv(k1S_slot(j), (NF*)nullptr) + tS) / 2;
            } else if (H2) {   // average_tendencies! (heun.jl:27-35) with k1 of stage 1 ; the base is the state at time n
                const uint32_t x = xring0 + (uint32_t)((j & (PF_ - 1)) * B * ES);
                tU = (ldsv(x, (NF*)nullptr) + tU) / 2;
                Ub = ldsv(x + 2 * PF_ * B * ES, (NF*)nullptr);
                if (RICH) { tS = (ldsv(x + PF_ * B * ES, (NF*)nullptr) + tS) / 2; sb = ldsv(x + 3 * PF_ * B * ES, (NF*)nullptr); }
                else sb = ldsv(ringS(j), (NF*)nullptr);   // NoFlow: the saturation is not a prognostic variable
            } else {
                if (H1) { A.oTU[o] = tU; if (RICH) A.oTS[o] = tS; }   // k1, before the Flux BCs
                Ub = ldsv(ringU(j), (NF*)nullptr); sb = ldsv(ringS(j), (NF*)nullptr);
            }
            if (j_is_nz) apply_top_flux(tU, tS);
            if (j_is_1) apply_bottom_flux(tU, tS);
            // ---- explicit step, abstract_timestepper.jl:113-141 ----
            const NF Un = Ub + tU * dt;
            NF sn = sb;
            if (RICH) {
                sn = sn + tS * dt;
                // ---- adjust_saturation_profile!, upward sweep (soil_hydrology.jl:192-199) ----
                sn = sn + carry;
                if (!j_is_nz) {
                    const NF e = Mx::pos(sn - 1);
                    sn -= e;
                    carry = FAST ? e * met.dzc(j) * met.rdzc(j + 1) : e * met.dzc(j) / met.dzc(j + 1);
                }
                if (FAST) neg_acc |= sign_word(sn); else if (sn < 0) neg_acc = -1;
            }
            // (fast math, every layer but the top one: the regular stores below ARE the raw values the slow path re-reads -- no
            //  surface excess, no clamp -- so the flag needs no branch here; the closure stores are overwritten later)
            // (recompute protocol, Heun stage 1: the stage state is not stored -- except its top layer, which the vegetation
            //  block of stage 2 reads through surface_kernel -- and a column that goes negative rebuilds it in its slow path)
            constexpr bool STORE = !(H1 && RC);
            if (!STORE && j_is_nz) A.yU[o] = Un;
            const bool raw_is_regular = FAST && (GEN ? false : !j_is_nz);
            if (RICH && !raw_is_regular && any_neg()) {
                // raw values for the slow path below (the downward sweep needs the whole profile)
                if (STORE) { A.yU[o] = Un; A.yS[o] = sn; }
            } else {
                if (RICH) {
                    // downward sweep with no deficit anywhere: sat += max(-sat, 0) (:201-208) is the identity
                    if (!FAST && j_ge_2) sn = sn + jmax(-sn, NF(0));
                    if (j_is_nz) {                         // top excess -> surface_excess_water (:210-214)
                        const NF e = Mx::pos(sn - 1);
                        sn -= e;
                        Sx_new += e * met.dzc(nz);
                    }
                    if (!FAST && j_is_1) sn = jmax(sn, NF(0));       // :216
                    if (STORE) stg(A.yS + o, sn);
                    else if (j_is_nz) A.yS[o] = sn;                  // (top layer, after the top excess went to the surface)
                    if (idx == 0 && (FAST ? below_one(sn) : sn < 1)) { idx = j; wt_new = met.zF(j); }   // compute_water_table!, kernel_utils.jl:7-16
                }
                if (STORE) stg(A.yU + o, Un);
                if (CLOSE || (LAND && has_veg(A))) {
                    NF Tc, lc;
                    energy_to_temperature<NF, FAST>(p, Un, sn, Tc, lc);
                    // vegetated LandModel: soil moisture limiting factor of the NEW state, for the surface block of the
                    // next evaluation (Integral(PAW * root_fraction / dz, dims = 3), accumulated bottom -> top)
                    if (LAND && has_veg(A))
                        wr(EF_BETA, rd(EF_BETA) + (FAST ? plant_available_water_fast(A.vp, p, sn, lc) * met.root(j)
                                                        : plant_available_water(A.vp, p, sn, lc) * met.root(j) / met.dzc(j) * met.dzc(j)));
                    if (CLOSE) {
                        stg(A.yT + o, Tc); stg(A.yL + o, lc);
                        if (j_is_nz && A.hio_out) A.hio_out[c] = Tc;   // ground temperature -> mapped host memory
                        // layers below the water table wait for it (written after the sweep)
                        if (RICH && idx != 0) stg(A.yP + o, pressure_head<NF, FAST, VG2>(p, sn, wt_new, met.zC(j), met.psiz(j)));
                    }
                }
            }
        }
        // ---- what later iterations need from this one ----
        wr(EF_T, Tn); wr(EF_KAP, kapn); wr(EF_QH, qhn); wr(EF_DQH, dqhn);
        if (RICH) { wr(EF_P, Pn); sts(kf_cur, Kfn); wr(EF_G, gn); wr(EF_QD, qdn); }
        const uint32_t t = kf_cur; kf_cur = kf_prv; kf_prv = t;
    };
    if (SPEC && nz >= 4) {
        iterate(1, IterTag<IP_M1>{}); iterate(2, IterTag<IP_M2>{}); iterate(3, IterTag<IP_M3>{});
        int m = 4;
#pragma unroll 1
        for (; m <= nz - DIST_; ++m) iterate(m, IterTag<IP_INNER>{});
#pragma unroll 1
        for (; m < nz; ++m) iterate(m, IterTag<IP_INNER_NP>{});
        iterate(nz, IterTag<IP_NZ>{}); iterate(nz + 1, IterTag<IP_HALO>{}); iterate(nz + 2, IterTag<IP_LAST>{});
    } else {
        int m = 1;
#pragma unroll 1
        while (m <= nz + 2) {
            if (m >= 4 && m <= nz - DIST_) {
#pragma unroll 1
                do { iterate(m, IterTag<IP_INNER>{}); ++m; } while (m <= nz - DIST_);
            } else {
                iterate(m, IterTag<IP_GEN>{});
                ++m;
            }
        }
    }
    if (LAND && has_veg(A) && !(RICH && any_neg())) A.ybeta[c] = rd(EF_BETA);
    if (!RICH) return;

    if (!any_neg()) {
        if (idx == 0) { idx = nz + 1; wt_new = met.zF(nz + 1); }   // all saturated: z of the surface (halo cell / fallback give the same)
        A.yWt[c] = wt_new;
        if (H1 && RC) A.hflag_out[c] = NF(0);   // stage 2 rebuilds the stage state of this column from the base state and k1
        if (H1) return;            // the stage state needs no surface excess water and no closure fields
        A.ySx[c] = Sx_new;
        // pressure head of the saturated zone below the water table: psi_m(sat >= 1) is a constant
        const NF psat = swrc_inverse<NF, FAST, VG2>(p, p.por, p.por);
        int64_t o = c;
        if (FAST) {
            // (wt - zC) + psat + (zC - zref) below the water table: one value for the whole saturated zone (the layer
            // centres cancel; the two forms differ by rounding only)
            const NF pconst = (wt_new - met.zF(nz + 1)) + psat;
#pragma unroll 1
            for (int k = 1; k < idx && k <= nz; ++k, o += ld) stg(A.yP + o, pconst);
        } else {
#pragma unroll 1
            for (int k = 1; k < idx && k <= nz; ++k, o += ld) A.yP[o] = Mx::mx(NF(0), wt_new - met.zC(k)) + psat + met.psiz(k);
        }
        return;
    }
    if (H1 && RC) { heun1_slow_column<NF, FAST, Met, LAND>(A, met, c); return; }
    // ---- slow path: a layer went negative. Downward sweep (soil_hydrology.jl:201-216) top -> bottom on the raw
    //      profile this thread just stored, then water table and closures bottom -> top. ----
    {
        NF carry_dn = NF(0);
#pragma unroll 1
        for (int k = nz; k >= 1; --k) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            NF s = A.yS[o];
            if (k < nz) s -= carry_dn;
            if (k >= 2) {
                const NF d = jmax(-s, NF(0));
                s += d;
                carry_dn = d * met.dzc(k) / met.dzc(k - 1);
            }
            if (k == nz) {
                const NF e = jmax(s - 1, NF(0));
                s -= e;
                Sx_new += e * met.dzc(nz);
            }
            if (k == 1) s = jmax(s, NF(0));
            A.yS[o] = s;
        }
        idx = 0;
#pragma unroll 1
        for (int k = 1; k <= nz; ++k) if (idx == 0 && A.yS[(int64_t)(k - 1) * ld + c] < 1) idx = k;
        if (idx == 0) idx = nz + 1;
        wt_new = met.zF(idx);
        A.yWt[c] = wt_new;
        if (H1 && !(LAND && has_veg(A))) return;
        if (!H1) A.ySx[c] = Sx_new;
        NF beta = NF(0);
#pragma unroll 1
        for (int k = 1; k <= nz; ++k) {
            const int64_t o = (int64_t)(k - 1) * ld + c;
            NF s = A.yS[o], U = A.yU[o], Tc, lc;
            energy_to_temperature<NF, FAST>(p, U, s, Tc, lc);
            if (LAND && has_veg(A))
                beta += FAST ? plant_available_water_fast(A.vp, p, s, lc) * met.root(k) : plant_available_water(A.vp, p, s, lc) * met.root(k) / met.dzc(k) * met.dzc(k);
            if (H1) continue;   // the stage state keeps no closure fields
            A.yT[o] = Tc; A.yL[o] = lc;
            if (k == nz && A.hio_out) A.hio_out[c] = Tc;
            A.yP[o] = pressure_head<NF, FAST, VG2>(p, s, wt_new, met.zC(k), met.psiz(k));
        }
        if (LAND && has_veg(A)) A.ybeta[c] = beta;
    }
}

}  // namespace trm
