// terrarium_b200.cu -- the C ABI of include/terrarium_b200.h on top of the fused stage kernels.
//
// Host orchestration only: device buffers in [layer][column] SoA, parameter conversion to the
// handle's number format, the ForwardEuler / Heun stage sequencing of
// src/timesteppers/forward_euler.jl:19-31 and heun.jl:37-71, host <-> device field copies.
// There is NO CPU implementation behind these entry points: without a CUDA device trm_create
// fails with TRM_ERR_NO_DEVICE.
#include <nvtx3/nvToolsExt.h>
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "kernel_set.h"

namespace {

using namespace trm;

thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }

#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(TRM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));               \
    } while (0)

template <class NF>
__global__ void copy_row_kernel(int64_t n, const NF* __restrict__ x, NF* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i];
}

// ---- NCCL, bound at run time to the library the host process uses (no link-time dependency) ----
struct NcclApi {
    struct Id { char b[128]; };   // ncclUniqueId, passed by value
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* lib = RTLD_DEFAULT;
        if (!dlsym(RTLD_DEFAULT, "ncclAllReduce")) {
            lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!lib) return a;
        }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
        a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy;
        return a;
    }();
    return api;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMin = 3;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
int nccl_fail(int rc, const char* what) {
    const NcclApi& n = nccl_api();
    return fail(TRM_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error " + std::to_string(rc)));
}

// diag_out = [energy, water, tmin, tmax, smin, smax, nan, ncol]  <->  sums [energy, water, nan, ncol], minima [tmin, smin, -tmax, -smax]
static __global__ void diag_pack_kernel(const double* __restrict__ d, double* __restrict__ q) {
    if (threadIdx.x == 0) { q[0] = d[0]; q[1] = d[1]; q[2] = d[6]; q[3] = d[7]; q[4] = d[2]; q[5] = d[4]; q[6] = -d[3]; q[7] = -d[5]; }
}
static __global__ void diag_unpack_kernel(const double* __restrict__ q, double* __restrict__ d) {
    if (threadIdx.x == 0) { d[0] = q[0]; d[1] = q[1]; d[6] = q[2]; d[7] = q[3]; d[2] = q[4]; d[4] = q[5]; d[3] = -q[6]; d[5] = -q[7]; }
}

struct HandleBase {
    virtual ~HandleBase() {}
    virtual int set_field(int id, const void* host, int64_t count) = 0;
    virtual int get_field(int id, void* host, int64_t count) = 0;
    virtual int field_ptr(int id, void** p, int64_t* ld, int32_t* nrows) = 0;
    virtual int field_view(int id, const void** p, int64_t* ld, int32_t* nrows) = 0;
    virtual int set_input_const(int id, double v) = 0;
    virtual int set_input_field(int id, const void* v) = 0;
    virtual int set_input_field_pair(int id, const void* v0, const void* v1) = 0;
    virtual int set_input_sinusoid(int id, const void* mean, const void* amp, const void* phase, double period, double lo, double hi) = 0;
    virtual int get_input(int id, void* host, int64_t count) = 0;
    virtual int accumulate(int id, double w) = 0;
    virtual int get_accumulated(int id, void* host, int64_t count, double scale, int reset) = 0;
    virtual int set_input_table(int id, int nt, const double* times, const void* values, int kind = TRM_SRC_TABLE) = 0;
    virtual int input_ptr(int id, void** p) = 0;
    virtual int initialize() = 0;
    virtual int step(double dt, int64_t n) = 0;
    virtual int aux() = 0;
    virtual int tendencies() = 0;
    virtual int diagnostics(trm_diag* out, double** dev) = 0;
    virtual int set_block(int b) = 0;
    virtual int set_input_field_async(int id, const void* v) = 0;
    virtual int get_field_async(int id, void* host, int64_t count) = 0;
    virtual int step_async(double dt, int64_t n) = 0;
    virtual int sync_all() = 0;
    virtual int set_ring_index(const int64_t* idx, int64_t nring) = 0;
    virtual int get_field_ring(int id, void* host, int64_t count, double fill) = 0;
    virtual int set_field_ring(int id, const void* host, int64_t count) = 0;
    virtual int bind_host_io(int in_id, const void* host_in, int field_id, void* host_out, int nslots) = 0;
    virtual int host_io_wait(int64_t iteration) = 0;
    virtual int reset_state() = 0;
    virtual int diagnostics_allreduce(trm_diag* out) = 0;
    void* nccl_comm = nullptr; bool nccl_owned = false;
    virtual void set_clock(double t, int64_t it) = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    double time = 0.0;   // holds an NF value
    double t_inputs = 0.0;   // clock time of the last update_inputs! (state_variables.jl:154-162)
    int64_t iteration = 0;
    int64_t launches = 0;
    float last_ms = 0.f;
};

template <class NF>
struct Handle : HandleBase {
    trm_config cfg{};
    int nz = 0; int64_t nc = 0, ld = 0;
    bool land = false, richards = false, heun = false, fast = false, veg = false;
    int phys = PHYS_NOFLOW;
    int block = 128;          // threads per block of the register-streaming stage kernel
    // which implementation runs ForwardEuler stages (env TRM_KERNEL = smem | stream):
    //   1 smem:   streaming kernel with the pipeline state in shared memory (euler_kernel.cuh), the default
    //   0 stream: register-resident streaming kernel (stage_kernel.cuh), which also serves every other mode
    int euler_impl = 1;
    const KernelSet* ks = nullptr;
    DevParams<NF> p{};
    std::vector<void*> allocs;
    NF* metrics = nullptr;
    NF cmet[MET_COUNT * CMET_STRIDE] = {};   // compact rows handed to the staged kernels inside their parameters
    // 3-D [nz][ld]
    NF *U = nullptr, *T = nullptr, *Lq = nullptr, *S = nullptr, *P = nullptr, *Kf = nullptr;
    NF *tU = nullptr, *tS = nullptr, *gU = nullptr, *gS = nullptr;   // tendencies, Heun stage state
    // 2-D [ld]
    NF *Sx = nullptr, *Wt = nullptr, *gWt = nullptr;
    NF* hflag = nullptr;   // Heun recompute protocol: columns whose stage state was stored by stage 1 (negative saturation)
    NF* land2d[10] = {nullptr};   // Ts, G, SWup, LWup, Rnet, Hs, Hl, Egnd, infil, runoff
    // vegetated LandModel: 2-D fields in VegField order (first three prognostic), Heun stage values and k1 of the
    // prognostic triple, plant available water [nz][ld], host copy of the static root fraction per layer
    NF* veg2d[VF_COUNT] = {nullptr};
    NF *gveg[3] = {nullptr}, *tveg[3] = {nullptr};
    NF* paw = nullptr;
    NF *beta = nullptr, *gbeta = nullptr;   // soil moisture limiting factor of the model state / of the Heun stage state
    bool beta_stale = true;                  // `beta` does not belong to the stored state: recompute it before the next surface launch
    VegParams<NF> vp{};
    std::vector<NF> rootf;
    struct Input { int kind = TRM_SRC_CONST; double cval = 0, period = 1, lo = -INFINITY, hi = INFINITY; int nt = 0;
                   NF *a = nullptr, *b = nullptr, *c = nullptr; std::vector<double> times;   // (time axis of a table: host side only)
                   // asynchronous per-column field inputs are double buffered: `a` is what enqueued steps read,
                   // `a2` receives the next upload; ev_free[i] fires when buffer i is no longer read by any step
                   NF* a2 = nullptr; cudaEvent_t ev_free[2] = {nullptr, nullptr}; cudaEvent_t ev_ready = nullptr; int front = 0; };
    Input in[TRM_IN_COUNT];
    bool initialized = false;
    bool aux_stale = true;   // stored T / liq / psi are not closure(U, sat): the next stage must read them
    bool debug_nancheck = false;   // env TERRARIUM_DEBUG=true, the reference's debug switch (src/diagnostics/debugging.jl:1)
    bool force_load = false; // tuning knob (env TRM_FORCE_LOAD_AUX=1): always read T / liq / psi instead of recomputing them
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;          // copy streams of the asynchronous entry points
    cudaEvent_t ev_staged = nullptr, ev_out_done = nullptr;
    // Heun stage 2 of the vegetated LandModel: the vegetation block on the stage state (k2 of the three vegetation prognostics
    // only) and the soil stage kernel do not depend on each other -- the surface launch runs beside the stage kernel on a
    // stream of its own (both are latency bound on small and medium domains: N145 111 -> 102 us per step ; 10 M columns 9.33 -> 9.17 ms)
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    NF* staging = nullptr; size_t staging_count = 0;
    int64_t* ring_index = nullptr; int64_t nring = 0;   // ring-grid position of every owned column (ColumnRingGrid mask)
    NF* ring_buf = nullptr; size_t ring_count = 0;
    bool timing_open = false;
    bool fallback_logged = false;   // the staged kernels do not apply (fields of 2^32 elements or more): said once on stderr
    void log_fallback() {
        if (fallback_logged) return;
        fallback_logged = true;
        std::fprintf(stderr, "terrarium_b200: nz * ld = %llu >= 2^32 elements per field: the shared-memory staged kernels use 32-bit element "
                             "offsets, running the register-streaming stage kernel instead (about 1.4x slower per step)\n",
                     (unsigned long long)nz * (unsigned long long)ld);
    }
    // per-step exchange through mapped host memory (trm_bind_host_io): step k -> k+1 reads input `in_id` from slot
    // k % nslots of `in_dev` and writes field `out_id` of the new state to the same slot of `out_dev`; ev[slot] fires when
    // that step has completed
    struct HostIO { int in_id = -1, out_id = -1, nslots = 0; const NF* in_dev = nullptr; NF* out_dev = nullptr;
                    std::vector<cudaEvent_t> ev; int64_t first_iter = 0; } hio;
    double* diag_partial = nullptr; double* diag_out = nullptr; int diag_blocks = 0;

    ~Handle() override {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        for (void* q : allocs) cudaFree(q);
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_out) cudaStreamSynchronize(s_out);
        if (s_aux) cudaStreamSynchronize(s_aux);
        for (cudaEvent_t e : {ev0, ev1, ev_staged, ev_out_done, ev_fork, ev_join}) if (e) cudaEventDestroy(e);
        for (Input& s : in) for (cudaEvent_t e : {s.ev_free[0], s.ev_free[1], s.ev_ready}) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : hio.ev) if (e) cudaEventDestroy(e);
        if (nccl_comm && nccl_owned && nccl_api().ok) nccl_api().CommDestroy(nccl_comm);
        for (cudaStream_t s : {stream, s_in, s_out, s_aux}) if (s) cudaStreamDestroy(s);
    }

    template <class X> int dalloc(X** out, size_t count, bool zero = true) {
        void* q = nullptr;
        CU(cudaMalloc(&q, count * sizeof(X)));
        allocs.push_back(q);
        if (zero) CU(cudaMemsetAsync(q, 0, count * sizeof(X), stream));
        *out = (X*)q;
        return TRM_OK;
    }
    void dfree(void* q) {
        for (size_t i = 0; i < allocs.size(); ++i) if (allocs[i] == q) { allocs.erase(allocs.begin() + i); break; }
        cudaFree(q);
    }

    int setup(const trm_config& c) {
        cfg = c; nz = c.nz; nc = c.ncol; device = c.device;
        land = c.model == TRM_MODEL_LAND; richards = c.hydrology == TRM_RICHARDS; heun = c.timestepper == TRM_HEUN;
        fast = c.math == TRM_MATH_FAST;
        if (c.vegetation != TRM_VEG_NONE && c.vegetation != TRM_VEG_CARBON) return fail(TRM_ERR_INVALID, "bad vegetation code");
        veg = land && c.vegetation == TRM_VEG_CARBON;   // (the field is ignored for a SoilModel)
        { const char* e = std::getenv("TRM_FORCE_LOAD_AUX"); force_load = e && e[0] == '1'; }
        { const char* e = std::getenv("TERRARIUM_DEBUG"); debug_nancheck = e && std::string(e) == "true"; }
        if (const char* e = std::getenv("TRM_KERNEL")) {
            const std::string k(e);
            if (k == "stream") euler_impl = 0; else if (k == "smem") euler_impl = 1;
            else return fail(TRM_ERR_INVALID, "TRM_KERNEL must be smem or stream");
        }
        phys = land ? (richards ? PHYS_LAND : PHYS_LAND_NOFLOW) : (richards ? PHYS_RICHARDS : PHYS_NOFLOW);
        ks = fast ? &kernels_fast() : &kernels_faithful();
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(TRM_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
        }
        if (device < 0 || device >= ndev) return fail(TRM_ERR_INVALID, "device ordinal out of range");
        CU(cudaSetDevice(device));
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&ev0)); CU(cudaEventCreate(&ev1));
        ld = (nc + 63) / 64 * 64;   // rows are 256 B (FP32) / 512 B (FP64) aligned

        // ---- grid metrics in NF, exactly as Oceananigans builds them from the NF faces:
        //      halo faces extend with the edge spacing, centres are face midpoints.
        std::vector<NF> m((size_t)MET_COUNT * MET_STRIDE, NF(0));
        NF *zF = m.data() + MET_ZF * MET_STRIDE, *zC = m.data() + MET_ZC * MET_STRIDE, *dzc = m.data() + MET_DZC * MET_STRIDE,
           *rdzc = m.data() + MET_RDZC * MET_STRIDE, *dzf = m.data() + MET_DZF * MET_STRIDE, *rdzf = m.data() + MET_RDZF * MET_STRIDE,
           *psiz = m.data() + MET_PSIZ * MET_STRIDE;
        for (int k = 1; k <= nz + 1; ++k) zF[k] = (NF)c.z_faces[k - 1];
        zF[0] = zF[1] - (zF[2] - zF[1]);
        zF[nz + 2] = zF[nz + 1] + (zF[nz + 1] - zF[nz]);
        for (int k = 0; k <= nz + 1; ++k) { zC[k] = (zF[k + 1] + zF[k]) / 2; dzc[k] = zF[k + 1] - zF[k]; rdzc[k] = 1 / dzc[k]; }
        for (int k = 1; k <= nz + 1; ++k) { dzf[k] = zC[k] - zC[k - 1]; rdzf[k] = 1 / dzf[k]; }
        for (int k = 0; k <= nz + 1; ++k) psiz[k] = zC[k] - zF[nz + 1];   // elevation head relative to the surface
        if (veg) {
            // root_fraction(rootdist, grid, ...), root_distribution.jl:47-56: density at the cell centres times the layer
            // thickness, normalised by its sum over the column
            NF* root = m.data() + MET_ROOT * MET_STRIDE;
            const NF a = (NF)c.params.root_a, b = (NF)c.params.root_b;
            NF sum = 0;
            for (int k = 1; k <= nz; ++k) { root[k] = NF(0.5) * (a * std::exp(a * zC[k]) + b * std::exp(b * zC[k])) * dzc[k]; sum += root[k]; }
            for (int k = 1; k <= nz; ++k) root[k] = root[k] / sum;
            rootf.assign(root + 1, root + nz + 1);
        }
        if (nz + 3 <= CMET_STRIDE)   // compact copy for the kernel parameters (MetricsC)
            for (int q = 0; q < MET_COUNT; ++q) std::copy_n(m.data() + (size_t)q * MET_STRIDE, nz + 3, cmet + (size_t)q * CMET_STRIDE);
        if (int rc = dalloc(&metrics, m.size())) return rc;
        CU(cudaMemcpyAsync(metrics, m.data(), m.size() * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));

        // ---- parameters in NF (same expression order as the reference constructors / kernels)
        const trm_params& q = c.params;
        NF por_m = (NF)q.mineral_porosity, por_o = (NF)q.organic_porosity;
        p.org = (NF)q.rho_soc / ((1 - por_o) * (NF)q.rho_org);      // homogeneous_strat.jl:34-45
        p.por = (1 - p.org) * por_m + p.org * por_o;                // homogeneous_strat.jl:52-61
        p.L = (NF)q.rho_w * (NF)q.Lsl;
        for (int i = 0; i < 5; ++i) { p.sqk[i] = std::sqrt((NF)q.kappa[i]); p.hc[i] = (NF)q.heatcap[i]; }
        p.Ksat = (NF)q.K_sat; p.vg_alpha = (NF)q.vg_alpha; p.vg_n = (NF)q.vg_n; p.bc_psis = (NF)q.bc_psis; p.bc_lambda = (NF)q.bc_lambda;
        p.theta_res = (NF)q.theta_res; p.Omega = (NF)q.impedance; p.vwcf = (NF)q.vwc_forcing;
        p.rho_a = (NF)q.rho_a; p.c_a = (NF)q.c_a; p.Llg = (NF)q.Llg; p.Tref = (NF)q.Tref; p.sigma = (NF)q.sigma; p.eps_mw = (NF)q.eps_mw;
        p.albedo = (NF)q.albedo; p.emis = (NF)q.emissivity; p.kappa_skin = (NF)q.kappa_skin; p.C_h = (NF)q.C_h; p.Vmin = (NF)q.min_windspeed;
        p.tau_r = (NF)q.tau_r; p.beta = (NF)q.evap_beta;
        p.rpor = 1 / p.por; p.neg_inv_alpha = -1 / p.vg_alpha;
        p.vg_k_exp1 = p.vg_n / (p.vg_n + 1); p.vg_k_exp2 = (p.vg_n - 1) / p.vg_n;
        { NF mm = 1 - 1 / p.vg_n; p.vg_inv_m_neg = -1 / mm; p.vg_inv_n = 1 / p.vg_n; }
        { NF solid = 1 - p.por, organic = solid * p.org, mineral = solid * (1 - p.org);
          p.hc_wi = p.hc[0] - p.hc[1]; p.hc_ia = p.hc[1] - p.hc[2]; p.hc_base = p.hc[2] * p.por + (p.hc[3] * mineral + p.hc[4] * organic);
          p.sqk_wi = p.sqk[0] - p.sqk[1]; p.sqk_ia = p.sqk[1] - p.sqk[2]; p.sqk_base = p.sqk[2] * p.por + (p.sqk[3] * mineral + p.sqk[4] * organic); }
        p.r_thspan = 1 / (p.por - p.theta_res); p.se_off = -p.theta_res * p.r_thspan;
        p.swrc = c.swrc; p.unsat_k = c.unsat_k; p.sat_halo = c.sat_halo; p.skin = c.skin;
        if (c.ground_resistance != TRM_GROUND_RES_CONSTANT && c.ground_resistance != TRM_GROUND_RES_SOIL_MOISTURE) return fail(TRM_ERR_INVALID, "bad ground_resistance code");
        p.ground_res = c.ground_resistance; p.th_fc = (NF)q.field_capacity;
        if ((c.albedo_kind | c.radiative | c.turbulent) & ~1 || c.reserved0 != 0) return fail(TRM_ERR_INVALID, "bad albedo_kind / radiative / turbulent code (or reserved0 != 0)");
        p.albedo_kind = c.albedo_kind; p.rad_kind = c.radiative; p.turb_kind = c.turbulent;
        p.vg_n_is_2 = (p.vg_n == NF(2)) ? 1 : 0;
        { const NF k = 1 / p.bc_lambda; const int ki = (int)std::lround((double)k); p.bc_k = (ki >= 1 && ki <= 8 && NF(ki) == k) ? ki : 0; }
        vp.th_fc = (NF)q.field_capacity; vp.th_wp = (NF)q.wilting_point; vp.C_mass = (NF)q.C_mass;
        vp.tau25 = (NF)q.tau25; vp.Kc25 = (NF)q.Kc25; vp.Ko25 = (NF)q.Ko25; vp.q10_tau = (NF)q.q10_tau; vp.q10_Kc = (NF)q.q10_Kc; vp.q10_Ko = (NF)q.q10_Ko;
        vp.alpha_leaf = (NF)q.alpha_leaf; vp.alpha_a = (NF)q.alpha_a; vp.alpha_C3 = (NF)q.alpha_C3; vp.cq = (NF)q.cq; vp.k_ext = (NF)q.k_ext;
        vp.T_CO2_high = (NF)q.T_CO2_high; vp.T_CO2_low = (NF)q.T_CO2_low; vp.T_photos_high = (NF)q.T_photos_high; vp.T_photos_low = (NF)q.T_photos_low;
        vp.theta_r = (NF)q.theta_r; vp.g1 = (NF)q.g1; vp.g_min = (NF)q.g_min; vp.cn_sapwood = (NF)q.cn_sapwood; vp.cn_root = (NF)q.cn_root; vp.aws = (NF)q.aws;
        vp.SLA = (NF)q.SLA; vp.awl = (NF)q.awl; vp.LAI_min = (NF)q.LAI_min; vp.LAI_max = (NF)q.LAI_max;
        vp.gamma_L = (NF)q.gamma_L; vp.gamma_R = (NF)q.gamma_R; vp.gamma_S = (NF)q.gamma_S; vp.nu_seed = (NF)q.nu_seed; vp.gamma_v = (NF)q.gamma_v_min;
        vp.r_paw_span = 1 / (vp.th_fc - vp.th_wp);
        vp.ln_q10_tau = std::log(vp.q10_tau); vp.ln_q10_Kc = std::log(vp.q10_Kc); vp.ln_q10_Ko = std::log(vp.q10_Ko);
        vp.ts_k1 = NF(2.0) * std::log(NF(1.0) / NF(0.99) - NF(1.0)) / (vp.T_CO2_low - vp.T_photos_low);
        vp.ts_k2 = NF(0.5) * (vp.T_CO2_low + vp.T_photos_low);
        vp.ts_k3 = std::log(NF(0.99) / NF(0.01)) / (vp.T_CO2_high - vp.T_photos_high);
        vp.alpha_int = (NF)q.alpha_int; vp.k_ext_can = (NF)q.k_ext_can; vp.w_can_max = (NF)q.w_can_max; vp.tau_w = (NF)q.tau_w; vp.C_can = (NF)q.C_can;

        // ---- fields
        const size_t n3 = (size_t)nz * ld;
        for (NF** f : {&U, &T, &Lq, &S}) if (int rc = dalloc(f, n3)) return rc;
        if (richards) { if (int rc = dalloc(&P, n3)) return rc; }
        if (int rc = dalloc(&Kf, (size_t)(nz + 1) * ld)) return rc;
        if (int rc = dalloc(&Sx, ld)) return rc;
        if (int rc = dalloc(&Wt, ld)) return rc;
        if (heun) {
            for (NF** f : {&tU, &gU}) if (int rc = dalloc(f, n3)) return rc;
            if (richards) { for (NF** f : {&tS, &gS}) if (int rc = dalloc(f, n3)) return rc; if (int rc = dalloc(&gWt, ld)) return rc; if (int rc = dalloc(&hflag, ld)) return rc; }
        }
        if (land) for (int i = 0; i < 10; ++i) if (int rc = dalloc(&land2d[i], ld)) return rc;
        if (veg) {
            for (int i = 0; i < VF_COUNT; ++i) if (int rc = dalloc(&veg2d[i], ld)) return rc;
            if (heun) for (int i = 0; i < 3; ++i) { if (int rc = dalloc(&gveg[i], ld)) return rc; if (int rc = dalloc(&tveg[i], ld)) return rc; }
            if (int rc = dalloc(&paw, n3)) return rc;
            if (int rc = dalloc(&beta, ld)) return rc;
            if (heun) { if (int rc = dalloc(&gbeta, ld)) return rc; }
        }
        // input defaults, prescribed_atmosphere.jl:89-99,147-149,192-195,220-224,10-14
        in[TRM_IN_AIR_TEMPERATURE].cval = 10; in[TRM_IN_AIR_PRESSURE].cval = 101325; in[TRM_IN_WINDSPEED].cval = 0.1;
        in[TRM_IN_SPECIFIC_HUMIDITY].cval = 1.0e-3; in[TRM_IN_SHORTWAVE_DOWN].cval = 300; in[TRM_IN_LONGWAVE_DOWN].cval = 50;
        in[TRM_IN_DAYTIME_LENGTH].cval = 12; in[TRM_IN_CO2].cval = 380;
        diag_blocks = (int)std::min<int64_t>((nc + 255) / 256, 148 * 8);
        if (int rc = dalloc(&diag_partial, (size_t)diag_blocks * 8)) return rc;
        if (int rc = dalloc(&diag_out, 8)) return rc;
        CU(cudaStreamSynchronize(stream));
        return TRM_OK;
    }

    // ------------------------------------------------------------------ field access
    struct FieldRef { NF* ptr; int nrows; bool writable; };
    FieldRef field(int id) {
        switch (id) {
            case TRM_F_INTERNAL_ENERGY: return {U, nz, true};
            case TRM_F_TEMPERATURE: return {T, nz, true};
            case TRM_F_LIQUID_WATER_FRACTION: return {Lq, nz, true};
            case TRM_F_SATURATION_WATER_ICE: return {S, nz, true};
            case TRM_F_PRESSURE_HEAD: return {P, nz, true};
            case TRM_F_HYDRAULIC_CONDUCTIVITY: return {Kf, nz + 1, false};
            case TRM_F_SURFACE_EXCESS_WATER: return {Sx, 1, true};
            case TRM_F_WATER_TABLE: return {Wt, 1, true};
            case TRM_F_GROUND_TEMPERATURE: return {T ? T + (size_t)(nz - 1) * ld : nullptr, 1, false};
            case TRM_F_TEND_INTERNAL_ENERGY: return {tU, nz, false};
            case TRM_F_TEND_SATURATION: return {tS, nz, false};
        }
        if (id >= TRM_F_SKIN_TEMPERATURE && id <= TRM_F_SURFACE_RUNOFF) return {land2d[id - TRM_F_SKIN_TEMPERATURE], 1, true};
        if (id >= TRM_F_CARBON_VEGETATION && id <= TRM_F_TRANSPIRATION) return {veg2d[id - TRM_F_CARBON_VEGETATION], 1, true};
        if (id == TRM_F_PLANT_AVAILABLE_WATER) return {paw, nz, false};
        return {nullptr, 0, false};
    }
    int set_field(int id, const void* host, int64_t count) override {
        FieldRef f = field(id);
        if (!f.ptr && (id == TRM_F_PRESSURE_HEAD || (id >= TRM_F_SKIN_TEMPERATURE && id <= TRM_F_SURFACE_RUNOFF)))
            return fail(TRM_ERR_INVALID, "field not defined for this model");
        if (!f.ptr || !f.writable) return fail(TRM_ERR_INVALID, "set_field: unknown or read-only field");
        if (count != (int64_t)f.nrows * nc) return fail(TRM_ERR_INVALID, f.nrows == 1 ? "set_field: count != ncol" : "set_field: count != nz*ncol");
        CU(cudaSetDevice(device));
        CU(cudaMemcpy2DAsync(f.ptr, ld * sizeof(NF), host, nc * sizeof(NF), nc * sizeof(NF), f.nrows, cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        aux_stale = true; beta_stale = true;
        return TRM_OK;
    }
    int get_field(int id, void* host, int64_t count) override {
        if (id == TRM_F_ROOT_FRACTION && veg) {   // static function of depth: the same profile in every column
            if (count != (int64_t)nz * nc) return fail(TRM_ERR_INVALID, "get_field: wrong element count");
            NF* h = (NF*)host;
            for (int k = 0; k < nz; ++k) std::fill(h + (size_t)k * nc, h + (size_t)(k + 1) * nc, rootf[k]);
            return TRM_OK;
        }
        FieldRef f = field(id);
        if (!f.ptr) return fail(TRM_ERR_INVALID, "get_field: unknown field or field not defined for this model");
        if (count != (int64_t)f.nrows * nc) return fail(TRM_ERR_INVALID, "get_field: wrong element count");
        CU(cudaSetDevice(device));
        CU(cudaMemcpy2DAsync(host, nc * sizeof(NF), f.ptr, ld * sizeof(NF), nc * sizeof(NF), f.nrows, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return TRM_OK;
    }
    int field_ptr(int id, void** q, int64_t* ld_out, int32_t* nrows) override { return field_borrow(id, q, ld_out, nrows, true); }
    int field_view(int id, const void** q, int64_t* ld_out, int32_t* nrows) override {
        void* w = nullptr;
        int rc = field_borrow(id, &w, ld_out, nrows, false);
        if (q) *q = w;
        return rc;
    }
    int field_borrow(int id, void** q, int64_t* ld_out, int32_t* nrows, bool may_write) {
        FieldRef f = field(id);
        if (!f.ptr) return fail(TRM_ERR_INVALID, "field_ptr: unknown field or field not defined for this model");
        if (q) *q = f.ptr;
        if (ld_out) *ld_out = ld;
        if (nrows) *nrows = f.nrows;
        // the caller may write through the pointer: treat the closure fields as user data from now on
        if (may_write && f.writable) { aux_stale = true; beta_stale = true; }
        return TRM_OK;
    }

    // ------------------------------------------------------------------ inputs
    int ensure(NF** q, size_t count) { if (*q) return TRM_OK; return dalloc(q, count); }
    int set_input_const(int id, double v) override { in[id].kind = TRM_SRC_CONST; in[id].cval = v; return TRM_OK; }
    int set_input_field(int id, const void* v) override {
        CU(cudaSetDevice(device));
        if (int rc = ensure(&in[id].a, ld)) return rc;
        CU(cudaMemcpyAsync(in[id].a, v, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        in[id].kind = TRM_SRC_FIELD;
        return TRM_OK;
    }
    int set_input_field_pair(int id, const void* v0, const void* v1) override {
        CU(cudaSetDevice(device));
        Input& s = in[id];
        if ((s.kind == TRM_SRC_TABLE || s.kind == TRM_SRC_RASTER) && s.a) { dfree(s.a); s.a = nullptr; }
        if (int rc = ensure(&s.a, ld)) return rc;
        if (int rc = ensure(&s.b, ld)) return rc;
        CU(cudaMemcpyAsync(s.a, v0, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(s.b, v1, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        s.kind = TRM_SRC_FIELD_PAIR;
        return TRM_OK;
    }
    int set_input_sinusoid(int id, const void* mean, const void* amp, const void* phase, double period, double lo, double hi) override {
        CU(cudaSetDevice(device));
        Input& s = in[id];
        if ((s.kind == TRM_SRC_TABLE || s.kind == TRM_SRC_RASTER) && s.a) { dfree(s.a); s.a = nullptr; }
        if (int rc = ensure(&s.a, ld)) return rc;
        if (int rc = ensure(&s.b, ld)) return rc;
        if (int rc = ensure(&s.c, ld)) return rc;
        CU(cudaMemcpyAsync(s.a, mean, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(s.b, amp, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(s.c, phase, nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        s.kind = TRM_SRC_SINUSOID; s.period = period; s.lo = lo; s.hi = hi;
        return TRM_OK;
    }
    int set_input_table(int id, int nt, const double* times, const void* values, int kind = TRM_SRC_TABLE) override {
        CU(cudaSetDevice(device));
        Input& s = in[id];
        if (s.a) { dfree(s.a); s.a = nullptr; }
        if (int rc = dalloc(&s.a, (size_t)nt * ld)) return rc;
        s.times.assign(times, times + nt);
        CU(cudaMemcpy2DAsync(s.a, ld * sizeof(NF), values, nc * sizeof(NF), nc * sizeof(NF), nt, cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        s.kind = kind; s.nt = nt;
        return TRM_OK;
    }
    int get_input(int id, void* host, int64_t count) override;
    int accumulate(int id, double w) override;
    int get_accumulated(int id, void* host, int64_t count, double scale, int reset) override;
    NF* acc[TRM_F_COUNT] = {nullptr};   // time-average accumulators [rows][ld], allocated on first use
    int input_ptr(int id, void** q) override {
        CU(cudaSetDevice(device));
        if (in[id].kind == TRM_SRC_TABLE || in[id].kind == TRM_SRC_RASTER) return fail(TRM_ERR_STATE, "input_ptr: input is a time series table");
        if (int rc = ensure(&in[id].a, ld)) return rc;
        if (in[id].kind == TRM_SRC_CONST) {   // materialise the constant so that the borrowed vector is meaningful
            std::vector<NF> v((size_t)nc, (NF)in[id].cval);
            CU(cudaMemcpyAsync(in[id].a, v.data(), nc * sizeof(NF), cudaMemcpyHostToDevice, stream));
            CU(cudaStreamSynchronize(stream));
        }
        if (in[id].kind != TRM_SRC_SINUSOID) in[id].kind = TRM_SRC_FIELD;
        *q = in[id].a;
        return TRM_OK;
    }

    // ------------------------------------------------------------------ launches
    void base_args(StageArgs<NF>& a) {
        std::memset(&a, 0, sizeof(a));
        a.ncol = nc; a.ld = ld; a.nz = nz; a.richards = richards ? 1 : 0;
        a.metrics = metrics; a.p = p;
        std::memcpy(a.cmet, cmet, sizeof(cmet));
        for (int s = 0; s < TRM_BC_NSLOTS; ++s) a.bc[s] = cfg.bc[s];
        for (int i = 0; i < TRM_IN_COUNT; ++i) {
            InputDesc<NF>& d = a.in[i]; const Input& s = in[i];
            d.kind = s.kind; d.nt = s.nt; d.cval = (NF)s.cval; d.period = s.period; d.lo = s.lo; d.hi = s.hi;
            d.a = s.a; d.b = s.b; d.c = s.c; d.ld = ld;
            if (s.kind == TRM_SRC_FIELD_PAIR) d.b = s.a;   // every evaluation at the step's start time; Heun stage 2: pair_stage2()
        }
        a.Kf = Kf;
        a.Ts = land2d[0]; a.G = land2d[1]; a.SWup = land2d[2]; a.LWup = land2d[3]; a.Rnet = land2d[4];
        a.Hs = land2d[5]; a.Hl = land2d[6]; a.Egnd = land2d[7]; a.infil = land2d[8]; a.runoff = land2d[9];
        a.veg = veg ? 1 : 0; a.vp = vp; a.paw = paw; a.xbeta = beta; a.ybeta = beta;
        for (int i = 0; i < VF_COUNT; ++i) a.veg2d[i] = veg2d[i];
        // ForwardEuler / auxiliary evaluations: evaluate on, and update, the model state in place
        for (int i = 0; i < 3; ++i) { a.vx[i] = veg2d[i]; a.vb[i] = veg2d[i]; a.vy[i] = veg2d[i]; a.vk1[i] = nullptr; a.vok1[i] = nullptr; }
    }
    // Position of clock time t on the time axis of a table input, with the reference's two update rules:
    // TABLE  = FieldTimeSeries[Time(t)] (input_sources.jl:165-171): flat up to the first / from the last node, else the
    //          bracket [n1, n2) with n2 = first node after t ;
    // RASTER = update_from_raster! (TerrariumRastersExt.jl:96-121): searchsorted ; on a node the node itself, beyond the
    //          axis the nearest end.
    static void bracket(const Input& s, double t, TimeBracket& k) {
        const std::vector<double>& tt = s.times;
        const int nt = (int)tt.size();
        k = TimeBracket{0, 0, 1, 0, 0.0, 1.0};
        if (nt == 0) return;
        if (s.kind == TRM_SRC_TABLE) {
            if (t <= tt.front()) { k.i1 = 0; return; }
            if (t >= tt.back()) { k.i1 = nt - 1; return; }
            const int n2 = (int)(std::upper_bound(tt.begin(), tt.end(), t) - tt.begin()), n1 = n2 - 1;
            k.i1 = n1; k.i2 = n2; k.flat = 0; k.e = t - tt[n1]; k.dtt = tt[n2] - tt[n1];
        } else {
            const int right = (int)(std::lower_bound(tt.begin(), tt.end(), t) - tt.begin()) + 1;   // first(searchsorted), 1-based
            const int left = (int)(std::upper_bound(tt.begin(), tt.end(), t) - tt.begin());        // last(searchsorted), 1-based
            if (left >= 1 && right <= nt && right > left) {
                k.i1 = left - 1; k.i2 = right - 1; k.flat = 0; k.e = t - tt[left - 1]; k.dtt = tt[right - 1] - tt[left - 1];
            } else if (left >= 1 && right <= nt) {
                k.i1 = right - 1;                       // exactly on a node (dt = 0 in the reference: the node value)
            } else {
                k.i1 = std::min(right, nt) - 1;          // beyond either end: flat
            }
        }
    }
    // call after t_x / t_b are set: table inputs get the brackets of both clock times
    void set_times(StageArgs<NF>& a) {
        for (int i = 0; i < TRM_IN_COUNT; ++i) {
            if (in[i].kind != TRM_SRC_TABLE && in[i].kind != TRM_SRC_RASTER) continue;
            bracket(in[i], (double)a.t_x, a.in[i].br[0]);
            bracket(in[i], (double)a.t_b, a.in[i].br[1]);
        }
    }
    // Heun stage 2: host-evaluated functions of time are read at t + dt where the tendencies are evaluated, at t for Flux BCs
    void pair_stage2(StageArgs<NF>& a) {
        for (int i = 0; i < TRM_IN_COUNT; ++i) if (in[i].kind == TRM_SRC_FIELD_PAIR) { a.in[i].a = in[i].b; a.in[i].b = in[i].a; }
    }
    void x_state(StageArgs<NF>& a) { a.xU = U; a.xS = S; a.xT = T; a.xL = Lq; a.xP = P; a.xWt = Wt; a.bU = U; a.bS = S; a.bSx = Sx; }
    void y_state(StageArgs<NF>& a) { a.yU = U; a.yS = S; a.yT = T; a.yL = Lq; a.yP = P; a.yWt = Wt; a.ySx = Sx; }
    int launch(int variant, const StageArgs<NF>& a);

    int check_bcs() {
        for (int s = 0; s < TRM_BC_NSLOTS; ++s) {
            int k = cfg.bc[s].kind;
            if (k < TRM_BC_DEFAULT || k > TRM_BC_FLUX) return fail(TRM_ERR_INVALID, "bad boundary condition kind");
            if (k != TRM_BC_DEFAULT && (cfg.bc[s].input < 0 || cfg.bc[s].input >= TRM_IN_COUNT)) return fail(TRM_ERR_INVALID, "bad boundary condition input id");
        }
        return TRM_OK;
    }

    int initialize() override;
    int step(double dt, int64_t n) override;
    int aux() override;
    int tendencies() override;
    int diagnostics(trm_diag* out, double** dev) override;
    int diagnostics_allreduce(trm_diag* out) override;
    double* diag_pack = nullptr;
    int set_block(int b) override {
        if (b < 32 || b > TRM_MAX_BLOCK || b % 32) return fail(TRM_ERR_INVALID, "block must be a multiple of 32 in [32, 128]");
        block = b;
        return TRM_OK;
    }
    int launch_euler(const StageArgs<NF>& a, int load_aux);
    int launch_surface(int what, const StageArgs<NF>& a, cudaStream_t on = nullptr);   // (default: the compute stream)
    // the staged kernels (euler_kernel.cuh) leave the LandModel surface block to surface_kernel; the generic streaming
    // kernel evaluates it inline
    // Heun recompute protocol (stage_kernel.cuh: heun_recompute): the Float64 staged kernels, both stages of a step alike
    bool heun_rc() const { return heun_recompute<NF>() && euler_impl == 1 && (uint64_t)nz * (uint64_t)ld < (1ull << 32); }
    bool split_surface() const { return land && euler_impl == 1 && (uint64_t)nz * (uint64_t)ld < (1ull << 32); }
    int enqueue_steps(double dt, int64_t n);
    // Small domains (warp_kernel.cuh): one warp per column, several steps per launch (LandModel: surface_kernel + one launch
    // per step, both Heun stages in it). Applies with nz <= 31 while the column count leaves the one-thread-per-column
    // kernels latency bound (less than about one wave of warps). Measured crossover with the streaming kernels
    // (profiles/r02_warp_crossover.txt): ~115 k columns in Float32 (Heun: ~200 k), ~80 k in Float64 (lanes = layers of one
    // column diverge where adjacent columns of one layer do not, and the Float64 fast-math sequences are longer): default
    // limit 114688 / 65536 columns. A launch that advances ONE step only -- LandModel (the surface block runs in between),
    // inputs whose descriptor changes per step, a bound host exchange, a caller that asks for one step per call -- has much
    // lower limits (see use_warp). TRM_WARP_COLS overrides the limit, TRM_WARP=0 switches the kernel off (tests compare
    // both within a process).
    bool steps_one_by_one() const {
        if (land || hio.nslots != 0) return true;
        for (int i = 0; i < TRM_IN_COUNT; ++i)
            if (in[i].kind == TRM_SRC_TABLE || in[i].kind == TRM_SRC_RASTER || in[i].kind == TRM_SRC_FIELD_PAIR) return true;
        return false;
    }
    bool use_warp(int64_t n) const {
        if (euler_impl != 1 || nz > 31) return false;
        const char* e = std::getenv("TRM_WARP");
        if (e && e[0] == '0') return false;
        const char* m = std::getenv("TRM_WARP_COLS");
        const bool single = n <= 1 || steps_one_by_one(), f32 = sizeof(NF) == 4;
        int64_t max_cols = f32 ? 114688 : 65536;
        if (single) {
            // measured (profiles/r02_warp_crossover.txt, r02_small_domains.csv): a one-step launch pays its fixed latencies
            // (metrics, tile, boundary inputs, closure) in every step -- ~23 us on 8192 columns against 2.6 us per step of a
            // 600-step launch -- and only wins where it replaces several streaming launches: Heun, and the LandModel under Heun
            if (heun) max_cols = land ? (f32 ? 32768 : 16384) : (f32 ? 16384 : 8192);
            else max_cols = f32 ? 8192 : 4096;
        }
        if (m) max_cols = std::atoll(m);
        return nc <= max_cols;
    }
    int enqueue_steps_warp(NF dt, int64_t n);
    cudaError_t call_warp(int nsteps, const StageArgs<NF>& a);
    int set_input_field_async(int id, const void* v) override;
    int get_field_async(int id, void* host, int64_t count) override;
    int step_async(double dt, int64_t n) override;
    int sync_all() override;
    int set_ring_index(const int64_t* idx, int64_t nring_) override;
    int get_field_ring(int id, void* host, int64_t count, double fill) override;
    int set_field_ring(int id, const void* host, int64_t count) override;
    int ring_buffer(size_t need);
    int bind_host_io(int in_id, const void* host_in, int field_id, void* host_out, int nslots) override;
    int host_io_wait(int64_t it) override;
    int reset_state() override;
    void set_clock(double t, int64_t it) override { time = (double)(NF)t; t_inputs = time; iteration = it; }
    // host exchange of the step that starts at `iteration`: input pointer of this step's slot, prefetch flag
    void apply_host_io(StageArgs<NF>& a, bool writes_final_state) {
        if (hio.nslots) {
            const int64_t slot = iteration % hio.nslots;
            if (hio.in_id >= 0) { a.in[hio.in_id].kind = TRM_SRC_FIELD; a.in[hio.in_id].a = hio.in_dev + slot * nc; }
            if (writes_final_state && hio.out_id == TRM_F_GROUND_TEMPERATURE) a.hio_out = hio.out_dev + slot * nc;
        }
        const trm_bc& b = a.bc[TRM_BC_TEMPERATURE_TOP];
        a.bct_pre = (b.kind == TRM_BC_VALUE || b.kind == TRM_BC_GRADIENT) && a.in[b.input].kind == TRM_SRC_FIELD && a.in[b.input].a ? 1 : 0;
    }
};

template <> int Handle<float>::launch(int variant, const StageArgs<float>& a) {
    cudaError_t e = ks->stage_f32(phys, variant, a, block, stream); ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("stage kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}
template <> int Handle<double>::launch(int variant, const StageArgs<double>& a) {
    cudaError_t e = ks->stage_f64(phys, variant, a, block, stream); ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("stage kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}

template <> int Handle<float>::launch_surface(int what, const StageArgs<float>& a, cudaStream_t on) {
    cudaError_t e = ks->surface_f32(what, a, on ? on : stream); ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("surface kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}
template <> int Handle<double>::launch_surface(int what, const StageArgs<double>& a, cudaStream_t on) {
    cudaError_t e = ks->surface_f64(what, a, on ? on : stream); ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("surface kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}

// stage launch on the shared-memory staged kernel; falls back to the generic streaming kernel when that one does
// not apply (fields of 2^32 elements or more) or when TRM_KERNEL=stream asks for it
template <> int Handle<float>::launch_euler(const StageArgs<float>& a, int load_aux) {
    const int generic = a.mode == MODE_EULER ? (load_aux ? VAR_EULER_LOAD : VAR_EULER_RECOMPUTE) : VAR_GENERIC;
    if (euler_impl != 1) return launch(generic, a);
    cudaError_t e = ks->euler_f32(phys, a.mode, load_aux, a, stream);
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); log_fallback(); return launch(generic, a); }
    ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("euler kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}
template <> int Handle<double>::launch_euler(const StageArgs<double>& a, int load_aux) {
    const int generic = a.mode == MODE_EULER ? (load_aux ? VAR_EULER_LOAD : VAR_EULER_RECOMPUTE) : VAR_GENERIC;
    if (euler_impl != 1) return launch(generic, a);
    cudaError_t e = ks->euler_f64(phys, a.mode, load_aux, a, stream);
    if (e == cudaErrorInvalidConfiguration) { cudaGetLastError(); log_fallback(); return launch(generic, a); }
    ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("euler kernel launch: ") + cudaGetErrorString(e));
    return TRM_OK;
}

template <class NF> cudaError_t call_init(const KernelSet* ks, Handle<NF>* h);
template <> cudaError_t call_init<float>(const KernelSet* ks, Handle<float>* h) {
    return ks->init_f32(h->nc, h->ld, h->nz, h->richards, h->metrics, h->p, h->U, h->S, h->T, h->Lq, h->P, h->Wt, h->Sx, h->stream);
}
template <> cudaError_t call_init<double>(const KernelSet* ks, Handle<double>* h) {
    return ks->init_f64(h->nc, h->ld, h->nz, h->richards, h->metrics, h->p, h->U, h->S, h->T, h->Lq, h->P, h->Wt, h->Sx, h->stream);
}

// initialize!(integrator) tail (model_integrator.jl:96-109 -> soil_model.jl:31-37 / land_model.jl:68-77)
template <class NF> int Handle<NF>::initialize() {
    CU(cudaSetDevice(device));
    if (int rc = check_bcs()) return rc;
    time = 0.0; iteration = 0; t_inputs = 0.0;   // reset!(clock); update_inputs! at t0
    cudaError_t e = call_init<NF>(ks, this); ++launches;
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("init kernel launch: ") + cudaGetErrorString(e));
    // compute_hydraulics! of initialize! (soil_hydrology.jl:113-117, soil_hydrology_rre.jl:33-47) runs before
    // the liquid fraction exists (the field is still zero there); it is recomputed by the first
    // update_state!, so the hydraulic conductivity field is materialised by trm_compute_auxiliary only.
    CU(cudaStreamSynchronize(stream));
    initialized = true; aux_stale = true; beta_stale = true;
    return TRM_OK;
}

template <> cudaError_t Handle<float>::call_warp(int nsteps, const StageArgs<float>& a) { return ks->warp_f32(phys, heun ? 1 : 0, nsteps, a, stream); }
template <> cudaError_t Handle<double>::call_warp(int nsteps, const StageArgs<double>& a) { return ks->warp_f64(phys, heun ? 1 : 0, nsteps, a, stream); }

// timestep! x n on a small domain: the warp-per-column kernel advances every column `chunk` steps per launch, both Heun stages
// included. Inputs whose descriptor changes from step to step on the host side (tables / rasters: time bracket ; host
// evaluated functions: value pair ; a mapped host exchange: ring slot) limit a launch to one step.
template <class NF> int Handle<NF>::enqueue_steps_warp(NF dt, int64_t n) {
    const bool per_step = steps_one_by_one();   // (LandModel: the surface block is a launch of its own before every step)
    int64_t done = 0;
    while (done < n) {
        const int64_t chunk = per_step ? 1 : std::min<int64_t>(n - done, 1 << 30);
        const NF t = (NF)time;
        const NF t1 = t + dt;
        if (land) {
            // surface block on the time-n state (land_model.jl:79-88): leaves the ground heat flux / infiltration that the warp
            // kernel applies as top Flux BCs; same arguments as the stage-1 / ForwardEuler surface launch of enqueue_steps
            if (veg && beta_stale) {
                StageArgs<NF> g; base_args(g); x_state(g);
                if (int rc = launch_surface(1, g)) return rc;
            }
            StageArgs<NF> a; base_args(a);
            a.dt = dt; a.mode = heun ? MODE_HEUN1 : MODE_EULER; a.t_x = t; a.t_b = t; x_state(a); y_state(a);
            a.load_aux = (aux_stale || force_load) ? 1 : 0;
            if (heun) {
                a.yU = gU; a.yS = gS; a.yWt = gWt; a.ySx = nullptr; a.oTU = tU; a.oTS = tS;
                for (int i = 0; i < 3; ++i) { a.vy[i] = gveg[i]; a.vok1[i] = tveg[i]; }
                a.ybeta = gbeta;
            }
            set_times(a);
            apply_host_io(a, !heun);
            if (int rc = launch_surface(0, a)) return rc;
        }
        StageArgs<NF> a; base_args(a);
        a.dt = dt; a.mode = heun ? MODE_HEUN1 : MODE_EULER; a.t_x = t; a.t_b = t; x_state(a); y_state(a);
        a.load_aux = (aux_stale || force_load) ? 1 : 0;
        a.stU = gU; a.stS = gS; a.sbeta = gbeta;   // (vegetated LandModel, Heun) top layer and factor of the stage state
        // time index of the descriptors inside this kernel: 0 = start of the step, 1 = start + dt (Heun stage 2)
        for (int i = 0; i < TRM_IN_COUNT; ++i) {
            if (in[i].kind == TRM_SRC_TABLE || in[i].kind == TRM_SRC_RASTER) {
                bracket(in[i], (double)t, a.in[i].br[0]);
                bracket(in[i], (double)t1, a.in[i].br[1]);
            }
            if (in[i].kind == TRM_SRC_FIELD_PAIR) { a.in[i].a = in[i].a; a.in[i].b = in[i].b; }
        }
        apply_host_io(a, true);
        cudaError_t e = call_warp((int)chunk, a); ++launches;
        if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("warp kernel launch: ") + cudaGetErrorString(e));
        if (land && veg && heun) {
            // Heun stage 2 of the vegetated model: the vegetation block on the stage state at t + dt, for k2 of canopy water,
            // vegetation carbon and area fraction only (heun.jl:45-58) -- it feeds nothing back into the soil step, so it runs
            // after the kernel that formed the stage state; same arguments as the stage-2 surface launch of enqueue_steps
            StageArgs<NF> b; base_args(b);
            b.dt = dt; b.mode = MODE_HEUN2; b.load_aux = 0; b.t_x = t1; b.t_b = t;
            b.xU = gU; b.xS = richards ? gS : S; b.xWt = gWt; b.bU = U; b.bS = S; b.bSx = Sx; b.k1U = tU; b.k1S = tS;
            for (int i = 0; i < 3; ++i) { b.vx[i] = gveg[i]; b.vk1[i] = tveg[i]; }
            b.xbeta = gbeta;
            y_state(b);
            set_times(b);
            pair_stage2(b);
            apply_host_io(b, true);
            if (int rc = launch_surface(0, b)) return rc;
        }
        if (hio.nslots) {
            const int64_t slot = iteration % hio.nslots;
            if (hio.out_id >= 0 && hio.out_id != TRM_F_GROUND_TEMPERATURE) {
                copy_row_kernel<NF><<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>(nc, field(hio.out_id).ptr, hio.out_dev + slot * nc);
                ++launches;
            }
            CU(cudaEventRecord(hio.ev[slot], stream));
        }
        aux_stale = false; beta_stale = !(land && veg);   // (the LandModel variant leaves the factor of the new state behind)
        for (int64_t i = 0; i < chunk; ++i) {   // tick!(clock, dt) in the clock's number format, like the kernel
            const NF ts = (NF)time;
            t_inputs = (double)ts;
            time = (double)(NF)(ts + dt); iteration += 1;
        }
        done += chunk;
    }
    return TRM_OK;
}

template <class NF> int Handle<NF>::enqueue_steps(double dt_, int64_t n) {
    if (!initialized) return fail(TRM_ERR_STATE, "trm_step before trm_initialize");
    if (n < 0) return fail(TRM_ERR_INVALID, "nsteps < 0");
    CU(cudaSetDevice(device));
    const NF dt = (NF)dt_;
    CU(cudaEventRecord(ev0, stream));
    if (use_warp(n)) {
        if (int rc = enqueue_steps_warp(dt, n)) return rc;
        CU(cudaEventRecord(ev1, stream));
        timing_open = true;
        return TRM_OK;
    }
    for (int64_t i = 0; i < n; ++i) {
        StageArgs<NF> a; base_args(a);
        const NF t = (NF)time;
        const NF t1 = t + dt;   // tick!(clock, dt) in the clock's number format
        a.dt = dt;
        const bool split = split_surface();
        if (split && veg && beta_stale) {   // soil moisture limiting factor of the stored state (first step / after a user write)
            StageArgs<NF> g; base_args(g); x_state(g);
            if (int rc = launch_surface(1, g)) return rc;
        }
        if (!heun) {   // forward_euler.jl:19-31
            a.mode = MODE_EULER; a.t_x = t; a.t_b = t; x_state(a); y_state(a);
            const bool load = aux_stale || force_load;
            a.load_aux = load ? 1 : 0;
            set_times(a);
            apply_host_io(a, true);
            if (split) { if (int rc = launch_surface(0, a)) return rc; }
            if (int rc = launch_euler(a, load ? 1 : 0)) return rc;
        } else {       // heun.jl:37-71
            a.mode = MODE_HEUN1; a.load_aux = aux_stale ? 1 : 0; a.t_x = t; a.t_b = t; x_state(a);
            a.yU = gU; a.yS = gS; a.yWt = gWt; a.ySx = nullptr; a.oTU = tU; a.oTS = tS;
            for (int i = 0; i < 3; ++i) { a.vy[i] = gveg[i]; a.vok1[i] = tveg[i]; }
            a.ybeta = gbeta;
            a.hflag_out = hflag;
            set_times(a);
            apply_host_io(a, false);
            if (split) { if (int rc = launch_surface(0, a)) return rc; }
            if (int rc = launch_euler(a, a.load_aux)) return rc;
            StageArgs<NF> b; base_args(b);
            b.dt = dt; b.mode = MODE_HEUN2; b.load_aux = 0; b.t_x = t1; b.t_b = t;
            b.xU = gU; b.xS = richards ? gS : S; b.xWt = gWt; b.bU = U; b.bS = S; b.bSx = Sx; b.k1U = tU; b.k1S = tS;
            const bool rc = heun_rc();
            // recompute protocol: the stage kernel streams the base state and k1 and rebuilds the stage state; only flagged
            // columns (and the top layer, for the surface block) read what stage 1 stored
            if (rc) { b.xU = U; b.xS = S; b.sU = gU; b.sS = gS; b.hflag_in = hflag; }
            for (int i = 0; i < 3; ++i) { b.vx[i] = gveg[i]; b.vk1[i] = tveg[i]; }
            b.xbeta = gbeta;
            y_state(b);
            set_times(b);
            pair_stage2(b);
            apply_host_io(b, true);
            // stage 2 only re-evaluates the vegetation block (k2 of canopy water, vegetation carbon and area fraction);
            // the bare-ground surface block of the stage state has no effect on the step (heun.jl:63-66)
            if (split && veg) {
                StageArgs<NF> bs = b;   // the surface block is evaluated on the stage state (its top layer is always stored)
                bs.xU = gU; bs.xS = richards ? gS : S;
                // it reads what stage 1 left (top layer and factor of the stage state, stage values and k1 of the vegetation
                // prognostics) and writes the new vegetation prognostics only; the stage kernel reads and writes none of
                // them: fork after stage 1, join before the next launch on the compute stream
                if (!s_aux) {
                    CU(cudaStreamCreateWithFlags(&s_aux, cudaStreamNonBlocking));
                    CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
                    CU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
                }
                CU(cudaEventRecord(ev_fork, stream));
                CU(cudaStreamWaitEvent(s_aux, ev_fork, 0));
                if (int rc2 = launch_surface(0, bs, s_aux)) return rc2;
                CU(cudaEventRecord(ev_join, s_aux));
            }
            if (int rc = launch_euler(b, 0)) return rc;
            if (split && veg) CU(cudaStreamWaitEvent(stream, ev_join, 0));
        }
        if (hio.nslots) {
            const int64_t slot = iteration % hio.nslots;
            if (hio.out_id >= 0 && hio.out_id != TRM_F_GROUND_TEMPERATURE) {   // any other 2-D field: one small copy kernel
                copy_row_kernel<NF><<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>(nc, field(hio.out_id).ptr, hio.out_dev + slot * nc);
                ++launches;
            }
            CU(cudaEventRecord(hio.ev[slot], stream));
        }
        beta_stale = !split;   // the staged kernels leave the factor of the new state behind; the generic kernel does not
        aux_stale = false;
        t_inputs = (double)t;   // the state's inputs were last updated at the start of this step
        time = (double)t1; iteration += 1;
    }
    CU(cudaEventRecord(ev1, stream));
    timing_open = true;
    return TRM_OK;
}

template <class NF> int Handle<NF>::step(double dt_, int64_t n) {
    nvtxRangePushA("trm_step");   // visible in nsys / ncu timelines around the stage launches of this call
    int rc = enqueue_steps(dt_, n);
    nvtxRangePop();
    if (rc) return rc;
    CU(cudaEventSynchronize(ev1));
    CU(cudaEventElapsedTime(&last_ms, ev0, ev1));
    timing_open = false;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TRM_ERR_CUDA, std::string("trm_step: ") + cudaGetErrorString(e));
    if (debug_nancheck) {
        // TERRARIUM_DEBUG=true (src/diagnostics/debugging.jl:1-47): the reference checks every kernel output for NaN;
        // here the state is checked once per trm_step call with the diagnostics reduction
        trm_diag d;
        if (int rc2 = diagnostics(&d, nullptr)) return rc2;
        if (d.nan_count > 0) return fail(TRM_ERR_STATE, "TERRARIUM_DEBUG: found " + std::to_string((long long)d.nan_count) + " non-finite values in internal_energy / temperature / saturation_water_ice after trm_step");
    }
    return TRM_OK;
}

template <class NF> int Handle<NF>::step_async(double dt_, int64_t n) { return enqueue_steps(dt_, n); }

// ---- per-step exchange with a host-side coupler through mapped, page-locked host memory ----
template <class NF> int Handle<NF>::bind_host_io(int in_id, const void* host_in, int field_id, void* host_out, int nslots) {
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize(stream));
    for (cudaEvent_t e : hio.ev) if (e) cudaEventDestroy(e);
    hio = HostIO{};
    if (!host_in && !host_out) return TRM_OK;   // unbind
    if (nslots < 1 || nslots > 64) return fail(TRM_ERR_INVALID, "bind_host_io: nslots must be in [1, 64]");
    if (host_in) {
        if (in_id < 0 || in_id >= TRM_IN_COUNT) return fail(TRM_ERR_INVALID, "bind_host_io: bad input id");
        if (in[in_id].kind == TRM_SRC_TABLE || in[in_id].kind == TRM_SRC_RASTER) return fail(TRM_ERR_STATE, "bind_host_io: input is a time series table");
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, const_cast<void*>(host_in), 0) != cudaSuccess) {
            cudaGetLastError();
            return fail(TRM_ERR_INVALID, "bind_host_io: host_in is not page-locked, mapped host memory (allocate it with trm_host_alloc / trm_host_alloc_ex)");
        }
        hio.in_id = in_id; hio.in_dev = (const NF*)d;
    }
    if (host_out) {
        FieldRef f = field(field_id);
        if (!f.ptr || f.nrows != 1) return fail(TRM_ERR_INVALID, "bind_host_io: the output must be a 2-D field of this model (e.g. ground_temperature, skin_temperature)");
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, host_out, 0) != cudaSuccess) {
            cudaGetLastError();
            return fail(TRM_ERR_INVALID, "bind_host_io: host_out is not page-locked, mapped host memory (allocate it with trm_host_alloc / trm_host_alloc_ex)");
        }
        hio.out_id = field_id; hio.out_dev = (NF*)d;
    }
    hio.nslots = nslots; hio.first_iter = iteration;
    hio.ev.assign((size_t)nslots, nullptr);
    for (int i = 0; i < nslots; ++i) CU(cudaEventCreateWithFlags(&hio.ev[i], cudaEventDisableTiming));
    return TRM_OK;
}
template <class NF> int Handle<NF>::host_io_wait(int64_t it) {
    if (!hio.nslots) return fail(TRM_ERR_STATE, "host_io_wait without trm_bind_host_io");
    if (it > iteration) return fail(TRM_ERR_INVALID, "host_io_wait: that step has not been enqueued yet");
    if (it <= hio.first_iter) return TRM_OK;   // produced before the binding
    CU(cudaSetDevice(device));
    // the step (it-1) -> it recorded ev[(it-1) % nslots]; if a later step has reused the slot, waiting for that one is
    // still correct (steps complete in order)
    CU(cudaEventSynchronize(hio.ev[(size_t)((it - 1) % hio.nslots)]));
    return TRM_OK;
}

// reset!(integrator.state) of initialize!(integrator) (model_integrator.jl:98): every field of the state goes back to
// zero before the initializers run (metrics and input sources are not state)
template <class NF> int Handle<NF>::reset_state() {
    CU(cudaSetDevice(device));
    const size_t n3 = (size_t)nz * ld;
    for (NF* f : {U, T, Lq, S, P, tU, tS, gU, gS, paw}) if (f) CU(cudaMemsetAsync(f, 0, n3 * sizeof(NF), stream));
    if (Kf) CU(cudaMemsetAsync(Kf, 0, (size_t)(nz + 1) * ld * sizeof(NF), stream));
    for (NF* f : {Sx, Wt, gWt, beta, gbeta}) if (f) CU(cudaMemsetAsync(f, 0, (size_t)ld * sizeof(NF), stream));
    for (NF* f : land2d) if (f) CU(cudaMemsetAsync(f, 0, (size_t)ld * sizeof(NF), stream));
    for (NF* f : veg2d) if (f) CU(cudaMemsetAsync(f, 0, (size_t)ld * sizeof(NF), stream));
    for (int i = 0; i < 3; ++i) for (NF* f : {gveg[i], tveg[i]}) if (f) CU(cudaMemsetAsync(f, 0, (size_t)ld * sizeof(NF), stream));
    for (int id = 0; id < TRM_F_COUNT; ++id) if (acc[id]) CU(cudaMemsetAsync(acc[id], 0, (size_t)field(id).nrows * ld * sizeof(NF), stream));
    CU(cudaStreamSynchronize(stream));
    time = 0.0; iteration = 0; t_inputs = 0.0;
    initialized = false; aux_stale = true; beta_stale = true;
    return TRM_OK;
}

template <class NF> int Handle<NF>::sync_all() {
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize(stream));
    if (s_in) CU(cudaStreamSynchronize(s_in));
    if (s_out) CU(cudaStreamSynchronize(s_out));
    if (timing_open) { CU(cudaEventElapsedTime(&last_ms, ev0, ev1)); timing_open = false; }
    return TRM_OK;
}

// ---- ColumnRingGrid conversions (src/grids/column_ring_grid.jl:102-149) on the device ----
template <class NF> int Handle<NF>::set_ring_index(const int64_t* idx, int64_t nring_) {
    if (nring_ < nc) return fail(TRM_ERR_INVALID, "set_ring_index: ring grid smaller than the column count");
    for (int64_t i = 0; i < nc; ++i) if (idx[i] < 0 || idx[i] >= nring_) return fail(TRM_ERR_INVALID, "set_ring_index: index outside the ring grid");
    CU(cudaSetDevice(device));
    if (!ring_index) { if (int rc = dalloc(&ring_index, (size_t)ld)) return rc; }
    CU(cudaMemcpyAsync(ring_index, idx, nc * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));
    nring = nring_;
    return TRM_OK;
}
template <class NF> int Handle<NF>::ring_buffer(size_t need) {
    if (need <= ring_count) return TRM_OK;
    if (ring_buf) dfree(ring_buf);
    ring_buf = nullptr; ring_count = 0;
    if (int rc = dalloc(&ring_buf, need, false)) return rc;
    ring_count = need;
    return TRM_OK;
}
template <class NF> int Handle<NF>::get_field_ring(int id, void* host, int64_t count, double fill) {
    if (!ring_index) return fail(TRM_ERR_STATE, "get_field_ring before trm_set_ring_index");
    FieldRef f = field(id);
    if (!f.ptr) return fail(TRM_ERR_INVALID, "get_field_ring: unknown field or field not defined for this model");
    if (count != (int64_t)f.nrows * nring) return fail(TRM_ERR_INVALID, "get_field_ring: count != nrows * nring");
    CU(cudaSetDevice(device));
    if (int rc = ring_buffer((size_t)count)) return rc;
    ring_fill_kernel<NF><<<(unsigned)std::min<int64_t>((count + 255) / 256, 148 * 16), 256, 0, stream>>>(count, ring_buf, (NF)fill);
    ring_scatter_kernel<NF><<<(unsigned)((nc + 127) / 128), 128, 0, stream>>>(nc, ld, f.nrows, nring, ring_index, f.ptr, ring_buf);
    launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, ring_buf, (size_t)count * sizeof(NF), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return TRM_OK;
}
template <class NF> int Handle<NF>::set_field_ring(int id, const void* host, int64_t count) {
    if (!ring_index) return fail(TRM_ERR_STATE, "set_field_ring before trm_set_ring_index");
    FieldRef f = field(id);
    if (!f.ptr || !f.writable) return fail(TRM_ERR_INVALID, "set_field_ring: unknown or read-only field");
    if (count != (int64_t)f.nrows * nring) return fail(TRM_ERR_INVALID, "set_field_ring: count != nrows * nring");
    CU(cudaSetDevice(device));
    if (int rc = ring_buffer((size_t)count)) return rc;
    CU(cudaMemcpyAsync(ring_buf, host, (size_t)count * sizeof(NF), cudaMemcpyHostToDevice, stream));
    ring_gather_kernel<NF><<<(unsigned)((nc + 127) / 128), 128, 0, stream>>>(nc, ld, f.nrows, nring, ring_index, ring_buf, f.ptr);
    ++launches;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(stream));
    aux_stale = true; beta_stale = true;
    return TRM_OK;
}

// Upload of a per-column input overlapped with the steps already enqueued: the copy goes to the back buffer on
// the copy-in stream; steps enqueued after this call wait for it and read the new buffer.
template <class NF> int Handle<NF>::set_input_field_async(int id, const void* v) {
    CU(cudaSetDevice(device));
    Input& s = in[id];
    if (s.kind == TRM_SRC_TABLE || s.kind == TRM_SRC_RASTER) return fail(TRM_ERR_STATE, "set_input_field_async: input is a time series table");
    if (!s_in) CU(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    if (int rc = ensure(&s.a, ld)) return rc;
    if (int rc = ensure(&s.a2, ld)) return rc;
    if (!s.ev_ready) {
        CU(cudaEventCreateWithFlags(&s.ev_ready, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&s.ev_free[i], cudaEventDisableTiming));
        CU(cudaStreamSynchronize(stream));   // the ensure() memsets above
    }
    // every step that reads the current front buffer has been enqueued: mark the point where it becomes free
    CU(cudaEventRecord(s.ev_free[s.front], stream));
    const int back = 1 - s.front;
    CU(cudaStreamWaitEvent(s_in, s.ev_free[back], 0));   // (never recorded yet == already complete)
    NF* dst = s.a2;
    CU(cudaMemcpyAsync(dst, v, nc * sizeof(NF), cudaMemcpyHostToDevice, s_in));
    CU(cudaEventRecord(s.ev_ready, s_in));
    CU(cudaStreamWaitEvent(stream, s.ev_ready, 0));
    std::swap(s.a, s.a2);   // `a` is what base_args hands to the kernels
    s.front = back;
    s.kind = TRM_SRC_FIELD;
    return TRM_OK;
}

// Download of a field overlapped with later steps: snapshot on the compute stream (device to device), then
// device to host on the copy-out stream.
template <class NF> int Handle<NF>::get_field_async(int id, void* host, int64_t count) {
    FieldRef f = field(id);
    if (!f.ptr) return fail(TRM_ERR_INVALID, "get_field_async: unknown field or field not defined for this model");
    if (count != (int64_t)f.nrows * nc) return fail(TRM_ERR_INVALID, "get_field_async: wrong element count");
    CU(cudaSetDevice(device));
    if (!s_out) {
        CU(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ev_staged, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_out_done, cudaEventDisableTiming));
    }
    const size_t need = (size_t)f.nrows * ld;
    if (need > staging_count) {
        CU(cudaStreamSynchronize(s_out));
        if (staging) dfree(staging);
        staging = nullptr;
        if (int rc = dalloc(&staging, need, false)) return rc;
        staging_count = need;
    }
    CU(cudaStreamWaitEvent(stream, ev_out_done, 0));   // the previous download is done with the staging buffer
    CU(cudaMemcpyAsync(staging, f.ptr, (f.nrows == 1 ? (size_t)nc : need) * sizeof(NF), cudaMemcpyDeviceToDevice, stream));
    CU(cudaEventRecord(ev_staged, stream));
    CU(cudaStreamWaitEvent(s_out, ev_staged, 0));
    if (f.nrows == 1) CU(cudaMemcpyAsync(host, staging, nc * sizeof(NF), cudaMemcpyDeviceToHost, s_out));
    else CU(cudaMemcpy2DAsync(host, nc * sizeof(NF), staging, ld * sizeof(NF), nc * sizeof(NF), f.nrows, cudaMemcpyDeviceToHost, s_out));
    CU(cudaEventRecord(ev_out_done, s_out));
    return TRM_OK;
}

// compute_auxiliary!(state, model): soil_coupled.jl:62-72 / land_model.jl:79-88
// input values as of the last update_inputs! (evaluated on the device with the same code the stage kernels use)
template <class NF>
__global__ void eval_input_kernel(int64_t ncol, InputDesc<NF> d, NF t, NF* out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncol) out[c] = eval_input_inline(d, c, t);
}
template <class NF> int Handle<NF>::get_input(int id, void* host, int64_t count) {
    if (count != nc) return fail(TRM_ERR_INVALID, "get_input: count != ncol");
    CU(cudaSetDevice(device));
    StageArgs<NF> a; base_args(a);
    if (int rc = ring_buffer((size_t)nc)) return rc;
    a.t_x = a.t_b = (NF)t_inputs;
    set_times(a);
    eval_input_kernel<NF><<<(unsigned)((nc + 127) / 128), 128, 0, stream>>>(nc, a.in[id], (NF)t_inputs, ring_buf);
    ++launches;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, ring_buf, nc * sizeof(NF), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return TRM_OK;
}

// time-averaged output: acc += w * field (one element per thread over the padded rows), scaled read-back
template <class NF>
__global__ void axpy_kernel(int64_t n, NF w, const NF* __restrict__ x, NF* __restrict__ y, int overwrite) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = overwrite ? w * x[i] : y[i] + w * x[i];
}
template <class NF> int Handle<NF>::accumulate(int id, double w) {
    FieldRef f = field(id);
    if (!f.ptr) return fail(TRM_ERR_INVALID, "accumulate: unknown field or field not defined for this model");
    CU(cudaSetDevice(device));
    const int64_t n = (int64_t)f.nrows * ld;
    if (!acc[id]) { if (int rc = dalloc(&acc[id], (size_t)n)) return rc; }
    axpy_kernel<NF><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, (NF)w, f.ptr, acc[id], 0);
    ++launches;
    CU(cudaGetLastError());
    return TRM_OK;
}
template <class NF> int Handle<NF>::get_accumulated(int id, void* host, int64_t count, double scale, int reset) {
    FieldRef f = field(id);
    if (!f.ptr || !acc[id]) return fail(TRM_ERR_INVALID, "get_accumulated: nothing accumulated for this field");
    if (count != (int64_t)f.nrows * nc) return fail(TRM_ERR_INVALID, "get_accumulated: wrong element count");
    CU(cudaSetDevice(device));
    const int64_t n = (int64_t)f.nrows * ld;
    if (int rc = ring_buffer((size_t)n)) return rc;
    axpy_kernel<NF><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, (NF)scale, acc[id], ring_buf, 1);
    ++launches;
    CU(cudaGetLastError());
    CU(cudaMemcpy2DAsync(host, nc * sizeof(NF), ring_buf, ld * sizeof(NF), nc * sizeof(NF), f.nrows, cudaMemcpyDeviceToHost, stream));
    if (reset) CU(cudaMemsetAsync(acc[id], 0, (size_t)n * sizeof(NF), stream));
    CU(cudaStreamSynchronize(stream));
    return TRM_OK;
}

template <class NF> int Handle<NF>::aux() {
    if (!initialized) return fail(TRM_ERR_STATE, "trm_compute_auxiliary before trm_initialize");
    CU(cudaSetDevice(device));
    StageArgs<NF> a; base_args(a);
    // compute_auxiliary! does not call update_inputs!: the input fields still hold the values of the last
    // update_state! (start of the last step), which is what timestep!(...; finalize = true) / run! see.
    a.mode = MODE_AUX; a.load_aux = 1; a.t_x = (NF)t_inputs; a.t_b = a.t_x; a.dt = 0; x_state(a);
    set_times(a);
    if (int rc = launch(VAR_GENERIC, a)) return rc;
    CU(cudaStreamSynchronize(stream));
    return TRM_OK;
}

// update_state!(...) with the tendencies (incl. Flux BCs) materialised
template <class NF> int Handle<NF>::tendencies() {
    if (!initialized) return fail(TRM_ERR_STATE, "trm_compute_tendencies before trm_initialize");
    CU(cudaSetDevice(device));
    const size_t n3 = (size_t)nz * ld;
    if (!tU) { if (int rc = dalloc(&tU, n3)) return rc; }
    if (richards && !tS) { if (int rc = dalloc(&tS, n3)) return rc; }
    StageArgs<NF> a; base_args(a);
    a.mode = MODE_TEND; a.load_aux = 1; a.t_x = (NF)time; a.t_b = a.t_x; a.dt = 0; x_state(a);
    set_times(a);
    t_inputs = time;
    a.oTU = tU; a.oTS = tS;
    if (int rc = launch(VAR_GENERIC, a)) return rc;
    CU(cudaStreamSynchronize(stream));
    return TRM_OK;
}

template <class NF> int Handle<NF>::diagnostics(trm_diag* out, double** dev) {
    CU(cudaSetDevice(device));
    diag_kernel<NF><<<diag_blocks, 256, 0, stream>>>(nc, ld, nz, metrics, p.por, U, T, S, richards ? Sx : nullptr, diag_partial);
    finish_diag<<<1, 256, 0, stream>>>(diag_blocks, diag_partial, (double)nc, diag_out);
    launches += 2;
    CU(cudaGetLastError());
    if (dev) *dev = diag_out;
    if (out) {
        double v[8];
        CU(cudaMemcpyAsync(v, diag_out, sizeof(v), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        out->energy = v[0]; out->water = v[1]; out->t_min = v[2]; out->t_max = v[3];
        out->sat_min = v[4]; out->sat_max = v[5]; out->nan_count = v[6]; out->ncol = v[7];
    }
    return TRM_OK;
}

template <class NF> int Handle<NF>::diagnostics_allreduce(trm_diag* out) {
    const NcclApi& n = nccl_api();
    if (!n.ok) return fail(TRM_ERR_UNSUPPORTED, "no NCCL library found (libnccl.so.2)");
    if (!nccl_comm) return fail(TRM_ERR_STATE, "trm_diagnostics_allreduce before trm_nccl_comm_init / trm_nccl_comm_adopt");
    if (int rc = diagnostics(nullptr, nullptr)) return rc;   // local partials -> diag_out (device)
    if (!diag_pack) { if (int rc = dalloc(&diag_pack, 8)) return rc; }
    diag_pack_kernel<<<1, 32, 0, stream>>>(diag_out, diag_pack);
    if (int rc = n.AllReduce(diag_pack, diag_pack, 4, kNcclFloat64, kNcclSum, nccl_comm, stream)) return nccl_fail(rc, "ncclAllReduce(sum)");
    if (int rc = n.AllReduce(diag_pack + 4, diag_pack + 4, 4, kNcclFloat64, kNcclMin, nccl_comm, stream)) return nccl_fail(rc, "ncclAllReduce(min)");
    diag_unpack_kernel<<<1, 32, 0, stream>>>(diag_pack, diag_out);
    launches += 2;
    CU(cudaGetLastError());
    double v[8];
    CU(cudaMemcpyAsync(v, diag_out, sizeof(v), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    out->energy = v[0]; out->water = v[1]; out->t_min = v[2]; out->t_max = v[3];
    out->sat_min = v[4]; out->sat_max = v[5]; out->nan_count = v[6]; out->ncol = v[7];
    return TRM_OK;
}

HandleBase* H(trm_handle* h) { return reinterpret_cast<HandleBase*>(h); }
bool bad_input(int id) { return id < 0 || id >= TRM_IN_COUNT; }

}  // namespace

extern "C" {

void trm_default_params(trm_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->mineral_porosity = 0.49; p->organic_porosity = 0.9; p->rho_soc = 0.0; p->rho_org = 1300.0;
    const double k[5] = {0.57, 2.2, 0.025, 3.8, 0.25};
    const double c[5] = {4.2e6, 1.9e6, 0.00125e6, 2.0e6, 2.5e6};
    for (int i = 0; i < 5; ++i) { p->kappa[i] = k[i]; p->heatcap[i] = c[i]; }
    p->rho_w = 1000.0; p->Lsl = 3.34e5; p->Llg = 2.257e6; p->rho_a = 1.293; p->c_a = 1005.7; p->Tref = 273.15;
    p->sigma = 5.6704e-8; p->eps_mw = 0.622;
    p->K_sat = 1.0e-5; p->vg_alpha = 1.0; p->vg_n = 2.0; p->bc_psis = 0.01; p->bc_lambda = 0.2; p->theta_res = 0.0;
    p->impedance = 7.0; p->vwc_forcing = 0.0;
    p->albedo = 0.3; p->emissivity = 0.97; p->kappa_skin = 2.0; p->C_h = 1.2e-3; p->min_windspeed = 0.01;
    p->tau_r = 3600.0; p->evap_beta = 1.0;
    p->field_capacity = 0.25; p->wilting_point = 0.05; p->C_mass = 12.0;
    p->tau25 = 2600.0; p->Kc25 = 30.0; p->Ko25 = 3.0e4; p->q10_tau = 0.57; p->q10_Kc = 2.1; p->q10_Ko = 1.2;
    p->alpha_leaf = 0.17; p->alpha_a = 0.5; p->alpha_C3 = 0.08; p->cq = 4.6e-6; p->k_ext = 0.5;
    p->T_CO2_high = 42.0; p->T_CO2_low = -4.0; p->T_photos_high = 30.0; p->T_photos_low = 15.0; p->theta_r = 0.7;
    p->g1 = 2.3; p->g_min = 0.5; p->cn_sapwood = 330.0; p->cn_root = 29.0; p->aws = 10.0;
    p->SLA = 10.0; p->awl = 2.0; p->LAI_min = 1.0; p->LAI_max = 6.0; p->gamma_L = 0.3; p->gamma_R = 0.3; p->gamma_S = 0.05;
    p->nu_seed = 0.001; p->gamma_v_min = 0.002; p->root_a = 7.0; p->root_b = 2.0;
    p->alpha_int = 0.2; p->k_ext_can = 0.5; p->w_can_max = 2.0e-4; p->tau_w = 86400.0; p->C_can = 0.006;
}

void trm_default_config(trm_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = TRM_ABI_VERSION; c->dtype = TRM_F64; c->model = TRM_MODEL_SOIL; c->timestepper = TRM_EULER;
    c->hydrology = TRM_NOFLOW; c->swrc = TRM_SWRC_BROOKSCOREY; c->unsat_k = TRM_UNSATK_LINEAR; c->sat_halo = TRM_HALO_ZERO;
    c->skin = TRM_SKIN_IMPLICIT; c->math = TRM_MATH_FAITHFUL;
    trm_default_params(&c->params);
}

const char* trm_last_error(void) { return g_err.c_str(); }
int trm_abi_version(void) { return TRM_ABI_VERSION; }

int trm_create(const trm_config* cfg, trm_handle** out) {
    if (!cfg || !out) return fail(TRM_ERR_INVALID, "null argument");
    if (cfg->abi_version != TRM_ABI_VERSION) return fail(TRM_ERR_INVALID, "ABI version mismatch");
    if (cfg->nz < 2 || cfg->nz > TRM_MAX_NZ || cfg->ncol < 1 || !cfg->z_faces) return fail(TRM_ERR_INVALID, "bad nz / ncol / z_faces");
    for (int k = 0; k < cfg->nz; ++k) if (!(cfg->z_faces[k + 1] > cfg->z_faces[k])) return fail(TRM_ERR_INVALID, "z_faces must increase");
    HandleBase* h = nullptr; int rc;
    if (cfg->dtype == TRM_F32) { auto* q = new Handle<float>(); rc = q->setup(*cfg); h = q; }
    else if (cfg->dtype == TRM_F64) { auto* q = new Handle<double>(); rc = q->setup(*cfg); h = q; }
    else return fail(TRM_ERR_INVALID, "bad dtype");
    if (rc != TRM_OK) { std::string keep = g_err; delete h; g_err = keep; return rc; }
    *out = reinterpret_cast<trm_handle*>(h);
    return TRM_OK;
}
int trm_destroy(trm_handle* h) { if (h) delete H(h); return TRM_OK; }
int trm_sync(trm_handle* h) {
    if (!h) return fail(TRM_ERR_INVALID, "null handle");
    return H(h)->sync_all();
}
int trm_field_ptr(trm_handle* h, int id, void** p, int64_t* ld, int32_t* nrows) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->field_ptr(id, p, ld, nrows); }
int trm_field_view(trm_handle* h, int id, const void** p, int64_t* ld, int32_t* nrows) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->field_view(id, p, ld, nrows); }
int trm_set_field(trm_handle* h, int id, const void* host, int64_t count) { if (!h || !host) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->set_field(id, host, count); }
int trm_get_field(trm_handle* h, int id, void* host, int64_t count) { if (!h || !host) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->get_field(id, host, count); }
int trm_set_input_const(trm_handle* h, int id, double v) { if (!h || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id"); return H(h)->set_input_const(id, v); }
int trm_set_input_field(trm_handle* h, int id, const void* v) { if (!h || !v || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id"); return H(h)->set_input_field(id, v); }
int trm_set_input_field_pair(trm_handle* h, int id, const void* v0, const void* v1) {
    if (!h || !v0 || !v1 || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id");
    return H(h)->set_input_field_pair(id, v0, v1);
}
int trm_set_input_sinusoid(trm_handle* h, int id, const void* mean, const void* amp, const void* phase, double period, double lo, double hi) {
    if (!h || !mean || !amp || !phase || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id");
    return H(h)->set_input_sinusoid(id, mean, amp, phase, period, lo, hi);
}
int trm_set_input_table(trm_handle* h, int id, int32_t nt, const double* times, const void* values) {
    if (!h || !times || !values || bad_input(id) || nt < 1) return fail(TRM_ERR_INVALID, "bad handle / input id / nt");
    return H(h)->set_input_table(id, nt, times, values);
}
int trm_set_input_raster(trm_handle* h, int id, int32_t nt, const double* times, const void* values) {
    if (!h || !times || !values || bad_input(id) || nt < 1) return fail(TRM_ERR_INVALID, "bad handle / input id / table");
    for (int i = 1; i < nt; ++i) if (!(times[i] > times[i - 1])) return fail(TRM_ERR_INVALID, "raster time axis must increase strictly");
    return H(h)->set_input_table(id, nt, times, values, TRM_SRC_RASTER);
}
int trm_get_input(trm_handle* h, int id, void* host, int64_t count) { if (!h || !host || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id"); return H(h)->get_input(id, host, count); }
int trm_input_ptr(trm_handle* h, int id, void** p) { if (!h || !p || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id"); return H(h)->input_ptr(id, p); }
int trm_initialize(trm_handle* h) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->initialize(); }
int trm_step(trm_handle* h, double dt, int64_t n) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->step(dt, n); }
int trm_compute_auxiliary(trm_handle* h) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->aux(); }
int trm_compute_tendencies(trm_handle* h) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->tendencies(); }
int trm_get_clock(trm_handle* h, double* t, int64_t* it) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); if (t) *t = H(h)->time; if (it) *it = H(h)->iteration; return TRM_OK; }
int trm_set_clock(trm_handle* h, double t, int64_t it) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); H(h)->set_clock(t, it); return TRM_OK; }
int trm_reset(trm_handle* h) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->reset_state(); }
int trm_bind_host_io(trm_handle* h, int input_id, const void* host_in, int field_id, void* host_out, int32_t nslots) {
    if (!h) return fail(TRM_ERR_INVALID, "null handle");
    return H(h)->bind_host_io(input_id, host_in, field_id, host_out, nslots);
}
int trm_host_io_wait(trm_handle* h, int64_t iteration) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->host_io_wait(iteration); }
int trm_diagnostics(trm_handle* h, trm_diag* out) { if (!h || !out) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->diagnostics(out, nullptr); }
int trm_diagnostics_device(trm_handle* h, double** dev) { if (!h || !dev) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->diagnostics(nullptr, dev); }
int trm_nccl_get_unique_id(void* id) {
    if (!id) return fail(TRM_ERR_INVALID, "null argument");
    const NcclApi& n = nccl_api();
    if (!n.ok) return fail(TRM_ERR_UNSUPPORTED, "no NCCL library found (libnccl.so.2)");
    if (int rc = n.GetUniqueId(id)) return nccl_fail(rc, "ncclGetUniqueId");
    return TRM_OK;
}
int trm_nccl_comm_init(trm_handle* h, int32_t nranks, int32_t rank, const void* id) {
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(TRM_ERR_INVALID, "bad handle / rank / id");
    const NcclApi& n = nccl_api();
    if (!n.ok) return fail(TRM_ERR_UNSUPPORTED, "no NCCL library found (libnccl.so.2)");
    HandleBase* b = H(h);
    if (b->nccl_comm) return fail(TRM_ERR_STATE, "the handle already has a communicator");
    if (cudaSetDevice(b->device) != cudaSuccess) return fail(TRM_ERR_CUDA, "cudaSetDevice");
    NcclApi::Id uid; std::memcpy(uid.b, id, sizeof(uid.b));
    void* comm = nullptr;
    if (int rc = n.CommInitRank(&comm, nranks, uid, rank)) return nccl_fail(rc, "ncclCommInitRank");
    b->nccl_comm = comm; b->nccl_owned = true;
    return TRM_OK;
}
int trm_nccl_comm_adopt(trm_handle* h, void* comm) {
    if (!h || !comm) return fail(TRM_ERR_INVALID, "null argument");
    if (!nccl_api().ok) return fail(TRM_ERR_UNSUPPORTED, "no NCCL library found (libnccl.so.2)");
    HandleBase* b = H(h);
    if (b->nccl_comm && b->nccl_owned) nccl_api().CommDestroy(b->nccl_comm);
    b->nccl_comm = comm; b->nccl_owned = false;
    return TRM_OK;
}
int trm_diagnostics_allreduce(trm_handle* h, trm_diag* out) { if (!h || !out) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->diagnostics_allreduce(out); }
int64_t trm_launch_count(trm_handle* h) { return h ? H(h)->launches : 0; }
int trm_last_step_ms(trm_handle* h, float* ms) { if (!h || !ms) return fail(TRM_ERR_INVALID, "null argument"); *ms = H(h)->last_ms; return TRM_OK; }
int trm_set_input_field_async(trm_handle* h, int id, const void* v) { if (!h || !v || bad_input(id)) return fail(TRM_ERR_INVALID, "bad handle / input id"); return H(h)->set_input_field_async(id, v); }
int trm_get_field_async(trm_handle* h, int id, void* host, int64_t count) { if (!h || !host) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->get_field_async(id, host, count); }
int trm_accumulate(trm_handle* h, int id, double w) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->accumulate(id, w); }
int trm_get_accumulated(trm_handle* h, int id, void* host, int64_t count, double scale, int32_t reset) {
    if (!h || !host) return fail(TRM_ERR_INVALID, "null argument");
    return H(h)->get_accumulated(id, host, count, scale, reset);
}
int trm_host_alloc(int64_t bytes, void** host) {
    if (!host || bytes <= 0) return fail(TRM_ERR_INVALID, "trm_host_alloc: bad arguments");
    cudaError_t e = cudaHostAlloc(host, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TRM_ERR_NO_DEVICE : TRM_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    return TRM_OK;
}
int trm_host_alloc_ex(int64_t bytes, int32_t flags, void** host) {
    if (!host || bytes <= 0 || (flags & ~TRM_HOST_WRITE_COMBINED)) return fail(TRM_ERR_INVALID, "trm_host_alloc_ex: bad arguments");
    unsigned f = cudaHostAllocPortable | cudaHostAllocMapped;
    if (flags & TRM_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    cudaError_t e = cudaHostAlloc(host, (size_t)bytes, f);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? TRM_ERR_NO_DEVICE : TRM_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    return TRM_OK;
}
int trm_host_free(void* host) {
    if (!host) return TRM_OK;
    cudaError_t e = cudaFreeHost(host);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(TRM_ERR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e)); }
    return TRM_OK;
}
int trm_step_async(trm_handle* h, double dt, int64_t n) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->step_async(dt, n); }
int trm_set_ring_index(trm_handle* h, const int64_t* ring_index, int64_t nring) { if (!h || !ring_index) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->set_ring_index(ring_index, nring); }
int trm_get_field_ring(trm_handle* h, int id, void* host, int64_t count, double fill) { if (!h || !host) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->get_field_ring(id, host, count, fill); }
int trm_set_field_ring(trm_handle* h, int id, const void* host, int64_t count) { if (!h || !host) return fail(TRM_ERR_INVALID, "null argument"); return H(h)->set_field_ring(id, host, count); }
int trm_set_block_size(trm_handle* h, int block) { if (!h) return fail(TRM_ERR_INVALID, "null handle"); return H(h)->set_block(block); }

}  // extern "C"
