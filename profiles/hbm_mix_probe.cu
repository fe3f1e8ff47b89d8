// HBM ceiling probe for the access mix of the stage kernel: per cell 2 values read (U, sat) and 5 written (U, sat, T, liq,
// psi), seven separate [layer][column] arrays of 2.4 GB each (10 M columns x 30 layers x Float64).
//   copy      : 1 read + 1 write stream, grid-stride (what MEASURED_PEAKS.json's hbm_gbs measures)
//   mix       : 2 read + 5 write streams, grid-stride over the flat arrays, 16 bytes per thread and access
//   columns   : 2 read + 5 write streams in the kernel's own order -- a block owns 128 (x2) adjacent columns and walks the
//               30 layers bottom to top (one 1 KB / 2 KB row segment per array and layer), evict-first stores
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hbm_mix_probe hbm_mix_probe.cu ; run: ./hbm_mix_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void copy_k(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void mix_k(const double2* __restrict__ r0, const double2* __restrict__ r1, double2* w0, double2* w1, double2* w2, double2* w3, double2* w4, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 u = r0[i], s = r1[i];
        const double2 x = make_double2(u.x + s.x, u.y * s.y);
        __stcs(w0 + i, x); __stcs(w1 + i, s); __stcs(w2 + i, u); __stcs(w3 + i, x); __stcs(w4 + i, s);
    }
}
template <int W>   // W = columns per thread (1: 8-byte accesses, 2: 16-byte accesses)
__global__ void __launch_bounds__(128) columns_k(const double* __restrict__ r0, const double* __restrict__ r1, double* w0, double* w1, double* w2, double* w3, double* w4, size_t ncol, size_t ld, int nz, int ahead) {
    const size_t c = ((size_t)blockIdx.x * 128 + threadIdx.x) * W;
    if (c >= ncol) return;
    double u[8][W], s[8][W];
    for (int k = 0; k < ahead; ++k)
        for (int q = 0; q < W; ++q) { u[k & 7][q] = r0[k * ld + c + q]; s[k & 7][q] = r1[k * ld + c + q]; }
    for (int k = 0; k < nz; ++k) {
        if (k + ahead < nz)
            for (int q = 0; q < W; ++q) { u[(k + ahead) & 7][q] = r0[(k + ahead) * ld + c + q]; s[(k + ahead) & 7][q] = r1[(k + ahead) * ld + c + q]; }
        for (int q = 0; q < W; ++q) {
            const double a = u[k & 7][q], b = s[k & 7][q];
            const size_t o = k * ld + c + q;
            __stcs(w0 + o, a + b); __stcs(w1 + o, b); __stcs(w2 + o, a); __stcs(w3 + o, a * b); __stcs(w4 + o, b - a);
        }
    }
}

// same walk, generic addressing: off(k, c) = (c / TW) * tile_stride + k * layer_stride + (c % TW) ; the seven arrays are
// p + f * field_stride. [layer][column] SoA: TW = ld, layer_stride = ld, tile_stride = 0, field_stride = nz * ld.
// Tiled per field: TW = 128, layer_stride = 128, tile_stride = nz * 128. Tiled over all fields ("AoSoA"): layer_stride = 7 * 128,
// tile_stride = nz * 7 * 128, field_stride = 128.
template <int BLK, int V, int AHEAD = 2>   // V = doubles per access (1 or 2) ; the layer loop is fully unrolled (registers, no local arrays)
__global__ void __launch_bounds__(BLK) walk_k(double* p, size_t ncol, int nz_, size_t TW, size_t layer_stride, size_t tile_stride, size_t field_stride, int ahead_, int plain_stores) {
    constexpr int nz = 30, ahead = AHEAD;
    const size_t c = ((size_t)blockIdx.x * BLK + threadIdx.x) * V;
    if (c >= ncol) return;
    const size_t base = (c / TW) * tile_stride + (c % TW);
    double u[8][V], s[8][V];
    auto ld = [&](int f, int k, double* v) {
        const double* a = p + f * field_stride + base + k * layer_stride;
        if (V == 2) { const double2 t = *reinterpret_cast<const double2*>(a); v[0] = t.x; v[V - 1] = t.y; } else v[0] = *a;
    };
    auto st = [&](int f, int k, const double* v) {
        double* a = p + f * field_stride + base + k * layer_stride;
        if (plain_stores) { if (V == 2) *reinterpret_cast<double2*>(a) = make_double2(v[0], v[V - 1]); else *a = v[0]; }
        else { if (V == 2) __stcs(reinterpret_cast<double2*>(a), make_double2(v[0], v[V - 1])); else __stcs(a, v[0]); }
    };
#pragma unroll
    for (int k = 0; k < ahead; ++k) { ld(0, k, u[k & 7]); ld(1, k, s[k & 7]); }
#pragma unroll
    for (int k = 0; k < nz; ++k) {
        if (k + ahead < nz) { ld(0, k + ahead, u[(k + ahead) & 7]); ld(1, k + ahead, s[(k + ahead) & 7]); }
        double x[V], y[V], z[V];
        for (int q = 0; q < V; ++q) { x[q] = u[k & 7][q] + s[k & 7][q]; y[q] = u[k & 7][q] * s[k & 7][q]; z[q] = s[k & 7][q] - u[k & 7][q]; }
        st(2, k, x); st(3, k, s[k & 7]); st(4, k, u[k & 7]); st(5, k, y); st(6, k, z);
    }
}

int main() {
    const size_t ncol = 10000000, ld = 10000000, nz = 30, n = ld * nz;
    double* p[7];
    for (int i = 0; i < 7; ++i) { CK(cudaMalloc(&p[i], n * 8)); CK(cudaMemset(p[i], 0, n * 8)); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto f, double bytes, const char* name) {
        for (int i = 0; i < 3; ++i) f();
        cudaEventRecord(e0);
        const int reps = 10;
        for (int i = 0; i < reps; ++i) f();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%-46s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms * 1e-6);
    };
    const int nb = 148 * 16;
    time([&] { copy_k<<<nb, 256>>>((const double2*)p[0], (double2*)p[1], n / 2); }, 2.0 * n * 8, "copy 1R+1W (2.4 GB arrays)");
    time([&] { mix_k<<<nb, 256>>>((const double2*)p[0], (const double2*)p[1], (double2*)p[2], (double2*)p[3], (double2*)p[4], (double2*)p[5], (double2*)p[6], n / 2); }, 7.0 * n * 8, "mix 2R+5W grid-stride");
    time([&] { mix_k<<<148 * 6, 128>>>((const double2*)p[0], (const double2*)p[1], (double2*)p[2], (double2*)p[3], (double2*)p[4], (double2*)p[5], (double2*)p[6], n / 2); }, 7.0 * n * 8, "mix 2R+5W grid-stride, 6 x 128 per SM");
    for (int ahead : {1, 2, 4}) {
        char name[96];
        snprintf(name, sizeof name, "columns 2R+5W, 1 col/thread, %d layers ahead", ahead);
        time([&] { columns_k<1><<<(unsigned)((ncol + 127) / 128), 128>>>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], ncol, ld, nz, ahead); }, 7.0 * n * 8, name);
        snprintf(name, sizeof name, "columns 2R+5W, 2 col/thread, %d layers ahead", ahead);
        time([&] { columns_k<2><<<(unsigned)((ncol / 2 + 127) / 128), 128>>>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], ncol, ld, nz, ahead); }, 7.0 * n * 8, name);
    }
    // one allocation for the layout variants
    for (int i = 0; i < 7; ++i) cudaFree(p[i]);
    double* q; CK(cudaMalloc(&q, 7 * n * 8)); CK(cudaMemset(q, 0, 7 * n * 8));
    const double B7 = 7.0 * n * 8;
    auto run = [&](auto kern, int blk, int v, size_t TW, size_t ls, size_t ts, size_t fs, int ahead, int plain, const char* name) {
        time([&] { kern<<<(unsigned)((ncol / v + blk - 1) / blk), blk>>>(q, ncol, (int)nz, TW, ls, ts, fs, ahead, plain); }, B7, name);
    };
    run(walk_k<128, 1>, 128, 1, ld, ld, 0, nz * ld, 2, 0, "walk SoA [layer][col], 128 thr x 8 B");
    run(walk_k<128, 1, 4>, 128, 1, ld, ld, 0, nz * ld, 4, 0, "walk SoA, 128 thr x 8 B, 4 ahead");
    run(walk_k<128, 1, 6>, 128, 1, ld, ld, 0, nz * ld, 6, 0, "walk SoA, 128 thr x 8 B, 6 ahead");
    run(walk_k<128, 2, 4>, 128, 2, ld, ld, 0, nz * ld, 4, 0, "walk SoA, 128 thr x 16 B, 4 ahead");
    run(walk_k<128, 1>, 128, 1, ld, ld, 0, nz * ld, 2, 1, "walk SoA, 128 thr x 8 B, plain stores");
    run(walk_k<128, 2>, 128, 2, ld, ld, 0, nz * ld, 2, 0, "walk SoA, 128 thr x 16 B");
    run(walk_k<256, 2>, 256, 2, ld, ld, 0, nz * ld, 2, 0, "walk SoA, 256 thr x 16 B");
    run(walk_k<512, 2>, 512, 2, ld, ld, 0, nz * ld, 2, 0, "walk SoA, 512 thr x 16 B");
    run(walk_k<1024, 2>, 1024, 2, ld, ld, 0, nz * ld, 2, 0, "walk SoA, 1024 thr x 16 B");
    run(walk_k<128, 1>, 128, 1, 128, 128, nz * 128, nz * ld, 2, 0, "walk tiled per field (128 col), 128 thr x 8 B");
    run(walk_k<128, 2>, 128, 2, 256, 256, nz * 256, nz * ld, 2, 0, "walk tiled per field (256 col), 128 thr x 16 B");
    run(walk_k<128, 1>, 128, 1, 128, 7 * 128, nz * 7 * 128, 128, 2, 0, "walk AoSoA [tile][layer][field][128], 8 B");
    run(walk_k<128, 2>, 128, 2, 256, 7 * 256, nz * 7 * 256, 256, 2, 0, "walk AoSoA [tile][layer][field][256], 16 B");
    run(walk_k<128, 2>, 128, 2, 256, 7 * 256, nz * 7 * 256, 256, 2, 1, "walk AoSoA [256], 16 B, plain stores");
    return 0;
}
