// hostio_probe.cu -- what limits the per-step host exchange of a coupled run on N GPUs of one box?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o hostio_probe profiles/hostio_probe.cu -lpthread
//   ./hostio_probe [ngpus] [MB per direction per GPU] [repetitions]
//
// One host thread per GPU, all GPUs at the same time (barrier before every timed region), per direction `MB` of
// page-locked host memory per repetition -- the traffic pattern of bench.py's end-to-end leg (10 M columns x 8 B
// split over the GPUs, in and out, every step). Legs:
//   h2d / d2h / both      cudaMemcpyAsync on one / two copy streams
//   kread / kwrite / kboth a kernel reading from / writing to MAPPED page-locked host memory (what trm_bind_host_io does)
//   kread_wc              same read from write-combined host memory
// Prints per-GPU and aggregate GB/s per direction. Timing: host clock around the region after a barrier, max over GPUs.
#include <cuda_runtime.h>
#include <pthread.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static pthread_barrier_t bar;
static int ngpus = 1, reps = 50;
static size_t bytes = 10u << 20;
static double results[16][8];   // [gpu][leg] seconds

__global__ void kcopy(const double* __restrict__ src, double* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void* worker(void* arg) {
    const int g = (int)(size_t)arg;
    CK(cudaSetDevice(g));
    const size_t n = bytes / 8;
    double *hin, *hout, *hwc, *din, *dout;
    CK(cudaHostAlloc(&hin, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&hout, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    CK(cudaHostAlloc(&hwc, bytes, cudaHostAllocPortable | cudaHostAllocMapped | cudaHostAllocWriteCombined));
    for (size_t i = 0; i < n; ++i) { hin[i] = (double)i; hwc[i] = (double)i; hout[i] = 0; }
    CK(cudaMalloc(&din, bytes)); CK(cudaMalloc(&dout, bytes));
    CK(cudaMemset(dout, 0, bytes));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    double *mhin, *mhout, *mhwc;
    CK(cudaHostGetDevicePointer(&mhin, hin, 0)); CK(cudaHostGetDevicePointer(&mhout, hout, 0)); CK(cudaHostGetDevicePointer(&mhwc, hwc, 0));
    const int blocks = 148 * 2, threads = 256;
    for (int leg = 0; leg < 7; ++leg) {
        for (int phase = 0; phase < 2; ++phase) {   // phase 0 = warm-up
            const int r = phase ? reps : 3;
            pthread_barrier_wait(&bar);
            const double t0 = now();
            for (int i = 0; i < r; ++i) {
                switch (leg) {
                    case 0: CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s1)); break;
                    case 1: CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s2)); break;
                    case 2: CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s1));
                            CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s2)); break;
                    case 3: kcopy<<<blocks, threads, 0, s1>>>(mhin, din, n); break;
                    case 4: kcopy<<<blocks, threads, 0, s2>>>(dout, mhout, n); break;
                    case 5: kcopy<<<blocks, threads, 0, s1>>>(mhin, din, n); kcopy<<<blocks, threads, 0, s2>>>(dout, mhout, n); break;
                    case 6: kcopy<<<blocks, threads, 0, s1>>>(mhwc, din, n); break;
                }
            }
            CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
            const double t1 = now();
            pthread_barrier_wait(&bar);
            if (phase) results[g][leg] = t1 - t0;
        }
    }
    return nullptr;
}

int main(int argc, char** argv) {
    if (argc > 1) ngpus = atoi(argv[1]);
    if (argc > 2) bytes = (size_t)atoi(argv[2]) << 20;
    if (argc > 3) reps = atoi(argv[3]);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ngpus > ndev) ngpus = ndev;
    pthread_barrier_init(&bar, nullptr, ngpus);
    std::vector<pthread_t> th(ngpus);
    for (int g = 0; g < ngpus; ++g) pthread_create(&th[g], nullptr, worker, (void*)(size_t)g);
    for (int g = 0; g < ngpus; ++g) pthread_join(th[g], nullptr);
    const char* names[7] = {"h2d (memcpy)", "d2h (memcpy)", "both (memcpy)", "kread (mapped)", "kwrite (mapped)", "kboth (mapped)", "kread (mapped, WC)"};
    printf("gpus %d, %zu MB per direction per GPU per repetition, %d repetitions\n", ngpus, bytes >> 20, reps);
    for (int leg = 0; leg < 7; ++leg) {
        double worst = 0;
        for (int g = 0; g < ngpus; ++g) worst = results[g][leg] > worst ? results[g][leg] : worst;
        const double per_dir = (double)bytes * reps / worst / 1e9;
        printf("%-20s %8.3f ms per repetition   %7.1f GB/s per GPU per direction   %7.1f GB/s aggregate per direction\n",
               names[leg], 1e3 * worst / reps, per_dir, per_dir * ngpus);
    }
    return 0;
}
