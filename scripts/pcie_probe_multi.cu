// Aggregate host -> device bandwidth of one box with 1, 2, 4, 8 GPUs copying AT THE SAME TIME (one host thread and one
// page-locked 1 GiB buffer per GPU): is the e2e path of N ranks bound by a platform ceiling (host memory / root complexes)
// rather than by anything in the library? Two legs per GPU count: plain cudaMemcpyAsync (copy engine) and SM loads of
// 2 KB pieces every 10 kB out of mapped host memory (the shape of the ASCII window pull).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -pthread -o scripts/pcie_probe_multi.bin scripts/pcie_probe_multi.cu
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void pull_ldg(const unsigned char* __restrict__ h, size_t L, int rows, int W, unsigned* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    unsigned acc = 0;
    for (int r = warp; r < rows; r += nW) {
        const uint4* p = reinterpret_cast<const uint4*>(h + (size_t)r * L);
        uint4 v[5];
#pragma unroll
        for (int i = 0; i < 5; i++) v[i] = (lane + 32 * i) * 16 < W ? __ldg(p + lane + 32 * i) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < 5; i++) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    int nDev = 0;
    CK(cudaGetDeviceCount(&nDev));
    const size_t bytes = 1ull << 30;
    std::vector<unsigned char*> host((size_t)nDev), dev((size_t)nDev), hostDev((size_t)nDev);
    std::vector<unsigned*> outp((size_t)nDev);
    for (int d = 0; d < nDev; d++) {
        CK(cudaSetDevice(d));
        CK(cudaHostAlloc((void**)&host[d], bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(host[d], 65, bytes);
        CK(cudaHostGetDevicePointer((void**)&hostDev[d], host[d], 0));
        CK(cudaMalloc((void**)&dev[d], bytes));
        CK(cudaMalloc((void**)&outp[d], 4));
    }
    printf("%d GPUs, %u host threads available\n", nDev, std::thread::hardware_concurrency());
    for (int leg = 0; leg < 2; leg++) {
        for (int n = 1; n <= nDev; n *= 2) {
            std::vector<double> gbs((size_t)n, 0.0);
            std::atomic<int> ready(0);
            std::atomic<int> go(0);
            double t0 = 0, t1 = 0;
            std::vector<std::thread> th;
            const int reps = 8;
            for (int d = 0; d < n; d++) {
                th.emplace_back([&, d]() {
                    CK(cudaSetDevice(d));
                    cudaStream_t st;
                    CK(cudaStreamCreate(&st));
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a));
                    CK(cudaEventCreate(&b));
                    const int W = 2064, rows = (int)(bytes / 10000) - 1;
                    // warm-up
                    if (leg == 0) CK(cudaMemcpyAsync(dev[d], host[d], bytes, cudaMemcpyHostToDevice, st));
                    else pull_ldg<<<296, 256, 0, st>>>(hostDev[d], 10000, rows, W, outp[d]);
                    CK(cudaStreamSynchronize(st));
                    ready.fetch_add(1);
                    while (!go.load()) {}
                    CK(cudaEventRecord(a, st));
                    for (int r = 0; r < reps; r++) {
                        if (leg == 0) CK(cudaMemcpyAsync(dev[d], host[d], bytes, cudaMemcpyHostToDevice, st));
                        else pull_ldg<<<296, 256, 0, st>>>(hostDev[d], 10000, rows, W, outp[d]);
                    }
                    CK(cudaEventRecord(b, st));
                    CK(cudaStreamSynchronize(st));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, a, b));
                    const double moved = leg == 0 ? (double)bytes * reps : (double)rows * W * reps;
                    gbs[(size_t)d] = moved / 1e6 / ms;
                });
            }
            while (ready.load() < n) {}
            t0 = now_s();
            go.store(1);
            for (auto& t : th) t.join();
            t1 = now_s();
            double sum = 0, mn = 1e30;
            for (double g : gbs) {
                sum += g;
                if (g < mn) mn = g;
            }
            printf("%s, %d GPUs at once: aggregate %.1f GB/s (slowest GPU %.1f GB/s, per-GPU mean %.1f), wall %.0f ms\n",
                   leg == 0 ? "cudaMemcpyAsync 1 GiB x 8   " : "SM loads of 2 KB pieces x 8 ", n, sum, mn, sum / n, (t1 - t0) * 1e3);
        }
    }
    return 0;
}
