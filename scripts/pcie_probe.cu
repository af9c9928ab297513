// PCIe host->device paths for the windowed pull: plain DMA, 2D DMA (uniform reads), batched DMA (ragged reads).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pcie_probe scripts/pcie_probe.cu && /tmp/pcie_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main() {
    const size_t nReads = 131072, L = 10000, W = 2064;  // tail window of read i + head window of read i+1 are adjacent
    const size_t hostBytes = nReads * L;
    unsigned char* h;
    CK(cudaHostAlloc((void**)&h, hostBytes, cudaHostAllocDefault));
    memset(h, 65, hostBytes);
    unsigned char* d;
    CK(cudaMalloc((void**)&d, hostBytes));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, st));
        CK(cudaMemcpyAsync(d, h, hostBytes, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("plain DMA        : %.1f MB in %.3f ms = %.1f GB/s\n", hostBytes / 1e6, ms, hostBytes / ms / 1e6);
    }
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a, st));
        CK(cudaMemcpy2DAsync(d, W, h + L - W / 2, L, W, nReads - 1, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(b, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("2D DMA %zu B rows: %.1f MB in %.3f ms = %.1f GB/s\n", W, (nReads - 1) * W / 1e6, ms, (nReads - 1) * W / ms / 1e6);
    }
#if CUDART_VERSION >= 12080
    {
        std::vector<void*> dsts(nReads - 1), srcs(nReads - 1);
        std::vector<size_t> sizes(nReads - 1, W);
        for (size_t i = 0; i + 1 < nReads; i++) {
            dsts[i] = d + i * W;
            srcs[i] = h + (i + 1) * L - W / 2;
        }
        cudaMemcpyAttributes attr;
        memset(&attr, 0, sizeof(attr));
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t attrIdx = 0, failIdx = 0;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(a, st));
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), nReads - 1, &attr, &attrIdx, 1, &failIdx, st);
            if (e != cudaSuccess) { printf("cudaMemcpyBatchAsync: %s\n", cudaGetErrorString(e)); break; }
            CK(cudaEventRecord(b, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("batched DMA      : %.1f MB in %.3f ms = %.1f GB/s\n", (nReads - 1) * W / 1e6, ms, (nReads - 1) * W / ms / 1e6);
        }
    }
#endif
    return 0;
}
