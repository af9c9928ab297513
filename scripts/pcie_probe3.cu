// Pulls of SMALL pieces (the windows of packed reads: 272-560 bytes every 2500) out of pinned host memory: what bounds
// them — piece size, alignment, or requests in flight? TMA bulk copies (one issuing thread per CTA, R slots), LDG.128
// warps and a strided 2D DMA, each over piece sizes 272 / 560 / 1040 / 2064 B at a 16 B and at a 128 B alignment.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pcie_probe3.bin scripts/pcie_probe3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void pull_ldg(const unsigned char* __restrict__ h, size_t L, int rows, int W, unsigned* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    unsigned acc = 0;
    const int per = (W + 15) / 16;  // 16-byte blocks per piece
    if (per <= 32) {                // several pieces per warp trip: lanes [0,per) piece a, [per,2per) piece b ...
        const int fit = 32 / per;
        const int sub = lane / per, off = lane % per;
        for (int r = warp * fit; r < rows; r += nW * fit) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (sub < fit && r + sub < rows) v = __ldg(reinterpret_cast<const uint4*>(h + (size_t)(r + sub) * L) + off);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    } else {
        for (int r = warp; r < rows; r += nW) {
            const uint4* p = reinterpret_cast<const uint4*>(h + (size_t)r * L);
            uint4 v[5];
#pragma unroll
            for (int i = 0; i < 5; i++) v[i] = (lane + 32 * i) * 16 < W ? __ldg(p + lane + 32 * i) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 5; i++) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int R>
__global__ void pull_tma(const unsigned char* __restrict__ h, size_t L, int rows, int W, unsigned* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[R];
    if (threadIdx.x != 0) return;
    const int slotBytes = (W + 127) / 128 * 128;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars), dst0 = (unsigned)__cvta_generic_to_shared(smem);
    for (int s = 0; s < R; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned acc = 0;
    int issued = 0, done = 0;
    const int first = blockIdx.x, step = gridDim.x;
    const int mine = first < rows ? (rows - first + step - 1) / step : 0;
    while (done < mine) {
        while (issued < mine && issued - done < R) {
            const int s = issued % R;
            const unsigned char* src = h + (size_t)(first + (size_t)issued * step) * L;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(W) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst0 + (unsigned)(s * slotBytes)),
                         "l"(src), "r"(W), "r"(bar0 + 8u * s)
                         : "memory");
            issued++;
        }
        const int s = done % R;
        const unsigned parity = (unsigned)(done / R) & 1u;
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(bar0 + 8u * s), "r"(parity)
                         : "memory");
        }
        acc += *reinterpret_cast<volatile unsigned*>(smem + (size_t)s * slotBytes);
        done++;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int R>
float run_tma(int ctas, const unsigned char* hd, size_t L, int rows, int W, unsigned* out, cudaEvent_t a, cudaEvent_t b) {
    const int slot = (W + 127) / 128 * 128;
    CK(cudaFuncSetAttribute(pull_tma<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a));
        pull_tma<R><<<ctas, 32, (size_t)R * slot>>>(hd, L, rows, W, out);
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
    }
    return ms;
}

int main() {
    const size_t L = 2500, nReads = 1000000;  // packed 10 kb reads, back to back
    const size_t hostBytes = nReads * L + 4096;
    unsigned char* h;
    CK(cudaHostAlloc((void**)&h, hostBytes, cudaHostAllocMapped));
    memset(h, 65, hostBytes);
    unsigned char* hd0;
    CK(cudaHostGetDevicePointer((void**)&hd0, h, 0));
    unsigned* out;
    CK(cudaMalloc((void**)&out, 4));
    unsigned char* dst;
    CK(cudaMalloc((void**)&dst, 2 * nReads * 2176));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float ms;
    const int Ws[4] = {272, 560, 1040, 2064};
    for (int wi = 0; wi < 4; wi++) {
        const int W = Ws[wi];
        // stride: every read for <= 560 B pieces (one merged tail+head piece per read); 4 reads for the larger ones
        const size_t stride = W <= 560 ? (W <= 272 ? L / 2 + 6 : L) : (W <= 1040 ? 2 * L : 4 * L);  // multiples of 4 B
        const int rows = (int)((nReads * L - 4096) / stride);
        const double mb = (double)rows * W / 1e6;
        for (int al = 0; al < 2; al++) {
            // al = 0: the pieces start at 16-byte aligned addresses that drift over the 128-byte lines (stride % 128 != 0)
            // al = 1: stride rounded to a multiple of 128: every piece starts on a 128-byte line
            const size_t st = al ? (stride + 127) / 128 * 128 : (stride + 15) / 16 * 16;
            const unsigned char* hd = hd0 + (al ? 0 : 16);
            printf("--- piece %d B every %zu B (%s), %d pieces = %.0f MB\n", W, st, al ? "128 B aligned" : "16 B aligned", rows, mb);
            const int rowsFit = (int)((nReads * L - 4096) / st);
            const int rws = rows < rowsFit ? rows : rowsFit;
            const double mbb = (double)rws * W / 1e6;
            for (int ctas = 148; ctas <= 148 * 8; ctas *= 2) {
                for (int rep = 0; rep < 2; rep++) {
                    CK(cudaEventRecord(a));
                    pull_ldg<<<ctas, 256>>>(hd, st, rws, W, out);
                    CK(cudaEventRecord(b));
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventElapsedTime(&ms, a, b));
                }
                printf("LDG.128 %4d CTAs x 8 warps        : %.3f ms = %.1f GB/s, %.1f M pieces/s\n", ctas, ms, mbb / ms, rws / ms / 1e3);
            }
            const int cs[5] = {16, 32, 64, 128, 296};
            for (int ci = 0; ci < 5; ci++) {
                ms = run_tma<8>(cs[ci], hd, st, rws, W, out, a, b);
                printf("TMA bulk %4d CTAs x  8 slots       : %.3f ms = %.1f GB/s, %.1f M pieces/s\n", cs[ci], ms, mbb / ms, rws / ms / 1e3);
                ms = run_tma<16>(cs[ci], hd, st, rws, W, out, a, b);
                printf("TMA bulk %4d CTAs x 16 slots       : %.3f ms = %.1f GB/s, %.1f M pieces/s\n", cs[ci], ms, mbb / ms, rws / ms / 1e3);
                ms = run_tma<48>(cs[ci], hd, st, rws, W, out, a, b);
                printf("TMA bulk %4d CTAs x 48 slots       : %.3f ms = %.1f GB/s, %.1f M pieces/s\n", cs[ci], ms, mbb / ms, rws / ms / 1e3);
            }
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(a));
                CK(cudaMemcpy2DAsync(dst, 2176, h + (al ? 0 : 16), st, W, rws, cudaMemcpyHostToDevice, 0));
                CK(cudaEventRecord(b));
                CK(cudaDeviceSynchronize());
                CK(cudaEventElapsedTime(&ms, a, b));
            }
            printf("2D DMA (cudaMemcpy2DAsync)          : %.3f ms = %.1f GB/s, %.1f M pieces/s\n", ms, mbb / ms, rws / ms / 1e3);
        }
    }
    // one plain copy of everything as the yardstick
    unsigned char* big;
    CK(cudaMalloc((void**)&big, 1ull << 30));
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaEventRecord(a));
        CK(cudaMemcpyAsync(big, h, 1ull << 30, cudaMemcpyHostToDevice, 0));
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
    }
    printf("plain 1 GiB cudaMemcpyAsync          : %.3f ms = %.1f GB/s\n", ms, 1073.74 / ms);
    return 0;
}
