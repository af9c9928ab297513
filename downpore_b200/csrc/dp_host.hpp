// downpore_b200 — host-side helpers shared by the translation units of the library (dp_api.cu: `map`; dp_overlap_api.cu:
// `overlap`): CUDA error check, grow-only device / mapped host buffers, event pair.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

// the thread-local message dp_last_error() returns (defined in dp_api.cu)
void dp_set_last_error(const char* msg);

namespace {

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " __FILE__ ":" + \
                                     std::to_string(__LINE__));                                               \
    } while (0)

template <class T>
struct DBuf {  // device buffer, grow-only
    T* p = nullptr;
    size_t cap = 0;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    void reserve(size_t n) {
        if (n <= cap) return;
        release();
        size_t want = n + n / 8 + 64;
        CK(cudaMalloc((void**)&p, want * sizeof(T)));
        cap = want;
    }
    void reserve_exact(size_t n) {  // (for a buffer that trades places with another: equal capacities end the growth)
        if (n <= cap) return;
        release();
        CK(cudaMalloc((void**)&p, n * sizeof(T)));
        cap = n;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

template <class T>
struct HBuf {  // page-locked host buffer mapped into the device address space (d = device view), grow-only
    T* p = nullptr;
    T* d = nullptr;
    size_t cap = 0;
    ~HBuf() {
        if (p) cudaFreeHost(p);
    }
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t want = n + n / 8 + 64;
        CK(cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&d, p, 0));
        cap = want;
    }
    void reserve_exact(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        CK(cudaHostAlloc((void**)&p, n * sizeof(T), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&d, p, 0));
        cap = n;
    }
};

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init() {
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
    }
    ~Timer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace
