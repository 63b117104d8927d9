// Shared host-side plumbing for libb200audio: status codes, error text, launch counting,
// device buffers.  sm_100a only; no CPU fallback anywhere.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200audio_internal.h"

namespace b2a {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B2A_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw b2a::Error(B2A_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define B2A_CHECK(cond, code, msg)                         \
    do {                                                   \
        if (!(cond)) throw b2a::Error((code), (msg));      \
    } while (0)

// Runs `fn`, maps exceptions to status codes (never lets one cross the C ABI).
template <class F>
static inline int32_t guarded(F&& fn) {
    try {
        fn();
        return B2A_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return B2A_ERR_GENERATION_FAILED;
    } catch (...) {
        set_last_error("unknown error");
        return B2A_ERR_GENERATION_FAILED;
    }
}

void require_device(int device);  // throws B2A_ERR_CUDA when there is no usable device

// RAII device buffer (cudaMalloc / cudaFree), grow-only resize.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    ~DBuf() { if (p) cudaFree(p); }
    void alloc(size_t count) {
        if (count <= n) return;
        if (p) cudaFree(p);
        p = nullptr;
        B2A_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void upload(const T* host, size_t count, cudaStream_t s = 0) {
        alloc(count);
        B2A_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

// Pinned host staging buffer, grow-only.
template <class T>
struct HBuf {
    T* p = nullptr;
    size_t n = 0;
    HBuf() = default;
    HBuf(const HBuf&) = delete;
    HBuf& operator=(const HBuf&) = delete;
    ~HBuf() { if (p) cudaFreeHost(p); }
    void alloc(size_t count) {
        if (count <= n) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        B2A_CUDA(cudaMallocHost(&p, count * sizeof(T)));
        n = count;
    }
};

// Named-tensor lookup over the b2a_tensor table passed across the ABI.
struct TensorTable {
    std::map<std::string, const b2a_tensor*> m;
    TensorTable(const b2a_tensor* t, int n) {
        for (int i = 0; i < n; ++i) m[t[i].name] = &t[i];
    }
    const b2a_tensor* find(const std::string& name) const {
        auto it = m.find(name);
        return it == m.end() ? nullptr : it->second;
    }
    const b2a_tensor& get(const std::string& name) const {
        auto* t = find(name);
        if (!t) throw Error(B2A_ERR_MODEL_NOT_INITIALIZED, "missing tensor: " + name);
        return *t;
    }
    static int64_t numel(const b2a_tensor& t) {
        int64_t n = 1;
        for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
        return n;
    }
    // fp32 copy of a tensor of dtype f32 or bf16
    std::vector<float> f32(const std::string& name, int64_t expect_numel = -1) const {
        const b2a_tensor& t = get(name);
        int64_t n = numel(t);
        if (expect_numel >= 0 && n != expect_numel)
            throw Error(B2A_ERR_MODEL_NOT_INITIALIZED, "bad shape for tensor: " + name);
        std::vector<float> v(n);
        if (t.dtype == B2A_DTYPE_F32) {
            memcpy(v.data(), t.data, n * sizeof(float));
        } else if (t.dtype == B2A_DTYPE_BF16) {
            const uint16_t* s = (const uint16_t*)t.data;
            for (int64_t i = 0; i < n; ++i) {
                uint32_t u = (uint32_t)s[i] << 16;
                memcpy(&v[i], &u, 4);
            }
        } else {
            throw Error(B2A_ERR_MODEL_NOT_INITIALIZED, "bad dtype for tensor: " + name);
        }
        return v;
    }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (PDL): every kernel of the decode step is launched with
// programmaticStreamSerializationAllowed so that its prologue (barrier init, TMEM allocation, and -- for the
// GEMMs -- the TMA prefetch of the first ring-full of WEIGHTS, which never depend on the previous kernel) overlaps
// with the tail of the kernel before it.  Device side: pdl_trigger() lets the next kernel start launching,
// pdl_wait() blocks until the previous kernel has completed and its writes are visible.  B2A_PDL=0 disables it.
bool pdl_enabled();

template <class... KArgs, class... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    B2A_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
    count_launch();
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// L2 prefetch of a byte range, split over `n_workers` threads in 8 KB bulk requests (cp.async.bulk.prefetch.L2: a hint, no
// completion tracking, no shared memory).  Used to keep HBM streaming the NEXT GEMM's weights while kernels that read little
// (norms, attention, epilogue-only phases) run: the decode step is weight-streaming bound and the weights never depend on
// activations.
struct L2Prefetch {
    const void* ptr;
    long long bytes;
};
__device__ __forceinline__ void l2_prefetch(const L2Prefetch& pf, int worker, int n_workers) {
    if (pf.ptr == nullptr) return;
    constexpr long long CH = 8192;
    const char* base = static_cast<const char*>(pf.ptr);
    for (long long off = (long long)worker * CH; off < pf.bytes; off += (long long)n_workers * CH) {
        const unsigned n = (unsigned)(pf.bytes - off < CH ? ((pf.bytes - off) & ~15ll) : CH);
        if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(n) : "memory");
    }
}
#endif

}  // namespace b2a
