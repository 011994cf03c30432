// Host-side context: raw weight store, device buffer registry, error plumbing.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace l2s {

struct HostTensor {
    std::vector<float> f;          // fp32 payload (int64 tensors are stored converted; only num_batches_tracked)
    std::vector<int64_t> shape;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct L2sError : std::runtime_error {
    int code;
    L2sError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define L2S_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::l2s::L2sError(2, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct Context {
    int device = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    std::string err;
    std::map<std::string, HostTensor> w;
    std::map<std::string, DevBuf> bufs;       // named device allocations (packed weights + workspaces)
    std::map<std::string, int64_t> meta;      // small integers produced by packing (sizes, counts)
    int64_t launches = 0;
    int committed = 0;
    cudaStream_t host_stream = nullptr, copy_stream = nullptr;   // l2s_infer_host: compute / clip-copy streams (created on first use)
    cudaEvent_t copy_done[2] = {nullptr, nullptr}, slot_done[2] = {nullptr, nullptr}, video_consumed[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    void* nccl_comm = nullptr; int world = 1, rank = 0;          // data-parallel gradient exchange (l2s_comm_init)
    int pw_min_rows = 16384;                  // streaming 1x1 kernel for GEMMs with at least this many rows (L2S_PW_MIN_ROWS)
#ifdef L2S_DEBUG
    // debug builds only (nvcc -DL2S_DEBUG): environment toggles that swap kernels for bisecting; a release library has ONE path
    bool use_pw = true;                       // L2S_PW=0: tcgen05 GEMM instead of the streaming mma.sync kernel for the trunk's 1x1 convolutions
    bool use_tc = true;                       // L2S_TC=0: exact-fp32 SIMT GEMMs
#else
    static constexpr bool use_pw = true, use_tc = true;
#endif
    // optional stage timing (CUDA events on the caller's stream), enabled by l2s_set_profiling
    bool profiling = false;
    struct Span { cudaEvent_t e0 = nullptr, e1 = nullptr; bool used = false; };
    std::map<std::string, Span> spans;
    void span_begin(const std::string& name, cudaStream_t s) {
        if (!profiling) return;
        Span& sp = spans[name];
        if (!sp.e0) { L2S_CUDA(cudaEventCreate(&sp.e0)); L2S_CUDA(cudaEventCreate(&sp.e1)); }
        L2S_CUDA(cudaEventRecord(sp.e0, s));
    }
    void span_end(const std::string& name, cudaStream_t s) {
        if (!profiling) return;
        Span& sp = spans[name];
        L2S_CUDA(cudaEventRecord(sp.e1, s));
        sp.used = true;
    }

    const HostTensor& W(const std::string& key) const {
        auto it = w.find(key);
        if (it == w.end()) throw L2sError(3, "missing weight: " + key);
        return it->second;
    }
    bool has(const std::string& key) const { return w.count(key) != 0; }

    // (Re)allocate a named device buffer of at least `bytes`; contents are zeroed on (re)allocation.
    void* buf(const std::string& name, size_t bytes) {
        DevBuf& b = bufs[name];
        if (b.bytes < bytes) {
            if (b.p) L2S_CUDA(cudaFree(b.p));
            b.p = nullptr; b.bytes = 0;
            size_t alloc = (bytes + 255) & ~size_t(255);
            L2S_CUDA(cudaMalloc(&b.p, alloc));
            L2S_CUDA(cudaMemset(b.p, 0, alloc));
            // cudaMemset on device memory is asynchronous on the legacy default stream, which does not order against the
            // caller's (possibly non-blocking) streams: without this, a copy enqueued on such a stream right after the
            // allocation can be overtaken by the zero fill (seen once per fresh l2s_infer_host_submit slot)
            L2S_CUDA(cudaDeviceSynchronize());
            b.bytes = alloc;
        }
        return b.p;
    }
    float* fbuf(const std::string& name, size_t floats) { return static_cast<float*>(buf(name, floats * sizeof(float))); }
    float* upload(const std::string& name, const std::vector<float>& v) {
        float* d = fbuf(name, v.size());
        L2S_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return d;
    }
    template <typename T>
    T* upload_raw(const std::string& name, const T* src, size_t count) {
        T* d = static_cast<T*>(buf(name, count * sizeof(T)));
        L2S_CUDA(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
        return d;
    }
    float* dev(const std::string& name) const {
        auto it = bufs.find(name);
        if (it == bufs.end() || !it->second.p) throw L2sError(1, "internal: buffer not packed: " + name + " (call l2s_commit_weights)");
        return static_cast<float*>(it->second.p);
    }
    void free_all() {
        for (auto& kv : bufs)
            if (kv.second.p) cudaFree(kv.second.p);
        bufs.clear();
        for (auto& kv : spans) {
            if (kv.second.e0) cudaEventDestroy(kv.second.e0);
            if (kv.second.e1) cudaEventDestroy(kv.second.e1);
        }
        spans.clear();
        for (int i = 0; i < 2; ++i) {
            if (copy_done[i]) cudaEventDestroy(copy_done[i]);
            if (slot_done[i]) cudaEventDestroy(slot_done[i]);
            if (video_consumed[i]) cudaEventDestroy(video_consumed[i]);
            copy_done[i] = slot_done[i] = video_consumed[i] = nullptr; slot_used[i] = false;
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (host_stream) cudaStreamDestroy(host_stream);
        copy_stream = host_stream = nullptr;
    }
};

}  // namespace l2s
