// Shared plumbing of the C-ABI translation units: error reporting, host/device pointer staging and
// a growable per-device workspace.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace cl {

extern thread_local std::string g_last_error;

inline int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CL_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ::cl::fail(-2, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

inline bool is_device_ptr(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Named, growable device buffers owned by the library: one set per (device, stream), so that calls issued on
// different streams never share intermediates (two solves in flight on one device used to overwrite each other's
// scores / hypotheses / error maps).  A buffer only ever grows; before it is replaced the owning stream is
// synchronised, so no queued kernel can still be reading the old allocation.
class Workspace {
public:
    cudaStream_t stream = nullptr;
    std::mutex mu;                  // held for the duration of a C-ABI call using this workspace
    std::vector<std::array<cudaEvent_t, 4>> timing_log;   // cl_dsac_timing: sample / score / refine boundaries per solve

    cudaError_t get(const std::string& name, size_t bytes, void** out)
    {
        Slot& s = slots_[name];
        if (s.cap < bytes) {
            if (s.ptr) {
                cudaError_t e = cudaStreamSynchronize(stream);
                if (e != cudaSuccess) return e;
                e = cudaFree(s.ptr);
                if (e != cudaSuccess) return e;
                s.ptr = nullptr;
                s.cap = 0;
            }
            size_t cap = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&s.ptr, cap);
            if (e != cudaSuccess) return e;
            s.cap = cap;
        }
        *out = s.ptr;
        return cudaSuccess;
    }
    void drop_timing()
    {
        for (auto& q : timing_log)
            for (cudaEvent_t e : q) cudaEventDestroy(e);
        timing_log.clear();
    }
    void release()
    {
        for (auto& kv : slots_)
            if (kv.second.ptr) cudaFree(kv.second.ptr);
        slots_.clear();
        drop_timing();
    }

private:
    struct Slot {
        void* ptr = nullptr;
        size_t cap = 0;
    };
    std::map<std::string, Slot> slots_;
};

// the workspace of (current device, stream); created on first use
Workspace& workspace_for(cudaStream_t stream);
// frees every workspace of the current device (synchronises the device first)
void release_workspaces();

// Stages host inputs into the workspace and copies host outputs back after the launch.
class Stager {
public:
    Stager(Workspace& ws, cudaStream_t stream) : ws_(ws), stream_(stream) {}

    // Device view of an input buffer (nullptr stays nullptr).
    template <typename T>
    cudaError_t in(const char* name, const T* p, size_t count, const T** dev)
    {
        if (!p || count == 0) { *dev = nullptr; return cudaSuccess; }
        if (is_device_ptr(p)) { *dev = p; return cudaSuccess; }
        void* d;
        cudaError_t e = ws_.get(name, count * sizeof(T), &d);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, stream_);
        *dev = static_cast<const T*>(d);
        return e;
    }

    // Device view of an output buffer; host buffers are filled by finish().  `always` allocates a
    // workspace buffer even when the caller passed nullptr (internal intermediate).
    template <typename T>
    cudaError_t out(const char* name, T* p, size_t count, T** dev, bool always = false)
    {
        if (!p && !always) { *dev = nullptr; return cudaSuccess; }
        if (p && is_device_ptr(p)) { *dev = p; return cudaSuccess; }
        void* d;
        cudaError_t e = ws_.get(name, count * sizeof(T), &d);
        if (e != cudaSuccess) return e;
        *dev = static_cast<T*>(d);
        if (p) pending_.push_back({p, d, count * sizeof(T)});
        return cudaSuccess;
    }

    // Device view of a buffer the kernels accumulate into: host contents are staged in and copied back by finish().
    template <typename T>
    cudaError_t inout(const char* name, T* p, size_t count, T** dev)
    {
        if (!p || count == 0) { *dev = nullptr; return cudaSuccess; }
        if (is_device_ptr(p)) { *dev = p; return cudaSuccess; }
        void* d;
        cudaError_t e = ws_.get(name, count * sizeof(T), &d);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, stream_);
        *dev = static_cast<T*>(d);
        pending_.push_back({p, d, count * sizeof(T)});
        return e;
    }

    // Copies pending host outputs back and synchronises the stream if there were any.
    cudaError_t finish()
    {
        for (auto& c : pending_) {
            cudaError_t e = cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, stream_);
            if (e != cudaSuccess) return e;
        }
        if (!pending_.empty()) return cudaStreamSynchronize(stream_);
        return cudaSuccess;
    }

private:
    struct Copy {
        void* host;
        void* dev;
        size_t bytes;
    };
    Workspace& ws_;
    cudaStream_t stream_;
    std::vector<Copy> pending_;
};

}  // namespace cl
