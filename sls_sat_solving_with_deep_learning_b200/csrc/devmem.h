// devmem.h -- device allocations of the library go through a PRIVATE stream-ordered pool.
//
// The reference-facing call (run_sls_python, rust/src/lib.rs:103-149) allocates and frees its working set on
// every call; with cudaMalloc / cudaFree that is a page-table update per buffer and showed up as 10-350 ms of
// jitter around a 180 ms kernel.  The pool keeps freed blocks mapped (up to the release threshold) so a repeated
// call of the same shape allocates without talking to the OS.  The pool is the library's own (cudaMemPoolCreate):
// the device's default pool and its release threshold, which belong to the host application, are left alone.
//
// Large buffers (GNN activations, HBM-tier chain state: more than 512 MB, up to tens of GB) do NOT go through the pool:
// growing the pool by gigabytes costs ~30 ms per GB, and keeping such blocks cached would take them away from the
// application's own allocator (torch uses cudaMalloc, which cannot see the pool's free memory).  They use plain
// cudaMalloc / cudaFree.  The limit was 32 MB until the chunked forward -> walk pipeline (pipeline.py) showed what that costs
// when two host threads use the library at once: a cudaMalloc / cudaFree pair of a 67 MB table per walk call stalled BOTH
// threads for 0.5-0.8 s in 1-3 of every 20 calls (cudaFree waits for the device and serialises with the other thread's
// allocations); with those blocks in the pool the same run has no such outliers (2300 instead of 1650-1930 instances/s).
//
// Freeing: dfree() does NOT synchronise the device per buffer.  Its contract is "no work that uses the block is in
// flight".  The synchronous entry points of the library end with a wait on their own work, and mark the time in between
// with InFlight (below): a dfree reached while the calling thread's work may still be running -- an error path --
// waits for the device ONCE and clears the mark.  sls_solver_free waits for the solver's last launch itself.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <unordered_set>

namespace slsb {

constexpr int kMaxDevices = 64;
constexpr uint64_t kPoolKeepBytes = 4ull << 30;  // freed memory above this goes back to the driver at the next sync
constexpr size_t kPoolMaxBlock = 512u << 20;      // larger requests bypass the pool (SLS_B200_POOL_MAX_BLOCK_MB)

struct DevMemState {
    std::mutex mu;
    cudaStream_t stream[kMaxDevices] = {};
    cudaMemPool_t pool[kMaxDevices] = {};
    bool ready[kMaxDevices] = {};
    std::unordered_set<void*> direct;  // blocks that came from cudaMalloc
};

inline DevMemState& devmem_state()
{
    static DevMemState st;
    return st;
}

// Stream the pool operations of `device` are ordered on (created on first use; the device must be current).
inline cudaError_t devmem_stream(int device, cudaStream_t* out, cudaMemPool_t* pool_out = nullptr)
{
    DevMemState& st = devmem_state();
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(st.mu);
    if (!st.ready[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&st.pool[device], &props);
        if (e != cudaSuccess) return e;
        uint64_t keep = kPoolKeepBytes;
        if (const char* f = std::getenv("SLS_B200_POOL_KEEP_MB")) keep = (uint64_t)std::strtoull(f, nullptr, 10) << 20;
        e = cudaMemPoolSetAttribute(st.pool[device], cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        e = cudaStreamCreateWithFlags(&st.stream[device], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        // SLS_B200_POOL_RESERVE_MB: grow the pool to this size once, now (one block, allocated and freed: it stays cached up to
        // the release threshold).  The pool otherwise grows on demand, and growing maps new memory -- which waits for the kernels
        // that are running: a host thread that uploads instances while another thread's walk kernel runs (pipeline.py) then
        // stalls for a kernel's length per growth step.
        if (const char* f = std::getenv("SLS_B200_POOL_RESERVE_MB")) {
            const size_t bytes = (size_t)std::strtoull(f, nullptr, 10) << 20;
            void* q = nullptr;
            if (bytes && cudaMallocFromPoolAsync(&q, bytes, st.pool[device], st.stream[device]) == cudaSuccess) {
                cudaFreeAsync(q, st.stream[device]);
                cudaStreamSynchronize(st.stream[device]);
            } else {
                cudaGetLastError();
            }
        }
        st.ready[device] = true;
    }
    *out = st.stream[device];
    if (pool_out) *pool_out = st.pool[device];
    return cudaSuccess;
}

// "work of the calling thread that uses library buffers may still be running": set by the synchronous entry points
// around their launches, cleared once they have waited for them.
inline bool& devmem_inflight()
{
    static thread_local bool f = false;
    return f;
}
struct InFlight {
    InFlight() { devmem_inflight() = true; }
    static void done() { devmem_inflight() = false; }
};

// Allocation on the current device, usable from any stream on return.
template <class T>
inline cudaError_t dmalloc(T** p, size_t bytes)
{
    *p = nullptr;
    static const size_t max_block = [] {
        const char* f = std::getenv("SLS_B200_POOL_MAX_BLOCK_MB");
        return f ? (size_t)std::strtoull(f, nullptr, 10) << 20 : kPoolMaxBlock;
    }();
    if (bytes > max_block) {
        void* q = nullptr;
        const cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) return e;
        DevMemState& st = devmem_state();
        std::lock_guard<std::mutex> lock(st.mu);
        st.direct.insert(q);
        *p = static_cast<T*>(q);
        return cudaSuccess;
    }
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    cudaStream_t st;
    cudaMemPool_t pool;
    e = devmem_stream(device, &st, &pool);
    if (e != cudaSuccess) return e;
    void* q = nullptr;
    e = cudaMallocFromPoolAsync(&q, bytes ? bytes : 1, pool, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    *p = static_cast<T*>(q);
    return cudaSuccess;
}

// The caller guarantees that no work using the block is in flight (see the header comment); a thread that is inside
// an InFlight region waits for the device once.  NULL is ignored.
inline void dfree(void* p)
{
    if (!p) return;
    if (devmem_inflight()) {
        cudaDeviceSynchronize();
        devmem_inflight() = false;
    }
    {
        DevMemState& st = devmem_state();
        std::unique_lock<std::mutex> lock(st.mu);
        if (st.direct.erase(p)) {
            lock.unlock();
            cudaFree(p);  // waits for the device like every cudaFree
            return;
        }
    }
    cudaPointerAttributes attr;
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (attr.device != prev) cudaSetDevice(attr.device);
    cudaStream_t st;
    if (devmem_stream(attr.device, &st) == cudaSuccess) cudaFreeAsync(p, st);
    if (attr.device != prev) cudaSetDevice(prev);
}

}  // namespace slsb
