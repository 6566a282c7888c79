// ubench.cu -- B200 micro-benchmarks behind the roofline denominators that MEASURED_PEAKS.json does not hold
// ("L2 peak must be measured on the box") and behind the latency model of the walk kernels.
//   library (what bench.py loads so that every roofline denominator is measured IN THE RUN, on the bench box):
//           nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/_build/libubench.so tools/ubench.cu
//           extern "C" int slsub_measure(char* json, size_t cap, int with_hbm)   -> 0 / CUDA error code
//   stand-alone: add -DUBENCH_MAIN, run tools/_build/ubench (prints the same JSON object)
// Bandwidth legs (all SMs, working set resident in the 126 MB L2 unless stated):
//   l2_stream      coalesced 16-byte loads over a 32 MB buffer, many passes
//   l2_sector32    random 32-byte sectors (one 256-bit load per lane) from a 32 MB table: the access of the
//                  bit-sliced clause evaluation (eval.cuh: one sector = the values of a variable in 256 assignments)
//   l2_rec64       random 64-byte records, four broadcast 16-byte loads per warp (walk_k3: clause record)
//   l2_rec16x12    random runs of 12 consecutive 16-byte records, one record per lane (walk_k3: occurrence records)
//   hbm_sector32   random 32-byte sectors from a 4 GB table (HBM tier of the walk, config 5)
// Latency legs (one warp on one SM, dependent chain, cycles per link):
//   lat_l2_hit, lat_lds, lat_shfl_up_add, lat_shfl_idx, lat_vote_popc, lat_redux_or, lat_redux_add, lat_atoms_ret
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <stdexcept>
#include <string>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            throw std::runtime_error(cudaGetErrorString(e_));                                   \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__global__ void stream_kernel(const uint4* __restrict__ buf, size_t n_quads, int passes, uint32_t* sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}

// one random 32-byte sector per lane per iteration, UNROLL independent loads in flight
template <int UNROLL>
__global__ void sector32_kernel(const uint4* __restrict__ buf, uint32_t n_sectors_mask, int iters, uint32_t* sink)
{
    uint32_t acc = 0;
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (int it = 0; it < iters; it++) {
        uint32_t v[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            s = s * 1664525u + 1013904223u;
            const uint4* p = buf + 2 * (size_t)(mix(s) & n_sectors_mask);
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]), "=r"(v[u][4]), "=r"(v[u][5]), "=r"(v[u][6]), "=r"(v[u][7])
                         : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc ^= v[u][j];
    }
    if (acc == 0x12345678u) *sink = acc;
}

// the walk's two accesses: a 64-byte record read by all lanes at one address, then `run` consecutive 16-byte records
__global__ void rec64_kernel(const uint4* __restrict__ buf, uint32_t n_rec_mask, int iters, uint32_t* sink)
{
    uint32_t acc = 0;
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / 32 + 1u);  // per warp
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const uint4* p = buf + 4 * (size_t)((mix(s) ^ acc) & n_rec_mask);  // dependent on the previous record, like a chain
        const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
        acc = a.x ^ b.y ^ c.z ^ d.w;  // the table holds zeros: the dependence is real, the value does not disturb the index
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void rec16_run_kernel(const uint4* __restrict__ buf, uint32_t n_rec_mask, int run, int iters, uint32_t* sink)
{
    uint32_t acc = 0;
    const int lane = threadIdx.x & 31;
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) / 32 + 1u);
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const size_t base = (mix(s) ^ acc) & n_rec_mask;
        const uint4 a = __ldg(buf + base + min(lane, run - 1));
        acc = __shfl_sync(0xffffffffu, a.x ^ a.w, 0);
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ---- latency chains: one warp, `iters` dependent links, clock64 around
enum { LAT_L2 = 0, LAT_LDS, LAT_SHFL_UP_ADD, LAT_SHFL_IDX, LAT_VOTE_POPC, LAT_REDUX_OR, LAT_REDUX_ADD, LAT_ATOMS, LAT_MATCH, LAT_N };

template <int LEG>
__global__ void latency_kernel(const uint32_t* __restrict__ chase, uint32_t n_mask, int iters, long long* cycles, uint32_t* sink)
{
    __shared__ uint32_t sm[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    uint32_t x = lane;
    if (LEG == LAT_L2)  // warm the pointer-chase ring into L2
        for (int i = 0; i < 8192; i++) x = __ldcg(chase + (x & n_mask));
    const long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < iters; i++) {
        if (LEG == LAT_L2) x = __ldcg(chase + (x & n_mask));
        if (LEG == LAT_LDS) x = sm[x & 1023];
        if (LEG == LAT_SHFL_UP_ADD) x += __shfl_up_sync(0xffffffffu, x, 1);
        if (LEG == LAT_SHFL_IDX) x = __shfl_sync(0xffffffffu, x, x & 31) + 1;
        if (LEG == LAT_VOTE_POPC) x = __popc(__ballot_sync(0xffffffffu, (x + lane) & 1)) + lane;
        if (LEG == LAT_REDUX_OR) x = __reduce_or_sync(0xffffffffu, x) + lane;
        if (LEG == LAT_REDUX_ADD) x = __reduce_add_sync(0xffffffffu, x & 0xff) + lane;
        if (LEG == LAT_ATOMS) x = atomicAdd(&sm[(x + lane) & 1023], 1u);
        if (LEG == LAT_MATCH) x = __match_any_sync(0xffffffffu, x & 7) + lane;
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[LEG] = t1 - t0;
    if (x == 0x12345678u) *sink = x;
}

template <class F>
static float time_ms(F&& launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    return best;
}

static std::string measure(bool with_hbm)
{
    std::string out;
    char tmp[512];
#define P(...) do { snprintf(tmp, sizeof tmp, __VA_ARGS__); out += tmp; } while (0)
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int sms = 0, clock_khz = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
    uint32_t* sink;
    CK(cudaMalloc(&sink, 4));
    const size_t small_bytes = 32u << 20;
    uint4* small;
    CK(cudaMalloc(&small, small_bytes));
    CK(cudaMemset(small, 0, small_bytes));
    P("{\"sms\": %d, \"sm_clock_mhz\": %.0f", sms, clock_khz / 1000.0);

    {  // L2 streaming
        const int passes = 64, blocks = sms * 8, threads = 256;
        const float ms = time_ms([&] { stream_kernel<<<blocks, threads>>>(small, small_bytes / 16, passes, sink); });
        P(", \"l2_stream_gbs\": %.1f", passes * (double)small_bytes / ms * 1e-6);
        const float ms2 = time_ms([&] { stream_kernel<<<blocks, threads>>>(small, small_bytes / 32, passes, sink); });
        P(", \"l2_stream_16mb_gbs\": %.1f", passes * (double)small_bytes / 2 / ms2 * 1e-6);
    }
    {  // L2 random 32-byte sectors
        const int iters = 256, blocks = sms * 8, threads = 256;
        const uint32_t mask = (uint32_t)(small_bytes / 32 - 1);
        const float ms4 = time_ms([&] { sector32_kernel<4><<<blocks, threads>>>(small, mask, iters, sink); });
        const float ms8 = time_ms([&] { sector32_kernel<8><<<blocks, threads>>>(small, mask, iters / 2, sink); });
        const double bytes = (double)blocks * threads * iters * 4 * 32;
        P(", \"l2_sector32_gbs_4_in_flight\": %.1f, \"l2_sector32_gbs_8_in_flight\": %.1f", bytes / ms4 * 1e-6, bytes / ms8 * 1e-6);
    }
    {  // the walk's record accesses, 28 warps per SM like the C2 launch, dependent per warp
        const int iters = 20000, blocks = sms * 7, threads = 128;
        const uint32_t mask64 = (uint32_t)(small_bytes / 64 - 1), mask16 = (uint32_t)(small_bytes / 16 / 2 - 1);
        const float msA = time_ms([&] { rec64_kernel<<<blocks, threads>>>(small, mask64, iters, sink); }, 3);
        const float msB = time_ms([&] { rec16_run_kernel<<<blocks, threads>>>(small, mask16, 12, iters, sink); }, 3);
        const double warps = (double)blocks * threads / 32;
        P(", \"l2_rec64_dependent\": {\"ns_per_link\": %.1f, \"gbs\": %.1f}", msA * 1e6 / iters, warps * iters * 64 / msA * 1e-6);
        P(", \"l2_rec16x12_dependent\": {\"ns_per_link\": %.1f, \"gbs\": %.1f}", msB * 1e6 / iters, warps * iters * 12 * 16 / msB * 1e-6);
    }
    if (with_hbm) {  // HBM random sectors
        const size_t big_bytes = 4ull << 30;
        uint4* big;
        CK(cudaMalloc(&big, big_bytes));
        CK(cudaMemset(big, 0, big_bytes));
        const int iters = 64, blocks = sms * 8, threads = 256;
        const uint32_t mask = (uint32_t)(big_bytes / 32 - 1);
        const float ms = time_ms([&] { sector32_kernel<8><<<blocks, threads>>>(big, mask, iters, sink); }, 3);
        P(", \"hbm_sector32_gbs\": %.1f", (double)blocks * threads * iters * 8 * 32 / ms * 1e-6);
        CK(cudaFree(big));
    }
    {  // latencies
        const uint32_t n = 1u << 20;  // 4 MB ring: L2, not L1
        std::vector<uint32_t> h(n);
        for (uint32_t i = 0; i < n; i++) h[i] = (uint32_t)(((uint64_t)i * 2654435761u + 12345u) & (n - 1));
        uint32_t* chase;
        CK(cudaMalloc(&chase, n * 4));
        CK(cudaMemcpy(chase, h.data(), n * 4, cudaMemcpyHostToDevice));
        long long* cyc;
        CK(cudaMalloc(&cyc, LAT_N * 8));
        const int iters = 4096;
        latency_kernel<LAT_L2><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_LDS><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_SHFL_UP_ADD><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_SHFL_IDX><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_VOTE_POPC><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_REDUX_OR><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_REDUX_ADD><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_ATOMS><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        latency_kernel<LAT_MATCH><<<1, 32>>>(chase, n - 1, iters, cyc, sink);
        CK(cudaDeviceSynchronize());
        long long hc[LAT_N];
        CK(cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
        const char* names[LAT_N] = {"l2_hit", "lds", "shfl_up_add", "shfl_idx_add", "vote_popc_add", "redux_or_add", "redux_add_add", "atoms_ret", "match_any_add"};
        P(", \"latency_cycles\": {");
        for (int i = 0, printed = 0; i < LAT_N; i++)
            if (hc[i] > 0) P("%s\"%s\": %.1f", printed++ ? ", " : "", names[i], (double)hc[i] / iters);  // (a leg the compiler folded away reports 0)
        P("}");
        CK(cudaFree(chase));
        CK(cudaFree(cyc));
    }
    P("}");
#undef P
    CK(cudaFree(small));
    CK(cudaFree(sink));
    return out;
}

extern "C" int slsub_measure(char* json, size_t cap, int with_hbm)
{
    try {
        const std::string s = measure(with_hbm != 0);
        if (s.size() + 1 > cap) return -1;
        std::memcpy(json, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::exception&) {
        return (int)cudaGetLastError() ? (int)cudaGetLastError() : 1;
    }
}

#ifdef UBENCH_MAIN
int main()
{
    static char buf[1 << 14];
    const int rc = slsub_measure(buf, sizeof buf, 1);
    if (rc) return 1;
    puts(buf);
    return 0;
}
#endif
