// api.cu -- the C ABI of include/sls_b200.h on top of the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <new>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/sls_b200.h"
#include "common.h"
#include "devmem.h"
#include "eval.cuh"
#include "init_sliced.cuh"
#include "instance.h"
#include "walk_generic.cuh"
#include "walk_k3.cuh"
#include "walk_k3d.cuh"
#include "walk_k3h.cuh"

using namespace slsb;

// ------------------------------------------------------------------ errors
static thread_local std::string g_error;

static int set_error(int code, const std::string& msg)
{
    g_error = msg;
    return code;
}

extern "C" const char* sls_last_error(void) { return g_error.c_str(); }

// shared with gnn_api.cu
int sls_set_error_internal(int code, const std::string& msg) { return set_error(code, msg); }

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return set_error(SLS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// ------------------------------------------------------------------ device
// (Loading the library changes no process-wide state: CUDA_DEVICE_MAX_CONNECTIONS, which the evaluation driver's waves of
// concurrent launches profit from, is set by that driver -- evaluate.py -- and only if the user has not set it.)

extern "C" int sls_device_count(int* count)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *count = 0;
        return set_error(SLS_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = c;
    return SLS_OK;
}

extern "C" int sls_set_device(int device)
{
    CUDA_TRY(cudaSetDevice(device));
    return SLS_OK;
}

extern "C" int sls_device_pool_reserve(uint64_t bytes)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaStream_t st;
    cudaMemPool_t pool;
    CUDA_TRY(devmem_stream(dev, &st, &pool));
    uint64_t keep = 0;
    CUDA_TRY(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    bytes = std::min<uint64_t>(bytes, keep);  // what lies above the threshold would be unmapped again at the next synchronisation
    uint64_t have = 0;
    CUDA_TRY(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &have));
    if (bytes <= have) return SLS_OK;
    // one block of the missing size, allocated and freed: the pool keeps it mapped and serves later requests out of it
    void* q = nullptr;
    CUDA_TRY(cudaMallocFromPoolAsync(&q, (size_t)(bytes - have), pool, st));
    CUDA_TRY(cudaFreeAsync(q, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SLS_OK;
}

// ---------------------------------------------------------------- instance
namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// The device image of an instance: every array back to back (256-byte aligned segments), built on the host
// without touching CUDA, so that the images of a batch can be built by several threads.
struct InstImage {
    std::vector<char> bytes;
    size_t off[7] = {};
};

void build_image(const HostInstance& h, InstImage& img)
{
    const size_t L = h.lits.size();
    const size_t n_orec = h.k3 ? L + 32 : 0;  // padded: a warp may over-read one full round
    const size_t seg_bytes[7] = {h.row_ptr.size() * 4, L * 4, h.occ_ptr.size() * 4, L * 4, h.has_duplicates ? h.canon.size() * 4 : 0,
                                 h.k3 ? h.m * sizeof(uint4) : 0, n_orec * sizeof(uint4)};
    size_t total = 0;
    for (int i = 0; i < 7; i++) {
        img.off[i] = total;
        total += align_up(std::max<size_t>(seg_bytes[i], 16), 256);
    }
    img.bytes.assign(total, 0);
    char* b = img.bytes.data();
    const void* src[5] = {h.row_ptr.data(), h.lits.data(), h.occ_ptr.data(), h.occ.data(), h.has_duplicates ? h.canon.data() : nullptr};
    for (int i = 0; i < 5; i++)
        if (seg_bytes[i]) std::memcpy(b + img.off[i], src[i], seg_bytes[i]);
    if (h.k3) {
        // clause records {l0,l1,l2,k}; occurrence records {clause, other a, other b, neg of the variable}
        uint4* crec = reinterpret_cast<uint4*>(b + img.off[5]);
        uint4* orec = reinterpret_cast<uint4*>(b + img.off[6]);
        std::vector<uint32_t> fill(h.occ_ptr.begin(), h.occ_ptr.end() - 1);
        for (uint32_t c = 0; c < h.m; c++) {
            const uint32_t s = h.row_ptr[c], k = h.row_ptr[c + 1] - s;
            uint32_t l[3] = {LIT_NONE, LIT_NONE, LIT_NONE};
            for (uint32_t j = 0; j < k; j++) l[j] = h.lits[s + j];
            crec[c] = make_uint4(l[0], l[1], l[2], k);
            for (uint32_t j = 0; j < k; j++) {
                uint32_t others[2] = {h.n << 1, h.n << 1};  // absent: positive literal of the dummy variable n (bit always 0)
                uint32_t t = 0;
                for (uint32_t q = 0; q < k; q++)
                    if (q != j) others[t++] = h.lits[s + q];
                const uint32_t lit = h.lits[s + j];
                orec[fill[lit >> 1]++] = make_uint4(c, others[0], others[1], lit & 1u);
            }
        }
        for (size_t i = L; i < n_orec; i++) orec[i] = make_uint4(0, h.n << 1, h.n << 1, 2);
    }
}

// One allocation and ONE host-to-device copy per instance.  The copy is issued on the memory pool's own non-blocking stream
// and the CALLER waits for that stream (once per instance, or once per group of a batch): an upload then neither waits for
// nor delays kernels of other host threads on the default stream -- pipeline.solve_with_oracle uploads the next chunk while
// the walk of the previous one runs.
int commit_image(sls_instance* inst, const InstImage& img, int dev, cudaStream_t up)
{
    const HostInstance& h = inst->host;
    char* blob = nullptr;
    CUDA_TRY(dmalloc(&blob, img.bytes.size()));
    cudaError_t e = cudaMemcpyAsync(blob, img.bytes.data(), img.bytes.size(), cudaMemcpyHostToDevice, up);
    if (e != cudaSuccess) {
        dfree(blob);
        return set_error(SLS_ERR_CUDA, std::string("cudaMemcpy(instance): ") + cudaGetErrorString(e));
    }
    inst->d_blob = blob;
    inst->blob_bytes = img.bytes.size();
    inst->device = dev;
    DevInst& d = inst->dev;
    d.n = h.n;
    d.m = h.m;
    d.nwords = (h.n + 32) / 32;  // one spare bit: the dummy variable n of the k<=3 path
    d.mwords = (h.m + 31) / 32;
    d.ngroups = (d.mwords + 31) / 32;
    d.off_v = (d.nwords + 7u) & ~7u;  // vbits rows start on a 32-byte sector (tile stores of init_sliced.cuh)
    d.off_l1 = d.off_v + ((d.mwords + 31u) & ~31u);  // vbits zero-padded to whole 32-word groups
    d.l1_pad = std::max(64u, (d.ngroups + 63u) & ~63u);
    d.l2_pad = d.l1_pad > 64 ? ((d.l1_pad / 64 + 63u) & ~63u) : 0u;
    d.off_l2 = d.off_l1 + d.l1_pad;
    d.slab_words = d.off_l2 + d.l2_pad;
    d.uniform_k = h.uniform_k;
    d.max_k = h.max_k;
    d.flags = (h.has_duplicates ? INST_HAS_DUPLICATES : 0u) | (h.k3 ? INST_K3 : 0u);
    d.row_ptr = reinterpret_cast<const uint32_t*>(blob + img.off[0]);
    d.lits = reinterpret_cast<const uint32_t*>(blob + img.off[1]);
    d.occ_ptr = reinterpret_cast<const uint32_t*>(blob + img.off[2]);
    d.occ = reinterpret_cast<const uint32_t*>(blob + img.off[3]);
    d.canon = h.has_duplicates ? reinterpret_cast<const uint32_t*>(blob + img.off[4]) : nullptr;
    d.crec = h.k3 ? reinterpret_cast<const uint4*>(blob + img.off[5]) : nullptr;
    d.orec = h.k3 ? reinterpret_cast<const uint4*>(blob + img.off[6]) : nullptr;
    return SLS_OK;
}

// Upload the instance.  Called lazily by the first compute entry point so that parsing / inspection works on a
// machine without a GPU.
int ensure_on_device(sls_instance* inst)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (inst->d_blob && inst->device == dev) return SLS_OK;
    if (inst->d_blob) {
        // the handle holds ONE device image; solvers keep raw pointers into it, so it cannot move while any exists
        if (inst->solver_refs.load() > 0)
            return set_error(SLS_ERR_ARG, "the instance is resident on device " + std::to_string(inst->device) +
                                              " and in use by a solver there; free the solver or use a second handle for device " +
                                              std::to_string(dev));
        dfree(inst->d_blob);
        inst->d_blob = nullptr;
    }
    InstImage img;
    build_image(inst->host, img);
    cudaStream_t up;
    CUDA_TRY(devmem_stream(dev, &up));
    const int rc = commit_image(inst, img, dev, up);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(up));  // the image is complete before any kernel of any stream may read it
    return SLS_OK;
}

// The same for a batch: the images of the instances that are not resident yet are built by up to 16 host threads,
// then committed one copy each (a batch of config 4 is 1250 instances of ~200 KB).
int ensure_batch_on_device(sls_instance* const* insts, uint32_t n_inst)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::vector<uint32_t> todo;
    std::unordered_set<const sls_instance*> seen;  // the same handle may appear more than once in a batch
    for (uint32_t i = 0; i < n_inst; i++) {
        if (!insts[i]) return set_error(SLS_ERR_ARG, "NULL instance handle in batch");
        if (insts[i]->d_blob && insts[i]->device == dev) continue;
        if (seen.insert(insts[i]).second) todo.push_back(i);
    }
    if (todo.empty()) return SLS_OK;
    const size_t chunk = 64;  // images alive at a time: bounds the host staging memory
    cudaStream_t up;
    CUDA_TRY(devmem_stream(dev, &up));
    std::vector<InstImage> imgs(std::min(chunk, todo.size()));
    for (size_t base = 0; base < todo.size(); base += chunk) {
        const size_t cnt = std::min(chunk, todo.size() - base);
        const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)std::thread::hardware_concurrency(), (size_t)16, cnt}));
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t k; (k = next.fetch_add(1)) < cnt;) build_image(insts[todo[base + k]]->host, imgs[k]);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        for (size_t k = 0; k < cnt; k++) {
            sls_instance* inst = insts[todo[base + k]];
            if (inst->d_blob) {
                if (inst->solver_refs.load() > 0)
                    return set_error(SLS_ERR_ARG, "an instance of the batch is in use by a solver on another device");
                dfree(inst->d_blob);
                inst->d_blob = nullptr;
            }
            int rc = commit_image(inst, imgs[k], dev, up);
            if (rc) return rc;
        }
        CUDA_TRY(cudaStreamSynchronize(up));  // before the staging images are rebuilt, and before anything reads the copies
    }
    return SLS_OK;
}

}  // namespace

int slsb_ensure_batch_on_device(sls_instance* const* insts, uint32_t n_inst) { return ensure_batch_on_device(insts, n_inst); }

namespace {

int finish_handle(sls_instance* inst, const std::string& err, sls_instance** out)
{
    if (!err.empty()) {
        delete inst;
        return set_error(SLS_ERR_PARSE, "parse error: " + err);
    }
    *out = inst;
    return SLS_OK;
}

}  // namespace

extern "C" int sls_instance_from_dimacs_text(const char* text, size_t len, sls_instance** out)
{
    if (!out) return set_error(SLS_ERR_ARG, "out is NULL");
    *out = nullptr;
    sls_instance* inst = new (std::nothrow) sls_instance();
    if (!inst) return set_error(SLS_ERR_NOMEM, "out of memory");
    return finish_handle(inst, parse_dimacs(text, len, inst->host), out);
}

extern "C" int sls_instance_from_dimacs(const char* path, sls_instance** out)
{
    if (!out) return set_error(SLS_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!path) return set_error(SLS_ERR_ARG, "path is NULL");
    FILE* f = std::fopen(path, "rb");
    if (!f) return set_error(SLS_ERR_IO, std::string("failed to read ") + path);
    std::string text;
    bool ok = true;
    if (std::fseek(f, 0, SEEK_END) == 0) {  // regular file: one read of the whole text
        const long size = std::ftell(f);
        std::rewind(f);
        if (size > 0) {
            text.resize(size_t(size));
            text.resize(std::fread(&text[0], 1, size_t(size), f));
        }
    }
    char chunk[1 << 16];  // whatever is left (pipes, files growing under us)
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
    ok = !std::ferror(f);
    std::fclose(f);
    if (!ok) return set_error(SLS_ERR_IO, std::string("failed to read ") + path);
    return sls_instance_from_dimacs_text(text.data(), text.size(), out);
}

extern "C" int sls_instance_from_csr(uint32_t n, uint32_t m, const uint64_t* row_ptr, const int32_t* signed_lits,
                                     sls_instance** out)
{
    if (!out || !row_ptr) return set_error(SLS_ERR_ARG, "NULL argument");
    *out = nullptr;
    const uint64_t L = row_ptr[m];
    if (L > 0xfffffff0ull) return set_error(SLS_ERR_ARG, "formula too large");
    sls_instance* inst = new (std::nothrow) sls_instance();
    if (!inst) return set_error(SLS_ERR_NOMEM, "out of memory");
    HostInstance& h = inst->host;
    h.n = n;
    h.m = m;
    h.row_ptr.resize(size_t(m) + 1);
    for (uint32_t c = 0; c <= m; c++) h.row_ptr[c] = uint32_t(row_ptr[c]);
    h.lits.resize(L);
    for (uint64_t i = 0; i < L; i++) {
        const int64_t s = signed_lits[i];
        const uint64_t v = uint64_t(s < 0 ? -s : s);
        if (s == 0 || v > n) {
            delete inst;
            return set_error(SLS_ERR_ARG, "literal out of range");
        }
        h.lits[i] = uint32_t((v - 1) << 1) | uint32_t(s < 0);
    }
    const std::string err = finalize_instance(h);
    if (!err.empty()) {
        delete inst;
        return set_error(SLS_ERR_ARG, err);
    }
    *out = inst;
    return SLS_OK;
}

extern "C" void sls_instance_free(sls_instance* inst)
{
    if (!inst) return;
    if (inst->d_blob) dfree(inst->d_blob);
    delete inst;
}

extern "C" uint32_t sls_instance_num_vars(const sls_instance* inst) { return inst->host.n; }
extern "C" uint32_t sls_instance_num_clauses(const sls_instance* inst) { return inst->host.m; }
extern "C" uint64_t sls_instance_num_lits(const sls_instance* inst) { return inst->host.lits.size(); }
extern "C" uint32_t sls_instance_max_clause_len(const sls_instance* inst) { return inst->host.max_k; }

extern "C" int sls_instance_get_csr(const sls_instance* inst, uint64_t* row_ptr, int32_t* signed_lits)
{
    const HostInstance& h = inst->host;
    if (row_ptr)
        for (uint32_t c = 0; c <= h.m; c++) row_ptr[c] = h.row_ptr[c];
    if (signed_lits)
        for (size_t i = 0; i < h.lits.size(); i++) {
            const int32_t v = int32_t(h.lits[i] >> 1) + 1;
            signed_lits[i] = (h.lits[i] & 1u) ? -v : v;
        }
    return SLS_OK;
}

// ------------------------------------------------------- clause evaluation
static thread_local double g_last_eval_ms = 0.0;  // device time of the calling thread's last sls_eval_assignments
static thread_local double g_last_eval_only_ms = 0.0;  // ... of its evaluation kernel alone (without the transposition)

extern "C" int sls_eval_assignments(sls_instance* inst, const uint8_t* assignments, uint64_t batch,
                                    uint32_t* violated_bits, uint32_t* energy)
{
    if (!inst || (!assignments && batch && inst->host.n)) return set_error(SLS_ERR_ARG, "NULL argument");
    if (batch == 0) return SLS_OK;
    int rc = ensure_on_device(inst);
    if (rc) return rc;
    const DevInst& I = inst->dev;
    if (I.m == 0) {
        if (energy) std::memset(energy, 0, batch * sizeof(uint32_t));
        return SLS_OK;
    }
    const uint64_t nslices = (batch + SLICE_W - 1) / SLICE_W;
    if (nslices > 65535) return set_error(SLS_ERR_ARG, "batch too large for one call (max 16776960 rows)");
    uint8_t* d_bytes = nullptr;
    uint32_t *d_tq = nullptr, *d_v = nullptr, *d_e = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm = nullptr;
    auto cleanup = [&] {
        dfree(d_bytes);
        dfree(d_tq);
        dfree(d_v);
        dfree(d_e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evm) cudaEventDestroy(evm);
    };
    const size_t n1 = std::max<uint32_t>(I.n, 1);
    // device rows of violated bits are padded to whole tiles so that a tile row is one aligned 32-byte sector
    const uint32_t ntiles = (I.mwords + TILE_WORDS - 1) / TILE_WORDS;
    const uint32_t pitch = ntiles * TILE_WORDS;
#define EV_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            cleanup();                                                                                   \
            return set_error(SLS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
        }                                                                                                \
    } while (0)
    EV_TRY(dmalloc(&d_bytes, batch * n1));
    EV_TRY(dmalloc(&d_tq, nslices * n1 * SLICE_Q * 4));
    EV_TRY(dmalloc(&d_e, batch * 4));
    if (violated_bits) EV_TRY(dmalloc(&d_v, batch * (size_t)pitch * 4));
    EV_TRY(cudaEventCreate(&ev0));
    EV_TRY(cudaEventCreate(&ev1));
    EV_TRY(cudaEventCreate(&evm));
    if (I.n) EV_TRY(cudaMemcpy(d_bytes, assignments, batch * (size_t)I.n, cudaMemcpyHostToDevice));
    InFlight mark;  // an error exit below waits for the device before its buffers are freed
    EV_TRY(cudaEventRecord(ev0));
    EV_TRY(cudaMemsetAsync(d_e, 0, batch * 4));
    if (I.n) {
        const uint64_t total = nslices * I.n;
        transpose_assignments_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_bytes, d_tq, I.n, batch, (uint32_t)nslices);
    }
    EV_TRY(cudaEventRecord(evm));
    {
        constexpr int WARPS = 4;
        // Small tiles-per-CTA: with many more CTAs per slice than the machine holds, the resident CTAs belong to
        // one or two adjacent slices (blockIdx.x runs fastest) and their tq arrays (32 B per variable) stay in L2.
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, inst->device);
        const uint64_t want_ctas = (uint64_t)sms * 32;
        uint32_t tiles_per_block = (uint32_t)std::max<uint64_t>(WARPS, (ntiles + want_ctas - 1) / want_ctas);
        tiles_per_block = (tiles_per_block + WARPS - 1) / WARPS * WARPS;
        if (d_v) {
            dim3 grid((ntiles + tiles_per_block - 1) / tiles_per_block, (unsigned)nslices);
            if (I.uniform_k && I.uniform_k <= 3)  // three pipelined literals, as in the energy-only kernel
                eval_sliced_kernel<WARPS, true, 3><<<grid, WARPS * 32>>>(I, d_tq, d_v, pitch, d_e, batch, tiles_per_block);
            else
                eval_sliced_kernel<WARPS, true, 4><<<grid, WARPS * 32>>>(I, d_tq, d_v, pitch, d_e, batch, tiles_per_block);
        } else {
            // energies only: vertical counters, flushed every 7 tiles -- a warp wants a long run of tiles (14 = two
            // flushes) as long as that still leaves every SM its 12 resident warps
            const uint64_t warps_machine = (uint64_t)sms * 12, ctas_machine = (uint64_t)sms * 3;  // 168 registers: 3 CTAs per SM
            uint32_t per_warp = (uint32_t)std::min<uint64_t>(14, std::max<uint64_t>(1, (uint64_t)ntiles * nslices / warps_machine));
            if (per_warp >= 7) {  // among 7..14 tiles per warp take the run length that wastes least of the last wave
                double best_eff = 0.0;
                for (uint32_t pw = 7; pw <= 14; pw++) {
                    const uint64_t ctas = (uint64_t)((ntiles + pw * WARPS - 1) / (pw * WARPS)) * nslices;
                    const double waves = (double)ctas / (double)ctas_machine, eff = waves / std::ceil(waves);
                    if (eff >= best_eff) best_eff = eff, per_warp = pw;
                }
            }
            tiles_per_block = per_warp * WARPS;
            dim3 grid((ntiles + tiles_per_block - 1) / tiles_per_block, (unsigned)nslices);
            if (I.uniform_k && I.uniform_k <= 3)  // three pipelined literals
                eval_energy_sliced_kernel<WARPS, 3><<<grid, WARPS * 32>>>(I, d_tq, d_e, batch, tiles_per_block);
            else
                eval_energy_sliced_kernel<WARPS, 4><<<grid, WARPS * 32>>>(I, d_tq, d_e, batch, tiles_per_block);
        }
    }
    EV_TRY(cudaGetLastError());
    EV_TRY(cudaEventRecord(ev1));
    EV_TRY(cudaEventSynchronize(ev1));
    InFlight::done();
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        g_last_eval_ms = ms;
        cudaEventElapsedTime(&ms, evm, ev1);
        g_last_eval_only_ms = ms;
    }
    if (energy) EV_TRY(cudaMemcpy(energy, d_e, batch * 4, cudaMemcpyDeviceToHost));
    if (violated_bits)
        EV_TRY(cudaMemcpy2D(violated_bits, (size_t)I.mwords * 4, d_v, (size_t)pitch * 4, (size_t)I.mwords * 4, batch,
                            cudaMemcpyDeviceToHost));
    cleanup();
#undef EV_TRY
    return SLS_OK;
}

extern "C" double sls_eval_last_kernel_ms(void) { return g_last_eval_ms; }
extern "C" double sls_eval_last_eval_kernel_ms(void) { return g_last_eval_only_ms; }

// ------------------------------------------------------- constraint graph (host)
extern "C" int sls_constraint_graph(uint32_t n_variables, uint32_t n_clauses, const int32_t* var_of, const int32_t* clause_of,
                                    uint64_t n_occurrences, int64_t* senders, int64_t* receivers, uint64_t capacity, uint64_t* n_edge)
{
    if (!n_edge || (n_occurrences && (!var_of || !clause_of))) return set_error(SLS_ERR_ARG, "NULL argument");
    for (uint64_t i = 0; i < n_occurrences; i++)
        if (var_of[i] < 0 || (uint32_t)var_of[i] >= n_variables || clause_of[i] < 0 || (uint32_t)clause_of[i] >= n_clauses)
            return set_error(SLS_ERR_ARG, "occurrence out of range");
    const bool fill = senders && receivers && capacity;
    if (!fill) {
        *n_edge = slsb::constraint_graph(n_variables, n_clauses, var_of, clause_of, n_occurrences, nullptr, nullptr);
        return SLS_OK;
    }
    const uint64_t need = slsb::constraint_graph(n_variables, n_clauses, var_of, clause_of, n_occurrences, nullptr, nullptr);
    *n_edge = need;
    if (need > capacity) return set_error(SLS_ERR_ARG, "constraint graph: capacity too small");
    slsb::constraint_graph(n_variables, n_clauses, var_of, clause_of, n_occurrences, senders, receivers);
    const int64_t off = 2 * (int64_t)n_variables;
    for (uint64_t e = 0; e < need; e++) senders[e] += off, receivers[e] += off;
    return SLS_OK;
}

// ------------------------------------------------------------------ solver
struct sls_solver {
    sls_instance* inst = nullptr;
    sls_run_params p{};
    WalkArgs args{};
    bool smem_state = false;
    bool k3 = false;   // walk_k3s: k<=3, state in shared memory
    bool k3h = false;  // walk_k3h: k<=3, state in HBM
    uint32_t wpb = 4;
    size_t smem_bytes = 0;
    int device = 0;
    int num_sms = 0;
    bool weights_set = false;
    bool thr_has_always_true = false;  // some weights_initialize entry is exactly 1 (Bernoulli's ALWAYS_TRUE threshold)
    std::string variant;
    // device buffers
    uint64_t* d_thr = nullptr;
    double* d_w = nullptr;
    uint4* d_crecx = nullptr;
    double* d_ftab = nullptr;    // generic kernel: static flip tables (build_flip_tables_kernel)
    double* d_cscale = nullptr;
    uint32_t* d_gstate = nullptr;
    uint32_t* d_best_bits = nullptr;
    uint32_t* d_flip_log = nullptr;
    uint32_t* d_nf_init = nullptr;   // HBM tier: initial energies from the sliced init
    uint32_t* d_tbits = nullptr;     // HBM tier: transposed assignments of one chunk of slices
    uint32_t tbits_slices = 0;
    uint64_t log_cap = 0;
    // per-run results + the error word: ONE device block (err | steps | flips | best_energy | found | has_best), fetched
    // with one copy (sls_solver_fetch); the pointers below point into it
    char* d_results = nullptr;
    size_t results_bytes = 0, off_steps = 0, off_flips = 0, off_be = 0, off_found = 0, off_hasbest = 0;
    uint8_t* d_found = nullptr;
    uint8_t* d_has_best = nullptr;
    uint64_t* d_steps = nullptr;
    uint32_t* d_best_energy = nullptr;
    uint64_t* d_flips = nullptr;
    uint32_t* d_traj = nullptr;
    int32_t* d_err = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_w = nullptr;  // ev_w: the weight tables are built (legacy default stream)
    cudaStream_t last_stream = nullptr;
    bool launched = false;  // a launch has been enqueued and not yet waited for
    int launches = 0;
    double last_ms = 0.0;
};

extern "C" void sls_solver_free(sls_solver* s)
{
    if (!s) return;
    // one wait for the solver's own work (its last launch and the weight-table build), then every buffer goes back to the pool
    if (s->launched && s->ev1) cudaEventSynchronize(s->ev1);
    if (s->ev_w) cudaEventSynchronize(s->ev_w);
    if (s->inst) s->inst->solver_refs.fetch_sub(1);
    dfree(s->d_thr);
    dfree(s->d_w);
    dfree(s->d_crecx);
    dfree(s->d_ftab);
    dfree(s->d_cscale);
    dfree(s->d_gstate);
    dfree(s->d_best_bits);
    dfree(s->d_flip_log);
    dfree(s->d_nf_init);
    dfree(s->d_tbits);
    dfree(s->d_results);
    dfree(s->d_traj);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_w) cudaEventDestroy(s->ev_w);
    delete s;
}

extern "C" int sls_solver_create(sls_instance* inst, const sls_run_params* params, sls_solver** out)
{
    if (!inst || !params || !out) return set_error(SLS_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (params->algo < 0 || params->algo > 3) return set_error(SLS_ERR_ALGO, "Unknown algorithm type");
    if (params->return_trajectories && params->nsteps + 1 > 0xffffffffull)
        return set_error(SLS_ERR_ARG, "nsteps too large for trajectories");
    int rc = ensure_on_device(inst);
    if (rc) return rc;
    sls_solver* s = new (std::nothrow) sls_solver();
    if (!s) return set_error(SLS_ERR_NOMEM, "out of memory");
    s->inst = inst;
    inst->solver_refs.fetch_add(1);
    s->p = *params;
    const DevInst& I = inst->dev;
    const uint64_t R = params->nruns;
    const size_t R1 = std::max<uint64_t>(R, 1);

#define SC_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            sls_solver_free(s);                                                                          \
            return set_error(SLS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
        }                                                                                                \
    } while (0)
    SC_TRY(cudaGetDevice(&s->device));
    int max_smem_optin = 0, smem_per_sm = 0;
    SC_TRY(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, s->device));
    SC_TRY(cudaDeviceGetAttribute(&max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
    SC_TRY(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, s->device));

    // chain state slab: assignment bits | violated bits | level-1 counters
    const size_t slab_bytes = size_t(I.slab_words) * 4;
    const size_t chain_bytes = slab_bytes + RING_WORDS * 4;  // + the chain's random-word ring (walk_k3.cuh)
    // Tier choice: keep the chain state in shared memory when enough chains fit per SM to cover
    // this call's share of chains (or at least 16 warps); otherwise the state lives in HBM.
    uint32_t wpb = 4;
    while (wpb > 1 && wpb * chain_bytes > (size_t)max_smem_optin) wpb >>= 1;
    bool smem_ok = wpb * chain_bytes <= (size_t)max_smem_optin;
    if (smem_ok) {
        const size_t per_cta = wpb * chain_bytes + 1024;  // 1 KB reserved per CTA
        const uint64_t ctas = std::min<uint64_t>(smem_per_sm / per_cta, 64 / wpb);
        const uint64_t chains_per_sm = ctas * wpb;
        const uint64_t needed = std::min<uint64_t>((R + s->num_sms - 1) / std::max(s->num_sms, 1), 16);
        smem_ok = chains_per_sm >= std::max<uint64_t>(needed, 1);
    }
    // test hooks: SLS_B200_FORCE_STATE=hbm|smem, SLS_B200_FORCE_GENERIC=1
    if (const char* f = std::getenv("SLS_B200_FORCE_STATE")) {
        if (std::strcmp(f, "hbm") == 0) smem_ok = false;
        if (std::strcmp(f, "smem") == 0 && wpb * chain_bytes <= (size_t)max_smem_optin) smem_ok = true;
    }
    s->smem_state = smem_ok;
    s->wpb = smem_ok ? wpb : 4;
    s->smem_bytes = smem_ok ? s->wpb * chain_bytes : 0;
    // the fast kernel: k<=3 instance, chain state in shared memory, at most 64 level-1 counters
    s->k3 = (I.flags & INST_K3) != 0 && s->smem_state && I.l1_pad == 64;
    if (const char* f = std::getenv("SLS_B200_FORCE_GENERIC"))
        if (f[0] == '1') s->k3 = false;
    // ... and its HBM-state sibling: same records, the slab prepared by the sliced init kernels (which need distinct
    // clause contents -- guaranteed for k<=3 instances -- and a non-empty formula), at most 64 x 64 level-1 counters
    s->k3h = (I.flags & INST_K3) != 0 && !s->smem_state && I.n > 0 && I.m > 0 && I.l1_pad <= 64u * 64u &&
             (I.l2_pad == 0 || I.l2_pad == 64);
    if (const char* f = std::getenv("SLS_B200_FORCE_GENERIC"))
        if (f[0] == '1') s->k3h = false;
    if (s->k3h) s->smem_bytes = size_t(s->wpb) * K3H_SMEM_WORDS * 4;  // random-word ring + level-2 counters per chain
    s->variant = s->k3 ? "walk_k3s<smem>" : s->k3h ? "walk_k3h<hbm>" : std::string("walk_generic<") + (s->smem_state ? "smem" : "hbm") + ">";

    SC_TRY(dmalloc(&s->d_thr, std::max<size_t>(I.n, 1) * 8));
    SC_TRY(dmalloc(&s->d_w, std::max<size_t>(I.n, 1) * 8));
    if (s->k3 || s->k3h) SC_TRY(dmalloc(&s->d_crecx, std::max<size_t>(I.m, 1) * CRECX_QUADS * sizeof(uint4)));
    if (!(s->k3 || s->k3h) && params->algo != SLS_ALGO_SCHOENING) {
        SC_TRY(dmalloc(&s->d_ftab, std::max<size_t>(inst->host.lits.size(), 1) * 8));
        SC_TRY(dmalloc(&s->d_cscale, std::max<size_t>(I.m, 1) * 8));
    }
    if (!s->smem_state) SC_TRY(dmalloc(&s->d_gstate, R1 * slab_bytes));
    SC_TRY(dmalloc(&s->d_best_bits, R1 * std::max<size_t>(I.nwords, 1) * 4));
    if (!s->smem_state && !(I.flags & INST_HAS_DUPLICATES) && I.n > 0 && I.m > 0) {
        // chains are initialised 256 at a time by the sliced kernels (init_sliced.cuh); the transposed
        // assignments of one chunk of slices (32 bytes per variable and slice) are kept in a scratch buffer
        // of at most ~1 GB
        const uint64_t nslices = (R + SLICE_W - 1) / SLICE_W;
        s->tbits_slices = (uint32_t)std::min<uint64_t>(nslices, std::max<uint64_t>(1, (1ull << 25) / I.n));
        SC_TRY(dmalloc(&s->d_tbits, size_t(s->tbits_slices) * I.n * SLICE_Q * 4));
        SC_TRY(dmalloc(&s->d_nf_init, R1 * 4));
    }
    if (!s->smem_state) {
        // flip log instead of cloning the assignment on every improvement (lib.rs:294) when the log
        // of the whole walk is no larger than four assignment copies
        const uint64_t per_step = params->algo == SLS_ALGO_MOSER ? std::max<uint32_t>(I.max_k, 1) : 1;
        if (params->nsteps < (1ull << 40) && params->nsteps * per_step + 1 <= 4ull * I.nwords) {
            s->log_cap = params->nsteps * per_step + 1;
            SC_TRY(dmalloc(&s->d_flip_log, R1 * s->log_cap * 4));
        }
    }
    s->off_steps = 256;
    s->off_flips = s->off_steps + align_up(R1 * 8, 256);
    s->off_be = s->off_flips + align_up(R1 * 8, 256);
    s->off_found = s->off_be + align_up(R1 * 4, 256);
    s->off_hasbest = s->off_found + align_up(R1, 256);
    s->results_bytes = s->off_hasbest + align_up(R1, 256);
    SC_TRY(dmalloc(&s->d_results, s->results_bytes));
    s->d_err = reinterpret_cast<int32_t*>(s->d_results);
    s->d_steps = reinterpret_cast<uint64_t*>(s->d_results + s->off_steps);
    s->d_flips = reinterpret_cast<uint64_t*>(s->d_results + s->off_flips);
    s->d_best_energy = reinterpret_cast<uint32_t*>(s->d_results + s->off_be);
    s->d_found = reinterpret_cast<uint8_t*>(s->d_results + s->off_found);
    s->d_has_best = reinterpret_cast<uint8_t*>(s->d_results + s->off_hasbest);
    if (params->return_trajectories) SC_TRY(dmalloc(&s->d_traj, R1 * (params->nsteps + 1) * 4));
    SC_TRY(cudaEventCreateWithFlags(&s->ev0, cudaEventDefault));
    SC_TRY(cudaEventCreateWithFlags(&s->ev1, cudaEventDefault));
    SC_TRY(cudaEventCreateWithFlags(&s->ev_w, cudaEventDisableTiming));
    if (s->smem_bytes) {
        SC_TRY(cudaFuncSetAttribute(walk_generic_kernel_for<true>(params->algo), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)s->smem_bytes));
        SC_TRY(cudaFuncSetAttribute(walk_k3s_kernel_for(params->algo, params->prob_flip_best, params->return_trajectories != 0),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    }
#undef SC_TRY

    WalkArgs& a = s->args;
    a.inst = I;
    a.thr_init = s->d_thr;
    a.w = s->d_w;
    a.crecx = s->d_crecx;
    a.ftab = s->d_ftab;
    a.cscale = s->d_cscale;
    a.algo = params->algo;
    a.mapping = params->pre_compute_mapping;
    a.want_traj = params->return_trajectories;
    a.pfb = params->prob_flip_best;
    a.nsteps = params->nsteps;
    a.nruns = R;
    a.seed = params->seed;
    a.chain_offset = params->chain_offset;
    a.gstate = s->d_gstate;
    a.warps_per_block = s->wpb;
    a.best_bits = s->d_best_bits;
    a.flip_log = s->d_flip_log;
    a.nf_init = s->d_nf_init;
    a.log_cap = s->log_cap;
    a.found = s->d_found;
    a.has_best = s->d_has_best;
    a.steps = s->d_steps;
    a.best_energy = s->d_best_energy;
    a.flips = s->d_flips;
    a.traj = s->d_traj;
    a.err = s->d_err;
    *out = s;
    return SLS_OK;
}

extern "C" int sls_solver_set_seed(sls_solver* s, uint64_t seed, uint64_t chain_offset)
{
    s->p.seed = s->args.seed = seed;
    s->p.chain_offset = s->args.chain_offset = chain_offset;
    return SLS_OK;
}

extern "C" int sls_solver_set_weights(sls_solver* s, const double* w_init, size_t w_init_len, const double* w_resample,
                                      size_t w_resample_len)
{
    const uint32_t n = s->inst->host.n;
    // the reference indexes both vectors by variable (lib.rs:209, :29): a short vector panics
    if (s->p.nruns > 0 && w_init_len < n) return set_error(SLS_ERR_WEIGHTS, "weights_initialize shorter than var_count");
    if (s->p.nruns > 0 && w_resample_len < n && s->p.algo != SLS_ALGO_SCHOENING)
        return set_error(SLS_ERR_WEIGHTS, "weights_resample shorter than var_count");
    if (s->p.nruns == 0) {
        s->weights_set = true;
        return SLS_OK;
    }
    // rand 0.8.5 Bernoulli::new(p): p_int = (p * 2^64) as u64; p == 1 is ALWAYS_TRUE; else p must be in [0,1)
    std::vector<uint64_t> thr(std::max<uint32_t>(n, 1));
    bool any_one = false;
    for (uint32_t i = 0; i < n; i++) {
        const double p = w_init[i];
        if (p >= 0.0 && p < 1.0)
            thr[i] = (uint64_t)(p * 18446744073709551616.0);
        else if (p == 1.0)
            thr[i] = ~0ull, any_one = true;
        else
            return set_error(SLS_ERR_WEIGHTS, "weights_initialize: p is outside range [0.0, 1.0]");
    }
    s->thr_has_always_true = any_one;
    // Nothing here waits for the device: the copies (pageable sources are staged before the call returns) and the table
    // builders are enqueued on the legacy default stream, and the next launch -- on whatever stream -- waits for ev_w.
    // A launch of this solver that is still running reads the old tables: wait for it first (stream order alone does not
    // cover a launch on a non-blocking stream).
    if (s->launched) CUDA_TRY(cudaEventSynchronize(s->ev1));
    CUDA_TRY(cudaMemcpyAsync(s->d_thr, thr.data(), size_t(n) * 8, cudaMemcpyHostToDevice, nullptr));
    if (w_resample_len >= n && n)
        CUDA_TRY(cudaMemcpyAsync(s->d_w, w_resample, size_t(n) * 8, cudaMemcpyHostToDevice, nullptr));
    else if (n)
        CUDA_TRY(cudaMemsetAsync(s->d_w, 0, size_t(n) * 8, nullptr));
    if (s->d_crecx && s->inst->host.m) {
        // 64-byte clause records of the k<=3 kernel: literals + occurrence ranges + resample weights
        const DevInst& I = s->inst->dev;
        build_crecx_kernel<<<(I.m + 255) / 256, 256>>>(I.crec, I.occ_ptr, s->d_w, s->d_crecx, I.m, I.n, s->p.algo);
        CUDA_TRY(cudaGetLastError());
    }
    if (s->d_ftab && s->inst->host.m) {
        const DevInst& I = s->inst->dev;
        build_flip_tables_kernel<<<(I.m + 127) / 128, 128>>>(I, s->d_w, s->p.algo, s->d_ftab, s->d_cscale);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(s->ev_w, nullptr));
    s->weights_set = true;
    return SLS_OK;
}

extern "C" int sls_solver_launch(sls_solver* s, void* cuda_stream)
{
    if (!s->weights_set) return set_error(SLS_ERR_ARG, "sls_solver_set_weights was not called");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    s->last_stream = st;
    s->launches = 0;
    if (s->p.nruns == 0) return SLS_OK;
    CUDA_TRY(cudaStreamWaitEvent(st, s->ev_w, 0));  // the weight tables of sls_solver_set_weights
    CUDA_TRY(cudaMemsetAsync(s->d_err, 0, 4, st));
    const unsigned blocks = (unsigned)((s->p.nruns + s->wpb - 1) / s->wpb);
    const unsigned threads = s->wpb * 32;
    CUDA_TRY(cudaEventRecord(s->ev0, st));
    if (s->d_nf_init) {
        // HBM tier: prepare every chain's slab with the sliced kernels
        const DevInst& I = s->args.inst;
        const uint64_t R = s->p.nruns;
        CUDA_TRY(cudaMemsetAsync(s->d_gstate, 0, R * size_t(I.slab_words) * 4, st));
        const uint32_t nslices = (uint32_t)((R + SLICE_W - 1) / SLICE_W);
        const uint32_t words_per_block = 16;
        constexpr int SCAN_WARPS = 4;
        for (uint32_t s0 = 0; s0 < nslices; s0 += s->tbits_slices) {
            const uint32_t cnt = std::min(s->tbits_slices, nslices - s0);
            dim3 g1((I.nwords + words_per_block - 1) / words_per_block, cnt);
            if (s->thr_has_always_true)
                init_assign_sliced_kernel<true><<<g1, 256, 0, st>>>(s->args, s->d_tbits, s0, words_per_block);
            else
                init_assign_sliced_kernel<false><<<g1, 256, 0, st>>>(s->args, s->d_tbits, s0, words_per_block);
            // one scan launch per slice: a slice's 32 n bytes of transposed assignments then stay in L2 for all of its gathers
            // (ncu, 32 slices in one grid: 8.8 GB of DRAM reads per 8192 chains; slice by slice: 3.8 GB = literals + assignments
            // + 40 MB per slice; the time barely moves -- 6.0 -> 5.7 ms -- the kernel is bound by its transposition, not by DRAM)
            for (uint32_t k = 0; k < cnt; k++) {
                dim3 g2((I.ngroups + SCAN_WARPS - 1) / SCAN_WARPS, 1);
                // (four pipelined literals also for 3-SAT: the instantiation with three -- one gather less per clause, the
                // form the evaluation kernels use -- measured 24.3 ms against 20.9 ms for the init of 8192 chains)
                init_scan_sliced_kernel<SCAN_WARPS, 4><<<g2, SCAN_WARPS * 32, 0, st>>>(s->args, s->d_tbits + (size_t)k * I.n * SLICE_Q, s0 + k);
            }
            s->launches += 1 + (int)cnt;
        }
        init_counters_kernel<<<(unsigned)((R * 32 + 127) / 128), 128, 0, st>>>(s->args);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    const bool use_k3 = s->k3 && walk_k3_supports(s->args);
    if (use_k3) {
        walk_k3s_kernel_for(s->p.algo, s->p.prob_flip_best, s->p.return_trajectories != 0)<<<blocks, threads, s->smem_bytes, st>>>(s->args);
    } else if (s->k3h && walk_k3h_supports(s->args)) {
        walk_k3h_kernel_for(s->p.algo, s->p.prob_flip_best, s->p.return_trajectories != 0)<<<blocks, threads, s->smem_bytes, st>>>(s->args);
    } else {
        if (s->smem_state)
            walk_generic_kernel_for<true>(s->p.algo)<<<blocks, threads, s->smem_bytes, st>>>(s->args);
        else
            walk_generic_kernel_for<false>(s->p.algo)<<<blocks, threads, 0, st>>>(s->args);
    }
    s->launches += 1;
    s->launched = true;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    return SLS_OK;
}

extern "C" int sls_solver_sync(sls_solver* s)
{
    if (s->p.nruns == 0 || !s->launched) return SLS_OK;
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_ms = ms;
    return SLS_OK;
}

extern "C" int sls_solver_last_kernel_ms(sls_solver* s, double* ms)
{
    *ms = s->last_ms;
    return SLS_OK;
}

extern "C" int sls_solver_launch_count(const sls_solver* s) { return s->launches; }
extern "C" const char* sls_solver_variant(const sls_solver* s) { return s->variant.c_str(); }

extern "C" int sls_solver_export_device(sls_solver* s, void* d_found, void* d_steps, void* d_best_energy, void* d_flips,
                                        void* cuda_stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t R = s->p.nruns;
    if (R == 0) return SLS_OK;
    if (d_found) CUDA_TRY(cudaMemcpyAsync(d_found, s->d_found, R, cudaMemcpyDeviceToDevice, st));
    if (d_steps) CUDA_TRY(cudaMemcpyAsync(d_steps, s->d_steps, R * 8, cudaMemcpyDeviceToDevice, st));
    if (d_best_energy) CUDA_TRY(cudaMemcpyAsync(d_best_energy, s->d_best_energy, R * 4, cudaMemcpyDeviceToDevice, st));
    if (d_flips) CUDA_TRY(cudaMemcpyAsync(d_flips, s->d_flips, R * 8, cudaMemcpyDeviceToDevice, st));
    return SLS_OK;
}

extern "C" int sls_solver_fetch(sls_solver* s, sls_run_result* res)
{
    const HostInstance& h = s->inst->host;
    const size_t R = s->p.nruns;
    res->best_energy = h.m;  // lib.rs:320
    res->has_best_assignment = 0;
    res->total_steps = 0;
    res->total_flips = 0;
    res->kernel_ms = s->last_ms;
    if (R == 0) return SLS_OK;
    int rc = sls_solver_sync(s);
    if (rc) return rc;
    res->kernel_ms = s->last_ms;
    // everything the reduction needs in ONE device-to-host copy
    std::vector<char> blk(s->results_bytes);
    CUDA_TRY(cudaMemcpy(blk.data(), s->d_results, s->results_bytes, cudaMemcpyDeviceToHost));
    int32_t err = 0;
    std::memcpy(&err, blk.data(), 4);
    if (err) {
        const char* msg = err == DERR_ALL_ZERO       ? "probsat: AllWeightsZero"
                          : err == DERR_EMPTY_CLAUSE ? "an empty clause was selected"
                          : err == DERR_WEIGHTS      ? "probsat: InvalidWeight"
                                                     : "clause too long for the Moser-Tardos kernel (k > 1024)";
        return set_error(err, msg);
    }
    const uint8_t* found = reinterpret_cast<const uint8_t*>(blk.data() + s->off_found);
    const uint8_t* has_best = reinterpret_cast<const uint8_t*>(blk.data() + s->off_hasbest);
    const uint64_t* steps = reinterpret_cast<const uint64_t*>(blk.data() + s->off_steps);
    const uint64_t* flips = reinterpret_cast<const uint64_t*>(blk.data() + s->off_flips);
    const uint32_t* be = reinterpret_cast<const uint32_t*>(blk.data() + s->off_be);
    // lib.rs:320-334: strict '<' arg-min in run order starting from formula.len()
    uint64_t best = h.m;
    size_t winner = R;
    for (size_t r = 0; r < R; r++) {
        if (be[r] < best) {
            best = be[r];
            winner = r;
        }
        res->total_steps += steps[r];
        res->total_flips += flips[r];
    }
    res->best_energy = best;
    if (winner < R && has_best[winner]) {
        res->has_best_assignment = 1;
        if (res->best_assignment) {
            const uint32_t nwords = s->inst->dev.nwords;
            std::vector<uint32_t> bits(std::max<uint32_t>(nwords, 1));
            CUDA_TRY(cudaMemcpy(bits.data(), s->d_best_bits + winner * nwords, size_t(nwords) * 4,
                                cudaMemcpyDeviceToHost));
            for (uint32_t v = 0; v < h.n; v++) res->best_assignment[v] = (bits[v >> 5] >> (v & 31)) & 1u;
        }
    }
    if (res->found) std::memcpy(res->found, found, R);
    if (res->steps) std::memcpy(res->steps, steps, R * 8);
    if (res->flips_run) std::memcpy(res->flips_run, flips, R * 8);
    if (res->best_energy_run)
        for (size_t r = 0; r < R; r++) res->best_energy_run[r] = be[r];
    if (s->p.return_trajectories) {
        if (res->traj)
            CUDA_TRY(cudaMemcpy(res->traj, s->d_traj, R * (s->p.nsteps + 1) * 4, cudaMemcpyDeviceToHost));
        if (res->traj_len)
            for (size_t r = 0; r < R; r++) res->traj_len[r] = steps[r] + 1;
    }
    return SLS_OK;
}

extern "C" int sls_solver_fetch_trajectories(sls_solver* s, uint32_t* packed, uint64_t capacity_words, uint64_t* offsets)
{
    if (!s->p.return_trajectories) return set_error(SLS_ERR_ARG, "the solver was created without return_trajectories");
    if (!offsets) return set_error(SLS_ERR_ARG, "NULL argument");
    const size_t R = s->p.nruns;
    offsets[0] = 0;
    if (R == 0) return SLS_OK;
    int rc = sls_solver_sync(s);
    if (rc) return rc;
    std::vector<uint64_t> steps(R);
    CUDA_TRY(cudaMemcpy(steps.data(), s->d_steps, R * 8, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < R; r++) offsets[r + 1] = offsets[r] + steps[r] + 1;
    if (!packed) return SLS_OK;
    const uint64_t total = offsets[R];
    if (capacity_words < total) return set_error(SLS_ERR_ARG, "trajectory buffer too small");
    uint64_t* d_off = nullptr;
    uint32_t* d_packed = nullptr;
    CUDA_TRY(dmalloc(&d_off, (R + 1) * 8));
    cudaError_t e = dmalloc(&d_packed, std::max<uint64_t>(total, 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_off, offsets, (R + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        uint64_t longest = 0;
        for (size_t r = 0; r < R; r++) longest = std::max(longest, steps[r] + 1);
        dim3 grid((unsigned)std::min<uint64_t>((longest + 255) / 256, 1024), (unsigned)std::min<uint64_t>(R, 65535));
        pack_trajectories_kernel<<<grid, 256>>>(s->d_traj, s->p.nsteps + 1, d_off, d_packed, R);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(packed, d_packed, total * 4, cudaMemcpyDeviceToHost);
    dfree(d_off);
    dfree(d_packed);
    if (e != cudaSuccess) return set_error(SLS_ERR_CUDA, std::string("fetch_trajectories: ") + cudaGetErrorString(e));
    return SLS_OK;
}

extern "C" int sls_solver_trajectory_stats(sls_solver* s, double* mean, double* median)
{
    if (!s->p.return_trajectories) return set_error(SLS_ERR_ARG, "the solver was created without return_trajectories");
    const size_t R = s->p.nruns;
    const uint64_t n1 = s->p.nsteps + 1;
    if (R == 0) return set_error(SLS_ERR_ARG, "no runs");
    int rc = sls_solver_sync(s);
    if (rc) return rc;
    double *d_mean = nullptr, *d_median = nullptr;
    cudaError_t e = cudaSuccess;
    if (mean) e = dmalloc(&d_mean, n1 * 8);
    if (e == cudaSuccess && median) e = dmalloc(&d_median, n1 * 8);
    if (e == cudaSuccess) {
        const uint64_t warps = n1;
        trajectory_stats_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256>>>(s->d_traj, n1, s->d_steps, (uint32_t)R, n1,
                                                                               s->inst->host.m, d_mean, d_median);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && mean) e = cudaMemcpy(mean, d_mean, n1 * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && median) e = cudaMemcpy(median, d_median, n1 * 8, cudaMemcpyDeviceToHost);
    dfree(d_mean);
    dfree(d_median);
    if (e != cudaSuccess) return set_error(SLS_ERR_CUDA, std::string("trajectory_stats: ") + cudaGetErrorString(e));
    return SLS_OK;
}

extern "C" int sls_run(sls_instance* inst, const double* w_init, size_t w_init_len, const double* w_resample,
                       size_t w_resample_len, const sls_run_params* params, sls_run_result* res)
{
    sls_solver* s = nullptr;
    int rc = sls_solver_create(inst, params, &s);
    if (rc) return rc;
    rc = sls_solver_set_weights(s, w_init, w_init_len, w_resample, w_resample_len);
    if (!rc) rc = sls_solver_launch(s, nullptr);
    if (!rc) rc = sls_solver_fetch(s, res);
    sls_solver_free(s);
    return rc;
}

// ------------------------------------------------------------------ batched solver
// Many instances, the same solver parameters, one launch per kernel variant: replaces the
// per-instance loop of python/src/evaluate_with_given_params.py:108-138 (oracle -> run_sls_python
// for every file, sequentially) for the throughput configs (BASELINE configs 1, 3, 4).
namespace {

struct Pool {
    char* base = nullptr;
    size_t size = 0;
    ~Pool() { dfree(base); }
};

struct ItemPlan {
    bool smem = false, k3 = false;
    uint32_t wpb = 4;
    size_t smem_bytes = 0;
    // element offsets into the pools
    size_t off_n = 0, off_bits = 0, off_run = 0, off_crecx = 0, off_gstate = 0, off_ftab = 0, off_cscale = 0;
};

}  // namespace

extern "C" int sls_run_batch(sls_instance* const* insts, uint32_t n_inst, const double* w_init, const double* w_resample,
                             const uint64_t* var_offsets, const uint64_t* instance_seeds, const sls_run_params* params,
                             sls_run_result* results)
{
    if ((n_inst && (!insts || !results || !var_offsets)) || !params) return set_error(SLS_ERR_ARG, "NULL argument");
    if (params->algo < 0 || params->algo > 3) return set_error(SLS_ERR_ALGO, "Unknown algorithm type");
    if (params->return_trajectories)
        return set_error(SLS_ERR_ARG, "sls_run_batch does not return trajectories; use sls_run per instance for them");
    const uint64_t R = params->nruns;
    for (uint32_t i = 0; i < n_inst; i++)
        if (!insts[i]) return set_error(SLS_ERR_ARG, "NULL instance handle in batch");
    for (uint32_t i = 0; i < n_inst; i++) {
        results[i].best_energy = insts[i]->host.m;
        results[i].has_best_assignment = 0;
        results[i].total_steps = results[i].total_flips = 0;
        results[i].kernel_ms = 0.0;
    }
    if (n_inst == 0 || R == 0) return SLS_OK;
    // SLS_B200_DEBUG_PHASES=1: wall-clock phases of the call on stderr (development aid for pipeline.py's Gantt view)
    static const bool dbg_phases = std::getenv("SLS_B200_DEBUG_PHASES") != nullptr;
    const auto ph0 = std::chrono::steady_clock::now();
    auto ph_ms = [ph0] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ph0).count(); };
    double ph[6] = {0, 0, 0, 0, 0, 0};
    struct PhasePrinter {  // declared before the pools: runs after their frees
        bool on;
        const double* ph;
        decltype(ph_ms)& now;
        uint32_t n;
        ~PhasePrinter()
        {
            if (on)
                std::fprintf(stderr, "sls_run_batch(%u): plan %.1f | weights+pools+H2D %.1f | enqueue %.1f | device wait %.1f | D2H+reduce %.1f | frees %.1f ms\n", n,
                             ph[0], ph[1] - ph[0], ph[2] - ph[1], ph[3] - ph[2], ph[4] - ph[3], now() - ph[4]);
        }
    } phase_printer{dbg_phases, ph, ph_ms, n_inst};

    int device = 0, num_sms = 0, max_smem_optin = 0, smem_per_sm = 0;
    CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));

    // ---- plan: kernel variant and pool offsets per instance
    std::vector<ItemPlan> plan(n_inst);
    size_t tot_n = 0, tot_bits = 0, tot_run = 0, tot_crecx = 0, tot_gstate = 0, tot_ftab = 0, tot_cscale = 0;
    const bool force_generic = [] { const char* f = std::getenv("SLS_B200_FORCE_GENERIC"); return f && f[0] == '1'; }();
    const char* force_state = std::getenv("SLS_B200_FORCE_STATE");
    {
        int rc = ensure_batch_on_device(insts, n_inst);
        if (rc) return rc;
    }
    for (uint32_t i = 0; i < n_inst; i++) {
        const DevInst& I = insts[i]->dev;
        ItemPlan& pl = plan[i];
        const size_t slab_bytes = size_t(I.slab_words) * 4 + RING_WORDS * 4;  // slab + random-word ring
        pl.wpb = 4;
        // in a batch the machine is filled by other instances, so shared memory is used whenever
        // at least 8 chains (2 CTAs) of this instance fit per SM
        pl.smem = pl.wpb * slab_bytes <= (size_t)max_smem_optin && 2 * (pl.wpb * slab_bytes + 1024) <= (size_t)smem_per_sm;
        if (force_state && std::strcmp(force_state, "hbm") == 0) pl.smem = false;
        pl.smem_bytes = pl.smem ? pl.wpb * slab_bytes : 0;
        pl.k3 = (I.flags & INST_K3) != 0 && pl.smem && I.l1_pad == 64 && !force_generic;
        pl.off_n = tot_n, pl.off_bits = tot_bits, pl.off_run = tot_run, pl.off_crecx = tot_crecx, pl.off_gstate = tot_gstate;
        tot_n += std::max<uint32_t>(I.n, 1);
        tot_bits += R * std::max<uint32_t>(I.nwords, 1);
        tot_run += R;
        if (pl.k3) tot_crecx += size_t(I.m) * CRECX_QUADS;
        pl.off_ftab = tot_ftab, pl.off_cscale = tot_cscale;
        if (!pl.k3 && params->algo != SLS_ALGO_SCHOENING) {
            tot_ftab += insts[i]->host.lits.size();
            tot_cscale += I.m;
        }
        if (!pl.smem) tot_gstate += R * size_t(I.slab_words);
    }

    ph[0] = ph_ms();  // plan (+ upload of instances that were not resident)
    // ---- weights: validate and convert on the host, one upload per pool
    std::vector<uint64_t> thr(tot_n, 0);
    std::vector<double> wres(tot_n, 0.0);
    for (uint32_t i = 0; i < n_inst; i++) {
        const uint32_t n = insts[i]->host.n;
        const double* wi = w_init + var_offsets[i];
        const double* wr = w_resample + var_offsets[i];
        if (var_offsets[i + 1] - var_offsets[i] < n) return set_error(SLS_ERR_WEIGHTS, "weights shorter than var_count");
        for (uint32_t v = 0; v < n; v++) {
            const double p = wi[v];
            if (p >= 0.0 && p < 1.0)
                thr[plan[i].off_n + v] = (uint64_t)(p * 18446744073709551616.0);
            else if (p == 1.0)
                thr[plan[i].off_n + v] = ~0ull;
            else
                return set_error(SLS_ERR_WEIGHTS, "weights_initialize: p is outside range [0.0, 1.0]");
            wres[plan[i].off_n + v] = wr[v];
        }
    }

    // per-run results and the per-instance error words share ONE pool (err | steps | flips | best | found | has_best):
    // one device-to-host copy brings all of them back
    Pool p_thr, p_w, p_crecx, p_ftab, p_cscale, p_bits, p_res, p_gstate;
    const size_t ro_steps = align_up(size_t(n_inst) * 4, 256), ro_flips = ro_steps + align_up(tot_run * 8, 256),
                 ro_best = ro_flips + align_up(tot_run * 8, 256), ro_found = ro_best + align_up(tot_run * 4, 256),
                 ro_hasbest = ro_found + align_up(tot_run, 256), res_bytes = ro_hasbest + align_up(tot_run, 256);
    auto alloc = [&](Pool& p, size_t bytes) -> cudaError_t {
        p.size = std::max<size_t>(bytes, 16);
        return dmalloc(&p.base, p.size);
    };
    InFlight mark;  // an error exit below waits for the device before the pools are freed
    CUDA_TRY(alloc(p_thr, tot_n * 8));
    CUDA_TRY(alloc(p_w, tot_n * 8));
    CUDA_TRY(alloc(p_crecx, tot_crecx * sizeof(uint4)));
    CUDA_TRY(alloc(p_ftab, tot_ftab * 8));
    CUDA_TRY(alloc(p_cscale, tot_cscale * 8));
    CUDA_TRY(alloc(p_bits, tot_bits * 4));
    CUDA_TRY(alloc(p_res, res_bytes));
    CUDA_TRY(alloc(p_gstate, tot_gstate * 4));
    CUDA_TRY(cudaMemcpyAsync(p_thr.base, thr.data(), tot_n * 8, cudaMemcpyHostToDevice, nullptr));
    CUDA_TRY(cudaMemcpyAsync(p_w.base, wres.data(), tot_n * 8, cudaMemcpyHostToDevice, nullptr));
    CUDA_TRY(cudaMemsetAsync(p_res.base, 0, ro_steps, nullptr));

    // ---- per-instance kernel arguments, grouped by variant: 0 = k3s, 1 = generic<smem>, 2 = generic<hbm>
    // (3 = k3d: two chains per warp, small k<=3 instances, 4 = four chains per warp for the smallest;
    // test hooks SLS_B200_NO_DUAL=1, SLS_B200_NO_QUAD=1)
    const bool no_dual = [] { const char* f = std::getenv("SLS_B200_NO_DUAL"); return f && f[0] == '1'; }();
    const bool no_quad = [] { const char* f = std::getenv("SLS_B200_NO_QUAD"); return f && f[0] == '1'; }();
    constexpr int NG = 5;
    std::vector<WalkArgs> items[NG];
    std::vector<uint32_t> cta_begin[NG], blk_begin[NG], owner[NG];
    size_t smem_max[NG] = {0, 0, 0, 0, 0};
    // Launch order inside a kernel: instances with the highest clause / variable ratio first.  Their chains are the ones
    // that run the whole step budget; started last, a single such CTA would keep the launch alive long after the rest of
    // the machine has drained.  (Outputs are addressed per instance, so the order is free.)
    std::vector<uint32_t> order(n_inst);
    for (uint32_t i = 0; i < n_inst; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        const HostInstance &hx = insts[x]->host, &hy = insts[y]->host;
        return (uint64_t)hx.m * std::max<uint32_t>(hy.n, 1) > (uint64_t)hy.m * std::max<uint32_t>(hx.n, 1);
    });
    for (uint32_t oi = 0; oi < n_inst; oi++) {
        const uint32_t i = order[oi];
        const DevInst& I = insts[i]->dev;
        const ItemPlan& pl = plan[i];
        WalkArgs a{};
        a.inst = I;
        a.thr_init = reinterpret_cast<uint64_t*>(p_thr.base) + pl.off_n;
        a.w = reinterpret_cast<double*>(p_w.base) + pl.off_n;
        a.crecx = pl.k3 ? reinterpret_cast<uint4*>(p_crecx.base) + pl.off_crecx : nullptr;
        a.algo = params->algo;
        a.mapping = params->pre_compute_mapping;
        a.want_traj = 0;
        a.pfb = params->prob_flip_best;
        a.nsteps = params->nsteps;
        a.nruns = R;
        a.seed = instance_seeds ? instance_seeds[i] : params->seed + i;  // identical to sls_run(inst_i, that seed)
        a.chain_offset = params->chain_offset;
        a.gstate = pl.smem ? nullptr : reinterpret_cast<uint32_t*>(p_gstate.base) + pl.off_gstate;
        a.warps_per_block = pl.wpb;
        a.best_bits = reinterpret_cast<uint32_t*>(p_bits.base) + pl.off_bits;
        a.found = reinterpret_cast<uint8_t*>(p_res.base + ro_found) + pl.off_run;
        a.has_best = reinterpret_cast<uint8_t*>(p_res.base + ro_hasbest) + pl.off_run;
        a.steps = reinterpret_cast<uint64_t*>(p_res.base + ro_steps) + pl.off_run;
        a.best_energy = reinterpret_cast<uint32_t*>(p_res.base + ro_best) + pl.off_run;
        a.flips = reinterpret_cast<uint64_t*>(p_res.base + ro_flips) + pl.off_run;
        a.traj = nullptr;
        a.err = reinterpret_cast<int32_t*>(p_res.base) + i;
        if (!pl.k3 && params->algo != SLS_ALGO_SCHOENING) {
            a.ftab = reinterpret_cast<double*>(p_ftab.base) + pl.off_ftab;
            a.cscale = reinterpret_cast<double*>(p_cscale.base) + pl.off_cscale;
        }
        // small k<=3 instances run two chains per warp (walk_k3d.cuh): 4 warps = 8 chains per CTA
        const bool dual = pl.k3 && !no_dual && walk_k3d_supports(a);
        const bool quad = dual && !no_quad && walk_k3q_supports(a);
        uint32_t chains_per_cta = pl.wpb;
        size_t smem_bytes = pl.smem_bytes;
        if (dual) {
            a.warps_per_block = 4;
            chains_per_cta = quad ? 16 : 8;
            smem_bytes = size_t(chains_per_cta) * (size_t(I.slab_words) + HRING_WORDS) * 4;
        }
        const int g = quad ? 4 : dual ? 3 : pl.k3 ? 0 : pl.smem ? 1 : 2;
        if (cta_begin[g].empty()) cta_begin[g].push_back(0), blk_begin[g].push_back(0);
        // per-weights tables (clause records / flip tables): built by one launch per group, below
        blk_begin[g].push_back(blk_begin[g].back() + ((a.crecx || a.ftab) ? (I.m + 255) / 256 : 0));
        cta_begin[g].push_back(cta_begin[g].back() + (uint32_t)((R + chains_per_cta - 1) / chains_per_cta));
        items[g].push_back(a);
        owner[g].push_back(i);
        smem_max[g] = std::max(smem_max[g], smem_bytes);
    }
    CUDA_TRY(cudaGetLastError());
    ph[1] = ph_ms();  // weights, pools, H2D, items

    struct Events {  // destroyed on every exit path
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        ~Events()
        {
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
        }
    } evs;
    CUDA_TRY(cudaEventCreate(&evs.ev0));
    CUDA_TRY(cudaEventCreate(&evs.ev1));
    cudaEvent_t ev0 = evs.ev0, ev1 = evs.ev1;
    Pool d_items[NG], d_cta[NG];
    for (int g = 0; g < NG; g++) {
        if (items[g].empty()) continue;
        const size_t nb = cta_begin[g].size();  // = blk_begin[g].size(): both tables in one allocation and one copy
        cta_begin[g].insert(cta_begin[g].end(), blk_begin[g].begin(), blk_begin[g].end());
        CUDA_TRY(alloc(d_items[g], items[g].size() * sizeof(WalkArgs)));
        CUDA_TRY(alloc(d_cta[g], 2 * nb * 4));
        CUDA_TRY(cudaMemcpyAsync(d_items[g].base, items[g].data(), items[g].size() * sizeof(WalkArgs), cudaMemcpyHostToDevice, nullptr));
        CUDA_TRY(cudaMemcpyAsync(d_cta[g].base, cta_begin[g].data(), 2 * nb * 4, cudaMemcpyHostToDevice, nullptr));
        cta_begin[g].resize(nb);
        if (blk_begin[g].back())
            build_tables_batch_kernel<<<blk_begin[g].back(), 256>>>(reinterpret_cast<const WalkArgs*>(d_items[g].base),
                                                                    reinterpret_cast<const uint32_t*>(d_cta[g].base) + nb, (uint32_t)items[g].size());
    }
    const WalkK3BatchKernel k3_batch = walk_k3s_batch_kernel_for(params->algo, params->prob_flip_best);
    if (smem_max[0]) CUDA_TRY(cudaFuncSetAttribute(k3_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max[0]));
    if (smem_max[1])
        CUDA_TRY(cudaFuncSetAttribute(walk_generic_batch_kernel_for<true>(params->algo), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max[1]));
    const WalkK3BatchKernel k3d_batch = walk_k3d_batch_kernel_for(params->algo, 16), k3q_batch = walk_k3d_batch_kernel_for(params->algo, 8);
    if (smem_max[3]) CUDA_TRY(cudaFuncSetAttribute(k3d_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max[3]));
    if (smem_max[4]) CUDA_TRY(cudaFuncSetAttribute(k3q_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max[4]));
    CUDA_TRY(cudaEventRecord(ev0));
    for (int g = 0; g < NG; g++) {
        if (items[g].empty()) continue;
        const WalkArgs* di = reinterpret_cast<const WalkArgs*>(d_items[g].base);
        const uint32_t* dc = reinterpret_cast<const uint32_t*>(d_cta[g].base);
        const uint32_t ni = (uint32_t)items[g].size(), blocks = cta_begin[g].back();
        if (g == 4)
            k3q_batch<<<blocks, 128, smem_max[4]>>>(di, dc, ni);
        else if (g == 3)
            k3d_batch<<<blocks, 128, smem_max[3]>>>(di, dc, ni);
        else if (g == 0)
            k3_batch<<<blocks, 128, smem_max[0]>>>(di, dc, ni);
        else if (g == 1)
            walk_generic_batch_kernel_for<true>(params->algo)<<<blocks, 128, smem_max[1]>>>(di, dc, ni);
        else
            walk_generic_batch_kernel_for<false>(params->algo)<<<blocks, 128, 0>>>(di, dc, ni);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1));
    ph[2] = ph_ms();  // tables + launches enqueued
    CUDA_TRY(cudaEventSynchronize(ev1));
    ph[3] = ph_ms();  // device done
    InFlight::done();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);

    // ---- results: one pooled D2H, per-instance reduction (lib.rs:320-334)
    std::vector<char> blk(res_bytes);
    CUDA_TRY(cudaMemcpy(blk.data(), p_res.base, res_bytes, cudaMemcpyDeviceToHost));
    const int32_t* err = reinterpret_cast<const int32_t*>(blk.data());
    for (uint32_t i = 0; i < n_inst; i++)
        if (err[i]) {
            const char* msg = err[i] == DERR_ALL_ZERO       ? "probsat: AllWeightsZero"
                              : err[i] == DERR_EMPTY_CLAUSE ? "an empty clause was selected"
                              : err[i] == DERR_WEIGHTS      ? "probsat: InvalidWeight"
                                                            : "clause too long for the Moser-Tardos kernel (k > 1024)";
            return set_error(err[i], std::string(msg) + " (instance " + std::to_string(i) + ")");
        }
    const uint8_t* found = reinterpret_cast<const uint8_t*>(blk.data() + ro_found);
    const uint8_t* has_best = reinterpret_cast<const uint8_t*>(blk.data() + ro_hasbest);
    const uint64_t* steps = reinterpret_cast<const uint64_t*>(blk.data() + ro_steps);
    const uint64_t* flips = reinterpret_cast<const uint64_t*>(blk.data() + ro_flips);
    const uint32_t* be = reinterpret_cast<const uint32_t*>(blk.data() + ro_best);
    // best assignments: one copy of the whole pool when it is small (a batch of small instances would otherwise pay
    // one synchronous copy per instance), else only the winning chain's words
    std::vector<uint32_t> bits, all_bits;
    const bool bits_at_once = tot_bits * 4 <= (64u << 20);
    if (bits_at_once) {
        all_bits.resize(std::max<size_t>(tot_bits, 1));
        CUDA_TRY(cudaMemcpy(all_bits.data(), p_bits.base, tot_bits * 4, cudaMemcpyDeviceToHost));
    }
    for (uint32_t i = 0; i < n_inst; i++) {
        const HostInstance& h = insts[i]->host;
        const ItemPlan& pl = plan[i];
        sls_run_result& r = results[i];
        r.kernel_ms = ms;
        uint64_t best = h.m;
        size_t winner = R;
        for (size_t q = 0; q < R; q++) {
            const size_t k = pl.off_run + q;
            if (be[k] < best) {
                best = be[k];
                winner = q;
            }
            r.total_steps += steps[k];
            r.total_flips += flips[k];
        }
        r.best_energy = best;
        if (winner < R && has_best[pl.off_run + winner]) {
            r.has_best_assignment = 1;
            if (r.best_assignment) {
                const uint32_t nwords = insts[i]->dev.nwords;
                const uint32_t* wb = all_bits.data() + pl.off_bits + winner * nwords;
                if (!bits_at_once) {
                    bits.resize(std::max<uint32_t>(nwords, 1));
                    CUDA_TRY(cudaMemcpy(bits.data(), reinterpret_cast<uint32_t*>(p_bits.base) + pl.off_bits + winner * nwords,
                                        size_t(nwords) * 4, cudaMemcpyDeviceToHost));
                    wb = bits.data();
                }
                for (uint32_t v = 0; v < h.n; v++) r.best_assignment[v] = (wb[v >> 5] >> (v & 31)) & 1u;
            }
        }
        if (r.found) std::memcpy(r.found, found + pl.off_run, R);
        if (r.steps) std::memcpy(r.steps, steps + pl.off_run, R * 8);
        if (r.flips_run) std::memcpy(r.flips_run, flips + pl.off_run, R * 8);
        if (r.best_energy_run)
            for (size_t q = 0; q < R; q++) r.best_energy_run[q] = be[pl.off_run + q];
    }
    ph[4] = ph_ms();
    return SLS_OK;
}
