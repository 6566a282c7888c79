// gnn_api.cu -- C ABI of the GNN oracle forward (include/sls_b200.h, section "oracle forward").
//
// Host side of the reference's oracle call path:
//   network.apply(params, graph)                 python/src/evaluate_with_given_params.py:116
//   LCG.get_graph + get_problem_from_cnf          python/src/sat_representations.py:607-666,
//                                                 python/src/sat_instances.py:50-73  (sls_gnn_forward_lcg)
//   LCG.get_model_probabilities                   python/src/sat_representations.py:775-780
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sls_b200.h"
#include "devmem.h"
#include "gnn_kernels.cuh"
#include "instance.h"

int sls_set_error_internal(int code, const std::string& msg);  // api.cu

namespace {

using namespace gnn;
using slsb::dfree;
using slsb::dmalloc;

int round_up(int x, int a) { return (x + a - 1) / a * a; }

// round-to-nearest (ties away, like cvt.rna.tf32.f32) to a 10-bit mantissa
float round_tf32_host(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}

struct DevMlp {
    int k1 = 0, h1 = 0, h2 = 0, k1_slabs = 0, k2_slabs = 0, n1 = 0, n2 = 0, k1_slabs_bf = 0, k2_slabs_bf = 0;
    uint16_t *w1_b0 = nullptr, *w1_b1 = nullptr, *w2_b0 = nullptr, *w2_b1 = nullptr;  // two bf16 pieces, 64 k-values per slab row
    uint16_t *w1_h0 = nullptr, *w1_h1 = nullptr, *w2_h0 = nullptr, *w2_h1 = nullptr;  // two fp16 pieces, same images
    float *w1_hi = nullptr, *w1_lo = nullptr, *b1 = nullptr, *w2_hi = nullptr, *w2_lo = nullptr, *b2 = nullptr, *ln_scale = nullptr,
          *ln_offset = nullptr;
};

#define G_TRY(expr)                                                                                          \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return sls_set_error_internal(SLS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// round to nearest even to bfloat16 (the top 16 bits of an fp32)
uint16_t bf16_bits_host(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
// round to nearest even to IEEE half, saturating at the largest finite value (what cvt.rn.satfinite.f16x2.f32 does on the device)
uint16_t f16_bits_host(float x)
{
    if (x > 65504.f) x = 65504.f;
    if (x < -65504.f) x = -65504.f;
    return __half_as_ushort(__float2half_rn(x));
}
float f16_value_host(uint16_t b) { return __half2float(__ushort_as_half(b)); }
float bf16_value_host(uint16_t b)
{
    const uint32_t u = (uint32_t)b << 16;
    float x;
    std::memcpy(&x, &u, 4);
    return x;
}

int upload_u16(const std::vector<uint16_t>& h, uint16_t** d)
{
    G_TRY(dmalloc(d, std::max<size_t>(h.size(), 1) * sizeof(uint16_t)));
    if (!h.empty()) G_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return SLS_OK;
}

int upload(const std::vector<float>& h, float** d)
{
    G_TRY(dmalloc(d, std::max<size_t>(h.size(), 1) * sizeof(float)));
    if (!h.empty()) G_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return SLS_OK;
}

}  // namespace

struct sls_gnn_model {
    int num_steps = 0, embed = 0, h1 = 0, h2 = 0;
    float *edge_embed_w = nullptr, *edge_embed_b = nullptr, *node_embed_w = nullptr, *node_embed_b = nullptr;
    std::vector<DevMlp> edge_mlp, node_mlp;
    float* readout_w = nullptr;
    float readout_b = 0.f;
    std::vector<void*> allocs;
    int launches = 0;
    bool prefer_tf32x3 = false;  // this model's default product when SLS_B200_GNN_PRECISION is unset (see sls_gnn_model_create)
    bool f16_range_ok = true;    // every weight and bias is finite and within fp16's range: fp16 pieces are the default
    bool profile = false;        // sls_gnn_set_profile: event pairs around every kernel class of a forward
    double prof_ms[4] = {0, 0, 0, 0};
    // Activation workspace (edge / node feature ping-pong buffers and the two aggregates): one block, grown on
    // demand and kept across calls, so a repeated forward of the same batch shape allocates nothing.
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // graph workspace of the batched LCG path (edge lists, segment CSRs, outputs), same policy
    void* gws = nullptr;
    size_t gws_bytes = 0;
};

extern "C" void sls_gnn_model_free(sls_gnn_model* m)
{
    if (!m) return;
    for (void* p : m->allocs) dfree(p);
    dfree(m->ws);
    dfree(m->gws);
    delete m;
}

namespace {

// haiku Linear stores w as [in][out]; the tensor-core B operand wants [out][in] (K-major), zero padded
int pack_mlp(sls_gnn_model* m, const float*& p, int k1, int h1, int h2, DevMlp& out)
{
    out.k1 = k1;
    out.h1 = h1;
    out.h2 = h2;
    out.k1_slabs = (k1 + SLAB_K - 1) / SLAB_K;
    out.k2_slabs = (h1 + SLAB_K - 1) / SLAB_K;
    out.n1 = round_up(h1, 16);
    out.n2 = round_up(h2, 16);
    const int k1p = out.k1_slabs * SLAB_K, k2p = out.k2_slabs * SLAB_K;
    // 3xTF32: w = hi + lo with hi = tf32(w), lo = tf32(w - hi).  Stored as shared-memory slab images:
    // slab s = W^T[n][32 s .. 32 s + 31] for all n, in the K-major SWIZZLE_128B layout the kernel's
    // UMMA descriptors expect (8-row groups of 1024 B, 16-byte unit index XOR n % 8), zero padded.
    auto img_index = [](int n_pad, int s, int n, int k_in_slab) {
        const int u = k_in_slab >> 2, e = k_in_slab & 3;
        return (size_t)s * n_pad * SLAB_K + (size_t)(n >> 3) * 256 + (size_t)(n & 7) * 32 + (size_t)((u ^ (n & 7)) << 2) + e;
    };
    // the same weights as two bf16 pieces b0 = bf16(w), b1 = bf16(w - b0): 64 k-values per 128-byte row, 16-byte unit = 8 values
    constexpr int SLAB_KB = 64;
    // one more k-value per GEMM: the kernel appends a constant 1 to every A row and the images hold the bias in that weight
    // row (b1 behind the k1 gathered columns, b2 behind the h1 hidden columns), so the bias is added by the tensor core and
    // neither the GEMM2 staging nor the epilogue loads it (two bf16 pieces like every weight: 16 mantissa bits)
    out.k1_slabs_bf = (k1 + 1 + SLAB_KB - 1) / SLAB_KB;
    out.k2_slabs_bf = (h1 + 1 + SLAB_KB - 1) / SLAB_KB;
    auto img_index_bf = [](int n_pad, int s, int n, int k_in_slab) {
        const int u = k_in_slab >> 3, e = k_in_slab & 7;
        return (size_t)s * n_pad * SLAB_KB + (size_t)(n >> 3) * 512 + (size_t)(n & 7) * 64 + (size_t)((u ^ (n & 7)) << 3) + e;
    };
    // GEMM1's A operand reaches tensor memory through tcgen05.st.16x256b (four lanes per row, see the kernel's producer): inside
    // every 64-k slab the k-value 32 h + 8 j + 4 q + 2 b + par (unit h, lane j of the row's quad, 8-column repeat q, column b of
    // the lane's pair, element par) sits at position 16 (2 h + q) + 4 j + 2 b + par of the MMA's k-order.  W1's slab images
    // carry the same order; W2 (A staged thread = row out of the first accumulator) keeps the natural one.
    auto w1_perm = [](int kk) {
        const int h = kk >> 5, r = kk & 31, j = r >> 3, q = (r >> 2) & 1, b = (r >> 1) & 1, par = r & 1;
        return 16 * (2 * h + q) + 4 * j + 2 * b + par;
    };
    std::vector<uint16_t> w1b0((size_t)out.n1 * out.k1_slabs_bf * SLAB_KB, 0), w1b1(w1b0.size(), 0);
    std::vector<uint16_t> w2b0((size_t)out.n2 * out.k2_slabs_bf * SLAB_KB, 0), w2b1(w2b0.size(), 0);
    std::vector<uint16_t> w1h0(w1b0.size(), 0), w1h1(w1b0.size(), 0), w2h0(w2b0.size(), 0), w2h1(w2b0.size(), 0);
    std::vector<uint16_t>* twin[4][2] = {{&w1b0, &w1h0}, {&w1b1, &w1h1}, {&w2b0, &w2h0}, {&w2b1, &w2h1}};
    auto f16_twin = [&](std::vector<uint16_t>& v) -> std::vector<uint16_t>& {
        for (auto& t : twin)
            if (t[0] == &v) return *t[1];
        return v;
    };
    // both 16-bit formats in one pass: w = p0 + p1 with p0 = round(w), p1 = round(w - p0); fp16 pieces carry 11 + 11 mantissa bits
    // (bf16: 8 + 8), at the price of fp16's range -- see sls_gnn_model_create for when they are the default
    auto put_bf = [&](std::vector<uint16_t>& p0, std::vector<uint16_t>& p1, size_t ix, float w) {
        p0[ix] = bf16_bits_host(w);
        p1[ix] = bf16_bits_host(w - bf16_value_host(p0[ix]));
        std::vector<uint16_t>&h0 = f16_twin(p0), &h1v = f16_twin(p1);
        h0[ix] = f16_bits_host(w);
        h1v[ix] = f16_bits_host(w - f16_value_host(h0[ix]));
        if (!(std::fabs(w) <= 65504.f)) m->f16_range_ok = false;
    };
    std::vector<float> w1h((size_t)out.n1 * k1p, 0.f), w1l(w1h.size(), 0.f), b1(out.n1, 0.f);
    std::vector<float> w2h((size_t)out.n2 * k2p, 0.f), w2l(w2h.size(), 0.f), b2(out.n2, 0.f);
    for (int k = 0; k < k1; k++)
        for (int n = 0; n < h1; n++) {
            const float w = p[(size_t)k * h1 + n], hi = round_tf32_host(w);
            const size_t ix = img_index(out.n1, k / SLAB_K, n, k % SLAB_K);
            w1h[ix] = hi;
            w1l[ix] = round_tf32_host(w - hi);
            put_bf(w1b0, w1b1, img_index_bf(out.n1, k / SLAB_KB, n, w1_perm(k % SLAB_KB)), w);
        }
    p += (size_t)k1 * h1;
    std::copy(p, p + h1, b1.begin());
    for (int n = 0; n < h1; n++) put_bf(w1b0, w1b1, img_index_bf(out.n1, k1 / SLAB_KB, n, w1_perm(k1 % SLAB_KB)), p[n]);
    p += h1;
    for (int k = 0; k < h1; k++)
        for (int n = 0; n < h2; n++) {
            const float w = p[(size_t)k * h2 + n], hi = round_tf32_host(w);
            const size_t ix = img_index(out.n2, k / SLAB_K, n, k % SLAB_K);
            w2h[ix] = hi;
            w2l[ix] = round_tf32_host(w - hi);
            put_bf(w2b0, w2b1, img_index_bf(out.n2, k / SLAB_KB, n, k % SLAB_KB), w);
        }
    p += (size_t)h1 * h2;
    std::copy(p, p + h2, b2.begin());
    for (int n = 0; n < h2; n++) put_bf(w2b0, w2b1, img_index_bf(out.n2, h1 / SLAB_KB, n, h1 % SLAB_KB), p[n]);
    p += h2;
    std::vector<float> sc(p, p + h2), of(p + h2, p + 2 * h2);
    {
        // What the next GEMM gathers are this LayerNorm's outputs and sums of them over a node's edges: |y| <= |scale| sqrt(h2)
        // + |offset|.  fp16 pieces are the model's default only while a few hundred such rows can be summed inside fp16's
        // range (any model near the initialisation: the bound is 14); otherwise the model keeps fp32's range with bf16 pieces.
        float smax = 0.f, omax = 0.f;
        for (float v : sc) smax = std::max(smax, std::fabs(v));
        for (float v : of) omax = std::max(omax, std::fabs(v));
        if (!(smax * std::sqrt((float)h2) + omax <= 256.f)) m->f16_range_ok = false;
    }
    p += 2 * h2;
    int rc;
    if ((rc = upload(w1h, &out.w1_hi)) || (rc = upload(w1l, &out.w1_lo)) || (rc = upload(b1, &out.b1)) || (rc = upload(w2h, &out.w2_hi)) ||
        (rc = upload(w2l, &out.w2_lo)) || (rc = upload(b2, &out.b2)) || (rc = upload(sc, &out.ln_scale)) ||
        (rc = upload(of, &out.ln_offset)))
        return rc;
    for (float* q : {out.w1_hi, out.w1_lo, out.b1, out.w2_hi, out.w2_lo, out.b2, out.ln_scale, out.ln_offset}) m->allocs.push_back(q);
    if ((rc = upload_u16(w1b0, &out.w1_b0)) || (rc = upload_u16(w1b1, &out.w1_b1)) || (rc = upload_u16(w2b0, &out.w2_b0)) ||
        (rc = upload_u16(w2b1, &out.w2_b1)))
        return rc;
    for (uint16_t* q : {out.w1_b0, out.w1_b1, out.w2_b0, out.w2_b1}) m->allocs.push_back(q);
    if ((rc = upload_u16(w1h0, &out.w1_h0)) || (rc = upload_u16(w1h1, &out.w1_h1)) || (rc = upload_u16(w2h0, &out.w2_h0)) ||
        (rc = upload_u16(w2h1, &out.w2_h1)))
        return rc;
    for (uint16_t* q : {out.w1_h0, out.w1_h1, out.w2_h0, out.w2_h1}) m->allocs.push_back(q);
    return SLS_OK;
}

uint64_t expected_blob_len(int steps, int D, int h1, int h2)
{
    uint64_t n = 2 * (2ull * D + D);
    for (int t = 0; t < steps; t++) {
        const int de = t == 0 ? D : h2, dn = t == 0 ? D : h2;
        const uint64_t k1e = de + 2ull * dn, k1n = dn + 2ull * h2;
        n += k1e * h1 + h1 + (uint64_t)h1 * h2 + h2 + 2ull * h2;
        n += k1n * h1 + h1 + (uint64_t)h1 * h2 + h2 + 2ull * h2;
    }
    return n + h2 + 1;
}

}  // namespace

extern "C" int sls_gnn_model_create(const float* blob, uint64_t blob_len, int32_t num_steps, int32_t embed_dim, int32_t h1, int32_t h2,
                                    sls_gnn_model** out)
{
    if (!blob || !out) return sls_set_error_internal(SLS_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (num_steps < 1 || embed_dim < 4 || embed_dim % 4 || h1 < 4 || h2 < 4 || h2 % 4 || h1 > MAX_N || h2 > MAX_N)
        return sls_set_error_internal(SLS_ERR_ARG, "unsupported network shape (need embed, h2 multiples of 4; h1, h2 <= 208)");
    if (blob_len != expected_blob_len(num_steps, embed_dim, h1, h2))
        return sls_set_error_internal(SLS_ERR_ARG, "parameter blob has the wrong length for this network shape");
    int dev = 0;
    G_TRY(cudaGetDevice(&dev));
    sls_gnn_model* m = new (std::nothrow) sls_gnn_model();
    if (!m) return sls_set_error_internal(SLS_ERR_NOMEM, "out of memory");
    m->num_steps = num_steps;
    m->embed = embed_dim;
    m->h1 = h1;
    m->h2 = h2;
    const float* p = blob;
    const int D = embed_dim;
    int rc = SLS_OK;
    auto up = [&](size_t n, float** d) {
        std::vector<float> h(p, p + n);
        p += n;
        int r = upload(h, d);
        if (!r) m->allocs.push_back(*d);
        return r;
    };
    if ((rc = up(2 * D, &m->edge_embed_w)) || (rc = up(D, &m->edge_embed_b)) || (rc = up(2 * D, &m->node_embed_w)) ||
        (rc = up(D, &m->node_embed_b))) {
        sls_gnn_model_free(m);
        return rc;
    }
    m->edge_mlp.resize(num_steps);
    m->node_mlp.resize(num_steps);
    for (int t = 0; t < num_steps && !rc; t++) {
        const int de = t == 0 ? D : h2, dn = t == 0 ? D : h2;
        rc = pack_mlp(m, p, de + 2 * dn, h1, h2, m->edge_mlp[t]);
        if (!rc) rc = pack_mlp(m, p, dn + 2 * h2, h1, h2, m->node_mlp[t]);
    }
    if (!rc) rc = up(h2, &m->readout_w);
    if (rc) {
        sls_gnn_model_free(m);
        return rc;
    }
    m->readout_b = *p;
    cudaError_t e = cudaSuccess;
    for (auto kern : {mlp_ln_persistent_ts_kernel<0, 1>, mlp_ln_persistent_ts_kernel<1, 1>, mlp_ln_persistent_ts_kernel<1, 2>,
                      mlp_ln_persistent_ts_kernel<2, 1>, mlp_ln_persistent_ts_kernel<2, 2>})
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES);
    if (e != cudaSuccess) {
        sls_gnn_model_free(m);
        return sls_set_error_internal(SLS_ERR_CUDA, std::string("cudaFuncSetAttribute(mlp_ln_kernel): ") + cudaGetErrorString(e));
    }
    *out = m;
    return SLS_OK;
}

namespace {

// two CTAs per cluster share every weight slab by multicast: 1-2 % on the bf16x2 kernel (15.98 -> 15.76 ms per forward), bit-identical
constexpr int GNN_DEFAULT_CLUSTER = 2;

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { dfree(p); }
    template <class T>
    T* as() { return static_cast<T*>(p); }
};

int launch_mlp(sls_gnn_model* m, const DevMlp& w, const float* s0, const int32_t* i0, int d0, const float* s1, const int32_t* i1, int d1,
               const float* s2, const int32_t* i2, int d2, int64_t rows, float* out)
{
    if (rows == 0) return SLS_OK;
    MlpArgs a{};
    a.src[0] = s0, a.src[1] = s1, a.src[2] = s2;
    a.idx[0] = i0, a.idx[1] = i1, a.idx[2] = i2;
    a.d[0] = d0, a.d[1] = d1, a.d[2] = d2;
    a.rows = (int32_t)rows;
    a.k1_slabs = w.k1_slabs, a.k2_slabs = w.k2_slabs, a.n1 = w.n1, a.n2 = w.n2, a.h2 = w.h2, a.h1 = w.h1;
    a.w1_hi = w.w1_hi, a.w1_lo = w.w1_lo, a.b1 = w.b1, a.w2_hi = w.w2_hi, a.w2_lo = w.w2_lo, a.b2 = w.b2;
    a.ln_scale = w.ln_scale, a.ln_offset = w.ln_offset;
    a.out = out;
    a.w1_b0 = w.w1_b0, a.w1_b1 = w.w1_b1, a.w2_b0 = w.w2_b0, a.w2_b1 = w.w2_b1;
    a.k1_slabs_bf = w.k1_slabs_bf, a.k2_slabs_bf = w.k2_slabs_bf;
    a.trace = nullptr;
#ifdef GNN_TRACE
    // development build: stamps of the launch number SLS_B200_GNN_TRACE_LAUNCH (counted over the process) go to the file
    // SLS_B200_GNN_TRACE names, after the launch has finished
    static int launch_no = 0;
    static unsigned long long* d_trace = nullptr;
    const char* tf = std::getenv("SLS_B200_GNN_TRACE");
    const char* tl = std::getenv("SLS_B200_GNN_TRACE_LAUNCH");
    const bool trace_this = tf && launch_no == (tl ? std::atoi(tl) : 0);
    launch_no++;
    if (trace_this) {
        if (!d_trace) cudaMalloc(&d_trace, TRACE_WORDS * 8);
        cudaMemset(d_trace, 0, TRACE_WORDS * 8);
        a.trace = d_trace;
    }
#endif
    const unsigned blocks = (unsigned)((rows + TILE_M - 1) / TILE_M);
    {
        // SLS_B200_GNN_CLUSTER = 1 | 2: CTAs per cluster sharing the weight stream by multicast (see the kernel)
        const char* f = std::getenv("SLS_B200_GNN_CLUSTER");  // read per launch: the tests switch it
        const int c = f ? std::atoi(f) : GNN_DEFAULT_CLUSTER;
        // SLS_B200_GNN_PRECISION = fp16x2 | bf16x2 | tf32x3.  Default fp16x2: two fp16 pieces per operand on kind::f16, three
        // product terms a0*b0 + a1*b0 + a0*b1 with fp32 accumulation -- 22 mantissa bits per operand at the cost of three
        // 16-bit MMAs (tools/precision_split_study.py: max |dp| 3e-6 against fp64, the fp32 restatement itself 2e-6; bf16
        // pieces 3e-5 at the same cost; 3xTF32 2e-6 at twice the tensor time).  fp16's range: activations beyond 65504
        // saturate piece by piece (exact up to 131008), weights beyond it switch the model to bf16x2 (pack_mlp).
        const char* fp = std::getenv("SLS_B200_GNN_PRECISION");
        const int prec = fp ? (std::strcmp(fp, "tf32x3") == 0 ? 0 : std::strcmp(fp, "bf16x2") == 0 ? 1 : 2)
                            : (m->prefer_tf32x3 ? 0 : m->f16_range_ok ? 2 : 1);
        if (prec == 2) a.w1_b0 = w.w1_h0, a.w1_b1 = w.w1_h1, a.w2_b0 = w.w2_h0, a.w2_b1 = w.w2_h1;
        const bool tf32x3 = prec == 0;
        const int cl = (!tf32x3 && c == 2) ? 2 : 1;
        void (*kern)(const MlpArgs) = tf32x3      ? mlp_ln_persistent_ts_kernel<0, 1>
                                      : prec == 1 ? (cl == 2 ? mlp_ln_persistent_ts_kernel<1, 2> : mlp_ln_persistent_ts_kernel<1, 1>)
                                                  : (cl == 2 ? mlp_ln_persistent_ts_kernel<2, 2> : mlp_ln_persistent_ts_kernel<2, 1>);
        const size_t smem = TS_SMEM_BYTES;
        if (cl == 1) {
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            kern<<<std::min<unsigned>(blocks, (unsigned)sms), MLPP_THREADS, smem>>>(a);
        } else {
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = (unsigned)cl, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
            cfg.blockDim = dim3(MLPP_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = 0;
            cfg.attrs = &attr, cfg.numAttrs = 1;
            // the grid is one wave of co-resident clusters (tiles are assigned statically): ask how many fit
            static int max_clusters = 0;
            if (max_clusters == 0) {
                cfg.gridDim = dim3((unsigned)cl);
                int n = 0;
                G_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
                if (n < 1) return sls_set_error_internal(SLS_ERR_CUDA, "no cluster of the oracle-forward kernel fits on this device");
                max_clusters = n;
            }
            const unsigned want = (blocks + (unsigned)cl - 1) / (unsigned)cl;
            cfg.gridDim = dim3(std::min<unsigned>(want, (unsigned)max_clusters) * (unsigned)cl);
            G_TRY(cudaLaunchKernelEx(&cfg, kern, a));
        }
    }
    m->launches++;
    G_TRY(cudaGetLastError());
#ifdef GNN_TRACE
    if (trace_this) {
        cudaDeviceSynchronize();
        std::vector<unsigned long long> h(TRACE_WORDS);
        cudaMemcpy(h.data(), d_trace, TRACE_WORDS * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = std::fopen(tf, "wb")) {
            const int hdr[4] = {a.k1_slabs_bf, a.k2_slabs_bf, (int)rows, TRACE_WORDS};
            std::fwrite(hdr, 4, 4, f);
            std::fwrite(h.data(), 8, h.size(), f);
            std::fclose(f);
        }
    }
#endif
    return SLS_OK;
}

// the whole forward on device-resident inputs; decoded_dev[N]
//
// Workspace (one block, rows of H floats unless stated):
//   eA [E] | sent [N] | recv [N] | zero [1] | eB [E]     one row-addressable array: the node update reads its two aggregates
//                                                         through row indices relative to eA (agg_index_kernel), which point at
//                                                         a summed row (degree >= 2), straight at the single edge row (degree 1)
//                                                         or at the zero row (degree 0: clause nodes send nothing, positive
//                                                         literals receive nothing -- 2/3 of an LCG graph's (node, direction)
//                                                         pairs, which round 1 wrote as zeros and read back)
//   hA [N] | hB [N] | e0 [E x D] | h0 [N x D] | four int32 index arrays [N]
// The new edges of step t are written to eA (t even) or eB (t odd); every kernel reads rows it does not write.
int forward_device(sls_gnn_model* m, int64_t N, int64_t E, const float* d_nodes, const float* d_edges, const uint8_t* d_node_type,
                   const uint8_t* d_edge_type, const int32_t* d_send,
                   const int32_t* d_recv, const int64_t* d_sptr, const int32_t* d_sids, const int64_t* d_rptr, const int32_t* d_rids,
                   float* d_decoded, cudaEvent_t ev0, cudaEvent_t ev1)
{
    const int D = m->embed, H = m->h2;
    if (2 * E + 2 * N + 1 > 0x7fffffffll) return sls_set_error_internal(SLS_ERR_ARG, "graph too large for one call");
    auto a256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t rowsH = (size_t)(2 * E + 2 * N + 1);
    const size_t b_e = a256(rowsH * H * 4), b_h = a256((size_t)std::max<int64_t>(N, 1) * H * 4), b_e0 = a256((size_t)std::max<int64_t>(E, 1) * D * 4),
                 b_h0 = a256((size_t)std::max<int64_t>(N, 1) * D * 4), b_ix = a256((size_t)std::max<int64_t>(N, 1) * 4);
    const size_t need = b_e + 2 * b_h + b_e0 + b_h0 + 4 * b_ix;
    if (need > m->ws_bytes) {
        dfree(m->ws);
        m->ws = nullptr, m->ws_bytes = 0;
        G_TRY(dmalloc(&m->ws, need));
        m->ws_bytes = need;
    }
    char* base = static_cast<char*>(m->ws);
    float* eA = (float*)base;
    float* sent = eA + (size_t)E * H;
    float* recv = sent + (size_t)N * H;
    float* zero = recv + (size_t)N * H;
    float* eB = zero + H;
    float* hA = (float*)(base + b_e);
    float* hB = (float*)(base + b_e + b_h);
    float* e0 = (float*)(base + b_e + 2 * b_h);
    float* h0 = (float*)(base + b_e + 2 * b_h + b_e0);
    int32_t* ix = (int32_t*)(base + b_e + 2 * b_h + b_e0 + b_h0);
    int32_t *sidxA = ix, *ridxA = (int32_t*)((char*)ix + b_ix), *sidxB = (int32_t*)((char*)ix + 2 * b_ix), *ridxB = (int32_t*)((char*)ix + 3 * b_ix);

    // optional per-class device times (sls_gnn_set_profile): 0 edge update, 1 node update, 2 segment sums, 3 everything else
    struct Rec {
        int cls;
        cudaEvent_t a, b;
    };
    std::vector<Rec> recs;
    auto timed = [&](int cls, auto&& launch) -> int {
        if (!m->profile) {
            launch();
            return SLS_OK;
        }
        Rec r{cls, nullptr, nullptr};
        G_TRY(cudaEventCreate(&r.a));
        G_TRY(cudaEventCreate(&r.b));
        G_TRY(cudaEventRecord(r.a));
        launch();
        G_TRY(cudaEventRecord(r.b));
        recs.push_back(r);
        return SLS_OK;
    };
    int rc = SLS_OK;
    if (ev0) G_TRY(cudaEventRecord(ev0));
    if ((rc = timed(3, [&] {
             cudaMemsetAsync(zero, 0, (size_t)H * 4);
             if (E && d_edge_type)
                 embed_onehot_kernel<<<(unsigned)((E * D + 255) / 256), 256>>>(d_edge_type, m->edge_embed_w, m->edge_embed_b, e0, E, D);
             else if (E)
                 embed_kernel<<<(unsigned)((E * D + 255) / 256), 256>>>(d_edges, m->edge_embed_w, m->edge_embed_b, e0, E, D);
             if (N && d_node_type)
                 embed_onehot_kernel<<<(unsigned)((N * D + 255) / 256), 256>>>(d_node_type, m->node_embed_w, m->node_embed_b, h0, N, D);
             else if (N)
                 embed_kernel<<<(unsigned)((N * D + 255) / 256), 256>>>(d_nodes, m->node_embed_w, m->node_embed_b, h0, N, D);
             if (N) agg_index_kernel<<<(unsigned)((N + 255) / 256), 256>>>(d_sptr, d_sids, d_rptr, d_rids, N, E, sidxA, ridxA, sidxB, ridxB);
         })))
        return rc;
    m->launches += 3;
    const float *e = e0, *h = h0;
    int de = D, dn = D;
    for (int t = 0; t < m->num_steps; t++) {
        const bool even = (t & 1) == 0;
        float* e_new = even ? eA : eB;
        float* h_new = even ? hA : hB;
        // edge update sees [edges | nodes[senders] | nodes[receivers]]  (jraph InteractionNetwork + concatenated_args)
        int lrc = SLS_OK;
        if ((rc = timed(0, [&] { lrc = launch_mlp(m, m->edge_mlp[t], e, nullptr, de, h, d_send, dn, h, d_recv, dn, E, e_new); })) || (rc = lrc)) return rc;
        e = e_new;
        de = H;
        // node update sees [nodes | segment_sum(new edges, senders) | segment_sum(new edges, receivers)]; only the sums over two
        // or more edges are materialised
        if (N) {
            const unsigned blocks = (unsigned)((N * 32 + 255) / 256);
            if ((rc = timed(2, [&] { segment_sum_kernel<<<blocks, 256>>>(e, d_sptr, d_sids, sent, d_rptr, d_rids, recv, N, H); })))
                return rc;
            m->launches += 1;
        }
        if ((rc = timed(1, [&] {
                 lrc = launch_mlp(m, m->node_mlp[t], h, nullptr, dn, eA, even ? sidxA : sidxB, H, eA, even ? ridxA : ridxB, H, N, h_new);
             })) ||
            (rc = lrc))
            return rc;
        h = h_new;
        dn = H;
    }
    if (N) {
        if ((rc = timed(3, [&] { readout_kernel<<<(unsigned)((N * 32 + 255) / 256), 256>>>(h, m->readout_w, m->readout_b, d_decoded, N, H); }))) return rc;
        m->launches++;
    }
    G_TRY(cudaGetLastError());
    if (ev1) G_TRY(cudaEventRecord(ev1));
    G_TRY(cudaStreamSynchronize(nullptr));
    if (m->profile) {
        for (double& x : m->prof_ms) x = 0.0;
        for (Rec& r : recs) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            m->prof_ms[r.cls] += ms;
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
    }
    return SLS_OK;
}

// CSR of edge ids per node, edges of one node in increasing edge id (stable counting sort)
void build_csr(int64_t N, int64_t E, const int32_t* key, std::vector<int64_t>& ptr, std::vector<int32_t>& ids)
{
    ptr.assign(N + 1, 0);
    for (int64_t i = 0; i < E; i++) ptr[key[i] + 1]++;
    for (int64_t v = 0; v < N; v++) ptr[v + 1] += ptr[v];
    ids.resize(std::max<int64_t>(E, 1));
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < E; i++) ids[fill[key[i]]++] = (int32_t)i;
}

template <class T>
int to_device(const std::vector<T>& h, DevBuf& d)
{
    G_TRY(dmalloc(&d.p, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty()) G_TRY(cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SLS_OK;
}

int forward_host_graph(sls_gnn_model* m, int64_t N, int64_t E, const float* nodes, const float* edges, const int32_t* senders,
                       const int32_t* receivers, float* decoded, DevBuf* keep_decoded, double* kernel_ms)
{
    for (int64_t i = 0; i < E; i++)
        if (senders[i] < 0 || senders[i] >= N || receivers[i] < 0 || receivers[i] >= N)
            return sls_set_error_internal(SLS_ERR_ARG, "edge endpoint out of range");
    std::vector<int64_t> sptr, rptr;
    std::vector<int32_t> sids, rids;
    build_csr(N, E, senders, sptr, sids);
    build_csr(N, E, receivers, rptr, rids);
    DevBuf d_nodes, d_edges, d_send, d_recv, d_sptr, d_sids, d_rptr, d_rids, d_dec;
    int rc;
    std::vector<float> hn(nodes, nodes + 2 * N), he(edges, edges + 2 * E);
    std::vector<int32_t> hs(senders, senders + E), hr(receivers, receivers + E);
    if ((rc = to_device(hn, d_nodes)) || (rc = to_device(he, d_edges)) || (rc = to_device(hs, d_send)) || (rc = to_device(hr, d_recv)) ||
        (rc = to_device(sptr, d_sptr)) || (rc = to_device(sids, d_sids)) || (rc = to_device(rptr, d_rptr)) ||
        (rc = to_device(rids, d_rids)))
        return rc;
    G_TRY(dmalloc(&d_dec.p, std::max<int64_t>(N, 1) * sizeof(float)));
    cudaEvent_t e0, e1;
    G_TRY(cudaEventCreate(&e0));
    G_TRY(cudaEventCreate(&e1));
    m->launches = 0;
    rc = forward_device(m, N, E, d_nodes.as<float>(), d_edges.as<float>(), nullptr, nullptr, d_send.as<int32_t>(), d_recv.as<int32_t>(),
                        d_sptr.as<int64_t>(), d_sids.as<int32_t>(), d_rptr.as<int64_t>(), d_rids.as<int32_t>(), d_dec.as<float>(), e0, e1);
    if (!rc) {
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (kernel_ms) *kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (decoded && N) G_TRY(cudaMemcpy(decoded, d_dec.p, N * sizeof(float), cudaMemcpyDeviceToHost));
    if (keep_decoded) std::swap(keep_decoded->p, d_dec.p);
    return SLS_OK;
}

}  // namespace

extern "C" int sls_gnn_forward_graph(sls_gnn_model* m, uint64_t n_node, uint64_t n_edge, const float* nodes, const float* edges,
                                     const int32_t* senders, const int32_t* receivers, float* decoded, double* kernel_ms)
{
    if (!m || (n_node && !nodes) || (n_edge && (!edges || !senders || !receivers)))
        return sls_set_error_internal(SLS_ERR_ARG, "NULL argument");
    if (n_node > 0x7fffffffull || n_edge > 0x7fffffffull) return sls_set_error_internal(SLS_ERR_ARG, "graph too large for one call");
    return forward_host_graph(m, (int64_t)n_node, (int64_t)n_edge, nodes, edges, senders, receivers, decoded, nullptr, kernel_ms);
}

namespace {

// One instance's slice of the batched LCG graph (LCG.get_graph, sat_representations.py:607-666):
// 2n literal nodes (2j = negative, 2j+1 = positive literal of variable j) then its non-empty
// clauses; one edge literal -> clause per literal occurrence in file order, then one edge
// positive -> negative literal per variable.  Also both segment CSRs, written directly:
// receivers are clause nodes (their edges are contiguous) or negative literals (the pair edge);
// senders are literal nodes (their occurrences in increasing edge id, then the pair edge).
struct LcgPlan {
    int64_t node_off = 0, edge_off = 0, var_off = 0;
    int64_t m_ne = 0, L_ne = 0;
    bool dup_var = false;
};

void lcg_count(const slsb::HostInstance& h, LcgPlan& pl)
{
    pl.m_ne = pl.L_ne = 0;
    for (uint32_t c = 0; c < h.m; c++) {
        const uint32_t s = h.row_ptr[c], e = h.row_ptr[c + 1];
        if (s == e) continue;  // sat_instances.py:50 drops empty clauses
        pl.m_ne++;
        pl.L_ne += e - s;
        for (uint32_t j = s + 1; j < e && !pl.dup_var; j++)
            for (uint32_t q = s; q < j; q++)
                if ((h.lits[q] >> 1) == (h.lits[j] >> 1)) pl.dup_var = true;  // sat_representations.py:647-650 asserts
    }
}

void lcg_fill(const slsb::HostInstance& h, const LcgPlan& pl, int32_t* send, int32_t* recv, uint8_t* node_type, uint8_t* edge_type,
              int64_t* sptr, int32_t* sids, int64_t* rptr, int32_t* rids, int64_t* var_node0)
{
    const int64_t n = h.n, no = pl.node_off, eo = pl.edge_off;
    const int64_t E = pl.L_ne + n;
    // node types: literal nodes one-hot index 1 (np.eye(2)[1] and np.eye(2)[-1] are both [0,1]), clause nodes 0
    for (int64_t j = 0; j < 2 * n; j++) node_type[no + j] = 1;
    for (int64_t j = 0; j < pl.m_ne; j++) node_type[no + 2 * n + j] = 0;
    // edges + receiver CSR + sender degree
    std::vector<int64_t> deg(2 * n + 1, 0);
    int64_t ei = 0, cn = 0;
    for (uint32_t c = 0; c < h.m; c++) {
        const uint32_t s = h.row_ptr[c], e = h.row_ptr[c + 1];
        if (s == e) continue;
        for (uint32_t j = s; j < e; j++, ei++) {
            const uint32_t lit = h.lits[j];
            const int64_t node = 2 * (int64_t)(lit >> 1) + ((lit & 1u) ? 0 : 1);
            send[eo + ei] = (int32_t)(no + node);
            recv[eo + ei] = (int32_t)(no + 2 * n + cn);
            edge_type[eo + ei] = 0;
            deg[node + 1]++;
        }
        cn++;
    }
    for (int64_t j = 0; j < n; j++) {
        send[eo + pl.L_ne + j] = (int32_t)(no + 2 * j + 1);
        recv[eo + pl.L_ne + j] = (int32_t)(no + 2 * j);
        edge_type[eo + pl.L_ne + j] = 1;
        deg[2 * j + 1 + 1]++;
        var_node0[pl.var_off + j] = no + 2 * j;
    }
    // receiver ids: the instance's block of rids is [pair edges (for literal nodes) | clause edges (for clause nodes)]
    //   literal node 2j   -> rids[eo + j]            = pair edge j      (rptr = eo + j .. eo + j + 1)
    //   literal node 2j+1 -> empty                    (rptr = eo + j + 1)
    //   clause node cn    -> rids[eo + n + first .. ] = its clause edges
    for (int64_t j = 0; j < 2 * n; j++) rptr[no + j] = eo + (j + 1) / 2;
    for (int64_t j = 0; j < n; j++) rids[eo + j] = (int32_t)(eo + pl.L_ne + j);
    ei = 0, cn = 0;
    for (uint32_t c = 0; c < h.m; c++) {
        const uint32_t s = h.row_ptr[c], e = h.row_ptr[c + 1];
        if (s == e) continue;
        rptr[no + 2 * n + cn] = eo + n + ei;
        for (uint32_t j = s; j < e; j++, ei++) rids[eo + n + ei] = (int32_t)(eo + ei);
        cn++;
    }
    // sender ids: counting sort by literal node, stable in edge id; clause nodes send nothing
    for (int64_t j = 0; j < 2 * n; j++) deg[j + 1] += deg[j];
    for (int64_t j = 0; j < 2 * n; j++) sptr[no + j] = eo + deg[j];
    for (int64_t j = 0; j < pl.m_ne; j++) sptr[no + 2 * n + j] = eo + E;
    std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
    for (int64_t k = 0; k < E; k++) sids[eo + fill[send[eo + k] - no]++] = (int32_t)(eo + k);
}

}  // namespace

namespace {

// The batched path for resident instances: the graph arrays are produced on the device (gnn_kernels.cuh,
// lcg_*_kernel) from the instances' CSRs in HBM, so nothing but two small per-instance tables goes up and only
// the probabilities come back.
int forward_lcg_device(sls_gnn_model* m, sls_instance* const* insts, uint32_t n_inst, const std::vector<LcgPlan>& plan, int64_t N, int64_t E,
                       int64_t V, float* prob, float* decoded, double* kernel_ms)
{
    int rc = slsb_ensure_batch_on_device(insts, n_inst);
    if (rc) return rc;
    std::vector<LcgItem> items(n_inst);
    std::vector<int64_t> begins(2 * (size_t)(n_inst + 1));  // clause_begin | var_begin
    int64_t C = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        const slsb::DevInst& d = insts[i]->dev;
        LcgItem& it = items[i];
        it.row_ptr = d.row_ptr, it.lits = d.lits, it.occ_ptr = d.occ_ptr, it.occ = d.occ;
        it.n = d.n, it.m = d.m, it.L = (uint32_t)insts[i]->host.lits.size(), it.pad = 0;
        it.node_off = plan[i].node_off, it.edge_off = plan[i].edge_off, it.var_off = plan[i].var_off, it.clause_off = C;
        begins[i] = C;
        begins[n_inst + 1 + i] = plan[i].var_off;
        C += d.m;
    }
    begins[n_inst] = C;
    begins[2 * (size_t)n_inst + 1] = V;

    auto a256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t bE4 = a256((size_t)E * 4), bN8 = a256((size_t)(N + 1) * 8), bV8 = a256((size_t)V * 8), bN1 = a256((size_t)N), bE1 = a256((size_t)E),
                 bN4 = a256((size_t)N * 4), bV4 = a256((size_t)V * 4), bItems = a256(items.size() * sizeof(LcgItem)),
                 bBegins = a256(begins.size() * 8);
    const size_t need = 4 * bE4 + 2 * bN8 + bV8 + bN1 + bE1 + bN4 + bV4 + bItems + bBegins;
    if (need > m->gws_bytes) {
        dfree(m->gws);
        m->gws = nullptr, m->gws_bytes = 0;
        G_TRY(dmalloc(&m->gws, need));
        m->gws_bytes = need;
    }
    char* p = static_cast<char*>(m->gws);
    auto take = [&p](size_t bytes) { char* q = p; p += bytes; return q; };
    LcgOut o{};
    o.send = (int32_t*)take(bE4), o.recv = (int32_t*)take(bE4), o.sids = (int32_t*)take(bE4), o.rids = (int32_t*)take(bE4);
    o.sptr = (int64_t*)take(bN8), o.rptr = (int64_t*)take(bN8), o.var_node0 = (int64_t*)take(bV8);
    o.node_type = (uint8_t*)take(bN1), o.edge_type = (uint8_t*)take(bE1);
    o.N = N, o.E = E;
    float* d_dec = (float*)take(bN4);
    float* d_p = (float*)take(bV4);
    LcgItem* d_items = (LcgItem*)take(bItems);
    int64_t* d_begins = (int64_t*)take(bBegins);
    G_TRY(cudaMemcpy(d_items, items.data(), items.size() * sizeof(LcgItem), cudaMemcpyHostToDevice));
    G_TRY(cudaMemcpy(d_begins, begins.data(), begins.size() * 8, cudaMemcpyHostToDevice));
    m->launches = 0;
    if (C) lcg_clauses_kernel<<<(unsigned)((C + 255) / 256), 256>>>(d_items, d_begins, n_inst, C, o);
    lcg_variables_kernel<<<(unsigned)((V + 127) / 128), 128>>>(d_items, d_begins + n_inst + 1, n_inst, V, o);
    G_TRY(cudaGetLastError());
    cudaEvent_t e0, e1;
    G_TRY(cudaEventCreate(&e0));
    G_TRY(cudaEventCreate(&e1));
    rc = forward_device(m, N, E, nullptr, nullptr, o.node_type, o.edge_type, o.send, o.recv, o.sptr, o.sids, o.rptr, o.rids, d_dec, e0, e1);
    m->launches += 2;
    if (!rc) {
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (kernel_ms) *kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (decoded && N) G_TRY(cudaMemcpy(decoded, d_dec, N * sizeof(float), cudaMemcpyDeviceToHost));
    if (prob && V) {
        pair_prob_kernel<<<(unsigned)((V + 255) / 256), 256>>>(d_dec, o.var_node0, d_p, V);
        m->launches++;
        G_TRY(cudaGetLastError());
        G_TRY(cudaMemcpy(prob, d_p, V * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return SLS_OK;
}

}  // namespace

namespace {
int forward_lcg_one(sls_gnn_model* m, sls_instance* const* insts, uint32_t n_inst, float* prob, float* decoded, double* kernel_ms);
}

// A batch is cut into chunks whose activation workspace stays below a budget (8 GB by default = ~380 instances of the
// config-4 shape, far more than it takes to fill the machine; SLS_B200_GNN_WS_MB overrides it): 10,000 instances in one call
// would otherwise ask for 208 GB.  Rows are independent of their tile neighbours, so chunking does not change a single bit.
extern "C" int sls_gnn_forward_lcg(sls_gnn_model* m, sls_instance* const* insts, uint32_t n_inst, float* prob, float* decoded,
                                   double* kernel_ms)
{
    if (!m || (n_inst && !insts)) return sls_set_error_internal(SLS_ERR_ARG, "NULL argument");
    size_t budget = (size_t)8 << 30;
    if (const char* e = std::getenv("SLS_B200_GNN_WS_MB")) budget = std::max<size_t>(1, (size_t)std::atoll(e)) << 20;
    const size_t wH = (size_t)m->h2 * sizeof(float), wD = (size_t)m->embed * sizeof(float);
    if (kernel_ms) *kernel_ms = 0.0;
    int launches = 0;
    uint32_t begin = 0;
    int64_t v_off = 0, n_off = 0;
    while (begin < n_inst || (n_inst == 0 && begin == 0)) {
        uint32_t end = begin;
        size_t need = 0;
        int64_t v_cnt = 0, n_cnt = 0;
        while (end < n_inst) {
            if (!insts[end]) return sls_set_error_internal(SLS_ERR_ARG, "NULL instance handle in batch");
            const slsb::HostInstance& h = insts[end]->host;
            const size_t nodes = 2 * (size_t)h.n + h.m, edges = h.lits.size() + h.n;
            // forward_device's workspace: two edge buffers, two aggregates and two node buffers of width H, the embeddings, indices
            const size_t bytes = (2 * edges + 4 * nodes) * wH + (edges + nodes) * wD + 16 * nodes;
            if (end > begin && need + bytes > budget) break;
            need += bytes;
            v_cnt += h.n;
            n_cnt += (int64_t)nodes;  // upper bound (empty clauses have no node); only used when no clause is empty
            end++;
        }
        double ms = 0.0;
        // decoded offsets need the exact node count of a chunk: with empty clauses in play a chunked call cannot place them
        if (decoded && (begin > 0 || end < n_inst))
            for (uint32_t i = begin; i < end; i++)
                for (uint32_t c = 0; c < insts[i]->host.m; c++)
                    if (insts[i]->host.row_ptr[c] == insts[i]->host.row_ptr[c + 1])
                        return sls_set_error_internal(SLS_ERR_ARG, "decoded output of a chunked batch needs instances without empty clauses");
        int rc = forward_lcg_one(m, insts + begin, end - begin, prob ? prob + v_off : nullptr, decoded ? decoded + n_off : nullptr, &ms);
        if (rc) return rc;
        if (kernel_ms) *kernel_ms += ms;
        launches += m->launches;
        v_off += v_cnt;
        n_off += n_cnt;
        begin = end;
        if (n_inst == 0) break;
    }
    m->launches = launches;
    return SLS_OK;
}

namespace {

int forward_lcg_one(sls_gnn_model* m, sls_instance* const* insts, uint32_t n_inst, float* prob, float* decoded, double* kernel_ms)
{
    // Graph construction is per instance and independent: plan (count) in parallel, prefix the
    // offsets, fill every instance's slice of the batched arrays in parallel.
    std::vector<LcgPlan> plan(n_inst);
    const unsigned T = std::max(1u, std::min(std::thread::hardware_concurrency(), std::min(32u, n_inst)));
    auto parallel_for = [&](auto&& fn) {
        std::atomic<uint32_t> next{0};
        auto work = [&] {
            for (uint32_t i; (i = next.fetch_add(1)) < n_inst;) fn(i);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    };
    parallel_for([&](uint32_t i) { lcg_count(insts[i]->host, plan[i]); });
    int64_t N = 0, E = 0, V = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        if (plan[i].dup_var) return sls_set_error_internal(SLS_ERR_ARG, "Multiple occurrences of single variable in constraint");
        plan[i].node_off = N, plan[i].edge_off = E, plan[i].var_off = V;
        N += 2 * (int64_t)insts[i]->host.n + plan[i].m_ne;
        E += plan[i].L_ne + insts[i]->host.n;
        V += insts[i]->host.n;
        if (N > 0x7fffffffll || E > 0x7fffffffll) return sls_set_error_internal(SLS_ERR_ARG, "batch too large for one call");
    }
    // test hook: SLS_B200_LCG_HOST=1 forces the host-built graph arrays
    bool device_build = V > 0 && !([] { const char* f = std::getenv("SLS_B200_LCG_HOST"); return f && f[0] == '1'; }());
    for (uint32_t i = 0; i < n_inst && device_build; i++) device_build = plan[i].m_ne == insts[i]->host.m;
    if (device_build) return forward_lcg_device(m, insts, n_inst, plan, N, E, V, prob, decoded, kernel_ms);

    std::vector<int32_t> send(std::max<int64_t>(E, 1)), recv(send.size()), sids(send.size()), rids(send.size());
    std::vector<uint8_t> node_type(std::max<int64_t>(N, 1)), edge_type(send.size());
    std::vector<int64_t> sptr(N + 1, 0), rptr(N + 1, 0), var_node0(std::max<int64_t>(V, 1));
    parallel_for([&](uint32_t i) {
        lcg_fill(insts[i]->host, plan[i], send.data(), recv.data(), node_type.data(), edge_type.data(), sptr.data(), sids.data(),
                 rptr.data(), rids.data(), var_node0.data());
    });
    sptr[N] = E;
    rptr[N] = E;

    DevBuf d_nt, d_et, d_send, d_recv, d_sptr, d_sids, d_rptr, d_rids, d_dec;
    int rc;
    if ((rc = to_device(node_type, d_nt)) || (rc = to_device(edge_type, d_et)) || (rc = to_device(send, d_send)) ||
        (rc = to_device(recv, d_recv)) || (rc = to_device(sptr, d_sptr)) || (rc = to_device(sids, d_sids)) ||
        (rc = to_device(rptr, d_rptr)) || (rc = to_device(rids, d_rids)))
        return rc;
    G_TRY(dmalloc(&d_dec.p, std::max<int64_t>(N, 1) * sizeof(float)));
    cudaEvent_t e0, e1;
    G_TRY(cudaEventCreate(&e0));
    G_TRY(cudaEventCreate(&e1));
    m->launches = 0;
    rc = forward_device(m, N, E, nullptr, nullptr, d_nt.as<uint8_t>(), d_et.as<uint8_t>(), d_send.as<int32_t>(), d_recv.as<int32_t>(),
                        d_sptr.as<int64_t>(), d_sids.as<int32_t>(), d_rptr.as<int64_t>(), d_rids.as<int32_t>(), d_dec.as<float>(), e0, e1);
    if (!rc) {
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (kernel_ms) *kernel_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (decoded && N) G_TRY(cudaMemcpy(decoded, d_dec.p, N * sizeof(float), cudaMemcpyDeviceToHost));
    if (prob && V) {
        DevBuf d_v0, d_p;
        if ((rc = to_device(var_node0, d_v0))) return rc;
        G_TRY(dmalloc(&d_p.p, V * sizeof(float)));
        pair_prob_kernel<<<(unsigned)((V + 255) / 256), 256>>>(d_dec.as<float>(), d_v0.as<int64_t>(), d_p.as<float>(), V);
        m->launches++;
        G_TRY(cudaGetLastError());
        G_TRY(cudaMemcpy(prob, d_p.p, V * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return SLS_OK;
}

}  // namespace

extern "C" int sls_gnn_launch_count(const sls_gnn_model* m) { return m ? m->launches : 0; }

extern "C" int sls_gnn_set_profile(sls_gnn_model* m, int on)
{
    if (!m) return sls_set_error_internal(SLS_ERR_ARG, "NULL argument");
    m->profile = on != 0;
    return SLS_OK;
}

extern "C" int sls_gnn_get_profile(const sls_gnn_model* m, double ms[4])
{
    if (!m || !ms) return sls_set_error_internal(SLS_ERR_ARG, "NULL argument");
    for (int i = 0; i < 4; i++) ms[i] = m->prof_ms[i];
    return SLS_OK;
}
