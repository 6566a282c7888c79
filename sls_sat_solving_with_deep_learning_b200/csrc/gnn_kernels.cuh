// gnn_kernels.cuh -- GNN oracle forward (LCG + interaction network) on sm_100a.
//
// Device restatement of the reference's oracle forward:
//   python/src/model.py:11-18    get_embedding (Linear 2->32 on nodes and edges)        -> embed_kernel
//   python/src/model.py:21-61    apply_interaction: per step, jraph.InteractionNetwork with
//                                update_fn = LayerNorm(ReLU(Linear(ReLU(Linear(concat))))) -> mlp_ln_kernel
//                                and segment_sum aggregation of the NEW edges per sender /
//                                receiver                                                -> segment_sum_kernel
//   python/src/model.py:144      readout Linear(H->1)                                     -> readout_kernel
//   python/src/sat_representations.py:775-780 LCG.get_model_probabilities (pair softmax)  -> pair_prob_kernel
//
// The two dense contractions of every update (K1 x H and H x H, K1 = 96 / 432 / 600, H = 200)
// run on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulators in TMEM,
// operands staged in shared memory in the canonical K-major SWIZZLE_128B layout (weights by TMA
// bulk copies of pre-swizzled slab images, activations by the threads), completion tracked with
// mbarriers via tcgen05.commit.  One CTA owns a tile of 128 rows (edges or nodes) and
// runs the whole update for it: gather -> GEMM1 -> bias+ReLU -> GEMM2 -> bias+ReLU -> LayerNorm ->
// global.  Nothing but the tile's inputs and its normalised output touches HBM; the hidden
// activation goes TMEM -> registers -> shared memory slab by slab as GEMM2 consumes it.
//
// Precision: a single TF32 product (10-bit mantissa) accumulates to |dp| = 1.7e-3 over the 20
// dense layers, above the 1e-3 contract, so every contraction is error-compensated ("3xTF32"):
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo
// (three MMAs into the same fp32 accumulator; the dropped lo*lo term is 2^-22 relative).  Weights
// are split on the host, activations while they are staged.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace gnn {

static constexpr int TILE_M = 128;       // rows per CTA = UMMA M
static constexpr int SLAB_K = 32;        // tf32 elements per 128-byte swizzle row
static constexpr int MAX_N = 208;        // padded output width (multiple of 16, >= hidden)
static constexpr int B_SLAB_BYTES = MAX_N * 128;    // 26 KB: one weight slab image (one 128-byte row per output column)
static constexpr int B_STAGE_BYTES = 2 * B_SLAB_BYTES;   // 52 KB: the two pieces of a slab (hi | lo, or b0 | b1)
// One mbarrier per unit index mod DONE_BARS, waited for by PARITY.  A parity wait for completion k of a barrier is sound only
// while the barrier stands at k or k + 1 completions: (a) the unit DONE_BARS before the awaited one must be known complete,
// (b) the unit DONE_BARS after it must be unable to complete before the waiter moves on.  The producers skip the GEMM2 units
// of a tile (the epilogue warps stage those), i.e. up to 2 * ceil(MAX_N / 32) = 14 units between two of their waits, and the
// issuer runs through those units without them: (a) needs DONE_BARS >= 14 + 2; (b) holds because the awaited unit's third
// successor is the one the waiter is about to stage.  (With 6 barriers a producer that reached its wait late found the
// barrier two completions ahead and slept until a completion that depended on itself.)
static constexpr int DONE_BARS = 16;
static constexpr int TMEM_COLS = 512;
static constexpr int PAR_BYTES = 3 * MAX_N * 4;          // b2 | ln_scale | ln_offset staged in shared memory
static constexpr int BAR_BYTES = 384;                    // mbarriers + the TMEM base address

struct MlpArgs {
    // up to three gathered input segments: row i of the A matrix is
    //   [ src0[idx0 ? idx0[i] : i][0:d0] | src1[idx1 ? idx1[i] : i][0:d1] | src2[...][0:d2] ]
    const float* src[3];
    const int32_t* idx[3];
    int32_t d[3];
    int32_t rows;        // number of rows (edges or nodes)
    int32_t k1_slabs;    // ceil((d0+d1+d2)/32)
    int32_t k2_slabs;    // ceil(h1/32)
    int32_t n1, n2;      // padded output widths of the two linears (multiples of 16, <= 208)
    int32_t h2;          // true output width (LayerNorm length)
    int32_t h1;          // true hidden width (the bf16 images carry b2 as weight row h1, b1 as weight row d0+d1+d2: see pack_mlp)
    // weights as shared-memory slab images: [k_slabs][n_pad * 32] floats, slab s holding W^T[n][32 s .. 32 s + 31]
    // already in the K-major SWIZZLE_128B layout (sw128_off), so one bulk copy per slab stages it
    const float* w1_hi;  // TF32-rounded weights
    const float* w1_lo;  // TF32-rounded residual w - hi
    const float* b1;     // [n1] zero padded
    const float* w2_hi;
    const float* w2_lo;
    const float* b2;     // [n2]
    const float* ln_scale;   // [h2]
    const float* ln_offset;  // [h2]
    float* out;          // [rows][h2]
    // the same weights as two bf16 pieces (w = b0 + b1), slab images of 64 k-values per 128-byte row, for the kind::f16 kernel
    const uint16_t* w1_b0;
    const uint16_t* w1_b1;
    const uint16_t* w2_b0;
    const uint16_t* w2_b1;
    int32_t k1_slabs_bf;  // ceil((d0+d1+d2)/64)
    int32_t k2_slabs_bf;  // ceil(h1/64)
    // development builds (-DGNN_TRACE): clock64 stamps of CTA 0's roles, see tools/gnn_trace.py; nullptr otherwise
    unsigned long long* trace;
};

#ifdef GNN_TRACE
#define GNN_STAMP(slot) do { if (P.trace && blockIdx.x == 0) P.trace[(slot)] = (unsigned long long)clock64(); } while (0)
#else
#define GNN_STAMP(slot) do { } while (0)
#endif
#ifndef GNN_TRACE_TILE0
#define GNN_TRACE_TILE0 3
#endif
static constexpr int TRACE_TILE0 = GNN_TRACE_TILE0, TRACE_TILES = 3;   // tiles of CTA 0 that are traced (-DGNN_TRACE_TILE0=50: mid-launch)
static constexpr int TRACE_ISSUER = 0, TRACE_PROD = 4096, TRACE_EPI = 8192, TRACE_WORDS = 8192 + 64;

// ---- small PTX helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{ .reg .pred p;\n"
        "WAIT_%=:\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "  @!p bra WAIT_%=;\n }"
        ::"r"(bar), "r"(parity) : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{ .reg .pred q; elect.sync _|q, 0xffffffff; selp.u32 %0, 1, 0, q; }" : "=r"(p));
    return p != 0;
}
// the same for a long wait (the epilogue warps wait a whole tile): sleep between polls instead of taking issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{ .reg .pred p;\n"
        "WAITR_%=:\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "  @p bra DONER_%=;\n"
        "  nanosleep.u32 256;\n"
        "  bra WAITR_%=;\n"
        "DONER_%=:\n }"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (contiguous bytes), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
// tcgen05.commit: the mbarrier is arrived on when all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrival delivered to the barrier at this offset in every CTA of the cluster named by cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask) : "memory");
}
// TMA bulk copy global -> the same shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier (same
// offset) counts the bytes
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{ .reg .pred p;\n"
        "  setp.ne.b32 p, %4, 0;\n"
        "  tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte
// offset = 1024 B between 8-row groups in [32,46), version 1 in [46,48), layout type 2 in [61,64)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// byte offset of (row r, 16-byte unit u) inside a K-major SWIZZLE_128B slab: rows are 128 B apart,
// 8-row groups 1024 B apart, the unit index is XORed with r % 8 (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t u) { return (r >> 3) * 1024u + (r & 7u) * 128u + ((u ^ (r & 7u)) << 4); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait (the caller overlaps it with other work and calls tmem_wait_ld before touching v), and its
// 32-column form
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
          "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]),
          "=f"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Round to the nearest TF32 value (10-bit mantissa), ties away from zero like cvt.rna.tf32.f32.  The tensor core ignores
// the low 13 mantissa bits of an fp32 operand, so operands are rounded explicitly before they are staged.  sm_100 has no
// conversion instruction for this: ptxas expands cvt.rna.tf32.f32 into the same add-and-mask plus an |x| < inf guard
// (four instructions); activations and residuals here are finite, and inf maps to inf without the guard anyway, so the
// two-instruction form is used -- the staging threads, not the tensor pipe, set the pace of the MLP kernel.
__device__ __forceinline__ float round_tf32(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// The residual x - hi of the 3xTF32 product.  It is exact in fp32 (at most 13 significant bits) and is handed to the tensor
// core as it is: the hardware drops its low mantissa bits, an error of 2^-11 of the residual = 2^-22 of x, below the fp32
// rounding of the accumulation itself -- and one instruction instead of three on the staging threads' path.
__device__ __forceinline__ float residual_tf32(float x, float hi) { return x - hi; }

// ---- the fused update: LayerNorm(ReLU(Linear2(ReLU(Linear1(gathered rows))))) -------------------
// One persistent CTA per SM walks over tiles of 128 rows (tile = blockIdx.x, += gridDim.x).  Roles:
//   warps 0-7  producers: stage the A operand unit by unit (GEMM1: gathered rows; GEMM2: ReLU(acc1 + b1) out of TMEM);
//   warp  8    one elected lane issues the weight copies (TMA bulk) and the MMAs;
//   warps 9-12 epilogue: LayerNorm(ReLU(acc2 + b2)) out of TMEM straight to global memory.
// Slabs are numbered globally (J) across tiles, so stage and barrier parities simply run on; two barriers hand acc2
// back and forth:
//   sAccFull (tcgen05.commit after the last slab of a tile) -> epilogue may read acc2,
//   sAccFree (128 epilogue threads, after their last TMEM read) -> the issuer may overwrite acc2 (first GEMM2 MMA of
//   the next tile).  acc1 needs no barrier: the producers read it (GEMM2 staging of tile t) before they stage, in
//   program order, the GEMM1 slabs whose MMAs overwrite it.
// (Round 1 also shipped two predecessors of this kernel -- one tile per CTA, and a persistent kernel with the A operand
// staged in shared memory -- behind environment switches; they are retired, see git history and profiles/r01_gnn_*.)
static constexpr int MLPP_PRODUCERS = 256;
static constexpr int MLPP_EPILOGUE0 = MLPP_PRODUCERS + 32;   // first epilogue thread
static constexpr int MLPP_ISSUER_WARP = MLPP_PRODUCERS / 32;
// warps 13-15: normalisers (pass B of the epilogue: no tensor memory involved, so any warp can do it).  13 warps are allocated
// as 16 anyway (the 128-register cap comes from there), so the three extra warps cost nothing.
static constexpr int MLPP_NORM0 = MLPP_EPILOGUE0 + 128;      // first normaliser thread
static constexpr int MLPP_NORM_WARPS = 3;
static constexpr int MLPP_THREADS = MLPP_NORM0 + 32 * MLPP_NORM_WARPS;

// ---- the same update with the A operand in TENSOR MEMORY ---------------------------------------------------------
// ncu of the kernel above (profiles/r01_gnn_mlp_ln_summary.md): the shared-memory data pipe (128 B per clock and SM) is what
// the 3xTF32 product saturates, not the tensor pipe -- per slab the twelve MMAs read 12 x (4 KB of A + 6.6 KB of B), the
// producers write 32 KB of A and the copy engine 52 KB of B: 1.5x what the pipe moves in the 1160 cycles the MMAs need, so
// the tensor pipe cannot be more than ~65 % busy however the roles are balanced (measured 47-50 %; cheaper staging
// code, deeper prefetch and multicast weight slabs all left the time unchanged).  Here the A operand never touches
// shared memory: the producers write the split rows with tcgen05.st into three 32-column stages next to the two
// accumulators (acc1 columns 0-207, acc2 208-415, A stages 416-511: 16 hi + 16 lo columns = 16 of a slab's 32 k-values)
// and the MMAs take A from tensor memory (tcgen05.mma [d], [a], b-desc).  Shared memory then only carries the weight
// slabs: 80 KB of B reads + 52 KB of B writes per slab, which fits under the MMA time, and the 64 KB of A stages become a
// fourth weight stage.
// A "unit" is half a slab (16 k-values, six MMAs).  Producer group g (warps 4g .. 4g+3, thread = row) stages the units
// 2J + g, i.e. the columns it already gathers; unit U lives in A stage U % 3 and is released by its own tcgen05.commit
// (sDone[U % 6]).
static constexpr int TS_B_STAGES = 4;
static constexpr int TS_A_STAGES = 3;
static constexpr int TS_A_COL = 416;     // first A-stage column; 32 columns per stage
static constexpr int TS_ACC2_COL = 208;
static constexpr int TS_OUT_BYTES = 4 * 32 * 64;  // epilogue staging: per warp 32 rows x 16 columns, see the store pass
static constexpr int TS_STAT_BYTES = 2 * TILE_M * 8;  // per tile parity: (rstd, -mean * rstd) of every row, epilogue -> normalisers
static constexpr int TS_SMEM_BYTES = TS_B_STAGES * B_STAGE_BYTES + BAR_BYTES + PAR_BYTES + MAX_N * 4 + TS_OUT_BYTES + TS_STAT_BYTES + 1024;  // + b1

// instruction descriptor for kind::f16 with bf16 operands: D = f32, A = B = bf16 (format 1), both K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f16 with fp16 operands: a_format = b_format = 0 (F16), fp32 accumulator
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
// x = a0 + a1 for a pair of k-values, packed two to a 32-bit column (the lower k in the low half): a0 = round(x), a1 =
// round(x - a0), both to nearest even; one packing conversion per pair and piece.  fp16 saturates at the largest finite value
// instead of producing infinities (the residual then carries what is left, exact up to twice the range).
template <bool F16>
__device__ __forceinline__ void split_pair16(float lo_k, float hi_k, uint32_t& a0, uint32_t& a1)
{
    if constexpr (F16) {
        float f_lo, f_hi;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(a0) : "f"(hi_k), "f"(lo_k));
        asm("{ .reg .b16 l, h;\n  mov.b32 {l, h}, %2;\n  cvt.f32.f16 %0, l;\n  cvt.f32.f16 %1, h; }" : "=f"(f_lo), "=f"(f_hi) : "r"(a0));
        const float re = lo_k - f_lo, ro = hi_k - f_hi;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(a1) : "f"(ro), "f"(re));
    } else {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a0) : "f"(hi_k), "f"(lo_k));
        const float re = lo_k - __uint_as_float(a0 << 16), ro = hi_k - __uint_as_float(a0 & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a1) : "f"(ro), "f"(re));
    }
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{ .reg .pred p;\n"
        "  setp.ne.b32 p, %4, 0;\n"
        "  tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{ .reg .pred p;\n"
        "  setp.ne.b32 p, %4, 0;\n"
        "  tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 TMEM lanes starting at the lane in taddr, 32 columns: per 8-column repeat c (registers 4c .. 4c+3) lane l supplies
// {row l / 4: columns 8c + 2 (l % 4), + 1;  row l / 4 + 8: the same columns}   (cute SM100_TMEM_STORE_16dp256b4x).  No wait.
__device__ __forceinline__ void tmem_st16x256b_x4(uint32_t taddr, const uint32_t* r)
{
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// BF = true: the products are formed from two bf16 pieces per operand on kind::f16 (x = a0 + a1, w = b0 + b1; a0*b0 + a1*b0 +
// a0*b1, fp32 accumulation) instead of 3xTF32.  A 128-byte weight row then holds 64 k-values and a unit 32 (two per TMEM
// column), so every MMA covers twice the k-range in the same time: half the MMAs, weight bytes and barrier hand-overs per
// tile.  16 mantissa bits per operand: max |dp| vs fp64 3.4e-5 in emulation, 2.7e-5 measured (tools/precision_split_study.py; single-pass
// TF32 1.75e-3, 3xTF32 1.9e-6), against the 1e-3 contract.
// CL = 2: the two CTAs of a cluster run their tiles in lock step; each requests half of every weight slab and the copy engine
// multicasts it into the same stage of both (as in mlp_ln_persistent_kernel<CL>): the weight stream, 70 % of this kernel's
// L2 -> SM traffic, is read from L2 once per pair.  A stage may be refilled when both CTAs' MMAs have drained it (sFreeB).
// PREC: 0 = 3xTF32 (kind::tf32), 1 = two bf16 pieces per operand, 2 = two fp16 pieces per operand (both kind::f16, same
// staging and images; only the conversions and the instruction descriptor's operand formats differ)
template <int PREC, int CL>
__global__ void __launch_bounds__(MLPP_THREADS, 1) mlp_ln_persistent_ts_kernel(const MlpArgs P)
{
    static_assert(CL == 1 || CL == 2, "cluster size");
    static_assert(PREC >= 0 && PREC <= 2, "precision");
    constexpr bool BF = PREC != 0;   // "two 16-bit pieces" path
    constexpr bool F16 = PREC == 2;
    constexpr int FREE_BARS = 2 * TS_B_STAGES;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
    constexpr int NV = BF ? 32 : 16;        // k-values per thread and unit
    constexpr int NQ = NV / 4;              // float4 per thread and slab
    constexpr int SLABK = 2 * NV;           // k-values per weight slab (one 128-byte row)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sBst = base;
    const uint32_t sFull = sBst + TS_B_STAGES * B_STAGE_BYTES;
    const uint32_t sDone = sFull + 8 * TS_B_STAGES;
    const uint32_t sArdy = sDone + 8 * DONE_BARS;
    const uint32_t sAccFull = sArdy + 8 * TS_A_STAGES;
    const uint32_t sAccFree = sAccFull + 8;
    const uint32_t sFreeB = sAccFree + 8;
    const uint32_t sStatF = sFreeB + 8 * FREE_BARS;   // [2] a tile's row statistics are in shared memory (4 epilogue warps arrive)
    const uint32_t sStatE = sStatF + 16;              // [2] ... and have been consumed (3 normaliser warps arrive)
    const uint32_t sAcc1Full = sStatE + 16;           // every GEMM1 MMA of the tile has completed (tcgen05.commit): acc1 may be read
    const uint32_t sTmem = sAcc1Full + 8;
    const uint32_t sPar = sFull + BAR_BYTES;
    const uint32_t sOut = sPar + PAR_BYTES + MAX_N * 4;
    const uint32_t sStat = sOut + TS_OUT_BYTES;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const bool producer = tid < MLPP_PRODUCERS;
    const bool epilogue = tid >= MLPP_EPILOGUE0 && tid < MLPP_EPILOGUE0 + 128;
    const int prow = tid & (TILE_M - 1);   // producers: row of the tile = TMEM lane
    const uint32_t grp = (uint32_t)tid >> 7;  // producers: which 16 of a slab's 32 columns, i.e. which units
    const int ntiles = (P.rows + TILE_M - 1) / TILE_M;
    // in a cluster every CTA runs the tile count of the first one (tiles past the end gather the last row and store nothing)
    const uint32_t my_tiles = CL > 1 ? (uint32_t)((ntiles + (int)gridDim.x - 1) / (int)gridDim.x)
                                     : ((int)blockIdx.x < ntiles ? (uint32_t)((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) : 0u);

    if (tid == 0) {
        for (int i = 0; i < TS_B_STAGES; i++) mbar_init(sFull + 8 * i, 1);
        for (int i = 0; i < DONE_BARS; i++) mbar_init(sDone + 8 * i, 1);
        for (int i = 0; i < TS_A_STAGES; i++) mbar_init(sArdy + 8 * i, MLPP_PRODUCERS / 64);  // the four warps of one group
        mbar_init(sAccFull, 1);
        mbar_init(sAccFree, 128);
        mbar_init(sAcc1Full, 1);
        for (int i = 0; i < 2; i++) mbar_init(sStatF + 8 * i, 4), mbar_init(sStatE + 8 * i, MLPP_NORM_WARPS);
        for (int i = 0; i < FREE_BARS; i++) mbar_init(sFreeB + 8 * i, CL);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < MAX_N; i += MLPP_THREADS) {
        const bool in = i < P.h2;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPar + i * 4), "f"(i < P.n2 ? __ldg(P.b2 + i) : 0.f) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPar + (MAX_N + i) * 4), "f"(in ? __ldg(P.ln_scale + i) : 0.f) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPar + (2 * MAX_N + i) * 4), "f"(in ? __ldg(P.ln_offset + i) : 0.f) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPar + (3 * MAX_N + i) * 4), "f"(i < P.n1 ? __ldg(P.b1 + i) : 0.f) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no multicast copy or commit may reach a CTA whose barriers are not initialised yet
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + (sTmem - base));
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t n1s = (uint32_t)(BF ? P.k1_slabs_bf : P.k1_slabs), n2s = (uint32_t)(BF ? P.k2_slabs_bf : P.k2_slabs), S = n1s + n2s;
    auto wait_unit = [&](uint32_t U) { mbar_wait(sDone + 8 * (U % DONE_BARS), (U / DONE_BARS) & 1u); };

    if (producer) {
        const int kA = P.d[0] + P.d[1] + P.d[2];
        // segment widths that are multiples of 8 floats (every shipped layer: 32 and 200): a 32-byte piece of a row never
        // straddles two segments and is fetched with one 256-bit load
        const bool wide = ((P.d[0] | P.d[1] | P.d[2]) & 7) == 0;
        // ---- GEMM1 operand, two gather geometries ------------------------------------------------------------------
        // BF (default): FOUR LANES PER ROW.  Lane (r8 = lane / 4, jq = lane % 4) of producer warp w fetches, for its rows
        // 32 (w % 4) + r8 + 8 i (i = 0..3), the 32 bytes [8 jq, 8 jq + 8) of the unit's 32 k-values: a load instruction of
        // the warp touches 8 rows x 128 contiguous bytes = 8 lines.  With thread = row (the tf32 path below) it touched 32
        // lines for the same bytes, and the L1 tag stage -- one line per cycle -- was what paced the kernel (~1000 tag
        // cycles per slab next to 1250 MMA cycles).  The quads match tcgen05.st's 16x256b shape (lane l holds two 32-bit
        // columns of rows l / 4 and l / 4 + 8 per 8-column repeat), so a unit goes to tensor memory with two stores of 16
        // rows; the k-values inside a unit land in a permuted column order, which the host bakes into the W1 slab images
        // (gnn_api.cu: w1_perm) -- the contraction does not care in which order k is summed as long as A and B agree.
        // !BF (tf32x3): thread = row, 32x32b stores (unchanged).
        const int lane = tid & 31;
        const int jq = lane & 3, r8 = lane >> 2;
        const int qrow0 = (warp & 3) * 32 + r8;  // BF: first of this lane's four rows inside the tile
        const float* rp[3] = {nullptr, nullptr, nullptr};   // !BF: this thread's row in each source
        int nidx[3] = {0, 0, 0};
        int qidx[3][4], qnext[3][4];                        // BF: row indices of the current / next tile (sources x rows)
        // Row indices are read one tile ahead of the rows they address: at the GEMM1 -> GEMM2 hand-over the pointers of
        // the next tile are formed from registers, not from a fresh load on the drain's critical path.
        auto fetch_idx = [&](int tile) {
            if constexpr (BF) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int my_row = min(tile * TILE_M + qrow0 + 8 * i, P.rows - 1);  // rows past the end re-read the last row; never stored
#pragma unroll
                    for (int j = 0; j < 3; j++) qnext[j][i] = (P.d[j] > 0 && P.idx[j]) ? __ldg(P.idx[j] + my_row) : my_row;
                }
            } else {
                const int my_row = min(tile * TILE_M + prow, P.rows - 1);
#pragma unroll
                for (int j = 0; j < 3; j++) nidx[j] = (P.d[j] > 0 && P.idx[j]) ? __ldg(P.idx[j] + my_row) : my_row;
            }
        };
        auto set_rows = [&]() {  // the tile whose indices were fetched last becomes the current one
            if constexpr (BF) {
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        qidx[j][i] = qnext[j][i];
                        // The row is gathered unit by unit over the whole tile; ask L2 for all of it now (the eight lanes
                        // that share a row -- four per producer group -- take alternate 128-byte lines).  Gathered rows come
                        // from HBM (the feature arrays are several times the L2).
                        if (P.d[j] > 0) {
                            const float* row = P.src[j] + (size_t)qidx[j][i] * P.d[j];
                            for (int o = ((int)grp * 4 + jq) * 32; o < P.d[j]; o += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                        }
                    }
            } else {
#pragma unroll
                for (int j = 0; j < 3; j++)
                    if (P.d[j] > 0) {
                        rp[j] = P.src[j] + (size_t)nidx[j] * P.d[j];
                        for (int o = (int)grp * 32; o < P.d[j]; o += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp[j] + o));
                    }
            }
        };
        // BF: x[i] = the 8 k-values [8 jq, 8 jq + 8) of unit (slab s, group grp) in row i of this lane
        auto load_q = [&](int s, float (*x)[8]) {
            const int col = s * SLABK + (int)grp * NV + 8 * jq;
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int e = 0; e < 8; e++) x[i][e] = 0.f;
                if (col >= kA) {
                    if (col == kA) x[i][0] = 1.f;  // the constant-one column: W1's image holds b1 in this row (wide: kA % 8 == 0)
                    continue;
                }
                if (wide) {
                    const int sj = col < P.d[0] ? 0 : col < P.d[0] + P.d[1] ? 1 : 2;
                    const int c = col - (sj == 0 ? 0 : sj == 1 ? P.d[0] : P.d[0] + P.d[1]);
                    const int ri = sj == 0 ? qidx[0][i] : sj == 1 ? qidx[1][i] : qidx[2][i];
                    const float* p = P.src[sj] + (size_t)ri * P.d[sj] + c;
                    // no L1 allocation: each byte is read once, by one thread
                    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=f"(x[i][0]), "=f"(x[i][1]), "=f"(x[i][2]), "=f"(x[i][3]), "=f"(x[i][4]), "=f"(x[i][5]), "=f"(x[i][6]),
                                   "=f"(x[i][7])
                                 : "l"(p));
                } else {  // odd widths: element by element (a piece may straddle two segments)
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int ce = col + e;
                        if (ce < kA) {
                            const int sj = ce < P.d[0] ? 0 : ce < P.d[0] + P.d[1] ? 1 : 2;
                            const int c = ce - (sj == 0 ? 0 : sj == 1 ? P.d[0] : P.d[0] + P.d[1]);
                            const int ri = sj == 0 ? qidx[0][i] : sj == 1 ? qidx[1][i] : qidx[2][i];
                            x[i][e] = __ldg(P.src[sj] + (size_t)ri * P.d[sj] + c);
                        } else if (ce == kA) {
                            x[i][e] = 1.f;
                        }
                    }
                }
            }
        };
        auto load_a1 = [&](int s, float4* x) {  // !BF: thread = row
            if (wide) {
#pragma unroll
                for (int u = 0; u < NQ; u += 2) {
                    const int col = s * SLABK + (int)grp * NV + 4 * u;
                    x[u] = x[u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col < kA) {
                        const float* p = col < P.d[0] ? rp[0] + col : col < P.d[0] + P.d[1] ? rp[1] + (col - P.d[0]) : rp[2] + (col - P.d[0] - P.d[1]);
                        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w), "=f"(x[u + 1].x), "=f"(x[u + 1].y),
                                       "=f"(x[u + 1].z), "=f"(x[u + 1].w)
                                     : "l"(p));
                    }
                }
                return;
            }
#pragma unroll
            for (int u = 0; u < NQ; u++) {
                const int col = s * SLABK + (int)grp * NV + 4 * u;
                x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col < kA) {
                    const float* p = col < P.d[0] ? rp[0] + col : col < P.d[0] + P.d[1] ? rp[1] + (col - P.d[0]) : rp[2] + (col - P.d[0] - P.d[1]);
                    x[u] = __ldg(reinterpret_cast<const float4*>(p));
                }
            }
        };
        // v[NV] -> the 32 columns of A stage U % 3: 16 hi | 16 lo (tf32), or 16 columns of a0 pairs | 16 of a1 pairs (bf16: the
        // lower k in the low half of a column; both pieces rounded to nearest even)
        // The split is register work and runs BEFORE the thread waits for its A stage; once the stage is free only the TMEM
        // store, its wait and the barrier arrival are left on the stage's turn-around path.
        auto split_pair = [](float lo_k, float hi_k, uint32_t& a0, uint32_t& a1) { split_pair16<F16>(lo_k, hi_k, a0, a1); };
        auto pack_unit = [&](const float* v, uint32_t* r) {
            if constexpr (BF) {
#pragma unroll
                for (int i = 0; i < 16; i++) split_pair(v[2 * i], v[2 * i + 1], r[i], r[16 + i]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float hi = round_tf32(v[i]);
                    r[i] = __float_as_uint(hi);
                    r[16 + i] = __float_as_uint(residual_tf32(v[i], hi));
                }
            }
        };
        // BF, GEMM1: the lane's four rows x four pairs -> register order of two tcgen05.st.16x256b.x4 (rows i = 0,1 go to TMEM
        // lanes r8, r8 + 8 of the warp's quadrant, rows i = 2,3 to lanes 16 + r8, 24 + r8).  Per 8-column repeat c a lane
        // supplies {row lo: columns 2 jq, 2 jq + 1; row hi: the same columns}; repeats 0-1 are the a0 block, 2-3 the a1 block.
        auto pack_quads = [&](const float (*x)[8], uint32_t* r) {
#pragma unroll
            for (int hs = 0; hs < 2; hs++)
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int i = 2 * hs + rr;
#pragma unroll
                    for (int pp = 0; pp < 4; pp++) {
                        uint32_t a0, a1;
                        split_pair(x[i][2 * pp], x[i][2 * pp + 1], a0, a1);
                        const int c = pp >> 1, b = pp & 1;
                        r[16 * hs + 4 * c + 2 * rr + b] = a0;
                        r[16 * hs + 8 + 4 * c + 2 * rr + b] = a1;
                    }
                }
        };
        auto store_unit = [&](uint32_t U, const uint32_t* r) {
            if (U >= (uint32_t)TS_A_STAGES) wait_unit(U - TS_A_STAGES);
            tc_fence_after();
            tmem_st32(tmem + lane_base + (uint32_t)(TS_A_COL + 32 * (U % TS_A_STAGES)), r);
        };
        auto store_quads = [&](uint32_t U, const uint32_t* r) {
            if (U >= (uint32_t)TS_A_STAGES) wait_unit(U - TS_A_STAGES);
            tc_fence_after();
            const uint32_t a = tmem + lane_base + (uint32_t)(TS_A_COL + 32 * (U % TS_A_STAGES));
            tmem_st16x256b_x4(a, r);
            tmem_st16x256b_x4(a + (16u << 16), r + 16);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        // tf32: even slabs of a tile in xa0, odd ones in xa1 (two slabs of lead); bf16: a slab lasts twice as long and holds
        // twice the registers, one buffer and one slab of lead
        constexpr int LEAD = BF ? 1 : 2;
        float4 xa0[BF ? 1 : NQ], xa1[BF ? 1 : NQ];
        float xq[BF ? 4 : 1][8];
        uint32_t J = 0;
        auto first_loads = [&]() {
            if constexpr (BF) {
                load_q(0, xq);
            } else {
                load_a1(0, xa0);
                load_a1(1, xa1);
            }
        };
        if (my_tiles > 0) {
            fetch_idx(blockIdx.x);
            set_rows();
            first_loads();
            fetch_idx((int)(blockIdx.x + gridDim.x));
        }
        uint32_t trace_it = 0;
        auto publish = [&](uint32_t U) {  // the unit is in its A stage: hand it to the issuer
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(sArdy + 8 * (U % TS_A_STAGES));
#ifdef GNN_TRACE
            if ((tid & 127) == 0 && trace_it >= (uint32_t)TRACE_TILE0 && trace_it < (uint32_t)(TRACE_TILE0 + TRACE_TILES))
                GNN_STAMP(TRACE_PROD + U - 2 * TRACE_TILE0 * S);
#endif
            J++;
        };
        // GEMM2's operand comes out of GEMM1's accumulator: NV columns of this thread's TMEM lane, requested without waiting
        // (zeros beyond the padded width)
        auto request_acc1 = [&](uint32_t j2, float* vg) {
            const int c0 = (int)j2 * SLABK + (int)grp * NV;
            if constexpr (BF) {
                if (c0 + 32 <= P.n1)
                    tmem_ld32_nowait(tmem + lane_base + (uint32_t)c0, vg);
                else if (c0 < P.n1)
                    tmem_ld16_nowait(tmem + lane_base + (uint32_t)c0, vg);
            } else {
                if (c0 < P.n1) tmem_ld16_nowait(tmem + lane_base + (uint32_t)c0, vg);
            }
        };
        for (uint32_t it = 0; it < my_tiles; it++) {
            const int next_tile = (int)(blockIdx.x + (it + 1) * gridDim.x);
            trace_it = it;
            // ---- GEMM1: gathered rows
            auto gemm1_unit = [&](uint32_t j, float4* xa) {
                const uint32_t U = 2 * J + grp;
                uint32_t r[32];
                if constexpr (BF) {
                    pack_quads(xq, r);
                    if (j + LEAD < n1s) load_q((int)j + LEAD, xq);
                    store_quads(U, r);
                } else {
                    float v[NV];
#pragma unroll
                    for (int u = 0; u < NQ; u++) v[4 * u] = xa[u].x, v[4 * u + 1] = xa[u].y, v[4 * u + 2] = xa[u].z, v[4 * u + 3] = xa[u].w;
                    pack_unit(v, r);
                    if (j + LEAD < n1s) load_a1((int)j + LEAD, xa);
                    store_unit(U, r);
                }
                publish(U);
            };
            if constexpr (BF) {
                for (uint32_t j = 0; j < n1s; j++) gemm1_unit(j, xa0);
            } else {
                for (uint32_t j = 0; j < n1s; j += 2) {
                    gemm1_unit(j, xa0);
                    if (j + 1 < n1s) gemm1_unit(j + 1, xa1);
                }
            }
            // ---- GEMM2's operand (ReLU of the first accumulator) is staged by the EPILOGUE warps (see there): the producers go
            // straight on to the next tile -- row pointers, first gathers, indices one tile further ahead -- and their unit
            // counter skips the GEMM2 units
            if (it + 1 < my_tiles) {
                set_rows();
                first_loads();
                fetch_idx(next_tile + (int)gridDim.x);
                // the next unit of this group waits for the A stage of a GEMM2 unit like for any other (DONE_BARS above)
            }
            J += n2s;
        }
    } else if (warp_u == MLPP_ISSUER_WARP) {
        const bool leader = elect_one();
        const uint32_t total = my_tiles * S;
        const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
        constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version 1 | SWIZZLE_128B
        auto desc = [&](uint32_t addr) -> uint64_t { return (uint64_t)DESC_HI << 32 | (((addr >> 4) & 0x3fffu) | (1u << 16)); };
        const uint32_t idesc1 = F16 ? idesc_f16(TILE_M, P.n1) : BF ? idesc_bf16(TILE_M, P.n1) : idesc_tf32(TILE_M, P.n1);
        const uint32_t idesc2 = F16 ? idesc_f16(TILE_M, P.n2) : BF ? idesc_bf16(TILE_M, P.n2) : idesc_tf32(TILE_M, P.n2);
        auto issue_b = [&](uint32_t J, uint32_t j) {  // j = J % S
            const bool g1 = j < n1s;
            const int n_pad = g1 ? P.n1 : P.n2;
            const uint32_t s = g1 ? j : j - n1s, bbytes = (uint32_t)n_pad * 128u;  // one 128-byte row per output column, either format
            const uint8_t* hi = reinterpret_cast<const uint8_t*>(BF ? (const void*)(g1 ? P.w1_b0 : P.w2_b0) : (const void*)(g1 ? P.w1_hi : P.w2_hi)) + (size_t)s * bbytes;
            const uint8_t* lo = reinterpret_cast<const uint8_t*>(BF ? (const void*)(g1 ? P.w1_b1 : P.w2_b1) : (const void*)(g1 ? P.w1_lo : P.w2_lo)) + (size_t)s * bbytes;
            const uint32_t bs = sBst + (J % TS_B_STAGES) * B_STAGE_BYTES, bar = sFull + 8 * (J % TS_B_STAGES);
            if (CL == 1) {
                if (J >= (uint32_t)TS_B_STAGES) wait_unit(2 * (J - TS_B_STAGES) + 1);  // both units of the stage's last slab are done
                if (leader) {
                    mbar_expect_tx(bar, 2 * bbytes);
                    bulk_g2s(bs, hi, bbytes, bar);
                    bulk_g2s(bs + B_SLAB_BYTES, lo, bbytes, bar);
                }
            } else {
                if (J >= (uint32_t)TS_B_STAGES) {  // ... in every CTA of the cluster
                    const uint32_t Jf = J - TS_B_STAGES;
                    mbar_wait(sFreeB + 8 * (Jf % FREE_BARS), (Jf / FREE_BARS) & 1u);
                }
                if (leader) {
                    const uint32_t part = bbytes / CL, off = crank * part;  // n_pad * 128 / CL: a multiple of 1024
                    mbar_expect_tx(bar, 2 * bbytes);  // own part + the part the other CTA sends
                    bulk_g2s_multicast(bs + off, hi + off, part, bar, CL_MASK);
                    bulk_g2s_multicast(bs + B_SLAB_BYTES + off, lo + off, part, bar, CL_MASK);
                }
            }
        };
        // weights travel PF slabs ahead; the stage a request overwrites was read two slabs ago, so the issuer does not
        // stall on the slab it has just queued
        constexpr uint32_t PF = TS_B_STAGES - 2;
        uint32_t jn = 0;
        for (uint32_t J = 0; J < PF && J < total; J++) {
            issue_b(J, jn);
            jn = jn + 1 == S ? 0 : jn + 1;
        }
        uint32_t j = 0, it = 0;
        for (uint32_t J = 0; J < total; J++) {
            const bool g1 = j < n1s;
            const uint32_t idesc = g1 ? idesc1 : idesc2;
            const uint32_t acc = tmem + (g1 ? 0u : (uint32_t)TS_ACC2_COL);
            const uint32_t first = g1 ? 0u : n1s;
#ifdef GNN_TRACE
            const bool tr = leader && it >= (uint32_t)TRACE_TILE0 && it < (uint32_t)(TRACE_TILE0 + TRACE_TILES);
            const uint32_t tb = TRACE_ISSUER + 8 * ((it - TRACE_TILE0) * S + j);
            if (tr) GNN_STAMP(tb + 0);
#endif
            mbar_wait(sFull + 8 * (J % TS_B_STAGES), (J / TS_B_STAGES) & 1u);
#ifdef GNN_TRACE
            if (tr) GNN_STAMP(tb + 1);
#endif
            if (j == n1s && it > 0) mbar_wait(sAccFree, (it - 1) & 1u);  // the epilogue of the previous tile has read acc2
#ifdef GNN_TRACE
            if (tr) GNN_STAMP(tb + 2);
#endif
            const uint32_t bHi = sBst + (J % TS_B_STAGES) * B_STAGE_BYTES;
            const uint64_t dBh = desc(bHi), dBl = desc(bHi + B_SLAB_BYTES);
#pragma unroll
            for (uint32_t h = 0; h < 2; h++) {
                const uint32_t U = 2 * J + h;
                mbar_wait(sArdy + 8 * (U % TS_A_STAGES), (U / TS_A_STAGES) & 1u);
#ifdef GNN_TRACE
                if (tr) GNN_STAMP(tb + 3 + 2 * h);
#endif
                tc_fence_after();
                const uint32_t a = tmem + (uint32_t)(TS_A_COL + 32 * (U % TS_A_STAGES));
                if (leader) {
#pragma unroll
                    for (uint32_t q = 0; q < 2; q++) {  // k-step 2h + q of the slab: 8 tf32 = 8 columns of A, 32 bytes of a B row
                        const uint32_t kk = 2 * h + q;
                        if constexpr (BF) {
                            umma_bf16_ts(acc, a + 8 * q, dBh + 2 * kk, idesc, (j != first || kk != 0) ? 1u : 0u);
                            umma_bf16_ts(acc, a + 16 + 8 * q, dBh + 2 * kk, idesc, 1u);
                            umma_bf16_ts(acc, a + 8 * q, dBl + 2 * kk, idesc, 1u);
                        } else {
                            umma_tf32_ts(acc, a + 8 * q, dBh + 2 * kk, idesc, (j != first || kk != 0) ? 1u : 0u);
                            umma_tf32_ts(acc, a + 16 + 8 * q, dBh + 2 * kk, idesc, 1u);
                            umma_tf32_ts(acc, a + 8 * q, dBl + 2 * kk, idesc, 1u);
                        }
                    }
                    umma_commit(sDone + 8 * (U % DONE_BARS));
                    if (CL > 1 && h == 1) umma_commit_multicast(sFreeB + 8 * (J % FREE_BARS), CL_MASK);
                    if (h == 1 && j == S - 1) umma_commit(sAccFull);
                    if (h == 1 && j == n1s - 1) umma_commit(sAcc1Full);
                }
#ifdef GNN_TRACE
                if (tr) GNN_STAMP(tb + 4 + 2 * h);
#endif
                __syncwarp();
            }
            if (J + PF < total) {
                issue_b(J + PF, jn);
                jn = jn + 1 == S ? 0 : jn + 1;
            }
#ifdef GNN_TRACE
            if (tr) GNN_STAMP(tb + 7);
#endif
            if (++j == S) j = 0, it++;
        }
        if (CL > 1) {  // every arrival the other CTA still sends to this one has landed before this CTA may exit
            for (uint32_t J = total > (uint32_t)FREE_BARS ? total - FREE_BARS : 0u; J < total; J++)
                mbar_wait(sFreeB + 8 * (J % FREE_BARS), (J / FREE_BARS) & 1u);
        }
    } else if (epilogue) {
        const int H = P.h2;
        auto lds4 = [](uint32_t a) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            return v;
        };
        // acc2 is read ONCE per tile.  Reading tensor memory is slow while the tensor pipe runs (measured ~900 cycles per
        // 32-column load), and GEMM2 of the next tile may only start when the epilogue has released acc2: with the three
        // passes of round 1 (sum, variance, normalise) the epilogue held the accumulator for 27,000 cycles per tile -- longer
        // than GEMM1 -- and the chain epilogue -> GEMM2 -> epilogue set the kernel's period (issuer trace: 9,500 of a tile's
        // 34,000 cycles waiting for acc2).  Now:
        //   pass A (thread = row = TMEM lane): per 32-column chunk ReLU(acc2 [+ b2]), row statistics, and the RAW values stored
        //           to their place in `out` (through shared memory, so that four lanes write 64 contiguous bytes of a row);
        //           after the last chunk acc2 is released;
        //   pass B (the three normaliser warps below): the tile's raw rows again, from L2 (they were written microseconds ago), a row
        //           at a time, fully coalesced, normalised with the row's statistics (handed over through shared memory) and the
        //           lane's own columns of scale / offset (registers), and overwritten.
        // One-pass statistics about a pivot: with d = x - c (c = the row's first value), mean = c + sum(d) / H and
        // var = (sum(d^2) - sum(d)^2 / H) / H; the pivot keeps the subtraction benign (|mean - c| is of the order of the
        // standard deviation), unlike the raw sum of squares.
        const int NCH = (P.n2 + 31) / 32;  // chunks of 32 columns; the last one may hold 16
        const uint32_t lane_e = (uint32_t)tid & 31u, wbase = sOut + (uint32_t)(warp & 3) * 2048u;
        float xb[32];
        auto request = [&](int ch) {
            const int c0 = 32 * ch;
            if (c0 + 32 <= P.n2)
                tmem_ld32_nowait(tmem + lane_base + (uint32_t)(TS_ACC2_COL + c0), xb);
            else
                tmem_ld16_nowait(tmem + lane_base + (uint32_t)(TS_ACC2_COL + c0), xb);
        };
        for (uint32_t it = 0; it < my_tiles; it++) {
            const int tile = (int)(blockIdx.x + it * gridDim.x);
            const int row0 = tile * TILE_M + (warp & 3) * 32;  // first row of this warp
            // ---- GEMM2's A operand: ReLU(acc1 [+ b1]) of this tile, thread = row = TMEM lane, both units of every slab.  These
            // warps are idle between their pass A of the previous tile (done well inside GEMM1 of this one) and this tile's
            // acc2; the producers, who staged these units in round 1, were what the issuer waited for (10,500 cycles per tile)
            // once the epilogue was off the critical path.  acc1 needs no barrier of its own: the issuer consumes units in order,
            // so GEMM1 of the next tile overwrites acc1 only after the last of these units -- i.e. its read -- is published.
            {
                const uint32_t Jt = it * S + n1s;   // global index of the tile's first GEMM2 slab
                // a barrier of its own, one completion per tile: the per-unit barriers are waited for by parity, which is only
                // sound for a waiter that has seen the previous completion of the same barrier -- this warp has not
                mbar_wait(sAcc1Full, it & 1u);
                tc_fence_after();
                for (uint32_t j2 = 0; j2 < n2s; j2++) {
#pragma unroll 1
                    for (uint32_t g = 0; g < 2; g++) {
                        const uint32_t U = 2 * (Jt + j2) + g;
                        float v[NV];
                        uint32_t r[32];
#pragma unroll
                        for (int half = 0; half < NV / 16; half++) {
                            const int c0 = (int)j2 * SLABK + (int)g * NV + 16 * half;
                            float* vh = v + 16 * half;
                            if (c0 < P.n1) {
                                tmem_ld16(tmem + lane_base + (uint32_t)c0, vh);
                                if constexpr (BF) {  // b1 came through the GEMM (constant-one column of A): only the ReLU is left
#pragma unroll
                                    for (int i = 0; i < 16; i++) vh[i] = fmaxf(vh[i], 0.f);
                                } else {
#pragma unroll
                                    for (int q = 0; q < 4; q++) {
                                        const float4 b = lds4(sPar + (3 * MAX_N + c0 + 4 * q) * 4);
                                        vh[4 * q] = fmaxf(vh[4 * q] + b.x, 0.f), vh[4 * q + 1] = fmaxf(vh[4 * q + 1] + b.y, 0.f);
                                        vh[4 * q + 2] = fmaxf(vh[4 * q + 2] + b.z, 0.f), vh[4 * q + 3] = fmaxf(vh[4 * q + 3] + b.w, 0.f);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; i++) vh[i] = 0.f;
                            }
                            if constexpr (BF) {  // the constant-one column behind the hidden activation: W2's image holds b2 in row h1
                                if (c0 <= P.h1 && P.h1 < c0 + 16) {
#pragma unroll
                                    for (int i = 0; i < 16; i++)
                                        if (c0 + i == P.h1) vh[i] = 1.f;
                                }
                            }
                        }
                        if constexpr (BF) {
#pragma unroll
                            for (int i = 0; i < 16; i++) {  // one packing conversion per pair and piece (see the producers' split_pair)
                                split_pair16<F16>(v[2 * i], v[2 * i + 1], r[i], r[16 + i]);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                const float hi = round_tf32(v[i]);
                                r[i] = __float_as_uint(hi);
                                r[16 + i] = __float_as_uint(residual_tf32(v[i], hi));
                            }
                        }
                        if (U >= (uint32_t)TS_A_STAGES) wait_unit(U - TS_A_STAGES);
                        tc_fence_after();
                        tmem_st32(tmem + lane_base + (uint32_t)(TS_A_COL + 32 * (U % TS_A_STAGES)), r);
                        tc_fence_before();
                        __syncwarp();
                        if (lane_e == 0) mbar_arrive(sArdy + 8 * (U % TS_A_STAGES));
                    }
                }
            }
            mbar_wait_relaxed(sAccFull, it & 1u);
#ifdef GNN_TRACE
            if (tid == MLPP_EPILOGUE0 && it >= (uint32_t)TRACE_TILE0 && it < (uint32_t)(TRACE_TILE0 + TRACE_TILES)) GNN_STAMP(TRACE_EPI + 4 * (it - TRACE_TILE0));
#endif
            tc_fence_after();
            float sd = 0.f, sd2 = 0.f, pivot = 0.f;
            request(0);
            for (int ch = 0; ch < NCH; ch++) {
                float x[32];
                tmem_wait_ld();
#ifdef GNN_TRACE
                const bool trc = (ch == 3 || ch == 4) && tid == MLPP_EPILOGUE0 && it == (uint32_t)TRACE_TILE0;
                if (trc) GNN_STAMP(TRACE_EPI + 32 + 8 * (ch - 3));
#endif
                if constexpr (BF) {  // b2 came through the GEMM
#pragma unroll
                    for (int i = 0; i < 32; i++) x[i] = fmaxf(xb[i], 0.f);
                } else {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (32 * ch + 4 * q < P.n2) {
                            const float4 b = lds4(sPar + (32 * ch + 4 * q) * 4);
                            x[4 * q] = fmaxf(xb[4 * q] + b.x, 0.f), x[4 * q + 1] = fmaxf(xb[4 * q + 1] + b.y, 0.f);
                            x[4 * q + 2] = fmaxf(xb[4 * q + 2] + b.z, 0.f), x[4 * q + 3] = fmaxf(xb[4 * q + 3] + b.w, 0.f);
                        }
                    }
                }
#ifdef GNN_TRACE
                if (trc) GNN_STAMP(TRACE_EPI + 33 + 8 * (ch - 3));
#endif
                if (ch + 1 < NCH) {
                    request(ch + 1);
                } else {  // the last read of acc2 has landed: GEMM2 of the next tile may overwrite it
                    tc_fence_before();
                    mbar_arrive(sAccFree);
#ifdef GNN_TRACE
                    if (tid == MLPP_EPILOGUE0 && it >= (uint32_t)TRACE_TILE0 && it < (uint32_t)(TRACE_TILE0 + TRACE_TILES)) GNN_STAMP(TRACE_EPI + 4 * (it - TRACE_TILE0) + 1);
#endif
                }
#ifdef GNN_TRACE
                if (trc) GNN_STAMP(TRACE_EPI + 34 + 8 * (ch - 3));
#endif
                if (ch == 0) pivot = x[0];
                {  // four partial sums per moment: the additions of a chunk are not one dependent chain
                    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const float d = (32 * ch + i < H) ? x[i] - pivot : 0.f;
                        s1[i & 3] += d;
                        s2[i & 3] = fmaf(d, d, s2[i & 3]);
                    }
                    sd += (s1[0] + s1[1]) + (s1[2] + s1[3]);
                    sd2 += (s2[0] + s2[1]) + (s2[2] + s2[3]);
                }
#ifdef GNN_TRACE
                if (trc) GNN_STAMP(TRACE_EPI + 35 + 8 * (ch - 3));
#endif
                // raw values to `out` through shared memory: each lane parks 16 columns of its row (four 16-byte units, unit
                // index XORed with (row / 2) % 4: conflict-free both ways), then four lanes store 64 contiguous bytes of one
                // row -- a store instruction covers 8 rows x 2 full sectors instead of 32 rows x half a sector
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int c16 = 32 * ch + 16 * half;
                    if (c16 < H) {
#pragma unroll
                        for (uint32_t u = 0; u < 4; u++)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wbase + lane_e * 64u + ((u ^ ((lane_e >> 1) & 3u)) << 4)),
                                         "f"(x[16 * half + 4 * u]), "f"(x[16 * half + 4 * u + 1]), "f"(x[16 * half + 4 * u + 2]), "f"(x[16 * half + 4 * u + 3])
                                         : "memory");
                        __syncwarp();
#pragma unroll
                        for (uint32_t k = 0; k < 4; k++) {
                            const uint32_t r = 8u * k + (lane_e >> 2), u = lane_e & 3u;
                            const float4 y = lds4(wbase + r * 64u + ((u ^ ((r >> 1) & 3u)) << 4));
                            const int grow = row0 + (int)r, col = c16 + 4 * (int)u;
                            if (col < H && grow < P.rows) *reinterpret_cast<float4*>(P.out + (size_t)grow * H + col) = y;
                        }
                        __syncwarp();
#ifdef GNN_TRACE
                        if (trc) GNN_STAMP(TRACE_EPI + 36 + half + 8 * (ch - 3));
#endif
                    }
                }
            }
#ifdef GNN_TRACE
            if (tid == MLPP_EPILOGUE0 && it >= (uint32_t)TRACE_TILE0 && it < (uint32_t)(TRACE_TILE0 + TRACE_TILES)) GNN_STAMP(TRACE_EPI + 4 * (it - TRACE_TILE0) + 2);
#endif
            const float md = sd / (float)H, mean = pivot + md;
            const float rstd = rsqrtf(fmaxf(sd2 / (float)H - md * md, 0.f) + 1e-5f);
            // ---- hand the rows to the normaliser warps: statistics into shared memory (double-buffered by tile parity), raw rows
            // already in `out`.  The barrier arrival releases this warp's global and shared stores to the CTA.
            {
                const uint32_t par = it & 1u;
                if (it >= 2) mbar_wait(sStatE + 8 * par, ((it >> 1) - 1u) & 1u);  // the normalisers are done with tile it - 2
                const uint32_t sa = sStat + par * (TILE_M * 8) + ((uint32_t)(warp & 3) * 32u + lane_e) * 8u;
                asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sa), "f"(rstd), "f"(-mean * rstd) : "memory");
                __threadfence_block();
                __syncwarp();
                if (lane_e == 0) mbar_arrive(sStatF + 8 * par);
            }
#ifdef GNN_TRACE
            if (tid == MLPP_EPILOGUE0 && it >= (uint32_t)TRACE_TILE0 && it < (uint32_t)(TRACE_TILE0 + TRACE_TILES)) GNN_STAMP(TRACE_EPI + 4 * (it - TRACE_TILE0) + 3);
#endif
        }
    }
    else if (tid >= MLPP_NORM0) {
        // ---- pass B of the epilogue on its own warps: the tile's 128 raw rows again, from L2 (ld.global.cg: they were written by the
        // epilogue warps of this CTA microseconds ago; the mbarrier orders those stores), a row at a time, fully coalesced, normalised
        // with the row's statistics (broadcast loads from shared memory) and this lane's own columns of scale / offset (registers),
        // and overwritten.  Warp h takes rows h, h + 3, ...; four rows in flight.
        const int H = P.h2, q4 = H >> 2;
        const uint32_t lane_n = (uint32_t)tid & 31u;
        const int hid = (tid - MLPP_NORM0) >> 5;
        auto lds4n = [](uint32_t a) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            return v;
        };
        float4 gsc[2], gof[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int c = 4 * ((int)lane_n + 32 * t);
            gsc[t] = c < H ? lds4n(sPar + (MAX_N + c) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            gof[t] = c < H ? lds4n(sPar + (2 * MAX_N + c) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool t0_ok = (int)lane_n < q4, t1_ok = (int)lane_n + 32 < q4;
        for (uint32_t it = 0; it < my_tiles; it++) {
            const int tile = (int)(blockIdx.x + it * gridDim.x);
            const uint32_t par = it & 1u;
            mbar_wait_relaxed(sStatF + 8 * par, (it >> 1) & 1u);
            const int nrows = min(TILE_M, P.rows - tile * TILE_M);  // <= 0 for a tile past the end (cluster padding)
            for (int r0 = hid; r0 < nrows; r0 += 4 * MLPP_NORM_WARPS) {
                float4 v[4][2];
                float2 st[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = min(r0 + k * MLPP_NORM_WARPS, nrows - 1);
                    const float4* rowp = reinterpret_cast<const float4*>(P.out + (size_t)(tile * TILE_M + r) * H);
                    if (t0_ok) v[k][0] = __ldcg(rowp + lane_n);
                    if (t1_ok) v[k][1] = __ldcg(rowp + lane_n + 32);
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(st[k].x), "=f"(st[k].y) : "r"(sStat + par * (TILE_M * 8) + (uint32_t)r * 8u));
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = r0 + k * MLPP_NORM_WARPS;
                    if (r < nrows) {
                        float4* rowp = reinterpret_cast<float4*>(P.out + (size_t)(tile * TILE_M + r) * H);
#pragma unroll
                        for (int t = 0; t < 2; t++)
                            if (t ? t1_ok : t0_ok) {  // y = ((v - mean) * rstd) * scale + offset as two FMAs per element
                                float4 y;
                                y.x = fmaf(fmaf(v[k][t].x, st[k].x, st[k].y), gsc[t].x, gof[t].x), y.y = fmaf(fmaf(v[k][t].y, st[k].x, st[k].y), gsc[t].y, gof[t].y);
                                y.z = fmaf(fmaf(v[k][t].z, st[k].x, st[k].y), gsc[t].z, gof[t].z), y.w = fmaf(fmaf(v[k][t].w, st[k].x, st[k].y), gsc[t].w, gof[t].w);
                                rowp[lane_n + 32 * t] = y;
                            }
                    }
                }
            }
            __syncwarp();
            if (lane_n == 0) mbar_arrive(sStatE + 8 * par);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// ---- small kernels -------------------------------------------------------------------------------
// Linear(2 -> D): out[i][c] = x[i][0]*w[0][c] + x[i][1]*w[1][c] + b[c]     (model.py:13-16)
__global__ void embed_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                             float* __restrict__ out, int64_t rows, int D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * D) return;
    const int64_t r = i / D;
    const int c = (int)(i % D);
    out[i] = x[2 * r] * w[c] + x[2 * r + 1] * w[D + c] + b[c];
}

// Same embedding when the 2-vector input is one-hot and given as its index (LCG: nodes are
// literal [0,1] or clause [1,0], edges clause-edge [1,0] or pair-edge [0,1]; sat_instances.py:65-73)
__global__ void embed_onehot_kernel(const uint8_t* __restrict__ type, const float* __restrict__ w, const float* __restrict__ b,
                                    float* __restrict__ out, int64_t rows, int D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * D) return;
    const int c = (int)(i % D);
    out[i] = w[(int)type[i / D] * D + c] + b[c];
}

// jraph.utils.segment_sum over a CSR of edge ids (edge order inside a segment = increasing edge id, i.e. the order a
// sequential scatter-add visits them): out[v] = sum_{j in [ptr[v], ptr[v+1])} e[ids[j]], for both directions of every node in
// one launch (nodes and edges of an instance are contiguous in a batch, so the two sweeps over an instance's edge rows fall
// into the same time window and the second is served by L2).  16-byte loads (lane l owns float4 l and l+32 of the row), edge ids fetched 32 at a time and broadcast by
// shuffle so that the row loads of consecutive edges are independent and eight are in flight per lane.  The additions of one
// output element happen in list order (deterministic, same values as a sequential scatter-add).
// A pair with fewer than two edges is not summed at all: the node update reads the zero row or the edge row itself through
// agg_index_kernel's indices (2/3 of an LCG graph's pairs: clause nodes send nothing, positive literals receive nothing,
// negative literals receive only their pair edge).
// (Round 2 also measured one thread per (pair, 16-byte column) -- no idle lanes on the second half of an 800-byte row -- at
// 4.2 ms per forward against 3.4 ms for this form: four loads in flight per thread were too few.)
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ e, const int64_t* __restrict__ ptr0,
                                                          const int32_t* __restrict__ ids0, float* __restrict__ out0,
                                                          const int64_t* __restrict__ ptr1, const int32_t* __restrict__ ids1,
                                                          float* __restrict__ out1, int64_t n_node, int H)
{
    // one warp per NODE, both directions one after the other: in an LCG graph a node needs at most one of the two sums (literals
    // send, clauses receive), so every warp of the launch has exactly one sum to do; with a warp per (node, direction) half of
    // the warps of every CTA retired at once and their slots stayed empty until the CTA's other warps were done
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n_node) return;
    const int q = H >> 2;  // float4 per row (H is a multiple of 4, at most 64 float4 = 256 columns)
    const bool has0 = lane < q, has1 = lane + 32 < q;
    // both directions' list bounds in one round trip (a clause node found "sends nothing" first and fetched its receive bounds in
    // a second trip).  Measured: no change (2.7 ms per forward either way) -- the kernel sits at 59 % of the copy peak with DRAM
    // traffic at the algorithmic minimum; what it reads are scattered 800-byte rows interleaved with an equal number of row writes
    const int64_t s0 = __ldg(ptr0 + v), t0 = __ldg(ptr0 + v + 1), s1 = __ldg(ptr1 + v), t1 = __ldg(ptr1 + v + 1);
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++) {
        const int64_t s = dir ? s1 : s0, t = dir ? t1 : t0;
        if (t - s < 2) continue;
        const int32_t* __restrict__ ids = dir ? ids1 : ids0;
        float* __restrict__ out = dir ? out1 : out0;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        for (int64_t base = s; base < t; base += 32) {
            const int cnt = (int)min((int64_t)32, t - base);
            const int32_t my_id = lane < cnt ? __ldg(ids + base + lane) : 0;
#pragma unroll 8
            for (int k = 0; k < cnt; k++) {
                const int32_t id = __shfl_sync(0xffffffffu, my_id, k);
                const float4* row = reinterpret_cast<const float4*>(e + (size_t)id * H);
                if (has0) {
                    const float4 x = __ldg(row + lane);
                    a0.x += x.x, a0.y += x.y, a0.z += x.z, a0.w += x.w;
                }
                if (has1) {
                    const float4 x = __ldg(row + lane + 32);
                    a1.x += x.x, a1.y += x.y, a1.z += x.z, a1.w += x.w;
                }
            }
        }
        float4* o = reinterpret_cast<float4*>(out + (size_t)v * H);
        if (has0) o[lane] = a0;
        if (has1) o[lane + 32] = a1;
    }
}

// Row indices of a node's two aggregates in the array [eA (E rows) | sent (N) | recv (N) | zero (1) | eB (E)] (gnn_api.cu,
// forward_device), one set per buffer the step's new edges live in (A = eA, B = eB): degree 0 -> the zero row, degree 1 ->
// the edge row itself, degree >= 2 -> the row segment_sum_flat_kernel writes for this node.
__global__ void agg_index_kernel(const int64_t* __restrict__ sptr, const int32_t* __restrict__ sids, const int64_t* __restrict__ rptr,
                                 const int32_t* __restrict__ rids, int64_t N, int64_t E, int32_t* __restrict__ sidxA,
                                 int32_t* __restrict__ ridxA, int32_t* __restrict__ sidxB, int32_t* __restrict__ ridxB)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int64_t zero = E + 2 * N, eb = zero + 1;
    {
        const int64_t s = sptr[v], d = sptr[v + 1] - s;
        const int64_t one = d == 1 ? sids[s] : 0;
        sidxA[v] = (int32_t)(d == 0 ? zero : d == 1 ? one : E + v);
        sidxB[v] = (int32_t)(d == 0 ? zero : d == 1 ? eb + one : E + v);
    }
    {
        const int64_t s = rptr[v], d = rptr[v + 1] - s;
        const int64_t one = d == 1 ? rids[s] : 0;
        ridxA[v] = (int32_t)(d == 0 ? zero : d == 1 ? one : E + N + v);
        ridxB[v] = (int32_t)(d == 0 ? zero : d == 1 ? eb + one : E + N + v);
    }
}

// readout Linear(H -> 1): one warp per node (model.py:144)
__global__ void readout_kernel(const float* __restrict__ h, const float* __restrict__ w, float b, float* __restrict__ out,
                               int64_t n_node, int H)
{
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n_node) return;
    float acc = 0.f;
    for (int c = lane; c < H; c += 32) acc += h[(size_t)v * H + c] * __ldg(w + c);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) out[v] = acc + b;
}

// LCG.get_model_probabilities: softmax over (decoded[2j], decoded[2j+1]) -> column 1
// (sat_representations.py:775-780); var_node0[j] = node index of the negative literal of variable j
__global__ void pair_prob_kernel(const float* __restrict__ decoded, const int64_t* __restrict__ var_node0, float* __restrict__ p,
                                 int64_t n_vars)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_vars) return;
    const int64_t v0 = var_node0[j];
    const float a = decoded[v0], b = decoded[v0 + 1];
    const float mx = fmaxf(a, b);
    const float ea = __expf(a - mx), eb = __expf(b - mx);
    p[j] = eb / (ea + eb);
}


// ---------------------------------------------------------------------------------------------
// LCG graph of a batch of resident instances, built on the device from their clause / occurrence CSRs
// (LCG.get_graph, sat_representations.py:607-666, with sat_instances.py:50-73; same node / edge numbering as the
// host builder in gnn_api.cu).  Instance i owns nodes [node_off, node_off + 2n + m): literal nodes 2j (negative) and
// 2j+1 (positive) then one node per clause; edges [edge_off, edge_off + L + n): one literal -> clause edge per
// literal occurrence in file order, then one positive -> negative edge per variable.  The two segment CSRs
// (edge ids per sender / per receiver, increasing edge id) are written directly.  Instances with empty clauses
// (which the reference drops) take the host path.
struct LcgItem {
    const uint32_t* row_ptr;
    const uint32_t* lits;
    const uint32_t* occ_ptr;
    const uint32_t* occ;
    uint32_t n, m, L, pad;
    int64_t node_off, edge_off, var_off, clause_off;  // clause_off: position of the first clause in the batch-wide clause numbering
};

__device__ __forceinline__ uint32_t lcg_find(const int64_t* __restrict__ begin, uint32_t n_items, int64_t x)
{
    uint32_t lo = 0, hi = n_items;  // invariant: begin[lo] <= x < begin[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(begin + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

struct LcgOut {
    int32_t *send, *recv, *sids, *rids;
    int64_t *sptr, *rptr, *var_node0;
    uint8_t *node_type, *edge_type;
    int64_t N, E;  // totals of the batch: sptr[N] = rptr[N] = E
};

// one thread per clause of the batch
__global__ void lcg_clauses_kernel(const LcgItem* __restrict__ items, const int64_t* __restrict__ clause_begin, uint32_t n_items,
                                   int64_t total_clauses, LcgOut o)
{
    const int64_t gc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gc >= total_clauses) return;
    const uint32_t it = lcg_find(clause_begin, n_items, gc);
    const LcgItem I = items[it];
    const uint32_t c = (uint32_t)(gc - I.clause_off);
    const int64_t no = I.node_off, eo = I.edge_off, n = I.n;
    const uint32_t s = __ldg(I.row_ptr + c), e = __ldg(I.row_ptr + c + 1);
    const int64_t node = no + 2 * n + c;
    o.node_type[node] = 0;
    o.rptr[node] = eo + n + s;
    o.sptr[node] = eo + I.L + n;  // clause nodes send nothing
    for (uint32_t j = s; j < e; j++) {
        const uint32_t lit = __ldg(I.lits + j);
        o.send[eo + j] = (int32_t)(no + 2 * (int64_t)(lit >> 1) + ((lit & 1u) ? 0 : 1));
        o.recv[eo + j] = (int32_t)node;
        o.edge_type[eo + j] = 0;
        o.rids[eo + n + j] = (int32_t)(eo + j);
    }
}

// one thread per variable of the batch
__global__ void lcg_variables_kernel(const LcgItem* __restrict__ items, const int64_t* __restrict__ var_begin, uint32_t n_items,
                                     int64_t total_vars, LcgOut o)
{
    const int64_t gv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gv >= total_vars) return;
    const uint32_t it = lcg_find(var_begin, n_items, gv);
    const LcgItem I = items[it];
    const uint32_t v = (uint32_t)(gv - I.var_off);
    const int64_t no = I.node_off, eo = I.edge_off;
    const int64_t neg = no + 2 * (int64_t)v, pos = neg + 1, pair = eo + I.L + v;
    o.node_type[neg] = 1;  // np.eye(2)[1] and np.eye(2)[-1] are both [0,1]
    o.node_type[pos] = 1;
    o.send[pair] = (int32_t)pos;
    o.recv[pair] = (int32_t)neg;
    o.edge_type[pair] = 1;
    o.var_node0[gv] = neg;
    if (gv == 0) o.sptr[o.N] = o.rptr[o.N] = o.E;
    // receivers: the negative literal receives the pair edge, the positive literal nothing
    o.rptr[neg] = eo + v;
    o.rptr[pos] = eo + v + 1;
    o.rids[eo + v] = (int32_t)pair;
    // senders: occurrences of the negative literal, then of the positive literal (increasing edge id = increasing
    // literal position, the order of the occurrence list), then the pair edge
    const uint32_t os = __ldg(I.occ_ptr + v), oe = __ldg(I.occ_ptr + v + 1);
    const int64_t base = eo + os + v;  // every earlier variable contributed its occurrences and one pair edge
    auto position = [&](uint32_t c) -> uint32_t {
        uint32_t j = __ldg(I.row_ptr + c);
        while ((__ldg(I.lits + j) >> 1) != v) j++;  // the variable occurs in the clause (exactly once: duplicates are refused)
        return j;
    };
    uint32_t nneg = 0;
    for (uint32_t q = os; q < oe; q++) nneg += __ldg(I.lits + position(__ldg(I.occ + q))) & 1u;
    o.sptr[neg] = base;
    o.sptr[pos] = base + nneg;
    uint32_t in = 0, ip = 0;
    for (uint32_t q = os; q < oe; q++) {
        const uint32_t j = position(__ldg(I.occ + q));
        if (__ldg(I.lits + j) & 1u)
            o.sids[base + in++] = (int32_t)(eo + j);
        else
            o.sids[base + nneg + ip++] = (int32_t)(eo + j);
    }
    o.sids[base + (oe - os)] = (int32_t)pair;
}

}  // namespace gnn
