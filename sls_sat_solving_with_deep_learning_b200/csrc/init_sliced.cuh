// init_sliced.cuh -- chain initialisation for the HBM tier, 256 chains at a time.
//
// Restates the reference's per-chain setup (rust/src/lib.rs:206-214: assignment[i] =
// gen_bool(weights_initialize[i]); violated = find_violated_clauses(assignment), :176-181) for
// instances whose chain state lives in HBM (BASELINE config 5: 4*10^6 clauses, 640 KB per chain).
// Doing the full clause scan inside the walk kernel costs one literal fetch and k scattered
// assignment reads per clause PER CHAIN.  Here a "slice" of SLICE_W = 256 chains is processed together
// (same layout and clause evaluation as eval.cuh):
//   init_assign_sliced_kernel : lane = chain; every lane evaluates its chain's Philox init stream,
//                               a ballot per variable yields one word of the transposed tq[slice][v]
//                               (8 words = the variable's value in the 256 chains) and 16 blocks yield the
//                               chain's own abits word;
//   init_scan_sliced_kernel   : lane = clause; one literal fetch + one 32-byte sector of tq per literal
//                               evaluates the clause for 256 chains with bitwise ops, register transposes
//                               turn 32x32 tiles into per-chain vbits words, staged in shared memory and
//                               written as whole sectors, plus the level-1 counters (the streaming,
//                               bandwidth-bound form of lib.rs:176-181);
//   init_counters_kernel      : warp = chain; level-2 counters and the initial energy.
// The walk kernel then starts from the prepared slab (WalkArgs::nf_init != nullptr).
// Not used when the instance has duplicate clause contents (canonical ids need the atomic path).
#pragma once
#include "common.h"
#include "eval.cuh"
#include "philox.cuh"

namespace slsb {

// grid: (word chunks, slices in this launch); block: 8 warps, warp q owns chains 32q .. 32q+31 of the slice
// and works through the block's assignment words w (variables 32w .. 32w+31).
// ONES: some threshold is Bernoulli's ALWAYS_TRUE (p == 1, rand 0.8.5): only then the draw needs the extra test -- the host
// knows (sls_solver_set_weights) and picks the instantiation.  Whole words (all 32 variables exist) take a path without bounds
// tests and fetch a pair's two thresholds with one 128-bit load; the kernel is bound by Philox's multiplies and by issue
// (ncu: math_pipe_throttle, not_selected), so what is trimmed are the ~28 instructions per Philox block that are not Philox.
template <bool ONES>
__global__ void __launch_bounds__(256) init_assign_sliced_kernel(const WalkArgs A, uint32_t* __restrict__ tq_all, uint32_t slice0,
                                                                 uint32_t words_per_block)
{
    const DevInst& I = A.inst;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const uint32_t slice = slice0 + blockIdx.y;
    const uint64_t run_raw = (uint64_t)slice * SLICE_W + 32u * q + lane;
    const bool valid = run_raw < A.nruns;
    const uint64_t run = valid ? run_raw : A.nruns - 1;
    const ChainKey ck(A.seed, A.chain_offset + run, STREAM_INIT);
    uint32_t* abits = A.gstate + run * (size_t)I.slab_words;
    uint32_t* tq = tq_all + (size_t)blockIdx.y * I.n * SLICE_Q;
    const uint32_t w_begin = blockIdx.x * words_per_block;
    const uint32_t w_end = min(I.nwords, w_begin + words_per_block);
    auto draw = [](uint64_t x, uint64_t t) { return ONES ? (t == ~0ull || x < t) : x < t; };
    for (uint32_t w = w_begin; w < w_end; w++) {
        uint32_t bits = 0, tmine = 0;
        if (w * 32u + 31u < I.n) {
            const ulonglong2* __restrict__ thr2 = reinterpret_cast<const ulonglong2*>(A.thr_init) + (size_t)w * 16u;  // 16-byte aligned pairs
#pragma unroll 4
            for (uint32_t b = 0; b < 16; b++) {
                const uint4 x = ck.block((uint64_t)w * 16u + b);
                const ulonglong2 t = __ldg(thr2 + b);
                const bool bit0 = draw((uint64_t)x.y << 32 | x.x, t.x), bit1 = draw((uint64_t)x.w << 32 | x.z, t.y);
                bits |= (uint32_t)bit0 << (2u * b) | (uint32_t)bit1 << (2u * b + 1u);
                const uint32_t tw0 = __ballot_sync(0xffffffffu, bit0 && valid);
                const uint32_t tw1 = __ballot_sync(0xffffffffu, bit1 && valid);
                if ((uint32_t)lane == 2u * b) tmine = tw0;
                if ((uint32_t)lane == 2u * b + 1u) tmine = tw1;
            }
        } else {
#pragma unroll 1
            for (uint32_t b = 0; b < 16; b++) {
                const uint32_t v0 = w * 32u + 2u * b;
                if (v0 >= I.n) break;  // warp-uniform
                const uint4 x = ck.block((uint64_t)(v0 >> 1));
                const bool bit0 = draw((uint64_t)x.y << 32 | x.x, __ldg(A.thr_init + v0));
                const bool bit1 = v0 + 1 < I.n && draw((uint64_t)x.w << 32 | x.z, __ldg(A.thr_init + v0 + 1));
                bits |= (uint32_t)bit0 << (2u * b) | (uint32_t)bit1 << (2u * b + 1u);
                const uint32_t tw0 = __ballot_sync(0xffffffffu, bit0 && valid);
                const uint32_t tw1 = __ballot_sync(0xffffffffu, bit1 && valid);
                if ((uint32_t)lane == 2u * b) tmine = tw0;
                if ((uint32_t)lane == 2u * b + 1u) tmine = tw1;
            }
        }
        if (valid) abits[w] = bits;
        const uint32_t v = w * 32u + lane;
        if (v < I.n) tq[(size_t)v * SLICE_Q + q] = tmine;  // the 8 warps of the block complete the sector of v
    }
}

// grid: (level-1 groups / WARPS, slices in this launch); each warp owns one level-1 group = 32 vbits
// words = 1024 clauses = 4 tiles and produces, for each of the slice's 256 chains, those words and the counter.
template <int WARPS, int NL = 4>
__global__ void __launch_bounds__(WARPS * 32) init_scan_sliced_kernel(const WalkArgs A, const uint32_t* __restrict__ tq_all, uint32_t slice0)
{
    __shared__ uint32_t stage[WARPS][SLICE_W][TILE_WORDS + 1];
    const DevInst& I = A.inst;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x * WARPS + warp;
    if (g >= I.ngroups) return;
    const uint32_t slice = slice0 + blockIdx.y;
    const uint64_t run0 = (uint64_t)slice * SLICE_W;
    const uint32_t valid = (uint32_t)min((uint64_t)SLICE_W, A.nruns - run0);
    const uint32_t* tq = tq_all + (size_t)blockIdx.y * I.n * SLICE_Q;
    uint32_t acc[SLICE_Q];  // lane b, entry q: violated clauses of chain run0 + 32q + b in this group
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) acc[q] = 0;
    ClauseRowsT<NL> rows(I, tq, g * 1024u + lane);  // the 32 rows of the group, pipelined across its four tiles
    for (uint32_t t = 0; t < 32 / TILE_WORDS; t++) {
        const uint32_t w0 = g * 32u + t * TILE_WORDS;  // first vbits word of the tile (whole groups exist in the slab)
#pragma unroll 1
        for (uint32_t wi = 0; wi < TILE_WORDS; wi++) {
            uint32_t viol[SLICE_Q];
            rows.next(viol);
            transpose32x8(viol, lane);
#pragma unroll
            for (uint32_t q = 0; q < SLICE_Q; q++) {
                acc[q] += __popc(viol[q]);
                stage[warp][32u * q + lane][wi] = viol[q];
            }
        }
        __syncwarp();
        const uint32_t j = lane & 7u;
        for (uint32_t r = 0; r < SLICE_W / 4; r++) {
            const uint32_t a = r * 4u + (lane >> 3);
            if (a < valid) __stcs(A.gstate + (run0 + a) * (size_t)I.slab_words + I.off_v + w0 + j, stage[warp][a][j]);
        }
        __syncwarp();
    }
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) {
        const uint32_t a = 32u * q + lane;
        if (a < valid) A.gstate[(run0 + a) * (size_t)I.slab_words + I.off_l1 + g] = acc[q];
    }
}

// one warp per chain: level-2 counters and the initial number of violated clauses
__global__ void __launch_bounds__(128) init_counters_kernel(const WalkArgs A)
{
    const DevInst& I = A.inst;
    const int lane = threadIdx.x & 31;
    const uint64_t run = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (run >= A.nruns) return;
    uint32_t* slab = A.gstate + run * (size_t)I.slab_words;
    const uint32_t* l1 = slab + I.off_l1;
    uint32_t total = 0;
    if (I.l2_pad) {
        for (uint32_t j = lane; j < I.l2_pad; j += 32) {
            uint32_t acc = 0;
            if (j * 64u < I.l1_pad)
                for (uint32_t i = 0; i < 64; i++) acc += l1[j * 64u + i];
            slab[I.off_l2 + j] = acc;
            total += acc;
        }
    } else {
        for (uint32_t i = lane; i < I.l1_pad; i += 32) total += l1[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (lane == 0) A.nf_init[run] = total;
}

}  // namespace slsb
