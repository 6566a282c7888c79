// init_sliced.cuh -- chain initialisation for the HBM tier, 32 chains at a time.
//
// Restates the reference's per-chain setup (rust/src/lib.rs:206-214: assignment[i] =
// gen_bool(weights_initialize[i]); violated = find_violated_clauses(assignment), :176-181) for
// instances whose chain state lives in HBM (BASELINE config 5: 4*10^6 clauses, 640 KB per chain).
// Doing the full clause scan inside the walk kernel costs one literal fetch and k scattered
// assignment reads per clause PER CHAIN.  Here a "slice" of 32 chains is processed together:
//   init_assign_sliced_kernel : lane = chain; every lane evaluates its chain's Philox init stream,
//                               a ballot per variable yields the transposed word tbits[slice][v]
//                               (bit c = value of v in chain 32*slice + c) and 16 blocks yield the
//                               chain's own abits word;
//   init_scan_sliced_kernel   : lane = clause; one literal fetch + one tbits word per literal
//                               evaluates the clause for 32 chains with bitwise ops, a 32x32
//                               ballot transpose turns it into per-chain vbits words and level-1
//                               counters (the streaming, bandwidth-bound form of lib.rs:176-181);
//   init_counters_kernel      : warp = chain; level-2 counters and the initial energy.
// The walk kernel then starts from the prepared slab (WalkArgs::nf_init != nullptr).
// Not used when the instance has duplicate clause contents (canonical ids need the atomic path).
#pragma once
#include "common.h"
#include "philox.cuh"

namespace slsb {

// grid: (word chunks, slices in this launch); block: 128 threads = 4 warps, each warp owns whole
// assignment words w (variables 32w .. 32w+31) of the slice's 32 chains.
__global__ void __launch_bounds__(128) init_assign_sliced_kernel(const WalkArgs A, uint32_t* __restrict__ tbits, uint32_t slice0,
                                                                 uint32_t words_per_block)
{
    const DevInst& I = A.inst;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slice = slice0 + blockIdx.y;
    const uint64_t run_raw = (uint64_t)slice * 32u + lane;
    const bool valid = run_raw < A.nruns;
    const uint64_t run = valid ? run_raw : A.nruns - 1;
    const ChainKey ck(A.seed, A.chain_offset + run, STREAM_INIT);
    uint32_t* abits = A.gstate + run * (size_t)I.slab_words;
    uint32_t* tb = tbits + (size_t)blockIdx.y * I.n;
    const uint32_t w_begin = blockIdx.x * words_per_block;
    const uint32_t w_end = min(I.nwords, w_begin + words_per_block);
    for (uint32_t w = w_begin + warp; w < w_end; w += 4) {
        uint32_t bits = 0, tmine = 0;
#pragma unroll 4
        for (uint32_t b = 0; b < 16; b++) {
            const uint32_t v0 = w * 32u + 2u * b;
            if (v0 >= I.n) break;  // warp-uniform
            const uint4 x = ck.block((uint64_t)(v0 >> 1));
            const uint64_t t0 = __ldg(A.thr_init + v0);
            const bool bit0 = t0 == ~0ull || ((uint64_t)x.y << 32 | x.x) < t0;
            bool bit1 = false;
            if (v0 + 1 < I.n) {
                const uint64_t t1 = __ldg(A.thr_init + v0 + 1);
                bit1 = t1 == ~0ull || ((uint64_t)x.w << 32 | x.z) < t1;
            }
            bits |= (uint32_t)bit0 << (2u * b) | (uint32_t)bit1 << (2u * b + 1u);
            const uint32_t tw0 = __ballot_sync(0xffffffffu, bit0 && valid);
            const uint32_t tw1 = __ballot_sync(0xffffffffu, bit1 && valid);
            if ((uint32_t)lane == 2u * b) tmine = tw0;
            if ((uint32_t)lane == 2u * b + 1u) tmine = tw1;
        }
        if (valid) abits[w] = bits;
        const uint32_t v = w * 32u + lane;
        if (v < I.n) tb[v] = tmine;  // one coalesced 128-byte store per assignment word
    }
}

// grid: (level-1 groups / 4, slices in this launch); each warp owns one level-1 group = 32 vbits
// words = 1024 clauses and produces, for each of the slice's 32 chains, those 32 words and the counter.
__global__ void __launch_bounds__(128) init_scan_sliced_kernel(const WalkArgs A, const uint32_t* __restrict__ tbits, uint32_t slice0)
{
    const DevInst& I = A.inst;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = blockIdx.x * 4 + warp;
    if (g >= I.ngroups) return;
    const uint32_t slice = slice0 + blockIdx.y;
    const uint64_t run = (uint64_t)slice * 32u + lane;  // this lane's chain when it stores
    const bool valid = run < A.nruns;
    const uint32_t* tb = tbits + (size_t)blockIdx.y * I.n;
    uint32_t* slab = A.gstate + (valid ? run : 0) * (size_t)I.slab_words;
    uint32_t acc = 0;
    const uint32_t wend = min(32u, I.mwords - g * 32u);
    for (uint32_t wi = 0; wi < wend; wi++) {
        const uint32_t w = g * 32u + wi;
        const uint32_t c = w * 32u + lane;
        uint32_t viol = 0;  // bit ch: clause c violated in chain 32*slice + ch
        if (c < I.m) {
            uint32_t s, e;
            if (I.uniform_k) {
                s = c * I.uniform_k;
                e = s + I.uniform_k;
            } else {
                s = __ldg(I.row_ptr + c);
                e = __ldg(I.row_ptr + c + 1);
            }
            viol = 0xffffffffu;  // empty clause: violated everywhere (lib.rs:13-18)
            for (uint32_t j = s; j < e && viol; j++) {
                const uint32_t l = __ldg(I.lits + j);
                const uint32_t av = __ldg(tb + (l >> 1));
                viol &= (l & 1u) ? av : ~av;  // literal false <=> value == neg bit
            }
        }
        uint32_t mine = 0;  // transpose: lane ch collects the word of chain ch
#pragma unroll
        for (int b = 0; b < 32; b++) {
            const uint32_t word = __ballot_sync(0xffffffffu, (viol >> b) & 1u);
            if (lane == b) mine = word;
        }
        if (valid) {
            slab[I.off_v + w] = mine;
            acc += __popc(mine);
        }
    }
    if (valid) slab[I.off_l1 + g] = acc;
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
