// walk_k3.cuh -- the fast walk kernel: one warp per chain, chain state in shared memory,
// instances whose clauses have at most 3 literals (random 3-SAT: BASELINE configs 1, 2, 4),
// no duplicate clause contents, no variable twice in a clause, at most 65,536 clauses.
//
// Same algorithm and the same random draws as walk_generic.cuh (reference
// rust/src/lib.rs:201-317, trajectory-identical); what changes is the data layout and the
// instruction diet, because this loop is issue-bound (profiles/r01_*):
//   crecx[c] = 64 bytes {lit0,lit1,lit2,k | occurrence ranges of the 3 variables | the clause's flip
//              probabilities, already reduced to what the rule needs}: everything a step needs in ONE
//              broadcast load (built per set of weights by build_crecx_kernel, see there);
//   orec[o]  = {clause, other lit a, other lit b, neg}: the clause content travels with the
//              occurrence entry, so re-evaluating the clauses of the flipped variable
//              (lib.rs:278-287) is one coalesced 16-byte load per lane and two shared-memory
//              bit tests -- no row_ptr / lits gathers.  An absent literal (k < 3) is encoded as
//              the positive literal of the dummy variable n, whose bit is always 0, so the
//              evaluation is branch-free;
//   shared memory is addressed through the `sm[]` array directly (LDS/ATOMS with an
//   immediate base), never through generic pointers.
// A step makes two dependent trips to L2 (crecx, orec); everything else is registers,
// shuffles and shared memory.
#pragma once
#include "common.h"
#include "philox.cuh"
#include "walk_generic.cuh"

namespace slsb {

static constexpr uint32_t CRECX_QUADS = 4;  // uint4 per clause record (64 bytes: two sectors)

__host__ __device__ inline bool walk_k3_supports(const WalkArgs& a)
{
    return a.inst.orec != nullptr && a.crecx != nullptr && a.inst.l1_pad == 64;
}

// The flip rule of a step only ever looks at a VIOLATED clause, and in a violated clause every literal is
// false: the value of variable j is the literal's negation bit.  The flip probabilities of lib.rs:29-30
// (w if the variable is false, 1-w if it is true) are therefore a static property of the clause, and so is
// everything probSAT derives from them (WeightedIndex's cumulative sums and UniformFloat's scale, lib.rs:71-72).
// They are computed here, once per set of weights, with the same double operations in the same order as the
// per-step code of walk_generic.cuh:
//   crecx[c] = { {l0,l1,l2, k | err<<8}, {os0,ol0,os1,ol1}, {os2,ol2, X}, {Y, Z} }   (X, Y, Z doubles)
//   probSAT:  X = scale, Y = cum0, Z = cum1;  err = DERR_WEIGHTS / DERR_ALL_ZERO if WeightedIndex::new fails
//   MT rules: X = fp0,   Y = fp1,  Z = fp2    (absent literals: 0)
__device__ __forceinline__ void build_crecx_clause(const uint4* __restrict__ crec, const uint32_t* __restrict__ occ_ptr,
                                                   const double* __restrict__ w, uint4* __restrict__ crecx, uint32_t c, uint32_t n, int algo)
{
    const uint4 cr = crec[c];
    uint32_t lit[3] = {cr.x, cr.y, cr.z};
    uint32_t os[3], ol[3];
    double fp[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        os[j] = ol[j] = 0;
        fp[j] = 0.0;
        if (lit[j] != LIT_NONE) {
            const uint32_t v = lit[j] >> 1;
            os[j] = occ_ptr[v];
            ol[j] = occ_ptr[v + 1] - os[j];
            const double wv = w[v];
            fp[j] = (lit[j] & 1u) ? 1.0 - wv : wv;  // negative literal false <=> variable true
        } else {
            lit[j] = n << 1;  // dummy variable n: always-false positive literal
        }
    }
    double X = fp[0], Y = fp[1], Z = fp[2];
    uint32_t err = 0;
    if (algo == ALGO_PROBSAT) {
        const double cum0 = fp[0], cum1 = cum0 + fp[1], total = cum1 + fp[2];
        if (!(fp[0] >= 0.0) || !(fp[1] >= 0.0) || !(fp[2] >= 0.0))
            err = DERR_WEIGHTS;
        else if (total == 0.0)
            err = DERR_ALL_ZERO;
        // UniformFloat::new(0, total): the largest scale with scale * (1 - 2^-52) < total
        double scale = total;
        if (!err)
            while (scale * (1.0 - 0x1p-52) >= total) scale = __longlong_as_double(__double_as_longlong(scale) - 1);
        X = scale;
        Y = cum0;
        Z = cum1;
        if (err) {  // the step that meets this clause still runs (see the walk): make it pick literal 0, which exists
            X = 0.0;
            Y = Z = __longlong_as_double(0x7ff0000000000000ll);
        }
    }
    uint4* out = crecx + CRECX_QUADS * (size_t)c;
    out[0] = make_uint4(lit[0], lit[1], lit[2], cr.w | err << 8);
    out[1] = make_uint4(os[0], ol[0], os[1], ol[1]);
    out[2] = make_uint4(os[2], ol[2], (uint32_t)__double2loint(X), (uint32_t)__double2hiint(X));
    out[3] = make_uint4((uint32_t)__double2loint(Y), (uint32_t)__double2hiint(Y), (uint32_t)__double2loint(Z),
                        (uint32_t)__double2hiint(Z));
}

__global__ void build_crecx_kernel(const uint4* __restrict__ crec, const uint32_t* __restrict__ occ_ptr,
                                   const double* __restrict__ w, uint4* __restrict__ crecx, uint32_t m, uint32_t n, int algo)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < m) build_crecx_clause(crec, occ_ptr, w, crecx, c, n, algo);
}

// Batched form (sls_run_batch): the tables of every item of a launch group in ONE launch.  blk_begin[i] = first 256-clause
// block of item i (n_items + 1 entries).  An item carries either a clause-record pool (k <= 3 kernels) or the generic
// kernels' flip tables; Schoening's rule needs neither.
__global__ void build_tables_batch_kernel(const WalkArgs* __restrict__ items, const uint32_t* __restrict__ blk_begin, uint32_t n_items)
{
    const uint32_t it = find_item(blk_begin, n_items, blockIdx.x);
    const WalkArgs& a = items[it];
    const uint32_t c = (blockIdx.x - __ldg(blk_begin + it)) * blockDim.x + threadIdx.x;
    if (c >= a.inst.m) return;
    if (a.crecx)
        build_crecx_clause(a.inst.crec, a.inst.occ_ptr, a.w, const_cast<uint4*>(a.crecx), c, a.inst.n, a.algo);
    else if (a.ftab)
        build_flip_tables_clause(a.inst, a.w, a.algo, const_cast<double*>(a.ftab), const_cast<double*>(a.cscale), c);
}

// inclusive warp scan, two instructions per level (shuffle with predicate + predicated add)
__device__ __forceinline__ uint32_t warp_incl_scan_p(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        asm volatile(
            "{ .reg .pred p; .reg .u32 t;\n"
            "  shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n"
            "  @p add.u32 %0, %0, t; }\n"
            : "+r"(v)
            : "r"(d));
    }
    return v;
}

__device__ __forceinline__ uint32_t ffs0(uint32_t mask) { return __ffs(mask) - 1; }
// index of the highest set bit (one FLO)
__device__ __forceinline__ uint32_t bfind(uint32_t mask)
{
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(mask));
    return r;
}

// Shared memory through explicit shared-space instructions on a 32-bit byte address: the hot loop
// must never see a generic pointer to shared memory (the generic->shared conversion costs an
// S2R SR_CgaCtaId + LEA per access on sm_100).
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
// the same store issued by the lanes with `on` set only, as ONE predicated instruction (no branch region)
__device__ __forceinline__ void sts32_if(uint32_t a, uint32_t v, bool on)
{
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(a), "r"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void atoms_xor(uint32_t a, uint32_t v)
{
    asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void atoms_add(uint32_t a, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// a : b on a predicate, as one SEL (the compiler otherwise turns three selects on one condition into a branch)
__device__ __forceinline__ uint32_t selp(bool p, uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.u32 %0, %1, %2, q; }" : "=r"(r) : "r"(a), "r"(b), "r"((uint32_t)p));
    return r;
}

// The chain's walk stream (same word positions, alignment rules and distributions as WalkStream /
// WarpStream in philox.cuh) buffered in a 128-word ring in shared memory: on a refill every lane
// evaluates one Philox block and stores its four words; a draw is then a single broadcast LDS.
static constexpr uint32_t RING_WORDS = 128;
struct RingStream {
    ChainKey ck;
    uint64_t base;  // word position of ring[0], multiple of 4
    uint32_t cur;   // cursor: shared-window byte address of the next word
    uint32_t ring;  // shared-window byte address of the ring (16-byte aligned)
    int lane;
    __device__ __forceinline__ RingStream(uint64_t seed, uint64_t chain, uint32_t ring_addr, int lane_)
        : ck(seed, chain, STREAM_WALK), base(0), cur(ring_addr), ring(ring_addr), lane(lane_)
    {
        fill();
    }
    // fewer than `words` words left in the ring?
    __device__ __forceinline__ bool low(uint32_t words) const { return cur > ring + (RING_WORDS - words) * 4u; }
    __device__ __forceinline__ void refill()
    {
        const uint32_t off = (cur - ring) >> 2;
        base += off & ~3u;
        cur = ring + (off & 3u) * 4u;
        fill();
    }
    __device__ __forceinline__ void fill()
    {
        uint2 k = ck.key;
        asm volatile("" : "+r"(k.x), "+r"(k.y));  // keep the key schedule inside this cold path
        const uint64_t b = (base >> 2) + lane;
        const uint4 x = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), ck.chain_lo, ck.chain_hi_stream), k);
        __syncwarp();  // earlier readers of the ring are done
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ring + lane * 16u), "r"(x.x), "r"(x.y), "r"(x.z), "r"(x.w) : "memory");
        __syncwarp();
    }
    // the *_unchecked draws rely on the caller having made sure the ring holds the words (top of the step)
    __device__ __forceinline__ uint32_t next_u32_unchecked()
    {
        const uint32_t v = lds32(cur);
        cur += 4u;
        return v;
    }
    __device__ __forceinline__ uint32_t next_u32()
    {
        if (low(1)) refill();
        return next_u32_unchecked();
    }
    __device__ __forceinline__ uint64_t next_u64_unchecked()
    {
        cur = (cur + 4u) & ~7u;  // 64-bit draws sit on even word positions
        const uint2 v = lds64(cur);
        cur += 8u;
        return (uint64_t)v.y << 32 | v.x;
    }
    __device__ __forceinline__ uint64_t next_u64()
    {
        cur = (cur + 4u) & ~7u;
        if (low(2)) refill();
        const uint2 v = lds64(cur);
        cur += 8u;
        return (uint64_t)v.y << 32 | v.x;
    }
    __device__ __forceinline__ void skip_u64() { cur = ((cur + 4u) & ~7u) + 8u; }
    __device__ __forceinline__ double next_f64() { return (double)(next_u64() >> 11) * 0x1p-53; }
    __device__ __forceinline__ double next_f64_unchecked() { return (double)(next_u64_unchecked() >> 11) * 0x1p-53; }
    __device__ __forceinline__ uint64_t uniform_u64(uint64_t range)
    {
        const uint64_t zone = (range << __clzll(range)) - 1ull;
        for (;;) {
            const uint64_t v = next_u64();
            if (v * range <= zone) return __umul64hi(v, range);
        }
    }
};

// `block` = index of this CTA among the CTAs of the item (instance) described by A.
// ALGO and GREEDY (prob_flip_best > 0) are compile-time: every instantiation carries only its own flip rule,
// which keeps the step loop short enough for the instruction cache and frees registers.
template <int ALGO, bool GREEDY, bool TRAJ>
__device__ __forceinline__ void walk_k3s_body(const WalkArgs& A, uint32_t block, uint32_t* sm)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t run = (uint64_t)block * A.warps_per_block + warp;
    if (run >= A.nruns) return;
    const DevInst& I = A.inst;
    // byte addresses in the shared window
    uint32_t sA = (uint32_t)__cvta_generic_to_shared(sm) + warp * I.slab_words * 4u;  // abits
    uint32_t sV = sA + I.off_v * 4u;   // vbits (zero-padded to a multiple of 32 words)
    uint32_t sL = sA + I.off_l1 * 4u;  // 64 level-1 counters
    // opaque to the optimiser: otherwise the shared-window base (S2UR SR_CgaCtaId + ULEA) is rematerialised
    // at several places inside the step loop instead of living in three registers
    asm volatile("" : "+r"(sA), "+r"(sV), "+r"(sL));
    const uint64_t chain = A.chain_offset + run;

    // literal code l = 2*var + neg is TRUE iff bit(var) != neg
    auto lit_true = [sA](uint32_t l) -> uint32_t { return (__funnelshift_r(lds32(sA + (l >> 6) * 4u), 0u, l >> 1) ^ l) & 1u; };

    // ---- chain init: assignment from the init stream (lib.rs:206-210) ...
    for (uint32_t i = lane; i < I.slab_words; i += 32) sts32(sA + i * 4u, 0u);
    __syncwarp();
    {
        const ChainKey ck(A.seed, chain, STREAM_INIT);
        for (uint32_t w = lane; w < I.nwords; w += 32) {
            uint32_t bits = 0;
#pragma unroll 4
            for (uint32_t b = 0; b < 16; b++) {
                const uint32_t v0 = w * 32u + 2u * b;
                if (v0 >= I.n) break;
                const uint4 x = ck.block((uint64_t)(v0 >> 1));
                const uint64_t t0 = __ldg(A.thr_init + v0);
                const uint64_t u0 = (uint64_t)x.y << 32 | x.x;
                bits |= (uint32_t)(t0 == ~0ull || u0 < t0) << (2u * b);
                if (v0 + 1 < I.n) {
                    const uint64_t t1 = __ldg(A.thr_init + v0 + 1);
                    const uint64_t u1 = (uint64_t)x.w << 32 | x.z;
                    bits |= (uint32_t)(t1 == ~0ull || u1 < t1) << (2u * b + 1u);
                }
            }
            sts32(sA + w * 4u, bits);
        }
        __syncwarp();
    }
    // ... and the full scan (lib.rs:213-214, find_violated_clauses :176-181): one clause per lane,
    // a ballot assembles each vbits word, the level-1 counter is the popcount sum of its 32 words
    uint32_t nf = 0;
    const uint4* __restrict__ crec = I.crec;
    for (uint32_t g = 0; g < I.ngroups; g++) {
        uint32_t acc = 0, mine = 0;
        const uint32_t wend = min(32u, I.mwords - g * 32u);
        for (uint32_t wi = 0; wi < wend; wi++) {
            const uint32_t c = (g * 32u + wi) * 32u + lane;
            bool viol = false;
            if (c < I.m) {
                const uint4 cr = __ldg(crec + c);
                viol = !lit_true(cr.x) && (cr.w < 2 || !lit_true(cr.y)) && (cr.w < 3 || !lit_true(cr.z));
            }
            const uint32_t word = __ballot_sync(FULL, viol);
            if ((uint32_t)lane == wi) mine = word;
            acc += __popc(word);
        }
        if ((uint32_t)lane < wend) sts32(sV + (g * 32u + lane) * 4u, mine);
        if (lane == 0) sts32(sL + g * 4u, acc);
        nf += acc;
    }
    __syncwarp();

    // the rings of the CTA's chains sit behind their slabs
    RingStream rng(A.seed, chain, (uint32_t)__cvta_generic_to_shared(sm) + (A.warps_per_block * I.slab_words + warp * RING_WORDS) * 4u, lane);
    uint32_t best = I.m;
    uint32_t* best_bits = A.best_bits + run * (size_t)I.nwords;
    auto save_best = [&]() {
        best = nf;
        for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = lds32(sA + w * 4u);
    };
    if (nf < best) save_best();
    uint32_t* traj = TRAJ ? A.traj + run * (A.nsteps + 1) : nullptr;
    if (TRAJ && lane == 0) traj[0] = nf;

    // probSAT, Schoening and the greedy move flip exactly one variable per step: no separate counter
    constexpr bool COUNT_FLIPS = ALGO == ALGO_MOSER || ALGO == ALGO_MOSER_SINGLE;
    uint64_t steps = 0, flips = 0;
    int err = 0;
    const bool mapping = A.mapping != 0;
    const double pfb = A.pfb;
    const uint64_t budget = A.nsteps;
    const uint4* __restrict__ crecx = A.crecx;
    const uint4* __restrict__ orec = I.orec;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // 64-bit step budget, 32-bit counter inside a chunk (the loop test and the increment stay single instructions)
    while (nf > 0 && steps < budget) {
        uint32_t lim = (uint32_t)min(budget - steps, (uint64_t)0x40000000u), i = 0;
        while (nf > 0 && i < lim) {
            i++;  // (the ring holds at least 16 words here: a step consumes at most 13 outside rejection loops)

            // ---- lib.rs:229: uniform pick among the violated clauses (UniformInt: widening multiply, the low half
            // is rejected above `zone` -- up to half of the draws when nf is just above a power of two)
            const uint32_t zone1 = nf << __clz(nf);  // rand's zone + 1 (top bit set, so no wrap)
            uint32_t pick = rng.next_u32_unchecked();
            while (pick * nf >= zone1) {
                if (rng.low(14)) rng.refill();
                pick = rng.next_u32_unchecked();
            }
            uint32_t r = __umulhi(pick, nf);

            // lib.rs:231: one f64 is drawn on every step, greedy or not
            bool greedy = false;
            if (GREEDY)
                greedy = rng.next_f64_unchecked() < pfb;
            else
                rng.skip_u64();
            // probSAT's single draw (lib.rs:72) does not depend on the clause: fetch and convert it now so that
            // the conversion overlaps the rank-select below
            double u_ps = 0.0;
            if (ALGO == ALGO_PROBSAT && !greedy) u_ps = (double)(rng.next_u64_unchecked() >> 12) * 0x1p-52;
            asm volatile("" : "+r"(r), "+d"(u_ps));  // pins the conversion before the rank-select instead of after the L2 trip

            // level A: lane owns counters 2*lane, 2*lane+1
            const uint2 cc = lds64(sL + lane * 8u);
            const uint32_t both = cc.x + cc.y;
            const uint32_t inclA = warp_incl_scan_p(both);
            const uint32_t srcA = __popc(__ballot_sync(FULL, r >= inclA));  // lanes below the crossing form a prefix
            r -= __shfl_sync(FULL, inclA - both, srcA);
            const uint32_t first = __shfl_sync(FULL, cc.x, srcA);
            const bool second = r >= first;
            const uint32_t g = 2 * srcA + (second ? 1u : 0u);
            if (second) r -= first;
            // level B: the 32 words of group g
            const uint32_t word = lds32(sV + (g * 32u + lane) * 4u);
            const uint32_t pc = __popc(word);
            const uint32_t inclB = warp_incl_scan_p(pc);
            const uint32_t srcB = __popc(__ballot_sync(FULL, r >= inclB));
            r -= __shfl_sync(FULL, inclB - pc, srcB);
            const uint32_t wsel = __shfl_sync(FULL, word, srcB);
            const bool hit = ((wsel >> lane) & 1u) && __popc(wsel & lt_mask) == r;
            const uint32_t c = (g * 32u + srcB) * 32u + bfind(__ballot_sync(FULL, hit));  // exactly one lane hits

            // ---- the clause record: 16-byte quads, same address in every lane, issued together
            const uint4* rec = crecx + CRECX_QUADS * (size_t)c;
            const uint4 q0 = __ldg(rec);
            const uint4 q1 = __ldg(rec + 1);
            const uint4 q2 = __ldg(rec + 2);
            const uint4 q3 = __ldg(rec + 3);
            const uint32_t k = q0.w & 0xffu;
            const double dX = __hiloint2double((int)q2.w, (int)q2.z);
            const double dY = __hiloint2double((int)q3.y, (int)q3.x);
            const double dZ = __hiloint2double((int)q3.w, (int)q3.z);

            // Current values of the clause's variables.  The clause is violated, so each of its literals is false:
            // the variable's value equals the literal's negation bit (the dummy literal of a short clause is
            // positive and its variable reads 0) -- no shared-memory read.
            const uint32_t a0 = q0.x & 1u, a1 = q0.y & 1u, a2 = q0.z & 1u;

            // Re-evaluation of every clause of a flipped variable (lib.rs:278-287): one occurrence record per lane.
            // A clause never holds a variable twice, so only OTHER variables' bits are read here.  Lanes past the end
            // of the list re-read its last record (same sector, same banks) and are masked out.
            auto update_round = [&](uint32_t os, uint32_t len, uint32_t av, uint32_t base) {
                const uint32_t o = base + lane;
                const bool act = o < len;
                const uint4 orc = __ldg(orec + (os + min(o, len - 1u)));  // one 32-bit index, one widening multiply-add
                const uint32_t va = sV + (orc.x >> 5) * 4u;
                const uint32_t vword = lds32(va);
                const uint32_t bit = 1u << (orc.x & 31);
                const bool viol = act & (av == orc.w) & !(lit_true(orc.y) | lit_true(orc.z));
                const bool cur = act & ((vword & bit) != 0);
                const bool ins = viol && !cur;  // HashSet::insert of a new element
                const bool rem = !viol && cur;  // HashSet::remove of a present element
                if (ins | rem) {
                    atoms_xor(va, bit);
                    atoms_add(sL + (orc.x >> 10) * 4u, ins ? 1u : 0xffffffffu);
                }
                nf += __popc(__ballot_sync(FULL, ins)) - __popc(__ballot_sync(FULL, rem));
            };
            auto update_var = [&](uint32_t os, uint32_t len, uint32_t av) {
                update_round(os, len, av, 0);
                if (len > 32)  // a variable with more than 32 occurrences: rare on the instances this kernel serves
                    for (uint32_t base = 32; base < len; base += 32) update_round(os, len, av, base);
                __syncwarp();
            };

            if (ALGO == ALGO_MOSER && !(GREEDY && greedy)) {
                // lib.rs:24-35: one f64 per literal, in literal order; every marked variable flips
                const double fp0 = dX, fp1 = dY, fp2 = dZ;
                uint32_t marks = 0;  // bit j <=> the variable of literal j flips in this step
                if (rng.next_f64_unchecked() < fp0) marks |= 1u;
                if (k > 1 && rng.next_f64_unchecked() < fp1) marks |= 2u;
                if (k > 2 && rng.next_f64_unchecked() < fp2) marks |= 4u;
                if (marks) {
                    // lib.rs:37-41: toggle the marked variables (distinct on this path); lane j owns literal j
                    const uint32_t lj = selp(lane == 0, q0.x, selp(lane == 1, q0.y, q0.z));
                    if (lane < 3 && ((marks >> lane) & 1u)) atoms_xor(sA + (lj >> 6) * 4u, 1u << ((lj >> 1) & 31u));
                    __syncwarp();  // several variables flip: the re-evaluation below reads the other ones
                    if (COUNT_FLIPS) flips += __popc(marks);
                    for (uint32_t mk = marks; mk; mk &= mk - 1) {
                        const uint32_t j = ffs0(mk);
                        update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)),
                                   selp(j == 0, a0, selp(j == 1, a1, a2)) ^ 1u);
                    }
                }
            } else {
                // every other rule flips at most one variable: literal j of the clause
                uint32_t j = 3;  // 3 = nothing to flip (MT-single without a mark)
                if (GREEDY && greedy) {
                    // lib.rs:236-256: score = violations after tentatively flipping the variable; first strict minimum
                    uint64_t min_viol = ~0ull;
                    const uint32_t aj[3] = {a0, a1, a2};
                    const uint32_t osj[3] = {q1.x, q1.z, q2.x}, olj[3] = {q1.y, q1.w, q2.y};
    #pragma unroll
                    for (int t = 0; t < 3; t++) {
                        if ((uint32_t)t < k) {
                            const uint32_t av = aj[t] ^ 1u;
                            uint32_t after = 0, before = 0;
                            for (uint32_t base = 0; base < olj[t]; base += 32) {
                                const uint32_t o = base + lane;
                                bool ca = false, cb = false;
                                if (o < olj[t]) {
                                    const uint4 orc = __ldg(orec + osj[t] + o);
                                    ca = (av == orc.w) && !lit_true(orc.y) && !lit_true(orc.z);
                                    cb = (lds32(sV + (orc.x >> 5) * 4u) >> (orc.x & 31)) & 1u;
                                }
                                after += __popc(__ballot_sync(FULL, ca));
                                before += __popc(__ballot_sync(FULL, cb));
                            }
                            const uint64_t viol = mapping ? (uint64_t)after : (uint64_t)nf - before + after;
                            if (viol < min_viol) {
                                min_viol = viol;
                                j = t;
                            }
                        }
                    }
                } else if (ALGO == ALGO_SCHOENING) {
                    uint32_t v;
                    const uint32_t zk = (k << __clz(k)) - 1u;  // lib.rs:52 clause.choose(rng)
                    do v = rng.next_u32(); while (v * k > zk);
                    j = __umulhi(v, k);
                } else {
                    if (ALGO == ALGO_PROBSAT) {
                        // lib.rs:71-72 on the precomputed record: WeightedIndex::new's verdict, then one uniform
                        // sample in [0, total) and partition_point(cum <= x) over the cumulative weights (x < total, and
                        // an absent literal repeats the previous sum, so no test on k is needed)
                        // (no branch on the verdict here: it would split the record's four loads into two dependent
                        // trips to L2.  A failed clause ends the walk through the step budget; the call then fails
                        // as a whole and nothing of this chain is returned.)
                        if (q0.w > 0xffu) {
                            err = (int)(q0.w >> 8);
                            lim = i;
                        }
                        const double x = u_ps * dX;
                        j = (dY <= x ? 1u : 0u) + (dZ <= x ? 1u : 0u);
                    } else {
                        // MT-single, lib.rs:81-99: mark like MT, then flip one uniformly chosen marked variable
                        uint32_t marks = 0;
                        if (rng.next_f64_unchecked() < dX) marks |= 1u;
                        if (k > 1 && rng.next_f64_unchecked() < dY) marks |= 2u;
                        if (k > 2 && rng.next_f64_unchecked() < dZ) marks |= 4u;
                        if (marks) j = select_in_word(marks, (uint32_t)rng.uniform_u64(__popc(marks)));
                    }
                }
                if (ALGO != ALGO_MOSER_SINGLE || j < 3) {
                    const uint32_t lit = selp(j == 0, q0.x, selp(j == 1, q0.y, q0.z));
                    // lib.rs:37-41.  Lane 0 alone reads and rewrites the word (one predicated store: no branch region, no
                    // atomic, and no assumption about the lanes running in lock step); the word is read again only after
                    // the barrier that closes the re-evaluation
                    const uint32_t wa = sA + (lit >> 6) * 4u;
                    sts32_if(wa, lds32(wa) ^ (1u << ((lit >> 1) & 31u)), lane == 0);
                    if (COUNT_FLIPS) flips++;
                    // no barrier: the re-evaluation never reads the flipped variable itself, and it ends with one
                    update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)), (lit & 1u) ^ 1u);
                }
            }

            // one rarely taken region for both housekeeping jobs: a new best assignment (lib.rs:292-296) and the
            // ring refill that guarantees the next step its 16 words
            if ((nf < best) | rng.low(16)) {
                if (nf < best) save_best();
                if (rng.low(16)) rng.refill();
            }
            if (TRAJ && lane == 0) traj[steps + i] = nf;       // lib.rs:298-300
        }
        steps += i;
        if (err) break;
    }

    if (lane == 0) {
        if (err) atomicCAS(A.err, 0, err);
        A.found[run] = nf == 0 && !err;
        A.has_best[run] = best < I.m;  // best starts at m (lib.rs:216) and every improvement is saved
        A.steps[run] = steps;
        A.best_energy[run] = best;
        A.flips[run] = COUNT_FLIPS ? flips : steps;
    }
}

template <int ALGO, bool GREEDY, bool TRAJ>
__global__ void __launch_bounds__(128, 7) walk_k3s_kernel(const WalkArgs A)
{
    extern __shared__ uint32_t sm[];
    walk_k3s_body<ALGO, GREEDY, TRAJ>(A, blockIdx.x, sm);
}

// the batched form never records trajectories (sls_run_batch refuses them)
template <int ALGO, bool GREEDY>
__global__ void __launch_bounds__(128, 9) walk_k3s_batch_kernel(const WalkArgs* __restrict__ items, const uint32_t* __restrict__ cta_begin,
                                                                uint32_t n_items)
{
    extern __shared__ uint32_t sm[];
    __shared__ WalkArgs item;
    const uint32_t it = find_item(cta_begin, n_items, blockIdx.x);
    load_item(item, items + it);
    walk_k3s_body<ALGO, GREEDY, false>(item, blockIdx.x - __ldg(cta_begin + it), sm);
}

// Host-side choice of the instantiation for (algo, prob_flip_best > 0, trajectories).
using WalkK3Kernel = void (*)(const WalkArgs);
using WalkK3BatchKernel = void (*)(const WalkArgs*, const uint32_t*, uint32_t);

template <bool GREEDY, bool TRAJ>
inline WalkK3Kernel walk_k3s_kernel_for_algo(int algo)
{
    switch (algo) {
        case ALGO_MOSER: return walk_k3s_kernel<ALGO_MOSER, GREEDY, TRAJ>;
        case ALGO_MOSER_SINGLE: return walk_k3s_kernel<ALGO_MOSER_SINGLE, GREEDY, TRAJ>;
        case ALGO_SCHOENING: return walk_k3s_kernel<ALGO_SCHOENING, GREEDY, TRAJ>;
        default: return walk_k3s_kernel<ALGO_PROBSAT, GREEDY, TRAJ>;
    }
}
inline WalkK3Kernel walk_k3s_kernel_for(int algo, double pfb, bool traj)
{
    if (pfb > 0.0) return traj ? walk_k3s_kernel_for_algo<true, true>(algo) : walk_k3s_kernel_for_algo<true, false>(algo);
    return traj ? walk_k3s_kernel_for_algo<false, true>(algo) : walk_k3s_kernel_for_algo<false, false>(algo);
}

template <bool GREEDY>
inline WalkK3BatchKernel walk_k3s_batch_kernel_for_algo(int algo)
{
    switch (algo) {
        case ALGO_MOSER: return walk_k3s_batch_kernel<ALGO_MOSER, GREEDY>;
        case ALGO_MOSER_SINGLE: return walk_k3s_batch_kernel<ALGO_MOSER_SINGLE, GREEDY>;
        case ALGO_SCHOENING: return walk_k3s_batch_kernel<ALGO_SCHOENING, GREEDY>;
        default: return walk_k3s_batch_kernel<ALGO_PROBSAT, GREEDY>;
    }
}
inline WalkK3BatchKernel walk_k3s_batch_kernel_for(int algo, double pfb)
{
    return pfb > 0.0 ? walk_k3s_batch_kernel_for_algo<true>(algo) : walk_k3s_batch_kernel_for_algo<false>(algo);
}

}  // namespace slsb
