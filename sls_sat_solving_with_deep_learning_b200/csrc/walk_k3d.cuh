// walk_k3d.cuh -- two chains per warp, for batches of small k<=3 instances (BASELINE configs 1 and 4).
//
// In a batch (thousands of instances x 100 chains) the machine runs many waves of chains, so what counts is the
// number of warp instructions per chain step, not the latency of one chain.  A small instance cannot feed 32 lanes:
// a variable has ~13 occurrences, its counters fit in a few lanes.  Here a warp carries TWO chains of the same
// instance, one per half-warp, in lock step: every collective is issued once with the full mask and each half reads
// its own 16 lanes of the result (width-16 shuffles, half of the ballot), so a step of both chains costs about one
// step's worth of instructions.  (Per-half lane masks were measured and rejected: every collective with a
// non-uniform mask compiles to a MATCH.ANY / BRA.DIV sequence, see DESIGN.md.)
//
// The same body with LANES = 8 carries FOUR chains per warp (quarter-warps); it serves the smallest instances (at most
// four level-1 counters = 4096 clauses: configs 1 and 4), where an occurrence list (~13 entries) takes two rounds of 8
// lanes and the 32 words of a counter group are four per lane.
//
// Same records (crecx / orec), same random streams and the same trajectory per chain as walk_k3.cuh.
// Restrictions (the caller falls back to walk_k3s_batch_kernel otherwise): at most 16 level-1 counters
// (16,384 clauses), rules "moser" and "probsat" without the greedy mix, no trajectories.
#pragma once
#include "common.h"
#include "philox.cuh"
#include "walk_k3.cuh"

namespace slsb {

static constexpr uint32_t HRING_WORDS = 64;  // a half-warp refills 16 Philox blocks at a time

__host__ __device__ inline bool walk_k3d_supports(const WalkArgs& a)
{
    return a.inst.orec != nullptr && a.crecx != nullptr && a.inst.ngroups <= 16 && a.pfb <= 0.0 && !a.want_traj &&
           (a.algo == ALGO_MOSER || a.algo == ALGO_PROBSAT);
}

__host__ __device__ inline bool walk_k3q_supports(const WalkArgs& a) { return walk_k3d_supports(a) && a.inst.ngroups <= 4; }

// inclusive scan over the LANES lanes of a sub-warp (shfl.up clamp operand: (32 - LANES) << 8)
template <int LANES>
__device__ __forceinline__ uint32_t sub_incl_scan(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
        asm volatile(
            "{ .reg .pred p; .reg .u32 t;\n"
            "  shfl.sync.up.b32 t|p, %0, %1, %2, 0xffffffff;\n"
            "  @p add.u32 %0, %0, t; }\n"
            : "+r"(v)
            : "r"(d), "n"((32 - LANES) << 8));
    }
    return v;
}

// The walk stream of one chain, buffered in a 64-word ring that the chain's 16 lanes refill together.
// Every method that contains a barrier is called by all 32 lanes of the warp (both chains), with `need` telling
// which half actually refills.
template <int LANES>
struct HalfRing {
    static constexpr int BLOCKS = 16 / LANES;  // Philox blocks per lane and refill
    ChainKey ck;
    uint64_t base;  // word position of ring[0], multiple of 4
    uint32_t cur;   // shared-window byte address of the next word
    uint32_t ring;  // shared-window byte address of the ring (16-byte aligned)
    int l16;
    __device__ __forceinline__ HalfRing(uint64_t seed, uint64_t chain, uint32_t ring_addr, int l16_)
        : ck(seed, chain, STREAM_WALK), base(0), cur(ring_addr), ring(ring_addr), l16(l16_)
    {
        fill_together(true);
    }
    __device__ __forceinline__ bool low(uint32_t words) const { return cur > ring + (HRING_WORDS - words) * 4u; }
    __device__ __forceinline__ void fill_together(bool need)
    {
        uint4 x[BLOCKS];
        if (need) {
            uint2 k = ck.key;
            asm volatile("" : "+r"(k.x), "+r"(k.y));  // keep the key schedule inside this cold path
#pragma unroll
            for (int i = 0; i < BLOCKS; i++) {
                const uint64_t b = (base >> 2) + l16 + i * LANES;
                x[i] = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(b >> 32), ck.chain_lo, ck.chain_hi_stream), k);
            }
        }
        __syncwarp();  // earlier readers of the rings are done
        if (need) {
#pragma unroll
            for (int i = 0; i < BLOCKS; i++)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ring + (l16 + i * LANES) * 16u), "r"(x[i].x), "r"(x[i].y), "r"(x[i].z),
                             "r"(x[i].w) : "memory");
        }
        __syncwarp();
    }
    __device__ __forceinline__ void refill_together(bool need)
    {
        if (need) {
            const uint32_t off = (cur - ring) >> 2;
            base += off & ~3u;
            cur = ring + (off & 3u) * 4u;
        }
        fill_together(need);
    }
    __device__ __forceinline__ uint32_t next_u32()
    {
        const uint32_t v = lds32(cur);
        cur += 4u;
        return v;
    }
    __device__ __forceinline__ uint64_t next_u64()
    {
        cur = (cur + 4u) & ~7u;  // 64-bit draws sit on even word positions
        const uint2 v = lds64(cur);
        cur += 8u;
        return (uint64_t)v.y << 32 | v.x;
    }
    __device__ __forceinline__ void skip_u64() { cur = ((cur + 4u) & ~7u) + 8u; }
    __device__ __forceinline__ double next_f64() { return (double)(next_u64() >> 11) * 0x1p-53; }
};

// `block` = index of this CTA among the CTAs of the item; a CTA of W warps carries 2W chains
template <int ALGO, int LANES>
__device__ __forceinline__ void walk_k3d_body(const WalkArgs& A, uint32_t block, uint32_t* sm)
{
    constexpr uint32_t CH = 32 / LANES;    // chains per warp
    constexpr uint32_t PER = 32 / LANES;   // bit positions / words / clauses a lane owns where a full warp owns one
    constexpr uint32_t SUBMASK = LANES == 16 ? 0xffffu : 0xffu;
    const int lane = threadIdx.x & 31;
    const int l16 = lane & (LANES - 1);    // lane within the chain's sub-warp
    const uint32_t half = (uint32_t)lane / LANES, hshift = half * LANES;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t slot = warp * CH + half;  // this chain's slot in the CTA
    const uint64_t run = ((uint64_t)block * A.warps_per_block + warp) * CH + half;
    const bool valid = run < A.nruns;
    if (!__any_sync(FULL, valid)) return;  // (warp-uniform)
    const DevInst& I = A.inst;
    const uint32_t nslots = A.warps_per_block * CH;
    uint32_t sA = (uint32_t)__cvta_generic_to_shared(sm) + slot * I.slab_words * 4u;  // abits
    uint32_t sV = sA + I.off_v * 4u;                                                   // vbits
    uint32_t sL = sA + I.off_l1 * 4u;                                                  // level-1 counters
    asm volatile("" : "+r"(sA), "+r"(sV), "+r"(sL));  // keep the three bases in registers (see walk_k3.cuh)
    const uint64_t chain = A.chain_offset + (valid ? run : 0);

    auto hballot = [hshift](bool p) -> uint32_t { return (__ballot_sync(FULL, p) >> hshift) & SUBMASK; };
    const uint32_t hmask = SUBMASK << hshift;
    auto hcount = [hmask](bool p) -> uint32_t { return __popc(__ballot_sync(FULL, p) & hmask); };  // votes in this half
    auto hshfl = [](uint32_t v, uint32_t src) -> uint32_t { return __shfl_sync(FULL, v, src, LANES); };
    // literal code l = 2*var + neg is TRUE iff bit(var) != neg
    auto lit_true = [sA](uint32_t l) -> uint32_t { return (__funnelshift_r(lds32(sA + (l >> 6) * 4u), 0u, l >> 1) ^ l) & 1u; };

    // ---- chain init: assignment from the init stream (lib.rs:206-210) ...
    for (uint32_t i = l16; i < I.slab_words; i += LANES) sts32(sA + i * 4u, 0u);
    __syncwarp();
    {
        const ChainKey ck(A.seed, chain, STREAM_INIT);
        for (uint32_t w = l16; w < I.nwords; w += LANES) {
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
    // ... and the full scan (lib.rs:213-214): a lane evaluates clauses l16, l16 + LANES, ... of every vbits word
    uint32_t nf = 0;
    const uint4* __restrict__ crec = I.crec;
    for (uint32_t g = 0; g < I.ngroups; g++) {
        uint32_t acc = 0;
        const uint32_t wend = min(32u, I.mwords - g * 32u);
        for (uint32_t wi = 0; wi < wend; wi++) {
            bool viol[PER];
#pragma unroll
            for (uint32_t h = 0; h < PER; h++) {
                const uint32_t c = (g * 32u + wi) * 32u + h * LANES + l16;
                viol[h] = false;
                if (c < I.m) {
                    const uint4 cr = __ldg(crec + c);
                    viol[h] = !lit_true(cr.x) && (cr.w < 2 || !lit_true(cr.y)) && (cr.w < 3 || !lit_true(cr.z));
                }
            }
            uint32_t word = 0;
#pragma unroll
            for (uint32_t h = 0; h < PER; h++) word |= hballot(viol[h]) << (h * LANES);
            if (l16 == 0) sts32(sV + (g * 32u + wi) * 4u, word);
            acc += __popc(word);
        }
        if (l16 == 0) sts32(sL + g * 4u, acc);
        nf += acc;
    }
    __syncwarp();
    if (!valid) nf = 0;  // an absent chain never steps

    // the rings of the CTA's chains sit behind their slabs
    HalfRing<LANES> rng(A.seed, chain, (uint32_t)__cvta_generic_to_shared(sm) + (nslots * I.slab_words + slot * HRING_WORDS) * 4u, l16);
    uint32_t best = I.m;
    uint32_t* best_bits = A.best_bits + (valid ? run : 0) * (size_t)I.nwords;
    auto save_best = [&]() {
        best = nf;
        for (uint32_t w = l16; w < I.nwords; w += LANES) best_bits[w] = lds32(sA + w * 4u);
    };
    if (valid && nf < best) save_best();

    uint64_t steps = 0, flips = 0;
    int err = 0;
    const uint64_t budget = A.nsteps;
    const uint4* __restrict__ crecx = A.crecx;
    const uint4* __restrict__ orec = I.orec;

    // Re-evaluation of the clauses of a flipped variable (lib.rs:278-287) for both chains at once: a lane owns one
    // occurrence record of its chain's variable; len = 0 for a half that has nothing to update this round.
    // (RPL = records per lane and round.  Two per lane for quarter-warps -- one trip to L2 for a typical list of ~13
    // entries instead of two dependent ones -- was measured and is slower: 5.35 vs 5.62 G flips/s on the C1-shaped
    // batch; the walk is bound by issue slots there, not by that latency.)
    constexpr uint32_t RPL = 1;
    auto update_var = [&](uint32_t os, uint32_t len, uint32_t av) {
        for (uint32_t base = 0; __any_sync(FULL, base < len); base += RPL * LANES) {
            uint4 orc[RPL];
            bool act[RPL];
#pragma unroll
            for (uint32_t h = 0; h < RPL; h++) {
                const uint32_t o = base + l16 + h * LANES;
                act[h] = o < len;
                orc[h] = __ldg(orec + (os + min(o, max(len, 1u) - 1u)));  // past the end: a valid record, masked
            }
#pragma unroll
            for (uint32_t h = 0; h < RPL; h++) {
                const uint32_t va = sV + (orc[h].x >> 5) * 4u;
                const uint32_t vword = lds32(va);
                const uint32_t bit = 1u << (orc[h].x & 31);
                const bool viol = act[h] & (av == orc[h].w) & !(lit_true(orc[h].y) | lit_true(orc[h].z));
                const bool cur = act[h] & ((vword & bit) != 0);
                const bool ins = viol && !cur;  // HashSet::insert of a new element
                const bool rem = !viol && cur;  // HashSet::remove of a present element
                if (ins | rem) {
                    atoms_xor(va, bit);
                    atoms_add(sL + (orc[h].x >> 10) * 4u, ins ? 1u : 0xffffffffu);
                }
                nf += hcount(ins) - hcount(rem);
            }
        }
        __syncwarp();
    };

    const bool few_groups = I.ngroups <= 4;
    uint32_t lt[PER];  // bits below position l16 + h * LANES
#pragma unroll
    for (uint32_t h = 0; h < PER; h++) lt[h] = (1u << (l16 + h * LANES)) - 1u;
    bool active = nf > 0 && budget > 0;  // per chain
    while (__any_sync(FULL, active)) {
        if (active) steps++;

        // ---- lib.rs:229: uniform pick among the violated clauses (UniformInt zone rejection, see walk_k3.cuh)
        const uint32_t nfs = active ? nf : 1u;
        const uint32_t zone1 = nfs << __clz(nfs);
        uint32_t pick = 0;
        bool rej = false;
        if (active) {
            pick = rng.next_u32();
            rej = pick * nfs >= zone1;
        }
        while (__any_sync(FULL, rej)) {
            const bool need = rej && rng.low(14);
            if (__any_sync(FULL, need)) rng.refill_together(need);
            if (rej) {
                pick = rng.next_u32();
                rej = pick * nfs >= zone1;
            }
        }
        uint32_t r = __umulhi(pick, nfs);

        // lib.rs:231: one f64 is drawn on every step (prob_flip_best = 0 here: never greedy)
        double u_ps = 0.0;
        if (active) {
            rng.skip_u64();
            if (ALGO == ALGO_PROBSAT) u_ps = (double)(rng.next_u64() >> 12) * 0x1p-52;
        }

        // ---- exact rank select: level-1 counter per lane -> 32 words (two per lane) -> bit
        uint32_t g;
        if (few_groups) {
            // at most four level-1 counters (4096 clauses): every lane reads them all and walks the prefix itself
            uint4 c4;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c4.x), "=r"(c4.y), "=r"(c4.z), "=r"(c4.w) : "r"(sL));
            const uint32_t p1 = c4.x, p2 = p1 + c4.y, p3 = p2 + c4.z;
            g = (r >= p1 ? 1u : 0u) + (r >= p2 ? 1u : 0u) + (r >= p3 ? 1u : 0u);
            r -= g == 0 ? 0u : g == 1 ? p1 : g == 2 ? p2 : p3;
            g = active ? g : 0u;  // (an idle chain has r = 0 and may see an empty set)
        } else if (LANES == 16) {
            const uint32_t cntA = lds32(sL + l16 * 4u);
            const uint32_t inclA = sub_incl_scan<LANES>(cntA);
            g = hcount(r >= inclA) & 15u;
            r -= hshfl(inclA - cntA, g);
        } else {
            g = 0;  // (not reached: the four-chain form is only chosen for instances with at most four counters)
        }
        // the lane's PER words of the group, their popcounts, and the scan over the sub-warp
        uint32_t ww[PER];
        if constexpr (LANES == 16) {
            const uint2 t = lds64(sV + (g * 32u + 2u * l16) * 4u);
            ww[0] = t.x, ww[1] = t.y;
        } else {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ww[0]), "=r"(ww[1]), "=r"(ww[2]), "=r"(ww[3]) : "r"(sV + (g * 32u + 4u * l16) * 4u));
        }
        uint32_t pc = 0;
#pragma unroll
        for (uint32_t h = 0; h < PER; h++) pc += __popc(ww[h]);
        const uint32_t inclB = sub_incl_scan<LANES>(pc);
        const uint32_t srcB = hcount(r >= inclB) & (LANES - 1);
        r -= hshfl(inclB - pc, srcB);
        // every lane walks its own words with the rank (meaningful in lane srcB only), then srcB's verdict is broadcast
        uint32_t widx = 0, wmine = ww[0], rr = r;
#pragma unroll
        for (uint32_t h = 0; h + 1 < PER; h++) {
            const uint32_t cnt = __popc(ww[h]);
            const bool next = widx == h && rr >= cnt;
            if (next) rr -= cnt, widx = h + 1, wmine = ww[h + 1];
        }
        const uint32_t wsel = hshfl(wmine, srcB);
        const uint32_t wr = hshfl(widx | rr << 2, srcB);
        r = wr >> 2;
        // the r-th set bit of wsel: a lane tests positions l16 + h * LANES, exactly one test of the sub-warp hits
        uint32_t hits = 0;
#pragma unroll
        for (uint32_t h = 0; h < PER; h++) {
            const bool hit = ((wsel >> (l16 + h * LANES)) & 1u) && __popc(wsel & lt[h]) == r;
            hits |= hballot(hit) << (h * LANES);
        }
        const uint32_t pos = bfind(hits);
        const uint32_t c = active ? (g * 32u + PER * srcB + (wr & 3u)) * 32u + (pos & 31u) : 0u;

        // ---- the clause record (see build_crecx_kernel)
        const uint4* rec = crecx + CRECX_QUADS * c;
        const uint4 q0 = __ldg(rec);
        const uint4 q1 = __ldg(rec + 1);
        const uint4 q2 = __ldg(rec + 2);
        const uint4 q3 = __ldg(rec + 3);
        const uint32_t k = q0.w & 0xffu;
        const double dX = __hiloint2double((int)q2.w, (int)q2.z);
        const double dY = __hiloint2double((int)q3.y, (int)q3.x);
        const double dZ = __hiloint2double((int)q3.w, (int)q3.z);
        // the clause is violated: a variable's value is its literal's negation bit
        const uint32_t a0 = q0.x & 1u, a1 = q0.y & 1u, a2 = q0.z & 1u;

        if (ALGO == ALGO_MOSER) {
            // lib.rs:24-35: one f64 per literal, in literal order; every marked variable flips
            uint32_t marks = 0;
            if (active) {
                if (rng.next_f64() < dX) marks |= 1u;
                if (k > 1 && rng.next_f64() < dY) marks |= 2u;
                if (k > 2 && rng.next_f64() < dZ) marks |= 4u;
            }
            // lib.rs:37-41: toggle the marked variables (distinct on this path); lane j of the half owns literal j
            const uint32_t lj = selp(l16 == 0, q0.x, selp(l16 == 1, q0.y, q0.z));
            if (l16 < 3 && ((marks >> l16) & 1u)) atoms_xor(sA + (lj >> 6) * 4u, 1u << ((lj >> 1) & 31u));
            __syncwarp();  // several variables flip: the re-evaluation below reads the other ones
            flips += __popc(marks);
            uint32_t mk = marks;
            while (__any_sync(FULL, mk != 0)) {
                const uint32_t j = mk ? ffs0(mk) : 0u;
                const uint32_t len = mk ? selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)) : 0u;
                update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), len, selp(j == 0, a0, selp(j == 1, a1, a2)) ^ 1u);
                mk &= mk - 1;
            }
        } else {
            // probSAT, lib.rs:71-72 on the precomputed record (see walk_k3.cuh)
            bool flip = active;
            if (active && q0.w > 0xffu) {
                err = (int)(q0.w >> 8);
                flip = false;
            }
            const double x = u_ps * dX;
            const uint32_t j = (dY <= x ? 1u : 0u) + (dZ <= x ? 1u : 0u);
            const uint32_t lit = selp(j == 0, q0.x, selp(j == 1, q0.y, q0.z));
            {
                // the first lane of the chain's sub-warp alone reads and rewrites the word (one predicated store; an idle
                // chain reads record 0's word and writes nothing); it is read again only after update_var's barrier
                const uint32_t wa = sA + (lit >> 6) * 4u;
                sts32_if(wa, lds32(wa) ^ (1u << ((lit >> 1) & 31u)), flip && l16 == 0);
            }
            update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), flip ? selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)) : 0u,
                       (lit & 1u) ^ 1u);
        }

        if (active && nf < best) save_best();  // lib.rs:292-296
        const bool need = active && rng.low(16);  // a step consumes at most 13 words outside rejection loops
        if (__any_sync(FULL, need)) rng.refill_together(need);
        active = active && nf > 0 && steps < budget && !err;
    }

    if (l16 == 0 && valid) {
        if (err) atomicCAS(A.err, 0, err);
        A.found[run] = nf == 0 && !err;
        A.has_best[run] = best < I.m;
        A.steps[run] = steps;
        A.best_energy[run] = best;
        A.flips[run] = ALGO == ALGO_MOSER ? flips : steps;
    }
}

template <int ALGO, int LANES>
__global__ void __launch_bounds__(128, 8) walk_k3d_batch_kernel(const WalkArgs* __restrict__ items, const uint32_t* __restrict__ cta_begin,
                                                                uint32_t n_items)
{
    extern __shared__ uint32_t sm[];
    __shared__ WalkArgs item;
    const uint32_t it = find_item(cta_begin, n_items, blockIdx.x);
    load_item(item, items + it);
    walk_k3d_body<ALGO, LANES>(item, blockIdx.x - __ldg(cta_begin + it), sm);
}

// lanes = 16: two chains per warp; lanes = 8: four
inline WalkK3BatchKernel walk_k3d_batch_kernel_for(int algo, int lanes)
{
    if (lanes == 8) return algo == ALGO_MOSER ? walk_k3d_batch_kernel<ALGO_MOSER, 8> : walk_k3d_batch_kernel<ALGO_PROBSAT, 8>;
    return algo == ALGO_MOSER ? walk_k3d_batch_kernel<ALGO_MOSER, 16> : walk_k3d_batch_kernel<ALGO_PROBSAT, 16>;
}

}  // namespace slsb
