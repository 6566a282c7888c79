// walk_k3h.cuh -- the fast walk for k<=3 instances whose chain state does NOT fit in shared memory
// (BASELINE config 5: n = 10^6, m = 4*10^6, 640 KB of state per chain, tens of thousands of chains per GPU).
//
// Same algorithm, same random draws, same records (crecx / orec, static flip probabilities) and the same
// trajectory as walk_k3.cuh; what differs is where the chain state lives and what that implies:
//   * abits / vbits / level-1 / level-2 counters are the chain's slab in HBM, prepared 256 chains at a time by
//     init_sliced.cuh; reads go to L2 (ld.global.cg: the words are modified by reductions performed in L2),
//     membership changes are red.global.xor / red.global.add;
//   * the rank select has three counter levels (64 level-2 counters of 65,536 clauses -> 64 level-1 counters of
//     1024 clauses -> 32 words -> bit), i.e. up to 4,194,304 clauses;
//   * a step is a chain of ~6 dependent round trips to L2 / HBM instead of 2, so throughput comes from the number of
//     chains in flight per SM (registers are the limit), not from the length of the instruction path;
//   * the best assignment is not cloned on every improvement (lib.rs:294; 125 KB per clone here): the chain logs the
//     variables it flips and the best assignment is rebuilt once at the end, as in walk_generic.cuh.
//   * the 64 level-2 counters (256 bytes) live in shared memory next to the chain's random-word ring (512 bytes):
//     one dependent trip to L2 / HBM less per step, and one reduction less per membership change.
#pragma once
#include "common.h"
#include "philox.cuh"
#include "walk_k3.cuh"

namespace slsb {

static constexpr uint32_t K3H_SMEM_WORDS = RING_WORDS + 64;  // per chain: random-word ring + level-2 counters

__host__ __device__ inline bool walk_k3h_supports(const WalkArgs& a)
{
    return a.inst.orec != nullptr && a.crecx != nullptr && a.gstate != nullptr && a.nf_init != nullptr &&
           a.inst.l1_pad <= 64u * 64u && (a.inst.l2_pad == 0 || a.inst.l2_pad == 64);
}

__device__ __forceinline__ uint2 ldcg64(const uint32_t* p)
{
    return __ldcg(reinterpret_cast<const uint2*>(p));
}
__device__ __forceinline__ void red_xor(uint32_t* p, uint32_t v) { asm volatile("red.global.xor.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int ALGO, bool GREEDY, bool TRAJ>
__device__ __forceinline__ void walk_k3h_body(const WalkArgs& A, uint32_t block, uint32_t* sm)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t run = (uint64_t)block * A.warps_per_block + warp;
    if (run >= A.nruns) return;
    const DevInst& I = A.inst;
    uint32_t* const pA = A.gstate + run * (size_t)I.slab_words;  // abits
    uint32_t* const pV = pA + I.off_v;                           // vbits
    uint32_t* const pL1 = pA + I.off_l1;
    const bool has_l2 = I.l2_pad != 0;
    // shared memory of the chain: ring | level-2 counters (byte addresses in the shared window)
    const uint32_t sRing = (uint32_t)__cvta_generic_to_shared(sm) + warp * K3H_SMEM_WORDS * 4u;
    const uint32_t sL2 = sRing + RING_WORDS * 4u;
    if (has_l2) {
        const uint2 c2 = ldcg64(pA + I.off_l2 + 2 * lane);  // prepared by init_counters_kernel
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sL2 + lane * 8u), "r"(c2.x), "r"(c2.y) : "memory");
    }
    __syncwarp();
    const uint64_t chain = A.chain_offset + run;

    // literal code l = 2*var + neg is TRUE iff bit(var) != neg
    auto lit_true = [pA](uint32_t l) -> uint32_t { return (__funnelshift_r(__ldcg(pA + (l >> 6)), 0u, l >> 1) ^ l) & 1u; };

    uint32_t nf = A.nf_init[run];  // the slab was prepared by init_sliced.cuh (lib.rs:206-214)

    RingStream rng(A.seed, chain, sRing, lane);
    uint32_t best = I.m;
    uint32_t* best_bits = A.best_bits + run * (size_t)I.nwords;
    uint32_t* flog = A.flip_log ? A.flip_log + run * (size_t)A.log_cap : nullptr;
    uint32_t nlog = 0, best_pos = 0;
    auto log_flip = [&](uint32_t v) {  // warp-uniform v
        if (flog) {
            if (lane == 0) flog[nlog] = v;
            nlog++;
        }
    };
    auto save_best = [&]() {
        best = nf;
        if (flog)
            best_pos = nlog;
        else
            for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = __ldcg(pA + w);
    };
    if (nf < best) save_best();
    uint32_t* traj = TRAJ ? A.traj + run * (A.nsteps + 1) : nullptr;
    if (TRAJ && lane == 0) traj[0] = nf;

    constexpr bool COUNT_FLIPS = ALGO == ALGO_MOSER || ALGO == ALGO_MOSER_SINGLE;
    uint64_t steps = 0, flips = 0;
    int err = 0;
    const bool mapping = A.mapping != 0;
    const double pfb = A.pfb;
    const uint64_t budget = A.nsteps;
    const uint4* __restrict__ crecx = A.crecx;
    const uint4* __restrict__ orec = I.orec;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // one level of the counter search: the lane owns the two adjacent counters `cc`; returns the index of the
    // counter that holds rank r and removes the ranks in front of it from r
    auto pick64 = [&](const uint2 cc, uint32_t& r) -> uint32_t {
        const uint32_t both = cc.x + cc.y;
        const uint32_t incl = warp_incl_scan_p(both);
        const uint32_t src = __popc(__ballot_sync(FULL, r >= incl));  // lanes below the crossing form a prefix
        r -= __shfl_sync(FULL, incl - both, src);
        const uint32_t first = __shfl_sync(FULL, cc.x, src);
        const bool second = r >= first;
        if (second) r -= first;
        return 2 * src + (second ? 1u : 0u);
    };

    while (nf > 0 && steps < budget) {
        uint32_t lim = (uint32_t)min(budget - steps, (uint64_t)0x40000000u), i = 0;
        while (nf > 0 && i < lim) {
            i++;  // (the ring holds at least 16 words here: a step consumes at most 13 outside rejection loops)

            // ---- lib.rs:229: uniform pick among the violated clauses
            const uint32_t zone1 = nf << __clz(nf);  // rand's UniformInt zone + 1
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
            double u_ps = 0.0;  // probSAT's single draw (lib.rs:72) does not depend on the clause
            if (ALGO == ALGO_PROBSAT && !greedy) u_ps = (double)(rng.next_u64_unchecked() >> 12) * 0x1p-52;

            // ---- exact rank select: level 2 -> level 1 -> words -> bit
            uint32_t gbase = 0;
            if (has_l2) gbase = pick64(lds64(sL2 + lane * 8u), r) * 64u;
            const uint32_t g = gbase + pick64(ldcg64(pL1 + gbase + 2 * lane), r);
            const uint32_t word = __ldcg(pV + g * 32u + lane);  // vbits is zero-padded to whole groups
            const uint32_t pc = __popc(word);
            const uint32_t inclB = warp_incl_scan_p(pc);
            const uint32_t srcB = __popc(__ballot_sync(FULL, r >= inclB));
            r -= __shfl_sync(FULL, inclB - pc, srcB);
            const uint32_t wsel = __shfl_sync(FULL, word, srcB);
            const bool hit = ((wsel >> lane) & 1u) && __popc(wsel & lt_mask) == r;
            const uint32_t c = (g * 32u + srcB) * 32u + bfind(__ballot_sync(FULL, hit));  // exactly one lane hits

            // ---- the clause record (see build_crecx_kernel): literals, occurrence ranges, static flip probabilities
            uint4 q0, q1, q2, q3;  // two 256-bit loads (sm_100): the 64-byte record is two sectors
            {
                const uint4* rec = crecx + CRECX_QUADS * (size_t)c;
                asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w)
                    : "l"(rec));
                asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w), "=r"(q3.x), "=r"(q3.y), "=r"(q3.z), "=r"(q3.w)
                    : "l"(rec + 2));
            }
            const uint32_t k = q0.w & 0xffu;
            const double dX = __hiloint2double((int)q2.w, (int)q2.z);
            const double dY = __hiloint2double((int)q3.y, (int)q3.x);
            const double dZ = __hiloint2double((int)q3.w, (int)q3.z);
            // the clause is violated: a variable's value is its literal's negation bit
            const uint32_t a0 = q0.x & 1u, a1 = q0.y & 1u, a2 = q0.z & 1u;

            // re-evaluation of every clause of a flipped variable (lib.rs:278-287), one occurrence record per lane
            auto update_var = [&](uint32_t os, uint32_t len, uint32_t av, bool single) {
                for (uint32_t base = 0; base < len; base += 32) {
                    const uint32_t o = base + lane;
                    const bool act = o < len;
                    const uint4 orc = __ldg(orec + os + min(o, len - 1u));  // past the end: the last record again, masked
                    uint32_t* const va = pV + (orc.x >> 5);
                    const uint32_t bit = 1u << (orc.x & 31);
                    // Every read here is a random sector of HBM, so only the needed ones are made.  The variable's
                    // literal in this clause either became true -- the clause is satisfied now, and it can only
                    // have been in the violated set before (one read of vbits) -- or became false: then the clause
                    // was satisfied by it before, i.e. not in the set, and it enters the set iff the other two
                    // literals are false (two reads of abits).  1.5 reads per occurrence instead of 3.
                    // (Only when this is the step's single flip: with several flips the set may already have been
                    // updated for this clause by an earlier variable of the same step.)
                    bool rem = false, ins = false;  // HashSet::remove of a present element / insert of a new one
                    if (single) {
                        const bool now_true = act & (av != orc.w), now_false = act & (av == orc.w);
                        if (now_true) rem = (__ldcg(va) & bit) != 0;
                        if (now_false) ins = !(lit_true(orc.y) | lit_true(orc.z));
                    } else {
                        const bool viol = act & (av == orc.w) & !(lit_true(orc.y) | lit_true(orc.z));
                        const bool cur = act & ((__ldcg(va) & bit) != 0);
                        ins = viol && !cur;
                        rem = !viol && cur;
                    }
                    if (ins | rem) {
                        const uint32_t delta = ins ? 1u : 0xffffffffu;
                        red_xor(va, bit);
                        red_add(pL1 + (orc.x >> 10), delta);
                        if (has_l2) atoms_add(sL2 + (orc.x >> 16) * 4u, delta);
                    }
                    nf += __popc(__ballot_sync(FULL, ins)) - __popc(__ballot_sync(FULL, rem));
                }
                __syncwarp();
            };
            auto toggle = [&](uint32_t lit) {  // lane 0 alone reads and rewrites the word (one predicated store)
                uint32_t* const wa = pA + (lit >> 6);
                const uint32_t nw = __ldcg(wa) ^ (1u << ((lit >> 1) & 31u));
                asm volatile("{ .reg .pred p; setp.eq.u32 p, %2, 0; @p st.global.cg.u32 [%0], %1; }" ::"l"(wa), "r"(nw), "r"((uint32_t)lane) : "memory");
                log_flip(lit >> 1);
            };

            if (ALGO == ALGO_MOSER && !(GREEDY && greedy)) {
                // lib.rs:24-35: one f64 per literal, in literal order; every marked variable flips
                uint32_t marks = 0;
                if (rng.next_f64_unchecked() < dX) marks |= 1u;
                if (k > 1 && rng.next_f64_unchecked() < dY) marks |= 2u;
                if (k > 2 && rng.next_f64_unchecked() < dZ) marks |= 4u;
                if (marks) {
                    // lib.rs:37-41: toggle the marked variables (distinct variables, but possibly the same word)
                    if (marks & 1u) toggle(q0.x);
                    if (marks & 2u) {
                        __syncwarp();
                        toggle(q0.y);
                    }
                    if (marks & 4u) {
                        __syncwarp();
                        toggle(q0.z);
                    }
                    __syncwarp();  // several variables flip: the re-evaluation below reads the other ones
                    if (COUNT_FLIPS) flips += __popc(marks);
                    for (uint32_t mk = marks; mk; mk &= mk - 1) {
                        const uint32_t j = ffs0(mk);
                        update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)),
                                   selp(j == 0, a0, selp(j == 1, a1, a2)) ^ 1u, (marks & (marks - 1)) == 0);
                    }
                }
            } else {
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
                                    cb = (__ldcg(pV + (orc.x >> 5)) >> (orc.x & 31)) & 1u;
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
                } else if (ALGO == ALGO_PROBSAT) {
                    // lib.rs:71-72 on the precomputed record (see walk_k3.cuh)
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
                if (ALGO != ALGO_MOSER_SINGLE || j < 3) {
                    const uint32_t lit = selp(j == 0, q0.x, selp(j == 1, q0.y, q0.z));
                    toggle(lit);  // lib.rs:37-41 (the re-evaluation never reads the flipped variable itself)
                    if (COUNT_FLIPS) flips++;
                    update_var(selp(j == 0, q1.x, selp(j == 1, q1.z, q2.x)), selp(j == 0, q1.y, selp(j == 1, q1.w, q2.y)), (lit & 1u) ^ 1u, true);
                }
            }

            if ((nf < best) | rng.low(16)) {
                if (nf < best) save_best();  // lib.rs:292-296
                if (rng.low(16)) rng.refill();
            }
            if (TRAJ && lane == 0) traj[steps + i] = nf;  // lib.rs:298-300
        }
        steps += i;
        if (err) break;
    }

    const bool has_best = best < I.m;  // best starts at m (lib.rs:216) and every improvement is saved
    if (flog && has_best) {
        // best assignment = current assignment with the flips logged after the best point undone
        __syncwarp();
        for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = __ldcg(pA + w);
        __syncwarp();
        for (uint32_t t = best_pos + lane; t < nlog; t += 32) {
            const uint32_t v = flog[t];
            atomicXor(&best_bits[v >> 5], 1u << (v & 31));
        }
    }
    if (lane == 0) {
        if (err) atomicCAS(A.err, 0, err);
        A.found[run] = nf == 0 && !err;
        A.has_best[run] = has_best;
        A.steps[run] = steps;
        A.best_energy[run] = best;
        A.flips[run] = COUNT_FLIPS ? flips : steps;
    }
}

template <int ALGO, bool GREEDY, bool TRAJ>
__global__ void __launch_bounds__(128, 8) walk_k3h_kernel(const WalkArgs A)
{
    extern __shared__ uint32_t sm[];
    walk_k3h_body<ALGO, GREEDY, TRAJ>(A, blockIdx.x, sm);
}

template <bool GREEDY, bool TRAJ>
inline WalkK3Kernel walk_k3h_kernel_for_algo(int algo)
{
    switch (algo) {
        case ALGO_MOSER: return walk_k3h_kernel<ALGO_MOSER, GREEDY, TRAJ>;
        case ALGO_MOSER_SINGLE: return walk_k3h_kernel<ALGO_MOSER_SINGLE, GREEDY, TRAJ>;
        case ALGO_SCHOENING: return walk_k3h_kernel<ALGO_SCHOENING, GREEDY, TRAJ>;
        default: return walk_k3h_kernel<ALGO_PROBSAT, GREEDY, TRAJ>;
    }
}
inline WalkK3Kernel walk_k3h_kernel_for(int algo, double pfb, bool traj)
{
    if (pfb > 0.0) return traj ? walk_k3h_kernel_for_algo<true, true>(algo) : walk_k3h_kernel_for_algo<true, false>(algo);
    return traj ? walk_k3h_kernel_for_algo<false, true>(algo) : walk_k3h_kernel_for_algo<false, false>(algo);
}

}  // namespace slsb
