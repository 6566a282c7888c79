// walk_generic.cuh -- one warp per chain, any clause shape.
//
// Device restatement of the reference's chain loop (rust/src/lib.rs:201-317):
//   chain init lib.rs:206-223, violated-clause pick :229, greedy branch :231-268,
//   MT resample :20-49, MT-single :77-101, Schoening :51-55, probSAT :57-75,
//   incremental update :278-287, best tracking :292-296, trajectory :298-300.
//
// Chain state (one "slab" per chain, in shared memory or in HBM):
//   abits [nwords]  assignment, 1 bit per variable
//   vbits [mwords]  violated set, 1 bit per canonical clause id
//   l1    [ngroups] number of set bits per 32 vbits words (1024 clauses)
// The violated-clause pick is an exact rank select over vbits in increasing clause id
// (prefix over l1 -> prefix over the 32 words of the group -> bit), so it is uniform over
// the DISTINCT violated clause contents exactly as HashSet::iter().choose() is.
#pragma once
#include "common.h"
#include "philox.cuh"

namespace slsb {

static constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// position of the r-th (0-based) set bit of w; requires r < popc(w)
__device__ __forceinline__ uint32_t select_in_word(uint32_t w, uint32_t r)
{
    uint32_t pos = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const uint32_t c = __popc((w >> pos) & ((1u << s) - 1u));
        if (r >= c) {
            r -= c;
            pos += s;
        }
    }
    return pos;
}

template <bool SMEM>
struct Chain {
    const DevInst& I;
    // Copies of the instance fields the step loop reads.  In the batched kernels the DevInst sits in shared memory
    // next to the chain state, and every shared-memory atomic of the walk may alias it as far as the compiler can
    // tell: through the reference each pointer was re-read from shared memory after every atomic (measured: ~20 % of
    // the step's instructions were those reloads and the generic->shared address conversions around them).
    const uint32_t* __restrict__ row_ptr;
    const uint32_t* __restrict__ lits;
    const uint32_t* __restrict__ occ_ptr;
    const uint32_t* __restrict__ occ;
    const uint32_t* __restrict__ canon;
    const uint32_t uniform_k, mwords;
    uint32_t* abits;
    uint32_t* vbits;
    uint32_t* l1;
    uint32_t* l2;  // nullptr unless the instance has more than 64 level-1 counters
    int lane;
    uint32_t numfalse;

    __device__ __forceinline__ Chain(const DevInst& inst, uint32_t* slab, int lane_)
        : I(inst), row_ptr(inst.row_ptr), lits(inst.lits), occ_ptr(inst.occ_ptr), occ(inst.occ), canon(inst.canon),
          uniform_k(inst.uniform_k), mwords(inst.mwords), abits(slab), vbits(slab + inst.off_v), l1(slab + inst.off_l1),
          l2(inst.l2_pad ? slab + inst.off_l2 : nullptr), lane(lane_), numfalse(0)
    {
    }

    // membership change of canonical clause id: level-1 and (if present) level-2 counters
    __device__ __forceinline__ void count(uint32_t id, uint32_t delta)
    {
        atomicAdd(&l1[id >> 10], delta);
        if (l2) atomicAdd(&l2[id >> 16], delta);
    }

    __device__ __forceinline__ uint32_t abit(uint32_t v) const { return (abits[v >> 5] >> (v & 31)) & 1u; }

    __device__ __forceinline__ void clause_range(uint32_t c, uint32_t& s, uint32_t& e) const
    {
        if (uniform_k) {
            s = c * uniform_k;
            e = s + uniform_k;
        } else {
            s = __ldg(row_ptr + c);
            e = __ldg(row_ptr + c + 1);
        }
    }

    // lib.rs:13-18 with an optional tentatively flipped variable (the greedy branch, lib.rs:239/255)
    __device__ __forceinline__ bool violated(uint32_t c, uint32_t flipped_var) const
    {
        uint32_t s, e;
        clause_range(c, s, e);
        for (uint32_t j = s; j < e; j++) {
            const uint32_t l = __ldg(lits + j);
            const uint32_t v = l >> 1;
            uint32_t a = abit(v);
            if (v == flipped_var) a ^= 1u;
            if (a != (l & 1u)) return false;  // literal is true
        }
        return true;
    }

    // find_violated_clauses (lib.rs:176-181): rebuild vbits / l1 / numfalse from scratch
    __device__ void full_scan()
    {
        numfalse = 0;
        for (uint32_t g = lane; g < I.l1_pad; g += 32) l1[g] = 0;
        __syncwarp();
        if (canon == nullptr) {
            for (uint32_t g = 0; g < I.ngroups; g++) {
                uint32_t acc = 0, mine = 0;
                const uint32_t wend = min(32u, mwords - g * 32u);
                for (uint32_t wi = 0; wi < wend; wi++) {
                    const uint32_t c = (g * 32u + wi) * 32u + lane;
                    const bool viol = c < I.m && violated(c, 0xffffffffu);
                    const uint32_t word = __ballot_sync(FULL, viol);
                    if ((uint32_t)lane == wi) mine = word;
                    acc += __popc(word);
                }
                if ((uint32_t)lane < wend) vbits[g * 32u + lane] = mine;
                if (lane == 0) l1[g] = acc;
                numfalse += acc;
            }
        } else {
            for (uint32_t w = lane; w < mwords; w += 32) vbits[w] = 0;
            __syncwarp();
            for (uint32_t base = 0; base < I.m; base += 32) {
                const uint32_t c = base + lane;
                bool fresh = false;
                if (c < I.m && violated(c, 0xffffffffu)) {
                    const uint32_t id = __ldg(canon + c);
                    const uint32_t bit = 1u << (id & 31);
                    const uint32_t old = atomicOr(&vbits[id >> 5], bit);
                    if (!(old & bit)) {
                        fresh = true;
                        atomicAdd(&l1[id >> 10], 1u);
                    }
                }
                numfalse += __popc(__ballot_sync(FULL, fresh));
            }
        }
        __syncwarp();
        if (l2) {
            for (uint32_t j = lane; j < I.l2_pad; j += 32) {
                uint32_t acc = 0;
                if (j * 64u < I.l1_pad)
                    for (uint32_t i = 0; i < 64; i++) acc += l1[j * 64u + i];
                l2[j] = acc;
            }
            __syncwarp();
        }
    }

    // One round of the counter search: `cnt` holds two adjacent counters per lane (64 per round,
    // zero-padded); finds the counter containing rank r, subtracts the ranks before it from r and
    // returns its index; when `more` rounds follow and r lies beyond this one, returns 64.
    __device__ __forceinline__ uint32_t pick64(const uint32_t* cnt, uint32_t& r, bool more) const
    {
        const uint2 cc = reinterpret_cast<const uint2*>(cnt)[lane];
        const uint32_t both = cc.x + cc.y;
        const uint32_t incl = warp_incl_scan(both, lane);
        if (more) {
            const uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (r >= tot) {
                r -= tot;
                return 64;
            }
        }
        const uint32_t src = __ffs(__ballot_sync(FULL, r < incl)) - 1;
        r -= __shfl_sync(FULL, incl - both, src);
        const uint32_t first = __shfl_sync(FULL, cc.x, src);
        if (r >= first) {
            r -= first;
            return 2 * src + 1;
        }
        return 2 * src;
    }

    // r-th violated clause in increasing canonical clause id; r < numfalse, warp-uniform.
    // Counter levels (each lane owns two adjacent counters, one 8-byte load, a warp scan finds the
    // counter): level 2 (65,536 clauses per counter, only for more than 65,536 clauses) -> level 1
    // (1024 clauses) -> the 32 words of the group (popc + warp scan) -> the bit (ballot).
    __device__ __forceinline__ uint32_t select(uint32_t r) const
    {
        uint32_t gb = 0;
        if (l2) {
            for (uint32_t jb = 0;; jb += 64) {
                const uint32_t j = pick64(l2 + jb, r, jb + 64 < I.l2_pad);
                if (j < 64) {
                    gb = (jb + j) * 64u;
                    break;
                }
            }
        }
        const uint32_t g = gb + pick64(l1 + gb, r, false);
        const uint32_t wi = g * 32u + lane;
        const uint32_t word = wi < mwords ? vbits[wi] : 0u;
        const uint32_t pc = __popc(word);
        const uint32_t incl = warp_incl_scan(pc, lane);
        const uint32_t src = __ffs(__ballot_sync(FULL, r < incl)) - 1;
        r -= __shfl_sync(FULL, incl - pc, src);
        const uint32_t wsel = __shfl_sync(FULL, word, src);
        // lane i holds the answer iff bit i is set and exactly r set bits lie below it
        const bool hit = ((wsel >> lane) & 1u) && __popc(wsel & ((1u << lane) - 1u)) == r;
        return (g * 32u + src) * 32u + (__ffs(__ballot_sync(FULL, hit)) - 1);
    }

    // lib.rs:278-287 for one flipped variable (assignment already holds the new value)
    __device__ __forceinline__ void update_var(uint32_t v)
    {
        const uint32_t os = __ldg(occ_ptr + v), oe = __ldg(occ_ptr + v + 1);
        for (uint32_t base = os; base < oe; base += 32) {
            const uint32_t o = base + lane;
            int delta = 0;
            if (o < oe) {
                const uint32_t cc = __ldg(occ + o);
                const bool viol = violated(cc, 0xffffffffu);
                const uint32_t id = canon ? __ldg(canon + cc) : cc;
                const uint32_t bit = 1u << (id & 31);
                if (viol) {
                    const uint32_t old = atomicOr(&vbits[id >> 5], bit);  // HashSet::insert
                    if (!(old & bit)) {
                        delta = 1;
                        count(id, 1u);
                    }
                } else {
                    const uint32_t old = atomicAnd(&vbits[id >> 5], ~bit);  // HashSet::remove
                    if (old & bit) {
                        delta = -1;
                        count(id, 0xffffffffu);
                    }
                }
            }
            numfalse += __popc(__ballot_sync(FULL, delta > 0));
            numfalse -= __popc(__ballot_sync(FULL, delta < 0));
        }
        __syncwarp();
    }

    // greedy score of flipping v (lib.rs:236-256)
    __device__ __forceinline__ uint64_t greedy_score(uint32_t v, bool mapping) const
    {
        const uint32_t os = __ldg(occ_ptr + v), oe = __ldg(occ_ptr + v + 1);
        uint32_t after = 0, before = 0;
        for (uint32_t base = os; base < oe; base += 32) {
            const uint32_t o = base + lane;
            bool cnt_after = false, cnt_before = false;
            if (o < oe) {
                const uint32_t cc = __ldg(occ + o);
                const bool viol = violated(cc, v);
                if (mapping) {
                    cnt_after = viol;  // every var_to_clauses entry counts (lib.rs:243-247)
                } else {
                    // |find_violated_clauses(after flip)|: distinct contents only (lib.rs:249)
                    const bool rep = (canon == nullptr || __ldg(canon + cc) == cc) &&
                                     !(o > os && __ldg(occ + o - 1) == cc);
                    cnt_after = rep && viol;
                    cnt_before = rep && ((vbits[cc >> 5] >> (cc & 31)) & 1u);
                }
            }
            after += __popc(__ballot_sync(FULL, cnt_after));
            before += __popc(__ballot_sync(FULL, cnt_before));
        }
        return mapping ? (uint64_t)after : (uint64_t)numfalse - before + after;
    }
};

// flip probability of lib.rs:29-30: (1-a)*w + a*(1-w)
__device__ __forceinline__ double flip_prob(uint32_t a, double w) { return a ? (1.0 - w) : w; }

// The flip rules only ever look at a VIOLATED clause, whose literals are all false: the value of a variable is
// its literal's negation bit, so the flip probabilities (lib.rs:29-30) are a static property of the literal
// position (see build_crecx_kernel in walk_k3.cuh for the k<=3 form).  One thread per clause, the same double
// operations in the same order as WeightedIndex::new / UniformFloat::new (lib.rs:71-72):
//   MT rules:  ftab[s+j] = flip probability of literal j
//   probSAT:   ftab[s+j] = sum of the flip probabilities of literals 0..j, cscale[c] = UniformFloat's scale, or
//              -(DERR_WEIGHTS | DERR_ALL_ZERO) when WeightedIndex::new fails on this clause
__device__ __forceinline__ void build_flip_tables_clause(const DevInst& I, const double* __restrict__ w, int algo, double* __restrict__ ftab,
                                                         double* __restrict__ cscale, uint32_t c)
{
    uint32_t s, e;
    if (I.uniform_k) {
        s = c * I.uniform_k;
        e = s + I.uniform_k;
    } else {
        s = I.row_ptr[c];
        e = I.row_ptr[c + 1];
    }
    double run = 0.0;
    bool invalid = false;
    for (uint32_t j = s; j < e; j++) {
        const uint32_t l = I.lits[j];
        const double fp = flip_prob(l & 1u, w[l >> 1]);
        if (algo == ALGO_PROBSAT) {
            invalid |= !(fp >= 0.0);
            run = (j == s) ? fp : run + fp;
            ftab[j] = run;
        } else {
            ftab[j] = fp;
        }
    }
    if (algo == ALGO_PROBSAT) {
        double scale = run;
        if (invalid)
            scale = -(double)DERR_WEIGHTS;
        else if (run == 0.0)
            scale = -(double)DERR_ALL_ZERO;  // also the empty clause; the walk reports that one first
        else
            while (scale * (1.0 - 0x1p-52) >= run) scale = __longlong_as_double(__double_as_longlong(scale) - 1);
        cscale[c] = scale;
    }
}

__global__ void build_flip_tables_kernel(const DevInst I, const double* __restrict__ w, int algo, double* __restrict__ ftab,
                                         double* __restrict__ cscale)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < I.m) build_flip_tables_clause(I, w, algo, ftab, cscale, c);
}

// ALGO is compile-time: every instantiation carries only its own flip rule (fewer registers, shorter loop)
template <bool SMEM, int ALGO>
__device__ __forceinline__ void walk_generic_body(const WalkArgs& A, uint32_t block, uint32_t* smem_state)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t run = (uint64_t)block * A.warps_per_block + warp;
    if (run >= A.nruns) return;
    const DevInst& I = A.inst;
    uint32_t* slab = SMEM ? smem_state + (size_t)warp * I.slab_words : A.gstate + run * (size_t)I.slab_words;
    Chain<SMEM> ch(I, slab, lane);
    const uint64_t chain = A.chain_offset + run;

    if (A.nf_init) {
        // the slab was prepared 32 chains at a time by init_sliced.cuh
        ch.numfalse = A.nf_init[run];
    } else {
    // ---- lib.rs:206-210: assignment[i] = gen_bool(weights_initialize[i]) on the init stream
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
                ch.abits[w] = bits;
            }
            __syncwarp();
        }
        ch.full_scan();  // lib.rs:213-214
    }

    WarpStream rng(A.seed, chain, lane);
    // (loop invariants out of the WalkArgs, which may live in shared memory: see Chain)
    const double* __restrict__ ftab = A.ftab;
    const double* __restrict__ cscale = A.cscale;
    const double pfb = A.pfb;
    const uint64_t nsteps = A.nsteps;
    const bool mapping = A.mapping != 0;

    uint32_t best = I.m;  // lib.rs:203
    bool has_best = false;
    uint32_t* best_bits = A.best_bits + run * (size_t)I.nwords;
    // lib.rs:294 clones the whole assignment on every improvement.  With a large instance in HBM
    // that copy would dominate the walk, so there the chain logs the variables it flips and the best
    // assignment is rebuilt once at the end: current assignment with the flips after the best point undone.
    uint32_t* flog = A.flip_log ? A.flip_log + run * (size_t)A.log_cap : nullptr;
    uint32_t nlog = 0, best_pos = 0;
    auto log_flip = [&](uint32_t v) {  // warp-uniform v
        if (flog) {
            if (lane == 0) flog[nlog] = v;
            nlog++;
        }
    };
    auto save_best = [&]() {
        best = ch.numfalse;
        has_best = true;
        if (flog)
            best_pos = nlog;
        else
            for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = ch.abits[w];
    };
    if (ch.numfalse < best) save_best();  // lib.rs:216-219
    uint32_t* traj = A.want_traj ? A.traj + run * (A.nsteps + 1) : nullptr;
    if (traj && lane == 0) traj[0] = ch.numfalse;  // lib.rs:221-223

    uint64_t steps = 0, flips = 0;
    int err = 0;
    uint32_t mark_mask = 0;  // lane i keeps the MT marks of literals [32 i, 32 i + 32) of the picked clause

    while (ch.numfalse > 0 && steps < nsteps) {  // lib.rs:225
        steps++;
        const uint32_t c = ch.select(rng.uniform_u32(ch.numfalse));  // lib.rs:229
        uint32_t s, e;
        ch.clause_range(c, s, e);
        const uint32_t k = e - s;
        const double ugreedy = rng.next_f64();  // lib.rs:231: drawn on every step

        if (ugreedy < pfb) {
            // ---- lib.rs:231-268 flip the literal with the fewest violations after the flip
            uint32_t best_v = 0;
            uint64_t min_viol = ~0ull;
            for (uint32_t j = 0; j < k; j++) {
                const uint32_t v = __ldg(ch.lits + s + j) >> 1;
                const uint64_t viol = ch.greedy_score(v, mapping);
                if (viol < min_viol) {
                    min_viol = viol;
                    best_v = v;
                }
            }
            if (I.n == 0) {
                err = DERR_EMPTY_CLAUSE;
                break;
            }
            if (lane == 0) ch.abits[best_v >> 5] ^= 1u << (best_v & 31);
            __syncwarp();
            flips++;
            log_flip(best_v);
            ch.update_var(best_v);
        } else if (ALGO == ALGO_MOSER || ALGO == ALGO_MOSER_SINGLE) {
            // ---- lib.rs:24-35 / :81-92: one f64 per literal, in literal order
            if (k > 1024) {
                err = DERR_ARG;
                break;
            }
            const uint64_t p0 = rng.reserve_u64(k);
            uint32_t nmarks = 0;
            for (uint32_t cb = 0; cb < k; cb += 32) {
                const uint32_t j = cb + lane;
                bool mark = false;
                if (j < k) mark = u64_to_f64_53(rng.ck.u64_at(p0 + 2ull * j)) < __ldg(ftab + s + j);
                const uint32_t mask = __ballot_sync(FULL, mark);
                if ((uint32_t)lane == (cb >> 5)) mark_mask = mask;
                nmarks += __popc(mask);
            }
            if (ALGO == ALGO_MOSER) {
                // lib.rs:37-41: toggle every marked variable (a variable marked twice toggles twice)
                for (uint32_t cb = 0; cb < k; cb += 32) {
                    const uint32_t mask = __shfl_sync(FULL, mark_mask, cb >> 5);
                    if ((mask >> lane) & 1u) {
                        const uint32_t v = __ldg(ch.lits + s + cb + lane) >> 1;
                        atomicXor(&ch.abits[v >> 5], 1u << (v & 31));
                    }
                }
                __syncwarp();
                flips += nmarks;
                for (uint32_t cb = 0; cb < k; cb += 32) {
                    uint32_t mask = __shfl_sync(FULL, mark_mask, cb >> 5);
                    while (mask) {
                        const uint32_t j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t v = __ldg(ch.lits + s + cb + j) >> 1;
                        log_flip(v);
                        ch.update_var(v);
                    }
                }
            } else if (nmarks) {
                // lib.rs:94-99: flip one uniformly chosen marked variable
                uint64_t pick = rng.uniform_u64(nmarks);
                uint32_t v = 0;
                for (uint32_t cb = 0; cb < k; cb += 32) {
                    const uint32_t mask = __shfl_sync(FULL, mark_mask, cb >> 5);
                    const uint32_t pc = __popc(mask);
                    if (pick < pc) {
                        v = __ldg(ch.lits + s + cb + select_in_word(mask, (uint32_t)pick)) >> 1;
                        break;
                    }
                    pick -= pc;
                }
                if (lane == 0) ch.abits[v >> 5] ^= 1u << (v & 31);
                __syncwarp();
                flips++;
                log_flip(v);
                ch.update_var(v);
            }
        } else if (ALGO == ALGO_SCHOENING) {
            // ---- lib.rs:51-55
            if (k == 0) {
                err = DERR_EMPTY_CLAUSE;
                break;
            }
            const uint32_t v = __ldg(ch.lits + s + rng.uniform_u32(k)) >> 1;
            if (lane == 0) ch.abits[v >> 5] ^= 1u << (v & 31);
            __syncwarp();
            flips++;
            log_flip(v);
            ch.update_var(v);
        } else {
            // ---- lib.rs:57-75 probSAT: WeightedIndex over the flip probabilities
            if (k == 0) {
                err = DERR_EMPTY_CLAUSE;
                break;
            }
            // WeightedIndex::new's verdict and UniformFloat's scale come precomputed (build_flip_tables_kernel)
            const double scale = __ldg(cscale + c);
            if (!(scale > 0.0)) {
                err = (int)(-scale);
                break;
            }
            // sample from 52 mantissa bits, then partition_point(cum <= x) over the first k-1 running sums
            const double x = ((double)(rng.next_u64() >> 12) * 0x1p-52) * scale;
            uint32_t pick = 0;
            for (uint32_t cb = 0; cb + 1 < k; cb += 32) {
                const uint32_t j = cb + lane;
                pick += __popc(__ballot_sync(FULL, j + 1 < k && __ldg(ftab + s + j) <= x));
            }
            const uint32_t v = __ldg(ch.lits + s + pick) >> 1;
            if (lane == 0) ch.abits[v >> 5] ^= 1u << (v & 31);
            __syncwarp();
            flips++;
            log_flip(v);
            ch.update_var(v);
        }

        if (ch.numfalse < best) save_best();            // lib.rs:292-296
        if (traj && lane == 0) traj[steps] = ch.numfalse;  // lib.rs:298-300
    }

    if (flog && has_best) {
        // best assignment = current assignment with the flips logged after the best point undone
        __syncwarp();
        for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = ch.abits[w];
        __syncwarp();
        for (uint32_t i = best_pos + lane; i < nlog; i += 32) {
            const uint32_t v = flog[i];
            atomicXor(&best_bits[v >> 5], 1u << (v & 31));
        }
    }
    if (lane == 0) {
        if (err) atomicCAS(A.err, 0, err);
        A.found[run] = ch.numfalse == 0 && !err;  // lib.rs:305-307
        A.has_best[run] = has_best;
        A.steps[run] = steps;
        A.best_energy[run] = best;
        A.flips[run] = flips;
    }
}

template <bool SMEM, int ALGO>
__global__ void __launch_bounds__(128, 4) walk_generic_kernel(const WalkArgs A)
{
    extern __shared__ uint32_t smem_state[];
    walk_generic_body<SMEM, ALGO>(A, blockIdx.x, smem_state);
}

// Batched form: many instances in one launch.  cta_begin[i] = first CTA of item i (n_items + 1
// entries); a CTA finds its item by binary search and works from a shared-memory copy of its
// WalkArgs, so thousands of 100-chain solves fill the machine instead of one small grid each.
__device__ __forceinline__ uint32_t find_item(const uint32_t* __restrict__ cta_begin, uint32_t n_items, uint32_t cta)
{
    uint32_t lo = 0, hi = n_items;  // invariant: cta_begin[lo] <= cta < cta_begin[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(cta_begin + mid) <= cta) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void load_item(WalkArgs& dst, const WalkArgs* src)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    for (uint32_t i = threadIdx.x; i < sizeof(WalkArgs) / 4; i += blockDim.x) d[i] = __ldg(s + i);
    __syncthreads();
}

template <bool SMEM, int ALGO>
__global__ void __launch_bounds__(128, 8) walk_generic_batch_kernel(const WalkArgs* __restrict__ items, const uint32_t* __restrict__ cta_begin,
                                                                 uint32_t n_items)
{
    extern __shared__ uint32_t smem_state[];
    __shared__ WalkArgs item;
    const uint32_t it = find_item(cta_begin, n_items, blockIdx.x);
    load_item(item, items + it);
    walk_generic_body<SMEM, ALGO>(item, blockIdx.x - __ldg(cta_begin + it), smem_state);
}

// Host-side choice of the instantiation.
using WalkGenericKernel = void (*)(const WalkArgs);
using WalkGenericBatchKernel = void (*)(const WalkArgs*, const uint32_t*, uint32_t);
template <bool SMEM>
inline WalkGenericKernel walk_generic_kernel_for(int algo)
{
    switch (algo) {
        case ALGO_MOSER: return walk_generic_kernel<SMEM, ALGO_MOSER>;
        case ALGO_MOSER_SINGLE: return walk_generic_kernel<SMEM, ALGO_MOSER_SINGLE>;
        case ALGO_SCHOENING: return walk_generic_kernel<SMEM, ALGO_SCHOENING>;
        default: return walk_generic_kernel<SMEM, ALGO_PROBSAT>;
    }
}
template <bool SMEM>
inline WalkGenericBatchKernel walk_generic_batch_kernel_for(int algo)
{
    switch (algo) {
        case ALGO_MOSER: return walk_generic_batch_kernel<SMEM, ALGO_MOSER>;
        case ALGO_MOSER_SINGLE: return walk_generic_batch_kernel<SMEM, ALGO_MOSER_SINGLE>;
        case ALGO_SCHOENING: return walk_generic_batch_kernel<SMEM, ALGO_SCHOENING>;
        default: return walk_generic_batch_kernel<SMEM, ALGO_PROBSAT>;
    }
}

}  // namespace slsb
