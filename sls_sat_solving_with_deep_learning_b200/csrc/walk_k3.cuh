// walk_k3.cuh -- one warp per chain, instances whose clauses have at most 3 literals
// (random 3-SAT: BASELINE configs 1, 2, 4, 5), no duplicate clause contents, no variable
// twice in a clause.  Same algorithm and the same random draws as walk_generic.cuh
// (reference rust/src/lib.rs:201-317); what changes is the memory layout, chosen so that one
// step makes two dependent trips to L2 and everything else stays in registers / shared memory:
//   crecx[c] = 64 bytes: {lit0, lit1, lit2, k | occ_start/occ_len of the three variables |
//              the three resample weights (f64)} -- everything the flip rule needs, one
//              broadcast load per step (built per call from the weights by build_crecx_kernel)
//   orec[o]  = {clause, other lit a, other lit b, neg} one coalesced 16-byte load per lane:
//              the clause content travels with the occurrence entry, so re-evaluating the
//              clauses of the flipped variable (lib.rs:278-287) needs no row_ptr / lits gathers.
#pragma once
#include "common.h"
#include "philox.cuh"
#include "walk_generic.cuh"

namespace slsb {

__host__ __device__ inline bool walk_k3_supports(const WalkArgs& a) { return a.inst.orec != nullptr && a.crecx != nullptr; }

// crecx[c] = { {l0,l1,l2,k}, {os0,ol0,os1,ol1}, {os2,ol2,w0.lo,w0.hi}, {w1.lo,w1.hi,w2.lo,w2.hi} }
__global__ void build_crecx_kernel(const uint4* __restrict__ crec, const uint32_t* __restrict__ occ_ptr,
                                   const double* __restrict__ w, uint4* __restrict__ crecx, uint32_t m)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint4 cr = crec[c];
    const uint32_t lit[3] = {cr.x, cr.y, cr.z};
    uint32_t os[3], ol[3];
    uint2 wb[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        os[j] = ol[j] = 0;
        wb[j] = make_uint2(0, 0);
        if (lit[j] != LIT_NONE) {
            const uint32_t v = lit[j] >> 1;
            os[j] = occ_ptr[v];
            ol[j] = occ_ptr[v + 1] - os[j];
            const double wv = w[v];
            wb[j] = make_uint2((uint32_t)__double2loint(wv), (uint32_t)__double2hiint(wv));
        }
    }
    uint4* out = crecx + 4 * (size_t)c;
    out[0] = cr;
    out[1] = make_uint4(os[0], ol[0], os[1], ol[1]);
    out[2] = make_uint4(os[2], ol[2], wb[0].x, wb[0].y);
    out[3] = make_uint4(wb[1].x, wb[1].y, wb[2].x, wb[2].y);
}

template <bool SMEM>
struct ChainK3 : Chain<SMEM> {
    using Base = Chain<SMEM>;
    using Base::abit;
    using Base::abits;
    using Base::I;
    using Base::l1;
    using Base::lane;
    using Base::numfalse;
    using Base::vbits;

    __device__ __forceinline__ ChainK3(const DevInst& inst, uint32_t* slab, int lane_) : Base(inst, slab, lane_) {}

    // literal l is false under the current assignment (an absent literal counts as false)
    __device__ __forceinline__ bool lit_false(uint32_t l) const
    {
        const uint32_t v = l == LIT_NONE ? 0u : l >> 1;
        return l == LIT_NONE || abit(v) == (l & 1u);
    }

    // lib.rs:278-287 for the flipped variable whose occurrence records are [os, os+len) and whose
    // NEW value is av.  No duplicates on this path, so membership changes are detected with a plain
    // read and only actual changes touch the bitmap / counters.
    __device__ __forceinline__ void update_var(uint32_t av, uint32_t os, uint32_t len)
    {
        for (uint32_t base = 0; base < len; base += 32) {
            const uint32_t o = base + lane;
            bool ins = false, rem = false;
            if (o < len) {
                const uint4 rec = __ldg(I.orec + os + o);
                const bool viol = (av == rec.w) && lit_false(rec.y) && lit_false(rec.z);
                const uint32_t bit = 1u << (rec.x & 31);
                const bool cur = (vbits[rec.x >> 5] & bit) != 0;
                ins = viol && !cur;  // HashSet::insert of a new element
                rem = !viol && cur;  // HashSet::remove of a present element
                if (ins != rem) {
                    atomicXor(&vbits[rec.x >> 5], bit);
                    atomicAdd(&l1[rec.x >> 10], ins ? 1u : 0xffffffffu);
                }
            }
            numfalse += __popc(__ballot_sync(FULL, ins));
            numfalse -= __popc(__ballot_sync(FULL, rem));
        }
        __syncwarp();
    }

    // lib.rs:236-256: violations after tentatively flipping the variable (current value a)
    __device__ __forceinline__ uint64_t greedy_score(uint32_t a, uint32_t os, uint32_t len, bool mapping) const
    {
        const uint32_t av = a ^ 1u;
        uint32_t after = 0, before = 0;
        for (uint32_t base = 0; base < len; base += 32) {
            const uint32_t o = base + lane;
            bool ca = false, cb = false;
            if (o < len) {
                const uint4 rec = __ldg(I.orec + os + o);
                ca = (av == rec.w) && lit_false(rec.y) && lit_false(rec.z);
                cb = (vbits[rec.x >> 5] >> (rec.x & 31)) & 1u;
            }
            after += __popc(__ballot_sync(FULL, ca));
            before += __popc(__ballot_sync(FULL, cb));
        }
        return mapping ? (uint64_t)after : (uint64_t)numfalse - before + after;
    }
};

template <bool SMEM>
__global__ void __launch_bounds__(128, 8) walk_k3_kernel(const WalkArgs A)
{
    extern __shared__ uint32_t smem_state[];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t run = (uint64_t)blockIdx.x * A.warps_per_block + warp;
    if (run >= A.nruns) return;
    const DevInst& I = A.inst;
    uint32_t* slab = SMEM ? smem_state + (size_t)warp * I.slab_words : A.gstate + run * (size_t)I.slab_words;
    ChainK3<SMEM> ch(I, slab, lane);
    const uint64_t chain = A.chain_offset + run;

    {  // lib.rs:206-210
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

    WarpStream rng(A.seed, chain, lane);
    ch.full_scan();  // lib.rs:213-214

    uint32_t best = I.m;
    bool has_best = false;
    uint32_t* best_bits = A.best_bits + run * (size_t)I.nwords;
    auto save_best = [&]() {
        best = ch.numfalse;
        has_best = true;
        for (uint32_t w = lane; w < I.nwords; w += 32) best_bits[w] = ch.abits[w];
    };
    if (ch.numfalse < best) save_best();
    uint32_t* traj = A.want_traj ? A.traj + run * (A.nsteps + 1) : nullptr;
    if (traj && lane == 0) traj[0] = ch.numfalse;

    uint64_t steps = 0, flips = 0;
    int err = 0;
    const bool mapping = A.mapping != 0;
    const int algo = A.algo;
    const bool greedy_possible = A.pfb > 0.0;
    const uint4* crecx = A.crecx;

    while (ch.numfalse > 0 && steps < A.nsteps) {
        steps++;
        const uint32_t c = ch.select(rng.uniform_u32(ch.numfalse));  // lib.rs:229
        // one 64-byte record, same address in every lane
        const uint4 q0 = __ldg(crecx + 4 * (size_t)c);
        const uint4 q1 = __ldg(crecx + 4 * (size_t)c + 1);
        const uint4 q2 = __ldg(crecx + 4 * (size_t)c + 2);
        const uint4 q3 = __ldg(crecx + 4 * (size_t)c + 3);
        const uint32_t k = q0.w;
        // lib.rs:231: one f64 is drawn on every step, greedy or not
        bool greedy = false;
        if (greedy_possible)
            greedy = rng.next_f64() < A.pfb;
        else
            rng.skip_u64();

        // current values of the clause's variables and their flip probabilities (lib.rs:29-30)
        const uint32_t a0 = ch.abit(q0.x >> 1);
        const uint32_t a1 = k > 1 ? ch.abit(q0.y >> 1) : 0u;
        const uint32_t a2 = k > 2 ? ch.abit(q0.z >> 1) : 0u;

        int pick = -1;       // single-flip rules: index of the literal whose variable flips
        uint32_t marks = 0;  // MT: bit j <=> literal j is resampled
        if (greedy) {
            uint64_t min_viol = ~0ull;
            const uint64_t s0 = ch.greedy_score(a0, q1.x, q1.y, mapping);
            if (s0 < min_viol) {
                min_viol = s0;
                pick = 0;
            }
            if (k > 1) {
                const uint64_t s1 = ch.greedy_score(a1, q1.z, q1.w, mapping);
                if (s1 < min_viol) {
                    min_viol = s1;
                    pick = 1;
                }
            }
            if (k > 2) {
                const uint64_t s2 = ch.greedy_score(a2, q2.x, q2.y, mapping);
                if (s2 < min_viol) {
                    min_viol = s2;
                    pick = 2;
                }
            }
        } else if (algo == ALGO_SCHOENING) {
            pick = (int)rng.uniform_u32(k);
        } else {
            const double fp0 = flip_prob(a0, __hiloint2double((int)q2.w, (int)q2.z));
            const double fp1 = k > 1 ? flip_prob(a1, __hiloint2double((int)q3.y, (int)q3.x)) : 0.0;
            const double fp2 = k > 2 ? flip_prob(a2, __hiloint2double((int)q3.w, (int)q3.z)) : 0.0;
            if (algo == ALGO_PROBSAT) {
                // WeightedIndex::new over fp[0..k), sampled with UniformFloat (lib.rs:71-72)
                const double cum0 = fp0;
                const double cum1 = k > 1 ? cum0 + fp1 : cum0;
                const double total = k > 2 ? cum1 + fp2 : cum1;
                if (!(fp0 >= 0.0) || !(fp1 >= 0.0) || !(fp2 >= 0.0)) {
                    err = DERR_WEIGHTS;
                    break;
                }
                if (total == 0.0) {
                    err = DERR_ALL_ZERO;
                    break;
                }
                double scale = total;
                while (scale * (1.0 - 0x1p-52) >= total) scale = __longlong_as_double(__double_as_longlong(scale) - 1);
                const double x = ((double)(rng.next_u64() >> 12) * 0x1p-52) * scale;
                pick = (k > 1 && cum0 <= x) ? ((k > 2 && cum1 <= x) ? 2 : 1) : 0;
            } else {
                // lib.rs:24-35 / :81-92: one f64 per literal, in literal order
                if (rng.next_f64() < fp0) marks |= 1u;
                if (k > 1 && rng.next_f64() < fp1) marks |= 2u;
                if (k > 2 && rng.next_f64() < fp2) marks |= 4u;
                if (algo == ALGO_MOSER_SINGLE && marks) {  // lib.rs:94-99
                    pick = (int)select_in_word(marks, (uint32_t)rng.uniform_u64(__popc(marks)));
                    marks = 0;
                }
            }
        }

        if (pick >= 0) marks = 1u << pick;
        if (marks) {
            // toggle the marked variables (distinct on this path), then re-evaluate their clauses
            if (lane < 3 && ((marks >> lane) & 1u)) {
                const uint32_t v = (lane == 0 ? q0.x : lane == 1 ? q0.y : q0.z) >> 1;
                atomicXor(&ch.abits[v >> 5], 1u << (v & 31));
            }
            __syncwarp();
            flips += __popc(marks);
            if (marks & 1u) ch.update_var(a0 ^ 1u, q1.x, q1.y);
            if (marks & 2u) ch.update_var(a1 ^ 1u, q1.z, q1.w);
            if (marks & 4u) ch.update_var(a2 ^ 1u, q2.x, q2.y);
        }

        if (ch.numfalse < best) save_best();               // lib.rs:292-296
        if (traj && lane == 0) traj[steps] = ch.numfalse;  // lib.rs:298-300
    }

    if (lane == 0) {
        if (err) atomicCAS(A.err, 0, err);
        A.found[run] = ch.numfalse == 0 && !err;
        A.has_best[run] = has_best;
        A.steps[run] = steps;
        A.best_energy[run] = best;
        A.flips[run] = flips;
    }
}

}  // namespace slsb
