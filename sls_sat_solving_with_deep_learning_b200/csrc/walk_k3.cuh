// walk_k3.cuh -- one warp per chain, instances whose clauses have at most 3 literals
// (random 3-SAT: BASELINE configs 1, 2, 4, 5), no duplicate clause contents, no variable
// twice in a clause.  Same algorithm and the same random draws as walk_generic.cuh
// (reference rust/src/lib.rs:201-317); what changes is the memory layout, chosen so that a
// step has three dependent trips to L2 and everything else stays in registers / shared memory:
//   crec[c]  = {lit0, lit1, lit2, k}                    one 16 B broadcast load per step
//   vrec[v]  = {weight (f64), occ_start, occ_len}       one 16 B broadcast load per literal
//   orec[o]  = {clause, other lit a, other lit b, neg}  one coalesced 16 B load per lane:
//              the clause content travels with the occurrence entry, so re-evaluating the
//              clauses of the flipped variable (lib.rs:278-287) needs no row_ptr / lits gathers.
#pragma once
#include "common.h"
#include "philox.cuh"
#include "walk_generic.cuh"

namespace slsb {

__host__ __device__ inline bool walk_k3_supports(const WalkArgs& a) { return a.inst.crec != nullptr && a.vrec != nullptr; }

template <bool SMEM>
struct ChainK3 : Chain<SMEM> {
    using Base = Chain<SMEM>;
    using Base::abit;
    using Base::I;
    using Base::l1;
    using Base::lane;
    using Base::numfalse;
    using Base::vbits;
    const uint4* vrec;

    __device__ __forceinline__ ChainK3(const DevInst& inst, uint32_t* slab, int lane_, const uint4* vrec_)
        : Base(inst, slab, lane_), vrec(vrec_)
    {
    }

    __device__ __forceinline__ bool lit_false(uint32_t l) const { return l == LIT_NONE || abit(l >> 1) == (l & 1u); }

    // lib.rs:278-287 for the flipped variable v whose occurrence records are [os, os+len)
    __device__ __forceinline__ void update_var(uint32_t v, uint32_t os, uint32_t len)
    {
        const uint32_t av = abit(v);
        for (uint32_t base = 0; base < len; base += 32) {
            const uint32_t o = base + lane;
            int delta = 0;
            if (o < len) {
                const uint4 rec = __ldg(I.orec + os + o);
                const bool viol = (av == rec.w) && lit_false(rec.y) && lit_false(rec.z);
                const uint32_t bit = 1u << (rec.x & 31);
                if (viol) {
                    const uint32_t old = atomicOr(&vbits[rec.x >> 5], bit);
                    if (!(old & bit)) {
                        delta = 1;
                        atomicAdd(&l1[rec.x >> 10], 1u);
                    }
                } else {
                    const uint32_t old = atomicAnd(&vbits[rec.x >> 5], ~bit);
                    if (old & bit) {
                        delta = -1;
                        atomicSub(&l1[rec.x >> 10], 1u);
                    }
                }
            }
            numfalse += __popc(__ballot_sync(FULL, delta > 0));
            numfalse -= __popc(__ballot_sync(FULL, delta < 0));
        }
        __syncwarp();
    }

    // lib.rs:236-256: violations after tentatively flipping v
    __device__ __forceinline__ uint64_t greedy_score(uint32_t v, uint32_t os, uint32_t len, bool mapping) const
    {
        const uint32_t av = abit(v) ^ 1u;
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

__device__ __forceinline__ double vrec_weight(const uint4& r) { return __hiloint2double((int)r.y, (int)r.x); }

template <bool SMEM>
__global__ void __launch_bounds__(128) walk_k3_kernel(const WalkArgs A)
{
    extern __shared__ uint32_t smem_state[];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t run = (uint64_t)blockIdx.x * A.warps_per_block + warp;
    if (run >= A.nruns) return;
    const DevInst& I = A.inst;
    uint32_t* slab = SMEM ? smem_state + (size_t)warp * A.slab_words : A.gstate + run * (size_t)A.slab_words;
    ChainK3<SMEM> ch(I, slab, lane, A.vrec);
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

    WalkStream rng(A.seed, chain);
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

    while (ch.numfalse > 0 && steps < A.nsteps) {
        steps++;
        const uint32_t c = ch.select(rng.uniform_u32(ch.numfalse));  // lib.rs:229
        const uint4 cr = __ldg(I.crec + c);
        const uint32_t k = cr.w;
        const double ugreedy = rng.next_f64();  // lib.rs:231
        // every lane fetches the (up to) three variable records: same address across the warp
        const uint32_t lit[3] = {cr.x, cr.y, cr.z};
        uint4 vr[3];
        double fp[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            vr[j] = make_uint4(0, 0, 0, 0);
            fp[j] = 0.0;
            if ((uint32_t)j < k) {
                vr[j] = __ldg(A.vrec + (lit[j] >> 1));
                fp[j] = flip_prob(ch.abit(lit[j] >> 1), vrec_weight(vr[j]));
            }
        }

        int pick = -1;         // single-flip algorithms: index of the literal to flip
        uint32_t marks = 0;    // MT: bit j set <=> literal j is resampled
        if (ugreedy < A.pfb) {
            uint64_t min_viol = ~0ull;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if ((uint32_t)j < k) {
                    const uint64_t viol = ch.greedy_score(lit[j] >> 1, vr[j].z, vr[j].w, mapping);
                    if (viol < min_viol) {
                        min_viol = viol;
                        pick = j;
                    }
                }
            }
        } else if (A.algo == ALGO_MOSER || A.algo == ALGO_MOSER_SINGLE) {
#pragma unroll
            for (int j = 0; j < 3; j++)
                if ((uint32_t)j < k && rng.next_f64() < fp[j]) marks |= 1u << j;
            if (A.algo == ALGO_MOSER_SINGLE && marks) {
                pick = (int)select_in_word(marks, (uint32_t)rng.uniform_u64(__popc(marks)));
                marks = 0;
            }
        } else if (A.algo == ALGO_SCHOENING) {
            pick = (int)rng.uniform_u32(k);
        } else {
            // WeightedIndex::new over fp[0..k), sampled with UniformFloat (lib.rs:71-72)
            bool invalid = false;
            double total = 0.0, cum0 = 0.0, cum1 = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if ((uint32_t)j < k) {
                    invalid |= !(fp[j] >= 0.0);
                    total = j == 0 ? fp[0] : total + fp[j];
                    if (j == 0) cum0 = total;
                    if (j == 1) cum1 = total;
                }
            }
            if (invalid) {
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
            pick = 0;
            if (k > 1 && cum0 <= x) pick = 1;
            if (k > 2 && cum0 <= x && cum1 <= x) pick = 2;
        }

        if (pick >= 0) {
            const uint32_t v = (pick == 0 ? lit[0] : pick == 1 ? lit[1] : lit[2]) >> 1;
            const uint32_t os = pick == 0 ? vr[0].z : pick == 1 ? vr[1].z : vr[2].z;
            const uint32_t ol = pick == 0 ? vr[0].w : pick == 1 ? vr[1].w : vr[2].w;
            if (lane == 0) ch.abits[v >> 5] ^= 1u << (v & 31);
            __syncwarp();
            flips++;
            ch.update_var(v, os, ol);
        } else if (marks) {
            if (lane < 3 && ((marks >> lane) & 1u)) {
                const uint32_t v = (lane == 0 ? lit[0] : lane == 1 ? lit[1] : lit[2]) >> 1;
                atomicXor(&ch.abits[v >> 5], 1u << (v & 31));
            }
            __syncwarp();
            flips += __popc(marks);
#pragma unroll
            for (int j = 0; j < 3; j++)
                if ((marks >> j) & 1u) ch.update_var(lit[j] >> 1, vr[j].z, vr[j].w);
        }

        if (ch.numfalse < best) save_best();
        if (traj && lane == 0) traj[steps] = ch.numfalse;
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
