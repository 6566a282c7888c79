// eval.cuh -- batched clause evaluation: violated bits + energy for B assignments.
//
// Device restatement of clause_is_violated / find_violated_clauses
// (rust/src/lib.rs:13-18, :176-181) and of {LCG,VCG}.get_violated_constraints
// (python/src/sat_representations.py:718-741, :353-373).
#pragma once
#include "common.h"

namespace slsb {

// Bit-sliced evaluation: assignments arrive as one byte per variable (the reference's Vec<bool> /
// int array).  A CTA owns one slice of 32 assignments and a range of clauses.
// tbits is the transposed assignment: tbits[slice][v] holds, in bit b, the value of variable v
// under assignment 32*slice + b, so one literal fetch serves 32 assignments with a single
// bitwise op.  Each lane evaluates one clause for all 32 assignments, then the 32x32 result
// tile is transposed with 32 ballots so that each assignment gets its violated-set word.
__global__ void transpose_assignments_kernel(const uint8_t* __restrict__ bytes, uint32_t* __restrict__ tbits,
                                             uint32_t n, uint64_t batch, uint32_t nslices)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nslices * n) return;
    const uint32_t slice = (uint32_t)(idx / n);
    const uint32_t v = (uint32_t)(idx % n);
    uint32_t word = 0;
    const uint64_t b0 = (uint64_t)slice * 32u;
    const uint32_t cnt = (uint32_t)min((uint64_t)32, batch - b0);
    for (uint32_t i = 0; i < cnt; i++) word |= (uint32_t)(bytes[(b0 + i) * n + v] != 0) << i;
    tbits[idx] = word;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) eval_sliced_kernel(const DevInst I, const uint32_t* __restrict__ tbits,
                                                                 uint32_t* __restrict__ vout,
                                                                 uint32_t* __restrict__ energy, uint64_t batch,
                                                                 uint32_t words_per_block)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t slice = blockIdx.y;
    const uint32_t* a = tbits + (size_t)slice * I.n;
    const uint64_t b0 = (uint64_t)slice * 32u;
    const uint32_t valid = (uint32_t)min((uint64_t)32, batch - b0);
    const uint32_t w_begin = blockIdx.x * words_per_block;
    const uint32_t w_end = min(I.mwords, w_begin + words_per_block);
    uint32_t my_energy = 0;  // lane b accumulates the energy of assignment b0 + b
    for (uint32_t w = w_begin + warp; w < w_end; w += WARPS) {
        const uint32_t c = w * 32u + lane;
        uint32_t viol = 0;  // bit b: clause c violated under assignment b0 + b
        if (c < I.m) {
            uint32_t s, e;
            if (I.uniform_k) {
                s = c * I.uniform_k;
                e = s + I.uniform_k;
            } else {
                s = __ldg(I.row_ptr + c);
                e = __ldg(I.row_ptr + c + 1);
            }
            viol = 0xffffffffu;
            for (uint32_t j = s; j < e && viol; j++) {
                const uint32_t l = __ldg(I.lits + j);
                const uint32_t av = __ldg(a + (l >> 1));
                // literal false  <=>  value == neg bit
                viol &= (l & 1u) ? av : ~av;
            }
        }
        // transpose: word for assignment b = ballot over lanes (clauses) of bit b
        uint32_t mine = 0;
#pragma unroll
        for (int b = 0; b < 32; b++) {
            const uint32_t word = __ballot_sync(0xffffffffu, (viol >> b) & 1u);
            if (lane == b) mine = word;
        }
        if ((uint32_t)lane < valid) {
            if (vout) vout[(b0 + lane) * I.mwords + w] = mine;
            my_energy += __popc(mine);
        }
    }
    if (energy && (uint32_t)lane < valid && my_energy) atomicAdd(energy + b0 + lane, my_energy);
}

// Trajectory rows are written at a fixed pitch (nsteps+1) while a run needs only steps+1 entries
// (rust/src/lib.rs:298-300 pushes one energy per executed step); pack the valid prefixes back to back
// so that only they cross PCIe.  offsets[r] = start of row r in the packed buffer (offsets[R] = total).
__global__ void pack_trajectories_kernel(const uint32_t* __restrict__ traj, uint64_t pitch, const uint64_t* __restrict__ offsets,
                                         uint32_t* __restrict__ packed, uint64_t nruns)
{
    for (uint64_t r = blockIdx.y; r < nruns; r += gridDim.y) {
        const uint64_t beg = offsets[r], len = offsets[r + 1] - beg;
        const uint32_t* src = traj + r * pitch;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x)
            packed[beg + i] = src[i];
    }
}

// Per-step statistics over the runs of one call, on the device: what the reference's evaluation does
// in NumPy on [n_runs, n_steps] arrays (python/src/evaluate_with_given_params.py:18-23, :141-156):
// every trajectory is zero-padded after its last entry (np.pad), then mean and median over runs per step.
// One warp per step; consecutive steps sit in consecutive warps so the strided reads share sectors.
// The median of R non-negative integers is found by bisection on the value with ballot counts.
__global__ void __launch_bounds__(256) trajectory_stats_kernel(const uint32_t* __restrict__ traj, uint64_t pitch,
                                                               const uint64_t* __restrict__ steps, uint32_t nruns, uint64_t nsteps1,
                                                               uint32_t vmax, double* __restrict__ mean, double* __restrict__ median)
{
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nsteps1) return;
    auto value = [&](uint32_t r) -> uint32_t { return (r < nruns && t <= __ldg(steps + r)) ? __ldg(traj + (uint64_t)r * pitch + t) : 0u; };
    // mean
    uint64_t sum = 0;
    for (uint32_t r = lane; r < nruns; r += 32) sum += value(r);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    // k-th smallest (0-based) by bisection: smallest x with count(v <= x) > k
    auto kth = [&](uint32_t k) -> uint32_t {
        uint32_t lo = 0, hi = vmax;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            uint32_t cnt = 0;
            for (uint32_t rb = 0; rb < nruns; rb += 32) {
                const uint32_t r = rb + lane;
                cnt += __popc(__ballot_sync(0xffffffffu, r < nruns && value(r) <= mid));
            }
            if (cnt > k)
                hi = mid;
            else
                lo = mid + 1;
        }
        return lo;
    };
    if (median) {
        const uint32_t a = kth((nruns - 1) / 2), b = (nruns & 1u) ? a : kth(nruns / 2);  // np.median averages the two middle values
        if (lane == 0) median[t] = 0.5 * ((double)a + (double)b);
    }
    if (lane == 0 && mean) mean[t] = (double)sum / (double)nruns;
}

}  // namespace slsb
