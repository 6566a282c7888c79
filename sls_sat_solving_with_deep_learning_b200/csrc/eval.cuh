// eval.cuh -- batched clause evaluation: violated bits + energy for B assignments.
//
// Device restatement of clause_is_violated / find_violated_clauses
// (rust/src/lib.rs:13-18, :176-181) and of {LCG,VCG}.get_violated_constraints
// (python/src/sat_representations.py:718-741, :353-373).
#pragma once
#include "common.h"

namespace slsb {

// ---------------------------------------------------------------------------------------------
// Bit-sliced evaluation, 256 assignments per literal fetch.
//
// A "slice" is a group of SLICE_W = 256 assignments (or chains, for the HBM-tier chain init of
// init_sliced.cuh).  tq[slice][v] holds the 256 values of variable v as 8 words = 32 bytes = exactly one
// memory sector, so the random gather a literal costs brings in nothing but useful bits (with the
// former 32-wide slices 7/8 of every gathered sector was discarded and the kernel ran at the L2's sector rate,
// not at its byte rate).  A lane owns one clause and evaluates it for the whole slice with bitwise ops;
// the 32 clauses of a warp form 32x32 bit tiles that are transposed in registers (5 shuffle rounds per tile
// instead of 32 ballots) so that each assignment gets its violated-set word; the words of 8 consecutive
// clause words are collected in registers and leave as one whole 32-byte sector per assignment (a 256-bit store).
static constexpr uint32_t SLICE_W = 256, SLICE_Q = SLICE_W / 32, TILE_WORDS = 8;

// Transposes of the eight 32x32 bit tiles a row of clauses holds (x[q], one tile row per lane): on return lane b
// holds column b of every tile (bit i = row i).  The eight tiles advance round by round, so eight independent
// shuffles are in flight at a time; transposing them one after the other exposes the full shuffle latency 40
// times per row (measured: 19 % issue utilisation, every sample on the instruction after a SHFL).
__device__ __forceinline__ void transpose32x8(uint32_t (&x)[SLICE_Q], int lane)
{
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        // m: the bit positions whose index has bit j clear
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const bool hi = (lane & j) != 0;
        // this lane's diagonal block stays.  The mask is made opaque to the compiler: left visible, `hi ? ~m : m` is folded
        // into every use as (immediate ^ lane mask) and the merge below becomes three LOP3 per word instead of one
        // (ncu, round 2: 1268 LOP3 per tile of 256 x 256 bits, 40 % of the kernel's instructions)
        uint32_t keep;
        asm("xor.b32 %0, %1, %2;" : "=r"(keep) : "r"(m), "r"(hi ? 0xffffffffu : 0u));
        // the partner's off-diagonal block moves across by a rotation (the bits that wrap around land in positions
        // `keep` masks out): one funnel shift instead of shift-left / shift-right / select
        const uint32_t rot = hi ? 32u - (uint32_t)j : (uint32_t)j;
        uint32_t y[SLICE_Q];
        // volatile: the eight shuffles of a round are issued back to back and wait together (left to itself the
        // compiler serialises tile pairs to save registers and every pair pays the shuffle latency)
#pragma unroll
        for (uint32_t q = 0; q < SLICE_Q; q++)
            asm volatile("shfl.sync.bfly.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(y[q]) : "r"(x[q]), "r"(j));
#pragma unroll
        for (uint32_t q = 0; q < SLICE_Q; q++) {
            const uint32_t moved = __funnelshift_l(y[q], y[q], rot);
            asm("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(x[q]) : "r"(moved), "r"(keep));  // keep ? x : moved, bit by bit
        }
    }
}

// The gathered sector of tq: asked to stay in L2 (every clause of the slice comes back to this array at random),
// while the literal stream and the outputs pass through with evict-first hints.
// (sm_100 has 256-bit global loads: the whole sector is one instruction.)
struct Sector {
    uint32_t w[SLICE_Q];
};
__device__ __forceinline__ Sector ldg_sector_keep(const uint32_t* p)
{
    Sector v;
    asm("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
        : "l"(p));
    return v;
}

// one literal's contribution: literal false <=> value == negation bit
__device__ __forceinline__ void and_literal(uint32_t (&viol)[SLICE_Q], uint32_t l, const Sector& a)
{
    const uint32_t flip = (l & 1u) ? 0u : 0xffffffffu;  // positive literal: false where the value is 0
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) viol[q] &= a.w[q] ^ flip;
}

// Clause rows of a tile, software-pipelined.  A lane owns clause 32*(w0 + wi) + lane in row wi.  While the caller
// transposes row wi, the four sector gathers of row wi+1 and the literal words of row wi+2 are in flight: the
// scan is bound by memory latency (DRAM for the literal stream, L2 for the gathers), not by any throughput.
// The first four literals of a clause travel through the pipeline; longer clauses finish with a plain loop.
template <int NL = 4>
struct ClauseRowsT {
    const DevInst& I;
    const uint32_t* __restrict__ tq;
    uint32_t c_next;             // clause whose literals are loaded next
    uint32_t s_cur, e_cur, s_nxt, e_nxt;
    bool ok_cur, ok_nxt;         // clause index < m
    uint32_t l_cur[NL], l_nxt[NL];
    Sector a[NL];

    __device__ __forceinline__ void fetch_lits(uint32_t c, uint32_t& s, uint32_t& e, bool& ok, uint32_t (&l)[NL])
    {
        ok = c < I.m;
        s = e = 0;
        if (ok) {
            if (I.uniform_k) {
                s = c * I.uniform_k;
                e = s + I.uniform_k;
            } else {
                s = __ldcs(I.row_ptr + c);
                e = __ldcs(I.row_ptr + c + 1);
            }
        }
#pragma unroll
        for (uint32_t t = 0; t < NL; t++) l[t] = e > s ? __ldcs(I.lits + min(s + t, e - 1u)) : 0u;  // tail: last literal again
    }
    __device__ __forceinline__ void gather(const uint32_t (&l)[NL])
    {
#pragma unroll
        for (uint32_t t = 0; t < NL; t++) a[t] = ldg_sector_keep(tq + (size_t)(l[t] >> 1) * SLICE_Q);
    }
    // c0 = this lane's clause in row 0
    __device__ __forceinline__ ClauseRowsT(const DevInst& I_, const uint32_t* tq_, uint32_t c0) : I(I_), tq(tq_)
    {
        fetch_lits(c0, s_cur, e_cur, ok_cur, l_cur);
        gather(l_cur);
        fetch_lits(c0 + 32u, s_nxt, e_nxt, ok_nxt, l_nxt);
        c_next = c0 + 64u;
    }
    // viol[q] bit b <=> this lane's clause of the current row is violated under assignment 32q + b; then advance
    __device__ __forceinline__ void next(uint32_t (&viol)[SLICE_Q])
    {
        const uint32_t init = ok_cur ? 0xffffffffu : 0u;  // an empty clause is violated everywhere (lib.rs:13-18)
#pragma unroll
        for (uint32_t q = 0; q < SLICE_Q; q++) viol[q] = init;
        if (e_cur > s_cur) {
#pragma unroll
            for (uint32_t t = 0; t < NL; t++) and_literal(viol, l_cur[t], a[t]);  // AND is idempotent: repeats are harmless
            for (uint32_t j = s_cur + NL; j < e_cur; j++) {
                const uint32_t l = __ldcs(I.lits + j);
                and_literal(viol, l, ldg_sector_keep(tq + (size_t)(l >> 1) * SLICE_Q));
            }
        }
        s_cur = s_nxt;
        e_cur = e_nxt;
        ok_cur = ok_nxt;
#pragma unroll
        for (uint32_t t = 0; t < NL; t++) l_cur[t] = l_nxt[t];
        gather(l_cur);
        fetch_lits(c_next, s_nxt, e_nxt, ok_nxt, l_nxt);
        c_next += 32u;
    }
};

using ClauseRows = ClauseRowsT<4>;

// bytes[b][v] (the reference's int / bool arrays, one byte per variable) -> tq[slice][v][q]
__global__ void transpose_assignments_kernel(const uint8_t* __restrict__ bytes, uint32_t* __restrict__ tq, uint32_t n, uint64_t batch,
                                             uint32_t nslices)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)nslices * n) return;
    const uint32_t slice = (uint32_t)(idx / n);
    const uint32_t v = (uint32_t)(idx % n);
    const uint64_t b0 = (uint64_t)slice * SLICE_W;
    uint32_t w[SLICE_Q];
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) {
        uint32_t word = 0;
        const uint64_t r0 = b0 + 32u * q;
        if (r0 < batch) {
            const uint32_t cnt = (uint32_t)min((uint64_t)32, batch - r0);
            for (uint32_t i = 0; i < cnt; i++) word |= (uint32_t)(bytes[(r0 + i) * n + v] != 0) << i;
        }
        w[q] = word;
    }
    uint4* out = reinterpret_cast<uint4*>(tq + idx * SLICE_Q);
    out[0] = make_uint4(w[0], w[1], w[2], w[3]);
    out[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// grid: (tile chunks, slices); a warp works through tiles of TILE_WORDS clause words x 256 assignments.
// vout rows have a pitch of `pitch` words (a multiple of 8, so that a tile row is one aligned sector).
// OUT = false: energies only (no staging buffer, twice the resident warps).
// (168 registers, three CTAs per SM.  Round 2 measured four and five CTAs per SM through launch bounds: 128 / 96 registers with
// 128 / 304 bytes of spills, 0.50 / 0.55 ms against 0.46 ms.)
// NL: literals fetched and gathered ahead per clause (3 for uniform k <= 3: no redundant fourth gather; else 4 + a tail loop)
template <int WARPS, bool OUT, int NL = 4>
__global__ void __launch_bounds__(WARPS * 32) eval_sliced_kernel(const DevInst I, const uint32_t* __restrict__ tq_all,
                                                                 uint32_t* __restrict__ vout, uint32_t pitch,
                                                                 uint32_t* __restrict__ energy, uint64_t batch, uint32_t tiles_per_block)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t slice = blockIdx.y;
    const uint32_t* tq = tq_all + (size_t)slice * I.n * SLICE_Q;
    const uint64_t b0 = (uint64_t)slice * SLICE_W;
    const uint32_t valid = (uint32_t)min((uint64_t)SLICE_W, batch - b0);
    const uint32_t ntiles = (I.mwords + TILE_WORDS - 1) / TILE_WORDS;
    const uint32_t t_begin = blockIdx.x * tiles_per_block;
    const uint32_t t_end = min(ntiles, t_begin + tiles_per_block);
    uint32_t my_energy[SLICE_Q];  // lane b, entry q: energy of assignment b0 + 32q + b
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) my_energy[q] = 0;
    // a warp owns a contiguous run of tiles, so the row pipeline runs through from its first to its last row
    const uint32_t per_warp = tiles_per_block / WARPS;  // tiles_per_block is a multiple of WARPS
    const uint32_t t0 = min(t_end, t_begin + warp * per_warp), t1 = min(t_end, t0 + per_warp);
    ClauseRowsT<NL> rows(I, tq, t0 * TILE_WORDS * 32u + lane);
    for (uint32_t t = t0; t < t1; t++) {
        // After the transposition lane b holds, for each of its eight assignments 32q + b, the violated-set word of the row.
        // The eight rows of a tile are kept in registers (acc[q][wi]) and leave as ONE 32-byte sector per assignment: a
        // 256-bit store per lane and q, no staging in shared memory (round 1 staged the tile: 64 shared-memory stores, 64
        // loads and 64 four-byte global stores per tile next to the 320 shuffles of the transpositions -- ncu: the kernel
        // was bound by the shared-memory / shuffle instruction queue, mio_throttle + short_scoreboard).
        uint32_t acc[OUT ? SLICE_Q : 1][OUT ? TILE_WORDS : 1];
#pragma unroll
        for (uint32_t wi = 0; wi < TILE_WORDS; wi++) {
            uint32_t viol[SLICE_Q];
            rows.next(viol);
            transpose32x8(viol, lane);
#pragma unroll
            for (uint32_t q = 0; q < SLICE_Q; q++) {
                my_energy[q] += __popc(viol[q]);
                if (OUT) acc[q][wi] = viol[q];
            }
        }
        if (OUT) {
            const uint32_t w = t * TILE_WORDS;  // pitch is a multiple of TILE_WORDS: the sector lies inside the row
#pragma unroll
            for (uint32_t q = 0; q < SLICE_Q; q++) {
                const uint32_t a = 32u * q + lane;
                if (a < valid)  // written once, never re-read here: evict-first
                    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(vout + (b0 + a) * (size_t)pitch + w), "r"(acc[q][0]),
                                 "r"(acc[q][1]), "r"(acc[q][2]), "r"(acc[q][3]), "r"(acc[q][4]), "r"(acc[q][5]), "r"(acc[q][6]), "r"(acc[q][7])
                                 : "memory");
            }
        }
    }
    if (energy) {
#pragma unroll
        for (uint32_t q = 0; q < SLICE_Q; q++) {
            const uint32_t a = 32u * q + lane;
            if (a < valid && my_energy[q]) atomicAdd(energy + b0 + a, my_energy[q]);
        }
    }
}

// Energies only (candidate energies, data_utils.py:123-130; sls_eval_assignments without the violated bits): no
// transposition per row.  A lane keeps, for its own clauses, a bit-sliced ("vertical") counter per assignment: six
// planes of 8 words, plane p holding bit p of the 256 counts.  Rows enter through carry-save adders, four rows per
// round (Harley-Seal: 14 three-input logic ops per word per four rows); after at most 60 rows the planes are
// transposed once (lane b then holds the plane bits of all 32 lanes for its assignments) and folded into the
// energies with popcounts.  What remains per row is the literal stream, the sector gathers and ~60 logic
// instructions: the kernel runs at the L2's random-sector rate (tools/ubench.cu measures that ceiling).
__device__ __forceinline__ void csa(uint32_t& sum_io, uint32_t& carry, uint32_t a, uint32_t b)
{
    const uint32_t s = sum_io;
    carry = (a & b) | (s & (a ^ b));  // one LOP3 (majority)
    sum_io = s ^ a ^ b;               // one LOP3
}

template <int WARPS, int NL>
__global__ void __launch_bounds__(WARPS * 32) eval_energy_sliced_kernel(const DevInst I, const uint32_t* __restrict__ tq_all,
                                                                        uint32_t* __restrict__ energy, uint64_t batch, uint32_t tiles_per_block)
{
    constexpr int NP = 6;               // planes: counts up to 63
    constexpr uint32_t FLUSH_TILES = 7; // 56 rows between flushes
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t slice = blockIdx.y;
    const uint32_t* tq = tq_all + (size_t)slice * I.n * SLICE_Q;
    const uint64_t b0 = (uint64_t)slice * SLICE_W;
    const uint32_t valid = (uint32_t)min((uint64_t)SLICE_W, batch - b0);
    const uint32_t ntiles = (I.mwords + TILE_WORDS - 1) / TILE_WORDS;
    const uint32_t t_begin = blockIdx.x * tiles_per_block;
    const uint32_t t_end = min(ntiles, t_begin + tiles_per_block);
    const uint32_t per_warp = tiles_per_block / WARPS;  // tiles_per_block is a multiple of WARPS
    const uint32_t t0 = min(t_end, t_begin + warp * per_warp), t1 = min(t_end, t0 + per_warp);
    uint32_t my_energy[SLICE_Q];  // lane b, entry q: energy of assignment b0 + 32q + b
    uint32_t P[NP][SLICE_Q];
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) {
        my_energy[q] = 0;
#pragma unroll
        for (int p = 0; p < NP; p++) P[p][q] = 0;
    }
    auto flush = [&]() {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            transpose32x8(P[p], lane);
#pragma unroll
            for (uint32_t q = 0; q < SLICE_Q; q++) {
                my_energy[q] += (uint32_t)__popc(P[p][q]) << p;
                P[p][q] = 0;
            }
        }
    };
    ClauseRowsT<NL> rows(I, tq, t0 * TILE_WORDS * 32u + lane);
    uint32_t since_flush = 0;
    for (uint32_t t = t0; t < t1; t++) {
        uint32_t A[SLICE_Q], C1[SLICE_Q];
#pragma unroll
        for (uint32_t wi = 0; wi < TILE_WORDS; wi++) {
            uint32_t viol[SLICE_Q];
            rows.next(viol);
#pragma unroll
            for (uint32_t q = 0; q < SLICE_Q; q++) {
                if ((wi & 1u) == 0) {
                    A[q] = viol[q];
                } else if ((wi & 3u) == 1) {
                    csa(P[0][q], C1[q], A[q], viol[q]);
                } else {
                    uint32_t c2, c3;
                    csa(P[0][q], c2, A[q], viol[q]);
                    csa(P[1][q], c3, C1[q], c2);
#pragma unroll
                    for (int p = 2; p < NP - 1; p++) {  // ripple the fours
                        const uint32_t tcarry = P[p][q] & c3;
                        P[p][q] ^= c3;
                        c3 = tcarry;
                    }
                    P[NP - 1][q] ^= c3;
                }
            }
        }
        if (++since_flush == FLUSH_TILES) {
            flush();
            since_flush = 0;
        }
    }
    if (since_flush) flush();
#pragma unroll
    for (uint32_t q = 0; q < SLICE_Q; q++) {
        const uint32_t a = 32u * q + lane;
        if (a < valid && my_energy[q]) atomicAdd(energy + b0 + a, my_energy[q]);
    }
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
