// common.h -- shared host/device definitions of the B200 SLS engine.
#pragma once
#include <cstddef>
#include <cstdint>

#include <vector_types.h>  // uint4 (host and device)

namespace slsb {

// Device view of one CNF instance.  The clause CSR replaces varisat's CnfFormula
// (reference rust/src/lib.rs:118), the occurrence CSR replaces var_to_clauses
// (lib.rs:186-199, clause ids in formula order, one entry per literal occurrence).
struct DevInst {
    uint32_t n;        // variables
    uint32_t m;        // clauses (formula.len(), duplicates included)
    uint32_t nwords;   // ceil((n+1)/32) assignment bitmap words; bit n is a dummy variable that stays 0
    uint32_t mwords;   // ceil(m/32)   violated-set bitmap words
    uint32_t ngroups;  // ceil(mwords/32) level-1 counters (1024 clauses each)
    // chain-state slab layout (words): abits at 0, vbits at off_v, level-1 counters at off_l1;
    // the counter array is zero-padded to l1_pad (a multiple of 64) entries so that a lane can
    // fetch its two counters with one 8-byte load
    uint32_t off_v, off_l1, l1_pad, slab_words;
    // second counter level (one counter per 64 level-1 counters = 65,536 clauses), only when
    // l1_pad > 64; zero-padded to l2_pad (a multiple of 64) entries at word offset off_l2
    uint32_t off_l2, l2_pad;
    uint32_t uniform_k;// clause length when all clauses share it, else 0
    uint32_t max_k;
    uint32_t flags;    // INST_*
    const uint32_t* row_ptr;  // [m+1]
    const uint32_t* lits;     // [L]  2*var + neg (varisat Lit code)
    const uint32_t* occ_ptr;  // [n+1]
    const uint32_t* occ;      // [L]  clause ids
    const uint32_t* canon;    // [m]  first clause with identical content; nullptr when all distinct
    // k<=3 fast path (nullptr when the instance does not qualify):
    const uint4* crec;        // [m]  {lit0, lit1, lit2, k}; absent literals = LIT_NONE
    const uint4* orec;        // [L]  per occurrence of var v: {clause id, other lit a, other lit b, neg(v in clause)}
};

enum : uint32_t { INST_HAS_DUPLICATES = 1u, INST_K3 = 2u };
static constexpr uint32_t LIT_NONE = 0xffffffffu;

enum : int { ALGO_MOSER = 0, ALGO_MOSER_SINGLE = 1, ALGO_SCHOENING = 2, ALGO_PROBSAT = 3 };

struct WalkArgs {
    DevInst inst;
    const uint64_t* thr_init;  // [n] Bernoulli thresholds of weights_initialize (lib.rs:208-210)
    const double* w;           // [n] weights_resample (input of the record / table builders; the walks read those)
    const uint4* crecx;        // [m][4] 64-byte clause record of the k<=3 kernels (walk_k3.cuh) or nullptr
    const double* ftab;        // [L]  generic kernel: per literal position, MT rules: flip probability; probSAT: running sum
    const double* cscale;      // [m]  generic kernel, probSAT: UniformFloat scale of the clause, or -(error code)
    int32_t algo;
    int32_t mapping;           // pre_compute_mapping (only changes the greedy score, lib.rs:242-250)
    int32_t want_traj;
    uint32_t reserved;
    double pfb;                // prob_flip_best
    uint64_t nsteps;
    uint64_t nruns;
    uint64_t seed;
    uint64_t chain_offset;
    uint32_t* gstate;          // [nruns][slab_words] chain state when it lives in HBM
    uint32_t warps_per_block;
    uint32_t pad2;
    uint32_t* best_bits;       // [nruns][nwords]
    uint32_t* nf_init;         // [nruns] initial energy when the slab was prepared by init_sliced.cuh, else nullptr
    uint32_t* flip_log;        // [nruns][log_cap] flipped variables (HBM tier with a bounded step budget) or nullptr
    uint64_t log_cap;
    uint8_t* found;            // [nruns]
    uint8_t* has_best;         // [nruns]
    uint64_t* steps;           // [nruns]
    uint32_t* best_energy;     // [nruns]
    uint64_t* flips;           // [nruns]
    uint32_t* traj;            // [nruns][nsteps+1] or nullptr
    int32_t* err;              // first error code raised by any chain
};

// error codes raised on the device (same numbering as include/sls_b200.h)
enum : int { DERR_WEIGHTS = 4, DERR_ALL_ZERO = 5, DERR_EMPTY_CLAUSE = 6, DERR_ARG = 7 };

}  // namespace slsb
