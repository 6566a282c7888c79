/*
 * sls_oracle.h -- CPU restatement of the reference's SLS solver path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The CUDA product path never links or calls it.
 *
 * What it restates (all file:line relative to /root/reference):
 *   rust/src/lib.rs:13-18    clause_is_violated
 *   rust/src/lib.rs:20-49    resample_clause            ("moser")
 *   rust/src/lib.rs:51-55    flip_literal               ("schoening")
 *   rust/src/lib.rs:57-75    flip_literal_probsat       ("probsat")
 *   rust/src/lib.rs:77-101   flip_literal_moser_single  ("moser_single")
 *   rust/src/lib.rs:165-335  run_sls (setup, chain loop, reduction)
 *   python/src/sat_representations.py:718-741  LCG.get_violated_constraints
 *
 * PARITY STATUS
 *   clause evaluation (orc_eval_clauses, orc_clause_is_violated):
 *     PINNED against the reference's known-answer tests
 *     (python/tests/test_sat_instances.py:69-76, :86-111, :121-135) and
 *     against every shipped *_sol.pkl (energy 0), see tests/test_oracle_*.py.
 *   the walk itself (orc_run_sls): PARITY UNPINNED.  The reference's tests
 *     never execute run_sls_python, the Rust crate cannot be built in this
 *     environment (no cargo/rustc, crates not vendored) and the reference
 *     seeds every chain from OS entropy (lib.rs:206), so no golden
 *     trajectory exists.  The restatement follows lib.rs line by line and is
 *     checked through invariants and distributional properties only.
 *
 * Third-party arithmetic restated from memory of the published sources
 * (not present under /root/reference, pinned in Cargo.toml:12-18):
 *   rand 0.8.5   Standard f64 (53 bits), Bernoulli/gen_bool (64-bit threshold),
 *                UniformInt::sample_single (widening multiply + zone rejection),
 *                UniformFloat (52-bit mantissa fill), gen_index (u32 path)
 *   rand_distr 0.4 / rand 0.8.5  WeightedIndex (sequential f64 cumulative
 *                sums, partition_point(cum <= x))
 *   varisat 0.2.2  DimacsParser (header counts are checked strictly)
 * The random *source* is not ChaCha12-from-OS-entropy (irreproducible by
 * construction) but a counter-based Philox4x32-10 stream per chain, so the
 * CUDA path can be compared trajectory-for-trajectory with this file.
 *
 * Data-structure note: the reference keeps the violated clauses in a
 * HashSet<&[Lit]> keyed by clause CONTENT (lib.rs:176-181): two clauses with
 * the same literal sequence are one set element.  Here the set is a bitmap
 * over canonical clause ids (canon[c] = first clause with identical content),
 * which has the same membership, the same cardinality and supports the same
 * uniform choice; `iter().choose()`'s element order (random SipHash keys in
 * the reference) is fixed to increasing canonical clause index.
 */
#ifndef SLS_ORACLE_H
#define SLS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ALGO_MOSER = 0, ORC_ALGO_MOSER_SINGLE = 1, ORC_ALGO_SCHOENING = 2, ORC_ALGO_PROBSAT = 3 };

enum {
    ORC_OK = 0,
    ORC_ERR_IO = 1,          /* lib.rs:115 "failed to read" */
    ORC_ERR_PARSE = 2,       /* lib.rs:118 "parse error" */
    ORC_ERR_ALGO = 3,        /* lib.rs:125 "Unknown algorithm type" */
    ORC_ERR_WEIGHTS = 4,     /* weights shorter than var_count / p outside [0,1] (lib.rs:209) */
    ORC_ERR_ALL_ZERO = 5,    /* WeightedIndex::new(..).unwrap() lib.rs:71 */
    ORC_ERR_EMPTY_CLAUSE = 6,/* choose() on an empty clause lib.rs:52 / :71 */
    ORC_ERR_ARG = 7
};

typedef struct orc_instance orc_instance;

/* varisat literal code: 2*var_index + (negative ? 1 : 0) */
int orc_instance_from_dimacs(const char *path, orc_instance **out);
int orc_instance_from_dimacs_text(const char *text, size_t len, orc_instance **out);
int orc_instance_from_csr(uint32_t n, uint32_t m, const uint64_t *row_ptr, const int32_t *signed_lits,
                          orc_instance **out);
void orc_instance_free(orc_instance *inst);
uint32_t orc_num_vars(const orc_instance *inst);
uint32_t orc_num_clauses(const orc_instance *inst);
uint64_t orc_num_lits(const orc_instance *inst);
uint32_t orc_num_distinct_clauses(const orc_instance *inst);
/* copies out the CSR (row_ptr has m+1 entries, lits L signed DIMACS literals) */
void orc_get_csr(const orc_instance *inst, uint64_t *row_ptr, int32_t *signed_lits);

/* lib.rs:13-18 on one clause; assignment is one byte (0/1) per variable */
int orc_clause_is_violated(const orc_instance *inst, uint32_t clause, const uint8_t *assignment);
/* sat_representations.py:718-741: per-clause 0/1, every clause counted on its own.
 * returns the number of violated clauses (duplicates counted separately). */
uint32_t orc_eval_clauses(const orc_instance *inst, const uint8_t *assignment, uint8_t *violated_out);
/* |HashSet| of lib.rs:176-181: distinct violated clause contents */
uint32_t orc_energy_distinct(const orc_instance *inst, const uint8_t *assignment);

/* Philox4x32-10 block (counter ctr[4], key[2]) -> out[4]; exposed for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the init assignment of one chain (lib.rs:208-210 with the Philox init stream) */
int orc_init_assignment(const orc_instance *inst, const double *w_init, uint64_t seed, uint64_t chain,
                        uint8_t *assignment_out);

typedef struct {
    int algo;                   /* ORC_ALGO_* */
    uint64_t nsteps;
    uint64_t nruns;
    int return_trajectories;
    int pre_compute_mapping;
    double prob_flip_best;
    uint64_t seed;              /* Philox key */
    uint64_t chain_offset;      /* global id of run 0 (multi-GPU sharding keeps streams stable) */
    int nthreads;               /* <=0: all online cores (mirrors rayon's global pool, lib.rs:201) */
} orc_run_params;

typedef struct {
    /* caller-allocated */
    uint8_t *best_assignment;   /* [n]; valid iff has_best_assignment */
    uint8_t *found;             /* [nruns] */
    uint64_t *steps;            /* [nruns] numstep_local */
    uint64_t *best_energy_run;  /* [nruns] or NULL */
    uint64_t *flips_run;        /* [nruns] or NULL: number of assignment toggles */
    uint32_t *traj;             /* [nruns][nsteps+1] or NULL */
    uint64_t *traj_len;         /* [nruns] or NULL */
    /* written */
    uint64_t best_energy;       /* lib.rs:320-325 */
    int has_best_assignment;    /* 0 <=> the reference returns an empty Vec */
    uint64_t total_steps;
    uint64_t total_flips;
} orc_run_result;

int orc_run_sls(const orc_instance *inst, const double *w_init, size_t w_init_len, const double *w_resample,
                size_t w_resample_len, const orc_run_params *params, orc_run_result *res);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
