/*
 * sls_b200.h -- C ABI of the B200-native SLS engine (libsls_b200.so).
 *
 * This is the drop-in boundary for the reference's Rust/pyo3 module
 * `moser_rust` (reference: rust/src/lib.rs).  Every entry point names the
 * reference interface it replaces (file:line relative to the reference
 * repository).  Plain pointers and sizes only: no C++ types, no torch types,
 * no exceptions cross this boundary.  All functions return SLS_OK (0) or an
 * SLS_ERR_* code; sls_last_error() gives the message for the calling thread.
 *
 * There is NO CPU fallback behind this interface: every compute entry point
 * runs hand-written sm_100a CUDA kernels and fails with SLS_ERR_CUDA when no
 * usable device is present.
 *
 * Randomness: the reference seeds each chain from OS entropy (lib.rs:206) and
 * has no seed argument.  Here every chain owns two counter-based
 * Philox4x32-10 streams keyed by (seed, global chain id); see DESIGN.md
 * "Random streams".  Results are therefore independent of the number of GPUs
 * a set of chains is sharded over.
 */
#ifndef SLS_B200_H
#define SLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:157-163 enum AlgoType; the strings of lib.rs:120-126 map to these */
enum { SLS_ALGO_MOSER = 0, SLS_ALGO_MOSER_SINGLE = 1, SLS_ALGO_SCHOENING = 2, SLS_ALGO_PROBSAT = 3 };

/* the reference signals all of these as Rust panics (pyo3 PanicException) */
enum {
    SLS_OK = 0,
    SLS_ERR_IO = 1,           /* lib.rs:115  .expect("failed to read") */
    SLS_ERR_PARSE = 2,        /* lib.rs:118  .expect("parse error") */
    SLS_ERR_ALGO = 3,         /* lib.rs:125  panic!("Unknown algorithm type") */
    SLS_ERR_WEIGHTS = 4,      /* lib.rs:209  gen_bool(p) with p outside [0,1]; weights shorter than var_count;
                                 lib.rs:71   WeightedIndex InvalidWeight */
    SLS_ERR_ALL_ZERO = 5,     /* lib.rs:71   WeightedIndex::new(..).unwrap() on all-zero weights */
    SLS_ERR_EMPTY_CLAUSE = 6, /* lib.rs:52 / :71 choose()/WeightedIndex on an empty clause */
    SLS_ERR_ARG = 7,
    SLS_ERR_CUDA = 8,         /* no device, launch or allocation failure */
    SLS_ERR_NOMEM = 9
};

const char *sls_last_error(void);

/* ---- device ------------------------------------------------------------ */
int sls_device_count(int *count);
int sls_set_device(int device); /* per calling thread, like cudaSetDevice */
/* Grow the library's private device memory pool (current device) so that `bytes` are mapped and cached, capped at the pool's
 * release threshold (4 GB).  The pool otherwise grows on demand; mapping new memory waits for the kernels that are running,
 * so a host thread that uploads instances while another thread's walk kernel runs stalls for a kernel's length per growth
 * step (pipeline.solve_with_oracle reserves its working set up front).  No reference counterpart: the reference has no device. */
int sls_device_pool_reserve(uint64_t bytes);

/* ---- instance: replaces DimacsParser::parse + CnfFormula (lib.rs:115-118)
 *      and the var_to_clauses mapping (lib.rs:186-199).  The clause CSR and the
 *      occurrence CSR live in HBM for the lifetime of the handle. ----------- */
typedef struct sls_instance sls_instance;

int sls_instance_from_dimacs(const char *path, sls_instance **out);
int sls_instance_from_dimacs_text(const char *text, size_t len, sls_instance **out);
/* row_ptr[m+1], signed DIMACS literals (+v / -v, v in 1..n) */
int sls_instance_from_csr(uint32_t n, uint32_t m, const uint64_t *row_ptr, const int32_t *signed_lits,
                          sls_instance **out);
void sls_instance_free(sls_instance *inst);
uint32_t sls_instance_num_vars(const sls_instance *inst);    /* CnfFormula::var_count, lib.rs:186 */
uint32_t sls_instance_num_clauses(const sls_instance *inst); /* CnfFormula::len, lib.rs:203 */
uint64_t sls_instance_num_lits(const sls_instance *inst);
uint32_t sls_instance_max_clause_len(const sls_instance *inst);
int sls_instance_get_csr(const sls_instance *inst, uint64_t *row_ptr, int32_t *signed_lits);

/* ---- clause evaluation: replaces clause_is_violated (lib.rs:13-18),
 *      find_violated_clauses (lib.rs:176-181) and
 *      {LCG,VCG}.get_violated_constraints (python/src/sat_representations.py:718-741,
 *      :353-373).  assignments: B rows of n bytes (0/1), host memory.
 *      violated_bits (optional): B rows of ceil(m/32) words, bit c%32 of word
 *      c/32 set <=> clause c violated.  energy (optional): violated clauses
 *      per row, every clause counted on its own (the Python semantics). ----- */
int sls_eval_assignments(sls_instance *inst, const uint8_t *assignments, uint64_t batch, uint32_t *violated_bits,
                         uint32_t *energy);
/* device time (ms, CUDA events: transposition + evaluation kernels, no copies) of the calling thread's last
 * sls_eval_assignments -- measurement only, no reference equivalent */
double sls_eval_last_kernel_ms(void);
/* ... and of its evaluation kernel alone (the byte -> bit-slice transposition of the assignments excluded) */
double sls_eval_last_eval_kernel_ms(void);

/* ---- clause neighbourhood of the LLL loss: replaces LCG.get_constraint_graph
 *      (python/src/sat_representations.py:668-715, scipy C^T C of the variable-clause incidence).  Host-side, no GPU
 *      needed.  var_of / clause_of: variable and clause index of every literal occurrence (the reference passes
 *      floor(senders/2) and receivers - 2n of the LCG edge list without its last n pair edges).  Writes, for every
 *      clause in increasing index, its neighbours in increasing index as senders[e] = clause + 2n, receivers[e] =
 *      neighbour + 2n (the LCG node numbering the reference returns); *n_edge = number of edges.  Call with
 *      capacity = 0 (buffers may be NULL) to size the buffers; SLS_ERR_ARG if capacity is too small. ----- */
int sls_constraint_graph(uint32_t n_variables, uint32_t n_clauses, const int32_t *var_of, const int32_t *clause_of,
                         uint64_t n_occurrences, int64_t *senders, int64_t *receivers, uint64_t capacity, uint64_t *n_edge);

/* ---- solver: replaces run_sls_python / run_sls (lib.rs:103-149, :165-335) */
typedef struct {
    int32_t algo;                /* SLS_ALGO_*            (algo_type_as_str) */
    int32_t return_trajectories; /* lib.rs:110 */
    int32_t pre_compute_mapping; /* lib.rs:111 */
    int32_t reserved;
    uint64_t nsteps;             /* lib.rs:108 */
    uint64_t nruns;              /* lib.rs:109 */
    double prob_flip_best;       /* lib.rs:112 */
    uint64_t seed;               /* Philox key (no reference equivalent: lib.rs:206 uses OS entropy) */
    uint64_t chain_offset;       /* global id of run 0 of this call (sharding over GPUs) */
} sls_run_params;

typedef struct {
    /* caller-allocated host buffers (NULL = not wanted) */
    uint8_t *best_assignment;  /* [n]      lib.rs:334 .0, valid iff has_best_assignment */
    uint8_t *found;            /* [nruns]  lib.rs:334 .1 */
    uint64_t *steps;           /* [nruns]  lib.rs:334 .3 numstep_local */
    uint64_t *best_energy_run; /* [nruns]  per-run best energy (lib.rs:292-296) */
    uint64_t *flips_run;       /* [nruns]  assignment toggles per run (throughput metric) */
    uint32_t *traj;            /* [nruns][nsteps+1] lib.rs:334 .4 when return_trajectories */
    uint64_t *traj_len;        /* [nruns]  valid prefix of each trajectory row (= steps+1) */
    /* written by the call */
    uint64_t best_energy;      /* lib.rs:334 .2 */
    int32_t has_best_assignment; /* 0 <=> the reference returns an empty best_assignment */
    int32_t reserved;
    uint64_t total_steps;
    uint64_t total_flips;
    double kernel_ms;          /* device time of the walk kernel (CUDA events) */
} sls_run_result;

/* one call = one run_sls_python call on an already parsed instance: host buffers in,
 * host buffers out, H2D/D2H inside. */
int sls_run(sls_instance *inst, const double *w_init, size_t w_init_len, const double *w_resample,
            size_t w_resample_len, const sls_run_params *params, sls_run_result *res);

/* batched form: n_inst parsed instances, the same parameters (params->nruns chains EACH, no
 * trajectories), one launch per kernel variant instead of one small grid per instance.  Replaces the
 * sequential per-file loop of python/src/evaluate_with_given_params.py:108-138.  w_init / w_resample
 * are the per-instance weight vectors concatenated; instance i owns [var_offsets[i], var_offsets[i+1]).
 * Instance i runs with seed instance_seeds[i] (or params->seed + i when instance_seeds is NULL), i.e.
 * results[i] is identical to sls_run(insts[i], ..., that seed): sharding a batch over processes does not
 * change any instance's result.  results[i] carries caller-allocated per-instance buffers like sls_run. */
int sls_run_batch(sls_instance *const *insts, uint32_t n_inst, const double *w_init, const double *w_resample,
                  const uint64_t *var_offsets, const uint64_t *instance_seeds, const sls_run_params *params,
                  sls_run_result *results);

/* resident form: state for params->nruns chains stays in HBM between launches */
typedef struct sls_solver sls_solver;

int sls_solver_create(sls_instance *inst, const sls_run_params *params, sls_solver **out);
void sls_solver_free(sls_solver *s);
int sls_solver_set_weights(sls_solver *s, const double *w_init, size_t w_init_len, const double *w_resample,
                           size_t w_resample_len);
int sls_solver_set_seed(sls_solver *s, uint64_t seed, uint64_t chain_offset);
/* asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream): init + walk of all chains */
int sls_solver_launch(sls_solver *s, void *cuda_stream);
int sls_solver_sync(sls_solver *s);
/* device time of the last launch (ms, CUDA events on the launch stream); call after sync */
int sls_solver_last_kernel_ms(sls_solver *s, double *ms);
int sls_solver_fetch(sls_solver *s, sls_run_result *res);
/* Trajectories of the last launch, packed: row r (steps[r]+1 energies, rust/src/lib.rs:298-300) starts at
 * offsets[r] in `packed`; offsets has nruns+1 entries.  Call once with packed == NULL to learn the total
 * (offsets[nruns]), then with a buffer of at least that many words.  Only the valid prefixes cross PCIe. */
int sls_solver_fetch_trajectories(sls_solver *s, uint32_t *packed, uint64_t capacity_words, uint64_t *offsets);
/* Per-step mean and median energy over the runs of the last launch, computed on the device: every
 * trajectory zero-padded to nsteps+1 entries, then np.mean / np.median over runs per step -- the reduction
 * of python/src/evaluate_with_given_params.py:18-23, :141-156 (before its division by the clause count).
 * mean / median: host arrays of nsteps+1 doubles (either may be NULL).  Needs return_trajectories. */
int sls_solver_trajectory_stats(sls_solver *s, double *mean, double *median);
/* device-to-device export of the per-chain results for a collective (all device pointers,
 * any may be NULL): found u8[nruns], steps u64[nruns], best_energy u32[nruns], flips u64[nruns] */
int sls_solver_export_device(sls_solver *s, void *d_found, void *d_steps, void *d_best_energy, void *d_flips,
                             void *cuda_stream);
/* number of kernels the last launch enqueued (for launch accounting) */
int sls_solver_launch_count(const sls_solver *s);
/* description of the kernel variant the solver picked, e.g. "walk<k3,smem>" */
const char *sls_solver_variant(const sls_solver *s);

/* ---- oracle forward: replaces network.apply(params, graph) for the shipped network type
 *      ("interaction" + LCG: python/src/model.py:127-144, call sites
 *      python/src/evaluate_with_given_params.py:116, :297, :306 and python/src/train_utils.py:279)
 *      and LCG.get_model_probabilities (python/src/sat_representations.py:756-780).
 *      The dense layers run on tcgen05 tensor cores with FP32 accumulation: every operand is split into two fp16
 *      pieces (22 mantissa bits; three kind::f16 product terms) -- probabilities within 1e-5 of an fp64 forward.
 *      A model whose weights or LayerNorm outputs leave fp16's range is created with two bf16 pieces instead
 *      (16 bits, fp32's range); SLS_B200_GNN_PRECISION = fp16x2 | bf16x2 | tf32x3 overrides per call. ------------- */
typedef struct sls_gnn_model sls_gnn_model;

/* blob = all parameters as float32 in haiku construction order (python/src/model.py:11-61):
 *   edge embed w[2][D], b[D]; node embed w[2][D], b[D];
 *   per step: edge MLP w1[K1e][h1], b1[h1], w2[h1][h2], b2[h2], ln_scale[h2], ln_offset[h2];
 *             node MLP w1[K1n][h1], ... (same order);   K1e = de + 2 dn, K1n = dn + 2 h2,
 *             de = dn = D in step 0 and h2 afterwards;
 *   readout w[h2], b[1].   Weights are [in][out] row-major exactly as haiku stores them. */
int sls_gnn_model_create(const float *blob, uint64_t blob_len, int32_t num_steps, int32_t embed_dim, int32_t h1,
                         int32_t h2, sls_gnn_model **out);
void sls_gnn_model_free(sls_gnn_model *m);
/* jraph.GraphsTuple in, decoded_nodes[N] out (host buffers): nodes[N][2], edges[E][2] (the one-hot
 * features of python/src/sat_instances.py:65-73), senders[E], receivers[E].  kernel_ms optional. */
int sls_gnn_forward_graph(sls_gnn_model *m, uint64_t n_node, uint64_t n_edge, const float *nodes, const float *edges,
                          const int32_t *senders, const int32_t *receivers, float *decoded, double *kernel_ms);
/* batched form: builds the LCG graphs of n_inst parsed instances (LCG.get_graph,
 * python/src/sat_representations.py:607-666), runs one forward over the concatenation and returns
 * p[sum n_i] = P(x = True) per variable (and optionally decoded[sum N_i], N_i = 2 n_i + m_i).  The instances
 * become resident on the current device (as for sls_run_batch) and the graph arrays are derived there from their
 * CSRs; a batch with an empty clause (which the reference drops, sat_instances.py:50) builds them on the host. */
int sls_gnn_forward_lcg(sls_gnn_model *m, sls_instance *const *insts, uint32_t n_inst, float *prob, float *decoded,
                        double *kernel_ms);
int sls_gnn_launch_count(const sls_gnn_model *m);
/* measurement only (no reference equivalent): with profiling on, every forward brackets its kernels with CUDA events by
 * class and sls_gnn_get_profile returns the device milliseconds of the LAST forward (of its last chunk, for a chunked
 * batch): ms[0] edge updates, ms[1] node updates, ms[2] segment sums, ms[3] everything else (embeddings, indices,
 * readout).  The events serialise nothing but cost a few microseconds per launch: leave it off in timed runs. */
int sls_gnn_set_profile(sls_gnn_model *m, int on);
int sls_gnn_get_profile(const sls_gnn_model *m, double ms[4]);

#ifdef __cplusplus
}
#endif
#endif
