/*
 * sls_oracle.c -- CPU restatement of rust/src/lib.rs (reference SLS solver).
 * TEST INFRASTRUCTURE ONLY; see sls_oracle.h for scope, citations and the
 * parity status ("clause evaluation pinned, walk parity unpinned").
 *
 * Plain C11 + pthreads.  Build: make -C oracle  (gcc -O3 -march=native).
 */
#define _GNU_SOURCE
#include "sls_oracle.h"

#include <errno.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------ errors */

static _Thread_local char g_err[512];

static int fail(int code, const char *fmt, const char *arg)
{
    snprintf(g_err, sizeof g_err, fmt, arg ? arg : "");
    return code;
}

const char *orc_last_error(void) { return g_err; }

/* ---------------------------------------------------------------- instance */

struct orc_instance {
    uint32_t n;         /* formula.var_count() */
    uint32_t m;         /* formula.len() */
    uint64_t L;
    uint64_t *row_ptr;  /* [m+1] */
    uint32_t *lits;     /* [L] varisat code 2*var + neg */
    uint32_t *canon;    /* [m] first clause with identical literal sequence */
    uint32_t m_distinct;
    uint64_t *occ_ptr;  /* [n+1] var_to_clauses (lib.rs:186-199), with multiplicity */
    uint32_t *occ;      /* [L] clause ids in formula order */
};

void orc_instance_free(orc_instance *inst)
{
    if (!inst) return;
    free(inst->row_ptr);
    free(inst->lits);
    free(inst->canon);
    free(inst->occ_ptr);
    free(inst->occ);
    free(inst);
}

uint32_t orc_num_vars(const orc_instance *inst) { return inst->n; }
uint32_t orc_num_clauses(const orc_instance *inst) { return inst->m; }
uint64_t orc_num_lits(const orc_instance *inst) { return inst->L; }
uint32_t orc_num_distinct_clauses(const orc_instance *inst) { return inst->m_distinct; }

void orc_get_csr(const orc_instance *inst, uint64_t *row_ptr, int32_t *signed_lits)
{
    memcpy(row_ptr, inst->row_ptr, (size_t)(inst->m + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < inst->L; i++) {
        int32_t v = (int32_t)(inst->lits[i] >> 1) + 1;
        signed_lits[i] = (inst->lits[i] & 1) ? -v : v;
    }
}

static uint64_t hash_clause(const uint32_t *l, uint64_t k)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ k;
    for (uint64_t i = 0; i < k; i++) {
        h ^= l[i];
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    return h;
}

/* canon[], occurrence lists; takes ownership of row_ptr / lits */
static int finish_instance(orc_instance *inst)
{
    uint32_t m = inst->m, n = inst->n;
    inst->canon = malloc(sizeof(uint32_t) * (m ? m : 1));
    inst->occ_ptr = calloc((size_t)n + 2, sizeof(uint64_t));
    inst->occ = malloc(sizeof(uint32_t) * (inst->L ? inst->L : 1));
    if (!inst->canon || !inst->occ_ptr || !inst->occ) return fail(ORC_ERR_ARG, "out of memory%s", NULL);

    /* content-keyed identity of clauses: HashSet<&[Lit]> compares slices by value */
    uint64_t cap = 16;
    while (cap < 2ull * m) cap <<= 1;
    uint32_t *table = malloc(sizeof(uint32_t) * cap);
    if (!table) return fail(ORC_ERR_ARG, "out of memory%s", NULL);
    memset(table, 0xff, sizeof(uint32_t) * cap);
    inst->m_distinct = 0;
    for (uint32_t c = 0; c < m; c++) {
        const uint32_t *l = inst->lits + inst->row_ptr[c];
        uint64_t k = inst->row_ptr[c + 1] - inst->row_ptr[c];
        uint64_t slot = hash_clause(l, k) & (cap - 1);
        for (;;) {
            uint32_t o = table[slot];
            if (o == 0xffffffffu) {
                table[slot] = c;
                inst->canon[c] = c;
                inst->m_distinct++;
                break;
            }
            uint64_t ko = inst->row_ptr[o + 1] - inst->row_ptr[o];
            if (ko == k && memcmp(inst->lits + inst->row_ptr[o], l, k * sizeof(uint32_t)) == 0) {
                inst->canon[c] = o;
                break;
            }
            slot = (slot + 1) & (cap - 1);
        }
    }
    free(table);

    /* var_to_clauses: for clause in formula, for lit in clause: push(clause)  (lib.rs:190-195) */
    for (uint64_t i = 0; i < inst->L; i++) inst->occ_ptr[(inst->lits[i] >> 1) + 1]++;
    for (uint32_t v = 0; v < n; v++) inst->occ_ptr[v + 1] += inst->occ_ptr[v];
    uint64_t *fill = malloc(sizeof(uint64_t) * ((size_t)n + 1));
    if (!fill) return fail(ORC_ERR_ARG, "out of memory%s", NULL);
    memcpy(fill, inst->occ_ptr, sizeof(uint64_t) * ((size_t)n + 1));
    for (uint32_t c = 0; c < m; c++)
        for (uint64_t j = inst->row_ptr[c]; j < inst->row_ptr[c + 1]; j++) inst->occ[fill[inst->lits[j] >> 1]++] = c;
    free(fill);
    return ORC_OK;
}

int orc_instance_from_csr(uint32_t n, uint32_t m, const uint64_t *row_ptr, const int32_t *signed_lits,
                          orc_instance **out)
{
    *out = NULL;
    orc_instance *inst = calloc(1, sizeof *inst);
    if (!inst) return fail(ORC_ERR_ARG, "out of memory%s", NULL);
    inst->n = n;
    inst->m = m;
    inst->L = row_ptr[m];
    inst->row_ptr = malloc(sizeof(uint64_t) * ((size_t)m + 1));
    inst->lits = malloc(sizeof(uint32_t) * (inst->L ? inst->L : 1));
    if (!inst->row_ptr || !inst->lits) {
        orc_instance_free(inst);
        return fail(ORC_ERR_ARG, "out of memory%s", NULL);
    }
    memcpy(inst->row_ptr, row_ptr, sizeof(uint64_t) * ((size_t)m + 1));
    for (uint64_t i = 0; i < inst->L; i++) {
        int32_t s = signed_lits[i];
        uint32_t v = (uint32_t)(s < 0 ? -(int64_t)s : s);
        if (s == 0 || v > n) {
            orc_instance_free(inst);
            return fail(ORC_ERR_ARG, "literal out of range%s", NULL);
        }
        inst->lits[i] = ((v - 1) << 1) | (s < 0);
    }
    int rc = finish_instance(inst);
    if (rc) {
        orc_instance_free(inst);
        return rc;
    }
    *out = inst;
    return ORC_OK;
}

/* DIMACS CNF.  Restates varisat 0.2.2's DimacsParser acceptance as used at
 * lib.rs:118: 'c' comment lines, an optional "p cnf <vars> <clauses>" header
 * whose counts are CHECKED (more variables than declared, or a different
 * number of clauses, is a parse error), clauses are 0-terminated and may span
 * lines, an unterminated last clause is an error.  var_count = header count
 * when a header is present, else the largest variable seen. */
int orc_instance_from_dimacs_text(const char *text, size_t len, orc_instance **out)
{
    *out = NULL;
    size_t cap_l = 1024, cap_c = 256;
    uint32_t *lits = malloc(sizeof(uint32_t) * cap_l);
    uint64_t *row_ptr = malloc(sizeof(uint64_t) * cap_c);
    if (!lits || !row_ptr) {
        free(lits);
        free(row_ptr);
        return fail(ORC_ERR_ARG, "out of memory%s", NULL);
    }
    uint64_t L = 0, m = 0, clause_start = 0;
    uint64_t max_var = 0;
    int have_header = 0;
    uint64_t hdr_vars = 0, hdr_clauses = 0;
    size_t i = 0;
    int line_start = 1;
    const char *why = NULL;
    row_ptr[0] = 0;
    while (i < len) {
        char ch = text[i];
        if (ch == '\n') {
            line_start = 1;
            i++;
            continue;
        }
        if (ch == ' ' || ch == '\t' || ch == '\r') {
            i++;
            continue;
        }
        if (line_start && ch == 'c') {
            while (i < len && text[i] != '\n') i++;
            continue;
        }
        if (line_start && ch == 'p') {
            if (have_header || m > 0 || L > clause_start) {
                why = "unexpected header";
                break;
            }
            char fmt[8] = {0};
            unsigned long long a = 0, b = 0;
            size_t e = i;
            while (e < len && text[e] != '\n') e++;
            char buf[256];
            size_t ll = e - i < sizeof buf - 1 ? e - i : sizeof buf - 1;
            memcpy(buf, text + i, ll);
            buf[ll] = 0;
            if (sscanf(buf, "p %7s %llu %llu", fmt, &a, &b) != 3 || strcmp(fmt, "cnf") != 0) {
                why = "invalid header";
                break;
            }
            have_header = 1;
            hdr_vars = a;
            hdr_clauses = b;
            i = e;
            continue;
        }
        line_start = 0;
        int neg = 0;
        if (ch == '-') {
            neg = 1;
            i++;
            if (i >= len) {
                why = "dangling '-'";
                break;
            }
            ch = text[i];
        }
        if (ch < '0' || ch > '9') {
            why = "unexpected character";
            break;
        }
        uint64_t v = 0;
        while (i < len && text[i] >= '0' && text[i] <= '9') {
            v = v * 10 + (uint64_t)(text[i] - '0');
            if (v > 0x7fffffffull) {
                why = "literal too large";
                break;
            }
            i++;
        }
        if (why) break;
        if (i < len && !(text[i] == ' ' || text[i] == '\t' || text[i] == '\r' || text[i] == '\n')) {
            why = "unexpected character";
            break;
        }
        if (v == 0) {
            if (neg) {
                why = "negative zero";
                break;
            }
            if (m + 2 > cap_c) {
                cap_c *= 2;
                row_ptr = realloc(row_ptr, sizeof(uint64_t) * cap_c);
            }
            m++;
            row_ptr[m] = L;
            clause_start = L;
        } else {
            if (L + 1 > cap_l) {
                cap_l *= 2;
                lits = realloc(lits, sizeof(uint32_t) * cap_l);
            }
            if (v > max_var) max_var = v;
            lits[L++] = (uint32_t)(((v - 1) << 1) | (uint64_t)neg);
        }
    }
    if (!why && L > clause_start) why = "unterminated clause";
    if (!why && have_header && max_var > hdr_vars) why = "more variables than the header declares";
    if (!why && have_header && m != hdr_clauses) why = "clause count differs from the header";
    if (!why && m > 0xfffffffeull) why = "too many clauses";
    if (why) {
        free(lits);
        free(row_ptr);
        return fail(ORC_ERR_PARSE, "parse error: %s", why);
    }
    orc_instance *inst = calloc(1, sizeof *inst);
    inst->n = (uint32_t)(have_header ? hdr_vars : max_var);
    inst->m = (uint32_t)m;
    inst->L = L;
    inst->row_ptr = row_ptr;
    inst->lits = lits;
    int rc = finish_instance(inst);
    if (rc) {
        orc_instance_free(inst);
        return rc;
    }
    *out = inst;
    return ORC_OK;
}

int orc_instance_from_dimacs(const char *path, orc_instance **out)
{
    *out = NULL;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(ORC_ERR_IO, "failed to read %s", path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = malloc((size_t)sz + 1);
    if (!buf || fread(buf, 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        free(buf);
        return fail(ORC_ERR_IO, "failed to read %s", path);
    }
    fclose(f);
    int rc = orc_instance_from_dimacs_text(buf, (size_t)sz, out);
    free(buf);
    return rc;
}

/* ------------------------------------------------------ clause evaluation */

/* lib.rs:13-18: all(lit.is_positive() != assignment[var]); empty clause => true */
static inline int clause_violated(const orc_instance *inst, uint32_t c, const uint8_t *a)
{
    for (uint64_t j = inst->row_ptr[c]; j < inst->row_ptr[c + 1]; j++) {
        uint32_t l = inst->lits[j];
        int positive = !(l & 1);
        if (positive == (a[l >> 1] != 0)) return 0;
    }
    return 1;
}

int orc_clause_is_violated(const orc_instance *inst, uint32_t clause, const uint8_t *assignment)
{
    return clause_violated(inst, clause, assignment);
}

uint32_t orc_eval_clauses(const orc_instance *inst, const uint8_t *assignment, uint8_t *violated_out)
{
    uint32_t e = 0;
    for (uint32_t c = 0; c < inst->m; c++) {
        int v = clause_violated(inst, c, assignment);
        if (violated_out) violated_out[c] = (uint8_t)v;
        e += (uint32_t)v;
    }
    return e;
}

uint32_t orc_energy_distinct(const orc_instance *inst, const uint8_t *assignment)
{
    uint32_t e = 0;
    for (uint32_t c = 0; c < inst->m; c++)
        if (inst->canon[c] == c && clause_violated(inst, c, assignment)) e++;
    return e;
}

/* ------------------------------------------------------------------ Philox */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

/* Per-chain streams.  counter = (block_lo, block_hi, chain_lo, chain_hi | stream<<30),
 * key = seed.  Stream 0 ("init"): block b holds the two u64 for variables 2b, 2b+1.
 * Stream 1 ("walk"): a sequence of u32 words consumed in reference draw order
 * (32-bit draws take one word, 64-bit draws two words at an even position). */
enum { STREAM_INIT = 0, STREAM_WALK = 1 };

typedef struct {
    uint32_t key[2];
    uint32_t chain_lo, chain_hi_stream;
    uint64_t pos;        /* next word */
    uint64_t cached_blk; /* block held in buf, ~0 = none */
    uint32_t buf[4];
} rng_stream;

static void stream_init(rng_stream *s, uint64_t seed, uint64_t chain, int stream)
{
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->chain_lo = (uint32_t)chain;
    s->chain_hi_stream = ((uint32_t)(chain >> 32) & 0x3fffffffu) | ((uint32_t)stream << 30);
    s->pos = 0;
    s->cached_blk = ~0ull;
}

static inline void stream_block(const rng_stream *s, uint64_t blk, uint32_t out[4])
{
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), s->chain_lo, s->chain_hi_stream};
    orc_philox4x32_10(ctr, s->key, out);
}

static inline uint32_t next_u32(rng_stream *s)
{
    uint64_t blk = s->pos >> 2;
    if (blk != s->cached_blk) {
        stream_block(s, blk, s->buf);
        s->cached_blk = blk;
    }
    return s->buf[s->pos++ & 3];
}

/* 64-bit draws sit on even word positions (an odd cursor skips one word), so a u64
 * never straddles two Philox blocks and lanes can fetch the draw of literal j directly. */
static inline uint64_t next_u64(rng_stream *s)
{
    s->pos = (s->pos + 1) & ~1ull;
    uint64_t lo = next_u32(s);
    uint64_t hi = next_u32(s);
    return lo | (hi << 32);
}

/* rand 0.8.5 Standard for f64: (next_u64() >> 11) * 2^-53 */
static inline double next_f64(rng_stream *s) { return (double)(next_u64(s) >> 11) * (1.0 / 9007199254740992.0); }

/* rand 0.8.5 UniformInt<u32>::sample_single(0, range): widening multiply, zone rejection */
static inline uint32_t uniform_u32(rng_stream *s, uint32_t range)
{
    uint32_t zone = (range << __builtin_clz(range)) - 1u;
    for (;;) {
        uint64_t p = (uint64_t)next_u32(s) * range;
        if ((uint32_t)p <= zone) return (uint32_t)(p >> 32);
    }
}

/* rand 0.8.5 UniformInt<usize>::sample_single on a 64-bit target */
static inline uint64_t uniform_u64(rng_stream *s, uint64_t range)
{
    uint64_t zone = (range << __builtin_clzll(range)) - 1ull;
    for (;;) {
        unsigned __int128 p = (unsigned __int128)next_u64(s) * range;
        if ((uint64_t)p <= zone) return (uint64_t)(p >> 64);
    }
}

/* rand 0.8.5 Bernoulli::new(p): p_int = (p * 2^64) as u64, p == 1.0 => ALWAYS_TRUE */
static int bernoulli_threshold(double p, uint64_t *t)
{
    if (!(p >= 0.0 && p < 1.0)) {
        if (p == 1.0) {
            *t = ~0ull;
            return 0;
        }
        return -1;
    }
    *t = (uint64_t)(p * 18446744073709551616.0);
    return 0;
}

/* init stream: variable i draws the u64 at words 2i, 2i+1 of stream 0 (lib.rs:208-210) */
static void init_assignment(const orc_instance *inst, const uint64_t *thr, uint64_t seed, uint64_t chain, uint8_t *a)
{
    rng_stream s;
    stream_init(&s, seed, chain, STREAM_INIT);
    for (uint32_t i = 0; i < inst->n; i++) {
        uint64_t v = next_u64(&s);
        a[i] = (thr[i] == ~0ull) || (v < thr[i]);
    }
}

int orc_init_assignment(const orc_instance *inst, const double *w_init, uint64_t seed, uint64_t chain,
                        uint8_t *assignment_out)
{
    uint64_t *thr = malloc(sizeof(uint64_t) * ((size_t)inst->n + 1));
    for (uint32_t i = 0; i < inst->n; i++)
        if (bernoulli_threshold(w_init[i], &thr[i])) {
            free(thr);
            return fail(ORC_ERR_WEIGHTS, "p is outside range [0.0, 1.0]%s", NULL);
        }
    init_assignment(inst, thr, seed, chain, assignment_out);
    free(thr);
    return ORC_OK;
}

/* ------------------------------------------------------------- one chain */

typedef struct {
    const orc_instance *inst;
    const uint64_t *thr_init;
    const double *w;
    const orc_run_params *p;
    orc_run_result *res;
    uint8_t *best_assign_all; /* [nruns][n] scratch for per-run best assignments (cloned lib.rs:294) */
    uint8_t *has_best_all;    /* [nruns] */
    atomic_ullong next_run;
    atomic_int error;
} run_ctx;

typedef struct {
    uint8_t *a;       /* assignment */
    uint64_t *set;    /* violated-set bitmap over canonical clause ids */
    uint32_t numfalse;
    uint32_t *flipped;
    size_t flipped_cap;
    double *wbuf;
} chain_ws;

static inline void set_insert(chain_ws *w, uint32_t id)
{
    uint64_t bit = 1ull << (id & 63);
    if (!(w->set[id >> 6] & bit)) {
        w->set[id >> 6] |= bit;
        w->numfalse++;
    }
}
static inline void set_remove(chain_ws *w, uint32_t id)
{
    uint64_t bit = 1ull << (id & 63);
    if (w->set[id >> 6] & bit) {
        w->set[id >> 6] &= ~bit;
        w->numfalse--;
    }
}

/* find_violated_clauses (lib.rs:176-181) */
static void set_rescan(const orc_instance *inst, chain_ws *w)
{
    memset(w->set, 0, sizeof(uint64_t) * (((size_t)inst->m + 63) / 64 + 1));
    w->numfalse = 0;
    for (uint32_t c = 0; c < inst->m; c++)
        if (clause_violated(inst, c, w->a)) set_insert(w, inst->canon[c]);
}

/* r-th (0-based) member in increasing canonical clause index */
static uint32_t set_select(const orc_instance *inst, const chain_ws *w, uint32_t r)
{
    size_t words = ((size_t)inst->m + 63) / 64;
    for (size_t i = 0; i < words; i++) {
        uint32_t pc = (uint32_t)__builtin_popcountll(w->set[i]);
        if (r < pc) {
            uint64_t x = w->set[i];
            while (r--) x &= x - 1;
            return (uint32_t)(i * 64 + (size_t)__builtin_ctzll(x));
        }
        r -= pc;
    }
    return 0xffffffffu; /* unreachable when r < numfalse */
}

static inline double flip_prob(const chain_ws *w, const double *weights, uint32_t idx)
{
    double fa = w->a[idx] ? 1.0 : 0.0;
    return (1.0 - fa) * weights[idx] + fa * (1.0 - weights[idx]);
}

static int run_chain(run_ctx *ctx, chain_ws *ws, uint64_t run)
{
    const orc_instance *inst = ctx->inst;
    const orc_run_params *p = ctx->p;
    orc_run_result *res = ctx->res;
    const uint32_t n = inst->n, m = inst->m;
    uint64_t chain = p->chain_offset + run;
    uint32_t *traj = (p->return_trajectories && res->traj) ? res->traj + run * (p->nsteps + 1) : NULL;
    uint64_t traj_n = 0;

    uint64_t best_energy_run = m; /* lib.rs:203 */
    uint8_t *best_a = ctx->best_assign_all + run * (size_t)n;
    int has_best = 0;

    init_assignment(inst, ctx->thr_init, p->seed, chain, ws->a); /* lib.rs:206-210 */
    rng_stream rng;
    stream_init(&rng, p->seed, chain, STREAM_WALK);

    uint64_t numstep = 0, flips = 0;
    set_rescan(inst, ws); /* lib.rs:213-214 */
    if (ws->numfalse < best_energy_run) { /* lib.rs:216-219 */
        best_energy_run = ws->numfalse;
        memcpy(best_a, ws->a, n);
        has_best = 1;
    }
    if (traj) traj[traj_n++] = ws->numfalse; /* lib.rs:221-223 */

    while (ws->numfalse > 0 && numstep < p->nsteps) { /* lib.rs:225 */
        numstep++;
        /* lib.rs:229 violated_clauses.iter().choose(&mut rng): gen_index(len) then nth */
        uint32_t c = set_select(inst, ws, uniform_u32(&rng, ws->numfalse));
        const uint32_t *cl = inst->lits + inst->row_ptr[c];
        uint32_t k = (uint32_t)(inst->row_ptr[c + 1] - inst->row_ptr[c]);
        size_t nflipped = 0;
        if ((size_t)k + 1 > ws->flipped_cap) {
            ws->flipped_cap = (size_t)k + 1;
            ws->flipped = realloc(ws->flipped, sizeof(uint32_t) * ws->flipped_cap);
            ws->wbuf = realloc(ws->wbuf, sizeof(double) * ws->flipped_cap);
        }

        if (next_f64(&rng) < p->prob_flip_best) { /* lib.rs:231-268 */
            uint32_t best_idx = 0;
            uint64_t min_viol = ~0ull;
            for (uint32_t j = 0; j < k; j++) {
                uint32_t idx = cl[j] >> 1;
                ws->a[idx] = !ws->a[idx];
                uint64_t viol = 0;
                if (p->pre_compute_mapping) {
                    for (uint64_t o = inst->occ_ptr[idx]; o < inst->occ_ptr[idx + 1]; o++)
                        viol += (uint64_t)clause_violated(inst, inst->occ[o], ws->a);
                } else {
                    viol = orc_energy_distinct(inst, ws->a);
                }
                ws->a[idx] = !ws->a[idx];
                if (viol < min_viol) {
                    min_viol = viol;
                    best_idx = idx;
                }
            }
            if (n == 0) return ORC_ERR_EMPTY_CLAUSE; /* assignment[0] out of bounds */
            ws->a[best_idx] = !ws->a[best_idx];
            ws->flipped[0] = best_idx;
            nflipped = 1;
        } else {
            switch (p->algo) {
            case ORC_ALGO_MOSER: /* lib.rs:20-49 */
                for (uint32_t j = 0; j < k; j++) {
                    uint32_t idx = cl[j] >> 1;
                    double fp = flip_prob(ws, ctx->w, idx);
                    if (next_f64(&rng) < fp) ws->flipped[nflipped++] = idx;
                }
                for (size_t j = 0; j < nflipped; j++) ws->a[ws->flipped[j]] = !ws->a[ws->flipped[j]];
                break;
            case ORC_ALGO_MOSER_SINGLE: /* lib.rs:77-101 */
                for (uint32_t j = 0; j < k; j++) {
                    uint32_t idx = cl[j] >> 1;
                    double fp = flip_prob(ws, ctx->w, idx);
                    if (next_f64(&rng) < fp) ws->flipped[nflipped++] = idx;
                }
                if (nflipped) {
                    uint32_t chosen = ws->flipped[uniform_u64(&rng, nflipped)];
                    ws->a[chosen] = !ws->a[chosen];
                    ws->flipped[0] = chosen;
                    nflipped = 1;
                }
                break;
            case ORC_ALGO_SCHOENING: /* lib.rs:51-55 */
                if (k == 0) return ORC_ERR_EMPTY_CLAUSE;
                {
                    uint32_t idx = cl[uniform_u32(&rng, k)] >> 1;
                    ws->a[idx] = !ws->a[idx];
                    ws->flipped[0] = idx;
                    nflipped = 1;
                }
                break;
            case ORC_ALGO_PROBSAT: /* lib.rs:57-75 */
            {
                if (k == 0) return ORC_ERR_EMPTY_CLAUSE; /* WeightedIndex::new: NoItem */
                /* WeightedIndex::new: cumulative sums in order, total */
                double total = 0.0;
                for (uint32_t j = 0; j < k; j++) {
                    double wj = flip_prob(ws, ctx->w, cl[j] >> 1);
                    if (!(wj >= 0.0)) return ORC_ERR_WEIGHTS; /* InvalidWeight */
                    if (j == 0)
                        total = wj;
                    else {
                        ws->wbuf[j - 1] = total;
                        total += wj;
                    }
                }
                if (total == 0.0) return ORC_ERR_ALL_ZERO;
                /* UniformFloat::new(0, total): shrink scale until scale*max_rand < high */
                double scale = total;
                const double max_rand = 1.0 - 2.220446049250313e-16;
                while (scale * max_rand >= total) {
                    uint64_t bits;
                    memcpy(&bits, &scale, 8);
                    bits--;
                    memcpy(&scale, &bits, 8);
                }
                /* sample: value0_1 from 52 mantissa bits */
                double x = ((double)(next_u64(&rng) >> 12) * (1.0 / 4503599627370496.0)) * scale;
                uint32_t pick = 0; /* partition_point(|w| w <= x) over k-1 cumulative weights */
                while (pick < k - 1 && ws->wbuf[pick] <= x) pick++;
                uint32_t idx = cl[pick] >> 1;
                ws->a[idx] = !ws->a[idx];
                ws->flipped[0] = idx;
                nflipped = 1;
                break;
            }
            default:
                return ORC_ERR_ALGO;
            }
        }
        flips += nflipped;

        if (p->pre_compute_mapping) { /* lib.rs:278-287 */
            for (size_t f = 0; f < nflipped; f++) {
                uint32_t idx = ws->flipped[f];
                for (uint64_t o = inst->occ_ptr[idx]; o < inst->occ_ptr[idx + 1]; o++) {
                    uint32_t cc = inst->occ[o];
                    if (clause_violated(inst, cc, ws->a))
                        set_insert(ws, inst->canon[cc]);
                    else
                        set_remove(ws, inst->canon[cc]);
                }
            }
        } else { /* lib.rs:288-290 */
            set_rescan(inst, ws);
        }

        if (ws->numfalse < best_energy_run) { /* lib.rs:292-296 */
            best_energy_run = ws->numfalse;
            memcpy(best_a, ws->a, n);
            has_best = 1;
        }
        if (traj) traj[traj_n++] = ws->numfalse; /* lib.rs:298-300 */
    }

    res->found[run] = ws->numfalse == 0; /* lib.rs:305-307 */
    res->steps[run] = numstep;
    if (res->best_energy_run) res->best_energy_run[run] = best_energy_run;
    if (res->flips_run) res->flips_run[run] = flips;
    if (res->traj_len) res->traj_len[run] = traj_n;
    ctx->has_best_all[run] = (uint8_t)has_best;
    return ORC_OK;
}

static void *worker(void *arg)
{
    run_ctx *ctx = arg;
    const orc_instance *inst = ctx->inst;
    chain_ws ws;
    memset(&ws, 0, sizeof ws);
    ws.a = malloc((size_t)inst->n + 1);
    ws.set = malloc(sizeof(uint64_t) * (((size_t)inst->m + 63) / 64 + 1));
    for (;;) {
        if (atomic_load(&ctx->error)) break;
        uint64_t run = atomic_fetch_add(&ctx->next_run, 1);
        if (run >= ctx->p->nruns) break;
        int rc = run_chain(ctx, &ws, run);
        if (rc) {
            int zero = 0;
            atomic_compare_exchange_strong(&ctx->error, &zero, rc);
            break;
        }
    }
    free(ws.a);
    free(ws.set);
    free(ws.flipped);
    free(ws.wbuf);
    return NULL;
}

int orc_run_sls(const orc_instance *inst, const double *w_init, size_t w_init_len, const double *w_resample,
                size_t w_resample_len, const orc_run_params *params, orc_run_result *res)
{
    if (params->algo < 0 || params->algo > 3) return fail(ORC_ERR_ALGO, "Unknown algorithm type%s", NULL);
    const uint32_t n = inst->n;
    res->best_energy = inst->m;
    res->has_best_assignment = 0;
    res->total_steps = 0;
    res->total_flips = 0;
    if (params->nruns == 0) return ORC_OK;
    /* index-out-of-bounds panics of lib.rs:209 / :29 */
    if (w_init_len < n) return fail(ORC_ERR_WEIGHTS, "weights_initialize shorter than var_count%s", NULL);
    if (w_resample_len < n && params->algo != ORC_ALGO_SCHOENING)
        return fail(ORC_ERR_WEIGHTS, "weights_resample shorter than var_count%s", NULL);

    uint64_t *thr = malloc(sizeof(uint64_t) * ((size_t)n + 1));
    for (uint32_t i = 0; i < n; i++)
        if (bernoulli_threshold(w_init[i], &thr[i])) {
            free(thr);
            return fail(ORC_ERR_WEIGHTS, "p is outside range [0.0, 1.0]%s", NULL);
        }

    run_ctx ctx;
    memset(&ctx, 0, sizeof ctx);
    ctx.inst = inst;
    ctx.thr_init = thr;
    ctx.w = w_resample;
    ctx.p = params;
    ctx.res = res;
    ctx.best_assign_all = malloc((size_t)params->nruns * ((size_t)n ? n : 1));
    ctx.has_best_all = calloc(params->nruns, 1);
    atomic_init(&ctx.next_run, 0);
    atomic_init(&ctx.error, 0);

    int nt = params->nthreads;
    if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if ((uint64_t)nt > params->nruns) nt = (int)params->nruns;
    if (nt == 1) {
        worker(&ctx);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nt);
        for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, worker, &ctx);
        for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
        free(th);
    }

    int rc = atomic_load(&ctx.error);
    if (!rc) {
        /* lib.rs:320-334: strict '<' arg-min in run order, starting from formula.len() */
        uint64_t best = inst->m;
        for (uint64_t r = 0; r < params->nruns; r++) {
            uint64_t e = res->best_energy_run ? res->best_energy_run[r] : 0;
            if (!res->best_energy_run) {
                /* recompute from the stored per-run best assignment */
                e = ctx.has_best_all[r] ? orc_energy_distinct(inst, ctx.best_assign_all + r * (size_t)n) : inst->m;
            }
            if (e < best) {
                best = e;
                if (res->best_assignment) memcpy(res->best_assignment, ctx.best_assign_all + r * (size_t)n, n);
                res->has_best_assignment = 1;
            }
            res->total_steps += res->steps[r];
            if (res->flips_run) res->total_flips += res->flips_run[r];
        }
        res->best_energy = best;
    } else {
        const char *msg = rc == ORC_ERR_ALL_ZERO       ? "AllWeightsZero%s"
                          : rc == ORC_ERR_EMPTY_CLAUSE ? "empty clause selected%s"
                          : rc == ORC_ERR_WEIGHTS      ? "InvalidWeight%s"
                                                       : "chain failed%s";
        fail(rc, msg, NULL);
    }
    free(thr);
    free(ctx.best_assign_all);
    free(ctx.has_best_all);
    return rc;
}
