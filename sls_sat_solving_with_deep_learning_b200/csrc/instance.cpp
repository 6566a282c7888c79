// instance.cpp -- DIMACS reader and CSR construction (host side of the instance handle).
//
// Replaces, for the B200 engine, what the reference does on every call of run_sls_python:
//   read_to_string + DimacsParser::parse   rust/src/lib.rs:115-118
//   var_to_clauses construction            rust/src/lib.rs:186-199
#include "instance.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstring>

namespace slsb {

namespace {

// A cursor over the DIMACS text.  varisat's parser is a byte-level state machine; the accepted
// language restated here: whitespace-separated signed decimal literals, clauses closed by `0`
// (possibly spanning lines), lines whose first byte is 'c' are comments, one optional header
// "p cnf <vars> <clauses>" before the first clause.
struct Cursor {
    const char* p;
    const char* end;
    bool at_line_start = true;

    bool done() const { return p >= end; }
    void skip_line()
    {
        while (p < end && *p != '\n') ++p;
    }
};

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

bool parse_header(Cursor& cur, uint64_t& vars, uint64_t& clauses)
{
    // cur.p points at 'p'
    const char* q = cur.p + 1;
    auto skip_sp = [&] {
        while (q < cur.end && is_space(*q)) ++q;
    };
    auto read_num = [&](uint64_t& v) {
        skip_sp();
        if (q >= cur.end || *q < '0' || *q > '9') return false;
        v = 0;
        while (q < cur.end && *q >= '0' && *q <= '9') {
            v = v * 10 + uint64_t(*q - '0');
            if (v > 0xffffffffull) return false;
            ++q;
        }
        return true;
    };
    skip_sp();
    if (cur.end - q < 3 || std::memcmp(q, "cnf", 3) != 0) return false;
    q += 3;
    if (q < cur.end && !is_space(*q)) return false;
    if (!read_num(vars) || !read_num(clauses)) return false;
    skip_sp();
    if (q < cur.end && *q != '\n') return false;
    cur.p = q;
    return true;
}

}  // namespace

std::string parse_dimacs(const char* text, size_t len, HostInstance& out)
{
    Cursor cur{text, text + len};
    std::vector<uint32_t>& lits = out.lits;
    std::vector<uint32_t>& row_ptr = out.row_ptr;
    lits.clear();
    row_ptr.assign(1, 0u);
    lits.reserve(len / 4);
    bool have_header = false;
    uint64_t hdr_vars = 0, hdr_clauses = 0, max_var = 0;

    while (!cur.done()) {
        const char c = *cur.p;
        if (c == '\n') {
            cur.at_line_start = true;
            ++cur.p;
            continue;
        }
        if (is_space(c)) {
            ++cur.p;
            continue;
        }
        if (cur.at_line_start && c == 'c') {
            cur.skip_line();
            continue;
        }
        if (cur.at_line_start && c == 'p') {
            if (have_header || row_ptr.size() > 1 || lits.size() > row_ptr.back()) return "unexpected header";
            if (!parse_header(cur, hdr_vars, hdr_clauses)) return "invalid header";
            have_header = true;
            continue;
        }
        cur.at_line_start = false;
        bool neg = false;
        if (c == '-') {
            neg = true;
            ++cur.p;
            if (cur.done()) return "unexpected end of input after '-'";
        }
        if (*cur.p < '0' || *cur.p > '9') return "unexpected character";
        uint64_t v = 0;
        while (!cur.done() && *cur.p >= '0' && *cur.p <= '9') {
            v = v * 10 + uint64_t(*cur.p - '0');
            if (v > 0x7fffffffull) return "literal index too large";
            ++cur.p;
        }
        if (!cur.done() && !is_space(*cur.p) && *cur.p != '\n') return "unexpected character";
        if (v == 0) {
            if (neg) return "unexpected character";
            if (lits.size() > 0xfffffff0ull) return "formula too large";
            row_ptr.push_back(uint32_t(lits.size()));
        } else {
            max_var = std::max(max_var, v);
            lits.push_back(uint32_t((v - 1) << 1) | uint32_t(neg));
        }
    }
    if (lits.size() > row_ptr.back()) return "unterminated clause";
    const uint64_t m = row_ptr.size() - 1;
    if (have_header && max_var > hdr_vars) return "formula has more variables than the header declares";
    if (have_header && m != hdr_clauses) return "formula clause count differs from the header";
    if (m > 0xfffffff0ull) return "formula too large";
    out.n = uint32_t(have_header ? hdr_vars : max_var);
    out.m = uint32_t(m);
    return finalize_instance(out);
}

std::string finalize_instance(HostInstance& h)
{
    const uint32_t n = h.n, m = h.m;
    const uint64_t L = h.lits.size();
    if (h.row_ptr.size() != size_t(m) + 1 || h.row_ptr.back() != L) return "inconsistent CSR";

    // clause shape
    h.max_k = 0;
    uint32_t min_k = 0xffffffffu;
    for (uint32_t c = 0; c < m; c++) {
        const uint32_t k = h.row_ptr[c + 1] - h.row_ptr[c];
        h.max_k = std::max(h.max_k, k);
        min_k = std::min(min_k, k);
    }
    h.uniform_k = (m > 0 && min_k == h.max_k) ? h.max_k : 0;

    // var_to_clauses (lib.rs:190-195): one entry per literal occurrence, clause order preserved
    h.occ_ptr.assign(size_t(n) + 1, 0u);
    for (uint64_t i = 0; i < L; i++) {
        const uint32_t v = h.lits[i] >> 1;
        if (v >= n) return "literal out of range";
        h.occ_ptr[v + 1]++;
    }
    for (uint32_t v = 0; v < n; v++) h.occ_ptr[v + 1] += h.occ_ptr[v];
    h.occ.resize(L);
    {
        std::vector<uint32_t> fill(h.occ_ptr.begin(), h.occ_ptr.end() - 1);
        for (uint32_t c = 0; c < m; c++)
            for (uint32_t j = h.row_ptr[c]; j < h.row_ptr[c + 1]; j++) h.occ[fill[h.lits[j] >> 1]++] = c;
    }

    // Clause identity by content: the reference keeps violated clauses in a HashSet<&[Lit]>
    // (lib.rs:176-181), so clauses with the same literal sequence are one set element.
    h.canon.resize(m);
    h.has_duplicates = false;
    {
        // open-addressing table of clause ids keyed by a content hash (linear probing, <= 50 % load)
        size_t cap = 16;
        while (cap < size_t(m) * 2) cap <<= 1;
        const uint32_t EMPTY = 0xffffffffu;
        std::vector<uint32_t> slot(cap, EMPTY);
        for (uint32_t c = 0; c < m; c++) {
            const uint32_t* l = h.lits.data() + h.row_ptr[c];
            const uint32_t k = h.row_ptr[c + 1] - h.row_ptr[c];
            uint64_t hash = 1469598103934665603ull ^ k;
            for (uint32_t j = 0; j < k; j++) hash = (hash ^ l[j]) * 1099511628211ull;
            size_t i = size_t(hash ^ (hash >> 29)) & (cap - 1);
            uint32_t rep = c;
            while (slot[i] != EMPTY) {
                const uint32_t o = slot[i];
                if (h.row_ptr[o + 1] - h.row_ptr[o] == k && std::equal(l, l + k, h.lits.data() + h.row_ptr[o])) {
                    rep = o;
                    break;
                }
                i = (i + 1) & (cap - 1);
            }
            h.canon[c] = rep;
            if (rep == c)
                slot[i] = c;
            else
                h.has_duplicates = true;
        }
    }

    // k<=3 fast path needs: short clauses, distinct contents, no variable twice in a clause
    h.k3 = m > 0 && h.max_k <= 3 && min_k >= 1 && !h.has_duplicates;
    if (h.k3) {
        for (uint32_t c = 0; c < m && h.k3; c++) {
            const uint32_t* l = h.lits.data() + h.row_ptr[c];
            const uint32_t k = h.row_ptr[c + 1] - h.row_ptr[c];
            for (uint32_t a = 0; a < k; a++)
                for (uint32_t b = a + 1; b < k; b++)
                    if ((l[a] >> 1) == (l[b] >> 1)) h.k3 = false;
        }
    }
    return std::string();
}

uint64_t constraint_graph(uint32_t n_vars, uint32_t n_clauses, const int32_t* var_of, const int32_t* clause_of, uint64_t n_occ,
                          int64_t* major, int64_t* minor)
{
    // clause -> variables and variable -> clauses by counting sort (stable in occurrence order)
    std::vector<uint64_t> cptr(n_clauses + 1, 0), vptr(n_vars + 1, 0);
    for (uint64_t i = 0; i < n_occ; i++) {
        cptr[clause_of[i] + 1]++;
        vptr[var_of[i] + 1]++;
    }
    for (uint32_t c = 0; c < n_clauses; c++) cptr[c + 1] += cptr[c];
    for (uint32_t v = 0; v < n_vars; v++) vptr[v + 1] += vptr[v];
    std::vector<uint32_t> cvar(n_occ), vcl(n_occ);
    {
        std::vector<uint64_t> cf(cptr.begin(), cptr.end() - 1), vf(vptr.begin(), vptr.end() - 1);
        for (uint64_t i = 0; i < n_occ; i++) {
            cvar[cf[clause_of[i]]++] = (uint32_t)var_of[i];
            vcl[vf[var_of[i]]++] = (uint32_t)clause_of[i];
        }
    }
    // neighbours of clause c = union over its variables of their clauses, without c; rows are independent
    const unsigned T = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    std::vector<uint64_t> rowcnt(n_clauses + 1, 0);
    auto sweep = [&](bool fill) {
        std::atomic<uint32_t> next{0};
        auto work = [&] {
            std::vector<uint32_t> stamp(n_clauses, 0), nb;
            for (;;) {
                const uint32_t c0 = next.fetch_add(256);
                if (c0 >= n_clauses) break;
                for (uint32_t c = c0; c < std::min(n_clauses, c0 + 256); c++) {
                    nb.clear();
                    for (uint64_t a = cptr[c]; a < cptr[c + 1]; a++) {
                        const uint32_t v = cvar[a];
                        for (uint64_t b = vptr[v]; b < vptr[v + 1]; b++) {
                            const uint32_t d = vcl[b];
                            if (d != c && stamp[d] != c + 1) {
                                stamp[d] = c + 1;
                                nb.push_back(d);
                            }
                        }
                    }
                    if (!fill) {
                        rowcnt[c + 1] = nb.size();
                    } else {
                        std::sort(nb.begin(), nb.end());
                        uint64_t o = rowcnt[c];
                        for (uint32_t d : nb) {
                            major[o] = c;
                            minor[o++] = d;
                        }
                    }
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    };
    sweep(false);
    for (uint32_t c = 0; c < n_clauses; c++) rowcnt[c + 1] += rowcnt[c];
    if (major && minor) sweep(true);
    return rowcnt[n_clauses];
}

}  // namespace slsb
