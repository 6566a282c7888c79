// instance.h -- host-side CNF instance: DIMACS reader, clause CSR, occurrence CSR, HBM upload.
#pragma once
#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include "common.h"

namespace slsb {

struct HostInstance {
    uint32_t n = 0, m = 0;
    uint32_t max_k = 0, uniform_k = 0;
    bool has_duplicates = false;
    bool k3 = false;                  // all clauses have <= 3 literals, no duplicate clause contents
    std::vector<uint32_t> row_ptr;    // [m+1]
    std::vector<uint32_t> lits;       // [L] 2*var+neg
    std::vector<uint32_t> occ_ptr;    // [n+1]
    std::vector<uint32_t> occ;        // [L]
    std::vector<uint32_t> canon;      // [m]
    uint64_t num_lits() const { return lits.size(); }
};

// Parses DIMACS CNF with the acceptance rules of varisat 0.2.2's DimacsParser as the reference
// uses it (rust/src/lib.rs:115-118).  Returns an empty string on success, else the parse error.
std::string parse_dimacs(const char* text, size_t len, HostInstance& out);
// Fills occurrence lists, canonical clause ids and shape flags from row_ptr/lits.
std::string finalize_instance(HostInstance& h);

// Clause neighbourhood of the Lovasz-Local-Lemma loss (LCG.get_constraint_graph, python/src/sat_representations.py:668-715:
// sparsity pattern of C^T C minus the diagonal): for every clause, in increasing clause index, the other clauses that share
// a variable with it, in increasing index.  var_of / clause_of: one entry per literal occurrence.  Returns the edge count;
// fills major / minor (clause indices, NOT yet offset by 2n) when they are non-null.
uint64_t constraint_graph(uint32_t n_vars, uint32_t n_clauses, const int32_t* var_of, const int32_t* clause_of, uint64_t n_occ,
                          int64_t* major, int64_t* minor);

}  // namespace slsb

// The opaque handle of include/sls_b200.h
struct sls_instance {
    slsb::HostInstance host;
    slsb::DevInst dev{};      // device pointers valid on `device`
    int device = -1;
    void* d_blob = nullptr;   // one allocation holding every device array
    size_t blob_bytes = 0;
    std::atomic<int> solver_refs{0};  // live sls_solver objects pointing into d_blob (the image may not move meanwhile)
};

// Makes every instance of a batch resident on the current device (defined in api.cu): images of the missing ones
// are built by several host threads, one allocation and one host-to-device copy each.
int slsb_ensure_batch_on_device(sls_instance* const* insts, uint32_t n_inst);
