// Additive Schwarz / block-Jacobi preconditioner on the device (schwarz.cu).
#pragma once
#include "api_internal.h"

// AdditiveSchwarzPreconditioner (math-solvers/src/preconditioners/schwarz.rs:31-38) of the rows one rank owns.
// Every subdomain lies inside the rank's row block, so M^-1 acts on the rank's slab of a vector without communication;
// its local solve -- ILU(0) of the extracted block, which on a dense block is LU without pivoting (schwarz.rs:252-305)
// followed by the two substitutions of schwarz.rs:348-380 -- is stored as the explicit inverse of the block, i.e. the
// substitutions applied to the unit vectors, so that apply() is one batched block GEMV streaming the inverses once.
struct bemb200_precond {
    bemb200_ctx* ctx = nullptr;
    uint64_t n = 0;            // global dimension
    uint64_t r0 = 0, r1 = 0;   // rows of this rank
    uint32_t nsub_global = 0;
    uint32_t nsub = 0;         // subdomains owned by this rank
    uint32_t min_size = 0, max_size = 0;
    uint64_t total_entries_global = 0;
    uint64_t entries = 0;      // sum of the local subdomain sizes
    uint64_t inv_elems = 0;    // sum of size^2
    bool disjoint = true;      // every local row in exactly one subdomain: apply writes z directly (weights are 1)
    // device arrays
    uint64_t* sub_off = nullptr;   // [nsub + 1] offsets into idx / sol
    uint64_t* inv_off = nullptr;   // [nsub + 1] offsets into inv
    uint32_t* idx = nullptr;       // [entries] local row index (global - r0) of every subdomain entry
    bemb::cplx* inv = nullptr;     // [inv_elems] row-major inverse blocks
    bemb::cplx* sol = nullptr;     // [entries] local solutions (overlapping subdomains only)
    uint64_t* dof_ptr = nullptr;   // [nloc + 1] entries of every local row, ordered by subdomain (overlapping only)
    uint64_t* dof_pos = nullptr;   // [entries] positions into sol
    double* weight = nullptr;      // [nloc] 1 / (number of subdomains containing the row)  (schwarz.rs:92-108)
    uint32_t* cta_sub = nullptr;   // [ncta] subdomain of every apply CTA
    uint32_t* cta_row = nullptr;   // [ncta] first local row of the CTA inside its subdomain
    uint32_t ncta = 0;
    bemb::cplx* tmp = nullptr;     // [nloc] scratch slab (input of apply inside GMRES)
    double factor_ms = 0.0;
};

namespace bemb {
// z_loc = M^-1 r_loc on this rank's slab (both nloc long, device); stream-ordered, no synchronisation
cudaError_t schwarz_apply_local(const bemb200_precond* p, const cplx* r_loc, cplx* z_loc, cudaStream_t s);
int schwarz_apply_launches(const bemb200_precond* p);
// Z_loc[nloc][S] = M^-1 R_loc[nloc][S] for S interleaved right-hand sides (batched GMRES); disjoint subdomains only
cudaError_t schwarz_apply_block_local(const bemb200_precond* p, const cplx* R_loc, cplx* Z_loc, int S, cudaStream_t s);
}  // namespace bemb
