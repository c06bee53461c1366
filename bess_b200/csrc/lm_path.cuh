// Resident PDAS path for the gaussian family (lm_path.cu): declarations shared with engine.cu.
#pragma once
#include <string>

#include "kernels.cuh"

namespace bess {

constexpr int LP_NT = 512;        // threads per CTA
constexpr int LP_CAP = 512;       // capacity of a chain's candidate list (== LP_NT: one rank-counting thread per candidate)
constexpr int LP_MAXSTEP = 64;    // path steps (sparsity level, ridge level) per launch
constexpr int LP_WPMAX = 128;     // widest column slice of a sweeper CTA, in column pairs
constexpr int LP_RPMAX = 32;      // most row phases of a sweeper CTA (bounds its serial partial reduction)
constexpr int LP_TRACE = 128;     // phases recorded per CTA in LpDesc::trace
constexpr int LP_KMAX = 64;       // largest support the in-kernel solver handles
constexpr int LP_NDBG = 32 + 4 * MAXC;  // 0..15 owner 0 phases, 16..31 sweeper 0 phases, then 4 words per chain owner

struct LpCand {
    double v;
    int idx;
    int pad;
};

// words of LpDesc::sync.  B1 / B2 have one counter per chain group.
enum { LP_SYNC_B1 = 0, LP_SYNC_B2 = 2, LP_SYNC_NCOMPLETE = 4, LP_SYNC_TERM = 5, LP_SYNC_ITERS = 6, LP_SYNC_ABORT = 7,
       LP_SYNC_FALLBACKS = 8, LP_SYNC_MERGED = 9, LP_SYNC_RANKDEF = 10, LP_SYNC_WORDS = 16 };

// One launch = `nsteps` path steps for the chains of one batch.  Passed by value.
struct LpDesc {
    int nsteps, nch, n_always, nsweep;
    int ns;          // column slots of an owner CTA (>= kcap)
    int hist_rows;   // rows of A_list an owner keeps (max_iter + 2)
    int T[LP_MAXSTEP];
    double lam[LP_MAXSTEP];
    int chain[MAXC];
    // chain groups: the sweepers alternate between the groups, so that the owners of one group work on their active sets
    // while the other group's sacrifices are swept (one group: plain alternation of sweep and owner phases)
    int ng;                // 1 or 2
    int gcount[2];         // chains per group
    int gchain[2][MAXC];   // chain id of slot f of group g
    int ogroup[MAXC];      // group / slot of the chain at position i of chain[]
    int oslot[MAXC];
    int fh;                // chain slots of a group's residual matrix (the sweeper's FT)
    double *Rg[2];         // [npad][fh] residual vectors of each group
    const int *always;     // [n_always] pinned columns (always_select)
    const double *y;       // [n] response after normalisation
    unsigned *sync;        // [LP_SYNC_WORDS], zeroed before every launch
    LpCand *cand;          // [MAXC][LP_CAP] candidates of the running iteration, by chain id
    int *ncand;            // [MAXC]
    double *pub;           // [MAXC][2] (candidate threshold, ridge level) an owner publishes for the sweepers
    double *tau;           // [MAXC] candidate threshold carried from launch to launch (a hint: never affects the result)
    int *res_i;            // [nsteps][nch][2 + kcap]: l, boundary ties, support
    double *res_d;         // [nsteps][nch][2 + kcap]: loss over all rows, loss over the chain's held-out rows, coefficients
    unsigned long long *dbg;  // [LP_NDBG] phase timers (clock ticks) of owner 0 and sweeper 0
    unsigned long long *trace;  // optional [MAXC + 1][LP_TRACE][2] globaltimer stamps (start, end) of every owner phase / sweeper-0 step
};

// Can the resident kernel run this problem (family, shapes, shared-memory budget)?  `why` <- reason when not.
bool lm_path_eligible(const Dev &d, int max_iter, int sm_count, std::string *why);
size_t lm_path_smem_bytes(const Dev &d, int max_iter, int sm_count);
int lm_path_slots(const Dev &d, int max_iter);
void launch_lm_path(const Dev &d, const LpDesc &desc, int sm_count, int max_iter, cudaStream_t st);
// chain slots per group for a batch of nch chains split into ng groups (a supported sweeper instantiation)
int lm_path_fh(int nch, int ng);
int lm_path_groups(int nch);

}  // namespace bess
