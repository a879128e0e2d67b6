// admm_types.cuh — device-side data layout of the consensus-ADMM trainer (path 2).
//
// Replaces the state of framework/LOne.java (fields :24-85, ADMMrunner :270-392). Differences in
// layout, none in meaning:
//   * every per-node weight vector is padded from dim = numPredictors+1 to dimp (multiple of 16
//     doubles) so rows are 128-byte aligned; padding is identically zero everywhere;
//   * z / zold / u are stored per (leaf,parent) EDGE ([E][dimp]) instead of the reference's dense
//     numNodes^2*dim arrays (LOne.java:80-85) whose non-edge entries are provably always zero;
//     a parentless leaf owns one self edge (parent = -1), as LOne.java:440,697,846 index it;
//   * the T row blocks of LOne.java:337-357 are materialised once, block-major.
#pragma once

#include <cstdint>

namespace sqw {

constexpr int kLbfgsM = 5;        // L-BFGS history length (LOne.java:868)
constexpr int kThreads = 352;     // threads per CTA of the x-update kernel: 8 DMMA + 2 softmax +
                                  // 1 producer warp; 65536 / 352 leaves 184 registers per thread
constexpr int kWarps = kThreads / 32;
// Per evaluation mode: threads per CTA, CTAs per SM the kernel is built for, DMMA warps per CTA.
// MODE 9 is the packed-count path with TWO CTAs per SM (192 threads: 4 DMMA + 2 helper warps, the
// same 168 registers per thread): while one CTA of an SM sits in the solver's exchange chain
// (latency bound, the DMMA pipe idle) the other one evaluates.
template <int MODE>
struct ModeCfg {
    static constexpr int NT = kThreads, MINB = 1, DW = 8;
};
template <>
struct ModeCfg<9> {
    static constexpr int NT = 192, MINB = 2, DW = 4;
};
constexpr int kProfSlots = 32;  // SQW_ADMM_PROFILE counters per CTA
constexpr int kNumDots = 32;      // scalar partials exchanged per evaluation (see admm_solver.cuh)
constexpr int kMaxEdgesPerLeaf = 32;
constexpr int kMaxFanIn = 64;

// framework/SeqUnwinderConfig.java:93-101
constexpr double kAlpha = 1.9;
constexpr double kAbsTol = 1e-2;
constexpr double kRelTol = 1e-2;
constexpr double kNodesTol = 1e-2;
constexpr double kPhoMax = 1000.0;

// Device-resident control block: the scalar fields of LOne / ADMMrunner that persist across
// kernel launches. Written by the consensus kernels, read by the x-update kernel and the host.
struct Ctrl {
    double pho;          // ADMM_pho
    double fold;         // ADMM_pho_fold
    double prev_primal;  // sum of history_primal[itr-1]
    double prev_dual;    // sum of history_dual[itr-1]
    int ran_admm;        // ranADMM
    int itr;             // ADMM_currItr_value
    int max_itr;         // ADMM_maxItr
    int done;            // this ADMMrunner.execute() has left its loop
    int fatal;           // NaN/Inf met (LBFGS.java:792-814 would have thrown)
    int lbfgs_errors;    // swallowed ExceptionWithIflag count (LOne.java:881-883)
    int log_rows;        // rows appended to the iteration log
    int log_cap;
    int outer;           // current outer iteration (for the log)
    int num_pass;        // outer convergence counters (LOne.java:167-174)
    int num_pass_relaxed;
    unsigned cons_count; // last-CTA-done counter of the consensus kernel
    long long total_evals;
};

struct AdmmDev {
    // ---- sizes ----
    int T;          // total ADMM blocks (= --threads), fixes the numerics
    int T_local;    // blocks owned by this rank
    int G;          // CTAs per block while every local block is active (= Gmax / T_local)
    int Gmax;       // CTAs of the x-update launch = slot stride of gpart / spart per block
    int gcap;       // most CTAs a re-grouped block may get in a later phase
    int nphase;     // x-update launches per ADMM iteration (re-grouping points)
    int nb;         // rows per block = floor(N / T)
    int dim;        // numPredictors + 1
    int dimp;       // padded row length (multiple of 16)
    int C;          // leaf classes
    int Cp;         // classes padded to a multiple of 8 (DMMA n-tile)
    int E;          // (leaf,parent) edges incl. self edges of parentless leaves
    int NN;         // numNodes
    int npad;       // C * dimp
    int upad;       // E * dimp
    int wstride;    // row stride of W in shared memory (bank-conflict free)
    int world;      // ranks
    int n_sm;       // SMs of the device (MODE 9: CTAs with blockIdx >= n_sm are an SM's second CTA)
    int skew_ns;    // MODE 9: start delay of the second-wave blocks (0 = none)
    double ridge;   // lambda

    // ---- training data, block-major ----
    const double *X;   // [T_local][nb][dimp]
    const int *y;      // [T_local][nb]
    const double *w;   // [T_local][nb]
    const int *block_ids;  // [T_local] global block index t of each local block
    // packed-count form of the same matrix (MODE 8, admm_eval_packed.cuh): counts as bytes with
    // column 0 == 1, and the column statistics that turn them into X = (n - mean) / sd
    const unsigned char *Xq;  // [T_local][nb][rsq]
    int rsq;                  // row stride in bytes (= dimp, a multiple of 16)
    const double *cmean;      // [dimp]  xMean (0 for the intercept, padding and sd == 0 columns)
    const double *cinvsd;     // [dimp]  1 / xSD (1 for the intercept and sd == 0 columns, 0 for padding)

    // ---- label hierarchy ----
    const int *edge_leaf;     // [E]
    const int *edge_par;      // [E], -1 for the self edge of a parentless leaf
    const int *leaf_edge_ptr; // [C+1] edges of leaf c are [ptr[c], ptr[c+1]) in parents-list order
    const int *edge_sum_order;// [E] edge ids sorted by leaf*NN+pid (LOne.java:519-522 sum order)

    // ---- consensus state (replicated on every rank) ----
    double *xbar;  // [npad]      LOne.x
    double *ubar;  // [upad]      LOne.u
    double *z;     // [upad]
    double *zold;  // [upad]
    double *smx;   // [NN*dimp]   LOne.sm_x
    double *smx_old; // [NN*dimp] snapshot for the outer convergence test

    // ---- per local block ----
    double *xt;    // [T_local][npad]  t_b_x / t_x
    double *ut;    // [T_local][upad]  t_b_u / t_u

    // ---- L-BFGS workspace per local block ----
    double *xcur;  // [T_local][npad] trial point
    double *g;     // [T_local][npad]
    double *gold;  // [T_local][npad] w[0..n) of LBFGS.java
    double *d;     // [T_local][npad] search direction
    double *wa;    // [T_local][npad] line-search origin (diag/wa in the reference)
    double *S;     // [T_local][m][npad]
    double *Y;     // [T_local][m][npad]
    double *R;     // [T_local][nb][Cp]   w_i * (softmax - onehot), between the two passes
    double *gpart; // [T_local][Gmax][npad]  per-CTA partial gradients
    double *spart; // [T_local][Gmax][kNumDots] per-CTA partial scalars
    unsigned *bar; // [nphase][T_local][32] group barrier counters (one per 128-byte line)
    struct BlockState *bstate; // [T_local] solver state carried between the phases of an iteration
    int *nfev;     // [T_local] evaluations done by the last x-update
    long long *prof; // optional [T_local*G][16] phase cycle counters (SQW_ADMM_PROFILE=1)

    // ---- consensus exchange buffers ----
    double *sum1;  // [npad + upad]  sum over blocks of x_t | u_t   (allreduced when world > 1)
    double *sum2;  // [E]            sum over blocks of ||x_t - z - sm_x[par]||^2 per edge
    double *edge_stats; // [E][4]    znorm^2 (or unorm^2 for self edges), dual^2, -, -
    double *leaf_xnorm; // [C]

    // ---- log ----
    Ctrl *ctrl;
    int *log_outer, *log_itr, *log_nfev;     // [log_cap], [log_cap], [log_cap][T_local]
    double *log_pho, *log_fold, *log_primal, *log_dual;
};

}  // namespace sqw
