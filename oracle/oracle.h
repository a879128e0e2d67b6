/*
 * oracle.h — CPU restatement of SeqUnwinder's two hot paths (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the parity checker, not the product. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The shipped path
 * (sequnwinder_b200/) never links, imports or falls back to anything here.
 *
 * PARITY UNPINNED: the reference (/root/reference, Java 8 + Weka + seqcode-core) ships no tests,
 * golden vectors or fixtures for these paths and cannot be compiled or run in this image
 * (no JDK, dependencies not vendored). Every function below is therefore a line-traceable
 * restatement of the cited reference lines, pinned only by hand-derived known-answer vectors
 * (tests/golden/). All file:line citations are relative to
 * /root/reference/src/org/seqcode/projects/sequnwinder/.
 *
 * Compile with -ffp-contract=off: Java never fuses a*b+c, so neither may this code.
 */
#ifndef SQW_ORACLE_H
#define SQW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- path 1: loaddata/MakeArff.java:323-377 (KmerCounter.printPeakSeqKmerRange) ---- */

/* numK = sum_{k=kmin..kmax} 4^k (MakeArff.java:325-328). */
int64_t orc_num_kmers(int kmin, int kmax);

/* counts: n x numK int32 row-major, zero-filled for rows containing 'N' (has_n[r]=1; the
 * reference omits such rows from tmpCounts.mat, MakeArff.java:354-355).
 * Returns 0, or -1 on a character outside ACGTN (seqcode-core seq2int throws). */
int orc_kmer_count(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   int32_t *counts, uint8_t *has_n);

/* ---- hills scan: motifs/Discrim.java:328-444 (findHills / findSeqlets) ---- */

/* For every sequence without 'N': every window of length min_m..max_m is scored with the k-mer
 * weights (kmer_weights: numK doubles in count-matrix column order) and the per-sequence list of
 * non-overlapping "hills" is built with the reference's iterator logic. mode 0 = findSeqlets
 * (ties keep both hills), 1 = findHills (ties replace). Hills of sequence r, in list order:
 * hill_start/len/score[r*cap + 0 .. counts[r]). Returns 0; -1 on a character outside ACGTN;
 * -4 if a sequence produced more than cap hills. */
int orc_hills_scan(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   const double *kmer_weights, int min_m, int max_m, double thresh, int mode,
                   int32_t cap, int32_t *hill_start, int32_t *hill_len, double *hill_score,
                   int32_t *counts, uint8_t *has_n);

/* ---- class structure: framework/Classifier.java:633-736 (ClassRelationStructure) ---- */
typedef struct {
    int32_t num_nodes;          /* allNodes.size() */
    int32_t num_layers;         /* -NL */
    /* one entry per design-file line, in file order */
    const int32_t *node_index;  /* pieces[0] */
    const int32_t *is_leaf;     /* pieces[2]==1 */
    const int32_t *layer;       /* pieces[5] */
    const int32_t *parent_ptr;  /* CSR over file-order entries, num_nodes+1 */
    const int32_t *parent_idx;  /* only recorded for leaves (Classifier.java:665) */
    const int32_t *child_ptr;   /* CSR, num_nodes+1 */
    const int32_t *child_idx;   /* only recorded for non-leaves (Classifier.java:672) */
} orc_structure;

/* ---- path 2: framework/LOne.java + LBFGS.java + Mcsrch.java ---- */
typedef struct {
    int64_t n_rows;             /* data.length */
    int32_t dim;                /* numPredictors+1, column 0 == 1 */
    int32_t num_classes;        /* C */
    const double *data;         /* n_rows x dim row-major, already standardised */
    const int32_t *cls;         /* Y */
    const double *weights;      /* instance weights */
    orc_structure st;
    double ridge;               /* -R   */
    double pho;                 /* -PHO */
    int32_t num_threads;        /* -threads (T = number of ADMM blocks) */
    int32_t admm_max_itr;       /* -A   */
    int32_t nodes_max_itr;      /* -S   */
    int32_t host_threads;       /* how many pthreads execute the T block updates (1 = serial) */
    int32_t java_sum_order;     /* 1: sum thread buffers in java.util.HashMap key order */
} orc_problem;

typedef struct {
    int32_t outer_iters;        /* outer (label) iterations executed */
    int32_t converged;          /* outer loop convergence flag */
    int32_t log_cap;            /* capacity (rows) of the arrays below */
    int32_t log_rows;           /* rows written: one per executed ADMM iteration */
    int32_t *log_outer;         /* [log_cap] outer iteration index */
    int32_t *log_itr;           /* [log_cap] ADMM iteration index k */
    double *log_pho;            /* [log_cap] rho used by the workers in iteration k */
    double *log_fold;           /* [log_cap] rho_fold used by the workers in iteration k */
    double *log_primal;         /* [log_cap] sum of history_primal[k] */
    double *log_dual;           /* [log_cap] sum of history_dual[k] */
    int32_t *log_nfev;          /* [log_cap x num_threads] f/g evaluations per block */
    int32_t *admm_iters;        /* [nodes_max_itr] ADMM_currItr_value at exit, per outer it */
    double final_pho;
    double final_fold;
    int32_t ran_admm;
    int64_t total_evals;        /* total (f,g) evaluations over all blocks */
    double seconds_admm;        /* wall seconds inside ADMMrunner.execute equivalents */
    int32_t lbfgs_errors;       /* count of swallowed ExceptionWithIflag (LOne.java:881-883) */
} orc_trace;

int orc_lone_execute(const orc_problem *p, double *x_inout, double *smx_inout, orc_trace *tr);

/* iteration order of a java.util.HashMap<String,...> holding keys "Thread0".."Thread{T-1}"
 * inserted in that order (LOne.java:399,411,427). order[] receives T thread ids. */
void orc_java_thread_order(int T, int32_t *order);

/* single (f,g) evaluation of LOne.OptObject (LOne.java:936-1049) on one block: exposed so the
 * GPU evaluation kernel can be checked call by call. z/u/smx use the dense reference layout. */
double orc_objective(const orc_problem *p, const double *blk_data, const int32_t *blk_cls,
                     const double *blk_w, int64_t nb, const double *cx, const double *smx,
                     const double *z, const double *tu, double pho);
void orc_gradient(const orc_problem *p, const double *blk_data, const int32_t *blk_cls,
                  const double *blk_w, int64_t nb, const double *cx, const double *smx,
                  const double *z, const double *tu, double pho, double *grad);

/* ---- L-BFGS as used by the worker (LOne.java:865-896): minimise a user callback ---- */
typedef double (*orc_fg_cb)(void *user, const double *x, double *g);
/* returns final iflag (0 ok, -1 line search failed); *nfev_out = number of f/g evaluations */
int orc_lbfgs_minimize(int n, int m, double *x, double eps, double xtol, orc_fg_cb cb,
                       void *user, int *nfev_out, int *iters_out);

/* ---- caller side: framework/Classifier.java:337-420, 452-477, 267-287 ---- */
/* data: n x dim with column 0 == 1, standardised in place. mean/sd: [dim]. */
int orc_standardize(double *data, int64_t n, int32_t dim, const double *weights,
                    double *mean, double *sd);
void orc_init_params(const int32_t *cls, int64_t n, int32_t dim, int32_t C,
                     const orc_structure *st, double *x, double *smx);
/* par: [dim x ncols] (m_Par layout, Classifier.java:334), from v: [ncols x dim] */
void orc_destandardize(const double *v, int32_t ncols, int32_t dim, const double *mean,
                       const double *sd, double *par);
/* prob: n x C; data_raw: n x dim with column 0 == 1 (Classifier.java:267-287) */
void orc_predict(const double *par, int32_t dim, int32_t C, const double *data_raw, int64_t n,
                 double *prob);

#ifdef __cplusplus
}
#endif
#endif
