/*
 * sequnwinder_b200.h — C ABI of libsequnwinder_b200.so (sm_100a CUDA), the drop-in for
 * SeqUnwinder's two data-parallel hot paths. Plain pointers and sizes only; every entry point
 * names the reference interface it replaces. Citations are relative to
 * /root/reference/src/org/seqcode/projects/sequnwinder/. The Java-side JNI stubs that bind these
 * symbols are jni/sequnwinder_b200_jni.c and the Java classes beside it, described in INTEGRATION.md.
 *
 * Conventions: every function returns SQW_OK (0) or a negative SQW_ERR_* code; the message for
 * the calling thread's last failure is sqw_last_error(). There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with SQW_ERR_NO_DEVICE.
 * "host" variants take host buffers and do their own H2D/D2H copies (this is what the Java
 * host calls); "_dev" variants take device pointers and a cudaStream_t (as void*) and never
 * synchronise, so featurisation -> training can stay resident in HBM.
 */
#ifndef SEQUNWINDER_B200_H
#define SEQUNWINDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQW_OK 0
#define SQW_ERR_INVALID_ARG (-1)
#define SQW_ERR_NO_DEVICE (-2)
#define SQW_ERR_CUDA (-3)
#define SQW_ERR_BAD_BASE (-4)     /* seqcode-core seq2int's IllegalArgumentException */
#define SQW_ERR_UNSUPPORTED (-5)
#define SQW_ERR_NUMERIC (-6)      /* NaN/Inf met by the solver (LBFGS.java:792-814 guards) */
#define SQW_ERR_NCCL (-7)
#define SQW_ERR_STATE (-8)        /* call sequence violated (e.g. execute before set_problem) */

#define SQW_ABI_VERSION 2

/* Environment hooks read by the library (none is needed in production; defaults in brackets):
 *   tests      SQW_ADMM_EVAL_MODE=1|2|6   force a more general fp64 evaluation path
 *              SQW_ADMM_NO_PACKED=1       count matrices always through the fp64 kernels
 *              SQW_ADMM_PACKED_MODE=9|10|12  packed counts with two CTAs per SM (built, measured slower) /
 *                                         always through the column-chunked streamed kernels (what long
 *                                         rows and more than 8 classes use anyway) / always panel by
 *                                         panel (what short rows that do not fit shared memory use)
 *              SQW_ADMM_PANEL_ROWS=n      MODE 12: at most n rows per panel [as many as fit]
 *              SQW_ADMM_UNFUSED=1         iteration tail as separate kernels / two collectives
 *   profiling  SQW_ADMM_PROFILE=1         per-phase timers of the x-update kernel on stderr
 *              SQW_ADMM_TAIL_PROFILE=1    multi-GPU: events between the kernels of the iteration tail
 *   tuning     SQW_ADMM_CAPS=a,b,c,d,e    evaluation caps of the x-update phases [8,12,16,20,26]
 *              SQW_ADMM_GCAP=n            most CTAs a re-grouped block may get [~ sqrt(rows per block)]
 *              SQW_ADMM_SKEW_NS=n         MODE 9 probe: start delay of an SM's second CTA [0]
 *              SQW_ADMM_L2_PERSIST_MB=n   persisting-L2 window over the fp64 matrix [off]
 *              SQW_KMER_WARP64=1          the older 64-bit warp kernel of path 1
 * (SQW_LIB_PATH, read by the Python mirror only, selects another build of this library for A/B runs.) */
int sqw_abi_version(void);
const char *sqw_last_error(void);
/* number of visible CUDA devices (0 when there is none); never fails */
int sqw_device_count(void);
/* bind the calling thread (and handles created afterwards) to a device */
int sqw_set_device(int device);
/* the calling thread's current device (the CUDA current device is per host thread: a worker
 * thread that creates handles must first sqw_set_device() the device of the thread that owns the
 * data); -1 when there is no device */
int sqw_get_device(void);

/* ------------------------------------------------------------------------------------------
 * Path 1 — k-mer featurisation.
 * Replaces loaddata/MakeArff.java:314-377 (KmerCounter.printPeakSeqKmerRange) and its second
 * copy utils/ScoreSites.java:105-116.
 * ---------------------------------------------------------------------------------------- */

/* numK = sum_{k=kmin..kmax} 4^k (MakeArff.java:325-328); -1 on invalid range */
int64_t sqw_num_kmers(int kmin, int kmax);

/* Host entry point.
 *   bases   : concatenated ASCII sequences (upper or lower case), sum of lengths bytes
 *   offsets : n+1 start offsets into bases (sequence r = bases[offsets[r] .. offsets[r+1]))
 *   counts  : out, n x numK int32 row-major; column = base_k + min(fwd, revcomp) with
 *             base_k = sum_{j<k} 4^j for j>=kmin (MakeArff.java:357-367)
 *   has_n   : out, n bytes; 1 when the sequence contains 'N' (MakeArff.java:354-355: the caller
 *             must omit the row; its counts are returned as zeros)
 * Returns SQW_ERR_BAD_BASE if any sequence holds a character outside ACGTN. */
int sqw_kmer_count(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   int32_t *counts, uint8_t *has_n);

/* Device-resident entry point: all pointers are device pointers, launch is asynchronous on
 * `stream`. max_len = longest sequence (bases). d_status: one int32, OR-ed with 1 when a
 * character outside ACGTN was met and with 4 when a row is longer than max_len (that row is
 * truncated to max_len; the host variant returns SQW_ERR_INVALID_ARG). The caller zeroes it and
 * reads it back. */
int sqw_kmer_count_dev(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n,
                       int64_t max_len, int kmin, int kmax, int32_t *d_counts, uint8_t *d_has_n,
                       int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------
 * Path 2 — consensus-ADMM trainer.
 * Replaces framework/LOne.java (an framework/Optimizer.java:5-27 plug-in) together with its
 * inner solver framework/LBFGS.java + Mcsrch.java. Call order mirrors
 * framework/Classifier.java:422-449:
 *   create -> set_structure -> set_params -> [set_comm] -> set_problem -> execute -> destroy
 * (set_problem comes last because the row blocking depends on T; the Java shim keeps the
 * caller's order by uploading inside execute())
 * ---------------------------------------------------------------------------------------- */
typedef struct sqw_admm sqw_admm;

/* new LOne(...) (LOne.java:110-114) */
int sqw_admm_create(sqw_admm **out);
/* clearOptimizer() (LOne.java:259-263) */
void sqw_admm_destroy(sqw_admm *h);

/* LOne(xinit, sm_xinit, data) + setClsMembership + setInstanceWeights + setNumClasses +
 * setNumPredictors (LOne.java:95-98,110-114).
 *   data : n_rows x dim row-major fp64, column 0 == 1, already standardised
 *          (Classifier.java:337-381); y: class of each row (0..num_classes-1); w: weights.
 * Host variant copies to the device; _dev variant borrows device pointers (they must stay
 * valid until destroy). Multi-GPU: every rank passes the FULL problem to the host variant
 * (only the rows of its own ADMM blocks are uploaded), see sqw_admm_set_comm. */
int sqw_admm_set_problem(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                         const double *data, const int32_t *y, const double *w);
int sqw_admm_set_problem_dev(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                             const double *d_data, const int32_t *d_y, const double *d_w);

/* The same problem in PACKED-COUNT form. SeqUnwinder's feature matrix is a matrix of k-mer
 * counts (MakeArff.java:347-368: every feature is an integer in 0..L-k+1) that
 * Classifier.buildClassifier standardises column by column (Classifier.java:362-381). Instead of
 * the standardised doubles the caller hands over the counts and the column statistics:
 *   counts : n_rows x (dim-1) uint8 row-major = m_Data[i][1..nR] BEFORE Classifier.java:373-381
 *   mean, sd : dim doubles each = xMean / xSD of Classifier.java:362-371 (entry 0: 0 and 1)
 * and the trainer evaluates LOne.OptObject (LOne.java:936-1049) on
 *   X[i][j] = (counts[i][j-1] - mean[j]) / sd[j]   (the raw count where sd[j] == 0, :376)
 * without ever materialising X: logits = counts . (W/sd) - sum_j mean_j W_j/sd_j and
 * G = (R^T counts - mean (x) sum_i R_i) / sd, fp64 throughout. The matrix in HBM and shared memory
 * is 8x smaller than the fp64 one, which turns the evaluation from HBM-bound into fp64-DMMA-bound
 * (DESIGN.md 2.3). Relative to the fp64 entry point the results differ by rounding only (X is never
 * rounded to fp64 element by element; the summation order differs as it does between any two
 * group sizes).
 * _dev variant: the dense device count matrix of sqw_kmer_count_dev (n_rows x numK int32) plus
 * the column plan of sqw_prepare_plan_dev (keep_idx[num_kept], mean / sd indexed by ORIGINAL
 * column); dim = num_kept + 1. rows_local = 0: d_counts / d_y / d_w hold all n_rows rows;
 * rows_local = 1 (multi-GPU, sqw_admm_set_comm): they hold only the rows of this rank's blocks,
 * i.e. the rows i < floor(n_rows / T) * T with (i mod T) mod world == rank in ascending order of i,
 * so that every rank featurises its own sequences only (mean / sd stay the statistics of ALL rows).
 * The packed kernels cover every shape. Up to 8 classes and rows up to 768 columns (BASELINE
 * configs[0..1]): a CTA's rows of its block stay in shared memory for the whole launch where they
 * fit, and pass through the same buffer panel by panel where they do not (handles with a CTA
 * budget, blocks of many rows). Longer rows, up to 24 classes (configs[2..3]):
 * streamed as byte tiles, column chunk by column chunk. The
 * call always succeeds for valid counts. Counts above 255 -> SQW_ERR_UNSUPPORTED. Synchronises
 * `stream`. */
int sqw_admm_set_problem_counts(sqw_admm *h, int64_t n_rows, int32_t dim, int32_t num_classes,
                                const uint8_t *counts, const double *mean, const double *sd,
                                const int32_t *y, const double *w);
int sqw_admm_set_problem_counts_dev(sqw_admm *h, int64_t n_rows, int64_t numK, int32_t num_classes,
                                    const int32_t *d_counts, const int32_t *d_keep_idx,
                                    int32_t num_kept, const double *d_mean, const double *d_sd,
                                    const int32_t *d_y, const double *d_w, int32_t rows_local,
                                    void *stream);

/* setClassStructure (LOne.java:99; Classifier.java:633-736), CSR arrays with one entry per
 * design-file line in file order. Parents are meaningful for leaves only, children for
 * non-leaves only (Classifier.java:665,672). Leaves must carry indices 0..num_classes-1. */
int sqw_admm_set_structure(sqw_admm *h, int32_t num_nodes, int32_t num_layers,
                           const int32_t *node_index, const int32_t *is_leaf,
                           const int32_t *layer, const int32_t *parent_ptr,
                           const int32_t *parent_idx, const int32_t *child_ptr,
                           const int32_t *child_idx);

/* setRidge, setPho, set_numThreads, setADMMmaxItrs, setSeqUnwinderMaxIts, setDebugMode
 * (LOne.java:92-104). num_threads is T, the number of ADMM row blocks — it fixes the numerics
 * (row i -> block i % T, rows >= floor(N/T)*T dropped, LOne.java:337-357), not the number of
 * host threads or GPUs. */
int sqw_admm_set_params(sqw_admm *h, double ridge, double pho, int32_t num_threads,
                        int32_t admm_max_itr, int32_t nodes_max_itr, int32_t debug);

/* Regularisation sweep (BASELINE config 5: --r in {0.1 .. 100} on one training matrix; the
 * reference would start a new JVM run per value, SeqUnwinderConfig.java:389 takes a scalar):
 * changes lambda on a handle whose problem is already uploaded, so that successive
 * sqw_admm_execute calls re-use the resident training matrix and workspace. Each execute starts
 * from the x / sm_x it is given with fresh z, u, rho (a new LOne per run, LOne.java:80-85), so a
 * sweep returns exactly what independent runs return. */
int sqw_admm_set_ridge(sqw_admm *h, double ridge);

/* Concurrent sweep: caps the CTAs of this handle's cooperative x-update launches (0 = every SM).
 * Handles with budgets that add up to the SM count, each driven by its own host thread, solve
 * independent problems (the lambda values of a sweep, the folds of a cross-validation) side by
 * side on one GPU: a block's group of CTAs shrinks, and with it the share of a solver trip that is
 * L2 exchange between CTAs. Call before sqw_admm_set_problem. The group size enters the
 * summation order of the partial gradients, so results agree with a full-GPU run to rounding
 * (1e-16 per evaluation), not bit for bit. No reference counterpart. */
int sqw_admm_set_cta_budget(sqw_admm *h, int32_t max_ctas);

/* Multi-GPU (one process per GPU): blocks t with t % world == rank live on this rank; the
 * consensus sums (LOne.java:395-417) become one ncclAllReduce per ADMM iteration.
 * nccl_id is the 128-byte ncclUniqueId produced by sqw_nccl_unique_id on rank 0 and
 * distributed by the caller (e.g. torch.distributed broadcast). world == 1 needs no call. */
int sqw_nccl_unique_id(uint8_t id_out[128]);
int sqw_admm_set_comm(sqw_admm *h, int32_t rank, int32_t world, const uint8_t nccl_id[128]);

/* initUandZ() + execute() + getX() + getsmX() (LOne.java:80-85,116-204,107-108).
 * x: num_classes*dim, sm_x: num_nodes*dim; both hold the initial values on entry
 * (Classifier.java:400-420) and the trained values on return, host memory. */
int sqw_admm_execute(sqw_admm *h, double *x_inout, double *smx_inout);

/* What the reference only prints in -DEBUG mode (LOne.java:129-190,609-610,749-751), plus the
 * counters the benchmarks need. Valid after execute. */
typedef struct {
    int32_t outer_iters;      /* label-consensus iterations executed (LOne.java:119) */
    int32_t converged;        /* outer convergence flag (LOne.java:176-180) */
    int32_t ran_admm;         /* LOne.ranADMM at exit */
    int32_t total_admm_iters; /* sum over outer iterations of ADMM_currItr_value at exit */
    int64_t total_evals;      /* (objective, gradient) evaluations executed, all local blocks */
    double final_pho;
    double final_fold;
    double seconds_admm;      /* device time inside the ADMM loops (CUDA events), seconds */
    double seconds_total;     /* device time of the whole execute */
    int64_t launches;         /* kernels launched by execute */
    double seconds_eval;      /* device time of the x-update kernels only */
    int64_t eval_bytes;       /* algorithmic bytes read by those kernels */
    /* which evaluation path ran (tests assert that the path a benchmark uses is the path they
     * compared with the oracle): */
    int32_t eval_mode;        /* 4/5 fp64 single-pass TMA tiles, 6 fp64 column-chunked, 1/2 fp64
                                 streaming, 8/9 packed counts (uint8) resident in shared memory,
                                 12 the same panel by panel, 10/11 packed counts column-chunked */
    int32_t ctas_per_block;   /* CTAs of a block's group while every local block is active */
    int32_t ring_stages;      /* tile stages of the TMA ring in shared memory (8 / 12: 8-row tiles of
                                 the row buffer = one panel) */
    int32_t tiles_per_cta;    /* tiles a CTA walks per evaluation pass; > ring_stages means the
                                 ring streams (refill path), <= means its rows stay resident */
} sqw_admm_stats;
int sqw_admm_get_stats(const sqw_admm *h, sqw_admm_stats *out);

/* Per-ADMM-iteration log: rows = number of consensus updates executed; arrays may be NULL.
 * outer/itr: iteration indices; pho/fold: values the x-update of that iteration used;
 * primal/dual: sums of history_primal/history_dual rows (LOne.java:519-522);
 * nfev: [rows x num_local_blocks] f/g evaluations per local block. */
int sqw_admm_get_log(const sqw_admm *h, int32_t cap, int32_t *rows_out, int32_t *outer,
                     int32_t *itr, double *pho, double *fold, double *primal, double *dual,
                     int32_t *nfev, int32_t *admm_iters_per_outer);

/* ------------------------------------------------------------------------------------------
 * The steps either side of the trainer, device resident. All pointers are
 * device pointers; `stream` is a cudaStream_t.
 * ---------------------------------------------------------------------------------------- */

/* f-1a. Column statistics of the dense count matrix (n x numK int32): which columns survive
 * weka RemoveUseless (non-constant over the rows; framework/Classifier.java:305-308), and the
 * instance-weighted mean / SD of every column (Classifier.java:337-372; total_weight = sum of w).
 * d_keep_idx receives the num_kept surviving column indices in ascending order (capacity numK);
 * d_mean / d_sd are indexed by ORIGINAL column (numK each). Synchronises the stream once. */
int sqw_prepare_plan_dev(const int32_t *d_counts, int64_t n, int64_t numK, const double *d_w,
                         double total_weight, int32_t *d_keep_idx, int32_t *num_kept_out,
                         double *d_mean, double *d_sd, void *stream);
/* The same in two steps, for rows sharded over ranks (path 1 shards by rows): every
 * rank computes the per-column statistics of ITS rows (min, max, sum w x, sum w x^2; numK each),
 * the caller combines the four arrays across ranks (min / max / sum / sum, e.g. an all-reduce) and
 * every rank finishes with the global total_weight and row count. sqw_prepare_plan_dev is
 * stats + finish on one rank. */
int sqw_prepare_stats_dev(const int32_t *d_counts, int64_t n, int64_t numK, const double *d_w,
                          int32_t *d_min, int32_t *d_max, double *d_swx, double *d_swxx, void *stream);
int sqw_prepare_finish_dev(const int32_t *d_min, const int32_t *d_max, const double *d_swx,
                           const double *d_swxx, int64_t numK, double total_weight, int64_t n_total,
                           int32_t *d_keep_idx, int32_t *num_kept_out, double *d_mean, double *d_sd,
                           void *stream);

/* f-1b. The standardised training matrix the trainer takes: d_X is n x (num_kept+1) fp64
 * row-major, column 0 == 1 (Classifier.java:344,374-381). Replaces the tmpCounts.mat -> CSVLoader
 * -> ARFF text round trip of loaddata/MakeArff.java:255-288. Asynchronous. */
int sqw_prepare_apply_dev(const int32_t *d_counts, int64_t n, int64_t numK, const int32_t *d_keep_idx,
                          int32_t num_kept, const double *d_mean, const double *d_sd, double *d_X,
                          void *stream);

/* f-2. Batched Classifier.evaluateProbability (Classifier.java:244-287) on raw count rows:
 * d_par is m_Par, dim x num_classes row-major (de-standardised weights, dim = num_kept+1);
 * d_prob (n x num_classes, may be NULL) and d_argmax (n, first maximum, may be NULL) are
 * outputs. Asynchronous. */
int sqw_predict_dev(const double *d_par, int32_t dim, int32_t num_classes, const int32_t *d_counts,
                    int64_t n, int64_t numK, const int32_t *d_keep_idx, double *d_prob,
                    int32_t *d_argmax, void *stream);

/* f-4. Hills scan: motifs/Discrim.java:328-444 (findHills :328-380, findSeqlets :382-444), the
 * first consumer of the trained weights. Every window of length min_m..max_m (SeqUnwinderConfig
 * minscanlen / maxscanlen, :429-430) of every sequence without 'N' is scored with the k-mer
 * weights of one label (kmer_weights: numK doubles in count-matrix column order,
 * Classifier.getWeights / SeqUnwinderConfig.getKmerBaseInd :198-204); scores are accumulated in
 * the reference's order and are bit-identical to it. The per-sequence hill list follows the
 * reference's iterator logic (:410-427): mode 0 = findSeqlets (an equal-scoring overlapping hill
 * is kept), mode 1 = findHills (it is replaced); thresh = hillsthresh (:431).
 *   hill_start / hill_len / hill_score : out, n x cap; entries [r*cap, r*cap + counts[r]) are the
 *                                        hills of sequence r in list order (start is 0-based in
 *                                        the sequence; findHills adds the region start, :355)
 *   counts : out, n;  has_n : out, n (1 = sequence skipped, :337-338 / :389-390)
 * The caller sorts by score (Collections.sort, stable, descending; :231-249) and cuts the
 * substrings (:432-438). Returns SQW_ERR_BAD_BASE for a character outside ACGTN and
 * SQW_ERR_UNSUPPORTED if a sequence yields more than cap hills. */
int sqw_hills_scan(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   const double *kmer_weights, int min_m, int max_m, double thresh, int mode,
                   int32_t cap, int32_t *hill_start, int32_t *hill_len, double *hill_score,
                   int32_t *counts, uint8_t *has_n);

/* Device-resident form: all pointers are device pointers, asynchronous on `stream`. max_len =
 * longest sequence. d_status: one int32, OR-ed with 2 for a character outside ACGTN, with 4 when a
 * sequence overflowed cap and with 8 when a row is longer than max_len (truncated; caller zeroes it
 * and reads it back). */
int sqw_hills_scan_dev(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n, int64_t max_len,
                       int kmin, int kmax, const double *d_kmer_weights, int min_m, int max_m,
                       double thresh, int mode, int32_t cap, int32_t *d_hill_start,
                       int32_t *d_hill_len, double *d_hill_score, int32_t *d_counts, uint8_t *d_has_n,
                       int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks exposed for parity tests and micro-benchmarks (same kernels execute() uses).
 * ---------------------------------------------------------------------------------------- */

/* One (objective, gradient) evaluation of LOne.OptObject (LOne.java:936-1049) for every local
 * block at the block weights cx[t] (host, [T_local x C*dim]); z/u in compact edge layout are
 * taken as zero, sm_x as given, i.e. only the data term + augmented term with z=u=0.
 * f_out: [T_local], g_out: [T_local x C*dim]. Requires set_problem/structure/params. */
int sqw_admm_eval(sqw_admm *h, const double *cx, const double *smx, double pho, double *f_out,
                  double *g_out);

/* L-BFGS + Moré–Thuente exactly as driven by LOne.java:865-896, on the test function
 * f(x) = sum_i 100 (x_{2i+1} - x_{2i}^2)^2 + (1 - x_{2i})^2 (device-side solver state machine,
 * n even). Returns iflag; nfev_out/iters_out as in the oracle. */
int sqw_lbfgs_rosenbrock(int32_t n, double *x_inout, double eps, int32_t *nfev_out,
                         int32_t *iters_out, int32_t *iflag_out);

#ifdef __cplusplus
}
#endif
#endif /* SEQUNWINDER_B200_H */
