// admm_kernels.cuh — the kernels of one consensus-ADMM iteration (path 2).
//
// Replaces framework/LOne.java: ADMMrun.run (:811-914, worker: u-update + L-BFGS x-update),
// ADMMrunner.execute's consensus step (:670-716), updateResiduals / norms (:418-510),
// updatePhoAndFold (:511-535), hasADMMConverged (:544-581), updateInternalNodes (:208-241) and
// the outer convergence test (:149-180).
//
// Per ADMM iteration the host enqueues, with no synchronisation:
//     admm_xupdate_kernel   cooperative, T_local groups of G CTAs, one group per ADMM block
//     admm_pack_kernel      sum over local blocks of x_t | u_t        [ncclAllReduce if world>1]
//     admm_consensus_kernel one CTA per (leaf,parent) edge: x-bar, u-bar, z-update, norms,
//                           dual residual, local part of the primal residual
//                                                                     [ncclAllReduce if world>1]
//     admm_decide_kernel    residual bookkeeping, rho adaptation, convergence -> Ctrl.done
// Once Ctrl.done is set the kernels of later (speculatively enqueued) iterations return at once.
//
// Reordering relative to the reference, results unchanged: hasADMMConverged(k-1) is evaluated
// right after iteration k-1's residuals instead of after iteration k's x-update, because it
// depends on nothing that x-update produces; the x-update whose result the reference computes and
// then discards on convergence (LOne.java:649-665,753-758) is therefore never run. The rho update
// of iteration k (:614-615) still precedes that test, as in the reference.
#pragma once

#include "admm_eval_tma.cuh"
#include "admm_eval_packed.cuh"
#include "admm_eval_packed_chunk.cuh"
#include "admm_solver.cuh"

namespace sqw {

__device__ __forceinline__ int ld_volatile_int(const int *p)
{
    return *reinterpret_cast<const volatile int *>(p);
}

// ------------------------------------------------------------------ x-update
// Evaluation modes of the x-update kernel:
//   MODE 4  fused single pass over TMA-staged 8-row tiles, warp specialised (TileEvalWS):
//           rows up to 672 columns, i.e. the k = 4..5 feature sets of the headline configs;
//   MODE 6  two passes over column-chunked TMA-staged tiles (TileEvalChunk): any row length;
//   MODE 1  two streaming passes from global memory, W in shared memory (admm_eval.cuh);
//   MODE 2  the same with W read from global memory (row lengths beyond shared memory).
//   MODE 8  the CTA's rows as byte counts resident in shared memory, three bulk phases per
//           evaluation (PackedEval): count matrices with rows up to 768 columns and C <= 8.
//   MODE 10 packed counts streamed as byte tiles (cp.async), column-chunked, two passes: count
//           matrices of any row length, up to 24 classes, blocks of any size (cfg-3, cfg-4).
//   MODE 11 MODE 10 for C = 8 NCT + 1 classes (2^k labels + one outgroup class, cfg-4): NCT class
//           tiles of DMMAs and the last class as DFMAs on the operands the DMMA path converted.
//   MODE 12 MODE 8 for CTAs that own more rows than shared memory holds (CTA budgets, blocks of
//           many short rows): the rows pass through the buffer panel by panel, one fetch per
//           evaluation, the gradient accumulators in registers across the panels.
//   MODE 9  MODE 8 with two CTAs of 192 threads per SM (4 DMMA + 2 helper warps each): the CTAs of
//           an SM belong to different ADMM blocks, so one evaluates while the other sits in its
//           solver's exchange chain.
template <int NCT, int MODE>
struct StreamEval {
    const AdmmDev &P;
    int blk, cta, G;
    double *smW, *red;
    __device__ void init(unsigned char *smem, const TileSmemLayout &, double *red_)
    {
        smW = reinterpret_cast<double *>(smem);
        red = red_;
    }
    __device__ double eval(const double *x)
    {
        return eval_data_term<NCT, MODE == 1>(P, blk, cta, G, x, smW, red);
    }
    __device__ void set_profile(long long *) {}
    __device__ void drain() {}
};

template <int NCT, int MODE>
struct EvalSelect {
    typedef StreamEval<NCT, MODE> type;
};
template <int NCT>
struct EvalSelect<NCT, 4> {
    typedef TileEvalWS<NCT, 24> type;
};
template <int NCT>
struct EvalSelect<NCT, 6> {   // column-chunked TMA tiles: any row length
    typedef TileEvalChunk<NCT> type;
};
template <int NCT>
struct EvalSelect<NCT, 5> {   // MODE 4 with 21 k-steps per DMMA warp (rows <= 672 columns)
    typedef TileEvalWS<NCT, 21> type;
};
template <int NCT>
struct EvalSelect<NCT, 8> {   // packed counts resident in shared memory (admm_eval_packed.cuh)
    typedef PackedEval<NCT, 8, kThreads> type;
};
template <int NCT>
struct EvalSelect<NCT, 10> {  // packed counts streamed from HBM, column-chunked (admm_eval_packed_chunk.cuh)
    typedef PackedChunkEval<NCT> type;
};
template <int NCT>
struct EvalSelect<NCT, 11> {  // the same for C = 8 NCT + 1 classes: the last class rides along as DFMAs
    typedef PackedChunkEval<NCT, true> type;
};
template <int NCT>
struct EvalSelect<NCT, 12> {  // packed counts streamed PANEL by panel through the resident form's buffer
    typedef PackedEval<NCT, 8, kThreads, true> type;
};
template <int NCT>
struct EvalSelect<NCT, 9> {   // the same with two CTAs of 192 threads per SM
    typedef PackedEval<NCT, ModeCfg<9>::DW, ModeCfg<9>::NT> type;
};

// Shared-memory copy of the (tiny) leaf -> edge -> parent tables: the augmented term of every
// coordinate walks them, and from global memory each hop is a dependent L1/L2 round trip.
struct StructCache {
    static constexpr int kMaxLeaves = 64, kMaxEdges = 256;
    int lep[kMaxLeaves + 1];
    int epar[kMaxEdges];
    const int *lep_p, *epar_p;
    // all threads; followed by a __syncthreads before first use (the evaluator's init has one)
    __device__ void init(const AdmmDev &P)
    {
        const bool fits = P.C <= kMaxLeaves && P.E <= kMaxEdges;
        if (fits) {
            for (int i = threadIdx.x; i <= P.C; i += blockDim.x)
                lep[i] = P.leaf_edge_ptr[i];
            for (int i = threadIdx.x; i < P.E; i += blockDim.x)
                epar[i] = P.edge_par[i];
        }
        if (threadIdx.x == 0) {
            lep_p = fits ? lep : P.leaf_edge_ptr;
            epar_p = fits ? epar : P.edge_par;
        }
    }
};

template <class Eval>
struct BlockProblem {
    const AdmmDev &P;
    int blk, cta, G;
    double pho;
    const double *ut;
    Eval &ev;
    const int *lep;    // leaf_edge_ptr / edge_par: shared-memory copies when they fit (StructCache)
    const int *epar;

    __device__ double eval_partial(const double *x) { return ev.eval(x); }

    // data term of coordinate k reduced over the group's phase-E partials (fixed order), in two
    // halves so that the caller can put other loads behind the first batch: data_issue starts the
    // loads of the first kBatch partials, data_finish sums them and any further batches.
    // 24 covers a group of 18 (8 blocks on 148 SMs) in one L2 round trip.
    static constexpr int kBatch = 24;
    __device__ void data_issue(int k, double (&v)[kBatch])
    {
        const double *gp = P.gpart + (size_t)blk * P.Gmax * P.npad + k;
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
            v[j] = (j < G) ? __ldcg(gp + (size_t)j * P.npad) : 0.0;
    }
    __device__ double data_finish(int k, double (&v)[kBatch])
    {
        const double *gp = P.gpart + (size_t)blk * P.Gmax * P.npad + k;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
            s += v[j];
        for (int c0 = kBatch; c0 < G; c0 += kBatch) {
#pragma unroll
            for (int j = 0; j < kBatch; ++j)
                v[j] = (c0 + j < G) ? __ldcg(gp + (size_t)(c0 + j) * P.npad) : 0.0;
#pragma unroll
            for (int j = 0; j < kBatch; ++j)
                s += v[j];
        }
        return s;
    }
    __device__ double data_term(int k)
    {
        double v[kBatch];
        data_issue(k, v);
        return data_finish(k, v);
    }
    // The same sum for a group of at most kSmallBatch CTAs (a handle with a CTA budget): same
    // order, a third of the registers - the solver's loop over a thread's further coordinates
    // keeps all of a coordinate's loads in flight instead of spilling the first ones to make room.
    static constexpr int kSmallBatch = 8;
    __device__ void data_issue_small(int k, double (&v)[kSmallBatch])
    {
        const double *gp = P.gpart + (size_t)blk * P.Gmax * P.npad + k;
#pragma unroll
        for (int j = 0; j < kSmallBatch; ++j)
            v[j] = (j < G) ? __ldcg(gp + (size_t)j * P.npad) : 0.0;
    }
    __device__ double data_finish_small(int, double (&v)[kSmallBatch])
    {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kSmallBatch; ++j)
            s += v[j];
        return s;
    }

    // augmented Lagrangian term rho*(x - sm_x[par] - z + u_t) of coordinate k for w >= 1
    // (LOne.java:982-998); faug accumulates rho/2 * (...)^2 (LOne.java:1031-1047). xk = x[k].
    // Reads only x and data that is constant once the u-update is complete.
    __device__ double aug_term(int k, double xk, double &faug)
    {
        double s = 0.0;
        const int leaf = k / P.dimp, w = k - leaf * P.dimp;
        if (w >= 1 && w < P.dim) {
            const int e0 = lep[leaf], e1 = lep[leaf + 1];
            // two edges per trip with every load issued before the first use: the term costs one
            // L2 round trip for the usual one or two parents per leaf (same arithmetic order)
            for (int e = e0; e < e1; e += 2) {
                const bool two = e + 1 < e1;
                const int par0 = epar[e], par1 = two ? epar[e + 1] : -1;
                const size_t o0 = (size_t)e * P.dimp + w, o1 = two ? o0 + P.dimp : o0;
                const double sx0 = (par0 >= 0) ? __ldg(P.smx + (size_t)par0 * P.dimp + w) : 0.0;
                const double sx1 = (par1 >= 0) ? __ldg(P.smx + (size_t)par1 * P.dimp + w) : 0.0;
                const double z0 = __ldg(P.z + o0), z1 = __ldg(P.z + o1);
                const double u0 = __ldcg(ut + o0), u1 = __ldcg(ut + o1);
                double v = xk;
                if (par0 >= 0)
                    v -= sx0;
                v = (v - z0) + u0;
                s += pho * v;
                faug += (pho / 2) * v * v;
                if (two) {
                    v = xk;
                    if (par1 >= 0)
                        v -= sx1;
                    v = (v - z1) + u1;
                    s += pho * v;
                    faug += (pho / 2) * v * v;
                }
            }
        }
        return s;
    }

    // The same term from inputs cached on chip (admm_solver.cuh, SQW_AUG_CACHE): aug_load fetches
    // (sm_x[par], z, u_t) of the coordinate's edges once per launch, false if it has more than two;
    // a parentless edge stores sm_x = +0.0 (x - 0.0 == x exactly). Same operations, same order.
    __device__ bool aug_load(int k, double (&a)[6], int &ne)
    {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            a[j] = 0.0;
        ne = 0;
        const int leaf = k / P.dimp, w = k - leaf * P.dimp;
        if (w >= 1 && w < P.dim) {
            const int e0 = lep[leaf], e1 = lep[leaf + 1];
            if (e1 - e0 > 2)
                return false;
            ne = e1 - e0;
            // (constant indices: a[] stays in registers; all loads are issued before any use)
            const size_t o = (size_t)e0 * P.dimp + w;
            const int par0 = (ne >= 1) ? epar[e0] : -1, par1 = (ne == 2) ? epar[e0 + 1] : -1;
            if (par0 >= 0)
                a[0] = __ldg(P.smx + (size_t)par0 * P.dimp + w);
            if (par1 >= 0)
                a[3] = __ldg(P.smx + (size_t)par1 * P.dimp + w);
            if (ne >= 1) {
                a[1] = __ldg(P.z + o);
                a[2] = __ldcg(ut + o);
            }
            if (ne == 2) {
                a[4] = __ldg(P.z + o + P.dimp);
                a[5] = __ldcg(ut + o + P.dimp);
            }
        }
        return true;
    }
    __device__ double aug_term_cached(double xk, const double (&a)[6], int ne, double &faug)
    {
        double s = 0.0;
        if (ne >= 1) {
            double v = xk - a[0];
            v = (v - a[1]) + a[2];
            s += pho * v;
            faug += (pho / 2) * v * v;
        }
        if (ne == 2) {
            double v = xk - a[3];
            v = (v - a[4]) + a[5];
            s += pho * v;
            faug += (pho / 2) * v * v;
        }
        return s;
    }

    __device__ double grad_coord(int k, double xk, double &faug)
    {
        const double a = aug_term(k, xk, faug);
        return data_term(k) + a;
    }
};

// One "phase" of the x-update of an ADMM iteration. The host enqueues P.nphase of these per
// iteration with growing evaluation caps: a block whose L-BFGS solve returns before the cap is
// marked finished, and the next phase spreads ALL CTAs of the launch over the blocks that are still
// iterating (the slowest block of an iteration needs 1.6x the mean number of evaluations at
// cfg-2; with a fixed partition 39 % of the SM time was idle). The re-grouping points depend only
// on evaluation counts, never on timing, so results stay reproducible run to run.
template <int NCT, int MODE>
__global__ void __launch_bounds__(ModeCfg<MODE>::NT, ModeCfg<MODE>::MINB)
admm_xupdate_kernel(AdmmDev P, TileSmemLayout L, int phase, int eval_cap)
{
    constexpr int kThreads = ModeCfg<MODE>::NT;   // (shadows the namespace constant: this kernel's CTA size)
    constexpr int kCap = (MODE == 9) ? kPackedSliceCap2 : kSliceCap;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ SolveSmem sm;
    __shared__ int s_blk, s_G;
    if (ld_volatile_int(&P.ctrl->done) || ld_volatile_int(&P.ctrl->fatal))
        return;
    // SQW_ADMM_PROFILE: thread 0 accumulates its time stamps in shared memory (a global
    // read-modify-write per stamp would put an L2 round trip into the next interval) and adds them
    // to the per-phase global counters once, at the end of the launch
    __shared__ long long s_prof[kProfSlots];
    if (P.prof) {
        P.prof += ((size_t)phase * gridDim.x + blockIdx.x) * kProfSlots;
        if (threadIdx.x < kProfSlots)
            s_prof[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        // the blocks still iterating, in local order; every CTA derives the same partition
        int act = 0;
        for (int b = 0; b < P.T_local; ++b)
            act += P.bstate[b].finished ? 0 : 1;
        s_blk = -1;
        s_G = 0;
        if (act > 0) {
            // (past gcap CTAs a group's reduction costs more than its shorter evaluation saves)
            const int G = min((int)gridDim.x / act, P.gcap);
            const int grp = (int)blockIdx.x / G;
            if (grp < act) {
                int seen = 0;
                for (int b = 0; b < P.T_local; ++b)
                    if (!P.bstate[b].finished && seen++ == grp)
                        s_blk = b;
                s_G = G;
            }
        }
    }
    __syncthreads();
    const int blk = s_blk, G = s_G;
    if (blk < 0)
        return;
    const int cta = (int)blockIdx.x % G;
    BlockState *bs = P.bstate + blk;
    const bool resume = bs->started != 0;
    const long long t_kernel0 = (P.prof && threadIdx.x == 0) ? sqw_now_ns() : 0;
    const double pho = P.ctrl->pho, fold = P.ctrl->fold;
    double *xt = P.xt + (size_t)blk * P.Cp * P.dimp;
    double *ut = P.ut + (size_t)blk * P.upad;
    __shared__ StructCache sc;
    sc.init(P);
    typedef typename EvalSelect<NCT, MODE>::type Eval;
    Eval ev{P, blk, cta, G};
    ev.init(smem_raw, L, sm.red);
    ev.set_profile(P.prof ? s_prof + 8 : nullptr);

    if (!resume) {
        // u-update (LOne.java:833-861): x-hat_t = a x_t + (1-a) z_old - a sm_x[par];
        // u_t += x-hat_t - z; u_t /= rho_fold.  Slice of the edge coordinates per CTA.
        const int usl = ((P.upad + G - 1) / G + 15) & ~15;
        const int u0 = min(P.upad, cta * usl), u1 = min(P.upad, u0 + usl);
        for (int i = u0 + threadIdx.x; i < u1; i += kThreads) {
            const int e = i / P.dimp, w = i - e * P.dimp;
            if (w >= P.dim)
                continue;
            const int leaf = P.edge_leaf[e], par = P.edge_par[e];
            double xhat = kAlpha * xt[(size_t)leaf * P.dimp + w] + (1 - kAlpha) * P.zold[i];
            if (par >= 0)
                xhat -= kAlpha * P.smx[(size_t)par * P.dimp + w];
            double u = ut[i] + xhat - P.z[i];
            ut[i] = u / fold;
        }
    } else {
        // resume an interrupted solve: scalar state back into shared memory
        for (int i = threadIdx.x; i < (int)(sizeof(LbState) / sizeof(int)); i += kThreads)
            reinterpret_cast<int *>(&sm.st)[i] = reinterpret_cast<const int *>(&bs->st)[i];
    }

    GroupBarrier gb{P.bar + ((size_t)phase * P.T_local + blk) * 32, 0u, G};
    if (MODE == 9 && P.skew_ns > 0 && (int)blockIdx.x - cta >= P.n_sm - G / 2) {
        // Two CTAs per SM only pay when the two evaluate at DIFFERENT times: the CTAs of the second
        // wave (block indices past the number of SMs share an SM with one of the first wave) start
        // half a trip late, after which the pairs stay in anti-phase - one in its evaluation
        // (DMMA pipe), the other in its solver's exchange chain (latency). Left alone, all blocks
        // start together and stay in step, each evaluation taking twice as long.
        const long long t_end = sqw_now_ns() + P.skew_ns;
        while (sqw_now_ns() < t_end)
            __nanosleep(500);
    }
    __syncthreads();
#if SQW_AUG_CACHE
    if (!resume)
        gb.sync();  // u_t complete in every CTA's slice before the solver caches its inputs
#endif
    BlockProblem<Eval> prob{P, blk, cta, G, pho, ut, ev, sc.lep_p, sc.epar_p};
    SolveWs ws;
    ws.n = P.npad;
    ws.x = xt;
    ws.g = P.g + (size_t)blk * P.npad;
    ws.gold = P.gold + (size_t)blk * P.npad;
    ws.d = P.d + (size_t)blk * P.npad;
    ws.wa = P.wa + (size_t)blk * P.npad;
    ws.S = P.S + (size_t)blk * kLbfgsM * P.npad;
    ws.Y = P.Y + (size_t)blk * kLbfgsM * P.npad;
    ws.spart = P.spart + (size_t)blk * P.Gmax * kNumDots;
    ws.G = G;
    ws.cta = cta;
    ws.prof = P.prof ? s_prof : nullptr;
    // modes 4/5 keep W in registers: the head of the dynamic shared memory is the slice cache
    ws.slice = ((MODE == 4 || MODE == 5 || MODE == 8 || MODE == 9 || MODE == 10 || MODE == 11 || MODE == 12) && NCT == 1)
                   ? reinterpret_cast<double *>(smem_raw) : nullptr;
    int status = 0;
    bool finished = false;
    const long long t_loop0 = (P.prof && threadIdx.x == 0) ? sqw_now_ns() : 0;
    const int evals = lbfgs_group_solve<BlockProblem<Eval>, kThreads, kCap>(
        prob, ws, gb, sm, 0.1, 10e-16, status, eval_cap, resume, &finished);  // eps, xtol: LOne.java:874-875
    const long long t_loop1 = (P.prof && threadIdx.x == 0) ? sqw_now_ns() : 0;
    ev.drain();
    if (P.prof && threadIdx.x == 0) {
        s_prof[12] += t_loop0 - t_kernel0;       // kernel prologue
        s_prof[8] += sqw_now_ns() - t_loop1;     // kernel epilogue
        for (int i = 0; i < kProfSlots; ++i)
            P.prof[i] += s_prof[i];
    }
    if (cta == 0) {
        // every CTA of the group holds the same scalar state; CTA 0 hands it to the next phase
        __syncthreads();
        if (!finished)
            for (int i = threadIdx.x; i < (int)(sizeof(LbState) / sizeof(int)); i += kThreads)
                reinterpret_cast<int *>(&bs->st)[i] = reinterpret_cast<const int *>(&sm.st)[i];
        if (threadIdx.x == 0) {
            bs->started = 1;
            if (finished) {
                bs->finished = 1;
                bs->status = status;
                P.nfev[blk] = evals;
                if (status == 1)
                    atomicAdd(&P.ctrl->lbfgs_errors, 1);
                if (status == 2)
                    atomicExch(&P.ctrl->fatal, 1);
            }
        }
    }
}

// ------------------------------------------------------------------ pack: sum over local blocks
// WITH_SQ: a third section sum_t x_t^2 per coordinate, so that the primal residual of the NEW z
// can be formed from all-reduced sums alone (one collective per iteration, see consensus_edge).
template <bool WITH_SQ>
__global__ void __launch_bounds__(256) admm_pack_kernel(AdmmDev P)
{
    if (ld_volatile_int(&P.ctrl->done) || ld_volatile_int(&P.ctrl->fatal))
        return;
    const int total = P.npad + P.upad + (WITH_SQ ? P.npad : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        if (i < P.npad) {
            for (int b = 0; b < P.T_local; ++b)  // local blocks are stored in summation order
                s += P.xt[(size_t)b * P.Cp * P.dimp + i];
        } else if (i < P.npad + P.upad) {
            for (int b = 0; b < P.T_local; ++b)
                s += P.ut[(size_t)b * P.upad + (i - P.npad)];
        } else {
            for (int b = 0; b < P.T_local; ++b) {
                const double v = P.xt[(size_t)b * P.Cp * P.dimp + (i - P.npad - P.upad)];
                s += v * v;
            }
        }
        P.sum1[i] = s;
    }
}

// ------------------------------------------------------------------ consensus, one CTA per edge
// LOne.java:670-716 restricted to edge e = (leaf, par): z_old <- z, x-bar, u-bar, x-hat, z,
// soft-threshold with kappa = 2 lambda / (rho T) (intercept included), znorm (from u for a self
// edge, :504), xnorm, dual residual term, and this rank's part of the primal residual term.
// FUSED: the sums over the local blocks are formed here (same order as admm_pack_kernel) instead
// of being read from sum1 - the single-GPU path, where no all-reduce sits between pack and consensus.
// MODE_ = 0: sums from sum1 (all-reduced), primal residual of this rank's blocks (a second
//            all-reduce of sum2 follows);
//         1: FUSED above (single GPU);
//         2: sums AND the whole job's primal residual from the all-reduced sum1, which then also
//            carries sum_t x_t^2: sum_t (x_t - a)^2 = sum_t x_t^2 - 2 a sum_t x_t + T a^2 with
//            a = z_new + sm_x[par]. One collective per ADMM iteration. The expansion cancels when
//            the blocks agree (x_t ~ a); the residual it feeds is compared with a tolerance of
//            1e-2 relative, the cancellation error is ~1e-12 of it (test_one_collective_tail...).
template <int MODE_>
__device__ __forceinline__ void consensus_edge(const AdmmDev &P, double *red)
{
    constexpr bool FUSED = MODE_ == 1;
    const int e = blockIdx.x;
    const int leaf = P.edge_leaf[e], par = P.edge_par[e];
    const bool first_edge = (P.leaf_edge_ptr[leaf] == e);
    const double pho = P.ctrl->pho;
    const double kappa = (2 * P.ridge) / (pho * P.T);
    const double Td = (double)P.T;
    double zn2 = 0, du2 = 0, xn2 = 0, pr2 = 0, cn2 = 0;
    for (int w = threadIdx.x; w < P.dim; w += blockDim.x) {
        const size_t o = (size_t)e * P.dimp + w, xo = (size_t)leaf * P.dimp + w;
        double sx, su;
        if (FUSED) {
            sx = 0.0;
            su = 0.0;
            for (int b = 0; b < P.T_local; ++b) {  // local blocks are stored in summation order
                sx += P.xt[(size_t)b * P.Cp * P.dimp + xo];
                su += P.ut[(size_t)b * P.upad + o];
            }
        } else {
            sx = P.sum1[xo];
            su = P.sum1[P.npad + o];
        }
        const double xb = sx / Td;
        const double ub = su / Td;
        const double zo = P.z[o];
        double smp = 0.0;
        double xhat = kAlpha * xb + (1 - kAlpha) * zo;
        if (par >= 0) {
            smp = P.smx[(size_t)par * P.dimp + w];
            xhat -= kAlpha * smp;
            cn2 += smp * smp;
        }
        double zn = xhat + ub;
        const double sg = (zn > 0.0) ? 1.0 : ((zn < 0.0) ? -1.0 : zn);
        zn = zn - sg * fmin(kappa, fabs(zn));
        P.zold[o] = zo;
        P.z[o] = zn;
        P.ubar[o] = ub;
        if (first_edge) {
            P.xbar[xo] = xb;
            xn2 += xb * xb;
        }
        zn2 += (par >= 0) ? zn * zn : ub * ub;
        const double dd = pho * (zn - zo);
        du2 += dd * dd;
        if (MODE_ == 2) {
            const double a = zn + smp;
            pr2 += (P.sum1[P.npad + P.upad + xo] - 2.0 * a * sx) + Td * a * a;
        } else {
            for (int b = 0; b < P.T_local; ++b) {
                const double r = P.xt[(size_t)b * P.Cp * P.dimp + xo] - zn - smp;
                pr2 += r * r;
            }
        }
    }
    zn2 = block_sum(zn2, red);
    du2 = block_sum(du2, red);
    xn2 = block_sum(xn2, red);
    pr2 = block_sum(pr2, red);
    cn2 = block_sum(cn2, red);
    if (threadIdx.x == 0) {
        P.edge_stats[e * 4 + 0] = zn2;
        P.edge_stats[e * 4 + 1] = du2;
        P.edge_stats[e * 4 + 2] = cn2;
        P.sum2[e] = pr2;
        if (first_edge)
            P.leaf_xnorm[leaf] = xn2;
    }
}

__global__ void __launch_bounds__(kThreads) admm_consensus_kernel(AdmmDev P)
{
    __shared__ double red[kWarps];
    if (ld_volatile_int(&P.ctrl->done) || ld_volatile_int(&P.ctrl->fatal))
        return;
    consensus_edge<0>(P, red);
}

// ------------------------------------------------------------------ decide
// The decision itself is the reference's serial bookkeeping and runs in one thread; but from
// global memory it was ~140 dependent loads (20 us per ADMM iteration under ncu). The CTA first
// stages every input in shared memory with one coalesced round trip, the other threads do the
// per-block resets meanwhile, and thread 0 works on the staged copies.
constexpr int kDecideThreads = 256;
// (all threads of ONE CTA; inputs written by other CTAs are read with ld.cg)
__device__ void decide_cta(const AdmmDev &P)
{
    constexpr int kML = StructCache::kMaxLeaves, kME = StructCache::kMaxEdges;
    const int kDecideThreads = (int)blockDim.x;
    __shared__ int s_lep[kML + 1], s_epar[kME], s_order[kME];
    __shared__ double s_xnorm[kML], s_sum2[kME], s_stats[4 * kME];
    __shared__ Ctrl s_ctrl;
    Ctrl *c = P.ctrl;
    const bool fits = P.C <= kML && P.E <= kME;
    const int tid = threadIdx.x;
    if (fits) {
        for (int i = tid; i <= P.C; i += kDecideThreads)
            s_lep[i] = P.leaf_edge_ptr[i];
        for (int i = tid; i < P.C; i += kDecideThreads)
            s_xnorm[i] = __ldcg(P.leaf_xnorm + i);
        for (int i = tid; i < P.E; i += kDecideThreads) {
            s_epar[i] = P.edge_par[i];
            s_order[i] = P.edge_sum_order[i];
            s_sum2[i] = __ldcg(P.sum2 + i);
        }
        for (int i = tid; i < 4 * P.E; i += kDecideThreads)
            s_stats[i] = __ldcg(P.edge_stats + i);
    }
    if (tid < (int)(sizeof(Ctrl) / sizeof(int)))
        reinterpret_cast<int *>(&s_ctrl)[tid] = reinterpret_cast<const int *>(c)[tid];
    // per-block resets for the next iteration (independent of the decision)
    for (int i = tid; i < P.T_local * P.nphase; i += kDecideThreads)
        P.bar[(size_t)i * 32] = 0u;
    for (int b = tid; b < P.T_local; b += kDecideThreads) {
        P.bstate[b].started = 0;
        P.bstate[b].finished = 0;
    }
    __syncthreads();
    if (tid != 0)
        return;
    const int *lep = fits ? s_lep : P.leaf_edge_ptr, *epar = fits ? s_epar : P.edge_par;
    const int *order = fits ? s_order : P.edge_sum_order;
    const double *lxn = fits ? s_xnorm : P.leaf_xnorm, *sum2 = fits ? s_sum2 : P.sum2;
    double *stats = fits ? s_stats : P.edge_stats;
    Ctrl &lc = s_ctrl;

    // updateResiduals (LOne.java:418-454): the accumulators run on across a leaf's parents and
    // the dual one is square-rooted in place.
    bool converged = true;
    const double abs_tol = sqrt((double)P.dim) * kAbsTol;
    const double rel = sqrt((double)P.T) * kRelTol;
    for (int leaf = 0; leaf < P.C; ++leaf) {
        double r_t = 0.0, s_t = 0.0;
        const int e0 = lep[leaf], e1 = lep[leaf + 1];
        const double xnorm = sqrt(lxn[leaf]);
        for (int e = e0; e < e1; ++e) {
            r_t += fmax(sum2[e], 0.0);   // (the one-collective expansion can round below zero)
            s_t += stats[e * 4 + 1];
            s_t = sqrt(s_t * P.T);
            const double primal = sqrt(r_t), dual = s_t;
            stats[e * 4 + 3] = primal;
            stats[e * 4 + 1] = dual;
            if (fits) {
                P.edge_stats[e * 4 + 3] = primal;
                P.edge_stats[e * 4 + 1] = dual;
            }
            // hasADMMConverged (LOne.java:544-581); unorm is always 0 when read (:650 vs :680)
            const double znorm = sqrt(stats[e * 4 + 0]);
            const double cnorm = sqrt(stats[e * 4 + 2]);
            const double mx = (epar[e] >= 0) ? fmax(xnorm, fmax(znorm, cnorm)) : fmax(xnorm, znorm);
            const double primal_tol = abs_tol + rel * mx;
            const double dual_tol = abs_tol + rel * lc.pho * 0.0;
            if (primal > primal_tol || dual > dual_tol)
                converged = false;
        }
    }
    double ps = 0.0, ds = 0.0;  // LOne.java:519-522 sums, in i = leaf*NN+pid order
    for (int i = 0; i < P.E; ++i) {
        const int e = order[i];
        ps += stats[e * 4 + 3];
        ds += stats[e * 4 + 1];
    }
    if (lc.log_rows < lc.log_cap) {
        const int r = lc.log_rows++;
        P.log_outer[r] = lc.outer;
        P.log_itr[r] = lc.itr;
        P.log_pho[r] = lc.pho;
        P.log_fold[r] = lc.fold;
        P.log_primal[r] = ps;
        P.log_dual[r] = ds;
        for (int b = 0; b < P.T_local; ++b)
            P.log_nfev[(size_t)r * P.T_local + b] = P.nfev[b];
    }
    for (int b = 0; b < P.T_local; ++b)
        lc.total_evals += P.nfev[b];
    lc.itr += 1;
    if (lc.itr < lc.max_itr) {
        // top of the next loop trip: updatePhoAndFold(itr-1) (LOne.java:511-535,614-615) ...
        if (!lc.ran_admm && lc.pho < kPhoMax) {
            const double mu = 10, tao = 2;
            if (ps > mu * ds) {
                const double old = lc.pho;
                lc.pho = fmin(kPhoMax, lc.pho * tao);
                lc.fold = lc.pho / old;
            } else if (ds > ps * mu) {
                lc.pho = lc.pho / tao;
                lc.fold = 1 / tao;
            } else {
                lc.fold = 1.0;
            }
        }
        // ... then hasADMMConverged(itr-1) (:649-665)
        if (converged) {
            lc.ran_admm = 1;
            lc.done = 1;
        }
    } else {
        lc.done = 1;
    }
    // publish the fields this kernel owns (lbfgs_errors / fatal / cons_count belong to others)
    c->pho = lc.pho;
    c->fold = lc.fold;
    c->ran_admm = lc.ran_admm;
    c->itr = lc.itr;
    c->log_rows = lc.log_rows;
    c->total_evals = lc.total_evals;
    c->done = lc.done;
}

__global__ void __launch_bounds__(kDecideThreads) admm_decide_kernel(AdmmDev P)
{
    if (blockIdx.x != 0)
        return;
    if (ld_volatile_int(&P.ctrl->done) || ld_volatile_int(&P.ctrl->fatal))
        return;
    decide_cta(P);
}

// Single-GPU iteration tail in ONE launch: pack + consensus per edge, and the CTA that finishes
// last (ticket counter in Ctrl) runs the decision. Same arithmetic in the same order as the three
// kernels above, two launches and their gaps less per ADMM iteration.
// MODE_ 1: single GPU (sums over the local blocks formed here); MODE_ 2: multi-GPU after ONE
// all-reduce of [sum x_t | sum u_t | sum x_t^2].
template <int MODE_>
__global__ void __launch_bounds__(kThreads) admm_consensus_fused_kernel(AdmmDev P)
{
    __shared__ double red[kWarps];
    __shared__ int s_last;
    if (ld_volatile_int(&P.ctrl->done) || ld_volatile_int(&P.ctrl->fatal))
        return;
    consensus_edge<MODE_>(P, red);
    if (threadIdx.x == 0) {
        __threadfence();  // this edge's statistics before the ticket
        const unsigned t = atomicAdd(&P.ctrl->cons_count, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last)
        return;
    if (threadIdx.x == 0) {
        P.ctrl->cons_count = 0u;
        __threadfence();
    }
    __syncthreads();
    decide_cta(P);
}

// ------------------------------------------------------------------ outer-iteration kernels
// new ADMMrunner() (LOne.java:360-391): t_x <- x, t_u <- u; iteration counter reset.
__global__ void __launch_bounds__(256) admm_begin_kernel(AdmmDev P, int outer, int snapshot)
{
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = tid; i < P.npad; i += stride)
        for (int b = 0; b < P.T_local; ++b)
            P.xt[(size_t)b * P.Cp * P.dimp + i] = P.xbar[i];
    for (int i = tid; i < P.upad; i += stride)
        for (int b = 0; b < P.T_local; ++b)
            P.ut[(size_t)b * P.upad + i] = P.ubar[i];
    if (snapshot)  // sm_x_old (LOne.java:122-126)
        for (int i = tid; i < P.NN * P.dimp; i += stride)
            P.smx_old[i] = P.smx[i];
    if (tid == 0) {
        P.ctrl->itr = 0;
        P.ctrl->done = 0;
        P.ctrl->outer = outer;
        P.ctrl->num_pass = 0;
        P.ctrl->num_pass_relaxed = 0;
        for (int b = 0; b < P.T_local; ++b) {
            P.bstate[b].started = 0;
            P.bstate[b].finished = 0;
            P.nfev[b] = 0;
            for (int ph = 0; ph < P.nphase; ++ph)
                P.bar[((size_t)ph * P.T_local + b) * 32] = 0u;
        }
    }
}

// sm_x[leaf] <- x-bar (LOne.java:753-758)
__global__ void __launch_bounds__(256) admm_copy_leaves_kernel(AdmmDev P)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.npad; i += gridDim.x * blockDim.x)
        P.smx[i] = P.xbar[i];
}

// updateNode (LOne.java:223-241) for the nodes of one layer: lower median over parents' and
// children's weights, intercept untouched. One thread per (node, w).
__global__ void __launch_bounds__(256)
admm_update_layer_kernel(AdmmDev P, const int *nodes, int n_nodes, const int *nbr_ptr,
                         const int *nbr_idx)
{
    const int per = P.dim - 1;
    const long long total = (long long)n_nodes * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ni = (int)(i / per), w = (int)(i - (long long)ni * per) + 1;
        const int node = nodes[ni];
        const int b = nbr_ptr[ni], cnt = nbr_ptr[ni + 1] - b;
        if (cnt <= 0)
            continue;  // xs.get(-1) would throw in the reference; never happens for its designs
        double xs[kMaxFanIn];
        for (int j = 0; j < cnt; ++j) {  // insertion sort (Collections.sort, ascending)
            const double v = P.smx[(size_t)nbr_idx[b + j] * P.dimp + w];
            int q = j;
            while (q > 0 && xs[q - 1] > v) {
                xs[q] = xs[q - 1];
                --q;
            }
            xs[q] = v;
        }
        P.smx[(size_t)node * P.dimp + w] = (cnt % 2 == 0) ? xs[cnt / 2 - 1] : xs[cnt / 2];
    }
}

// LOne.java:152-174: per node delta = ||sm_x_old - sm_x||, target = NODES_tol * ||sm_x||.
__global__ void __launch_bounds__(kThreads) admm_outer_check_kernel(AdmmDev P)
{
    __shared__ double red[kWarps];
    const int node = blockIdx.x;
    double d2 = 0, n2 = 0;
    for (int w = threadIdx.x; w < P.dim; w += blockDim.x) {
        const size_t o = (size_t)node * P.dimp + w;
        const double a = P.smx[o], df = P.smx_old[o] - a;
        d2 += df * df;
        n2 += a * a;
    }
    d2 = block_sum(d2, red);
    n2 = block_sum(n2, red);
    if (threadIdx.x == 0) {
        const double delta = sqrt(d2), target = kNodesTol * sqrt(n2);
        if (delta < target)
            atomicAdd(&P.ctrl->num_pass, 1);
        if (delta < target * 2 / kNodesTol)
            atomicAdd(&P.ctrl->num_pass_relaxed, 1);
    }
}

// Gather rows into the block-major padded layout (LOne.java:337-357): local block b holds rows
// i = t + T*s of the caller's row-major matrix, t = block_ids[b].
__global__ void __launch_bounds__(256)
admm_block_rows_kernel(const double *src, const int *ysrc, const double *wsrc, int dim, int T,
                       const int *block_ids, int T_local, int nb, int dimp, double *X, int *y,
                       double *w)
{
    const long long rows = (long long)T_local * nb;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const int b = (int)(r / nb), s = (int)(r - (long long)b * nb);
        const long long i = (long long)block_ids[b] + (long long)T * s;
        const double *in = src + i * dim;
        double *out = X + r * dimp;
        for (int j = threadIdx.x; j < dimp; j += blockDim.x)
            out[j] = (j < dim) ? in[j] : 0.0;
        if (threadIdx.x == 0) {
            y[r] = ysrc[i];
            w[r] = wsrc[i];
        }
    }
}

// The same gather from a COUNT matrix (MakeArff.java:347-368 rows before Classifier.java:373-381):
// src row i holds the counts of the features (column col_idx[j-1] when col_idx is given, else
// j-1) for j = 1..dim-1; destination column 0 is the constant 1 (Classifier.java:340).
//   PACKED   bytes, row stride `stride` (MODE 8); a count above 255 raises status bit 0
//   !PACKED  fp64 (n - mean[j]) / sd[j] where sd[j] != 0 (Classifier.java:373-381), row stride `stride`
// rows_local: src / ysrc / wsrc hold only this rank's rows in ascending global order, i.e. local
// row s * T_local + t / world for block t (the rank owns the blocks t = rank mod world).
template <typename SrcT, bool PACKED>
__global__ void __launch_bounds__(256)
admm_gather_counts_kernel(const SrcT *src, long long src_stride, const int *col_idx, const int *ysrc,
                          const double *wsrc, int dim, int T, const int *block_ids, int T_local,
                          int nb, int world, int rows_local, const double *mean, const double *sd,
                          int stride, void *Xout, int *y, double *w, int *status)
{
    const long long rows = (long long)T_local * nb;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const int b = (int)(r / nb), s = (int)(r - (long long)b * nb);
        const int t = block_ids[b];
        const long long i = rows_local ? (long long)s * T_local + t / world : (long long)t + (long long)T * s;
        const SrcT *in = src + i * src_stride;
        for (int j = threadIdx.x; j < stride; j += blockDim.x) {
            long long v = 0;
            if (j == 0)
                v = 1;
            else if (j < dim)
                v = (long long)in[col_idx ? col_idx[j - 1] : j - 1];
            if (PACKED) {
                if (v < 0 || v > 255) {
                    atomicOr(status, 1);
                    v = 0;
                }
                static_cast<unsigned char *>(Xout)[r * stride + j] = (unsigned char)v;
            } else {
                double x = (double)v;
                if (j >= 1 && j < dim && sd[j] != 0.0)
                    x = (x - mean[j]) / sd[j];
                static_cast<double *>(Xout)[r * stride + j] = x;
            }
        }
        if (threadIdx.x == 0) {
            y[r] = ysrc[i];
            w[r] = wsrc[i];
        }
    }
}

// mean_p / invsd_p [dimp] of the packed form from xMean / xSD: per feature j (src index
// col_idx[j-1] + ... when the statistics are indexed by original column, else j)
__global__ void __launch_bounds__(256)
admm_packed_stats_kernel(const double *mean, const double *sd, const int *col_idx, int dim, int dimp,
                         double *mean_p, double *invsd_p, double *mean_f, double *sd_f)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < dimp; j += gridDim.x * blockDim.x) {
        double m = 0.0, s = 0.0, is = 0.0;
        if (j == 0) {
            s = 1.0;
            is = 1.0;
        } else if (j < dim) {
            const int k = col_idx ? col_idx[j - 1] : j;
            m = mean[k];
            s = sd[k];
            if (s != 0.0) {
                is = 1.0 / s;
            } else {  // Classifier.java:376: a column with zero SD keeps its raw values
                m = 0.0;
                is = 1.0;
            }
        }
        mean_p[j] = m;
        invsd_p[j] = is;
        mean_f[j] = (j >= 1 && j < dim) ? mean[col_idx ? col_idx[j - 1] : j] : 0.0;
        sd_f[j] = (j >= 1 && j < dim) ? s : (j == 0 ? 1.0 : 0.0);
    }
}

// ------------------------------------------------------------------ test kernels
// One evaluation per local block at the current x_t with the device's z / u_t / sm_x state.
template <int NCT, int MODE>
__global__ void __launch_bounds__(ModeCfg<MODE>::NT, ModeCfg<MODE>::MINB)
admm_eval_kernel(AdmmDev P, TileSmemLayout L, double pho, double *f_out)
{
    constexpr int kThreads = ModeCfg<MODE>::NT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ SolveSmem sm;
    const int blk = blockIdx.x / P.G, cta = blockIdx.x - blk * P.G;
    double *xt = P.xt + (size_t)blk * P.Cp * P.dimp;
    GroupBarrier gb{P.bar + blk * 32, 0u, P.G};
    typedef typename EvalSelect<NCT, MODE>::type Eval;
    Eval ev{P, blk, cta, P.G};
    ev.init(smem_raw, L, sm.red);
    BlockProblem<Eval> prob{P, blk, cta, P.G, pho, P.ut + (size_t)blk * P.upad, ev,
                            P.leaf_edge_ptr, P.edge_par};
    double fpart = prob.eval_partial(xt);
    if (MODE == 4 || MODE == 5 || MODE == 6 || MODE == 8 || MODE == 9 || MODE == 10 || MODE == 11 || MODE == 12) {
        // further evaluations exercise the ring wrap-around / resident path; results must agree
        for (int rep = 0; rep < 2; ++rep) {
            __syncthreads();
            const double f2 = prob.eval_partial(xt);
            if (f2 != fpart)
                fpart = nan("");
        }
    }
    fpart = block_sum<kThreads / 32>(fpart, sm.red);
    ev.drain();
    gb.sync();
    const int n = P.npad;
    const int sl = ((n + P.G - 1) / P.G + 15) & ~15;
    const int k0 = min(n, cta * sl), k1 = min(n, k0 + sl);
    double faug = 0.0;
    for (int k = k0 + threadIdx.x; k < k1; k += kThreads)
        P.g[(size_t)blk * n + k] = prob.grad_coord(k, __ldcg(xt + k), faug);
    const double fa = block_sum<kThreads / 32>(faug, sm.red);
    if (threadIdx.x == 0)
        atomicAdd(f_out + blk, fpart + fa);
}

struct RosenProblem {
    int n;
    __device__ double eval_partial(const double *) { return 0.0; }
    const double *x;
    static constexpr int kBatch = 1, kSmallBatch = 1;
    __device__ void data_issue(int, double (&)[kBatch]) {}
    __device__ double data_finish(int, double (&)[kBatch]) { return 0.0; }
    __device__ void data_issue_small(int, double (&)[kSmallBatch]) {}
    __device__ double data_finish_small(int, double (&)[kSmallBatch]) { return 0.0; }
    __device__ bool aug_load(int, double (&)[6], int &) { return false; }
    __device__ double aug_term_cached(double, const double (&)[6], int, double &) { return 0.0; }
    __device__ double aug_term(int k, double, double &facc)
    {
        const int i = k >> 1;
        const double a = __ldcg(x + 2 * i), b = __ldcg(x + 2 * i + 1);
        const double t1 = b - a * a;
        if ((k & 1) == 0) {
            facc += 100.0 * t1 * t1 + (1.0 - a) * (1.0 - a);
            return -400.0 * a * t1 - 2.0 * (1.0 - a);
        }
        return 200.0 * t1;
    }
};

__global__ void __launch_bounds__(kThreads, 1)
lbfgs_rosenbrock_kernel(SolveWs ws, unsigned *bar, double eps, int *out)
{
    __shared__ SolveSmem sm;
    extern __shared__ __align__(16) double rosen_slice[];
    ws.slice = rosen_slice;
    RosenProblem prob{ws.n, ws.x};
    ws.cta = blockIdx.x;
    GroupBarrier gb{bar, 0u, ws.G};
    int status = 0;
    const int evals = lbfgs_group_solve(prob, ws, gb, sm, eps, 10e-16, status);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = evals;
        out[1] = status;
    }
}

}  // namespace sqw
