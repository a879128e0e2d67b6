// admm_solver.cuh — device-resident L-BFGS (m = 5) with the Moré–Thuente line search, run by a
// group of CTAs without any host round trip.
//
// Replaces framework/LBFGS.java:338-592 and framework/Mcsrch.java:173-369,442-631 as driven by
// LOne.ADMMrun.run (framework/LOne.java:865-896): fresh solver per x-update, m = 5, eps = 0.1,
// xtol = 1e-15, no convergence test before the first line search, and the reference's local
// constants (stpmin 1e-5, stpmax 1e20, gtol 0.9, ftol 1e-4, maxfev 200, TWO_THIRDS 0.66, steps
// clipped to 1000, every line-search exit code 2..6 reported as success).
//
// B200-first restructuring (same mathematics, different schedule):
//   * the reference's reverse-communication state machine becomes straight-line control flow that
//     every CTA of the group executes redundantly on identical reduced scalars, so all CTAs take
//     identical branches with no broadcast;
//   * the two-loop recursion's 2m sequential dot products would cost 2m group barriers per
//     iteration (1.2 us each on B200, tools/microbench); instead every evaluation's reduce phase
//     produces, in ONE pass over the CTA's coordinate slice, all inner products the next
//     direction needs (g, y_new, d against the stored s_i / y_i), the Gram matrices S^T Y and
//     Y^T Y are maintained incrementally, and the recursion runs on scalars. The new direction is
//     then a single fused linear combination  d = -gamma g - sum gamma alpha_i y_i + sum beta_i s_i;
//   * three group barriers per evaluation: partial gradients visible -> scalars visible -> trial
//     point visible.
#pragma once

#include "admm_eval.cuh"

namespace sqw {

// ------------------------------------------------------------------ group barrier
// All G CTAs of a group are co-resident (cooperative launch). Monotonic counter, release/acquire
// at gpu scope; data exchanged across the barrier is read with ld.global.cg.
struct GroupBarrier {
    unsigned *ctr;
    unsigned target;
    int G;
    // Split form: arrive() publishes this CTA's writes, wait() returns once all G CTAs have
    // arrived. Loads that do not depend on the other CTAs' current phase go in between, so their
    // L2 latency overlaps the barrier's.
    __device__ __forceinline__ void arrive()
    {
        target += (unsigned)G;
        __syncthreads();
        if (threadIdx.x == 0)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    }
    __device__ __forceinline__ void wait()
    {
        // (lane 0 of four warps polling, the first to see the count telling the rest through shared
        // memory, was measured: 60.8 vs 59.9 us per evaluation round - more polls on the counter's
        // line slow the arrivals down)
        if (threadIdx.x == 0) {
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < target);
        }
        __syncthreads();
    }
    __device__ __forceinline__ void sync()
    {
        arrive();
        wait();
    }
};

// ------------------------------------------------------------------ Moré–Thuente, scalar part
constexpr double kGtol = 0.9;      // LBFGS.java:116
constexpr double kStpMin = 1e-5;   // LBFGS.java:125
constexpr double kStpMax = 1e20;   // LBFGS.java:134
constexpr int kMaxFev = 200;       // LBFGS.java:136
constexpr double kFtol = 1e-4;     // LBFGS.java:398
constexpr double kTwoThirds = 0.66;  // Mcsrch.java:38

struct McState {
    int infoc, brackt, stage1, nfev;
    double dginit, dgtest, finit;
    double stx, fx, dgx, sty, fy, dgy;
    double stmin, stmax, width, width1;
};

// Mcsrch.java:621-631 (value < 0 throws in the reference -> fatal)
__device__ __forceinline__ void set_stp(double &stp, double value, int &fatal)
{
    if (value < 0.0 || isnan(value)) {
        fatal = 1;
        return;
    }
    stp = value > 1000.0 ? 1000.0 : value;
}

__device__ __forceinline__ double max3(double x, double y, double z)
{
    return x < y ? (y < z ? z : y) : (x < z ? z : x);
}

// Mcsrch.java:442-619
__device__ inline void mcstep(double &stx, double &fx, double &dx, double &sty, double &fy,
                              double &dy, double &stp, double fp, double dp, int &brackt,
                              double stpmin, double stpmax, int &info, int &fatal)
{
    info = 0;
    if ((brackt && (stp <= fmin(stx, sty) || stp >= fmax(stx, sty))) ||
        dx * (stp - stx) >= 0.0 || stpmax < stpmin)
        return;
    const double sgnd = dp * (dx / fabs(dx));
    double gamma, p, q, r, s, stpc, stpf, stpq, theta;
    bool bound;
    if (fp > fx) {
        info = 1;
        bound = true;
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx)
            gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2;
        brackt = 1;
    } else if (sgnd < 0.0) {
        info = 2;
        bound = false;
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx)
            gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp))
            stpf = stpc;
        else
            stpf = stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        info = 3;
        bound = true;
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = max3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx)
            gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = stp + r * (stx - stp);
        else if (stp > stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt)
            stpf = (fabs(stp - stpc) < fabs(stp - stpq)) ? stpc : stpq;
        else
            stpf = (fabs(stp - stpc) > fabs(stp - stpq)) ? stpc : stpq;
    } else {
        info = 4;
        bound = false;
        if (brackt) {
            theta = 3 * (fp - fy) / (sty - stp) + dy + dp;
            s = max3(fabs(theta), fabs(dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty)
                gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp;
        fy = fp;
        dy = dp;
    } else {
        if (sgnd < 0.0) {
            sty = stx;
            fy = fx;
            dy = dx;
        }
        stx = stp;
        fx = fp;
        dx = dp;
    }
    stpf = fmin(stpmax, stpf);
    stpf = fmax(stpmin, stpf);
    set_stp(stp, stpf, fatal);
    if (brackt && bound) {
        const double possible = stx + kTwoThirds * (sty - stx);
        if (sty > stx)
            set_stp(stp, fmin(possible, stp), fatal);
        else
            set_stp(stp, fmax(possible, stp), fatal);
    }
}

// First entry of mcsrch (Mcsrch.java:178-230). Returns false for the "improper input / not a
// descent direction" exits that leave info == 0 (LBFGS.java:542-549 then throws).
__device__ inline bool mc_begin(McState &mc, double f, double dginit, double stp, double xtol)
{
    mc.infoc = 1;
    if (stp <= 0.0 || xtol < 0.0)
        return false;
    mc.dginit = dginit;
    if (!(dginit < 0.0))
        return false;
    mc.brackt = 0;
    mc.stage1 = 1;
    mc.nfev = 0;
    mc.finit = f;
    mc.dgtest = kFtol * dginit;
    mc.width = kStpMax - kStpMin;
    mc.width1 = mc.width / 0.5;
    mc.stx = 0.0;
    mc.fx = f;
    mc.dgx = dginit;
    mc.sty = 0.0;
    mc.fy = f;
    mc.dgy = dginit;
    return true;
}

// Top of the mcsrch loop (Mcsrch.java:233-270): interval, clamps, unusual-termination step.
__device__ inline void mc_next_step(McState &mc, double &stp, double xtol, int &fatal)
{
    if (mc.brackt) {
        mc.stmin = fmin(mc.stx, mc.sty);
        mc.stmax = fmax(mc.stx, mc.sty);
    } else {
        mc.stmin = mc.stx;
        mc.stmax = stp + 4.0 * (stp - mc.stx);
    }
    set_stp(stp, fmax(stp, kStpMin), fatal);
    set_stp(stp, fmin(stp, kStpMax), fatal);
    if ((mc.brackt && (stp <= mc.stmin || stp >= mc.stmax ||
                       mc.stmax - mc.stmin <= xtol * mc.stmax)) ||
        mc.nfev >= kMaxFev - 1 || mc.infoc == 0)
        set_stp(stp, mc.stx, fatal);
}

// After an evaluation at the trial step (Mcsrch.java:273-368). Returns true when the search is
// over (the reference rewrites every exit code to 1); otherwise stp holds the next raw step.
__device__ inline bool mc_after_eval(McState &mc, double f, double dg, double &stp, double xtol,
                                     int &fatal)
{
    int info = 0;
    mc.nfev = mc.nfev + 1;
    const double ftest1 = mc.finit + stp * mc.dgtest;
    if ((mc.brackt && (stp <= mc.stmin || stp >= mc.stmax)) || mc.infoc == 0)
        info = 6;
    if (stp == kStpMax && f <= ftest1 && dg <= mc.dgtest)
        info = 5;
    if (stp == kStpMin && (f > ftest1 || dg >= mc.dgtest))
        info = 4;
    if (mc.nfev >= kMaxFev)
        info = 3;
    if (mc.brackt && mc.stmax - mc.stmin <= xtol * mc.stmax)
        info = 2;
    if (f <= ftest1 && fabs(dg) <= kGtol * (-mc.dginit))
        info = 1;
    if (info != 0)
        return true;

    if (mc.stage1 && f <= ftest1 && dg >= fmin(kFtol, kGtol) * mc.dginit)
        mc.stage1 = 0;

    if (mc.stage1 && f <= mc.fx && f > ftest1) {
        const double fm = f - stp * mc.dgtest;
        double fxm = mc.fx - mc.stx * mc.dgtest;
        double fym = mc.fy - mc.sty * mc.dgtest;
        const double dgm = dg - mc.dgtest;
        double dgxm = mc.dgx - mc.dgtest;
        double dgym = mc.dgy - mc.dgtest;
        mcstep(mc.stx, fxm, dgxm, mc.sty, fym, dgym, stp, fm, dgm, mc.brackt, mc.stmin, mc.stmax,
               mc.infoc, fatal);
        mc.fx = fxm + mc.stx * mc.dgtest;
        mc.fy = fym + mc.sty * mc.dgtest;
        mc.dgx = dgxm + mc.dgtest;
        mc.dgy = dgym + mc.dgtest;
    } else {
        mcstep(mc.stx, mc.fx, mc.dgx, mc.sty, mc.fy, mc.dgy, stp, f, dg, mc.brackt, mc.stmin,
               mc.stmax, mc.infoc, fatal);
    }
    if (mc.brackt) {
        if (fabs(mc.sty - mc.stx) >= kTwoThirds * mc.width1)
            set_stp(stp, (mc.sty + mc.stx) / 2.0, fatal);
        mc.width1 = mc.width;
        mc.width = fabs(mc.sty - mc.stx);
    }
    return false;
}

// ------------------------------------------------------------------ scalar ids of one reduce
enum DotId : int {
    D_F = 0,    // objective (data partial of the CTA + augmented term of its slice)
    D_GD = 1,   // g . d
    D_GG = 2,   // g . g
    D_XX = 3,   // x . x
    D_GYN = 4,  // g . y_new          (y_new = g - g_old)
    D_YNYN = 5, // y_new . y_new
    D_DYN = 6,  // d . y_new
    D_GS = 7,   // +i : g . s_i
    D_GY = 12,  // +i : g . y_i
    D_YNS = 17, // +i : y_new . s_i
    D_YNY = 22, // +i : y_new . y_i
    D_DY = 27,  // +i : d . y_i
};
static_assert(D_DY + kLbfgsM == kNumDots, "dot table");

enum Action : int { ACT_LINESEARCH = 0, ACT_NEWDIR = 1, ACT_FINISH = 2, ACT_FAIL = 3 };

// Scalar L-BFGS state (identical in every CTA of the group). The Gram matrices are kept in AGE
// order (index 0 = newest pair) so that every loop below has compile-time bounds and the whole
// state lives in registers of the one thread that runs the scalar step; slot <-> age is
// slot(a) = (point - 1 - a) mod m.
struct LbState {
    int iter, point, hist, nfun, evals;
    double stp, f;
    McState mc;
    double SY[kLbfgsM][kLbfgsM];  // s_(age i) . y_(age j)
    double YY[kLbfgsM][kLbfgsM];  // y_(age i) . y_(age j)
    double rho[kLbfgsM];
};

// What thread 0 tells the rest of the CTA after the scalar step.
struct Decision {
    int action;
    int slot;            // history slot written on ACT_NEWDIR
    int hist;            // number of valid pairs after the update
    double stp_accept;   // step that scales d into s_new
    double stp;          // trial step for the next evaluation
    double cg;           // d_new = cg * g + sum cs[i] * s_i + sum cy[i] * y_i   (i = slot)
    double cs[kLbfgsM], cy[kLbfgsM];
    unsigned valid_mask; // slots taking part in d_new / in the next reduce (bit i)
};

// Slots holding valid pairs: the `hist` most recent ones before `point`.
__device__ __forceinline__ unsigned hist_mask(int point, int hist)
{
    unsigned m = 0;
    for (int j = 0; j < hist; ++j)
        m |= 1u << ((point - 1 - j + 2 * kLbfgsM) % kLbfgsM);
    return m;
}

// The scalar step after an evaluation whose reduced scalars are dots[] (shared memory). Mirrors
// the tail of LBFGS.lbfgs (LBFGS.java:531-590) followed by the head of its next loop trip
// (:407-527), with the two-loop recursion evaluated on the Gram matrices.
#ifndef SQW_SCALAR_NOINLINE
#define SQW_SCALAR_NOINLINE 0
#endif
#ifndef SQW_PROF_SPLIT
#define SQW_PROF_SPLIT 0   // probe: time the partial-gradient loads and the augmented term apart
#endif
#if SQW_SCALAR_NOINLINE
__device__ __noinline__ void lbfgs_scalar_step(
#else
__device__ __forceinline__ void lbfgs_scalar_step(
#endif
    LbState &st, const double *dots, double eps,
                                                  double xtol, Decision &dec, int &fatal, bool first)
{
    constexpr int m = kLbfgsM;
    st.evals++;
    st.f = dots[D_F];
    if (first) {
        // LBFGS.java:345-403: iter = 1, d = -g, stp1 = 1/gnorm
        st.iter = 1;
        st.point = 0;
        st.hist = 0;
        st.nfun = 1;
        const double gnorm = sqrt(dots[D_GG]);
        const double stp1 = 1.0 / gnorm;
        dec.cg = -1.0;
        dec.valid_mask = 0;
        dec.slot = -1;
        dec.hist = 0;
        dec.stp_accept = 0.0;
        st.stp = 0.0;
        set_stp(st.stp, (stp1 < 1.0) ? stp1 : 1.0, fatal);
        if (!mc_begin(st.mc, st.f, -dots[D_GG], st.stp, xtol)) {
            dec.action = ACT_FAIL;
            return;
        }
        mc_next_step(st.mc, st.stp, xtol, fatal);
        dec.action = ACT_NEWDIR;
        dec.stp = st.stp;
        return;
    }

    if (!mc_after_eval(st.mc, st.f, dots[D_GD], st.stp, xtol, fatal)) {
        mc_next_step(st.mc, st.stp, xtol, fatal);
        dec.action = ACT_LINESEARCH;
        dec.stp = st.stp;
        return;
    }

    // line search finished: s = stp*d, y = g - g_old stored at slot p (LBFGS.java:551-563).
    // Old pairs age by one (the oldest drops out when the history is full); the new pair is age 0.
    const int p = st.point;
    const double stp = st.stp;
    st.nfun += st.mc.nfev;
    const int nold = min(st.hist, m - 1);  // old pairs that survive
#pragma unroll
    for (int i = m - 1; i >= 1; --i) {
#pragma unroll
        for (int j = m - 1; j >= 1; --j) {
            st.SY[i][j] = st.SY[i - 1][j - 1];
            st.YY[i][j] = st.YY[i - 1][j - 1];
        }
        st.rho[i] = st.rho[i - 1];
    }
    double sg[m], yg[m];
#pragma unroll
    for (int a = 0; a < m - 1; ++a) {
        // old pair of (old) age a sits in slot (p - 1 - a) mod m and becomes age a + 1
        int slot = p - 1 - a;
        slot += (slot < 0) ? m : 0;
        const bool ok = a < nold;
        st.SY[0][a + 1] = ok ? stp * dots[D_DY + slot] : 0.0;   // s_new . y_old
        st.SY[a + 1][0] = ok ? dots[D_YNS + slot] : 0.0;        // s_old . y_new
        st.YY[0][a + 1] = st.YY[a + 1][0] = ok ? dots[D_YNY + slot] : 0.0;
        sg[a + 1] = ok ? dots[D_GS + slot] : 0.0;
        yg[a + 1] = ok ? dots[D_GY + slot] : 0.0;
    }
    st.SY[0][0] = stp * dots[D_DYN];
    st.YY[0][0] = dots[D_YNYN];
    sg[0] = stp * dots[D_GD];
    yg[0] = dots[D_GYN];
    st.hist = nold + 1;
    st.point = (p + 1 == m) ? 0 : p + 1;
    dec.slot = p;
    dec.hist = st.hist;
    dec.stp_accept = stp;

    // termination test (LBFGS.java:565-572)
    const double gnorm = sqrt(dots[D_GG]);
    const double xnorm = fmax(1.0, sqrt(dots[D_XX]));
    if (gnorm / xnorm <= eps) {
        dec.action = ACT_FINISH;
        return;
    }

    // next trip of the loop (LBFGS.java:407-506): H0 = ys/yy, two-loop recursion on scalars
    st.iter++;
    const double ys = st.SY[0][0], yy = st.YY[0][0];
    double diag = ys / yy;
    if (isnan(diag) || isinf(diag)) {  // safeDiv (LBFGS.java:804-814)
        if (ys == 0.0 && yy == 0.0)
            diag = 1.0;
        else
            fatal = 1;
    }
    st.rho[0] = 1.0 / ys;
    const int hist = st.hist;
    double alpha[m], beta[m];
    // first loop, newest -> oldest: q = -g - sum_{b<a} alpha_b y_b ; alpha_a = rho_a * (s_a . q)
#pragma unroll
    for (int a = 0; a < m; ++a) {
        double sq = -sg[a];
#pragma unroll
        for (int b = 0; b < a; ++b)
            sq -= alpha[b] * st.SY[a][b];
        alpha[a] = (a < hist) ? st.rho[a] * sq : 0.0;
    }
    // second loop, oldest -> newest: r = diag*q + sum_{b>a} beta_b s_b ; beta_a = alpha_a - rho_a (y_a . r)
#pragma unroll
    for (int a = m - 1; a >= 0; --a) {
        double yq = -yg[a];
#pragma unroll
        for (int b = 0; b < m; ++b)
            yq -= alpha[b] * st.YY[a][b];
        double yr = diag * yq;
#pragma unroll
        for (int b = m - 1; b > a; --b)
            yr += beta[b] * st.SY[b][a];
        beta[a] = (a < hist) ? alpha[a] - st.rho[a] * yr : 0.0;
    }
    dec.cg = -diag;
    double dginit = -diag * dots[D_GG];
    unsigned valid = 0;
#pragma unroll
    for (int a = 0; a < m; ++a) {
        int slot = st.point - 1 - a;
        slot += (slot < 0) ? m : 0;
        if (a < hist) {
            dec.cs[slot] = beta[a];
            dec.cy[slot] = -diag * alpha[a];
            dginit += beta[a] * sg[a] + (-diag * alpha[a]) * yg[a];
            valid |= 1u << slot;
        }
    }
    dec.valid_mask = valid;
    if (isnan(dginit) || isinf(dginit))
        fatal = 1;

    // LBFGS.java:505-527: nfev = 0, stp = 1, new line search
    st.stp = 0.0;
    set_stp(st.stp, 1.0, fatal);
    if (!mc_begin(st.mc, st.f, dginit, st.stp, xtol)) {
        dec.action = ACT_FAIL;
        return;
    }
    mc_next_step(st.mc, st.stp, xtol, fatal);
    dec.action = ACT_NEWDIR;
    dec.stp = st.stp;
}

// Solver state of one ADMM block carried from one x-update launch ("phase") of an ADMM iteration
// to the next: the launches re-partition the CTAs among the blocks that are still iterating.
struct BlockState {
    LbState st;
    int started;   // u-update done and first evaluation taken
    int finished;  // L-BFGS returned (iflag 0 or line-search failure)
    int status;    // 0 ok, 1 line search failed, 2 numeric trouble
    int pad;
};

// ------------------------------------------------------------------ group solver
// Workspace of one solve (global memory, length n each; S/Y hold kLbfgsM vectors).
struct SolveWs {
    int n;
    double *x;     // in: start point, out: result (trial point during the solve)
    double *g, *gold, *d, *wa, *S, *Y;
    double *spart; // [G][kNumDots]
    int G, cta;
    long long *prof; // optional (shared memory): ns in E, barrier1, R, barrier2, S, U, barrier3, evals, ...
    double *slice;   // optional shared memory, kSliceVectors * CAP doubles: the CTA's coordinate
                     // slice of x, d, g_old, wa, S, Y stays on chip for the whole launch
};

#ifndef SQW_AUG_FIRST
#define SQW_AUG_FIRST 1   // phase R: the augmented term's loads ahead of the partials' (see the loop)
#endif
#ifndef SQW_AUG_CACHE
#define SQW_AUG_CACHE 0   // keep the augmented term's inputs (sm_x, z, u_t of a coordinate's edges) on chip too:
                          // 6 of a coordinate's 24 loads per trip less; measured 59.8 vs 59.9 us per evaluation
                          // round at cfg-2, i.e. no gain for 15 KB of shared memory - left off
#endif
#if SQW_AUG_CACHE
constexpr int kSliceCap = 304;                       // coordinates per CTA the slice cache holds
constexpr int kAugVectors = 6;                       // (sm_x[par], z, u_t) of up to two edges per coordinate
#else
constexpr int kSliceCap = 320;
constexpr int kAugVectors = 0;
#endif
constexpr int kSliceVectors = 4 + 2 * kLbfgsM;       // x, d, g_old, wa, S[m], Y[m]
constexpr size_t kSolveSliceBytes = (size_t)(kSliceVectors + kAugVectors) * kSliceCap * sizeof(double);

struct SolveSmem {
    double dots[kNumDots];
    double red[kWarps];
    double wred[kWarps][kNumDots];  // per-warp totals of the 32 scalar partials
    Decision dec;
    LbState st;       // scalar solver state: shared memory keeps it out of every thread's registers
    int fatal;
};

// Sum 32 per-lane values across the warp for 32 different quantities at once: after the call
// lane l holds the warp total of v[l] in v[0]. A butterfly that halves the number of live values
// per stage: 31 shuffles instead of 32 x 5, in a fixed order.
__device__ __forceinline__ void warp_transpose_sum(double (&v)[kNumDots], int lane)
{
    static_assert(kNumDots == 32, "one quantity per lane");
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const double send = upper ? v[i] : v[i + off];
            const double keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// Problem concept:
//   double eval_partial(const double *x)   phase E, all threads; returns a per-thread value whose
//                                          sum over the CTA is the CTA's partial data term f
//   double aug_term(int k, double xk, double &faug)   part of the gradient of coordinate k that
//                                          depends only on x (xk = x[k]) and on data that is
//                                          constant during the solve; adds its objective part to faug
//   kBatch, data_issue(k, v), data_finish(k, v)   the part reduced over the group's phase-E
//                                          partials: issue starts the first batch of loads,
//                                          finish sums them (and any further batches)
// Runs to completion; returns the number of evaluations (identical in all CTAs). status: 0 ok,
// 1 line search failed (ExceptionWithIflag in the reference), 2 numeric trouble.
// eval_cap: stop (unfinished, state kept in sm.st) once that many evaluations have been made in
// total; resume: sm.st already holds the state of an interrupted solve. finished tells which.
//
// Latency notes (B200, cfg-2, measured with SQW_ADMM_PROFILE). A trip of the loop is
// E | barrier | R | barrier | S | U | barrier, and every phase but E is a handful of dependent L2
// round trips (~1 us each for lines another SM has just written). So
//  * the per-coordinate operands of R and U (x, d, g_old, wa, S, Y of the CTA's slice) live in
//    shared memory when the slice fits (ws.slice): global memory only receives the write-through
//    copy the next launch resumes from, and R's only global loads are the group's partials;
//  * U keeps its operands in registers from R;
//  * the 32 scalar partials are reduced by a transposing butterfly (31 shuffles per warp).
struct CoordRegs {
    double x, d, go, wa, gaug, faug;
    double s[kLbfgsM], y[kLbfgsM];
};

template <class Problem, int NT = kThreads, int CAP = kSliceCap>
__device__ int lbfgs_group_solve(Problem &prob, const SolveWs &ws, GroupBarrier &gb,
                                 SolveSmem &sm, double eps, double xtol, int &status,
                                 int eval_cap = 0x7fffffff, bool resume = false,
                                 bool *finished = nullptr)
{
    const int n = ws.n, G = ws.G, cta = ws.cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // coordinate slice of this CTA: whole 128-byte lines, so no line is shared between CTAs
    const int sl = ((n + G - 1) / G + 15) & ~15;
    const int k0 = min(n, cta * sl), k1 = min(n, k0 + sl);
    const int kf = k0 + (int)threadIdx.x;  // this thread's first coordinate
    const bool mine = kf < k1;
    LbState &st = sm.st;
    if (!resume) {
        for (int i = threadIdx.x; i < (int)(sizeof(LbState) / sizeof(int)); i += NT)
            reinterpret_cast<int *>(&st)[i] = 0;
    }
    // slice cache: thread t owns slot t of every vector (CAP <= NT: one coordinate each)
    static_assert(CAP <= NT, "one slice coordinate per thread");
    const bool cached = ws.slice != nullptr && k1 - k0 <= CAP;
    double *const cx = ws.slice + threadIdx.x;             // + v * CAP: vector v
    double *const cd = cx + CAP, *const cgo = cx + 2 * CAP, *const cwa = cx + 3 * CAP;
    double *const cS = cx + 4 * CAP, *const cY = cx + (4 + kLbfgsM) * CAP;
    if (cached && mine) {
        *cx = __ldcg(ws.x + kf);
        if (resume) {
            *cd = ws.d[kf];
            *cgo = ws.gold[kf];
            *cwa = ws.wa[kf];
#pragma unroll
            for (int i = 0; i < kLbfgsM; ++i) {
                cS[i * CAP] = ws.S[(size_t)i * n + kf];
                cY[i * CAP] = ws.Y[(size_t)i * n + kf];
            }
        }
    }
    // the inputs of the augmented term are constant during the solve: with SQW_AUG_CACHE they are
    // read once per launch; naug = number of cached edges of this thread's coordinate, -1 = not cached
    int naug = -1;
#if SQW_AUG_CACHE
    double *const caug = cx + kSliceVectors * CAP;
    if (cached && mine) {
        double a[6];
        int ne = 0;
        if (prob.aug_load(kf, a, ne)) {
#pragma unroll
            for (int j = 0; j < 6; ++j)
                caug[j * CAP] = a[j];
            naug = ne;
        }
    }
#endif
    __syncthreads();
    // history slots to reduce against (excludes the slot being replaced) / slot y_new is written to
    int point = resume ? st.point : 0;
    unsigned valid = resume ? (hist_mask(st.point, st.hist) & ~(1u << st.point)) : 0u;
    bool first = !resume;
    bool done = false;
    status = 0;

    // operands of coordinate k for phase R (and U): everything but the phase-E partials
    auto load_coord = [&](int k, CoordRegs &c, bool from_cache, bool with_aug = true) {
        c.d = c.go = c.wa = 0.0;
#pragma unroll
        for (int i = 0; i < kLbfgsM; ++i)
            c.s[i] = c.y[i] = 0.0;
        if (from_cache) {
            c.x = *cx;
            if (!first) {
                c.d = *cd;
                c.go = *cgo;
                c.wa = *cwa;
#pragma unroll
                for (int i = 0; i < kLbfgsM; ++i) {
                    if ((valid >> i) & 1u) {
                        c.s[i] = cS[i * CAP];
                        c.y[i] = cY[i * CAP];
                    }
                }
            }
        } else {
            c.x = __ldcg(ws.x + k);
            if (!first) {
                c.d = ws.d[k];
                c.go = ws.gold[k];
                c.wa = ws.wa[k];
#pragma unroll
                for (int i = 0; i < kLbfgsM; ++i) {
                    if ((valid >> i) & 1u) {
                        c.s[i] = ws.S[(size_t)i * n + k];
                        c.y[i] = ws.Y[(size_t)i * n + k];
                    }
                }
            }
        }
        c.faug = 0.0;
#if SQW_AUG_CACHE
        if (from_cache && naug >= 0) {
            double a[6];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                a[j] = caug[j * CAP];
            c.gaug = prob.aug_term_cached(c.x, a, naug, c.faug);
            return;
        }
#endif
        c.gaug = with_aug ? prob.aug_term(k, c.x, c.faug) : 0.0;
    };

    for (int trip = 0;; ++trip) {
        // ---- phase E: data-term partials of this CTA at the trial point
        long long tk0 = 0, tk1 = 0, tk2 = 0;
        if (ws.prof && threadIdx.x == 0)
            tk0 = sqw_now_ns();
#define SQW_PROF(slot)                                   \
    if (ws.prof && threadIdx.x == 0) {                   \
        tk1 = sqw_now_ns();                              \
        ws.prof[slot] += tk1 - tk0;                      \
        if (trip == 0)                                   \
            ws.prof[24 + slot] += tk1 - tk0;             \
        tk0 = tk1;                                       \
    }
#define SQW_PROF2(slot)                                  \
    if (ws.prof && threadIdx.x == 0) {                   \
        const long long t2 = sqw_now_ns();               \
        ws.prof[16 + slot] += t2 - tk2;                  \
        tk2 = t2;                                        \
    }
        const double fpart = prob.eval_partial(ws.x);
        SQW_PROF(0)
        gb.sync();
        SQW_PROF(1)

        // ---- phase R: one pass over my coordinate slice produces the reduced gradient, y_new
        // and every scalar partial
        double pd[kNumDots];
#pragma unroll
        for (int i = 0; i < kNumDots; ++i)
            pd[i] = 0.0;
        double *ynew = ws.Y + (size_t)point * n;
        auto reduce_coord = [&](int k, const CoordRegs &c, double gdata, double &gk_out,
                                double &yn_out) {
            const double gk = gdata + c.gaug;
            pd[D_F] += c.faug;
            pd[D_GG] += gk * gk;
            pd[D_XX] += c.x * c.x;
            double yn = 0.0;
            if (!first) {
                yn = gk - c.go;
                ynew[k] = yn;
                pd[D_GD] += gk * c.d;
                pd[D_GYN] += gk * yn;
                pd[D_YNYN] += yn * yn;
                pd[D_DYN] += c.d * yn;
#pragma unroll
                for (int i = 0; i < kLbfgsM; ++i) {
                    if ((valid >> i) & 1u) {
                        pd[D_GS + i] += gk * c.s[i];
                        pd[D_GY + i] += gk * c.y[i];
                        pd[D_YNS + i] += yn * c.s[i];
                        pd[D_YNY + i] += yn * c.y[i];
                        pd[D_DY + i] += c.d * c.y[i];
                    }
                }
            }
            gk_out = gk;
            yn_out = yn;
        };
        // the loads of the partials are issued first, the coordinate's other operands behind them
        CoordRegs c0;
        double g0 = 0.0, yn0 = 0.0;
        tk2 = tk0;
        if (mine) {
            double v[Problem::kBatch];
#if SQW_AUG_FIRST
            // The augmented term reads (sm_x[par], z, u_t) of the coordinate's one or two edges. Left
            // inside load_coord, ptxas sinks these loads behind the wait for the partials and issues
            // them in three dependent groups (sm_x | z, u | z', u': ncu source page).
            // Fetched here, ahead of the partials, they share the partials' round trip; same
            // operations in the same order. Measured A/B at cfg-2: 2.838 -> 2.822 s (sm_x and z hit
            // L1, so only the two u_t round trips were saved); a 49-CTA handle 6.20 -> 6.24 s.
            double aug_in[6];
            int aug_ne = 0;
            const bool aug_split = prob.aug_load(kf, aug_in, aug_ne);
#endif
            prob.data_issue(kf, v);
#if SQW_PROF_SPLIT
            // probe build: serialise the two groups of loads and stamp after each has arrived
            const double dsum = prob.data_finish(kf, v);
            if (dsum == 1.2345e300)
                tk2 += 1;
            SQW_PROF2(6)
            load_coord(kf, c0, cached);
            if (c0.gaug == 1.2345e300)
                tk2 += 1;
            SQW_PROF2(7)
            reduce_coord(kf, c0, dsum, g0, yn0);
#elif SQW_AUG_FIRST
            // (the loads of the augmented term were issued ahead of the partials', see above)
            load_coord(kf, c0, cached, !aug_split);
            if (aug_split)
                c0.gaug = prob.aug_term_cached(c0.x, aug_in, aug_ne, c0.faug);
            reduce_coord(kf, c0, prob.data_finish(kf, v), g0, yn0);
#else
            load_coord(kf, c0, cached);
            reduce_coord(kf, c0, prob.data_finish(kf, v), g0, yn0);
#endif
            if (!cached)
                ws.g[kf] = g0;
        }
        if (G <= Problem::kSmallBatch) {
            // (a small group = a long slice: 2 - 3 coordinates per thread at 6 CTAs per block)
            for (int k = kf + NT; k < k1; k += NT) {
                CoordRegs c;
                double v[Problem::kSmallBatch], gk, yn;
                prob.data_issue_small(k, v);
                load_coord(k, c, false);
                reduce_coord(k, c, prob.data_finish_small(k, v), gk, yn);
                ws.g[k] = gk;
            }
        } else {
            for (int k = kf + NT; k < k1; k += NT) {  // (never with the slice cache)
                CoordRegs c;
                double v[Problem::kBatch], gk, yn;
                prob.data_issue(k, v);
                load_coord(k, c, false);
                reduce_coord(k, c, prob.data_finish(k, v), gk, yn);
                ws.g[k] = gk;
            }
        }
        pd[D_F] += fpart;
        SQW_PROF2(0)
        warp_transpose_sum(pd, lane);
        sm.wred[warp][lane] = pd[0];
        SQW_PROF2(1)
        __syncthreads();
        SQW_PROF2(2)
        if (warp == 0) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w)
                acc += sm.wred[w][lane];
            __stcg(ws.spart + (size_t)cta * kNumDots + lane, acc);
        }
        SQW_PROF2(3)
        SQW_PROF(2)
        gb.sync();
        SQW_PROF(3)
        tk2 = tk0;

        // ---- phase S: every CTA reduces the G partials identically and takes the same decision
        if (warp == 0) {
            double s = 0.0;
            for (int cb = 0; cb < G; cb += 32) {
                double v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = (cb + j < G) ? __ldcg(ws.spart + (size_t)(cb + j) * kNumDots + lane) : 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    s += v[j];
            }
            sm.dots[lane] = s;
            const int bad = __any_sync(0xffffffffu, isnan(s) || isinf(s));
            __syncwarp();
            SQW_PROF2(4)
            if (lane == 0) {
                int fatal = bad;
                if (!fatal)
                    lbfgs_scalar_step(st, sm.dots, eps, xtol, sm.dec, fatal, first);
                sm.fatal = fatal;
            }
            SQW_PROF2(5)
        }
        __syncthreads();
        SQW_PROF(4)
        const Decision &dec = sm.dec;
        if (sm.fatal) {
            status = 2;
            done = true;
            break;
        }
        if (dec.action == ACT_FINISH || dec.action == ACT_FAIL) {
            status = (dec.action == ACT_FAIL) ? 1 : 0;
            done = true;
            break;
        }
        if (dec.action == ACT_LINESEARCH) {
            if (mine) {
                const double xn = c0.wa + dec.stp * c0.d;
                ws.x[kf] = xn;
                if (cached)
                    *cx = xn;
            }
            for (int k = kf + NT; k < k1; k += NT)
                ws.x[k] = ws.wa[k] + dec.stp * ws.d[k];
        } else {  // ACT_NEWDIR
            if (mine) {
                // operands still in registers from phase R; the pair being stored at dec.slot is
                // (stp_accept * d, y_new)
                const double snew = dec.stp_accept * c0.d;
                if (dec.slot >= 0) {
                    ws.S[(size_t)dec.slot * n + kf] = snew;
                    if (cached) {
                        cS[dec.slot * CAP] = snew;
                        cY[dec.slot * CAP] = yn0;
                    }
                }
                double dn = dec.cg * g0;
#pragma unroll
                for (int i = 0; i < kLbfgsM; ++i)
                    if ((dec.valid_mask >> i) & 1u) {
                        const double si = (i == dec.slot) ? snew : c0.s[i];
                        const double yi = (i == dec.slot) ? yn0 : c0.y[i];
                        dn += dec.cs[i] * si + dec.cy[i] * yi;
                    }
                const double xn = c0.x + dec.stp * dn;
                ws.gold[kf] = g0;
                ws.wa[kf] = c0.x;
                ws.d[kf] = dn;
                ws.x[kf] = xn;
                if (cached) {
                    *cgo = g0;
                    *cwa = c0.x;
                    *cd = dn;
                    *cx = xn;
                }
            }
            for (int k = kf + NT; k < k1; k += NT) {
                const double gk = ws.g[k];
                if (dec.slot >= 0)
                    ws.S[(size_t)dec.slot * n + k] = dec.stp_accept * ws.d[k];
                double dn = dec.cg * gk;
                for (int i = 0; i < kLbfgsM; ++i)
                    if ((dec.valid_mask >> i) & 1u)
                        dn += dec.cs[i] * ws.S[(size_t)i * n + k] + dec.cy[i] * ws.Y[(size_t)i * n + k];
                const double xk = ws.x[k];
                ws.gold[k] = gk;
                ws.wa[k] = xk;
                ws.d[k] = dn;
                ws.x[k] = xk + dec.stp * dn;
            }
            // slots for the next reduce: valid pairs except the one the next y_new overwrites
            point = (dec.slot < 0) ? 0 : (dec.slot + 1) % kLbfgsM;
            valid = dec.valid_mask & ~(1u << point);
        }
        first = false;
        SQW_PROF(5)
        gb.sync();
        SQW_PROF(6)
        if (ws.prof && threadIdx.x == 0) {
            ws.prof[7] += 1;
            if (trip == 0)
                ws.prof[31] += 1;
        }
        if (st.evals >= eval_cap)  // identical in every CTA; the trial point for the next
            break;                 // evaluation is already in place
    }
#undef SQW_PROF
#undef SQW_PROF2
    if (finished)
        *finished = done;
    int evals = 0;
    if (threadIdx.x == 0)
        sm.dec.hist = st.evals;
    __syncthreads();
    evals = sm.dec.hist;
    __syncthreads();
    return evals;
}

}  // namespace sqw
