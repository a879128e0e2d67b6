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
    __device__ __forceinline__ void sync()
    {
        target += (unsigned)G;
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < target);
        }
        __syncthreads();
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

// Scalar L-BFGS state (identical in every CTA of the group).
struct LbState {
    int iter, point, hist, nfun, evals;
    double stp, f;
    McState mc;
    double SY[kLbfgsM][kLbfgsM];  // s_i . y_j
    double YY[kLbfgsM][kLbfgsM];  // y_i . y_j
    double sg[kLbfgsM], yg[kLbfgsM], rho[kLbfgsM];
};

// What thread 0 tells the rest of the CTA after the scalar step.
struct Decision {
    int action;
    int slot;            // history slot written on ACT_NEWDIR
    int hist;            // number of valid pairs after the update
    double stp_accept;   // step that scales d into s_new
    double stp;          // trial step for the next evaluation
    double cg;           // d_new = cg * g + sum cs[i] * s_i + sum cy[i] * y_i
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

// The scalar step after an evaluation whose reduced scalars are dots[]. Mirrors the tail of
// LBFGS.lbfgs (LBFGS.java:531-590) followed by the head of its next loop trip (:407-527).
__device__ inline void lbfgs_scalar_step(LbState &st, const double *dots, double eps, double xtol,
                                         Decision &dec, int &fatal, bool first)
{
    const int m = kLbfgsM;
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

    // line search finished: s = stp*d, y = g - g_old stored at slot p (LBFGS.java:551-563)
    const int p = st.point;
    const double stp = st.stp;
    const unsigned old = hist_mask(st.point, st.hist) & ~(1u << p);
    st.nfun += st.mc.nfev;
    for (int i = 0; i < m; ++i) {
        if (!((old >> i) & 1u))
            continue;
        st.SY[p][i] = stp * dots[D_DY + i];
        st.SY[i][p] = dots[D_YNS + i];
        st.YY[p][i] = st.YY[i][p] = dots[D_YNY + i];
        st.sg[i] = dots[D_GS + i];
        st.yg[i] = dots[D_GY + i];
    }
    st.SY[p][p] = stp * dots[D_DYN];
    st.YY[p][p] = dots[D_YNYN];
    st.sg[p] = stp * dots[D_GD];
    st.yg[p] = dots[D_GYN];
    st.hist = min(st.hist + 1, m);
    st.point = (p + 1) % m;
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
    const double ys = st.SY[p][p], yy = st.YY[p][p];
    double diag = ys / yy;
    if (isnan(diag) || isinf(diag)) {  // safeDiv (LBFGS.java:804-814)
        if (ys == 0.0 && yy == 0.0)
            diag = 1.0;
        else
            fatal = 1;
    }
    st.rho[p] = 1.0 / ys;
    const unsigned valid = hist_mask(st.point, st.hist);
    double alpha[kLbfgsM], beta[kLbfgsM];
    for (int i = 0; i < m; ++i)
        alpha[i] = beta[i] = 0.0;
    // first loop: newest -> oldest; q = -g - sum_{processed} alpha_l y_l
    for (int j = 0; j < st.hist; ++j) {
        const int cp = (st.point - 1 - j + 2 * m) % m;
        double sq = -st.sg[cp];
        for (int jj = 0; jj < j; ++jj) {
            const int l = (st.point - 1 - jj + 2 * m) % m;
            sq -= alpha[l] * st.SY[cp][l];
        }
        alpha[cp] = st.rho[cp] * sq;
    }
    // second loop: oldest -> newest; r = diag*q + sum_{processed} beta_l s_l
    for (int j = st.hist - 1; j >= 0; --j) {
        const int cp = (st.point - 1 - j + 2 * m) % m;
        double yq = -st.yg[cp];
        for (int jj = 0; jj < st.hist; ++jj) {
            const int l = (st.point - 1 - jj + 2 * m) % m;
            yq -= alpha[l] * st.YY[cp][l];
        }
        double yr = diag * yq;
        for (int jj = st.hist - 1; jj > j; --jj) {
            const int l = (st.point - 1 - jj + 2 * m) % m;
            yr += beta[l] * st.SY[l][cp];
        }
        beta[cp] = alpha[cp] - st.rho[cp] * yr;
    }
    dec.cg = -diag;
    double dginit = -diag * dots[D_GG];
    for (int i = 0; i < m; ++i) {
        dec.cs[i] = beta[i];
        dec.cy[i] = -diag * alpha[i];
        if ((valid >> i) & 1u)
            dginit += dec.cs[i] * st.sg[i] + dec.cy[i] * st.yg[i];
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

// ------------------------------------------------------------------ group solver
// Workspace of one solve (all global memory, length n each; S/Y hold kLbfgsM vectors).
struct SolveWs {
    int n;
    double *x;     // in: start point, out: result (trial point during the solve)
    double *g, *gold, *d, *wa, *S, *Y;
    double *spart; // [G][kNumDots]
    int G, cta;
};

struct SolveSmem {
    double dots[kNumDots];
    double red[kWarps];
    Decision dec;
    int fatal;
};

// Problem concept:
//   double eval_partial(const double *x)        phase E, all threads; returns the CTA's partial f
//   double grad_coord(int k, const double *x, double &faug)   reduced gradient of coordinate k
// Runs to completion; returns the number of evaluations (identical in all CTAs). status: 0 ok,
// 1 line search failed (ExceptionWithIflag in the reference), 2 numeric trouble.
template <class Problem>
__device__ int lbfgs_group_solve(Problem &prob, const SolveWs &ws, GroupBarrier &gb,
                                 SolveSmem &sm, double eps, double xtol, int &status)
{
    const int n = ws.n, G = ws.G, cta = ws.cta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // coordinate slice of this CTA: whole 128-byte lines, so no line is shared between CTAs
    const int sl = ((n + G - 1) / G + 15) & ~15;
    const int k0 = min(n, cta * sl), k1 = min(n, k0 + sl);
    LbState st;  // meaningful in thread 0 only
    st.evals = 0;
    st.point = 0;
    st.hist = 0;
    unsigned valid = 0;   // history slots to reduce against (excludes the slot being replaced)
    int point = 0;        // slot y_new is written to
    bool first = true;
    status = 0;

    for (;;) {
        // ---- phase E: data-term partials of this CTA at the trial point
        const double fpart = prob.eval_partial(ws.x);
        gb.sync();

        // ---- phase R: reduced gradient of my slice, y_new, and all scalar partials
        double faug = 0.0;
        double *ynew = ws.Y + (size_t)point * n;
        for (int k = k0 + threadIdx.x; k < k1; k += kThreads) {
            const double gk = prob.grad_coord(k, ws.x, faug);
            ws.g[k] = gk;
            if (!first)
                ynew[k] = gk - ws.gold[k];
        }
        const double faug_sum = block_sum(faug, sm.red);  // also orders g / y_new stores
        for (int id = warp; id < kNumDots; id += kWarps) {
            const double *a = nullptr, *b = nullptr;
            bool need = true;
            if (id == D_F) {
                need = false;
            } else if (id == D_GG) {
                a = ws.g, b = ws.g;
            } else if (id == D_XX) {
                a = ws.x, b = ws.x;
            } else if (first) {
                need = false;
            } else if (id == D_GD) {
                a = ws.g, b = ws.d;
            } else if (id == D_GYN) {
                a = ws.g, b = ynew;
            } else if (id == D_YNYN) {
                a = ynew, b = ynew;
            } else if (id == D_DYN) {
                a = ws.d, b = ynew;
            } else {
                const int i = (id - D_GS) % kLbfgsM, kind = (id - D_GS) / kLbfgsM;
                need = (valid >> i) & 1u;
                const double *si = ws.S + (size_t)i * n, *yi = ws.Y + (size_t)i * n;
                a = (kind <= 1) ? ws.g : (kind <= 3 ? ynew : ws.d);
                b = (kind == 0 || kind == 2) ? si : yi;
            }
            double acc = 0.0;
            if (need)
                for (int k = k0 + lane; k < k1; k += 32)
                    acc += a[k] * b[k];
            acc = warp_sum(acc);
            if (lane == 0)
                __stcg(ws.spart + (size_t)cta * kNumDots + id,
                       id == D_F ? fpart + faug_sum : acc);
        }
        gb.sync();

        // ---- phase S: every CTA reduces the G partials identically and takes the same decision
        if (warp == 0) {
            double s = 0.0;
            for (int c = 0; c < G; ++c)
                s += __ldcg(ws.spart + (size_t)c * kNumDots + lane);
            sm.dots[lane] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int fatal = 0;
            for (int i = 0; i < kNumDots; ++i)
                if (isnan(sm.dots[i]) || isinf(sm.dots[i]))
                    fatal = 1;
            if (!fatal)
                lbfgs_scalar_step(st, sm.dots, eps, xtol, sm.dec, fatal, first);
            sm.fatal = fatal;
        }
        __syncthreads();
        const Decision &dec = sm.dec;
        if (sm.fatal) {
            status = 2;
            break;
        }
        if (dec.action == ACT_FINISH || dec.action == ACT_FAIL) {
            status = (dec.action == ACT_FAIL) ? 1 : 0;
            break;
        }
        if (dec.action == ACT_LINESEARCH) {
            for (int k = k0 + threadIdx.x; k < k1; k += kThreads)
                ws.x[k] = ws.wa[k] + dec.stp * ws.d[k];
        } else {  // ACT_NEWDIR
            for (int k = k0 + threadIdx.x; k < k1; k += kThreads) {
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
        gb.sync();
    }
    int evals = 0;
    if (threadIdx.x == 0)
        sm.dec.hist = st.evals;
    __syncthreads();
    evals = sm.dec.hist;
    __syncthreads();
    return evals;
}

}  // namespace sqw
