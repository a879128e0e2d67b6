/*
 * lbfgs_oracle.c — CPU restatement of framework/LBFGS.java:338-592,720-814 and
 * framework/Mcsrch.java:173-369,442-631 (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Reverse-communication Nocedal L-BFGS with the Moré–Thuente line search, including the
 * reference's local modifications, which are part of the numerics:
 *   stpmin = 1e-5 (LBFGS.java:125), stpmax = 1e20 (:134), gtol = 0.9 (:116), ftol = 1e-4 (:398),
 *   maxfev = 200 (:136), TWO_THIRDS = 0.66 (Mcsrch.java:38), steps clipped to <= 1000
 *   (Mcsrch.java:621-631), every line-search termination code 2..6 rewritten to 1
 *   (Mcsrch.java:313-316).
 * Only the diagco == false path is restated: LOne always passes false (LOne.java:880).
 * The MY_MAX_STEP block (LBFGS.java:511-521) computes a dot product and then only ever assigns
 * stpmax = 1e20, its initial value; it is omitted.
 * Java exceptions become return codes: ORC_LB_IFLAG_EXC for ExceptionWithIflag (caught and
 * ignored by LOne.java:881-883), ORC_LB_FATAL for IllegalArgumentException (NaN/Inf guards,
 * uncaught in the reference).
 */
#include "oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double LB_GTOL = 0.9;     /* LBFGS.java:116 */
static const double LB_STPMIN = 1e-5;  /* LBFGS.java:125 */
static const double LB_STPMAX = 1e20;  /* LBFGS.java:134 */
static const double MC_ONE_HALF = 0.5, MC_TWO_THIRDS = 0.66, MC_FOUR = 4; /* Mcsrch.java:38 */

static int bad_num(double v) { return isnan(v) || isinf(v); }        /* LBFGS.java:788-790 */

/* LBFGS.java:792-802 */
static double safe_mult(double a, double b, int *fatal)
{
    double r = a * b;
    if (bad_num(r)) {
        if (a == 0 || b == 0)
            r = 0;
        else
            *fatal = 1;
    }
    return r;
}

/* LBFGS.java:804-814 */
static double safe_div(double a, double b, int *fatal)
{
    double r = a / b;
    if (bad_num(r)) {
        if (a == 0 && b == 0)
            r = 1;
        else
            *fatal = 1;
    }
    return r;
}

/* LBFGS.java:776-786 */
static double ddot(int n, const double *dx, const double *dy, int *fatal)
{
    double result = 0;
    for (int i = 0; i < n; i++)
        result += safe_mult(dx[i], dy[i], fatal);
    return result;
}

/* LBFGS.java:720-745 */
static void daxpy(int n, double da, const double *dx, double *dy, int *fatal)
{
    if (n <= 0)
        return;
    if (da == 0)
        return;
    for (int i = 0; i < n; i++)
        dy[i] += safe_mult(da, dx[i], fatal);
}

/* Mcsrch.java:621-631 */
static void set_stp(double *stp, double value, int *fatal)
{
    if (value < 0) {
        *fatal = 1;
        return;
    }
    if (value > 1000)
        *stp = 1000;
    else
        *stp = value;
}

static double max3(double x, double y, double z)                      /* Mcsrch.java:44-46 */
{
    return x < y ? (y < z ? z : y) : (x < z ? z : x);
}

static double sqr(double x) { return x * x; }

/* Mcsrch.java:442-619. Operates on the (st,f,d) triples passed by pointer. */
static void mcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy,
                   double *stp, double fp, double dp, int *brackt, double stpmin, double stpmax,
                   int *info, int *fatal)
{
    int bound;
    double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;

    *info = 0;
    if ((*brackt && (*stp <= fmin(*stx, *sty) || *stp >= fmax(*stx, *sty))) ||
        *dx * (*stp - *stx) >= 0.0 || stpmax < stpmin)
        return;

    sgnd = dp * (*dx / fabs(*dx));

    if (fp > *fx) {
        /* case 1: higher function value; minimum bracketed (:530-553) */
        *info = 1;
        bound = 1;
        theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = max3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt(sqr(theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx)
            gamma = -gamma;
        p = (gamma - *dx) + theta;
        q = ((gamma - *dx) + gamma) + dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2) * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2;
        *brackt = 1;
    } else if (sgnd < 0.0) {
        /* case 2: lower value, derivatives of opposite sign (:554-576) */
        *info = 2;
        bound = 0;
        theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = max3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt(sqr(theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx)
            gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp))
            stpf = stpc;
        else
            stpf = stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) {
        /* case 3: lower value, same sign, derivative magnitude decreases (:577-620) */
        *info = 3;
        bound = 1;
        theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = max3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt(fmax(0, sqr(theta / s) - (*dx / s) * (dp / s)));
        if (*stp > *stx)
            gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (*dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            if (fabs(*stp - stpc) < fabs(*stp - stpq))
                stpf = stpc;
            else
                stpf = stpq;
        } else {
            if (fabs(*stp - stpc) > fabs(*stp - stpq))
                stpf = stpc;
            else
                stpf = stpq;
        }
    } else {
        /* case 4: lower value, same sign, derivative magnitude does not decrease (:621-646) */
        *info = 4;
        bound = 0;
        if (*brackt) {
            theta = 3 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = max3(fabs(theta), fabs(*dy), fabs(dp));
            gamma = s * sqrt(sqr(theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty)
                gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }

    /* interval update (:651-664) */
    if (fp > *fx) {
        *sty = *stp;
        *fy = fp;
        *dy = dp;
    } else {
        if (sgnd < 0.0) {
            *sty = *stx;
            *fy = *fx;
            *dy = *dx;
        }
        *stx = *stp;
        *fx = fp;
        *dx = dp;
    }

    /* new step, safeguarded (:668-680) */
    stpf = fmin(stpmax, stpf);
    stpf = fmax(stpmin, stpf);
    set_stp(stp, stpf, fatal);

    if (*brackt && bound) {
        double possibleStep = *stx + MC_TWO_THIRDS * (*sty - *stx);
        if (*sty > *stx)
            set_stp(stp, fmin(possibleStep, *stp), fatal);
        else
            set_stp(stp, fmax(possibleStep, *stp), fatal);
    }
}

/* Mcsrch.java:173-369. s = search direction (w + is0 in the reference), wa = diag. */
static void mcsrch(orc_mcsrch *mc, int n, double *x, double f, const double *g, const double *s,
                   double *stp, double ftol, double xtol, int maxfev, int *info, int *nfev,
                   double *wa, int *fatal)
{
    if (*info != -1) {
        mc->infoc = 1;
        if (n <= 0 || *stp <= 0 || ftol < 0 || LB_GTOL < 0 || xtol < 0 || LB_STPMIN < 0 ||
            LB_STPMAX < LB_STPMIN || maxfev <= 0)
            return;

        mc->dginit = 0;
        for (int j = 0; j < n; j++)
            mc->dginit += g[j] * s[j];
        if (mc->dginit >= 0)
            return; /* "The search direction is not a descent direction." info stays 0 */

        mc->brackt = 0;
        mc->stage1 = 1;
        *nfev = 0;
        mc->finit = f;
        mc->dgtest = ftol * mc->dginit;
        mc->width = LB_STPMAX - LB_STPMIN;
        mc->width1 = mc->width / MC_ONE_HALF;
        for (int j = 0; j < n; j++)
            wa[j] = x[j];

        mc->stx = 0;
        mc->fx = mc->finit;
        mc->dgx = mc->dginit;
        mc->sty = 0;
        mc->fy = mc->finit;
        mc->dgy = mc->dginit;
    }

    for (;;) {
        if (*info != -1) {
            if (mc->brackt) {
                mc->stmin = fmin(mc->stx, mc->sty);
                mc->stmax = fmax(mc->stx, mc->sty);
            } else {
                mc->stmin = mc->stx;
                mc->stmax = *stp + MC_FOUR * (*stp - mc->stx);
            }

            set_stp(stp, fmax(*stp, LB_STPMIN), fatal);
            set_stp(stp, fmin(*stp, LB_STPMAX), fatal);

            if ((mc->brackt && (*stp <= mc->stmin || *stp >= mc->stmax ||
                                mc->stmax - mc->stmin <= xtol * mc->stmax)) ||
                *nfev >= maxfev - 1 || mc->infoc == 0)
                set_stp(stp, mc->stx, fatal);

            for (int j = 0; j < n; j++)
                x[j] = wa[j] + *stp * s[j];

            *info = -1;
            return;
        }

        *info = 0;
        *nfev = *nfev + 1;
        mc->dg = 0;
        for (int j = 0; j < n; j++)
            mc->dg = mc->dg + g[j] * s[j];

        mc->ftest1 = mc->finit + *stp * mc->dgtest;

        if ((mc->brackt && (*stp <= mc->stmin || *stp >= mc->stmax)) || mc->infoc == 0)
            *info = 6;
        if (*stp == LB_STPMAX && f <= mc->ftest1 && mc->dg <= mc->dgtest)
            *info = 5;
        if (*stp == LB_STPMIN && (f > mc->ftest1 || mc->dg >= mc->dgtest))
            *info = 4;
        if (*nfev >= maxfev)
            *info = 3;
        if (mc->brackt && mc->stmax - mc->stmin <= xtol * mc->stmax)
            *info = 2;
        if (f <= mc->ftest1 && fabs(mc->dg) <= LB_GTOL * (-mc->dginit))
            *info = 1;

        if (*info != 0) {
            *info = 1; /* Mcsrch.java:313-316: every termination reported as success */
            return;
        }

        if (mc->stage1 && f <= mc->ftest1 && mc->dg >= fmin(ftol, LB_GTOL) * mc->dginit)
            mc->stage1 = 0;

        if (mc->stage1 && f <= mc->fx && f > mc->ftest1) {
            mc->fm = f - *stp * mc->dgtest;
            mc->fxm = mc->fx - mc->stx * mc->dgtest;
            mc->fym = mc->fy - mc->sty * mc->dgtest;
            mc->dgm = mc->dg - mc->dgtest;
            mc->dgxm = mc->dgx - mc->dgtest;
            mc->dgym = mc->dgy - mc->dgtest;

            mcstep(&mc->stx, &mc->fxm, &mc->dgxm, &mc->sty, &mc->fym, &mc->dgym, stp, mc->fm,
                   mc->dgm, &mc->brackt, mc->stmin, mc->stmax, &mc->infoc, fatal);

            mc->fx = mc->fxm + mc->stx * mc->dgtest;
            mc->fy = mc->fym + mc->sty * mc->dgtest;
            mc->dgx = mc->dgxm + mc->dgtest;
            mc->dgy = mc->dgym + mc->dgtest;
        } else {
            mcstep(&mc->stx, &mc->fx, &mc->dgx, &mc->sty, &mc->fy, &mc->dgy, stp, f, mc->dg,
                   &mc->brackt, mc->stmin, mc->stmax, &mc->infoc, fatal);
        }

        if (mc->brackt) {
            if (fabs(mc->sty - mc->stx) >= MC_TWO_THIRDS * mc->width1)
                set_stp(stp, (mc->sty + mc->stx) / 2.0, fatal);
            mc->width1 = mc->width;
            mc->width = fabs(mc->sty - mc->stx);
        }
    }
}

void orc_lbfgs_init(orc_lbfgs *lb, int n, int m)
{
    memset(lb, 0, sizeof(*lb));
    lb->n = n;
    lb->m = m;
    lb->maxfev = 200;                                   /* LBFGS.java:136 */
    lb->w = (double *)calloc((size_t)n * (2 * m + 1) + 2 * m, sizeof(double));
}

void orc_lbfgs_free(orc_lbfgs *lb)
{
    free(lb->w);
    lb->w = NULL;
}

/* LBFGS.java:338-592 with diagco == false. */
int orc_lbfgs_step(orc_lbfgs *lb, double *x, double f, const double *g, double *diag, double eps,
                   double xtol, int *iflag)
{
    const int n = lb->n, m = lb->m;
    double *w = lb->w;
    int execute_entire_while_loop = 0;
    int fatal = 0;

    if (*iflag == 0) {
        lb->iter = 0;
        if (n <= 0 || m <= 0) {
            *iflag = -3;
            return ORC_LB_IFLAG_EXC;
        }
        lb->nfun = 1;
        lb->point = 0;
        lb->finish = 0;
        for (int i = 0; i < n; i++)
            diag[i] = 1;
        lb->ispt = n + 2 * m;
        lb->iypt = lb->ispt + n * m;
        for (int i = 0; i < n; i++)
            w[lb->ispt + i] = -g[i] * diag[i];
        lb->gnorm = sqrt(ddot(n, g, g, &fatal));
        lb->stp1 = 1 / lb->gnorm;
        lb->ftol = 0.0001;
        execute_entire_while_loop = 1;
    }

    for (;;) {
        if (execute_entire_while_loop) {
            lb->iter = lb->iter + 1;
            lb->info = 0;
            lb->bound = lb->iter - 1;
            if (lb->iter != 1) {
                if (lb->iter > m)
                    lb->bound = m;
                lb->ys = ddot(n, w + lb->iypt + lb->npt, w + lb->ispt + lb->npt, &fatal);
                lb->yy = ddot(n, w + lb->iypt + lb->npt, w + lb->iypt + lb->npt, &fatal);
                {
                    double d = safe_div(lb->ys, lb->yy, &fatal);
                    for (int i = 0; i < n; i++)
                        diag[i] = d;
                }
            }
        }

        if (execute_entire_while_loop || *iflag == 2) {
            if (lb->iter != 1) {
                lb->cp = lb->point;
                if (lb->point == 0)
                    lb->cp = m;
                w[n + lb->cp - 1] = 1 / lb->ys;

                for (int i = 0; i < n; i++) {
                    w[i] = -g[i];
                    if (bad_num(w[i]))
                        fatal = 1;
                }

                lb->cp = lb->point;
                for (int i = 1; i <= lb->bound; i++) {
                    lb->cp = lb->cp - 1;
                    if (lb->cp == -1)
                        lb->cp = m - 1;
                    lb->sq = ddot(n, w + lb->ispt + lb->cp * n, w, &fatal);
                    lb->inmc = n + m + lb->cp + 1;
                    lb->iycn = lb->iypt + lb->cp * n;
                    w[lb->inmc - 1] = w[n + lb->cp] * lb->sq;
                    daxpy(n, -w[lb->inmc - 1], w + lb->iycn, w, &fatal);
                }

                for (int i = 0; i < n; i++)
                    w[i] *= diag[i];

                for (int i = 1; i <= lb->bound; i++) {
                    lb->yr = ddot(n, w + lb->iypt + lb->cp * n, w, &fatal);
                    lb->beta = safe_mult(w[n + lb->cp], lb->yr, &fatal);
                    lb->inmc = n + m + lb->cp + 1;
                    lb->beta = w[lb->inmc - 1] - lb->beta;
                    if (bad_num(lb->beta))
                        fatal = 1;
                    lb->iscn = lb->ispt + lb->cp * n;
                    daxpy(n, lb->beta, w + lb->iscn, w, &fatal);
                    lb->cp = lb->cp + 1;
                    if (lb->cp == m)
                        lb->cp = 0;
                }

                for (int i = 0; i < n; i++)
                    w[lb->ispt + lb->point * n + i] = w[i];
            }

            lb->nfev = 0;
            set_stp(&lb->stp, (lb->iter == 1 && lb->stp1 < 1) ? lb->stp1 : 1, &fatal);

            for (int i = 0; i < n; i++)
                w[i] = g[i];
        }
        if (fatal)
            return ORC_LB_FATAL;

        mcsrch(&lb->mc, n, x, f, g, w + lb->ispt + lb->point * n, &lb->stp, lb->ftol, xtol,
               lb->maxfev, &lb->info, &lb->nfev, diag, &fatal);
        if (fatal)
            return ORC_LB_FATAL;

        if (lb->info == -1) {
            *iflag = 1;
            return ORC_LB_OK;
        }

        if (lb->info != 1) {
            *iflag = -1;
            return ORC_LB_IFLAG_EXC; /* "Line search failed" (LBFGS.java:542-549) */
        }

        lb->nfun += lb->nfev;
        lb->npt = lb->point * n;

        for (int i = 0; i < n; i++) {
            w[lb->ispt + lb->npt + i] *= lb->stp;
            w[lb->iypt + lb->npt + i] = g[i] - w[i];
        }

        lb->point++;
        if (lb->point == m)
            lb->point = 0;

        lb->gnorm = sqrt(ddot(n, g, g, &fatal));
        lb->xnorm = sqrt(ddot(n, x, x, &fatal));
        lb->xnorm = fmax(1.0, lb->xnorm);
        if (fatal)
            return ORC_LB_FATAL;

        if (lb->gnorm / lb->xnorm <= eps)
            lb->finish = 1;

        if (lb->finish) {
            *iflag = 0;
            return ORC_LB_OK;
        }

        execute_entire_while_loop = 1;
    }
}

/* The worker's driver loop, LOne.java:865-896, over a user callback. */
int orc_lbfgs_minimize(int n, int m, double *x, double eps, double xtol, orc_fg_cb cb, void *user,
                       int *nfev_out, int *iters_out)
{
    orc_lbfgs lb;
    int iflag = 0, nfev = 0, rc;
    double *g = (double *)calloc((size_t)n, sizeof(double));
    double *diag = (double *)calloc((size_t)n, sizeof(double));
    orc_lbfgs_init(&lb, n, m);

    double f = cb(user, x, g);
    nfev++;
    rc = orc_lbfgs_step(&lb, x, f, g, diag, eps, xtol, &iflag);
    while (rc == ORC_LB_OK && iflag == 1) {
        f = cb(user, x, g);
        nfev++;
        rc = orc_lbfgs_step(&lb, x, f, g, diag, eps, xtol, &iflag);
    }
    if (nfev_out)
        *nfev_out = nfev;
    if (iters_out)
        *iters_out = lb.iter;
    orc_lbfgs_free(&lb);
    free(g);
    free(diag);
    if (rc == ORC_LB_FATAL)
        return -100;
    return iflag;
}
