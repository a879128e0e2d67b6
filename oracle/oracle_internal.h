/* oracle_internal.h — solver state shared between the oracle's C files (TEST INFRASTRUCTURE
 * ONLY; see oracle.h). Fields mirror the instance fields of framework/Mcsrch.java:27-37 and
 * framework/LBFGS.java:148-159. */
#ifndef SQW_ORACLE_INTERNAL_H
#define SQW_ORACLE_INTERNAL_H

#include "oracle.h"

enum { ORC_LB_OK = 0, ORC_LB_IFLAG_EXC = 1, ORC_LB_FATAL = 2 };

typedef struct {
    int infoc;
    double dg, dgm, dginit, dgtest, dgx, dgxm, dgy, dgym, finit, ftest1, fm, fx, fxm, fy, fym;
    double stx, sty, stmin, stmax, width, width1;
    int brackt, stage1;
} orc_mcsrch;

typedef struct {
    int n, m, maxfev;
    double gnorm, stp1, ftol, stp, ys, yy, sq, yr, beta, xnorm;
    int iter, nfun, point, ispt, iypt, info, bound, npt, cp, nfev, inmc, iycn, iscn;
    int finish;
    double *w;
    orc_mcsrch mc;
} orc_lbfgs;

void orc_lbfgs_init(orc_lbfgs *lb, int n, int m);
void orc_lbfgs_free(orc_lbfgs *lb);
/* one reverse-communication call of LBFGS.lbfgs; returns ORC_LB_* */
int orc_lbfgs_step(orc_lbfgs *lb, double *x, double f, const double *g, double *diag, double eps,
                   double xtol, int *iflag);

#endif
