/*
 * lone_oracle.c — CPU restatement of framework/LOne.java (TEST INFRASTRUCTURE ONLY; see
 * oracle.h). Every quirk listed in SURVEY.md Appendix A is reproduced on purpose:
 *   - rows >= floor(N/T)*T dropped, interleaved blocking i % T        (LOne.java:337-357)
 *   - kappa = 2*lambda/(rho*T), thresholds the intercept too           (:709,536-541)
 *   - augmented term skips the intercept                                (:988,994,1037,1043)
 *   - u-update precedes the x-update and divides by a possibly stale rho_fold (:833-861)
 *   - convergence of iteration k-1 tested after the k-th x-update, which is then discarded
 *                                                                       (:649-665,753-758)
 *   - residual accumulators not reset between a leaf's parents, in-place sqrt (:422-438)
 *   - unorm is always 0 when read (:650 vs :680); parentless-leaf znorm reads u (:504)
 *   - z, zold, u, rho, rho_fold, ranADMM persist across outer iterations (:56-64,27-28,51)
 *   - outer relaxed convergence after it > NODES_maxItr/2, lower median  (:172-180,237)
 * Dense numNodes^2*dim z/u arrays are kept exactly as the reference allocates them (:80-85).
 * The T Java threads are independent between the two barriers (:618-645), so they are executed
 * as a loop over blocks (optionally spread over host_threads pthreads; results are identical).
 */
#include "oracle_internal.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* framework/SeqUnwinderConfig.java:93-101 */
static const double ADMM_ALPHA = 1.9;
static const double ADMM_ABSTOL = 1E-2;
static const double ADMM_RELTOL = 1E-2;
static const double NODES_TOL = 1E-2;
static const double ADMM_PHO_MAX = 1000;

typedef struct {
    int index, is_leaf, layer;
    int n_parents, n_children;
    const int32_t *parents, *children;
} node_t;

typedef struct {
    const orc_problem *p;
    int dim, C, NN, T;
    int64_t zlen, block_size;
    node_t *nodes;          /* by nodeIndex */
    int *leafs, n_leafs;    /* classStructure.leafs, file order */
    int **layers, *layer_n; /* classStructure.layers */
    /* LOne fields */
    double *x, *sm_x, *z, *zold, *u;
    double pho, pho_fold;
    int ran_admm;
    /* ADMMrunner fields */
    double **t_data, **t_weights;
    int32_t **t_cls;
    double **t_x, **t_u;     /* published by the workers */
    double **t_b_x, **t_b_u; /* worker-private copies */
    int *nfev;
    double *h_primal, *h_dual, *h_xnorm, *h_unorm, *h_znorm; /* [maxItr][NN*NN] / [maxItr][NN] */
    int32_t *order;          /* summation order over thread buffers */
    int lbfgs_errors, fatal;
} lone_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* java.lang.String.hashCode + java.util.HashMap bucket order (Java 8). */
void orc_java_thread_order(int T, int32_t *order)
{
    int cap = 16;
    while (T > (cap * 3) / 4) /* resize when ++size > 0.75*cap; splits keep relative order */
        cap *= 2;
    int n = 0;
    for (int b = 0; b < cap; b++) {
        for (int t = 0; t < T; t++) {
            char key[32];
            snprintf(key, sizeof key, "Thread%d", t);
            uint32_t h = 0;
            for (const char *c = key; *c; c++)
                h = 31u * h + (uint32_t)(unsigned char)*c;
            h ^= (h >> 16);
            if ((int)(h & (uint32_t)(cap - 1)) == b)
                order[n++] = t;
        }
    }
}

/* LOne.java:1007-1049 */
static double objective_fn(const lone_t *L, const double *data, const int32_t *cls,
                           const double *wts, int64_t nb, const double *c_x, const double *t_b_u,
                           double *exp_buf)
{
    const int dim = L->dim, C = L->C, NN = L->NN;
    double nll = 0.0;
    for (int64_t i = 0; i < nb; i++) {
        const double *row = data + i * dim;
        for (int off = 0; off < C; off++) {
            int index = off * dim;
            exp_buf[off] = 0.0;
            for (int j = 0; j < dim; j++)
                exp_buf[off] += row[j] * c_x[index + j];
        }
        double max = exp_buf[0]; /* weka Utils.maxIndex: first maximum */
        for (int off = 1; off < C; off++)
            if (exp_buf[off] > max)
                max = exp_buf[off];
        double denom = 0;
        double num = exp_buf[cls[i]] - max;
        for (int off = 0; off < C; off++)
            denom += exp(exp_buf[off] - max);
        nll -= wts[i] * (num - log(denom));
    }
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        int nOffset = n->index * dim;
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                int64_t pOffset = (int64_t)pid * dim;
                int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)pid * dim);
                for (int w = 1; w < dim; w++) {
                    double e = c_x[nOffset + w] - L->sm_x[pOffset + w] - L->z[zOffset + w] +
                               t_b_u[zOffset + w];
                    nll += (L->pho / 2) * e * e;
                }
            }
        } else {
            int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)n->index * dim);
            for (int w = 1; w < dim; w++) {
                double e = c_x[nOffset + w] - L->z[zOffset + w] + t_b_u[zOffset + w];
                nll += (L->pho / 2) * e * e;
            }
        }
    }
    return nll;
}

/* LOne.java:936-1000 */
static void gradient_fn(const lone_t *L, const double *data, const int32_t *cls,
                        const double *wts, int64_t nb, const double *c_x, const double *t_b_u,
                        double *num, double *grad)
{
    const int dim = L->dim, C = L->C, NN = L->NN;
    memset(grad, 0, sizeof(double) * (size_t)C * dim);
    for (int64_t i = 0; i < nb; i++) {
        const double *row = data + i * dim;
        for (int off = 0; off < C; off++) {
            double e = 0.0;
            int index = off * dim;
            for (int j = 0; j < dim; j++)
                e += row[j] * c_x[index + j];
            num[off] = e;
        }
        double max = num[0];
        for (int off = 1; off < C; off++)
            if (num[off] > max)
                max = num[off];
        double denom = 0.0;
        for (int off = 0; off < C; off++) {
            num[off] = exp(num[off] - max);
            denom += num[off];
        }
        for (int off = 0; off < C; off++) /* weka Utils.normalize */
            num[off] /= denom;
        for (int off = 0; off < C; off++) {
            int index = off * dim;
            double firstTerm = wts[i] * num[off];
            for (int q = 0; q < dim; q++)
                grad[index + q] += firstTerm * row[q];
        }
        for (int q = 0; q < dim; q++)
            grad[cls[i] * dim + q] -= wts[i] * row[q];
    }
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        int nOffset = n->index * dim;
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                int64_t pOffset = (int64_t)pid * dim;
                int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)pid * dim);
                for (int w = 1; w < dim; w++)
                    grad[nOffset + w] += L->pho * (c_x[nOffset + w] - L->sm_x[pOffset + w] -
                                                   L->z[zOffset + w] + t_b_u[zOffset + w]);
            }
        } else {
            int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)n->index * dim);
            for (int w = 1; w < dim; w++)
                grad[nOffset + w] +=
                    L->pho * (c_x[nOffset + w] - L->z[zOffset + w] + t_b_u[zOffset + w]);
        }
    }
}

/* One pass of ADMMrun.run's loop body, LOne.java:830-907, for block t. */
static void worker_step(lone_t *L, int t)
{
    const int dim = L->dim, NN = L->NN, C = L->C;
    const int n = C * dim;
    double *t_b_x = L->t_b_x[t], *t_b_u = L->t_b_u[t];
    double *xhat = (double *)calloc((size_t)L->zlen, sizeof(double));
    double *grad = (double *)malloc(sizeof(double) * (size_t)n);
    double *diag = (double *)calloc((size_t)n, sizeof(double));
    double *tmp = (double *)malloc(sizeof(double) * (size_t)C);

    for (int li = 0; li < L->n_leafs; li++) { /* :835-852 */
        const node_t *nd = &L->nodes[L->leafs[li]];
        int nOffset = nd->index * dim;
        if (nd->n_parents > 0) {
            for (int pi = 0; pi < nd->n_parents; pi++) {
                int pid = nd->parents[pi];
                int64_t zOffset = ((int64_t)nd->index * NN * dim) + ((int64_t)pid * dim);
                int64_t pOffset = (int64_t)pid * dim;
                for (int w = 0; w < dim; w++)
                    xhat[zOffset + w] = ADMM_ALPHA * (t_b_x[nOffset + w]) +
                                        (1 - ADMM_ALPHA) * L->zold[zOffset + w] -
                                        ADMM_ALPHA * L->sm_x[pOffset + w];
            }
        } else {
            int64_t zOffset = ((int64_t)nd->index * NN * dim) + ((int64_t)nd->index * dim);
            for (int w = 0; w < dim; w++)
                xhat[zOffset + w] =
                    ADMM_ALPHA * t_b_x[nOffset + w] + (1 - ADMM_ALPHA) * L->zold[zOffset + w];
        }
    }
    for (int64_t i = 0; i < L->zlen; i++) /* :854-856 */
        t_b_u[i] = t_b_u[i] + xhat[i] - L->z[i];
    for (int64_t i = 0; i < L->zlen; i++) /* :859-861 */
        t_b_u[i] = t_b_u[i] / L->pho_fold;

    /* :865-896 */
    int iflag = 0, nfev = 0, rc;
    orc_lbfgs lb;
    orc_lbfgs_init(&lb, n, 5);
    const double eps = 0.1, xtol = 10e-16;
    double obj = objective_fn(L, L->t_data[t], L->t_cls[t], L->t_weights[t], L->block_size, t_b_x,
                              t_b_u, tmp);
    gradient_fn(L, L->t_data[t], L->t_cls[t], L->t_weights[t], L->block_size, t_b_x, t_b_u, tmp,
                grad);
    nfev++;
    rc = orc_lbfgs_step(&lb, t_b_x, obj, grad, diag, eps, xtol, &iflag);
    while (rc == ORC_LB_OK && iflag == 1) {
        obj = objective_fn(L, L->t_data[t], L->t_cls[t], L->t_weights[t], L->block_size, t_b_x,
                           t_b_u, tmp);
        gradient_fn(L, L->t_data[t], L->t_cls[t], L->t_weights[t], L->block_size, t_b_x, t_b_u,
                    tmp, grad);
        nfev++;
        rc = orc_lbfgs_step(&lb, t_b_x, obj, grad, diag, eps, xtol, &iflag);
    }
    if (rc == ORC_LB_IFLAG_EXC)
        __sync_fetch_and_add(&L->lbfgs_errors, 1);
    if (rc == ORC_LB_FATAL)
        L->fatal = 1;
    orc_lbfgs_free(&lb);
    L->nfev[t] = nfev;

    memcpy(L->t_x[t], t_b_x, sizeof(double) * (size_t)n);          /* :898-902 */
    memcpy(L->t_u[t], t_b_u, sizeof(double) * (size_t)L->zlen);    /* :903-907 */
    free(xhat);
    free(grad);
    free(diag);
    free(tmp);
}

typedef struct {
    lone_t *L;
    int tid, stride;
} wk_arg;

static void *worker_thread(void *a)
{
    wk_arg *w = (wk_arg *)a;
    for (int t = w->tid; t < w->L->T; t += w->stride)
        worker_step(w->L, t);
    return NULL;
}

static void run_workers(lone_t *L)
{
    int H = L->p->host_threads;
    if (H > L->T)
        H = L->T;
    if (H <= 1) {
        for (int t = 0; t < L->T; t++)
            worker_step(L, t);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)H);
    wk_arg *args = (wk_arg *)malloc(sizeof(wk_arg) * (size_t)H);
    for (int i = 0; i < H; i++) {
        args[i].L = L;
        args[i].tid = i;
        args[i].stride = H;
        pthread_create(&th[i], NULL, worker_thread, &args[i]);
    }
    for (int i = 0; i < H; i++)
        pthread_join(th[i], NULL);
    free(th);
    free(args);
}

static double l2_norm_smx(const lone_t *L, int nodeIndex) /* LOne.java:243-251 */
{
    double norm = 0;
    int64_t offset = (int64_t)nodeIndex * L->dim;
    for (int w = 0; w < L->dim; w++)
        norm += L->sm_x[offset + w] * L->sm_x[offset + w];
    return sqrt(norm);
}

static void update_ubar(lone_t *L) /* :395-406 */
{
    for (int64_t i = 0; i < L->zlen; i++) {
        L->u[i] = 0;
        for (int k = 0; k < L->T; k++)
            L->u[i] += L->t_u[L->order[k]][i];
        L->u[i] = L->u[i] / L->T;
    }
}

static void update_xbar(lone_t *L) /* :407-417 */
{
    int n = L->C * L->dim;
    for (int i = 0; i < n; i++) {
        L->x[i] = 0;
        for (int k = 0; k < L->T; k++)
            L->x[i] += L->t_x[L->order[k]][i];
        L->x[i] = L->x[i] / L->T;
    }
}

static void update_residuals(lone_t *L, int itr) /* :418-454 */
{
    const int dim = L->dim, NN = L->NN;
    double *hp = L->h_primal + (int64_t)itr * NN * NN, *hd = L->h_dual + (int64_t)itr * NN * NN;
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        double r_t = 0.0, s_t = 0.0; /* declared outside the parent loop: they accumulate */
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)pid * dim);
                for (int k = 0; k < L->T; k++) {
                    const double *tx = L->t_x[L->order[k]];
                    for (int w = 0; w < dim; w++) {
                        double d = tx[n->index * dim + w] - L->z[zOffset + w] -
                                   L->sm_x[(int64_t)pid * dim + w];
                        r_t += d * d;
                    }
                }
                for (int w = 0; w < dim; w++) {
                    double d = L->pho * (L->z[zOffset + w] - L->zold[zOffset + w]);
                    s_t += d * d;
                }
                s_t = sqrt(s_t * L->T);
                hp[n->index * NN + pid] = sqrt(r_t);
                hd[n->index * NN + pid] = s_t;
            }
        } else {
            int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)n->index * dim);
            for (int k = 0; k < L->T; k++) {
                const double *tx = L->t_x[L->order[k]];
                for (int w = 0; w < dim; w++) {
                    double d = tx[n->index * dim + w] - L->z[zOffset + w];
                    r_t += d * d;
                }
            }
            for (int w = 0; w < dim; w++) {
                double d = L->pho * (L->z[zOffset + w] - L->zold[zOffset + w]);
                s_t += d * d;
            }
            s_t = sqrt(s_t * L->T);
            hp[n->index * NN + n->index] = sqrt(r_t);
            hd[n->index * NN + n->index] = s_t;
        }
    }
}

static void update_unorm(lone_t *L, int itr) /* :455-476 */
{
    const int dim = L->dim, NN = L->NN;
    double *hu = L->h_unorm + (int64_t)itr * NN * NN;
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                double unorm = 0.0;
                int64_t uOffset = ((int64_t)n->index * NN * dim) + ((int64_t)pid * dim);
                for (int w = 0; w < dim; w++)
                    unorm += L->u[uOffset + w] * L->u[uOffset + w];
                hu[n->index * NN + pid] = sqrt(unorm);
            }
        } else {
            double unorm = 0;
            int64_t uOffset = ((int64_t)n->index * NN * dim) + ((int64_t)n->index * dim);
            for (int w = 0; w < dim; w++)
                unorm += L->u[uOffset + w] * L->u[uOffset + w];
            hu[n->index * NN + n->index] = sqrt(unorm);
        }
    }
}

static void update_xnorm(lone_t *L, int itr) /* :477-487 */
{
    const int dim = L->dim;
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        int xOffset = n->index * dim;
        double xnorm = 0.0;
        for (int w = 0; w < dim; w++)
            xnorm += L->x[xOffset + w] * L->x[xOffset + w];
        L->h_xnorm[(int64_t)itr * L->NN + n->index] = sqrt(xnorm);
    }
}

static void update_znorm(lone_t *L, int itr) /* :488-510 */
{
    const int dim = L->dim, NN = L->NN;
    double *hz = L->h_znorm + (int64_t)itr * NN * NN;
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                double znorm = 0.0;
                int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)pid * dim);
                for (int w = 0; w < dim; w++)
                    znorm += L->z[zOffset + w] * L->z[zOffset + w];
                hz[n->index * NN + pid] = sqrt(znorm);
            }
        } else {
            double znorm = 0;
            int64_t zOffset = ((int64_t)n->index * NN * dim) + ((int64_t)n->index * dim);
            for (int w = 0; w < dim; w++)
                znorm += L->u[zOffset + w] * L->u[zOffset + w]; /* sic: reads u (:504) */
            hz[n->index * NN + n->index] = sqrt(znorm);
        }
    }
}

static void update_pho_and_fold(lone_t *L, int A_itr) /* :511-535 */
{
    const int NN = L->NN;
    double primal_residuals = 0, dual_residuals = 0;
    const double mu = 10, tao = 2;
    for (int i = 0; i < NN * NN; i++) {
        primal_residuals += L->h_primal[(int64_t)A_itr * NN * NN + i];
        dual_residuals += L->h_dual[(int64_t)A_itr * NN * NN + i];
    }
    if (primal_residuals > mu * dual_residuals) {
        double old_pho = L->pho;
        L->pho = fmin(ADMM_PHO_MAX, L->pho * tao);
        L->pho_fold = L->pho / old_pho;
    } else if (dual_residuals > primal_residuals * mu) {
        L->pho = L->pho / tao;
        L->pho_fold = 1 / tao;
    } else {
        L->pho_fold = 1.0;
    }
}

static void update_z(lone_t *L, double pho) /* :536-541 */
{
    for (int64_t i = 0; i < L->zlen; i++) {
        double zi = L->z[i];
        double sg = (zi > 0) ? 1.0 : ((zi < 0) ? -1.0 : zi); /* Math.signum */
        L->z[i] = zi - sg * fmin(pho, fabs(zi));
    }
}

static int has_admm_converged(lone_t *L, int itr) /* :544-581 */
{
    const int dim = L->dim, NN = L->NN, T = L->T;
    int converged = 1;
    double *primal_tol = (double *)calloc((size_t)NN * NN, sizeof(double));
    double *dual_tol = (double *)calloc((size_t)NN * NN, sizeof(double));
    for (int li = 0; li < L->n_leafs; li++) {
        const node_t *n = &L->nodes[L->leafs[li]];
        double xnorm = L->h_xnorm[(int64_t)itr * NN + n->index];
        if (n->n_parents > 0) {
            for (int pi = 0; pi < n->n_parents; pi++) {
                int pid = n->parents[pi];
                int zOffset = (n->index * NN) + pid;
                double znorm = L->h_znorm[(int64_t)itr * NN * NN + zOffset];
                double unorm = L->h_unorm[(int64_t)itr * NN * NN + zOffset];
                double cnorm = l2_norm_smx(L, pid);
                primal_tol[zOffset] = sqrt((double)dim) * ADMM_ABSTOL +
                                      sqrt((double)T) * ADMM_RELTOL * fmax(xnorm, fmax(znorm, cnorm));
                dual_tol[zOffset] =
                    sqrt((double)dim) * ADMM_ABSTOL + sqrt((double)T) * ADMM_RELTOL * L->pho * unorm;
            }
        } else {
            int zOffset = (n->index * NN) + n->index;
            double znorm = L->h_znorm[(int64_t)itr * NN * NN + zOffset];
            double unorm = L->h_unorm[(int64_t)itr * NN * NN + zOffset];
            primal_tol[zOffset] =
                sqrt((double)dim) * ADMM_ABSTOL + sqrt((double)T) * ADMM_RELTOL * fmax(xnorm, znorm);
            dual_tol[zOffset] =
                sqrt((double)dim) * ADMM_ABSTOL + sqrt((double)T) * ADMM_RELTOL * L->pho * unorm;
        }
    }
    for (int i = 0; i < NN * NN; i++) {
        if (L->h_primal[(int64_t)itr * NN * NN + i] > primal_tol[i])
            converged = 0;
        if (L->h_dual[(int64_t)itr * NN * NN + i] > dual_tol[i])
            converged = 0;
        if (!converged)
            break;
    }
    free(primal_tol);
    free(dual_tol);
    return converged;
}

/* new ADMMrunner() + ADMMrunner.execute(), LOne.java:330-392,595-760. Returns iterations. */
static int admm_run(lone_t *L, int outer, orc_trace *tr)
{
    const int dim = L->dim, NN = L->NN, T = L->T, maxItr = L->p->admm_max_itr;
    const int n = L->C * dim;

    for (int t = 0; t < T; t++) { /* :360-391 and ADMMrun ctor :794-806 */
        memcpy(L->t_x[t], L->x, sizeof(double) * (size_t)n);
        memcpy(L->t_u[t], L->u, sizeof(double) * (size_t)L->zlen);
        memcpy(L->t_b_x[t], L->t_x[t], sizeof(double) * (size_t)n);
        memcpy(L->t_b_u[t], L->t_u[t], sizeof(double) * (size_t)L->zlen);
    }
    memset(L->h_primal, 0, sizeof(double) * (size_t)maxItr * NN * NN); /* initHistories :322 */
    memset(L->h_dual, 0, sizeof(double) * (size_t)maxItr * NN * NN);
    memset(L->h_unorm, 0, sizeof(double) * (size_t)maxItr * NN * NN);
    memset(L->h_znorm, 0, sizeof(double) * (size_t)maxItr * NN * NN);
    memset(L->h_xnorm, 0, sizeof(double) * (size_t)maxItr * NN);

    double *xhat = (double *)malloc(sizeof(double) * (size_t)L->zlen);
    int itr = 0;
    while (itr < maxItr) { /* :608 */
        if (itr > 0 && !L->ran_admm && L->pho < ADMM_PHO_MAX) /* :614-615 */
            update_pho_and_fold(L, itr - 1);

        double pho_used = L->pho, fold_used = L->pho_fold;
        run_workers(L); /* barriers :618-645 */
        for (int t = 0; t < T; t++)
            tr->total_evals += L->nfev[t];
        if (L->fatal)
            break;

        if (itr > 0 && has_admm_converged(L, itr - 1)) { /* :649-665 */
            L->ran_admm = 1;
            break;
        }

        memcpy(L->zold, L->z, sizeof(double) * (size_t)L->zlen); /* :670-672 */
        update_xbar(L);                                          /* :675 */
        update_ubar(L);                                          /* :677 */
        if (itr > 0)
            update_unorm(L, itr - 1);                            /* :679-680 */
        update_xnorm(L, itr);                                    /* :682 */

        memset(xhat, 0, sizeof(double) * (size_t)L->zlen);      /* :685-702 */
        for (int li = 0; li < L->n_leafs; li++) {
            const node_t *nd = &L->nodes[L->leafs[li]];
            int nOffset = nd->index * dim;
            if (nd->n_parents > 0) {
                for (int pi = 0; pi < nd->n_parents; pi++) {
                    int pid = nd->parents[pi];
                    int64_t zOffset = ((int64_t)nd->index * NN * dim) + ((int64_t)pid * dim);
                    int64_t pOffset = (int64_t)pid * dim;
                    for (int w = 0; w < dim; w++)
                        xhat[zOffset + w] = ADMM_ALPHA * (L->x[nOffset + w]) +
                                            (1 - ADMM_ALPHA) * L->zold[zOffset + w] -
                                            ADMM_ALPHA * L->sm_x[pOffset + w];
                }
            } else {
                int64_t zOffset = ((int64_t)nd->index * NN * dim) + ((int64_t)nd->index * dim);
                for (int w = 0; w < dim; w++)
                    xhat[zOffset + w] =
                        ADMM_ALPHA * L->x[nOffset + w] + (1 - ADMM_ALPHA) * L->zold[zOffset + w];
            }
        }
        for (int64_t i = 0; i < L->zlen; i++) /* :704-706 */
            L->z[i] = xhat[i] + L->u[i];

        update_z(L, (2 * L->p->ridge) / (L->pho * T)); /* :709 */
        update_znorm(L, itr);                          /* :711 */
        update_residuals(L, itr);                      /* :714 */

        if (tr->log_rows < tr->log_cap) {
            int r = tr->log_rows++;
            double ps = 0, ds = 0;
            for (int i = 0; i < NN * NN; i++) {
                ps += L->h_primal[(int64_t)itr * NN * NN + i];
                ds += L->h_dual[(int64_t)itr * NN * NN + i];
            }
            tr->log_outer[r] = outer;
            tr->log_itr[r] = itr;
            tr->log_pho[r] = pho_used;
            tr->log_fold[r] = fold_used;
            tr->log_primal[r] = ps;
            tr->log_dual[r] = ds;
            for (int t = 0; t < T; t++)
                tr->log_nfev[(int64_t)r * T + t] = L->nfev[t];
        }
        itr++; /* :716 */
    }
    free(xhat);

    for (int li = 0; li < L->n_leafs; li++) { /* :753-758 */
        int nOffset = L->nodes[L->leafs[li]].index * dim;
        for (int w = 0; w < dim; w++)
            L->sm_x[nOffset + w] = L->x[nOffset + w];
    }
    return itr;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

static void update_node(lone_t *L, const node_t *n) /* :223-241 */
{
    const int dim = L->dim;
    int64_t nOffset = (int64_t)n->index * dim;
    int cnt = n->n_parents + n->n_children;
    double *xs = (double *)malloc(sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
    for (int w = 1; w < dim; w++) {
        int k = 0;
        for (int i = 0; i < n->n_parents; i++)
            xs[k++] = L->sm_x[(int64_t)n->parents[i] * dim + w];
        for (int i = 0; i < n->n_children; i++)
            xs[k++] = L->sm_x[(int64_t)n->children[i] * dim + w];
        qsort(xs, (size_t)k, sizeof(double), cmp_double);
        L->sm_x[nOffset + w] = (k % 2 == 0) ? xs[(k / 2) - 1] : xs[k / 2];
    }
    free(xs);
}

static void update_internal_nodes(lone_t *L) /* :208-221 */
{
    const int numLayers = L->p->st.num_layers;
    for (int l = 1; l < numLayers; l += 2)
        for (int i = 0; i < L->layer_n[l]; i++)
            update_node(L, &L->nodes[L->layers[l][i]]);
    for (int l = 2; l < numLayers; l += 2)
        for (int i = 0; i < L->layer_n[l]; i++)
            update_node(L, &L->nodes[L->layers[l][i]]);
}

static int build_structure(lone_t *L)
{
    const orc_structure *st = &L->p->st;
    const int NN = st->num_nodes;
    L->nodes = (node_t *)calloc((size_t)NN, sizeof(node_t));
    L->leafs = (int *)calloc((size_t)NN, sizeof(int));
    L->layers = (int **)calloc((size_t)st->num_layers, sizeof(int *));
    L->layer_n = (int *)calloc((size_t)st->num_layers, sizeof(int));
    for (int l = 0; l < st->num_layers; l++)
        L->layers[l] = (int *)calloc((size_t)NN, sizeof(int));
    for (int e = 0; e < NN; e++) { /* file order, Classifier.java:658-683 */
        int idx = st->node_index[e];
        if (idx < 0 || idx >= NN || st->layer[e] < 0 || st->layer[e] >= st->num_layers)
            return -1;
        node_t *nd = &L->nodes[idx];
        nd->index = idx;
        nd->is_leaf = st->is_leaf[e];
        nd->layer = st->layer[e];
        nd->n_parents = nd->is_leaf ? st->parent_ptr[e + 1] - st->parent_ptr[e] : 0;
        nd->parents = st->parent_idx + st->parent_ptr[e];
        nd->n_children = nd->is_leaf ? 0 : st->child_ptr[e + 1] - st->child_ptr[e];
        nd->children = st->child_idx + st->child_ptr[e];
        if (nd->is_leaf) {
            if (idx >= L->C)
                return -1;
            L->leafs[L->n_leafs++] = idx;
        }
        L->layers[nd->layer][L->layer_n[nd->layer]++] = idx;
    }
    return 0;
}

static void lone_free(lone_t *L)
{
    for (int t = 0; t < L->T; t++) {
        if (L->t_data) free(L->t_data[t]);
        if (L->t_weights) free(L->t_weights[t]);
        if (L->t_cls) free(L->t_cls[t]);
        if (L->t_x) free(L->t_x[t]);
        if (L->t_u) free(L->t_u[t]);
        if (L->t_b_x) free(L->t_b_x[t]);
        if (L->t_b_u) free(L->t_b_u[t]);
    }
    free(L->t_data); free(L->t_weights); free(L->t_cls);
    free(L->t_x); free(L->t_u); free(L->t_b_x); free(L->t_b_u);
    free(L->nfev); free(L->order);
    free(L->h_primal); free(L->h_dual); free(L->h_xnorm); free(L->h_unorm); free(L->h_znorm);
    free(L->z); free(L->zold); free(L->u);
    if (L->layers)
        for (int l = 0; l < L->p->st.num_layers; l++)
            free(L->layers[l]);
    free(L->layers); free(L->layer_n); free(L->leafs); free(L->nodes);
}

static int lone_setup(lone_t *L, const orc_problem *p, double *x, double *smx)
{
    memset(L, 0, sizeof(*L));
    L->p = p;
    L->dim = p->dim;
    L->C = p->num_classes;
    L->NN = p->st.num_nodes;
    L->T = p->num_threads;
    L->zlen = (int64_t)L->NN * L->NN * L->dim;
    L->x = x;
    L->sm_x = smx;
    L->pho = p->pho;      /* setPho, LOne.java:103 */
    L->pho_fold = 1.0;    /* :28 */
    L->ran_admm = 0;      /* :51 */
    if (L->T < 1 || L->dim < 1 || L->C < 1)
        return -1;
    if (build_structure(L) != 0)
        return -1;
    L->z = (double *)calloc((size_t)L->zlen, sizeof(double)); /* initUandZ :80-85 */
    L->zold = (double *)calloc((size_t)L->zlen, sizeof(double));
    L->u = (double *)calloc((size_t)L->zlen, sizeof(double));
    const int T = L->T, dim = L->dim;
    L->block_size = p->n_rows / T; /* :337 */
    L->t_data = (double **)calloc((size_t)T, sizeof(double *));
    L->t_weights = (double **)calloc((size_t)T, sizeof(double *));
    L->t_cls = (int32_t **)calloc((size_t)T, sizeof(int32_t *));
    L->t_x = (double **)calloc((size_t)T, sizeof(double *));
    L->t_u = (double **)calloc((size_t)T, sizeof(double *));
    L->t_b_x = (double **)calloc((size_t)T, sizeof(double *));
    L->t_b_u = (double **)calloc((size_t)T, sizeof(double *));
    L->nfev = (int *)calloc((size_t)T, sizeof(int));
    L->order = (int32_t *)calloc((size_t)T, sizeof(int32_t));
    for (int t = 0; t < T; t++) {
        size_t bs = (size_t)(L->block_size > 0 ? L->block_size : 1);
        L->t_data[t] = (double *)malloc(sizeof(double) * bs * dim);
        L->t_weights[t] = (double *)malloc(sizeof(double) * bs);
        L->t_cls[t] = (int32_t *)malloc(sizeof(int32_t) * bs);
        L->t_x[t] = (double *)malloc(sizeof(double) * (size_t)L->C * dim);
        L->t_b_x[t] = (double *)malloc(sizeof(double) * (size_t)L->C * dim);
        L->t_u[t] = (double *)malloc(sizeof(double) * (size_t)L->zlen);
        L->t_b_u[t] = (double *)malloc(sizeof(double) * (size_t)L->zlen);
    }
    /* :338-357 (the per-ADMMrunner deep copy always yields the same blocks; done once) */
    for (int64_t i = 0; i < L->block_size * T; i++) {
        int t = (int)(i % T);
        int64_t s = i / T;
        L->t_weights[t][s] = p->weights[i];
        L->t_cls[t][s] = p->cls[i];
        memcpy(L->t_data[t] + s * dim, p->data + i * dim, sizeof(double) * (size_t)dim);
    }
    if (p->java_sum_order)
        orc_java_thread_order(T, L->order);
    else
        for (int t = 0; t < T; t++)
            L->order[t] = t;
    const size_t hist = (size_t)p->admm_max_itr * L->NN * L->NN;
    L->h_primal = (double *)calloc(hist, sizeof(double));
    L->h_dual = (double *)calloc(hist, sizeof(double));
    L->h_unorm = (double *)calloc(hist, sizeof(double));
    L->h_znorm = (double *)calloc(hist, sizeof(double));
    L->h_xnorm = (double *)calloc((size_t)p->admm_max_itr * L->NN, sizeof(double));
    return 0;
}

/* LOne.execute(), LOne.java:116-204 */
int orc_lone_execute(const orc_problem *p, double *x_inout, double *smx_inout, orc_trace *tr)
{
    lone_t L;
    if (lone_setup(&L, p, x_inout, smx_inout) != 0) {
        lone_free(&L);
        return -1;
    }
    const int dim = L.dim, NN = L.NN;
    const int NODES_maxItr = p->nodes_max_itr;
    double *sm_x_old = (double *)malloc(sizeof(double) * (size_t)NN * dim);
    double *deltas = (double *)malloc(sizeof(double) * (size_t)NN);
    double *targets = (double *)malloc(sizeof(double) * (size_t)NN);
    tr->outer_iters = 0;
    tr->converged = 0;
    tr->log_rows = 0;
    tr->total_evals = 0;
    tr->seconds_admm = 0;

    for (int it = 0; it < NODES_maxItr; it++) {
        memcpy(sm_x_old, L.sm_x, sizeof(double) * (size_t)NN * dim); /* :122-126 */

        double t0 = now_s();
        int iters = admm_run(&L, it, tr); /* :133-134 */
        tr->seconds_admm += now_s() - t0;
        if (tr->admm_iters)
            tr->admm_iters[it] = iters;
        tr->outer_iters = it + 1;
        if (L.fatal)
            break;

        update_internal_nodes(&L); /* :142 */

        for (int nidx = 0; nidx < NN; nidx++) { /* :152-165 */
            double diff = 0.0;
            for (int w = 0; w < dim; w++) {
                double d = sm_x_old[(int64_t)nidx * dim + w] - L.sm_x[(int64_t)nidx * dim + w];
                diff += d * d;
            }
            deltas[nidx] = sqrt(diff);
            targets[nidx] = NODES_TOL * l2_norm_smx(&L, nidx);
        }
        int numPass = 0, numPass_relaxed = 0; /* :167-174 */
        for (int q = 0; q < NN; q++) {
            if (deltas[q] < targets[q])
                numPass++;
            if (deltas[q] < targets[q] * 2 / NODES_TOL)
                numPass_relaxed++;
        }
        int converged = 0; /* :176-180 */
        if (numPass == NN)
            converged = 1;
        else if (numPass_relaxed == NN && it > NODES_maxItr / 2)
            converged = 1;
        if (converged) {
            tr->converged = 1;
            break;
        }
    }
    tr->final_pho = L.pho;
    tr->final_fold = L.pho_fold;
    tr->ran_admm = L.ran_admm;
    tr->lbfgs_errors = L.lbfgs_errors;
    int rc = L.fatal ? -2 : 0;
    free(sm_x_old);
    free(deltas);
    free(targets);
    lone_free(&L);
    return rc;
}

/* Single-evaluation entry points (dense reference layout for z/u). */
double orc_objective(const orc_problem *p, const double *blk_data, const int32_t *blk_cls,
                     const double *blk_w, int64_t nb, const double *cx, const double *smx,
                     const double *z, const double *tu, double pho)
{
    lone_t L;
    memset(&L, 0, sizeof L);
    L.p = p;
    L.dim = p->dim;
    L.C = p->num_classes;
    L.NN = p->st.num_nodes;
    L.T = 0;
    if (build_structure(&L) != 0)
        return NAN;
    L.sm_x = (double *)smx;
    L.z = (double *)z;
    L.pho = pho;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)L.C);
    double f = objective_fn(&L, blk_data, blk_cls, blk_w, nb, cx, tu, tmp);
    free(tmp);
    L.z = NULL;
    lone_free(&L);
    return f;
}

void orc_gradient(const orc_problem *p, const double *blk_data, const int32_t *blk_cls,
                  const double *blk_w, int64_t nb, const double *cx, const double *smx,
                  const double *z, const double *tu, double pho, double *grad)
{
    lone_t L;
    memset(&L, 0, sizeof L);
    L.p = p;
    L.dim = p->dim;
    L.C = p->num_classes;
    L.NN = p->st.num_nodes;
    L.T = 0;
    if (build_structure(&L) != 0)
        return;
    L.sm_x = (double *)smx;
    L.z = (double *)z;
    L.pho = pho;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)L.C);
    gradient_fn(&L, blk_data, blk_cls, blk_w, nb, cx, tu, tmp, grad);
    free(tmp);
    L.z = NULL;
    lone_free(&L);
}
