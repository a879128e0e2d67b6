/*
 * classifier_oracle.c — CPU restatement of the caller side of the trainer boundary,
 * framework/Classifier.java:337-420 (statistics, standardisation, null-model init),
 * :452-477 (de-standardisation) and :267-287 (evaluateProbability).
 * TEST INFRASTRUCTURE ONLY; see oracle.h.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Classifier.java:337-381. data has column 0 == 1 already; returns -1 for the
 * "Sum of weights of instances less than 1" exception (:359-362). */
int orc_standardize(double *data, int64_t n, int32_t dim, const double *weights, double *mean,
                    double *sd)
{
    double totWeights = 0;
    for (int j = 0; j < dim; j++) {
        mean[j] = 0;
        sd[j] = 0;
    }
    for (int64_t i = 0; i < n; i++) {
        totWeights += weights[i];
        for (int j = 1; j < dim; j++) {
            double x = data[i * dim + j];
            mean[j] += weights[i] * x;
            sd[j] += weights[i] * x * x;
        }
    }
    if ((totWeights <= 1) && (n > 1))
        return -1;
    mean[0] = 0;
    sd[0] = 1;
    for (int j = 1; j < dim; j++) {
        mean[j] = mean[j] / totWeights;
        if (totWeights > 1)
            sd[j] = sqrt(fabs(sd[j] - totWeights * mean[j] * mean[j]) / (totWeights - 1));
        else
            sd[j] = 0;
    }
    for (int64_t i = 0; i < n; i++)
        for (int j = 0; j < dim; j++)
            if (sd[j] != 0)
                data[i * dim + j] = (data[i * dim + j] - mean[j]) / sd[j];
    return 0;
}

/* Classifier.java:383-420: class counts are unweighted; label counts are subtree sums. */
void orc_init_params(const int32_t *cls, int64_t n, int32_t dim, int32_t C,
                     const orc_structure *st, double *x, double *smx)
{
    const int NN = st->num_nodes;
    double *sY = (double *)calloc((size_t)C, sizeof(double));
    double *sm_sY = (double *)calloc((size_t)NN, sizeof(double));
    for (int64_t i = 0; i < n; i++)
        sY[cls[i]]++;
    for (int e = 0; e < NN; e++) /* leaves first (:388-390) */
        if (st->is_leaf[e])
            sm_sY[st->node_index[e]] = sY[st->node_index[e]];
    for (int l = 1; l < st->num_layers; l++) /* :391-397, file order inside a layer */
        for (int e = 0; e < NN; e++)
            if (st->layer[e] == l && !st->is_leaf[e])
                for (int c = st->child_ptr[e]; c < st->child_ptr[e + 1]; c++)
                    sm_sY[st->node_index[e]] = sm_sY[st->node_index[e]] + sm_sY[st->child_idx[c]];
    for (int p = 0; p < C; p++) { /* :400-410 */
        int64_t offset = (int64_t)p * dim;
        x[offset] = log(sY[p] + 1.0) - log(sY[C - 1] + 1.0);
        for (int q = 1; q < dim; q++)
            x[offset + q] = 0.0;
    }
    for (int p = 0; p < NN; p++) { /* :412-420 */
        int64_t offset = (int64_t)p * dim;
        smx[offset] = log(sm_sY[p] + 1.0) - log(sm_sY[C - 1] + 1.0);
        for (int q = 1; q < dim; q++)
            smx[offset + q] = 0.0;
    }
    free(sY);
    free(sm_sY);
}

/* Classifier.java:452-477. par is [dim][ncols] like m_Par / sm_Par. */
void orc_destandardize(const double *v, int32_t ncols, int32_t dim, const double *mean,
                       const double *sd, double *par)
{
    for (int i = 0; i < ncols; i++) {
        par[0 * ncols + i] = v[(int64_t)i * dim];
        for (int j = 1; j < dim; j++) {
            par[(int64_t)j * ncols + i] = v[(int64_t)i * dim + j];
            if (sd[j] != 0) {
                par[(int64_t)j * ncols + i] /= sd[j];
                par[0 * ncols + i] -= par[(int64_t)j * ncols + i] * mean[j];
            }
        }
    }
}

/* Classifier.java:267-287 */
void orc_predict(const double *par, int32_t dim, int32_t C, const double *data_raw, int64_t n,
                 double *prob)
{
    double *v = (double *)malloc(sizeof(double) * (size_t)C);
    for (int64_t r = 0; r < n; r++) {
        const double *data = data_raw + r * dim;
        for (int j = 0; j < C; j++) {
            v[j] = 0;
            for (int k = 0; k < dim; k++)
                v[j] += par[(int64_t)k * C + j] * data[k];
        }
        for (int m = 0; m < C; m++) {
            double sum = 0;
            for (int q = 0; q < C; q++)
                sum += exp(v[q] - v[m]);
            prob[r * C + m] = 1 / (sum);
        }
    }
    free(v);
}
