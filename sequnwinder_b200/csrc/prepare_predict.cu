// prepare_predict.cu — the steps either side of the trainer, device resident.
//
// f-1  counts -> training matrix: replaces the text round trip loaddata/MakeArff.java:255-288
//      (tmpCounts.mat -> Weka CSVLoader -> ARFF) and framework/Classifier.java:300-381:
//      RemoveUseless (drop columns that are constant over the training rows), instance-weighted
//      mean / SD with the (W-1) denominator and abs() guard, in-place standardisation, column 0 = 1.
// f-2  batched prediction: framework/Classifier.java:244-287 (distributionForInstance /
//      evaluateProbability) for many rows at once, straight from the dense count matrix.
//
// All sums are two-level with fixed chunking, so results are reproducible run to run; they differ
// from the reference's sequential loops only in rounding (checked against the oracle in tests/).
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace sqw {

constexpr int kRowChunks = 64;  // fixed number of row chunks for the column statistics

// ---- f-1, step 1: per column, per row chunk: min, max, sum w x, sum w x^2 -------------------
__global__ void __launch_bounds__(256)
column_stats_kernel(const int32_t *counts, const double *w, int64_t n, int64_t numK, int32_t *cmin,
                    int32_t *cmax, double *swx, double *swxx)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    if (col >= numK)
        return;
    const int64_t r0 = n * chunk / kRowChunks, r1 = n * (chunk + 1) / kRowChunks;
    int32_t mn = INT32_MAX, mx = INT32_MIN;
    double a = 0.0, b = 0.0;
    for (int64_t r = r0; r < r1; ++r) {  // consecutive threads read consecutive columns: coalesced
        const int32_t v = counts[r * numK + col];
        const double wi = w[r], x = (double)v;
        mn = min(mn, v);
        mx = max(mx, v);
        a += wi * x;
        b += wi * x * x;
    }
    cmin[(int64_t)chunk * numK + col] = mn;
    cmax[(int64_t)chunk * numK + col] = mx;
    swx[(int64_t)chunk * numK + col] = a;
    swxx[(int64_t)chunk * numK + col] = b;
}

// ---- f-1, step 2: combine the row chunks in fixed order ----------------------------------------
__global__ void __launch_bounds__(256)
column_combine_kernel(const int32_t *cmin, const int32_t *cmax, const double *swx, const double *swxx,
                      int64_t numK, int32_t *omin, int32_t *omax, double *oswx, double *oswxx)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= numK)
        return;
    int32_t mn = INT32_MAX, mx = INT32_MIN;
    double a = 0.0, b = 0.0;
    for (int c = 0; c < kRowChunks; ++c) {
        mn = min(mn, cmin[(int64_t)c * numK + col]);
        mx = max(mx, cmax[(int64_t)c * numK + col]);
        a += swx[(int64_t)c * numK + col];
        b += swxx[(int64_t)c * numK + col];
    }
    omin[col] = mn;
    omax[col] = mx;
    oswx[col] = a;
    oswxx[col] = b;
}

// ---- f-1, step 2b: keep flag, mean, sd per column from the combined statistics ----------------
__global__ void __launch_bounds__(256)
column_finish_kernel(const int32_t *cmin, const int32_t *cmax, const double *swx, const double *swxx,
                     int64_t numK, double tot_w, int32_t *keep_flag, double *mean, double *sd)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= numK)
        return;
    keep_flag[col] = (cmin[col] != cmax[col]) ? 1 : 0;  // weka RemoveUseless on a numeric attribute
    const double m = swx[col] / tot_w;                    // Classifier.java:366
    mean[col] = m;
    sd[col] = (tot_w > 1) ? sqrt(fabs(swxx[col] - tot_w * m * m) / (tot_w - 1)) : 0.0;  // :367-371
}

// ---- f-1, step 3: standardised matrix [n][F+1], column 0 = 1 (Classifier.java:344,374-381) ----
__global__ void __launch_bounds__(256)
standardize_kernel(const int32_t *counts, int64_t n, int64_t numK, const int32_t *keep_idx, int32_t F,
                   const double *mean, const double *sd, double *X)
{
    const int64_t dim = (int64_t)F + 1;
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        const int32_t *row = counts + r * numK;
        double *out = X + r * dim;
        for (int64_t j = threadIdx.x; j < dim; j += blockDim.x) {
            if (j == 0) {
                out[0] = 1.0;
            } else {
                const int32_t c = keep_idx[j - 1];
                const double s = sd[c];
                const double x = (double)row[c];
                out[j] = (s != 0.0) ? (x - mean[c]) / s : x;
            }
        }
    }
}

// ---- f-2: evaluateProbability for a batch, one warp per row -------------------------------------
template <int CMAX>
__global__ void __launch_bounds__(256)
predict_kernel(const double *par, int32_t dim, int32_t C, const int32_t *counts, int64_t n,
               int64_t numK, const int32_t *keep_idx, double *prob, int32_t *argmax)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int32_t *row = counts + r * numK;
        double v[CMAX];
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            v[j] = 0.0;
        for (int k = lane; k < dim; k += 32) {  // v_j = sum_k par[k][j] * data[k], data[0] = 1
            const double x = (k == 0) ? 1.0 : (double)row[keep_idx[k - 1]];
            const double *pk = par + (size_t)k * C;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < C)
                    v[j] += pk[j] * x;
        }
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            for (int o = 16; o > 0; o >>= 1)
                v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        // prob_m = 1 / sum_n exp(v_n - v_m) (Classifier.java:278-284); weka takes the first maximum
        double best = -1.0;
        int besti = 0;
        for (int m = 0; m < C; ++m) {
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < CMAX; ++q)
                if (q < C)
                    sum += exp(v[q] - v[m]);
            const double p = 1.0 / sum;
            if (lane == 0 && prob)
                prob[r * C + m] = p;
            if (p > best) {
                best = p;
                besti = m;
            }
        }
        if (lane == 0 && argmax)
            argmax[r] = besti;
    }
}

}  // namespace sqw

extern "C" {

// Per column over the caller's rows: min, max, sum w x, sum w x^2 (the row chunks combined in
// fixed order). With the rows sharded over ranks every rank calls this on its own rows and the
// caller combines the four arrays across ranks (min / max / sum) before sqw_prepare_finish_dev.
int sqw_prepare_stats_dev(const int32_t *d_counts, int64_t n, int64_t numK, const double *d_w,
                          int32_t *d_min, int32_t *d_max, double *d_swx, double *d_swxx, void *stream_)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!d_counts || !d_w || !d_min || !d_max || !d_swx || !d_swxx || n < 1 || numK < 1)
        return set_error(SQW_ERR_INVALID_ARG, "invalid argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    // one scratch allocation (stream ordered), released on every path
    const size_t cells = (size_t)kRowChunks * numK;
    unsigned char *scratch = nullptr;
    SQW_CUDA_CHECK(cudaMallocAsync((void **)&scratch, cells * (2 * sizeof(int32_t) + 2 * sizeof(double)), st));
    double *swx = reinterpret_cast<double *>(scratch), *swxx = swx + cells;
    int32_t *cmin = reinterpret_cast<int32_t *>(swxx + cells), *cmax = cmin + cells;
    const dim3 grid((unsigned)((numK + 255) / 256), kRowChunks);
    column_stats_kernel<<<grid, 256, 0, st>>>(d_counts, d_w, n, numK, cmin, cmax, swx, swxx);
    column_combine_kernel<<<(unsigned)((numK + 255) / 256), 256, 0, st>>>(cmin, cmax, swx, swxx, numK,
                                                                          d_min, d_max, d_swx, d_swxx);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess)
        return set_error(SQW_ERR_CUDA, "prepare stats failed: %s", cudaGetErrorString(e));
    return SQW_OK;
}

// keep flag (weka RemoveUseless: min != max), mean and SD per column from the combined statistics;
// the kept column indices are written in ascending order. Synchronises the stream.
int sqw_prepare_finish_dev(const int32_t *d_min, const int32_t *d_max, const double *d_swx,
                           const double *d_swxx, int64_t numK, double total_weight, int64_t n_total,
                           int32_t *d_keep_idx, int32_t *num_kept_out, double *d_mean, double *d_sd,
                           void *stream_)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!d_min || !d_max || !d_swx || !d_swxx || !d_keep_idx || !num_kept_out || !d_mean || !d_sd || numK < 1)
        return set_error(SQW_ERR_INVALID_ARG, "invalid argument");
    if ((total_weight <= 1) && (n_total > 1))  // Classifier.java:359-362
        return set_error(SQW_ERR_INVALID_ARG, "Sum of weights of instances less than 1, please reweight!");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    int32_t *flag = nullptr;
    SQW_CUDA_CHECK(cudaMallocAsync((void **)&flag, (size_t)numK * sizeof(int32_t), st));
    column_finish_kernel<<<(unsigned)((numK + 255) / 256), 256, 0, st>>>(d_min, d_max, d_swx, d_swxx, numK,
                                                                         total_weight, flag, d_mean, d_sd);
    std::vector<int32_t> hflag((size_t)numK);
    cudaError_t e = cudaMemcpyAsync(hflag.data(), flag, (size_t)numK * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(flag, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        return set_error(SQW_ERR_CUDA, "prepare plan failed: %s", cudaGetErrorString(e));
    std::vector<int32_t> keep;
    for (int64_t c = 0; c < numK; ++c)
        if (hflag[(size_t)c])
            keep.push_back((int32_t)c);
    *num_kept_out = (int32_t)keep.size();
    if (!keep.empty())
        SQW_CUDA_CHECK(cudaMemcpyAsync(d_keep_idx, keep.data(), keep.size() * sizeof(int32_t),
                                       cudaMemcpyHostToDevice, st));
    SQW_CUDA_CHECK(cudaStreamSynchronize(st));
    return SQW_OK;
}

int sqw_prepare_plan_dev(const int32_t *d_counts, int64_t n, int64_t numK, const double *d_w,
                         double total_weight, int32_t *d_keep_idx, int32_t *num_kept_out,
                         double *d_mean, double *d_sd, void *stream_)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!d_counts || !d_w || !d_keep_idx || !num_kept_out || !d_mean || !d_sd || n < 1 || numK < 1)
        return set_error(SQW_ERR_INVALID_ARG, "invalid argument");
    if ((total_weight <= 1) && (n > 1))  // Classifier.java:359-362
        return set_error(SQW_ERR_INVALID_ARG, "Sum of weights of instances less than 1, please reweight!");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    unsigned char *scratch = nullptr;
    SQW_CUDA_CHECK(cudaMallocAsync((void **)&scratch, (size_t)numK * (2 * sizeof(int32_t) + 2 * sizeof(double)), st));
    double *swx = reinterpret_cast<double *>(scratch), *swxx = swx + numK;
    int32_t *cmin = reinterpret_cast<int32_t *>(swxx + numK), *cmax = cmin + numK;
    rc = sqw_prepare_stats_dev(d_counts, n, numK, d_w, cmin, cmax, swx, swxx, stream_);
    if (rc == SQW_OK)
        rc = sqw_prepare_finish_dev(cmin, cmax, swx, swxx, numK, total_weight, n, d_keep_idx, num_kept_out,
                                    d_mean, d_sd, stream_);
    cudaFreeAsync(scratch, st);
    return rc;
}

int sqw_prepare_apply_dev(const int32_t *d_counts, int64_t n, int64_t numK, const int32_t *d_keep_idx,
                          int32_t num_kept, const double *d_mean, const double *d_sd, double *d_X,
                          void *stream_)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!d_counts || !d_keep_idx || !d_mean || !d_sd || !d_X || n < 1 || num_kept < 0)
        return set_error(SQW_ERR_INVALID_ARG, "invalid argument");
    const int grid = (int)std::min<int64_t>(n, (int64_t)kNumSMs * 16);
    standardize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        d_counts, n, numK, d_keep_idx, num_kept, d_mean, d_sd, d_X);
    SQW_CUDA_CHECK(cudaGetLastError());
    return SQW_OK;
}

int sqw_predict_dev(const double *d_par, int32_t dim, int32_t num_classes, const int32_t *d_counts,
                    int64_t n, int64_t numK, const int32_t *d_keep_idx, double *d_prob,
                    int32_t *d_argmax, void *stream_)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if (!d_par || !d_counts || !d_keep_idx || dim < 1 || num_classes < 1 || n < 0)
        return set_error(SQW_ERR_INVALID_ARG, "invalid argument");
    if (num_classes > 32)
        return set_error(SQW_ERR_UNSUPPORTED, "more than 32 classes");
    if (n == 0)
        return SQW_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)kNumSMs * 8);
    if (num_classes <= 8)
        predict_kernel<8><<<grid, 256, 0, st>>>(d_par, dim, num_classes, d_counts, n, numK, d_keep_idx,
                                                d_prob, d_argmax);
    else if (num_classes <= 16)
        predict_kernel<16><<<grid, 256, 0, st>>>(d_par, dim, num_classes, d_counts, n, numK,
                                                 d_keep_idx, d_prob, d_argmax);
    else
        predict_kernel<32><<<grid, 256, 0, st>>>(d_par, dim, num_classes, d_counts, n, numK,
                                                 d_keep_idx, d_prob, d_argmax);
    SQW_CUDA_CHECK(cudaGetLastError());
    return SQW_OK;
}

}  // extern "C"
