// hills_scan.cu — SURVEY.md 8 f-4: the sliding-window scan of a trained k-mer model over the
// training sequences (motifs/Discrim.java:328-444, findHills / findSeqlets) on sm_100a.
//
// For every sequence without 'N', every window of length l = minM..maxM at every offset i is
// scored with sum_k sum_j weight[base_k + canonical(kmer at i+j)] — accumulated from 0.0 in the
// reference's order (k ascending, then j ascending, Discrim.java:394-405), so the scores are
// bit-identical to the reference's — and a per-sequence list of "hills" is maintained with the
// reference's iterator logic (Discrim.java:410-427): walking the list in insertion order, an
// overlapping hill with a lower score is removed (`<` in findSeqlets, `<=` in findHills) until
// an overlapping hill with a higher score ends the walk and vetoes the candidate; survivors with
// score > hillsThresh are appended.
//
// Kernel design: one warp per sequence, persistent grid-stride loop.
//   * the bases are packed to 2-bit codes in shared memory (kmer_pack.cuh, as path 1);
//   * the weight of the canonical k-mer at every position is gathered once per k into shared
//     memory (numK doubles of weights: L1/L2 resident), so a window score is l-k+1 conflict-free
//     LDS + DADD per k;
//   * the lanes score 32 consecutive windows at a time; a ballot finds the candidates above the
//     threshold (a candidate at or below it can neither be added nor remove anything: every hill
//     in the list scores above the threshold, so such a candidate meets `higher` at its first
//     overlap), and each candidate walks the list warp-wide: two ballots give the lower / higher
//     overlapping hills, the first higher one truncates the removal set, a ballot-prefix
//     compaction rewrites the list in place;
//   * the finished list goes to [n][cap] outputs with coalesced stores.
#include "common.cuh"
#include "kmer_pack.cuh"

#include <algorithm>
#include <vector>

namespace sqw {

struct HillsParams {
    const uint8_t *bases;
    const int64_t *offsets;
    int64_t n;
    int kmin, kmax;
    const double *weights;
    int min_m, max_m;
    double thresh;
    int mode;  // 0 findSeqlets (ties keep both), 1 findHills (ties replace)
    int cap;
    int32_t *start, *len;
    double *score;
    int32_t *counts;
    uint8_t *has_n;
    int32_t *status;  // OR of kFlagBad (bad base) and kFlagOverflow
    int max_len;
    int seq_words;    // 64-bit words per packed sequence incl. pad word
    size_t warp_bytes;
};

enum : int { kFlagOverflow = 4 };
constexpr int kHillsWarps = 8;

__global__ void __launch_bounds__(kHillsWarps * 32) hills_scan_kernel(HillsParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nk = p.kmax - p.kmin + 1;
    unsigned char *mine = smem + (size_t)warp * p.warp_bytes;
    double *pw = reinterpret_cast<double *>(mine);                              // [nk][max_len]
    double *lscore = pw + (size_t)nk * p.max_len;                               // [cap]
    unsigned long long *words = reinterpret_cast<unsigned long long *>(lscore + p.cap);
    int32_t *lstart = reinterpret_cast<int32_t *>(words + p.seq_words);         // [cap]
    int32_t *lend = lstart + p.cap;                                             // [cap]
    int *flag = reinterpret_cast<int *>(lend + p.cap);

    const int64_t total_warps = (int64_t)gridDim.x * kHillsWarps;
    for (int64_t r = (int64_t)blockIdx.x * kHillsWarps + warp; r < p.n; r += total_warps) {
        const int64_t beg = p.offsets[r];
        int len = (int)(p.offsets[r + 1] - beg);
        if (len > p.max_len) {   // would overrun pw[] / words: truncate and report (status value 8)
            if (lane == 0)
                atomicOr(p.status, 8);
            len = p.max_len;
        }
        if (lane == 0)
            *flag = 0;
        __syncwarp();
        pack_sequence(p.bases + beg, len, words, lane, 32, flag);
        __syncwarp();
        const int flags = *flag;
        if (flags & kFlagN) {  // Discrim.java:337-338 / :389-390: sequences with N are skipped
            if (lane == 0) {   // (before seq2int could throw on another character)
                p.counts[r] = 0;
                p.has_n[r] = 1;
            }
            continue;
        }
        if (flags & kFlagBad) {
            if (lane == 0) {
                atomicOr(p.status, kFlagBad);
                p.counts[r] = 0;
                p.has_n[r] = 0;
            }
            continue;
        }
        // weight of the canonical k-mer starting at every position, per k
        {
            int64_t base = 0;
            for (int k = p.kmin; k <= p.kmax; ++k) {
                double *row = pw + (size_t)(k - p.kmin) * p.max_len;
                for (int pos = lane; pos + k <= len; pos += 32)
                    row[pos] = __ldg(p.weights + base + (int64_t)canonical_code(words, pos, k));
                base += (int64_t)1 << (2 * k);
            }
        }
        __syncwarp();

        int cnt = 0;           // list size (warp-uniform)
        bool overflow = false;
        for (int l = p.min_m; l <= p.max_m; ++l) {
            const int nwin = len - l + 1;
            for (int i0 = 0; i0 < nwin; i0 += 32) {
                const int i = i0 + lane;
                double s = 0.0;
                if (i < nwin) {
                    for (int k = p.kmin; k <= p.kmax; ++k) {
                        const double *row = pw + (size_t)(k - p.kmin) * p.max_len + i;
                        for (int j = 0; j < l - k + 1; ++j)
                            s = s + row[j];
                    }
                }
                unsigned cand = __ballot_sync(0xffffffffu, i < nwin && s > p.thresh);
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const double cs = __shfl_sync(0xffffffffu, s, b);
                    const int hs = i0 + b, he = hs + l - 1;
                    // walk the list in insertion order (Discrim.java:411-424)
                    bool add = true, stop = false;
                    int out = 0;
                    for (int c0 = 0; c0 < cnt; c0 += 32) {
                        const int e = c0 + lane;
                        const bool valid = e < cnt;
                        int st = 0, en = 0;
                        double sv = 0.0;
                        if (valid) {
                            st = lstart[e];
                            en = lend[e];
                            sv = lscore[e];
                        }
                        const bool ov = valid && st <= he && en >= hs;
                        const bool lower = ov && (p.mode ? (sv <= cs) : (sv < cs));
                        const bool higher = ov && !lower && sv > cs;
                        unsigned bl = __ballot_sync(0xffffffffu, lower);
                        const unsigned bh = __ballot_sync(0xffffffffu, higher);
                        if (stop) {
                            bl = 0;
                        } else if (bh) {
                            bl &= (1u << (__ffs(bh) - 1)) - 1u;  // the walk ends at the first higher hill
                            stop = true;
                            add = false;
                        }
                        const bool keep = valid && !((bl >> lane) & 1u);
                        const unsigned kb = __ballot_sync(0xffffffffu, keep);
                        const int pos = out + __popc(kb & ((1u << lane) - 1u));
                        if (keep && pos != e) {  // pos <= e and all reads of this chunk are done
                            lstart[pos] = st;
                            lend[pos] = en;
                            lscore[pos] = sv;
                        }
                        __syncwarp();
                        out += __popc(kb);
                    }
                    cnt = out;
                    if (add) {  // cs > thresh holds for every candidate (Discrim.java:425)
                        if (cnt < p.cap) {
                            if (lane == 0) {
                                lstart[cnt] = hs;
                                lend[cnt] = he;
                                lscore[cnt] = cs;
                            }
                            ++cnt;
                        } else {
                            overflow = true;
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (overflow && lane == 0)
            atomicOr(p.status, kFlagOverflow);
        for (int e = lane; e < cnt; e += 32) {
            p.start[r * p.cap + e] = lstart[e];
            p.len[r * p.cap + e] = lend[e] - lstart[e] + 1;
            p.score[r * p.cap + e] = lscore[e];
        }
        if (lane == 0) {
            p.counts[r] = cnt;
            p.has_n[r] = 0;
        }
        __syncwarp();
    }
}

static int64_t hills_num_kmers(int kmin, int kmax)
{
    if (kmin < 1 || kmax < kmin || kmax > 15)
        return -1;
    int64_t s = 0;
    for (int k = kmin; k <= kmax; ++k)
        s += (int64_t)1 << (2 * k);
    return s;
}

static int launch_hills(HillsParams p, cudaStream_t stream)
{
    if (p.n == 0)
        return SQW_OK;
    const int nk = p.kmax - p.kmin + 1;
    p.seq_words = p.max_len / 32 + 2;
    size_t wb = (size_t)nk * p.max_len * 8 + (size_t)p.cap * 8 + (size_t)p.seq_words * 8 +
                (size_t)p.cap * 8 + 16;
    wb = (wb + 15) & ~(size_t)15;
    p.warp_bytes = wb;
    const size_t smem = wb * kHillsWarps;
    int dev = 0, smem_optin = 0, n_sm = 0;  // (attribute queries: cudaGetDeviceProperties takes ms)
    SQW_CUDA_CHECK(cudaGetDevice(&dev));
    SQW_CUDA_CHECK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SQW_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (smem > (size_t)smem_optin)
        return set_error(SQW_ERR_UNSUPPORTED,
                         "hills scan: sequences of %d bases with %d k values and cap %d need %zu bytes of "
                         "shared memory per CTA (limit %zu)",
                         p.max_len, nk, p.cap, smem, (size_t)smem_optin);
    SQW_CUDA_CHECK(cudaFuncSetAttribute(hills_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    int per_sm = 1;
    SQW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hills_scan_kernel,
                                                                 kHillsWarps * 32, smem));
    per_sm = std::max(per_sm, 1);
    const int64_t need = (p.n + kHillsWarps - 1) / kHillsWarps;
    const int grid = (int)std::min<int64_t>(need, (int64_t)n_sm * per_sm);
    hills_scan_kernel<<<grid, kHillsWarps * 32, smem, stream>>>(p);
    SQW_CUDA_CHECK(cudaGetLastError());
    return SQW_OK;
}

static int check_hills_args(int64_t n, int kmin, int kmax, int min_m, int max_m, int32_t cap)
{
    if (hills_num_kmers(kmin, kmax) < 0)
        return set_error(SQW_ERR_INVALID_ARG, "invalid k range [%d,%d] (1 <= kmin <= kmax <= 15)", kmin,
                         kmax);
    if (min_m < 1 || max_m < min_m)
        return set_error(SQW_ERR_INVALID_ARG, "invalid window lengths [%d,%d]", min_m, max_m);
    if (cap < 1 || n < 0)
        return set_error(SQW_ERR_INVALID_ARG, "cap must be >= 1 and n >= 0");
    return SQW_OK;
}

}  // namespace sqw

extern "C" {

int sqw_hills_scan_dev(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n, int64_t max_len,
                       int kmin, int kmax, const double *d_kmer_weights, int min_m, int max_m,
                       double thresh, int mode, int32_t cap, int32_t *d_hill_start,
                       int32_t *d_hill_len, double *d_hill_score, int32_t *d_counts, uint8_t *d_has_n,
                       int32_t *d_status, void *stream)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if ((rc = check_hills_args(n, kmin, kmax, min_m, max_m, cap)) != SQW_OK)
        return rc;
    if (max_len < 0 || max_len > (1 << 20))
        return set_error(SQW_ERR_INVALID_ARG, "max_len out of range");
    HillsParams p{};
    p.bases = d_bases;
    p.offsets = d_offsets;
    p.n = n;
    p.kmin = kmin;
    p.kmax = kmax;
    p.weights = d_kmer_weights;
    p.min_m = min_m;
    p.max_m = max_m;
    p.thresh = thresh;
    p.mode = mode ? 1 : 0;
    p.cap = cap;
    p.start = d_hill_start;
    p.len = d_hill_len;
    p.score = d_hill_score;
    p.counts = d_counts;
    p.has_n = d_has_n;
    p.status = d_status;
    p.max_len = (int)std::max<int64_t>(max_len, 1);
    return launch_hills(p, static_cast<cudaStream_t>(stream));
}

int sqw_hills_scan(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   const double *kmer_weights, int min_m, int max_m, double thresh, int mode,
                   int32_t cap, int32_t *hill_start, int32_t *hill_len, double *hill_score,
                   int32_t *counts, uint8_t *has_n)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    if ((rc = check_hills_args(n, kmin, kmax, min_m, max_m, cap)) != SQW_OK)
        return rc;
    if (n > 0 && (!offsets || !kmer_weights || !hill_start || !hill_len || !hill_score || !counts || !has_n))
        return set_error(SQW_ERR_INVALID_ARG, "null buffer");
    if (n == 0)
        return SQW_OK;
    int64_t max_len = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t len = offsets[r + 1] - offsets[r];
        if (len < 0)
            return set_error(SQW_ERR_INVALID_ARG, "offsets must be non-decreasing");
        max_len = std::max(max_len, len);
    }
    const int64_t numK = hills_num_kmers(kmin, kmax);
    const int64_t nbases = offsets[n] - offsets[0];
    uint8_t *d_bases = nullptr, *d_hasn = nullptr;
    int64_t *d_off = nullptr;
    double *d_w = nullptr, *d_score = nullptr;
    int32_t *d_start = nullptr, *d_len = nullptr, *d_counts = nullptr, *d_status = nullptr;
    std::vector<int64_t> rel((size_t)n + 1);
    for (int64_t r = 0; r <= n; ++r)
        rel[(size_t)r] = offsets[r] - offsets[0];
    int result = SQW_OK;
    auto cleanup = [&]() {
        cudaFree(d_bases);
        cudaFree(d_hasn);
        cudaFree(d_off);
        cudaFree(d_w);
        cudaFree(d_score);
        cudaFree(d_start);
        cudaFree(d_len);
        cudaFree(d_counts);
        cudaFree(d_status);
    };
#define HS_TRY(call)                                                                         \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            result = set_error(SQW_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));       \
            cleanup();                                                                       \
            return result;                                                                   \
        }                                                                                    \
    } while (0)
    const size_t slots = (size_t)n * (size_t)cap;
    HS_TRY(cudaMalloc((void **)&d_bases, (size_t)std::max<int64_t>(nbases, 1)));
    HS_TRY(cudaMalloc((void **)&d_off, ((size_t)n + 1) * sizeof(int64_t)));
    HS_TRY(cudaMalloc((void **)&d_w, (size_t)numK * sizeof(double)));
    HS_TRY(cudaMalloc((void **)&d_start, slots * sizeof(int32_t)));
    HS_TRY(cudaMalloc((void **)&d_len, slots * sizeof(int32_t)));
    HS_TRY(cudaMalloc((void **)&d_score, slots * sizeof(double)));
    HS_TRY(cudaMalloc((void **)&d_counts, (size_t)n * sizeof(int32_t)));
    HS_TRY(cudaMalloc((void **)&d_hasn, (size_t)n));
    HS_TRY(cudaMalloc((void **)&d_status, sizeof(int32_t)));
    HS_TRY(cudaMemset(d_status, 0, sizeof(int32_t)));
    HS_TRY(cudaMemset(d_start, 0, slots * sizeof(int32_t)));
    HS_TRY(cudaMemset(d_len, 0, slots * sizeof(int32_t)));
    HS_TRY(cudaMemset(d_score, 0, slots * sizeof(double)));
    if (nbases > 0)
        HS_TRY(cudaMemcpy(d_bases, bases + offsets[0], (size_t)nbases, cudaMemcpyHostToDevice));
    HS_TRY(cudaMemcpy(d_off, rel.data(), ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    HS_TRY(cudaMemcpy(d_w, kmer_weights, (size_t)numK * sizeof(double), cudaMemcpyHostToDevice));
    rc = sqw_hills_scan_dev(d_bases, d_off, n, max_len, kmin, kmax, d_w, min_m, max_m, thresh, mode, cap,
                            d_start, d_len, d_score, d_counts, d_hasn, d_status, nullptr);
    if (rc != SQW_OK) {
        cleanup();
        return rc;
    }
    HS_TRY(cudaDeviceSynchronize());
    int32_t status = 0;
    HS_TRY(cudaMemcpy(&status, d_status, sizeof(int32_t), cudaMemcpyDeviceToHost));
    HS_TRY(cudaMemcpy(hill_start, d_start, slots * sizeof(int32_t), cudaMemcpyDeviceToHost));
    HS_TRY(cudaMemcpy(hill_len, d_len, slots * sizeof(int32_t), cudaMemcpyDeviceToHost));
    HS_TRY(cudaMemcpy(hill_score, d_score, slots * sizeof(double), cudaMemcpyDeviceToHost));
    HS_TRY(cudaMemcpy(counts, d_counts, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    HS_TRY(cudaMemcpy(has_n, d_hasn, (size_t)n, cudaMemcpyDeviceToHost));
#undef HS_TRY
    cleanup();
    if (status & kFlagBad)
        return set_error(SQW_ERR_BAD_BASE, "Nonstandard nucleotide character encountered (not ACGTN)");
    if (status & kFlagOverflow)
        return set_error(SQW_ERR_UNSUPPORTED, "a sequence produced more than cap = %d hills", cap);
    if (status & 8)
        return set_error(SQW_ERR_INVALID_ARG, "a sequence is longer than max_len");
    return SQW_OK;
}

}  // extern "C"
