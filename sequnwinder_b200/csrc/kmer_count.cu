// kmer_count.cu — path 1: reverse-complement-collapsed k-mer count matrix on sm_100a.
//
// Replaces loaddata/MakeArff.java:323-377 (KmerCounter.printPeakSeqKmerRange): for every
// sequence and every k in [kmin,kmax], every window increments column
// base_k + min(seq2int(window), seq2int(revcomp(window))) of a dense numK-wide int row
// (MakeArff.java:357-367). Encoding A0 C1 G2 T3, first base most significant.
//
// Kernel design (HBM-write bound: 4*numK bytes out vs ~L bytes in per sequence):
//   * a group of threads (>= one warp) owns one sequence; the ASCII bases are loaded coalesced,
//     mapped to 2-bit codes and packed MSB-first into 64-bit words with two warp REDUX.OR
//     reductions per 32 bases, staged in shared memory;
//   * a window of any k <= 32 is then two LDS.64 + a funnel shift; its reverse complement for
//     every k at once is brev + pair swap of the complemented word, so forward and reverse codes
//     for all k in [kmin,kmax] cost one extraction per position;
//   * counts go to a per-sequence shared-memory histogram with ATOMS (spread addresses);
//   * the finished row is streamed to HBM with 16-byte st.global.cs stores and the histogram is
//     re-zeroed in the same pass.
// Rows containing 'N' are flagged and written as zeros (the Java caller omits them,
// MakeArff.java:354-355); any other character raises the status bit (seq2int would throw).
#include "common.cuh"
#include "kmer_pack.cuh"

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace sqw {

struct KmerParams {
    const uint8_t *bases;
    const int64_t *offsets;
    int64_t n;
    int kmin, kmax;
    int64_t numK;
    int32_t *counts;
    uint8_t *has_n;
    int32_t *status;
    int rows_per_block;
    int seq_words;  // 64-bit words reserved per sequence (incl. one pad word)
    int64_t max_len;  // the caller's bound on a row's length: the shared-memory buffers are sized from it
};

// A row longer than the caller's max_len would overrun its packed words / histogram: it is
// truncated to max_len and status bit 2 (value 4) is raised - the host wrappers turn that into
// SQW_ERR_INVALID_ARG, a _dev caller sees it in its status word.
constexpr int kStatusTooLong = 4;
__device__ __forceinline__ int64_t clamp_len(const KmerParams &p, int64_t len)
{
    if (len > p.max_len) {
        atomicOr(p.status, kStatusTooLong);
        return p.max_len;
    }
    return len;
}

// One window position: forward code = top 2k bits of `comb`, reverse-complement code = low 2k
// bits of rev2_64(~comb). Target is either shared (ATOMS) or global (RED) memory.
template <typename Ptr>
__device__ __forceinline__ void count_position(const unsigned long long *__restrict__ words,
                                               int64_t i, int64_t len, int kmin, int kmax,
                                               Ptr hist)
{
    const int64_t w = i >> 5;
    const int s = (int)(i & 31) << 1;
    const unsigned long long hi = words[w], lo = words[w + 1];
    const unsigned long long comb = s ? ((hi << s) | (lo >> (64 - s))) : hi;
    const unsigned long long rc_all = rev2_64(~comb);
    int64_t base = 0;
    for (int k = kmin; k <= kmax; ++k) {
        if (i + k <= len) {
            const unsigned long long fwd = comb >> (64 - 2 * k);
            const unsigned long long rc = (k == 32) ? rc_all : (rc_all & ((1ull << (2 * k)) - 1ull));
            atomicAdd(hist + base + (int64_t)(fwd < rc ? fwd : rc), 1);
        }
        base += (int64_t)1 << (2 * k);
    }
}

// Shared-memory histogram path: rows_per_block sequences per CTA, blockDim/rows_per_block
// threads each, persistent grid-stride loop over row batches.
__global__ void __launch_bounds__(256) kmer_count_smem_kernel(KmerParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int R = p.rows_per_block;
    const int tpr = blockDim.x / R;
    const int g = threadIdx.x / tpr;
    const int t = threadIdx.x - g * tpr;
    const int64_t numK = p.numK;
    int32_t *hist_all = reinterpret_cast<int32_t *>(smem);
    unsigned long long *words_all =
        reinterpret_cast<unsigned long long *>(smem + (size_t)R * numK * sizeof(int32_t));
    int *flags = reinterpret_cast<int *>(words_all + (size_t)R * p.seq_words);
    int32_t *hist = hist_all + (size_t)g * numK;
    unsigned long long *words = words_all + (size_t)g * p.seq_words;

    {
        int4 *h4 = reinterpret_cast<int4 *>(hist_all);
        const int64_t n4 = (int64_t)R * numK / 4;
        for (int64_t i = threadIdx.x; i < n4; i += blockDim.x)
            h4[i] = make_int4(0, 0, 0, 0);
    }

    const int64_t nbatches = (p.n + R - 1) / R;
    for (int64_t b = blockIdx.x; b < nbatches; b += gridDim.x) {
        const int64_t row = b * R + g;
        const bool valid = row < p.n;
        int64_t beg = 0, len = 0;
        if (t == 0)
            flags[g] = 0;
        __syncthreads();  // histogram zeroed / re-zeroed, flags cleared
        if (valid) {
            beg = p.offsets[row];
            len = clamp_len(p, p.offsets[row + 1] - beg);
            pack_sequence(p.bases + beg, len, words, t, tpr, &flags[g]);
        }
        __syncthreads();
        const int fl = flags[g];
        if (valid && fl == 0) {
            for (int64_t i = t; i + p.kmin <= len; i += tpr)
                count_position(words, i, len, p.kmin, p.kmax, hist);
        }
        __syncthreads();
        if (valid) {
            int4 *src = reinterpret_cast<int4 *>(hist);
            int4 *dst = reinterpret_cast<int4 *>(p.counts + row * numK);
            const int n4 = (int)(numK >> 2);
            for (int i = t; i < n4; i += tpr) {
                const int4 v = src[i];
                __stcs(dst + i, v);
                if (v.x | v.y | v.z | v.w)
                    src[i] = make_int4(0, 0, 0, 0);
            }
            if (t == 0) {
                p.has_n[row] = (fl & kFlagN) ? 1 : 0;
                if ((fl & kFlagBad) && !(fl & kFlagN))
                    atomicOr(p.status, 1);
            }
        }
    }
}

// One warp per sequence, no CTA-wide barrier: for short histograms (k = 4..5: 5 KB per row) every
// warp walks its own rows (pack -> count -> store) and only ever synchronises with itself, so
// the eight warps of a CTA drift apart and their pack / count / store phases overlap. Measured
// at cfg-2 (20 000 x 1 280): 31 us against 37 us with the CTA-synchronous kernel above.
__global__ void __launch_bounds__(256) kmer_count_warp_kernel(KmerParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warps = blockDim.x >> 5;
    const int g = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    const int64_t numK = p.numK;
    int32_t *hist = reinterpret_cast<int32_t *>(smem) + (size_t)g * numK;
    unsigned long long *words =
        reinterpret_cast<unsigned long long *>(smem + (size_t)warps * numK * sizeof(int32_t)) +
        (size_t)g * p.seq_words;
    int *flags = reinterpret_cast<int *>(
        reinterpret_cast<unsigned long long *>(smem + (size_t)warps * numK * sizeof(int32_t)) +
        (size_t)warps * p.seq_words);
    const int n4 = (int)(numK >> 2);
    int4 *h4 = reinterpret_cast<int4 *>(hist);
    for (int i = t; i < n4; i += 32)
        h4[i] = make_int4(0, 0, 0, 0);
    for (int64_t row = (int64_t)blockIdx.x * warps + g; row < p.n; row += (int64_t)gridDim.x * warps) {
        if (t == 0)
            flags[g] = 0;
        __syncwarp();
        const int64_t beg = p.offsets[row];
        const int64_t len = clamp_len(p, p.offsets[row + 1] - beg);
        pack_sequence(p.bases + beg, len, words, t, 32, &flags[g]);
        __syncwarp();
        const int fl = flags[g];
        if (fl == 0)
            for (int64_t i = t; i + p.kmin <= len; i += 32)
                count_position(words, i, len, p.kmin, p.kmax, hist);
        __syncwarp();
        int4 *dst = reinterpret_cast<int4 *>(p.counts + row * numK);
        for (int i = t; i < n4; i += 32) {
            const int4 v = h4[i];
            __stcs(dst + i, v);
            if (v.x | v.y | v.z | v.w)
                h4[i] = make_int4(0, 0, 0, 0);
        }
        if (t == 0) {
            p.has_n[row] = (fl & kFlagN) ? 1 : 0;
            if ((fl & kFlagBad) && !(fl & kFlagN))
                atomicOr(p.status, 1);
        }
        __syncwarp();
    }
}

// 32-bit variant of the warp-per-sequence kernel for kmax <= 16 (the API caps k at 15) with the k
// range known at compile time (NK = kmax - kmin + 1 <= 5). Same results; what changes is the
// instruction count and the exposed latency per row, which bounded the kernel above at the
// k = 4..5 shape (ncu / CUDA events on B200: 4.2 TB/s of 6.56, ~750 warp instructions per row):
//   * codes and the N / bad-character flags come from a 256-entry table in shared memory, the
//     bytes of up to 8 chunks of 32 bases are loaded before the first one is used (one exposed
//     global latency per row instead of one per chunk), the flags are reduced in registers;
//   * the packed stream is kept as 32-bit words (16 bases each, first base most significant): a
//     window of k <= 16 bases is two LDS.32 + one funnel shift, its reverse complement one BREV +
//     pair swap, and every per-k step (shift, mask, min, bounds test, address) is 32-bit;
//   * the histogram is re-zeroed unconditionally while the row is streamed out.
template <int NK>
__global__ void __launch_bounds__(256) kmer_count_warp32_kernel(KmerParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned char lut[256];
    const int warps = blockDim.x >> 5;
    const int g = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    const int numK = (int)p.numK;
    const int words32 = 2 * p.seq_words;  // 32-bit words reserved per sequence (>= 2 pad words)
    int32_t *hist = reinterpret_cast<int32_t *>(smem) + (size_t)g * numK;
    unsigned *words = reinterpret_cast<unsigned *>(smem + (size_t)warps * numK * sizeof(int32_t)) +
                      (size_t)g * words32;
    {
        // table: bits 0-1 code (A0 C1 G2 T3, either case), bit 2 'N' / 'n', bit 3 anything else
        const unsigned ch = threadIdx.x & 255u, up = ch & 0xDFu;
        unsigned code = (up >> 1) & 3u;
        code ^= (code >> 1);
        const bool ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
        lut[ch] = ok ? (unsigned char)code : (unsigned char)((up == 'N') ? 4u : 8u);
    }
    const int n4 = numK >> 2;
    int4 *h4 = reinterpret_cast<int4 *>(hist);
    for (int i = t; i < n4; i += 32)
        h4[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    const int kmin = p.kmin;
    // shift of the forward code / mask of the reverse-complement code / histogram offset per k
    int fsh[NK], boff[NK];
    unsigned msk[NK];
#pragma unroll
    for (int j = 0, b = 0; j < NK; ++j) {
        const int k = kmin + j;
        fsh[j] = 32 - 2 * k;
        msk[j] = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
        boff[j] = b;
        b += 1 << (2 * k);
    }
    const unsigned sh = 30u - 2u * (unsigned)(t & 15);
    for (int64_t row = (int64_t)blockIdx.x * warps + g; row < p.n; row += (int64_t)gridDim.x * warps) {
        const int64_t beg = p.offsets[row];
        const int len = (int)clamp_len(p, p.offsets[row + 1] - beg);
        const uint8_t *seq = p.bases + beg;
        const int nchunks = (len + 31) >> 5;
        unsigned flags = 0;
        for (int c0 = 0; c0 < nchunks; c0 += 8) {
            unsigned ch[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int pos = ((c0 + u) << 5) + t;
                ch[u] = (pos < len) ? (unsigned)__ldg(seq + pos) : (unsigned)'A';
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (c0 + u < nchunks) {  // warp uniform
                    const unsigned e = lut[ch[u]];
                    flags |= e;
                    const unsigned v = (e & 3u) << sh;
                    const unsigned hi = __reduce_or_sync(0xffffffffu, t < 16 ? v : 0u);
                    const unsigned lo = __reduce_or_sync(0xffffffffu, t >= 16 ? v : 0u);
                    if (t == 0)
                        *reinterpret_cast<uint2 *>(words + 2 * (c0 + u)) = make_uint2(hi, lo);
                }
            }
        }
        if (t == 0)
            *reinterpret_cast<uint2 *>(words + 2 * nchunks) = make_uint2(0u, 0u);  // pad for the funnel shift
        const unsigned fl = __reduce_or_sync(0xffffffffu, flags) >> 2;  // bit 0: N, bit 1: bad
        __syncwarp();
        if (fl == 0) {
            for (int i = t; i + kmin <= len; i += 32) {
                const int w = i >> 4;
                const unsigned cur = words[w], nxt = words[w + 1];
                const unsigned comb = __funnelshift_l(nxt, cur, (i & 15) << 1);
                unsigned rc = __brev(~comb);
                rc = ((rc >> 1) & 0x55555555u) | ((rc & 0x55555555u) << 1);
#pragma unroll
                for (int j = 0; j < NK; ++j) {
                    if (i + kmin + j <= len) {
                        const unsigned f = comb >> fsh[j], r = rc & msk[j];
                        atomicAdd(hist + boff[j] + (int)min(f, r), 1);
                    }
                }
            }
        }
        __syncwarp();
        int4 *dst = reinterpret_cast<int4 *>(p.counts + row * p.numK);
        for (int i = t; i < n4; i += 32) {
            __stcs(dst + i, h4[i]);
            h4[i] = make_int4(0, 0, 0, 0);
        }
        if (t == 0) {
            p.has_n[row] = (fl & 1u) ? 1 : 0;
            if ((fl & 2u) && !(fl & 1u))
                atomicOr(p.status, 1);
        }
        __syncwarp();
    }
}

// Fallback for histograms that do not fit in shared memory (numK*4 > ~200 KB, i.e. k >= 8):
// the output is cleared with a memset and one warp per sequence issues global RED.ADDs.
__global__ void __launch_bounds__(256) kmer_count_global_kernel(KmerParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warps = blockDim.x >> 5;
    const int g = threadIdx.x >> 5;
    const int t = threadIdx.x & 31;
    unsigned long long *words =
        reinterpret_cast<unsigned long long *>(smem) + (size_t)g * p.seq_words;
    int *flags = reinterpret_cast<int *>(reinterpret_cast<unsigned long long *>(smem) +
                                         (size_t)warps * p.seq_words);
    for (int64_t row = (int64_t)blockIdx.x * warps + g; row < p.n;
         row += (int64_t)gridDim.x * warps) {
        if (t == 0)
            flags[g] = 0;
        __syncwarp();
        const int64_t beg = p.offsets[row];
        const int64_t len = clamp_len(p, p.offsets[row + 1] - beg);
        pack_sequence(p.bases + beg, len, words, t, 32, &flags[g]);
        __syncwarp();
        const int fl = flags[g];
        if (fl == 0) {
            int32_t *out = p.counts + row * p.numK;
            for (int64_t i = t; i + p.kmin <= len; i += 32)
                count_position(words, i, len, p.kmin, p.kmax, out);
        }
        if (t == 0) {
            p.has_n[row] = (fl & kFlagN) ? 1 : 0;
            if ((fl & kFlagBad) && !(fl & kFlagN))
                atomicOr(p.status, 1);
        }
        __syncwarp();
    }
}

static int64_t num_kmers(int kmin, int kmax)
{
    if (kmin < 1 || kmax < kmin || kmax > 15)
        return -1;
    int64_t numK = 0;
    for (int k = kmin; k <= kmax; ++k)
        numK += (int64_t)1 << (2 * k);
    return numK;
}

// cudaFuncSetAttribute + the occupancy query cost tens of microseconds of host time, as much as
// the kernel itself at cfg-2: remember them per (kernel, shared-memory size) and device.
template <typename Kernel>
static int cached_occupancy(Kernel kernel, int which, int threads, size_t smem, int *per_sm)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int, size_t>, int> cache;
    int dev = 0;
    SQW_CUDA_CHECK(cudaGetDevice(&dev));
    static std::map<std::pair<int, int>, size_t> attr_set;  // (device, kernel) -> size last set
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = attr_set[std::make_pair(dev, which)];
    if (cur != smem) {  // the attribute is per function: re-set it when the size changes
        SQW_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    const auto key = std::make_tuple(dev, which, smem);
    auto it = cache.find(key);
    if (it == cache.end()) {
        int v = 1;
        SQW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem));
        it = cache.emplace(key, std::max(v, 1)).first;
    }
    *per_sm = it->second;
    return SQW_OK;
}

static int launch_kmer(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n,
                       int64_t max_len, int kmin, int kmax, int32_t *d_counts, uint8_t *d_has_n,
                       int32_t *d_status, cudaStream_t stream)
{
    const int64_t numK = num_kmers(kmin, kmax);
    if (numK < 0)
        return set_error(SQW_ERR_INVALID_ARG, "invalid k range [%d,%d] (1 <= kmin <= kmax <= 15)",
                         kmin, kmax);
    if (n < 0 || max_len < 0)
        return set_error(SQW_ERR_INVALID_ARG, "negative size");
    if (n == 0)
        return SQW_OK;
    if ((reinterpret_cast<uintptr_t>(d_counts) & 15) != 0)
        return set_error(SQW_ERR_INVALID_ARG, "counts must be 16-byte aligned");

    KmerParams p;
    p.bases = d_bases;
    p.offsets = d_offsets;
    p.n = n;
    p.kmin = kmin;
    p.kmax = kmax;
    p.numK = numK;
    p.counts = d_counts;
    p.has_n = d_has_n;
    p.status = d_status;
    p.max_len = max_len;
    p.seq_words = (int)((max_len + 31) / 32 + 1);

    const size_t seq_bytes = (size_t)p.seq_words * 8;
    const size_t hist_bytes = (size_t)numK * 4;
    const size_t kMaxSmem = 200 * 1024;
    const int threads = 256;

    if (hist_bytes + seq_bytes + 16 <= kMaxSmem && (numK % 4) == 0) {
        // rows per CTA: aim for <= ~48 KB per CTA (4-5 CTAs per SM overlap pack / count / store)
        int R = 8;
        while (R > 1 && (size_t)R * (hist_bytes + seq_bytes) + 64 > 56 * 1024)
            R >>= 1;
        if (R == 8) {
            // one warp per row: warp-synchronous kernel, 8 rows in flight per CTA
            p.rows_per_block = 8;
            const size_t wsmem = 8 * (hist_bytes + seq_bytes) + 8 * sizeof(int) + 16;
            int per_sm_w = 1;
            const int64_t nblk = (n + 7) / 8;
            const int nk = kmax - kmin + 1;
            if (nk <= 5 && max_len < (1 << 30) && !getenv("SQW_KMER_WARP64")) {
                // k range known at compile time, 32-bit stream (see kmer_count_warp32_kernel)
#define SQW_LAUNCH_W32(NK)                                                                      \
    case NK: {                                                                                  \
        int rc = cached_occupancy(kmer_count_warp32_kernel<NK>, 10 + NK, threads, wsmem, &per_sm_w); \
        if (rc != SQW_OK)                                                                       \
            return rc;                                                                          \
        const int gridw = (int)std::min<int64_t>(nblk, (int64_t)kNumSMs * per_sm_w);            \
        kmer_count_warp32_kernel<NK><<<gridw, threads, wsmem, stream>>>(p);                     \
        break;                                                                                  \
    }
                switch (nk) {
                    SQW_LAUNCH_W32(1)
                    SQW_LAUNCH_W32(2)
                    SQW_LAUNCH_W32(3)
                    SQW_LAUNCH_W32(4)
                    SQW_LAUNCH_W32(5)
                }
#undef SQW_LAUNCH_W32
                SQW_CUDA_CHECK(cudaGetLastError());
                return SQW_OK;
            }
            int rc = cached_occupancy(kmer_count_warp_kernel, 0, threads, wsmem, &per_sm_w);
            if (rc != SQW_OK)
                return rc;
            const int gridw = (int)std::min<int64_t>(nblk, (int64_t)kNumSMs * per_sm_w);
            kmer_count_warp_kernel<<<gridw, threads, wsmem, stream>>>(p);
            SQW_CUDA_CHECK(cudaGetLastError());
            return SQW_OK;
        }
        p.rows_per_block = R;
        const size_t smem = (size_t)R * (hist_bytes + seq_bytes) + (size_t)R * sizeof(int) + 16;
        int per_sm = 1;
        int rc = cached_occupancy(kmer_count_smem_kernel, 1, threads, smem, &per_sm);
        if (rc != SQW_OK)
            return rc;
        const int64_t nbatches = (n + R - 1) / R;
        const int grid = (int)std::min<int64_t>(nbatches, (int64_t)kNumSMs * per_sm);
        kmer_count_smem_kernel<<<grid, threads, smem, stream>>>(p);
    } else {
        if ((size_t)(threads / 32) * seq_bytes > kMaxSmem)
            return set_error(SQW_ERR_UNSUPPORTED, "sequence of %lld bases is too long",
                             (long long)max_len);
        SQW_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, (size_t)n * numK * 4, stream));
        p.rows_per_block = threads / 32;
        const size_t smem = (size_t)(threads / 32) * (seq_bytes + sizeof(int)) + 16;
        SQW_CUDA_CHECK(cudaFuncSetAttribute(kmer_count_global_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        const int64_t nb = (n + threads / 32 - 1) / (threads / 32);
        const int grid = (int)std::min<int64_t>(nb, (int64_t)kNumSMs * 8);
        kmer_count_global_kernel<<<grid, threads, smem, stream>>>(p);
    }
    SQW_CUDA_CHECK(cudaGetLastError());
    return SQW_OK;
}


// ------------------------------------------------------------------------------------------------
// Host-buffer path (what the Java side calls through JNI: MakeArff.java:347-368 with the sequences
// and the int[n][numK] matrix in host memory).
//
// The dense matrix is 4 numK bytes per row on the host side of PCIe (5 120 B at k = 4..5) but only
// the canonical columns (k-mer <= reverse complement: 648 of 1 280) can be non-zero, and a count
// never exceeds the number of windows of a row. So when max_len - kmin + 1 <= 255 the device
// compacts the rows to one byte per canonical column, the copy goes into a pinned staging buffer
// (648 B per row instead of 5 120 B through pageable memory), and host threads expand it into the
// caller's dense int32 rows (zero-fill + scatter) while the next chunk is in flight. Bit-exact
// with the dense path by construction. Buffers, streams and the pinned staging ring are
// per-thread and persist between calls.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kmer_compact_kernel(const int32_t *dense, int64_t rows, int64_t numK, const int32_t *canon, int ncanon,
                    uint8_t *out)
{
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const int32_t *in = dense + r * numK;
        uint8_t *o = out + r * ncanon;
        for (int j = threadIdx.x; j < ncanon; j += blockDim.x)
            o[j] = (uint8_t)in[canon[j]];
    }
}

struct KmerHostCtx {
    int device = -1;
    cudaStream_t st[2] = {nullptr, nullptr};
    uint8_t *d_bases[2] = {nullptr, nullptr};
    int64_t *d_off[2] = {nullptr, nullptr};
    int32_t *d_counts[2] = {nullptr, nullptr};
    uint8_t *d_hasn[2] = {nullptr, nullptr};
    uint8_t *d_compact[2] = {nullptr, nullptr};
    uint8_t *h_stage[2] = {nullptr, nullptr};   // pinned: compact rows | has_n
    int64_t *h_off[2] = {nullptr, nullptr};     // pinned: chunk-relative offsets
    size_t cap_bases = 0, cap_rows = 0, cap_dense = 0, cap_stage = 0;
    int32_t *d_status = nullptr;
    int32_t *d_canon = nullptr;
    std::vector<int32_t> canon;
    int ck_min = -1, ck_max = -1;

    void release()
    {
        if (device < 0)
            return;
        for (int i = 0; i < 2; ++i) {
            if (d_bases[i]) cudaFree(d_bases[i]);
            if (d_off[i]) cudaFree(d_off[i]);
            if (d_counts[i]) cudaFree(d_counts[i]);
            if (d_hasn[i]) cudaFree(d_hasn[i]);
            if (d_compact[i]) cudaFree(d_compact[i]);
            if (h_stage[i]) cudaFreeHost(h_stage[i]);
            if (h_off[i]) cudaFreeHost(h_off[i]);
            if (st[i]) cudaStreamDestroy(st[i]);
            d_bases[i] = nullptr; d_off[i] = nullptr; d_counts[i] = nullptr; d_hasn[i] = nullptr;
            d_compact[i] = nullptr; h_stage[i] = nullptr; h_off[i] = nullptr; st[i] = nullptr;
        }
        if (d_status) cudaFree(d_status);
        if (d_canon) cudaFree(d_canon);
        d_status = nullptr; d_canon = nullptr;
        cap_bases = cap_rows = cap_dense = cap_stage = 0;
        ck_min = ck_max = -1;
        device = -1;
        cudaGetLastError();
    }
    ~KmerHostCtx() { release(); }  // (at process exit the calls fail harmlessly)
};

static void canonical_columns(int kmin, int kmax, std::vector<int32_t> &out)
{
    out.clear();
    int64_t base = 0;
    for (int k = kmin; k <= kmax; ++k) {
        const int64_t nk = (int64_t)1 << (2 * k);
        for (int64_t v = 0; v < nk; ++v) {
            int64_t rcv = 0, t = v;
            for (int i = 0; i < k; ++i) {  // reverse complement: reverse the base order, 3 - base
                rcv = (rcv << 2) | (3 - (t & 3));
                t >>= 2;
            }
            if (v <= rcv)
                out.push_back((int32_t)(base + v));
        }
        base += nk;
    }
}

// Rows [a, b) of a compact chunk into the caller's dense int32 rows. The dense row is built in a
// cache-resident scratch row (zero-fill + scatter of the canonical columns) and leaves with
// non-temporal 16-byte stores: the caller's matrix (102 MB at cfg-2) is written once and never read
// here, so the stores should not first pull every line into the cache (measured on the host alone,
// 8 threads: 30 against 20 GB/s with memset + scatter in place).
static void expand_range(const uint8_t *stage, int64_t a, int64_t b, const int32_t *canon, int ncanon,
                         int64_t numK, int32_t *out, std::vector<int32_t> &scratch)
{
#if defined(__SSE2__)
    if (numK * 4 <= (256 << 10) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (numK & 3) == 0) {
        scratch.resize((size_t)numK);
        int32_t *tmp = scratch.data();
        for (int64_t r = a; r < b; ++r) {
            memset(tmp, 0, (size_t)numK * sizeof(int32_t));
            const uint8_t *in = stage + r * ncanon;
            for (int j = 0; j < ncanon; ++j)
                tmp[canon[j]] = in[j];
            __m128i *o = reinterpret_cast<__m128i *>(out + r * numK);
            for (int64_t c = 0; c < numK / 4; ++c)
                _mm_stream_si128(o + c, _mm_loadu_si128(reinterpret_cast<const __m128i *>(tmp) + c));
        }
        _mm_sfence();
        return;
    }
#endif
    for (int64_t r = a; r < b; ++r) {
        int32_t *o = out + r * numK;
        memset(o, 0, (size_t)numK * sizeof(int32_t));
        const uint8_t *in = stage + r * ncanon;
        for (int j = 0; j < ncanon; ++j)
            o[canon[j]] = in[j];
    }
}

// Host threads of the expansion: created once per process and parked on a condition variable
// (creating 15 threads per call cost a cfg-2 call up to a third of its 1.6 - 3 ms). One job at a
// time; worker t of the first nt takes part in a job.
class ExpandPool {
public:
    explicit ExpandPool(int n)
    {
        for (int t = 0; t < n; ++t)
            workers_.emplace_back([this, t] { loop(t); });
    }
    ~ExpandPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &w : workers_)
            w.join();
    }
    int size() const { return (int)workers_.size(); }
    template <class F>
    void run(int nt, const F &f)
    {
        std::lock_guard<std::mutex> one_job(job_mu_);
        std::unique_lock<std::mutex> lk(mu_);
        job_ = [&f](int t) { f(t); };
        active_ = nt;
        pending_ = nt;
        ++generation_;
        cv_work_.notify_all();
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_)
                return;
            seen = generation_;
            if (t >= active_)
                continue;
            lk.unlock();
            job_(t);           // (job_ stays valid until pending_ reaches 0)
            lk.lock();
            if (--pending_ == 0)
                cv_done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_work_, cv_done_;
    std::function<void(int)> job_;
    int active_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

// `rows` compact rows into the caller's dense int32 rows on host threads: at least 2 MB of dense
// output per thread, and one core left to the thread that drives the GPU (measured on the pool's
// 16-core hosts, cfg-2: 1.6 ms per call with 15 threads, 1.7 with 16, 2.5 with 8).
static void expand_rows(const uint8_t *stage, int64_t rows, const std::vector<int32_t> &canon,
                        int64_t numK, int32_t *out)
{
    const int ncanon = (int)canon.size();
    const int64_t bytes = rows * numK * 4;
    static ExpandPool pool((int)std::min(31u, std::max(2u, std::thread::hardware_concurrency()) - 1));
    const int nt = (int)std::min<int64_t>(pool.size(), std::max<int64_t>(1, bytes >> 21));
    if (nt <= 1) {
        std::vector<int32_t> scratch;
        expand_range(stage, 0, rows, canon.data(), ncanon, numK, out, scratch);
        return;
    }
    pool.run(nt, [&](int t) {
        thread_local std::vector<int32_t> scratch;
        expand_range(stage, rows * t / nt, rows * (t + 1) / nt, canon.data(), ncanon, numK, out, scratch);
    });
}

static int host_kmer_count(const char *bases, const int64_t *offsets, int64_t n, int64_t max_len,
                           int kmin, int kmax, int64_t numK, int32_t *counts, uint8_t *has_n)
{
    static thread_local KmerHostCtx ctx;
    int dev = 0;
    SQW_CUDA_CHECK(cudaGetDevice(&dev));
    if (ctx.device != dev) {
        ctx.release();
        ctx.device = dev;
    }
    const bool compact = (max_len - kmin + 1) <= 255;
    if (ctx.ck_min != kmin || ctx.ck_max != kmax) {
        canonical_columns(kmin, kmax, ctx.canon);
        if (ctx.d_canon) cudaFree(ctx.d_canon);
        ctx.d_canon = nullptr;
        SQW_CUDA_CHECK(cudaMalloc(&ctx.d_canon, ctx.canon.size() * sizeof(int32_t)));
        SQW_CUDA_CHECK(cudaMemcpy(ctx.d_canon, ctx.canon.data(), ctx.canon.size() * sizeof(int32_t),
                                  cudaMemcpyHostToDevice));
        ctx.ck_min = kmin;
        ctx.ck_max = kmax;
    }
    const int64_t ncanon = (int64_t)ctx.canon.size();
    // Row chunks: the dense device rows of a chunk <= 512 MiB
    const int64_t row_bytes = numK * 4;
    int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)512 << 20) / row_bytes);
    chunk_rows = std::min(chunk_rows, n);
    int64_t max_chunk_bases = 1;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
        const int64_t r1 = std::min(n, r0 + chunk_rows);
        max_chunk_bases = std::max(max_chunk_bases, offsets[r1] - offsets[r0]);
    }
    // (re)allocate: grow only
    if (!ctx.d_status)
        SQW_CUDA_CHECK(cudaMalloc(&ctx.d_status, sizeof(int32_t)));
    const size_t stage_row = compact ? (size_t)ncanon + 1 : 1;   // + has_n byte per row
    const bool grow = (size_t)max_chunk_bases > ctx.cap_bases || (size_t)chunk_rows > ctx.cap_rows ||
                      (size_t)chunk_rows * row_bytes > ctx.cap_dense ||
                      (size_t)chunk_rows * stage_row > ctx.cap_stage;
    if (grow) {
        const size_t cb = std::max(ctx.cap_bases, (size_t)max_chunk_bases);
        const size_t cr = std::max(ctx.cap_rows, (size_t)chunk_rows);
        const size_t cd = std::max(ctx.cap_dense, (size_t)chunk_rows * row_bytes);
        const size_t cs = std::max(ctx.cap_stage, (size_t)chunk_rows * std::max<size_t>(stage_row, (size_t)ncanon + 1));
        const int device = ctx.device;
        std::vector<int32_t> canon_keep = ctx.canon;
        ctx.release();
        ctx.device = device;
        ctx.canon = canon_keep;
        SQW_CUDA_CHECK(cudaMalloc(&ctx.d_status, sizeof(int32_t)));
        SQW_CUDA_CHECK(cudaMalloc(&ctx.d_canon, ctx.canon.size() * sizeof(int32_t)));
        SQW_CUDA_CHECK(cudaMemcpy(ctx.d_canon, ctx.canon.data(), ctx.canon.size() * sizeof(int32_t),
                                  cudaMemcpyHostToDevice));
        ctx.ck_min = kmin;
        ctx.ck_max = kmax;
        for (int i = 0; i < 2; ++i) {
            SQW_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx.st[i], cudaStreamNonBlocking));
            SQW_CUDA_CHECK(cudaMalloc(&ctx.d_bases[i], cb));
            SQW_CUDA_CHECK(cudaMalloc(&ctx.d_off[i], (cr + 1) * sizeof(int64_t)));
            SQW_CUDA_CHECK(cudaMalloc(&ctx.d_counts[i], cd));
            SQW_CUDA_CHECK(cudaMalloc(&ctx.d_hasn[i], cr));
            SQW_CUDA_CHECK(cudaMalloc(&ctx.d_compact[i], cs));
            SQW_CUDA_CHECK(cudaMallocHost((void **)&ctx.h_stage[i], cs));
            SQW_CUDA_CHECK(cudaMallocHost((void **)&ctx.h_off[i], (cr + 1) * sizeof(int64_t)));
        }
        ctx.cap_bases = cb;
        ctx.cap_rows = cr;
        ctx.cap_dense = cd;
        ctx.cap_stage = cs;
    }
    SQW_CUDA_CHECK(cudaMemsetAsync(ctx.d_status, 0, sizeof(int32_t), ctx.st[0]));
    SQW_CUDA_CHECK(cudaStreamSynchronize(ctx.st[0]));

    // chunk i is enqueued (H2D, count, compact, D2H into pinned memory) before chunk i-1 is
    // expanded on the host, so the copies of one chunk overlap the host work of the other
    // (Splitting a chunk's compact copy into pieces with an event each, so that the expansion of
    // piece s overlaps the copy of piece s+1, was measured and gained nothing at cfg-2 - 1.8 against
    // 1.6 ms per call: the copy engine and the expanding threads share the host's memory bandwidth.)
    struct Pending { int64_t r0 = 0, rows = 0; bool live = false; } pend[2];
    auto finish = [&](int b) -> int {
        if (!pend[b].live)
            return SQW_OK;
        SQW_CUDA_CHECK(cudaStreamSynchronize(ctx.st[b]));
        const int64_t r0 = pend[b].r0, rows = pend[b].rows;
        if (compact) {
            memcpy(has_n + r0, ctx.h_stage[b] + (size_t)rows * ncanon, (size_t)rows);
            expand_rows(ctx.h_stage[b], rows, ctx.canon, numK, counts + r0 * numK);
        }
        pend[b].live = false;
        return SQW_OK;
    };
    int buf = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, buf ^= 1) {
        int rc = finish(buf);  // the previous use of this buffer pair
        if (rc != SQW_OK)
            return rc;
        const int64_t r1 = std::min(n, r0 + chunk_rows);
        const int64_t rows = r1 - r0;
        const int64_t b0 = offsets[r0], nb = offsets[r1] - offsets[r0];
        cudaStream_t st = ctx.st[buf];
        for (int64_t r = 0; r <= rows; ++r)
            ctx.h_off[buf][r] = offsets[r0 + r] - b0;
        if (nb > 0)
            SQW_CUDA_CHECK(cudaMemcpyAsync(ctx.d_bases[buf], bases + b0, (size_t)nb, cudaMemcpyHostToDevice, st));
        SQW_CUDA_CHECK(cudaMemcpyAsync(ctx.d_off[buf], ctx.h_off[buf], (size_t)(rows + 1) * sizeof(int64_t),
                                       cudaMemcpyHostToDevice, st));
        rc = launch_kmer(ctx.d_bases[buf], ctx.d_off[buf], rows, max_len, kmin, kmax, ctx.d_counts[buf],
                         ctx.d_hasn[buf], ctx.d_status, st);
        if (rc != SQW_OK)
            return rc;
        if (compact) {
            const int grid = (int)std::min<int64_t>(rows, (int64_t)kNumSMs * 16);
            kmer_compact_kernel<<<grid, 256, 0, st>>>(ctx.d_counts[buf], rows, numK, ctx.d_canon, (int)ncanon,
                                                      ctx.d_compact[buf]);
            SQW_CUDA_CHECK(cudaGetLastError());
            SQW_CUDA_CHECK(cudaMemcpyAsync(ctx.h_stage[buf], ctx.d_compact[buf], (size_t)rows * ncanon,
                                           cudaMemcpyDeviceToHost, st));
            SQW_CUDA_CHECK(cudaMemcpyAsync(ctx.h_stage[buf] + (size_t)rows * ncanon, ctx.d_hasn[buf], (size_t)rows,
                                           cudaMemcpyDeviceToHost, st));
        } else {
            // rows too long for byte counts: the dense rows go straight into the caller's buffer
            SQW_CUDA_CHECK(cudaMemcpyAsync(counts + r0 * numK, ctx.d_counts[buf], (size_t)rows * row_bytes,
                                           cudaMemcpyDeviceToHost, st));
            SQW_CUDA_CHECK(cudaMemcpyAsync(has_n + r0, ctx.d_hasn[buf], (size_t)rows, cudaMemcpyDeviceToHost, st));
        }
        pend[buf].r0 = r0;
        pend[buf].rows = rows;
        pend[buf].live = true;
    }
    for (int b = 0; b < 2; ++b) {
        int rc = finish(buf ^ b);   // oldest first
        if (rc != SQW_OK)
            return rc;
    }
    int32_t status = 0;
    SQW_CUDA_CHECK(cudaMemcpy(&status, ctx.d_status, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (status & 1)
        return set_error(SQW_ERR_BAD_BASE, "Nonstandard nucleotide character encountered");
    if (status & kStatusTooLong)
        return set_error(SQW_ERR_INVALID_ARG, "a sequence is longer than max_len");
    return SQW_OK;
}

}  // namespace sqw

extern "C" {

int64_t sqw_num_kmers(int kmin, int kmax) { return sqw::num_kmers(kmin, kmax); }

int sqw_kmer_count_dev(const uint8_t *d_bases, const int64_t *d_offsets, int64_t n,
                       int64_t max_len, int kmin, int kmax, int32_t *d_counts, uint8_t *d_has_n,
                       int32_t *d_status, void *stream)
{
    int rc = sqw::require_device();
    if (rc != SQW_OK)
        return rc;
    return sqw::launch_kmer(d_bases, d_offsets, n, max_len, kmin, kmax, d_counts, d_has_n,
                            d_status, static_cast<cudaStream_t>(stream));
}

int sqw_kmer_count(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   int32_t *counts, uint8_t *has_n)
{
    using namespace sqw;
    int rc = require_device();
    if (rc != SQW_OK)
        return rc;
    const int64_t numK = num_kmers(kmin, kmax);
    if (numK < 0)
        return set_error(SQW_ERR_INVALID_ARG, "invalid k range [%d,%d] (1 <= kmin <= kmax <= 15)",
                         kmin, kmax);
    if (n < 0 || (n > 0 && (!offsets || !counts || !has_n)))
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
    return host_kmer_count(bases, offsets, n, max_len, kmin, kmax, numK, counts, has_n);
}

}  // extern "C"
