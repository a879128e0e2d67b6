// kmer_pack.cuh — sequence packing shared by the k-mer count (path 1) and hills scan kernels:
// ASCII bases -> 2-bit codes (A0 C1 G2 T3, first base most significant, the seq2int encoding of
// loaddata/MakeArff.java:362-363) packed MSB-first into 64-bit words, and the reverse-complement
// trick that turns a window into its canonical (min of forward / reverse-complement) code.
#pragma once
#include <cstdint>

namespace sqw {

enum : int { kFlagN = 1, kFlagBad = 2 };

__device__ __forceinline__ unsigned long long rev2_64(unsigned long long v)
{
    // reverse the order of the 32 two-bit groups of v
    v = __brevll(v);
    return ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
}

// Pack one sequence into MSB-first 2-bit words; returns flags seen by this thread's warp loop.
__device__ __forceinline__ void pack_sequence(const uint8_t *__restrict__ seq, int64_t len,
                                              unsigned long long *__restrict__ words, int t,
                                              int tpr, int *flag_slot)
{
    const int lane = t & 31;
    const int wg = t >> 5;
    const int nwarps = tpr >> 5;
    const int64_t nchunks = (len + 31) >> 5;
    int flags = 0;
    for (int64_t c = wg; c < nchunks; c += nwarps) {
        const int64_t pos = (c << 5) + lane;
        unsigned ch = (pos < len) ? (unsigned)seq[pos] : (unsigned)'A';
        const unsigned up = ch & 0xDFu;
        unsigned code = (up >> 1) & 3u;  // A0 C1 G3 T2
        code ^= (code >> 1);             // A0 C1 G2 T3
        const bool ok = (up == 'A') | (up == 'C') | (up == 'G') | (up == 'T');
        if (!ok) {
            flags |= (up == 'N') ? kFlagN : kFlagBad;
            code = 0;
        }
        const unsigned sh = 30u - 2u * (unsigned)(lane & 15);
        const unsigned hi = __reduce_or_sync(0xffffffffu, lane < 16 ? code << sh : 0u);
        const unsigned lo = __reduce_or_sync(0xffffffffu, lane >= 16 ? code << sh : 0u);
        if (lane == 0)
            words[c] = ((unsigned long long)hi << 32) | (unsigned long long)lo;
    }
    if (t == 0)
        words[nchunks] = 0ull;  // pad word read by the funnel shift of the last windows
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags)
        atomicOr(flag_slot, flags);
}

// Canonical code of the k-mer starting at base i (k <= 32): min(seq2int(w), seq2int(revcomp(w))).
// words must hold one pad word after the last base.
__device__ __forceinline__ unsigned long long canonical_code(const unsigned long long *__restrict__ words,
                                                             int64_t i, int k)
{
    const int64_t w = i >> 5;
    const int s = (int)(i & 31) << 1;
    const unsigned long long hi = words[w], lo = words[w + 1];
    const unsigned long long comb = s ? ((hi << s) | (lo >> (64 - s))) : hi;
    const unsigned long long rc_all = rev2_64(~comb);
    const unsigned long long fwd = comb >> (64 - 2 * k);
    const unsigned long long rc = (k == 32) ? rc_all : (rc_all & ((1ull << (2 * k)) - 1ull));
    return fwd < rc ? fwd : rc;
}

}  // namespace sqw
