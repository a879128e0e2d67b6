/*
 * kmer_oracle.c — CPU restatement of loaddata/MakeArff.java:323-377 (TEST INFRASTRUCTURE ONLY;
 * see oracle.h). Deliberately literal: one substring, one reverse complement and two
 * seq2int() calls per window, exactly as the reference does it.
 *
 * Third-party semantics restated (seqcode-core, un-vendored, version unpinned; build.xml:11-13):
 *   RegionFileUtilities.seq2int : A=0 C=1 G=2 T=3, first base most significant, throws on any
 *                                 other character (call sites MakeArff.java:362-363);
 *   SequenceUtils.reverseComplement : reverse, A<->T, C<->G (call site MakeArff.java:361).
 * Local corroboration of the encoding: utils/SimulateBindingSite.java:132-147.
 */
#include "oracle.h"
#include <string.h>

int64_t orc_num_kmers(int kmin, int kmax)
{
    /* MakeArff.java:325-328 */
    int64_t numK = 0;
    for (int k = kmin; k <= kmax; k++)
        numK = numK + ((int64_t)1 << (2 * k));
    return numK;
}

/* seqcode-core RegionFileUtilities.seq2int; -1 plays the role of the thrown exception */
static int64_t seq2int(const char *s, int len)
{
    int64_t v = 0;
    for (int i = 0; i < len; i++) {
        int64_t cur;
        switch (s[i]) {
        case 'A': cur = 0; break;
        case 'C': cur = 1; break;
        case 'G': cur = 2; break;
        case 'T': cur = 3; break;
        default: return -1;
        }
        v = v << 2;
        v += cur;
    }
    return v;
}

/* seqcode-core SequenceUtils.reverseComplement */
static void reverse_complement(const char *s, int len, char *out)
{
    for (int i = 0; i < len; i++) {
        char c = s[len - 1 - i];
        switch (c) {
        case 'A': c = 'T'; break;
        case 'C': c = 'G'; break;
        case 'G': c = 'C'; break;
        case 'T': c = 'A'; break;
        default: break;
        }
        out[i] = c;
    }
}

int orc_kmer_count(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   int32_t *counts, uint8_t *has_n)
{
    int64_t numK = orc_num_kmers(kmin, kmax);
    char rev[64];
    if (kmax > 31 || kmin < 1 || kmin > kmax)
        return -2;
    for (int64_t r = 0; r < n; r++) {
        int32_t *kmerCounts = counts + r * numK;
        const char *seq = bases + offsets[r];
        int64_t len = offsets[r + 1] - offsets[r];
        /* MakeArff.java:348-349 */
        for (int64_t i = 0; i < numK; i++)
            kmerCounts[i] = 0;
        /* MakeArff.java:354-355: seq.contains("N") -> row skipped */
        has_n[r] = 0;
        if (memchr(seq, 'N', (size_t)len) != NULL) {
            has_n[r] = 1;
            continue;
        }
        int64_t ind = 0;
        for (int k = kmin; k <= kmax; k++) {                   /* :358 */
            for (int64_t i = 0; i < (len - k + 1); i++) {       /* :359 */
                const char *currK = seq + i;                    /* :360 */
                reverse_complement(currK, k, rev);              /* :361 */
                int64_t currKInt = seq2int(currK, k);           /* :362 */
                int64_t revCurrKInt = seq2int(rev, k);          /* :363 */
                if (currKInt < 0 || revCurrKInt < 0)
                    return -1;
                int64_t kmer = currKInt < revCurrKInt ? currKInt : revCurrKInt; /* :364 */
                kmerCounts[ind + kmer]++;                       /* :365 */
            }
            ind = ind + ((int64_t)1 << (2 * k));                /* :367 */
        }
    }
    return 0;
}
