/*
 * hills_oracle.c — CPU restatement of motifs/Discrim.java:328-444 (findHills / findSeqlets):
 * the sliding-window scan of a trained k-mer model over the training sequences that picks the
 * "hills" handed to the clustering / MEME steps. TEST INFRASTRUCTURE ONLY; see
 * oracle.h. PARITY UNPINNED: the reference holds no test or fixture for this path and cannot run
 * here; the restatement is line-traceable to the citations below.
 *
 * Deliberately literal: for every window length l = minM..maxM and every offset i the score is
 * accumulated from 0.0 in the reference's order (k ascending, then window offset j ascending,
 * Discrim.java:394-405), one substring / reverse complement / two seq2int per k-mer, and the
 * per-sequence "mountains" list is maintained with the reference's iterator logic
 * (Discrim.java:410-427), including its quirks:
 *   - hills removed before an overlapping higher-scoring hill is met stay removed even though the
 *     candidate is then not added;
 *   - findSeqlets removes on `currScore < score` (ties keep both hills, :417), findHills on
 *     `currScore <= score` (:361);
 *   - the threshold test `score > hillsThresh` comes after the list walk (:371, :425).
 * Third-party semantics restated (seqcode-core, un-vendored, version unpinned): seq2int /
 * reverseComplement as in kmer_oracle.c; Region.overlaps on one chromosome with inclusive
 * coordinates: a.start <= b.end && a.end >= b.start.
 * SeqUnwinderConfig.getKmerBaseInd (framework/SeqUnwinderConfig.java:198-204): columns of the
 * k-mers shorter than k come first.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static int64_t h_seq2int(const char *s, int len)
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

static void h_reverse_complement(const char *s, int len, char *out)
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

typedef struct {
    int32_t start, end; /* inclusive, Region(genome, chrom, i, i+l-1) */
    double score;
} hill_t;

int orc_hills_scan(const char *bases, const int64_t *offsets, int64_t n, int kmin, int kmax,
                   const double *kmer_weights, int min_m, int max_m, double thresh, int mode,
                   int32_t cap, int32_t *hill_start, int32_t *hill_len, double *hill_score,
                   int32_t *counts, uint8_t *has_n)
{
    char rev[64];
    if (kmax > 31 || kmin < 1 || kmin > kmax || min_m < 1 || max_m < min_m || cap < 0)
        return -2;
    for (int64_t r = 0; r < n; r++) {
        const char *seq = bases + offsets[r];
        const int64_t len = offsets[r + 1] - offsets[r];
        counts[r] = 0;
        has_n[r] = 0;
        if (memchr(seq, 'N', (size_t)len) != NULL) { /* Discrim.java:337-338, :389-390 */
            has_n[r] = 1;
            continue;
        }
        size_t mcap = 64, msize = 0;
        hill_t *mountains = (hill_t *)malloc(mcap * sizeof(hill_t));
        if (!mountains)
            return -3;
        for (int l = min_m; l <= max_m; l++) {                    /* :392 */
            for (int64_t i = 0; i < (len - l + 1); i++) {         /* :393 */
                const char *motif = seq + i;                      /* :394 */
                double score = 0.0;
                for (int k = kmin; k <= kmax; k++) {              /* :396 */
                    /* getKmerBaseInd(currk), SeqUnwinderConfig.java:198-204 */
                    int64_t baseInd = 0;
                    for (int kk = kmin; kk < k; kk++)
                        baseInd += (int64_t)1 << (2 * kk);
                    for (int j = 0; j < l - k + 1; j++) {         /* :397 */
                        const char *currk = motif + j;
                        h_reverse_complement(currk, k, rev);
                        int64_t currKInt = h_seq2int(currk, k);
                        int64_t revCurrKInt = h_seq2int(rev, k);
                        if (currKInt < 0 || revCurrKInt < 0) {
                            free(mountains);
                            return -1;
                        }
                        int64_t kmer = currKInt < revCurrKInt ? currKInt : revCurrKInt;
                        score = score + kmer_weights[baseInd + kmer]; /* :404 */
                    }
                }
                const int32_t hs = (int32_t)i, he = (int32_t)(i + l - 1); /* :409 */
                /* :411-424: walk the list in insertion order while `add` holds */
                int add = 1;
                size_t idx = 0;
                while (idx < msize && add) {
                    const hill_t cur = mountains[idx];
                    const int ov = cur.start <= he && cur.end >= hs;
                    const int lower = mode ? (cur.score <= score) : (cur.score < score);
                    if (ov && lower) {
                        memmove(mountains + idx, mountains + idx + 1, (msize - idx - 1) * sizeof(hill_t));
                        msize--;          /* it.remove(); the next element moves into idx */
                        add = 1;
                    } else if (ov && cur.score > score) {
                        add = 0;
                    } else {
                        idx++;
                    }
                }
                if (add && score > thresh) {                      /* :425 */
                    if (msize == mcap) {
                        mcap *= 2;
                        hill_t *m2 = (hill_t *)realloc(mountains, mcap * sizeof(hill_t));
                        if (!m2) {
                            free(mountains);
                            return -3;
                        }
                        mountains = m2;
                    }
                    mountains[msize].start = hs;
                    mountains[msize].end = he;
                    mountains[msize].score = score;
                    msize++;
                }
            }
        }
        if (msize > (size_t)cap) {
            free(mountains);
            return -4; /* caller's per-sequence capacity too small */
        }
        for (size_t m = 0; m < msize; m++) {                      /* :432-438 / ret.addAll */
            hill_start[r * cap + (int64_t)m] = mountains[m].start;
            hill_len[r * cap + (int64_t)m] = mountains[m].end - mountains[m].start + 1;
            hill_score[r * cap + (int64_t)m] = mountains[m].score;
        }
        counts[r] = (int32_t)msize;
        free(mountains);
    }
    return 0;
}
