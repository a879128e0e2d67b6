package org.seqcode.projects.sequnwinder.loaddata;

import java.util.List;

/**
 * Native replacement of the counting loop of MakeArff.KmerCounter.printPeakSeqKmerRange
 * (MakeArff.java:347-368) and of its second copy in utils/ScoreSites.java:105-116. Java 8.
 *
 * Use inside printPeakSeqKmerRange (header, weights and labels stay as they are,
 * MakeArff.java:336-344,369-374):
 *
 *     KmerCounterNative.Counts kc = KmerCounterNative.count(seqs, kmin, kmax);   // all rows, one call
 *     for (int r = 0; r < seqs.size(); r++) {
 *         if (kc.hasN[r] != 0) continue;                       // MakeArff.java:354-355
 *         ... append kc.counts[r * kc.numK + i] for i < numK, then weight and label ...
 *     }
 *
 * Sequences must be upper case (SeqUnwinderConfig.java:296 / the toUpperCase() at MakeArff.java:351
 * already guarantee it; the library folds case anyway). A character outside ACGTN raises
 * IllegalArgumentException, as seqcode-core's seq2int does.
 */
public final class KmerCounterNative {
    static {
        System.loadLibrary("sequnwinder_b200_jni");
    }

    public static final class Counts {
        public final int numK;
        public final int[] counts; // n x numK row-major; rows with hasN != 0 are zero
        public final byte[] hasN;

        Counts(int numK, int[] counts, byte[] hasN) {
            this.numK = numK;
            this.counts = counts;
            this.hasN = hasN;
        }
    }

    private KmerCounterNative() { }

    /** numK = sum over k of 4^k (MakeArff.java:325-328). */
    public static int numKmers(int kmin, int kmax) {
        return (int) nativeNumKmers(kmin, kmax);
    }

    public static Counts count(List<String> seqs, int kmin, int kmax) {
        int n = seqs.size();
        long total = 0;
        for (String s : seqs)
            total += s.length();
        if (total > Integer.MAX_VALUE - 8)
            throw new IllegalArgumentException("call in row chunks: more than 2^31 bases");
        byte[] bases = new byte[(int) total];
        long[] offsets = new long[n + 1];
        int pos = 0;
        for (int r = 0; r < n; r++) {
            String s = seqs.get(r);
            for (int i = 0; i < s.length(); i++)
                bases[pos++] = (byte) s.charAt(i);
            offsets[r + 1] = pos;
        }
        int numK = numKmers(kmin, kmax);
        if ((long) n * numK > Integer.MAX_VALUE - 8)
            throw new IllegalArgumentException("call in row chunks: count matrix exceeds a Java array");
        int[] counts = new int[n * numK];
        byte[] hasN = new byte[n];
        int rc = nativeKmerCount(bases, offsets, kmin, kmax, counts, hasN);
        if (rc == -4)
            throw new IllegalArgumentException("Nonstandard nucleotide character encountered");
        if (rc != 0)
            throw new RuntimeException("libsequnwinder_b200: sqw_kmer_count failed with code " + rc);
        return new Counts(numK, counts, hasN);
    }

    private static native int nativeKmerCount(byte[] bases, long[] offsets, int kmin, int kmax, int[] countsOut,
            byte[] hasNOut);

    private static native long nativeNumKmers(int kmin, int kmax);
}
