package org.seqcode.projects.sequnwinder.motifs;

/**
 * Native replacement of the scan loops of Discrim.KmerModelScanner.findHills (Discrim.java:328-380,
 * mode 1) and findSeqlets (:382-444, mode 0). The caller keeps the stable sort by score
 * (:231-249) and builds the substrings / Regions from (start, len). Java 8.
 *
 * Hills of sequence r, in the reference's list order, are entries [r*cap, r*cap + counts[r]) of
 * start / len / score; start is 0-based in the sequence (findHills adds the region start, :355).
 */
public final class HillsScanNative {
    static {
        System.loadLibrary("sequnwinder_b200_jni");
    }

    private HillsScanNative() { }

    /** @return 0, or a negative SQW_ERR_* code (-5: more than cap hills in a sequence) */
    public static int scan(byte[] bases, long[] offsets, int kmin, int kmax, double[] kmerWeights, int minM,
            int maxM, double hillsThresh, int mode, int cap, int[] startOut, int[] lenOut, double[] scoreOut,
            int[] countsOut, byte[] hasNOut) {
        return nativeHillsScan(bases, offsets, kmin, kmax, kmerWeights, minM, maxM, hillsThresh, mode, cap,
                startOut, lenOut, scoreOut, countsOut, hasNOut);
    }

    private static native int nativeHillsScan(byte[] bases, long[] offsets, int kmin, int kmax,
            double[] kmerWeights, int minM, int maxM, double hillsThresh, int mode, int cap, int[] startOut,
            int[] lenOut, double[] scoreOut, int[] countsOut, byte[] hasNOut);
}
