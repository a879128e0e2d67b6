package org.seqcode.projects.sequnwinder.framework;

import java.util.ArrayList;
import java.util.List;

import org.seqcode.projects.sequnwinder.framework.Classifier.ClassRelationStructure;
import org.seqcode.projects.sequnwinder.framework.Classifier.ClassRelationStructure.Node;

/**
 * B200 drop-in for {@link LOne}: the same {@link Optimizer} plug-in, the same setters in the same
 * order (Classifier.java:422-449), trained by libsequnwinder_b200.so through
 * jni/sequnwinder_b200_jni.c. Java 8.
 *
 * Selection (two added lines at Classifier.java:424-425):
 *     else if (m_OptimizationType.equals("CUDA")) opt = new LOneCuda(x, sm_x, m_Data);
 * and, for the packed-count path, before the in-place standardisation at Classifier.java:373-381:
 *     byte[] rawCounts = LOneCuda.packCounts(m_Data);      // null if a value is not in 0..255
 *     ... opt = new LOneCuda(x, sm_x, m_Data); ((LOneCuda) opt).setCounts(rawCounts, xMean, xSD);
 *
 * The native handle lives only inside the native call, so the Serializable deep copies Weka
 * makes of the Classifier per cross-validation fold never see a native pointer.
 */
public class LOneCuda extends Optimizer {
    static {
        System.loadLibrary("sequnwinder_b200_jni"); // links libsequnwinder_b200.so
    }

    // same fields and defaults as LOne.java:24-51
    public int ADMM_maxItr = 500;
    public int ADMM_numThreads = 5;
    public int BGFS_maxIts = -1;
    public double ADMM_pho = 1.7;
    protected int numPredictors;
    protected int numClasses;
    protected int numNodes;
    protected int NODES_maxItr = 10;
    protected double regularization;
    protected boolean sm_Debug;
    protected ClassRelationStructure classStructure;
    public double[] x;
    public double[] sm_x;
    public double[][] data;
    protected double[] weights;
    protected int[] cls;
    // packed-count path (optional): raw k-mer counts + the statistics Classifier.java:362-371 made
    protected byte[] counts;
    protected double[] xMean;
    protected double[] xSD;
    // placement (no reference counterpart): the GPU to train on (-1 = the thread's current device)
    // and the CTA budget of this training (0 = every SM)
    protected int device = -1;
    protected int ctaBudget = 0;

    /** LOne.java:110-114 */
    public LOneCuda(double[] xinit, double[] sm_xinit, double[][] d) {
        x = xinit;
        sm_x = sm_xinit;
        data = d;
    }

    // LOne.java:92-104
    public void setBGFSmaxItrs(int m) { BGFS_maxIts = m; }
    public void setADMMmaxItrs(int m) { ADMM_maxItr = m; }
    public void setSeqUnwinderMaxIts(int m) { NODES_maxItr = m; }
    public void setInstanceWeights(double[] w) { weights = w; }
    public void setClsMembership(int[] c) { cls = c; }
    public void setNumPredictors(int p) { numPredictors = p; }
    public void setNumClasses(int c) { numClasses = c; }
    public void setClassStructure(ClassRelationStructure rel) { classStructure = rel; setNumNodes(rel.allNodes.size()); }
    public void setNumNodes(int n) { numNodes = n; }
    public void setRidge(double r) { regularization = r; }
    public void setDebugMode(boolean debug) { sm_Debug = debug; }
    public void setPho(double ph) { ADMM_pho = ph; }
    public void set_numThreads(int nt) { ADMM_numThreads = nt; }
    /** z, zold and u are zeroed on the device by execute() (LOne.java:80-85). */
    public void initUandZ() { }
    public double[] getX() { return x; }
    public double[] getsmX() { return sm_x; }
    /** LOne.java:259-263 */
    public void clearOptimizer() { data = null; sm_x = null; x = null; counts = null; }

    /**
     * Packed-count path: rawCounts is nC x nR row-major (the feature values of m_Data[i][1..nR]
     * BEFORE Classifier.java:373-381 standardised them), mean / sd are xMean / xSD (nR+1 entries,
     * entry 0 = 0 / 1). The trainer then evaluates with X[i][j] = (count - mean[j]) / sd[j]
     * (the raw count where sd[j] == 0, Classifier.java:376) without materialising the matrix.
     */
    public void setCounts(byte[] rawCounts, double[] mean, double[] sd) {
        counts = rawCounts;
        xMean = mean;
        xSD = sd;
    }

    /**
     * Several trainings side by side on one GPU: give each LOneCuda a share of the 148 SMs
     * (148 / slots) and call execute() from `slots` Java threads - the folds of --x (the model on all
     * sites + one per fold) or the values of a regularisation sweep. Measured on B200 at cfg-2, 3
     * folds: 12.5 s one training after the other, 8.7 s with four slots. Each training returns what
     * a single run with the same budget returns.
     */
    public void setCtaBudget(int maxCtas) { ctaBudget = maxCtas; }

    /** The GPU this training runs on: a JVM worker thread starts on CUDA device 0. */
    public void setDevice(int dev) { device = dev; }

    /** nC x (nR+1) un-standardised m_Data -> nC x nR bytes, or null if a value is not 0..255. */
    public static byte[] packCounts(double[][] raw) {
        int nC = raw.length, nR = raw[0].length - 1;
        byte[] out = new byte[nC * nR];
        for (int i = 0; i < nC; i++)
            for (int j = 1; j <= nR; j++) {
                double v = raw[i][j];
                int c = (int) v;
                if (c != v || c < 0 || c > 255)
                    return null;
                out[i * nR + (j - 1)] = (byte) c;
            }
        return out;
    }

    public void execute() throws Exception {
        // class structure -> CSR arrays, one entry per design-file line, layer by layer
        // (Classifier.java:648-683 appends to layers.get(l) in file order)
        List<Node> order = new ArrayList<Node>();
        for (List<Node> layer : classStructure.layers)
            order.addAll(layer);
        int n = order.size();
        int[] nodeIndex = new int[n], isLeaf = new int[n], layerOf = new int[n];
        int[] parentPtr = new int[n + 1], childPtr = new int[n + 1];
        List<Integer> p = new ArrayList<Integer>(), c = new ArrayList<Integer>();
        int e = 0;
        for (int l = 0; l < classStructure.numLayers; l++)
            for (Node nd : classStructure.layers.get(l)) {
                nodeIndex[e] = nd.nodeIndex;
                isLeaf[e] = nd.isLeaf ? 1 : 0;
                layerOf[e] = l;
                p.addAll(nd.parents);
                c.addAll(nd.children);
                parentPtr[e + 1] = p.size();
                childPtr[e + 1] = c.size();
                e++;
            }
        int nC = data.length, dim = data[0].length;
        int rc;
        if (counts != null) {
            rc = nativeTrainCounts(nodeIndex, isLeaf, layerOf, parentPtr, toArray(p), childPtr, toArray(c),
                    classStructure.numLayers, counts, xMean, xSD, nC, dim, numClasses, cls, weights,
                    regularization, ADMM_pho, ADMM_numThreads, ADMM_maxItr, NODES_maxItr, sm_Debug ? 1 : 0,
                    device, ctaBudget, x, sm_x);
        } else {
            // flatten double[][] -> double[] row-major (Classifier.java:326: m_Data is nC x (nR+1))
            double[] flat = new double[nC * dim];
            for (int i = 0; i < nC; i++)
                System.arraycopy(data[i], 0, flat, i * dim, dim);
            rc = nativeTrain(nodeIndex, isLeaf, layerOf, parentPtr, toArray(p), childPtr, toArray(c),
                    classStructure.numLayers, flat, nC, dim, numClasses, cls, weights, regularization,
                    ADMM_pho, ADMM_numThreads, ADMM_maxItr, NODES_maxItr, sm_Debug ? 1 : 0, device, ctaBudget,
                    x, sm_x);
        }
        if (rc != 0)
            throw new Exception("libsequnwinder_b200: " + lastError());
    }

    private static int[] toArray(List<Integer> l) {
        int[] a = new int[l.size()];
        for (int i = 0; i < a.length; i++)
            a[i] = l.get(i);
        return a;
    }

    private static native int nativeTrain(int[] nodeIndex, int[] isLeaf, int[] layer, int[] parentPtr,
            int[] parentIdx, int[] childPtr, int[] childIdx, int numLayers, double[] data, long nRows, int dim,
            int numClasses, int[] cls, double[] weights, double ridge, double pho, int threads, int admmMaxItr,
            int nodesMaxItr, int debug, int device, int ctaBudget, double[] xInOut, double[] smxInOut);

    private static native int nativeTrainCounts(int[] nodeIndex, int[] isLeaf, int[] layer, int[] parentPtr,
            int[] parentIdx, int[] childPtr, int[] childIdx, int numLayers, byte[] counts, double[] mean,
            double[] sd, long nRows, int dim, int numClasses, int[] cls, double[] weights, double ridge,
            double pho, int threads, int admmMaxItr, int nodesMaxItr, int debug, int device, int ctaBudget,
            double[] xInOut, double[] smxInOut);

    private static native String lastError();
}
