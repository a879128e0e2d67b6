/*
 * sequnwinder_b200_jni.c — the JNI shim (Java 8) between SeqUnwinder's Java host code and the C
 * ABI of libsequnwinder_b200.so (include/sequnwinder_b200.h). One native method per replaced
 * loop; no callbacks into Java, no JVM threads attached, no native pointer stored in a Java
 * object (Weka deep-copies the Classifier per cross-validation fold by serialisation,
 * framework/Classifier.java:38).
 *
 * Java side: jni/LOneCuda.java, jni/KmerCounterNative.java, jni/HillsScanNative.java.
 * Reference call sites (relative to /root/reference/src/org/seqcode/projects/sequnwinder/):
 *   framework/Classifier.java:422-449   construct optimizer, setters, execute, getX / getsmX
 *   loaddata/MakeArff.java:240-241,347-368   KmerCounter.printPeakSeqKmerRange counting loop
 *   utils/ScoreSites.java:105-116        the scorer's copy of that loop
 *   motifs/Discrim.java:328-444          findHills / findSeqlets scan loops
 *
 * Build (needs a JDK for the real <jni.h>):
 *   make -C jni JNI_INCLUDE=$JAVA_HOME/include         -> jni/libsequnwinder_b200_jni.so
 * Type check without a JDK (this image): make -C jni check   (uses jni/stub/jni.h)
 */
#include <jni.h>
#include <stddef.h>

#include "sequnwinder_b200.h"

#define SQW_JNI(cls, name) Java_org_seqcode_projects_sequnwinder_##cls##_##name

/* ------------------------------------------------------------------ trainer (LOneCuda) */

JNIEXPORT jstring JNICALL SQW_JNI(framework_LOneCuda, lastError)(JNIEnv *env, jclass klass)
{
    (void)klass;
    return (*env)->NewStringUTF(env, sqw_last_error());
}

/* set_structure + set_params from Java arrays; returns the new handle in *out */
/* device >= 0: the GPU this training runs on (CUDA's current device is per host thread, and a JVM
 * worker thread starts on device 0); ctaBudget > 0: at most that many CTAs per x-update launch, so
 * that several LOneCuda.execute() calls on different Java threads train side by side on one GPU
 * (the folds of a cross-validation, a regularisation sweep) */
static jint open_handle(JNIEnv *env, jintArray nodeIndex, jintArray isLeaf, jintArray layer,
                        jintArray parentPtr, jintArray parentIdx, jintArray childPtr,
                        jintArray childIdx, jint numLayers, jdouble ridge, jdouble pho, jint threads,
                        jint admmMaxItr, jint nodesMaxItr, jint debug, jint device, jint ctaBudget,
                        sqw_admm **out)
{
    sqw_admm *h = NULL;
    jint rc = (device >= 0) ? sqw_set_device(device) : SQW_OK;
    if (rc != SQW_OK)
        return rc;
    rc = sqw_admm_create(&h);
    if (rc != SQW_OK)
        return rc;
    const jsize numNodes = (*env)->GetArrayLength(env, nodeIndex);
    jint *ni = (*env)->GetIntArrayElements(env, nodeIndex, NULL);
    jint *il = (*env)->GetIntArrayElements(env, isLeaf, NULL);
    jint *ly = (*env)->GetIntArrayElements(env, layer, NULL);
    jint *pp = (*env)->GetIntArrayElements(env, parentPtr, NULL);
    jint *pi = (*env)->GetIntArrayElements(env, parentIdx, NULL);
    jint *cp = (*env)->GetIntArrayElements(env, childPtr, NULL);
    jint *ci = (*env)->GetIntArrayElements(env, childIdx, NULL);
    if (ni && il && ly && pp && pi && cp && ci)
        rc = sqw_admm_set_structure(h, numNodes, numLayers, ni, il, ly, pp, pi, cp, ci);
    else
        rc = SQW_ERR_INVALID_ARG;
    if (ci) (*env)->ReleaseIntArrayElements(env, childIdx, ci, JNI_ABORT);
    if (cp) (*env)->ReleaseIntArrayElements(env, childPtr, cp, JNI_ABORT);
    if (pi) (*env)->ReleaseIntArrayElements(env, parentIdx, pi, JNI_ABORT);
    if (pp) (*env)->ReleaseIntArrayElements(env, parentPtr, pp, JNI_ABORT);
    if (ly) (*env)->ReleaseIntArrayElements(env, layer, ly, JNI_ABORT);
    if (il) (*env)->ReleaseIntArrayElements(env, isLeaf, il, JNI_ABORT);
    if (ni) (*env)->ReleaseIntArrayElements(env, nodeIndex, ni, JNI_ABORT);
    if (rc == SQW_OK)
        rc = sqw_admm_set_params(h, ridge, pho, threads, admmMaxItr, nodesMaxItr, debug);
    if (rc == SQW_OK && ctaBudget > 0)
        rc = sqw_admm_set_cta_budget(h, ctaBudget);   /* before set_problem: it sizes the groups */
    if (rc != SQW_OK) {
        sqw_admm_destroy(h);
        h = NULL;
    }
    *out = h;
    return rc;
}

/* execute with x / sm_x in place (LOne mutates the caller's arrays, LOne.java:107-114) */
static jint run_and_close(JNIEnv *env, sqw_admm *h, jdoubleArray xInOut, jdoubleArray smxInOut)
{
    jdouble *x = (*env)->GetDoubleArrayElements(env, xInOut, NULL);
    jdouble *s = (*env)->GetDoubleArrayElements(env, smxInOut, NULL);
    jint rc = (x && s) ? sqw_admm_execute(h, x, s) : SQW_ERR_INVALID_ARG;
    if (s) (*env)->ReleaseDoubleArrayElements(env, smxInOut, s, 0); /* mode 0: copy back */
    if (x) (*env)->ReleaseDoubleArrayElements(env, xInOut, x, 0);
    sqw_admm_destroy(h);
    return rc;
}

/* LOneCuda.execute() with the standardised double matrix (any feature values). */
JNIEXPORT jint JNICALL SQW_JNI(framework_LOneCuda, nativeTrain)(
    JNIEnv *env, jclass klass, jintArray nodeIndex, jintArray isLeaf, jintArray layer,
    jintArray parentPtr, jintArray parentIdx, jintArray childPtr, jintArray childIdx, jint numLayers,
    jdoubleArray data, jlong nRows, jint dim, jint numClasses, jintArray cls, jdoubleArray weights,
    jdouble ridge, jdouble pho, jint threads, jint admmMaxItr, jint nodesMaxItr, jint debug,
    jint device, jint ctaBudget, jdoubleArray xInOut, jdoubleArray smxInOut)
{
    (void)klass;
    sqw_admm *h = NULL;
    jint rc = open_handle(env, nodeIndex, isLeaf, layer, parentPtr, parentIdx, childPtr, childIdx,
                          numLayers, ridge, pho, threads, admmMaxItr, nodesMaxItr, debug, device,
                          ctaBudget, &h);
    if (rc != SQW_OK)
        return rc;
    jint *y = (*env)->GetIntArrayElements(env, cls, NULL);
    jdouble *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
    /* the training matrix: critical section, no copy, no JNI calls until the release */
    jdouble *X = (y && w) ? (jdouble *)(*env)->GetPrimitiveArrayCritical(env, data, NULL) : NULL;
    rc = X ? sqw_admm_set_problem(h, nRows, dim, numClasses, X, y, w) : SQW_ERR_INVALID_ARG;
    if (X) (*env)->ReleasePrimitiveArrayCritical(env, data, X, JNI_ABORT);
    if (w) (*env)->ReleaseDoubleArrayElements(env, weights, w, JNI_ABORT);
    if (y) (*env)->ReleaseIntArrayElements(env, cls, y, JNI_ABORT);
    if (rc != SQW_OK) {
        sqw_admm_destroy(h);
        return rc;
    }
    return run_and_close(env, h, xInOut, smxInOut);
}

/* LOneCuda.execute() with the k-mer COUNT matrix (bytes) + xMean / xSD: the packed-count path. */
JNIEXPORT jint JNICALL SQW_JNI(framework_LOneCuda, nativeTrainCounts)(
    JNIEnv *env, jclass klass, jintArray nodeIndex, jintArray isLeaf, jintArray layer,
    jintArray parentPtr, jintArray parentIdx, jintArray childPtr, jintArray childIdx, jint numLayers,
    jbyteArray counts, jdoubleArray mean, jdoubleArray sd, jlong nRows, jint dim, jint numClasses,
    jintArray cls, jdoubleArray weights, jdouble ridge, jdouble pho, jint threads, jint admmMaxItr,
    jint nodesMaxItr, jint debug, jint device, jint ctaBudget, jdoubleArray xInOut,
    jdoubleArray smxInOut)
{
    (void)klass;
    sqw_admm *h = NULL;
    jint rc = open_handle(env, nodeIndex, isLeaf, layer, parentPtr, parentIdx, childPtr, childIdx,
                          numLayers, ridge, pho, threads, admmMaxItr, nodesMaxItr, debug, device,
                          ctaBudget, &h);
    if (rc != SQW_OK)
        return rc;
    jint *y = (*env)->GetIntArrayElements(env, cls, NULL);
    jdouble *w = (*env)->GetDoubleArrayElements(env, weights, NULL);
    jdouble *mu = (*env)->GetDoubleArrayElements(env, mean, NULL);
    jdouble *sg = (*env)->GetDoubleArrayElements(env, sd, NULL);
    jbyte *c = (y && w && mu && sg) ? (jbyte *)(*env)->GetPrimitiveArrayCritical(env, counts, NULL) : NULL;
    rc = c ? sqw_admm_set_problem_counts(h, nRows, dim, numClasses, (const uint8_t *)c, mu, sg, y, w)
           : SQW_ERR_INVALID_ARG;
    if (c) (*env)->ReleasePrimitiveArrayCritical(env, counts, c, JNI_ABORT);
    if (sg) (*env)->ReleaseDoubleArrayElements(env, sd, sg, JNI_ABORT);
    if (mu) (*env)->ReleaseDoubleArrayElements(env, mean, mu, JNI_ABORT);
    if (w) (*env)->ReleaseDoubleArrayElements(env, weights, w, JNI_ABORT);
    if (y) (*env)->ReleaseIntArrayElements(env, cls, y, JNI_ABORT);
    if (rc != SQW_OK) {
        sqw_admm_destroy(h);
        return rc;
    }
    return run_and_close(env, h, xInOut, smxInOut);
}

/* ------------------------------------------------------------------ k-mer counter */

/* MakeArff.KmerCounter / ScoreSites counting loop: counts is n x numK, hasN is n. */
JNIEXPORT jint JNICALL SQW_JNI(loaddata_KmerCounterNative, nativeKmerCount)(
    JNIEnv *env, jclass klass, jbyteArray bases, jlongArray offsets, jint kmin, jint kmax,
    jintArray counts, jbyteArray hasN)
{
    (void)klass;
    const jsize n = (*env)->GetArrayLength(env, offsets) - 1;
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    if (!off)
        return SQW_ERR_INVALID_ARG;
    jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, bases, NULL);
    jint *c = (jint *)(*env)->GetPrimitiveArrayCritical(env, counts, NULL);
    jbyte *hn = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, hasN, NULL);
    jint rc = SQW_ERR_INVALID_ARG;
    if (b && c && hn)
        rc = sqw_kmer_count((const char *)b, (const int64_t *)off, n, kmin, kmax, (int32_t *)c,
                            (uint8_t *)hn);
    if (hn) (*env)->ReleasePrimitiveArrayCritical(env, hasN, hn, 0);
    if (c) (*env)->ReleasePrimitiveArrayCritical(env, counts, c, 0);
    if (b) (*env)->ReleasePrimitiveArrayCritical(env, bases, b, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    return rc; /* SQW_ERR_BAD_BASE -> the caller throws IllegalArgumentException like seq2int */
}

JNIEXPORT jlong JNICALL SQW_JNI(loaddata_KmerCounterNative, nativeNumKmers)(JNIEnv *env, jclass klass,
                                                                            jint kmin, jint kmax)
{
    (void)env;
    (void)klass;
    return (jlong)sqw_num_kmers(kmin, kmax);
}

/* ------------------------------------------------------------------ hills scan */

JNIEXPORT jint JNICALL SQW_JNI(motifs_HillsScanNative, nativeHillsScan)(
    JNIEnv *env, jclass klass, jbyteArray bases, jlongArray offsets, jint kmin, jint kmax,
    jdoubleArray kmerWeights, jint minM, jint maxM, jdouble thresh, jint mode, jint cap,
    jintArray start, jintArray len, jdoubleArray score, jintArray counts, jbyteArray hasN)
{
    (void)klass;
    const jsize n = (*env)->GetArrayLength(env, offsets) - 1;
    jlong *off = (*env)->GetLongArrayElements(env, offsets, NULL);
    jdouble *w = (*env)->GetDoubleArrayElements(env, kmerWeights, NULL);
    jint *st = (*env)->GetIntArrayElements(env, start, NULL);
    jint *ln = (*env)->GetIntArrayElements(env, len, NULL);
    jdouble *sc = (*env)->GetDoubleArrayElements(env, score, NULL);
    jint *ct = (*env)->GetIntArrayElements(env, counts, NULL);
    jbyte *hn = (*env)->GetByteArrayElements(env, hasN, NULL);
    jint rc = SQW_ERR_INVALID_ARG;
    if (off && w && st && ln && sc && ct && hn) {
        jbyte *b = (jbyte *)(*env)->GetPrimitiveArrayCritical(env, bases, NULL);
        if (b) {
            rc = sqw_hills_scan((const char *)b, (const int64_t *)off, n, kmin, kmax, w, minM, maxM,
                                thresh, mode, cap, (int32_t *)st, (int32_t *)ln, sc, (int32_t *)ct,
                                (uint8_t *)hn);
            (*env)->ReleasePrimitiveArrayCritical(env, bases, b, JNI_ABORT);
        }
    }
    if (hn) (*env)->ReleaseByteArrayElements(env, hasN, hn, 0);
    if (ct) (*env)->ReleaseIntArrayElements(env, counts, ct, 0);
    if (sc) (*env)->ReleaseDoubleArrayElements(env, score, sc, 0);
    if (ln) (*env)->ReleaseIntArrayElements(env, len, ln, 0);
    if (st) (*env)->ReleaseIntArrayElements(env, start, st, 0);
    if (w) (*env)->ReleaseDoubleArrayElements(env, kmerWeights, w, JNI_ABORT);
    if (off) (*env)->ReleaseLongArrayElements(env, offsets, off, JNI_ABORT);
    return rc; /* SQW_ERR_UNSUPPORTED: more than cap hills in a sequence -> retry with a larger cap */
}
