/*
 * jni/stub/jni.h — a MINIMAL stand-in for the JDK's <jni.h>, written for this repository so that
 * jni/sequnwinder_b200_jni.c can be type-checked in an image that has no JDK (`make -C jni check`).
 *
 * It declares only the types, macros and JNIEnv members the shim uses, with the JNI
 * specification's names and signatures. The member ORDER of the function table is NOT the real
 * one, so objects compiled against this header must never be loaded by a JVM: build the real
 * library with the JDK's header (`make -C jni JNI_INCLUDE=$JAVA_HOME/include`).
 */
#ifndef SQW_JNI_STUB_H
#define SQW_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1
#define JNI_OK 0

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

struct sqw_stub_jobject;
typedef struct sqw_stub_jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jbyteArray;
typedef jarray jdoubleArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);
    jstring (*NewStringUTF)(JNIEnv *env, const char *utf);
    jint *(*GetIntArrayElements)(JNIEnv *env, jintArray array, jboolean *isCopy);
    jlong *(*GetLongArrayElements)(JNIEnv *env, jlongArray array, jboolean *isCopy);
    jbyte *(*GetByteArrayElements)(JNIEnv *env, jbyteArray array, jboolean *isCopy);
    jdouble *(*GetDoubleArrayElements)(JNIEnv *env, jdoubleArray array, jboolean *isCopy);
    void (*ReleaseIntArrayElements)(JNIEnv *env, jintArray array, jint *elems, jint mode);
    void (*ReleaseLongArrayElements)(JNIEnv *env, jlongArray array, jlong *elems, jint mode);
    void (*ReleaseByteArrayElements)(JNIEnv *env, jbyteArray array, jbyte *elems, jint mode);
    void (*ReleaseDoubleArrayElements)(JNIEnv *env, jdoubleArray array, jdouble *elems, jint mode);
    void *(*GetPrimitiveArrayCritical)(JNIEnv *env, jarray array, jboolean *isCopy);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *env, jarray array, void *carray, jint mode);
};

#endif /* SQW_JNI_STUB_H */
