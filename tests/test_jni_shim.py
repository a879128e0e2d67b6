"""The JNI shim as code: jni/sequnwinder_b200_jni.c must compile (against the stub
<jni.h>: there is no JDK in this image), export exactly the Java_* symbols the `native` methods of
jni/*.java resolve to, with matching argument counts and types, and LOneCuda.java must implement
the reference's Optimizer plug-in (framework/Optimizer.java:5-27) method for method."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI = os.path.join(ROOT, "jni")
REF = "/root/reference/src/org/seqcode/projects/sequnwinder"

JAVA_TO_JNI = {"int": "jint", "long": "jlong", "double": "jdouble", "byte": "jbyte",
               "int[]": "jintArray", "long[]": "jlongArray", "double[]": "jdoubleArray",
               "byte[]": "jbyteArray", "String": "jstring"}


def mangle(s: str) -> str:
    return s.replace("_", "_1").replace(".", "_").replace("$", "_00024")


def java_natives():
    """{jni symbol: (return type, [arg types])} parsed from the `native` declarations."""
    out = {}
    for fn in sorted(os.listdir(JNI)):
        if not fn.endswith(".java"):
            continue
        src = open(os.path.join(JNI, fn)).read()
        pkg = re.search(r"^package\s+([\w.]+);", src, re.M).group(1)
        cls = fn[:-5]
        for m in re.finditer(r"native\s+([\w\[\]]+)\s+(\w+)\s*\(([^)]*)\)\s*;", src, re.S):
            ret, name, args = m.group(1), m.group(2), m.group(3)
            types = [" ".join(a.split()[:-1]) for a in args.split(",") if a.strip()]
            out[f"Java_{mangle(pkg)}_{mangle(cls)}_{mangle(name)}"] = (ret, types)
    return out


def c_definitions():
    """{symbol: (return type, [arg types])} of the JNIEXPORT functions in the shim."""
    src = open(os.path.join(JNI, "sequnwinder_b200_jni.c")).read()
    out = {}
    for m in re.finditer(r"JNIEXPORT\s+(\w+)\s+JNICALL\s+SQW_JNI\((\w+),\s*(\w+)\)\s*\(([^)]*)\)", src, re.S):
        ret, cls, name, args = m.groups()
        types = [" ".join(a.replace("*", " * ").split()[:-1]).replace(" *", "*") for a in args.split(",")]
        assert types[0] == "JNIEnv*" and types[1] == "jclass", (name, types[:2])
        out[f"Java_org_seqcode_projects_sequnwinder_{cls}_{name}"] = (ret, types[2:])
    return out


def test_shim_compiles_and_exports_what_java_declares():
    r = subprocess.run(["make", "-C", JNI, "check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exported = set(open(os.path.join(JNI, "_check", "symbols.txt")).read().split())
    natives = java_natives()
    assert exported == set(natives), (exported ^ set(natives))
    cdefs = c_definitions()
    assert set(cdefs) == exported
    for sym, (ret, args) in natives.items():
        cret, cargs = cdefs[sym]
        assert JAVA_TO_JNI[ret] == cret, sym
        assert [JAVA_TO_JNI[a] for a in args] == cargs, sym


def test_shim_only_calls_declared_abi():
    """Every sqw_* the shim calls is declared in include/sequnwinder_b200.h."""
    src = open(os.path.join(JNI, "sequnwinder_b200_jni.c")).read()
    hdr = open(os.path.join(ROOT, "include", "sequnwinder_b200.h")).read()
    called = set(re.findall(r"\b(sqw_\w+)\s*\(", src))
    declared = set(re.findall(r"\b(sqw_\w+)\s*\(", hdr))
    assert called and called <= declared, called - declared


def test_lonecuda_implements_the_reference_optimizer():
    """Same abstract methods as framework/Optimizer.java:5-27 and the LOne-specific setters that
    Classifier.java:428-441 calls; the class sits in the reference's package."""
    src = open(os.path.join(JNI, "LOneCuda.java")).read()
    assert "package org.seqcode.projects.sequnwinder.framework;" in src
    assert re.search(r"class\s+LOneCuda\s+extends\s+Optimizer", src)
    want = ["setBGFSmaxItrs", "setSeqUnwinderMaxIts", "setInstanceWeights", "setClsMembership",
            "setNumPredictors", "setNumClasses", "setClassStructure", "setNumNodes", "setRidge",
            "setDebugMode", "execute", "getX", "getsmX", "clearOptimizer",
            "setADMMmaxItrs", "setPho", "set_numThreads", "initUandZ"]       # LOne.java:92-104
    opt_java = os.path.join(REF, "framework", "Optimizer.java")
    if os.path.exists(opt_java):    # the reference tree exists in the build container only
        ref = open(opt_java).read()
        abstract = re.findall(r"public\s+abstract\s+[\w\[\]]+\s+(\w+)\s*\(", ref)
        assert set(abstract) <= set(want)
        cls_java = open(os.path.join(REF, "framework", "Classifier.java")).read()
        for name in ("setADMMmaxItrs", "setPho", "set_numThreads", "initUandZ"):
            assert name in cls_java
    for name in want:
        assert re.search(r"public\s+[\w\[\]]+\s+%s\s*\(" % name, src), name
    assert "transient long" not in src and "long handle" not in src   # no native pointer in a field
