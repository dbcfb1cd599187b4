"""Builds the in-tree CUDA library (sm_100a only) with nvcc.  No GPU is needed to build."""
import json
import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libctrlhair_b200.so"
SOURCES = ["conv_igemm.cu", "aux_kernels.cu", "generator.cu", "mlp.cu", "zencoder.cu", "shape.cu", "ct_train.cu", "blend.cu", "bisenet.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-cudart", "static",
]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


LAST_BUILD = {}


def build_library(force=False, verbose=False):
    """Compiles stale objects and relinks.  What was compiled and what was reused (objects ship to the GPU box with the
    snapshot, so a build there normally reuses everything) is returned in LAST_BUILD and written to lib/BUILD_INFO.json."""
    os.makedirs(LIBDIR, exist_ok=True)
    compiled, reused, t_start = [], [], time.time()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "ctrlhair_b200.h"))
    objs = []
    nvcc = _nvcc()
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
            compiled.append(src)
        else:
            reused.append(src)
        objs.append(o)
    out = lib_path()
    if force or _stale(out, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + \
              ["-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        linked = True
    else:
        linked = False
    LAST_BUILD.clear()
    LAST_BUILD.update({"compiled": compiled, "reused": reused, "linked": linked, "seconds": round(time.time() - t_start, 1),
                       "flags": NVCC_FLAGS, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())})
    if compiled or linked or not os.path.exists(os.path.join(LIBDIR, "BUILD_INFO.json")):
        with open(os.path.join(LIBDIR, "BUILD_INFO.json"), "w") as f:
            json.dump(LAST_BUILD, f)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
