"""Builds a variant of the library for in-box A/B runs: current csrc with some files replaced.

  python tools/build_variant.py NAME [path_in_csrc=GIT_REV ...] [path_in_csrc=@/abs/file ...]
-> variants/libNAME.so   (variants/ is git-ignored; it travels to the GPU box with gpurun)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import build as b  # noqa: E402


def main():
    name, subs = sys.argv[1], sys.argv[2:]
    src = "/tmp/variant_%s/ctrlhair_b200/csrc" % name
    shutil.rmtree("/tmp/variant_%s" % name, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "ctrlhair_b200", "csrc"), src)
    shutil.copytree(os.path.join(ROOT, "include"), "/tmp/variant_%s/include" % name)
    for s in subs:
        path, rev = s.split("=", 1)
        if rev.startswith("@"):
            shutil.copy(rev[1:], os.path.join(src, path))
        else:
            data = subprocess.check_output(["git", "show", "%s:ctrlhair_b200/csrc/%s" % (rev, path)], cwd=ROOT)
            open(os.path.join(src, path), "wb").write(data)
    objs = []
    for f in b.SOURCES:
        o = os.path.join(src, f.replace(".cu", ".o"))
        extra = os.environ.get("VARIANT_FLAGS", "").split()   # e.g. VARIANT_FLAGS="-DCHB_EXP_NOSTG"
        subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + extra + ["-c", os.path.join(src, f), "-o", o])
        objs.append(o)
    os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
    out = os.path.join(ROOT, "variants", "lib%s.so" % name)
    subprocess.check_call([b._nvcc(), "-shared", "-cudart", "static", "-o", out] + objs + ["-lpthread", "-ldl", "-lrt"])
    print(out)


if __name__ == "__main__":
    main()
