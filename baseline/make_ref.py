"""Recipe: stage the UNMODIFIED reference hot-path tree into git-ignored baseline/_ref/ so that it travels to the GPU box.

  python baseline/make_ref.py            # copies from /root/reference (or $CTRLHAIR_REFERENCE)

The reference (XuyangGuo/CtrlHair) is a script tree without packaging (no setup.py / pyproject), so there is nothing
to `pip install`; the reference arm of bench.py (`--impl reference`) and its `reference_gpu` row import the staged
copy instead.  Only the directories the SEAN generator path imports are staged — byte for byte, no edits:

  sean_codes/   generator / architecture / normalization / sync_batchnorm / pix2pix_model (generator.py:72-109)
  util/         `import util.util` in sean_codes/models/networks/__init__.py:10
  imgs/         the 50 example faces BASELINE.json config 3 names (ui/backend.py:67-106 reads such files)

baseline/_ref/ is listed in .gitignore (it never enters history) and NOT in .gpurunignore (it ships with the snapshot,
like the built .so).  The digest file written next to the copy lets bench.py state which tree it timed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("CTRLHAIR_REFERENCE", "/root/reference")
SUBTREES = ["sean_codes", "util", "imgs"]


def tree_digest(root):
    h = hashlib.sha256()
    n = 0
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames.sort()
        for fn in sorted(filenames):
            if fn.endswith(".pyc") or fn == "DIGEST.json":
                continue
            p = os.path.join(dirpath, fn)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as f:
                h.update(f.read())
            n += 1
    return h.hexdigest(), n


def staged():
    return os.path.isdir(os.path.join(DEST, "sean_codes", "models", "networks"))


def make(force=False):
    if staged() and not force:
        return True          # e.g. on the GPU box: the staged copy travelled with the snapshot, /root/reference does not exist
    if not os.path.isdir(os.path.join(SRC, "sean_codes")):
        return False
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for sub in SUBTREES:
        shutil.copytree(os.path.join(SRC, sub), os.path.join(DEST, sub), ignore=ignore)
    digest, n = tree_digest(DEST)
    with open(os.path.join(DEST, "DIGEST.json"), "w") as f:
        json.dump({"source": SRC, "subtrees": SUBTREES, "files": n, "sha256": digest}, f)
    return True


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    print("baseline/_ref %s" % ("staged" if ok else "NOT staged: %s has no sean_codes/" % SRC))
    sys.exit(0 if ok else 1)
