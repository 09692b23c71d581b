"""Stage the reference's own Python implementation of the shift-layer path under ``oracle/_ref/``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (Image-Processing-Systems-Laboratory/DeepInPainting) is
pure Python, so "building" it is staging its ``models/`` and ``util/`` modules where the GPU box can import them:
``/root/reference`` exists only in the build container, ``oracle/_ref/`` is git-ignored (never part of the history)
but travels with the working tree.  Nothing is modified: files are byte-identical copies, their SHA-256 is recorded
in ``oracle/_ref/MANIFEST.json``.

    python oracle/build_ref.py            # needs /root/reference (or $IPSR_REFERENCE)

Users: ``bench.py --impl reference`` (times the reference's real CPU path, ``cpu_baseline.kind = "reference"``),
``tests/test_gpu_refnet.py`` (the reference's own ``models/networks.py`` with our three modules switched in) and
``oracle/make_golden.py``.  The product never imports anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SUBDIRS = ("models", "util")


def reference_root():
    return os.environ.get("IPSR_REFERENCE", "/root/reference")


def build(verbose=True):
    """Returns the staged directory, or None when the reference tree is not present (GPU box: prebuilt files)."""
    src = reference_root()
    if not os.path.isdir(os.path.join(src, "models")):
        if verbose:
            print("oracle/build_ref: %s not present; keeping %s as is" % (src, DEST))
        return DEST if os.path.isdir(os.path.join(DEST, "models")) else None
    manifest = {}
    for sub in SUBDIRS:
        os.makedirs(os.path.join(DEST, sub), exist_ok=True)
        for name in sorted(os.listdir(os.path.join(src, sub))):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(src, sub, name), os.path.join(DEST, sub, name)
            shutil.copyfile(s, d)
            with open(d, "rb") as fh:
                manifest["%s/%s" % (sub, name)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "Image-Processing-Systems-Laboratory/DeepInPainting (unmodified)", "files": manifest}, fh, indent=1)
    if verbose:
        print("oracle/build_ref: staged %d files under %s" % (len(manifest), DEST))
    return DEST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
