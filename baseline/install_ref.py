#!/usr/bin/env python
"""Installs the UNMODIFIED reference's hot-path modules into the git-ignored `baseline/_ref/`.

    python baseline/install_ref.py [--src /root/reference] [--force]

The reference (NguyendaiAI-pixta/face_detection_multi_scale) is a plain script tree with no setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to build: this recipe copies
the files of the hot path byte for byte instead -- `multi_scale_face_detector.py`, `detect.py`, the `models/`,
`utils/` and `cfg/` trees it imports (python / yaml only) and four of `data/images/*.jpg` -- and writes a manifest
with their SHA-256 digests.  Nothing is edited.  `baseline/_ref/` is listed in .gitignore (reference sources never
enter this repository's history) but not in .gpurunignore, so the tree travels to the GPU box like the built .so
files: there `bench.py --impl reference`, `bench.py`'s `cpu_baseline` leg and the `-m gpu` reference tests import
it through `baseline/refshim.py`.  `__graft_entry__.build()` runs this where `/root/reference` exists.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["multi_scale_face_detector.py", "detect.py", "requirements.txt"]
TREES = [("models", (".py", ".yaml")), ("utils", (".py",)), ("cfg", (".yaml",))]
IMAGES = 4  # data/images/*.jpg, smallest first: real crowd photographs for the from-files tests


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def install(src="/root/reference", force=False):
    """Returns the destination directory, or None when there is no reference tree to install from."""
    if not os.path.isfile(os.path.join(src, "multi_scale_face_detector.py")):
        return DEST if os.path.isfile(os.path.join(DEST, "MANIFEST.json")) else None
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    copied = {}

    def put(rel):
        d = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
        copied[rel] = _sha(d)

    for rel in FILES:
        if os.path.isfile(os.path.join(src, rel)):
            put(rel)
    for tree, exts in TREES:
        for root, dirs, files in os.walk(os.path.join(src, tree)):
            dirs[:] = [x for x in dirs if x != "__pycache__"]
            for f in sorted(files):
                if f.endswith(exts):
                    put(os.path.relpath(os.path.join(root, f), src))
    img_dir = os.path.join(src, "data", "images")
    if os.path.isdir(img_dir):
        jpgs = sorted((os.path.getsize(os.path.join(img_dir, f)), f) for f in os.listdir(img_dir) if f.lower().endswith(".jpg"))
        for _, f in jpgs[:IMAGES]:
            put(os.path.join("data", "images", f))
    with open(manifest_path, "w") as f:
        json.dump({"source": src, "files": copied, "note": "byte-for-byte copies of the reference; nothing edited"}, f, indent=1)
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    d = install(a.src, a.force)
    print(d if d else "no reference tree at %s" % a.src)
    sys.exit(0 if d else 1)
