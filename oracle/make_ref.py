"""Recipe: place the reference's own loss modules under ``oracle/_ref/`` so that they travel to the GPU box.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` exists only in the build container; the
GPU box receives a snapshot of this repository.  ``oracle/_ref/`` is git-ignored (never part of the history, no
reference source is committed) but NOT gpurun-ignored, so what this recipe puts there is what ``bench.py --impl
reference`` and ``tests/test_oracle_vs_reference.py`` execute on the box: the UNMODIFIED reference files
(mono/model/{mono_fm,mono_baseline,mono_fm_joint,mono_fm_joint_inpaint,mono_autoencoder}/*.py, registry.py -- the
modules SURVEY.md section 8c lists plus what they import).  A manifest with the SHA-256 of every file is written next
to them; ``ref_loader`` refuses a tree whose manifest does not match.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("TDL_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
PACKAGES = ["mono_fm", "mono_baseline", "mono_fm_joint", "mono_fm_joint_inpaint", "mono_autoencoder"]
FILES = ["registry.py"]


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def make_ref(verbose=False):
    src_model = os.path.join(SRC_ROOT, "mono", "model")
    if not os.path.isdir(src_model):
        return None                                   # not the build container: keep whatever travelled here
    dst_model = os.path.join(DST_ROOT, "mono", "model")
    if os.path.isdir(DST_ROOT):
        shutil.rmtree(DST_ROOT)
    os.makedirs(dst_model)
    manifest = {}
    for pkg in PACKAGES:
        for name in sorted(os.listdir(os.path.join(src_model, pkg))):
            if name.endswith(".py"):
                os.makedirs(os.path.join(dst_model, pkg), exist_ok=True)
                shutil.copyfile(os.path.join(src_model, pkg, name), os.path.join(dst_model, pkg, name))
                manifest[f"mono/model/{pkg}/{name}"] = _sha(os.path.join(dst_model, pkg, name))
    for name in FILES:
        shutil.copyfile(os.path.join(src_model, name), os.path.join(dst_model, name))
        manifest[f"mono/model/{name}"] = _sha(os.path.join(dst_model, name))
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files from {SRC_ROOT}")
    return DST_ROOT


def verify(root=DST_ROOT):
    """True when every file listed in the manifest is present and unmodified."""
    mpath = os.path.join(root, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    files = json.load(open(mpath))["files"]
    return all(os.path.isfile(os.path.join(root, rel)) and _sha(os.path.join(root, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    print(make_ref(verbose=True), "verified" if verify() else "NOT verified")
