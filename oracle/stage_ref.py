#!/usr/bin/env python
"""Stages the UNMODIFIED reference package for the CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE).

    python oracle/stage_ref.py        # copies /root/reference/machupX -> oracle/_ref/machupX, byte for byte

MachUpX is pure Python: nothing compiles, so "building" the reference means putting its package where the GPU box can import
it.  /root/reference does not exist on that box; oracle/_ref/ is git-ignored (no reference source ever enters the history) but
travels with the repository snapshot like the built liblls.so does.  The copy is verified file by file (sha256) and the manifest
oracle/_ref/STAGED.json records what was staged.  oracle/ref_runner.py imports the staged package together with the stand-ins
in oracle/ref_stubs (airfoil_db / matplotlib / numpy-stl are not installable offline; SURVEY.md section 8c: 61 of the
reference's 62 non-CLI tests pass on them).  Called by __graft_entry__.build() whenever /root/reference is present."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/machupX"
DST = os.path.join(HERE, "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage(src=SRC, dst=DST):
    """Returns the manifest, or None when the reference tree is not present (GPU box: the staged copy is used as it is)."""
    if not os.path.isdir(src):
        return None
    pkg = os.path.join(dst, "machupX")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(dst, exist_ok=True)
    shutil.copytree(src, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    files = {}
    for root, _, names in os.walk(src):
        for name in sorted(names):
            if name.endswith(".pyc"):
                continue
            a = os.path.join(root, name)
            rel = os.path.relpath(a, src)
            b = os.path.join(pkg, rel)
            assert _sha(a) == _sha(b), "staged copy differs from the reference: " + rel
            files[rel] = _sha(a)
    version = "?"
    try:
        with open(os.path.join(os.path.dirname(src), "setup.py")) as fh:
            for line in fh:
                if "version" in line and "=" in line:
                    version = line.split("=")[1].strip().strip(",").strip("'\"")
                    break
    except OSError:
        pass
    manifest = {"source": src, "package": "machupX", "version": version, "files": files,
                "note": "byte-for-byte copy of the reference package; never committed (oracle/_ref/ is git-ignored)"}
    with open(os.path.join(dst, "STAGED.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


def staged_path():
    """Directory to put on sys.path to import the staged reference, or None."""
    return DST if os.path.isfile(os.path.join(DST, "machupX", "scene.py")) else None


if __name__ == "__main__":
    m = stage()
    if m is None:
        print("reference tree not present at %s; nothing staged" % SRC)
        sys.exit(0 if staged_path() else 1)
    print("staged %d files of machupX %s into %s" % (len(m["files"]), m["version"], DST))
