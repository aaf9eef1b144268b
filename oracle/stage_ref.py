#!/usr/bin/env python
"""Stages the UNMODIFIED reference package for the CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE).

    python oracle/stage_ref.py        # copies /root/reference/machupX -> oracle/_ref/machupX and /root/reference/test ->
                                      # oracle/_ref/test, byte for byte

MachUpX is pure Python: nothing compiles, so "building" the reference means putting its package where the GPU box can import
it.  /root/reference does not exist on that box; oracle/_ref/ is git-ignored (no reference source ever enters the history) but
travels with the repository snapshot like the built liblls.so does.  The copy is verified file by file (sha256) and the manifest
oracle/_ref/STAGED.json records what was staged.  oracle/ref_runner.py imports the staged package together with the stand-ins
in oracle/ref_stubs (airfoil_db / matplotlib / numpy-stl are not installable offline; SURVEY.md section 8c: 61 of the
reference's 62 non-CLI tests pass on them).  The reference's own test-suite (test/, with its JSON / csv fixtures) is staged
the same way: tests/test_zz_reference_suite.py runs it UNMODIFIED against machupx_b200 (`import machupX` resolved to this
package), on CPU with the oracle stand-in devices and on the GPU box with the CUDA path.  Called by __graft_entry__.build()
whenever /root/reference is present."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/machupX"
SRC_TESTS = "/root/reference/test"
DST = os.path.join(HERE, "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _copy_verified(src, dst_dir):
    """Byte-for-byte copy of a directory tree; returns {relative path: sha256}."""
    if os.path.isdir(dst_dir):
        shutil.rmtree(dst_dir)
    shutil.copytree(src, dst_dir, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache"))
    files = {}
    for root, dirs, names in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".pytest_cache")]
        for name in sorted(names):
            if name.endswith(".pyc"):
                continue
            a = os.path.join(root, name)
            rel = os.path.relpath(a, src)
            assert _sha(a) == _sha(os.path.join(dst_dir, rel)), "staged copy differs from the reference: " + rel
            files[rel] = _sha(a)
    return files


def stage(src=SRC, dst=DST, src_tests=SRC_TESTS):
    """Returns the manifest, or None when the reference tree is not present (GPU box: the staged copy is used as it is)."""
    if not os.path.isdir(src):
        return None
    os.makedirs(dst, exist_ok=True)
    files = _copy_verified(src, os.path.join(dst, "machupX"))
    test_files = _copy_verified(src_tests, os.path.join(dst, "test")) if os.path.isdir(src_tests) else {}
    version = "?"
    try:
        with open(os.path.join(os.path.dirname(src), "setup.py")) as fh:
            for line in fh:
                if "version" in line and "=" in line:
                    version = line.split("=")[1].strip().strip(",").strip("'\"")
                    break
    except OSError:
        pass
    manifest = {"source": src, "package": "machupX", "version": version, "files": files, "test_files": test_files,
                "note": "byte-for-byte copy of the reference package; never committed (oracle/_ref/ is git-ignored)"}
    with open(os.path.join(dst, "STAGED.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


def staged_path():
    """Directory to put on sys.path to import the staged reference, or None."""
    return DST if os.path.isfile(os.path.join(DST, "machupX", "scene.py")) else None


def staged_tests():
    """Directory holding the staged copy of the reference's test-suite (its `test/`), or None."""
    path = os.path.join(DST, "test")
    return path if os.path.isfile(os.path.join(path, "input_for_testing.json")) else None


if __name__ == "__main__":
    m = stage()
    if m is None:
        print("reference tree not present at %s; nothing staged" % SRC)
        sys.exit(0 if staged_path() else 1)
    print("staged %d package files and %d test-suite files of machupX %s into %s"
          % (len(m["files"]), len(m["test_files"]), m["version"], DST))
