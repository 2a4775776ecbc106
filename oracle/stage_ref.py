"""Stages the UNMODIFIED reference into git-ignored oracle/_ref/ so that it travels to the GPU box.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python: there is nothing to compile, the
"build recipe" of oracle/_ref is a byte-for-byte copy of every .py file under General/ and Applications/
of the checkout at /root/reference (572 KB; notebooks, data files and LFS pointers are not copied).  The
copy is listed in .gitignore (it never enters the history) but NOT in .gpurunignore, like a built .so.

    python oracle/stage_ref.py            # run by __graft_entry__.build() where /root/reference exists

oracle/_ref/MANIFEST.json records the SHA-256 of every staged file; `verify()` re-hashes them, so a test on
the GPU box can state that what it ran against is the reference as published (tests/test_reference_cuda.py).
oracle/ref_shim.py imports the reference from /root/reference when it exists and from oracle/_ref otherwise.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("RETINA_REFERENCE_ROOT", "/root/reference")
PACKAGES = ("General", "Applications")


def _sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _python_files(root):
    out = []
    for pkg in PACKAGES:
        for d, _, files in os.walk(os.path.join(root, pkg)):
            for f in sorted(files):
                if f.endswith(".py"):
                    out.append(os.path.relpath(os.path.join(d, f), root))
    return sorted(out)


def stage(source=SOURCE, dest=DEST):
    """Copies the reference's Python sources to `dest` (replacing a previous copy) and writes the manifest.
    Returns the manifest dict, or None when the checkout is absent (then an existing copy is left alone)."""
    if not os.path.isdir(os.path.join(source, "Applications")):
        return None
    files = _python_files(source)
    tmp = dest + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    manifest = {"source": source, "files": {}}
    for rel in files:
        dst = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source, rel), dst)
        manifest["files"][rel] = _sha256(dst)
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(dest, ignore_errors=True)
    os.rename(tmp, dest)
    return manifest


def staged(dest=DEST):
    return os.path.isfile(os.path.join(dest, "MANIFEST.json")) and os.path.isdir(os.path.join(dest, "Applications"))


def verify(dest=DEST, against=None):
    """Re-hashes the staged files against the manifest (and, when `against` names a checkout that exists, against
    the files there).  Returns the number of files; raises on any difference."""
    with open(os.path.join(dest, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, digest in manifest["files"].items():
        got = _sha256(os.path.join(dest, rel))
        if got != digest:
            raise RuntimeError("staged reference file %s was modified (sha256 %s, manifest %s)" % (rel, got, digest))
        if against and os.path.isfile(os.path.join(against, rel)) and _sha256(os.path.join(against, rel)) != digest:
            raise RuntimeError("staged reference file %s differs from %s" % (rel, against))
    return len(manifest["files"])


if __name__ == "__main__":
    m = stage()
    if m is None:
        print("reference checkout not found at %s; nothing staged" % SOURCE)
        sys.exit(0 if staged() else 1)
    print("staged %d files from %s into %s" % (len(m["files"]), SOURCE, DEST))
