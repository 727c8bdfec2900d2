"""Recipe: place the UNMODIFIED reference sources under git-ignored ``oracle/_ref/`` (TEST INFRASTRUCTURE).

    python -m oracle.fetch_ref            (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python, so there is nothing to compile: the "build output" of this recipe is a verbatim
copy of the reference's ``*.py`` / ``config/*.yaml`` files (388 KB), byte for byte, with a manifest of their
sha256 digests.  ``oracle/_ref/`` is listed in ``.gitignore`` (the sources never enter the history) but not in
``.gpurunignore``, so the copy travels to the GPU box exactly like the built ``libnrw.so`` does; there it is
what ``bench.py --impl reference``, ``cpu_baseline`` and ``reference_torch_gpu`` execute (``kind: "reference"``),
and what ``tests/test_gpu_dropin.py`` instantiates ``NeuconWSystem`` from.  Nothing in the product path
(``neuralrecon-w_b200/``) ever imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NRW_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yaml", ".yml")
SKIP_DIRS = {".git", "assets", "__pycache__", "deeplabv3_config"}


def fetch(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "rendering", "renderer.py")):
        if verbose:
            print(f"fetch_ref: {SRC} not present; keeping whatever oracle/_ref already holds", file=sys.stderr)
        return False
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        rel = os.path.relpath(root, SRC)
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            src = os.path.join(root, f)
            dst = os.path.join(DST, rel, f) if rel != "." else os.path.join(DST, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            data = open(src, "rb").read()
            manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(data).hexdigest()
            if not (os.path.isfile(dst) and open(dst, "rb").read() == data):
                shutil.copyfile(src, dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"fetch_ref: {len(manifest)} reference files -> {DST}")
    return True


def verify():
    """True when every file of the manifest is present and unmodified (sha256)."""
    mpath = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    files = json.load(open(mpath))["files"]
    for rel, digest in files.items():
        p = os.path.join(DST, rel)
        if not os.path.isfile(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != digest:
            return False
    return True


if __name__ == "__main__":
    ok = fetch()
    sys.exit(0 if ok or verify() else 1)
