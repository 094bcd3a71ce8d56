#!/usr/bin/env python3
"""Recipe for oracle/_ref: the UNMODIFIED reference package, for the reference arms of bench.py, the hook tests and the
drop-in example tests on the GPU box (where /root/reference does not exist).

TEST / BENCH INFRASTRUCTURE ONLY - nothing under muggled_dpt_b200/ imports it.

The reference (heyoeyo/muggled_dpt) is pure Python. `pip install --target` of it cannot run in this image (its build
backend, hatchling, is not in the offline wheelhouse), so this script builds what the install would have produced, as ONE
archive: oracle/_ref/muggled_dpt_reference.zip holds the `muggled_dpt` package, file for file - minus the demo-only UI
toolkit (`demo_helpers/toadui`, `demo_helpers/3dviewer`) that the hot path never imports - plus the two
`simple_examples/` scripts the drop-in test runs with the import swapped. Python imports the package straight from the
archive (zipimport). oracle/_ref/ is git-ignored (no reference source ever enters this repository's history) but not
gpurun-ignored, so the archive travels to the GPU box like a built .so.

usage: python oracle/build_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DEST, "muggled_dpt_reference.zip")
SKIP_DIRS = {"__pycache__", "toadui", "3dviewer"}
EXAMPLES = ("depth_prediction.py", "internal_features.py")


def build_ref(src: str = "/root/reference", quiet: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards (False: no source tree here and no earlier archive)."""
    pkg = os.path.join(src, "muggled_dpt")
    if not os.path.isdir(pkg):
        return ref_available()
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    n = 0
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for root, dirs, files in os.walk(pkg):
            dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
            for f in sorted(files):
                if f.endswith(".py"):
                    full = os.path.join(root, f)
                    z.write(full, os.path.relpath(full, src))
                    n += 1
        for name in EXAMPLES:
            full = os.path.join(src, "simple_examples", name)
            if os.path.exists(full):
                z.write(full, os.path.join("simple_examples", name))
                n += 1
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write(f"muggled_dpt_reference.zip: {n} unmodified files archived by oracle/build_ref.py from {src} "
                "(git-ignored build output, not repository source)\n")
    if not quiet:
        print(f"oracle/_ref/muggled_dpt_reference.zip: {n} files from {src}")
    return True


def ref_available() -> bool:
    return os.path.isfile(ARCHIVE)


def import_reference():
    """imports the reference package from the archive under oracle/_ref (never from /root/reference); returns its
    make_dpt module"""
    if not ref_available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if ARCHIVE not in sys.path:
        sys.path.insert(0, ARCHIVE)
    from muggled_dpt import make_dpt

    return make_dpt


def read_example(name: str) -> str:
    """source text of one of the reference's simple_examples scripts (from the archive)"""
    with zipfile.ZipFile(ARCHIVE) as z:
        return z.read("simple_examples/" + name).decode()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ok = build_ref(ap.parse_args().src)
    sys.exit(0 if ok else 1)
