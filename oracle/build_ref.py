#!/usr/bin/env python3
"""Recipe for oracle/_ref: the UNMODIFIED reference package, for the reference arm of bench.py and for validating the
oracle restatement on the GPU box (where /root/reference does not exist).

TEST / BENCH INFRASTRUCTURE ONLY - nothing under muggled_dpt_b200/ imports it.

The reference (heyoeyo/muggled_dpt) is pure Python. `pip install --target` of it cannot run in this image (its build
backend, hatchling, is not in the offline wheelhouse), so this script does exactly what installing its wheel would do:
it places the `muggled_dpt` package directory, file for file, under oracle/_ref/ - minus the demo-only UI toolkit
(`demo_helpers/toadui`, `demo_helpers/3dviewer`) that the hot path never imports. It also places the two
`simple_examples/` scripts the drop-in test runs with the import swapped. oracle/_ref/ is git-ignored (no reference
source ever enters this repository's history) but not gpurun-ignored, so it travels to the GPU box like a built .so.

usage: python oracle/build_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def build_ref(src: str = "/root/reference", quiet: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards (False: no source tree here and no earlier copy)."""
    pkg = os.path.join(src, "muggled_dpt")
    if not os.path.isdir(pkg):
        return os.path.isdir(os.path.join(DEST, "muggled_dpt"))
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    skip = shutil.ignore_patterns("__pycache__", "*.pyc", "toadui", "3dviewer", "*.md")
    shutil.copytree(pkg, os.path.join(DEST, "muggled_dpt"), ignore=skip)
    ex = os.path.join(src, "simple_examples")
    if os.path.isdir(ex):
        os.makedirs(os.path.join(DEST, "simple_examples"))
        for name in ("depth_prediction.py", "internal_features.py"):
            if os.path.exists(os.path.join(ex, name)):
                shutil.copy2(os.path.join(ex, name), os.path.join(DEST, "simple_examples", name))
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write(f"copied by oracle/build_ref.py from {src} (unmodified; git-ignored build output, not repository source)\n")
    if not quiet:
        n = sum(len(fs) for _, _, fs in os.walk(DEST))
        print(f"oracle/_ref: {n} files from {src}")
    return True


def ref_available() -> bool:
    return os.path.isdir(os.path.join(DEST, "muggled_dpt"))


def import_reference():
    """imports the reference package from oracle/_ref (never from /root/reference) and returns the module"""
    if not ref_available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import muggled_dpt  # noqa: F401
    from muggled_dpt import make_dpt

    return make_dpt


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ok = build_ref(ap.parse_args().src)
    sys.exit(0 if ok else 1)
