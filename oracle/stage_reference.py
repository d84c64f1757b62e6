"""Stages the UNMODIFIED reference package into the git-ignored oracle/_ref/ so that it travels to the GPU box.

    python oracle/stage_reference.py          (also run by __graft_entry__.build() when /root/reference is present)

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (multimodallearning/convexAdam) is pure Python over torch: there
is nothing to compile, "building" oracle/_ref means copying `src/convexAdam/*.py` byte for byte from where the sources
lie (/root/reference) into oracle/_ref/src/convexAdam/ -- and the three image fixtures of the reference's own end-to-end
test into oracle/_ref/tests/ -- together with a manifest of their SHA-256 digests.  oracle/_ref/
is listed in .gitignore (never in history) but not in .gpurunignore (it ships with the snapshot like our own .so files).
Who may use it: bench.py's `--impl reference` arm and its `reference_torch_cuda` informational key (the reference's own
code on the box's host cores / on the same B200), and oracle/validate_vs_reference.py.  Nothing under convexadam_b200/.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CONVEXADAM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "convex_adam_utils.py", "convex_adam_MIND.py", "convex_adam_nnUNet.py", "apply_convex.py",
         "convex_adam_translation.py")
# the image fixtures of the reference's own end-to-end test (tests/test_convex_adam_mind.py:45-85): inputs and the golden
# warped image, for tests/test_fixture_io_free.py
FIXTURES = ("tests/input/10000/10000_1000000_t2w.mha", "tests/input/10000/10000_1000000_adc.mha",
            "tests/output-expected/10000/10000_1000000_adc_warped.mha")


def stage(verbose=True):
    src_pkg = os.path.join(SRC, "src", "convexAdam")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"[stage_reference] {src_pkg} not found: nothing staged (the GPU box uses what was staged in the build container)")
        return None
    dst_pkg = os.path.join(DST, "src", "convexAdam")
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for f in FILES:
        s = os.path.join(src_pkg, f)
        if not os.path.exists(s):
            continue
        shutil.copyfile(s, os.path.join(dst_pkg, f))
        with open(s, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    for rel in FIXTURES:
        src_f = os.path.join(SRC, rel)
        if os.path.exists(src_f):
            os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
            shutil.copyfile(src_f, os.path.join(DST, rel))
            with open(src_f, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    lic = os.path.join(SRC, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(DST, "LICENSE"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_pkg, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"[stage_reference] {len(manifest)} files -> {dst_pkg}")
    return dst_pkg


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
