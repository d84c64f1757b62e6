"""Import the UNMODIFIED reference (multimodallearning/convexAdam) in a container that lacks
nibabel / SimpleITK.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used (a) by tests/golden/make_golden*.py to generate the committed golden
vectors, (b) by oracle validation scripts and (c) by bench.py's reference arm (`--impl reference`) and its
`reference_torch_cuda` key, which time the reference's own code.  `/root/reference` does not exist on the GPU box:
there the byte-for-byte copy staged by oracle/stage_reference.py under the git-ignored oracle/_ref/ is used.  Nothing in
the `-m gpu` tests or smoke() imports this module.

Protocol follows SURVEY.md §8(c): two stub modules in sys.modules, `/root/reference/src` on sys.path.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")    # oracle/stage_reference.py (git-ignored copy)


def _root():
    """The reference where its sources lie, else the byte-for-byte copy staged under oracle/_ref/ (the GPU box)."""
    for r in (os.environ.get("CONVEXADAM_REFERENCE"), "/root/reference", _STAGED):
        if r and os.path.isdir(os.path.join(r, "src", "convexAdam")):
            return r
    return None


REFERENCE_ROOT = _root() or "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "convexAdam"))


def import_reference():
    """Returns (convex_adam_utils, convex_adam_MIND) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "nibabel" not in sys.modules:
        nib = types.ModuleType("nibabel")

        class Nifti1Image:  # noqa: D401 - dummy, only used for isinstance checks
            pass

        nib.Nifti1Image = Nifti1Image
        sys.modules["nibabel"] = nib
    if "SimpleITK" not in sys.modules:
        sitk = types.ModuleType("SimpleITK")

        class Image:
            pass

        sitk.Image = Image
        sys.modules["SimpleITK"] = sitk
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import convexAdam.convex_adam_utils as ref_utils  # type: ignore
    import convexAdam.convex_adam_MIND as ref_mind  # type: ignore

    return ref_utils, ref_mind
