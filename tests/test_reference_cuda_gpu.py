"""The CUDA path next to the reference's OWN torch-CUDA path on the same GPU (SURVEY App. B.5 iii: the probes the
survey could not run).  Uses the byte-for-byte copy of the reference staged under the git-ignored oracle/_ref/
(oracle/stage_reference.py); skipped where it is absent.  The reference's CUDA path is not bit-deterministic
(avg_pool3d_backward atomics) and its default dtype is half, so these are measured distances with loose bounds: the
parity gates proper are the CPU-fp32 golden vectors (tests/test_config1_parity.py, tests/test_pipeline_gpu.py).
"""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_smooth=0, grid_sp_adam=2, ic=True)


@pytest.fixture(scope="module")
def ref_mind():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not staged (oracle/_ref)")
    torch.backends.cudnn.allow_tf32 = False            # the reference's MINDSSC is built from conv3d calls
    torch.backends.cuda.matmul.allow_tf32 = False
    return ref_shim.import_reference()[1]


def pair():
    from convexadam_b200.synth import make_pair
    return make_pair((64, 64, 64), seed=21, disp_scale=0.08)


def test_fp32_vs_reference_cuda_fp32(ref_mind, golden):
    """ours(fp32) vs reference CUDA fp32, and both against the reference's CPU fp32 golden field, 1 and 15 iterations."""
    from convexadam_b200.convex_adam_MIND import convex_adam_pt
    dev = torch.device("cuda:0")
    fixed, moving = pair()
    g = golden("config1_ns")
    for n in (1, 15):
        ours = convex_adam_pt(fixed, moving, selected_niter=n, dtype=torch.float32, device=dev, **KW)
        ref = ref_mind.convex_adam_pt(fixed, moving, selected_niter=n, dtype=torch.float32, device=dev, **KW)
        ref2 = ref_mind.convex_adam_pt(fixed, moving, selected_niter=n, dtype=torch.float32, device=dev, **KW)
        cpu = g[f"field_n{n}"]
        d_or = float(np.abs(ours - ref).max())
        print(f"[cuda] 64^3 niter={n}: ours vs reference-CUDA-fp32 {d_or:.3e}; reference-CUDA-fp32 vs reference-CPU-fp32 "
              f"{float(np.abs(ref[::2, ::2, ::2] - cpu).max()):.3e}; ours vs reference-CPU-fp32 {float(np.abs(ours[::2, ::2, ::2] - cpu).max()):.3e}; "
              f"reference-CUDA run-to-run {float(np.abs(ref - ref2).max()):.3e} voxel")
        assert np.isfinite(ours).all() and d_or <= 0.25         # sanity bound only; the numbers are the result


def test_default_dtype_warns_and_distance_to_reference_fp16(ref_mind):
    """dtype=float16 (the default argument): fp32 arithmetic + half rounding of the result, announced by a one-time
    RuntimeWarning; the distance to the reference's genuine half-precision CUDA run is measured and printed."""
    import convexadam_b200.convex_adam_MIND as M
    dev = torch.device("cuda:0")
    fixed, moving = pair()
    M._FP16_WARNED = False
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        ours16 = M.convex_adam_pt(fixed, moving, selected_niter=15, device=dev, **KW)
    assert any(issubclass(w.category, RuntimeWarning) and "float16" in str(w.message) for w in rec)
    ours32 = M.convex_adam_pt(fixed, moving, selected_niter=15, dtype=torch.float32, device=dev, **KW)
    assert (ours16 == ours32.astype(np.float16).astype(np.float64)).all()
    ref16 = ref_mind.convex_adam_pt(fixed, moving, selected_niter=15, device=dev, **KW)        # dtype default: half
    ref32 = ref_mind.convex_adam_pt(fixed, moving, selected_niter=15, dtype=torch.float32, device=dev, **KW)
    d = np.abs(ours16 - ref16)
    print(f"[cuda] 64^3 niter=15 default dtype: ours(fp32 compute, half output) vs reference-CUDA-fp16: max {float(d.max()):.3e}, "
          f"mean {float(d.mean()):.3e} voxel; reference-CUDA-fp16 vs reference-CUDA-fp32: max {float(np.abs(ref16 - ref32).max()):.3e}, "
          f"mean {float(np.abs(ref16 - ref32).mean()):.3e} voxel")
    assert np.isfinite(ref16).all()
