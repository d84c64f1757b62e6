"""Host-side orchestration (convexadam_b200/pipeline.py) exercised on the CPU-emulation build of the kernels and
compared with (a) the oracle's restatement of convex_adam_pt and (b) the reference's committed golden field.
CPU only; the same comparisons run on the real GPU in tests/test_pipeline_gpu.py."""
import numpy as np
import pytest
import torch

from convexadam_b200 import pipeline
from convexadam_b200._native import Lib
from oracle import ca_oracle as orc
from tests import kernel_checks as kc
from tests.emu.build_emu import build

DEV = torch.device("cpu")


@pytest.fixture(scope="module")
def lib():
    return Lib(build(), expect_cuda=False)


def run(lib, fixed, moving, **kw):
    prm = pipeline.Params(**kw)
    disp = pipeline.register_mind(torch.from_numpy(fixed.copy()), torch.from_numpy(moving.copy()), prm, lib)
    return lib.pack_field(disp).numpy()


@pytest.mark.parametrize("ic", [True, False])
def test_small_pair_vs_oracle(lib, ic):
    from convexadam_b200.synth import make_pair
    f, m = make_pair((24, 20, 28), seed=3, disp_scale=0.12)
    kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=2, selected_niter=3, selected_smooth=0,
              grid_sp_adam=2, ic=ic)
    got = run(lib, f.numpy(), m.numpy(), **kw)
    want = orc.convex_adam_pt(f.numpy(), m.numpy(), 1, 2, 1.25, 2, 2, 3, 0, 2, ic)
    assert got.shape == want.shape == (24, 20, 28, 3)
    assert np.abs(got - want).max() <= 1e-4


def test_two_pool_factors_and_smoothing(lib):
    from convexadam_b200.synth import make_pair
    f, m = make_pair((24, 24, 30), seed=4, disp_scale=0.2)
    kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=3, disp_hw=2, selected_niter=2, selected_smooth=3,
              grid_sp_adam=2, ic=True)
    got = run(lib, f.numpy(), m.numpy(), **kw)
    want = orc.convex_adam_pt(f.numpy(), m.numpy(), 1, 2, 1.25, 3, 2, 2, 3, 2, True)
    assert np.abs(got - want).max() <= 1e-4


def test_lambda_zero_skips_adam(lib):
    from convexadam_b200.synth import make_pair
    f, m = make_pair((16, 16, 20), seed=5)
    kw = dict(mind_r=1, mind_d=2, lambda_weight=0.0, grid_sp=2, disp_hw=2, selected_niter=5, selected_smooth=0,
              grid_sp_adam=2, ic=True)
    got = run(lib, f.numpy(), m.numpy(), **kw)
    want = orc.convex_adam_pt(f.numpy(), m.numpy(), 1, 2, 0.0, 2, 2, 5, 0, 2, True)
    assert np.abs(got - want).max() <= 1e-5


def test_golden_convex_stage(lib, golden):
    kc.check_golden_convex_stage(lib, DEV, golden("pair_gs2_hw3"), 3)


def test_golden_field_niter1(lib, golden):
    # niter=1 returns the convex-stage field smoothed by B^3; <= 1e-5 voxel vs the reference
    g = golden("pair_gs2_hw3")
    got = run(lib, g["fixed"], g["moving"], mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=3,
              selected_niter=1, selected_smooth=0, grid_sp_adam=2, ic=True)
    assert np.abs(got[::2, ::2, ::2] - g["field_niter1"]).max() <= 1e-5


def test_niter_zero_raises(lib):
    z = np.zeros((8, 8, 8), np.float32)
    with pytest.raises(ValueError):
        run(lib, z, z, lambda_weight=1.0, grid_sp=2, disp_hw=1, selected_niter=0)
