"""Runs the product kernel SOURCES under the CUDA-on-CPU shim (tests/emu/) against the oracle.  CPU only, small
sizes.  This is a debugging aid for indexing/ordering -- the parity tests proper are tests/test_kernels_gpu.py."""
import pytest
import torch

from convexadam_b200._native import Lib
from tests import kernel_checks as kc
from tests.emu.build_emu import build

DEV = torch.device("cpu")


@pytest.fixture(scope="module")
def lib():
    return Lib(build(), expect_cuda=False)


def test_product_binding_rejects_emulation_build(lib):
    from convexadam_b200._native import NativeError
    with pytest.raises(NativeError):
        Lib(lib.path, expect_cuda=True)


def test_avgpool(lib): kc.check_avgpool(lib, DEV)
def test_mind(lib): kc.check_mind(lib, DEV, shape=(14, 12, 18))
def test_mind_rows(lib): kc.check_mind_rows(lib, DEV, shapes=((9, 7, 11), (6, 13, 130)))
def test_correlate(lib): kc.check_correlate(lib, DEV, cases=((12, 4, 5, 7, 2), (16, 3, 4, 8, 1), (3, 4, 4, 5, 3)))
def test_coupled(lib): kc.check_coupled(lib, DEV, h=4, w=5, d=8)
def test_certified(lib): kc.check_certified(lib, DEV, cases=((12, 4, 5, 8, 2), (3, 4, 5, 7, 1)))
def test_mesh_and_axes(lib): kc.check_mesh_and_axes(lib, DEV)
def test_inverse_consistency(lib): kc.check_inverse_consistency(lib, DEV, shape=(6, 7, 9))
def test_trilinear_affine_pack(lib): kc.check_trilinear_affine_pack(lib, DEV)
def test_onehot(lib): kc.check_onehot(lib, DEV)
def test_adam(lib): kc.check_adam(lib, DEV, shape=(6, 7, 8), steps=(1, 3))
def test_correlate_slab(lib): kc.check_correlate_slab(lib, DEV)
def test_cc_pieces(lib): kc.check_cc_pieces(lib, DEV)
def test_adam_window(lib): kc.check_adam_window(lib, DEV)
def test_adam_pools(lib): kc.check_adam_pools(lib, DEV, shape=(7, 6, 8))
def test_apply_convex(lib): kc.check_apply_convex(lib, DEV)


def test_kernel_timing_binding(lib):
    """The measurement aid of bench.py: the slot names map onto the C enum and the calls are harmless no-ops in this build."""
    lib.kernel_timing(("raw_ssd", "box_filters", "select_pass"))
    kc.check_certified(lib, DEV, cases=((3, 4, 5, 7, 1),), tie_case=False)
    assert lib.kernel_times() == {}
    lib.kernel_timing(())
    with pytest.raises(ValueError):
        lib.kernel_timing(("no_such_kernel",))


def test_random_small_shapes(lib):
    """Exact and certified correlate / coupled_convex on randomly drawn degenerate shapes (poisoned workspaces)."""
    shapes = kc.random_small_shapes(123, 10)
    kc.check_certified(lib, DEV, cases=shapes, tie_case=False)
    kc.check_correlate(lib, DEV, cases=shapes[:6])


def test_field_mean(lib): kc.check_field_mean(lib, DEV)


def test_certified_fused_kernel(lib, monkeypatch):
    """The one-kernel form of ca_correlate_fast (csrc/correlate_fused.cu; opt-in through CA_CORR_FUSED=1): tiles that cross
    the volume faces, several row / z tiles, d not a multiple of 4, every disp_hw the kernel is instantiated for."""
    monkeypatch.setenv("CA_CORR_FUSED", "1")
    kc.check_certified(lib, DEV, cases=((12, 4, 5, 8, 2), (3, 4, 5, 7, 1), (7, 3, 14, 35, 2), (4, 3, 4, 6, 4), (2, 2, 3, 5, 6)))
    kc.check_certified(lib, DEV, cases=kc.random_small_shapes(55, 6), tie_case=False)


def test_certified_pipe_kernel(lib, monkeypatch):
    """The producer / consumer form of ca_correlate_fast (csrc/correlate_pipe.cu; opt-in through CA_CORR_PIPE=1): CTAs of one
    launch that wait for each other through global counters (blocks of the emulation run on several OS threads), several
    row tiles and H blocks, small rings."""
    import ctypes
    fn = lib.dll.ca_debug_corr_pipe_error
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_int]
    monkeypatch.setenv("CA_CORR_PIPE", "1")
    fn(1)
    kc.check_certified(lib, DEV, cases=((12, 4, 5, 8, 2), (7, 3, 14, 36, 2), (5, 9, 35, 20, 3), (3, 21, 7, 8, 2), (3, 30, 20, 16, 1),
                                        (3, 12, 20, 128, 1)))     # the last: rows of one warp of quads, T_D inside the producers
    for knobs in ({"CA_PIPE_PB": "4", "CA_PIPE_R": "12"}, {"CA_PIPE_PB": "9", "CA_PIPE_R": "13", "CA_PIPE_NW": "1"},
                  {"CA_PIPE_PB": "60", "CA_PIPE_NW": "3"}, {"CA_PIPE_HANDOFF": "1", "CA_PIPE_PB": "5"}, {"CA_PIPE_TDP": "0"}):
        for k, v in knobs.items():
            monkeypatch.setenv(k, v)
        kc.check_certified(lib, DEV, cases=((7, 3, 14, 36, 2), (3, 21, 7, 8, 2), (2, 21, 7, 128, 2)), tie_case=False)
        for k in knobs:
            monkeypatch.delenv(k)
    assert fn(1) == 0, "a hand-off wait of the producer / consumer kernel gave up"


def test_mask_fill(lib): kc.check_mask_fill(lib, DEV)
