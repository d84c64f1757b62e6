"""GPU parity tests proper: every kernel of libconvexadam_b200 (through the C ABI) against the CPU oracle on the
same seeded inputs, against the reference's committed golden vectors, and -- at the BASELINE full size --
through size-independent properties.  Run with `-m gpu` on the B200 box."""
import numpy as np
import pytest
import torch

from tests import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from convexadam_b200._native import get_lib
    return get_lib()


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_avgpool(lib, dev): kc.check_avgpool(lib, dev)
def test_mind(lib, dev): kc.check_mind(lib, dev, shape=(40, 34, 46))
def test_mind_rows(lib, dev): kc.check_mind_rows(lib, dev, shapes=((9, 7, 11), (40, 34, 46), (6, 13, 130), (20, 18, 300)))
def test_mesh_and_axes(lib, dev): kc.check_mesh_and_axes(lib, dev)
def test_inverse_consistency(lib, dev): kc.check_inverse_consistency(lib, dev, shape=(20, 18, 22))
def test_trilinear_affine_pack(lib, dev): kc.check_trilinear_affine_pack(lib, dev)
def test_onehot(lib, dev): kc.check_onehot(lib, dev)
def test_adam(lib, dev): kc.check_adam(lib, dev, shape=(20, 18, 22), steps=(1, 2, 7))
def test_adam_16ch(lib, dev): kc.check_adam(lib, dev, C=16, shape=(12, 14, 16), steps=(1, 4))


def test_correlate(lib, dev):
    kc.check_correlate(lib, dev, cases=((12, 10, 9, 14, 3), (16, 6, 7, 8, 1), (3, 5, 6, 9, 4), (12, 12, 10, 16, 2),
                                        (12, 6, 5, 7, 5), (12, 8, 6, 12, 6), (4, 4, 4, 4, 7), (12, 5, 5, 8, 8)))


def test_correlate_slab(lib, dev):
    kc.check_correlate_slab(lib, dev)
    kc.check_correlate_slab(lib, dev, C=12, h=20, w=14, d=24, hw=3, slabs=((0, 10), (10, 20), (7, 8), (18, 20)))


def test_cc_pieces(lib, dev): kc.check_cc_pieces(lib, dev, C=12, h=11, w=9, d=12, hw=3, cut=5)
def test_adam_window(lib, dev): kc.check_adam_window(lib, dev, C=12, shape=(44, 10, 16))
def test_adam_pools(lib, dev): kc.check_adam_pools(lib, dev)


def test_coupled(lib, dev):
    kc.check_coupled(lib, dev, h=12, w=10, d=16, hw=3)
    kc.check_coupled(lib, dev, h=6, w=7, d=9, hw=6)


def test_apply_convex(lib, dev):
    kc.check_apply_convex(lib, dev, shape=(20, 17, 23))


def test_apply_convex_matches_scipy_and_api(lib, dev):
    """The public drop-in (convexadam_b200.apply_convex.apply_convex) against scipy.ndimage.map_coordinates itself -- the
    call the reference makes (apply_convex.py:23) -- on a 96x80x128 volume."""
    from scipy.ndimage import map_coordinates
    from convexadam_b200.apply_convex import apply_convex
    rs = np.random.RandomState(3)
    H, W, D = 96, 80, 128
    mov = (rs.rand(H, W, D) * 1000).astype(np.float32)
    disp = (rs.randn(H, W, D, 3) * 4).astype(np.float64)
    ident = np.meshgrid(np.arange(H), np.arange(W), np.arange(D), indexing="ij")
    # numpy input: validate_image casts to float64 (utils.py:276), so the reference interpolates and returns float64
    want = map_coordinates(mov.astype(np.float64), disp.transpose(3, 0, 1, 2) + ident, order=1)
    got = apply_convex(disp, mov)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * 1000                 # float64 sums in a different corner order
    # torch float32 input keeps its dtype, as in the reference
    want32 = map_coordinates(mov, disp.transpose(3, 0, 1, 2) + ident, order=1)
    got32 = apply_convex(torch.from_numpy(disp), torch.from_numpy(mov))
    assert got32.dtype == np.float32 and (got32 == want32).mean() > 0.9999 and np.abs(got32 - want32).max() <= 1e-4


def test_field_mean(lib, dev): kc.check_field_mean(lib, dev, shape=(33, 20, 41))


def test_random_small_shapes(lib, dev):
    shapes = kc.random_small_shapes(321, 16, max_hw=5)
    kc.check_certified(lib, dev, cases=shapes, tie_case=False)
    kc.check_correlate(lib, dev, cases=shapes[:8])


def test_certified(lib, dev):
    kc.check_certified(lib, dev, cases=((12, 10, 9, 16, 3), (16, 6, 7, 8, 1), (3, 5, 6, 9, 4), (12, 8, 6, 12, 6), (12, 7, 9, 10, 2)))


@pytest.fixture
def fused_correlate(monkeypatch):
    """CA_CORR_FUSED=1: ca_correlate_fast runs as ONE kernel (csrc/correlate_fused.cu: tensor-map TMA staging, raw SSD, both
    box filters and the H march fused; opt-in, read on every call)."""
    monkeypatch.setenv("CA_CORR_FUSED", "1")


def test_certified_fused_kernel(lib, dev, fused_correlate):
    kc.check_certified(lib, dev, cases=((12, 10, 9, 16, 3), (16, 6, 7, 8, 1), (3, 5, 6, 9, 4), (12, 8, 6, 12, 6), (12, 7, 9, 10, 2),
                                        (12, 9, 30, 70, 2), (5, 5, 14, 33, 5)))
    kc.check_certified(lib, dev, cases=kc.random_small_shapes(77, 8, max_hw=6), tie_case=False)


def test_full_size_fused_kernel_certified_equals_exact(lib, full_volume, fused_correlate):
    """The fused kernel's volume at BASELINE config 2 size: inside the bound, and the same certified selections."""
    test_full_size_certified_equals_exact(lib, full_volume)


@pytest.fixture
def pipe_correlate(monkeypatch):
    """CA_CORR_PIPE=1: ca_correlate_fast runs as ONE launch of producer / consumer CTAs that hand the raw SSD planes over through
    a ring in L2 (csrc/correlate_pipe.cu; opt-in, read on every call)."""
    monkeypatch.setenv("CA_CORR_PIPE", "1")


def _pipe_error_flag(lib):
    import ctypes
    fn = lib.dll.ca_debug_corr_pipe_error
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_int]
    return fn(1)


def test_certified_pipe_kernel(lib, dev, pipe_correlate):
    """Shapes with several row tiles, several H blocks, ragged last tiles and every ring / block / chunk knob: rows of a multiple
    of 4 voxels go through the kernel, the others fall back to the two-kernel path (same checks)."""
    _pipe_error_flag(lib)
    n0 = lib.launch_count()
    kc.check_certified(lib, dev, cases=((12, 10, 9, 16, 3), (16, 6, 7, 8, 1), (3, 30, 6, 12, 4), (12, 8, 6, 12, 6), (12, 40, 35, 64, 2),
                                        (12, 9, 30, 72, 2), (5, 5, 14, 36, 5), (4, 26, 50, 20, 1), (12, 30, 21, 128, 3), (5, 9, 40, 128, 6)))
    assert lib.launch_count() > n0
    assert _pipe_error_flag(lib) == 0, "a hand-off wait of the producer / consumer kernel gave up"


@pytest.mark.parametrize("knobs", [{"CA_PIPE_PB": "4", "CA_PIPE_R": "12"}, {"CA_PIPE_PB": "9", "CA_PIPE_R": "13", "CA_PIPE_NW": "1"},
                                   {"CA_PIPE_PB": "60", "CA_PIPE_NW": "3"}, {"CA_PIPE_PB": "8", "CA_PIPE_NW": "2", "CA_PIPE_R": "40"},
                                   {"CA_PIPE_TDP": "0"}, {"CA_PIPE_HANDOFF": "1"}, {"CA_PIPE_HANDOFF": "1", "CA_PIPE_PB": "5", "CA_PIPE_NW": "2"}])
def test_certified_pipe_kernel_knobs(lib, dev, pipe_correlate, monkeypatch, knobs):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    _pipe_error_flag(lib)
    kc.check_certified(lib, dev, cases=((12, 40, 35, 64, 2), (6, 21, 18, 24, 3), (7, 20, 19, 128, 2)), tie_case=False)
    assert _pipe_error_flag(lib) == 0


def _pipe_config(lib):
    import ctypes
    cfg = (ctypes.c_int * 6)()
    lib.dll.ca_debug_corr_pipe_config(cfg)
    return dict(zip(("NW", "PB", "R", "tdp", "handoff", "launches"), cfg))


def test_full_size_pipe_kernel_certified_equals_exact(lib, full_volume, pipe_correlate):
    """The producer / consumer kernel's volume at BASELINE config 2 size: inside the bound, and the same certified selections."""
    _pipe_error_flag(lib)
    n0 = _pipe_config(lib)["launches"]
    test_full_size_certified_equals_exact(lib, full_volume)
    assert _pipe_error_flag(lib) == 0
    cfg = _pipe_config(lib)
    print("producer / consumer kernel ran with", cfg)
    assert cfg["launches"] > n0 and cfg["tdp"] == 1, "the full-size volume did not go through k_corr_pipe (T_D inside the producers)"


def test_correlate_rejects_bad_arguments(lib, dev):
    from convexadam_b200._native import NativeError
    f = torch.zeros(12, 4, 4, 4, device=dev)
    with pytest.raises(NativeError):
        lib.correlate(f, f, 9)          # disp_hw outside 1..8
    with pytest.raises(ValueError):
        lib.correlate(f, f[:, :2], 2)
    with pytest.raises(NativeError):
        lib.correlate(f.cpu(), f.cpu(), 2)   # no CPU path


@pytest.mark.parametrize("name,hw", [("pair_gs2_hw3", 3), ("pair_defaults", 4)])
def test_golden_convex_stage(lib, dev, golden, name, hw):
    kc.check_golden_convex_stage(lib, dev, golden(name), hw)


# ---- BASELINE.json config 2 sizes: size-independent properties ------------------------------------------------
FULL = dict(C=12, h=96, w=80, d=128, hw=6)


@pytest.fixture(scope="module")
def full_volume(lib, dev):
    from oracle import ca_oracle as orc
    g = torch.Generator().manual_seed(5)
    C, h, w, d, hw = FULL["C"], FULL["h"], FULL["w"], FULL["d"], FULL["hw"]
    low = torch.rand(2, C, h // 4 + 2, w // 4 + 2, d // 4 + 2, generator=g)
    smooth = torch.nn.functional.interpolate(low, size=(h, w, d), mode="trilinear", align_corners=True)
    feats = (0.7 * smooth + 0.3 * torch.rand(2, C, h, w, d, generator=g)).float().contiguous()
    fix, mov = feats[0], feats[1]
    ssd, am = lib.correlate(fix.to(dev), mov.to(dev), hw)
    torch.cuda.synchronize()
    return fix, mov, ssd, am, orc


def test_full_size_correlate_point_samples(full_volume):
    """1500 random (voxel,label) entries of the 8.6 GB volume, bit-exact against the oracle's 5x5x5 point evaluation
    (faces, edges and corners of the volume and of the search window included)."""
    fix, mov, ssd, am, orc = full_volume
    C, h, w, d, hw = FULL["C"], FULL["h"], FULL["w"], FULL["d"], FULL["hw"]
    L = (2 * hw + 1) ** 3
    rs = np.random.RandomState(0)
    pts = [(rs.randint(L), rs.randint(h), rs.randint(w), rs.randint(d)) for _ in range(1400)]
    pts += [(l, x, y, z) for l in (0, L - 1, L // 2) for x in (0, h - 1) for y in (0, w - 1) for z in (0, 1, d - 2, d - 1)]
    idx = torch.tensor(pts, device=ssd.device)
    got = ssd[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]].cpu().numpy()
    fn, mn = fix.numpy(), mov.numpy()
    for (l, x, y, z), gv in zip(pts, got):
        want = np.float32(orc.correlate_point(fn, mn, hw, l, x, y, z))
        assert np.float32(gv).view(np.uint32) == want.view(np.uint32), f"ssd[{l},{x},{y},{z}] = {gv!r} vs oracle {want!r}"


def test_full_size_argmin_is_first_minimum(full_volume):
    fix, mov, ssd, am, orc = full_volume
    ref = torch.empty_like(am)
    n_chunk = 8
    hs = ssd.shape[1] // n_chunk
    for i in range(n_chunk):                               # torch.argmin == first minimum on CUDA as well
        ref[i * hs:(i + 1) * hs] = torch.argmin(ssd[:, i * hs:(i + 1) * hs], 0)
    assert bool((ref == am).all())


def test_full_size_coupled_sampled_columns(lib, full_volume):
    """Each of the six passes at 400 random voxels: argmin of ssd + coeff*|mesh - soft|^2 recomputed on the host in
    the reference's op order from the downloaded cost columns and the kernel's own previous disp_soft."""
    fix, mov, ssd, am, orc = full_volume
    hw = FULL["hw"]
    L, h, w, d = ssd.shape
    mesh = lib.disp_mesh(hw, ssd.device)
    soft, sel = lib.coupled_convex(ssd, am, mesh, hw, want_sel=True)
    rs = np.random.RandomState(1)
    vox = np.stack([rs.randint(1, h - 1, 400), rs.randint(1, w - 1, 400), rs.randint(1, d - 1, 400)], 1)
    vt = torch.tensor(vox, device=ssd.device)
    cols = ssd[:, vt[:, 0], vt[:, 1], vt[:, 2]].cpu().numpy()           # (L, 400)
    mesh_n = mesh.cpu().numpy()
    sel_all = torch.cat([am[None], sel], 0).cpu().numpy()                # (7,h,w,d): selection entering pass j
    coeffs = np.array([0.003, 0.01, 0.03, 0.1, 0.3, 1], np.float32)
    for j in range(6):
        prev = sel_all[j]
        for k, (x, y, z) in enumerate(vox):
            nb = prev[x - 1:x + 2, y - 1:y + 2, z - 1:z + 2].reshape(-1)
            s = np.zeros(3, np.float32)
            for l in nb:                                                 # ascending (dH,dW,dD) 27-tap sum, / 27
                s = (s + mesh_n[:, l]).astype(np.float32)
            s = (s / np.float32(27)).astype(np.float32)
            dd = (mesh_n - s[:, None]).astype(np.float32)
            q = (dd * dd).astype(np.float32)
            dist = ((q[0] + q[1]).astype(np.float32) + q[2]).astype(np.float32)
            cost = (cols[:, k] + (coeffs[j] * dist).astype(np.float32)).astype(np.float32)
            assert int(np.argmin(cost)) == int(sel_all[j + 1][x, y, z]), f"pass {j} voxel {(x, y, z)}"
    assert torch.isfinite(soft).all()


def test_full_size_identity_pair(lib, dev):
    """fixed == moving: the zero-displacement label has cost exactly 0 in the interior and wins everywhere it can."""
    C, h, w, d, hw = 12, 48, 40, 64, 6
    g = torch.Generator().manual_seed(9)
    f = torch.rand(C, h, w, d, generator=g).to(dev)
    ssd, am = lib.correlate(f, f, hw)
    S = 2 * hw + 1
    centre = hw * S * S + hw * S + hw
    assert float(ssd[centre].abs().max()) == 0.0
    assert bool((am == centre).all())
    soft = lib.coupled_convex(ssd, am, lib.disp_mesh(hw, dev), hw)
    assert float(soft.abs().max()) <= 1e-6


def test_full_size_certified_equals_exact(lib, full_volume):
    """BASELINE config 2 size: the approximate volume stays inside its stated bound (every one of the 2.16e9 entries), and
    argmin, the six selections and disp_soft of the certified path are those of the exact path, bit for bit."""
    fix, mov, ssd, am, orc = full_volume
    hw = FULL["hw"]
    dev = ssd.device
    mesh = lib.disp_mesh(hw, dev)
    soft_e, sel_e = lib.coupled_convex(ssd, am, mesh, hw, want_sel=True)
    f, m = fix.to(dev), mov.to(dev)
    ssd_f, _ = lib.correlate(f, m, hw, fast=True)
    bound = 98 * 2.0 ** -24
    worst = 0.0
    for l0 in range(0, ssd.shape[0], 169):
        a, b = ssd[l0:l0 + 169].double(), ssd_f[l0:l0 + 169].double()
        worst = max(worst, float(((b - a).abs() / a.clamp_min(1e-30)).max()))
    assert worst <= bound, f"fast cost volume: max relative deviation {worst:.3e} > {bound:.3e}"
    soft_c, sel_c, flagged = lib.coupled_convex_certified(ssd_f, f, m, mesh, hw, want_sel=True, want_flagged=True)
    assert bool((sel_c[0] == am).all()), "certified argmin"
    for j in range(6):
        assert bool((sel_c[j + 1] == sel_e[j]).all()), f"certified selection of pass {j}"
    assert bool((soft_c.view(torch.int32) == soft_e.view(torch.int32)).all())
    fl = flagged.cpu().numpy()
    print(f"full size: max relative deviation of the fast volume {worst:.3e}; voxels re-decided exactly per selection {fl.tolist()} of {am.numel()}")
    assert fl.sum() > 0 and fl.max() < am.numel() // 50


def test_mask_fill(lib, dev): kc.check_mask_fill(lib, dev, shapes=((12, 10, 14), (8, 16, 6), (20, 18, 22), (48, 40, 64)))
