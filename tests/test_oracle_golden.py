"""Pins the CPU oracle (oracle/ca_oracle.c) to golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import ca_oracle as orc

CASES = {
    "pair_gs2_hw3": dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=3, grid_sp_adam=2, smooth=0),
    "pair_defaults": dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4, grid_sp_adam=2, smooth=3),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())


@pytest.fixture(scope="module", params=list(CASES))
def case(request, golden):
    g = golden(request.param)
    return request.param, CASES[request.param], g


def test_kwargs_recorded(case):
    name, kw, g = case
    rec = dict(zip([str(k) for k in g["kw_keys"]], g["kw_vals"]))
    assert rec["grid_sp"] == kw["grid_sp"] and rec["disp_hw"] == kw["disp_hw"] and rec["selected_smooth"] == kw["smooth"]


def test_mind_and_pool(case):
    # utils.py:24-68 + MIND.py:118-119.  Tolerance 3e-7 abs on features in [0,1]: torch uses SLEEF expf and a
    # float cascade mean, the oracle libm expf and a double mean.
    name, kw, g = case
    for which in ("fix", "mov"):
        img = g["fixed" if which == "fix" else "moving"]
        f = orc.mindssc(img, kw["mind_r"], kw["mind_d"])
        assert np.abs(f.reshape(-1)[g[f"mind_{which}_idx"]] - g[f"mind_{which}_val"]).max() <= 3e-7
        pooled = orc.avgpool3d(f, kw["grid_sp"], kw["grid_sp"])
        assert np.abs(pooled - g[f"{which}_s"]).max() <= 3e-7


def test_mesh_bit_exact(case):
    name, kw, g = case
    assert bits_equal(orc.disp_mesh(kw["disp_hw"]), g["mesh"])


@pytest.fixture(scope="module")
def cost_volumes(case):
    name, kw, g = case
    C, h, w, d = g["fix_s"].shape
    S = 2 * kw["disp_hw"] + 1
    tail = (h * S * S * w * d) % 32      # torch-CPU sum(0) tail on this class of host, see ca_oracle.c
    fwd = orc.correlate(g["fix_s"], g["mov_s"], kw["disp_hw"], tail=tail)
    bwd = orc.correlate(g["mov_s"], g["fix_s"], kw["disp_hw"], tail=tail)
    return fwd, bwd


def test_correlate_bit_exact(case, cost_volumes):
    name, kw, g = case
    (ssd, am), (ssd_, am_) = cost_volumes
    assert sha(ssd) == str(g["ssd_sha"]) and sha(ssd_) == str(g["ssd__sha"])
    assert (am == g["argmin"]).all() and (am_ == g["argmin_"]).all()
    assert bits_equal(ssd.reshape(-1)[g["ssd_idx"]], g["ssd_val"])


def test_correlate_canonical_order_same_argmin(case, cost_volumes):
    # the CUDA kernel sums channels strictly sequentially (tail=0); only the last n%32 columns may differ by an ulp
    name, kw, g = case
    ssd0, am0 = orc.correlate(g["fix_s"], g["mov_s"], kw["disp_hw"], tail=0)
    assert (am0 == g["argmin"]).all()
    assert int((ssd0.view(np.uint32) != cost_volumes[0][0].view(np.uint32)).sum()) <= 32 * 27 * 27
    assert np.abs(ssd0 - cost_volumes[0][0]).max() <= 2e-6


def test_correlate_point_matches_volume(case, cost_volumes):
    name, kw, g = case
    ssd0, _ = orc.correlate(g["fix_s"], g["mov_s"], kw["disp_hw"], tail=0)
    L, h, w, d = ssd0.shape
    rs = np.random.RandomState(0)
    for _ in range(200):
        l, x, y, z = rs.randint(L), rs.randint(h), rs.randint(w), rs.randint(d)
        v = orc.correlate_point(g["fix_s"], g["mov_s"], kw["disp_hw"], l, x, y, z)
        assert np.float32(v).view(np.uint32) == ssd0[l, x, y, z].view(np.uint32)


def test_coupled_convex_bit_exact(case, cost_volumes):
    name, kw, g = case
    (ssd, am), (ssd_, am_) = cost_volumes
    assert bits_equal(orc.coupled_convex(ssd, am, g["mesh"]), g["soft"])
    assert bits_equal(orc.coupled_convex(ssd_, am_, g["mesh"]), g["soft_"])


def test_inverse_consistency_bit_exact(case):
    name, kw, g = case
    _, h, w, d = g["soft"].shape
    scale = (np.array([h - 1, w - 1, d - 1], np.float32) / np.float32(2)).reshape(3, 1, 1, 1)
    a = np.ascontiguousarray((g["soft"] / scale)[::-1])
    b = np.ascontiguousarray((g["soft_"] / scale)[::-1])
    ice, _ = orc.inverse_consistency(a, b, 15)
    assert bits_equal(ice, g["ice"])
    H, W, D = [int(v) for v in g["shape"]]
    hr = orc.trilinear(np.ascontiguousarray(ice[::-1] * scale * np.float32(kw["grid_sp"])), (H, W, D))
    assert np.abs(hr[:, ::2, ::2, ::2] - g["disp_hr0"]).max() <= 4e-6 * max(1.0, np.abs(g["disp_hr0"]).max())


def test_adam_teacher_forced(case):
    # SURVEY 7.2 P2.  |dp| <= |dg|/eps with eps=1e-8, lr=1: the first step (v=0) amplifies fp32 gradient noise.
    name, kw, g = case
    H, W, D = [int(v) for v in g["shape"]]
    gsa = kw["grid_sp_adam"]
    if kw["grid_sp"] != gsa:
        pytest.skip("Adam-grid features not stored for this case")
    pf, pm = g["fix_s"], g["mov_s"]
    ks = sorted({int(k[4:].split("_")[0]) for k in g.files if k.startswith("adam")})
    for k in ks:
        st = orc.AdamState(g[f"adam{k}_p"])
        st.m[:] = g[f"adam{k}_m"]; st.v[:] = g[f"adam{k}_v"]; st.step = k
        u, gp = orc.adam_step(st, pf, pm, kw["lambda_weight"], want_grad=True)
        gr = g[f"adam{k}_g"]
        assert np.abs(gp - gr).max() <= 5e-6 * np.abs(gr).max()
        assert np.abs(st.p - g[f"adam{k}_p_next"]).max() <= (2e-4 if k == 0 else 2e-5)


def test_full_path_vs_reference_field(case):
    # SURVEY 7.2 P1 (niter=1: convex stage + B^3, <=1e-5 voxel) and P3 (free-running, niter<=15 pre-chaos, <=1e-3)
    name, kw, g = case
    stages = {}
    niter = int(g["niter"])
    f = orc.convex_adam_pt(g["fixed"], g["moving"], kw["mind_r"], kw["mind_d"], kw["lambda_weight"], kw["grid_sp"],
                           kw["disp_hw"], 1, kw["smooth"], kw["grid_sp_adam"], True, stages)
    assert (stages["argmin"] == g["argmin"]).all() and (stages["argmin_"] == g["argmin_"]).all()
    assert np.abs(f[::2, ::2, ::2] - g["field_niter1"]).max() <= 1e-5
    f = orc.convex_adam_pt(g["fixed"], g["moving"], kw["mind_r"], kw["mind_d"], kw["lambda_weight"], kw["grid_sp"],
                           kw["disp_hw"], niter, kw["smooth"], kw["grid_sp_adam"], True)
    assert np.abs(f[::2, ::2, ::2] - g["field"]).max() <= 1e-3


def test_apply_convex_oracle_vs_scipy():
    """apply_convex.py:23 calls scipy.ndimage.map_coordinates(order=1); the oracle's numpy restatement against scipy
    itself (the third-party library where the arithmetic lives, present in this image): float32 images exact, float64
    within a few ulp (corner accumulation order), out-of-volume points zero."""
    import numpy as np
    from scipy.ndimage import map_coordinates
    from oracle import ca_oracle as orc
    rs = np.random.RandomState(11)
    H, W, D = 12, 10, 14
    disp = (rs.randn(H, W, D, 3) * 3).astype(np.float64)
    disp[0, 0, 0] = 0
    disp[-1, -1, -1] = 0
    ident = np.meshgrid(np.arange(H), np.arange(W), np.arange(D), indexing="ij")
    for dt in (np.float32, np.float64):
        mov = (rs.rand(H, W, D) * 50).astype(dt)
        want = map_coordinates(mov, disp.transpose(3, 0, 1, 2) + ident, order=1)
        got = orc.apply_convex(disp, mov)
        assert got.dtype == want.dtype
        assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= (0.0 if dt == np.float32 else 1e-13 * 50)
    assert (want == 0).any() and (want != 0).any()
