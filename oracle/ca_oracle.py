"""ctypes/numpy front end of oracle/ca_oracle.c -- the CPU checker.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under convexadam_b200/ imports this module.

All arrays are C-contiguous numpy float32 (int64 for label indices); shapes follow the reference
(`/root/reference/src/convexAdam/convex_adam_utils.py`, `convex_adam_MIND.py`).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libca_oracle.so")
    src = os.path.join(_HERE, "ca_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libca_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_correlate_point.restype = C.c_float
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.c_void_p)


def avgpool3d(x, k, stride, pad=0):
    """x: (..., T, H, W).  F.avg_pool3d(x, k, stride=stride, padding=pad)."""
    x, xp = _f(x)
    T, H, W = x.shape[-3:]
    nc = int(np.prod(x.shape[:-3], dtype=np.int64))
    o = [(s + 2 * pad - k) // stride + 1 for s in (T, H, W)]
    out = np.empty(x.shape[:-3] + tuple(o), np.float32)
    lib().orc_avgpool3d(xp, C.c_int64(nc), T, H, W, k, stride, pad, out.ctypes.data_as(C.c_void_p))
    return out


def mind_pairs():
    a = np.zeros((12, 3), np.int32)
    b = np.zeros((12, 3), np.int32)
    lib().orc_mind_pairs(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    return a, b


def mindssc(img, radius=2, dilation=2):
    """img (H,W,D) -> (12,H,W,D); reference MINDSSC (utils.py:24-68)."""
    img, ip = _f(img)
    H, W, D = img.shape
    out = np.empty((12, H, W, D), np.float32)
    lib().orc_mindssc(ip, H, W, D, radius, dilation, out.ctypes.data_as(C.c_void_p))
    return out


def correlate(fix, mov, disp_hw, tail=0, want_argmin=True):
    """fix, mov (C,h,w,d) -> ssd (L,h,w,d), argmin (h,w,d) int64; reference correlate (utils.py:72-89)."""
    fix, fp = _f(fix)
    mov, mp = _f(mov)
    Cn, h, w, d = fix.shape
    S = 2 * disp_hw + 1
    ssd = np.empty((S ** 3, h, w, d), np.float32)
    am = np.empty((h, w, d), np.int64) if want_argmin else None
    lib().orc_correlate(fp, mp, Cn, h, w, d, disp_hw, ssd.ctypes.data_as(C.c_void_p),
                        am.ctypes.data_as(C.c_void_p) if want_argmin else None, C.c_int64(tail))
    return ssd, am


def correlate_point(fix, mov, disp_hw, label, x, y, z):
    fix, fp = _f(fix)
    mov, mp = _f(mov)
    Cn, h, w, d = fix.shape
    return float(lib().orc_correlate_point(fp, mp, Cn, h, w, d, disp_hw, C.c_int64(int(label)), int(x), int(y), int(z)))


def coupled_convex(ssd, argmin, mesh, want_sel=False):
    """ssd (L,h,w,d), argmin (h,w,d), mesh (3,L) -> soft (3,h,w,d) [, sel (7,h,w,d)]; utils.py:93-109."""
    ssd, sp = _f(ssd)
    argmin, ap = _i64(argmin)
    mesh, mp = _f(np.asarray(mesh).reshape(3, -1))
    L, h, w, d = ssd.shape
    soft = np.empty((3, h, w, d), np.float32)
    sel = np.empty((7, h, w, d), np.int64) if want_sel else None
    lib().orc_coupled_convex(sp, ap, mp, h, w, d, C.c_int64(L), soft.ctypes.data_as(C.c_void_p),
                             sel.ctypes.data_as(C.c_void_p) if want_sel else None)
    return (soft, sel) if want_sel else soft


def grid_sample3d(inp, grid):
    """inp (C,iD,iH,iW), grid (oD,oH,oW,3) xyz -> (C,oD,oH,oW); bilinear, zeros, align_corners=False."""
    inp, ip = _f(inp)
    grid, gp = _f(grid)
    Cn, iD, iH, iW = inp.shape
    oD, oH, oW = grid.shape[:3]
    out = np.empty((Cn, oD, oH, oW), np.float32)
    lib().orc_grid_sample3d(ip, Cn, iD, iH, iW, gp, oD, oH, oW, out.ctypes.data_as(C.c_void_p))
    return out


def _linspace_pm1(n):
    """torch.linspace(-1, 1, n) on CPU, fp32, bit-exact: the kernel (RangeFactoriesKernel.cpp) evaluates
    start + step*i (i < n//2) / end - step*(n-1-i) with the multiply-add CONTRACTED to one fma (the file
    is built with -mfma); an fp32 fma equals the fp64 evaluation rounded once (the product is exact)."""
    step = np.float32(np.float32(2.0) / np.float32(n - 1))
    i = np.arange(n)
    lo = (np.float64(step) * i + np.float64(-1.0)).astype(np.float32)
    hi = (np.float64(1.0) - np.float64(step) * (n - 1 - i)).astype(np.float32)
    return np.where(i < n // 2, lo, hi).astype(np.float32)


def identity_axes(h, w, d):
    """1-D axes of F.affine_grid(eye(3,4)[None], (1,1,h,w,d), align_corners=False): (gx[d], gy[w], gz[h]).

    torch: linspace(-1, 1, n) * (n - 1) / n (AffineGridGenerator.cpp linspace_from_neg_one).
    """
    def ax(n):
        if n <= 1:
            return np.zeros(n, np.float32)
        return (_linspace_pm1(n) * np.float32(n - 1) / np.float32(n)).astype(np.float32)
    return ax(d), ax(w), ax(h)


def disp_mesh(disp_hw):
    """disp_mesh_t of convex_adam_MIND.py:127 as (3, L) fp32: hw * affine_grid(align_corners=True) axes.

    Row c holds the displacement along array axis c of label l = kD*S^2 + kW*S + kH.
    """
    S = 2 * disp_hw + 1
    i = np.arange(S)
    base = _linspace_pm1(S)
    v = (base * np.float32(disp_hw)).astype(np.float32)  # theta = hw * I: x*hw + 0 + 0 + 0
    # affine_grid output (1,S,S,S,3) with [...,0] = x over the LAST size dim; permute(0,4,1,2,3).reshape(3,-1)
    # => row 0 varies with the last (fastest) index of l, row 2 with the slowest.
    kD, kW, kH = np.meshgrid(i, i, i, indexing="ij")
    mesh = np.stack([v[kH].reshape(-1), v[kW].reshape(-1), v[kD].reshape(-1)], 0)
    return np.ascontiguousarray(mesh, np.float32)


def inverse_consistency(f1, f2, iters):
    """f1, f2: (3,h,w,d) normalised, channel 0 = x (last axis); utils.py:114-129."""
    f1, p1 = _f(f1)
    f2, p2 = _f(f2)
    _, h, w, d = f1.shape
    gx, gy, gz = identity_axes(h, w, d)
    o1 = np.empty_like(f1)
    o2 = np.empty_like(f2)
    lib().orc_inverse_consistency(p1, p2, h, w, d, iters, gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p),
                                  gz.ctypes.data_as(C.c_void_p), o1.ctypes.data_as(C.c_void_p),
                                  o2.ctypes.data_as(C.c_void_p))
    return o1, o2


def trilinear(x, size):
    """x (C,h,w,d) -> (C,*size); F.interpolate(mode='trilinear', align_corners=False)."""
    x, xp = _f(x)
    Cn, ih, iw, idd = x.shape
    out = np.empty((Cn,) + tuple(size), np.float32)
    lib().orc_trilinear(xp, Cn, ih, iw, idd, size[0], size[1], size[2], out.ctypes.data_as(C.c_void_p))
    return out


class AdamState:
    """State of the Adam loop at convex_adam_MIND.py:155-179 (p = net[0].weight, exp_avg, exp_avg_sq)."""

    def __init__(self, p0):
        self.p = np.ascontiguousarray(p0, np.float32).copy()
        self.m = np.zeros_like(self.p)
        self.v = np.zeros_like(self.p)
        self.step = 0


def adam_step(st, fix, mov, lambda_weight, want_grad=False):
    """One iteration; returns u = B^3(p) of this iteration's forward (and dL/dp if asked)."""
    fix, fp = _f(fix)
    mov, mp = _f(mov)
    Cn, ha, wa, da = fix.shape
    gx, gy, gz = identity_axes(ha, wa, da)
    u = np.empty_like(st.p)
    g = np.empty_like(st.p) if want_grad else None
    st.step += 1
    lib().orc_adam_step(st.p.ctypes.data_as(C.c_void_p), st.m.ctypes.data_as(C.c_void_p),
                        st.v.ctypes.data_as(C.c_void_p), fp, mp, Cn, ha, wa, da, C.c_float(lambda_weight),
                        gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p), gz.ctypes.data_as(C.c_void_p),
                        st.step, u.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p) if want_grad else None)
    return (u, g) if want_grad else u


def convex_adam_pt(img_fixed, img_moving, mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=6, disp_hw=4,
                   selected_niter=80, selected_smooth=0, grid_sp_adam=2, ic=True, stages=None):
    """Whole path of convex_adam_MIND.py:64-202 in fp32 (the CPU dtype policy, :89-91), no mask.

    Returns the (H,W,D,3) float64 field; `stages` (a dict) receives intermediates when given.
    """
    imgf = np.ascontiguousarray(img_fixed, np.float32)
    imgm = np.ascontiguousarray(img_moving, np.float32)
    H, W, D = imgf.shape
    ff = mindssc(imgf, mind_r, mind_d)
    fm = mindssc(imgm, mind_r, mind_d)
    out = field_from_features(ff, fm, (H, W, D), lambda_weight, grid_sp, disp_hw, selected_niter, selected_smooth,
                              grid_sp_adam, ic, stages)
    return out


def field_from_features(ff, fm, shape, lambda_weight, grid_sp, disp_hw, selected_niter, selected_smooth,
                        grid_sp_adam, ic, stages=None):
    H, W, D = shape
    st = stages if stages is not None else {}
    fs = avgpool3d(ff, grid_sp, grid_sp)
    ms = avgpool3d(fm, grid_sp, grid_sp)
    h, w, d = fs.shape[1:]
    mesh = disp_mesh(disp_hw)
    ssd, am = correlate(fs, ms, disp_hw)
    soft = coupled_convex(ssd, am, mesh)
    st.update(fix_s=fs, mov_s=ms, argmin=am, soft=soft)
    if ic:
        scale = (np.array([h - 1, w - 1, d - 1], np.float32) / np.float32(2)).reshape(3, 1, 1, 1)
        ssd_, am_ = correlate(ms, fs, disp_hw)
        soft_ = coupled_convex(ssd_, am_, mesh)
        a = np.ascontiguousarray((soft / scale)[::-1])
        b = np.ascontiguousarray((soft_ / scale)[::-1])
        ice, _ = inverse_consistency(a, b, 15)
        st.update(argmin_=am_, soft_=soft_, ice=ice)
        disp_hr = trilinear(np.ascontiguousarray(ice[::-1] * scale * np.float32(grid_sp)), (H, W, D))
    else:
        disp_hr = soft
    st["disp_hr0"] = disp_hr
    if lambda_weight > 0:
        pf = avgpool3d(ff, grid_sp_adam, grid_sp_adam)
        pm = avgpool3d(fm, grid_sp_adam, grid_sp_adam)
        ha, wa, da = pf.shape[1:]
        disp_lr = trilinear(disp_hr, (ha, wa, da))
        state = AdamState(disp_lr / np.float32(grid_sp_adam))
        u = None
        for _ in range(selected_niter):
            u = adam_step(state, pf, pm, lambda_weight)
        disp_hr = trilinear(u * np.float32(grid_sp_adam), (H, W, D))
        if selected_smooth > 0:
            k = selected_smooth
            for _ in range(3):
                disp_hr = avgpool3d(disp_hr, k, 1, k // 2)
        st.update(adam_state=state, u=u)
    return np.stack([disp_hr[0], disp_hr[1], disp_hr[2]], 3).astype(np.float64)


def apply_convex(disp, moving):
    """apply_convex (reference apply_convex.py:13-24) = scipy.ndimage.map_coordinates(moving, disp^T + identity, order=1,
    mode='constant', cval=0): float64 linear interpolation at x + disp(x); 0 where a coordinate leaves [0, size-1]; the
    upper neighbour index clamped.  scipy (third-party, where the arithmetic lives) is restated here in numpy;
    tests/test_oracle_golden.py pins it against scipy itself."""
    moving = np.asarray(moving)
    disp = np.asarray(disp, np.float64)
    H, W, D = moving.shape
    ident = np.meshgrid(np.arange(H), np.arange(W), np.arange(D), indexing="ij")
    c = [disp[..., k] + ident[k] for k in range(3)]
    inside = np.ones(moving.shape, bool)
    for k, n in enumerate((H, W, D)):
        inside &= ~((c[k] < 0) | (c[k] > n - 1))
    f = [np.floor(np.where(inside, c[k], 0.0)) for k in range(3)]
    t = [np.where(inside, c[k], 0.0) - f[k] for k in range(3)]
    i0 = [f[k].astype(np.int64) for k in range(3)]
    i1 = [np.minimum(i0[k] + 1, n - 1) for k, n in enumerate((H, W, D))]
    m64 = moving.astype(np.float64)
    v = np.zeros(moving.shape, np.float64)
    for corner in range(8):
        sel = [(corner >> 2) & 1, (corner >> 1) & 1, corner & 1]
        idx = tuple(i1[k] if sel[k] else i0[k] for k in range(3))
        w = np.ones(moving.shape, np.float64)
        w = (t[0] if sel[0] else 1.0 - t[0]) * (t[1] if sel[1] else 1.0 - t[1]) * (t[2] if sel[2] else 1.0 - t[2])
        v += m64[idx] * w
    out_dtype = moving.dtype if moving.dtype in (np.float32, np.float64) else np.float64
    return np.where(inside, v, 0.0).astype(out_dtype)


def translation_from_field(field, spacing_zyx, seg=None):
    """convex_adam_translation.py:88-104: mean of the (H,W,D,3) field (inside seg > 0 when given), divided by the moving
    spacing, rounded to whole voxels, back to mm, reversed to (x,y,z)."""
    field = np.asarray(field, np.float64)
    t = field[np.asarray(seg) > 0].mean(axis=0) if seg is not None else field.mean(axis=(0, 1, 2))
    spacing = np.asarray(spacing_zyx, np.float64)
    return t, tuple(list((np.round(t / spacing, decimals=0) * spacing)[::-1]))


def fill_outside_mask(img, mask):
    """The use_mask pre-processing of extract_features (reference convex_adam_MIND.py:36-54) for one image, restated with
    numpy and -- like the reference itself -- scipy's distance transform: erode the mask with the replicate-padded 3x3x3
    mean (> 0.9), nearest inside voxel on the 2x subsampled grid, trilinear 2x upsampling, original intensities inside.
    Returns (filled image, eroded mask, EDT indices (3,h,w,d))."""
    from scipy.ndimage import distance_transform_edt as edt
    img = np.ascontiguousarray(img, np.float32)
    mask = np.ascontiguousarray(mask, np.float32)
    H, W, D = img.shape
    padded = np.pad(mask, 1, mode="edge")
    s = np.zeros_like(mask)
    for dx in range(3):                 # torch avg_pool3d: ascending (dH, dW, dD) fp32 sum, then a true division
        for dy in range(3):
            for dz in range(3):
                s = s + padded[dx:dx + H, dy:dy + W, dz:dz + D]
    m = ((s / np.float32(27)) > np.float32(0.9)).astype(np.float32)
    _, idx = edt(m[::2, ::2, ::2] == 0, return_indices=True)
    sub = img[::2, ::2, ::2].reshape(-1)[idx[0] * D // 2 * W // 2 + idx[1] * D // 2 + idx[2]]
    h2, w2, d2 = sub.shape
    filled = trilinear(sub[None].astype(np.float32), (2 * h2, 2 * w2, 2 * d2))[0]
    return np.where(m != 0, img, filled), m, idx.astype(np.int32)
