"""ctypes binding of libconvexadam_b200.so (include/convexadam_b200.h) over torch tensors.

PyTorch is plumbing here: device memory, the current stream, host<->device copies.  All arithmetic on the hot
path happens inside the library's sm_100a kernels.  There is NO fallback: if the library is missing, was not
built for CUDA, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes as C
import os
from functools import lru_cache

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libconvexadam_b200.so")

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_SIGS = {
    "ca_abi_version": (C.c_int, []),
    "ca_last_error": (C.c_char_p, []),
    "ca_is_cuda_build": (C.c_int, []),
    "ca_launch_count": (C.c_ulonglong, []),
    "ca_kernel_timing": (None, [_i]),
    "ca_kernel_times": (_i, [C.POINTER(C.c_double), C.POINTER(_i), _i]),
    "ca_mindssc_workspace_bytes": (_sz, [_i, _i, _i]),
    "ca_mindssc_pooled": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "ca_avgpool3d": (_i, [_vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ca_onehot_pooled": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp, _vp]),
    "ca_correlate_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ca_correlate_min_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ca_correlate": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_correlate_slab_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "ca_correlate_slab": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_cc_select_workspace_bytes": (_sz, [_i64]),
    "ca_cc_blockmin": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_cc_select": (_i, [_vp, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ca_cc_smooth": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ca_correlate_slab_fast": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ca_cc_certified_workspace_bytes": (_sz, [_i64]),
    "ca_cc_blockmin_certified": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ca_cc_select_certified": (_i, [_vp, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ca_adam_window_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "ca_adam_run_window": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                _i, _vp]),
    "ca_coupled_convex_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ca_coupled_convex": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_correlate_fast": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ca_coupled_convex_certified_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ca_coupled_convex_certified": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ca_inverse_consistency_workspace_bytes": (_sz, [_i, _i, _i]),
    "ca_inverse_consistency": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ca_field_affine": (_i, [_vp, _vp, _i64, _i, C.POINTER(_f), C.POINTER(_f), _f, _vp]),
    "ca_trilinear": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _f, _vp]),
    "ca_adam_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ca_adam_run": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ca_adam_run_gauss": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_adam_run_pools": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "ca_pack_field": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "ca_apply_convex": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ca_field_mean_workspace_bytes": (_sz, []),
    "ca_field_mean": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _sz, _vp]),
    "ca_mask_erode": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "ca_nearest_inside_workspace_bytes": (_sz, [_i, _i, _i]),
    "ca_nearest_inside": (_i, [_vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
}
EXPORTS = tuple(_SIGS)
ABI_VERSION = 1


class NativeError(RuntimeError):
    pass


@lru_cache(maxsize=64)
def identity_axes_cpu(h, w, d):
    """1-D axes (x over d, y over w, z over h) of F.affine_grid(eye, (1,1,h,w,d), align_corners=False), computed
    by torch on the host exactly as the reference does (utils.py:121, MIND.py:160)."""
    g = F.affine_grid(torch.eye(3, 4).unsqueeze(0), (1, 1, h, w, d), align_corners=False)[0]
    return g[0, 0, :, 0].contiguous(), g[0, :, 0, 1].contiguous(), g[:, 0, 0, 2].contiguous()


@lru_cache(maxsize=16)
def disp_mesh_cpu(disp_hw):
    """disp_mesh_t of MIND.py:127 as a (3, L) fp32 CPU tensor."""
    S = 2 * disp_hw + 1
    m = F.affine_grid(disp_hw * torch.eye(3, 4).unsqueeze(0), (1, 1, S, S, S), align_corners=True)
    return m.permute(0, 4, 1, 2, 3).reshape(3, -1).contiguous()


class Lib:
    """One loaded build of the C ABI.  `expect_cuda=True` (the product) refuses the CPU-emulation test build."""

    def __init__(self, path=LIB_PATH, expect_cuda=True):
        if not os.path.exists(path):
            raise NativeError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(nvcc, sm_100a). convexadam_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError as e:
                raise NativeError(f"{path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if self.dll.ca_abi_version() != ABI_VERSION:
            raise NativeError(f"ABI mismatch: library {self.dll.ca_abi_version()} != binding {ABI_VERSION}")
        self.is_cuda = bool(self.dll.ca_is_cuda_build())
        if expect_cuda and not self.is_cuda:
            raise NativeError(f"{path} is not a CUDA build")
        self._consts = {}

    def launch_count(self) -> int:
        return int(self.dll.ca_launch_count())

    KERNEL_SLOTS = ("raw_ssd", "box_filters", "blockmin_argmin", "select_pass", "exact_redecision", "adam_grad", "adam_box3",
                    "corr_fused", "corr_pipe")

    def kernel_timing(self, slots=()):
        """Bracket the launches of the named kernels (KERNEL_SLOTS) with CUDA events on the launch stream; () = off."""
        mask = 0
        for name in slots:
            mask |= 1 << self.KERNEL_SLOTS.index(name)
        self.dll.ca_kernel_timing(mask)

    def kernel_times(self):
        """{slot: (total ms, launches)} of everything recorded since kernel_timing(); synchronise first."""
        n = len(self.KERNEL_SLOTS)
        ms, cnt = (C.c_double * n)(), (_i * n)()
        rc = self.dll.ca_kernel_times(ms, cnt, n)
        if rc < 0:
            self._chk(rc)
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(self.KERNEL_SLOTS) if cnt[i]}

    # -- helpers -----------------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise NativeError(f"libconvexadam_b200 error {rc}: {self.dll.ca_last_error().decode()}")

    def _t(self, t, dtype=torch.float32):
        if not isinstance(t, torch.Tensor):
            raise TypeError("expected a torch.Tensor")
        if self.is_cuda and not t.is_cuda:
            raise NativeError("convexadam_b200 kernels need CUDA tensors (there is no CPU path)")
        if not self.is_cuda and t.is_cuda:
            raise NativeError("the CPU-emulation test build got a CUDA tensor")
        if t.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {t.dtype}")
        return t.contiguous()

    def _stream(self, ref):
        return C.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream) if self.is_cuda else C.c_void_p(0)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    poison_workspaces = False   # tests: fill every workspace with 0xFF (NaN patterns) so that any read of unwritten scratch shows

    def _ws(self, nbytes, ref):
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=ref.device)
        if self.poison_workspaces:
            ws.fill_(255)
        return ws

    def const(self, key, maker, device):
        k = (key, str(device))
        if k not in self._consts:
            self._consts[k] = tuple(x.to(device) for x in maker())
            # first use only: the pageable host-to-device copies above are ordered on the creating stream alone, while the
            # cached tensors are later read from any stream (pipelined batches, the two-stream convex stage)
            if torch.device(device).type == "cuda":
                torch.cuda.current_stream(device).synchronize()
        return self._consts[k]

    def identity_axes(self, h, w, d, device):
        return self.const(("axes", h, w, d), lambda: identity_axes_cpu(h, w, d), device)

    def disp_mesh(self, disp_hw, device):
        return self.const(("mesh", disp_hw), lambda: (disp_mesh_cpu(disp_hw),), device)[0]

    # -- stages ------------------------------------------------------------------------------------------
    def mindssc_pooled(self, img, radius, dilation, gs_a, gs_b=0):
        img = self._t(img)
        H, W, D = img.shape
        out_a = torch.empty((12, H // gs_a, W // gs_a, D // gs_a), dtype=torch.float32, device=img.device)
        out_b = torch.empty((12, H // gs_b, W // gs_b, D // gs_b), dtype=torch.float32, device=img.device) if gs_b else None
        nb = self.dll.ca_mindssc_workspace_bytes(H, W, D)
        ws = self._ws(nb, img)
        self._chk(self.dll.ca_mindssc_pooled(self._p(img), H, W, D, radius, dilation, gs_a, self._p(out_a), gs_b,
                                            self._p(out_b), self._p(ws), nb, self._stream(img)))
        return out_a, out_b

    def avgpool3d(self, x, k, stride, pad=0):
        x = self._t(x)
        T, H, W = x.shape[-3:]
        nc = 1
        for s in x.shape[:-3]:
            nc *= s
        o = [(s + 2 * pad - k) // stride + 1 for s in (T, H, W)]
        out = torch.empty(tuple(x.shape[:-3]) + tuple(o), dtype=torch.float32, device=x.device)
        self._chk(self.dll.ca_avgpool3d(self._p(x), nc, T, H, W, k, stride, pad, self._p(out), self._stream(x)))
        return out

    def onehot_pooled(self, labels, lut, weight, gs):
        labels = self._t(labels)
        lut = self._t(lut, torch.int32)
        weight = self._t(weight)
        H, W, D = labels.shape
        Cn = weight.numel()
        out = torch.empty((Cn, H // gs, W // gs, D // gs), dtype=torch.float32, device=labels.device)
        self._chk(self.dll.ca_onehot_pooled(self._p(labels), H, W, D, self._p(lut), lut.numel(), self._p(weight), Cn, gs,
                                           self._p(out), self._stream(labels)))
        return out

    def correlate(self, fix, mov, disp_hw, want_argmin=True, workspace_bytes=None, out=None, fast=False):
        """fast=True: the approximate volume of ca_correlate_fast (no argmin; feed it to coupled_convex_certified)."""
        fix = self._t(fix)
        mov = self._t(mov)
        if fix.shape != mov.shape or fix.dim() != 4:
            raise ValueError("correlate: fix/mov must both be (C,h,w,d)")
        Cn, h, w, d = fix.shape
        L = (2 * disp_hw + 1) ** 3
        ssd = out if out is not None else torch.empty((L, h, w, d), dtype=torch.float32, device=fix.device)
        am = torch.empty((h, w, d), dtype=torch.int64, device=fix.device) if (want_argmin and not fast) else None
        nb = workspace_bytes or self.dll.ca_correlate_workspace_bytes(Cn, h, w, d, disp_hw)
        ws = self._ws(nb, fix)
        if fast:
            self._chk(self.dll.ca_correlate_fast(self._p(fix), self._p(mov), Cn, h, w, d, disp_hw, self._p(ssd), self._p(ws), nb,
                                                self._stream(fix)))
            return ssd, None
        self._chk(self.dll.ca_correlate(self._p(fix), self._p(mov), Cn, h, w, d, disp_hw, self._p(ssd), self._p(am),
                                       self._p(ws), nb, self._stream(fix)))
        return ssd, am

    def correlate_slab(self, fix, mov, disp_hw, x_begin, x_end, want_argmin=True, out=None, fast=False):
        """correlate for the H planes [x_begin, x_end) of the coarse grid only (fix/mov: whole (C,h,w,d) volumes).
        fast=True: the approximate volume of the certified mode (no argmin)."""
        fix = self._t(fix)
        mov = self._t(mov)
        Cn, h, w, d = fix.shape
        L = (2 * disp_hw + 1) ** 3
        hs = x_end - x_begin
        ssd = out if out is not None else torch.empty((L, hs, w, d), dtype=torch.float32, device=fix.device)
        am = torch.empty((hs, w, d), dtype=torch.int64, device=fix.device) if (want_argmin and not fast) else None
        nb = self.dll.ca_correlate_slab_workspace_bytes(Cn, h, w, d, disp_hw, x_begin, x_end)
        ws = self._ws(nb, fix)
        if fast:
            self._chk(self.dll.ca_correlate_slab_fast(self._p(fix), self._p(mov), Cn, h, w, d, disp_hw, x_begin, x_end, self._p(ssd),
                                                     self._p(ws), nb, self._stream(fix)))
            return ssd, None
        self._chk(self.dll.ca_correlate_slab(self._p(fix), self._p(mov), Cn, h, w, d, disp_hw, x_begin, x_end, self._p(ssd),
                                            self._p(am), self._p(ws), nb, self._stream(fix)))
        return ssd, am

    def cc_blockmin(self, ssd, disp_hw, want_argmin=True):
        """Block minima (S*S planes) and slab minima (S more planes) of a slab's cost volume, (S*S+S,hs,w,d), and its plain
        argmin, in one pass."""
        ssd = self._t(ssd)
        n = ssd[0].numel()
        S = 2 * disp_hw + 1
        bmin = torch.empty((S * S + S,) + tuple(ssd.shape[1:]), dtype=torch.float32, device=ssd.device)
        am = torch.empty(ssd.shape[1:], dtype=torch.int64, device=ssd.device) if want_argmin else None
        nb = self.dll.ca_cc_select_workspace_bytes(n)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_cc_blockmin(self._p(ssd), n, disp_hw, self._p(bmin), self._p(am), self._p(ws), nb, self._stream(ssd)))
        return bmin, am

    def cc_select(self, ssd, disp_hw, mesh=None, soft=None, coeff=0.0, bmin=None, prev=None):
        """One pass of coupled_convex over a slab: ssd (L,hs,w,d), soft (3,hs,w,d) -> labels (hs,w,d) int64.
        bmin (from cc_blockmin) + prev (the previous selection) switch on the exact pruning."""
        ssd = self._t(ssd)
        n = ssd[0].numel()
        sel = torch.empty(ssd.shape[1:], dtype=torch.int64, device=ssd.device)
        mesh = self._t(mesh.reshape(3, -1)) if mesh is not None else None
        soft = self._t(soft) if soft is not None else None
        bmin = self._t(bmin) if bmin is not None else None
        prev = self._t(prev, torch.int64) if prev is not None else None
        nb = self.dll.ca_cc_select_workspace_bytes(n)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_cc_select(self._p(ssd), n, disp_hw, self._p(mesh), self._p(soft), float(coeff), self._p(bmin),
                                       self._p(prev), self._p(sel), self._p(ws), nb, self._stream(ssd)))
        return sel

    def cc_blockmin_certified(self, ssd, disp_hw, mesh, fix, mov, x_begin):
        """Certified counterpart of cc_blockmin for a slab of the approximate volume: (bmin (S*S+2*S,hs,w,d), exact argmin)."""
        ssd = self._t(ssd)
        fix = self._t(fix)
        mov = self._t(mov)
        mesh = self._t(mesh.reshape(3, -1))
        n = ssd[0].numel()
        S = 2 * disp_hw + 1
        Cn, h, w, d = fix.shape
        bmin = torch.empty((S * S + 2 * S,) + tuple(ssd.shape[1:]), dtype=torch.float32, device=ssd.device)
        am = torch.empty(ssd.shape[1:], dtype=torch.int64, device=ssd.device)
        nb = self.dll.ca_cc_certified_workspace_bytes(n)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_cc_blockmin_certified(self._p(ssd), n, disp_hw, self._p(mesh), self._p(bmin), self._p(am), self._p(fix),
                                                   self._p(mov), Cn, h, w, d, x_begin, self._p(ws), nb, self._stream(ssd)))
        return bmin, am

    def cc_select_certified(self, ssd, disp_hw, mesh, soft, coeff, bmin, prev, fix, mov, x_begin):
        """Certified counterpart of cc_select(..., bmin=, prev=) for a slab of the approximate volume."""
        ssd = self._t(ssd)
        fix = self._t(fix)
        mov = self._t(mov)
        mesh = self._t(mesh.reshape(3, -1))
        soft = self._t(soft)
        bmin = self._t(bmin)
        prev = self._t(prev, torch.int64)
        n = ssd[0].numel()
        Cn, h, w, d = fix.shape
        sel = torch.empty(ssd.shape[1:], dtype=torch.int64, device=ssd.device)
        nb = self.dll.ca_cc_certified_workspace_bytes(n)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_cc_select_certified(self._p(ssd), n, disp_hw, self._p(mesh), self._p(soft), float(coeff), self._p(bmin),
                                                 self._p(prev), self._p(sel), self._p(fix), self._p(mov), Cn, h, w, d, x_begin,
                                                 self._p(ws), nb, self._stream(ssd)))
        return sel

    def cc_smooth(self, sel_window, mesh, disp_hw):
        """box3(mesh[:, sel]) for the inner planes of a (P,w,d) label window (negative label = outside the volume)."""
        sel_window = self._t(sel_window, torch.int64)
        mesh = self._t(mesh.reshape(3, -1))
        P, w, d = sel_window.shape
        soft = torch.empty((3, P - 2, w, d), dtype=torch.float32, device=sel_window.device)
        self._chk(self.dll.ca_cc_smooth(self._p(sel_window), self._p(mesh), P, w, d, disp_hw, self._p(soft),
                                       self._stream(sel_window)))
        return soft

    def coupled_convex(self, ssd, argmin, mesh, disp_hw, want_sel=False):
        ssd = self._t(ssd)
        argmin = self._t(argmin, torch.int64) if argmin is not None else None      # None: taken inside, no extra pass
        mesh = self._t(mesh.reshape(3, -1))
        L, h, w, d = ssd.shape
        if L != (2 * disp_hw + 1) ** 3 or mesh.shape[1] != L:
            raise ValueError("coupled_convex: label count mismatch")
        soft = torch.empty((3, h, w, d), dtype=torch.float32, device=ssd.device)
        sel = torch.empty((6, h, w, d), dtype=torch.int64, device=ssd.device) if want_sel else None
        nb = self.dll.ca_coupled_convex_workspace_bytes(h, w, d, disp_hw)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_coupled_convex(self._p(ssd), self._p(argmin), self._p(mesh), h, w, d, disp_hw, self._p(soft),
                                            self._p(sel), self._p(ws), nb, self._stream(ssd)))
        return (soft, sel) if want_sel else soft

    def coupled_convex_certified(self, ssd, fix, mov, mesh, disp_hw, want_sel=False, want_flagged=False):
        """coupled_convex (argmin included) on the approximate volume of correlate(fast=True); selections proven equal
        to the exact path's.  Returns soft[, sel (7,h,w,d): argmin + six passes][, flagged (7,) int32 counts]."""
        ssd = self._t(ssd)
        fix = self._t(fix)
        mov = self._t(mov)
        mesh = self._t(mesh.reshape(3, -1))
        L, h, w, d = ssd.shape
        if L != (2 * disp_hw + 1) ** 3 or mesh.shape[1] != L or tuple(fix.shape[1:]) != (h, w, d) or fix.shape != mov.shape:
            raise ValueError("coupled_convex_certified: shape mismatch")
        soft = torch.empty((3, h, w, d), dtype=torch.float32, device=ssd.device)
        sel = torch.empty((7, h, w, d), dtype=torch.int64, device=ssd.device) if want_sel else None
        flg = torch.zeros((8,), dtype=torch.int32, device=ssd.device) if want_flagged else None
        nb = self.dll.ca_coupled_convex_certified_workspace_bytes(h, w, d, disp_hw)
        ws = self._ws(nb, ssd)
        self._chk(self.dll.ca_coupled_convex_certified(self._p(ssd), self._p(fix), self._p(mov), fix.shape[0], self._p(mesh), h, w, d,
                                                      disp_hw, self._p(soft), self._p(sel), self._p(flg), self._p(ws), nb,
                                                      self._stream(ssd)))
        out = (soft,)
        if want_sel:
            out += (sel,)
        if want_flagged:
            out += (flg[:7],)
        return out if len(out) > 1 else soft

    def inverse_consistency(self, f1, f2, iters):
        f1 = self._t(f1)
        f2 = self._t(f2)
        _, h, w, d = f1.shape
        gx, gy, gz = self.identity_axes(h, w, d, f1.device)
        o1 = torch.empty_like(f1)
        o2 = torch.empty_like(f2)
        nb = self.dll.ca_inverse_consistency_workspace_bytes(h, w, d)
        ws = self._ws(nb, f1)
        self._chk(self.dll.ca_inverse_consistency(self._p(f1), self._p(f2), h, w, d, iters, self._p(gx), self._p(gy),
                                                 self._p(gz), self._p(o1), self._p(o2), self._p(ws), nb, self._stream(f1)))
        return o1, o2

    def field_affine(self, x, flip=False, div=(1.0, 1.0, 1.0), mul1=(1.0, 1.0, 1.0), mul2=1.0):
        x = self._t(x)
        out = torch.empty_like(x)
        n = x.numel() // 3
        self._chk(self.dll.ca_field_affine(self._p(x), self._p(out), n, int(flip), (_f * 3)(*div), (_f * 3)(*mul1), mul2,
                                          self._stream(x)))
        return out

    def trilinear(self, x, size, post_mul=1.0):
        x = self._t(x)
        Cn, ih, iw, idd = x.shape
        out = torch.empty((Cn,) + tuple(size), dtype=torch.float32, device=x.device)
        self._chk(self.dll.ca_trilinear(self._p(x), Cn, ih, iw, idd, self._p(out), size[0], size[1], size[2], post_mul,
                                       self._stream(x)))
        return out

    def mask_erode(self, mask):
        """(avg3(mask) > 0.9).float() with the replicate-padded 3x3x3 mean of MIND.py:38-42; mask (H,W,D) float32."""
        mask = self._t(mask)
        H, W, D = mask.shape
        out = torch.empty_like(mask)
        self._chk(self.dll.ca_mask_erode(self._p(mask), H, W, D, self._p(out), self._stream(mask)))
        return out

    def nearest_inside(self, inside):
        """(3,h,w,d) int32 coordinates of the nearest voxel with inside != 0 (scipy's distance_transform_edt indices)."""
        inside = self._t(inside)
        h, w, d = inside.shape
        idx = torch.empty((3, h, w, d), dtype=torch.int32, device=inside.device)
        nb = self.dll.ca_nearest_inside_workspace_bytes(h, w, d)
        ws = self._ws(nb, inside)
        self._chk(self.dll.ca_nearest_inside(self._p(inside), h, w, d, self._p(idx), self._p(ws), nb, self._stream(inside)))
        return idx

    def adam_run(self, fix, mov, lambda_weight, p, m, v, first_step, niter, want_grad=False, pools=None, gauss=None):
        """pools: None = the reference's B^3 (three 3-wide average pools); a sequence of odd widths = the pool-sequence smoother of
        the self-configuring search (kovesi_spline, convexAdam_hyper_util.py:475-488).  gauss: a sequence of an odd number of
        filter weights = its GaussianSmoothing (hyper_util.py:454-473; `search.gaussian_weights`)."""
        fix = self._t(fix)
        mov = self._t(mov)
        for s in (p, m, v):
            if not s.is_contiguous() or s.dtype != torch.float32:
                raise TypeError("adam_run: p, m, v must be contiguous float32 (updated in place)")
            self._t(s)
        Cn, ha, wa, da = fix.shape
        gx, gy, gz = self.identity_axes(ha, wa, da, fix.device)
        u = torch.empty_like(p)
        g = torch.empty_like(p) if want_grad else None
        nb = self.dll.ca_adam_workspace_bytes(Cn, ha, wa, da)
        ws = self._ws(nb, fix)
        if gauss is not None:
            if pools is not None:
                raise ValueError("adam_run: pools and gauss exclude each other")
            wts = (C.c_float * len(gauss))(*[float(w) for w in gauss])
            self._chk(self.dll.ca_adam_run_gauss(self._p(fix), self._p(mov), Cn, ha, wa, da, lambda_weight, self._p(p), self._p(m),
                                                self._p(v), first_step, niter, self._p(gx), self._p(gy), self._p(gz),
                                                C.cast(wts, C.c_void_p), len(gauss), self._p(u), self._p(g),
                                                self._p(ws), nb, self._stream(fix)))
            return (u, g) if want_grad else u
        if pools is not None:
            widths = (C.c_int * len(pools))(*[int(w) for w in pools])
            self._chk(self.dll.ca_adam_run_pools(self._p(fix), self._p(mov), Cn, ha, wa, da, lambda_weight, self._p(p), self._p(m),
                                                self._p(v), first_step, niter, self._p(gx), self._p(gy), self._p(gz),
                                                C.cast(widths, C.c_void_p), len(pools), self._p(u), self._p(g),
                                                self._p(ws), nb, self._stream(fix)))
            return (u, g) if want_grad else u
        self._chk(self.dll.ca_adam_run(self._p(fix), self._p(mov), Cn, ha, wa, da, lambda_weight, self._p(p), self._p(m),
                                      self._p(v), first_step, niter, self._p(gx), self._p(gy), self._p(gz), self._p(u),
                                      self._p(g), self._p(ws), nb, self._stream(fix)))
        return (u, g) if want_grad else u

    def adam_window(self, fix, mov, lambda_weight, lo, hi, p, m, v, first_step, niter, ws=None, features_ready=False):
        """Adam iterations on the H planes [lo, hi) of the grid (p, m, v: (3, hi-lo, wa, da), updated in place).
        Returns (u, ws): pass `ws` back with features_ready=True to skip the feature re-layout on later calls."""
        fix = self._t(fix)
        mov = self._t(mov)
        for s in (p, m, v):
            if not s.is_contiguous() or s.dtype != torch.float32:
                raise TypeError("adam_window: p, m, v must be contiguous float32 (updated in place)")
            self._t(s)
        Cn, ha, wa, da = fix.shape
        if tuple(p.shape) != (3, hi - lo, wa, da):
            raise ValueError("adam_window: state shape does not match the window")
        gx, gy, gz = self.identity_axes(ha, wa, da, fix.device)
        u = torch.empty_like(p)
        nb = self.dll.ca_adam_window_workspace_bytes(Cn, ha, wa, da, lo, hi)
        if ws is None:
            ws = self._ws(nb, fix)
            features_ready = False
        self._chk(self.dll.ca_adam_run_window(self._p(fix), self._p(mov), Cn, ha, wa, da, lo, hi, lambda_weight, self._p(p),
                                             self._p(m), self._p(v), first_step, niter, self._p(gx), self._p(gy), self._p(gz),
                                             self._p(u), self._p(None), self._p(ws), nb, int(features_ready),
                                             self._stream(fix)))
        return u, ws

    def pack_field(self, field, round_half=False, out=None):
        field = self._t(field)
        _, H, W, D = field.shape
        if out is None:
            out = torch.empty((H, W, D, 3), dtype=torch.float64, device=field.device)
        self._chk(self.dll.ca_pack_field(self._p(field), H, W, D, int(round_half), self._p(out), self._stream(field)))
        return out


    def field_mean(self, field, mask=None, round_half=False):
        """(mean displacement per channel, voxel count) of a (3,H,W,D) f32 field over mask != 0 (uint8) or everywhere:
        a (4,) float64 device tensor."""
        field = self._t(field)
        n = field[0].numel()
        if mask is not None:
            mask = self._t(mask, torch.uint8)
            if mask.numel() != n:
                raise ValueError("field_mean: mask must have one entry per voxel")
        out = torch.empty(4, dtype=torch.float64, device=field.device)
        nb = self.dll.ca_field_mean_workspace_bytes()
        ws = self._ws(nb, field)
        self._chk(self.dll.ca_field_mean(self._p(field), self._p(mask), n, int(round_half), self._p(out), self._p(ws), nb,
                                        self._stream(field)))
        return out

    def apply_convex(self, disp, moving):
        """moving (H,W,D) f32/f64 warped by disp (H,W,D,3) f64 (apply_convex.py:13-24)."""
        disp = self._t(disp, torch.float64)
        if moving.dtype not in (torch.float32, torch.float64):
            raise TypeError("apply_convex: moving must be float32 or float64")
        moving = self._t(moving, moving.dtype)
        H, W, D = moving.shape
        if tuple(disp.shape) != (H, W, D, 3):
            raise ValueError("apply_convex: disp must be (H,W,D,3) matching moving")
        out = torch.empty_like(moving)
        self._chk(self.dll.ca_apply_convex(self._p(disp), self._p(moving), int(moving.dtype == torch.float64), H, W, D, self._p(out),
                                          self._stream(moving)))
        return out


_LIB = None


def get_lib() -> Lib:
    """The product library: in-tree CUDA build only.  Raises loudly when it (or a GPU) is missing."""
    global _LIB
    if _LIB is None:
        if not torch.cuda.is_available():
            raise NativeError("convexadam_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback")
        _LIB = Lib(LIB_PATH, expect_cuda=True)
    return _LIB
