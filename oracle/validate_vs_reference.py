"""Checks every function of the CPU oracle against the UNMODIFIED reference run in this container.

TEST INFRASTRUCTURE.  Only runnable where /root/reference exists (not on the GPU box).  The committed
golden vectors (tests/golden/) carry the same evidence to places where the reference is absent.
Usage: python oracle/validate_vs_reference.py
"""
import sys
import os
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ca_oracle as orc  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

ref, ref_mind = import_reference()
torch.set_num_threads(1)
cpu = torch.device("cpu")
ok = True


def report(name, a, b, tol=0.0):
    global ok
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype.kind in "iu":
        bad = int((a != b).sum()); good = bad == 0
        print(f"{'PASS' if good else 'FAIL'} {name}: int mismatches {bad}/{a.size}")
    else:
        diff = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
        nbit = int((a.view(np.uint32) != b.view(np.uint32)).sum()) if a.dtype == np.float32 and b.dtype == np.float32 else -1
        good = diff <= tol
        print(f"{'PASS' if good else 'FAIL'} {name}: max|diff| {diff:.3e} (tol {tol:g}), bit-mismatches {nbit}/{a.size}")
    ok &= good


g = torch.Generator().manual_seed(1)

# --- avg_pool3d variants -------------------------------------------------------------------------
x = torch.randn(3, 13, 11, 17, generator=g)
for (k, s, p) in [(3, 1, 1), (2, 2, 0), (6, 6, 0), (5, 1, 2), (4, 1, 2), (3, 3, 0)]:
    report(f"avgpool k{k}s{s}p{p}", orc.avgpool3d(x.numpy(), k, s, p), F.avg_pool3d(x[None], k, stride=s, padding=p)[0].numpy())

# --- mesh / identity axes -------------------------------------------------------------------------
for hw in range(1, 9):
    S = 2 * hw + 1
    m = F.affine_grid(hw * torch.eye(3, 4).unsqueeze(0), (1, 1, S, S, S), align_corners=True).permute(0, 4, 1, 2, 3).reshape(3, -1)
    report(f"disp_mesh hw={hw}", orc.disp_mesh(hw), m.numpy())
for (h, w, d) in [(5, 7, 9), (16, 14, 18), (96, 80, 128), (32, 32, 32), (2, 3, 4)]:
    idg = F.affine_grid(torch.eye(3, 4).unsqueeze(0), (1, 1, h, w, d), align_corners=False)[0]
    gx, gy, gz = orc.identity_axes(h, w, d)
    report(f"identity axes {(h,w,d)} x", gx, idg[0, 0, :, 0].numpy())
    report(f"identity axes {(h,w,d)} y", gy, idg[0, :, 0, 1].numpy())
    report(f"identity axes {(h,w,d)} z", gz, idg[:, 0, 0, 2].numpy())
    full = np.stack(np.broadcast_arrays(gx[None, None, :], gy[None, :, None], gz[:, None, None]), -1)
    report(f"identity full {(h,w,d)}", full.astype(np.float32), idg.numpy())

# --- correlate ------------------------------------------------------------------------------------
def check_correlate(C, h, w, d, hw, gs=2):
    fix = torch.rand(1, C, h, w, d, generator=g)
    mov = torch.rand(1, C, h, w, d, generator=g)
    ssd_r, am_r = ref.correlate(fix, mov, hw, gs, (h * gs, w * gs, d * gs), C)
    S = 2 * hw + 1
    n = h * S * S * w * d
    tail = n % 32
    for t in sorted({0, tail}):
        ssd_o, am_o = orc.correlate(fix[0].numpy(), mov[0].numpy(), hw, tail=t)
        nb = int((ssd_o.view(np.uint32) != ssd_r.numpy().view(np.uint32)).sum())
        print(f"   correlate C{C} {(h,w,d)} hw{hw}: n%64={n%64} n%16={n%16} tail={t}: ssd bit mismatches {nb}, argmin mismatches {int((am_o != am_r.numpy()).sum())}")
    return fix, mov, ssd_r, am_r

check_correlate(12, 6, 8, 8, 2)      # n = 6*25*64  multiple of 64
check_correlate(12, 5, 7, 9, 2)      # ragged
check_correlate(16, 4, 5, 6, 1)
check_correlate(3, 5, 6, 7, 3)
sys.stdout.flush()

# --- coupled_convex -------------------------------------------------------------------------------
C, h, w, d, hw = 12, 8, 8, 8, 3
fix = torch.rand(1, C, h, w, d, generator=g); mov = torch.rand(1, C, h, w, d, generator=g)
ssd_r, am_r = ref.correlate(fix, mov, hw, 2, (2 * h, 2 * w, 2 * d), C)
S = 2 * hw + 1
mesh = F.affine_grid(hw * torch.eye(3, 4).unsqueeze(0), (1, 1, S, S, S), align_corners=True).permute(0, 4, 1, 2, 3).reshape(3, -1, 1)
soft_r = ref.coupled_convex(ssd_r, am_r, mesh, 2, (2 * h, 2 * w, 2 * d))
ssd_o, am_o = orc.correlate(fix[0].numpy(), mov[0].numpy(), hw, tail=0)
report("correlate ssd (8^3,hw3)", ssd_o, ssd_r.numpy())
report("correlate argmin", am_o, am_r.numpy())
soft_o = orc.coupled_convex(ssd_r.numpy(), am_r.numpy(), mesh.numpy().reshape(3, -1))
report("coupled_convex soft", soft_o, soft_r[0].numpy())

# --- grid_sample / inverse_consistency -------------------------------------------------------------
inp = torch.randn(1, 3, 7, 9, 11, generator=g)
grid = torch.rand(1, 6, 5, 8, 3, generator=g) * 2.4 - 1.2
report("grid_sample3d", orc.grid_sample3d(inp[0].numpy(), grid[0].numpy()), F.grid_sample(inp, grid, align_corners=False)[0].numpy())
f1 = torch.randn(1, 3, 8, 10, 12, generator=g) * 0.1
f2 = torch.randn(1, 3, 8, 10, 12, generator=g) * 0.1
r1, r2 = ref.inverse_consistency(f1, f2, iter=15)
o1, o2 = orc.inverse_consistency(f1[0].numpy(), f2[0].numpy(), 15)
report("inverse_consistency d1", o1, r1[0].numpy())
report("inverse_consistency d2", o2, r2[0].numpy())

# --- trilinear --------------------------------------------------------------------------------------
x = torch.randn(1, 3, 8, 7, 9, generator=g)
for size in [(16, 14, 18), (17, 15, 19), (4, 3, 4), (8, 7, 9), (48, 42, 54)]:
    report(f"trilinear -> {size}", orc.trilinear(x[0].numpy(), size), F.interpolate(x, size=size, mode="trilinear", align_corners=False)[0].numpy(), tol=4e-6)

# --- MINDSSC -----------------------------------------------------------------------------------------
img = torch.randn(20, 18, 22, generator=g) * 100
for (r, dl) in [(1, 2), (2, 2), (1, 1)]:
    m_r = ref.MINDSSC(img[None, None], r, dl, device=cpu)[0].numpy()
    report(f"MINDSSC r{r} d{dl}", orc.mindssc(img.numpy(), r, dl), m_r, tol=3e-7)

# --- Adam iteration (autograd of the reference formulation, MIND.py:163-179) --------------------------
def ref_adam(p0, pf, pm, lam, niter):
    ha, wa, da = pf.shape[-3:]
    net = torch.nn.Sequential(torch.nn.Conv3d(3, 1, (ha, wa, da), bias=False))
    net[0].weight.data[:] = p0
    opt = torch.optim.Adam(net.parameters(), lr=1)
    grid0 = F.affine_grid(torch.eye(3, 4).unsqueeze(0), (1, 1, ha, wa, da), align_corners=False)
    hist = []
    for it in range(niter):
        opt.zero_grad()
        ds = F.avg_pool3d(F.avg_pool3d(F.avg_pool3d(net[0].weight, 3, stride=1, padding=1), 3, stride=1, padding=1), 3, stride=1, padding=1).permute(0, 2, 3, 4, 1)
        reg = lam * ((ds[0, :, 1:, :] - ds[0, :, :-1, :]) ** 2).mean() + lam * ((ds[0, 1:, :, :] - ds[0, :-1, :, :]) ** 2).mean() + lam * ((ds[0, :, :, 1:] - ds[0, :, :, :-1]) ** 2).mean()
        scale = torch.tensor([(ha - 1) / 2, (wa - 1) / 2, (da - 1) / 2]).unsqueeze(0)
        gd = grid0.view(-1, 3).float() + ((ds.view(-1, 3)) / scale).flip(1).float()
        smp = F.grid_sample(pm.float(), gd.view(1, ha, wa, da, 3), align_corners=False, mode="bilinear")
        cost = (smp - pf).pow(2).mean(1) * 12
        loss = cost.mean()
        (loss + reg).backward()
        stt = opt.state.get(net[0].weight, {})
        hist.append(dict(p=net[0].weight.data.clone(), g=net[0].weight.grad.clone(), u=ds.detach().permute(0, 4, 1, 2, 3).clone(),
                         m=stt["exp_avg"].clone() if stt else torch.zeros_like(p0), v=stt["exp_avg_sq"].clone() if stt else torch.zeros_like(p0)))
        opt.step()
        hist[-1]["p_next"] = net[0].weight.data.clone()
    return hist

for Cn in (12, 16):
    ha, wa, da = 10, 12, 14
    pf = torch.rand(1, Cn, ha, wa, da, generator=g); pm = torch.rand(1, Cn, ha, wa, da, generator=g)
    p0 = torch.randn(1, 3, ha, wa, da, generator=g) * 1.5
    hist = ref_adam(p0, pf, pm, 1.25, 6)
    st = orc.AdamState(p0[0].numpy())
    for it, hrec in enumerate(hist):
        # teacher forcing: start every step from the reference's (p, m, v)
        st.p[:] = hrec["p"][0].numpy(); st.m[:] = hrec["m"][0].numpy(); st.v[:] = hrec["v"][0].numpy(); st.step = it
        u, gp = orc.adam_step(st, pf[0].numpy(), pm[0].numpy(), 1.25, want_grad=True)
        gr = hrec["g"][0].numpy()
        report(f"adam C{Cn} it{it} u (teacher-forced)", u, hrec["u"][0].numpy())
        rel = float(np.abs(gp - gr).max() / np.abs(gr).max())
        print(f"   grad rel-to-max err {rel:.2e}  (max|g| {np.abs(gr).max():.2e})")
        ok &= rel < 5e-6
        # |dp| <= |dg| / eps (Adam, lr=1, eps=1e-8) -- the reference's own sensitivity; see DESIGN.md
        report(f"adam C{Cn} it{it} p_next (teacher-forced)", st.p, hrec["p_next"][0].numpy(), tol=2e-4)
    print()

print("ALL OK" if ok else "SOME CHECKS FAILED")
sys.exit(0 if ok else 1)
