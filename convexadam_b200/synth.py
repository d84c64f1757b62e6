"""Seeded synthetic volume pairs: CT-like smooth anatomy + noise, moving = fixed warped by a
smooth random field.  CPU torch generator => identical inputs in this container and on the GPU box."""
import torch
import torch.nn.functional as F


def make_pair(shape, seed=0, disp_scale=0.04, noise=15.0):
    """Returns (fixed, moving) float32 CPU tensors of `shape` = (H, W, D)."""
    H, W, D = shape
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(1, 1, H // 8 + 2, W // 8 + 2, D // 8 + 2, generator=g)
    fixed = F.interpolate(low, size=(H, W, D), mode="trilinear", align_corners=True)[0, 0] * 300.0
    fixed = fixed + noise * torch.randn(H, W, D, generator=g)
    fld = F.interpolate(torch.randn(1, 3, 4, 4, 4, generator=g) * disp_scale, size=(H, W, D), mode="trilinear",
                        align_corners=True)
    ident = F.affine_grid(torch.eye(3, 4).unsqueeze(0), (1, 1, H, W, D), align_corners=True)
    moving = F.grid_sample(fixed[None, None], ident + fld.permute(0, 2, 3, 4, 1), align_corners=True,
                           padding_mode="border")[0, 0]
    moving = moving + 0.5 * noise * torch.randn(H, W, D, generator=g)
    return fixed.contiguous().float(), moving.contiguous().float()


def make_label_pair(shape, n_labels=16, seed=0):
    """Two label maps (float tensors holding integers 0..n_labels-1), all labels present (config 4)."""
    H, W, D = shape
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(1, n_labels, H // 8 + 2, W // 8 + 2, D // 8 + 2, generator=g)
    fix = F.interpolate(low, size=(H, W, D), mode="trilinear", align_corners=True)[0].argmax(0)
    low2 = low + 0.35 * torch.randn(low.shape, generator=g)
    mov = F.interpolate(low2, size=(H, W, D), mode="trilinear", align_corners=True)[0].argmax(0)
    return fix.float(), mov.float()
