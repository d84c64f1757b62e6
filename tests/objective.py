"""The Adam stage's objective (reference convex_adam_MIND.py:166-177: sampled-feature SSD + diffusion regulariser),
re-evaluated in float64 on the CPU for a given field.  Test infrastructure: the parity ladder's top rung (SURVEY 7.2
P3) compares the objective the CUDA path reaches after N free-running iterations with the reference's."""
import numpy as np
import torch
import torch.nn.functional as F


def registration_objective(u, fix_a, mov_a, lambda_weight):
    """u: (3,ha,wa,da) displacement on the Adam grid in Adam-grid voxels (`disp_sample`, channel c = array axis c);
    fix_a, mov_a: (C,ha,wa,da) features on that grid.  Returns data term + regulariser as a Python float."""
    u = torch.as_tensor(np.asarray(u), dtype=torch.float64)
    fa = torch.as_tensor(np.asarray(fix_a), dtype=torch.float64)
    ma = torch.as_tensor(np.asarray(mov_a), dtype=torch.float64)
    _, ha, wa, da = u.shape
    reg = sum(((u.narrow(ax, 1, n - 1) - u.narrow(ax, 0, n - 1)) ** 2).mean() for ax, n in ((1, ha), (2, wa), (3, da)))
    ident = F.affine_grid(torch.eye(3, 4, dtype=torch.float64)[None], (1, 1, ha, wa, da), align_corners=False)
    half_extent = torch.tensor([(ha - 1) / 2, (wa - 1) / 2, (da - 1) / 2], dtype=torch.float64)
    # normalised by (size-1)/2 but sampled with align_corners=False; grid's last axis is (x=D, y=W, z=H)
    offs = (u.permute(1, 2, 3, 0) / half_extent).flip(-1)
    warped = F.grid_sample(ma[None], ident + offs[None], align_corners=False, mode="bilinear")[0]
    data = ((warped - fa) ** 2).mean(0).mul(12).mean()
    return float(data + lambda_weight * reg)
