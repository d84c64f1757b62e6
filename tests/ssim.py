"""3-D structural similarity with an 11^3 Gaussian window (sigma 1.5), the metric of the reference's fixture test
(tests/helper_functions.py:102-146 there; the constants C1 = 0.01^2, C2 = 0.03^2 and zero padding are part of the metric).
Test infrastructure."""
import torch
import torch.nn.functional as F


def ssim3d(a, b, window_size=11, sigma=1.5):
    """a, b: (D,H,W) arrays/tensors -> mean SSIM (float)."""
    a = torch.as_tensor(a, dtype=torch.float32)[None, None]
    b = torch.as_tensor(b, dtype=torch.float32)[None, None]
    x = torch.arange(window_size, dtype=torch.float32) - window_size // 2
    g = torch.exp(-x ** 2 / (2 * sigma ** 2))
    g = g / g.sum()
    win = (g[:, None, None] * g[None, :, None] * g[None, None, :])[None, None]
    pad = window_size // 2
    mu_a, mu_b = F.conv3d(a, win, padding=pad), F.conv3d(b, win, padding=pad)
    var_a = F.conv3d(a * a, win, padding=pad) - mu_a ** 2
    var_b = F.conv3d(b * b, win, padding=pad) - mu_b ** 2
    cov = F.conv3d(a * b, win, padding=pad) - mu_a * mu_b
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return float((((2 * mu_a * mu_b + c1) * (2 * cov + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (var_a + var_b + c2))).mean())
