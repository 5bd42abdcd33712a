"""ORACLE (test infrastructure, not product code) — the reference's per-image restoration metrics, restated.

Follows image_sample.py:671-680 (`evaluate_constraint`): sample -> add(1).div(2).clamp(0,1); mse = mean over (C,H,W);
psnr = 10 log10(1/mse); x_hat = 2 sample - 1; cons_orig = ||x_hat - batch_x||_1 with batch_x = 2 x_orig - 1 (:627).
Only tests/ may import this.

Parity of `restoration_metrics` is unpinned: `evaluate_constraint` needs a data loader, PNG output and FID files and
cannot be run stand-alone, so those few lines are pinned by inspection of the cited source only (they are torch
one-liners).  `ssim3d` (below) IS pinned, bit for bit, against the reference's own basicsr code."""
import torch


def restoration_metrics(sample, x_orig):
    s = sample.add(1).div(2).clamp(0, 1)
    mse = torch.mean((s - x_orig) ** 2, dim=(1, 2, 3))
    psnr = 10 * torch.log10(1 / mse)
    x_hat = 2 * s - 1.0
    batch_x = 2 * x_orig - 1.0
    cons_orig = torch.linalg.vector_norm(x_hat - batch_x, ord=1, dim=(1, 2, 3))
    return dict(mse=mse, psnr=psnr, const_orig=cons_orig, image=s)


# ------------------------------------------------------------------------------------------------ SSIM
# image_sample.py:571-582 `ssim_fn`: both images are rounded to uint8, then basicsr `calculate_ssim(..., crop_border=0,
# test_y_channel=False)` with its default ssim3d=True (basicsr/metrics/psnr_ssim.py:251-337): `_ssim_3d` (:171-208) runs
# ONE 11 x 11 x 11 Gaussian (sigma 1.5, cv2.getGaussianKernel) over the [H, W, 3] volume — the channel axis is filtered
# too — with replicate padding, in float32 (an nn.Conv3d), and averages the SSIM map over all H*W*3 entries.
# PINNED against the unmodified basicsr code run on the CPU (tests/test_oracle_vs_reference.py::test_ssim and
# tests/golden/ssim_r32.pt); the reference's own `.cuda()` calls are redirected to the CPU for that.

def gaussian_window(size=11, sigma=1.5):
    """cv2.getGaussianKernel(size, sigma) for sigma > 0: exp(-(i - (size-1)/2)^2 / (2 sigma^2)), normalised, float64."""
    i = torch.arange(size, dtype=torch.float64) - (size - 1) / 2
    g = torch.exp(-(i ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim3d(sample01, orig01):
    """Per-image SSIM of `ssim_fn`: inputs [B, 3, H, W] in [0, 1]."""
    import torch.nn.functional as F
    g = gaussian_window()
    w3 = (g[:, None, None] * g[None, :, None] * g[None, None, :]).float()[None, None]  # (H, W, C) taps
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2

    def blur(v):  # v: [H, W, 3]
        return F.conv3d(F.pad(v[None, None], (5, 5, 5, 5, 5, 5), mode="replicate"), w3)[0, 0]

    out = []
    for s, o in zip(sample01, orig01):
        a = torch.round(s * 255).to(torch.uint8).permute(1, 2, 0).double().float()
        b = torch.round(o * 255).to(torch.uint8).permute(1, 2, 0).double().float()
        mu1, mu2 = blur(a), blur(b)
        mu1_sq, mu2_sq, mu12 = mu1 ** 2, mu2 ** 2, mu1 * mu2
        s1, s2, s12 = blur(a ** 2) - mu1_sq, blur(b ** 2) - mu2_sq, blur(a * b) - mu12
        m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
        out.append(float(m.mean()))
    return torch.tensor(out, dtype=torch.float64)
