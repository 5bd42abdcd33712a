"""ORACLE (test infrastructure, not product code) — the reference's per-image restoration metrics, restated.

Follows image_sample.py:671-680 (`evaluate_constraint`): sample -> add(1).div(2).clamp(0,1); mse = mean over (C,H,W);
psnr = 10 log10(1/mse); x_hat = 2 sample - 1; cons_orig = ||x_hat - batch_x||_1 with batch_x = 2 x_orig - 1 (:627).
Only tests/ may import this.

Parity unpinned: `evaluate_constraint` needs a data loader, PNG output and FID files and cannot be run stand-alone, so
these few lines are pinned by inspection of the cited source only (they are torch one-liners)."""
import torch


def restoration_metrics(sample, x_orig):
    s = sample.add(1).div(2).clamp(0, 1)
    mse = torch.mean((s - x_orig) ** 2, dim=(1, 2, 3))
    psnr = 10 * torch.log10(1 / mse)
    x_hat = 2 * s - 1.0
    batch_x = 2 * x_orig - 1.0
    cons_orig = torch.linalg.vector_norm(x_hat - batch_x, ord=1, dim=(1, 2, 3))
    return dict(mse=mse, psnr=psnr, const_orig=cons_orig, image=s)
