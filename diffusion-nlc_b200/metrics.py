"""Restoration metrics of a finished batch, on the device (first slice).

The reference's `evaluate_constraint` (image_sample.py:671-684) clamps the sampler output to [0,1], computes per-image
MSE / PSNR and the L1 constraint residuals on the host and re-reads PNG files for FID.  Here MSE, PSNR, the constraint
residuals (`Constraint.loss`) and the distance to the ground truth are computed from the device tensors by libnlc_b200
kernels, and their sums are all-reduced across the ranks of a sharded run (parallel.reduce_metric_sums).  `ssim_fn`
(image_sample.py:571-582: uint8 rounding + basicsr's 3-D Gaussian-window SSIM, psnr_ssim.py:171-208, which the
reference evaluates image by image through the CPU, numpy and five cuDNN conv3d calls) is one tiled kernel over the
batch.  FID (InceptionV3) is outside this slice."""
import ctypes as C

import torch

from . import _lib, parallel
from .svd_operators import _stream


def ssim_fn(sample, orig):
    """Per-image SSIM of two [B,3,H,W] batches in [0,1], as image_sample.py:571-582 computes it; a device tensor [B]."""
    dev = orig.device if orig.is_cuda else sample.device
    assert dev.type == "cuda", "ssim_fn runs on the GPU (there is no CPU path in this package)"
    s = sample.to(dev, torch.float32).contiguous()
    o = orig.to(dev, torch.float32).contiguous()
    assert s.shape == o.shape and s.dim() == 4 and s.shape[1] == 3, "ssim_fn expects [B,3,H,W]"
    B, _, H, W = s.shape
    lib = _lib.lib()
    ws = torch.empty(int(lib.nlc_ssim3d_ws(B, H, W)), dtype=torch.uint8, device=dev)
    out = torch.empty(B, device=dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    _lib.check(lib.nlc_ssim3d(_lib.ctx(idx), s.data_ptr(), o.data_ptr(), B, H, W, ws.data_ptr(), out.data_ptr(),
                              _stream()))
    return out


def restoration_metrics(sample, x_orig, constraint=None, y=None, return_image=False, ssim=False):
    """sample: sampler output [B,C,H,W] in [-1,1] coordinates (device or CPU); x_orig: ground truth in [0,1].
    Returns a dict of per-sample device tensors: mse, psnr (image_sample.py:674-675), const_orig (:680), and with a
    Constraint_Function + measurement y also const_f / const_b (:679); `image` = clamp((sample+1)/2, 0, 1) on request."""
    dev = x_orig.device if x_orig.is_cuda else sample.device
    assert dev.type == "cuda", "restoration_metrics runs on the GPU (there is no CPU path in this package)"
    s = sample.to(dev, torch.float32).contiguous()
    o = x_orig.to(dev, torch.float32).contiguous()
    B, n = s.shape[0], s[0].numel()
    mse, l1 = torch.empty(B, device=dev), torch.empty(B, device=dev)
    img = torch.empty_like(s) if return_image or constraint is not None or ssim else None
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    _lib.check(_lib.lib().nlc_image_metrics(_lib.ctx(idx), s.data_ptr(), o.data_ptr(), B, n,
                                            C.c_void_p(img.data_ptr()) if img is not None else None, mse.data_ptr(),
                                            l1.data_ptr(), _stream()))
    out = dict(mse=mse, psnr=10 * torch.log10(1 / mse), const_orig=l1)
    if constraint is not None:
        x_hat = 2 * img - 1.0  # image_sample.py:677
        f, b = constraint.loss(x_hat, y)
        out["const_f"], out["const_b"] = f.to(dev), b.to(dev)
    if ssim:  # image_sample.py:676
        out["ssim"] = ssim_fn(img, o)
    if return_image:
        out["image"] = img
    return out


def reduce_means(metrics, keys=("mse", "psnr", "ssim", "const_orig", "const_f", "const_b")):
    """Global means over all ranks' samples: one all-reduce of (sums..., count)."""
    present = [k for k in keys if k in metrics]
    sums = torch.stack([metrics[k].double().sum() for k in present] +
                       [torch.tensor(float(metrics[present[0]].numel()), dtype=torch.float64,
                                     device=metrics[present[0]].device)])
    sums = parallel.reduce_metric_sums(sums)
    return {k: (sums[i] / sums[-1]).item() for i, k in enumerate(present)}
