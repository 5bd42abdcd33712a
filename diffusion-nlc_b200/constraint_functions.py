"""Host-side mirror of the reference's constraint factory and projection glue.

`svd_constraint` keeps the signature of src/constraint_functions.py:206-294 and returns the nlc_b200 operator for
the task name; `Constraint_Function`, `affine_svd` and `get_constraint_function` mirror image_sample.py:282-342,
376-379 and 359-405 for the `--constraint_proj svd` path (the reference's own branch raises UnboundLocalError,
SURVEY §0.6; this implements what it intends: functions/svd_operators.py semantics with the zero-guarded
pseudo-inverse).  Mask files are absent from the reference tree, so `inpainting_box` (a centred zero box, the
index convention of src/constraint_functions.py:220-240) is provided next to the file-based names.
"""
import os
from functools import partial

import numpy as np
import torch

from . import svd_operators as ops_svd


def _bicubic_kernel(factor):
    # src/constraint_functions.py:252-269
    def cubic(x, a=-0.5):
        ax = abs(x)
        if ax <= 1:
            return (a + 2) * ax ** 3 - (a + 3) * ax ** 2 + 1
        if 1 < ax < 2:
            return a * ax ** 3 - 5 * a * ax ** 2 + 8 * a * ax - 4 * a
        return 0

    k = np.zeros(factor * 4)
    for i in range(factor * 4):
        k[i] = cubic((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
    k = torch.from_numpy(k / np.sum(k)).float()
    return k / k.sum()


def _gauss_kernel(taps, sigma):
    pdf = lambda x: torch.exp(torch.Tensor([-0.5 * (x / sigma) ** 2]))
    k = torch.Tensor([pdf(i) for i in range(-(taps // 2), taps // 2 + 1)])
    return k / k.sum()


def _missing_from_mask(mask_flat):
    r = torch.nonzero(mask_flat == 0).long().reshape(-1) * 3
    return torch.cat([r, r + 1, r + 2], dim=0)


def svd_constraint(fn, fn_scale=4, device="cuda", base_mask_dir="store/inp_masks", image_size=256, channels=3,
                   perm=None, mask=None):
    """Task name -> operator (src/constraint_functions.py:206-294).  `perm` (WH-CS) and `mask` (inpainting) let the
    caller supply what the reference draws / loads; defaults reproduce the reference (torch.randperm on the
    current generator; files under base_mask_dir)."""
    if fn == "cs_walshhadamard":
        if perm is None:
            perm = torch.randperm(image_size ** 2)
        return ops_svd.WalshHadamardCS(channels, image_size, round(fn_scale), perm, device)
    if "inpainting" in fn:
        if mask is not None:
            missing = _missing_from_mask(torch.as_tensor(mask).reshape(-1))
        elif fn == "inpainting_box":
            m = torch.ones(image_size, image_size)
            q = image_size // 4
            m[q:3 * q, q:3 * q] = 0
            missing = _missing_from_mask(m.reshape(-1))
        elif fn == "inpainting_random":
            r = torch.randperm(image_size ** 2)[:image_size ** 2 // 2].long() * 3
            missing = torch.cat([r, r + 1, r + 2], dim=0)
        else:
            name = {"inpainting_lolcat": "inp_masks/lolcat_extra.npy", "inpainting_lorem": "inp_masks/lorem3.npy",
                    "inpainting_ddnm": os.path.join(base_mask_dir, "mask.npy"),
                    "inpainting_half": os.path.join(base_mask_dir, "mask_half.npy")}.get(fn)
            if name is not None:
                missing = _missing_from_mask(torch.from_numpy(np.load(name)).reshape(-1))
            else:
                r = torch.load(os.path.join(base_mask_dir, "mask_random.pt"))
                missing = torch.cat([r, r + 1, r + 2], dim=0)
        return ops_svd.Inpainting(channels, image_size, missing, device)
    if fn == "colorization":
        return ops_svd.Colorization(image_size, device)
    if fn == "sr_averagepooling":
        return ops_svd.SuperResolution(channels, image_size, int(fn_scale), device)
    if fn == "sr_bicubic":
        factor = int(fn_scale)
        k = _bicubic_kernel(factor)
        return ops_svd.SRConv(k / k.sum(), channels, image_size, device, stride=factor)
    if fn == "deblur_uni":
        return ops_svd.Deblurring(torch.Tensor([1 / 9] * 9), channels, image_size, device)
    if fn == "deblur_gauss":
        return ops_svd.Deblurring(_gauss_kernel(5, 10), channels, image_size, device)
    if fn == "deblur_aniso":  # src/constraint_functions.py:280-292: 9 taps, sigma 1 along rows, sigma 20 along columns
        k1, k2 = _gauss_kernel(9, 1), _gauss_kernel(9, 20)
        return ops_svd.Deblurring2D(k1, k2, channels, image_size, device)
    if fn == "cs_blockbased":  # src/constraint_functions.py:212-215
        return ops_svd.CS(channels, image_size, fn_scale, device)
    if fn == "denoising":  # :242-244
        return ops_svd.Denoising(channels, image_size, device)
    return None


def affine_svd(x0_t, y, lambda_t=None, A_funcs=None):
    """x0 - A^+(A x0 - y) (image_sample.py:376-379); `lambda_t` is ignored there too."""
    return A_funcs.project(x0_t, y)


class Constraint_Function:
    """image_sample.py:282-342 for proj='svd'."""

    def __init__(self, deg, A_funcs, constraint_fn, proj="svd", channels=3, image_size=256, lr=1.0):
        self.deg, self.A_funcs = deg, A_funcs
        self.A, self.Ap = A_funcs.A, A_funcs.A_pinv
        self.constraint_fn = constraint_fn
        self.proj, self.channels, self.image_size, self.lr = proj, channels, image_size, lr
        self._xhat_cache = None

    def transform(self, x):
        return self.A(x.reshape(x.shape[0], -1))

    def inv_transform(self, y):
        b = y.shape[0]
        shape = (b, self.channels, self.image_size, self.image_size)
        Apy = self.Ap(y).view(shape)
        if self.deg[:6] == "deblur":
            Apy = y.view(shape)
        elif self.deg == "colorization":
            Apy = y.view(b, 1, self.image_size, self.image_size).repeat(1, 3, 1, 1)
        elif self.deg == "inpainting":
            Apy = Apy + self.Ap(self.A(torch.ones_like(Apy).reshape(b, -1))).reshape(shape) - 1
        return Apy

    def loss(self, x, y):
        """(||A x - y||_1, ||inv_transform(y) - x||_1) per sample, on the CPU like the reference."""
        fwd, bwd = self.loss_device(x, y)
        return fwd.cpu(), bwd.cpu()

    def loss_device(self, x, y):
        """The same two per-sample losses left on the device: no host read, so a loop that tracks its best x0 with
        nlc_best_update can be captured in a CUDA graph (ExperimentDiffusion.denoise_loop, `constrain_loss_device`)."""
        y_hat = self.transform(x)
        # y is fixed for a whole batch: A^+ y is cached against the tensor OBJECT (held, so its address cannot be reused
        # by another measurement) and its in-place version counter (a refilled buffer invalidates the entry)
        c = self._xhat_cache
        if c is None or c[0] is not y or c[1] != y._version:
            c = self._xhat_cache = (y, y._version, self.inv_transform(y))
        x_hat = c[2]
        fwd = ops_svd.l1_diff_rows(y_hat, y.reshape(y.shape[0], -1))
        bwd = ops_svd.l1_diff_rows(x_hat, x)
        return fwd, bwd


def get_constraint_function(constraint, constraint_scale=4.0, device="cuda", image_size=256, channels=3,
                            constraint_lr=10, base_mask_dir="store/inp_masks", perm=None, mask=None):
    """image_sample.py:359-405, 'svd' branch, with explicit arguments instead of the argparse namespace."""
    A_funcs = svd_constraint(constraint, fn_scale=constraint_scale, device=device, base_mask_dir=base_mask_dir,
                             image_size=image_size, channels=channels, perm=perm, mask=mask)
    if A_funcs is None:
        raise ValueError("unknown constraint %r" % constraint)
    fn = partial(affine_svd, A_funcs=A_funcs)
    return Constraint_Function(constraint, A_funcs, fn, proj="svd", channels=channels, image_size=image_size,
                               lr=constraint_lr)
