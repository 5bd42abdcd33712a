"""Device-side restoration metrics (first slice) against the oracle restatement of
image_sample.py:671-680.  fp32 sums over up to 196 608 elements in a different order than torch's: 1e-5 relative; the
clamped image itself is bit-exact."""
import pytest
import torch

from oracle import metrics as OM

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (2, 3, 256, 256), (5, 3, 17, 9)])
def test_restoration_metrics_match_the_reference_formulas(shape):
    from nlc_b200 import constraint_functions as CF, metrics as M
    g = torch.Generator().manual_seed(41)
    x_orig = torch.rand(shape, generator=g)
    sample = (2 * x_orig - 1) + 0.3 * torch.randn(shape, generator=g)  # leaves [-1,1] in places: the clamp matters
    ref = OM.restoration_metrics(sample, x_orig)
    got = M.restoration_metrics(sample.to(dev), x_orig.to(dev), return_image=True)
    assert torch.equal(got["image"].cpu(), ref["image"])
    for k in ("mse", "psnr", "const_orig"):
        assert ((got[k].cpu() - ref[k]).abs() / ref[k].abs()).max() < 1e-5, k
    if shape[2] == shape[3] and shape[2] >= 64:
        con = CF.get_constraint_function("sr_averagepooling", constraint_scale=4.0, device=dev, image_size=shape[2])
        y = con.transform((2 * x_orig - 1).to(dev))
        full = M.restoration_metrics(sample.to(dev), x_orig.to(dev), constraint=con, y=y)
        f, b = con.loss((2 * ref["image"] - 1).to(dev), y)
        assert torch.allclose(full["const_f"].cpu(), f.cpu(), rtol=1e-6) and torch.allclose(full["const_b"].cpu(), b.cpu(), rtol=1e-6)
        means = M.reduce_means(full)
        assert abs(means["psnr"] - ref["psnr"].double().mean().item()) < 1e-4
