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


def test_ssim_against_reference_golden(golden_dir):
    """nlc_ssim3d against the reference's own ssim_fn values (tests/golden/ssim.pt: image_sample.py:571-582 -> basicsr
    `_ssim_3d`).  The reference filters in fp32 with an 11^3 window, the kernel separably: 1e-4 absolute on an index in
    [0, 1]; identical images give exactly 1."""
    import os
    from nlc_b200 import metrics as M
    g = torch.load(os.path.join(golden_dir, "ssim.pt"), weights_only=True)
    for size in g.values():
        for name, case in size.items():
            got = M.ssim_fn(case["sample"].to(dev), case["orig"].to(dev)).cpu().double()
            assert (got - case["ssim"]).abs().max() < 1e-4, (name, got, case["ssim"])
            if name == "same":
                assert torch.equal(got, torch.ones_like(got))


def test_ssim_at_256_against_the_oracle():
    from nlc_b200 import metrics as M
    gen = torch.Generator().manual_seed(43)
    B, H = 4, 256
    orig = torch.nn.functional.avg_pool2d(torch.rand(B, 3, H + 6, H + 6, generator=gen), 7, 1)
    sample = (orig + 0.04 * torch.randn(B, 3, H, H, generator=gen)).clamp(0, 1)
    want = OM.ssim3d(sample, orig)
    got = M.ssim_fn(sample.to(dev), orig.to(dev)).cpu().double()
    assert (got - want).abs().max() < 1e-4, (got, want)
    full = M.restoration_metrics((2 * sample - 1).to(dev), orig.to(dev), ssim=True)
    assert (full["ssim"].cpu().double() - want).abs().max() < 1e-4
    assert abs(M.reduce_means(full)["ssim"] - want.mean().item()) < 1e-4
