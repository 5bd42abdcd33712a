"""GPU parity of the DDNM operators (rows P0-P6): bit-level agreement with the reference's golden outputs at R=32
(tolerance 2e-6 absolute: fp32, FMA contraction only) and size-independent properties at the benchmark size
R=256 (A A^+ y = y, projection feasibility |A x^ - y|, idempotence)."""
import os

import pytest
import torch

from oracle import operators as O

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _build(name, R, C, missing=None, perm=None):
    from nlc_b200 import svd_operators as P
    if name == "inpainting":
        return P.Inpainting(C, R, missing, dev)
    if name == "colorization":
        return P.Colorization(R, dev)
    if name == "sr_averagepooling":
        return P.SuperResolution(C, R, 4, dev)
    if name == "cs_walshhadamard":
        return P.WalshHadamardCS(C, R, 4, perm, dev)
    if name == "sr_bicubic":
        return P.SRConv(O.bicubic_kernel(4), C, R, dev, stride=4)
    if name == "deblur_aniso":
        return P.Deblurring2D(*O.aniso_kernels(), C, R, dev)
    return P.Deblurring(O.gauss_kernel(), C, R, dev)


NAMES = ["inpainting", "colorization", "sr_averagepooling", "cs_walshhadamard", "sr_bicubic", "deblur_gauss",
         "deblur_aniso"]


@pytest.mark.parametrize("name", NAMES)
def test_against_reference_golden(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, "operators2_r32.pt" if name == "deblur_aniso" else "operators_r32.pt"),
                   weights_only=True)
    op = _build(name, 32, 3, g.get("missing"), g.get("perm"))
    y = op.A(g["x"].to(dev))
    tol = 2e-6 if name in ("inpainting", "colorization", "sr_averagepooling", "cs_walshhadamard") else 2e-4
    assert (y.cpu() - g[name]["A"]).abs().max() < tol
    assert (op.At(g[name]["A"].to(dev)).cpu() - g[name]["At"]).abs().max() < tol
    assert (op.A_pinv(g[name]["A"].to(dev)).cpu() - g[name]["A_pinv"]).abs().max() < tol * 50
    assert (op.project(g["x0"].to(dev), g[name]["A"].to(dev)).cpu() - g[name]["project"]).abs().max() < tol * 50


@pytest.mark.parametrize("name", NAMES)
def test_properties_at_256(name):
    R, C, B = 256, 3, 4
    gen = torch.Generator().manual_seed(9)
    mask = torch.ones(R, R)
    mask[64:192, 64:192] = 0  # centred 128 x 128 box
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(R * R, generator=gen)
    op = _build(name, R, C, missing, perm)
    x = (torch.rand(B, C * R * R, generator=gen) * 2 - 1).to(dev)
    x0 = torch.randn(B, C, R, R, generator=gen).to(dev)
    y = op.A(x)
    assert y.shape == (B, op.ydim)
    assert (op.A(op.A_pinv(y)) - y).abs().max() < 2e-5          # A A^+ y = y
    p = op.project(x0, y)
    assert (op.A(p) - y).abs().max() < 5e-5                     # feasibility after one projection
    assert (op.project(p, y) - p).abs().max() < 5e-4            # idempotence
    # adjoint identity <A x, y> = <x, At y>
    lhs = (op.A(x) * y).sum(dim=1)
    rhs = (x * op.At(y)).sum(dim=1)
    assert ((lhs - rhs).abs() / lhs.abs().clamp_min(1.0)).max() < 1e-3


def test_constraint_function_glue():
    from nlc_b200 import constraint_functions as CF
    R, B = 64, 2
    gen = torch.Generator().manual_seed(10)
    x = (torch.rand(B, 3, R, R, generator=gen) * 2 - 1).to(dev)
    for task in ("sr_averagepooling", "colorization", "inpainting_box", "deblur_gauss", "sr_bicubic",
                 "cs_walshhadamard"):
        con = CF.get_constraint_function(task, constraint_scale=4.0, device=dev, image_size=R,
                                         perm=torch.randperm(R * R, generator=gen))
        y = con.transform(x)
        x0 = torch.randn(B, 3, R, R, generator=gen).to(dev)
        p = con.constraint_fn(x0, y)
        assert p.shape == x0.shape
        fwd, bwd = con.loss(p.clamp(-1, 1), y)
        assert fwd.device.type == "cpu" and fwd.shape == (B,) and bwd.shape == (B,)
        f0, _ = con.loss(x, y)
        assert f0.max() < 1e-2  # the ground truth satisfies its own measurement
    with pytest.raises(NotImplementedError):
        CF.svd_constraint("cs_blockbased")
