"""GPU parity of the DDNM operators (rows P0-P6): bit-level agreement with the reference's golden outputs at R=32
(tolerance 2e-6 absolute: fp32, FMA contraction only) and size-independent properties at the benchmark size
R=256 (A A^+ y = y, projection feasibility |A x^ - y|, idempotence)."""
import math
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
                 "cs_walshhadamard", "cs_blockbased", "denoising"):
        scale = 0.25 if task == "cs_blockbased" else 4.0  # block CS takes the kept fraction (cs_ratio)
        con = CF.get_constraint_function(task, constraint_scale=scale, device=dev, image_size=R,
                                         perm=torch.randperm(R * R, generator=gen))
        y = con.transform(x)
        x0 = torch.randn(B, 3, R, R, generator=gen).to(dev)
        p = con.constraint_fn(x0, y)
        assert p.shape == x0.shape
        fwd, bwd = con.loss(p.clamp(-1, 1), y)
        assert fwd.device.type == "cpu" and fwd.shape == (B,) and bwd.shape == (B,)
        f0, _ = con.loss(x, y)
        assert f0.max() < 1e-2  # the ground truth satisfies its own measurement
    assert CF.svd_constraint("no_such_task") is None


def test_block_cs_and_general_a(golden_dir):
    """functions/svd_operators.py CS (:101-160) and GeneralA (:173-208) against the reference's golden outputs
    (tests/golden/operators3.pt), then CS at 256 x 256 through properties."""
    from nlc_b200 import svd_operators as P
    g = torch.load(os.path.join(golden_dir, "operators3.pt"), weights_only=True)
    V = O.hadamard_basis(1024, 7)
    for key, op in (("cs", P.CS(3, 64, 0.25, dev, V_small=V)), ("general", P.GeneralA(g["general"]["Amat"], dev))):
        c = g[key]
        assert op.ydim == c["A"].shape[1]
        for got, want in ((op.A(c["x"].to(dev)), c["A"]), (op.At(c["A"].to(dev)), c["At"]),
                          (op.A_pinv(c["A"].to(dev)), c["A_pinv"]), (op.A_pinv_eta(c["A"].to(dev), 0.1), c["A_pinv_eta"]),
                          (op.project(c["x0"].to(dev), c["A"].to(dev)), c["project"])):
            assert (got.cpu().reshape(want.shape) - want).abs().max() <= 1e-5 * want.abs().max(), key
        with pytest.raises(NotImplementedError):
            op.Lambda(c["x"].to(dev), 0.9, 0.1, 0.3, 0.85)
        # the plain DDNM step (functions/svd_ddnm.py:52-62) composed from the same pieces
        gen = torch.Generator().manual_seed(3)
        xt, et, z = (torch.randn(c["x0"].shape, generator=gen).to(dev) for _ in range(3))
        x0, xn = op.ddnm_step(xt, et, z, c["A"].to(dev), 0.5, 0.6, 0.85, None)
        want0 = (xt - et * math.sqrt(1 - 0.5)) / math.sqrt(0.5)
        assert (x0 - want0).abs().max() < 1e-5
        st = math.sqrt(1 - 0.6)
        wantn = math.sqrt(0.6) * op.project(x0, c["A"].to(dev)).reshape(x0.shape) + st * 0.85 * z + \
            st * math.sqrt(1 - 0.85 ** 2) * et
        assert (xn - wantn).abs().max() < 1e-5
    R, B = 256, 4
    torch.manual_seed(11)
    op = P.CS(3, R, 0.25, dev)  # the reference's own construction: random Gaussian matrix, its right singular vectors
    x = (torch.rand(B, 3 * R * R) * 2 - 1).to(dev)
    x0 = torch.randn(B, 3, R, R).to(dev)
    y = op.A(x)
    assert y.shape == (B, 3 * 64 * 256)
    assert (op.A(op.A_pinv(y)) - y).abs().max() < 2e-5
    p = op.project(x0, y)
    assert (op.A(p) - y).abs().max() < 5e-5
    assert (op.project(p, y) - p).abs().max() < 5e-5
    lhs, rhs = (op.A(x) * y).sum(dim=1), (x * op.At(y)).sum(dim=1)
    assert ((lhs - rhs).abs() / lhs.abs().clamp_min(1.0)).max() < 1e-3


@pytest.mark.parametrize("R", [64, 128, 256, 512])
def test_cluster_fwht_equals_two_kernel_fwht(R):
    """The one-kernel transform (thread-block cluster, the plane in distributed shared memory; fwht_cluster.cu) runs the
    same butterflies in the same stage order as the two-kernel one: results are bit-identical, for every entry point
    that goes through it; at R = 128 both are also checked against the oracle's spectral restatement."""
    from nlc_b200 import svd_operators as P
    gen = torch.Generator().manual_seed(R)
    B, C = 3, 3
    perm = torch.randperm(R * R, generator=gen)
    op = P.WalshHadamardCS(C, R, 4, perm, dev)
    x = (torch.rand(B, C * R * R, generator=gen) * 2 - 1).to(dev)
    x0 = torch.randn(B, C, R, R, generator=gen).to(dev)
    et, z = torch.randn(B, 2 * C, R, R, generator=gen).to(dev), torch.randn(B, C, R, R, generator=gen).to(dev)

    def run():
        y = op.A(x)
        return [y, op.At(y), op.project(x0, y), op.Lambda(x, 0.8, 0.2, 0.1, 0.85), op.Lambda_noise(x, 0.8, 0.2, 0.3, 0.85, x0),
                *op.ddnm_step(x0, et, z, y, 0.5, 0.6, 0.85, None), *op.ddnm_step(x0, et, z, y, 0.5, 0.6, 0.85, 0.1)]

    old = os.environ.get("NLC_FWHT_CLUSTER")
    try:
        os.environ["NLC_FWHT_CLUSTER"] = "1"
        a = run()
        os.environ["NLC_FWHT_CLUSTER"] = "0"
        b = run()
    finally:
        if old is None:
            os.environ.pop("NLC_FWHT_CLUSTER", None)
        else:
            os.environ["NLC_FWHT_CLUSTER"] = old
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    if R == 128:
        orc = O.WalshHadamardCS(C, R, 4, perm)
        y = orc.A(x.cpu())
        assert (a[0].cpu() - y).abs().max() < 2e-6
        assert (a[2].cpu() - orc.project(x0.cpu(), y)).abs().max() < 1e-5
