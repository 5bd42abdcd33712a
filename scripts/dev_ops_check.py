"""Development check on a B200: DDNM operator kernels against the CPU oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import svd_operators as P
from oracle import operators as O

dev = torch.device("cuda:0")


def rel(a, b):
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-30)).item()


def check(name, prod, orc, R, B=3, C=3):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, C * R * R, generator=g) * 2 - 1
    x0 = torch.randn(B, C, R, R, generator=g)
    y = orc.A(x.clone())
    res = []
    res.append(("A", prod.A(x.to(dev)).cpu(), y))
    res.append(("At", prod.At(y.to(dev)).cpu(), orc.At(y.clone())))
    res.append(("Apinv", prod.A_pinv(y.to(dev)).cpu(), orc.A_pinv(y.clone())))
    pr = orc.project(x0, y)
    got = prod.project(x0.to(dev), y.to(dev)).cpu()
    res.append(("project", got, pr))
    resid = (orc.A(got.reshape(B, -1)) - y).abs().max().item()
    print("%-10s R=%-4d " % (name, R) + "  ".join("%s rel %.2e max %.2e" % (n, rel(a, b), (a - b).abs().max().item())
                                                   for n, a, b in res) + "  |A x^ - y| %.2e" % resid, flush=True)


if __name__ == "__main__":
    for R in (32, 64, 256):
        C = 3
        mask = torch.ones(R, R)
        mask[R // 4:3 * R // 4, R // 4:3 * R // 4] = 0
        mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
        missing = torch.cat([mr, mr + 1, mr + 2])
        check("inpaint", P.Inpainting(C, R, missing, dev), O.Inpainting(C, R, missing), R)
        check("color", P.Colorization(R, dev), O.Colorization(R), R)
        check("sr_avg4", P.SuperResolution(C, R, 4, dev), O.SuperResolution(C, R, 4), R)
        perm = torch.randperm(R * R, generator=torch.Generator().manual_seed(3))
        check("whcs4", P.WalshHadamardCS(C, R, 4, perm, dev), O.WalshHadamardCS(C, R, 4, perm), R)
        k = O.bicubic_kernel(4)
        check("sr_bicubic", P.SRConv(k.clone(), C, R, dev, stride=4), O.SRConv(k.clone(), C, R, 4), R)
        gk = O.gauss_kernel()
        check("deblur", P.Deblurring(gk.clone(), C, R, dev), O.Deblurring(gk.clone(), C, R), R)
