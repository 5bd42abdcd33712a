"""One eager call of every operator entry point at the benchmark size (R = 256, batch 64) for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on \
        -k regex:'color_vec|sr_vec|inpaint_back_vec|identity_vec|fwht_rows|fwht_cols|needle_ddnm|color_ddnm|mask_ddnm|mix_kernel' \
        -f -o gpurun_out/TAG_ncu_operators python scripts/op_ncu.py
(after the same command has exited 0 without ncu).  The summary goes to profiles/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nlc_b200  # noqa: E402,F401
from nlc_b200 import svd_operators as P  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    R, C = 256, 3
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1)
    mask = torch.ones(R, R)
    mask[R // 4:3 * R // 4, R // 4:3 * R // 4] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    ops = [P.Colorization(R, dev), P.SuperResolution(C, R, 4, dev), P.Inpainting(C, R, torch.cat([mr, mr + 1, mr + 2]), dev),
           P.Denoising(C, R, dev), P.WalshHadamardCS(C, R, 4, torch.randperm(R * R, generator=gen), dev)]
    xt = torch.randn(B, C, R, R, device=dev)
    et = torch.randn(B, 2 * C, R, R, device=dev)
    z = torch.randn(B, C, R, R, device=dev)
    for op in ops:
        y = op.A(xt.reshape(B, -1))
        op.project(xt, y)
        op.ddnm_step(xt, et, z, y, 0.5, 0.6, 0.85, None)
        op.ddnm_step(xt, et, z, y, 0.5, 0.6, 0.85, 0.1)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
