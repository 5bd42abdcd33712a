"""FID statistics on the device (nlc_b200.fid, csrc/fid.cu) against the oracle restatement of pytorch_fid (oracle/fid.py,
pinned against torchvision's Inception3 modules) and the golden fixture tests/golden/fid_tiny.pt.

Tolerances (max-norm relative on the 2048 pool3 features): fp32 mode 1e-4, tf32 / fp16 1e-2, bf16 5e-2; the glue kernels
(preprocess, im2col, pooling, statistics) against torch / numpy to fp32 round-off."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fid as OF
from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
TOL = {"fp32": 1e-4, "tf32": 1e-2, "fp16": 1e-2, "bf16": 5e-2}


def _maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def sd():
    return weights.fid_inception_state_dict(seed=7)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "fid_tiny.pt"), weights_only=True)


def _map(t4, dtype=torch.float32):
    """NHWC tensor [B,H,W,C] -> fid._Map over a zero-padded [M_pad, C] matrix."""
    from nlc_b200.fid import _Map, _ceil
    B, H, W, C = t4.shape
    m = torch.zeros(_ceil(B * H * W, 128), _ceil(C, 8), device=dev, dtype=dtype)
    m[:B * H * W, :C] = t4.reshape(-1, C).to(dev, dtype)
    return _Map(m, B, H, W, C)


def test_preprocess_matches_torch():
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    g = torch.Generator().manual_seed(1)
    s = torch.randn(2, 3, 40, 56, generator=g) * 0.7
    want = 2 * F.interpolate(OF.png_round_trip((s + 1) / 2), size=(299, 299), mode="bilinear", align_corners=False) - 1
    y = _map(torch.zeros(2, 299, 299, 3))
    ops.fid_preprocess(s.to(dev), True, True, True, True, 299, y, NLC_F32X3)
    got = y.t[:2 * 299 * 299, :3].reshape(2, 299, 299, 3).permute(0, 3, 1, 2).cpu()
    assert (got - want).abs().max() < 2e-6
    # no resize, no quantisation, images already in [0, 1]
    x = torch.rand(1, 3, 17, 17, generator=g)
    y = _map(torch.zeros(1, 17, 17, 3))
    ops.fid_preprocess(x.to(dev), False, False, False, True, 17, y, NLC_F32X3)
    assert torch.equal(y.t[:289, :3].reshape(1, 17, 17, 3).permute(0, 3, 1, 2).cpu(), 2 * x - 1)


@pytest.mark.parametrize("C,k,stride,pad", [(3, (3, 3), (2, 2), (0, 0)), (16, (1, 7), (1, 1), (0, 3)), (8, (7, 1), (1, 1), (3, 0)),
                                            (24, (5, 5), (1, 1), (2, 2)), (5, (3, 3), (1, 1), (1, 1))])
def test_im2col_matches_unfold(C, k, stride, pad):
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    from nlc_b200.fid import _ceil
    x = torch.randn(2, C, 13, 11, generator=torch.Generator().manual_seed(2))
    cols = F.unfold(x, k, padding=pad, stride=stride)  # [B, C*kh*kw, L], channel-major
    Ho = (13 + 2 * pad[0] - k[0]) // stride[0] + 1
    Wo = (11 + 2 * pad[1] - k[1]) // stride[1] + 1
    want = cols.view(2, C, k[0] * k[1], Ho * Wo).permute(0, 3, 2, 1).reshape(2 * Ho * Wo, k[0] * k[1] * C)  # tap-major
    K_pad, M_pad = _ceil(want.shape[1], 32), _ceil(want.shape[0], 128)
    out = torch.full((M_pad, K_pad), float("nan"), device=dev)
    ops.im2col_nhwc(_map(x.permute(0, 2, 3, 1)), k[0], k[1], stride, pad, out, NLC_F32X3)
    assert torch.equal(out[:want.shape[0], :want.shape[1]].cpu(), want)
    assert (out[want.shape[0]:] == 0).all() and (out[:, want.shape[1]:] == 0).all()


@pytest.mark.parametrize("stride,pad,mode", [(2, 0, 0), (1, 1, 0), (1, 1, 1)])
def test_pooling_matches_torch(stride, pad, mode):
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    x = torch.randn(2, 12, 9, 11, generator=torch.Generator().manual_seed(3))
    want = F.max_pool2d(x, 3, stride=stride, padding=pad) if mode == 0 else \
        F.avg_pool2d(x, 3, stride=stride, padding=pad, count_include_pad=False)
    y = _map(torch.zeros(2, want.shape[2], want.shape[3], 12))
    ops.pool2d(_map(x.permute(0, 2, 3, 1)), stride, pad, mode, y, NLC_F32X3)
    got = y.t[:want.shape[0] * want.shape[2] * want.shape[3], :12].reshape(2, want.shape[2], want.shape[3], 12)
    assert (got.permute(0, 3, 1, 2).cpu() - want).abs().max() < 1e-6
    f = torch.empty(2, 12, device=dev)
    ops.global_avgpool(_map(x.permute(0, 2, 3, 1)), f, NLC_F32X3)
    assert (f.cpu() - x.mean(dim=(2, 3))).abs().max() < 1e-6


def test_statistics_accumulate_in_fp64():
    from nlc_b200.fid import FidStatistics
    rng = np.random.default_rng(4)
    feats = (rng.normal(size=(70, 40)) + 3.0).astype(np.float32)
    st = FidStatistics(dims=40, device=dev)
    for lo in range(0, 70, 32):  # uneven batches
        st.update(torch.from_numpy(feats[lo:lo + 32]).to(dev))
    mu, sigma = st.finalize()
    m, s = OF.statistics(feats)
    assert st.count == 70 and np.abs(mu - m).max() < 1e-12 and np.abs(sigma - s).max() < 1e-10


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_inception_features(sd, gold, prec):
    from nlc_b200.fid import InceptionV3
    net = InceptionV3(precision=prec, device=dev).load_state_dict(sd)
    f = net(gold["x"].to(dev))
    assert f.shape == (2, 2048) and _maxrel(f.cpu(), gold["features"]) < TOL[prec]
    fs = net.features_of_samples(gold["samples"].to(dev))
    assert _maxrel(fs.cpu(), gold["features_of_samples"]) < TOL[prec]
    # a second call at another batch size and image size re-plans and stays consistent with the oracle
    x = torch.rand(3, 3, 32, 32, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = OF.inception_features(sd, x)
    assert _maxrel(net(x.to(dev)).cpu(), want) < TOL[prec]


def test_fid_of_samples_end_to_end(sd):
    """fid_helper: the reference's fid_fn for device tensors against the oracle pipeline on the same samples."""
    import types
    from nlc_b200.fid import InceptionV3, fid_helper
    g = torch.Generator().manual_seed(12)
    samples = torch.randn(12, 3, 32, 32, generator=g) * 0.5
    target = torch.rand(12, 3, 32, 32, generator=g)
    with torch.no_grad():
        ft = OF.inception_features(sd, target)[:, :64]
        fs = OF.inception_features(sd, OF.png_round_trip((samples + 1) / 2))
    # (12 samples cannot carry a 2048 x 2048 covariance, and a 2048 x 2048 sqrtm is tens of seconds of host time: both sides
    # are evaluated on the first 64 features - the device statistics are cut to the same block)
    m1, s1 = OF.statistics(ft.numpy())
    m2, s2 = OF.statistics(fs[:, :64].numpy())
    want = OF.frechet_distance(m1, s1, m2, s2)
    net = InceptionV3(precision="fp32", device=dev).load_state_dict(sd)
    exp = types.SimpleNamespace()
    fid_helper(exp, (np.zeros(2048), np.eye(2048)), net, batch_size=5)
    st = exp.fid_stats()
    for b in samples.split(5):
        st.update(net.features_of_samples(b.to(dev)))
    mu, sigma = st.finalize()
    from nlc_b200.fid import calculate_frechet_distance
    got = calculate_frechet_distance(m1, s1, mu[:64], sigma[:64, :64])
    assert abs(got - want) <= 1e-3 * abs(want) + 1e-6
    assert callable(exp.fid_fn) and st.count == 12
