"""TEST INFRASTRUCTURE (oracle): the FID pipeline of the reference, restated on the CPU.

The reference computes FID through the third-party package `pytorch_fid` (imports: src/experiments.py:15-16,
result_evaluater.py:14; call sites: src/experiments.py:210-226 `fid_helper` -> `compute_statistics_of_path` +
`calculate_frechet_distance`, image_sample.py:566,703, result_evaluater.py:24-27).  The package is NOT vendored in the
reference tree, is not installed in this image and the reference pins no version (no requirements file); this file restates
its published algorithm (pytorch-fid 0.3.0, `inception.py` / `fid_score.py`):

  * `InceptionV3([3], resize_input=True, normalize_input=True)`: bilinear resize to 299 x 299 (align_corners=False),
    x -> 2x - 1, torchvision's Inception3 trunk up to Mixed_7c with the "FID Inception" patches (FIDInceptionA / C / E_1:
    the 3x3 average pool of the pool branch excludes the zero padding, count_include_pad=False; FIDInceptionE_2: that
    branch uses a MAX pool), adaptive average pool -> 2048 features;
  * activations -> mu = mean, sigma = np.cov(rowvar=False) (unbiased), in float64;
  * Frechet distance |mu1 - mu2|^2 + tr(s1) + tr(s2) - 2 tr(sqrtm(s1 s2)) with the eps-regularised retry and the
    imaginary-part check;
  * the image round trip of the reference's drivers: samples are written with torchvision `save_image` (x*255 + 0.5, clamp,
    truncate to uint8) and read back by `ImagePathDataset` (ToTensor: / 255) - `png_round_trip`.

Pinning: `inception_features` is checked against torchvision's own `Inception3` modules carrying those four patches
(tests/test_oracle_fid.py, live wherever torchvision is importable) and through tests/golden/fid_tiny.pt; statistics and
the Frechet distance against numpy / scipy directly.  The pretrained FID weights (pt_inception-2015-12-05) cannot be
fetched here: all tests use seeded random weights in the torchvision state_dict layout
(`nlc_b200.synthetic_weights.fid_inception_state_dict`).
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # torchvision BasicConv2d: BatchNorm2d(eps=0.001)


def _cbr(sd, p, x, stride=1, padding=0):
    """BasicConv2d: conv (no bias) -> BatchNorm2d(eval) -> ReLU."""
    x = F.conv2d(x, sd[p + ".conv.weight"], None, stride=stride, padding=padding)
    x = F.batch_norm(x, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"],
                     training=False, eps=BN_EPS)
    return F.relu(x)


def _inception_a(sd, p, x):  # FIDInceptionA
    b1 = _cbr(sd, p + ".branch1x1", x)
    b5 = _cbr(sd, p + ".branch5x5_2", _cbr(sd, p + ".branch5x5_1", x), padding=2)
    b3 = _cbr(sd, p + ".branch3x3dbl_1", x)
    b3 = _cbr(sd, p + ".branch3x3dbl_3", _cbr(sd, p + ".branch3x3dbl_2", b3, padding=1), padding=1)
    bp = _cbr(sd, p + ".branch_pool", F.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False))
    return torch.cat([b1, b5, b3, bp], 1)


def _inception_b(sd, p, x):  # Mixed_6a (unpatched)
    b3 = _cbr(sd, p + ".branch3x3", x, stride=2)
    bd = _cbr(sd, p + ".branch3x3dbl_2", _cbr(sd, p + ".branch3x3dbl_1", x), padding=1)
    bd = _cbr(sd, p + ".branch3x3dbl_3", bd, stride=2)
    return torch.cat([b3, bd, F.max_pool2d(x, 3, stride=2)], 1)


def _inception_c(sd, p, x):  # FIDInceptionC
    b1 = _cbr(sd, p + ".branch1x1", x)
    b7 = _cbr(sd, p + ".branch7x7_1", x)
    b7 = _cbr(sd, p + ".branch7x7_2", b7, padding=(0, 3))
    b7 = _cbr(sd, p + ".branch7x7_3", b7, padding=(3, 0))
    bd = _cbr(sd, p + ".branch7x7dbl_1", x)
    bd = _cbr(sd, p + ".branch7x7dbl_2", bd, padding=(3, 0))
    bd = _cbr(sd, p + ".branch7x7dbl_3", bd, padding=(0, 3))
    bd = _cbr(sd, p + ".branch7x7dbl_4", bd, padding=(3, 0))
    bd = _cbr(sd, p + ".branch7x7dbl_5", bd, padding=(0, 3))
    bp = _cbr(sd, p + ".branch_pool", F.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False))
    return torch.cat([b1, b7, bd, bp], 1)


def _inception_d(sd, p, x):  # Mixed_7a (unpatched)
    b3 = _cbr(sd, p + ".branch3x3_2", _cbr(sd, p + ".branch3x3_1", x), stride=2)
    b7 = _cbr(sd, p + ".branch7x7x3_1", x)
    b7 = _cbr(sd, p + ".branch7x7x3_2", b7, padding=(0, 3))
    b7 = _cbr(sd, p + ".branch7x7x3_3", b7, padding=(3, 0))
    b7 = _cbr(sd, p + ".branch7x7x3_4", b7, stride=2)
    return torch.cat([b3, b7, F.max_pool2d(x, 3, stride=2)], 1)


def _inception_e(sd, p, x, max_pool):  # FIDInceptionE_1 (average pool without padding) / FIDInceptionE_2 (max pool)
    b1 = _cbr(sd, p + ".branch1x1", x)
    b3 = _cbr(sd, p + ".branch3x3_1", x)
    b3 = torch.cat([_cbr(sd, p + ".branch3x3_2a", b3, padding=(0, 1)), _cbr(sd, p + ".branch3x3_2b", b3, padding=(1, 0))], 1)
    bd = _cbr(sd, p + ".branch3x3dbl_2", _cbr(sd, p + ".branch3x3dbl_1", x), padding=1)
    bd = torch.cat([_cbr(sd, p + ".branch3x3dbl_3a", bd, padding=(0, 1)),
                    _cbr(sd, p + ".branch3x3dbl_3b", bd, padding=(1, 0))], 1)
    pooled = F.max_pool2d(x, 3, stride=1, padding=1) if max_pool else \
        F.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False)
    bp = _cbr(sd, p + ".branch_pool", pooled)
    return torch.cat([b1, b3, bd, bp], 1)


def png_round_trip(x):
    """Images in [0, 1] as the FID network of the reference sees them: written by torchvision.utils.save_image
    (mul 255, add 0.5, clamp to [0, 255], truncate to uint8) and read back by pytorch_fid's ImagePathDataset (ToTensor)."""
    return (x * 255 + 0.5).clamp(0, 255).to(torch.uint8).to(torch.float32) / 255.0


def inception_features(sd, x, resize_input=True, normalize_input=True):
    """pytorch_fid InceptionV3([3]) forward on images x in [0, 1], [B,3,H,W] -> [B, 2048] pool3 features."""
    if resize_input:
        x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
    if normalize_input:
        x = 2 * x - 1
    x = _cbr(sd, "Conv2d_1a_3x3", x, stride=2)
    x = _cbr(sd, "Conv2d_2a_3x3", x)
    x = _cbr(sd, "Conv2d_2b_3x3", x, padding=1)
    x = F.max_pool2d(x, 3, stride=2)
    x = _cbr(sd, "Conv2d_3b_1x1", x)
    x = _cbr(sd, "Conv2d_4a_3x3", x)
    x = F.max_pool2d(x, 3, stride=2)
    for name in ("Mixed_5b", "Mixed_5c", "Mixed_5d"):
        x = _inception_a(sd, name, x)
    x = _inception_b(sd, "Mixed_6a", x)
    for name in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"):
        x = _inception_c(sd, name, x)
    x = _inception_d(sd, "Mixed_7a", x)
    x = _inception_e(sd, "Mixed_7b", x, max_pool=False)
    x = _inception_e(sd, "Mixed_7c", x, max_pool=True)
    return F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)


def statistics(act):
    """fid_score.calculate_activation_statistics: mean and unbiased covariance of [N, D] activations, float64 numpy."""
    act = np.asarray(act, dtype=np.float64)
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid_score.calculate_frechet_distance."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    try:  # (pytorch_fid calls sqrtm(..., disp=False) -> (matrix, error estimate); SciPy >= 1.16 dropped the argument)
        covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
    except TypeError:
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


# --------------------------------------------------------------------------------------------------------------------------
def torchvision_fid_inception(sd):
    """The same network assembled from torchvision's own Inception3 modules with pytorch_fid's four patches applied as
    forward overrides - the independent implementation the restatement above is pinned against."""
    import torchvision
    from torchvision.models import inception as tvi

    class FIDInceptionA(tvi.InceptionA):
        def forward(self, x):
            b1 = self.branch1x1(x)
            b5 = self.branch5x5_2(self.branch5x5_1(x))
            b3 = self.branch3x3dbl_3(self.branch3x3dbl_2(self.branch3x3dbl_1(x)))
            bp = self.branch_pool(F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False))
            return torch.cat([b1, b5, b3, bp], 1)

    class FIDInceptionC(tvi.InceptionC):
        def forward(self, x):
            b1 = self.branch1x1(x)
            b7 = self.branch7x7_3(self.branch7x7_2(self.branch7x7_1(x)))
            bd = self.branch7x7dbl_5(self.branch7x7dbl_4(self.branch7x7dbl_3(self.branch7x7dbl_2(self.branch7x7dbl_1(x)))))
            bp = self.branch_pool(F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False))
            return torch.cat([b1, b7, bd, bp], 1)

    class FIDInceptionE(tvi.InceptionE):
        use_max = False

        def forward(self, x):
            b1 = self.branch1x1(x)
            b3 = self.branch3x3_1(x)
            b3 = torch.cat([self.branch3x3_2a(b3), self.branch3x3_2b(b3)], 1)
            bd = self.branch3x3dbl_2(self.branch3x3dbl_1(x))
            bd = torch.cat([self.branch3x3dbl_3a(bd), self.branch3x3dbl_3b(bd)], 1)
            pooled = F.max_pool2d(x, kernel_size=3, stride=1, padding=1) if self.use_max else \
                F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False)
            return torch.cat([b1, b3, bd, self.branch_pool(pooled)], 1)

    net = torchvision.models.Inception3(num_classes=1008, aux_logits=False, transform_input=False, init_weights=False)
    net.Mixed_5b, net.Mixed_5c, net.Mixed_5d = FIDInceptionA(192, 32), FIDInceptionA(256, 64), FIDInceptionA(288, 64)
    net.Mixed_6b, net.Mixed_6c = FIDInceptionC(768, 128), FIDInceptionC(768, 160)
    net.Mixed_6d, net.Mixed_6e = FIDInceptionC(768, 160), FIDInceptionC(768, 192)
    net.Mixed_7b, net.Mixed_7c = FIDInceptionE(1280), FIDInceptionE(2048)
    net.Mixed_7c.use_max = True
    missing = net.load_state_dict(sd, strict=False)
    assert all(k.startswith("fc.") or k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
    net.eval()

    def features(x, resize_input=True, normalize_input=True):
        if resize_input:
            x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
        if normalize_input:
            x = 2 * x - 1
        x = net.Conv2d_2b_3x3(net.Conv2d_2a_3x3(net.Conv2d_1a_3x3(x)))
        x = F.max_pool2d(x, kernel_size=3, stride=2)
        x = net.Conv2d_4a_3x3(net.Conv2d_3b_1x1(x))
        x = F.max_pool2d(x, kernel_size=3, stride=2)
        for m in (net.Mixed_5b, net.Mixed_5c, net.Mixed_5d, net.Mixed_6a, net.Mixed_6b, net.Mixed_6c, net.Mixed_6d,
                  net.Mixed_6e, net.Mixed_7a, net.Mixed_7b, net.Mixed_7c):
            x = m(x)
        return F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)

    return features
