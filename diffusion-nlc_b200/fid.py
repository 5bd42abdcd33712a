"""FID statistics on the device.

The reference computes FID with the third-party `pytorch_fid` package after writing every sample to disk as a PNG
(`fid_helper`, src/experiments.py:210-226: InceptionV3([3]) -> compute_statistics_of_path -> calculate_frechet_distance;
callers image_sample.py:566,703 and result_evaluater.py:24-27).  Here the samples never leave the GPU:

  * `InceptionV3`: pytorch_fid's FID Inception trunk (torchvision Inception3 up to Mixed_7c with the FIDInceptionA / C / E_1
    / E_2 pooling patches) -> 2048 pool3 features.  Every BasicConv2d is one tensor-core GEMM (`nlc_conv_tc`, tcgen05) over
    a patch matrix written by `nlc_im2col_nhwc` (1x1 convolutions read the activation directly), with the eval-mode
    BatchNorm folded into weight and bias and the ReLU in the conv epilogue; the four branches of a Mixed block write their
    channel slices of one concat buffer.  The image spatial sizes (299, 149, 147, 73, 71, 35, 17, 8) are not powers of two,
    so the GEMM rows are plain pixel indices padded to a multiple of 128.  The PNG round trip of the reference's drivers
    (8-bit quantisation) and the 299 x 299 bilinear resize are one kernel (`nlc_fid_preprocess`).
  * `FidStatistics`: sum and outer-product accumulation of the features in fp64 on the device (`nlc_cov_accumulate`),
    all-reduced over the ranks of a sharded run -> (mu, Sigma) exactly as np.mean / np.cov(rowvar=False).
  * `calculate_frechet_distance`: the host-side formula of pytorch_fid (scipy.linalg.sqrtm, once per evaluation - as in the
    reference), and `fid_helper` which gives an experiment object the reference's `fid_fn`.

The pretrained FID weights (pt_inception-2015-12-05) are a state_dict in torchvision's layout: `load_state_dict` takes it
as is.  Oracle: oracle/fid.py (pinned against torchvision's own modules).
"""
import math

import numpy as np
import torch

from . import ops
from ._lib import NLC_BF16, NLC_F16, NLC_F32X3
from .engine import PRECISIONS
from .ops import Act

BN_EPS = 1e-3
DIMS = 2048

# (name, (kh, kw), (sh, sw), (ph, pw)) of the stem; the Mixed blocks are wired in InceptionV3._build_plan
_A = ("Mixed_5b", "Mixed_5c", "Mixed_5d")
_C = ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e")


def _ceil(x, m):
    return (x + m - 1) // m * m


class _Conv:
    """BasicConv2d (conv without bias -> BatchNorm2d(eps 1e-3, eval) -> ReLU) as GEMM weights [Cout_pad, K_pad] + bias."""

    def __init__(self, sd, name, op_dtype, device, stride=(1, 1), pad=(0, 0)):
        w = sd[name + ".conv.weight"].detach().float()
        g, b = sd[name + ".bn.weight"].detach().float(), sd[name + ".bn.bias"].detach().float()
        m, v = sd[name + ".bn.running_mean"].detach().float(), sd[name + ".bn.running_var"].detach().float()
        s = g / torch.sqrt(v + BN_EPS)
        w = w * s[:, None, None, None]
        self.name, self.stride, self.pad = name, stride, pad
        self.cout, self.cin, self.kh, self.kw = w.shape
        chunk = 64 if op_dtype in (NLC_BF16, NLC_F16) else 32
        self.K = self.kh * self.kw * self.cin
        self.K_pad, self.cout_pad = _ceil(self.K, chunk), _ceil(self.cout, 64)
        k = torch.zeros(self.cout_pad, self.K_pad, dtype=torch.float32)
        k[:self.cout, :self.K] = w.permute(0, 2, 3, 1).reshape(self.cout, self.K)
        k = k.to(device)
        if op_dtype in (NLC_BF16, NLC_F16):
            self.w = k.to(ops.OP_DTYPES[op_dtype]).contiguous()
        else:
            self.w = k if op_dtype == NLC_F32X3 else ops.round_tf32_(k)
        bias = torch.zeros(self.cout_pad, dtype=torch.float32)
        bias[:self.cout] = b - m * s
        self.bias = bias.to(device)


class _Map:
    """A feature map: channels [c0, c0 + C) of a [M_pad, ld] operand-dtype matrix whose row m is pixel (n, h, w)."""

    def __init__(self, t, B, H, W, C, c0=0):
        self.t, self.B, self.H, self.W, self.C, self.c0 = t, B, H, W, C, c0

    ld = property(lambda s: s.t.shape[1])
    M = property(lambda s: s.B * s.H * s.W)
    ptr = property(lambda s: s.t.data_ptr() + s.c0 * s.t.element_size())

    def as_act(self, C=None):
        """View for nlc_conv_tc: [1, M_pad / 128, 128, ld] (rows of 128 pixels tile as BH = 1)."""
        return Act(self.t.view(1, self.t.shape[0] // 128, 128, self.t.shape[1]), self.c0, self.C if C is None else C)


class InceptionV3:
    """pytorch_fid `InceptionV3([3], resize_input=True, normalize_input=True)` on the device: images in [0, 1] (or samples in
    [-1, 1] through `features_of_samples`) -> [B, 2048] fp32 features."""

    def __init__(self, precision="fp16", device="cuda", resize_input=True, normalize_input=True):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % list(PRECISIONS))
        self.device = torch.device(device)
        self.op_dtype = PRECISIONS[precision]
        self.op_torch = ops.OP_DTYPES[self.op_dtype]
        self.resize_input, self.normalize_input = resize_input, normalize_input
        self._plans, self._bufs, self._loaded = {}, {}, False
        self._patch_numel = 0

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, strict=True):
        names = sorted({k[:-len(".conv.weight")] for k in sd if k.endswith(".conv.weight") and not k.startswith("AuxLogits")})
        strides = {"Conv2d_1a_3x3": (2, 2), "Mixed_6a.branch3x3": (2, 2), "Mixed_6a.branch3x3dbl_3": (2, 2),
                   "Mixed_7a.branch3x3_2": (2, 2), "Mixed_7a.branch7x7x3_4": (2, 2)}
        self.convs = {}
        for n in names:
            kh, kw = sd[n + ".conv.weight"].shape[2:]
            pad = ((kh - 1) // 2, (kw - 1) // 2)
            if n in ("Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_4a_3x3") or n in strides:
                pad = (0, 0)  # the 'valid' convolutions of the stem and the stride-2 reductions
            self.convs[n] = _Conv(sd, n, self.op_dtype, self.device, stride=strides.get(n, (1, 1)), pad=pad)
        need = ["Conv2d_1a_3x3", "Conv2d_2a_3x3", "Conv2d_2b_3x3", "Conv2d_3b_1x1", "Conv2d_4a_3x3", "Mixed_7c.branch_pool"]
        missing = [n for n in need if n not in self.convs]
        if missing:
            raise KeyError("InceptionV3 state_dict lacks %s" % missing)
        self._loaded, self._plans = True, {}
        return self

    # ------------------------------------------------------------------ plan
    def _buf(self, tag, rows, cols):
        """Zero-initialised [rows, cols] operand matrix owned by one feature map of one plan."""
        t = self._bufs.get(tag)
        if t is None or t.shape != (rows, cols):
            t = torch.zeros(rows, cols, device=self.device, dtype=self.op_torch)
            self._bufs[tag] = t
        return t

    def _patches(self, rows, cols):
        """[rows, cols] view of the one patch-matrix workspace all im2col'ed convolutions share (stream order makes the
        reuse safe); sized for the largest request seen while planning, allocated at the first run."""
        t = self._bufs.get("patches")
        if t is None or t.numel() < self._patch_numel:
            t = torch.zeros(self._patch_numel, device=self.device, dtype=self.op_torch)
            self._bufs["patches"] = t
        return t[:rows * cols].view(rows, cols)

    def _plan(self, B, H, W):
        key = (B, H, W)
        if key not in self._plans:
            assert self._loaded, "load_state_dict() first"
            self._plans[key] = self._build_plan(B, H, W)
        return self._plans[key]

    def _build_plan(self, B, H, W):
        dt, dev = self.op_dtype, self.device
        steps = []
        P = {"x": torch.empty(B, 3, H, W, device=dev, dtype=torch.float32), "opts": [0, 0],
             "feat": torch.empty(B, DIMS, device=dev, dtype=torch.float32)}
        R = 299 if self.resize_input else H
        assert self.resize_input or H == W, "without resizing the network takes square images"
        uid = [0]

        def new_map(tag, Bn, Hh, Ww, C, ld=None):
            uid[0] += 1
            return _Map(self._buf("%s.%d.%d" % (tag, uid[0], B), _ceil(Bn * Hh * Ww, 128), _ceil(ld or C, 8)), Bn, Hh, Ww, C)

        x0 = new_map("in", B, R, R, 3, ld=8)
        steps.append(lambda: ops.fid_preprocess(P["x"], P["opts"][0], P["opts"][1], self.resize_input, self.normalize_input, R,
                                                x0, dt))

        def conv(x, name, out=None):
            """x: _Map -> BasicConv2d `name` -> `out` (a channel slice of a concat buffer) or a fresh map."""
            cv = self.convs[name]
            assert cv.cin == x.C, (name, cv.cin, x.C)
            Ho = (x.H + 2 * cv.pad[0] - cv.kh) // cv.stride[0] + 1
            Wo = (x.W + 2 * cv.pad[1] - cv.kw) // cv.stride[1] + 1
            if out is None:
                out = new_map(name, x.B, Ho, Wo, cv.cout, ld=cv.cout_pad)
            assert (out.H, out.W, out.C) == (Ho, Wo, cv.cout) and out.c0 + cv.cout_pad <= out.ld, name
            M_pad = out.t.shape[0]
            chunk = 64 if dt in (NLC_BF16, NLC_F16) else 32
            direct = (cv.kh == 1 and cv.kw == 1 and cv.stride == (1, 1) and x.C % chunk == 0 and x.c0 % 8 == 0
                      and x.t.shape[0] == M_pad)
            if direct:  # 1x1: the activation matrix is the GEMM operand (K = C = K_pad)
                a, K = x.as_act(), x.C
                o = out.as_act(cv.cout_pad)
                steps.append(lambda: ops.conv_tc([a], [(0, 0, 0, 0, K)], cv.w, cv.cout_pad, 1, M_pad // 128, 128, dt,
                                                 bias=cv.bias, out_op=o, relu=True))
                return out
            self._patch_numel = max(self._patch_numel, M_pad * cv.K_pad)
            o = out.as_act(cv.cout_pad)

            def run():
                patches = self._patches(M_pad, cv.K_pad)
                ops.im2col_nhwc(x, cv.kh, cv.kw, cv.stride, cv.pad, patches, dt)
                ops.conv_tc([Act(patches.view(1, M_pad // 128, 128, cv.K_pad))], [(0, 0, 0, 0, cv.K_pad)], cv.w, cv.cout_pad, 1,
                            M_pad // 128, 128, dt, bias=cv.bias, out_op=o, relu=True)

            steps.append(run)
            return out

        def pool(x, stride, pad, mode, out=None):
            Ho, Wo = (x.H + 2 * pad - 3) // stride + 1, (x.W + 2 * pad - 3) // stride + 1
            if out is None:
                out = new_map("pool", x.B, Ho, Wo, x.C)
            steps.append(lambda: ops.pool2d(x, stride, pad, mode, out, dt))
            return out

        def cat_buffer(tag, like, Hh, Ww, widths):
            """Concat buffer whose slices are written in increasing channel order: a branch's zero-padded GEMM columns (Cout
            rounded up to 64) spill into the next slice - rewritten afterwards by its own producer - or into row slack."""
            total = sum(widths)
            ld = max(total, max(sum(widths[:i]) + _ceil(wd, 64) for i, wd in enumerate(widths)))
            full = new_map(tag, like.B, Hh, Ww, total, ld=ld)
            offs = [sum(widths[:i]) for i in range(len(widths))]
            return full, [_Map(full.t, like.B, Hh, Ww, wd, c0) for wd, c0 in zip(widths, offs)]

        x = conv(conv(conv(x0, "Conv2d_1a_3x3"), "Conv2d_2a_3x3"), "Conv2d_2b_3x3")
        x = pool(x, 2, 0, 0)
        x = conv(conv(x, "Conv2d_3b_1x1"), "Conv2d_4a_3x3")
        x = pool(x, 2, 0, 0)
        for name in _A:  # FIDInceptionA
            pf = self.convs[name + ".branch_pool"].cout
            full, (s1, s5, s3, sp) = cat_buffer(name, x, x.H, x.W, [64, 64, 96, pf])
            conv(x, name + ".branch1x1", s1)
            conv(conv(x, name + ".branch5x5_1"), name + ".branch5x5_2", s5)
            conv(conv(conv(x, name + ".branch3x3dbl_1"), name + ".branch3x3dbl_2"), name + ".branch3x3dbl_3", s3)
            conv(pool(x, 1, 1, 1), name + ".branch_pool", sp)
            x = full
        Ho = (x.H - 3) // 2 + 1  # Mixed_6a
        full, (s3, sd, sp) = cat_buffer("Mixed_6a", x, Ho, Ho, [384, 96, x.C])
        conv(x, "Mixed_6a.branch3x3", s3)
        conv(conv(conv(x, "Mixed_6a.branch3x3dbl_1"), "Mixed_6a.branch3x3dbl_2"), "Mixed_6a.branch3x3dbl_3", sd)
        pool(x, 2, 0, 0, sp)
        x = full
        for name in _C:  # FIDInceptionC
            full, (s1, s7, sd, sp) = cat_buffer(name, x, x.H, x.W, [192, 192, 192, 192])
            conv(x, name + ".branch1x1", s1)
            conv(conv(conv(x, name + ".branch7x7_1"), name + ".branch7x7_2"), name + ".branch7x7_3", s7)
            y = conv(x, name + ".branch7x7dbl_1")
            for i in (2, 3, 4):
                y = conv(y, name + ".branch7x7dbl_%d" % i)
            conv(y, name + ".branch7x7dbl_5", sd)
            conv(pool(x, 1, 1, 1), name + ".branch_pool", sp)
            x = full
        Ho = (x.H - 3) // 2 + 1  # Mixed_7a
        full, (s3, s7, sp) = cat_buffer("Mixed_7a", x, Ho, Ho, [320, 192, x.C])
        conv(conv(x, "Mixed_7a.branch3x3_1"), "Mixed_7a.branch3x3_2", s3)
        y = conv(x, "Mixed_7a.branch7x7x3_1")
        y = conv(conv(y, "Mixed_7a.branch7x7x3_2"), "Mixed_7a.branch7x7x3_3")
        conv(y, "Mixed_7a.branch7x7x3_4", s7)
        pool(x, 2, 0, 0, sp)
        x = full
        for name, mode in (("Mixed_7b", 1), ("Mixed_7c", 0)):  # FIDInceptionE_1 (average pool) / FIDInceptionE_2 (max pool)
            full, (s1, s3a, s3b, sda, sdb, sp) = cat_buffer(name, x, x.H, x.W, [320, 384, 384, 384, 384, 192])
            conv(x, name + ".branch1x1", s1)
            y = conv(x, name + ".branch3x3_1")
            conv(y, name + ".branch3x3_2a", s3a)
            conv(y, name + ".branch3x3_2b", s3b)
            y = conv(conv(x, name + ".branch3x3dbl_1"), name + ".branch3x3dbl_2")
            conv(y, name + ".branch3x3dbl_3a", sda)
            conv(y, name + ".branch3x3dbl_3b", sdb)
            conv(pool(x, 1, 1, mode), name + ".branch_pool", sp)
            x = full
        assert x.C == DIMS
        last = x
        steps.append(lambda: ops.global_avgpool(last, P["feat"], dt))
        P["steps"] = steps
        return P

    # ------------------------------------------------------------------ execution
    def _run(self, x, from_pm1, quantize):
        assert x.dim() == 4 and x.shape[1] == 3, "images are [B, 3, H, W]"
        P = self._plan(*[int(x.shape[i]) for i in (0, 2, 3)])
        P["x"].copy_(x)
        P["opts"][0], P["opts"][1] = int(from_pm1), int(quantize)
        for fn in P["steps"]:
            fn()
        return P["feat"]

    def __call__(self, x):
        """Images in [0, 1] (what pytorch_fid's model is fed) -> [B, 2048] features (a fresh tensor)."""
        return self._run(x, False, False).clone()

    def features_of_samples(self, samples):
        """Sampler outputs in [-1, 1] -> features, through the reference drivers' image round trip: add(1).div(2).clamp(0,1)
        (image_sample.py:560), save_image's 8-bit quantisation and ImagePathDataset's ToTensor - without the disk."""
        return self._run(samples, True, True).clone()

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class FidStatistics:
    """mu, Sigma of pool3 features accumulated on the device in fp64 (fid_score.calculate_activation_statistics)."""

    def __init__(self, dims=DIMS, device="cuda"):
        self.dims, self.device = dims, torch.device(device)
        self.sum = torch.zeros(dims, device=self.device, dtype=torch.float64)
        self.outer = torch.zeros(dims, dims, device=self.device, dtype=torch.float64)
        self.count = 0

    def update(self, feats):
        feats = feats.to(self.device, torch.float32).contiguous()
        assert feats.dim() == 2 and feats.shape[1] == self.dims
        ops.cov_accumulate(feats, self.sum, self.outer)
        self.count += feats.shape[0]
        return self

    def all_reduce(self, group=None):
        """Sum the partial statistics of the ranks of a sharded run (NCCL; no-op without a process group)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            n = torch.tensor([float(self.count)], device=self.device, dtype=torch.float64)
            for t in (self.sum, self.outer, n):
                dist.all_reduce(t, group=group)
            self.count = int(round(n.item()))
        return self

    def finalize(self):
        """(mu, sigma) as float64 numpy arrays: np.mean(act, 0), np.cov(act, rowvar=False)."""
        n = float(self.count)
        mu = self.sum / n
        sigma = (self.outer - n * torch.outer(mu, mu)) / (n - 1.0)
        return mu.cpu().numpy(), sigma.cpu().numpy()


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """pytorch_fid fid_score.calculate_frechet_distance (host side, once per evaluation, as in the reference)."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    if mu1.shape != mu2.shape or sigma1.shape != sigma2.shape:
        raise ValueError("training and test statistics have different shapes")
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
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def fid_helper(experiment, fid_target, inception, batch_size=128):
    """The reference's `ExperimentDiffusion.fid_helper` (src/experiments.py:210-226) for device tensors: gives `experiment`
    a `fid_fn(samples)` that returns the FID of sampler outputs in [-1, 1] ([N,3,H,W] tensor or an iterable of batches)
    against the target statistics - `fid_target` is the reference's .npz path (keys mu, sigma) or a (mu, sigma) pair - and a
    `fid_stats()` factory for drivers that accumulate batch by batch (image_sample.evaluate_*)."""
    if isinstance(fid_target, str):
        with np.load(fid_target) as f:
            m1, s1 = f["mu"][:], f["sigma"][:]
    else:
        m1, s1 = fid_target

    def fid_of(stats):
        m2, s2 = stats.all_reduce().finalize()
        return calculate_frechet_distance(m1, s1, m2, s2)

    def calc_fid(samples):
        stats = FidStatistics(DIMS, inception.device)
        batches = samples.split(batch_size) if torch.is_tensor(samples) else samples
        for b in batches:
            stats.update(inception.features_of_samples(b.to(inception.device, torch.float32)))
        return fid_of(stats)

    experiment.fid_fn = calc_fid
    experiment.fid_stats = lambda: FidStatistics(DIMS, inception.device)
    experiment.fid_inception = inception
    experiment.fid_of = fid_of
    return calc_fid
