"""Host-side mirror of the reference's DDNM operators (functions/svd_operators.py) on libnlc_b200 kernels.

Each class keeps the reference constructor and exposes `A`, `At`, `A_pinv` on `[B, C*R*R]` rows (NCHW-flattened
images, as the reference's callers pass them: image_sample.py:376-379) plus the fused `project(x0, y)` =
x0 - A_pinv(A(x0) - y).  The small SVDs are taken on the CPU with the same torch calls as the reference
(`torch.svd(..., some=False)`), then uploaded once; nothing else runs in torch.

Covered: Inpainting (:324-359), Colorization (:627-667), SuperResolution (:479-533), WalshHadamardCS (:211-251),
SRConv (:851-931), Deblurring (:934-1014), Deblurring2D (:1094-1165), Denoising (:442-476), and the DDNM+ family
(SURVEY §8f rank 2): `A_pinv_eta` (:82-91), `Lambda` / `Lambda_noise` of every class that defines them (SRConv and
Deblurring2D raise NotImplementedError as the reference's base class does, :93-97) and the fused reverse step
`ddnm_step` used by svd_ddnm.py.  The spectral-domain accessors `U/Ut/V/Vt/singulars/add_zeros` are not exposed: the
kernels work in closed form on the image.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

INPAINT, COLOR, SR_AVG, WHCS, SEPARABLE, DENOISE, BLOCKCS, GENERAL = 1, 2, 3, 4, 5, 6, 7, 8


def _f32(v):
    """A scalar as the fp32 value the reference computes with (0-dim tensors at its call sites)."""
    return float(torch.as_tensor(v, dtype=torch.float32))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class A_functions:
    """Base: owns the nlc_op handle and the per-batch workspace."""

    def __init__(self, desc, keep, device):
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ctx = _lib.ctx(idx)
        h = C.c_void_p()
        _lib.check(_lib.lib().nlc_op_create(self._ctx, C.byref(desc), C.byref(h)))
        del keep  # host arrays only had to outlive nlc_op_create
        self._h = h
        self.ydim = int(_lib.lib().nlc_op_ydim(h))
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().nlc_op_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _workspace(self, B):
        need = int(_lib.lib().nlc_op_ws(self._h, B))
        if need == 0:
            return None
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        return C.c_void_p(self._ws.data_ptr())

    def _rows(self, v, n):
        v = v.reshape(v.shape[0], -1)
        assert v.shape[1] == n, "expected rows of length %d, got %d" % (n, v.shape[1])
        return v.contiguous().float()

    def A(self, vec):
        x = self._rows(vec, self.xdim)
        y = torch.empty(x.shape[0], self.ydim, device=x.device)
        _lib.check(_lib.lib().nlc_op_A(self._h, x.data_ptr(), x.shape[0], y.data_ptr(), self._workspace(x.shape[0]),
                                       _stream()))
        return y

    def At(self, vec):
        y = self._rows(vec, self.ydim)
        x = torch.empty(y.shape[0], self.xdim, device=y.device)
        _lib.check(_lib.lib().nlc_op_At(self._h, y.data_ptr(), y.shape[0], x.data_ptr(), self._workspace(y.shape[0]),
                                        _stream()))
        return x

    def A_pinv(self, vec):
        y = self._rows(vec, self.ydim)
        x = torch.empty(y.shape[0], self.xdim, device=y.device)
        _lib.check(_lib.lib().nlc_op_Apinv(self._h, y.data_ptr(), y.shape[0], x.data_ptr(),
                                           self._workspace(y.shape[0]), _stream()))
        return x

    def A_pinv_eta(self, vec, eta):
        """V diag(s / (s^2 + eta)) U^T (functions/svd_operators.py:82-91)."""
        y = self._rows(vec, self.ydim)
        x = torch.empty(y.shape[0], self.xdim, device=y.device)
        _lib.check(_lib.lib().nlc_op_Apinv_eta(self._h, y.data_ptr(), y.shape[0], float(eta), x.data_ptr(),
                                               self._workspace(y.shape[0]), _stream()))
        return x

    _has_lambda = True

    def _coef(self, a, sigma_y, sigma_t, eta):
        if not self._has_lambda:
            raise NotImplementedError()  # as the reference's base class (functions/svd_operators.py:93-97)
        return _lib.DdnmCoef(a=_f32(a), sigma_t=_f32(sigma_t), sigma_y=float(sigma_y), eta=float(eta))

    def Lambda(self, vec, a, sigma_y, sigma_t, eta):
        """V (lambda o V^T vec), Eq. 17 of DDNM+ (e.g. functions/svd_operators.py:535-570)."""
        c = self._coef(a, sigma_y, sigma_t, eta)
        v = self._rows(vec, self.xdim)
        out = torch.empty_like(v)
        _lib.check(_lib.lib().nlc_op_lambda(self._h, v.data_ptr(), v.shape[0], C.byref(c), out.data_ptr(),
                                            self._workspace(v.shape[0]), _stream()))
        return out

    def Lambda_noise(self, vec, a, sigma_y, sigma_t, eta, epsilon):
        """V (d1 o P vec) + V (d2 o P epsilon), Eq. 51 (e.g. :572-623)."""
        c = self._coef(a, sigma_y, sigma_t, eta)
        v = self._rows(vec, self.xdim)
        e = self._rows(epsilon, self.xdim)
        out = torch.empty_like(v)
        _lib.check(_lib.lib().nlc_op_lambda_noise(self._h, v.data_ptr(), e.data_ptr(), v.shape[0], C.byref(c),
                                                  out.data_ptr(), self._workspace(v.shape[0]), _stream()))
        return out

    def ddnm_step(self, xt, et, z, y, at, at_next, eta, sigma_y=None):
        """One fused reverse step of functions/svd_ddnm.py (:40-66 when sigma_y is None, :101-132 otherwise):
        returns (x0_t, x_next).  `et` may be the [B, 2C, R, R] output of a learned-variance head: its first C
        channels are read in place."""
        if sigma_y is not None and not self._has_lambda:
            raise NotImplementedError()
        B = xt.shape[0]
        x = self._rows(xt, self.xdim)
        zz = self._rows(z, self.xdim)
        yy = self._rows(y, self.ydim)
        e = et.reshape(B, -1).contiguous().float()
        assert e.shape[1] >= self.xdim
        x0 = torch.empty_like(x)
        xn = torch.empty_like(x)
        _lib.check(_lib.lib().nlc_ddnm_step(self._h, x.data_ptr(), e.data_ptr(), e.shape[1], zz.data_ptr(), yy.data_ptr(),
                                            B, _f32(at), _f32(at_next), float(eta),
                                            0.0 if sigma_y is None else float(sigma_y), 0 if sigma_y is None else 1,
                                            x0.data_ptr(), xn.data_ptr(), self._workspace(B), _stream()))
        return x0.view(xt.shape), xn.view(xt.shape)

    def project(self, x0, y, out=None):
        """x0 - A_pinv(A(x0) - y) in one fused pass; keeps x0's shape."""
        x = self._rows(x0, self.xdim)
        yy = self._rows(y, self.ydim)
        o = torch.empty_like(x) if out is None else out
        _lib.check(_lib.lib().nlc_op_project(self._h, x.data_ptr(), yy.data_ptr(), x.shape[0], o.data_ptr(),
                                             self._workspace(x.shape[0]), _stream()))
        return o.view(x0.shape)


def _fptr(t):
    return t.data_ptr()


class Inpainting(A_functions):
    def __init__(self, channels, img_dim, missing_indices, device):
        self.channels, self.img_dim = channels, img_dim
        self.xdim = channels * img_dim ** 2
        miss = missing_indices.detach().cpu().to(torch.int64).contiguous()
        d = _lib.OpDesc(task=INPAINT, channels=channels, R=img_dim, ratio=1, idx_host=_fptr(miss), n_idx=miss.numel())
        super().__init__(d, (miss,), device)


class Denoising(A_functions):
    """A = I (functions/svd_operators.py:442-476)."""

    def __init__(self, channels, img_dim, device):
        self.channels, self.img_dim = channels, img_dim
        self.xdim = channels * img_dim ** 2
        super().__init__(_lib.OpDesc(task=DENOISE, channels=channels, R=img_dim, ratio=1), (), device)


class CS(A_functions):
    """Block-wise compressed sensing (functions/svd_operators.py:101-160).  The reference draws a 1024 x 1024 Gaussian
    matrix and keeps its right singular vectors; the same draw (global CPU generator) is made here unless `V_small` is
    given, and the SVD runs on the CPU (the reference's runs where `device` says: any orthogonal basis is a valid
    instance of the operator, the bases differ)."""
    _has_lambda = False

    def __init__(self, channels, img_dim, ratio, device, V_small=None):
        E = 32
        assert img_dim % E == 0
        self.channels, self.img_dim, self.y_dim, self.ratio = channels, img_dim, img_dim // E, E
        self.xdim = channels * img_dim ** 2
        if V_small is None:
            _, _, V_small = torch.svd(torch.randn(E ** 2, E ** 2), some=False)
        V_small = V_small.detach().cpu().float().contiguous()
        self.cs_size = int(E * E * ratio)
        d = _lib.OpDesc(task=BLOCKCS, channels=channels, R=img_dim, ratio=E, V_small_host=_fptr(V_small),
                        m_small=self.cs_size)
        super().__init__(d, (V_small,), device)


class GeneralA(A_functions):
    """Dense SVD of an arbitrary (small) matrix A [ny, nx] (functions/svd_operators.py:173-208)."""
    _has_lambda = False

    def __init__(self, A, device=None):
        device = A.device if device is None else device
        U, s, V = torch.svd(A.detach().cpu().float(), some=False)
        s = s.clone()
        s[s < 1e-3] = 0
        U, s, V = U.contiguous(), s.contiguous(), V.contiguous()
        self.xdim = V.shape[0]
        d = _lib.OpDesc(task=GENERAL, channels=1, R=1, ratio=1, n_idx=V.shape[0], U_small_host=_fptr(U),
                        V_small_host=_fptr(V), sing_small_host=_fptr(s), m_small=U.shape[0])
        super().__init__(d, (U, s, V), device)


class Colorization(A_functions):
    def __init__(self, img_dim, device):
        self.channels, self.img_dim = 3, img_dim
        self.xdim = 3 * img_dim ** 2
        A = torch.Tensor([[0.3333, 0.3334, 0.3333]])
        U, S, V = torch.svd(A, some=False)
        U, S, V = U.contiguous(), S.contiguous(), V.contiguous()
        d = _lib.OpDesc(task=COLOR, channels=3, R=img_dim, ratio=1, U_small_host=_fptr(U), V_small_host=_fptr(V),
                        sing_small_host=_fptr(S))
        super().__init__(d, (U, S, V), device)


class SuperResolution(A_functions):
    def __init__(self, channels, img_dim, ratio, device):
        assert img_dim % ratio == 0
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.xdim = channels * img_dim ** 2
        A = torch.Tensor([[1 / ratio ** 2] * ratio ** 2])
        U, S, V = torch.svd(A, some=False)
        U, S, V = U.contiguous(), S.contiguous(), V.contiguous()
        d = _lib.OpDesc(task=SR_AVG, channels=channels, R=img_dim, ratio=ratio, U_small_host=_fptr(U),
                        V_small_host=_fptr(V), sing_small_host=_fptr(S))
        super().__init__(d, (U, S, V), device)


class WalshHadamardCS(A_functions):
    def __init__(self, channels, img_dim, ratio, perm, device):
        self.channels, self.img_dim, self.ratio = channels, img_dim, ratio
        self.xdim = channels * img_dim ** 2
        p = perm.detach().cpu().to(torch.int64).contiguous()
        d = _lib.OpDesc(task=WHCS, channels=channels, R=img_dim, ratio=ratio, idx_host=_fptr(p), n_idx=p.numel())
        super().__init__(d, (p,), device)


def _separable(self, U_s, V_s, mult, pinv, channels, img_dim, m, device, U2_s=None, V2_s=None, lambda_sing=None):
    U_s, V_s = U_s.contiguous().float(), V_s.contiguous().float()
    mult, pinv = mult.contiguous().float(), pinv.contiguous().float()
    d = _lib.OpDesc(task=SEPARABLE, channels=channels, R=img_dim, ratio=1, U_small_host=_fptr(U_s),
                    V_small_host=_fptr(V_s), m_small=m, mult_host=_fptr(mult), pinv_mult_host=_fptr(pinv))
    keep = [U_s, V_s, mult, pinv]
    self._has_lambda = lambda_sing is not None
    if lambda_sing is not None:  # singular value per spectral position for Lambda / Lambda_noise
        lambda_sing = lambda_sing.contiguous().float()
        d.lambda_sing_host = _fptr(lambda_sing)
        keep.append(lambda_sing)
    if U2_s is not None:  # different right-hand factors (Deblurring2D)
        U2_s, V2_s = U2_s.contiguous().float(), V2_s.contiguous().float()
        d.U_small2_host, d.V_small2_host = _fptr(U2_s), _fptr(V2_s)
        keep += [U2_s, V2_s]
    A_functions.__init__(self, d, tuple(keep), device)


def _zero_guarded_inverse(s):
    f = 1.0 / s
    f[s == 0] = 0.0
    return f


class SRConv(A_functions):
    """Separable strided blur (bicubic SR).  A x = U_s ((s s^T) o (V_s^T X V_s)[:m,:m]) U_s^T per channel."""

    def __init__(self, kernel, channels, img_dim, device, stride=1):
        self.channels, self.img_dim, self.ratio = channels, img_dim, stride
        self.xdim = channels * img_dim ** 2
        m = img_dim // stride
        kernel = kernel.detach().cpu().float()
        half = kernel.shape[0] // 2
        A_small = torch.zeros(m, img_dim)
        for i in range(stride // 2, img_dim + stride // 2, stride):
            for j in range(i - half, i + half):
                je = j
                if je < 0:
                    je = -je - 1  # reflective padding
                if je >= img_dim:
                    je = (img_dim - 1) - (je - img_dim)
                A_small[i // stride, je] += kernel[j - i + half]
        U_s, s, V_s = torch.svd(A_small, some=False)
        s = s.clone()
        s[s < 3e-2] = 0
        sing = torch.matmul(s.reshape(m, 1), s.reshape(1, m)).reshape(m * m)
        mult = sing.repeat(channels, 1)
        pinv = _zero_guarded_inverse(sing).repeat(channels, 1)
        _separable(self, U_s, V_s, mult, pinv, channels, img_dim, m, device)


class Deblurring(A_functions):
    """Separable blur with zero padding.  The reference pairs the sorted singular values with the interleaved
    (position, channel) spectral entries through `_singulars.repeat(1, 3)` (functions/svd_operators.py:995-996,
    1006-1014); the multiplier tables below reproduce exactly that pairing."""

    def __init__(self, kernel, channels, img_dim, device, ZERO=3e-2):
        self.channels, self.img_dim = channels, img_dim
        self.xdim = channels * img_dim ** 2
        R = img_dim
        kernel = kernel.detach().cpu().float()
        half = kernel.shape[0] // 2
        A_small = torch.zeros(R, R)
        for i in range(R):
            for j in range(i - half, i + half):
                if 0 <= j < R:
                    A_small[i, j] = kernel[j - i + half]
        U_s, s, V_s = torch.svd(A_small, some=False)
        s_orig = s.clone()
        s = s.clone()
        s[s < ZERO] = 0
        big = torch.matmul(s.reshape(R, 1), s.reshape(1, R)).reshape(R * R)
        big_sorted, perm = big.sort(descending=True)
        full = big_sorted.repeat(1, channels).reshape(-1)  # [S, S, S] concatenated, as in the reference
        mult = torch.empty(channels, R * R)
        mult[:, perm] = full.reshape(R * R, channels).t()
        pinv = torch.empty(channels, R * R)
        pinv[:, perm] = _zero_guarded_inverse(full).reshape(R * R, channels).t()
        # Lambda pairs spectral position q with the un-thresholded product at perm[q], the same for every channel
        # (:957-966, 1021-1033): in image-spectral order that is simply outer(s_orig, s_orig)
        lam_s = torch.matmul(s_orig.reshape(R, 1), s_orig.reshape(1, R)).reshape(R * R)
        _separable(self, U_s, V_s, mult, pinv, channels, R, R, device, lambda_sing=lam_s)


class Deblurring2D(A_functions):
    """Anisotropic blur (functions/svd_operators.py:1094-1165): kernel1 along the rows, kernel2 along the columns, zero
    padding, the same singular-value pairing as Deblurring."""

    def __init__(self, kernel1, kernel2, channels, img_dim, device, ZERO=3e-2):
        self.channels, self.img_dim = channels, img_dim
        self.xdim = channels * img_dim ** 2
        R = img_dim

        def small(kernel):
            kernel = kernel.detach().cpu().float()
            half = kernel.shape[0] // 2
            A = torch.zeros(R, R)
            for i in range(R):
                for j in range(i - half, i + half):
                    if 0 <= j < R:
                        A[i, j] = kernel[j - i + half]
            U, s, V = torch.svd(A, some=False)
            s = s.clone()
            s[s < ZERO] = 0
            return U, s, V

        U1, s1, V1 = small(kernel1)
        U2, s2, V2 = small(kernel2)
        big = torch.matmul(s1.reshape(R, 1), s2.reshape(1, R)).reshape(R * R)
        big_sorted, perm = big.sort(descending=True)
        full = big_sorted.repeat(1, channels).reshape(-1)
        mult = torch.empty(channels, R * R)
        mult[:, perm] = full.reshape(R * R, channels).t()
        pinv = torch.empty(channels, R * R)
        pinv[:, perm] = _zero_guarded_inverse(full).reshape(R * R, channels).t()
        _separable(self, U1, V1, mult, pinv, channels, R, R, device, U2_s=U2, V2_s=V2)


def l1_diff_rows(a, b):
    """Per-sample ||a - b||_1 (Constraint_Function.loss, image_sample.py:325-333)."""
    a2 = a.reshape(a.shape[0], -1).contiguous().float()
    b2 = b.reshape(b.shape[0], -1).contiguous().float()
    out = torch.empty(a2.shape[0], device=a2.device)
    idx = a2.device.index if a2.device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.lib().nlc_l1_diff_rows(_lib.ctx(idx), a2.data_ptr(), b2.data_ptr(), a2.shape[0], a2.shape[1],
                                           out.data_ptr(), _stream()))
    return out
