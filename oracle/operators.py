"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's DDNM operators and of
the constraint projection.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned by tests/test_oracle_vs_reference.py (live against functions/svd_operators.py when /root/reference is
present) and by tests/golden/operators_*.pt.

Every operator is kept in the reference's own spectral form  A = U Sigma V^T,  A^T = V Sigma^T U^T,
A^+ = V Sigma^+ U^T  (functions/svd_operators.py:52-80) with explicit `Vt / V / Ut / U / singulars` steps, so this
file restates the algorithm rather than the closed forms the CUDA kernels use.
  Inpainting :324-359   Colorization :627-667   SuperResolution :479-533   WalshHadamardCS :211-251
  SRConv :851-931       Deblurring :934-1014    Deblurring2D :1094-1165    projection image_sample.py:376-379
DDNM+ (noisy measurements): `A_pinv_eta` :82-91 and the per-class `Lambda` / `Lambda_noise`
(:253-320, :361-439, :464-476, :535-623, :669-736, :1016-1091) are restated once, on the base class, as
  Lambda(v)          = V( lambda o V^T v )
  Lambda_noise(v, e) = V( d1 o P v ) + V( d2 o P e )        P = the re-ordering half of V^T, without its rotation
with the three per-component tables of `ddnm_tables`; `Denoising` :442-476 and the block-wise `CS` :101-160 and
`GeneralA` :173-208 operators follow the same spectral form.
"""
import torch


def ddnm_tables(s, a, sigma_y, sigma_t, eta):
    """The per-spectral-component factors of DDNM+ for zero-padded singular values `s` (one entry per component of
    V^T x): lambda of Eq. 17 and the noise scales d1 (fresh noise), d2 (predicted noise) of Eq. 51, as the reference's
    Lambda / Lambda_noise build them (e.g. functions/svd_operators.py:256-270, 282-303).  `a` and `sigma_t` are fp32
    scalars (0-dim tensors at the reference's call site, functions/svd_ddnm.py:121-132), `sigma_y` and `eta` Python
    floats; every product below is therefore taken in fp32, in the reference's order."""
    a = torch.as_tensor(a, dtype=torch.float32).reshape(())
    st = torch.as_tensor(sigma_t, dtype=torch.float32).reshape(())
    one = torch.ones_like(s)
    root = (1 - eta ** 2) ** 0.5
    inv = 1.0 / s
    inv[s == 0] = 0.0
    lam, d1, d2 = one.clone(), one * st * eta, one * st * root
    if a != 0 and sigma_y != 0:
        thr = a * sigma_y * inv
        below, above, null = st < thr, st > thr, s == 0
        lam = torch.where(below, s * st * root / a / sigma_y, lam)
        resid = st ** 2 - a ** 2 * sigma_y ** 2 * inv ** 2
        d1 = torch.where(above, torch.sqrt(torch.where(above, resid, torch.zeros_like(resid))), d1)
        d2 = torch.where(below | above, torch.zeros_like(d2), d2)
        d1 = torch.where(null, one * st * eta, d1)
        d2 = torch.where(null, one * st * root, d2)
    return lam, d1, d2


class SpectralOp:
    def A(self, x):
        t = self.Vt(x)
        s = self.singulars()
        return self.U(s * t[:, :s.shape[0]])

    def At(self, y):
        t = self.Ut(y)
        s = self.singulars()
        return self.V(self.add_zeros(s * t[:, :s.shape[0]]))

    def A_pinv(self, y):
        t = self.Ut(y).clone()
        s = self.singulars()
        f = 1.0 / s
        f[s == 0] = 0.0
        t[:, :s.shape[0]] = t[:, :s.shape[0]] * f
        return self.V(self.add_zeros(t))

    def project(self, x0, y):
        b = x0.shape[0]
        return x0 - self.A_pinv(self.A(x0.reshape(b, -1)) - y.reshape(b, -1)).reshape(x0.shape)

    # ---- DDNM+ -----------------------------------------------------------------------------------------------
    def A_pinv_eta(self, y, eta):
        """Regularised pseudo-inverse V diag(s / (s^2 + eta)) U^T (functions/svd_operators.py:82-91)."""
        t = self.Ut(y).clone()
        s = self.singulars()
        t[:, :s.shape[0]] = t[:, :s.shape[0]] * (s / (s * s + eta))
        return self.V(self.add_zeros(t))

    def lambda_singulars(self):
        """Singular value of every component of Vt(x), zero on the null space."""
        s = self.singulars()
        out = torch.zeros(self.C * self.R * self.R)
        out[:s.shape[0]] = s
        return out

    def reorder(self, x):
        """P x: the re-ordering half of V^T (image entries put in the spectral order, no rotation): what Lambda_noise
        applies to the noise vectors.  Equals Vt where V is a permutation."""
        return self.Vt(x)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):
        lam, _, _ = ddnm_tables(self.lambda_singulars(), a, sigma_y, sigma_t, eta)
        return self.V(lam * self.Vt(v.reshape(v.shape[0], -1)))

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):
        _, d1, d2 = ddnm_tables(self.lambda_singulars(), a, sigma_y, sigma_t, eta)
        b = v.shape[0]
        return self.V(d1 * self.reorder(v.reshape(b, -1))) + self.V(d2 * self.reorder(eps.reshape(b, -1)))


class Inpainting(SpectralOp):
    def __init__(self, channels, R, missing):
        self.C, self.R = channels, R
        n = channels * R * R
        self.missing = missing.long()
        mask = torch.ones(n, dtype=torch.bool)
        mask[self.missing] = False
        self.kept = torch.nonzero(mask).reshape(-1)

    def _interleave(self, x):  # NCHW flat -> pixel-major (pixel*C + c)
        b = x.shape[0]
        return x.reshape(b, self.C, -1).transpose(1, 2).reshape(b, -1)

    def Vt(self, x):
        t = self._interleave(x)
        return torch.cat([t[:, self.kept], t[:, self.missing]], dim=1)

    def V(self, v):
        b = v.shape[0]
        out = torch.zeros_like(v)
        out[:, self.kept] = v[:, :self.kept.numel()]
        out[:, self.missing] = v[:, self.kept.numel():]
        return out.reshape(b, -1, self.C).transpose(1, 2).reshape(b, -1)

    def U(self, v):
        return v.reshape(v.shape[0], -1)

    Ut = U

    def singulars(self):
        return torch.ones(self.kept.numel())

    def add_zeros(self, v):
        out = torch.zeros(v.shape[0], self.C * self.R * self.R)
        out[:, :v.shape[1]] = v
        return out


class _Needle(SpectralOp):
    """Shared by Colorization (needle = the 3 channels of a pixel) and SuperResolution (needle = an r x r patch):
    a 1 x K matrix per needle, SVD'ed once."""

    def _svd(self, row):
        self.Us, self.Ss, self.Vs = torch.svd(torch.Tensor([row]), some=False)

    def U(self, v):
        return self.Us[0, 0] * v.reshape(v.shape[0], -1)

    Ut = U


class Colorization(_Needle):
    def __init__(self, R):
        self.C, self.R = 3, R
        self._svd([0.3333, 0.3334, 0.3333])

    def Vt(self, x):
        b = x.shape[0]
        needles = x.reshape(b, 3, -1).transpose(1, 2)  # [b, pixel, c]
        rot = torch.matmul(self.Vs.t(), needles.reshape(-1, 3, 1)).reshape(b, -1, 3)
        return rot.transpose(1, 2).reshape(b, -1)  # component-major

    def V(self, v):
        b = v.shape[0]
        needles = v.reshape(b, 3, -1).transpose(1, 2)
        rot = torch.matmul(self.Vs, needles.reshape(-1, 3, 1)).reshape(b, -1, 3)
        return rot.transpose(1, 2).reshape(b, -1)

    def singulars(self):
        return self.Ss.repeat(self.R * self.R)

    def reorder(self, x):  # channel k of a pixel stands in for component k (functions/svd_operators.py:698-699)
        return x.reshape(x.shape[0], -1)

    def add_zeros(self, v):
        out = torch.zeros(v.shape[0], 3 * self.R * self.R)
        out[:, :self.R * self.R] = v.reshape(v.shape[0], -1)
        return out


class SuperResolution(_Needle):
    def __init__(self, channels, R, ratio):
        self.C, self.R, self.r, self.yd = channels, R, ratio, R // ratio
        self._svd([1 / ratio ** 2] * ratio ** 2)

    def _patches(self, x):
        b, r, yd = x.shape[0], self.r, self.yd
        return x.reshape(b, self.C, yd, r, yd, r).permute(0, 1, 2, 4, 3, 5).reshape(b, self.C, yd * yd, r * r)

    def _spectral_order(self, comp):  # [b, C, patches, r*r] -> component 0 of every patch first, the rest interleaved
        b, r, yd = comp.shape[0], self.r, self.yd
        out = torch.zeros(b, self.C * self.R * self.R)
        n0 = self.C * yd * yd
        out[:, :n0] = comp[..., 0].reshape(b, n0)
        for k in range(r * r - 1):
            out[:, n0 + k::r * r - 1] = comp[..., k + 1].reshape(b, n0)
        return out

    def Vt(self, x):
        b, r, yd = x.shape[0], self.r, self.yd
        img = self._patches(x)
        rot = torch.matmul(self.Vs.t(), img.reshape(-1, r * r, 1)).reshape(b, self.C, yd * yd, r * r)
        return self._spectral_order(rot)

    def reorder(self, x):  # patch entry k stands in for component k (functions/svd_operators.py:575-581)
        return self._spectral_order(self._patches(x))

    def V(self, v):
        b, r, yd = v.shape[0], self.r, self.yd
        n0 = self.C * yd * yd
        comp = torch.zeros(b, self.C, yd * yd, r * r)
        comp[..., 0] = v[:, :n0].reshape(b, self.C, -1)
        for k in range(r * r - 1):
            comp[..., k + 1] = v[:, n0 + k::r * r - 1].reshape(b, self.C, -1)
        rot = torch.matmul(self.Vs, comp.reshape(-1, r * r, 1)).reshape(b, self.C, yd, yd, r, r)
        return rot.permute(0, 1, 2, 4, 3, 5).reshape(b, -1)

    def singulars(self):
        return self.Ss.repeat(self.C * self.yd * self.yd)

    def add_zeros(self, v):
        v = v.reshape(v.shape[0], -1)
        out = torch.zeros(v.shape[0], v.shape[1] * self.r ** 2)
        out[:, :v.shape[1]] = v
        return out


class WalshHadamardCS(SpectralOp):
    def __init__(self, channels, R, ratio, perm):
        self.C, self.R, self.ratio, self.perm = channels, R, ratio, perm.long()

    def fwht(self, v):
        b, n = v.shape[0], self.R * self.R
        a = v.reshape(b, self.C, n).clone()
        h = 1
        while h < n:
            a = a.reshape(b, self.C, -1, 2 * h)
            lo, hi = a[..., :h].clone(), a[..., h:].clone()
            a = torch.cat([lo + hi, lo - hi], dim=-1)
            h *= 2
        return a.reshape(b, self.C, n) / self.R

    def Vt(self, x):
        b = x.shape[0]
        return self.fwht(x)[:, :, self.perm].transpose(1, 2).reshape(b, -1)

    def V(self, v):
        b = v.shape[0]
        t = torch.zeros(b, self.C, self.R * self.R)
        t[:, :, self.perm] = v.reshape(b, -1, self.C).transpose(1, 2)
        return self.fwht(t).reshape(b, -1)

    def reorder(self, x):  # functions/svd_operators.py:277-280: the permutation without the transform
        b = x.shape[0]
        return x.reshape(b, self.C, -1)[:, :, self.perm].transpose(1, 2).reshape(b, -1)

    def U(self, v):
        return v.reshape(v.shape[0], -1)

    Ut = U

    def singulars(self):
        return torch.ones(self.C * self.R * self.R // self.ratio)

    def add_zeros(self, v):
        out = torch.zeros(v.shape[0], self.C * self.R * self.R)
        out[:, :v.shape[1]] = v.reshape(v.shape[0], -1)
        return out


class _Separable(SpectralOp):
    def _left(self, M, v, dim):
        return torch.matmul(M, v.reshape(-1, dim, dim))

    def _sandwich(self, L, img, Rm, b):
        # L @ img @ Rm per (sample, channel); img given as [b, C, rows, cols]
        return torch.matmul(torch.matmul(L, img.reshape(b * self.C, img.shape[-2], img.shape[-1])), Rm)


class _NoLambda:
    """The reference defines no Lambda / Lambda_noise for these classes (the base raises, :93-97)."""

    def Lambda(self, *a, **k):
        raise NotImplementedError()

    def Lambda_noise(self, *a, **k):
        raise NotImplementedError()


class SRConv(_NoLambda, _Separable):
    def __init__(self, kernel, channels, R, stride):
        self.C, self.R, self.ratio = channels, R, stride
        m = self.m = R // stride
        A_small = torch.zeros(m, R)
        half = kernel.shape[0] // 2
        for i in range(stride // 2, R + stride // 2, stride):
            for j in range(i - half, i + half):
                je = -j - 1 if j < 0 else ((R - 1) - (j - R) if j >= R else j)
                A_small[i // stride, je] += kernel[j - i + half]
        self.Us, s, self.Vs = torch.svd(A_small, some=False)
        s[s < 3e-2] = 0
        self.s_big = torch.outer(s, s).reshape(m * m)
        first = [R * i + j for i in range(m) for j in range(m)]
        rest = [R * i + j for i in range(m) for j in range(m, R)]
        self.perm = torch.tensor(first + rest, dtype=torch.long)

    def Vt(self, x):
        b, R = x.shape[0], self.R
        t = self._sandwich(self.Vs.t(), x.reshape(b, self.C, R, R), self.Vs, b).reshape(b, self.C, -1)
        head = t[:, :, self.perm]
        t = torch.cat([head, t[:, :, self.perm.numel():]], dim=2)
        return t.transpose(1, 2).reshape(b, -1)

    def V(self, v):
        b, R, P = v.shape[0], self.R, self.perm.numel()
        vv = v.reshape(b, R * R, self.C)
        t = torch.zeros(b, R * R, self.C)
        t[:, self.perm, :] = vv[:, :P, :]
        t[:, P:, :] = vv[:, P:, :]
        img = t.transpose(1, 2).reshape(b, self.C, R, R)
        return self._sandwich(self.Vs, img, self.Vs.t(), b).reshape(b, -1)

    def U(self, v):
        b, m = v.shape[0], self.m
        img = v.reshape(b, m * m, self.C).transpose(1, 2).reshape(b, self.C, m, m)
        return self._sandwich(self.Us, img, self.Us.t(), b).reshape(b, -1)

    def Ut(self, y):
        b, m = y.shape[0], self.m
        t = self._sandwich(self.Us.t(), y.reshape(b, self.C, m, m), self.Us, b).reshape(b, self.C, -1)
        return t.transpose(1, 2).reshape(b, -1)

    def singulars(self):
        return self.s_big.repeat_interleave(3).reshape(-1)

    def add_zeros(self, v):
        v = v.reshape(v.shape[0], -1)
        out = torch.zeros(v.shape[0], v.shape[1] * self.ratio ** 2)
        out[:, :v.shape[1]] = v
        return out


class Deblurring(_Separable):
    def __init__(self, kernel, channels, R, zero=3e-2):
        self.C, self.R = channels, R
        A_small = torch.zeros(R, R)
        half = kernel.shape[0] // 2
        for i in range(R):
            for j in range(i - half, i + half):
                if 0 <= j < R:
                    A_small[i, j] = kernel[j - i + half]
        self.Us, s, self.Vs = torch.svd(A_small, some=False)
        s_orig = s.clone()
        s[s < zero] = 0
        self.s_sorted, self.perm = torch.outer(s, s).reshape(R * R).sort(descending=True)
        self.s_orig_sorted = torch.outer(s_orig, s_orig).reshape(R * R)[self.perm]  # un-thresholded (:957-966)

    def lambda_singulars(self):  # Lambda pairs position q with the un-thresholded value, for all channels (:1021,1033)
        return self.s_orig_sorted.repeat_interleave(self.C)

    def reorder(self, x):  # :1045-1049
        b = x.shape[0]
        return x.reshape(b, self.C, -1)[:, :, self.perm].transpose(1, 2).reshape(b, -1)

    def _to_spec(self, M, x):
        b, R = x.shape[0], self.R
        t = self._sandwich(M.t(), x.reshape(b, self.C, R, R), M, b).reshape(b, self.C, -1)
        return t[:, :, self.perm].transpose(1, 2).reshape(b, -1)

    def _from_spec(self, M, v):
        b, R = v.shape[0], self.R
        t = torch.zeros(b, R * R, self.C)
        t[:, self.perm, :] = v.reshape(b, R * R, self.C)
        return self._sandwich(M, t.transpose(1, 2).reshape(b, self.C, R, R), M.t(), b).reshape(b, -1)

    def Vt(self, x):
        return self._to_spec(self.Vs, x)

    def V(self, v):
        return self._from_spec(self.Vs, v)

    def Ut(self, y):
        return self._to_spec(self.Us, y)

    def U(self, v):
        return self._from_spec(self.Us, v)

    def singulars(self):
        return self.s_sorted.repeat(1, 3).reshape(-1)  # concatenated, not interleaved: as in the reference

    def add_zeros(self, v):
        return v.reshape(v.shape[0], -1)


class Deblurring2D(_NoLambda, _Separable):
    """Anisotropic blur (functions/svd_operators.py:1094-1165): kernel1 acts on the rows (left matrices U1, V1), kernel2 on
    the columns (right matrices U2, V2); singular values s1 (x) s2 sorted descending, pairing as in Deblurring."""

    def __init__(self, kernel1, kernel2, channels, R, zero=3e-2):
        self.C, self.R = channels, R

        def small(kernel):
            A = torch.zeros(R, R)
            half = kernel.shape[0] // 2
            for i in range(R):
                for j in range(i - half, i + half):
                    if 0 <= j < R:
                        A[i, j] = kernel[j - i + half]
            return torch.svd(A, some=False)

        self.U1, s1, self.V1 = small(kernel1)
        self.U2, s2, self.V2 = small(kernel2)
        s1[s1 < zero] = 0
        s2[s2 < zero] = 0
        self.s_sorted, self.perm = torch.matmul(s1.reshape(R, 1), s2.reshape(1, R)).reshape(R * R).sort(descending=True)

    def _to_spec(self, L, Rm, x):
        b, R = x.shape[0], self.R
        t = self._sandwich(L.t(), x.reshape(b, self.C, R, R), Rm, b).reshape(b, self.C, -1)
        return t[:, :, self.perm].transpose(1, 2).reshape(b, -1)

    def _from_spec(self, L, Rm, v):
        b, R = v.shape[0], self.R
        t = torch.zeros(b, R * R, self.C)
        t[:, self.perm, :] = v.reshape(b, R * R, self.C)
        return self._sandwich(L, t.transpose(1, 2).reshape(b, self.C, R, R), Rm.t(), b).reshape(b, -1)

    def Vt(self, x):
        return self._to_spec(self.V1, self.V2, x)

    def V(self, v):
        return self._from_spec(self.V1, self.V2, v)

    def Ut(self, y):
        return self._to_spec(self.U1, self.U2, y)

    def U(self, v):
        return self._from_spec(self.U1, self.U2, v)

    def singulars(self):
        return self.s_sorted.repeat(1, 3).reshape(-1)

    def add_zeros(self, v):
        return v.reshape(v.shape[0], -1)


class Denoising(SpectralOp):
    """A = I (functions/svd_operators.py:442-476): all singular values are 1, so Lambda / Lambda_noise are scalars."""

    def __init__(self, channels, R):
        self.C, self.R = channels, R

    def Vt(self, x):
        return x.reshape(x.shape[0], -1).clone()

    V = U = Ut = add_zeros = Vt

    def singulars(self):
        return torch.ones(self.C * self.R * self.R)

    def Lambda(self, v, a, sigma_y, sigma_t, eta):  # :464-469 (no a / sigma_y zero guard there)
        a, st = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(sigma_t, dtype=torch.float32)
        if st < a * sigma_y:
            return v * (st * (1 - eta ** 2) ** 0.5 / a / sigma_y).item()
        return v

    def Lambda_noise(self, v, a, sigma_y, sigma_t, eta, eps):  # :471-476
        a, st = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(sigma_t, dtype=torch.float32)
        if st >= a * sigma_y:
            return v * torch.sqrt(st ** 2 - a ** 2 * sigma_y ** 2).item()
        return v * st * eta


class CS(_NoLambda, SpectralOp):
    """Block-wise compressed sensing (functions/svd_operators.py:101-160): every E x E patch (E = 32) is rotated by the
    transpose of an E^2 x E^2 orthogonal `V_small` (the right singular vectors of a random matrix, drawn by the
    constructor there; passed in here) and its first `cs_size = int(E^2 ratio)` components are the measurement; all
    singular values are 1.  Spectral order: the kept components of all patches first, then the dropped ones."""

    def __init__(self, channels, R, ratio, V_small, E=32):
        self.C, self.R, self.E, self.yd = channels, R, E, R // E
        self.Vs = V_small
        self.cs = int(E * E * ratio)

    def _patches(self, x):
        b, E, yd = x.shape[0], self.E, self.yd
        return x.reshape(b, self.C, yd, E, yd, E).permute(0, 1, 2, 4, 3, 5).reshape(b, self.C, yd * yd, E * E)

    def Vt(self, x):
        b, E = x.shape[0], self.E
        rot = torch.matmul(self.Vs.t(), self._patches(x).reshape(-1, E * E, 1)).reshape(b, self.C, -1, E * E)
        return torch.cat([rot[..., :self.cs].reshape(b, -1), rot[..., self.cs:].reshape(b, -1)], dim=1)

    def V(self, v):
        b, E, yd = v.shape[0], self.E, self.yd
        n_kept = self.C * yd * yd * self.cs
        comp = torch.cat([v[:, :n_kept].reshape(b, self.C * yd * yd, self.cs),
                          v[:, n_kept:].reshape(b, self.C * yd * yd, E * E - self.cs)], dim=2)
        rot = torch.matmul(self.Vs, comp.reshape(-1, E * E, 1)).reshape(b, self.C, yd, yd, E, E)
        return rot.permute(0, 1, 2, 4, 3, 5).reshape(b, -1)

    def U(self, v):
        return v.reshape(v.shape[0], -1)

    Ut = U

    def singulars(self):
        return torch.ones(self.cs * self.C * self.yd * self.yd)

    def add_zeros(self, v):
        out = torch.zeros(v.shape[0], self.C * self.R * self.R)
        out[:, :v.shape[1]] = v.reshape(v.shape[0], -1)
        return out


class GeneralA(_NoLambda, SpectralOp):
    """Dense SVD of an arbitrary small matrix A [ny, nx] (functions/svd_operators.py:173-208); singular values below
    1e-3 are dropped."""

    def __init__(self, A):
        self.Um, self.s, self.Vm = torch.svd(A, some=False)
        self.s = self.s.clone()
        self.s[self.s < 1e-3] = 0

    @staticmethod
    def _mv(M, v):
        return torch.matmul(M, v.reshape(v.shape[0], -1, 1)).reshape(v.shape[0], M.shape[0])

    def V(self, v):
        return self._mv(self.Vm, v)

    def Vt(self, v):
        return self._mv(self.Vm.t(), v)

    def U(self, v):
        return self._mv(self.Um, v)

    def Ut(self, v):
        return self._mv(self.Um.t(), v)

    def singulars(self):
        return self.s

    def add_zeros(self, v):
        out = torch.zeros(v.shape[0], self.Vm.shape[0])
        out[:, :self.Um.shape[0]] = v.reshape(v.shape[0], -1)
        return out


def hadamard_basis(n, seed):
    """A reproducible n x n orthogonal matrix with entries +-1/sqrt(n) (exact in fp32 for n a power of 4): a Sylvester
    Hadamard matrix with seeded row signs and column order.  Stands in for CS's random right singular vectors in the
    golden fixtures, where a 1024 x 1024 matrix is too large to store and an SVD is not reproducible across machines."""
    H = torch.ones(1, 1)
    while H.shape[0] < n:
        H = torch.cat([torch.cat([H, H], dim=1), torch.cat([H, -H], dim=1)], dim=0)
    g = torch.Generator().manual_seed(seed)
    signs = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    return (signs[:, None] * H[:, torch.randperm(n, generator=g)]) / n ** 0.5


def aniso_kernels():
    """src/constraint_functions.py:280-292: 9 taps, sigma 1 (rows) and sigma 20 (columns), each normalised."""
    def k(sigma):
        pdf = lambda x: torch.exp(torch.Tensor([-0.5 * (x / sigma) ** 2]))
        kk = torch.Tensor([pdf(i) for i in range(-4, 5)])
        return kk / kk.sum()
    return k(1), k(20)


def constraint_inv_transform(op, deg, y, channels, R):
    """Constraint_Function.inv_transform for proj='svd' (image_sample.py:312-323)."""
    b = y.shape[0]
    Apy = op.A_pinv(y).view(b, channels, R, R)
    if deg[:6] == "deblur":
        Apy = y.view(b, channels, R, R)
    elif deg == "colorization":
        Apy = y.view(b, 1, R, R).repeat(1, 3, 1, 1)
    elif deg == "inpainting":
        Apy = Apy + op.A_pinv(op.A(torch.ones_like(Apy).reshape(b, -1))).reshape(*Apy.shape) - 1
    return Apy


def constraint_loss(op, deg, x, y, channels, R):
    """Constraint_Function.loss (image_sample.py:325-333): per-sample L1 of (A x - y) and of (inv_transform(y) - x)."""
    b = x.shape[0]
    y_hat = op.A(x.reshape(b, -1))
    x_hat = constraint_inv_transform(op, deg, y, channels, R)
    fwd = torch.linalg.vector_norm(y_hat - y.reshape(b, -1), ord=1, dim=1)
    bwd = torch.linalg.vector_norm((x_hat - x).reshape(b, -1), ord=1, dim=1)
    return fwd, bwd


def bicubic_kernel(factor):
    """src/constraint_functions.py:252-269."""
    import numpy as np

    def cubic(x, a=-0.5):
        ax = abs(x)
        if ax <= 1:
            return (a + 2) * ax ** 3 - (a + 3) * ax ** 2 + 1
        if 1 < ax < 2:
            return a * ax ** 3 - 5 * a * ax ** 2 + 8 * a * ax - 4 * a
        return 0

    k = np.zeros(factor * 4)
    for i in range(factor * 4):
        k[i] = cubic((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
    k = torch.from_numpy(k / np.sum(k)).float()
    return k / k.sum()


def gauss_kernel(sigma=10):
    """src/constraint_functions.py:274-279."""
    pdf = lambda x: torch.exp(torch.Tensor([-0.5 * (x / sigma) ** 2]))
    k = torch.Tensor([pdf(-2), pdf(-1), pdf(0), pdf(1), pdf(2)])
    return k / k.sum()
