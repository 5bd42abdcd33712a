"""Synthetic, seeded state_dicts with exactly the key names and shapes of the reference modules, so the benchmark (random-init
weights of the named architectures: there is no network for checkpoints) and the parity tests need neither checkpoints
nor /root/reference.  Lives in the product package because bench.py's GPU arm builds its networks from it; the oracle and
the tests reach it through the alias `oracle/weights.py`.

Key layouts follow the reference constructors: src/unet_ddim.py:214-321 (UNetModel) and :493-519 (SigmaModel).
tests/test_oracle_vs_reference.py checks names and shapes against the real modules when the reference is present.
Initialisation follows SURVEY §8(d): no all-zero parameter, non-trivial BatchNorm running statistics.
"""
import torch


class _Init:
    def __init__(self, seed, shapes_only=False):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = {}
        self.shapes_only = shapes_only  # record torch.Size instead of drawing (layout checks of the big nets)

    def randn(self, *shape):
        if self.shapes_only:
            return torch.empty(*shape, device="meta")
        return torch.randn(*shape, generator=self.g)

    def conv(self, name, cin, cout, k, gain=1.0):
        fan_in = cin * k * k
        self.sd[name + ".weight"] = self.randn(cout, cin, k, k) * (gain / fan_in ** 0.5)
        self.sd[name + ".bias"] = self.randn(cout) * 0.02

    def linear(self, name, cin, cout, gain=1.0):
        self.sd[name + ".weight"] = self.randn(cout, cin) * (gain / cin ** 0.5)
        self.sd[name + ".bias"] = self.randn(cout) * 0.02

    def norm(self, name, c):
        self.sd[name + ".weight"] = 1.0 + 0.1 * self.randn(c)
        self.sd[name + ".bias"] = 0.05 * self.randn(c)


def _resblock(I, p, cin, cout, temb_ch=None):
    I.norm(p + "norm1", cin)
    I.conv(p + "conv1", cin, cout, 3)
    if temb_ch is not None:
        I.linear(p + "temb_proj", temb_ch, cout)
    I.norm(p + "norm2", cout)
    I.conv(p + "conv2", cout, cout, 3, gain=0.5)
    if cin != cout:
        I.conv(p + "nin_shortcut", cin, cout, 1)


def _attn(I, p, c):
    I.norm(p + "norm", c)
    for n in ("q", "k", "v"):
        I.conv(p + n, c, c, 1)
    I.conv(p + "proj_out", c, c, 1, gain=0.5)


def ddim_unet_state_dict(image_size, in_channels, model_channels, out_channels, num_res_blocks,
                         attention_resolutions, channel_mult, seed=0):
    """Same keys/shapes as src.unet_ddim.UNetModel(...).state_dict()."""
    I = _Init(seed)
    ch, temb_ch = model_channels, 4 * model_channels
    I.linear("temb.dense.0", ch, temb_ch)
    I.linear("temb.dense.1", temb_ch, temb_ch)
    I.conv("conv_in", in_channels, ch, 3)
    res = image_size
    in_mult = (1,) + tuple(channel_mult)
    L = len(channel_mult)
    block_in = None
    for lv in range(L):
        block_in = ch * in_mult[lv]
        block_out = ch * channel_mult[lv]
        for ib in range(num_res_blocks):
            _resblock(I, "down.%d.block.%d." % (lv, ib), block_in, block_out, temb_ch)
            block_in = block_out
            if res in attention_resolutions:
                _attn(I, "down.%d.attn.%d." % (lv, ib), block_in)
        if lv != L - 1:
            I.conv("down.%d.downsample.conv" % lv, block_in, block_in, 3)
            res //= 2
    _resblock(I, "mid.block_1.", block_in, block_in, temb_ch)
    _attn(I, "mid.attn_1.", block_in)
    _resblock(I, "mid.block_2.", block_in, block_in, temb_ch)
    for lv in reversed(range(L)):
        block_out = ch * channel_mult[lv]
        skip_in = ch * channel_mult[lv]
        for ib in range(num_res_blocks + 1):
            if ib == num_res_blocks:
                skip_in = ch * in_mult[lv]
            _resblock(I, "up.%d.block.%d." % (lv, ib), block_in + skip_in, block_out, temb_ch)
            block_in = block_out
            if res in attention_resolutions:
                _attn(I, "up.%d.attn.%d." % (lv, ib), block_in)
        if lv != 0:
            I.conv("up.%d.upsample.conv" % lv, block_in, block_in, 3)
            res *= 2
    I.norm("norm_out", block_in)
    I.conv("conv_out", block_in, out_channels, 3)
    return I.sd


def ddim_sigma_state_dict(dim, channels, n_blocks, seed=1, fc_dim=128):
    """Same keys/shapes as src.unet_ddim.SigmaModel(dim, channels, n_blocks).state_dict()."""
    I = _Init(seed)
    idx = 0
    d = dim
    for i in range(n_blocks):
        if d % 2 != 0:
            d += 1
        idx += 1  # ConstantPad2d / Identity
        _resblock(I, "down_layer.%d." % idx, channels, channels)
        idx += 1
        if i == 0:
            _attn(I, "down_layer.%d." % idx, channels)
            idx += 1
        I.conv("down_layer.%d.conv" % idx, channels, channels, 3)
        idx += 1
        d //= 2
    hidden = channels * d * d
    I.linear("fc_layer.1", hidden, fc_dim)
    I.sd["fc_layer.2.weight"] = 1.0 + 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.bias"] = 0.05 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_mean"] = 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_var"] = 0.5 + torch.rand(fc_dim, generator=I.g)
    I.sd["fc_layer.2.num_batches_tracked"] = torch.tensor(0)
    I.linear("final_mlp", fc_dim, 1, gain=0.3)
    return I.sd


# The benchmark / parity configurations of BASELINE.json
CONFIGS = {
    # c1: CIFAR-10-shaped unet_ddim, 32x32
    "c1": dict(unet=dict(image_size=32, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                         attention_resolutions=(16,), channel_mult=(1, 2, 2, 2)),
               sigma=dict(dim=4, channels=256, n_blocks=2)),
    # c2: CelebA-64 unet_ddim
    "c2": dict(unet=dict(image_size=64, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                         attention_resolutions=(16,), channel_mult=(1, 2, 2, 2, 4)),
               sigma=dict(dim=4, channels=512, n_blocks=2)),
    # tiny: fast CPU-side shape for unit tests (same topology, narrow)
    "tiny": dict(unet=dict(image_size=16, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=1,
                           attention_resolutions=(8,), channel_mult=(1, 2)),
                 sigma=dict(dim=8, channels=256, n_blocks=2)),
}


# ------------------------------------------------------------------------------------------------ ADM (src/unet_adm.py)
def _adm_res(I, p, cin, cout, emb_ch=None, scale_shift=False):
    I.norm(p + "in_layers.0", cin)
    I.conv(p + "in_layers.2", cin, cout, 3)
    if emb_ch is not None:
        I.linear(p + "emb_layers.1", emb_ch, 2 * cout if scale_shift else cout, gain=0.5)
    I.norm(p + "out_layers.0", cout)
    I.conv(p + "out_layers.3", cout, cout, 3, gain=0.5)  # zero_module in the reference: re-drawn
    if cin != cout:
        I.conv(p + "skip_connection", cin, cout, 1)


def _adm_attn(I, p, c):
    I.norm(p + "norm", c)
    I.sd[p + "qkv.weight"] = I.randn(3 * c, c, 1) / c ** 0.5
    I.sd[p + "qkv.bias"] = I.randn(3 * c) * 0.02
    I.sd[p + "proj_out.weight"] = I.randn(c, c, 1) * (0.5 / c ** 0.5)
    I.sd[p + "proj_out.bias"] = I.randn(c) * 0.02


def adm_unet_state_dict(image_size, model_channels, num_res_blocks, channel_mult, attention_resolutions,
                        out_channels=6, use_scale_shift_norm=True, resblock_updown=True, seed=0, shapes_only=False,
                        **_):
    """Same keys/shapes as src.unet_adm.UNetModel(...).state_dict() (attention_resolutions are the reference's
    downsample factors `ds`, src/script_util.py:170-172)."""
    I = _Init(seed, shapes_only)
    mc, emb = model_channels, 4 * model_channels
    I.linear("time_embed.0", mc, emb)
    I.linear("time_embed.2", emb, emb)
    ch = int(channel_mult[0] * mc)
    I.conv("input_blocks.0.0", 3, ch, 3)
    chans, ds, idx = [ch], 1, 1
    L = len(channel_mult)
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            _adm_res(I, "input_blocks.%d.0." % idx, ch, int(mult * mc), emb, use_scale_shift_norm)
            ch = int(mult * mc)
            if ds in attention_resolutions:
                _adm_attn(I, "input_blocks.%d.1." % idx, ch)
            chans.append(ch)
            idx += 1
        if level != L - 1:
            if resblock_updown:
                _adm_res(I, "input_blocks.%d.0." % idx, ch, ch, emb, use_scale_shift_norm)
            else:
                I.conv("input_blocks.%d.0.op" % idx, ch, ch, 3)
            chans.append(ch)
            idx += 1
            ds *= 2
    _adm_res(I, "middle_block.0.", ch, ch, emb, use_scale_shift_norm)
    _adm_attn(I, "middle_block.1.", ch)
    _adm_res(I, "middle_block.2.", ch, ch, emb, use_scale_shift_norm)
    idx = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            _adm_res(I, "output_blocks.%d.0." % idx, ch + ich, int(mc * mult), emb, use_scale_shift_norm)
            ch = int(mc * mult)
            j = 1
            if ds in attention_resolutions:
                _adm_attn(I, "output_blocks.%d.%d." % (idx, j), ch)
                j += 1
            if level and i == num_res_blocks:
                if resblock_updown:
                    _adm_res(I, "output_blocks.%d.%d." % (idx, j), ch, ch, emb, use_scale_shift_norm)
                else:
                    I.conv("output_blocks.%d.%d.conv" % (idx, j), ch, ch, 3)
                ds //= 2
            idx += 1
    I.norm("out.0", ch)
    I.conv("out.2", ch, out_channels, 3, gain=0.5)
    if shapes_only:
        return {k: v.shape for k, v in I.sd.items()}
    return I.sd


def adm_sigma_state_dict(dim, channels, n_blocks, seed=1, fc_dim=128):
    """Same keys/shapes as src.unet_adm.SigmaModel(dim, channels, n_blocks).state_dict()."""
    I = _Init(seed)
    idx, d = 0, dim
    for i in range(n_blocks):
        if d % 2 != 0:
            d += 1
        idx += 1
        _adm_res(I, "down_layer.%d." % idx, channels, channels)
        idx += 1
        if i == 0:
            _adm_attn(I, "down_layer.%d." % idx, channels)
            idx += 1
        I.conv("down_layer.%d.op" % idx, channels, channels, 3)
        idx += 1
        d //= 2
    I.linear("fc_layer.1", channels * d * d, fc_dim)
    I.sd["fc_layer.2.weight"] = 1.0 + 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.bias"] = 0.05 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_mean"] = 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_var"] = 0.5 + torch.rand(fc_dim, generator=I.g)
    I.sd["fc_layer.2.num_batches_tracked"] = torch.tensor(0)
    I.linear("final_mlp", fc_dim, 1, gain=0.3)
    return I.sd


ADM_CONFIGS = {
    # c4/c5: ImageNet-256 ADM (256x256_diffusion_uncond): attention at 32/16/8 -> ds 8,16,32
    "adm256": dict(image_size=256, model_channels=256, num_res_blocks=2, channel_mult=(1, 1, 2, 2, 4, 4),
                   attention_resolutions=(8, 16, 32), num_head_channels=64, num_heads=4, out_channels=6,
                   use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False,
                   sigma=dict(dim=8, channels=1024, n_blocks=2)),
    # same topology, two levels, for unit tests
    "adm_tiny": dict(image_size=32, model_channels=128, num_res_blocks=1, channel_mult=(1, 2),
                     attention_resolutions=(2,), num_head_channels=64, num_heads=4, out_channels=6,
                     use_scale_shift_norm=True, resblock_updown=True, use_new_attention_order=False,
                     sigma=dict(dim=16, channels=256, n_blocks=2)),
    # no scale-shift, conv resampling, new attention order, fixed head count: the other code paths
    "adm_alt": dict(image_size=32, model_channels=128, num_res_blocks=1, channel_mult=(1, 2),
                    attention_resolutions=(1, 2), num_head_channels=-1, num_heads=2, out_channels=3,
                    use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=True,
                    sigma=dict(dim=16, channels=256, n_blocks=2)),
}


# ------------------------------------------------------------------------------------------ EDM (src/edm_networks.py)
def _edm_block(I, p, cin, cout, emb_ch=None, attention=False, up=False, down=False):
    I.norm(p + "norm0", cin)
    I.conv(p + "conv0", cin, cout, 3)
    if up or down:
        I.sd[p + "conv0.resample_filter"] = torch.full((1, 1, 2, 2), 0.25)
    if emb_ch is not None:
        I.linear(p + "affine", emb_ch, cout, gain=0.5)
    I.norm(p + "norm1", cout)
    I.conv(p + "conv1", cout, cout, 3, gain=0.5)  # init_weight 1e-5 in the reference: re-drawn
    if cin != cout or up or down:
        I.conv(p + "skip", cin, cout, 1)
        if up or down:
            I.sd[p + "skip.resample_filter"] = torch.full((1, 1, 2, 2), 0.25)
    if attention:
        I.norm(p + "norm2", cout)
        I.conv(p + "qkv", cout, 3 * cout, 1)
        I.conv(p + "proj", cout, cout, 1, gain=0.5)


def edm_unet_state_dict(img_resolution, in_channels, out_channels, model_channels, channel_mult, num_blocks,
                        attn_resolutions, channel_mult_emb=4, seed=0, **_):
    """Same keys/shapes as src.edm_networks.SongUNet(...) (second definition, DDPM++ configuration)."""
    I = _Init(seed)
    emb = model_channels * channel_mult_emb
    I.linear("map_layer0", model_channels, emb)
    I.linear("map_layer1", emb, emb)
    cout = in_channels
    skips = []
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            cin, cout = cout, model_channels
            I.conv("enc.%dx%d_conv" % (res, res), cin, cout, 3)
        else:
            _edm_block(I, "enc.%dx%d_down." % (res, res), cout, cout, emb, down=True)
        skips.append(cout)
        for idx in range(num_blocks):
            cin, cout = cout, model_channels * mult
            _edm_block(I, "enc.%dx%d_block%d." % (res, res, idx), cin, cout, emb, attention=res in attn_resolutions)
            skips.append(cout)
    L = len(channel_mult)
    for level, mult in reversed(list(enumerate(channel_mult))):
        res = img_resolution >> level
        if level == L - 1:
            _edm_block(I, "dec.%dx%d_in0." % (res, res), cout, cout, emb, attention=True)
            _edm_block(I, "dec.%dx%d_in1." % (res, res), cout, cout, emb)
        else:
            _edm_block(I, "dec.%dx%d_up." % (res, res), cout, cout, emb, up=True)
        for idx in range(num_blocks + 1):
            cin, cout = cout + skips.pop(), model_channels * mult
            _edm_block(I, "dec.%dx%d_block%d." % (res, res, idx), cin, cout, emb,
                       attention=(idx == num_blocks and res in attn_resolutions))
        if level == 0:
            I.norm("dec.%dx%d_aux_norm" % (res, res), cout)
            I.conv("dec.%dx%d_aux_conv" % (res, res), cout, out_channels, 3, gain=0.5)
    return I.sd


def _dhariwal_block(I, p, cin, cout, emb_ch, attention=False, up=False, down=False):
    I.norm(p + "norm0", cin)
    I.conv(p + "conv0", cin, cout, 3)
    if up or down:
        I.sd[p + "conv0.resample_filter"] = torch.full((1, 1, 2, 2), 0.25)
    I.linear(p + "affine", emb_ch, 2 * cout, gain=0.5)  # adaptive_scale: [scale | shift]
    I.norm(p + "norm1", cout)
    I.conv(p + "conv1", cout, cout, 3, gain=0.5)  # init_zero in the reference: re-drawn
    if cin != cout:
        I.conv(p + "skip", cin, cout, 1)
    if up or down:  # (a same-width resampling block has a weight-less skip that only carries the filter buffer)
        I.sd[p + "skip.resample_filter"] = torch.full((1, 1, 2, 2), 0.25)
    if attention:
        I.norm(p + "norm2", cout)
        I.conv(p + "qkv", cout, 3 * cout, 1)
        I.conv(p + "proj", cout, cout, 1, gain=0.5)


def dhariwal_unet_state_dict(img_resolution, in_channels, out_channels, model_channels, channel_mult, num_blocks,
                             attn_resolutions, channel_mult_emb=4, seed=0, **_):
    """Same keys/shapes as src.edm_networks.DhariwalUNet(...).state_dict() (unconditional: label_dim = augment_dim = 0)."""
    I = _Init(seed)
    emb = model_channels * channel_mult_emb
    I.linear("map_layer0", model_channels, emb)
    I.linear("map_layer1", emb, emb)
    cout = in_channels
    skips = []
    for level, mult in enumerate(channel_mult):
        res = img_resolution >> level
        if level == 0:
            cin, cout = cout, model_channels * mult
            I.conv("enc.%dx%d_conv" % (res, res), cin, cout, 3)
        else:
            _dhariwal_block(I, "enc.%dx%d_down." % (res, res), cout, cout, emb, down=True)
        skips.append(cout)
        for idx in range(num_blocks):
            cin, cout = cout, model_channels * mult
            _dhariwal_block(I, "enc.%dx%d_block%d." % (res, res, idx), cin, cout, emb, attention=res in attn_resolutions)
            skips.append(cout)
    L = len(channel_mult)
    for level, mult in reversed(list(enumerate(channel_mult))):
        res = img_resolution >> level
        if level == L - 1:
            _dhariwal_block(I, "dec.%dx%d_in0." % (res, res), cout, cout, emb, attention=True)
            _dhariwal_block(I, "dec.%dx%d_in1." % (res, res), cout, cout, emb)
        else:
            _dhariwal_block(I, "dec.%dx%d_up." % (res, res), cout, cout, emb, up=True)
        for idx in range(num_blocks + 1):
            cin, cout = cout + skips.pop(), model_channels * mult
            _dhariwal_block(I, "dec.%dx%d_block%d." % (res, res, idx), cin, cout, emb, attention=res in attn_resolutions)
    I.norm("out_norm", cout)
    I.conv("out_conv", cout, out_channels, 3, gain=0.5)
    return I.sd


def edm_sigma_state_dict(dim, channels, n_blocks, seed=1, fc_dim=128):
    """Same keys/shapes as src.edm_networks.SigmaModel(dim, channels, n_blocks).state_dict()."""
    I = _Init(seed)
    idx, d = 0, dim
    for i in range(n_blocks):
        if d % 2 != 0:
            d += 1
        idx += 1
        _edm_block(I, "down_layer.%d." % idx, channels, channels, None, attention=i % 2 == 0)
        idx += 1
        I.conv("down_layer.%d.conv" % idx, channels, channels, 3)
        idx += 1
        d //= 2
    I.linear("fc_layer.1", channels * d * d, fc_dim)
    I.sd["fc_layer.2.weight"] = 1.0 + 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.bias"] = 0.05 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_mean"] = 0.1 * I.randn(fc_dim)
    I.sd["fc_layer.2.running_var"] = 0.5 + torch.rand(fc_dim, generator=I.g)
    I.sd["fc_layer.2.num_batches_tracked"] = torch.tensor(0)
    I.linear("final_mlp", fc_dim, 1, gain=0.3)
    return I.sd


DHARIWAL_CONFIGS = {
    # the ADM architecture of the EDM code base at 64x64 with 128-wide levels (the ImageNet-64 checkpoint's 192-wide levels
    # give 6 / 18 channels per GroupNorm group, which the GroupNorm kernels do not take: multiples of 4 only)
    "dhariwal64": dict(img_resolution=64, in_channels=3, out_channels=3, model_channels=128, channel_mult=(1, 2, 3, 4),
                       num_blocks=3, attn_resolutions=(32, 16, 8)),
    # two levels, one block per level, for unit tests
    "dhariwal_tiny": dict(img_resolution=16, in_channels=3, out_channels=3, model_channels=128, channel_mult=(1, 2),
                          num_blocks=1, attn_resolutions=(8,)),
}
EDM_CONFIGS = {
    # c3: EDM ffhq-64 DDPM++
    "edm64": dict(img_resolution=64, in_channels=3, out_channels=3, model_channels=128, channel_mult=(1, 2, 2, 2),
                  num_blocks=4, attn_resolutions=(16,), sigma=dict(dim=8, channels=256, n_blocks=2)),
    # same topology, two levels, one block per level, for unit tests
    "edm_tiny": dict(img_resolution=16, in_channels=3, out_channels=3, model_channels=128, channel_mult=(1, 2),
                     num_blocks=1, attn_resolutions=(8,), sigma=dict(dim=8, channels=256, n_blocks=2)),
}


# ----------------------------------------------------------------------------------------------------------------------
# FID InceptionV3 (pytorch_fid's pt_inception-2015-12-05 weights are a state_dict in torchvision's Inception3 layout; they
# cannot be fetched here, so the FID tests and the oracle use this seeded stand-in with the same keys and shapes).
def fid_inception_layers():
    """(name, cin, cout, kh, kw) of every BasicConv2d of the FID Inception trunk (torchvision Inception3 up to Mixed_7c)."""
    L = [("Conv2d_1a_3x3", 3, 32, 3, 3), ("Conv2d_2a_3x3", 32, 32, 3, 3), ("Conv2d_2b_3x3", 32, 64, 3, 3),
         ("Conv2d_3b_1x1", 64, 80, 1, 1), ("Conv2d_4a_3x3", 80, 192, 3, 3)]
    for n, cin, pf in (("Mixed_5b", 192, 32), ("Mixed_5c", 256, 64), ("Mixed_5d", 288, 64)):
        L += [(n + ".branch1x1", cin, 64, 1, 1), (n + ".branch5x5_1", cin, 48, 1, 1), (n + ".branch5x5_2", 48, 64, 5, 5),
              (n + ".branch3x3dbl_1", cin, 64, 1, 1), (n + ".branch3x3dbl_2", 64, 96, 3, 3),
              (n + ".branch3x3dbl_3", 96, 96, 3, 3), (n + ".branch_pool", cin, pf, 1, 1)]
    L += [("Mixed_6a.branch3x3", 288, 384, 3, 3), ("Mixed_6a.branch3x3dbl_1", 288, 64, 1, 1),
          ("Mixed_6a.branch3x3dbl_2", 64, 96, 3, 3), ("Mixed_6a.branch3x3dbl_3", 96, 96, 3, 3)]
    for n, c7 in (("Mixed_6b", 128), ("Mixed_6c", 160), ("Mixed_6d", 160), ("Mixed_6e", 192)):
        L += [(n + ".branch1x1", 768, 192, 1, 1), (n + ".branch7x7_1", 768, c7, 1, 1), (n + ".branch7x7_2", c7, c7, 1, 7),
              (n + ".branch7x7_3", c7, 192, 7, 1), (n + ".branch7x7dbl_1", 768, c7, 1, 1),
              (n + ".branch7x7dbl_2", c7, c7, 7, 1), (n + ".branch7x7dbl_3", c7, c7, 1, 7),
              (n + ".branch7x7dbl_4", c7, c7, 7, 1), (n + ".branch7x7dbl_5", c7, 192, 1, 7),
              (n + ".branch_pool", 768, 192, 1, 1)]
    L += [("Mixed_7a.branch3x3_1", 768, 192, 1, 1), ("Mixed_7a.branch3x3_2", 192, 320, 3, 3),
          ("Mixed_7a.branch7x7x3_1", 768, 192, 1, 1), ("Mixed_7a.branch7x7x3_2", 192, 192, 1, 7),
          ("Mixed_7a.branch7x7x3_3", 192, 192, 7, 1), ("Mixed_7a.branch7x7x3_4", 192, 192, 3, 3)]
    for n, cin in (("Mixed_7b", 1280), ("Mixed_7c", 2048)):
        L += [(n + ".branch1x1", cin, 320, 1, 1), (n + ".branch3x3_1", cin, 384, 1, 1),
              (n + ".branch3x3_2a", 384, 384, 1, 3), (n + ".branch3x3_2b", 384, 384, 3, 1),
              (n + ".branch3x3dbl_1", cin, 448, 1, 1), (n + ".branch3x3dbl_2", 448, 384, 3, 3),
              (n + ".branch3x3dbl_3a", 384, 384, 1, 3), (n + ".branch3x3dbl_3b", 384, 384, 3, 1),
              (n + ".branch_pool", cin, 192, 1, 1)]
    return L


def fid_inception_state_dict(seed=1):
    """Seeded weights in the layout of pytorch_fid's FID Inception (conv.weight + bn.{weight,bias,running_mean,running_var});
    He-scaled so that activations stay O(1) through the 94 layers, non-trivial BatchNorm statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cin, cout, kh, kw in fid_inception_layers():
        sd[name + ".conv.weight"] = torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5
        sd[name + ".bn.weight"] = 1.0 + 0.2 * (torch.rand(cout, generator=g) - 0.5)
        sd[name + ".bn.bias"] = 0.1 * torch.randn(cout, generator=g)
        sd[name + ".bn.running_mean"] = 0.1 * torch.randn(cout, generator=g)
        sd[name + ".bn.running_var"] = 0.5 + torch.rand(cout, generator=g)
    return sd
