"""Host-side mirror of the reference's model factories (src/script_util.py): the same keyword arguments / config objects
in, the nlc_b200 network classes out, so that the YAML `model:` sections and `args.json` files the reference's drivers
read (image_sample.py:112-141, edm_image_sample.py:140-147) construct the CUDA models unchanged.  SURVEY §8f rank 4.

`precision` ("bf16" | "fp16" | "tf32" | "fp32") and `device` are the only additions."""
from . import edm_networks, unet_adm, unet_ddim


def _channel_mult(channel_mult, image_size):
    # src/script_util.py:158-172
    if channel_mult == "":
        table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4),
                 32: (1, 2, 2, 2)}
        if image_size not in table:
            raise ValueError(f"unsupported image size: {image_size}")
        return table[image_size]
    if isinstance(channel_mult, str):
        return tuple(int(m) for m in channel_mult.split(","))
    return tuple(channel_mult)


def create_sigma_eps_model(image_size, num_channels, num_res_blocks, channel_mult="", learn_sigma=False,
                           class_cond=False, use_checkpoint=False, attention_resolutions="16", num_heads=1,
                           num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0.0,
                           resblock_updown=False, use_fp16=False, use_new_attention_order=False, sigma_block=2,
                           sigma_dropout=0.0, use_sigma_fp16=False, precision="bf16", device="cuda", **kwargs):
    """ADM UNet + sigma-model (src/script_util.py:136-206).  Returns (eps_model, sigma_model, feat_shape)."""
    if class_cond:
        raise NotImplementedError("class-conditional ADM networks are outside the sampling path (SURVEY section 8a)")
    mult = _channel_mult(channel_mult, image_size)
    attention_ds = tuple(image_size // int(res) for res in str(attention_resolutions).split(","))
    eps_model = unet_adm.UNetModel(
        image_size=image_size, in_channels=3, model_channels=num_channels, out_channels=(3 if not learn_sigma else 6),
        num_res_blocks=num_res_blocks, attention_resolutions=attention_ds, channel_mult=mult, num_heads=num_heads,
        num_head_channels=num_head_channels, use_scale_shift_norm=use_scale_shift_norm,
        resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order,
        feat_layer=kwargs.get("feat_layer", 1), precision=precision, device=device)
    inp_channels = int(num_channels * mult[-1])
    inp_dim = int(image_size * 0.5 ** (len(mult) - 1))
    sigma_model = unet_adm.SigmaModel(dim=inp_dim, channels=inp_channels, n_blocks=sigma_block, out_dim=1,
                                      dropout=sigma_dropout, num_heads=num_heads, num_head_channels=num_head_channels,
                                      use_new_attention_order=use_new_attention_order, precision=precision,
                                      device=device)
    return eps_model, sigma_model, (inp_channels, inp_dim, inp_dim)


def create_simple_sigma_eps_model(config, precision="bf16", device="cuda"):
    """DDIM UNet (src/unet_simple.py `Model(config)`, the architecture of src/unet_ddim.py) + sigma-model
    (src/script_util.py:209-219).  `config` has the reference's YAML layout: config.model.{ch, out_ch, ch_mult,
    num_res_blocks, attn_resolutions, dropout, in_channels, resamp_with_conv, sigma_block, sigma_dropout},
    config.data.image_size."""
    m = config.model
    mult = tuple(m.ch_mult)
    eps_model = unet_ddim.UNetModel(
        image_size=config.data.image_size, in_channels=m.in_channels, model_channels=m.ch, out_channels=m.out_ch,
        num_res_blocks=m.num_res_blocks, attention_resolutions=tuple(m.attn_resolutions), dropout=m.dropout,
        channel_mult=mult, conv_resample=getattr(m, "resamp_with_conv", True),
        feat_layer=int(getattr(m, "feat_layer", 0) != 0), precision=precision, device=device)
    inp_channels = int(m.ch * mult[-1])
    inp_dim = int(config.data.image_size * 0.5 ** (len(mult) - 1))
    sigma_model = unet_ddim.SigmaModel(dim=inp_dim, channels=inp_channels, n_blocks=m.sigma_block, out_dim=1,
                                       dropout=m.sigma_dropout, precision=precision, device=device)
    return eps_model, sigma_model, (inp_channels, inp_dim, inp_dim)


def create_edm_sigma_eps_model(img_resolution, in_channels, out_channels, augment_dim=0, model_channels=128,
                               channel_mult=(1, 2, 2, 2), channel_mult_emb=4, num_blocks=4, attn_resolutions=(16,),
                               dropout=0.10, embedding_type="positional", encoder_type="standard",
                               decoder_type="standard", resample_filter=(1, 1), sigma_block=2, sigma_dropout=0.0,
                               precision="bf16", device="cuda", **kwargs):
    """EDM SongUNet (DDPM++) + sigma-model (src/script_util.py:222-270)."""
    eps_model = edm_networks.SongUNet(
        img_resolution=img_resolution, in_channels=in_channels, out_channels=out_channels, label_dim=0,
        augment_dim=augment_dim, model_channels=model_channels, channel_mult=list(channel_mult),
        channel_mult_emb=channel_mult_emb, num_blocks=num_blocks, attn_resolutions=list(attn_resolutions),
        dropout=dropout, embedding_type=embedding_type, channel_mult_noise=1, encoder_type=encoder_type,
        decoder_type=decoder_type, resample_filter=list(resample_filter), precision=precision, device=device)
    inp_channels = int(model_channels * channel_mult[-1])
    inp_dim = int(img_resolution * 0.5 ** (len(channel_mult) - 1))
    sigma_model = edm_networks.SigmaModel(dim=inp_dim, channels=inp_channels, n_blocks=sigma_block, out_dim=1,
                                          dropout=sigma_dropout, resample_filter=list(resample_filter),
                                          precision=precision, device=device)
    return eps_model, sigma_model, (inp_channels, inp_dim, inp_dim)
