"""`X.from_reference(reference_module)` — the first thing INTEGRATION.md section 1 shows — against the constructor + load_state_dict
path every GPU parity test uses: for each of the six network classes the model built from a LIVE instance of the
reference's own module must come out with the same inferred hyper-parameters and bit-identical packed device weights as
the one built from keyword arguments.  CPU-only (construction and weight packing are host code; planning and running need
the GPU and are covered by the -m gpu parity tests of the kwargs path).  Needs /root/reference; skipped where it is absent."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

import refimport  # noqa: E402
from oracle import weights  # noqa: E402

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")


def _tensors(obj, seen=None, prefix=""):
    """Every tensor reachable from a model object's attributes (weight holders are plain objects / dicts / lists)."""
    seen = set() if seen is None else seen
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if torch.is_tensor(obj):
        yield prefix, obj
    elif isinstance(obj, dict):
        for k, v in obj.items():
            yield from _tensors(v, seen, "%s[%r]" % (prefix, k))
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            yield from _tensors(v, seen, "%s[%d]" % (prefix, i))
    elif hasattr(obj, "__dict__") and type(obj).__module__.startswith(("nlc_b200", "diffusion")):
        for k, v in vars(obj).items():
            if k in ("eng", "_plans"):
                continue
            yield from _tensors(v, seen, prefix + "." + k)


def _plain(v):
    return v is None or isinstance(v, (int, float, bool, str)) or (isinstance(v, (tuple, list)) and all(_plain(e) for e in v))


def _scalars(obj, seen=None, prefix=""):
    """Every plain attribute (ints, strings, tuples of them) of a model object and of its nested weight holders."""
    seen = set() if seen is None else seen
    if id(obj) in seen or torch.is_tensor(obj):
        return
    seen.add(id(obj))
    if isinstance(obj, dict):
        for k, v in obj.items():
            yield from _scalars(v, seen, "%s[%r]" % (prefix, k))
    elif isinstance(obj, (list, tuple)) and not _plain(obj):
        for i, v in enumerate(obj):
            yield from _scalars(v, seen, "%s[%d]" % (prefix, i))
    elif hasattr(obj, "__dict__") and type(obj).__module__.startswith(("nlc_b200", "diffusion")):
        for k, v in vars(obj).items():
            if k in ("eng", "_plans"):
                continue
            if _plain(v):
                yield prefix + "." + k, (list(v) if isinstance(v, tuple) else v)
            else:
                yield from _scalars(v, seen, prefix + "." + k)


def _same_model(a, b, equivalent=()):
    """`equivalent`: top-level constructor arguments that may be spelled differently for the same network (the reference's
    sigma-model keeps only the resulting head count of its attention block: num_heads / num_head_channels)."""
    ta, tb = dict(_tensors(a)), dict(_tensors(b))
    assert ta.keys() == tb.keys() and len(ta) > 4
    for k in ta:
        assert ta[k].shape == tb[k].shape and ta[k].dtype == tb[k].dtype and torch.equal(ta[k], tb[k]), k
    sa, sb = dict(_scalars(a)), dict(_scalars(b))
    for k in equivalent:
        sa.pop("." + k, None), sb.pop("." + k, None)
    assert sa == sb


@pytest.mark.parametrize("prec", ["fp16", "tf32"])
def test_ddim_from_reference(prec):
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    R = refimport.load()
    cfg = weights.CONFIGS["tiny"]
    u, sg = cfg["unet"], cfg["sigma"]
    sd, ssd = weights.ddim_unet_state_dict(**u, seed=3), weights.ddim_sigma_state_dict(**sg, seed=4)
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(sd)
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(ssd)
    _same_model(UNetModel.from_reference(net, precision=prec, device="cpu"),
                UNetModel(**u, precision=prec, device="cpu").load_state_dict(sd))
    _same_model(SigmaModel.from_reference(snet, dim=sg["dim"], precision=prec, device="cpu"),
                SigmaModel(**sg, precision=prec, device="cpu").load_state_dict(ssd))


@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_adm_from_reference(name):
    import make_golden
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    cfg, sg, sd, ssd, net, snet = make_golden.adm_reference_modules(name)
    keys = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
    _same_model(UNetModel.from_reference(net, precision="fp16", device="cpu"),
                UNetModel(in_channels=3, precision="fp16", device="cpu", **{k: cfg[k] for k in keys}).load_state_dict(sd))
    _same_model(SigmaModel.from_reference(snet, dim=sg["dim"], precision="fp16", device="cpu"),
                SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                           num_head_channels=cfg["num_head_channels"],
                           use_new_attention_order=cfg["use_new_attention_order"], precision="fp16",
                           device="cpu").load_state_dict(ssd), equivalent=("num_heads", "num_head_channels"))


def test_edm_from_reference():
    import make_golden
    from nlc_b200.edm_networks import SigmaModel, SongUNet
    cfg, sg, sd, ssd, net, snet = make_golden.edm_reference_modules("edm_tiny")
    _same_model(SongUNet.from_reference(net, precision="fp16", device="cpu"),
                SongUNet(precision="fp16", device="cpu", **cfg).load_state_dict(sd))
    _same_model(SigmaModel.from_reference(snet, dim=sg["dim"], precision="fp16", device="cpu"),
                SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision="fp16",
                           device="cpu").load_state_dict(ssd))
