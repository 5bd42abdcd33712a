"""Checkpoint ingestion (CPU): the DDIM list-with-EMA format (run_image_experiment.py:195-209) and EDM
network pickles (edm_image_sample.py:152-156) against what the reference's own loading code produces."""
import io
import pickle
import sys
import types

import pytest
import torch

import nlc_b200  # noqa: F401
from nlc_b200 import checkpoints as CK


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 4, 3)
        self.bn = torch.nn.BatchNorm2d(4)
        self.frozen = torch.nn.Linear(2, 2)
        for p in self.frozen.parameters():
            p.requires_grad = False


def test_ddim_list_checkpoint_with_ema(tmp_path):
    torch.manual_seed(0)
    net = torch.nn.DataParallel(_Net())
    ema = {n: torch.randn_like(p) for n, p in net.module.named_parameters() if p.requires_grad}
    path = tmp_path / "model.ckpt"
    torch.save([net.state_dict(), {"opt": 1}, 3, 1234, ema], path)
    # what run_image_experiment.py:195-209 does
    ref = torch.nn.DataParallel(_Net())
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    ref.load_state_dict(ckpt[0], strict=True)
    ref = ref.module
    for name, param in ref.named_parameters():
        if param.requires_grad:
            param.data.copy_(ckpt[-1][name].data)
    trainable = {n for n, p in _Net().named_parameters() if p.requires_grad}
    got = CK.eps_state_dict(CK.load_state_dict(str(path), weights_only=False), trainable=trainable)
    want = ref.state_dict()
    assert list(got) == list(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    plain = {"a": torch.ones(2)}
    assert CK.eps_state_dict(plain) is plain


def _fake_persistence():
    """A stand-in for the vendored torch_utils.persistence: objects pickle as (_reconstruct_persistent_obj, (meta,))."""
    mod = types.ModuleType("torch_utils.persistence")

    def _reconstruct_persistent_obj(meta):
        raise AssertionError("the checkpoint reader must not execute the pickle's own reconstruction code")

    _reconstruct_persistent_obj.__module__ = "torch_utils.persistence"
    _reconstruct_persistent_obj.__qualname__ = "_reconstruct_persistent_obj"
    mod._reconstruct_persistent_obj = _reconstruct_persistent_obj
    pkg = types.ModuleType("torch_utils")
    pkg.persistence = mod
    return pkg, mod


def test_edm_pickle_without_executing_its_source():
    pkg, mod = _fake_persistence()
    sys.modules["torch_utils"], sys.modules["torch_utils.persistence"] = pkg, mod

    class Persistent:
        def __init__(self, class_name, state):
            self.meta = dict(type="class", version=6, module_src="raise SystemExit('executed!')", class_name=class_name,
                             state=state)

        def __reduce__(self):
            return (mod._reconstruct_persistent_obj, (self.meta,))

    def module_state(params=None, buffers=None, modules=None, **extra):
        st = dict(training=False, _parameters=dict(params or {}), _buffers=dict(buffers or {}),
                  _non_persistent_buffers_set=set(), _modules=dict(modules or {}))
        st.update(extra)
        return st

    torch.manual_seed(1)
    w, b, rf = torch.nn.Parameter(torch.randn(4, 3, 3, 3)), torch.nn.Parameter(torch.randn(4)), torch.ones(2, 2)
    conv = Persistent("Conv2d", module_state(params=dict(weight=w, bias=b), buffers=dict(resample_filter=rf)))
    lin = Persistent("Linear", module_state(params=dict(weight=torch.nn.Parameter(torch.randn(5, 4)), bias=None)))
    enc = torch.nn.ModuleDict()  # a real torch container between persistent objects, as in SongUNet.enc
    enc.__dict__["_modules"]["32x32_conv"] = conv
    unet = Persistent("SongUNet", module_state(modules=dict(map_layer0=lin, enc=enc), img_resolution=32))
    precond = Persistent("EDMPrecond", module_state(modules=dict(model=unet), sigma_data=0.5))
    try:
        blob = pickle.dumps(dict(ema=precond, loss_fn=None))
    finally:
        del sys.modules["torch_utils"], sys.modules["torch_utils.persistence"]
    sd = CK.edm_state_dict(io.BytesIO(blob))
    assert list(sd) == ["map_layer0.weight", "enc.32x32_conv.weight", "enc.32x32_conv.bias", "enc.32x32_conv.resample_filter"]
    assert torch.equal(sd["enc.32x32_conv.weight"], w) and torch.equal(sd["enc.32x32_conv.resample_filter"], rf)

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))

    with pytest.raises(pickle.UnpicklingError):
        CK.edm_state_dict(io.BytesIO(pickle.dumps(dict(ema=Evil()))))


def test_edm_pickle_written_by_the_reference_itself():
    """A network pickle produced with the reference's own classes and its vendored torch_utils.persistence (a persistent
    EDMPrecond around the module's plain SongUNet, persistent layers inside): the reader returns exactly
    `pickle.load(f)['ema'].model.state_dict()` (edm_image_sample.py:152-156) without importing src.* or running the
    embedded source."""
    import importlib
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    import refimport
    if not refimport.available():
        pytest.skip("reference tree not present")
    refimport.load()
    EN = importlib.import_module("src.edm_networks")
    torch.manual_seed(0)
    net = EN.EDMPrecond(img_resolution=16, img_channels=3, model_type="SongUNet", model_channels=32, channel_mult=[1, 2],
                        num_blocks=1, attn_resolutions=[8])
    blob = pickle.dumps(dict(ema=net, loss_fn=None))
    want = net.model.state_dict()
    got = CK.edm_state_dict(io.BytesIO(blob))
    assert list(got) == list(want) and all(torch.equal(got[k], want[k]) for k in want)
    for hostile in ((__import__("os").system, ("true",)), (eval, ("1+1",)), (torch.hub.load, ("x", "y"))):
        class Evil:
            def __reduce__(self, _h=hostile):
                return _h
        with pytest.raises(pickle.UnpicklingError):
            CK.edm_state_dict(io.BytesIO(pickle.dumps(dict(ema=Evil()))))


def test_restricted_unpickler_refuses_helper_and_dotted_name_bypasses():
    """The allow-list is exact (module, name) pairs: torch._utils._import_dotted_name (which resolves ANY callable by
    name) and protocol-4 dotted attribute paths must be refused, not resolved."""
    import torch._utils

    class ViaHelper:
        def __reduce__(self):
            return (torch._utils._import_dotted_name, ("os.getcwd",))

    with pytest.raises(pickle.UnpicklingError):
        CK.edm_state_dict(io.BytesIO(pickle.dumps(dict(ema=ViaHelper()))))
    # protocol 4 STACK_GLOBAL with a dotted name: ('torch._utils', 'sys.modules')
    payload = (b"\x80\x04" + b"\x8c\x0ctorch._utils" + b"\x8c\x0bsys.modules" + b"\x93" + b".")
    with pytest.raises(pickle.UnpicklingError):
        CK._RestrictedUnpickler(io.BytesIO(payload)).load()
    for mod, name in (("torch._utils", "_import_dotted_name"), ("torch.storage", "_load_from_bytes_evil"),
                      ("torch._tensor", "Tensor"), ("numpy.core.multiarray", "frombuffer")):
        with pytest.raises(pickle.UnpicklingError):
            CK._RestrictedUnpickler(io.BytesIO(b"")).find_class(mod, name)


def test_load_eps_model_uses_the_weights_only_loader(tmp_path):
    class Evil:
        def __reduce__(self):
            return (eval, ("1+1",))

    class Sink:
        def load_state_dict(self, sd):
            self.sd = sd
            return self

    good = tmp_path / "good.pt"
    torch.save({"w": torch.arange(4.0)}, good)
    assert torch.equal(CK.load_eps_model(Sink(), str(good)).sd["w"], torch.arange(4.0))
    bad = tmp_path / "bad.pt"
    torch.save({"w": Evil()}, bad)
    with pytest.raises(pickle.UnpicklingError):
        CK.load_eps_model(Sink(), str(bad))
