"""Checkpoint ingestion: the three file formats the reference's drivers read, turned into the plain
`state_dict` the nlc_b200 network classes take (`load_state_dict` keeps the reference's key names).

* guided-diffusion / ADM `.pt`: a plain state dict (image_sample.py:757-771 via src/dist_util.py:57-77, which only adds an
  MPI broadcast around `torch.load`).
* DDIM `.ckpt`: a list `[model_state, optimizer_state, epoch, step, ema_state]` written from an `nn.DataParallel` model;
  the reference loads element 0 and then overwrites every trainable parameter with the EMA copy in the last element
  (run_image_experiment.py:195-209).
* EDM network pickles: `pickle.load(f)['ema']` is an `EDMPrecond` whose `.model` is the SongUNet
  (edm_image_sample.py:152-156).  Those pickles embed the defining module's SOURCE and re-execute it on load
  (torch_utils/persistence.py); here they are read with a restricted unpickler that never executes it: persistent objects
  come back as inert records of their `__dict__`, from which the parameter / buffer tree is walked; ordinary classes of
  the training code (`src.*`, `training.*`) come back as inert stand-ins in the same way.
"""
import collections
import io
import pickle

import torch


def load_state_dict(path, **kwargs):
    """src/dist_util.py:57-77 without the MPI broadcast (one process per GPU reads its own copy)."""
    kwargs.setdefault("map_location", "cpu")
    with open(path, "rb") as f:
        return torch.load(io.BytesIO(f.read()), **kwargs)


def _strip_module(sd):
    return collections.OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())


def eps_state_dict(ckpt, trainable=None):
    """The state dict the reference ends up with after run_image_experiment.py:195-209: a plain dict is returned as is; for
    a DDIM list checkpoint the DataParallel prefix is dropped and the EMA values replace the parameters (`trainable`:
    optional set of names to restrict the overwrite to, the reference's `requires_grad` filter; buffers are never in the EMA
    dict)."""
    if not isinstance(ckpt, (list, tuple)):
        return ckpt
    sd = _strip_module(ckpt[0])
    ema = _strip_module(ckpt[-1])
    for name, value in ema.items():
        if name in sd and (trainable is None or name in trainable):
            sd[name] = value.detach().clone() if torch.is_tensor(value) else value
    return sd


def load_eps_model(model, ckpt_file, trust_pickle=False):
    """`model.load_state_dict` from a `.pt` / `.ckpt` file, EMA weights applied (run_image_experiment.py:190-212).
    The file is read with torch's weights-only loader (tensors, containers and plain scalars - all these formats hold);
    `trust_pickle=True` is the explicit opt-in to the full unpickler for a checkpoint that carries other objects."""
    try:
        ckpt = load_state_dict(ckpt_file, weights_only=True)
    except pickle.UnpicklingError:
        if not trust_pickle:
            raise
        ckpt = load_state_dict(ckpt_file, weights_only=False)
    return model.load_state_dict(eps_state_dict(ckpt))


# ------------------------------------------------------------------------------------------------ EDM pickles
class _Record:
    """What a persistent object (torch_utils.persistence) is read back as: its class name and its `__dict__`."""

    def __init__(self, meta):
        self.class_name = meta.get("class_name")
        self.state = meta.get("state") or {}

    def __getattr__(self, name):  # obj.model, obj.sigma_data ... as on the live object
        state = self.__dict__.get("state", {})
        if name in state:
            return state[name]
        if name in state.get("_modules", {}):
            return state["_modules"][name]
        raise AttributeError(name)


class _AttrDict(dict):
    __getattr__ = dict.get


class _RestrictedUnpickler(pickle.Unpickler):
    # what a pickled module tree needs to come back as data: EXACT (module, name) pairs only - a whole-module allow-list
    # would hand out helpers such as torch._utils._import_dotted_name, which resolves any callable by name
    _ALLOWED_NAMES = {
        "collections": {"OrderedDict"},
        "builtins": {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "complex", "bytes", "bytearray",
                     "str", "slice", "range", "object"},
        "copyreg": {"_reconstructor"},
        "_codecs": {"encode"},
        "numpy": {"ndarray", "dtype"},
        "numpy.core.multiarray": {"_reconstruct", "scalar"},
        "numpy._core.multiarray": {"_reconstruct", "scalar"},
        "torch._utils": {"_rebuild_tensor", "_rebuild_tensor_v2", "_rebuild_parameter", "_rebuild_parameter_with_state"},
        "torch._tensor": {"_rebuild_from_type_v2"},
        "torch.nn.parameter": {"Parameter"},
        "torch": {"Size", "device", "Tensor", "FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage", "LongStorage",
                  "IntStorage", "BoolStorage", "ByteStorage", "UntypedStorage", "float32", "float16", "bfloat16", "float64",
                  "int64", "int32", "bool", "uint8"},
    }
    # torch.nn container / layer classes of a pickled module tree are never instantiated: inert stand-ins, like the
    # training code's own classes
    # model-code namespaces: classes from these are NOT imported; they come back as inert stand-ins that only hold the
    # pickled __dict__ (a plain nn.Module of the training code, e.g. this repository's own non-persistent
    # src.edm_networks.SongUNet inside a persistent EDMPrecond)
    _MODEL_PREFIXES = ("src", "training", "torch_utils", "dnnlib")

    def find_class(self, module, name):
        if module == "torch_utils.persistence" and name == "_reconstruct_persistent_obj":
            return _Record
        if module.startswith("dnnlib") and name == "EasyDict":
            return _AttrDict
        if module == "torch.storage" and name == "_load_from_bytes":
            return _load_storage_from_bytes  # torch's own helper would torch.load() the blob with the full unpickler
        if "." in name:  # protocol-4 dotted attribute paths ("sys.modules", "os.system" ...) are never needed
            raise pickle.UnpicklingError("refusing the dotted name %s.%s from a checkpoint" % (module, name))
        if name in self._ALLOWED_NAMES.get(module, ()):
            return super().find_class(module, name)
        if module.split(".")[0] in self._MODEL_PREFIXES or module.startswith("torch.nn.modules."):
            return type(str(name), (_Inert,), {"__module__": "nlc_b200.checkpoints.stub." + module})
        raise pickle.UnpicklingError("refusing to import %s.%s from a checkpoint" % (module, name))


def _load_storage_from_bytes(b):
    """torch.storage._load_from_bytes with the weights-only unpickler (storages of plain tensors need nothing else)."""
    return torch.load(io.BytesIO(b), weights_only=True)


class _Inert:
    """Stand-in for a class of the training code: holds whatever state the pickle sets, runs nothing."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


def _module_state(obj):
    return obj.state if isinstance(obj, _Record) else obj.__dict__


def module_state_dict(obj, prefix=""):
    """`nn.Module.state_dict()` of a module tree read back as records: parameters, persistent buffers, sub-modules."""
    st = _module_state(obj)
    sd = collections.OrderedDict()
    for name, p in (st.get("_parameters") or {}).items():
        if p is not None:
            sd[prefix + name] = p.detach()
    skip = st.get("_non_persistent_buffers_set") or set()
    for name, b in (st.get("_buffers") or {}).items():
        if b is not None and name not in skip:
            sd[prefix + name] = b
    for name, sub in (st.get("_modules") or {}).items():
        if sub is not None:
            sd.update(module_state_dict(sub, prefix + name + "."))
    return sd


def edm_state_dict(path_or_file, key="ema"):
    """State dict of the SongUNet inside an EDM network pickle (edm_image_sample.py:152-156:
    `pickle.load(f)['ema'].model.state_dict()`), without executing the source code the pickle carries."""
    if hasattr(path_or_file, "read"):
        data = _RestrictedUnpickler(path_or_file).load()
    else:
        with open(path_or_file, "rb") as f:
            data = _RestrictedUnpickler(f).load()
    net = data[key]
    return module_state_dict(net.model)
