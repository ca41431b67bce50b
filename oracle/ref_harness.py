"""Imports the UNMODIFIED reference (/root/reference) in the build container -- TEST INFRASTRUCTURE ONLY.

The reference needs packages that are not installed here (pytorch_lightning, clip, fire, skimage, pycocoevalcap,
the Salesforce BLIP checkout).  None of them is on the caption-generation path itself, so they are replaced by
empty import stubs; the reference's own files are imported as they lie (PYTHONDONTWRITEBYTECODE: the tree is
read-only).  Used by tools/make_golden.py to produce tests/golden/*.pt and by bench.py's CPU arms (`--impl reference`, `cpu_baseline`).
/root/reference does not exist on the GPU box: there the same modules are imported from oracle/_ref, the sourceless bytecode that
`__graft_entry__.build()` compiles from the reference tree (oracle/build_ref.py); without either, `available()` is False and
bench.py times the oracle's restatement instead.
"""
import os
import sys
import types

COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py: sourceless .pyc


def _pick_root() -> str:
    env = os.environ.get("CLIPCAP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/layers"):
        return "/root/reference"
    return COMPILED_ROOT


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "layers"))


def kind() -> str:
    """"source": the tree itself (build container); "compiled": oracle/_ref, the same modules byte-compiled by build()."""
    return "compiled" if os.path.abspath(REFERENCE_ROOT) == COMPILED_ROOT else "source"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch.nn as nn
    sys.dont_write_bytecode = True

    class _HParams(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class LightningModule(nn.Module):
        def save_hyperparameters(self, ignore=()):
            import inspect
            frame = inspect.currentframe().f_back
            args = dict(frame.f_locals)
            args.pop("self", None)
            args.pop("__class__", None)
            kwargs = args.pop("kwargs", {})
            args.update(kwargs)
            for k in ignore:
                args.pop(k, None)
            object.__setattr__(self, "_hp", _HParams(args))

        @property
        def hparams(self):
            return self._hp

    if "pytorch_lightning" not in sys.modules:
        pl = _stub("pytorch_lightning", LightningModule=LightningModule, Trainer=object, Callback=object)
        _stub("pytorch_lightning.utilities")
        _stub("pytorch_lightning.utilities.deepspeed", convert_zero_checkpoint_to_fp32_state_dict=None)
        pl.utilities = sys.modules["pytorch_lightning.utilities"]
    if "clip" not in sys.modules:
        clip = _stub("clip", load=None, tokenize=None)
        clip.model = _stub("clip.model", VisionTransformer=object, CLIP=object)
    if "fire" not in sys.modules:
        _stub("fire", Fire=lambda f: f)
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.io = _stub("skimage.io")
    if "pycocoevalcap" not in sys.modules:
        _stub("pycocoevalcap")
        _stub("pycocoevalcap.eval", Bleu=None, Meteor=None, Rouge=None, Cider=None, Spice=None, PTBTokenizer=None)
    if "models" not in sys.modules:
        _stub("models")
        _stub("models.blip", blip_decoder=None)
        _stub("models.blip_itm", blip_itm=None)
    if "BLIP" not in sys.modules:
        _stub("BLIP")
        _stub("BLIP.models")
        _stub("BLIP.models.blip", blip_decoder=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    install_stubs()
    import importlib
    ns = types.SimpleNamespace()
    ns.layers = importlib.import_module("layers")
    ns.lms = importlib.import_module("lms")
    ns.model = importlib.import_module("model")
    ns.inference = importlib.import_module("inference")
    ns.evaluate_model = importlib.import_module("evaluate_model")
    ns.sampling = importlib.import_module("sampling")
    return ns
