"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by tests/golden/make_golden.py to generate golden vectors and by
CPU tests that compare the oracle with the live reference when /root/reference exists.  Never
imported by the product path, bench.py or the -m gpu tests (the reference does not exist on
the GPU box).  Missing third-party packages are replaced by inert `sys.modules` shims
(SURVEY.md §8c): facenet_pytorch (vgg_face.py:6), fairseq (audio.py:6), matplotlib.pyplot
(vgg_face.py:2).
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")      # oracle/stage_reference.sh (git-ignored)
REF_ROOT = os.environ.get("L2S_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/model/modules") else _STAGED)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "modules"))


def _shim(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        __import__(name)
        return sys.modules[name]
    except Exception:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m


def import_reference():
    """Returns (Decoder, VideoExtractor, SpeakerEncoder, Lip2Speech) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import torch.nn as nn

    class _NoFace(nn.Module):            # stands in for InceptionResnetV1 (out of scope, needs network)
        def __init__(self, **kw):
            super().__init__()
            self.last_linear = nn.Linear(1, 1)          # vgg_face.py:19-20 touches these two attributes in its constructor
            self.last_bn = nn.BatchNorm1d(1)

    _shim("facenet_pytorch", InceptionResnetV1=_NoFace)
    _shim("fairseq")
    _shim("matplotlib")
    _shim("matplotlib.pyplot", winter=None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from model.modules.decoder import Decoder
    from model.modules.video import VideoExtractor
    from model.modules.audio import SpeakerEncoder
    from model.model import Lip2Speech
    return Decoder, VideoExtractor, SpeakerEncoder, Lip2Speech
