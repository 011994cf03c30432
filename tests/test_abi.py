"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/l2s_b200.h declares (no compute without a GPU), and the product path fails loudly off-GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lip2speech_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "l2s_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(l2s_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/l2s_b200.h but not exported"


def test_python_export_list_matches_header():
    from lip2speech_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_version(lib):
    assert lib.l2s_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    h = ctypes.c_void_p()
    assert lib.l2s_create(ctypes.byref(h), 0) != 0
    assert lib.l2s_last_error(None)
    from lip2speech_b200 import modules
    dec = modules.Decoder(seed=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        dec.inference(torch.zeros(1, 29, 1024), torch.zeros(1, 29, 256))


def test_mirror_state_dict_keys_match_reference_manifest():
    import json
    from lip2speech_b200 import modules
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "state_manifest.json")))
    for name, cls in (("decoder", modules.Decoder), ("encoder", modules.VideoExtractor), ("speaker_encoder", modules.SpeakerEncoder)):
        sd = cls().state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == man[name]
    top = modules.Lip2Speech().state_dict()
    assert set(top) == {"encoder." + k for k in man["encoder"]} | {"decoder." + k for k in man["decoder"]}


def test_mirror_loads_reference_style_checkpoint(weights):
    """A demo.py-style checkpoint (encoder.*, decoder.*, speaker_encoder.*) splits and loads strictly."""
    from lip2speech_b200 import modules
    sd = dict(weights)
    spk = {k[len("speaker_encoder."):]: sd.pop(k) for k in list(sd) if k.startswith("speaker_encoder.")}   # demo.py:33-36
    net = modules.get_network("test")
    net.load_state_dict(sd, strict=True)
    modules.SpeakerEncoder(state_dict=spk)
    assert torch.equal(net.decoder.state_dict()["Q.1.w"], weights["decoder.Q.1.w"])
