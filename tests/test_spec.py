"""The programmatically generated state_dict key set == the reference's."""
import json
import os

import pytest

from lip2speech_b200 import spec

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def manifest():
    with open(os.path.join(HERE, "golden", "state_manifest.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name,fn,count", [("decoder", spec.decoder_spec, 191), ("encoder", spec.encoder_spec, 337),
                                           ("speaker_encoder", spec.speaker_spec, 16)])
def test_keys_and_shapes(manifest, name, fn, count):
    mine = spec.manifest(fn())
    assert len(mine) == count
    assert list(mine.keys()) == sorted(mine.keys(), key=list(mine.keys()).index)
    assert mine == manifest[name]


def test_param_counts():
    n = lambda s: sum(int(__import__("numpy").prod(sh)) for k, (sh, kind) in s.items() if not spec.is_buffer(k))
    assert n(spec.decoder_spec()) == 37_285_512      # SURVEY §2.1 [measured]
    assert n(spec.encoder_spec()) == 1_151_324
