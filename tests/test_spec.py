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


def test_train_member_cache_follows_the_module():
    """modules._members walks a module tree once and afterwards re-reads the tensors from their owners: same keys / objects /
    order as named_parameters() and named_buffers(), also after nn.Module._apply (.to(), .double(), ...) replaced the buffer
    objects and after load_state_dict."""
    import torch
    from lip2speech_b200 import modules
    net = modules.get_network("train")
    for mod, pre in ((net.decoder, "decoder."), (net.encoder, "encoder.")):
        for attempt in range(2):
            keys, params, bufs = modules._members(mod, pre)
            ref = list(mod.named_parameters())
            assert keys == [pre + n for n, _ in ref] and all(a is b for a, (_, b) in zip(params, ref))
            refb = list(mod.named_buffers())
            assert [n for n, _ in bufs] == [n for n, _ in refb] and all(a is b for (_, a), (_, b) in zip(bufs, refb))
            mod.double().float()                        # _apply: new buffer objects, parameters updated in place
            mod.load_state_dict(mod.state_dict())
    assert "_l2s_members" in net.decoder.__dict__


def test_train_noise_draw_shapes_and_rates():
    """TrainNoise.draw: one bulk draw per site; shapes as l2s_decoder_train_fwd expects them, keep rates 1 - p of the reference's
    dropout sites (decoder.py:308,363,312,152-154), at most int(tf_ratio * M) teacher-forced steps are REPLACED by predictions."""
    import torch
    from lip2speech_b200 import modules
    torch.manual_seed(0)
    B, T, M = 4, 29, 60
    n = modules.TrainNoise.draw(B, T, M, 0.5, "cpu")
    assert n.tf_mask.shape == (M,) and n.tf_mask.dtype == torch.bool and int(n.tf_mask.sum()) <= int(0.5 * M)
    assert n.gumbel.shape == (B * spec.content_min_t(T), spec.VOCAB)
    assert n.prenet.shape == (M, B, 256) and n.attn.shape == (M, B, T) and n.lstm.shape == (M, B, 512)
    assert [tuple(t.shape) for t in n.post] == [(B, c, M) for c in (512, 512, 512, 512, 80)]
    for t, keep in ((n.prenet, 0.8), (n.attn, 0.9), (n.lstm, 0.9), (n.post[0], 0.5)):
        assert set(t.unique().tolist()) <= {0.0, 1.0} and abs(float(t.mean()) - keep) < 0.02
