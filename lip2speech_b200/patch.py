"""Run the reference's OWN classes on the B200 backend: `patch()` rebinds the hot-path methods of the reference's
`VideoExtractor`, `SpeakerEncoder` and `Decoder` (model/modules/{video,audio,decoder}.py) to the C-ABI calls, so
`model.Lip2Speech.inference()/forward()`, `demo.py` and `evaluate.py` run unchanged (INTEGRATION.md
option B).  The reference modules keep owning their parameters (ordinary nn.Parameters under the reference's
state_dict keys — the same keys the mirror modules use), the backend re-packs them whenever a tensor changes.

    from model import model                      # the reference, unmodified
    import lip2speech_b200.patch as b200
    b200.patch()                                 # once, after the reference modules are imported
    net = model.get_network('test').cuda()       # everything below model.py now runs in libl2s_b200.so

`Postnet.forward` is not patched on its own: it is only reached through `Decoder.forward/inference`, which are replaced
as a whole.  In train() mode the same methods run the train-mode CUDA forward as autograd nodes (BatchNorm batch statistics,
every dropout site, BPTT), so `loss.backward()` of train.py:184 fills the reference modules' own p.grad.  With the patched
reference classes the feature dropout of model.py:26 stays the reference's own F.dropout call.  `unpatch()` restores the
PyTorch bodies.
"""
from __future__ import annotations

import sys

from . import _lib, modules, spec

_saved = {}


def _video_forward(self, x):
    if self.training:                          # train.py:167 -> model.py:26: BatchNorm batch statistics + autograd on the CUDA train path
        return modules.video_forward_train(self, "encoder.", x)
    be = _lib.backend(modules._device_index(self))
    be.sync_module(self, "encoder.", _lib.PART_VIDEO)
    return be.video_fwd(x, getattr(self, "precision", _lib.PRECISION_FP32))


def _decoder_sync(self):
    be = _lib.backend(modules._device_index(self))
    be.sync_module(self, "decoder.", _lib.PART_DECODER)
    return be


def _decoder_inference(self, encoder_outputs, face_features, return_attention_map=False, gumbel_noise=None):
    be = _decoder_sync(self)
    B, T = encoder_outputs.shape[:2]
    if gumbel_noise is None:                                  # the draw F.gumbel_softmax makes at decoder.py:257
        gumbel_noise = modules.Decoder.draw_gumbel(self, B, T, encoder_outputs.device)
    return be.decoder_infer(encoder_outputs, face_features[:, 0], gumbel_noise, spec.MAX_DECODER_STEPS, return_attention_map)


def _decoder_forward(self, encoder_outputs, face_features, mels, text_lengths, output_lengths, tf_ratio, gumbel_noise=None):
    if self.training:                          # train.py:167: dropouts, BatchNorm batch statistics, BPTT — one autograd node
        return modules.decoder_forward_train(self, "decoder.", encoder_outputs, face_features, mels, tf_ratio)
    be = _decoder_sync(self)
    B, T = encoder_outputs.shape[:2]
    if gumbel_noise is None:
        gumbel_noise = modules.Decoder.draw_gumbel(self, B, T, encoder_outputs.device)
    mask = modules.Decoder.teacher_forcing_mask(float(tf_ratio), mels.shape[2])
    out_mel, out_post, out_stop, out_attn, out_dis = be.decoder_forward(encoder_outputs, face_features[:, 0], gumbel_noise, mels, mask)
    return [out_mel, out_post, out_stop, face_features[:, 0], out_attn, out_dis]


def _speaker_forward(self, utterances, hidden_init=None):
    return modules.SpeakerEncoder.forward(self, utterances, hidden_init)


def _speaker_inference(self, x):
    return modules.SpeakerEncoder.inference(self, x)


def _find(name, attr):
    mod = sys.modules.get(name)
    return getattr(mod, attr, None) if mod is not None else None


def patch(video_cls=None, speaker_cls=None, decoder_cls=None):
    """Rebind the reference classes (looked up in sys.modules unless passed).  Returns the list of patched methods."""
    video_cls = video_cls or _find("model.modules.video", "VideoExtractor")
    speaker_cls = speaker_cls or _find("model.modules.audio", "SpeakerEncoder")
    decoder_cls = decoder_cls or _find("model.modules.decoder", "Decoder")
    if video_cls is None and speaker_cls is None and decoder_cls is None:
        raise RuntimeError("lip2speech_b200.patch: import the reference first (`from model import model`) or pass its classes")
    done = []
    for cls, name, fn in ((video_cls, "forward", _video_forward), (speaker_cls, "forward", _speaker_forward),
                          (speaker_cls, "inference", _speaker_inference), (decoder_cls, "inference", _decoder_inference),
                          (decoder_cls, "forward", _decoder_forward)):
        if cls is None:
            continue
        key = (cls, name)
        if key not in _saved:
            _saved[key] = cls.__dict__.get(name)
        setattr(cls, name, fn)
        done.append(f"{cls.__module__}.{cls.__name__}.{name}")
    return done


def unpatch():
    for (cls, name), fn in list(_saved.items()):
        if fn is None:
            delattr(cls, name)
        else:
            setattr(cls, name, fn)
        del _saved[(cls, name)]
