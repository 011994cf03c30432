"""Host-side mirror of the reference's nn.Module boundary (reference model/model.py:13-72,
model/modules/video.py:26-87, audio.py:110-150, decoder.py:274-444).

Same class names, method signatures, return values, error behaviour and `state_dict` keys as the
reference — so checkpoints written by the reference's train.py load with `strict=True` — but the
modules own *only parameters*: every `forward` / `inference` body is one call into the C ABI
(include/l2s_b200.h).  There is no PyTorch implementation of the math in this package; if the CUDA
library is missing or the input is not on a B200, the call raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, spec


class ParamTree(nn.Module):
    """Registers parameters/buffers under dotted reference key names (e.g. `K.0.conv.1.0.weight`) by
    building the intermediate containers, so `state_dict()` reproduces the reference's keys."""

    def __init__(self, tensor_spec, seed=None):
        super().__init__()
        for key, (shape, kind) in tensor_spec.items():
            if seed is None:
                t = torch.zeros(shape, dtype=torch.int64 if kind == spec.BN_N else torch.float32)
            else:
                t = spec.seeded_tensor(spec.canonical_key(key, shape), shape, kind, seed)
            self._register(key, t, spec.is_buffer(key))

    def _register(self, key, tensor, is_buffer):
        *path, leaf = key.split(".")
        mod = self
        for name in path:
            if name not in mod._modules:
                mod.add_module(name, nn.Module())
            mod = mod._modules[name]
        if is_buffer:
            mod.register_buffer(leaf, tensor)
        else:
            mod.register_parameter(leaf, nn.Parameter(tensor, requires_grad=tensor.is_floating_point()))


def _device_index(module: nn.Module) -> int:
    p = next(module.parameters())
    if p.device.type != "cuda":
        raise RuntimeError(f"{type(module).__name__}: parameters are on {p.device}; lip2speech_b200 runs on a CUDA (sm_100a) "
                           "device only — call .cuda() (there is no CPU fallback)")
    return p.device.index or 0


class TrainNoise:
    """Every random draw of one train-mode Decoder.forward as explicit tensors: tf_mask [M] (host bool),
    gumbel [B*minT,501], and KEEP masks (1 = kept) prenet [M,B,256], attn [M,B,T], lstm [M,B,512], post 5 x [B,C,M]."""

    def __init__(self, tf_mask, gumbel, prenet, attn, lstm, post, video_drop=None):
        self.tf_mask, self.gumbel, self.prenet, self.attn, self.lstm, self.post, self.video_drop = tf_mask, gumbel, prenet, attn, lstm, post, video_drop

    @staticmethod
    def draw(B, T, M, tf_ratio, device):
        """Draws in the reference's order (gumbel first, then per step: coin flip on the CPU generator, prenet / attention /
        LSTM dropout, then the five postnet dropouts) from torch's global generators."""
        bern = lambda shape, keep: torch.empty(shape, device=device).bernoulli_(keep)
        gumbel = -torch.empty(B * spec.content_min_t(T), spec.VOCAB, device=device).exponential_().log()
        coins = (torch.rand(M) > tf_ratio).tolist()                                   # decoder.py:355-357 (CPU generator, like np/torch.rand(1) there)
        tf, consumed = [], 0
        for c in coins:
            use = c and consumed < int(tf_ratio * M)
            consumed += int(use)
            tf.append(use)
        prenet, attn, lstm = bern((M, B, 256), 0.8), bern((M, B, T), 0.9), bern((M, B, 512), 0.9)
        post = [bern((B, c, M), 0.5) for c in (512, 512, 512, 512, 80)]
        return TrainNoise(torch.tensor(tf, dtype=torch.bool), gumbel, prenet, attn, lstm, post)


def _bind_params(be, tag, keys, params):
    """Bind parameter and gradient memory at addresses that stay the same from step to step (what lets the library replay its
    captured launch graphs); re-binds only when something moved.  Two modes:
    direct  — every trainable parameter already has a dense fp32 .grad (e.g. ClipAdamW's flat buffer, or any optimizer after
              zero_grad(set_to_none=False)): the backward pass accumulates straight into p.grad, as AccumulateGrad would, and the
              autograd node returns no parameter gradients (saves one add kernel per parameter);
    scratch — otherwise: one persistent flat buffer per (backend, module kind); backward hands autograd views of a copy.
    Returns (scratch or None, offsets, sizes)."""
    def dense(p):
        g = p.grad
        return g is not None and g.dtype == torch.float32 and g.device == p.device and g.is_contiguous() and g.shape == p.shape
    cache = be.__dict__.setdefault("_train_grads", {})
    hit = cache.get(tag)
    if all((not p.requires_grad) or dense(p) for p in params):
        sig = ("direct", tuple(keys), tuple(p.data_ptr() for p in params), tuple(p.grad.data_ptr() if p.requires_grad else 0 for p in params))
        if hit is None or hit[0] != sig:
            for k, p in zip(keys, params):
                be.train_bind(k, p.detach(), p.grad if p.requires_grad else None)
            cache[tag] = (sig, None)
        return None, None, None
    sizes = [p.numel() for p in params]
    offs, total = [], 0
    for n in sizes:
        offs.append(total); total += (n + 3) // 4 * 4
    sig = ("scratch", tuple(keys), tuple(p.data_ptr() for p in params))
    if hit is None or hit[0] != sig:
        scratch = torch.zeros(total, device=params[0].device)
        for k, p, o, n in zip(keys, params, offs, sizes):
            be.train_bind(k, p.detach(), scratch[o:o + n])
        hit = cache[tag] = (sig, scratch)
    return hit[1], offs, sizes


def _grad_views(scratch, offs, sizes, params):
    """Fresh storage for what autograd receives (AccumulateGrad may keep the tensor as p.grad; the persistent buffer is rewritten
    by the next backward).  Direct mode: the gradients are already in p.grad."""
    if scratch is None:
        return [None] * len(params)
    out = scratch.clone()
    return [out[o:o + n].view_as(p) for o, n, p in zip(offs, sizes, params)]


class _DecoderTrainFn(torch.autograd.Function):
    """Decoder.forward in train mode as ONE autograd node: forward = l2s_decoder_train_fwd, backward = l2s_decoder_train_bwd.
    The parameters are inputs of the node, so `loss.backward()` (train.py:184) fills p.grad of whatever optimizer owns them."""

    @staticmethod
    def forward(ctx, be, noise, keys, visual, spk, mels, *params):
        scratch, offs, sizes = _bind_params(be, "decoder", keys, params)
        outs = be.decoder_train_fwd(visual, spk, mels, noise, want_input_grads=True)
        ctx.be, ctx.scratch, ctx.layout, ctx.params, ctx.BT = be, scratch, (offs, sizes), params, (visual.shape[0], visual.shape[1])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(outs[3])
        return outs

    @staticmethod
    def backward(ctx, g_mel, g_post, g_stop, g_attn, g_dis):
        if ctx.scratch is not None:
            ctx.scratch.zero_()
        g_visual, g_spk = ctx.be.decoder_train_bwd(g_mel, g_post, g_stop, g_dis, *ctx.BT)
        return (None, None, None, g_visual, g_spk, None, *_grad_views(ctx.scratch, *ctx.layout, ctx.params))


class _VideoTrainFn(torch.autograd.Function):
    """VideoExtractor.forward in train mode (+ the feature dropout of model.py:26) as one autograd node."""

    @staticmethod
    def forward(ctx, be, drop_mask, keys, video, *params):
        scratch, offs, sizes = _bind_params(be, "encoder", keys, params)
        ctx.be, ctx.scratch, ctx.layout, ctx.params = be, scratch, (offs, sizes), params
        return be.video_train_fwd(video, drop_mask)

    @staticmethod
    def backward(ctx, g_feat):
        if ctx.scratch is not None:
            ctx.scratch.zero_()
        ctx.be.video_train_bwd(g_feat)
        return (None, None, None, None, *_grad_views(ctx.scratch, *ctx.layout, ctx.params))


def _members(module, prefix):
    """(keys, parameters, named buffers) of `module`.  named_parameters() / named_buffers() cost ~1 ms per call on these trees,
    which a 25 ms train step would pay twice per forward — so the tree is walked once per module object and only the (owner, name)
    slots are remembered: the tensors are re-read from their owners on every call (.to() / .cuda() replace buffer objects, an
    optimizer may replace parameters).  Registering NEW submodules after the first train-mode forward needs
    `del module._l2s_members`."""
    slots = module.__dict__.get("_l2s_members")
    if slots is None:
        pslots, bslots = [], []
        for mname, m in module.named_modules():
            dot = mname + "." if mname else ""
            pslots += [(prefix + dot + n, m, n) for n, p in m._parameters.items() if p is not None]
            bslots += [(dot + n, m, n) for n, b in m._buffers.items() if b is not None]
        slots = (pslots, bslots)
        object.__setattr__(module, "_l2s_members", slots)
    pslots, bslots = slots
    return [k for k, _, _ in pslots], [m._parameters[n] for _, m, n in pslots], [(k, m._buffers[n]) for k, m, n in bslots]


def _bind_buffers(be, module, prefix):
    named = _members(module, prefix)[2]
    floats = [(prefix + n, b) for n, b in named if b.is_floating_point()]
    cache = be.__dict__.setdefault("_train_buffers", {})
    sig = tuple((k, b.data_ptr()) for k, b in floats)
    if cache.get(prefix) != sig:
        for k, b in floats:
            be.train_bind(k, b, None)                        # pos_table; BatchNorm running statistics (updated in place)
        cache[prefix] = sig
    counters = [b for n, b in named if n.endswith("num_batches_tracked")]
    if counters:
        torch._foreach_add_(counters, 1)                     # nn.BatchNorm*d bookkeeping in train()


def video_forward_train(module, prefix, x, drop_mask=None):
    """Train-mode VideoExtractor.forward for any module owning the reference's encoder parameters under the reference's names."""
    be = _lib.backend(_device_index(module))
    _bind_buffers(be, module, prefix)
    keys, params, _ = _members(module, prefix)
    return _VideoTrainFn.apply(be, drop_mask, keys, x, *params)


def decoder_forward_train(module, prefix, encoder_outputs, face_features, mels, tf_ratio, noise=None):
    """Train-mode Decoder.forward for any module that owns the reference's decoder parameters under the reference's names
    (the mirror `Decoder` below, or the reference's own class after lip2speech_b200.patch.patch())."""
    be = _lib.backend(_device_index(module))
    B, T = encoder_outputs.shape[:2]
    M = mels.shape[2]
    if noise is None:
        noise = TrainNoise.draw(B, T, M, float(tf_ratio), encoder_outputs.device)
    _bind_buffers(be, module, prefix)
    keys, params, _ = _members(module, prefix)
    spk = face_features[:, 0]
    out_mel, out_post, out_stop, out_attn, out_dis = _DecoderTrainFn.apply(be, noise, keys, encoder_outputs, spk, mels, *params)
    return [out_mel, out_post, out_stop, spk, out_attn, out_dis]


class VideoExtractor(ParamTree):
    """Conv3d stem + per-frame ShuffleNetV2 x1.0 trunk + L2 norm (reference video.py:26-87)."""

    def __init__(self, seed=None):
        super().__init__(spec.encoder_spec(), seed)
        self.backend_out = spec.VIDEO_FEAT
        self.precision = _lib.PRECISION_FP32

    def forward(self, x, drop_mask=None):
        if self.training:                      # BatchNorm batch statistics + autograd (drop_mask: the dropout of model.py:26, fused in)
            return video_forward_train(self, "encoder.", x, drop_mask)
        be = _lib.backend(_device_index(self))
        be.sync_module(self, "encoder.", _lib.PART_VIDEO)
        return be.video_fwd(x, self.precision)


class SpeakerEncoder(ParamTree):
    """GE2E-style speaker encoder (reference audio.py:110-150).  `forward` returns the raw (pre-ReLU)
    embedding, `inference` the normalised one, exactly like the reference."""

    def __init__(self, state_dict=None, seed=None):
        super().__init__(spec.speaker_spec(), seed)
        for p in self.parameters():
            p.requires_grad_(False)
        if state_dict is not None:
            self.load_state_dict(state_dict, strict=True)

    def forward(self, utterances, hidden_init=None):
        if hidden_init is not None:
            raise NotImplementedError("hidden_init is never passed by the reference's callers (demo.py:84)")
        be = _lib.backend(_device_index(self))
        be.sync_module(self, "speaker_encoder.", _lib.PART_SPEAKER)
        return be.speaker_fwd(utterances, normalize=False)

    def inference(self, x):
        if self.training:
            self.eval()
        be = _lib.backend(_device_index(self))
        be.sync_module(self, "speaker_encoder.", _lib.PART_SPEAKER)
        return be.speaker_fwd(x, normalize=True)


class Decoder(ParamTree):
    """Bi-LSTM encoder, K/V MultiHopConv, Content slots, 300-step AR loop, Postnet (reference decoder.py:274-444)."""

    def __init__(self, seed=None):
        super().__init__(spec.decoder_spec(), seed)
        self.n_mel_channels = spec.N_MELS
        self.max_decoder_steps = spec.MAX_DECODER_STEPS

    def _sync(self):
        be = _lib.backend(_device_index(self))
        be.sync_module(self, "decoder.", _lib.PART_DECODER)
        return be

    def draw_gumbel(self, n_clips: int, T: int, device) -> torch.Tensor:
        """g = -log(Exp(1)) from torch's global generator on `device` — the draw F.gumbel_softmax makes at
        decoder.py:257 (it is the first RNG consumer of Decoder.inference, so a seeded run consumes the
        same stream as the reference on the same device)."""
        rows = n_clips * spec.content_min_t(T)
        return -torch.empty(rows, spec.VOCAB, device=device, dtype=torch.float32).exponential_().log()

    def inference(self, encoder_outputs, face_features, return_attention_map=False, gumbel_noise=None):
        """encoder_outputs [B,T,1024]; face_features [B,T,256] (only [:,0] is read, decoder.py:385).
        Returns (mel_post [B,80,300], output_lengths [B] int64[, attention [B,300,T]])."""
        be = self._sync()
        B, T = encoder_outputs.shape[:2]
        if gumbel_noise is None:
            gumbel_noise = self.draw_gumbel(B, T, encoder_outputs.device)
        spk = face_features[:, 0]
        return be.decoder_infer(encoder_outputs, spk, gumbel_noise, self.max_decoder_steps, return_attention_map)

    def postnet_forward(self, x):
        """Postnet.forward (eval): x [B,80,L] -> [B,80,L] (reference decoder.py:143-156)."""
        return self._sync().postnet_fwd(x, add_residual=False)

    @staticmethod
    def teacher_forcing_mask(tf_ratio: float, steps: int) -> torch.Tensor:
        """Per-step teacher-forcing decisions exactly as decoder.py:355-357 makes them: one `torch.rand(1)` from the global
        CPU generator per step; teacher input when rand > tf_ratio and fewer than int(tf_ratio*steps) were consumed."""
        mask, consumed = [], 0
        for _ in range(steps):
            use = bool(torch.rand(1) > tf_ratio) and consumed < int(tf_ratio * steps)
            consumed += int(use)
            mask.append(use)
        return torch.tensor(mask, dtype=torch.bool)

    def forward(self, encoder_outputs, face_features, mels, text_lengths, output_lengths, tf_ratio, gumbel_noise=None, train_noise=None):
        """Decoder.forward (decoder.py:320-379).  train(): one autograd node over the train-mode CUDA forward / backward
        (BatchNorm batch statistics, every dropout site, BPTT; `train_noise` = explicit TrainNoise, drawn if omitted).  EVAL mode — the path evaluate.py:38 takes.  Returns
        [outputs [B,80,M], post [B,80,M], stop_logits [B,M,1], face [B,256], attention logits (pre-softmax) [B,M,T],
        content_dis [B*minT,501]]."""
        if self.training:
            return decoder_forward_train(self, "decoder.", encoder_outputs, face_features, mels, tf_ratio, noise=train_noise)
        be = self._sync()
        B, T = encoder_outputs.shape[:2]
        if gumbel_noise is None:
            gumbel_noise = self.draw_gumbel(B, T, encoder_outputs.device)      # content.encode draws first (decoder.py:340)
        mask = self.teacher_forcing_mask(float(tf_ratio), mels.shape[2])
        out_mel, out_post, out_stop, out_attn, out_dis = be.decoder_forward(encoder_outputs, face_features[:, 0], gumbel_noise, mels, mask)
        return [out_mel, out_post, out_stop, face_features[:, 0], out_attn, out_dis]


class Lip2Speech(nn.Module):
    """Top-level glue (reference model.py:13-59).  `vgg_face` (FaceRecognizer, out of scope: third-party
    InceptionResnetV1) may be attached by the caller; `inference` needs it only when no speaker embedding
    is given, as in the reference."""

    def __init__(self, seed=None, vgg_face: nn.Module | None = None):
        super().__init__()
        if vgg_face is not None:
            self.vgg_face = vgg_face
        self.encoder = VideoExtractor(seed)
        self.decoder = Decoder(seed)

    def inference(self, video_frames, face_frames, speaker_embedding=None, gumbel_noise=None, **kwargs):
        with torch.no_grad():
            video_features = self.encoder(video_frames)
            if speaker_embedding is None:
                if not hasattr(self, "vgg_face"):
                    raise RuntimeError("no speaker_embedding given and no vgg_face module attached (model.py:47-50)")
                face_features = self.vgg_face.inference(face_frames[:, 0, :, :, :])
            else:
                face_features = speaker_embedding
            N, T, C = video_features.shape
            face_features = face_features.unsqueeze(1).repeat(1, T, 1)
            visual_features = torch.cat([video_features, face_features], dim=2)
            return self.decoder.inference(visual_features, face_features, gumbel_noise=gumbel_noise, **kwargs)

    def forward(self, video_frames, face_frames, audio_frames, melspecs, video_lengths, audio_lengths, melspec_lengths, tf_ratio,
                speaker_embedding=None, train_noise=None):
        """model.py:23-40.  eval(): what evaluate.py:38 calls.  train(): what train.py:167 calls — train-mode video frontend,
        feature dropout (model.py:26), train-mode decoder, all on the CUDA train path with autograd nodes, so
        `loss.backward()` (train.py:184) produces the gradients of every encoder.* / decoder.* parameter.  The face embedding
        comes from the attached `vgg_face` module (third-party InceptionResnetV1, frozen: not in the optimizer, train.py:102-104)
        exactly as in the reference, or from `speaker_embedding` [B,256] when given."""
        if speaker_embedding is None:
            if not hasattr(self, "vgg_face"):
                raise RuntimeError("Lip2Speech.forward needs the vgg_face module (model.py:31) or a speaker_embedding; attach the reference's FaceRecognizer")
            speaker_embedding = self.vgg_face.inference(face_frames[:, 0, :, :, :])
        if self.training:
            B, T = video_frames.shape[0], video_frames.shape[2]
            keep = train_noise.video_drop if train_noise is not None and train_noise.video_drop is not None else \
                torch.empty(B, T, 768, device=video_frames.device).bernoulli_(0.9)               # F.dropout(..., 0.1, training)
            video_features = self.encoder(video_frames, keep)
            face_features = speaker_embedding.unsqueeze(1).repeat(1, T, 1)
            visual_features = torch.cat([video_features, face_features], dim=2)
            outputs = self.decoder(visual_features, face_features, melspecs, video_lengths, melspec_lengths, tf_ratio, train_noise=train_noise)
            return outputs + [video_lengths]
        with torch.no_grad():
            video_features = self.encoder(video_frames)                    # F.dropout(..., training=False) is the identity
            N, T, C = video_features.shape
            face_features = speaker_embedding.unsqueeze(1).repeat(1, T, 1)
            visual_features = torch.cat([video_features, face_features], dim=2)
            outputs = self.decoder(visual_features, face_features, melspecs, video_lengths, melspec_lengths, tf_ratio)
        return outputs + [video_lengths]


def get_network(mode: str, seed=None) -> Lip2Speech:
    """reference model.py:62-72"""
    assert mode in ("train", "test")
    model = Lip2Speech(seed)
    return model.train() if mode == "train" else model.eval()


def demo_span(backend: "_lib.Backend", video, wav, gumbel, steps=300, precision=_lib.PRECISION_FP32):
    """speaker_encoder.inference + net.inference on device tensors in one C-ABI call (demo.py:84-86)."""
    return backend.infer(video, wav, gumbel, steps, precision)
