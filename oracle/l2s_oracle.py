"""CPU oracle for the Lip2Speech inference hot path — TEST INFRASTRUCTURE ONLY.

A functional restatement (torch CPU fp32, no nn.Module, no reference import) of what the
reference computes from raw `state_dict` tensors.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product
path (`lip2speech_b200/`) never does and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against the *reference modules themselves* run in the build container
(`tests/golden/make_golden.py` imports /root/reference, loads the same seeded weights with
strict=True, injects the same gumbel noise, and stores the reference outputs under
`tests/golden/`).  `tests/test_oracle_golden.py` checks this file against those vectors.
Third-party arithmetic (torch 1.9 in the reference's requirements vs torch 2.11 here) is
"parity unpinned" in the sense of SURVEY.md §8c: the oracle is torch-2.11 CPU fp32 semantics.

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class _BnMode:
    """Train-mode switch for every BatchNorm of the restatement (oracle/train_oracle.py flips it): batch statistics
    (biased variance) normalise, and the updated running statistics (momentum 0.1, unbiased variance, as
    nn.BatchNorm*d does in train()) are collected in `updated` keyed by state_dict name."""
    train = False
    updated = None


def _bn(sd, name, x):
    """BatchNorm{1,2,3}d: (x-mean)/sqrt(var+eps)*gamma+beta with running (eval) or batch (train) statistics."""
    if _BnMode.train:
        rm, rv = sd[name + ".running_mean"].detach().clone(), sd[name + ".running_var"].detach().clone()
        y = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], True, 0.1, BN_EPS)
        if _BnMode.updated is not None:
            _BnMode.updated[name + ".running_mean"], _BnMode.updated[name + ".running_var"] = rm, rv
        return y
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, BN_EPS)


def psine(x, w, channel_dim=-1):
    """PSine: sin(x) * w with w per channel (decoder.py:43-70)."""
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    return torch.sin(x) * w.view(shape)


def _lstm(x, h0, c0, sd, name, layers, bidirectional=False):
    """nn.LSTM(batch_first=True) through the same ATen op the reference dispatches to
    (gate order i,f,g,o)."""
    flat = []
    for l in range(layers):
        for sfx in ([""] + (["_reverse"] if bidirectional else [])):
            flat += [sd[f"{name}.weight_ih_l{l}{sfx}"], sd[f"{name}.weight_hh_l{l}{sfx}"],
                     sd[f"{name}.bias_ih_l{l}{sfx}"], sd[f"{name}.bias_hh_l{l}{sfx}"]]
    # train flag: the dropout probability is 0 here, so it changes no value; cuDNN refuses a backward pass without it (only the
    # eager-CUDA baseline of bench.py --config c3 differentiates this op on the GPU)
    train = x.is_cuda and torch.is_grad_enabled() and any(t.requires_grad for t in flat)
    out, h, c = torch._VF.lstm(x, (h0, c0), flat, True, layers, 0.0, train, bidirectional, True)
    return out, h, c


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """Explicit cell (used by tests to cross-check _lstm): i,f,g,o split."""
    g = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, gg, o = g.chunk(4, dim=-1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


# ----------------------------------------------------------------------------------------
# video frontend  (model/modules/video.py:68-87, shufflenetv2.py:26-104,151-152)
# ----------------------------------------------------------------------------------------

def video_stem(sd, x, p="encoder."):
    """Conv3d(3->24,(5,7,7),s(1,2,2),p(2,3,3)) -> BN3d -> PReLU -> MaxPool3d((1,3,3),(1,2,2),(0,1,1))
    (video.py:68-72).  x [B,3,T,H,W] -> [B,24,T,H/4,W/4]."""
    y = F.conv3d(x, sd[p + "frontend3D.0.weight"], None, (1, 2, 2), (2, 3, 3))
    y = _bn(sd, p + "frontend3D.1", y)
    y = F.prelu(y, sd[p + "frontend3D.2.weight"])
    return F.max_pool3d(y, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def _shuffle2(x):
    """channel_shuffle(groups=2): out[2j+g] = in[g*C/2 + j] (shufflenetv2.py:26-40)."""
    n, c, h, w = x.shape
    return x.view(n, 2, c // 2, h, w).transpose(1, 2).reshape(n, c, h, w)


def _branch2(sd, p, x, stride):
    y = F.relu(_bn(sd, p + "banch2.1", F.conv2d(x, sd[p + "banch2.0.weight"])))
    wdw = sd[p + "banch2.3.weight"]
    y = _bn(sd, p + "banch2.4", F.conv2d(y, wdw, None, stride, 1, 1, wdw.shape[0]))
    return F.relu(_bn(sd, p + "banch2.6", F.conv2d(y, sd[p + "banch2.5.weight"])))


def _branch1(sd, p, x):
    wdw = sd[p + "banch1.0.weight"]
    y = _bn(sd, p + "banch1.1", F.conv2d(x, wdw, None, 2, 1, 1, wdw.shape[0]))
    return F.relu(_bn(sd, p + "banch1.3", F.conv2d(y, sd[p + "banch1.2.weight"])))


def video_trunk(sd, x, p="encoder."):
    """16 InvertedResidual blocks + conv_last + AvgPool2d(3) (shufflenetv2.py:42-104,151-152;
    wired at video.py:62-65).  x [N,24,h,w] -> [N,768]."""
    blk = 0
    while (p + f"trunk.0.{blk}.banch2.0.weight") in sd:
        q = p + f"trunk.0.{blk}."
        if (q + "banch1.0.weight") in sd:           # stride-2, two branches (shufflenetv2.py:101-102)
            x = torch.cat((_branch1(sd, q, x), _branch2(sd, q, x, 2)), 1)
        else:                                        # stride-1, split + pass-through (97-100)
            half = x.shape[1] // 2
            x = torch.cat((x[:, :half], _branch2(sd, q, x[:, half:], 1)), 1)
        x = _shuffle2(x)
        blk += 1
    x = F.relu(_bn(sd, p + "trunk.1.1", F.conv2d(x, sd[p + "trunk.1.0.weight"])))
    x = F.avg_pool2d(x, 3)
    return x.reshape(x.shape[0], -1)


def video_features(sd, x, p="encoder."):
    """VideoExtractor.forward (video.py:76-87): [B,3,T,H,W] -> L2-normalised [B,T,768]."""
    b = x.shape[0]
    y = video_stem(sd, x, p)
    t = y.shape[2]
    y = y.transpose(1, 2).reshape(b * t, y.shape[1], y.shape[3], y.shape[4])   # video.py:20-23
    y = video_trunk(sd, y, p).view(b, t, -1)
    return F.normalize(y, p=2, dim=2)


# ----------------------------------------------------------------------------------------
# speaker encoder  (model/modules/audio.py:110-150)
# ----------------------------------------------------------------------------------------

def speaker_melspec(sd, wav, p=""):
    """torchaudio MelSpectrogram(16k, n_fft=400, hop=160, n_mels=40, power=2, center/reflect,
    HTK, norm=None), no log (audio.py:124,133): wav [B,S] -> [B,40,1+S//160]."""
    win = sd[p + "mel_spec.spectrogram.window"]
    n_fft = win.numel()
    spec = torch.stft(wav, n_fft, hop_length=160, win_length=n_fft, window=win, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = spec.real ** 2 + spec.imag ** 2                 # [B,201,F]
    return torch.matmul(power.transpose(1, 2), sd[p + "mel_spec.mel_scale.fb"]).transpose(1, 2)


def speaker_forward(sd, wav, p=""):
    """SpeakerEncoder.forward: raw (pre-ReLU) embedding [B,256] (audio.py:132-142)."""
    mel = speaker_melspec(sd, wav, p).permute(0, 2, 1)
    b = wav.shape[0]
    z = torch.zeros(3, b, 256, device=wav.device)
    _, h, _ = _lstm(mel, z, z.clone(), sd, p + "lstm", 3)
    return F.linear(h[-1], sd[p + "linear.weight"], sd[p + "linear.bias"])


def speaker_inference(sd, wav, p=""):
    """SpeakerEncoder.inference: normalize(relu(forward)) (audio.py:144-150)."""
    return F.normalize(F.relu(speaker_forward(sd, wav, p)), p=2, dim=1)


# ----------------------------------------------------------------------------------------
# decoder  (model/modules/decoder.py)
# ----------------------------------------------------------------------------------------

def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def multihop(sd, p, x):
    """MultiHopConv (decoder.py:159-196): x [B,512,T] -> [B,512,T]."""
    feats = [x]
    for j in range(4):
        w = sd[f"{p}conv.{j}.0.weight"]
        y = F.conv1d(x, w, sd[f"{p}conv.{j}.0.bias"], padding=w.shape[2] // 2)
        feats.append(F.silu(_bn(sd, f"{p}conv.{j}.1", y)))
    return F.conv1d(torch.cat(feats, 1), sd[p + "bottleneck.weight"], sd[p + "bottleneck.bias"])


def content_encode(sd, p, x, gumbel_noise, tau=0.1):
    """Content.encode (decoder.py:239-260).  x [B,512,T]; gumbel_noise [B*minT,501] is the
    explicit g = -log(Exp(1)) draw.  Returns key [B,256,minT], value [B,minT,256],
    content_dis [B*minT,501]."""
    feats = [x]
    min_t = x.shape[-1]
    for j in range(4):
        w = sd[f"{p}agg.{j}.0.weight"]
        y = F.conv1d(x, w, sd[f"{p}agg.{j}.0.bias"], stride=w.shape[2])
        y = F.silu(_bn(sd, f"{p}agg.{j}.1", y))
        min_t = min(min_t, y.shape[-1])
        feats.append(y)
    z = torch.cat([F.adaptive_avg_pool1d(f, min_t) for f in feats], 1)
    w = F.conv1d(z, sd[p + "bottleneck.weight"], sd[p + "bottleneck.bias"]).permute(0, 2, 1)
    key = F.silu(_lin(sd, p + "K.2", F.silu(_lin(sd, p + "K.0", w)))).permute(0, 2, 1)
    w = F.silu(_lin(sd, p + "location_fc.0", w))
    w = F.silu(_lin(sd, p + "location_fc.2", w))
    w = F.silu(_lin(sd, p + "location_fc.4", w))
    n, t, _ = w.shape
    w_y = w.reshape(-1, w.shape[-1])
    zs = ((w_y + gumbel_noise) / tau).softmax(-1)                 # F.gumbel_softmax, soft (257)
    value = (zs @ sd[p + "word_embeddings"]).view(n, t, -1)
    return key, value, F.softmax(w_y, dim=-1)


def postnet(sd, x, p="decoder.postnet."):
    """Postnet.forward in eval mode (decoder.py:143-156): x [B,80,L] -> [B,80,L] (no residual)."""
    n = 5
    for i in range(n - 1):
        res = x
        y = F.conv1d(x, sd[f"{p}convolutions.{i}.0.conv.weight"], sd[f"{p}convolutions.{i}.0.conv.bias"], padding=2)
        y = _bn(sd, f"{p}convolutions.{i}.1", y)
        y = psine(y, sd[f"{p}sin_activation.{i}.w"], channel_dim=1)
        x = y + res if i != 0 else y
    y = F.conv1d(x, sd[f"{p}convolutions.{n-1}.0.conv.weight"], sd[f"{p}convolutions.{n-1}.0.conv.bias"], padding=2)
    return _bn(sd, f"{p}convolutions.{n-1}.1", y)


def decoder_preloop(sd, enc_in, face, gumbel_noise, p="decoder."):
    """Decoder.inference lines 383-406: everything before the autoregressive loop.
    enc_in [B,T,1024], face [B,256]."""
    b, t, _ = enc_in.shape
    residual = F.conv1d(enc_in.permute(0, 2, 1), sd[p + "residual_bottleneck.weight"],
                        sd[p + "residual_bottleneck.bias"]).permute(0, 2, 1)
    enc_site = psine(_lin(sd, p + "encoder_site.0.linear_layer", face), sd[p + "encoder_site.1.w"])
    att_site = psine(_lin(sd, p + "attention_site.0.linear_layer", face), sd[p + "attention_site.1.w"])
    h0 = enc_site.unsqueeze(0).repeat(2, 1, 1)
    rnn_out, hidden, cell = _lstm(enc_in, h0, h0.clone(), sd, p + "encoder_rnn", 1, bidirectional=True)
    enc_cell = _lin(sd, p + "E_C.linear_layer", torch.cat([cell[0], cell[1]], -1))
    enc = _lin(sd, p + "encoder_proj.linear_layer", rnn_out) + att_site.unsqueeze(1) + residual
    pos = sd[p + "positional_encodings.pos_table"][:, :t].permute(0, 2, 1)        # [1,512,T]
    enc_ct = enc.permute(0, 2, 1)
    k = psine(multihop(sd, p + "K.0.", enc_ct), sd[p + "K.1.w"], channel_dim=1) + pos
    v = (psine(multihop(sd, p + "V.0.", enc_ct), sd[p + "V.1.w"], channel_dim=1) + pos).permute(0, 2, 1)
    ckey, cval, cdis = content_encode(sd, p + "content.", enc_ct, gumbel_noise)
    return dict(enc=enc, hidden=hidden, enc_cell=enc_cell, k=k, v=v, ckey=ckey, cval=cval, cdis=cdis)


def decoder_steps(sd, pre, steps, p="decoder.", return_attention=False):
    """The autoregressive loop, decoder.py:403-435.  Returns outputs [B,steps,80], lengths [B]."""
    hidden = pre["hidden"]
    b = hidden.shape[1]
    h = [hidden[0], hidden[1]]
    c = [torch.zeros_like(h[0]), torch.zeros_like(h[0])]                 # cell.fill_(0), line 406
    k, v, ckey, cval, enc_cell = pre["k"], pre["v"], pre["ckey"], pre["cval"], pre["enc_cell"]
    pos = sd[p + "positional_encodings.pos_table"][0]
    temp, ctemp = sd[p + "temperature"], sd[p + "content.temperature"]
    ys = sd[p + "BOS"].reshape(1, -1).repeat(b, 1)
    outputs = torch.zeros(b, steps, ys.shape[1], device=ys.device)
    lengths = torch.full((b,), steps, dtype=torch.int64, device=ys.device)
    attn = []
    wl = [(sd[f"{p}decoder_rnn.weight_ih_l{l}"], sd[f"{p}decoder_rnn.weight_hh_l{l}"],
           sd[f"{p}decoder_rnn.bias_ih_l{l}"], sd[f"{p}decoder_rnn.bias_hh_l{l}"]) for l in range(2)]
    for i in range(steps):
        y = psine(_lin(sd, p + "prenet.0.linear_layer", ys), sd[p + "prenet.1.w"])
        y = psine(_lin(sd, p + "prenet.3.linear_layer", y), sd[p + "prenet.4.w"])
        q = psine(_lin(sd, p + "Q.0.linear_layer", torch.cat(h, 1)), sd[p + "Q.1.w"]) + pos[i]
        a = torch.softmax(torch.bmm((q * temp).unsqueeze(1), k), dim=-1)          # [B,1,T]
        if return_attention:
            attn.append(a)
        o = _lin(sd, p + "attention_proj.linear_layer", torch.bmm(a, v).squeeze(1))
        y = y + o
        cq = F.silu(_lin(sd, p + "content.Q.0", torch.cat(c, 1))).unsqueeze(1)    # decoder.py:262-271
        ca = torch.softmax(torch.bmm(cq * ctemp, ckey), dim=-1)
        co = torch.bmm(ca, cval).squeeze(1)
        x = torch.cat([co, y], -1)
        h[0], c[0] = lstm_cell(x, h[0], c[0], *wl[0])
        h[1], c[1] = lstm_cell(h[0], h[1], c[1], *wl[1])
        ys = _lin(sd, p + "fc_out.linear_layer", h[1])
        outputs[:, i] = ys
        stop = _lin(sd, p + "stop_token_layer.linear_layer", torch.cat([h[1], enc_cell], 1))
        hit = (torch.sigmoid(stop[:, 0]) > 0.5) & (lengths == steps)
        lengths[hit] = i + 1
    return outputs, lengths, (torch.cat(attn, 1) if return_attention else None)


def teacher_forcing_mask(tf_ratio: float, steps: int, generator=None):
    """The per-step teacher-forcing decisions of Decoder.forward (decoder.py:355-357): one `torch.rand(1)` draw from the
    CPU generator per step; teacher input is used when rand > tf_ratio AND fewer than int(tf_ratio*steps) were consumed."""
    mask, consumed = [], 0
    for _ in range(steps):
        use = bool(torch.rand(1, generator=generator) > tf_ratio) and consumed < int(tf_ratio * steps)
        consumed += int(use)
        mask.append(use)
    return torch.tensor(mask, dtype=torch.bool)


def decoder_forward(sd, enc_in, face_tiled, mels, tf_mask, gumbel_noise, p="decoder."):
    """Decoder.forward in EVAL mode (decoder.py:320-379; dropouts inactive), teacher-forcing decisions given explicitly.
    mels [B,80,M].  Returns [outputs [B,80,M], post [B,80,M], stop_logits [B,M,1], face [B,256],
    attention logits (PRE-softmax) [B,M,T], content_dis [B*minT,501]]."""
    pre = decoder_preloop(sd, enc_in, face_tiled[:, 0], gumbel_noise, p)
    b, m = mels.shape[0], mels.shape[2]
    hidden = pre["hidden"]
    h = [hidden[0], hidden[1]]
    c = [torch.zeros_like(h[0]), torch.zeros_like(h[0])]
    k, v, ckey, cval, enc_cell = pre["k"], pre["v"], pre["ckey"], pre["cval"], pre["enc_cell"]
    pos = sd[p + "positional_encodings.pos_table"][0]
    temp, ctemp = sd[p + "temperature"], sd[p + "content.temperature"]
    ys = sd[p + "BOS"].reshape(1, -1).repeat(b, 1)
    teacher = torch.cat([ys.unsqueeze(1), mels.permute(0, 2, 1)], dim=1)       # [B, M+1, 80]
    outputs = torch.zeros(b, m, ys.shape[1], device=ys.device)
    stops = torch.zeros(b, m, 1, device=ys.device)
    attn = []
    wl = [(sd[f"{p}decoder_rnn.weight_ih_l{l}"], sd[f"{p}decoder_rnn.weight_hh_l{l}"],
           sd[f"{p}decoder_rnn.bias_ih_l{l}"], sd[f"{p}decoder_rnn.bias_hh_l{l}"]) for l in range(2)]
    for i in range(m):
        if bool(tf_mask[i]):
            ys = teacher[:, i]
        y = psine(_lin(sd, p + "prenet.0.linear_layer", ys), sd[p + "prenet.1.w"])
        y = psine(_lin(sd, p + "prenet.3.linear_layer", y), sd[p + "prenet.4.w"])
        q = psine(_lin(sd, p + "Q.0.linear_layer", torch.cat(h, 1)), sd[p + "Q.1.w"]) + pos[i]
        a = torch.bmm((q * temp).unsqueeze(1), k)                                  # logits [B,1,T] (stored pre-softmax, 364)
        attn.append(a)
        o = _lin(sd, p + "attention_proj.linear_layer", torch.bmm(torch.softmax(a, dim=-1), v).squeeze(1))
        y = y + o
        cq = F.silu(_lin(sd, p + "content.Q.0", torch.cat(c, 1))).unsqueeze(1)
        co = torch.bmm(torch.softmax(torch.bmm(cq * ctemp, ckey), dim=-1), cval).squeeze(1)
        x = torch.cat([co, y], -1)
        h[0], c[0] = lstm_cell(x, h[0], c[0], *wl[0])
        h[1], c[1] = lstm_cell(h[0], h[1], c[1], *wl[1])
        ys = _lin(sd, p + "fc_out.linear_layer", h[1])
        outputs[:, i] = ys
        stops[:, i] = _lin(sd, p + "stop_token_layer.linear_layer", torch.cat([h[1], enc_cell], 1))
    outputs = outputs.permute(0, 2, 1)
    post = postnet(sd, outputs, p + "postnet.") + outputs
    return [outputs, post, stops, face_tiled[:, 0], torch.cat(attn, 1), pre["cdis"]]


def decoder_inference(sd, enc_in, face_tiled, gumbel_noise, steps=300, return_attention=False, p="decoder."):
    """Decoder.inference (decoder.py:382-444).  enc_in [B,T,1024]; face_tiled [B,T,256] (only
    [:,0] is used, line 385).  Returns mel_post [B,80,steps], lengths [B] (, attn [B,steps,T])."""
    pre = decoder_preloop(sd, enc_in, face_tiled[:, 0], gumbel_noise, p)
    outputs, lengths, attn = decoder_steps(sd, pre, steps, p, return_attention)
    outputs = outputs.permute(0, 2, 1)
    mel = postnet(sd, outputs, p + "postnet.") + outputs
    return (mel, lengths, attn) if return_attention else (mel, lengths)


def lip2speech_inference(sd, video, speaker_embedding, gumbel_noise, steps=300, return_attention=False):
    """Lip2Speech.inference with a given speaker embedding (model.py:43-59)."""
    feat = video_features(sd, video, "encoder.")
    t = feat.shape[1]
    face = speaker_embedding.unsqueeze(1).repeat(1, t, 1)
    visual = torch.cat([feat, face], dim=2)
    return decoder_inference(sd, visual, face, gumbel_noise, steps, return_attention, "decoder.")


def demo_span(sd, sd_spk, video, wav, gumbel_noise, steps=300):
    """The hot span of demo.py:84-86: speaker_encoder.inference + net.inference."""
    emb = speaker_inference(sd_spk, wav)
    return lip2speech_inference(sd, video, emb, gumbel_noise, steps)


def gumbel_noise(n_rows: int, vocab: int = 501, generator=None) -> torch.Tensor:
    """g = -log(E), E~Exp(1): what F.gumbel_softmax draws internally."""
    return -torch.empty(n_rows, vocab).exponential_(generator=generator).log()
