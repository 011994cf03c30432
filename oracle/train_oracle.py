"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's train-step tail, used as the checker for
csrc/train_step.cuh.  Only tests/ may import this.

* `loss_forward` restates train_utils/losses.py:35-79 (Loss.forward) in plain torch; pinned bit for bit (values and
  gradients) against the unmodified reference module by tests/test_train_oracle_vs_reference.py where /root/reference exists.
* `decoder_forward_train` / `lip2speech_forward_train` restate Decoder.forward (decoder.py:320-379) and Lip2Speech.forward
  (model.py:23-40) in TRAIN mode as pure functions of the state_dict tensors — BatchNorm batch statistics, and every RNG site
  of SURVEY A.4 as an explicit input (`TrainNoise`).  torch.autograd on these functions gives the reference gradients.
  `reference_noise` replays the reference's own draws (same order, same generator calls), so a seeded run of the UNMODIFIED
  reference modules is reproduced bit for bit: tests/test_train_oracle_vs_reference.py pins outputs and gradients that way.
* `clip_adamw_steps` runs the reference's own calls — torch.nn.utils.clip_grad_norm_ (train.py:191) and
  torch.optim.AdamW(lr, weight_decay, amsgrad=True) (train.py:102-104,193) — on CPU copies; torch is the third-party
  arithmetic the reference pins (requirements.txt:10), so parity is against torch 2.11 CPU fp32 semantics
  ("parity unpinned" in SURVEY.md §8c's sense).
"""
import contextlib

import torch
import torch.nn.functional as F

from . import l2s_oracle as O


def loss_forward(model_output, targets):
    """losses.py:35-79.  model_output = [mel, mel_post, gate_logits [B,M,1], face, attn, content_dis [rows,501], ...];
    targets = (mel_target [B,80,M], gate_target [B,M]).  Returns the dict of the four loss terms."""
    mel_target, gate_target = targets[0], targets[1]
    gate_target = gate_target.view(-1, 1)                                   # 43
    mel_out, mel_post, gate_out = model_output[0], model_output[1], model_output[2].view(-1, 1)   # 45-49
    qy = model_output[5]                                                    # 69
    log_ratio = torch.log(qy * qy.shape[-1] + 1e-20)                        # 71
    return {
        "KLD": torch.sum(qy * log_ratio, dim=-1).mean(),                    # 72-73
        "mel_loss": F.mse_loss(mel_out, mel_target),                        # 75
        "postnet_mel_loss": 10 * F.mse_loss(mel_post, mel_target),          # 76
        "gate_loss": F.binary_cross_entropy_with_logits(gate_out, gate_target),   # 77
    }


def clip_adamw_steps(params, grads_per_step, lr=1e-4, weight_decay=1e-6, max_norm=1.0, world_grads=None):
    """params: list of CPU tensors (copied).  grads_per_step: list (steps) of lists (params) of gradient tensors — when
    `world_grads` ranks are given instead ([rank][step][param]) their mean is the gradient (data-parallel semantics).
    Returns (final params, list of pre-clip gradient norms)."""
    ps = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt = torch.optim.AdamW(ps, lr=lr, weight_decay=weight_decay, amsgrad=True)
    norms = []
    steps = len(grads_per_step) if world_grads is None else len(world_grads[0])
    for s in range(steps):
        for i, p in enumerate(ps):
            if world_grads is None:
                p.grad = grads_per_step[s][i].detach().clone()
            else:
                p.grad = torch.stack([wg[s][i] for wg in world_grads]).sum(0) / len(world_grads)
        norms.append(torch.nn.utils.clip_grad_norm_(ps, max_norm))
        opt.step()
    return [p.detach() for p in ps], norms


# ----------------------------------------------------------------------------------------------------------------------
# train-mode forward with explicit noise
# ----------------------------------------------------------------------------------------------------------------------
class TrainNoise:
    """Every random draw of one train-mode forward, as KEEP masks (1 = kept, 0 = dropped; the 1/(1-p) scale is
    applied by the consumer) and the gumbel tensor:
      video_drop [B,T,768] (model.py:26, p=0.1) | gumbel [B*minT,501] (decoder.py:257) | tf_mask [M] bool (355-357)
      prenet [M,B,256] (308, p=0.2) | attn [M,B,T] (363, p=0.1) | lstm [M,B,512] (312: nn.LSTM inter-layer dropout, p=0.1)
      post: 5 x [B,C,M] (152,154, p=0.5; C = 512,512,512,512,80)"""

    def __init__(self, video_drop, gumbel, tf_mask, prenet, attn, lstm, post):
        self.video_drop, self.gumbel, self.tf_mask, self.prenet, self.attn, self.lstm, self.post = video_drop, gumbel, tf_mask, prenet, attn, lstm, post

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return TrainNoise(mv(self.video_drop), mv(self.gumbel), self.tf_mask, mv(self.prenet), mv(self.attn), mv(self.lstm), [mv(t) for t in self.post])


def reference_noise(B, T, M, tf_ratio, with_video=True, generator=None):
    """Replays, in the reference's order and with the reference's generator calls, every draw of a train-mode
    Lip2Speech.forward (CPU): F.dropout / nn.Dropout / nn.LSTM dropout all reach at::dropout -> empty_like(x).bernoulli_(1-p)
    (non-fused CPU path); gumbel_softmax draws empty_like(logits).exponential_(); decoder.py:355 draws torch.rand(1) per step.
    After torch.manual_seed(s) (or with a seeded `generator`) the masks equal the ones the unmodified reference consumes."""
    from lip2speech_b200 import spec
    g = generator
    bern = lambda shape, keep: torch.empty(shape).bernoulli_(keep, generator=g)
    video_drop = bern((B, T, 768), 0.9) if with_video else None
    min_t = spec.content_min_t(T)
    gumbel = -torch.empty(B * min_t, 501).exponential_(generator=g).log()
    tf, prenet, attn, lstm, consumed = [], [], [], [], 0
    for _ in range(M):
        use = bool(torch.rand(1, generator=g) > tf_ratio) and consumed < int(tf_ratio * M)
        consumed += int(use)
        tf.append(use)
        prenet.append(bern((B, 1, 256), 0.8)[:, 0])
        attn.append(bern((B, 1, T), 0.9)[:, 0])
        lstm.append(bern((1, B, 512), 0.9)[0])
    post = [bern((B, c, M), 0.5) for c in (512, 512, 512, 512, 80)]
    return TrainNoise(video_drop, gumbel, torch.tensor(tf, dtype=torch.bool), torch.stack(prenet), torch.stack(attn), torch.stack(lstm), post)


@contextlib.contextmanager
def bn_train(collect=None):
    O._BnMode.train, O._BnMode.updated = True, collect
    try:
        yield
    finally:
        O._BnMode.train, O._BnMode.updated = False, None


def _drop(x, keep, p):
    return x * (keep / (1.0 - p))


def decoder_forward_train(sd, enc_in, face_tiled, mels, noise, p="decoder.", bn_updates=None):
    """Decoder.forward in TRAIN mode (decoder.py:320-379).  enc_in [B,T,1024], face_tiled [B,T,256], mels [B,80,M].
    Returns [outputs [B,80,M], post [B,80,M], stop_logits [B,M,1], face [B,256], attention logits (pre-softmax, AFTER the
    logit dropout of line 363) [B,M,T], content_dis [B*minT,501]]."""
    with bn_train(bn_updates):
        pre = O.decoder_preloop(sd, enc_in, face_tiled[:, 0], noise.gumbel, p)
        b, m = mels.shape[0], mels.shape[2]
        hidden = pre["hidden"]
        h = [hidden[0], hidden[1]]
        c = [torch.zeros_like(h[0]), torch.zeros_like(h[0])]                       # cell.fill_(0): no gradient path (347)
        k, v, ckey, cval, enc_cell = pre["k"], pre["v"], pre["ckey"], pre["cval"], pre["enc_cell"]
        pos = sd[p + "positional_encodings.pos_table"][0]
        temp, ctemp = sd[p + "temperature"], sd[p + "content.temperature"]
        ys = sd[p + "BOS"].reshape(1, -1).repeat(b, 1)
        teacher = torch.cat([ys.unsqueeze(1), mels.permute(0, 2, 1)], dim=1)       # [B, M+1, 80]
        outputs, stops, attn = [], [], []
        wl = [(sd[f"{p}decoder_rnn.weight_ih_l{l}"], sd[f"{p}decoder_rnn.weight_hh_l{l}"],
               sd[f"{p}decoder_rnn.bias_ih_l{l}"], sd[f"{p}decoder_rnn.bias_hh_l{l}"]) for l in range(2)]
        for i in range(m):
            if bool(noise.tf_mask[i]):
                ys = teacher[:, i]
            y = O.psine(O._lin(sd, p + "prenet.0.linear_layer", ys), sd[p + "prenet.1.w"])
            y = _drop(y, noise.prenet[i], 0.2)                                      # nn.Dropout(0.2), line 308
            y = O.psine(O._lin(sd, p + "prenet.3.linear_layer", y), sd[p + "prenet.4.w"])
            q = O.psine(O._lin(sd, p + "Q.0.linear_layer", torch.cat(h, 1)), sd[p + "Q.1.w"]) + pos[i]
            a = _drop(torch.bmm((q * temp).unsqueeze(1), k), noise.attn[i].unsqueeze(1), 0.1)   # line 363
            attn.append(a)
            o = O._lin(sd, p + "attention_proj.linear_layer", torch.bmm(torch.softmax(a, dim=-1), v).squeeze(1))
            y = y + o
            cq = F.silu(O._lin(sd, p + "content.Q.0", torch.cat(c, 1))).unsqueeze(1)
            co = torch.bmm(torch.softmax(torch.bmm(cq * ctemp, ckey), dim=-1), cval).squeeze(1)
            x = torch.cat([co, y], -1)
            h0, c0 = O.lstm_cell(x, h[0], c[0], *wl[0])
            h1, c1 = O.lstm_cell(_drop(h0, noise.lstm[i], 0.1), h[1], c[1], *wl[1])   # inter-layer dropout: layer 1's INPUT only
            h, c = [h0, h1], [c0, c1]
            ys = O._lin(sd, p + "fc_out.linear_layer", h1)
            outputs.append(ys)
            stops.append(O._lin(sd, p + "stop_token_layer.linear_layer", torch.cat([h1, enc_cell], 1)))
        outputs = torch.stack(outputs, 1).permute(0, 2, 1)
        post = postnet_train(sd, outputs, noise.post, p + "postnet.") + outputs
    return [outputs, post, torch.stack(stops, 1), face_tiled[:, 0], torch.cat(attn, 1), pre["cdis"]]


def postnet_train(sd, x, masks, p="decoder.postnet."):
    """Postnet.forward in train mode (decoder.py:143-156): batch-stat BN, dropout 0.5 after every layer."""
    n = 5
    for i in range(n - 1):
        res = x
        y = F.conv1d(x, sd[f"{p}convolutions.{i}.0.conv.weight"], sd[f"{p}convolutions.{i}.0.conv.bias"], padding=2)
        y = O._bn(sd, f"{p}convolutions.{i}.1", y)
        y = O.psine(y, sd[f"{p}sin_activation.{i}.w"], channel_dim=1)
        x = y + res if i != 0 else y
        x = _drop(x, masks[i], 0.5)
    y = F.conv1d(x, sd[f"{p}convolutions.{n-1}.0.conv.weight"], sd[f"{p}convolutions.{n-1}.0.conv.bias"], padding=2)
    return _drop(O._bn(sd, f"{p}convolutions.{n-1}.1", y), masks[n - 1], 0.5)


def lip2speech_forward_train(sd, video, face_emb, mels, noise, bn_updates=None):
    """Lip2Speech.forward in TRAIN mode (model.py:23-40) with the face embedding given (vgg_face is frozen and out of scope:
    its parameters are not in the optimizer, train.py:102-104).  video [B,3,T,H,W], face_emb [B,256]."""
    with bn_train(bn_updates):
        feat = O.video_features(sd, video, "encoder.")
    feat = _drop(feat, noise.video_drop, 0.1)                                       # model.py:26
    t = feat.shape[1]
    face = face_emb.unsqueeze(1).repeat(1, t, 1)
    visual = torch.cat([feat, face], dim=2)
    return decoder_forward_train(sd, visual, face, mels, noise, "decoder.", bn_updates)
