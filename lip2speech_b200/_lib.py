"""ctypes binding of lip2speech_b200/lib/libl2s_b200.so (C ABI: include/l2s_b200.h).

There is no fallback: if the library is missing or fails, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libl2s_b200.so")

PART_VIDEO, PART_SPEAKER, PART_DECODER = 1, 2, 4
PRECISION_FP32, PRECISION_BF16 = 0, 1
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # datasets/lrw/dataset.py:84-85

EXPORTS = ("l2s_version", "l2s_create", "l2s_destroy", "l2s_last_error", "l2s_bind_weight", "l2s_commit_weights",
           "l2s_video_fwd", "l2s_speaker_fwd", "l2s_decoder_infer", "l2s_decoder_forward", "l2s_postnet_fwd", "l2s_infer", "l2s_infer_host", "l2s_infer_host_submit", "l2s_infer_host_wait",
           "l2s_video_fwd_u8", "l2s_infer_u8", "l2s_infer_host_submit_u8",
           "l2s_train_bind", "l2s_train_set_graphs", "l2s_decoder_train_fwd", "l2s_decoder_train_bwd", "l2s_video_train_fwd", "l2s_video_train_bwd",
           "l2s_vocoder", "l2s_estoi",
           "l2s_launch_count", "l2s_debug_read", "l2s_set_profiling", "l2s_span_ms",
           "l2s_loss_fwd_bwd", "l2s_nccl_unique_id", "l2s_comm_init", "l2s_comm_destroy", "l2s_allreduce_grads", "l2s_clip_adamw_step")
NCCL_UNIQUE_ID_BYTES = 128

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (no compute, no GPU needed) and declare the prototypes."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m lip2speech_b200.build` "
                               "(there is no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        vp, i, i64p, fp = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p
        lib.l2s_version.restype = i
        lib.l2s_create.argtypes = [C.POINTER(vp), i]
        lib.l2s_destroy.argtypes = [vp]; lib.l2s_destroy.restype = None
        lib.l2s_last_error.argtypes = [vp]; lib.l2s_last_error.restype = C.c_char_p
        lib.l2s_bind_weight.argtypes = [vp, C.c_char_p, vp, i64p, i, i, i]
        lib.l2s_commit_weights.argtypes = [vp, i]
        lib.l2s_video_fwd.argtypes = [vp, fp, i, i, i, i, fp, i, vp]
        lib.l2s_speaker_fwd.argtypes = [vp, fp, i, i, fp, i, vp]
        lib.l2s_decoder_infer.argtypes = [vp, fp, fp, fp, i, i, i, fp, vp, fp, vp]
        lib.l2s_decoder_forward.argtypes = [vp, fp, fp, fp, fp, vp, i, i, i, fp, fp, fp, fp, fp, vp]
        lib.l2s_postnet_fwd.argtypes = [vp, fp, i, i, fp, i, vp]
        lib.l2s_infer.argtypes = [vp, fp, fp, fp, i, i, i, i, i, i, fp, vp, i, vp]
        lib.l2s_infer_host.argtypes = [vp, fp, fp, fp, i, i, i, i, i, i, fp, vp, i]
        lib.l2s_infer_host_submit.argtypes = [vp, i, fp, fp, fp, i, i, i, i, i, i, fp, vp, i]
        lib.l2s_infer_host_wait.argtypes = [vp, i]
        lib.l2s_video_fwd_u8.argtypes = [vp, vp, vp, i, i, i, i, fp, i, vp]
        lib.l2s_infer_u8.argtypes = [vp, vp, vp, fp, fp, i, i, i, i, i, i, fp, vp, i, vp]
        lib.l2s_infer_host_submit_u8.argtypes = [vp, i, vp, vp, fp, fp, i, i, i, i, i, i, fp, vp, i]
        lib.l2s_train_bind.argtypes = [vp, C.c_char_p, vp, vp, C.c_int64]
        lib.l2s_train_set_graphs.argtypes = [vp, i]
        lib.l2s_decoder_train_fwd.argtypes = [vp, fp, fp, fp, vp, fp, fp, fp, fp, C.POINTER(vp), i, i, i, i, fp, fp, fp, fp, fp, vp]
        lib.l2s_decoder_train_bwd.argtypes = [vp, fp, fp, fp, fp, fp, fp, vp]
        lib.l2s_video_train_fwd.argtypes = [vp, fp, fp, i, i, i, i, fp, vp]
        lib.l2s_video_train_bwd.argtypes = [vp, fp, vp]
        lib.l2s_vocoder.argtypes = [vp, fp, fp, i, i, i, C.c_float, fp, vp]
        lib.l2s_estoi.argtypes = [vp, fp, fp, i, i, vp, vp]
        lib.l2s_launch_count.argtypes = [vp]; lib.l2s_launch_count.restype = C.c_int64
        lib.l2s_set_profiling.argtypes = [vp, i]
        lib.l2s_span_ms.argtypes = [vp, C.c_char_p]; lib.l2s_span_ms.restype = C.c_double
        lib.l2s_debug_read.argtypes = [vp, C.c_char_p, fp, C.c_int64]; lib.l2s_debug_read.restype = C.c_int64
        f = C.c_float
        lib.l2s_loss_fwd_bwd.argtypes = [vp, fp, fp, fp, fp, fp, fp, i, i, i, fp, fp, fp, fp, fp, vp]
        lib.l2s_nccl_unique_id.argtypes = [vp, i]
        lib.l2s_comm_init.argtypes = [vp, vp, i, i, i]
        lib.l2s_comm_destroy.argtypes = [vp]
        lib.l2s_allreduce_grads.argtypes = [vp, fp, C.c_int64, f, fp, vp]
        d = C.c_double
        lib.l2s_clip_adamw_step.argtypes = [vp, fp, fp, fp, fp, fp, C.c_int64, fp, d, d, d, d, d, d, i, vp]
        _lib = lib
        return lib


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous() or t.device != device:
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


class Backend:
    """One l2s_ctx on one CUDA device."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("lip2speech_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load()
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        rc = self.lib.l2s_create(C.byref(h), device)
        if rc != 0:
            raise RuntimeError("l2s_create failed: " + self.lib.l2s_last_error(None).decode())
        self.h = h
        self._sig = {}
        self._generation = 0
        self.world = 1          # data-parallel size of the library's communicator (comm_init)

    def close(self):
        if getattr(self, "h", None):
            self.lib.l2s_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): " + self.lib.l2s_last_error(self.h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- weights ---------------------------------------------------------------------------------
    def bind_state_dict(self, sd, prefix: str, part: int):
        """Bind every tensor of `sd` under `prefix + key` and commit `part`."""
        for k, t in sd.items():
            t = t.detach()
            if t.dtype == torch.int64:
                dtype = 1
            else:
                dtype = 0
                if t.dtype != torch.float32:
                    t = t.float()
            t = t.contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            rc = self.lib.l2s_bind_weight(self.h, (prefix + k).encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dtype,
                                          1 if t.is_cuda else 0)
            self._check(rc, f"l2s_bind_weight({prefix + k})")
        self._check(self.lib.l2s_commit_weights(self.h, part), "l2s_commit_weights")

    def sync_module(self, module: torch.nn.Module, prefix: str, part: int):
        """(Re)bind when any parameter/buffer changed: data_ptr, in-place version counter, or an explicit
        invalidate_weights() (raw-pointer writers such as ClipAdamW.step() change neither of the first two)."""
        sd = module.state_dict(keep_vars=True)
        sig = (self._generation,) + tuple((k, t.data_ptr(), t._version) for k, t in sd.items())
        if self._sig.get(prefix) != sig:
            self.bind_state_dict(sd, prefix, part)
            self._sig[prefix] = sig

    def invalidate_weights(self):
        """Parameters were modified behind autograd's back (a kernel wrote through raw pointers, or `p.data` was edited):
        the next forward re-binds and re-packs them."""
        self._generation += 1

    # ---- forward calls ---------------------------------------------------------------------------
    def video_fwd(self, video: torch.Tensor, precision: int = PRECISION_FP32) -> torch.Tensor:
        video = _f32c(video, self.device)
        B, Cc, T, H, W = video.shape
        assert Cc == 3
        out = torch.empty(B, T, 768, device=self.device, dtype=torch.float32)
        self._check(self.lib.l2s_video_fwd(self.h, video.data_ptr(), B, T, H, W, out.data_ptr(), precision, self._stream()), "l2s_video_fwd")
        return out

    def speaker_fwd(self, wav: torch.Tensor, normalize: bool) -> torch.Tensor:
        wav = _f32c(wav, self.device)
        B, S = wav.shape
        out = torch.empty(B, 256, device=self.device, dtype=torch.float32)
        self._check(self.lib.l2s_speaker_fwd(self.h, wav.data_ptr(), B, S, out.data_ptr(), int(normalize), self._stream()), "l2s_speaker_fwd")
        return out

    def decoder_infer(self, visual, spk, gumbel, steps: int = 300, return_attention: bool = False):
        visual, spk, gumbel = _f32c(visual, self.device), _f32c(spk, self.device), _f32c(gumbel, self.device)
        B, T, D = visual.shape
        assert D == 1024 and spk.shape == (B, 256)
        mel = torch.empty(B, 80, steps, device=self.device, dtype=torch.float32)
        lengths = torch.empty(B, device=self.device, dtype=torch.int64)
        attn = torch.empty(B, steps, T, device=self.device, dtype=torch.float32) if return_attention else None
        self._check(self.lib.l2s_decoder_infer(self.h, visual.data_ptr(), spk.data_ptr(), gumbel.data_ptr(), B, T, steps, mel.data_ptr(),
                                               C.c_void_p(lengths.data_ptr()), attn.data_ptr() if attn is not None else None,
                                               self._stream()), "l2s_decoder_infer")
        return (mel, lengths, attn) if return_attention else (mel, lengths)

    def decoder_forward(self, visual, spk, gumbel, mels, tf_mask):
        """Decoder.forward, eval flavour.  tf_mask: host bool/uint8 tensor [M]."""
        visual, spk, gumbel, mels = (_f32c(t, self.device) for t in (visual, spk, gumbel, mels))
        B, T, _ = visual.shape
        M = mels.shape[2]
        min_t = min([T] + [(T - k) // k + 1 for k in (1, 3, 5, 7)])
        mask = tf_mask.to(torch.uint8).cpu().contiguous()
        assert mask.numel() == M and mels.shape[:2] == (B, 80)
        out_mel = torch.empty(B, 80, M, device=self.device)
        out_post = torch.empty(B, 80, M, device=self.device)
        out_stop = torch.empty(B, M, 1, device=self.device)
        out_attn = torch.empty(B, M, T, device=self.device)
        out_dis = torch.empty(B * min_t, 501, device=self.device)
        self._check(self.lib.l2s_decoder_forward(self.h, visual.data_ptr(), spk.data_ptr(), gumbel.data_ptr(), mels.data_ptr(),
                                                 C.c_void_p(mask.data_ptr()), B, T, M, out_mel.data_ptr(), out_post.data_ptr(),
                                                 out_stop.data_ptr(), out_attn.data_ptr(), out_dis.data_ptr(), self._stream()),
                    "l2s_decoder_forward")
        torch.cuda.current_stream(self.device).synchronize()      # `mask` (host) must outlive the async copy
        return out_mel, out_post, out_stop, out_attn, out_dis

    def postnet_fwd(self, x: torch.Tensor, add_residual: bool = False) -> torch.Tensor:
        x = _f32c(x, self.device)
        B, Cc, L = x.shape
        assert Cc == 80
        out = torch.empty_like(x)
        self._check(self.lib.l2s_postnet_fwd(self.h, x.data_ptr(), B, L, out.data_ptr(), int(add_residual), self._stream()), "l2s_postnet_fwd")
        return out

    def infer(self, video, wav, gumbel, steps: int = 300, precision: int = PRECISION_FP32):
        video, wav, gumbel = _f32c(video, self.device), _f32c(wav, self.device), _f32c(gumbel, self.device)
        B, _, T, H, W = video.shape
        mel = torch.empty(B, 80, steps, device=self.device, dtype=torch.float32)
        lengths = torch.empty(B, device=self.device, dtype=torch.int64)
        self._check(self.lib.l2s_infer(self.h, video.data_ptr(), wav.data_ptr(), gumbel.data_ptr(), B, T, H, W, wav.shape[1], steps,
                                       mel.data_ptr(), C.c_void_p(lengths.data_ptr()), precision, self._stream()), "l2s_infer")
        return mel, lengths

    def infer_host(self, video, wav, gumbel, mel_out, lengths_out, steps: int = 300, precision: int = PRECISION_FP32):
        """All arguments are HOST tensors (ideally pinned); synchronous."""
        B, _, T, H, W = video.shape
        for t in (video, wav, gumbel, mel_out):
            assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
        self._check(self.lib.l2s_infer_host(self.h, video.data_ptr(), wav.data_ptr(), gumbel.data_ptr(), B, T, H, W, wav.shape[1], steps,
                                            mel_out.data_ptr(), C.c_void_p(lengths_out.data_ptr()), precision), "l2s_infer_host")

    def infer_host_submit(self, slot, video, wav, gumbel, mel_out, lengths_out, steps: int = 300, precision: int = PRECISION_FP32):
        """Asynchronous form: enqueue copies + compute for staging slot 0 / 1; the HOST tensors must stay alive (and should be
        pinned) until infer_host_wait(slot) returns.  Two slots in flight overlap batch i+1's clip copy with batch i's compute."""
        B, _, T, H, W = video.shape
        for t in (video, wav, gumbel, mel_out):
            assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
        self._check(self.lib.l2s_infer_host_submit(self.h, slot, video.data_ptr(), wav.data_ptr(), gumbel.data_ptr(), B, T, H, W, wav.shape[1],
                                                   steps, mel_out.data_ptr(), C.c_void_p(lengths_out.data_ptr()), precision), "l2s_infer_host_submit")

    # ---- raw uint8 frames [B,T,H,W,3] RGB (datasets/lrw/dataset.py:20-24); /255 + Normalize fused on the device -------------
    @staticmethod
    def _mean_std(mean, std):
        return (C.c_float * 6)(*[float(x) for x in tuple(mean) + tuple(std)])

    def video_fwd_u8(self, frames: torch.Tensor, precision: int = PRECISION_FP32, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
        assert frames.dtype == torch.uint8 and frames.dim() == 5 and frames.shape[-1] == 3
        frames = frames.to(self.device).contiguous()
        B, T, H, W, _ = frames.shape
        out = torch.empty(B, T, 768, device=self.device, dtype=torch.float32)
        self._check(self.lib.l2s_video_fwd_u8(self.h, C.c_void_p(frames.data_ptr()), self._mean_std(mean, std), B, T, H, W, out.data_ptr(),
                                              precision, self._stream()), "l2s_video_fwd_u8")
        return out

    def infer_u8(self, frames, wav, gumbel, steps: int = 300, precision: int = PRECISION_FP32, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        assert frames.dtype == torch.uint8 and frames.dim() == 5 and frames.shape[-1] == 3
        frames = frames.to(self.device).contiguous()
        wav, gumbel = _f32c(wav, self.device), _f32c(gumbel, self.device)
        B, T, H, W, _ = frames.shape
        mel = torch.empty(B, 80, steps, device=self.device, dtype=torch.float32)
        lengths = torch.empty(B, device=self.device, dtype=torch.int64)
        self._check(self.lib.l2s_infer_u8(self.h, C.c_void_p(frames.data_ptr()), self._mean_std(mean, std), wav.data_ptr(), gumbel.data_ptr(),
                                          B, T, H, W, wav.shape[1], steps, mel.data_ptr(), C.c_void_p(lengths.data_ptr()), precision,
                                          self._stream()), "l2s_infer_u8")
        return mel, lengths

    def infer_host_submit_u8(self, slot, frames, wav, gumbel, mel_out, lengths_out, steps: int = 300, precision: int = PRECISION_FP32,
                             mean=IMAGENET_MEAN, std=IMAGENET_STD):
        """Host uint8 frames [B,T,H,W,3] (ideally pinned) + host fp32 wav / gumbel; pairs with infer_host_wait(slot)."""
        B, T, H, W, _ = frames.shape
        assert frames.device.type == "cpu" and frames.dtype == torch.uint8 and frames.is_contiguous()
        for t in (wav, gumbel, mel_out):
            assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
        self._check(self.lib.l2s_infer_host_submit_u8(self.h, slot, C.c_void_p(frames.data_ptr()), self._mean_std(mean, std), wav.data_ptr(),
                                                      gumbel.data_ptr(), B, T, H, W, wav.shape[1], steps, mel_out.data_ptr(),
                                                      C.c_void_p(lengths_out.data_ptr()), precision), "l2s_infer_host_submit_u8")

    def infer_host_wait(self, slot):
        self._check(self.lib.l2s_infer_host_wait(self.h, slot), "l2s_infer_host_wait")

    # ---- train-step tail (train.py:167-193) --------------------------------------------------------
    def loss_fwd_bwd(self, mel_out, mel_post, gate_logits, content_dis, mel_target, gate_target, want_grads: bool = True):
        """Loss.forward (train_utils/losses.py:35-79) -> (losses[4] = KLD, mel, 10*postnet, gate; grads or None)."""
        mel_out, mel_post, gate_logits, content_dis, mel_target, gate_target = (
            _f32c(t, self.device) for t in (mel_out, mel_post, gate_logits, content_dis, mel_target, gate_target))
        B, _, M = mel_out.shape
        assert mel_post.shape == mel_out.shape == mel_target.shape and gate_logits.numel() == B * M == gate_target.numel()
        assert content_dis.shape[-1] == 501
        rows = content_dis.numel() // 501
        losses = torch.empty(4, device=self.device)
        grads = [torch.empty_like(t) for t in (mel_out, mel_post, gate_logits, content_dis)] if want_grads else [None] * 4
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check(self.lib.l2s_loss_fwd_bwd(self.h, ptr(mel_out), ptr(mel_post), ptr(gate_logits), ptr(content_dis), ptr(mel_target),
                                              ptr(gate_target), B, M, rows, ptr(losses), *(ptr(g) for g in grads), self._stream()),
                    "l2s_loss_fwd_bwd")
        return losses, (grads if want_grads else None)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), NCCL_UNIQUE_ID_BYTES)
        self._check(self.lib.l2s_comm_init(self.h, buf, NCCL_UNIQUE_ID_BYTES, rank, world), "l2s_comm_init")
        self.world = world

    def comm_destroy(self):
        self._check(self.lib.l2s_comm_destroy(self.h), "l2s_comm_destroy")
        self.world = 1

    def allreduce_grads(self, flat_grads: torch.Tensor, scale: float, sqnorm_out: torch.Tensor):
        assert flat_grads.is_cuda and flat_grads.dtype == torch.float32 and flat_grads.is_contiguous()
        self._check(self.lib.l2s_allreduce_grads(self.h, C.c_void_p(flat_grads.data_ptr()), flat_grads.numel(), scale,
                                                 C.c_void_p(sqnorm_out.data_ptr()), self._stream()), "l2s_allreduce_grads")

    def clip_adamw_step(self, p, g, m, v, vmax, sqnorm, max_norm, lr, beta1, beta2, eps, weight_decay, step):
        for t in (p, g, m, v, vmax):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == p.numel()
        self._check(self.lib.l2s_clip_adamw_step(self.h, *(C.c_void_p(t.data_ptr()) for t in (p, g, m, v, vmax)), p.numel(),
                                                 C.c_void_p(sqnorm.data_ptr()) if sqnorm is not None else None, max_norm, lr, beta1, beta2,
                                                 eps, weight_decay, step, self._stream()), "l2s_clip_adamw_step")

    # ---- train-mode forward / backward (train.py:167,184) --------------------------------------------------------------
    def train_bind(self, key: str, param: torch.Tensor, grad):
        """Register caller-owned parameter memory (and the gradient memory backward accumulates into; None = frozen)."""
        assert param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()
        assert grad is None or (grad.is_cuda and grad.dtype == torch.float32 and grad.is_contiguous() and grad.numel() == param.numel())
        self._check(self.lib.l2s_train_bind(self.h, key.encode(), C.c_void_p(param.data_ptr()),
                                            C.c_void_p(grad.data_ptr()) if grad is not None else None, param.numel()), f"l2s_train_bind({key})")

    def train_set_graphs(self, enabled: bool):
        """CUDA-graph replay of the train-mode forward / backward launch sequences (on by default; results are bit-identical)."""
        self._check(self.lib.l2s_train_set_graphs(self.h, int(bool(enabled))), "l2s_train_set_graphs")

    def decoder_train_fwd(self, visual, spk, mels, noise, want_input_grads=True):
        """noise: object with tf_mask [M] (host bool), gumbel, prenet [M,B,256], attn [M,B,T], lstm [M,B,512], post (5 x [B,C,M])
        — KEEP masks, float32 on the device.  Returns (out_mel, out_post, out_stop [B,M,1], attn_logits [B,M,T], content_dis)."""
        visual, spk, mels = (_f32c(t, self.device) for t in (visual, spk, mels))
        B, T, _ = visual.shape
        M = mels.shape[2]
        min_t = min([T] + [(T - k) // k + 1 for k in (1, 3, 5, 7)])
        mask = noise.tf_mask.to(torch.uint8).cpu().contiguous()
        gum, pm, am, lm = (_f32c(t, self.device) for t in (noise.gumbel, noise.prenet, noise.attn, noise.lstm))
        post = [_f32c(t, self.device) for t in noise.post]
        assert mask.numel() == M and pm.shape == (M, B, 256) and am.shape == (M, B, T) and lm.shape == (M, B, 512) and gum.shape == (B * min_t, 501)
        assert [tuple(t.shape) for t in post] == [(B, c, M) for c in (512, 512, 512, 512, 80)]
        out_mel = torch.empty(B, 80, M, device=self.device); out_post = torch.empty(B, 80, M, device=self.device)
        out_stop = torch.empty(B, M, 1, device=self.device); out_attn = torch.empty(B, M, T, device=self.device)
        out_dis = torch.empty(B * min_t, 501, device=self.device)
        pp = (C.c_void_p * 5)(*[t.data_ptr() for t in post])
        self._check(self.lib.l2s_decoder_train_fwd(self.h, visual.data_ptr(), spk.data_ptr(), mels.data_ptr(), C.c_void_p(mask.data_ptr()),
                                                   gum.data_ptr(), pm.data_ptr(), am.data_ptr(), lm.data_ptr(), pp, B, T, M, int(want_input_grads),
                                                   out_mel.data_ptr(), out_post.data_ptr(), out_stop.data_ptr(), out_attn.data_ptr(),
                                                   out_dis.data_ptr(), self._stream()), "l2s_decoder_train_fwd")
        return out_mel, out_post, out_stop, out_attn, out_dis

    def decoder_train_bwd(self, g_mel, g_post, g_stop, g_dis, B, T, want_input_grads=True):
        gs = [None if g is None else _f32c(g, self.device) for g in (g_mel, g_post, g_stop, g_dis)]
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        g_visual = torch.empty(B, T, 1024, device=self.device) if want_input_grads else None
        g_spk = torch.empty(B, 256, device=self.device) if want_input_grads else None
        self._check(self.lib.l2s_decoder_train_bwd(self.h, *(ptr(g) for g in gs), ptr(g_visual), ptr(g_spk), self._stream()), "l2s_decoder_train_bwd")
        return g_visual, g_spk

    def video_train_fwd(self, video, drop_mask=None):
        """VideoExtractor.forward in train mode (+ model.py:26 dropout when drop_mask [B,T,768] KEEP mask is given)."""
        video = _f32c(video, self.device)
        B, Cc, T, H, W = video.shape
        assert Cc == 3
        if drop_mask is not None:
            drop_mask = _f32c(drop_mask, self.device)
            assert drop_mask.shape == (B, T, 768)
        out = torch.empty(B, T, 768, device=self.device)
        self._check(self.lib.l2s_video_train_fwd(self.h, video.data_ptr(), drop_mask.data_ptr() if drop_mask is not None else None, B, T, H, W,
                                                 out.data_ptr(), self._stream()), "l2s_video_train_fwd")
        return out

    def video_train_bwd(self, g_feat):
        g_feat = _f32c(g_feat, self.device)
        self._check(self.lib.l2s_video_train_bwd(self.h, g_feat.data_ptr(), self._stream()), "l2s_video_train_bwd")

    # ---- after the path: vocoder + ESTOI (demo.py:89-90, evaluate.py:41-45) ----------------------------------------------------
    def bind_vocoder(self, inv_mel: torch.Tensor):
        """inv_mel [513,80]: the operator of torchaudio's InverseMelScale (see lip2speech_b200.audio.inverse_mel_operator)."""
        t = inv_mel.detach().float().contiguous().cpu()
        shape = (C.c_int64 * 2)(*t.shape)
        self._check(self.lib.l2s_bind_weight(self.h, b"vocoder.inv_mel", C.c_void_p(t.data_ptr()), shape, 2, 0, 0), "l2s_bind_weight(vocoder.inv_mel)")

    def vocoder(self, mel, init_angles=None, n_iter: int = 32, momentum: float = 0.99):
        """mel [B,80,L] log-mel -> waveform [B,(L-1)*256].  init_angles: complex [B,513,L] (GriffinLim's rand_init draw) or None."""
        mel = _f32c(mel, self.device)
        B, M, L = mel.shape
        assert M == 80
        ia = None
        if init_angles is not None:
            ia = torch.view_as_real(init_angles.to(self.device).to(torch.complex64).contiguous()).contiguous()
            assert ia.shape == (B, 513, L, 2)
        wav = torch.empty(B, (L - 1) * 256, device=self.device)
        self._check(self.lib.l2s_vocoder(self.h, mel.data_ptr(), ia.data_ptr() if ia is not None else None, B, L, n_iter, momentum,
                                         wav.data_ptr(), self._stream()), "l2s_vocoder")
        return wav

    def estoi(self, clean, processed):
        """pystoi.stoi(clean, processed, 16000, extended=True) for every row of [B,S] -> [B] float64."""
        clean, processed = _f32c(clean, self.device), _f32c(processed, self.device)
        assert clean.shape == processed.shape and clean.dim() == 2
        out = torch.empty(clean.shape[0], dtype=torch.float64, device=self.device)
        self._check(self.lib.l2s_estoi(self.h, clean.data_ptr(), processed.data_ptr(), clean.shape[0], clean.shape[1],
                                       C.c_void_p(out.data_ptr()), self._stream()), "l2s_estoi")
        return out

    def set_profiling(self, enabled: bool):
        self.lib.l2s_set_profiling(self.h, int(enabled))

    def span_ms(self, name: str) -> float:
        return float(self.lib.l2s_span_ms(self.h, name.encode()))

    def launch_count(self) -> int:
        return int(self.lib.l2s_launch_count(self.h))

    def debug_flag(self, name: str) -> int:
        return int(self.lib.l2s_debug_read(self.h, ("flag." + name).encode(), None, 0))

    def debug_read(self, name: str, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        out = torch.empty(n, dtype=torch.float32)
        got = self.lib.l2s_debug_read(self.h, name.encode(), out.data_ptr(), n)
        if got < 0:
            raise KeyError(name)
        assert got == n, (name, got, n)
        return out.view(*shape)


_backends = {}



def backend(device: int = 0) -> Backend:
    """Process-wide context per device."""
    if device not in _backends:
        _backends[device] = Backend(device)
    return _backends[device]
