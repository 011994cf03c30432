/* lip2speech_b200 — C ABI of the B200-native Lip2Speech hot path.
 *
 * The reference (Chris10M/Lip2Speech) has no plugin / operator / FFI interface: the hot path sits
 * behind Python nn.Module methods and state_dict key names (reference model/model.py:13-59).  This
 * header is the drop-in boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); each
 * entry point names the reference method it replaces.  Plain pointers and sizes only; no C++ or torch
 * types cross the boundary.  All device work is enqueued on the caller's stream; the library never
 * synchronises on its own except inside l2s_commit_weights, l2s_create/destroy and l2s_infer_host.
 *
 * Conventions: every function returns 0 on success, non-zero on error; l2s_last_error(ctx) returns a
 * description.  There is no CPU fallback and no other backend: an unsupported shape is an error.
 * A context belongs to one device and is not thread-safe.
 */
#ifndef L2S_B200_H
#define L2S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct l2s_ctx l2s_ctx;

#define L2S_OK 0
#define L2S_ERR_INVALID 1
#define L2S_ERR_CUDA 2
#define L2S_ERR_MISSING_WEIGHT 3

#define L2S_DTYPE_F32 0
#define L2S_DTYPE_I64 1

#define L2S_PART_VIDEO 1   /* encoder.*          (model/modules/video.py, shufflenetv2.py) */
#define L2S_PART_SPEAKER 2 /* speaker_encoder.*  (model/modules/audio.py:110-150)          */
#define L2S_PART_DECODER 4 /* decoder.*          (model/modules/decoder.py:274-444)        */

#define L2S_PRECISION_FP32 0 /* fp32-grade everywhere: 3xTF32 tensor-core products with fp32 accumulation + fp32 FMAs (parity mode) */
#define L2S_PRECISION_BF16 1 /* bf16 tensor-core stem (config 2 of BASELINE.json); decoder stays fp32 */

int l2s_version(void);

/* Creates a context on CUDA device `device` (replaces `.to(device)` of model.get_network, model.py:62-72). */
int l2s_create(l2s_ctx** out, int device);
void l2s_destroy(l2s_ctx* ctx);
const char* l2s_last_error(const l2s_ctx* ctx); /* ctx may be NULL: last creation error */

/* Hands one state_dict entry to the library under its reference key name, e.g.
 * "decoder.Q.0.linear_layer.weight", "encoder.frontend3D.0.weight", "speaker_encoder.lstm.weight_ih_l0"
 * (replaces nn.Module.load_state_dict, demo.py:38 / audio.py:129).  `ptr` may be host (on_device=0)
 * or device (on_device=1) memory; the library copies it.  Unknown keys (e.g. vgg_face.*) are accepted
 * and ignored by the kernels. */
int l2s_bind_weight(l2s_ctx* ctx, const char* key, const void* ptr, const int64_t* shape, int ndim,
                    int dtype, int on_device);

/* Folds eval-mode BatchNorm into the convolutions, repacks weights for the kernels and uploads them.
 * `parts` is a bitmask of L2S_PART_*.  Must be called after binding and before the forward calls;
 * call again after re-binding changed parameters. */
int l2s_commit_weights(l2s_ctx* ctx, int parts);

/* VideoExtractor.forward (video.py:76-87): video [B,3,T,H,W] fp32 NCDHW (device) -> out_feat [B,T,768]
 * L2-normalised (device).  H, W multiples of 32 or 88. */
int l2s_video_fwd(l2s_ctx* ctx, const float* video, int B, int T, int H, int W, float* out_feat,
                  int precision, void* stream);

/* SpeakerEncoder.forward / .inference (audio.py:132-150): wav [B,S] fp32 (device) -> emb [B,256].
 * normalize=0: raw linear output (forward); normalize=1: F.normalize(relu(.)) (inference). */
int l2s_speaker_fwd(l2s_ctx* ctx, const float* wav, int B, int S, float* emb, int normalize, void* stream);

/* Decoder.inference (decoder.py:382-444): visual [B,T,1024], spk [B,256] (= face_features[:,0]),
 * gumbel [B*minT,501] explicit g=-log(Exp(1)) noise of F.gumbel_softmax (decoder.py:257), all device fp32.
 * Writes mel_post [B,80,steps], lengths [B] int64 and, if non-NULL, attn [B,steps,T] (post-softmax). */
int l2s_decoder_infer(l2s_ctx* ctx, const float* visual, const float* spk, const float* gumbel, int B, int T,
                      int steps, float* mel_post, int64_t* lengths, float* attn, void* stream);

/* Decoder.forward in EVAL mode (decoder.py:320-379; the path evaluate.py:38 takes): M = mels.shape[2] steps, teacher
 * forcing decided by the caller per step (tf_mask[i] = 1: step i consumes the teacher frame — BOS for i = 0, mels[:,:,i-1]
 * otherwise — exactly the coin flips of decoder.py:355-357; HOST pointer, M bytes).  Outputs (device; any but out_post may
 * be NULL): out_mel [B,80,M] (pre-postnet), out_post [B,80,M], out_stop [B,M] raw stop logits, out_attn_logits [B,M,T]
 * PRE-softmax, out_content_dis [B*minT,501].  Train-mode dropout / batch-norm statistics are not implemented. */
int l2s_decoder_forward(l2s_ctx* ctx, const float* visual, const float* spk, const float* gumbel, const float* mels,
                        const unsigned char* tf_mask, int B, int T, int M, float* out_mel, float* out_post, float* out_stop,
                        float* out_attn_logits, float* out_content_dis, void* stream);

/* Postnet.forward in eval mode (decoder.py:143-156): x [B,80,L] -> out [B,80,L]; add_residual=1 returns
 * postnet(x)+x as Decoder.inference does (decoder.py:438-439). */
int l2s_postnet_fwd(l2s_ctx* ctx, const float* x, int B, int L, float* out, int add_residual, void* stream);

/* Lip2Speech.inference with a voice embedding — the hot span of demo.py:84-86 — on DEVICE buffers:
 * speaker encoder (normalised) -> video frontend -> tile+concat (model.py:52-55) -> decoder. */
int l2s_infer(l2s_ctx* ctx, const float* video, const float* wav, const float* gumbel, int B, int T, int H,
              int W, int S, int steps, float* mel_post, int64_t* lengths, int precision, void* stream);

/* Same span on HOST buffers (pinned or pageable): copies inputs host->device, runs, copies mel_post and
 * lengths back and synchronises.  This is the end-to-end call bench.py times as `e2e`. */
int l2s_infer_host(l2s_ctx* ctx, const float* video, const float* wav, const float* gumbel, int B, int T,
                   int H, int W, int S, int steps, float* mel_post, int64_t* lengths, int precision);

/* Pipelined form of l2s_infer_host for a caller that streams batches (the loop of demo.py / evaluate.py): `submit` enqueues
 * copies + compute + result copies for staging slot 0 or 1 and returns; `wait` blocks until that slot's mel_post / lengths
 * are in the caller's buffers.  With two submissions in flight the clip copy of batch i+1 overlaps the compute of batch i.
 * The caller's buffers must stay valid (and pinned, for the copies to be asynchronous) until `wait` returns. */
int l2s_infer_host_submit(l2s_ctx* ctx, int slot, const float* video, const float* wav, const float* gumbel, int B, int T,
                          int H, int W, int S, int steps, float* mel_post, int64_t* lengths, int precision);
int l2s_infer_host_wait(l2s_ctx* ctx, int slot);

/* ---- raw-frame input (the step before the path: datasets/lrw/dataset.py:20-24 loadframes -> 82-86 face_resize) ----------------
 * `frames` is uint8 [B,T,H,W,3] RGB, exactly what loadframes() returns per clip; the dataset's `im.float() / 255.0` followed by
 * Normalize(mean, std) is applied on the device while the stem's input rows are written (same fp32 arithmetic, IEEE division:
 * results are bit-identical to feeding the normalised fp32 tensor).  mean_std: HOST pointer to {mean[3], std[3]}
 * (LRW: 0.485,0.456,0.406 / 0.229,0.224,0.225).  A quarter of the bytes cross PCIe; the fp32 NCDHW tensor never exists. */
int l2s_video_fwd_u8(l2s_ctx* ctx, const unsigned char* frames, const float* mean_std, int B, int T, int H, int W, float* out_feat,
                     int precision, void* stream);
int l2s_infer_u8(l2s_ctx* ctx, const unsigned char* frames, const float* mean_std, const float* wav, const float* gumbel, int B, int T,
                 int H, int W, int S, int steps, float* mel_post, int64_t* lengths, int precision, void* stream);
/* Host-buffer form (frames, wav, gumbel, mel_post, lengths on the host); pairs with l2s_infer_host_wait. */
int l2s_infer_host_submit_u8(l2s_ctx* ctx, int slot, const unsigned char* frames, const float* mean_std, const float* wav,
                             const float* gumbel, int B, int T, int H, int W, int S, int steps, float* mel_post, int64_t* lengths,
                             int precision);

/* ---- train-step tail (train.py:167-193) ---------------------------------------------------------------------------
 * Loss, the data-parallel gradient exchange (the only collective of the path) and the optimizer step on FLAT fp32 buffers
 * (the train-mode forward / backward of the model is further down: l2s_decoder_train_fwd / _bwd). */

/* Loss.forward (train_utils/losses.py:35-79) and the gradient of sum(losses) w.r.t. the model outputs, all device fp32:
 * mel_out, mel_post, mel_target [B,80,M]; gate_logits, gate_target [B,M]; content_dis [rows,501].
 * losses[4] = {KLD, mel_loss, postnet_mel_loss (x10), gate_loss}; g_* (same shapes as the outputs) may be NULL. */
int l2s_loss_fwd_bwd(l2s_ctx* ctx, const float* mel_out, const float* mel_post, const float* gate_logits,
                     const float* content_dis, const float* mel_target, const float* gate_target, int B, int M, int rows,
                     float* losses, float* g_mel, float* g_post, float* g_gate, float* g_content_dis, void* stream);

/* Data-parallel communicator (one process per GPU; NCCL over NVLink/NVSwitch, loaded with dlopen("libnccl.so.2")).
 * Rank 0 obtains the id with l2s_nccl_unique_id (128 bytes) and hands it to the other ranks by any side channel
 * (the Python host uses torch.distributed); every rank then calls l2s_comm_init. */
int l2s_nccl_unique_id(void* out, int nbytes);
int l2s_comm_init(l2s_ctx* ctx, const void* unique_id, int nbytes, int rank, int world);
int l2s_comm_destroy(l2s_ctx* ctx);

/* Between loss.backward() (train.py:184) and clip_grad_norm_ (191): flat_grads <- scale * sum over ranks (in place; with
 * no communicator or world == 1 only the scaling), and sqnorm_out[0] <- ||flat_grads||^2 after scaling (device scalar, no
 * host sync).  scale = 1/world reproduces the gradient of the global-batch mean loss. */
int l2s_allreduce_grads(l2s_ctx* ctx, float* flat_grads, int64_t n, float scale, float* sqnorm_out, void* stream);

/* clip_grad_norm_(max_norm) (train.py:191; max_norm <= 0: no clipping) + AdamW(amsgrad=True) step (train.py:102-104,193)
 * on flat device buffers p, g, m, v, vmax of n floats; sqnorm = ||g||^2 from l2s_allreduce_grads; step = 1, 2, ...
 * g is left clipped, as clip_grad_norm_ leaves p.grad. */
int l2s_clip_adamw_step(l2s_ctx* ctx, float* p, float* g, float* m, float* v, float* vmax, int64_t n, const float* sqnorm,
                        double max_norm, double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                        void* stream);

/* ---- train-mode forward + backward (train.py:167 net(...), :184 loss.backward()) -----------------------------------------
 * Parameters are NOT copied for training: l2s_train_bind registers, under the reference's state_dict key ("decoder.Q.0.
 * linear_layer.weight", ...), the caller's fp32 DEVICE parameter memory and the gradient memory the backward pass
 * accumulates into (grad may be NULL: frozen tensor / buffer; BatchNorm running_mean / running_var buffers, when bound, are
 * updated in place by the forward pass as nn.BatchNorm*d does in train()).  Pointers stay owned by the caller and must stay
 * valid; re-binding a key replaces it.  An optimizer step therefore needs no re-bind / re-pack. */
int l2s_train_bind(l2s_ctx* ctx, const char* key, float* param, float* grad, int64_t numel);
/* The train-mode forward / backward calls below launch ~5 K small dependent kernels per step.  With graphs enabled (the default) the
 * second call with the same shapes captures that launch sequence as a CUDA graph and later calls replay it (inputs are staged
 * into library memory first, so caller tensors may move between steps; re-binding a key to different memory drops the graphs).
 * Results are bit-identical either way.  enabled = 0 runs every call eagerly and frees the captured graphs. */
int l2s_train_set_graphs(l2s_ctx* ctx, int enabled);

/* Decoder.forward in TRAIN mode (decoder.py:320-379): BatchNorm batch statistics, and every random draw of the reference as an
 * explicit device input so that results are reproducible and checkable: tf_mask (HOST, M bytes: the coin flips of :355-357),
 * gumbel [B*minT,501] (:257), and KEEP masks (1.0 = kept, 0.0 = dropped; the 1/(1-p) scale is applied here): prenet_mask
 * [M,B,256] (:308, p=0.2), attn_mask [M,B,T] (:363, p=0.1), lstm_mask [M,B,512] (:312, nn.LSTM inter-layer dropout, p=0.1),
 * post_masks[5] each [B,C,M] (:152-154, p=0.5; C = 512,512,512,512,80).  visual [B,T,1024], spk [B,256], mels [B,80,M].
 * Outputs (any may be NULL): out_mel / out_post [B,80,M], out_stop [B,M], out_attn_logits [B,M,T] (pre-softmax, after the
 * logit dropout), out_content_dis [B*minT,501].  Activations are kept inside the context until l2s_decoder_train_bwd. */
int l2s_decoder_train_fwd(l2s_ctx* ctx, const float* visual, const float* spk, const float* mels, const unsigned char* tf_mask,
                          const float* gumbel, const float* prenet_mask, const float* attn_mask, const float* lstm_mask,
                          const float* const* post_masks, int B, int T, int M, int want_input_grads, float* out_mel, float* out_post,
                          float* out_stop, float* out_attn_logits, float* out_content_dis, void* stream);
/* Backward of the last l2s_decoder_train_fwd: g_* = gradient of the scalar objective w.r.t. out_mel, out_post, out_stop,
 * out_content_dis (NULL = zero; l2s_loss_fwd_bwd produces exactly these).  Parameter gradients are ACCUMULATED into the bound
 * gradient memory; g_visual [B,T,1024] / g_spk [B,256] (optional, need want_input_grads=1) receive the input gradients. */
int l2s_decoder_train_bwd(l2s_ctx* ctx, const float* g_mel, const float* g_post, const float* g_stop, const float* g_content_dis,
                          float* g_visual, float* g_spk, void* stream);

/* VideoExtractor.forward in TRAIN mode (video.py:76-87: Conv3d stem, BatchNorm batch statistics everywhere, PReLU, MaxPool3d,
 * the 16 ShuffleNetV2 blocks, conv_last, AvgPool2d, L2 norm) followed by the feature dropout of model.py:26 when drop_mask
 * (KEEP mask [B,T,768], p = 0.1) is non-NULL.  video [B,3,T,H,W] fp32 NCDHW -> out_feat [B,T,768].  Parameters under
 * "encoder.*" come from l2s_train_bind.  l2s_video_train_bwd consumes the gradient w.r.t. out_feat (e.g. the first 768
 * columns of l2s_decoder_train_bwd's g_visual) and accumulates the parameter gradients. */
int l2s_video_train_fwd(l2s_ctx* ctx, const float* video, const float* drop_mask, int B, int T, int H, int W, float* out_feat, void* stream);
int l2s_video_train_bwd(l2s_ctx* ctx, const float* g_feat, void* stream);

/* ---- after the path: vocoder and metric (demo.py:89-90, evaluate.py:41-45) ----------------------------
 * MelSpec2Audio (datasets/spectograms.py:76-95): mel [B,80,L] log-mel (device) -> exp -> InverseMelScale -> GriffinLim(n_fft
 * 1024, hop 256, power 2, momentum) -> wav [B,(L-1)*256] (device).  The inverse mel operator is the weight "vocoder.inv_mel"
 * [513,80] (l2s_bind_weight; = the least-squares operator torchaudio's InverseMelScale applies: lstsq(fb^T, I)).  init_angles:
 * the complex tensor GriffinLim(rand_init=True) draws, [B,513,L] viewed as real [B,513,L,2] (device), or NULL for rand_init=False.
 * The STFT / inverse STFT of every iteration are DFT GEMMs on the tcgen05 kernel (3xTF32). */
int l2s_vocoder(l2s_ctx* ctx, const float* mel, const float* init_angles, int B, int L, int n_iter, float momentum, float* wav, void* stream);
/* pystoi.stoi(clean, processed, 16000, extended=True) for B signal pairs [B,S] fp32 at 16 kHz (device) -> out[B] fp64 (device). */
int l2s_estoi(l2s_ctx* ctx, const float* clean, const float* processed, int B, int S, double* out, void* stream);

/* Number of kernels this library has launched on ctx since creation (bench.py `gpu_launches`). */
int64_t l2s_launch_count(const l2s_ctx* ctx);

/* Stage timing for bench.py's roofline: when enabled, CUDA events are recorded on the caller's stream around
 * the stages "speaker", "video", "preloop", "decode_loop" (the persistent step kernel alone) and "postnet".
 * l2s_span_ms synchronises on the closing event and returns the last duration in ms (-1 if never recorded). */
int l2s_set_profiling(l2s_ctx* ctx, int enabled);
double l2s_span_ms(l2s_ctx* ctx, const char* name);

/* Debug/test access to the most recent intermediate tensors (device->host copy, synchronising).
 * Names: "dec.K" [B,T,512], "dec.V" [B,T,512], "dec.enc_cell" [B,512], "dec.ckey" [B,minT,256],
 * "dec.cval" [B,minT,256], "dec.outputs" [B,steps,80], "video.stem" [B*T,H/4,W/4,24].
 * Returns the number of floats available (copies min(n, available)), or -1 if unknown.
 * "flag.<name>" copies nothing and returns an integer fact about the last call: "flag.dec3" = 1 when the decode loop ran
 * on the stage-pipelined kernel (B <= 32), 0 when it ran on the row-partitioned one (B > 32). */
int64_t l2s_debug_read(l2s_ctx* ctx, const char* name, float* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* L2S_B200_H */
