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

#define L2S_PRECISION_FP32 0 /* exact fp32 FMA everywhere (parity mode) */
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
 * on the stage-pipelined kernel (8 < B <= 32), 0 when it ran on the row-partitioned one. */
int64_t l2s_debug_read(l2s_ctx* ctx, const char* name, float* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* L2S_B200_H */
