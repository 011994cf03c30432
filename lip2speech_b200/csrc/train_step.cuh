// Train-step tail of the reference (train.py:167-193): the reconstruction loss and its gradient, and everything between
// loss.backward() and the next forward: gradient exchange over the data-parallel ranks (the one collective of the path,
// SURVEY.md §8e), clip_grad_norm_(1.0) and the AdamW(amsgrad) update.  All of it is HBM-bound streaming over flat fp32
// buffers (38.44 M parameters): float4 accesses, grids sized in multiples of the SM count, two passes over the gradient.
//
//   pass 1 (grad_scale_sqnorm_kernel):  g <- g * scale (1/world after the NCCL sum), per-block partial sums of g^2
//           + sqnorm_finish_kernel:     fixed-order sum of the partials -> ||g||^2 on the device (no host sync)
//   pass 2 (clip_adamw_kernel):         clip coefficient from ||g||, then p, m, v, vmax updated in one read-modify-write
// Algorithmic bytes per parameter: pass 1 reads 4 + writes 4; pass 2 reads 20 (g, p, m, v, vmax) + writes 20
// (g clipped, p, m, v, vmax) = 48 B -> 1.85 GB per step = 0.28 ms at the measured 6.5 TB/s.
#pragma once
#include "common.cuh"

namespace l2s {

constexpr int TS_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = lane < TS_THREADS / 32 ? sh[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;                                                // valid in thread 0
}

// g <- g*scale ; partial[blockIdx] = sum over this block's elements of (g*scale)^2.  A thread adds at most a few dozen
// float4 in fp32, blocks are combined in double in a fixed order: deterministic for a fixed grid.
__global__ void __launch_bounds__(TS_THREADS) grad_scale_sqnorm_kernel(float* __restrict__ g, size_t n, float scale, double* __restrict__ partial) {
    __shared__ float sh[TS_THREADS / 32];
    const size_t n4 = n / 4;
    float4* g4 = reinterpret_cast<float4*>(g);
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        g4[i] = v;
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail (n not a multiple of 4)
        const size_t i = n4 * 4 + threadIdx.x;
        const float v = g[i] * scale;
        g[i] = v;
        acc += v * v;
    }
    const float s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}

// one warp: lane l adds partial[l], partial[l+32], ... in order, then a shuffle tree (fixed order)
__device__ __forceinline__ double warp_sum_partials(const double* __restrict__ partial, int nblocks, int stride, int offset) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[(size_t)i * stride + offset];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

__global__ void sqnorm_finish_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ sqnorm_out) {
    const double s = warp_sum_partials(partial, nblocks, 1, 0);
    if (threadIdx.x == 0) *sqnorm_out = (float)s;
}

struct AdamWParams {
    float lr, beta1, beta2, eps, weight_decay, max_norm;
    float bc1, bc2_sqrt;          // 1 - beta1^t, sqrt(1 - beta2^t)   (host, double precision)
    float step_size;              // lr / (1 - beta1^t)
    float omb1, omb2, decay;      // 1 - beta1, 1 - beta2, 1 - lr*wd: formed in DOUBLE on the host and rounded once, as torch does
                                  // (1.0f - 0.999f = 0.00099998712 is 1.3e-5 away from (float)0.001)
};

// torch.nn.utils.clip_grad_norm_ (clip_coef = max_norm / (norm + 1e-6), clamped to 1) followed by torch.optim.AdamW
// with amsgrad=True, in torch's operation order (train.py:102-104,191-193):
//   p *= 1 - lr*wd ; m += (g - m)(1 - b1) ; v = b2 v + (1 - b2) g^2 ; vmax = max(vmax, v) ;
//   p -= (lr / bc1) * m / (sqrt(vmax) / sqrt(bc2) + eps)
__device__ __forceinline__ void adamw_one(float& p, float& g, float& m, float& v, float& vmax, float coef, const AdamWParams& a) {
    g *= coef;
    p *= a.decay;
    m = m + (g - m) * a.omb1;
    v = v * a.beta2 + a.omb2 * g * g;
    vmax = fmaxf(vmax, v);
    const float denom = sqrtf(vmax) / a.bc2_sqrt + a.eps;
    p = p - a.step_size * (m / denom);
}

__global__ void __launch_bounds__(TS_THREADS) clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                                 float* __restrict__ vmax, size_t n, const float* __restrict__ sqnorm, AdamWParams a) {
    float coef = 1.0f;
    if (a.max_norm > 0.f) {
        const float norm = sqrtf(*sqnorm);
        coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);
    }
    const size_t n4 = n / 4;
    float4 *p4 = reinterpret_cast<float4*>(p), *g4 = reinterpret_cast<float4*>(g), *m4 = reinterpret_cast<float4*>(m),
           *v4 = reinterpret_cast<float4*>(v), *x4 = reinterpret_cast<float4*>(vmax);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i], xx = x4[i];
        adamw_one(pp.x, gg.x, mm.x, vv.x, xx.x, coef, a);
        adamw_one(pp.y, gg.y, mm.y, vv.y, xx.y, coef, a);
        adamw_one(pp.z, gg.z, mm.z, vv.z, xx.z, coef, a);
        adamw_one(pp.w, gg.w, mm.w, vv.w, xx.w, coef, a);
        p4[i] = pp; g4[i] = gg; m4[i] = mm; v4[i] = vv; x4[i] = xx;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        adamw_one(p[i], g[i], m[i], v[i], vmax[i], coef, a);
    }
}

// ---- Loss.forward (train_utils/losses.py:35-79) and d(loss)/d(outputs) -----------------------------------------------
//   losses[0] KLD  = mean_rows sum_j q log(q*501 + 1e-20)                       (69-73)
//   losses[1] mel  = MSE(mel_out, target)                                        (75)
//   losses[2] post = 10 * MSE(mel_post, target)                                  (76)
//   losses[3] gate = BCEWithLogits(gate_logits, gate_target)                     (77)
// Gradients of sum(losses) w.r.t. mel_out, mel_post, gate_logits and content_dis are written when the pointers are set.
// part[blockIdx][4] partial sums (double), finished by loss_finish_kernel in a fixed order.
__global__ void __launch_bounds__(TS_THREADS) loss_partial_kernel(const float* __restrict__ mel_out, const float* __restrict__ mel_post,
        const float* __restrict__ target, size_t n_mel, const float* __restrict__ gate_logits, const float* __restrict__ gate_target, size_t n_gate,
        const float* __restrict__ dis, size_t n_dis, int vocab, float* __restrict__ g_mel, float* __restrict__ g_post, float* __restrict__ g_gate,
        float* __restrict__ g_dis, double* __restrict__ part) {
    __shared__ float sh[TS_THREADS / 32];
    double s_mel = 0.0, s_post = 0.0, s_gate = 0.0, s_kld = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const float k_mel = 2.0f / (float)n_mel, k_post = 20.0f / (float)n_mel, k_gate = 1.0f / (float)n_gate;
    const float k_kld = 1.0f / (float)(n_dis / vocab);
    for (size_t i = i0; i < n_mel; i += stride) {
        const float t = target[i], d0 = mel_out[i] - t, d1 = mel_post[i] - t;
        s_mel += (double)(d0 * d0); s_post += (double)(d1 * d1);
        if (g_mel) g_mel[i] = k_mel * d0;
        if (g_post) g_post[i] = k_post * d1;
    }
    for (size_t i = i0; i < n_gate; i += stride) {
        const float x = gate_logits[i], y = gate_target[i];
        // BCEWithLogits: max(x,0) - x*y + log(1 + exp(-|x|))
        s_gate += (double)(fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))));
        if (g_gate) g_gate[i] = k_gate * (sigmoidf_acc(x) - y);
    }
    for (size_t i = i0; i < n_dis; i += stride) {
        const float q = dis[i], u = q * (float)vocab + 1e-20f, lr = logf(u);
        s_kld += (double)(q * lr);
        if (g_dis) g_dis[i] = k_kld * (lr + q * (float)vocab / u);
    }
    double sums[4] = {s_kld, s_mel, s_post, s_gate};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float hi = (float)sums[j], lo = (float)(sums[j] - (double)hi);
        const float a = block_sum_256(hi, sh);
        __syncthreads();
        const float b = block_sum_256(lo, sh);
        __syncthreads();
        if (threadIdx.x == 0) part[(size_t)blockIdx.x * 4 + j] = (double)a + (double)b;
    }
}

__global__ void loss_finish_kernel(const double* __restrict__ part, int nblocks, size_t n_mel, size_t n_gate, size_t n_rows, float* __restrict__ losses) {
    double s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = warp_sum_partials(part, nblocks, 4, j);
    if (threadIdx.x == 0) {
        losses[0] = (float)(s[0] / (double)n_rows);
        losses[1] = (float)(s[1] / (double)n_mel);
        losses[2] = (float)(10.0 * s[2] / (double)n_mel);
        losses[3] = (float)(s[3] / (double)n_gate);
    }
}

}  // namespace l2s
