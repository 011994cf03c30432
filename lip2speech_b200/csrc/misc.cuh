// Small kernels around the recurrent cores: speaker-encoder mel front end, layout changes between
// row-major [B][F] and feature-major [F][Bpad], adaptive average pooling, gumbel-softmax slot values.
#pragma once
#include "common.cuh"

namespace l2s {

// torchaudio MelSpectrogram(16 kHz, n_fft=400, hop=160, n_mels=40, power=2, center, reflect pad, no log)
// (reference audio.py:124,133).  One CTA per (clip, frame): windowed 400-point DFT by direct summation with
// a shared twiddle table (exact index arithmetic mod 400), |.|^2, then the [201,40] filterbank.
// wav [B][S] -> mel [B][F][40], F = 1 + S/160.
__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                      const float* __restrict__ fb, float* __restrict__ mel, int S, int F) {
    constexpr int NFFT = 400, HOP = 160, NBIN = 201, NMEL = 40;
    __shared__ float xw[NFFT];
    __shared__ float tc[NFFT], ts[NFFT];
    __shared__ float pw[NBIN];
    const int b = blockIdx.x / F, f = blockIdx.x % F;
    const int tid = threadIdx.x;
    for (int n = tid; n < NFFT; n += blockDim.x) {
        int idx = f * HOP + n - NFFT / 2;
        if (idx < 0) idx = -idx;
        if (idx >= S) idx = 2 * (S - 1) - idx;
        xw[n] = wav[(size_t)b * S + idx] * window[n];
        float sv, cv;
        sincospif(2.0f * (float)n / (float)NFFT, &sv, &cv);
        tc[n] = cv; ts[n] = sv;
    }
    __syncthreads();
    for (int k = tid; k < NBIN; k += blockDim.x) {
        float re = 0.f, im = 0.f;
        int j = 0;
        for (int n = 0; n < NFFT; ++n) {
            re = fmaf(xw[n], tc[j], re);
            im = fmaf(xw[n], ts[j], im);
            j += k; if (j >= NFFT) j -= NFFT;
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    for (int m = tid; m < NMEL; m += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < NBIN; ++k) a = fmaf(pw[k], __ldg(fb + k * NMEL + m), a);
        mel[((size_t)b * F + f) * NMEL + m] = a;
    }
}

// GEMM form of the same spectrogram (used when the tcgen05 path is on): the 400-point DFT of all B*F frames is one
// [B*F, 400] x [400, 402] GEMM against a cos | sin table (3xTF32), 5x faster than the per-frame direct summation above.
// frames[b*F + f][n] = wav[b][reflect(f*160 + n - 200)] * window[n]
__global__ void stft_frames_kernel(const float* __restrict__ wav, const float* __restrict__ window, float* __restrict__ frames, int B, int S, int F) {
    constexpr int NFFT = 400, HOP = 160;
    const size_t total = (size_t)B * F * NFFT;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = i % NFFT; const size_t r = i / NFFT;
        const int f = r % F, b = r / F;
        int idx = f * HOP + n - NFFT / 2;
        if (idx < 0) idx = -idx;
        if (idx >= S) idx = 2 * (S - 1) - idx;
        frames[i] = wav[(size_t)b * S + idx] * __ldg(window + n);
    }
}
// spec [rows][404]: columns 0..200 = Re, 201..401 = Im -> mel[rows][40] = (Re^2 + Im^2) . fb[201][40]; one warp per frame
__global__ void __launch_bounds__(256) power_mel_kernel(const float* __restrict__ spec, int lds, const float* __restrict__ fb, float* __restrict__ mel, int rows) {
    constexpr int NBIN = 201, NMEL = 40;
    __shared__ float pw[8][NBIN + 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r < rows) {
        const float* sp = spec + (size_t)r * lds;
        for (int k = lane; k < NBIN; k += 32) { const float re = sp[k], im = sp[NBIN + k]; pw[warp][k] = re * re + im * im; }
    }
    __syncwarp();
    if (r < rows) {
        for (int m = lane; m < NMEL; m += 32) {
            float a = 0.f;
            for (int k = 0; k < NBIN; ++k) a = fmaf(pw[warp][k], __ldg(fb + k * NMEL + m), a);
            mel[(size_t)r * NMEL + m] = a;
        }
    }
}

// dst[f][b] = src[b*lds + off + f]   (row-major -> feature-major), zero for b >= B.
__global__ void rows_to_fm_kernel(const float* __restrict__ src, int lds, int off, float* __restrict__ dst, int F, int B, int Bpad) {
    const int total = F * Bpad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int f = i / Bpad, b = i % Bpad;
        dst[i] = b < B ? src[(size_t)b * lds + off + f] : 0.f;
    }
}
// dst[b*ldd + off + f] = src[f][b]
__global__ void fm_to_rows_kernel(const float* __restrict__ src, int Bpad, float* __restrict__ dst, int ldd, int off, int F, int B) {
    const int total = F * B;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int b = i / F, f = i % F;
        dst[(size_t)b * ldd + off + f] = src[(size_t)f * Bpad + b];
    }
}

// Speaker-encoder head: emb[b] = W h[:,b] + bias (raw), optionally normalize(relu(.)) (reference audio.py:139-150).
// h is feature-major [256][Bpad].  One CTA (256 threads) per clip.
__global__ void __launch_bounds__(256) speaker_head_kernel(const float* __restrict__ h, int Bpad, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ out, int normalize) {
    __shared__ float hs[256];
    __shared__ float wsum[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    hs[tid] = h[(size_t)tid * Bpad + b];
    __syncthreads();
    float a = bias[tid];
    const float4* wr = reinterpret_cast<const float4*>(W + (size_t)tid * 256);
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
        float4 w = __ldg(wr + k);
        a = fmaf(w.x, hs[4 * k], a); a = fmaf(w.y, hs[4 * k + 1], a);
        a = fmaf(w.z, hs[4 * k + 2], a); a = fmaf(w.w, hs[4 * k + 3], a);
    }
    if (normalize) {
        a = fmaxf(a, 0.f);
        float ss = warp_sum(a * a);
        if ((tid & 31) == 0) wsum[tid >> 5] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += wsum[i];
        a = a / fmaxf(sqrtf(tot), 1e-12f);
    }
    out[(size_t)b * 256 + tid] = a;
}

// F.adaptive_avg_pool1d over time on time-major activations: in [B][L][ldi] (C channels) -> out[b][i][off + c],
// bin i = [floor(i*L/m), ceil((i+1)*L/m))  (reference decoder.py:247).
__global__ void adaptive_pool_kernel(const float* __restrict__ in, int ldi, int L, float* __restrict__ out, int ldo, int off,
                                     int m, int C, int B) {
    const size_t total = (size_t)B * m * C;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int c = idx % C; size_t r = idx / C;
        int i = r % m; int b = r / m;
        int s = (i * L) / m, e = ((i + 1) * L + m - 1) / m;
        float a = 0.f;
        for (int t = s; t < e; ++t) a += in[((size_t)b * L + t) * ldi + c];
        out[((size_t)b * m + i) * ldo + off + c] = a / (float)(e - s);
    }
}

// Content slots (reference decoder.py:253-260): z = softmax((logits + g)/tau); value = z @ E[V,256];
// optionally dis = softmax(logits).  One CTA (256 threads) per row of logits [rows][V].
__global__ void __launch_bounds__(256) gumbel_value_kernel(const float* __restrict__ logits, const float* __restrict__ g, float inv_tau,
                                                           const float* __restrict__ E, float* __restrict__ value,
                                                           float* __restrict__ dis, int V) {
    extern __shared__ float z[];            // [V]
    __shared__ float wred[8];
    __shared__ float bc;
    const int row = blockIdx.x, tid = threadIdx.x;
    for (int pass = 0; pass < (dis ? 2 : 1); ++pass) {
        float mx = -INFINITY;
        for (int v = tid; v < V; v += 256) {
            float x = logits[(size_t)row * V + v];
            x = pass == 0 ? (x + g[(size_t)row * V + v]) * inv_tau : x;
            z[v] = x;
            mx = fmaxf(mx, x);
        }
        mx = warp_max(mx);
        if ((tid & 31) == 0) wred[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) { float m2 = wred[0]; for (int i = 1; i < 8; ++i) m2 = fmaxf(m2, wred[i]); bc = m2; }
        __syncthreads();
        mx = bc;
        float sum = 0.f;
        for (int v = tid; v < V; v += 256) { float e = expf(z[v] - mx); z[v] = e; sum += e; }
        sum = warp_sum(sum);
        __syncthreads();
        if ((tid & 31) == 0) wred[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) { float s2 = 0.f; for (int i = 0; i < 8; ++i) s2 += wred[i]; bc = s2; }
        __syncthreads();
        const float inv = 1.0f / bc;
        if (pass == 0) {
            for (int v = tid; v < V; v += 256) z[v] *= inv;
            __syncthreads();
            float a = 0.f;
            for (int v = 0; v < V; ++v) a = fmaf(z[v], __ldg(E + (size_t)v * 256 + tid), a);
            value[(size_t)row * 256 + tid] = a;
        } else {
            for (int v = tid; v < V; v += 256) dis[(size_t)row * V + v] = z[v] * inv;
        }
        __syncthreads();
    }
}

__global__ void fill_i64_kernel(long long* p, long long v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace l2s
