// The two steps right after the hot path in demo.py:89-90 / evaluate.py:41-45, on the GPU and batched:
//   * vocoder  = datasets/spectograms.py:76-95 MelSpec2Audio: exp -> InverseMelScale -> GriffinLim(n_fft 1024, hop 256)
//   * ESTOI    = pystoi.stoi(clean, processed, 16000, extended=True)  (Jensen & Taal 2016), the metric of evaluate.py:45
// The STFT / inverse STFT of Griffin-Lim are real-DFT GEMMs on the tcgen05 kernel (3xTF32, analysis / synthesis windows folded
// into the tables); everything else is small fused kernels.  ESTOI runs in fp64 (tiny work; matches the fp64 numpy restatement
// in oracle/audio_metrics.py to rounding).
#pragma once
#include "common.cuh"

namespace l2s {

constexpr int VOC_NFFT = 1024, VOC_HOP = 256, VOC_BINS = 513;
constexpr int VOC_IM = 516;            // column of the imaginary parts inside a spectrum row (16-byte aligned)
constexpr int VOC_LD = 1032;           // row length of a spectrum: [re 513 | 3 pad | im 513 | 3 pad]

// melrows[(b*L + l)][m] = exp(mel[b][m][l])          (spectral_de_normalize, spectograms.py:92)
__global__ void voc_exp_rows_kernel(const float* __restrict__ mel, float* __restrict__ rows, int B, int L) {
    const size_t total = (size_t)B * 80 * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = i % 80; size_t r = i / 80;
        const int l = r % L; const int b = r / L;
        rows[i] = expf(mel[((size_t)b * 80 + m) * L + l]);
    }
}
// mag = sqrt(relu(x))   (InverseMelScale clamps at 0; GriffinLim(power=2) takes the square root)   rows x 513 inside ld 516
__global__ void voc_mag_kernel(float* __restrict__ x, int rows) {
    const size_t total = (size_t)rows * VOC_IM;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = i % VOC_IM;
        x[i] = f < VOC_BINS ? sqrtf(fmaxf(x[i], 0.f)) : 0.f;
    }
}
// angles[(b*L + l)][f | IM + f] <- init[b][f][l][re, im]   (torch complex [B,513,L] viewed as real), or (1, 0)
__global__ void voc_init_angles_kernel(const float* __restrict__ init, float* __restrict__ ang, float* __restrict__ tprev, int B, int L) {
    const size_t total = (size_t)B * L * VOC_LD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % VOC_LD; const size_t r = i / VOC_LD;
        const int l = r % L; const int b = r / L;
        float v = 0.f;
        if (c < VOC_BINS) v = init ? init[(((size_t)b * VOC_BINS + c) * L + l) * 2] : 1.f;
        else if (c >= VOC_IM && c < VOC_IM + VOC_BINS) v = init ? init[(((size_t)b * VOC_BINS + (c - VOC_IM)) * L + l) * 2 + 1] : 0.f;
        ang[i] = v; tprev[i] = 0.f;
    }
}
// spec = mag * angles
__global__ void voc_apply_kernel(const float* __restrict__ mag, const float* __restrict__ ang, float* __restrict__ spec, int rows) {
    const size_t total = (size_t)rows * VOC_LD;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = i % VOC_LD; const size_t r = i / VOC_LD;
        const int f = c < VOC_IM ? c : c - VOC_IM;
        spec[i] = f < VOC_BINS ? mag[r * VOC_IM + f] * ang[i] : 0.f;
    }
}
// torch.istft's overlap-add (center=True, length=None): frames [(b*L + t)][1024] already carry the synthesis window;
// out[b][j] = sum_t frames[t][p - 256 t] / sum_t w[p - 256 t]^2 with p = j + 512, j < (L-1)*256
__global__ void voc_ola_kernel(const float* __restrict__ frames, const float* __restrict__ win, float* __restrict__ out, int B, int L) {
    const int n = (L - 1) * VOC_HOP;
    const size_t total = (size_t)B * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = i % n; const int b = i / n;
        const int p = j + VOC_NFFT / 2;
        float acc = 0.f, env = 0.f;
        const int t1 = min(L - 1, p / VOC_HOP), t0 = max(0, (p - VOC_NFFT + VOC_HOP) / VOC_HOP);
        for (int t = t0; t <= t1; ++t) {
            const int k = p - t * VOC_HOP;
            if (k < 0 || k >= VOC_NFFT) continue;
            acc += frames[((size_t)b * L + t) * VOC_NFFT + k];
            const float w = win[k];
            env = fmaf(w, w, env);
        }
        out[i] = env > 1e-11f ? acc / env : acc;
    }
}
// torch.stft's framing (center=True, reflect pad 512): xf[(b*L + t)][k] = x_pad[256 t + k]; the analysis window is in the DFT table
__global__ void voc_frames_kernel(const float* __restrict__ x, float* __restrict__ xf, int B, int L) {
    const int n = (L - 1) * VOC_HOP;
    const size_t total = (size_t)B * L * VOC_NFFT;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = i % VOC_NFFT; const size_t r = i / VOC_NFFT;
        const int t = r % L; const int b = r / L;
        int src = t * VOC_HOP + k - VOC_NFFT / 2;
        if (src < 0) src = -src;
        if (src >= n) src = 2 * (n - 1) - src;
        xf[i] = x[(size_t)b * n + src];
    }
}
// Griffin-Lim phase update with momentum (torchaudio.functional.griffinlim): a = rebuilt - tprev * m/(1+m); a /= |a| + 1e-16
__global__ void voc_angle_kernel(const float* __restrict__ rebuilt, float* __restrict__ tprev, float* __restrict__ ang, int rows, float mfac) {
    const size_t total = (size_t)rows * VOC_BINS;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = i % VOC_BINS; const size_t r = i / VOC_BINS;
        const size_t ire = r * VOC_LD + f, iim = ire + VOC_IM;
        const float re = rebuilt[ire], im = rebuilt[iim];
        const float ar = re - tprev[ire] * mfac, ai = im - tprev[iim] * mfac;
        const float d = sqrtf(ar * ar + ai * ai) + 1e-16f;
        ang[ire] = ar / d; ang[iim] = ai / d;
        tprev[ire] = re; tprev[iim] = im;
    }
}

// ---- ESTOI (fp64) --------------------------------------------------------------------------------------------------------
constexpr int ES_FS = 10000, ES_FRAME = 256, ES_HOP = 128, ES_NFFT = 512, ES_BANDS = 15, ES_SEG = 30, ES_TAPS = 161, ES_HALF = 80;

struct EstoiConst {                 // device tables
    const double* h;                // [161] resample_poly FIR (firwin kaiser 5.0, x up)
    const double* win;              // [256] hanning(258)[1:-1]
    const int* band_lo; const int* band_hi;   // [15] one-third-octave bin ranges [lo, hi)
};

// scipy.signal.resample_poly(x, 5, 8): y[j] = y_full[j + n_pre_remove], y_full[k] = sum_i hp[i] xu[8k - i]
__global__ void es_resample_kernel(const float* __restrict__ x, double* __restrict__ y, int B, int S, int n_out, const double* __restrict__ h) {
    constexpr int up = 5, down = 8;
    const int n_pre_pad = down - ES_HALF % down;                       // 8
    const int n_pre_remove = (ES_HALF + n_pre_pad) / down;             // 11
    const size_t total = (size_t)B * n_out;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = i % n_out; const int b = i / n_out;
        const int k = j + n_pre_remove;
        const int hp_len = ES_TAPS + n_pre_pad;
        // 8k - idx = 5m  ->  idx = 8k - 5m in [n_pre_pad, hp_len)
        int m_lo = (down * k - hp_len + 1 + up - 1) / up; if (m_lo < 0) m_lo = 0;
        int m_hi = (down * k - n_pre_pad) / up; if (m_hi > S - 1) m_hi = S - 1;
        double acc = 0.0;
        for (int m = m_lo; m <= m_hi; ++m) {
            const int idx = down * k - up * m - n_pre_pad;
            if (idx >= 0 && idx < ES_TAPS) acc += h[idx] * (double)x[(size_t)b * S + m];
        }
        y[i] = acc;
    }
}
// One CTA per clip: windowed frames of the clean signal -> energies -> keep mask (within 40 dB of the loudest) -> compacted
// overlap-add of the kept frames of BOTH signals (pystoi remove_silent_frames).  Returns the new length in n_kept[b].
__global__ void __launch_bounds__(256) es_silent_kernel(const double* __restrict__ x, const double* __restrict__ y, int n, const double* __restrict__ win,
                                                        double* __restrict__ xs, double* __restrict__ ys, int* __restrict__ n_kept, int max_frames) {
    extern __shared__ double es_sm[];
    double* energy = es_sm;                          // [max_frames]
    int* pos = reinterpret_cast<int*>(energy + max_frames);            // [max_frames] compacted index or -1
    __shared__ double red[256];
    __shared__ int kept;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* xb = x + (size_t)b * n; const double* yb = y + (size_t)b * n;
    const int nf = n >= ES_FRAME ? 1 + (n - ES_FRAME) / ES_HOP : 0;
    for (int f = 0; f < nf; ++f) {
        const double v = xb[f * ES_HOP + tid] * win[tid];
        red[tid] = v * v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
        if (tid == 0) energy[f] = 20.0 * log10(sqrt(red[0]) + 2.220446049250313e-16);
        __syncthreads();
    }
    if (tid == 0) {
        double mx = -1e300;
        for (int f = 0; f < nf; ++f) mx = fmax(mx, energy[f]);
        int k = 0;
        for (int f = 0; f < nf; ++f) pos[f] = ((mx - 40.0 - energy[f]) < 0.0) ? k++ : -1;
        kept = k;
        n_kept[b] = k ? (k - 1) * ES_HOP + ES_FRAME : 0;
    }
    __syncthreads();
    const int nk = kept ? (kept - 1) * ES_HOP + ES_FRAME : 0;
    double* xo = xs + (size_t)b * n; double* yo = ys + (size_t)b * n;
    for (int i = tid; i < n; i += 256) { xo[i] = 0.0; yo[i] = 0.0; }
    __syncthreads();
    // every output sample belongs to at most two kept frames (hop = frame / 2): frames are added in order
    for (int f = 0; f < nf; ++f) {
        if (pos[f] < 0) continue;
        const int o = pos[f] * ES_HOP + tid;
        xo[o] += xb[f * ES_HOP + tid] * win[tid];
        yo[o] += yb[f * ES_HOP + tid] * win[tid];
        __syncthreads();
    }
    (void)nk;
}
// One CTA per (clip, frame): 512-point DFT of a windowed 256-sample frame -> one-third-octave band magnitudes
// tob[b][band][frame] = sqrt(sum_{bins in band} |X|^2)
__global__ void __launch_bounds__(256) es_bands_kernel(const double* __restrict__ xs, int n, const int* __restrict__ n_kept, const double* __restrict__ win,
                                                       const int* __restrict__ lo, const int* __restrict__ hi, double* __restrict__ tob, int max_frames) {
    __shared__ double fr[ES_FRAME];
    __shared__ double pw[ES_NFFT / 2 + 1];
    const int b = blockIdx.y, f = blockIdx.x, tid = threadIdx.x;
    const int len = n_kept[b];
    const int nf = len >= ES_FRAME ? 1 + (len - ES_FRAME) / ES_HOP : 0;
    if (f >= nf) return;
    fr[tid] = xs[(size_t)b * n + f * ES_HOP + tid] * win[tid];
    __syncthreads();
    for (int k = tid; k <= ES_NFFT / 2; k += 256) {
        double re = 0.0, im = 0.0;
        for (int t = 0; t < ES_FRAME; ++t) {
            const int ph = (k * t) & (ES_NFFT - 1);                   // exact argument reduction
            double s, c;
            sincospi(2.0 * (double)ph / (double)ES_NFFT, &s, &c);
            re += fr[t] * c; im -= fr[t] * s;
        }
        pw[k] = re * re + im * im;
    }
    __syncthreads();
    if (tid < ES_BANDS) {
        double a = 0.0;
        for (int k = lo[tid]; k < hi[tid]; ++k) a += pw[k];
        tob[((size_t)b * ES_BANDS + tid) * max_frames + f] = sqrt(a);
    }
}
// One CTA per clip: mean over the 30-frame segments of the correlation of the row+column normalised segments
__global__ void __launch_bounds__(512) es_score_kernel(const double* __restrict__ xt, const double* __restrict__ yt, const int* __restrict__ n_kept,
                                                       double* __restrict__ out, int max_frames) {
    __shared__ double xsg[ES_BANDS][ES_SEG], ysg[ES_BANDS][ES_SEG];
    __shared__ double acc;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int len = n_kept[b];
    const int nf = len >= ES_FRAME ? 1 + (len - ES_FRAME) / ES_HOP : 0;
    if (nf < ES_SEG) { if (tid == 0) out[b] = 1e-5; return; }            // pystoi: not enough frames
    if (tid == 0) acc = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int m = ES_SEG; m <= nf; ++m) {
        __syncthreads();
        if (tid < ES_BANDS * ES_SEG) {
            const int r = tid / ES_SEG, c = tid % ES_SEG;
            xsg[r][c] = xt[((size_t)b * ES_BANDS + r) * max_frames + m - ES_SEG + c];
            ysg[r][c] = yt[((size_t)b * ES_BANDS + r) * max_frames + m - ES_SEG + c];
        }
        __syncthreads();
        if (tid < 2 * ES_BANDS) {                                     // rows: subtract mean, divide by norm
            double (*sg)[ES_SEG] = tid < ES_BANDS ? xsg : ysg;
            const int r = tid % ES_BANDS;
            double mu = 0.0;
            for (int c = 0; c < ES_SEG; ++c) mu += sg[r][c];
            mu /= ES_SEG;
            double nn = 0.0;
            for (int c = 0; c < ES_SEG; ++c) { sg[r][c] -= mu; nn += sg[r][c] * sg[r][c]; }
            nn = sqrt(nn) + eps;
            for (int c = 0; c < ES_SEG; ++c) sg[r][c] /= nn;
        }
        __syncthreads();
        if (tid < 2 * ES_SEG) {                                       // columns likewise
            double (*sg)[ES_SEG] = tid < ES_SEG ? xsg : ysg;
            const int c = tid % ES_SEG;
            double mu = 0.0;
            for (int r = 0; r < ES_BANDS; ++r) mu += sg[r][c];
            mu /= ES_BANDS;
            double nn = 0.0;
            for (int r = 0; r < ES_BANDS; ++r) { sg[r][c] -= mu; nn += sg[r][c] * sg[r][c]; }
            nn = sqrt(nn) + eps;
            for (int r = 0; r < ES_BANDS; ++r) sg[r][c] /= nn;
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int r = 0; r < ES_BANDS; ++r)
                for (int c = 0; c < ES_SEG; ++c) s += xsg[r][c] * ysg[r][c];
            acc += s / ES_SEG;
        }
    }
    __syncthreads();
    if (tid == 0) out[b] = acc / (double)(nf - ES_SEG + 1);
}

}  // namespace l2s
