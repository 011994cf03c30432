// C ABI of lip2speech_b200 (include/l2s_b200.h): context, weight binding, and the orchestration of the
// CUDA kernels for the three reference modules on the inference hot path.
#include "../../include/l2s_b200.h"

#include "audio.cuh"
#include "context.h"
#include "decode.cuh"
#include "decode3.cuh"
#include "gemm.cuh"
#include "lstm.cuh"
#include "misc.cuh"
#include "pack.h"
#include "pw_mma.cuh"
#include "tc_gemm.cuh"
#include "train_model.cuh"
#include "train_step.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include "video.cuh"

using namespace l2s;

struct l2s_ctx {
    Context c;
    std::map<std::string, tr::Param> train_params;      // caller-owned parameter / gradient memory (l2s_train_bind)
    uint64_t train_bind_gen = 1;                        // bumped whenever a binding changes address or size: captured graphs are stale
    tr::DecoderTrain dec_train;
    tr::VideoTrain video_train;
};

static std::string g_create_err;

#define API_BEGIN try {
#define API_END(ctxp)                                                         \
    }                                                                         \
    catch (const L2sError& e) { (ctxp)->c.err = e.what(); return e.code; }    \
    catch (const std::exception& e) { (ctxp)->c.err = e.what(); return L2S_ERR_INVALID; } \
    return L2S_OK;

static inline void check_launch(Context& c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw L2sError(L2S_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    c.launches++;
}

static inline int ew_grid(size_t total, int threads = 256) {
    size_t g = (total + threads - 1) / threads;
    return (int)std::min<size_t>(std::max<size_t>(g, 1), 148 * 16);
}

static void run_gemm(Context& c, GemmParams p, cudaStream_t s, const char* what) {
    cudaError_t e = launch_gemm(p, s);
    if (e != cudaSuccess) throw L2sError(L2S_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    c.launches++;
}

static void run_tc(Context& c, const TcOperands& o, const TcParams& p, cudaStream_t s, const char* what, bool gather = false);
static inline TcParams tc_defaults();

// Few rows, long K (the pre-loop's site / content / E_C layers at B=32: 32-928 rows): the tcgen05 GEMM would run on 2-32
// CTAs, each streaming megabytes of weights alone.  Split K over the whole chip with the streaming mma.sync kernel
// (pw_mma.cuh) and add the partial sums in a fixed order.  Returns false when the shape does not qualify.
struct SplitKExtra { int L = 1; const float* addrow = nullptr; const float* addpos = nullptr; int ldpos = 0; const float* resid = nullptr; int ldr = 0; };
static bool splitk_small_m(Context& c, const float* A, int lda, const std::string& wname, const float* b, float* C, int ldc, int M, int N, int K,
                           int act, const float* act_w, cudaStream_t s, const char* what, const SplitKExtra& ex = SplitKExtra{}) {
    if (!c.use_tc || !c.use_pw || M > 1024 || (size_t)N * K < 32768 || (lda % 4) || (K % 4) || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    const int bn = (N <= 32) ? 32 : (N <= 64 ? 64 : 128);
    if (ceil_div(M, TC_BM) * ceil_div(N, bn) > 48) return false;             // enough tiles for the tcgen05 kernel
    // the split depends on N and K only: a row's result must not depend on how many other rows (clips) share the batch
    const int nslices = ceil_div(N, PW_NSLICE);
    const int want = std::max(1, 2 * c.num_sms / nslices);
    int ksplit = round_up(ceil_div(K, want), 16);
    ksplit = std::min(std::max(ksplit, 64), 224);
    const int nsplits = ceil_div(K, ksplit);
    PwParams p{};
    p.A = A; p.lda = lda; p.M = M; p.N = N; p.Kc = K;
    p.Whi = c.dev(wname + ".hi"); p.Wlo = c.dev(wname + ".lo"); p.kcp = (int)c.meta.at(wname + ".kcp");
    p.cstride = 1; p.ksplit = ksplit;
    p.partial = c.fbuf("ws.pw.partial", (size_t)nsplits * M * N);
    const char* err = launch_pw_mma(p, c.num_sms, s);
    if (err) throw L2sError(L2S_ERR_CUDA, std::string(what) + " (split-K mma): " + err);
    c.launches++;
    pw_reduce_kernel<<<ew_grid((size_t)M * N), 256, 0, s>>>(p.partial, nsplits, M, N, b, act, act_w, C, ldc, ex.L, ex.addrow, ex.addpos, ex.ldpos, ex.resid, ex.ldr);
    check_launch(c, what);
    return true;
}

// plain linear: C[M,N] = act(A[M,K] W[N,K]^T + b).  `wname` selects the packed weight: wname.hi/.lo (tcgen05 3xTF32
// path, needs 16-byte aligned rows for TMA) or wname.w (exact-fp32 SIMT path).
static void linear(Context& c, const float* A, int lda, const std::string& wname, const float* b, float* C, int ldc, int M, int N, int K,
                   int act, const float* act_w, cudaStream_t s, const char* what) {
    if (splitk_small_m(c, A, lda, wname, b, C, ldc, M, N, K, act, act_w, s, what)) return;
    if (c.use_tc && (lda % 4) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
        const int kcp = (int)c.meta.at(wname + ".kcp");
        TcOperands o{A, K, M, lda, c.dev(wname + ".hi"), c.dev(wname + ".lo"), kcp};
        TcParams p = tc_defaults();
        p.M = M; p.N = N; p.Kc = K; p.Kcp = kcp; p.Lp_in = M; p.L = M; p.Lp_out = M;
        p.C = C; p.ldc = ldc; p.bias = b; p.act = act; p.act_w = act_w;
        run_tc(c, o, p, s, what);
        return;
    }
    GemmParams p = gemm_defaults();
    p.A = A; p.lda = lda; p.W = c.dev(wname + ".w"); p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.Kc = K; p.bias = b; p.act = act; p.act_w = act_w;
    run_gemm(c, p, s, what);
}

// Dense (1-tap) GEMM described SIMT-style (GemmParams with W unset): tcgen05 path when TMA alignment allows, else SIMT.
static bool pw_eligible(const Context& c, const GemmParams& g) {
    return c.use_tc && c.use_pw && g.taps == 1 && g.stride == 1 && !g.stem && g.Kc <= 240 && g.M >= c.pw_min_rows && (g.act == ACT_NONE || g.act == ACT_RELU) &&
           !g.addrow && !g.addpos && !g.resid && !g.transposed && (g.lda % 4) == 0 && (g.Kc % 4) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
}

// pass_x / pass_ld: fuse the pass-through half of a stride-1 block into the store (only valid when pw_eligible(g))
static void gemm_auto(Context& c, GemmParams g, const std::string& wname, cudaStream_t s, const char* what,
                      const float* pass_x = nullptr, int pass_ld = 0) {
    if (g.L_out == 0) { g.L_out = g.M; g.L_in = g.M; }
    // tall-skinny pointwise convolutions of the trunk (K, N <= 240, >= 16 K rows: stages 2 and 3): streaming mma.sync kernel
    // (measured 50 vs 70 us at 12x12, 37 vs 40 us at 6x6; at 3x3 — 8 K rows — the tcgen05 GEMM is faster, 31 vs 40 us)
    if (pw_eligible(c, g)) {
        PwParams p{};
        p.pass_x = pass_x; p.pass_ld = pass_ld;
        p.A = g.A; p.lda = g.lda; p.M = g.M; p.N = g.N; p.Kc = g.Kc;
        p.Whi = c.dev(wname + ".hi"); p.Wlo = c.dev(wname + ".lo"); p.kcp = (int)c.meta.at(wname + ".kcp");
        p.bias = g.bias; p.relu = g.act == ACT_RELU; p.C = g.C; p.ldc = g.ldc;
        p.cstride = g.cstride; p.coff = g.coff; p.chalf = g.chalf; p.chp = g.chp;
        const char* err = launch_pw_mma(p, c.num_sms, s);
        if (err) throw L2sError(L2S_ERR_CUDA, std::string(what) + " (pointwise mma): " + err);
        c.launches++;
        return;
    }
    if (pass_x) throw L2sError(L2S_ERR_INVALID, std::string(what) + ": internal: fused pass-through requested on a non-streaming GEMM");
    // few rows, long K, plain channel order (encoder_proj, the K / V bottlenecks at B=32: 8 x 4 tiles would occupy 32 SMs):
    // split K over the chip; the reduction kernel applies the row / position / residual terms of the epilogue
    if (g.taps == 1 && g.stride == 1 && !g.stem && !g.transposed && g.cstride == 1 && g.coff == 0 && g.chalf == 0 ) {
        SplitKExtra ex; ex.L = g.L_out; ex.addrow = g.addrow; ex.addpos = g.addpos; ex.ldpos = g.ldpos; ex.resid = g.resid; ex.ldr = g.ldr;
        if (splitk_small_m(c, g.A, g.lda, wname, g.bias, g.C, g.ldc, g.M, g.N, g.Kc, g.act, g.act_w, s, what, ex)) return;
    }
    if (c.use_tc && g.taps == 1 && g.stride == 1 && !g.stem && (g.lda % 4) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0) {
        const int kcp = (int)c.meta.at(wname + ".kcp");
        TcOperands o{g.A, g.Kc, g.M, g.lda, c.dev(wname + ".hi"), c.dev(wname + ".lo"), kcp};
        TcParams p = tc_defaults();
        p.M = g.M; p.N = g.N; p.Kc = g.Kc; p.Kcp = kcp; p.Lp_in = g.L_out; p.L = g.L_out; p.Lp_out = g.L_out;
        p.C = g.C; p.ldc = g.ldc; p.bias = g.bias; p.act = g.act; p.act_w = g.act_w; p.addrow = g.addrow;
        p.addpos = g.addpos; p.ldpos = g.ldpos; p.resid = g.resid; p.ldr = g.ldr;
        p.cstride = g.cstride; p.coff = g.coff; p.chalf = g.chalf; p.chp = g.chp; p.transposed = g.transposed;
        run_tc(c, o, p, s, what);
        return;
    }
    g.W = c.dev(wname + ".w");
    run_gemm(c, g, s, what);
}

// Where the clips come from.  kind 0: the caller's [B,3,T,H,W] fp32 NCDHW tensor, already normalised (what
// train_collate_fn_pad hands to the model, datasets/__init__.py:7-46).  kind 1: raw decoded frames uint8 [B,T,H,W,3] RGB (what
// loadframes returns, datasets/lrw/dataset.py:20-24); the dataset's `im.float() / 255.0` + Normalize(mean, std)
// (datasets/lrw/dataset.py:82-86) is applied while the space-to-depth rows are written, so a quarter of the bytes cross PCIe
// and HBM and the fp32 NCDHW tensor never exists.
struct VideoSrc {
    const void* p = nullptr;
    int kind = 0;
    float mean[3] = {0.f, 0.f, 0.f}, stdv[3] = {1.f, 1.f, 1.f};
};

// The 12 space-to-depth values of output position (b, t, hq, wq): index = (ph*2 + pw)*3 + ci.
__device__ __forceinline__ void s2d_fetch(const VideoSrc& v, int b, int t, int hq, int wq, int T, int H, int W, float (&vals)[12]) {
    if (v.kind == 0) {
        const float* base = static_cast<const float*>(v.p);
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float* src = base + ((((size_t)b * 3 + ci) * T + t) * H + 2 * hq) * W + 2 * wq;
            const float2 top = *reinterpret_cast<const float2*>(src);
            const float2 bot = *reinterpret_cast<const float2*>(src + W);
            vals[0 + ci] = top.x; vals[3 + ci] = top.y; vals[6 + ci] = bot.x; vals[9 + ci] = bot.y;
        }
    } else {
        // two pixels x RGB = 6 contiguous bytes per row (offset 6*wq: 2-byte aligned)
        const unsigned char* base = static_cast<const unsigned char*>(v.p) + ((((size_t)b * T + t) * H + 2 * hq) * W + 2 * wq) * 3;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const unsigned short* row = reinterpret_cast<const unsigned short*>(base + (size_t)r * W * 3);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const unsigned short u = __ldg(row + j);
                const int i0 = 6 * r + 2 * j, i1 = i0 + 1;
                // exactly the dataset's arithmetic: (u / 255.0 - mean) / std in fp32, IEEE division
                vals[i0] = __fdiv_rn(__fdiv_rn((float)(u & 0xff), 255.0f) - v.mean[i0 % 3], v.stdv[i0 % 3]);
                vals[i1] = __fdiv_rn(__fdiv_rn((float)(u >> 8), 255.0f) - v.mean[i1 % 3], v.stdv[i1 % 3]);
            }
        }
    }
}

// clips -> space-to-depth, zero-padded rows [B][T+4][H/2+3][W/2+3][12]; channel = (ph*2+pw)*3 + ci.
__global__ void s2d_pad_kernel(const VideoSrc v, float* __restrict__ xs, float* __restrict__ xl, int B, int T, int H, int W, int Tp, int Hpp, int Wpp) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * T * Ho * Wo;           // one thread per space-to-depth position: 12 contiguous outputs
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int wq = i % Wo; size_t r = i / Wo;
        int hq = r % Ho; r /= Ho;
        int t = r % T; int b = r / T;
        float vals[12];
        s2d_fetch(v, b, t, hq, wq, T, H, W, vals);
        const size_t o = ((((size_t)b * Tp + t + 2) * Hpp + hq + 2) * Wpp + wq + 2) * 12;
        float hi[12], lo[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {                     // 3xTF32 operand split done once here: hi = top 19 bits, lo = exact remainder
            hi[j] = __uint_as_float(__float_as_uint(vals[j]) & 0xFFFFE000u);
            lo[j] = vals[j] - hi[j];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            *reinterpret_cast<float4*>(xs + o + 4 * q) = make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            *reinterpret_cast<float4*>(xl + o + 4 * q) = make_float4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
    }
}

// bf16 flavour: rows of 16 bf16 per position (12 channels + 4 zeros), so a tap window (4 positions) is 128 contiguous bytes.
__global__ void s2d_pad_bf16_kernel(const VideoSrc v, uint4* __restrict__ xs, int B, int T, int H, int W, int Tp, int Hpp, int Wpp) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * T * Ho * Wo;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int wq = i % Wo; size_t r = i / Wo;
        int hq = r % Ho; r /= Ho;
        int t = r % T; int b = r / T;
        float vals[12];
        s2d_fetch(v, b, t, hq, wq, T, H, W, vals);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 6; ++j)
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(vals[2 * j + 1]), "f"(vals[2 * j]));
        pk[6] = pk[7] = 0u;
        const size_t o = ((((size_t)b * Tp + t + 2) * Hpp + hq + 2) * Wpp + wq + 2) * 2;      // two 16-byte pieces per position
        xs[o] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        xs[o + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

// ------------------------------------------------------------------------------------------------
// video frontend
// ------------------------------------------------------------------------------------------------
static void video_forward_chunk(Context& c, const VideoSrc& video, int B, int T, int H, int W, float* out_feat, int precision, cudaStream_t s) {
    if (B <= 0 || T <= 0 || (H & 3) || (W & 3)) throw L2sError(L2S_ERR_INVALID, "video_fwd: bad shape");
    if (precision != L2S_PRECISION_FP32 && precision != L2S_PRECISION_BF16) throw L2sError(L2S_ERR_INVALID, "video_fwd: unknown precision");
    if (precision == L2S_PRECISION_BF16 && !c.use_tc) throw L2sError(L2S_ERR_INVALID, "video_fwd: the bf16 stem needs the tcgen05 path (L2S_TC=0 is set)");
    const int N = B * T;
    const int Ho = H / 2, Wo = W / 2;                 // Conv3d stride (1,2,2), pad 3, k 7
    const int Hp = (Ho - 1) / 2 + 1, Wp = (Wo - 1) / 2 + 1;   // MaxPool 3x3 s2 p1
    float* stem = c.fbuf("ws.v.stem", (size_t)N * Ho * Wo * 24);
    if (precision == L2S_PRECISION_BF16) {
        // BASELINE config 2 ("bf16 frontend"): the Conv3d stem multiplies bf16 operands (fp32 accumulation in TMEM); the
        // trunk stays on the 3xTF32 path (a bf16 trunk moves the features by 2e-3, outside the parity bound)
        const int Tp = T + 4, Hpp = Ho + 3, Wpp = Wo + 3;
        const size_t rows = (size_t)B * Tp * Hpp * Wpp;
        const size_t slack_rows = (size_t)2 * Hpp * Wpp + 2 * Wpp + 2 + TC_BM + 8;
        const size_t bytes = (rows + 2 * slack_rows) * 32;
        char* xs0 = static_cast<char*>(c.buf("ws.v.s2d16", bytes));
        const int64_t sig = ((int64_t)B << 40) ^ ((int64_t)T << 24) ^ ((int64_t)H << 12) ^ W;
        if (c.meta["ws.v.s2d16.layout"] != sig) {
            L2S_CUDA(cudaMemsetAsync(xs0, 0, bytes, s));
            c.meta["ws.v.s2d16.layout"] = sig;
        }
        char* xs = xs0 + slack_rows * 32;
        s2d_pad_bf16_kernel<<<ew_grid((size_t)B * T * Ho * Wo), 256, 0, s>>>(video, reinterpret_cast<uint4*>(xs), B, T, H, W, Tp, Hpp, Wpp);
        check_launch(c, "space-to-depth (bf16)");
        TcParams p = tc_defaults();
        p.M = (int)rows; p.N = 24; p.Kc = 64; p.Kcp = 64; p.C = stem; p.ldc = 24;
        p.bias = c.dev("v.stem.b"); p.act = ACT_PRELU; p.act_w = c.dev("v.stem.prelu");
        for (int kt = 0; kt < 5; ++kt)
            for (int jh = 0; jh < 4; ++jh) p.tap_shift[kt * 4 + jh] = (kt - 2) * Hpp * Wpp + (jh - 2) * Wpp - 2;
        p.stem = 1; p.sT = T; p.sTp = Tp; p.sHp = Hpp; p.sWp = Wpp; p.sHo = Ho; p.sWo = Wo;
        p.Lp_in = 1; p.L = 1; p.Lp_out = 1;
        const char* err = launch_tc_stem_bf16(xs, c.dev("v.stem.tc16"), 20, p, s);
        if (err) throw L2sError(L2S_ERR_CUDA, std::string("stem conv3d (bf16 tcgen05): ") + err);
        c.launches++;
    } else if (c.use_tc) {
        // space-to-depth + zero padding, then a 20-tap implicit GEMM (K = 4 w-taps x 12 channels per tap) on tcgen05
        const int Tp = T + 4, Hpp = Ho + 3, Wpp = Wo + 3;
        const size_t rows = (size_t)B * Tp * Hpp * Wpp;
        // zero slack before/after so the tap-shifted gathers of the first/last tiles never leave the allocation
        const size_t slack_rows = (size_t)2 * Hpp * Wpp + 2 * Wpp + 2 + TC_BM + 8;
        const size_t plane = (rows + 2 * slack_rows) * 12;
        float* xs0 = c.fbuf("ws.v.s2d", 2 * plane);          // [hi | lo]
        float* xs = xs0 + slack_rows * 12;
        float* xl = xs + plane;
        const int64_t sig = ((int64_t)B << 40) ^ ((int64_t)T << 24) ^ ((int64_t)H << 12) ^ W;
        if (c.meta["ws.v.s2d.layout"] != sig) {
            L2S_CUDA(cudaMemsetAsync(xs0, 0, 2 * plane * sizeof(float), s));
            c.meta["ws.v.s2d.layout"] = sig;
        }
        s2d_pad_kernel<<<ew_grid((size_t)B * T * Ho * Wo), 256, 0, s>>>(video, xs, xl, B, T, H, W, Tp, Hpp, Wpp);
        check_launch(c, "space-to-depth");
        const int kcp = (int)c.meta.at("v.stem.tc.kcp");
        TcOperands o{xs, 48, (int)rows, 12, c.dev("v.stem.tc.hi"), c.dev("v.stem.tc.lo"), 20 * kcp};
        TcParams p = tc_defaults();
        p.M = (int)rows; p.N = 24; p.Kc = 48; p.Kcp = kcp; p.taps = 20; p.C = stem; p.ldc = 24;
        p.bias = c.dev("v.stem.b"); p.act = ACT_PRELU; p.act_w = c.dev("v.stem.prelu");
        p.use_shift_table = 1;
        for (int kt = 0; kt < 5; ++kt)
            for (int jh = 0; jh < 4; ++jh) p.tap_shift[kt * 4 + jh] = (kt - 2) * Hpp * Wpp + (jh - 2) * Wpp - 2;
        p.ga_unchecked = 1; p.ga_Alo = xl;
        p.stem = 1; p.sT = T; p.sTp = Tp; p.sHp = Hpp; p.sWp = Wpp; p.sHo = Ho; p.sWo = Wo;
        p.Lp_in = 1; p.L = 1; p.Lp_out = 1;
        run_tc(c, o, p, s, "stem conv3d", /*gather=*/true);
    } else {
        GemmParams p = gemm_defaults();
        if (video.kind != 0) throw L2sError(L2S_ERR_INVALID, "video_fwd: uint8 frames need the tcgen05 path");
        p.stem = 1; p.A = static_cast<const float*>(video.p); p.W = c.dev("v.stem.w"); p.C = stem; p.ldc = 24;
        p.M = N * Ho * Wo; p.N = 24; p.Kc = 735; p.taps = 1; p.L_out = p.M; p.L_in = p.M;
        p.bias = c.dev("v.stem.b"); p.act = ACT_PRELU; p.act_w = c.dev("v.stem.prelu");
        p.T = T; p.H = H; p.Wd = W; p.Ho = Ho; p.Wo = Wo;
        run_gemm(c, p, s, "stem conv3d");
    }
    // activation ping-pong and branch temporaries, sized for the largest stage
    int h = Hp, w = Wp;
    const size_t rows0 = (size_t)N * h * w;
    float* xa = c.fbuf("ws.v.xa", rows0 * 32);     // >= N*24*24*24 and >= N*12*12*120
    float* xb = c.fbuf("ws.v.xb", rows0 * 32);
    float* t1 = c.fbuf("ws.v.t1", rows0 * 64);     // dw outputs (quarter resolution) / pw1 outputs (full res, hp<=60 at stage 2)
    float* t2 = c.fbuf("ws.v.t2", rows0 * 64);
    {
        size_t total = rows0 * 6;
        maxpool3x3s2_kernel<<<ew_grid(total), 256, 0, s>>>(stem, xa, N, Ho, Wo, 24, Hp, Wp);
        check_launch(c, "maxpool");
    }
    float* x = xa; float* y = xb;
    const int nblk = (int)c.meta.at("v.nblocks");
    for (int i = 0; i < nblk; ++i) {
        const std::string n = "v.b" + std::to_string(i) + ".";
        const int down = (int)c.meta.at(n + "down"), cin_phys = (int)c.meta.at(n + "cin_phys");
        const int half = (int)c.meta.at(n + "half"), hp = (int)c.meta.at(n + "hp");
        const int cph = 2 * hp;
        if (down) {
            const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
            const size_t rin = (size_t)N * h * w, rout = (size_t)N * ho * wo;
            // branch1: dw s2 (all input channels) -> 1x1 -> logical channel 2j
            dwconv3x3_kernel<<<ew_grid(rout * (cin_phys / 4)), 256, 0, s>>>(x, cin_phys, 0, t1, cin_phys, 0, c.dev(n + "b1dw.w"), c.dev(n + "b1dw.b"),
                                                                          N, h, w, cin_phys, 2, ho, wo);
            check_launch(c, "b1 dw");
            GemmParams p = gemm_defaults();
            p.A = t1; p.lda = cin_phys; p.bias = c.dev(n + "b1pw.b"); p.act = ACT_RELU;
            p.C = y; p.ldc = cph; p.M = (int)rout; p.N = half; p.Kc = cin_phys; p.cstride = 2; p.coff = 0; p.chalf = half; p.chp = hp;
            gemm_auto(c, p, n + "b1pw", s, "b1 pw");
            // branch2: 1x1 (full res) -> dw s2 -> 1x1 -> logical channel 2j+1
            p = gemm_defaults();
            p.A = x; p.lda = cin_phys; p.bias = c.dev(n + "b2pw1.b"); p.act = ACT_RELU;
            p.C = t2; p.ldc = hp; p.M = (int)rin; p.N = half; p.Kc = cin_phys;
            gemm_auto(c, p, n + "b2pw1", s, "b2 pw1");
            dwconv3x3_kernel<<<ew_grid(rout * (hp / 4)), 256, 0, s>>>(t2, hp, 0, t1, hp, 0, c.dev(n + "b2dw.w"), c.dev(n + "b2dw.b"), N, h, w, hp, 2, ho, wo);
            check_launch(c, "b2 dw");
            p = gemm_defaults();
            p.A = t1; p.lda = hp; p.bias = c.dev(n + "b2pw2.b"); p.act = ACT_RELU;
            p.C = y; p.ldc = cph; p.M = (int)rout; p.N = half; p.Kc = hp; p.cstride = 2; p.coff = 1; p.chalf = half; p.chp = hp;
            gemm_auto(c, p, n + "b2pw2", s, "b2 pw2");
            h = ho; w = wo;
        } else {
            const size_t rows = (size_t)N * h * w;
            GemmParams p = gemm_defaults();
            p.A = x + hp; p.lda = cph; p.bias = c.dev(n + "b2pw1.b"); p.act = ACT_RELU;
            p.C = t2; p.ldc = hp; p.M = (int)rows; p.N = half; p.Kc = hp;
            gemm_auto(c, p, n + "b2pw1", s, "b2 pw1");
            if (w % 4 == 0)
                dwconv3x3_s1_strip_kernel<4><<<ew_grid(rows / 4 * (hp / 4)), 256, 0, s>>>(t2, hp, 0, t1, hp, 0, c.dev(n + "b2dw.w"), c.dev(n + "b2dw.b"), N, h, w, hp);
            else if (w % 3 == 0)
                dwconv3x3_s1_strip_kernel<3><<<ew_grid(rows / 3 * (hp / 4)), 256, 0, s>>>(t2, hp, 0, t1, hp, 0, c.dev(n + "b2dw.w"), c.dev(n + "b2dw.b"), N, h, w, hp);
            else
                dwconv3x3_kernel<<<ew_grid(rows * (hp / 4)), 256, 0, s>>>(t2, hp, 0, t1, hp, 0, c.dev(n + "b2dw.w"), c.dev(n + "b2dw.b"), N, h, w, hp, 1, h, w);
            check_launch(c, "b2 dw");
            p = gemm_defaults();
            p.A = t1; p.lda = hp; p.bias = c.dev(n + "b2pw2.b"); p.act = ACT_RELU;
            p.C = y; p.ldc = cph; p.M = (int)rows; p.N = half; p.Kc = hp; p.cstride = 2; p.coff = 1; p.chalf = half; p.chp = hp;
            if (pw_eligible(c, p)) {
                gemm_auto(c, p, n + "b2pw2", s, "b2 pw2 + pass-through", x, cph);      // x1 -> even channels inside the GEMM's store
            } else {
                shuffle_passthrough_kernel<<<ew_grid(rows * half), 256, 0, s>>>(x, y, rows, cph, half, hp);
                check_launch(c, "passthrough");
                gemm_auto(c, p, n + "b2pw2", s, "b2 pw2");
            }
        }
        std::swap(x, y);
    }
    if (h != 3 || w != 3) throw L2sError(L2S_ERR_INVALID, "video_fwd: trunk output must be 3x3 (H,W in {88,96}) for AvgPool2d(3)");
    const int Kl = (int)c.meta.at("v.last.k"), Nl = (int)c.meta.at("v.last.n");
    float* last = c.fbuf("ws.v.last", (size_t)N * 9 * Nl);
    linear(c, x, Kl, "v.last", c.dev("v.last.b"), last, Nl, N * 9, Nl, Kl, ACT_RELU, nullptr, s, "conv_last");
    avgpool_l2norm_kernel<<<N, 256, Nl * sizeof(float), s>>>(last, out_feat, 9, Nl);
    check_launch(c, "avgpool_l2norm");
}

// Whole-batch entry: clips are independent in the frontend (eval BatchNorm), so large batches run in chunks of 32 clips —
// workspaces stay bounded (the unfused stem output is 16.6 MB per T=75 clip) and row counts stay far from 2^31.
constexpr int VIDEO_CHUNK = 32;
static void video_forward(Context& c, const VideoSrc& video, int B, int T, int H, int W, float* out_feat, int precision, cudaStream_t s) {
    if (B <= 0) throw L2sError(L2S_ERR_INVALID, "video_fwd: bad shape");
    const size_t clip_elems = (size_t)3 * T * H * W;
    for (int b0 = 0; b0 < B; b0 += VIDEO_CHUNK) {
        VideoSrc v = video;
        v.p = video.kind == 0 ? static_cast<const void*>(static_cast<const float*>(video.p) + (size_t)b0 * clip_elems)
                              : static_cast<const void*>(static_cast<const unsigned char*>(video.p) + (size_t)b0 * clip_elems);
        video_forward_chunk(c, v, std::min(VIDEO_CHUNK, B - b0), T, H, W, out_feat + (size_t)b0 * T * 768, precision, s);
    }
}
static inline VideoSrc video_f32(const float* p) { VideoSrc v; v.p = p; v.kind = 0; return v; }
static inline VideoSrc video_u8(const unsigned char* p, const float* mean_std) {
    VideoSrc v; v.p = p; v.kind = 1;
    for (int i = 0; i < 3; ++i) { v.mean[i] = mean_std[i]; v.stdv[i] = mean_std[3 + i]; }
    return v;
}

// ------------------------------------------------------------------------------------------------
// persistent LSTM launch helper
// ------------------------------------------------------------------------------------------------
// hist[time slot][plane][H][Bpad] is zero-filled ("not written") and slot 0 / the cell state get the initial state:
// h0 rows [B][lds] for every plane (the encoder Bi-LSTM's site embedding, decoder.py:325) or zeros (h0 == nullptr)
__global__ void lstm_init_kernel(const float* __restrict__ h0, int lds, float* __restrict__ hist0, float* __restrict__ cbuf, int planes, int H, int B, int Bpad) {
    const size_t total = (size_t)planes * H * Bpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i % Bpad; const int f = (i / Bpad) % H;
        const float v = (h0 && b < B) ? h0[(size_t)b * lds + f] : 0.f;
        lstm_st(hist0 + i, v);
        cbuf[i] = v;
    }
}

static float* lstm_history(Context& c, const std::string& name, const float* h0, int lds, float* cbuf, int T, int planes, int H, int B, int Bpad, cudaStream_t s) {
    const size_t slot = (size_t)planes * H * Bpad;
    float* hist = c.fbuf(name, (size_t)(T + 1) * slot);
    L2S_CUDA(cudaMemsetAsync(hist, 0, (size_t)(T + 1) * slot * sizeof(float), s));
    lstm_init_kernel<<<ew_grid(slot), 256, 0, s>>>(h0, lds, hist, cbuf, planes, H, B, Bpad);
    check_launch(c, "lstm initial state");
    return hist;
}

static void launch_lstm(Context& c, LstmParams lp, cudaStream_t s) {
    lp.abort_word = static_cast<unsigned*>(c.buf("ws.d.abort", 256));
    const size_t smem = lstm_smem_bytes(lp.H);
    if ((int64_t)smem > c.meta["attr.lstm.smem"]) {          // raised when needed, not set on every launch
        L2S_CUDA(cudaFuncSetAttribute(lstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c.meta["attr.lstm.smem"] = (int64_t)smem;
    }
    const int grid = c.num_sms;
    void* args[] = {&lp};
    L2S_CUDA(cudaLaunchCooperativeKernel((void*)lstm_persistent_kernel, dim3(grid), dim3(MV_THREADS), args, smem, s));
    c.launches++;
}

// ------------------------------------------------------------------------------------------------
// speaker encoder
// ------------------------------------------------------------------------------------------------
static void speaker_forward(Context& c, const float* wav, int B, int S, float* emb, int normalize, cudaStream_t s) {
    if (B <= 0 || S < 401) throw L2sError(L2S_ERR_INVALID, "speaker_fwd: bad shape");
    const int F = 1 + S / 160, H = 256, L = 3;
    const int Bpad = round_up(B, 32);
    float* mel = c.fbuf("ws.s.mel", (size_t)B * F * 40);
    if (c.use_tc) {
        // STFT as one [B*F, 400] x [400, 402] GEMM on tcgen05 (3xTF32) + power / filterbank kernel
        float* frames = c.fbuf("ws.s.frames", (size_t)B * F * 400);
        float* specb = c.fbuf("ws.s.spec", (size_t)B * F * 404);
        stft_frames_kernel<<<ew_grid((size_t)B * F * 400), 256, 0, s>>>(wav, c.dev("s.window"), frames, B, S, F);
        check_launch(c, "stft frames");
        linear(c, frames, 400, "s.dft", nullptr, specb, 404, B * F, 402, 400, ACT_NONE, nullptr, s, "stft dft");
        power_mel_kernel<<<ceil_div(B * F, 8), 256, 0, s>>>(specb, 404, c.dev("s.fb"), mel, B * F);
        check_launch(c, "power + mel filterbank");
    } else {
        melspec_kernel<<<B * F, 256, 0, s>>>(wav, c.dev("s.window"), c.dev("s.fb"), mel, S, F);
        check_launch(c, "melspec");
    }
    float* xproj = c.fbuf("ws.s.xproj", (size_t)B * F * 4 * H);
    linear(c, mel, 40, "s.wih0", c.dev("s.b0"), xproj, 4 * H, B * F, 4 * H, 40, ACT_NONE, nullptr, s, "speaker xproj");
    const size_t plane = (size_t)H * Bpad;
    float* cbuf = c.fbuf("ws.s.c", L * plane);
    float* hist = lstm_history(c, "ws.s.hist", nullptr, 0, cbuf, F, L, H, B, Bpad, s);      // zero initial state (audio.py:135)
    LstmParams lp{};
    lp.xproj = xproj; lp.ldx = 4 * H; lp.wpk = c.dev("s.lstm.w");
    lp.blocks = reinterpret_cast<const LstmBlock*>(c.dev("s.lstm.blocks"));
    lp.hist = hist; lp.cbuf = cbuf; lp.out = nullptr; lp.ldo = 0;
    lp.T = F; lp.B = B; lp.Bpad = Bpad; lp.H = H; lp.L = L; lp.dirs = 1;
    launch_lstm(c, lp, s);
    const float* hfinal = hist + (size_t)F * L * plane + (size_t)(L - 1) * plane;          // last layer after the last frame
    speaker_head_kernel<<<B, 256, 0, s>>>(hfinal, Bpad, c.dev("s.lin.w"), c.dev("s.lin.b"), emb, normalize);
    check_launch(c, "speaker head");
}

// ------------------------------------------------------------------------------------------------
// postnet (time-major rows [B*L][C])
// ------------------------------------------------------------------------------------------------
static void run_tc(Context& c, const TcOperands& o, const TcParams& p, cudaStream_t s, const char* what, bool gather) {
    const char* err = launch_tc_gemm(o, p, s, gather);
    if (err) throw L2sError(L2S_ERR_CUDA, std::string(what) + " (tcgen05 gemm): " + err);
    c.launches++;
}

static inline TcParams tc_defaults() {
    TcParams p{};
    p.taps = 1; p.cstride = 1;
    return p;
}

// dst[b][P + t][:] = src[b][t][:]   (compact rows -> zero-padded per-sequence layout)
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int L, int Lp, int P, int C4) {
    const size_t total = (size_t)B * L * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = i % C4; size_t r = i / C4;
        int t = r % L; int b = r / L;
        reinterpret_cast<float4*>(dst)[((size_t)b * Lp + P + t) * C4 + c4] = reinterpret_cast<const float4*>(src)[i];
    }
}

// dst[b][t][:] = src[b][t][:] for t < Lk (drops the tail rows of every sequence so that rows become contiguous k-groups)
__global__ void trim_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int T, int Lk, int C4) {
    const size_t total = (size_t)B * Lk * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c4 = i % C4; size_t r = i / C4;
        int t = r % Lk; int b = r / Lk;
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[((size_t)b * T + t) * C4 + c4];
    }
}

// (Re)allocate a padded-layout buffer; its pad rows must be zero, so it is cleared whenever the layout changes.
static float* padded_buf(Context& c, const std::string& name, int B, int Lp, int C, cudaStream_t s) {
    float* p = c.fbuf(name, (size_t)B * Lp * C);
    const int64_t sig = ((int64_t)B << 40) ^ ((int64_t)Lp << 20) ^ C;
    auto it = c.meta.find(name + ".layout");
    if (it == c.meta.end() || it->second != sig) {
        L2S_CUDA(cudaMemsetAsync(p, 0, (size_t)B * Lp * C * sizeof(float), s));
        c.meta[name + ".layout"] = sig;
    }
    return p;
}

// Postnet on the tensor cores: 5 x (k=5 Conv1d as 5-tap implicit GEMM, BN folded, PSine / residual epilogue).
static void postnet_rows_tc(Context& c, const float* x_rows /*[B*L][80]*/, int B, int L, float* out_bcl, bool add_residual, cudaStream_t s) {
    const int P = 2, Lp = L + 2 * P, Mp = B * Lp;
    float* xin = padded_buf(c, "ws.p.xin", B, Lp, 80, s);
    float* bufs[2] = {padded_buf(c, "ws.p.ta", B, Lp, 512, s), padded_buf(c, "ws.p.tb", B, Lp, 512, s)};
    pad_rows_kernel<<<ew_grid((size_t)B * L * 20), 256, 0, s>>>(x_rows, xin, B, L, Lp, P, 20);
    check_launch(c, "pad rows");
    const float* in = xin; int cin = 80;
    for (int i = 0; i < 5; ++i) {
        const std::string n = "d.post" + std::to_string(i);
        const int kcp = (int)c.meta.at(n + ".kcp");
        TcOperands o{in, cin, Mp, cin, c.dev(n + ".hi"), c.dev(n + ".lo"), 5 * kcp};
        TcParams p = tc_defaults();
        p.M = Mp; p.Kc = cin; p.Kcp = kcp; p.taps = 5; p.pad = 2;
        p.Lp_in = Lp; p.P_in = P; p.L = L; p.Lp_out = Lp; p.P_out = P;
        p.bias = c.dev(n + ".b");
        if (i < 4) {
            p.N = 512; p.C = bufs[i & 1]; p.ldc = 512; p.act = ACT_PSINE; p.act_w = c.dev(n + ".psw");
            if (i != 0) { p.resid = in; p.ldr = 512; }
        } else {
            p.N = 80; p.C = out_bcl; p.transposed = 1; p.act = ACT_NONE;
            if (add_residual) { p.resid = xin; p.ldr = 80; }
        }
        run_tc(c, o, p, s, "postnet conv");
        in = bufs[i & 1]; cin = 512;
    }
}

static void postnet_rows(Context& c, const float* x_rows /*[B*L][80]*/, int B, int L, float* out_bcl /*[B][80][L]*/, bool add_residual, cudaStream_t s) {
    if (c.use_tc) { postnet_rows_tc(c, x_rows, B, L, out_bcl, add_residual, s); return; }
    const int M = B * L;
    float* a = c.fbuf("ws.p.a", (size_t)M * 512);
    float* b = c.fbuf("ws.p.b", (size_t)M * 512);
    const float* in = x_rows; int cin = 80;
    float* bufs[2] = {a, b};
    for (int i = 0; i < 5; ++i) {
        const std::string n = "d.post" + std::to_string(i);
        GemmParams p = gemm_defaults();
        p.A = in; p.lda = cin; p.W = c.dev(n + ".w"); p.bias = c.dev(n + ".b");
        p.M = M; p.Kc = cin; p.taps = 5; p.pad = 2; p.stride = 1; p.L_out = L; p.L_in = L;
        if (i < 4) {
            p.N = 512; p.C = bufs[i & 1]; p.ldc = 512; p.act = ACT_PSINE; p.act_w = c.dev(n + ".psw");
            if (i != 0) { p.resid = in; p.ldr = 512; }
        } else {
            p.N = 80; p.C = out_bcl; p.transposed = 1; p.act = ACT_NONE;
            if (add_residual) { p.resid = x_rows; p.ldr = 80; }
        }
        run_gemm(c, p, s, "postnet conv");
        in = bufs[i & 1]; cin = 512;
    }
}

// ------------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------------
struct DecoderForwardIO {          // Decoder.forward flavour (eval mode); null pointer = plain inference
    const float* mels = nullptr;                 // [B][80][M] teacher frames
    const unsigned char* tf_mask = nullptr;      // [M] host bytes
    float* out_mel = nullptr;                    // [B][80][M] decoder outputs before the postnet
    float* out_stop = nullptr;                   // [B][M]
    float* out_attn_logits = nullptr;            // [B][M][T]
    float* out_content_dis = nullptr;            // [B*minT][501]
};

// rows [B][L][C] -> x [B][C][L]
__global__ void rows_to_bcl_kernel(const float* __restrict__ rows, float* __restrict__ x, int B, int C, int L) {
    const size_t total = (size_t)B * C * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int t = i % L; size_t r = i / L;
        int ch = r % C; int b = r / C;
        x[i] = rows[((size_t)b * L + t) * C + ch];
    }
}
// teacher rows: tr[(b*M + i)][:] = (i == 0) ? BOS : mels[b][:][i-1]      (decoder.py:343: cat([BOS, mels]))
__global__ void teacher_rows_kernel(const float* __restrict__ mels, const float* __restrict__ bos, float* __restrict__ tr, int B, int M) {
    const size_t total = (size_t)B * M * 80;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int ch = i % 80; size_t r = i / 80;
        int st = r % M; int b = r / M;
        tr[i] = st == 0 ? bos[ch] : mels[((size_t)b * 80 + ch) * M + st - 1];
    }
}
// p1t[(i*256 + f)*Bpad + b] = rows[(b*M + i)*256 + f]
__global__ void p1_teacher_fm_kernel(const float* __restrict__ rows, float* __restrict__ p1t, int B, int Bpad, int M) {
    const size_t total = (size_t)M * 256 * Bpad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int b = i % Bpad; size_t r = i / Bpad;
        int f = r % 256; int st = r / 256;
        p1t[i] = b < B ? rows[((size_t)b * M + st) * 256 + f] : 0.f;
    }
}

static void decoder_run(Context& c, const float* visual, const float* spk, const float* gumbel, int B, int T, int steps,
                        float* mel_post, int64_t* lengths, float* attn, const DecoderForwardIO& fw, cudaStream_t s) {
    if (B <= 0 || T < 7 || T > 300 || steps <= 0 || steps > 300) throw L2sError(L2S_ERR_INVALID, "decoder_infer: need 7<=T<=300, 1<=steps<=300 (pos_table has 300 rows)");
    if (!gumbel) throw L2sError(L2S_ERR_INVALID, "decoder_infer: gumbel noise tensor is required");
    const int M = B * T, Bpad = round_up(B, 32);
    int minT = T;
    const int cks[4] = {1, 3, 5, 7};
    int Lc[4];
    for (int j = 0; j < 4; ++j) { Lc[j] = (T - cks[j]) / cks[j] + 1; minT = std::min(minT, Lc[j]); }
    const float* pos = c.dev("d.pos");

    // ---- encoder pre-loop (decoder.py:383-394) --------------------------------------------------
    c.span_begin("preloop", s);
    float* resid = c.fbuf("ws.d.resid", (size_t)M * 512);
    linear(c, visual, 1024, "d.resid", c.dev("d.resid.b"), resid, 512, M, 512, 1024, ACT_NONE, nullptr, s, "residual_bottleneck");
    float* encsite = c.fbuf("ws.d.encsite", (size_t)B * 512);
    float* attsite = c.fbuf("ws.d.attsite", (size_t)B * 512);
    linear(c, spk, 256, "d.encsite", c.dev("d.encsite.b"), encsite, 512, B, 512, 256, ACT_PSINE, c.dev("d.encsite.psw"), s, "encoder_site");
    linear(c, spk, 256, "d.attsite", c.dev("d.attsite.b"), attsite, 512, B, 512, 256, ACT_PSINE, c.dev("d.attsite.psw"), s, "attention_site");
    float* xproj = c.fbuf("ws.d.xproj", (size_t)M * 4096);
    linear(c, visual, 1024, "d.ernn.wih", c.dev("d.ernn.b"), xproj, 4096, M, 4096, 1024, ACT_NONE, nullptr, s, "encoder_rnn xproj");
    const size_t plane = (size_t)512 * Bpad;
    float* ec = c.fbuf("ws.d.ec", 2 * plane);
    float* eh = lstm_history(c, "ws.d.ehist", encsite, 512, ec, T, 2, 512, B, Bpad, s);    // h0 = c0 = site embedding, both directions
    float* rnn_out = c.fbuf("ws.d.rnnout", (size_t)M * 1024);
    {
        LstmParams lp{};
        lp.xproj = xproj; lp.ldx = 4096; lp.wpk = c.dev("d.ernn.w");
        lp.blocks = reinterpret_cast<const LstmBlock*>(c.dev("d.ernn.blocks"));
        lp.hist = eh; lp.cbuf = ec; lp.out = rnn_out; lp.ldo = 1024;
        lp.T = T; lp.B = B; lp.Bpad = Bpad; lp.H = 512; lp.L = 1; lp.dirs = 2;
        launch_lstm(c, lp, s);
    }
    const float* hfinal = eh + (size_t)T * 2 * plane;            // slot T: feature-major [h_fwd ; h_bwd] = decoder (h0 ; h1)
    float* ccat = c.fbuf("ws.d.ccat", (size_t)B * 1024);
    fm_to_rows_kernel<<<ew_grid((size_t)1024 * B), 256, 0, s>>>(ec, Bpad, ccat, 1024, 0, 1024, B);
    check_launch(c, "c_n -> rows");
    float* enc_cell = c.fbuf("ws.d.enc_cell", (size_t)B * 512);
    linear(c, ccat, 1024, "d.ec", c.dev("d.ec.b"), enc_cell, 512, B, 512, 1024, ACT_NONE, nullptr, s, "E_C");
    float* enc = c.fbuf("ws.d.enc", (size_t)M * 512);
    {
        GemmParams p = gemm_defaults();
        p.A = rnn_out; p.lda = 1024; p.bias = c.dev("d.encproj.b"); p.C = enc; p.ldc = 512;
        p.M = M; p.N = 512; p.Kc = 1024; p.L_out = T; p.L_in = T; p.addrow = attsite; p.resid = resid; p.ldr = 512;
        gemm_auto(c, p, "d.encproj", s, "encoder_proj");
    }
    // ---- K / V (MultiHopConv + PSine + positions, decoder.py:396-399) -----------------------------
    float* cat = c.fbuf("ws.d.cat", (size_t)M * 5120);      // [M][K half: enc | 4 convs][V half: enc | 4 convs]
    float* Kmem = c.fbuf("ws.d.K", (size_t)M * 512);
    float* Vmem = c.fbuf("ws.d.V", (size_t)M * 512);
    L2S_CUDA(cudaMemcpy2DAsync(cat, 5120 * sizeof(float), enc, 512 * sizeof(float), 512 * sizeof(float), M, cudaMemcpyDeviceToDevice, s));
    L2S_CUDA(cudaMemcpy2DAsync(cat + 2560, 5120 * sizeof(float), enc, 512 * sizeof(float), 512 * sizeof(float), M, cudaMemcpyDeviceToDevice, s));
    const int mks[4] = {1, 3, 7, 11};
    const int PM = 5, Tp = T + 2 * PM;                       // zero-padded per-sequence layout for the tap-shifted TMA boxes
    float* encp = nullptr;
    if (c.use_tc) {
        encp = padded_buf(c, "ws.d.encp", B, Tp, 512, s);
        pad_rows_kernel<<<ew_grid((size_t)M * 128), 256, 0, s>>>(enc, encp, B, T, Tp, PM, 128);
        check_launch(c, "pad enc rows");
    }
    for (int j = 0; j < 4; ++j) {                            // K and V branches of conv j in one launch (N = 1024)
        const std::string wn = "d.KV.c" + std::to_string(j);
        const int coff = 512 * (j + 1);
        if (c.use_tc) {
            const int kcp = (int)c.meta.at(wn + ".kcp");
            TcOperands o{encp, 512, B * Tp, 512, c.dev(wn + ".hi"), c.dev(wn + ".lo"), mks[j] * kcp};
            TcParams p = tc_defaults();
            p.M = B * Tp; p.N = 1024; p.Kc = 512; p.Kcp = kcp; p.taps = mks[j]; p.pad = mks[j] / 2;
            p.Lp_in = Tp; p.P_in = PM; p.L = T; p.Lp_out = T; p.P_out = 0;
            p.C = cat; p.ldc = 5120; p.coff = coff; p.chalf = coff + 512; p.chp = 2560 + coff;
            p.bias = c.dev(wn + ".b"); p.act = ACT_SILU;
            run_tc(c, o, p, s, "multihop conv");
        } else {
            GemmParams p = gemm_defaults();
            p.A = enc; p.lda = 512; p.W = c.dev(wn + ".w"); p.bias = c.dev(wn + ".b");
            p.C = cat; p.ldc = 5120; p.coff = coff; p.chalf = coff + 512; p.chp = 2560 + coff;
            p.M = M; p.N = 1024; p.Kc = 512; p.taps = mks[j]; p.pad = mks[j] / 2;
            p.L_out = T; p.L_in = T; p.act = ACT_SILU;
            run_gemm(c, p, s, "multihop conv");
        }
    }
    for (int kv = 0; kv < 2; ++kv) {
        const std::string n = kv == 0 ? "d.K" : "d.V";
        GemmParams p = gemm_defaults();
        p.A = cat + 2560 * kv; p.lda = 5120; p.bias = c.dev(n + ".bn.b"); p.C = kv == 0 ? Kmem : Vmem; p.ldc = 512;
        p.M = M; p.N = 512; p.Kc = 2560; p.L_out = T; p.L_in = T; p.act = ACT_PSINE; p.act_w = c.dev(n + ".psw"); p.addpos = pos; p.ldpos = 512;
        gemm_auto(c, p, n + ".bn", s, "multihop bottleneck");
    }
    // ---- Content.encode (decoder.py:239-260) -----------------------------------------------------
    const int Mc = B * minT;
    float* ccat2 = c.fbuf("ws.d.ccat2", (size_t)Mc * 2560);
    float* ctmp = c.fbuf("ws.d.ctmp", (size_t)M * 512);
    adaptive_pool_kernel<<<ew_grid((size_t)Mc * 512), 256, 0, s>>>(enc, 512, T, ccat2, 2560, 0, minT, 512, B);
    check_launch(c, "adaptive pool");
    float* ctrim = c.fbuf("ws.d.ctrim", (size_t)M * 512);
    for (int j = 0; j < 4; ++j) {
        const std::string wn = "d.cagg" + std::to_string(j);
        if (c.use_tc) {
            // kernel == stride: output i reads input rows k*i .. k*i+k-1, i.e. one k*512-wide row of the (trimmed) sequence
            const float* src = enc;
            if (cks[j] > 1) {
                trim_rows_kernel<<<ew_grid((size_t)B * Lc[j] * cks[j] * 128), 256, 0, s>>>(enc, ctrim, B, T, Lc[j] * cks[j], 128);
                check_launch(c, "trim rows");
                src = ctrim;
            }
            linear(c, src, cks[j] * 512, wn, c.dev(wn + ".b"), ctmp, 512, B * Lc[j], 512, cks[j] * 512, ACT_SILU, nullptr, s, "content agg conv");
        } else {
            GemmParams p = gemm_defaults();
            p.A = enc; p.lda = 512; p.W = c.dev(wn + ".w"); p.bias = c.dev(wn + ".b");
            p.C = ctmp; p.ldc = 512; p.M = B * Lc[j]; p.N = 512; p.Kc = 512; p.taps = cks[j]; p.pad = 0; p.stride = cks[j];
            p.L_out = Lc[j]; p.L_in = T; p.act = ACT_SILU;
            run_gemm(c, p, s, "content agg conv");
        }
        adaptive_pool_kernel<<<ew_grid((size_t)Mc * 512), 256, 0, s>>>(ctmp, 512, Lc[j], ccat2, 2560, 512 * (j + 1), minT, 512, B);
        check_launch(c, "adaptive pool");
    }
    float* cw = c.fbuf("ws.d.cw", (size_t)Mc * 256);
    float* cu = c.fbuf("ws.d.cu", (size_t)Mc * 256);
    float* cv2 = c.fbuf("ws.d.cv2", (size_t)Mc * 256);
    float* ckey = c.fbuf("ws.d.ckey", (size_t)Mc * 256);
    float* cval = c.fbuf("ws.d.cval", (size_t)Mc * 256);
    float* clog = c.fbuf("ws.d.clog", (size_t)Mc * 501);
    linear(c, ccat2, 2560, "d.cbn", c.dev("d.cbn.b"), cw, 256, Mc, 256, 2560, ACT_NONE, nullptr, s, "content bottleneck");
    linear(c, cw, 256, "d.ck0", c.dev("d.ck0.b"), cu, 256, Mc, 256, 256, ACT_SILU, nullptr, s, "content K.0");
    linear(c, cu, 256, "d.ck2", c.dev("d.ck2.b"), ckey, 256, Mc, 256, 256, ACT_SILU, nullptr, s, "content K.2");
    linear(c, cw, 256, "d.cloc0", c.dev("d.cloc0.b"), cu, 256, Mc, 256, 256, ACT_SILU, nullptr, s, "location_fc.0");
    linear(c, cu, 256, "d.cloc2", c.dev("d.cloc2.b"), cv2, 256, Mc, 256, 256, ACT_SILU, nullptr, s, "location_fc.2");
    linear(c, cv2, 256, "d.cloc4", c.dev("d.cloc4.b"), clog, 501, Mc, 501, 256, ACT_SILU, nullptr, s, "location_fc.4");
    gumbel_value_kernel<<<Mc, 256, 501 * sizeof(float), s>>>(clog, gumbel, 1.0f / 0.1f, c.dev("d.cemb"), cval, fw.out_content_dis, 501);
    check_launch(c, "gumbel value");
    // ---- stop-token constant, initial state --------------------------------------------------------
    float* stopc = c.fbuf("ws.d.stopc", (size_t)B);
    linear(c, enc_cell, 512, "d.stop2", nullptr, stopc, 1, B, 1, 512, ACT_NONE, nullptr, s, "stop const");
    float* outputs = c.fbuf("ws.d.outputs", (size_t)B * steps * 80);
    fill_i64_kernel<<<ceil_div(B, 256), 256, 0, s>>>(reinterpret_cast<long long*>(lengths), (long long)steps, B);
    check_launch(c, "lengths init");

    // ---- the 300-step loop: one persistent cooperative kernel per 32 clips -----------------------
    {
        if (c.meta.at("d.step3.ok") != 1) throw L2sError(L2S_ERR_INVALID, "decoder: the stage-pipelined decode kernel needs >= 148 SMs");
        DecodeParams dp{};
        dp.pos = pos;
        dp.temp = c.W("decoder.temperature").f[0]; dp.ctemp = c.W("decoder.content.temperature").f[0];
        dp.Bpad = Bpad; dp.T = T; dp.minT = minT; dp.steps = steps;
        const unsigned char* tfm = nullptr; const float* p1t = nullptr;
        if (fw.mels) {
            // teacher frames -> prenet layer 1 for every step (one GEMM), kept feature-major for the step kernel
            float* tr = c.fbuf("ws.d.trows", (size_t)B * steps * 80);
            teacher_rows_kernel<<<ew_grid((size_t)B * steps * 80), 256, 0, s>>>(fw.mels, c.dev("d.bos"), tr, B, steps);
            check_launch(c, "teacher rows");
            float* p1r = c.fbuf("ws.d.p1rows", (size_t)B * steps * 256);
            linear(c, tr, 80, "d.prenet0", c.dev("d.prenet0.b"), p1r, 256, B * steps, 256, 80, ACT_PSINE, c.dev("d.prenet0.psw"), s, "prenet.0 (teacher)");
            float* p1tb = c.fbuf("ws.d.p1t", (size_t)steps * 256 * Bpad);
            p1_teacher_fm_kernel<<<ew_grid((size_t)steps * 256 * Bpad), 256, 0, s>>>(p1r, p1tb, B, Bpad, steps);
            check_launch(c, "p1 teacher fm");
            unsigned char* dmask = static_cast<unsigned char*>(c.buf("ws.d.tfmask", 512));
            L2S_CUDA(cudaMemcpyAsync(dmask, fw.tf_mask, steps, cudaMemcpyHostToDevice, s));
            tfm = dmask; p1t = p1tb;
        }
        // exchange buffers of tagged fp32 words (decode3.cuh), zero-filled before every call: tag 0 = "never written"
        const size_t xfloats = (size_t)(2 * 1024 + 1024 + 256 + 1024 + 512 + 256) * Bpad;
        float* xbase = c.fbuf("ws.d.xchg", xfloats);
        L2S_CUDA(cudaMemsetAsync(xbase, 0, xfloats * sizeof(float), s));
        float* xS = xbase;
        float* xC = xS + 2 * 2 * plane;
        float* xP1 = xC + 2 * plane;
        float* xXD = xP1 + (size_t)256 * Bpad;
        float* xQ = xXD + (size_t)1024 * Bpad;
        float* xCQ = xQ + (size_t)512 * Bpad;
        unsigned* abortw = static_cast<unsigned*>(c.buf("ws.d.abort", 256));
        // initial state: hidden = the Bi-LSTM's final (h_fwd ; h_bwd), cell.fill_(0) (decoder.py:392,406)
        d3_init_state_kernel<<<ew_grid(2 * plane), 256, 0, s>>>(hfinal, xS, xC, Bpad);
        check_launch(c, "initial decoder state");
        float* vsplit = c.fbuf("ws.d.Vsplit", (size_t)B * T * 512);
        float* cvsplit = c.fbuf("ws.d.cvsplit", (size_t)B * minT * 256);
        split_halves_kernel<<<ew_grid((size_t)B * T * 512), 256, 0, s>>>(Vmem, vsplit, B, T, 256);
        check_launch(c, "V halves");
        split_halves_kernel<<<ew_grid((size_t)B * minT * 256), 256, 0, s>>>(cval, cvsplit, B, minT, 128);
        check_launch(c, "content value halves");
        const size_t smem = (size_t)c.meta.at("d.step3.smem");
        if ((int64_t)smem > c.meta["attr.dec3.smem"]) {
            L2S_CUDA(cudaFuncSetAttribute(decode3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            c.meta["attr.dec3.smem"] = (int64_t)smem;
        }
        const int chunk = D3_CG * D3_NG;
        c.meta["dbg.dec3"] = 1;
        c.span_end("preloop", s);
        c.span_begin("decode_loop", s);
        // The kernel decodes up to 32 clips per launch; larger batches run in consecutive 32-clip chunks (pre-loop and
        // postnet stay whole-batch GEMMs).  A chunk sees its own clips through shifted pointers; the group-major planes keep
        // the whole batch's stride (Bpad).
        for (int b0 = 0; b0 < B; b0 += chunk) {
            const int g0 = b0 / D3_CG;
            Decode3Params q{};
            q.d = dp;
            q.d.B = std::min(chunk, B - b0);
            q.S = xS + (size_t)g0 * 1024 * D3_CG; q.Cst = xC + (size_t)g0 * 1024 * D3_CG;
            q.P1 = xP1 + (size_t)g0 * 256 * D3_CG; q.XD = xXD + (size_t)g0 * 1024 * D3_CG;
            q.Q = xQ + (size_t)b0 * 512; q.CQ = xCQ + (size_t)b0 * 256;
            q.abort_word = abortw;
            q.d.Kmem = Kmem + (size_t)b0 * T * 512; q.d.Vmem = Vmem + (size_t)b0 * T * 512;
            q.d.ckey = ckey + (size_t)b0 * minT * 256; q.d.cval = cval + (size_t)b0 * minT * 256;
            q.d.stop_const = stopc + b0;
            q.d.outputs = outputs + (size_t)b0 * steps * 80; q.d.lengths = reinterpret_cast<long long*>(lengths) + b0;
            q.d.attn = attn ? attn + (size_t)b0 * steps * T : nullptr;
            q.d.tf_mask = tfm; q.d.p1_teacher = p1t ? p1t + b0 : nullptr;
            q.d.stop_out = fw.out_stop ? fw.out_stop + (size_t)b0 * steps : nullptr;
            q.d.attn_logits = fw.out_attn_logits ? fw.out_attn_logits + (size_t)b0 * steps * T : nullptr;
            q.Vsplit = vsplit + (size_t)b0 * T * 512; q.cvsplit = cvsplit + (size_t)b0 * minT * 256;
            q.kv_smem = (size_t)(512 + 320 + 256 + 32) + (size_t)2 * (T * 768 + minT * 384) <= (size_t)c.meta.at("d.step3.wimg_floats") ? 1 : 0;
            q.passes = reinterpret_cast<const Dec3Pass*>(c.dev("d.step3.passes"));
            q.role = reinterpret_cast<const int*>(c.dev("d.step3.role"));
            q.job = reinterpret_cast<const int*>(c.dev("d.step3.job"));
            q.wimg = c.dev("d.step3.wimg"); q.wimg_floats = (int)c.meta.at("d.step3.wimg_floats");
            q.timing = c.profiling ? c.fbuf("ws.d.timing3", (size_t)c.num_sms * D3_TIMING_SLOTS) : nullptr;
            void* args[] = {&q};
            L2S_CUDA(cudaLaunchCooperativeKernel((void*)decode3_kernel, dim3(c.num_sms), dim3(MV_THREADS), args, smem, s));
            c.launches++;
        }
        c.span_end("decode_loop", s);
    }
    c.span_begin("postnet", s);
    // ---- postnet + residual (decoder.py:437-439) --------------------------------------------------
    postnet_rows(c, outputs, B, steps, mel_post, true, s);
    c.span_end("postnet", s);
    if (fw.out_mel) {
        rows_to_bcl_kernel<<<ew_grid((size_t)B * steps * 80), 256, 0, s>>>(outputs, fw.out_mel, B, 80, steps);
        check_launch(c, "outputs -> [B,80,M]");
    }
    c.meta["dbg.B"] = B; c.meta["dbg.T"] = T; c.meta["dbg.minT"] = minT; c.meta["dbg.steps"] = steps;
}

static void decoder_infer(Context& c, const float* visual, const float* spk, const float* gumbel, int B, int T, int steps,
                          float* mel_post, int64_t* lengths, float* attn, cudaStream_t s) {
    decoder_run(c, visual, spk, gumbel, B, T, steps, mel_post, lengths, attn, DecoderForwardIO{}, s);
}

// visual[b,t,:] = [feat[b,t,0:768], emb[b,0:256]]   (model.py:52-55)
__global__ void concat_visual_kernel(const float* __restrict__ feat, const float* __restrict__ emb, float* __restrict__ visual, int B, int T) {
    const size_t total = (size_t)B * T * 1024;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int col = i % 1024; size_t r = i / 1024;
        int b = r / T;
        visual[i] = col < 768 ? feat[r * 768 + col] : emb[(size_t)b * 256 + col - 768];
    }
}
// x [B][C][L] -> rows [B][L][C]
__global__ void bcl_to_rows_kernel(const float* __restrict__ x, float* __restrict__ rows, int B, int C, int L) {
    const size_t total = (size_t)B * C * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int ch = i % C; size_t r = i / C;
        int t = r % L; int b = r / L;
        rows[i] = x[((size_t)b * C + ch) * L + t];
    }
}

// video_ready: optional event after which `video` is valid on the device (the host entry point copies the clips on a
// second stream while the speaker encoder, which only needs the waveforms, already runs)
static void infer_device(Context& c, const VideoSrc& video, const float* wav, const float* gumbel, int B, int T, int H, int W, int S,
                         int steps, float* mel_post, int64_t* lengths, int precision, cudaStream_t s, cudaEvent_t video_ready = nullptr) {
    float* emb = c.fbuf("ws.i.emb", (size_t)B * 256);
    float* feat = c.fbuf("ws.i.feat", (size_t)B * T * 768);
    float* visual = c.fbuf("ws.i.visual", (size_t)B * T * 1024);
    c.span_begin("speaker", s);
    speaker_forward(c, wav, B, S, emb, 1, s);
    c.span_end("speaker", s);
    if (video_ready) L2S_CUDA(cudaStreamWaitEvent(s, video_ready, 0));
    c.span_begin("video", s);
    video_forward(c, video, B, T, H, W, feat, precision, s);
    c.span_end("video", s);
    concat_visual_kernel<<<ew_grid((size_t)B * T * 1024), 256, 0, s>>>(feat, emb, visual, B, T);
    check_launch(c, "concat visual");
    decoder_infer(c, visual, emb, gumbel, B, T, steps, mel_post, lengths, nullptr, s);
}

static void destroy_comm_quiet(Context& c);

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int l2s_version(void) { return 100; }

int l2s_create(l2s_ctx** out, int device) {
    if (!out) return L2S_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) {
        g_create_err = e != cudaSuccess ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e) : "no such CUDA device";
        return L2S_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        return L2S_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_err = "lip2speech_b200 is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return L2S_ERR_CUDA;
    }
    l2s_ctx* ctx = new l2s_ctx();
    ctx->c.device = device;
    ctx->c.num_sms = prop.multiProcessorCount;
    ctx->c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
#ifdef L2S_DEBUG
    if (const char* e = getenv("L2S_TC")) ctx->c.use_tc = (e[0] != '0');
    if (const char* e = getenv("L2S_PW")) ctx->c.use_pw = (e[0] != '0');
    if (const char* e = getenv("L2S_PW_MIN_ROWS")) ctx->c.pw_min_rows = std::max(1, atoi(e));
#endif
    *out = ctx;
    return L2S_OK;
}

void l2s_destroy(l2s_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaDeviceSynchronize();
    destroy_comm_quiet(ctx->c);
    ctx->dec_train.release();
    ctx->video_train.release();
    ctx->c.free_all();
    delete ctx;
}

const char* l2s_last_error(const l2s_ctx* ctx) { return ctx ? ctx->c.err.c_str() : g_create_err.c_str(); }

int l2s_bind_weight(l2s_ctx* ctx, const char* key, const void* ptr, const int64_t* shape, int ndim, int dtype, int on_device) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    if (!key || !ptr || ndim < 0 || ndim > 8) throw L2sError(L2S_ERR_INVALID, "bind_weight: bad arguments");
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const int64_t n = t.numel();
    t.f.resize((size_t)n);
    if (dtype == L2S_DTYPE_F32) {
        if (on_device) L2S_CUDA(cudaMemcpy(t.f.data(), ptr, n * sizeof(float), cudaMemcpyDeviceToHost));
        else std::memcpy(t.f.data(), ptr, n * sizeof(float));
    } else if (dtype == L2S_DTYPE_I64) {
        std::vector<int64_t> tmp((size_t)n);
        if (on_device) L2S_CUDA(cudaMemcpy(tmp.data(), ptr, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
        else std::memcpy(tmp.data(), ptr, n * sizeof(int64_t));
        for (int64_t i = 0; i < n; ++i) t.f[i] = (float)tmp[i];
    } else {
        throw L2sError(L2S_ERR_INVALID, std::string("bind_weight: unsupported dtype for ") + key);
    }
    ctx->c.w[key] = std::move(t);
    if (std::strcmp(key, "vocoder.inv_mel") == 0) ctx->c.meta.erase("voc.ready");      // re-pack the vocoder tables on next use
    API_END(ctx)
}

int l2s_commit_weights(l2s_ctx* ctx, int parts) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    L2S_CUDA(cudaDeviceSynchronize());
    if (parts & L2S_PART_VIDEO) pack_video(ctx->c);
    if (parts & L2S_PART_SPEAKER) pack_speaker(ctx->c);
    if (parts & L2S_PART_DECODER) pack_decoder(ctx->c);
    ctx->c.committed |= parts;
    L2S_CUDA(cudaDeviceSynchronize());
    API_END(ctx)
}

static void need(l2s_ctx* ctx, int part, const char* what) {
    if (!(ctx->c.committed & part)) throw L2sError(L2S_ERR_MISSING_WEIGHT, std::string(what) + ": weights not committed (l2s_commit_weights)");
    L2S_CUDA(cudaSetDevice(ctx->c.device));
}

int l2s_video_fwd(l2s_ctx* ctx, const float* video, int B, int T, int H, int W, float* out_feat, int precision, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO, "video_fwd");
    video_forward(ctx->c, video_f32(video), B, T, H, W, out_feat, precision, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_speaker_fwd(l2s_ctx* ctx, const float* wav, int B, int S, float* emb, int normalize, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_SPEAKER, "speaker_fwd");
    speaker_forward(ctx->c, wav, B, S, emb, normalize, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_decoder_infer(l2s_ctx* ctx, const float* visual, const float* spk, const float* gumbel, int B, int T, int steps,
                      float* mel_post, int64_t* lengths, float* attn, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_DECODER, "decoder_infer");
    decoder_infer(ctx->c, visual, spk, gumbel, B, T, steps, mel_post, lengths, attn, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_decoder_forward(l2s_ctx* ctx, const float* visual, const float* spk, const float* gumbel, const float* mels,
                        const unsigned char* tf_mask, int B, int T, int M, float* out_mel, float* out_post, float* out_stop,
                        float* out_attn_logits, float* out_content_dis, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_DECODER, "decoder_forward");
    if (!mels || !tf_mask || !out_post) throw L2sError(L2S_ERR_INVALID, "decoder_forward: mels, tf_mask and out_post are required");
    DecoderForwardIO fw;
    fw.mels = mels; fw.tf_mask = tf_mask; fw.out_mel = out_mel; fw.out_stop = out_stop;
    fw.out_attn_logits = out_attn_logits; fw.out_content_dis = out_content_dis;
    int64_t* lens = static_cast<int64_t*>(ctx->c.buf("ws.d.fwlens", (size_t)B * sizeof(int64_t)));
    decoder_run(ctx->c, visual, spk, gumbel, B, T, M, out_post, lens, nullptr, fw, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_postnet_fwd(l2s_ctx* ctx, const float* x, int B, int L, float* out, int add_residual, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_DECODER, "postnet_fwd");
    if (B <= 0 || L <= 0) throw L2sError(L2S_ERR_INVALID, "postnet_fwd: bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    float* rows = ctx->c.fbuf("ws.p.rows", (size_t)B * L * 80);
    bcl_to_rows_kernel<<<ew_grid((size_t)B * L * 80), 256, 0, s>>>(x, rows, B, 80, L);
    check_launch(ctx->c, "bcl->rows");
    postnet_rows(ctx->c, rows, B, L, out, add_residual != 0, s);
    API_END(ctx)
}

int l2s_infer(l2s_ctx* ctx, const float* video, const float* wav, const float* gumbel, int B, int T, int H, int W, int S, int steps,
              float* mel_post, int64_t* lengths, int precision, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO | L2S_PART_SPEAKER | L2S_PART_DECODER, "infer");
    infer_device(ctx->c, video_f32(video), wav, gumbel, B, T, H, W, S, steps, mel_post, lengths, precision, (cudaStream_t)stream);
    API_END(ctx)
}

// Host-buffer entry points.  Two slots of device staging buffers: `submit` enqueues H2D copies (the 100 MB clip tensor on
// a second stream, overlapped with the speaker encoder — which needs only the waveforms — and, when the caller keeps two
// submissions in flight, with the previous batch's compute), the whole span, and the D2H copies of the results; `wait`
// blocks until that slot's results are in the caller's buffers.
static void infer_host_submit(Context& c, int slot, const VideoSrc& video, const float* wav, const float* gumbel, int B, int T, int H, int W, int S,
                              int steps, float* mel_post, int64_t* lengths, int precision) {
    if (slot < 0 || slot > 1) throw L2sError(L2S_ERR_INVALID, "infer_host: slot must be 0 or 1");
    int minT = T;
    for (int k : {1, 3, 5, 7}) minT = std::min(minT, (T - k) / k + 1);
    const size_t nv = (size_t)B * 3 * T * H * W, nw = (size_t)B * S, ng = (size_t)B * minT * 501, nm = (size_t)B * 80 * steps;
    const std::string sfx = slot ? ".1" : "";
    // grow every staging buffer of this slot before anything is enqueued (a reallocation synchronises the device)
    const size_t vbytes = nv * (video.kind == 0 ? sizeof(float) : 1);
    void* dv = c.buf("ws.h.video" + sfx, vbytes); float* dw = c.fbuf("ws.h.wav" + sfx, nw); float* dg = c.fbuf("ws.h.gumbel" + sfx, ng);
    float* dm = c.fbuf("ws.h.mel" + sfx, nm);
    int64_t* dl = static_cast<int64_t*>(c.buf("ws.h.len" + sfx, (size_t)B * sizeof(int64_t)));
    if (!c.host_stream) {
        L2S_CUDA(cudaStreamCreateWithFlags(&c.host_stream, cudaStreamNonBlocking));
        L2S_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            L2S_CUDA(cudaEventCreateWithFlags(&c.copy_done[i], cudaEventDisableTiming));
            L2S_CUDA(cudaEventCreateWithFlags(&c.slot_done[i], cudaEventDisableTiming));
            L2S_CUDA(cudaEventCreateWithFlags(&c.video_consumed[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t s = c.host_stream;
    // the copy stream must not overwrite this slot's clip buffer while an earlier submission still reads it
    if (c.slot_used[slot]) L2S_CUDA(cudaStreamWaitEvent(c.copy_stream, c.video_consumed[slot], 0));
    L2S_CUDA(cudaMemcpyAsync(dv, video.p, vbytes, cudaMemcpyHostToDevice, c.copy_stream));
    L2S_CUDA(cudaEventRecord(c.copy_done[slot], c.copy_stream));
    L2S_CUDA(cudaMemcpyAsync(dw, wav, nw * sizeof(float), cudaMemcpyHostToDevice, s));
    L2S_CUDA(cudaMemcpyAsync(dg, gumbel, ng * sizeof(float), cudaMemcpyHostToDevice, s));
    VideoSrc dsrc = video; dsrc.p = dv;
    infer_device(c, dsrc, dw, dg, B, T, H, W, S, steps, dm, dl, precision, s, c.copy_done[slot]);
    L2S_CUDA(cudaEventRecord(c.video_consumed[slot], s));
    L2S_CUDA(cudaMemcpyAsync(mel_post, dm, nm * sizeof(float), cudaMemcpyDeviceToHost, s));
    L2S_CUDA(cudaMemcpyAsync(lengths, dl, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    L2S_CUDA(cudaEventRecord(c.slot_done[slot], s));
    c.slot_used[slot] = true;
}

int l2s_infer_host_submit(l2s_ctx* ctx, int slot, const float* video, const float* wav, const float* gumbel, int B, int T, int H, int W, int S,
                          int steps, float* mel_post, int64_t* lengths, int precision) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO | L2S_PART_SPEAKER | L2S_PART_DECODER, "infer_host_submit");
    infer_host_submit(ctx->c, slot, video_f32(video), wav, gumbel, B, T, H, W, S, steps, mel_post, lengths, precision);
    API_END(ctx)
}

int l2s_infer_host_wait(l2s_ctx* ctx, int slot) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (slot < 0 || slot > 1 || !c.slot_used[slot]) throw L2sError(L2S_ERR_INVALID, "infer_host_wait: nothing was submitted to this slot");
    L2S_CUDA(cudaEventSynchronize(c.slot_done[slot]));
    API_END(ctx)
}

int l2s_video_fwd_u8(l2s_ctx* ctx, const unsigned char* frames, const float* mean_std, int B, int T, int H, int W, float* out_feat, int precision, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO, "video_fwd_u8");
    if (!frames || !mean_std) throw L2sError(L2S_ERR_INVALID, "video_fwd_u8: frames and mean_std are required");
    video_forward(ctx->c, video_u8(frames, mean_std), B, T, H, W, out_feat, precision, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_infer_u8(l2s_ctx* ctx, const unsigned char* frames, const float* mean_std, const float* wav, const float* gumbel, int B, int T, int H, int W,
                 int S, int steps, float* mel_post, int64_t* lengths, int precision, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO | L2S_PART_SPEAKER | L2S_PART_DECODER, "infer_u8");
    if (!frames || !mean_std) throw L2sError(L2S_ERR_INVALID, "infer_u8: frames and mean_std are required");
    infer_device(ctx->c, video_u8(frames, mean_std), wav, gumbel, B, T, H, W, S, steps, mel_post, lengths, precision, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_infer_host_submit_u8(l2s_ctx* ctx, int slot, const unsigned char* frames, const float* mean_std, const float* wav, const float* gumbel,
                             int B, int T, int H, int W, int S, int steps, float* mel_post, int64_t* lengths, int precision) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO | L2S_PART_SPEAKER | L2S_PART_DECODER, "infer_host_submit_u8");
    if (!frames || !mean_std) throw L2sError(L2S_ERR_INVALID, "infer_host_submit_u8: frames and mean_std are required");
    infer_host_submit(ctx->c, slot, video_u8(frames, mean_std), wav, gumbel, B, T, H, W, S, steps, mel_post, lengths, precision);
    API_END(ctx)
}

int l2s_infer_host(l2s_ctx* ctx, const float* video, const float* wav, const float* gumbel, int B, int T, int H, int W, int S, int steps,
                   float* mel_post, int64_t* lengths, int precision) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    need(ctx, L2S_PART_VIDEO | L2S_PART_SPEAKER | L2S_PART_DECODER, "infer_host");
    infer_host_submit(ctx->c, 0, video_f32(video), wav, gumbel, B, T, H, W, S, steps, mel_post, lengths, precision);
    L2S_CUDA(cudaEventSynchronize(ctx->c.slot_done[0]));
    API_END(ctx)
}

// ---- train-step tail ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl_api() {
    static NcclApi a;
    if (!a.h) {
        // resolve into a local table and publish it only when every symbol is there: a failed attempt must not leave a
        // half-filled static behind (the next call would jump through a null pointer)
        NcclApi t;
        t.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!t.h) throw L2sError(L2S_ERR_CUDA, std::string("dlopen(libnccl.so.2): ") + dlerror());
        auto sym = [&](const char* n) {
            void* p = dlsym(t.h, n);
            if (!p) { dlclose(t.h); throw L2sError(L2S_ERR_CUDA, std::string("libnccl.so.2 lacks ") + n); }
            return p;
        };
        t.GetUniqueId = reinterpret_cast<decltype(t.GetUniqueId)>(sym("ncclGetUniqueId"));
        t.CommInitRank = reinterpret_cast<decltype(t.CommInitRank)>(sym("ncclCommInitRank"));
        t.AllReduce = reinterpret_cast<decltype(t.AllReduce)>(sym("ncclAllReduce"));
        t.CommDestroy = reinterpret_cast<decltype(t.CommDestroy)>(sym("ncclCommDestroy"));
        t.GetErrorString = reinterpret_cast<decltype(t.GetErrorString)>(sym("ncclGetErrorString"));
        a = t;
    }
    return a;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw L2sError(L2S_ERR_CUDA, std::string(what) + ": " + nccl_api().GetErrorString(r));
}
int ts_grid(Context& c, size_t n4) {
    const size_t want = (n4 + TS_THREADS - 1) / TS_THREADS;
    const size_t cap = (size_t)c.num_sms * 8;                 // a multiple of the SM count; each thread streams several float4
    return (int)std::max<size_t>(1, std::min(want, cap));
}
}  // namespace

}  // extern "C"
static void destroy_comm_quiet(Context& c) {
    if (!c.nccl_comm) return;
    try { nccl_api().CommDestroy(static_cast<ncclComm_t>(c.nccl_comm)); } catch (...) {}
    c.nccl_comm = nullptr;
}
extern "C" {

int l2s_nccl_unique_id(void* out, int nbytes) {
    if (!out || nbytes < (int)sizeof(ncclUniqueId)) return L2S_ERR_INVALID;
    try {
        ncclUniqueId id;
        nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(out, &id, sizeof(id));
    } catch (const std::exception& e) { g_create_err = e.what(); return L2S_ERR_CUDA; }
    return L2S_OK;
}

int l2s_comm_init(l2s_ctx* ctx, const void* unique_id, int nbytes, int rank, int world) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!unique_id || nbytes < (int)sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) throw L2sError(L2S_ERR_INVALID, "comm_init: bad arguments");
    if (c.nccl_comm) throw L2sError(L2S_ERR_INVALID, "comm_init: communicator already initialised");
    L2S_CUDA(cudaSetDevice(c.device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    nccl_check(nccl_api().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    c.nccl_comm = comm; c.world = world; c.rank = rank;
    API_END(ctx)
}

int l2s_comm_destroy(l2s_ctx* ctx) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (c.nccl_comm) { nccl_check(nccl_api().CommDestroy(static_cast<ncclComm_t>(c.nccl_comm)), "ncclCommDestroy"); c.nccl_comm = nullptr; c.world = 1; c.rank = 0; }
    API_END(ctx)
}

int l2s_allreduce_grads(l2s_ctx* ctx, float* flat_grads, int64_t n, float scale, float* sqnorm_out, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!flat_grads || n <= 0 || !sqnorm_out) throw L2sError(L2S_ERR_INVALID, "allreduce_grads: bad arguments");
    if (reinterpret_cast<uintptr_t>(flat_grads) & 15) throw L2sError(L2S_ERR_INVALID, "allreduce_grads: the flat buffer must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    L2S_CUDA(cudaSetDevice(c.device));
    if (c.nccl_comm && c.world > 1)
        nccl_check(nccl_api().AllReduce(flat_grads, flat_grads, (size_t)n, ncclFloat32, ncclSum, static_cast<ncclComm_t>(c.nccl_comm), s), "ncclAllReduce");
    const int grid = ts_grid(c, (size_t)n / 4);
    double* part = static_cast<double*>(c.buf("ws.t.part", (size_t)c.num_sms * 8 * 4 * sizeof(double)));
    grad_scale_sqnorm_kernel<<<grid, TS_THREADS, 0, s>>>(flat_grads, (size_t)n, scale, part);
    check_launch(c, "grad scale + norm");
    sqnorm_finish_kernel<<<1, 32, 0, s>>>(part, grid, sqnorm_out);
    check_launch(c, "grad norm finish");
    API_END(ctx)
}

int l2s_clip_adamw_step(l2s_ctx* ctx, float* p, float* g, float* m, float* v, float* vmax, int64_t n, const float* sqnorm,
                        double max_norm, double lr, double beta1, double beta2, double eps, double weight_decay, int step, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!p || !g || !m || !v || !vmax || n <= 0 || step < 1 || (max_norm > 0.0 && !sqnorm)) throw L2sError(L2S_ERR_INVALID, "clip_adamw_step: bad arguments");
    for (const void* q : {(const void*)p, (const void*)g, (const void*)m, (const void*)v, (const void*)vmax})
        if (reinterpret_cast<uintptr_t>(q) & 15) throw L2sError(L2S_ERR_INVALID, "clip_adamw_step: flat buffers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    L2S_CUDA(cudaSetDevice(c.device));
    // hyper-parameters arrive as doubles (Python floats): every derived constant is formed in double and rounded once, as torch does
    AdamWParams a{(float)lr, (float)beta1, (float)beta2, (float)eps, (float)weight_decay, (float)max_norm,
                  (float)(1.0 - std::pow(beta1, (double)step)), (float)std::sqrt(1.0 - std::pow(beta2, (double)step)),
                  (float)(lr / (1.0 - std::pow(beta1, (double)step))), (float)(1.0 - beta1), (float)(1.0 - beta2), (float)(1.0 - lr * weight_decay)};
    clip_adamw_kernel<<<ts_grid(c, (size_t)n / 4), TS_THREADS, 0, s>>>(p, g, m, v, vmax, (size_t)n, sqnorm, a);
    check_launch(c, "clip + AdamW");
    API_END(ctx)
}

int l2s_loss_fwd_bwd(l2s_ctx* ctx, const float* mel_out, const float* mel_post, const float* gate_logits, const float* content_dis,
                     const float* mel_target, const float* gate_target, int B, int M, int rows, float* losses,
                     float* g_mel, float* g_post, float* g_gate, float* g_content_dis, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!mel_out || !mel_post || !gate_logits || !content_dis || !mel_target || !gate_target || !losses || B <= 0 || M <= 0 || rows <= 0)
        throw L2sError(L2S_ERR_INVALID, "loss_fwd_bwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    L2S_CUDA(cudaSetDevice(c.device));
    const size_t n_mel = (size_t)B * 80 * M, n_gate = (size_t)B * M, n_dis = (size_t)rows * 501;
    const int grid = ts_grid(c, std::max(n_mel, n_dis) / 4);      // a thread handles ~4 elements of the longest array
    double* part = static_cast<double*>(c.buf("ws.t.part", (size_t)c.num_sms * 8 * 4 * sizeof(double)));
    loss_partial_kernel<<<grid, TS_THREADS, 0, s>>>(mel_out, mel_post, mel_target, n_mel, gate_logits, gate_target, n_gate, content_dis, n_dis, 501,
                                                    g_mel, g_post, g_gate, g_content_dis, part);
    check_launch(c, "loss partial sums");
    loss_finish_kernel<<<1, 32, 0, s>>>(part, grid, n_mel, n_gate, (size_t)rows, losses);
    check_launch(c, "loss finish");
    API_END(ctx)
}

// ---- vocoder + ESTOI (the steps after the path: demo.py:89-90, evaluate.py:41-45) -------------------------------------------
static void pack_audio_tables(Context& c) {
    if (c.meta.count("voc.ready")) return;
    const HostTensor& inv = c.W("vocoder.inv_mel");           // [513][80] = pinv of the mel filterbank (InverseMelScale's lstsq operator)
    if (inv.numel() != (int64_t)VOC_BINS * 80) throw L2sError(L2S_ERR_INVALID, "vocoder.inv_mel must be [513,80]");
    c.upload("voc.inv.w", inv.f);
    upload_tc(c, "voc.inv", inv.f, VOC_BINS, 1, 80);
    std::vector<float> win(VOC_NFFT);
    for (int n = 0; n < VOC_NFFT; ++n) win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / VOC_NFFT));       // torch.hann_window (periodic)
    c.upload("voc.win", win);
    // analysis: rows f (re) and IM + f (im) of X[f] = sum_n w[n] x[n] e^{-2 pi i f n / N}
    std::vector<float> dft((size_t)VOC_LD * VOC_NFFT, 0.f);
    // synthesis: x[n] = w[n]/N * sum_f c_f (Re X_f cos - Im X_f sin), c_0 = c_{N/2} = 1, else 2 (imaginary parts of DC / Nyquist ignored)
    std::vector<float> idft((size_t)VOC_NFFT * VOC_LD, 0.f);
    for (int f = 0; f < VOC_BINS; ++f)
        for (int n = 0; n < VOC_NFFT; ++n) {
            const int j = (int)(((long long)f * n) % VOC_NFFT);                  // exact argument reduction
            const double a = 2.0 * M_PI * (double)j / VOC_NFFT, cs = std::cos(a), sn = std::sin(a);
            dft[(size_t)f * VOC_NFFT + n] = (float)(win[n] * cs);
            dft[(size_t)(VOC_IM + f) * VOC_NFFT + n] = (float)(-(double)win[n] * sn);
            const double cf = (f == 0 || f == VOC_NFFT / 2) ? 1.0 : 2.0;
            idft[(size_t)n * VOC_LD + f] = (float)(win[n] * cf * cs / VOC_NFFT);
            idft[(size_t)n * VOC_LD + VOC_IM + f] = (f == 0 || f == VOC_NFFT / 2) ? 0.f : (float)(-(double)win[n] * cf * sn / VOC_NFFT);
        }
    c.upload("voc.dft.w", dft); upload_tc(c, "voc.dft", dft, VOC_LD, 1, VOC_NFFT);
    c.upload("voc.idft.w", idft); upload_tc(c, "voc.idft", idft, VOC_NFFT, 1, VOC_LD);
    c.meta["voc.ready"] = 1;
}

static void pack_estoi_tables(Context& c) {
    if (c.meta.count("es.ready")) return;
    {
        auto bessel_i0 = [](double x) { double s = 1.0, t = 1.0; for (int k = 1; k < 60; ++k) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; } return s; };
        std::vector<double> h(ES_TAPS);
        const double fc = 1.0 / 8.0, alpha = (ES_TAPS - 1) / 2.0;
        double sum = 0.0;
        for (int n = 0; n < ES_TAPS; ++n) {
            const double m = n - alpha, r = m / alpha;
            const double kw = bessel_i0(5.0 * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(5.0);
            const double x = fc * m, sinc = (x == 0.0) ? 1.0 : std::sin(M_PI * x) / (M_PI * x);
            h[n] = fc * sinc * kw; sum += h[n];
        }
        for (auto& v : h) v = v / sum * 5.0;                                     // firwin normalisation (unit DC gain) x up
        c.upload_raw("es.h", h.data(), h.size());
        std::vector<double> w(ES_FRAME);
        for (int n = 0; n < ES_FRAME; ++n) w[n] = 0.5 - 0.5 * std::cos(2.0 * M_PI * (n + 1) / (ES_FRAME + 1));      // np.hanning(258)[1:-1]
        c.upload_raw("es.win", w.data(), w.size());
        std::vector<int> lo(ES_BANDS), hi(ES_BANDS);
        for (int i = 0; i < ES_BANDS; ++i) {                                      // pystoi thirdoct(10000, 512, 15, 150)
            const double fl = 150.0 * std::pow(2.0, (2.0 * i - 1) / 6.0), fh = 150.0 * std::pow(2.0, (2.0 * i + 1) / 6.0);
            auto nearest = [](double target) {
                int best = 0; double bd = 1e300;
                for (int k = 0; k <= ES_NFFT / 2; ++k) { const double f = (double)ES_FS * k / ES_NFFT, d = (f - target) * (f - target); if (d < bd) { bd = d; best = k; } }
                return best;
            };
            lo[i] = nearest(fl); hi[i] = nearest(fh);
        }
        c.upload_raw("es.lo", lo.data(), lo.size()); c.upload_raw("es.hi", hi.data(), hi.size());
    }
    c.meta["es.ready"] = 1;
}

// ---- train-mode forward / backward ------------------------------------------------------------------------------------
int l2s_train_bind(l2s_ctx* ctx, const char* key, float* param, float* grad, int64_t numel) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    if (!key || !param || numel <= 0) throw L2sError(L2S_ERR_INVALID, "train_bind: bad arguments");
    tr::Param p; p.v = param; p.g = grad; p.n = numel;
    tr::Param& slot = ctx->train_params[key];
    if (slot.v != p.v || slot.g != p.g || slot.n != p.n) { slot = p; ++ctx->train_bind_gen; }
    API_END(ctx)
}

int l2s_train_set_graphs(l2s_ctx* ctx, int enabled) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    ctx->dec_train.use_graphs = ctx->video_train.use_graphs = enabled != 0;
    if (!enabled) { L2S_CUDA(cudaDeviceSynchronize()); ctx->dec_train.drop_graphs(); ctx->video_train.drop_graphs(); }
    API_END(ctx)
}

int l2s_decoder_train_fwd(l2s_ctx* ctx, const float* visual, const float* spk, const float* mels, const unsigned char* tf_mask, const float* gumbel,
                          const float* prenet_mask, const float* attn_mask, const float* lstm_mask, const float* const* post_masks, int B, int T, int M,
                          int want_input_grads, float* out_mel, float* out_post, float* out_stop, float* out_attn_logits, float* out_content_dis,
                          void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    if (!visual || !spk || !mels || !tf_mask || !gumbel || !prenet_mask || !attn_mask || !lstm_mask || !post_masks)
        throw L2sError(L2S_ERR_INVALID, "decoder_train_fwd: every input and every noise tensor is required");
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    tr::DecoderTrainIO io{};
    io.visual = visual; io.spk = spk; io.mels = mels; io.tf_mask = tf_mask; io.gumbel = gumbel;
    io.prenet_mask = prenet_mask; io.attn_mask = attn_mask; io.lstm_mask = lstm_mask;
    for (int i = 0; i < 5; ++i) { if (!post_masks[i]) throw L2sError(L2S_ERR_INVALID, "decoder_train_fwd: five postnet masks are required"); io.post_mask[i] = post_masks[i]; }
    io.out_mel = out_mel; io.out_post = out_post; io.out_stop = out_stop; io.out_attn_logits = out_attn_logits; io.out_content_dis = out_content_dis;
    ctx->dec_train.forward(ctx->c, ctx->train_params, ctx->train_bind_gen, io, B, T, M, want_input_grads != 0, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_decoder_train_bwd(l2s_ctx* ctx, const float* g_mel, const float* g_post, const float* g_stop, const float* g_content_dis, float* g_visual,
                          float* g_spk, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    ctx->dec_train.backward(ctx->c, g_mel, g_post, g_stop, g_content_dis, g_visual, g_spk, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_vocoder(l2s_ctx* ctx, const float* mel, const float* init_angles, int B, int L, int n_iter, float momentum, float* wav, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!mel || !wav || B <= 0 || L < 3 || n_iter < 0 || momentum < 0.f || momentum >= 1.f) throw L2sError(L2S_ERR_INVALID, "vocoder: bad arguments");
    L2S_CUDA(cudaSetDevice(c.device));
    pack_audio_tables(c);
    cudaStream_t s = (cudaStream_t)stream;
    const int R = B * L, n = (L - 1) * VOC_HOP;
    float* melrows = c.fbuf("ws.voc.melrows", (size_t)R * 80);
    float* mag = c.fbuf("ws.voc.mag", (size_t)R * VOC_IM);
    float* ang = c.fbuf("ws.voc.ang", (size_t)R * VOC_LD);
    float* tprev = c.fbuf("ws.voc.tprev", (size_t)R * VOC_LD);
    float* spec = c.fbuf("ws.voc.spec", (size_t)R * VOC_LD);
    float* frames = c.fbuf("ws.voc.frames", (size_t)R * VOC_NFFT);
    float* inv = c.fbuf("ws.voc.inv", (size_t)B * n);
    voc_exp_rows_kernel<<<ew_grid((size_t)R * 80), 256, 0, s>>>(mel, melrows, B, L);
    check_launch(c, "vocoder exp");
    linear(c, melrows, 80, "voc.inv", nullptr, mag, VOC_IM, R, VOC_BINS, 80, ACT_NONE, nullptr, s, "inverse mel scale");
    voc_mag_kernel<<<ew_grid((size_t)R * VOC_IM), 256, 0, s>>>(mag, R);
    check_launch(c, "vocoder magnitude");
    voc_init_angles_kernel<<<ew_grid((size_t)R * VOC_LD), 256, 0, s>>>(init_angles, ang, tprev, B, L);
    check_launch(c, "vocoder initial phase");
    const float mfac = momentum / (1.f + momentum);
    for (int it = 0; it <= n_iter; ++it) {
        float* out = it == n_iter ? wav : inv;
        voc_apply_kernel<<<ew_grid((size_t)R * VOC_LD), 256, 0, s>>>(mag, ang, spec, R);
        check_launch(c, "vocoder apply phase");
        linear(c, spec, VOC_LD, "voc.idft", nullptr, frames, VOC_NFFT, R, VOC_NFFT, VOC_LD, ACT_NONE, nullptr, s, "inverse DFT");
        voc_ola_kernel<<<ew_grid((size_t)B * n), 256, 0, s>>>(frames, c.dev("voc.win"), out, B, L);
        check_launch(c, "overlap-add");
        if (it == n_iter) break;
        voc_frames_kernel<<<ew_grid((size_t)R * VOC_NFFT), 256, 0, s>>>(inv, frames, B, L);
        check_launch(c, "stft frames");
        linear(c, frames, VOC_NFFT, "voc.dft", nullptr, spec, VOC_LD, R, VOC_LD, VOC_NFFT, ACT_NONE, nullptr, s, "DFT");
        voc_angle_kernel<<<ew_grid((size_t)R * VOC_BINS), 256, 0, s>>>(spec, tprev, ang, R, mfac);
        check_launch(c, "phase update");
    }
    API_END(ctx)
}

int l2s_estoi(l2s_ctx* ctx, const float* clean, const float* processed, int B, int S, double* out, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    Context& c = ctx->c;
    if (!clean || !processed || !out || B <= 0 || S < 1024) throw L2sError(L2S_ERR_INVALID, "estoi: bad arguments");
    L2S_CUDA(cudaSetDevice(c.device));
    pack_estoi_tables(c);
    cudaStream_t s = (cudaStream_t)stream;
    const int n10 = (int)(((long long)S * 5 + 7) / 8);
    const int max_frames = n10 >= ES_FRAME ? 1 + (n10 - ES_FRAME) / ES_HOP : 1;
    double* x10 = static_cast<double*>(c.buf("ws.es.x10", (size_t)B * n10 * 8));
    double* y10 = static_cast<double*>(c.buf("ws.es.y10", (size_t)B * n10 * 8));
    double* xs = static_cast<double*>(c.buf("ws.es.xs", (size_t)B * n10 * 8));
    double* ys = static_cast<double*>(c.buf("ws.es.ys", (size_t)B * n10 * 8));
    double* xt = static_cast<double*>(c.buf("ws.es.xt", (size_t)B * ES_BANDS * max_frames * 8));
    double* yt = static_cast<double*>(c.buf("ws.es.yt", (size_t)B * ES_BANDS * max_frames * 8));
    int* nk = static_cast<int*>(c.buf("ws.es.nk", (size_t)B * 4));
    const double* h = reinterpret_cast<const double*>(c.dev("es.h"));
    const double* win = reinterpret_cast<const double*>(c.dev("es.win"));
    const int* lo = reinterpret_cast<const int*>(c.dev("es.lo")); const int* hi = reinterpret_cast<const int*>(c.dev("es.hi"));
    es_resample_kernel<<<ew_grid((size_t)B * n10), 256, 0, s>>>(clean, x10, B, S, n10, h);
    check_launch(c, "estoi resample");
    es_resample_kernel<<<ew_grid((size_t)B * n10), 256, 0, s>>>(processed, y10, B, S, n10, h);
    check_launch(c, "estoi resample");
    es_silent_kernel<<<B, 256, (size_t)max_frames * 12 + 16, s>>>(x10, y10, n10, win, xs, ys, nk, max_frames);
    check_launch(c, "estoi silent frames");
    es_bands_kernel<<<dim3(max_frames, B), 256, 0, s>>>(xs, n10, nk, win, lo, hi, xt, max_frames);
    check_launch(c, "estoi bands");
    es_bands_kernel<<<dim3(max_frames, B), 256, 0, s>>>(ys, n10, nk, win, lo, hi, yt, max_frames);
    check_launch(c, "estoi bands");
    es_score_kernel<<<B, 512, 0, s>>>(xt, yt, nk, out, max_frames);
    check_launch(c, "estoi score");
    API_END(ctx)
}

int l2s_video_train_fwd(l2s_ctx* ctx, const float* video, const float* drop_mask, int B, int T, int H, int W, float* out_feat, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    if (!video || !out_feat) throw L2sError(L2S_ERR_INVALID, "video_train_fwd: video and out_feat are required");
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    ctx->video_train.forward(ctx->c, ctx->train_params, ctx->train_bind_gen, video, drop_mask, B, T, H, W, out_feat, (cudaStream_t)stream);
    API_END(ctx)
}

int l2s_video_train_bwd(l2s_ctx* ctx, const float* g_feat, void* stream) {
    if (!ctx) return L2S_ERR_INVALID;
    API_BEGIN
    if (!g_feat) throw L2sError(L2S_ERR_INVALID, "video_train_bwd: g_feat is required");
    L2S_CUDA(cudaSetDevice(ctx->c.device));
    ctx->video_train.backward(ctx->c, g_feat, (cudaStream_t)stream);
    API_END(ctx)
}

int64_t l2s_launch_count(const l2s_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int l2s_set_profiling(l2s_ctx* ctx, int enabled) {
    if (!ctx) return L2S_ERR_INVALID;
    ctx->c.profiling = enabled != 0;
    return L2S_OK;
}

double l2s_span_ms(l2s_ctx* ctx, const char* name) {
    if (!ctx || !name) return -1.0;
    auto it = ctx->c.spans.find(name);
    if (it == ctx->c.spans.end() || !it->second.used) return -1.0;
    cudaSetDevice(ctx->c.device);
    if (cudaEventSynchronize(it->second.e1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, it->second.e0, it->second.e1) != cudaSuccess) return -1.0;
    return (double)ms;
}

int64_t l2s_debug_read(l2s_ctx* ctx, const char* name, float* out, int64_t n) {
    if (!ctx || !name) return -1;
    Context& c = ctx->c;
    auto get = [&](const char* k) { auto it = c.meta.find(k); return it == c.meta.end() ? (int64_t)0 : it->second; };
    if (std::strcmp(name, "flag.dec_abort") == 0) {          // 1: a wait inside the decode kernel gave up (decode3.cuh: LLWait) — results invalid
        auto it = c.bufs.find("ws.d.abort");
        if (it == c.bufs.end() || !it->second.p) return 0;
        unsigned v = 0;
        cudaSetDevice(c.device); cudaDeviceSynchronize();
        cudaMemcpy(&v, it->second.p, sizeof(v), cudaMemcpyDeviceToHost);
        return (int64_t)v;
    }
    if (std::strncmp(name, "flag.", 5) == 0) return get((std::string("dbg.") + (name + 5)).c_str());
    const int64_t B = get("dbg.B"), T = get("dbg.T"), minT = get("dbg.minT"), steps = get("dbg.steps");
    struct { const char* name; const char* buf; int64_t count; } tab[] = {
        {"dec.K", "ws.d.K", B * T * 512}, {"dec.V", "ws.d.V", B * T * 512}, {"dec.enc_cell", "ws.d.enc_cell", B * 512},
        {"dec.enc", "ws.d.enc", B * T * 512}, {"dec.rnn_out", "ws.d.rnnout", B * T * 1024},
        {"dec.ckey", "ws.d.ckey", B * minT * 256}, {"dec.cval", "ws.d.cval", B * minT * 256},
        {"dec.outputs", "ws.d.outputs", B * steps * 80}, {"dec.clog", "ws.d.clog", B * minT * 501},
        {"dec.timing3", "ws.d.timing3", (int64_t)c.num_sms * D3_TIMING_SLOTS},
    };
    for (auto& t : tab) {
        if (std::strcmp(t.name, name) == 0) {
            auto it = c.bufs.find(t.buf);
            if (it == c.bufs.end() || !it->second.p) return -1;
            cudaSetDevice(c.device);
            cudaDeviceSynchronize();
            int64_t k = std::min<int64_t>(n, t.count);
            if (out && k > 0) cudaMemcpy(out, it->second.p, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost);
            return t.count;
        }
    }
    return -1;
}

}  // extern "C"
