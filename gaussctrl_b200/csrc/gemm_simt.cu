// Bring-up / comparison GEMM (mma.sync m16n8k16, implicit conv gather), the direct small-channel convolution,
// im2col for stride-2 convs, and the public conv2d dispatcher.
#include <stdlib.h>
#include <string.h>

#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

int gcb_gemm_tc_supported(int B, int H, int W, int Cin, int Cout, int ksize, int act);
void gcb_gemm_tc_set_schedule(int schedule);
int gcb_gemm_tc_launch(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                       const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                       int direct_epilogue, cudaStream_t stream);

namespace {

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 64x64 tile, BK = 32, 4 warps (2x2), each warp 32x32.  K index = tap*Cin + c.
constexpr int TM = 64, TN = 64, TK = 32, PADK = TK + 8;

__global__ void __launch_bounds__(128) gemm_mma_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                       const __half* __restrict__ bias,
                                                       const __half* __restrict__ rowvec, int rowvec_ld,
                                                       const __half* __restrict__ residual, __half* __restrict__ y,
                                                       int B, int H, int W, int Cin, int Cout, int ksize, int act) {
    __shared__ __align__(16) __half As[TM][PADK];
    __shared__ __align__(16) __half Bs[TN][PADK];
    const int M = B * H * W, K = ksize * ksize * Cin, HW = H * W;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[i][j][t] = 0.f;
    const int pad = ksize / 2;
    for (int k0 = 0; k0 < K; k0 += TK) {
        // A: 64 rows x 32 k = 256 chunks of 8 halves; 128 threads x 2
        for (int c = threadIdx.x; c < TM * (TK / 8); c += 128) {
            const int r = c / (TK / 8), kc = (c % (TK / 8)) * 8;
            const int m = m0 + r, k = k0 + kc;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (m < M && k < K) {
                const int tap = k / Cin, ci = k - tap * Cin;  // Cin % 8 == 0 => a chunk never straddles taps
                const int b = m / HW, rem = m - b * HW, h = rem / W, ww = rem - h * W;
                const int hh = h + tap / ksize - pad, w2 = ww + tap % ksize - pad;
                if (hh >= 0 && hh < H && w2 >= 0 && w2 < W)
                    v = *reinterpret_cast<const uint4*>(x + ((long long)(b * H + hh) * W + w2) * Cin + ci);
            }
            *reinterpret_cast<uint4*>(&As[r][kc]) = v;
        }
        for (int c = threadIdx.x; c < TN * (TK / 8); c += 128) {
            const int r = c / (TK / 8), kc = (c % (TK / 8)) * 8;
            const int n = n0 + r, k = k0 + kc;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (n < Cout && k < K) v = *reinterpret_cast<const uint4*>(w + (long long)n * K + k);
            *reinterpret_cast<uint4*>(&Bs[r][kc]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk += 16) {
            uint32_t a[2][4], bfr[4][2];
            const int kq = kk + (lane & 3) * 2, rq = lane >> 2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i][0] = *reinterpret_cast<const uint32_t*>(&As[wm + i * 16 + rq][kq]);
                a[i][1] = *reinterpret_cast<const uint32_t*>(&As[wm + i * 16 + rq + 8][kq]);
                a[i][2] = *reinterpret_cast<const uint32_t*>(&As[wm + i * 16 + rq][kq + 8]);
                a[i][3] = *reinterpret_cast<const uint32_t*>(&As[wm + i * 16 + rq + 8][kq + 8]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bfr[j][0] = *reinterpret_cast<const uint32_t*>(&Bs[wn + j * 8 + rq][kq]);
                bfr[j][1] = *reinterpret_cast<const uint32_t*>(&Bs[wn + j * 8 + rq][kq + 8]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_16816(acc[i][j], a[i], bfr[j]);
        }
        __syncthreads();
    }
    const int ldy = Cout;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int m = m0 + wm + i * 16 + (lane >> 2) + (t >= 2 ? 8 : 0);
                const int n = n0 + wn + j * 8 + (lane & 3) * 2 + (t & 1);
                if (m < M && n < Cout) {
                    float v = acc[i][j][t];
                    if (bias) v += __half2float(bias[n]);
                    if (rowvec) v += __half2float(rowvec[(long long)(m / HW) * rowvec_ld + n]);
                    if (act == GCB_ACT_SILU) v = silu_f(v);
                    if (residual) v += __half2float(residual[(long long)m * ldy + n]);
                    y[(long long)m * ldy + n] = __float2half_rn(v);
                }
            }
}

// One thread per (pixel, output channel); weights [Cout, taps*Cin].
__global__ void conv_direct_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                   const __half* __restrict__ bias, const __half* __restrict__ residual,
                                   __half* __restrict__ y, int B, int H, int W, int Cin, int Cout, int ksize,
                                   int stride, int pad_lo, int Ho, int Wo, int act) {
    const long long total = (long long)B * Ho * Wo * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long pix = i / Cout;
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
        float acc = bias ? __half2float(bias[co]) : 0.f;
        const __half* wr = w + (long long)co * ksize * ksize * Cin;
        for (int kh = 0; kh < ksize; ++kh) {
            const int hi = ho * stride + kh - pad_lo;
            if (hi < 0 || hi >= H) continue;
            for (int kw = 0; kw < ksize; ++kw) {
                const int wi = wo * stride + kw - pad_lo;
                if (wi < 0 || wi >= W) continue;
                const __half* xp = x + ((long long)(b * H + hi) * W + wi) * Cin;
                const __half* wp = wr + (kh * ksize + kw) * Cin;
                for (int ci = 0; ci < Cin; ++ci) acc += __half2float(xp[ci]) * __half2float(wp[ci]);
            }
        }
        if (act == GCB_ACT_SILU) acc = silu_f(acc);
        if (residual) acc += __half2float(residual[i]);
        y[i] = __float2half_rn(acc);
    }
}

// 3x3 stride-1 pad-1 convolution with a tiny input channel count (conv_in 4->320 of UNet/ControlNet, VAE conv_in):
// one CTA = 64 output pixels x all Cout; weights [K=9*Cin][Cout] and the pixels' input patches live in shared memory;
// a thread owns two adjacent output channels (half2 weights, coalesced half2 stores).
constexpr int SC_PIX = 64;
__global__ void __launch_bounds__(256) conv_small_cin_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                             const __half* __restrict__ bias,
                                                             const __half* __restrict__ residual,
                                                             __half* __restrict__ y, int B, int H, int W, int Cin,
                                                             int Cout, int act) {
    extern __shared__ __align__(16) uint8_t sc_smem[];
    const int K = 9 * Cin, cp = Cout / 2;
    __half2* ws = reinterpret_cast<__half2*>(sc_smem);                     // [K][cp]
    float* xs = reinterpret_cast<float*>(sc_smem + (size_t)K * cp * 4);    // [SC_PIX][K]
    const long long npix = (long long)B * H * W;
    const long long p0 = (long long)blockIdx.x * SC_PIX;
    for (int i = threadIdx.x; i < K * cp; i += blockDim.x) {
        const int k = i / cp, c2 = i % cp;
        ws[i] = __halves2half2(w[(long long)(2 * c2) * K + k], w[(long long)(2 * c2 + 1) * K + k]);
    }
    for (int i = threadIdx.x; i < SC_PIX * K; i += blockDim.x) {
        const int pl = i / K, k = i % K;
        const long long pix = p0 + pl;
        float v = 0.f;
        if (pix < npix) {
            const int tap = k / Cin, ci = k % Cin;
            const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
            const long long b = pix / ((long long)W * H);
            const int hi = hq + tap / 3 - 1, wi = wq + tap % 3 - 1;
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __half2float(x[((b * H + hi) * W + wi) * Cin + ci]);
        }
        xs[i] = v;
    }
    __syncthreads();
    const int groups = blockDim.x / cp;  // pixel groups working in parallel
    const int g = threadIdx.x / cp, c2 = threadIdx.x % cp;
    if (g >= groups) return;
    float2 bv = make_float2(0.f, 0.f);
    if (bias) bv = __half22float2(*reinterpret_cast<const __half2*>(bias + 2 * c2));
    for (int pl = g; pl < SC_PIX; pl += groups) {
        const long long pix = p0 + pl;
        if (pix >= npix) break;
        float a0 = bv.x, a1 = bv.y;
        const float* xr = xs + pl * K;
        for (int k = 0; k < K; ++k) {
            const float2 wf = __half22float2(ws[k * cp + c2]);
            a0 = fmaf(xr[k], wf.x, a0);
            a1 = fmaf(xr[k], wf.y, a1);
        }
        if (act == GCB_ACT_SILU) {
            a0 = silu_f(a0);
            a1 = silu_f(a1);
        }
        if (residual) {
            const float2 r = __half22float2(*reinterpret_cast<const __half2*>(residual + pix * Cout + 2 * c2));
            a0 += r.x;
            a1 += r.y;
        }
        *reinterpret_cast<__half2*>(y + pix * Cout + 2 * c2) = __floats2half2_rn(a0, a1);
    }
}

__global__ void im2col3x3_s2_kernel(const __half* __restrict__ x, __half* __restrict__ col, int B, int H, int W, int C,
                                    int pad_lo, int Ho, int Wo) {
    const int cv = C / 8;
    const long long total = (long long)B * Ho * Wo * 9 * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int tap = (int)(r % 9);
        r /= 9;
        const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
        const int hi = ho * 2 + tap / 3 - pad_lo, wi = wo * 2 + tap % 3 - pad_lo;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (hi >= 0 && hi < H && wi >= 0 && wi < W)
            v = *reinterpret_cast<const uint4*>(x + ((long long)(b * H + hi) * W + wi) * C + c8 * 8);
        *reinterpret_cast<uint4*>(col + (r * 9 + tap) * (long long)C + c8 * 8) = v;
    }
}

// 3x3 / stride 1 / pad 1 patches of a 4-channel tensor (the latents: conv_in of UNet and ControlNet) as rows of 40
// halves: 9 taps x 4 channels in (kh, kw, c) order + 4 zero columns, so that conv_in runs as a K = 40 GEMM on the
// tensor cores (the SIMT direct convolution needed 640 us for 72 x 64 x 64 pixels: 20x its HBM floor).
__global__ void im2col3x3_c4_kernel(const __half* __restrict__ x, __half* __restrict__ col, int B, int H, int W) {
    const long long total = (long long)B * H * W;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        const int wo = (int)(r % W), ho = (int)((r / W) % H);
        const long long img = r / ((long long)W * H) * H;
        uint2 t[10];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int hi = ho + tap / 3 - 1, wi = wo + tap % 3 - 1;
            t[tap] = make_uint2(0, 0);
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) t[tap] = *reinterpret_cast<const uint2*>(x + ((img + hi) * W + wi) * 4);
        }
        t[9] = make_uint2(0, 0);
        uint4* o = reinterpret_cast<uint4*>(col + r * 40);
#pragma unroll
        for (int j = 0; j < 5; ++j) o[j] = make_uint4(t[2 * j].x, t[2 * j].y, t[2 * j + 1].x, t[2 * j + 1].y);
    }
}

}  // namespace

extern "C" int gcb_im2col3x3_c4_nhwc(const void* x, void* col, int B, int H, int W, void* stream) {
    GCB_CHECK_ARG(x && col && B > 0 && H > 0 && W > 0, "bad arguments");
    const long long total = (long long)B * H * W;
    const int blocks = (int)((total + 255) / 256 < 148ll * 16 ? (total + 255) / 256 : 148ll * 16);
    im2col3x3_c4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)col, B, H, W);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_conv2d_nhwc_fwd(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                                   const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize,
                                   int act, int impl, void* stream) {
    GCB_CHECK_ARG(x && w && y, "null tensor pointer");
    GCB_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad shape B=%d H=%d W=%d Cin=%d Cout=%d", B, H, W,
                  Cin, Cout);
    GCB_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3 (got %d)", ksize);
    GCB_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0, "Cin (%d) and Cout (%d) must be multiples of 8; use gcb_conv2d_direct",
                  Cin, Cout);
    GCB_CHECK_ARG((long long)B * H * W < (1ll << 31), "M too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (const char* e = getenv("GCB_FORCE_GEMM_IMPL")) impl = atoi(e);
    if (impl == GCB_GEMM_TCGEN05 || impl == GCB_GEMM_TCGEN05_DIRECT || impl == GCB_GEMM_TCGEN05_PERSISTENT ||
        impl == GCB_GEMM_TCGEN05_ONE_TILE) {
        // schedule of the tcgen05 kernel: by shape unless an A/B tool forces one
        gcb_gemm_tc_set_schedule(impl == GCB_GEMM_TCGEN05_PERSISTENT ? 1 : impl == GCB_GEMM_TCGEN05_ONE_TILE ? 0 : -1);
        if (!gcb_gemm_tc_supported(B, H, W, Cin, Cout, ksize, act)) {
            gcb_set_error("tcgen05 path does not support B=%d H=%d W=%d Cin=%d Cout=%d k=%d act=%d", B, H, W, Cin, Cout,
                          ksize, act);
            return GCB_ERR_UNSUPPORTED;
        }
        return gcb_gemm_tc_launch(x, w, bias, rowvec, rowvec_ld, residual, y, B, H, W, Cin, Cout, ksize, act,
                                  impl == GCB_GEMM_TCGEN05_DIRECT, st);
    }
    GCB_CHECK_ARG(impl == GCB_GEMM_MMA_SYNC, "unknown impl %d", impl);
    GCB_CHECK_ARG(act != GCB_ACT_GEGLU, "GEGLU epilogue exists only on the tcgen05 path");
    dim3 grid(gcb_cdiv(Cout, TN), gcb_cdiv((long long)B * H * W, TM));
    gemm_mma_kernel<<<grid, 128, 0, st>>>((const __half*)x, (const __half*)w, (const __half*)bias,
                                          (const __half*)rowvec, rowvec_ld, (const __half*)residual, (__half*)y, B, H,
                                          W, Cin, Cout, ksize, act);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_conv2d_direct_nhwc_fwd(const void* x, const void* w, const void* bias, const void* residual, void* y,
                                          int B, int H, int W, int Cin, int Cout, int ksize, int stride, int pad_lo,
                                          int pad_hi, int act, void* stream) {
    GCB_CHECK_ARG(x && w && y, "null tensor pointer");
    GCB_CHECK_ARG(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    GCB_CHECK_ARG(stride == 1 || stride == 2, "stride must be 1 or 2");
    GCB_CHECK_ARG(act == GCB_ACT_NONE || act == GCB_ACT_SILU, "unsupported act");
    const int Ho = (H + pad_lo + pad_hi - ksize) / stride + 1, Wo = (W + pad_lo + pad_hi - ksize) / stride + 1;
    if (ksize == 3 && stride == 1 && pad_lo == 1 && pad_hi == 1 && Cin <= 8 && Cout % 2 == 0 && Cout >= 64 && Cout <= 512) {
        const int cp = Cout / 2, K = 9 * Cin;
        const int threads_sc = cp >= 256 ? cp : (256 / cp) * cp;
        const size_t smem = (size_t)K * cp * 4 + (size_t)SC_PIX * K * 4;
        static size_t configured = 0;
        if (smem > configured) {
            GCB_CUDA(cudaFuncSetAttribute(conv_small_cin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            configured = 100 * 1024;
        }
        const long long npix = (long long)B * H * W;
        conv_small_cin_kernel<<<(unsigned)((npix + SC_PIX - 1) / SC_PIX), threads_sc, smem, (cudaStream_t)stream>>>(
            (const __half*)x, (const __half*)w, (const __half*)bias, (const __half*)residual, (__half*)y, B, H, W, Cin,
            Cout, act);
        GCB_LAUNCH_CHECK();
        return GCB_OK;
    }
    const long long total = (long long)B * Ho * Wo * Cout;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads < 148ll * 64 ? (total + threads - 1) / threads : 148ll * 64);
    conv_direct_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        (const __half*)x, (const __half*)w, (const __half*)bias, (const __half*)residual, (__half*)y, B, H, W, Cin,
        Cout, ksize, stride, pad_lo, Ho, Wo, act);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_im2col3x3_s2_nhwc(const void* x, void* col, int B, int H, int W, int C, int pad_lo, int pad_hi,
                                     void* stream) {
    GCB_CHECK_ARG(x && col && C % 8 == 0, "im2col needs C %% 8 == 0 (C=%d)", C);
    const int Ho = (H + pad_lo + pad_hi - 3) / 2 + 1, Wo = (W + pad_lo + pad_hi - 3) / 2 + 1;
    const long long total = (long long)B * Ho * Wo * 9 * (C / 8);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads < 148ll * 32 ? (total + threads - 1) / threads : 148ll * 32);
    im2col3x3_s2_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)col, B, H, W, C, pad_lo,
                                                                      Ho, Wo);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
