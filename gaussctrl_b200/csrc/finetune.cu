// 3DGS fine-tune step after the edit (gc_trainer.py:257-301 -> SplatfactoModel.get_loss_dict ->
// loss.backward() -> Adam): the photometric loss with its gradient, and the Adam update of the Gaussian parameters.
//
//   main_loss = (1 - l) * mean|gt - pred| + l * (1 - SSIM(gt, pred))        l = ssim_lambda (0.2)
//   SSIM: 11-tap Gaussian (sigma 1.5), separable VALID blur of pred, gt, pred^2, gt^2, pred*gt per channel,
//         ssim = (2 mu_p mu_g + C1)(2 s_pg + C2) / ((mu_p^2 + mu_g^2 + C1)(s_p + s_g + C2)), mean over the
//         (H-10) x (W-10) x C map.
// Forward and backward in four streaming passes over channels-last fp32 images [H,W,C] (HBM/L2-bound: 512^2 x 3 floats
// = 3 MB per map, every map stays in the 126 MB L2 between passes):
//   1. horizontal blur of the five products                       -> hb [5][H][Wo][C]
//   2. vertical blur, ssim value and its partials w.r.t. the three blurred maps that depend on pred
//      (mu_p, E[p^2], E[p g]); deterministic per-block partial sums -> dm [3][Ho][Wo][C], part_ssim[blocks]
//   3. transposed vertical blur of the partials                   -> tb [3][H][Wo][C]
//   4. transposed horizontal blur, chain rule to pred, + L1 term  -> v_pred [H][W][C], part_l1[blocks]
//   5. fixed-order sum of the partials                            -> loss_out[3] = (main_loss, L1, ssim)
// No atomics: the loss value and the gradient are bit-reproducible run to run.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int WIN = 11;
constexpr int LOSS_THREADS = 256;
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct Window {
    float w[WIN];
};

// pytorch_msssim._fspecial_gauss_1d(11, 1.5), evaluated in fp32 like torch does
Window make_window() {
    Window win;
    float sum = 0.f;
    for (int i = 0; i < WIN; ++i) {
        const float c = (float)(i - WIN / 2);
        win.w[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
        sum += win.w[i];
    }
    for (int i = 0; i < WIN; ++i) win.w[i] /= sum;
    return win;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < LOSS_THREADS / 32; ++i) t += red[i];  // fixed order
    return t;  // valid in thread 0
}

// pass 1: one thread per (y, xo, c); row stride W*C, tap stride C
__global__ void __launch_bounds__(LOSS_THREADS)
ssim_hblur_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ hb, int H, int W,
                  int C, const Window win) {
    const int Wo = W - (WIN - 1);
    const long long plane = (long long)H * Wo * C;
    const long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (i >= plane) return;
    const int c = (int)(i % C);
    const int xo = (int)((i / C) % Wo);
    const int y = (int)(i / ((long long)C * Wo));
    const long long src = ((long long)y * W + xo) * C + c;
    float sp = 0.f, sg = 0.f, spp = 0.f, sgg = 0.f, spg = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
        const float p = pred[src + (long long)k * C], g = gt[src + (long long)k * C], w = win.w[k];
        sp = fmaf(w, p, sp);
        sg = fmaf(w, g, sg);
        spp = fmaf(w, p * p, spp);
        sgg = fmaf(w, g * g, sgg);
        spg = fmaf(w, p * g, spg);
    }
    hb[i] = sp;
    hb[plane + i] = sg;
    hb[2 * plane + i] = spp;
    hb[3 * plane + i] = sgg;
    hb[4 * plane + i] = spg;
}

// pass 2: one thread per (yo, xo, c)
__global__ void __launch_bounds__(LOSS_THREADS)
ssim_vblur_map_kernel(const float* __restrict__ hb, float* __restrict__ dm, float* __restrict__ part_ssim, int H, int W,
                      int C, const Window win) {
    __shared__ float red[LOSS_THREADS / 32];
    const int Wo = W - (WIN - 1), Ho = H - (WIN - 1);
    const long long row = (long long)Wo * C;
    const long long plane_in = (long long)H * row, plane_out = (long long)Ho * row;
    const long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    float s = 0.f;
    if (i < plane_out) {
        float mp = 0.f, mg = 0.f, epp = 0.f, egg = 0.f, epg = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const long long j = i + (long long)k * row;  // (yo + k, xo, c)
            const float w = win.w[k];
            mp = fmaf(w, hb[j], mp);
            mg = fmaf(w, hb[plane_in + j], mg);
            epp = fmaf(w, hb[2 * plane_in + j], epp);
            egg = fmaf(w, hb[3 * plane_in + j], egg);
            epg = fmaf(w, hb[4 * plane_in + j], epg);
        }
        const float sp = epp - mp * mp, sg = egg - mg * mg, spg = epg - mp * mg;
        const float A1 = 2.f * mp * mg + SSIM_C1, A2 = 2.f * spg + SSIM_C2;
        const float B1 = mp * mp + mg * mg + SSIM_C1, B2 = sp + sg + SSIM_C2;
        const float rB1 = 1.f / B1, rB2 = 1.f / B2;
        s = A1 * A2 * rB1 * rB2;
        // partials of s w.r.t. mu_p (E[p^2], E[pg] held fixed), E[p^2], E[pg]
        dm[i] = 2.f * mg * (A2 - A1) * rB1 * rB2 - 2.f * mp * s * (rB1 - rB2);
        dm[plane_out + i] = -s * rB2;
        dm[2 * plane_out + i] = 2.f * A1 * rB1 * rB2;
    }
    const float t = block_sum(s, red);
    if (threadIdx.x == 0) part_ssim[blockIdx.x] = t;
}

// pass 3: one thread per (y, xo, c): tb[y] = sum_k w[k] * dm[y - k], 0 <= y-k < Ho
__global__ void __launch_bounds__(LOSS_THREADS)
ssim_vblur_t_kernel(const float* __restrict__ dm, float* __restrict__ tb, int H, int W, int C, const Window win) {
    const int Wo = W - (WIN - 1), Ho = H - (WIN - 1);
    const long long row = (long long)Wo * C;
    const long long plane_in = (long long)Ho * row, plane_out = (long long)H * row;
    const long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (i >= plane_out) return;
    const int y = (int)(i / row);
    const long long col = i - (long long)y * row;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
        const int yo = y - k;
        if (yo >= 0 && yo < Ho) {
            const long long j = (long long)yo * row + col;
            const float w = win.w[k];
            a0 = fmaf(w, dm[j], a0);
            a1 = fmaf(w, dm[plane_in + j], a1);
            a2 = fmaf(w, dm[2 * plane_in + j], a2);
        }
    }
    tb[i] = a0;
    tb[plane_out + i] = a1;
    tb[2 * plane_out + i] = a2;
}

// pass 4: one thread per (y, x, c)
__global__ void __launch_bounds__(LOSS_THREADS)
ssim_hblur_t_grad_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ tb,
                         float* __restrict__ v_pred, float* __restrict__ part_l1, int H, int W, int C, const Window win,
                         float ssim_scale, float l1_scale) {
    __shared__ float red[LOSS_THREADS / 32];
    const int Wo = W - (WIN - 1);
    const long long n = (long long)H * W * C;
    const long long plane = (long long)H * Wo * C;
    const long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x;
    float ad = 0.f;
    if (i < n) {
        const int c = (int)(i % C);
        const int x = (int)((i / C) % W);
        const int y = (int)(i / ((long long)C * W));
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const int xo = x - k;
            if (xo >= 0 && xo < Wo) {
                const long long j = ((long long)y * Wo + xo) * C + c;
                const float w = win.w[k];
                a0 = fmaf(w, tb[j], a0);
                a1 = fmaf(w, tb[plane + j], a1);
                a2 = fmaf(w, tb[2 * plane + j], a2);
            }
        }
        const float p = pred[i], g = gt[i];
        const float d = p - g;
        ad = fabsf(d);
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        v_pred[i] = ssim_scale * (a0 + 2.f * p * a1 + g * a2) + l1_scale * sgn;
    }
    const float t = block_sum(ad, red);
    if (threadIdx.x == 0) part_l1[blockIdx.x] = t;
}

// pass 5: one block, fixed summation order
__global__ void __launch_bounds__(LOSS_THREADS)
loss_finalize_kernel(const float* __restrict__ part_ssim, int n_ssim, const float* __restrict__ part_l1, int n_l1,
                     float inv_n_ssim, float inv_n_l1, float lambda, float* __restrict__ loss_out) {
    __shared__ float red[LOSS_THREADS / 32];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < n_ssim; i += LOSS_THREADS) a += part_ssim[i];
    for (int i = threadIdx.x; i < n_l1; i += LOSS_THREADS) b += part_l1[i];
    const float sa = block_sum(a, red);
    __syncthreads();
    const float sb = block_sum(b, red);
    if (threadIdx.x == 0) {
        const float ssim = sa * inv_n_ssim, l1 = sb * inv_n_l1;
        loss_out[0] = (1.f - lambda) * l1 + lambda * (1.f - ssim);
        loss_out[1] = l1;
        loss_out[2] = ssim;
    }
}

struct LossLayout {
    long long hb, dm, tb, part_ssim, part_l1, total;  // float offsets
    int blocks_map, blocks_img;
};

LossLayout loss_layout(int H, int W, int C) {
    LossLayout L;
    const long long Wo = W - (WIN - 1), Ho = H - (WIN - 1);
    L.blocks_map = gcb_cdiv(Ho * Wo * C, LOSS_THREADS);
    L.blocks_img = gcb_cdiv((long long)H * W * C, LOSS_THREADS);
    L.hb = 0;
    L.dm = L.hb + 5 * H * Wo * C;
    L.tb = L.dm + 3 * Ho * Wo * C;
    L.part_ssim = L.tb + 3 * H * Wo * C;
    L.part_l1 = L.part_ssim + L.blocks_map;
    L.total = L.part_l1 + L.blocks_img;
    return L;
}

// ------------------------------------------------------------------------------------------------------- Adam
constexpr int ADAM_MAX_TENSORS = 8;
struct AdamBatch {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
    float step_size[ADAM_MAX_TENSORS];  // lr / (1 - beta1^t)
};

// torch.optim.Adam (single-tensor path, no weight decay / amsgrad):
//   m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;  p -= step_size * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float one_minus_b1, float b2,
                                            float one_minus_b2, float rsqrt_bc2, float eps, float step_size) {
    m = m + (g - m) * one_minus_b1;
    v = v * b2 + one_minus_b2 * g * g;
    const float denom = sqrtf(v) * rsqrt_bc2 + eps;
    p = p - step_size * (m / denom);
}

// blockIdx.y = tensor; 28 B of traffic per parameter (read p,g,m,v, write p,m,v): HBM-bound, float4 lanes
__global__ void __launch_bounds__(256)
adam_kernel(const AdamBatch t, float one_minus_b1, float b2, float one_minus_b2, float rsqrt_bc2, float eps) {
    const int ti = blockIdx.y;
    float* __restrict__ p = t.p[ti];
    const float* __restrict__ g = t.g[ti];
    float* __restrict__ m = t.m[ti];
    float* __restrict__ v = t.v[ti];
    const long long n = t.n[ti];
    const float ss = t.step_size[ti];
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_update(pp.x, gg.x, mm.x, vv.x, one_minus_b1, b2, one_minus_b2, rsqrt_bc2, eps, ss);
        adam_update(pp.y, gg.y, mm.y, vv.y, one_minus_b1, b2, one_minus_b2, rsqrt_bc2, eps, ss);
        adam_update(pp.z, gg.z, mm.z, vv.z, one_minus_b1, b2, one_minus_b2, rsqrt_bc2, eps, ss);
        adam_update(pp.w, gg.w, mm.w, vv.w, one_minus_b1, b2, one_minus_b2, rsqrt_bc2, eps, ss);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_update(pp, g[i], mm, vv, one_minus_b1, b2, one_minus_b2, rsqrt_bc2, eps, ss);
        p[i] = pp;
        m[i] = mm;
        v[i] = vv;
    }
}

}  // namespace

extern "C" size_t gcb_l1_ssim_workspace_bytes(int H, int W, int C) {
    if (H < WIN || W < WIN || C <= 0) return 0;
    return (size_t)loss_layout(H, W, C).total * sizeof(float);
}

extern "C" int gcb_l1_ssim_loss_fwd_bwd(const float* pred, const float* gt, int H, int W, int C, float ssim_lambda,
                                        float* loss_out, float* v_pred, void* workspace, size_t workspace_bytes,
                                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GCB_CHECK_ARG(pred && gt && loss_out && v_pred && workspace, "l1_ssim_loss: null pointer");
    GCB_CHECK_ARG(H >= WIN && W >= WIN && C > 0, "l1_ssim_loss: image %dx%dx%d smaller than the %d-tap window", H, W, C, WIN);
    const LossLayout L = loss_layout(H, W, C);
    if (workspace_bytes < (size_t)L.total * sizeof(float)) {
        gcb_set_error("l1_ssim_loss: workspace %zu < %zu bytes", workspace_bytes, (size_t)L.total * sizeof(float));
        return GCB_ERR_WORKSPACE;
    }
    float* ws = (float*)workspace;
    const Window win = make_window();
    const long long Wo = W - (WIN - 1), Ho = H - (WIN - 1);
    const long long n_map = Ho * Wo * C, n_img = (long long)H * W * C, n_h = (long long)H * Wo * C;
    const int blocks_h = gcb_cdiv(n_h, LOSS_THREADS);
    ssim_hblur_kernel<<<blocks_h, LOSS_THREADS, 0, stream>>>(pred, gt, ws + L.hb, H, W, C, win);
    GCB_LAUNCH_CHECK();
    ssim_vblur_map_kernel<<<L.blocks_map, LOSS_THREADS, 0, stream>>>(ws + L.hb, ws + L.dm, ws + L.part_ssim, H, W, C, win);
    GCB_LAUNCH_CHECK();
    ssim_vblur_t_kernel<<<blocks_h, LOSS_THREADS, 0, stream>>>(ws + L.dm, ws + L.tb, H, W, C, win);
    GCB_LAUNCH_CHECK();
    // d main_loss / d ssim_map = -lambda / n_map;  d main_loss / d |p-g| = (1-lambda) / n_img
    ssim_hblur_t_grad_kernel<<<L.blocks_img, LOSS_THREADS, 0, stream>>>(pred, gt, ws + L.tb, v_pred, ws + L.part_l1, H, W, C,
                                                                       win, -ssim_lambda / (float)n_map,
                                                                       (1.f - ssim_lambda) / (float)n_img);
    GCB_LAUNCH_CHECK();
    loss_finalize_kernel<<<1, LOSS_THREADS, 0, stream>>>(ws + L.part_ssim, L.blocks_map, ws + L.part_l1, L.blocks_img,
                                                        1.f / (float)n_map, 1.f / (float)n_img, ssim_lambda, loss_out);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_adam_step(int n_tensors, void* const* h_params, const void* const* h_grads, void* const* h_exp_avg,
                             void* const* h_exp_avg_sq, const long long* h_numel, const double* h_lr, double beta1,
                             double beta2, double eps, int step, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GCB_CHECK_ARG(n_tensors >= 0 && h_params && h_grads && h_exp_avg && h_exp_avg_sq && h_numel && h_lr,
                  "adam_step: null pointer");
    GCB_CHECK_ARG(step >= 1, "adam_step: step counts from 1 (got %d)", step);
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const float rsqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const int sms = gcb_sm_count();
    for (int base = 0; base < n_tensors; base += ADAM_MAX_TENSORS) {
        AdamBatch b;
        memset(&b, 0, sizeof(b));
        const int cnt = n_tensors - base < ADAM_MAX_TENSORS ? n_tensors - base : ADAM_MAX_TENSORS;
        long long nmax = 0;
        for (int i = 0; i < cnt; ++i) {
            GCB_CHECK_ARG(h_params[base + i] && h_grads[base + i] && h_exp_avg[base + i] && h_exp_avg_sq[base + i],
                          "adam_step: null tensor %d", base + i);
            GCB_CHECK_ARG(((uintptr_t)h_params[base + i] % 16) == 0 && ((uintptr_t)h_grads[base + i] % 16) == 0 &&
                              ((uintptr_t)h_exp_avg[base + i] % 16) == 0 && ((uintptr_t)h_exp_avg_sq[base + i] % 16) == 0,
                          "adam_step: tensor %d is not 16-byte aligned", base + i);
            b.p[i] = (float*)h_params[base + i];
            b.g[i] = (const float*)h_grads[base + i];
            b.m[i] = (float*)h_exp_avg[base + i];
            b.v[i] = (float*)h_exp_avg_sq[base + i];
            b.n[i] = h_numel[base + i];
            b.step_size[i] = (float)(h_lr[base + i] / bc1);
            if (b.n[i] > nmax) nmax = b.n[i];
        }
        if (nmax == 0) continue;
        // grid.x: enough float4 lanes for the largest tensor, capped at 8 CTAs per SM (grid-stride covers the rest)
        long long gx = (nmax / 4 + 255) / 256;
        if (gx < 1) gx = 1;
        if (gx > (long long)sms * 8) gx = (long long)sms * 8;
        adam_kernel<<<dim3((unsigned)gx, (unsigned)cnt), 256, 0, stream>>>(b, (float)(1.0 - beta1), (float)beta2,
                                                                            (float)(1.0 - beta2), rsqrt_bc2, (float)eps);
        GCB_LAUNCH_CHECK();
    }
    return GCB_OK;
}
