// Elementwise / data-movement kernels of the denoising loop (all HBM-bound, 128-bit vectorised, grid-stride
// over a multiple of the SM count).  Call sites replaced: see include/gaussctrl_b200.h.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

inline int ew_blocks(long long work_items, int threads = 256) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = (long long)gcb_sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

#define GRID_STRIDE(i, n) \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void silu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n) {
    const long long nv = n / 8;
    GRID_STRIDE(i, nv) {
        const uint4 r = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_half2(w[t]);
            o[t] = pack_half2(silu_f(f.x), silu_f(f.y));
        }
        reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    GRID_STRIDE(j, n - nv * 8) { y[nv * 8 + j] = __float2half_rn(silu_f(__half2float(x[nv * 8 + j]))); }
}

__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ y,
                           long long n, float alpha, float beta) {
    const long long nv = n / 8;
    GRID_STRIDE(i, nv) {
        const uint4 ra = reinterpret_cast<const uint4*>(a)[i];
        const uint4 rb = reinterpret_cast<const uint4*>(b)[i];
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 fa = unpack_half2(wa[t]), fb = unpack_half2(wb[t]);
            o[t] = pack_half2(alpha * fa.x + beta * fb.x, alpha * fa.y + beta * fb.y);
        }
        reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    GRID_STRIDE(j, n - nv * 8) {
        const long long k = nv * 8 + j;
        y[k] = __float2half_rn(alpha * __half2float(a[k]) + beta * __half2float(b[k]));
    }
}

__global__ void geglu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long M, int C) {
    const int cv = C / 8;
    GRID_STRIDE(i, M * cv) {
        const long long m = i / cv;
        const int v = (int)(i % cv);
        const uint4 rv = *reinterpret_cast<const uint4*>(x + m * 2 * C + v * 8);
        const uint4 rg = *reinterpret_cast<const uint4*>(x + m * 2 * C + C + v * 8);
        const uint32_t wv[4] = {rv.x, rv.y, rv.z, rv.w}, wg[4] = {rg.x, rg.y, rg.z, rg.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 fv = unpack_half2(wv[t]), fg = unpack_half2(wg[t]);
            o[t] = pack_half2(fv.x * gelu_erf_f(fg.x), fv.y * gelu_erf_f(fg.y));
        }
        *reinterpret_cast<uint4*>(y + m * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
    const int cv = C / 8;
    const long long total = (long long)B * 2 * H * 2 * W * cv;
    GRID_STRIDE(i, total) {
        const int v = (int)(i % cv);
        long long r = i / cv;
        const int wo = (int)(r % (2 * W));
        r /= 2 * W;
        const int ho = (int)(r % (2 * H));
        const int b = (int)(r / (2 * H));
        reinterpret_cast<uint4*>(y)[i] =
            *reinterpret_cast<const uint4*>(x + (((long long)b * H + ho / 2) * W + wo / 2) * C + v * 8);
    }
}

// diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32 math -> fp16
__global__ void temb_kernel(const float* __restrict__ t, __half* __restrict__ y, int B, int dim) {
    const int half = dim / 2;
    GRID_STRIDE(i, (long long)B * half) {
        const int b = (int)(i / half), j = (int)(i % half);
        const float freq = expf(-9.210340371976184f * (float)j / (float)half);
        const float a = t[b] * freq;
        y[(long long)b * dim + j] = __float2half_rn(cosf(a));
        y[(long long)b * dim + half + j] = __float2half_rn(sinf(a));
    }
}

__global__ void nchw_to_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int C, int HW) {
    GRID_STRIDE(i, (long long)B * C * HW) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int p = (int)(r % HW), b = (int)(r / HW);
        y[i] = x[((long long)b * C + c) * HW + p];
    }
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int C, int HW) {
    GRID_STRIDE(i, (long long)B * C * HW) {
        const int p = (int)(i % HW);
        const long long r = i / HW;
        const int c = (int)(r % C), b = (int)(r / C);
        y[i] = x[((long long)b * HW + p) * C + c];
    }
}

// 32x32 smem-tiled transpose: x [rows, cols] -> y [cols, rows], batched over blockIdx.z
__global__ void transpose_kernel(const __half* __restrict__ x, __half* __restrict__ y, int rows, int cols) {
    __shared__ __half tile[32][33];
    const long long boff = (long long)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = x[boff + (long long)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) y[boff + (long long)c * rows + r] = tile[threadIdx.x][j];
    }
}

// coef (device fp32[4]): sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)
__global__ void cfg_ddim_kernel(const __half* __restrict__ eu, const __half* __restrict__ ec,
                                const __half* __restrict__ x, __half* __restrict__ xo, long long n, float g,
                                const float* __restrict__ coef) {
    const float sa = coef[0], s1a = coef[1], sp = coef[2], s1p = coef[3];
    GRID_STRIDE(i, n) {
        float e = __half2float(eu[i]);
        if (ec) {
            // fp16 arithmetic order of diffusers: uncond + g * (text - uncond), each op rounded to fp16
            const __half d = __float2half_rn(__half2float(ec[i]) - e);
            const __half gd = __float2half_rn(g * __half2float(d));
            e = __half2float(__float2half_rn(e + __half2float(gd)));
        }
        const float xv = __half2float(x[i]);
        const float x0 = (xv - s1a * e) / sa;
        xo[i] = __float2half_rn(sp * x0 + s1p * e);
    }
}

__global__ void postprocess_kernel(const __half* __restrict__ img, const float* __restrict__ mask,
                                   const __half* __restrict__ uned, float* __restrict__ out, long long npix) {
    GRID_STRIDE(i, npix * 3) {
        const long long p = i / 3;
        // (image / 2 + 0.5).clamp(0, 1) in fp16 as the pipeline's image processor does, then the mask composite
        float v = __half2float(__float2half_rn(__half2float(__float2half_rn(__half2float(img[i]) * 0.5f)) + 0.5f));
        v = fminf(fmaxf(v, 0.f), 1.f);
        if (mask) {
            const float m = mask[p];
            v = v * m + __half2float(uned[i]) * (1.f - m);
        }
        out[i] = v;
    }
}

__global__ void disparity_max_kernel(const float* __restrict__ depth, float* __restrict__ ws, int HW, int round_f16) {
    // one CTA per image: max of 1/(d+1e-5)
    const int b = blockIdx.x;
    float m = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        float d = depth[(long long)b * HW + i];
        float v;
        if (round_f16) {
            const __half dh = __float2half_rn(d);
            const __half s = __float2half_rn(__half2float(dh) + 1e-5f);
            v = __half2float(__float2half_rn(1.f / __half2float(s)));
        } else {
            v = 1.f / (d + 1e-5f);
        }
        m = fmaxf(m, v);
    }
    __shared__ float sm[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        v = warp_max(v);
        if (threadIdx.x == 0) ws[b] = v;
    }
}
__global__ void disparity_apply_kernel(const float* __restrict__ depth, const float* __restrict__ ws,
                                       __half* __restrict__ out, int B, int HW, int round_f16) {
    GRID_STRIDE(i, (long long)B * HW) {
        const int b = (int)(i / HW);
        const float d = depth[i];
        __half r;
        if (round_f16) {
            const __half dh = __float2half_rn(d);
            const __half s = __float2half_rn(__half2float(dh) + 1e-5f);
            const __half v = __float2half_rn(1.f / __half2float(s));
            r = __float2half_rn(__half2float(v) / ws[b]);
        } else {
            r = __float2half_rn((1.f / (d + 1e-5f)) / ws[b]);
        }
        out[i * 3 + 0] = r;
        out[i * 3 + 1] = r;
        out[i * 3 + 2] = r;
    }
}

// one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                           int rows, int cols, float scale_log2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const __half* xr = x + row * cols;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, __half2float(xr[c]));
    m = warp_max(m) * scale_log2;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += exp2f(__half2float(xr[c]) * scale_log2 - m);
    const float inv = 1.f / warp_sum(s);
    for (int c = lane; c < cols; c += 32)
        y[row * cols + c] = __float2half_rn(exp2f(__half2float(xr[c]) * scale_log2 - m) * inv);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int gcb_silu_fwd(const void* x, void* y, long long n, void* stream) {
    GCB_CHECK_ARG(x && y && n >= 0, "bad args");
    if (n == 0) return GCB_OK;
    silu_kernel<<<ew_blocks(n / 8 + 1), 256, 0, ST>>>((const __half*)x, (__half*)y, n);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_add_fwd(const void* a, const void* b, void* y, long long n, float alpha, float beta, void* stream) {
    GCB_CHECK_ARG(a && b && y && n >= 0, "bad args");
    if (n == 0) return GCB_OK;
    add_kernel<<<ew_blocks(n / 8 + 1), 256, 0, ST>>>((const __half*)a, (const __half*)b, (__half*)y, n, alpha, beta);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_geglu_fwd(const void* x, void* y, int M, int C, void* stream) {
    GCB_CHECK_ARG(x && y && C % 8 == 0, "GEGLU needs C %% 8 == 0");
    geglu_kernel<<<ew_blocks((long long)M * C / 8), 256, 0, ST>>>((const __half*)x, (__half*)y, M, C);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_upsample_nearest2x_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream) {
    GCB_CHECK_ARG(x && y && C % 8 == 0, "upsample needs C %% 8 == 0");
    upsample2x_kernel<<<ew_blocks((long long)B * H * W * 4 * C / 8), 256, 0, ST>>>((const __half*)x, (__half*)y, B, H, W,
                                                                                  C);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_timestep_embedding(const float* timesteps, int B, int dim, void* y, void* stream) {
    GCB_CHECK_ARG(timesteps && y && dim % 2 == 0, "bad args");
    temb_kernel<<<ew_blocks((long long)B * dim / 2), 256, 0, ST>>>(timesteps, (__half*)y, B, dim);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_nchw_to_nhwc_f16(const void* x, void* y, int B, int C, int H, int W, void* stream) {
    GCB_CHECK_ARG(x && y, "null pointer");
    nchw_to_nhwc_kernel<<<ew_blocks((long long)B * C * H * W), 256, 0, ST>>>((const __half*)x, (__half*)y, B, C, H * W);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
extern "C" int gcb_nhwc_to_nchw_f16(const void* x, void* y, int B, int C, int H, int W, void* stream) {
    GCB_CHECK_ARG(x && y, "null pointer");
    nhwc_to_nchw_kernel<<<ew_blocks((long long)B * C * H * W), 256, 0, ST>>>((const __half*)x, (__half*)y, B, C, H * W);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_transpose_f16(const void* x, void* y, int batch, int rows, int cols, void* stream) {
    GCB_CHECK_ARG(x && y && batch > 0 && batch < 65536, "bad args");
    dim3 grid(gcb_cdiv(cols, 32), gcb_cdiv(rows, 32), batch);
    transpose_kernel<<<grid, dim3(32, 8), 0, ST>>>((const __half*)x, (__half*)y, rows, cols);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_out, long long n,
                                 float guidance, const float* coef, void* stream) {
    GCB_CHECK_ARG(eps_uncond && x && x_out && coef, "null pointer");
    cfg_ddim_kernel<<<ew_blocks(n), 256, 0, ST>>>((const __half*)eps_uncond, (const __half*)eps_cond, (const __half*)x,
                                                  (__half*)x_out, n, guidance, coef);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_postprocess_composite(const void* img, const float* mask, const void* unedited_f16, float* out, int B,
                                         int H, int W, void* stream) {
    GCB_CHECK_ARG(img && out && (!mask || unedited_f16), "bad args");
    const long long npix = (long long)B * H * W;
    postprocess_kernel<<<ew_blocks(npix * 3), 256, 0, ST>>>((const __half*)img, mask, (const __half*)unedited_f16, out,
                                                            npix);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_depth_to_disparity(const float* depth, void* disp_f16, float* workspace, int B, int HW,
                                      int round_f16_first, void* stream) {
    GCB_CHECK_ARG(depth && disp_f16 && workspace, "null pointer");
    disparity_max_kernel<<<B, 1024, 0, ST>>>(depth, workspace, HW, round_f16_first);
    GCB_LAUNCH_CHECK();
    disparity_apply_kernel<<<ew_blocks((long long)B * HW), 256, 0, ST>>>(depth, workspace, (__half*)disp_f16, B, HW,
                                                                        round_f16_first);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_softmax_rows_fwd(const void* x, void* y, int rows, int cols, float scale, void* stream) {
    GCB_CHECK_ARG(x && y, "null pointer");
    softmax_rows_kernel<<<gcb_cdiv(rows, 8), 256, 0, ST>>>((const __half*)x, (__half*)y, rows, cols,
                                                           scale * 1.4426950408889634f);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
