// GroupNorm (+SiLU, + channel-concat of two inputs) and LayerNorm over channels-last fp16 tensors, fp32 statistics.
// Replaces torch.nn.GroupNorm/F.silu in ResnetBlock2D / Transformer2DModel.norm / conv_norm_out and
// torch.nn.LayerNorm in BasicTransformerBlock (diffusers 0.26.0, invoked from gc_pipeline.py:142-145, 209-219).
// Both are HBM-bound: GroupNorm = 2 reads + 1 write of the tensor (stats pass is L2-resident for the second read
// at UNet sizes), LayerNorm = 1 read + 1 write.  Deterministic: partial sums are reduced in a fixed order.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int GN_MAX_GROUPS = 32;
constexpr int GN_MAX_CHUNKS = 256;

__device__ __forceinline__ const __half* gn_src(const __half* x1, const __half* x2, int C1, int C2, long long pix,
                                                int c) {
    return c < C1 ? x1 + pix * C1 + c : x2 + pix * C2 + (c - C1);
}

// partial[b][chunk][g][2] = (sum, sumsq) over the chunk's pixels.
// Thread (row r, column v) owns the 8 channels [8v, 8v+8) of pixels r, r+R, r+2R, ... of the chunk: 16-byte coalesced
// loads, four (sum, sumsq) pairs in registers (groups always own whole channel pairs).  The block reduction walks the
// [R][C/2] cells in a fixed order: deterministic, no atomics.
constexpr int GN_THREADS = 320;
constexpr int GN_MAX_PAIRS = 1280;  // C <= 2560

__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2,
                                                              float* __restrict__ partial, int HW, int C1, int C2,
                                                              int groups, int pix_per_chunk) {
    __shared__ float s_sum[GN_MAX_PAIRS], s_sq[GN_MAX_PAIRS];
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const int C = C1 + C2, cv = C / 8, cpg = C / groups;
    const int R = cv >= GN_THREADS ? 1 : GN_THREADS / cv;  // pixel rows handled per pass
    const int p0 = chunk * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
    for (int i = tid; i < C / 2; i += GN_THREADS) {
        s_sum[i] = 0.f;
        s_sq[i] = 0.f;
    }
    __syncthreads();
    // R passes over smem (one row of threads at a time) keep the accumulation order fixed
    for (int v0 = 0; v0 < cv; v0 += GN_THREADS) {
        const int r = cv >= GN_THREADS ? 0 : tid / cv;
        const int v = cv >= GN_THREADS ? v0 + tid : tid % cv;
        float sum[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
        const bool active = r < R && v < cv;
        if (active) {
            const int c = v * 8;
            for (int p = p0 + r; p < p1; p += R) {
                const long long pix = (long long)b * HW + p;
                const uint4 raw = *reinterpret_cast<const uint4*>(gn_src(x1, x2, C1, C2, pix, c));
                const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 f = unpack_half2(w[t]);
                    sum[t] += f.x + f.y;
                    sq[t] += f.x * f.x + f.y * f.y;
                }
            }
        }
        for (int rr = 0; rr < R; ++rr) {
            if (active && r == rr) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    s_sum[v * 4 + t] += sum[t];
                    s_sq[v * 4 + t] += sq[t];
                }
            }
            __syncthreads();
        }
    }
    if (tid < groups) {
        const int c0 = tid * cpg / 2, c1 = (tid + 1) * cpg / 2;
        float sum = 0.f, sq = 0.f;
        for (int col = c0; col < c1; ++col) {
            sum += s_sum[col];
            sq += s_sq[col];
        }
        float* o = partial + (((long long)b * gridDim.x + chunk) * groups + tid) * 2;
        o[0] = sum;
        o[1] = sq;
    }
}

// y = act(x * a[c] + b[c]) with a = rstd * gamma, b = beta - mean * a.
// Prologue: eight threads per group add up the chunk partials (fixed order: deterministic, independent of the batch).
// Body: thread (pixel lane ty, channel vector tx) keeps the a/b of its 8 channels in REGISTERS and walks the CTA's
// pixels ty, ty + L, ... with 16-byte coalesced loads and stores - no shared-memory reads and no integer division
// per element (round 2: 2 LDS per value + one division per 16 bytes held the kernel at 2.8 TB/s).
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x1, const __half* __restrict__ x2,
                                                       const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                                       const float* __restrict__ partial, __half* __restrict__ y, int HW,
                                                       int C1, int C2, int groups, int nchunks, int pix_per_cta,
                                                       float eps, int silu) {
    __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int C = C1 + C2, cpg = C / groups;
    {
        const int g = tid >> 3, part = tid & 7;
        float sum = 0.f, sq = 0.f;
        if (g < groups) {
#pragma unroll 4
            for (int ch = part; ch < nchunks; ch += 8) {
                const float2 pp = *reinterpret_cast<const float2*>(partial + (((long long)b * nchunks + ch) * groups + g) * 2);
                sum += pp.x;
                sq += pp.y;
            }
        }
#pragma unroll
        for (int o = 4; o >= 1; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (g < groups && part == 0) {
            const float n = (float)HW * (float)cpg;
            const float mean = sum / n;
            const float var = fmaxf(sq / n - mean * mean, 0.f);
            s_mean[g] = mean;
            s_rstd[g] = rsqrtf(var + eps);
        }
    }
    __syncthreads();
    const int cv = C / 8;
    const int cvb = cv < 256 ? cv : 256;   // channel vectors per pass (C = 2560: two passes)
    const int nl = 256 / cvb;              // pixel lanes
    const int tx = tid % cvb, ty = tid / cvb;
    if (ty >= nl) return;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    for (int v = tx; v < cv; v += cvb) {
        const int c = v * 8;
        float a[8], bb[8];
        {
            const uint4 gr = *reinterpret_cast<const uint4*>(gamma + c), br = *reinterpret_cast<const uint4*>(beta + c);
            const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w}, bw[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 gf = unpack_half2(gw[t]), bf = unpack_half2(bw[t]);
                const int g0 = (c + 2 * t) / cpg;   // groups own whole channel pairs
                a[2 * t] = s_rstd[g0] * gf.x;
                a[2 * t + 1] = s_rstd[g0] * gf.y;
                bb[2 * t] = bf.x - s_mean[g0] * a[2 * t];
                bb[2 * t + 1] = bf.y - s_mean[g0] * a[2 * t + 1];
            }
        }
        const __half* src = c < C1 ? x1 + ((long long)b * HW) * C1 + c : x2 + ((long long)b * HW) * C2 + (c - C1);
        const int ld = c < C1 ? C1 : C2;
        __half* dst = y + ((long long)b * HW) * C + c;
#pragma unroll 4
        for (int pix = p0 + ty; pix < p1; pix += nl) {
            const uint4 raw = *reinterpret_cast<const uint4*>(src + (long long)pix * ld);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            uint32_t o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_half2(w[t]);
                float r0 = f.x * a[2 * t] + bb[2 * t];
                float r1 = f.y * a[2 * t + 1] + bb[2 * t + 1];
                if (silu) {
                    r0 = silu_f(r0);
                    r1 = silu_f(r1);
                }
                o[t] = pack_half2(r0, r1);
            }
            *reinterpret_cast<uint4*>(dst + (long long)pix * C) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// one warp per row, row held in registers (C <= 8*32*MAXV)
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                                        const __half* __restrict__ beta, __half* __restrict__ y, int M,
                                                        int C, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= M) return;
    const int cv = C / 8;
    uint4 raw[MAXV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < cv) {
            raw[i] = *reinterpret_cast<const uint4*>(x + row * C + v * 8);
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_half2(w[t]);
                sum += f.x + f.y;
            }
        }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < cv) {
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_half2(w[t]);
                sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < cv) {
            const uint4 gr = *reinterpret_cast<const uint4*>(gamma + v * 8);
            const uint4 br = *reinterpret_cast<const uint4*>(beta + v * 8);
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
            const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w};
            const uint32_t bw[4] = {br.x, br.y, br.z, br.w};
            uint32_t o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_half2(w[t]), gg = unpack_half2(gw[t]), bb = unpack_half2(bw[t]);
                o[t] = pack_half2((f.x - mean) * rstd * gg.x + bb.x, (f.y - mean) * rstd * gg.y + bb.y);
            }
            *reinterpret_cast<uint4*>(y + row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace

extern "C" size_t gcb_groupnorm_workspace_bytes(int B, int groups) {
    return (size_t)B * GN_MAX_CHUNKS * groups * 2 * sizeof(float);
}

extern "C" int gcb_groupnorm_nhwc_fwd(const void* x1, const void* x2, const void* gamma, const void* beta, void* y,
                                      int B, int HW, int C1, int C2, int groups, float eps, int silu, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    const int C = C1 + C2;
    GCB_CHECK_ARG(x1 && gamma && beta && y && workspace, "null pointer");
    GCB_CHECK_ARG(C2 == 0 || x2, "x2 is NULL but C2=%d", C2);
    GCB_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "bad groups=%d for C=%d", groups, C);
    GCB_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && (C / groups) % 2 == 0 && C <= 2560, "channel counts C1=%d C2=%d unsupported", C1, C2);
    if (workspace_bytes < gcb_groupnorm_workspace_bytes(B, groups)) {
        gcb_set_error("groupnorm workspace too small");
        return GCB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // chunking depends on HW only (never on B) so that a row's statistics are bit-identical in any batch
    int ppc = HW <= 16384 ? gcb_cdiv(HW, 64) : gcb_cdiv(HW, GN_MAX_CHUNKS);
    if (ppc < 16) ppc = 16;
    const int nchunks = gcb_cdiv(HW, ppc);
    gn_stats_kernel<<<dim3(nchunks, B), GN_THREADS, 0, st>>>((const __half*)x1, (const __half*)x2, (float*)workspace, HW, C1,
                                                      C2, groups, ppc);
    GCB_LAUNCH_CHECK();
    // apply: 64 KB of the tensor per CTA
    int pix_per_cta = 32768 / C;
    if (pix_per_cta < 8) pix_per_cta = 8;
    gn_apply_kernel<<<dim3(gcb_cdiv(HW, pix_per_cta), B), 256, 0, st>>>(
        (const __half*)x1, (const __half*)x2, (const __half*)gamma, (const __half*)beta, (const float*)workspace,
        (__half*)y, HW, C1, C2, groups, nchunks, pix_per_cta, eps, silu);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

// C = 40 L (320: L = 8, 640: L = 16): L lanes per row, five 16-byte vectors per lane, 32 / L rows per warp - every lane
// carries the same load (one warp per 640-byte row left 24 lanes with a single vector: 3.7 TB/s at C = 320).
template <int L>
__global__ void __launch_bounds__(256) layernorm5_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                                         const __half* __restrict__ beta, __half* __restrict__ y, int M,
                                                         float eps) {
    constexpr int C = 40 * L, RPW = 32 / L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / L, j = lane % L;
    const long long row = ((long long)blockIdx.x * 8 + warp) * RPW + sub;
    const bool ok = row < M;
    uint4 raw[5];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        raw[i] = ok ? *reinterpret_cast<const uint4*>(x + row * C + (j + i * L) * 8) : make_uint4(0, 0, 0, 0);
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_half2(w[t]);
            sum += f.x + f.y;
        }
    }
#pragma unroll
    for (int o = L / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_half2(w[t]);
            sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
        }
    }
#pragma unroll
    for (int o = L / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
    if (!ok) return;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int v = j + i * L;
        const uint4 gr = *reinterpret_cast<const uint4*>(gamma + v * 8);
        const uint4 br = *reinterpret_cast<const uint4*>(beta + v * 8);
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w};
        const uint32_t bw[4] = {br.x, br.y, br.z, br.w};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 f = unpack_half2(w[t]), g = unpack_half2(gw[t]), b = unpack_half2(bw[t]);
            o[t] = pack_half2((f.x - mean) * rstd * g.x + b.x, (f.y - mean) * rstd * g.y + b.y);
        }
        *reinterpret_cast<uint4*>(y + row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

extern "C" int gcb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int M, int C, float eps,
                                 void* stream) {
    GCB_CHECK_ARG(x && gamma && beta && y, "null pointer");
    GCB_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * 8, "LayerNorm C=%d unsupported (multiple of 8, <= 2048)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = gcb_cdiv(M, 8);
    if (C == 320) {
        layernorm5_kernel<8><<<gcb_cdiv(M, 32), 256, 0, st>>>((const __half*)x, (const __half*)gamma, (const __half*)beta,
                                                              (__half*)y, M, eps);
    } else if (C == 640) {
        layernorm5_kernel<16><<<gcb_cdiv(M, 16), 256, 0, st>>>((const __half*)x, (const __half*)gamma, (const __half*)beta,
                                                               (__half*)y, M, eps);
    } else if (C <= 512)
        layernorm_kernel<2><<<blocks, 256, 0, st>>>((const __half*)x, (const __half*)gamma, (const __half*)beta,
                                                    (__half*)y, M, C, eps);
    else if (C <= 1280)
        layernorm_kernel<5><<<blocks, 256, 0, st>>>((const __half*)x, (const __half*)gamma, (const __half*)beta,
                                                    (__half*)y, M, C, eps);
    else
        layernorm_kernel<8><<<blocks, 256, 0, st>>>((const __half*)x, (const __half*)gamma, (const __half*)beta,
                                                    (__half*)y, M, C, eps);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
