// CLIP text-encoder pieces that the UNet kernels do not already cover (the reference reaches this
// model through diffusers' encode_prompt inside `self.pipe(prompt=..., negative_prompt=...)`, gc_pipeline.py:142-145
// and :209-219): token + position embedding gather, causal self-attention for one short sequence (T <= 128, d = 64),
// quick-GELU.  The projections and LayerNorms run on gemm_tc.cu / norm.cu.  This model runs once per prompt pair
// (2 x 77 tokens): the kernels are written for simplicity and exactness of the fp32 softmax, not for throughput.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int CA_MAX_T = 128;  // sequence length limit (CLIP: 77)
constexpr int CA_D = 64;       // head dim (CLIP ViT-L/14 text tower: 768 / 12)

// out[b,t,:] = tok[ids[b,t],:] + pos[t,:]      one CTA per token, 8 halves per thread
__global__ void __launch_bounds__(128)
embed_tokens_kernel(const int32_t* __restrict__ ids, const __half* __restrict__ tok, const __half* __restrict__ pos,
                    __half* __restrict__ out, int T, int C, int vocab) {
    const int bt = blockIdx.x;
    const int t = bt % T;
    int id = ids[bt];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const uint4* a = reinterpret_cast<const uint4*>(tok + (long long)id * C);
    const uint4* b = reinterpret_cast<const uint4*>(pos + (long long)t * C);
    uint4* o = reinterpret_cast<uint4*>(out + (long long)bt * C);
    for (int i = threadIdx.x; i < C / 8; i += blockDim.x) {
        uint4 va = a[i], vb = b[i], vo;
        const __half2* ha = reinterpret_cast<const __half2*>(&va);
        const __half2* hb = reinterpret_cast<const __half2*>(&vb);
        __half2* ho = reinterpret_cast<__half2*>(&vo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __half22float2(ha[j]), fb = __half22float2(hb[j]);
            ho[j] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
        }
        o[i] = vo;
    }
}

// y = x * sigmoid(1.702 x), 8 halves per thread
__global__ void __launch_bounds__(256)
quick_gelu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n8) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 v = reinterpret_cast<const uint4*>(x)[i], o;
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(hv[j]);
            ho[j] = __floats2half2_rn(f.x / (1.f + __expf(-1.702f * f.x)), f.y / (1.f + __expf(-1.702f * f.y)));
        }
        reinterpret_cast<uint4*>(y)[i] = o;
    }
}

// Causal self-attention of one (batch row, head): thread i owns query row i; K and V of the head sit in shared memory
// (every thread reads the same K/V row at the same time: a broadcast, no bank conflicts); online softmax in fp32.
__global__ void __launch_bounds__(CA_MAX_T)
attn_causal_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v, int ld,
                   __half* __restrict__ out, int ld_out, int T, float scale_log2) {
    __shared__ __align__(16) __half sk[CA_MAX_T][CA_D];
    __shared__ __align__(16) __half sv[CA_MAX_T][CA_D];
    const int head = blockIdx.x, b = blockIdx.y;
    const long long row0 = (long long)b * T;
    // stage K and V: T rows x 8 uint4 each
    for (int i = threadIdx.x; i < T * (CA_D / 8); i += blockDim.x) {
        const int r = i / (CA_D / 8), c = i % (CA_D / 8);
        reinterpret_cast<uint4*>(&sk[r][0])[c] = reinterpret_cast<const uint4*>(k + (row0 + r) * ld + head * CA_D)[c];
        reinterpret_cast<uint4*>(&sv[r][0])[c] = reinterpret_cast<const uint4*>(v + (row0 + r) * ld + head * CA_D)[c];
    }
    __syncthreads();
    const int i = threadIdx.x;
    if (i >= T) return;
    float qr[CA_D], acc[CA_D];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q + (row0 + i) * ld + head * CA_D);
#pragma unroll
        for (int c = 0; c < CA_D / 8; ++c) {
            const uint4 u = qp[c];
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                qr[c * 8 + 2 * j] = f.x * scale_log2;
                qr[c * 8 + 2 * j + 1] = f.y * scale_log2;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CA_D; ++c) acc[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j <= i; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CA_D; c += 2) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sk[j][c]));
            s = fmaf(qr[c], f.x, s);
            s = fmaf(qr[c + 1], f.y, s);
        }
        const float mn = fmaxf(m, s);
        const float alpha = exp2f(m - mn);  // m = -inf on the first key: exp2f(-inf) = 0
        const float p = exp2f(s - mn);
        l = l * alpha + p;
#pragma unroll
        for (int c = 0; c < CA_D; c += 2) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sv[j][c]));
            acc[c] = fmaf(acc[c], alpha, p * f.x);
            acc[c + 1] = fmaf(acc[c + 1], alpha, p * f.y);
        }
        m = mn;
    }
    const float inv = 1.f / l;
    uint4* op = reinterpret_cast<uint4*>(out + (row0 + i) * ld_out + head * CA_D);
#pragma unroll
    for (int c = 0; c < CA_D / 8; ++c) {
        uint4 o;
        __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[c * 8 + 2 * j] * inv, acc[c * 8 + 2 * j + 1] * inv);
        op[c] = o;
    }
}

}  // namespace

extern "C" int gcb_embed_tokens_f16(const int32_t* ids, const void* tok_emb, const void* pos_emb, void* out, int B, int T,
                                    int C, int vocab, void* stream) {
    GCB_CHECK_ARG(ids && tok_emb && pos_emb && out, "embed_tokens: null pointer");
    GCB_CHECK_ARG(B > 0 && T > 0 && C > 0 && C % 8 == 0 && vocab > 0, "embed_tokens: bad shape B=%d T=%d C=%d", B, T, C);
    embed_tokens_kernel<<<B * T, 128, 0, (cudaStream_t)stream>>>(ids, (const __half*)tok_emb, (const __half*)pos_emb,
                                                               (__half*)out, T, C, vocab);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_quick_gelu_fwd(const void* x, void* y, long long n, void* stream) {
    GCB_CHECK_ARG(x && y, "quick_gelu: null pointer");
    GCB_CHECK_ARG(n >= 0 && n % 8 == 0, "quick_gelu: n=%lld must be a multiple of 8", n);
    if (n == 0) return GCB_OK;
    const long long n8 = n / 8;
    long long blocks = (n8 + 255) / 256;
    const long long cap = (long long)gcb_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    quick_gelu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, n8);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_attn_causal_fwd(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int B,
                                   int T, int heads, int d, float scale, void* stream) {
    GCB_CHECK_ARG(q && k && v && out, "attn_causal: null pointer");
    GCB_CHECK_ARG(d == CA_D, "attn_causal: head dim %d not built (64 only)", d);
    GCB_CHECK_ARG(T > 0 && T <= CA_MAX_T, "attn_causal: sequence length %d out of range (1..%d)", T, CA_MAX_T);
    GCB_CHECK_ARG(B > 0 && heads > 0 && ld_qkv % 8 == 0 && ld_out % 8 == 0 && ld_qkv >= heads * d && ld_out >= heads * d,
                  "attn_causal: bad strides ld_qkv=%d ld_out=%d", ld_qkv, ld_out);
    GCB_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0,
                  "attn_causal: pointers must be 16-byte aligned");
    attn_causal_kernel<<<dim3(heads, B), CA_MAX_T, 0, (cudaStream_t)stream>>>(
        (const __half*)q, (const __half*)k, (const __half*)v, ld_qkv, (__half*)out, ld_out, T,
        scale * 1.4426950408889634f);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
