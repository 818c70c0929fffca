// Multi-source attention, register-resident flash formulation on mma.sync (bring-up / cross-check path and the
// path used for shapes the tcgen05 kernel does not cover: text cross-attention Nk=77, d=160 levels).
//
//   out[b, i, h, :] = sum_s w_s * softmax_j( q[b,i,h,:] . K_s[j,h,:] * scale ) V_s[j,h,:]
//
// which is CrossViewAttnProcessor's 5 passes (gaussctrl/utils.py:88-117) with s = {self, ref0..ref3},
// w = {c, (1-c)/4 ...}; the softmax of every source is independent, exactly as compute_attn (utils.py:25-37).
// One CTA = 64 query rows of one (batch row, head); 4 warps x 16 rows; keys streamed in BN-key tiles through a
// two-stage cp.async ring; probabilities never leave registers (the reference materialises [B*8, N, N] fp16).
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_SRC = 8;
constexpr int BMQ = 64;

struct AttnParams {
    const __half* q;
    const __half* k;
    const __half* v;
    const __half* k2;
    const __half* v2;
    __half* out;
    int ld_q, ld_kv, ld_kv2, ld_out;
    int B, Nq, Nk, heads, vstride;
    int n_src_total;         // stride of src_index per batch row
    int n_act;               // active (non-zero weight) sources
    int src_id[MAX_SRC];     // original source slot of each active source
    float weight[MAX_SRC];
    const int32_t* src_index;
    float scale_log2;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int D, int BN>
__global__ void __launch_bounds__(128) attn_mma_kernel(const AttnParams p) {
    constexpr int DP = (D + 15) / 16 * 16;  // K extent of Q K^T, zero padded
    constexpr int LDS = DP + 8;             // smem row stride (halves): conflict-free fragment loads
    constexpr int KS = DP / 16;             // k-steps of Q K^T
    constexpr int NT_S = BN / 8;            // n-tiles of S
    constexpr int NT_O = D / 8;             // n-tiles of O
    constexpr int CH = D / 8;               // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __half* Qs = reinterpret_cast<__half*>(smem_raw);
    __half* Ks = Qs + BMQ * LDS;            // [2][BN][LDS]
    __half* Vs = Ks + 2 * BN * LDS;         // [2][BN][LDS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BMQ;
    const int g = lane >> 2, t4 = lane & 3;

    // zero the pad columns [D, LDS) of Q and K once (cp.async only ever writes columns [0, D))
    if (LDS > D) {
        constexpr int PADW = LDS - D;
        for (int i = tid; i < (BMQ + 2 * BN) * PADW; i += 128) {
            const int r = i / PADW, c = D + i % PADW;
            Qs[r * LDS + c] = __float2half(0.f);  // Qs and Ks are contiguous: rows >= BMQ land in Ks
        }
    }
    // Q tile
    {
        const __half* qg = p.q + ((long long)b * p.Nq + q0) * p.ld_q + head * D;
        for (int i = tid; i < BMQ * CH; i += 128) {
            const int r = i / CH, c = i % CH;
            cp_async16(smem_u32(&Qs[r * LDS + c * 8]), qg + (long long)r * p.ld_q + c * 8, q0 + r < p.Nq);
        }
        cp_async_commit();
    }
    float oacc[NT_O][4];
#pragma unroll
    for (int j = 0; j < NT_O; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) oacc[j][t] = 0.f;
    const int nkt = (p.Nk + BN - 1) / BN;
    bool q_loaded = false;
    uint32_t qf[KS][4];

    for (int si = 0; si < p.n_act; ++si) {
        const int sidx = p.src_index[b * p.n_src_total + p.src_id[si]];
        const __half* kg;
        const __half* vg;
        int ldkv;
        if (sidx >= 0) {
            kg = p.k + (long long)sidx * p.Nk * p.ld_kv + head * D;
            vg = p.v + (long long)sidx * p.Nk * p.ld_kv + head * p.vstride;
            ldkv = p.ld_kv;
        } else {
            const long long r2 = -(long long)(sidx + 1);
            kg = p.k2 + r2 * p.Nk * p.ld_kv2 + head * D;
            vg = p.v2 + r2 * p.Nk * p.ld_kv2 + head * p.vstride;
            ldkv = p.ld_kv2;
        }
        auto load_kv = [&](int kt, int stage) {
            const int k0 = kt * BN;
            __half* kd = Ks + stage * BN * LDS;
            __half* vd = Vs + stage * BN * LDS;
            for (int i = tid; i < BN * CH; i += 128) {
                const int r = i / CH, c = i % CH;
                const bool ok = k0 + r < p.Nk;
                const long long off = (long long)(ok ? k0 + r : 0) * ldkv + c * 8;
                cp_async16(smem_u32(&kd[r * LDS + c * 8]), kg + off, ok);
                cp_async16(smem_u32(&vd[r * LDS + c * 8]), vg + off, ok);
            }
            cp_async_commit();
        };
        load_kv(0, 0);
        float o[NT_O][4];
#pragma unroll
        for (int j = 0; j < NT_O; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t) o[j][t] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kt = 0; kt < nkt; ++kt) {
            const int stage = kt & 1;
            if (kt + 1 < nkt) {
                load_kv(kt + 1, stage ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            if (!q_loaded) {
                q_loaded = true;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const __half* qb = Qs + (warp * 16 + g) * LDS + ks * 16 + t4 * 2;
                    qf[ks][0] = *reinterpret_cast<const uint32_t*>(qb);
                    qf[ks][1] = *reinterpret_cast<const uint32_t*>(qb + 8 * LDS);
                    qf[ks][2] = *reinterpret_cast<const uint32_t*>(qb + 8);
                    qf[ks][3] = *reinterpret_cast<const uint32_t*>(qb + 8 * LDS + 8);
                }
            }
            const __half* kb = Ks + stage * BN * LDS;
            const __half* vb = Vs + stage * BN * LDS;
            float s[NT_S][4];
#pragma unroll
            for (int j = 0; j < NT_S; ++j) {
#pragma unroll
                for (int t = 0; t < 4; ++t) s[j][t] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const __half* kp = kb + (j * 8 + g) * LDS + ks * 16 + t4 * 2;
                    mma16816(s[j], qf[ks], *reinterpret_cast<const uint32_t*>(kp),
                             *reinterpret_cast<const uint32_t*>(kp + 8));
                }
            }
            // scale into the exp2 domain, mask the key tail
            const int kbase = kt * BN;
            const bool tail = kbase + BN > p.Nk;
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < NT_S; ++j) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v = s[j][t] * p.scale_log2;
                    if (tail && (kbase + j * 8 + t4 * 2 + (t & 1)) >= p.Nk) v = -INFINITY;
                    s[j][t] = v;
                }
                mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
                mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
            const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);  // first tile: exp2(-inf) = 0
            m0 = mn0;
            m1 = mn1;
            float rs0 = 0.f, rs1 = 0.f;
            uint32_t pf[NT_S][2];
#pragma unroll
            for (int j = 0; j < NT_S; ++j) {
                const float p0 = exp2f(s[j][0] - mn0), p1 = exp2f(s[j][1] - mn0);
                const float p2 = exp2f(s[j][2] - mn1), p3 = exp2f(s[j][3] - mn1);
                // the row sum uses the fp16-rounded probabilities that the PV product sees
                const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                rs0 += f01.x + f01.y;
                rs1 += f23.x + f23.y;
                pf[j][0] = *reinterpret_cast<const uint32_t*>(&h01);
                pf[j][1] = *reinterpret_cast<const uint32_t*>(&h23);
            }
            l0 = l0 * c0 + rs0;
            l1 = l1 * c1 + rs1;
#pragma unroll
            for (int j = 0; j < NT_O; ++j) {
                o[j][0] *= c0;
                o[j][1] *= c0;
                o[j][2] *= c1;
                o[j][3] *= c1;
            }
            // O += P V : A fragments of P come straight from the S accumulators (two n-tiles = one k-step)
#pragma unroll
            for (int kk = 0; kk < BN / 16; ++kk) {
                const uint32_t a[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
                const uint32_t vrow = smem_u32(vb + (kk * 16 + (lane & 15)) * LDS);
#pragma unroll
                for (int j = 0; j < NT_O; ++j) {
                    uint32_t b0, b1;
                    ldmatrix_x2_trans(b0, b1, vrow + j * 16);
                    mma16816(o[j], a, b0, b1);
                }
            }
            __syncthreads();  // everyone done with this stage before it is refilled
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float w0 = p.weight[si] / l0, w1 = p.weight[si] / l1;
#pragma unroll
        for (int j = 0; j < NT_O; ++j) {
            oacc[j][0] += o[j][0] * w0;
            oacc[j][1] += o[j][1] * w0;
            oacc[j][2] += o[j][2] * w1;
            oacc[j][3] += o[j][3] * w1;
        }
    }
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __half* og = p.out + (long long)b * p.Nq * p.ld_out + head * D + t4 * 2;
#pragma unroll
    for (int j = 0; j < NT_O; ++j) {
        if (r0 < p.Nq)
            *reinterpret_cast<__half2*>(og + (long long)r0 * p.ld_out + j * 8) = __floats2half2_rn(oacc[j][0], oacc[j][1]);
        if (r1 < p.Nq)
            *reinterpret_cast<__half2*>(og + (long long)r1 * p.ld_out + j * 8) = __floats2half2_rn(oacc[j][2], oacc[j][3]);
    }
}

template <int D, int BN>
int launch(const AttnParams& p, cudaStream_t st) {
    constexpr int DP = (D + 15) / 16 * 16, LDS = DP + 8;
    const size_t smem = (size_t)(BMQ + 4 * BN) * LDS * 2;
    static unsigned long long configured = 0;   // one bit per device ordinal
    if (gcb_first_use_on_device(configured)) {
        GCB_CUDA(cudaFuncSetAttribute(attn_mma_kernel<D, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(gcb_cdiv(p.Nq, BMQ), p.heads, p.B);
    attn_mma_kernel<D, BN><<<grid, 128, smem, st>>>(p);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

}  // namespace

int gcb_attn_mma_launch(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2,
                        const void* v2, int ld_kv2, void* out, int ld_out, int B, int Nq, int Nk, int heads, int d,
                        int vstride, int n_src, const int32_t* src_index, const float* h_src_weight, float scale,
                        cudaStream_t stream) {
    AttnParams p;
    p.q = (const __half*)q;
    p.k = (const __half*)k;
    p.v = (const __half*)v;
    p.k2 = (const __half*)k2;
    p.v2 = (const __half*)v2;
    p.out = (__half*)out;
    p.ld_q = ld_q;
    p.ld_kv = ld_kv;
    p.ld_kv2 = ld_kv2;
    p.ld_out = ld_out;
    p.B = B;
    p.Nq = Nq;
    p.Nk = Nk;
    p.heads = heads;
    p.vstride = vstride;
    p.n_src_total = n_src;
    p.n_act = 0;
    for (int s = 0; s < n_src; ++s)
        if (h_src_weight[s] != 0.f) {
            p.src_id[p.n_act] = s;
            p.weight[p.n_act] = h_src_weight[s];
            ++p.n_act;
        }
    p.src_index = src_index;
    p.scale_log2 = scale * 1.4426950408889634f;
    switch (d) {
        case 40: return launch<40, 64>(p, stream);
        case 64: return launch<64, 64>(p, stream);
        case 80: return launch<80, 64>(p, stream);
        case 160: return launch<160, 32>(p, stream);
        default:
            gcb_set_error("attention head dim %d not built (40, 64, 80, 160)", d);
            return GCB_ERR_UNSUPPORTED;
    }
}

int gcb_attn_tc_supported(int Nq, int Nk, int heads, int d);
int gcb_attn_tc_launch(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2, const void* v2,
                       int ld_kv2, void* out, int ld_out, int B, int Nq, int Nk, int heads, int d, int vstride, int n_src,
                       const int32_t* src_index, const float* h_src_weight, float scale, cudaStream_t stream);

extern "C" int gcb_attn_multi_fwd(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2,
                                  const void* v2, int ld_kv2, void* out, int ld_out, int B, int Nq, int Nk, int heads,
                                  int d, int v_head_stride, int n_src, const int32_t* src_index,
                                  const float* h_src_weight, float scale, int impl, void* stream) {
    GCB_CHECK_ARG(q && k && v && out && src_index && h_src_weight, "null pointer");
    GCB_CHECK_ARG(B > 0 && Nq > 0 && Nk > 0 && heads > 0 && heads < 65536 && B < 65536, "bad shape");
    GCB_CHECK_ARG(n_src >= 1 && n_src <= MAX_SRC, "n_src=%d out of range (1..%d)", n_src, MAX_SRC);
    GCB_CHECK_ARG(ld_q % 8 == 0 && ld_kv % 8 == 0 && ld_out % 2 == 0 && (k2 == nullptr || ld_kv2 % 8 == 0),
                  "row strides must keep 16-byte alignment");
    GCB_CHECK_ARG(d % 8 == 0, "head dim must be a multiple of 8");
    GCB_CHECK_ARG(v_head_stride >= d && v_head_stride % 8 == 0, "v_head_stride=%d must be >= d and a multiple of 8",
                  v_head_stride);
    if (const char* e = getenv("GCB_FORCE_ATTN_IMPL")) impl = atoi(e);
    if (impl == GCB_ATTN_AUTO) impl = gcb_attn_tc_supported(Nq, Nk, heads, d) ? GCB_ATTN_TCGEN05 : GCB_ATTN_MMA_SYNC;
    if (impl == GCB_ATTN_TCGEN05) {
        if (!gcb_attn_tc_supported(Nq, Nk, heads, d)) {
            gcb_set_error("tcgen05 attention does not support Nq=%d Nk=%d d=%d", Nq, Nk, d);
            return GCB_ERR_UNSUPPORTED;
        }
        return gcb_attn_tc_launch(q, ld_q, k, v, ld_kv, k2, v2, ld_kv2, out, ld_out, B, Nq, Nk, heads, d, v_head_stride,
                                  n_src, src_index, h_src_weight, scale, (cudaStream_t)stream);
    }
    GCB_CHECK_ARG(impl == GCB_ATTN_MMA_SYNC, "unknown attention impl %d", impl);
    return gcb_attn_mma_launch(q, ld_q, k, v, ld_kv, k2, v2, ld_kv2, out, ld_out, B, Nq, Nk, heads, d, v_head_stride, n_src,
                               src_index, h_src_weight, scale, (cudaStream_t)stream);
}
