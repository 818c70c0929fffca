// Multi-source (cross-view) attention on the 5th-gen tensor cores for head dims 40 and 80 (the SD1.x 64x64 and 32x32
// levels: ~90 % of the attention FLOPs of the hot path).
//
//   out[b, i, h, :] = sum_s w_s * softmax_j( q[b,i,h,:] . K_s[j,h,:] * scale ) V_s[j,h,:]      (utils.py:25-37, 88-117)
//
// One CTA = NSLOT 128-row query slots of one (batch row, head).  Warp roles (32 * (4 NSLOT + 2) threads):
//   warps 4t..4t+3  : softmax of slot t - one query row per thread, S read from TMEM (tcgen05.ld), exp2 on the
//                     MUFU, row sums in fp32 registers, P written back to TMEM (tcgen05.st) as packed halves: the
//                     A operand of P V.  Probabilities never touch shared memory or HBM.
//   warp 4 NSLOT    : TMA producer (Q once; K and V tiles of BN keys through mbarrier rings)
//   warp 4 NSLOT + 1: TMEM allocator + tcgen05.mma issuer (one elected lane):
//                       S   = Q K^T    M128 x N(BN) x K(DK)   both operands K-major, 128B-swizzled 64-column TMA boxes
//                       O  += P V      M128 x N(NV) x K(BN)   A = P from TMEM, B = V MN-major straight from its row-major tile
// Head dim 40: two slots, BN=128, DK=48 - the Q/K boxes are 64 columns wide; columns 40..47 of Q are zeroed in smem so
//   the neighbouring head's columns that ride along in K contribute nothing; NV=48 (columns 40..47 of O are don't-care).
//   The kernel is bound by the MUFU (exp2: 16/clk/SM); each scheduler's MUFU is fed by the two softmax warps resident on
//   it, one per slot.  Measured alternatives (profiles/r2_attn_ab.md): a third slot with BN=64 (TMEM: 3 x (64 S + 32 P +
//   48 O) = 432 columns) runs at 405 TFLOP/s against 530 - and so does BN=64 with two slots, i.e. the per-tile fixed
//   cost of the smaller tile eats the gain; a four-piece P store 495; issuing a slot's next QK^T after its own PV 442.
// Head dim 80: two slots, BN=64, DK=80 (two boxes: 64 + 16 used columns), NV=80 (two MN atoms of V).
// Online softmax with lazy rescaling (threshold 2^8): the O correction (TMEM load-scale-store) is rare.  Sources are
// processed back to back; at the end of each source the slot folds O * w_s / l into an fp32 accumulator (shared
// memory for d=40, TMEM for d=80 - whichever the budget of 512 columns / 227 KB leaves room for).
// ncu history (profiles/): r1a P through shared memory + a separate "P x ones" row-sum MMA, every MMA wrapped in an
// ELECT waterfall loop: 277 TFLOP/s; r1c (this structure): 440-500 TFLOP/s at d=40.
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

#include <type_traits>

namespace {

constexpr int MAX_SRC = 8;
constexpr int BM = 128;  // query rows per slot
constexpr float RESCALE_THRESHOLD = 8.f;
#ifndef GCB_ATTN_LAG
#define GCB_ATTN_LAG 0
#endif
#ifndef GCB_ATTN_SPLIT_LD
#define GCB_ATTN_SPLIT_LD 0
#endif
#ifndef GCB_ATTN_SPLIT_ST
#define GCB_ATTN_SPLIT_ST 1
#endif
// The non-MUFU part of a tile (BN = 128 only; profiles/r1i_attn_source_summary.md), A/B in one process with
// tools/ab_attn_libs.py at B=24, N=4096, d=40 (profiles/r1o_ab_attn_split.txt; outputs bit-identical in all four builds):
//   SPLIT_ST (ON): the first 64 packed probabilities are stored (tcgen05.st) after half of the exponentials; the p_free
//             spin in between also splits the basic block, so ptxas can no longer sink all 64 packs behind the last
//             ex2: 487.5 -> 538.1 TFLOP/s (+10 %).
//   SPLIT_LD (off): the second half of S still loading from TMEM while the row max of the first half is taken:
//             484.7 TFLOP/s alone, 535.7 with SPLIT_ST - no gain, the TMEM-load latency is not what the max phase waits on.
//   SPLIT_ST = 4 (not measured yet - the round's GPU budget ended): four pieces of 16 packed registers; the extra block
//             boundaries come from re-waiting the already completed p_free phase (returns at once, but is a branch).
#ifndef GCB_ATTN_PINGPONG
#define GCB_ATTN_PINGPONG 0
#endif
// MUFU ping-pong (two-slot kernels): the softmax warps of slot 0 and slot 1 that share a scheduler (same TMEM lane
// quarter) pass a token through a pair of named barriers so that only ONE of them is in its exponential phase at a
// time - the other does its non-MUFU work (wait for S, TMEM load, row max, pack, store) meanwhile.  Without it the two
// exponential phases mostly coincide (each then runs at half the MUFU rate) and so do the non-MUFU phases (pipe idle).
constexpr bool ATTN_PINGPONG = GCB_ATTN_PINGPONG != 0;
constexpr bool ATTN_SPLIT_LD = GCB_ATTN_SPLIT_LD != 0, ATTN_SPLIT_ST = GCB_ATTN_SPLIT_ST != 0;
constexpr bool ATTN_SPLIT_ST4 = GCB_ATTN_SPLIT_ST == 4;
constexpr int ATTN_LAG = GCB_ATTN_LAG;  // 0 = off, 1 = slot 0 signals after its row max, 2 = after half of its exponentials

template <int D_>
struct Cfg;
template <>
struct Cfg<40> {
#if defined(GCB_ATTN40_THREE_SLOTS)   // A/B (r2j): three slots, BN = 64 - 405 TFLOP/s against 530 for the default
    static constexpr int D = 40, BN = 64, DK = 48, NV = 48, BOXES = 1, NSLOT = 3, STAGES = 6;
    static constexpr uint32_t TM_S = 0, TM_P = 192, TM_O = 288, TM_O_STRIDE = 48, TM_ACC = 0;
#elif defined(GCB_ATTN40_BN64_TWO)    // A/B (r2j): BN = 64 with two slots - also 405: the tile size costs, not the slots
    static constexpr int D = 40, BN = 64, DK = 48, NV = 48, BOXES = 1, NSLOT = 2, STAGES = 6;
    static constexpr uint32_t TM_S = 0, TM_P = 192, TM_O = 288, TM_O_STRIDE = 48, TM_ACC = 0;
#else
    static constexpr int D = 40, BN = 128, DK = 48, NV = 48, BOXES = 1, NSLOT = 2, STAGES = 4;
    static constexpr uint32_t TM_S = 0, TM_O = 256, TM_O_STRIDE = 64, TM_P = 384, TM_ACC = 0;
#endif
    static constexpr bool ZERO_Q_PAD = true, ACC_IN_TMEM = false;
    // exponentials per 8 evaluated by the FMA-pipe polynomial instead of the MUFU.  Measured on B200 (r1): 3/8 makes
    // the kernel SLOWER (546 -> 485 TFLOP/s): 3-register FFMA/FADD issue at half rate per SM sub-partition, so the
    // ~7-instruction polynomial costs more FMA-pipe time than the MUFU slot it frees.  Kept at 0.
    static constexpr int POLY_OF_8 = 0;
#ifdef GCB_ATTN_EXP_F16X2
    static constexpr bool EXP_F16X2 = true;
#else
    static constexpr bool EXP_F16X2 = false;
#endif
};
template <>
struct Cfg<80> {
    static constexpr int D = 80, BN = 64, DK = 80, NV = 80, BOXES = 2, NSLOT = 2, STAGES = 4;
    static constexpr bool ZERO_Q_PAD = false, ACC_IN_TMEM = true;
    static constexpr int POLY_OF_8 = 0;
    static constexpr bool EXP_F16X2 = false;
    static constexpr uint32_t TM_S = 0, TM_O = 128, TM_O_STRIDE = 80, TM_P = 288, TM_ACC = 352;
};

struct AttnTcParams {
    __half* out;
    int ld_out;
    int Nq, Nk, heads;
    int vstride;      // elements between heads in a V row
    int l_from_o;     // V column D holds 1.0: row sums come out of P V (column D of O)
    int n_src_total, n_act;
    int src_id[MAX_SRC];
    float weight[MAX_SRC];
    const int32_t* src_index;
    float scale_log2;
};

template <int D_>
struct __align__(1024) Smem {
    using C = Cfg<D_>;
    static constexpr uint32_t QBOX = BM * 128, KVBOX = C::BN * 128;
    static constexpr int STAGES = C::STAGES, NSLOT = C::NSLOT;
    uint8_t q[NSLOT][C::BOXES][QBOX];
    uint8_t k[STAGES][C::BOXES][KVBOX];
    uint8_t v[STAGES][C::BOXES][KVBOX];
    float acc[C::ACC_IN_TMEM ? 1 : NSLOT][C::ACC_IN_TMEM ? 1 : C::D][C::ACC_IN_TMEM ? 4 : BM];  // [slot][column][row]
    uint64_t q_full, q_ready;
    uint64_t k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
    uint64_t s_full[NSLOT], s_free[NSLOT], p_ready[NSLOT], p_free[NSLOT], o_free[NSLOT];
    uint64_t lag_bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_f32(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax, max rel. error 7.5e-5 - well below the fp16
// rounding of P): used for a fraction of the elements so the MUFU (16 ex2/clk/SM) stops being the limiter at d=40.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float fi = x + 12582912.f;  // 1.5 * 2^23: round(x) lands in the low mantissa bits
    const float f = x - (fi - 12582912.f);  // [-0.5, 0.5]
    float r = fmaf(f, 0.0551716648f, 0.2426111251f);
    r = fmaf(r, f, 0.6932609677f);
    r = fmaf(r, f, 0.9999280572f);
    return __int_as_float(__float_as_int(r) + (__float_as_int(fi) << 23));
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
    uint32_t y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
    return y;
}
// mbarrier arrive whose address carries a (zero) data dependency on `dep`, so that neither nvcc nor ptxas can move it
// ahead of the instruction that produces `dep`:  (dep >> 31) & (~dep >> 31) == 0 for every dep.
__device__ __forceinline__ void lag_arrive(uint32_t bar, uint32_t dep) {
    asm volatile(
        "{\n\t.reg .b32 a, b;\n\t"
        "shr.u32 a, %1, 31;\n\t"
        "not.b32 b, %1;\n\t"
        "shr.u32 b, b, 31;\n\t"
        "and.b32 a, a, b;\n\t"
        "add.u32 a, a, %0;\n\t"
        "mbarrier.arrive.shared::cta.b64 _, [a];\n\t}"
        ::"r"(bar), "r"(dep)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// tcgen05.ld / st 32x32b.x32 straight into / from r[OFF .. OFF+31] (no staging copies)
template <int OFF, int NR>
__device__ __forceinline__ void tmem_ld32_into(uint32_t taddr, uint32_t (&r)[NR]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
          "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
          "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]),
          "=r"(r[OFF + 17]), "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]),
          "=r"(r[OFF + 22]), "=r"(r[OFF + 23]), "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]),
          "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]), "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
        : "r"(taddr)
        : "memory");
}
template <int OFF, int NR>
__device__ __forceinline__ void tmem_st32_from(uint32_t taddr, const uint32_t (&r)[NR]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]), "r"(r[OFF + 3]), "r"(r[OFF + 4]), "r"(r[OFF + 5]),
          "r"(r[OFF + 6]), "r"(r[OFF + 7]), "r"(r[OFF + 8]), "r"(r[OFF + 9]), "r"(r[OFF + 10]), "r"(r[OFF + 11]),
          "r"(r[OFF + 12]), "r"(r[OFF + 13]), "r"(r[OFF + 14]), "r"(r[OFF + 15]), "r"(r[OFF + 16]), "r"(r[OFF + 17]),
          "r"(r[OFF + 18]), "r"(r[OFF + 19]), "r"(r[OFF + 20]), "r"(r[OFF + 21]), "r"(r[OFF + 22]), "r"(r[OFF + 23]),
          "r"(r[OFF + 24]), "r"(r[OFF + 25]), "r"(r[OFF + 26]), "r"(r[OFF + 27]), "r"(r[OFF + 28]), "r"(r[OFF + 29]),
          "r"(r[OFF + 30]), "r"(r[OFF + 31])
        : "memory");
}

template <int OFF, int NR>
__device__ __forceinline__ void tmem_st16_from(uint32_t taddr, const uint32_t (&r)[NR]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]), "r"(r[OFF + 3]), "r"(r[OFF + 4]), "r"(r[OFF + 5]),
          "r"(r[OFF + 6]), "r"(r[OFF + 7]), "r"(r[OFF + 8]), "r"(r[OFF + 9]), "r"(r[OFF + 10]), "r"(r[OFF + 11]),
          "r"(r[OFF + 12]), "r"(r[OFF + 13]), "r"(r[OFF + 14]), "r"(r[OFF + 15])
        : "memory");
}

template <int D_>
__global__ void __launch_bounds__(32 * (4 * Cfg<D_>::NSLOT + 2), 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
               const __grid_constant__ CUtensorMap tmV2, const AttnTcParams p) {
    using C = Cfg<D_>;
    using SM = Smem<D_>;
    constexpr int D = C::D, BN = C::BN, BOXES = C::BOXES, NSLOT = C::NSLOT, STAGES = C::STAGES;
    constexpr int W_TMA = 4 * NSLOT, W_MMA = 4 * NSLOT + 1;
    constexpr int KSTEPS = C::DK / 16, PV_STEPS = BN / 16, PCOLS = BN / 2;
    constexpr uint32_t STAGE_BYTES = BOXES * SM::KVBOX;
    extern __shared__ uint8_t smem_raw[];
    SM& sm = *reinterpret_cast<SM*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int nkt = p.Nk / BN;
    const int T = p.n_act * nkt;  // (source, key tile) pairs, processed in order
    // the last CTA of a (row, head) may own fewer than NSLOT slots (Nq is a multiple of 128, not of NSLOT * 128)
    const int nslot = min(NSLOT, p.Nq / BM - qt * NSLOT);

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&sm.q_full), 1);
        mbar_init(smem_u32(&sm.q_ready), 128 * nslot);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&sm.k_full[i]), 1);
            mbar_init(smem_u32(&sm.k_empty[i]), 1);
            mbar_init(smem_u32(&sm.v_full[i]), 1);
            mbar_init(smem_u32(&sm.v_empty[i]), 1);
        }
        for (int t = 0; t < NSLOT; ++t) {
            mbar_init(smem_u32(&sm.s_full[t]), 1);
            mbar_init(smem_u32(&sm.s_free[t]), 128);
            mbar_init(smem_u32(&sm.p_ready[t]), 128);
            mbar_init(smem_u32(&sm.p_free[t]), 1);
            mbar_init(smem_u32(&sm.o_free[t]), 128);
        }
        mbar_init(smem_u32(&sm.lag_bar), 128);
        mbar_fence_init();
    }
    if (warp == W_MMA) {
        tmem_alloc(smem_u32(&sm.tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == W_TMA) {
        // ===================================================================== TMA producer (whole warp walks the loop,
        // one elected lane issues)
        if (elect_one_sync()) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            const uint32_t qf = smem_u32(&sm.q_full);
            mbar_expect_tx(qf, nslot * BOXES * SM::QBOX);
            const int qrow = b * p.Nq + qt * NSLOT * BM;
            for (int t = 0; t < nslot; ++t)
#pragma unroll
                for (int bx = 0; bx < BOXES; ++bx)
                    tma_load_2d(smem_u32(sm.q[t][bx]), &tmQ, qf, head * D + bx * 64, qrow + t * BM);
        }
        __syncwarp();
        for (int i = 0; i < T; ++i) {
            const int s = i / nkt, j = i - s * nkt;
            const int sidx = p.src_index[b * p.n_src_total + p.src_id[s]];
            const bool second = sidx < 0;
            const int row = (second ? -(sidx + 1) : sidx) * p.Nk + j * BN;
            const int st = i % STAGES;
            const uint32_t par = (((uint32_t)(i / STAGES)) & 1u) ^ 1u;
            mbar_wait(smem_u32(&sm.k_empty[st]), par);
            if (elect_one_sync()) {
                const uint32_t fb = smem_u32(&sm.k_full[st]);
                mbar_expect_tx(fb, STAGE_BYTES);
#pragma unroll
                for (int bx = 0; bx < BOXES; ++bx)
                    tma_load_2d(smem_u32(sm.k[st][bx]), second ? &tmK2 : &tmK, fb, head * D + bx * 64, row);
            }
            __syncwarp();
            mbar_wait(smem_u32(&sm.v_empty[st]), par);
            if (elect_one_sync()) {
                const uint32_t fb = smem_u32(&sm.v_full[st]);
                mbar_expect_tx(fb, STAGE_BYTES);
#pragma unroll
                for (int bx = 0; bx < BOXES; ++bx)
                    tma_load_2d(smem_u32(sm.v[st][bx]), second ? &tmV2 : &tmV, fb, head * p.vstride + bx * 64, row);
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        // ===================================================================== MMA issuer (whole warp waits, one elected
        // lane issues tcgen05.mma / tcgen05.commit)
        const uint32_t idesc_qk = make_idesc_f16(BM, BN, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(BM, C::NV, 0, 1);  // B = V, MN-major
        mbar_wait(smem_u32(&sm.q_ready), 0);
        tc_fence_after();
        auto issue_qk = [&](int t, int i) {
            const int st = i % STAGES;
            if (t == 0) mbar_wait(smem_u32(&sm.k_full[st]), ((uint32_t)(i / STAGES)) & 1u);
            if (i > 0) mbar_wait(smem_u32(&sm.s_free[t]), ((uint32_t)(i - 1)) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    // 16 halves = 32 B per k-step inside a 64-column box (4 k-steps per box)
                    const uint64_t qd = make_smem_desc(smem_u32(sm.q[t][ks >> 2]) + (uint32_t)((ks & 3) * 32), 16, 1024, 2);
                    const uint64_t kd = make_smem_desc(smem_u32(sm.k[st][ks >> 2]) + (uint32_t)((ks & 3) * 32), 16, 1024, 2);
                    tc_mma_ss(tmem + C::TM_S + (uint32_t)(t * BN), qd, kd, idesc_qk, (uint32_t)(ks != 0));
                }
                tc_commit(smem_u32(&sm.s_full[t]));
                if (t == nslot - 1) tc_commit(smem_u32(&sm.k_empty[st]));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int i) {
            const int s = i / nkt, j = i - s * nkt;
            const int st = i % STAGES;
            if (t == 0) mbar_wait(smem_u32(&sm.v_full[st]), ((uint32_t)(i / STAGES)) & 1u);
            mbar_wait(smem_u32(&sm.p_ready[t]), ((uint32_t)i) & 1u);
            if (j == 0 && s > 0) mbar_wait(smem_u32(&sm.o_free[t]), ((uint32_t)(s - 1)) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t o_t = tmem + C::TM_O + (uint32_t)(t * C::TM_O_STRIDE);
                const uint32_t p_t = tmem + C::TM_P + (uint32_t)(t * PCOLS);
                // V row-major tile = MN-major B: 8-key groups 1024 B apart (SBO), 64-column atoms one box apart (LBO)
                const uint64_t vd0 = make_smem_desc(smem_u32(sm.v[st][0]), SM::KVBOX, 1024, 2);
#pragma unroll
                for (int kk = 0; kk < PV_STEPS; ++kk)
                    tc_mma_ts(o_t, p_t + (uint32_t)(kk * 8), vd0 + (uint64_t)(kk * 128), idesc_pv, (uint32_t)((j | kk) != 0));
                tc_commit(smem_u32(&sm.p_free[t]));
                if (t == nslot - 1) tc_commit(smem_u32(&sm.v_empty[st]));
            }
            __syncwarp();
        };
        for (int t = 0; t < nslot; ++t) issue_qk(t, 0);
        // Issue order.  qk(1, i+1) sits behind pv(0, i), which blocks until slot 0 has finished the exponentials of tile
        // i: slot 1 therefore starts every tile a fixed lag after slot 0.  That coupling is deliberate - with both QK^T
        // products issued ahead of the PVs (tried in r1j) the slots fall into step, their non-MUFU phases coincide and
        // the kernel is 6 % SLOWER (518 -> 486 TFLOP/s).  GCB_ATTN_LAG > 0 lengthens the lag instead: qk(1, i+1)
        // additionally waits until slot 0 has reached a given point of tile i+1 (lag_bar).  Also measured (r1k) and
        // rejected: 513 -> 427 (signal after the row max) / 434 TFLOP/s (after half of the exponentials) at d=40,
        // 729 -> 535 at d=80.  The round-1 order below is the measured optimum of the three; kept at 0.
        if (ATTN_LAG) mbar_wait(smem_u32(&sm.lag_bar), 0);
        for (int i = 0; i < T; ++i) {
#if defined(GCB_ATTN_ORDER) && GCB_ATTN_ORDER == 1
            // A/B: slot t's next QK^T goes out right after ITS OWN PV (it only needs s_free, signalled early in the tile)
            for (int t = 0; t < nslot; ++t) {
                issue_pv(t, i);
                if (i + 1 < T) issue_qk(t, i + 1);
            }
#else
            for (int t = 0; t < nslot; ++t) {
                if (i + 1 < T) {
                    if (ATTN_LAG && t == 1) mbar_wait(smem_u32(&sm.lag_bar), ((uint32_t)(i + 1)) & 1u);
                    issue_qk(t, i + 1);
                }
                issue_pv(t, i);
            }
#endif
        }
    } else if ((warp >> 2) < nslot) {
        // ===================================================================== softmax slots
        const int t = warp >> 2;         // slot
        const int wq = warp & 3;         // TMEM lane quarter
        const int row = wq * 32 + lane;  // row inside the slot
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        const uint32_t s_t = tmem + lane_base + C::TM_S + (uint32_t)(t * BN);
        const uint32_t o_t = tmem + lane_base + C::TM_O + (uint32_t)(t * C::TM_O_STRIDE);
        const uint32_t p_t = tmem + lane_base + C::TM_P + (uint32_t)(t * PCOLS);
        const uint32_t acc_t = tmem + lane_base + C::TM_ACC + (uint32_t)(t * 80);
        float* acc_row = &sm.acc[C::ACC_IN_TMEM ? 0 : t][0][C::ACC_IN_TMEM ? 0 : row];
        mbar_wait(smem_u32(&sm.q_full), 0);
        if (C::ZERO_Q_PAD) {
            // zero Q columns 40..47: 16-byte chunk 5 of the 128B-swizzled row
            uint8_t* qrow = sm.q[t][0] + (row >> 3) * 1024 + (row & 7) * 128 + ((5 ^ (row & 7)) * 16);
            *reinterpret_cast<uint4*>(qrow) = make_uint4(0, 0, 0, 0);
            fence_proxy_async();
        }
        mbar_arrive(smem_u32(&sm.q_ready));
        // named barriers 1..8: "slot t, quarter wq may run its exponentials"; slot 0 goes first
        const bool pingpong = ATTN_PINGPONG && NSLOT == 2 && nslot == 2;
        const uint32_t bar_mine = 1u + (uint32_t)(t * 4 + wq), bar_other = 1u + (uint32_t)((t ^ 1) * 4 + wq);
        if (pingpong && t == 1) named_bar_arrive(bar_other, 64);

        for (int s = 0; s < p.n_act; ++s) {
            float m = -INFINITY;  // running reference max, exp2 domain
            float l = 0.f;        // running row sum (fp32, of the un-rounded probabilities)
            for (int j = 0; j < nkt; ++j) {
                const int i = s * nkt + j;
                mbar_wait(smem_u32(&sm.s_full[t]), ((uint32_t)i) & 1u);
                tc_fence_after();
                uint32_t sr[BN];
                tmem_ld32_into<0>(s_t, sr);
                tmem_ld32_into<32>(s_t + 32, sr);
                float mxa, mxb, mxc, mxd;
                if (ATTN_SPLIT_LD && BN == 128) {
                    tc_wait_ld();  // first 64 columns are in registers
                    tmem_ld32_into<(BN == 128 ? 64 : 0)>(s_t + 64, sr);
                    tmem_ld32_into<(BN == 128 ? 96 : 0)>(s_t + 96, sr);
                    mxa = __uint_as_float(sr[0]), mxb = __uint_as_float(sr[1]), mxc = __uint_as_float(sr[2]),
                    mxd = __uint_as_float(sr[3]);
#pragma unroll
                    for (int e = 4; e < 64; e += 8) {
                        mxa = fmax3(mxa, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                        mxb = fmax3(mxb, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                        if (e + 4 < 64) {
                            mxc = fmax3(mxc, __uint_as_float(sr[e + 4]), __uint_as_float(sr[e + 5]));
                            mxd = fmax3(mxd, __uint_as_float(sr[e + 6]), __uint_as_float(sr[e + 7]));
                        }
                    }
                    tc_wait_ld();
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sm.s_free[t]));
#pragma unroll
                    for (int e = 64; e < BN; e += 8) {
                        mxa = fmax3(mxa, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                        mxb = fmax3(mxb, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                        mxc = fmax3(mxc, __uint_as_float(sr[e + 4]), __uint_as_float(sr[e + 5]));
                        mxd = fmax3(mxd, __uint_as_float(sr[e + 6]), __uint_as_float(sr[e + 7]));
                    }
                } else {
                if (BN == 128) {
                    tmem_ld32_into<(BN == 128 ? 64 : 0)>(s_t + 64, sr);
                    tmem_ld32_into<(BN == 128 ? 96 : 0)>(s_t + 96, sr);
                }
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(smem_u32(&sm.s_free[t]));
                // tile max (raw scores; scale > 0 so max commutes with scaling): four independent chains
                mxa = __uint_as_float(sr[0]), mxb = __uint_as_float(sr[1]), mxc = __uint_as_float(sr[2]),
                mxd = __uint_as_float(sr[3]);
#pragma unroll
                for (int e = 4; e < BN; e += 8) {
                    mxa = fmax3(mxa, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                    mxb = fmax3(mxb, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                    if (e + 4 < BN) {
                        mxc = fmax3(mxc, __uint_as_float(sr[e + 4]), __uint_as_float(sr[e + 5]));
                        mxd = fmax3(mxd, __uint_as_float(sr[e + 6]), __uint_as_float(sr[e + 7]));
                    }
                }
                }
                const float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)) * p.scale_log2;
                bool waited = (i == 0);
                if (j == 0) {
                    m = mx;
                } else {
                    const bool grow = mx > m + RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, grow)) {
                        // rare: O(i-1) must be complete before it is rescaled
                        mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                        tc_fence_after();
                        waited = true;
                        const float alpha = grow ? exp2f(m - mx) : 1.f;
                        l *= alpha;
#pragma unroll
                        for (int c = 0; c < C::NV / 16; ++c) {
                            uint32_t ov[16];
                            tmem_ld_32x32b_x16(o_t + (uint32_t)(c * 16), ov);
                            tc_wait_ld();
#pragma unroll
                            for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                            tmem_st_32x32b_x16(o_t + (uint32_t)(c * 16), ov);
                        }
                        tc_wait_st();
                        if (grow) m = mx;
                    }
                }
                // p = 2^(s*scale - m), kept in registers as packed halves (reusing the score registers) ...
                const float negm = -m;
                if (ATTN_LAG == 1 && t == 0) lag_arrive(smem_u32(&sm.lag_bar), __float_as_uint(mx));
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                const bool sum_here = !p.l_from_o;
                auto exp_block = [&](auto e_begin, auto e_end) {
#pragma unroll
                for (int e = decltype(e_begin)::value; e < decltype(e_end)::value; e += 2) {
                    const float x0 = fmaf(__uint_as_float(sr[2 * e]), p.scale_log2, negm);
                    const float x1 = fmaf(__uint_as_float(sr[2 * e + 1]), p.scale_log2, negm);
                    const float x2 = fmaf(__uint_as_float(sr[2 * e + 2]), p.scale_log2, negm);
                    const float x3 = fmaf(__uint_as_float(sr[2 * e + 3]), p.scale_log2, negm);
                    if (C::EXP_F16X2) {
                        // packed-half exponentials: row sums then come from the ones column / are summed from halves
                        const uint32_t h01 = ex2_f16x2(cvt_f16x2(x0, x1)), h23 = ex2_f16x2(cvt_f16x2(x2, x3));
                        if (sum_here) {
                            const float2 f01 = unpack_half2(h01), f23 = unpack_half2(h23);
                            l0 += f01.x;
                            l1 += f01.y;
                            l2 += f23.x;
                            l3 += f23.y;
                        }
                        sr[e] = h01;
                        sr[e + 1] = h23;
                        continue;
                    }
                    // d=40 is exp-bound: POLY_OF_8 of every 8 exponentials run on the FMA pipes instead of the MUFU
                    const float p0 = ex2_f32(x0);
                    const float p1 = (C::POLY_OF_8 >= 4 || (C::POLY_OF_8 >= 2 && (e & 2))) ? ex2_poly(x1) : ex2_f32(x1);
                    const float p2 = ex2_f32(x2);
                    const float p3 = (C::POLY_OF_8 >= 3 || (C::POLY_OF_8 >= 1 && (e & 2))) ? ex2_poly(x3) : ex2_f32(x3);
                    if (sum_here) {
                        l0 += p0;
                        l1 += p1;
                        l2 += p2;
                        l3 += p3;
                    }
                    sr[e] = cvt_f16x2(p0, p1);
                    sr[e + 1] = cvt_f16x2(p2, p3);
                }
                };
                using std::integral_constant;
                if (ATTN_SPLIT_ST4 && BN == 128) {
                    const uint32_t pf = smem_u32(&sm.p_free[t]), pf_par = ((uint32_t)(i - 1)) & 1u;
                    exp_block(integral_constant<int, 0>{}, integral_constant<int, BN / 8>{});            // keys 0..31
                    mbar_wait(pf, pf_par);  // the real wait (i == 0: parity 1 of a fresh barrier passes at once)
                    tc_fence_after();
                    tmem_st16_from<0>(p_t, sr);
                    exp_block(integral_constant<int, BN / 8>{}, integral_constant<int, BN / 4>{});       // keys 32..63
                    mbar_wait(pf, pf_par);  // completed phase: only a basic-block boundary
                    tmem_st16_from<(BN == 128 ? 16 : 0)>(p_t + 16, sr);
                    exp_block(integral_constant<int, BN / 4>{}, integral_constant<int, 3 * BN / 8>{});   // keys 64..95
                    mbar_wait(pf, pf_par);
                    tmem_st16_from<(BN == 128 ? 32 : 0)>(p_t + 32, sr);
                    exp_block(integral_constant<int, 3 * BN / 8>{}, integral_constant<int, BN / 2>{});   // keys 96..127
                    l += (l0 + l1) + (l2 + l3);
                    tmem_st16_from<(BN == 128 ? 48 : 0)>(p_t + 48, sr);
                } else if (ATTN_SPLIT_ST && BN == 128) {
                    if (pingpong) named_bar_sync(bar_mine, 64);
                    exp_block(integral_constant<int, 0>{}, integral_constant<int, BN / 4>{});   // keys 0..63
                    if (!waited) {
                        mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                        tc_fence_after();
                    }
                    tmem_st32_from<0>(p_t, sr);                                                   // packed pairs 0..31
                    exp_block(integral_constant<int, BN / 4>{}, integral_constant<int, BN / 2>{});  // keys 64..127
                    // hand the MUFU to the paired warp (slot 1 keeps the token after its very last tile: slot 0 is done)
                    if (pingpong && !(t == 1 && i == T - 1)) named_bar_arrive(bar_other, 64);
                    l += (l0 + l1) + (l2 + l3);
                    tmem_st32_from<(BN == 128 ? 32 : 0)>(p_t + 32, sr);
                } else if (ATTN_SPLIT_ST && BN == 64) {
                    if (pingpong) named_bar_sync(bar_mine, 64);
                    exp_block(integral_constant<int, 0>{}, integral_constant<int, BN / 4>{});   // keys 0..31
                    if (!waited) {
                        mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                        tc_fence_after();
                    }
                    tmem_st16_from<0>(p_t, sr);                                                   // packed pairs 0..15
                    exp_block(integral_constant<int, BN / 4>{}, integral_constant<int, BN / 2>{});  // keys 32..63
                    if (pingpong && !(t == 1 && i == T - 1)) named_bar_arrive(bar_other, 64);
                    l += (l0 + l1) + (l2 + l3);
                    tmem_st16_from<(BN == 64 ? 16 : 0)>(p_t + 16, sr);
                } else {
                exp_block(integral_constant<int, 0>{}, integral_constant<int, BN / 2>{});
                if (ATTN_LAG == 2 && t == 0) lag_arrive(smem_u32(&sm.lag_bar), sr[BN / 4 - 1]);
                l += (l0 + l1) + (l2 + l3);
                // ... so that the wait for P(i-1) to be consumed by its P V product overlaps the exponentials
                if (!waited) {
                    mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                    tc_fence_after();
                }
                // A operand of P V in TMEM: lane = query row, column e = keys (2e, 2e+1) packed
                tmem_st32_from<0>(p_t, sr);
                if (BN == 128) tmem_st32_from<(BN == 128 ? 32 : 0)>(p_t + 32, sr);
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&sm.p_ready[t]));
            }
            // ---- end of source: acc += w_s * O / l
            const int ilast = s * nkt + nkt - 1;
            mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)ilast) & 1u);
            tc_fence_after();
            if (p.l_from_o) {
                uint32_t lv[16];
                tmem_ld_32x32b_x16(o_t + (uint32_t)(D / 16 * 16), lv);  // the 16-column chunk that holds column D
                tc_wait_ld();
                l = __uint_as_float(lv[D % 16]);
            }
            const float wl = p.weight[s] / l;
#pragma unroll
            for (int c = 0; c < C::NV / 16; ++c) {
                uint32_t ov[16], av[16];
                tmem_ld_32x32b_x16(o_t + (uint32_t)(c * 16), ov);
                if (C::ACC_IN_TMEM && s > 0) tmem_ld_32x32b_x16(acc_t + (uint32_t)(c * 16), av);
                tc_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int col = c * 16 + e;
                    if (C::ACC_IN_TMEM) {
                        av[e] = __float_as_uint(fmaf(__uint_as_float(ov[e]), wl, s > 0 ? __uint_as_float(av[e]) : 0.f));
                    } else if (col < D) {
                        acc_row[col * BM] = fmaf(__uint_as_float(ov[e]), wl, s > 0 ? acc_row[col * BM] : 0.f);
                    }
                }
                if (C::ACC_IN_TMEM) tmem_st_32x32b_x16(acc_t + (uint32_t)(c * 16), av);
            }
            if (C::ACC_IN_TMEM) tc_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&sm.o_free[t]));
        }
        // ---- store the row: D halves = D/8 x 16 B
        const long long grow_ = (long long)b * p.Nq + (qt * NSLOT + t) * BM + row;
        uint4* dst = reinterpret_cast<uint4*>(p.out + grow_ * p.ld_out + head * D);
#pragma unroll
        for (int c = 0; c < (D + 15) / 16; ++c) {
            float o16[16];
            if (C::ACC_IN_TMEM) {
                uint32_t av[16];
                tmem_ld_32x32b_x16(acc_t + (uint32_t)(c * 16), av);
                tc_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) o16[e] = __uint_as_float(av[e]);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) o16[e] = (c * 16 + e < D) ? acc_row[(c * 16 + e) * BM] : 0.f;
            }
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
                if (c * 16 + h8 * 8 < D) {
                    uint4 o;
                    o.x = pack_half2(o16[h8 * 8 + 0], o16[h8 * 8 + 1]);
                    o.y = pack_half2(o16[h8 * 8 + 2], o16[h8 * 8 + 3]);
                    o.z = pack_half2(o16[h8 * 8 + 4], o16[h8 * 8 + 5]);
                    o.w = pack_half2(o16[h8 * 8 + 6], o16[h8 * 8 + 7]);
                    dst[c * 2 + h8] = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

int encode_rows(CUtensorMap* tm, const void* base, int ld, long long rows, int width, int box_rows) {
    const uint64_t dims[2] = {(uint64_t)width, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)ld * 2};
    const uint32_t box[2] = {64, (uint32_t)box_rows};
    return gcb_encode_tma(tm, base, 2, dims, strides, box, 1);
}

template <int D_>
int launch(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2, const void* v2, int ld_kv2,
           int B, int Nq, int heads, AttnTcParams& p, cudaStream_t stream) {
    using C = Cfg<D_>;
    // ones column: only where the P V tile has spare columns (NV > D) and the caller laid V out with padded heads
    p.l_from_o = (C::NV > C::D && p.vstride >= C::NV) ? 1 : 0;
    // tensor maps: inner extent = heads*d columns from each base pointer (columns beyond are zero-filled by TMA, so the
    // 64-wide boxes of the last head never read past their tensor).  The batch-row count of the K/V buffers is not
    // part of the ABI: rows are addressed through src_index, the row extent is left open.
    const int width = heads * C::D;
    const long long big = 1ll << 31;
    CUtensorMap tmQ, tmK, tmV, tmK2, tmV2;
    int rc;
    if ((rc = encode_rows(&tmQ, q, ld_q, (long long)B * Nq, width, BM))) return rc;
    if ((rc = encode_rows(&tmK, k, ld_kv, big, width, C::BN))) return rc;
    const int vwidth = heads * p.vstride;
    if ((rc = encode_rows(&tmV, v, ld_kv, big, vwidth, C::BN))) return rc;
    if (k2) {
        if ((rc = encode_rows(&tmK2, k2, ld_kv2, big, width, C::BN))) return rc;
        if ((rc = encode_rows(&tmV2, v2, ld_kv2, big, vwidth, C::BN))) return rc;
    } else {
        tmK2 = tmK;
        tmV2 = tmV;
    }
    const size_t smem = sizeof(Smem<D_>) + 1024;
    static unsigned long long configured = 0;   // one bit per device ordinal
    if (gcb_first_use_on_device(configured)) {
        GCB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid((Nq / BM + C::NSLOT - 1) / C::NSLOT, heads, B);
    attn_tc_kernel<D_><<<grid, 32 * (4 * C::NSLOT + 2), smem, stream>>>(tmQ, tmK, tmV, tmK2, tmV2, p);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

}  // namespace

int gcb_attn_tc_supported(int Nq, int Nk, int heads, int d) {
    (void)heads;
    if (Nq % BM != 0) return 0;
    if (d == 40) return Nk % Cfg<40>::BN == 0 && Nk >= Cfg<40>::BN;
    if (d == 80) return Nk % Cfg<80>::BN == 0 && Nk >= Cfg<80>::BN;
    return 0;
}

int gcb_attn_tc_launch(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2, const void* v2,
                       int ld_kv2, void* out, int ld_out, int B, int Nq, int Nk, int heads, int d, int vstride, int n_src,
                       const int32_t* src_index, const float* h_src_weight, float scale, cudaStream_t stream) {
    AttnTcParams p;
    memset(&p, 0, sizeof(p));
    p.out = (__half*)out;
    p.ld_out = ld_out;
    p.Nq = Nq;
    p.Nk = Nk;
    p.heads = heads;
    p.vstride = vstride;
    p.n_src_total = n_src;
    for (int s = 0; s < n_src; ++s)
        if (h_src_weight[s] != 0.f) {
            p.src_id[p.n_act] = s;
            p.weight[p.n_act] = h_src_weight[s];
            ++p.n_act;
        }
    if (p.n_act == 0) {
        gcb_set_error("all source weights are zero");
        return GCB_ERR_INVALID;
    }
    p.src_index = src_index;
    p.scale_log2 = scale * 1.4426950408889634f;
    GCB_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
                      ((uintptr_t)out % 16) == 0 && ld_out % 8 == 0,
                  "tcgen05 attention needs 16-byte aligned q/k/v/out");
    if (d == 40) return launch<40>(q, ld_q, k, v, ld_kv, k2, v2, ld_kv2, B, Nq, heads, p, stream);
    if (d == 80) return launch<80>(q, ld_q, k, v, ld_kv, k2, v2, ld_kv2, B, Nq, heads, p, stream);
    gcb_set_error("tcgen05 attention: head dim %d not built", d);
    return GCB_ERR_UNSUPPORTED;
}
