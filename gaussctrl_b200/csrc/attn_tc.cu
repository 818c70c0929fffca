// Multi-source (cross-view) attention on the 5th-gen tensor cores for head dims 40 and 80 (the SD1.x 64x64 and 32x32
// levels: ~90 % of the attention FLOPs of the hot path).
//
//   out[b, i, h, :] = sum_s w_s * softmax_j( q[b,i,h,:] . K_s[j,h,:] * scale ) V_s[j,h,:]      (utils.py:25-37, 88-117)
//
// One CTA = NSLOT 128-row query slots of one (batch row, head).  Warp roles (32 * (4 RS NSLOT + 2) threads, RS =
// threads per query row):
//   warps 4 RS t ..   : softmax of slot t - S read from TMEM (tcgen05.ld), exp2, P written back to TMEM (tcgen05.st) as
//                       packed halves: the A operand of P V.  Probabilities never touch shared memory or HBM.
//                       RS = 1: one query row per thread.  RS = 2 (head dim 40): the two warps that own a TMEM lane
//                       quarter split the tile's keys (64 each) and exchange the tile max through shared memory + a
//                       64-thread named barrier - four softmax warps per scheduler instead of two.
//   warp 4 RS NSLOT   : TMA producer (Q once; K and V tiles of BN keys through mbarrier rings)
//   warp 4 RS NSLOT + 1: TMEM allocator + tcgen05.mma issuer (one elected lane):
//                       S   = Q K^T    M128 x N(BN) x K(DK)   both operands K-major, 128B-swizzled 64-column TMA boxes
//                       O  += P V      M128 x N(NV) x K(BN)   A = P from TMEM, B = V MN-major straight from its row-major tile
// Head dim 40: two slots, BN=128, DK=48 - the Q/K boxes are 64 columns wide; columns 40..47 of Q are zeroed in smem so
//   the neighbouring head's columns that ride along in K contribute nothing; NV=48 (columns 40..47 of O are don't-care).
//   With every exponential on the MUFU (exp2: 16/clk/SM) and one thread per row the kernel sat at 530 TFLOP/s: 73 % of
//   its MUFU floor, the two softmax warps of a scheduler leaving the pipe idle whenever both were outside their
//   exponential phase (profiles/r1i_attn_source_summary.md).  Since (profiles/r3_attn.md): (a) 2-3 of every 8 key pairs
//   take a packed-half polynomial on the FMA/ALU pipes instead (ex2_hpoly; the row sums then come out of the P V
//   product through a ones column of V) - 583 TFLOP/s; (b) one MMA-issuing warp per slot - 634; (c) scale and running
//   reference folded into the Q K^T product (Cfg<40>::SHIFT_IN_MMA) - 680.  Two threads per row (four softmax warps per
//   scheduler) is written (ROWSPLIT = 2) but loses.
//   Measured and rejected before (profiles/r2_attn_ab.md): a third slot with BN=64 405 TFLOP/s, BN=64 with two slots
//   405, four-piece P store 495, MUFU ping-pong 500, other MMA issue orders 442-486, ex2.approx.f16x2 (two MUFU ops
//   in SASS) 428, fp32 polynomial 485, setmaxnreg 232/40 split 529 (r3d).
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

template <int D_>
struct Cfg;
template <>
struct Cfg<40> {
    static constexpr int D = 40, BN = 128, DK = 48, NV = 48, BOXES = 1, NSLOT = 2, STAGES = 4;
    static constexpr uint32_t TM_S = 0, TM_O = 256, TM_O_STRIDE = 64, TM_P = 384, TM_ACC = 0;
    static constexpr bool ZERO_Q_PAD = true, ACC_IN_TMEM = false;
    // threads per query row: each takes BN / ROWSPLIT keys of every tile (see the header).  Two threads per row (four
    // softmax warps per scheduler, 96 registers per thread) measured 545-570 TFLOP/s against 620-634 for one: 19 % more
    // instructions per key (per-warp fixed work twice, the max exchange) at the same ~64 % issue-slot use (r3e/r3g/r3j).
#ifdef GCB_ATTN40_ROWSPLIT
    static constexpr int ROWSPLIT = GCB_ATTN40_ROWSPLIT;
#else
    static constexpr int ROWSPLIT = 1;
#endif
    // key pairs out of every 8 whose exponentials run as ONE packed-half polynomial on the FMA / ALU pipes (ex2_hpoly)
    // instead of two MUFU.EX2 + one F2FP; only when the row sums come out of the P V product (V ones column).
    // 2 / 3 / 4 of 8 measure the same (628 / 623 / 617 TFLOP/s, r3h; with the shift in the MMA 673 / 662 / 674, r3s), 5 of
    // 8: 575, 6 of 8: 480 (HFMA2 issues at half rate).  2 of 8 = the fewest instructions of the three.
#ifdef GCB_ATTN_HPOLY
    static constexpr int HPOLY_OF_8 = GCB_ATTN_HPOLY;
#else
    static constexpr int HPOLY_OF_8 = 2;
#endif
    // The scale and the running reference m are folded into the Q K^T product: Q is multiplied by scale * log2(e) once
    // per CTA (in shared memory), and a fourth k-step multiplies Q's columns 48..63 - column 48 holds -m of the row, an
    // fp16-exact value - with a constant tile whose column 48 is 1.0.  S then arrives in TMEM as s * scale - m: the 128
    // FFMAs per row and tile that applied scale and shift (a fifth of the softmax warps' instructions) disappear.  m
    // changes only at the first tile of a source or when a tile's maximum exceeds it by 2^8 (the lazy-rescale rule):
    // the thread then shifts that tile's scores itself, rewrites its -m in shared memory and only afterwards releases S
    // for the next Q K^T.  633 -> 662-674 TFLOP/s (profiles/r3s_ab_attn_shift.txt).  The tile at which m moves runs a second
    // copy of the exponential code that subtracts the difference (modifying the 128 score registers in place in a rare
    // branch made ptxas spill ~25 of them on EVERY tile: 425 TFLOP/s, r3r).
#if defined(GCB_ATTN_NO_SHIFT_MMA)
    static constexpr bool SHIFT_IN_MMA = false;
#else
    static constexpr bool SHIFT_IN_MMA = ROWSPLIT == 1;
#endif
};
template <>
struct Cfg<80> {
    static constexpr int D = 80, BN = 64, DK = 80, NV = 80, BOXES = 2, NSLOT = 2, STAGES = 4;
    static constexpr bool ZERO_Q_PAD = false, ACC_IN_TMEM = true;
    static constexpr int ROWSPLIT = 1, HPOLY_OF_8 = 0;
    static constexpr bool SHIFT_IN_MMA = false;
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
    static constexpr int STAGES = C::STAGES, NSLOT = C::NSLOT, RS = C::ROWSPLIT;
    uint8_t q[NSLOT][C::BOXES][QBOX];
    uint8_t k[STAGES][C::BOXES][KVBOX];
    uint8_t v[STAGES][C::BOXES][KVBOX];
    uint8_t ones[C::SHIFT_IN_MMA ? KVBOX : 1024];   // K-box-shaped constant tile: column 48 = 1.0, columns 49..63 = 0
    float acc[C::ACC_IN_TMEM ? 1 : NSLOT][C::ACC_IN_TMEM ? 1 : C::D][C::ACC_IN_TMEM ? 4 : BM];  // [slot][column][row]
    // row split: the tile max (by tile parity) and the row sum (by source parity) of each thread of a row
    float xmax[RS > 1 ? NSLOT : 1][2][RS > 1 ? RS : 1][RS > 1 ? BM : 4];
    float xsum[RS > 1 ? NSLOT : 1][2][RS > 1 ? RS : 1][RS > 1 ? BM : 4];
    uint64_t q_full, q_ready;
    uint64_t k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
    uint64_t s_full[NSLOT], s_free[NSLOT], p_ready[NSLOT], p_free[NSLOT], o_free[NSLOT];
    uint32_t tmem_base;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_f32(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
    uint32_t y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
    return y;
}
// 2^x for a PAIR of keys in packed-half arithmetic: 11 FMA/ALU-pipe instructions per pair, no MUFU, and the result is
// already the packed fp16 pair the P V product consumes.
//   h  = fp16x2(x), clamped at -15            (x <= ~8: the lazy-rescale threshold; below -15 the probability is 0)
//   fi = h + 1551   (= 1536 + 15: in [1024, 2048) the fp16 ulp is 1, so fi = 1551 + round(h), exactly)
//   f  = h - (fi - 1551)  in [-0.5, 0.5], exact
//   2^n: the low 5 bits of fi's bit pattern 0x660F + n are n + 15 = the biased fp16 exponent: (fi << 10) & 0x7C00
//   2^f: degree-3 Horner in fp16, coefficients chosen for the fp16 evaluation: over EVERY fp16 f in [-0.5, 0.5] the max
//        rel. error is 5.6e-4, rms 2.05e-4, mean -3e-8 - the correctly rounded fp16 of the exact value has 4.9e-4 /
//        2.03e-4 (tools/fit_hpoly.py)
__device__ __forceinline__ uint32_t ex2_hpoly(float lo, float hi) {
    uint32_t h, r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
    asm("{\n\t.reg .b32 h, fi, n, f, t, r;\n\t"
        "max.f16x2 h, %1, %2;\n\t"
        "add.rn.f16x2 fi, h, %3;\n\t"
        "sub.rn.f16x2 n, fi, %3;\n\t"
        "sub.rn.f16x2 f, h, n;\n\t"
        "shl.b32 t, fi, 10;\n\t"
        "and.b32 t, t, 0x7C007C00;\n\t"
        "fma.rn.f16x2 r, f, %4, %5;\n\t"
        "fma.rn.f16x2 r, r, f, %6;\n\t"
        "fma.rn.f16x2 r, r, f, %7;\n\t"
        "mul.rn.f16x2 %0, r, t;\n\t}"
        : "=r"(r)
        : "r"(h), "r"(0xCB80CB80u), "r"(0x660F660Fu), "r"(0x2B082B08u), "r"(0x33C033C0u), "r"(0x398C398Cu),
          "r"(0x3C003C00u));
    return r;
}
__device__ __forceinline__ constexpr bool hpoly_pair(int pair, int of8) { return of8 > 0 && ((pair * of8) & 7) < of8; }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
// NR/32 x tcgen05.ld.x32 of NR consecutive columns
template <int NR>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[NR]) {
    static_assert(NR == 32 || NR == 64 || NR == 128, "32, 64 or 128 score columns per thread");
    tmem_ld32_into<0>(taddr, r);
    if constexpr (NR >= 64) tmem_ld32_into<(NR >= 64 ? 32 : 0)>(taddr + 32, r);
    if constexpr (NR >= 128) {
        tmem_ld32_into<(NR >= 128 ? 64 : 0)>(taddr + 64, r);
        tmem_ld32_into<(NR >= 128 ? 96 : 0)>(taddr + 96, r);
    }
}
// NP packed registers (= 2 NP keys) of P, stored from r[OFF ..]
template <int OFF, int NP, int NR>
__device__ __forceinline__ void tmem_st_packed(uint32_t taddr, const uint32_t (&r)[NR]) {
    static_assert(NP == 16 || NP == 32, "16 or 32 packed registers per store");
    if constexpr (NP == 32) tmem_st32_from<OFF>(taddr, r);
    else tmem_st16_from<OFF>(taddr, r);
}

// One MMA-issuing warp per slot (default) or one for the whole CTA (GCB_ATTN_ONE_MMA_WARP: A/B).  With a single in-order
// issuer the QK^T of slot 0's next tile queues behind the P V of slot 1's previous one, which waits for slot 1's
// softmax: the ncu source view (profiles/r3_attn.md) showed the softmax warps spending 21 % of their time waiting
// for S.  A warp per slot issues S(t, i+1) as soon as slot t has read S(t, i).
#ifdef GCB_ATTN_ONE_MMA_WARP
constexpr bool ATTN_MMA_PER_SLOT = false;
#else
constexpr bool ATTN_MMA_PER_SLOT = true;
#endif
template <int D_>
constexpr int attn_threads() {
    return 32 * (4 * Cfg<D_>::NSLOT * Cfg<D_>::ROWSPLIT + 1 + (ATTN_MMA_PER_SLOT ? Cfg<D_>::NSLOT : 1));
}

template <int D_>
__global__ void __launch_bounds__(attn_threads<D_>(), 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
               const __grid_constant__ CUtensorMap tmV2, const AttnTcParams p) {
    using C = Cfg<D_>;
    using SM = Smem<D_>;
    constexpr int D = C::D, BN = C::BN, BOXES = C::BOXES, NSLOT = C::NSLOT, STAGES = C::STAGES, RS = C::ROWSPLIT;
    constexpr int NSW = 4 * RS;                       // softmax warps per slot
    constexpr int W_TMA = NSW * NSLOT, W_MMA = NSW * NSLOT + 1;
    constexpr int KSTEPS = C::DK / 16 + (C::SHIFT_IN_MMA ? 1 : 0), PV_STEPS = BN / 16, PCOLS = BN / 2;
    static_assert(!C::SHIFT_IN_MMA || (C::DK == 48 && BN == BM && RS == 1), "shift step = k-step 3 of the first 64-column box");
    constexpr int CW = BN / RS;                       // score columns (keys) of a tile per softmax thread
    constexpr uint32_t STAGE_BYTES = BOXES * SM::KVBOX;
    constexpr uint32_t SLOT_THREADS = 32 * NSW;
    static_assert(RS == 1 || RS == 2, "one or two threads per query row");
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
        mbar_init(smem_u32(&sm.q_ready), SLOT_THREADS * nslot);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&sm.k_full[i]), 1);
            mbar_init(smem_u32(&sm.k_empty[i]), ATTN_MMA_PER_SLOT ? nslot : 1);   // one commit per issuing warp
            mbar_init(smem_u32(&sm.v_full[i]), 1);
            mbar_init(smem_u32(&sm.v_empty[i]), ATTN_MMA_PER_SLOT ? nslot : 1);
        }
        for (int t = 0; t < NSLOT; ++t) {
            mbar_init(smem_u32(&sm.s_full[t]), 1);
            mbar_init(smem_u32(&sm.s_free[t]), SLOT_THREADS);
            mbar_init(smem_u32(&sm.p_ready[t]), SLOT_THREADS);
            mbar_init(smem_u32(&sm.p_free[t]), 1);
            mbar_init(smem_u32(&sm.o_free[t]), SLOT_THREADS);
        }
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
    } else if (warp >= W_MMA) {
        // ===================================================================== MMA issuer(s) (whole warp waits, one
        // elected lane issues tcgen05.mma / tcgen05.commit)
        const int my_slot = warp - W_MMA;   // per-slot issuers: the slot this warp serves
        const uint32_t idesc_qk = make_idesc_f16(BM, BN, 0, 0);
        const uint32_t idesc_pv = make_idesc_f16(BM, C::NV, 0, 1);  // B = V, MN-major
        mbar_wait(smem_u32(&sm.q_ready), 0);
        tc_fence_after();
        auto issue_qk = [&](int t, int i) {
            const int st = i % STAGES;
            if (ATTN_MMA_PER_SLOT || t == 0) mbar_wait(smem_u32(&sm.k_full[st]), ((uint32_t)(i / STAGES)) & 1u);
            if (i > 0) mbar_wait(smem_u32(&sm.s_free[t]), ((uint32_t)(i - 1)) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    // 16 halves = 32 B per k-step inside a 64-column box (4 k-steps per box)
                    const uint64_t qd = make_smem_desc(smem_u32(sm.q[t][ks >> 2]) + (uint32_t)((ks & 3) * 32), 16, 1024, 2);
                    // the shift step (ks == 3) multiplies Q's columns 48..63 with the constant tile instead of K
                    const uint32_t kbox = (C::SHIFT_IN_MMA && ks == KSTEPS - 1) ? smem_u32(sm.ones) : smem_u32(sm.k[st][ks >> 2]);
                    const uint64_t kd = make_smem_desc(kbox + (uint32_t)((ks & 3) * 32), 16, 1024, 2);
                    tc_mma_ss(tmem + C::TM_S + (uint32_t)(t * BN), qd, kd, idesc_qk, (uint32_t)(ks != 0));
                }
                tc_commit(smem_u32(&sm.s_full[t]));
                if (ATTN_MMA_PER_SLOT || t == nslot - 1) tc_commit(smem_u32(&sm.k_empty[st]));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t, int i) {
            const int s = i / nkt, j = i - s * nkt;
            const int st = i % STAGES;
            if (ATTN_MMA_PER_SLOT || t == 0) mbar_wait(smem_u32(&sm.v_full[st]), ((uint32_t)(i / STAGES)) & 1u);
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
                if (ATTN_MMA_PER_SLOT || t == nslot - 1) tc_commit(smem_u32(&sm.v_empty[st]));
            }
            __syncwarp();
        };
        if (ATTN_MMA_PER_SLOT) {
            if (my_slot < nslot) {
                issue_qk(my_slot, 0);
                for (int i = 0; i < T; ++i) {
                    if (i + 1 < T) issue_qk(my_slot, i + 1);   // needs only K(i+1) and slot's read of S(i)
                    issue_pv(my_slot, i);
                }
            }
        } else {
            for (int t = 0; t < nslot; ++t) issue_qk(t, 0);
            // One issuer for both slots: qk(1, i+1) sits behind pv(0, i), which blocks until slot 0 has finished the
            // exponentials of tile i.  Other single-issuer orders measured in rounds 1-2 (profiles/r2_attn_ab.md): both
            // QK^T products ahead of the PVs -6 %, an extra lag barrier -17 %, a slot's next QK^T after its own PV -17 %.
            for (int i = 0; i < T; ++i) {
                for (int t = 0; t < nslot; ++t) {
                    if (i + 1 < T) issue_qk(t, i + 1);
                    issue_pv(t, i);
                }
            }
        }
    } else if (warp / NSW < nslot) {
        // ===================================================================== softmax slots
        const int t = warp / NSW;                 // slot
        const int wq = warp & 3;                  // TMEM lane quarter (NSW is a multiple of 4)
        const int ch = (warp % NSW) >> 2;         // which part of the tile's keys (row split)
        const int row = wq * 32 + lane;           // row inside the slot
        const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
        const uint32_t s_t = tmem + lane_base + C::TM_S + (uint32_t)(t * BN + ch * CW);
        const uint32_t o_t = tmem + lane_base + C::TM_O + (uint32_t)(t * C::TM_O_STRIDE);
        const uint32_t p_t = tmem + lane_base + C::TM_P + (uint32_t)(t * PCOLS + ch * (CW / 2));
        const uint32_t acc_t = tmem + lane_base + C::TM_ACC + (uint32_t)(t * 80);
        float* acc_row = &sm.acc[C::ACC_IN_TMEM ? 0 : t][0][C::ACC_IN_TMEM ? 0 : row];
        // named barrier of the RS warps that share this slot's TMEM lane quarter (ids 1 .. 4 NSLOT)
        const uint32_t pair_bar = 1u + (uint32_t)(t * 4 + wq);
        mbar_wait(smem_u32(&sm.q_full), 0);
        // this thread's Q row in the 128B-swizzled box: 16-byte chunk j at (j ^ (row & 7)) * 16
        uint8_t* const qrow = sm.q[t][0] + (row >> 3) * 1024 + (row & 7) * 128;
        auto qchunk = [&](int j) { return reinterpret_cast<uint4*>(qrow + ((j ^ (row & 7)) * 16)); };
        if constexpr (C::SHIFT_IN_MMA) {
            // Q *= scale * log2(e) (columns 0..39), columns 40..63 = 0 (column 48 will hold -m: m starts at 0)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                uint4 v = *qchunk(j);
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_half2(w[e]);
                    w[e] = pack_half2(f.x * p.scale_log2, f.y * p.scale_log2);
                }
                *qchunk(j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
#pragma unroll
            for (int j = 5; j < 8; ++j) *qchunk(j) = make_uint4(0, 0, 0, 0);
            if (t == 0) {   // the constant tile, one row per thread of slot 0 (BN == BM rows)
                uint8_t* orow = sm.ones + (row >> 3) * 1024 + (row & 7) * 128;
                *reinterpret_cast<uint4*>(orow + ((6 ^ (row & 7)) * 16)) = make_uint4(0x00003C00u, 0, 0, 0);   // 1.0, 0, ...
                *reinterpret_cast<uint4*>(orow + ((7 ^ (row & 7)) * 16)) = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async();
        } else if (C::ZERO_Q_PAD && ch == 0) {
            // zero Q columns 40..47 (chunk 5)
            *qchunk(5) = make_uint4(0, 0, 0, 0);
            fence_proxy_async();
        }
        mbar_arrive(smem_u32(&sm.q_ready));
        // combine a per-thread value over the RS threads of a row (same order in every thread: bit-identical)
        auto row_max = [&](float v, int par) {
            if constexpr (RS > 1) {
                // buffer `par` is rewritten two tiles later, behind the barrier of the tile in between, which the other
                // thread only reaches after it has read this one
                sm.xmax[t][par][ch][row] = v;
                named_bar_sync(pair_bar, 32 * RS);
                return fmaxf(sm.xmax[t][par][0][row], sm.xmax[t][par][1][row]);
            } else {
                return v;
            }
        };
        auto row_sum = [&](float v, int par) {
            if constexpr (RS > 1) {
                sm.xsum[t][par][ch][row] = v;
                named_bar_sync(pair_bar, 32 * RS);
                return sm.xsum[t][par][0][row] + sm.xsum[t][par][1][row];
            } else {
                return v;
            }
        };

        for (int s = 0; s < p.n_act; ++s) {
            // SHIFT_IN_MMA: the reference the Q K^T product subtracts (fp16-exact; -m_q sits in Q's column 48).  Every
            // source starts from 0 - its result does not depend on the sources before it (a source listed twice at half
            // the weight gives bit-identical output to listing it once)
            float m_q = 0.f;
            float m = -INFINITY;  // running reference max, exp2 domain
            float l = 0.f;        // running row sum of this thread's keys (fp32, of the un-rounded probabilities)
            for (int j = 0; j < nkt; ++j) {
                const int i = s * nkt + j;
                mbar_wait(smem_u32(&sm.s_full[t]), ((uint32_t)i) & 1u);
                tc_fence_after();
                uint32_t sr[CW];
                tmem_ld_cols(s_t, sr);
                tc_wait_ld();
                if constexpr (!C::SHIFT_IN_MMA) {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sm.s_free[t]));
                }
                // tile max (raw scores; scale > 0 so max commutes with scaling): four independent chains
                float mxa = __uint_as_float(sr[0]), mxb = __uint_as_float(sr[1]), mxc = __uint_as_float(sr[2]),
                      mxd = __uint_as_float(sr[3]);
#pragma unroll
                for (int e = 4; e < CW; e += 8) {
                    mxa = fmax3(mxa, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                    mxb = fmax3(mxb, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                    if (e + 4 < CW) {
                        mxc = fmax3(mxc, __uint_as_float(sr[e + 4]), __uint_as_float(sr[e + 5]));
                        mxd = fmax3(mxd, __uint_as_float(sr[e + 6]), __uint_as_float(sr[e + 7]));
                    }
                }
                bool waited = (i == 0);
                float delta = 0.f;    // SHIFT_IN_MMA: what this tile's scores still have to be shifted by (0 except when
                bool shifted = false; // the reference moved at this tile; warp-uniform flag)
                // rare: O(i-1) must be complete before it is rescaled by alpha
                auto rescale_o = [&](float alpha) {
                    mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                    tc_fence_after();
                    waited = true;
                    l *= alpha;
                    if (ch == 0) {   // one warp of the row rescales O; P V (i) waits for every warp's p_ready
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
                    }
                };
                if constexpr (C::SHIFT_IN_MMA) {
                    // the scores are already s * scale - m_q.  New reference at the first tile of a source and when the
                    // tile maximum exceeds the current one by 2^8: an fp16-exact value, so that Q's column 48 holds it exactly
                    const float mx = fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd));
                    const bool move = (j == 0) || (mx > RESCALE_THRESHOLD);
                    if (__builtin_expect(__any_sync(0xffffffffu, move) != 0, 0)) {
                        const float m_new = move ? __half2float(__float2half_rn(m_q + mx)) : m_q;
                        delta = m_new - m_q;   // difference of two fp16 values: exact
                        shifted = true;        // this tile's exponentials take the copy of the code that subtracts delta
                        if (j > 0) rescale_o(exp2f(-delta));
                        if (move) {
                            m_q = m_new;
                            *reinterpret_cast<__half*>(qchunk(6)) = __float2half_rn(-m_new);
                            fence_proxy_async();
                        }
                    }
                    if (j == nkt - 1) {   // the first tile of the next source is multiplied with a zero reference again
                        *reinterpret_cast<__half*>(qchunk(6)) = __float2half_rn(0.f);
                        fence_proxy_async();
                    }
                    // S is released only now: the next Q K^T of this slot must see the new -m
                    tc_fence_before();
                    mbar_arrive(smem_u32(&sm.s_free[t]));
                } else {
                const float mx = row_max(fmaxf(fmaxf(mxa, mxb), fmaxf(mxc, mxd)), i & 1) * p.scale_log2;
                if (j == 0) {
                    m = mx;
                } else {
                    // every thread of a row sees the same mx and m: the RS warps of a lane quarter take the same branch
                    const bool grow = mx > m + RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, grow)) {
                        rescale_o(grow ? exp2f(m - mx) : 1.f);
                        if (grow) m = mx;
                    }
                }
                }
                // p = 2^(s*scale - m), kept in registers as packed halves (reusing the score registers) ...
                const float negm = -m;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                // HP: packed-half polynomial for part of the pairs (row sums from the ones column: nothing to add up)
                auto exp_block = [&](auto hp_tag, auto sh_tag, auto e_begin, auto e_end) {
                    constexpr bool HP = decltype(hp_tag)::value, SH = decltype(sh_tag)::value;
#pragma unroll
                    for (int e = decltype(e_begin)::value; e < decltype(e_end)::value; e += 2) {
                        float x0 = __uint_as_float(sr[2 * e]), x1 = __uint_as_float(sr[2 * e + 1]);
                        float x2 = __uint_as_float(sr[2 * e + 2]), x3 = __uint_as_float(sr[2 * e + 3]);
                        if constexpr (SH) {   // the tile at which the reference moved: scores are relative to the old one
                            x0 -= delta;
                            x1 -= delta;
                            x2 -= delta;
                            x3 -= delta;
                        }
                        if constexpr (!C::SHIFT_IN_MMA) {
                            x0 = fmaf(x0, p.scale_log2, negm);
                            x1 = fmaf(x1, p.scale_log2, negm);
                            x2 = fmaf(x2, p.scale_log2, negm);
                            x3 = fmaf(x3, p.scale_log2, negm);
                        }
                        if constexpr (HP) {
                            if (hpoly_pair(e, C::HPOLY_OF_8)) sr[e] = ex2_hpoly(x0, x1);
                            else sr[e] = cvt_f16x2(ex2_f32(x0), ex2_f32(x1));
                            if (hpoly_pair(e + 1, C::HPOLY_OF_8)) sr[e + 1] = ex2_hpoly(x2, x3);
                            else sr[e + 1] = cvt_f16x2(ex2_f32(x2), ex2_f32(x3));
                        } else {
                            const float p0 = ex2_f32(x0), p1 = ex2_f32(x1), p2 = ex2_f32(x2), p3 = ex2_f32(x3);
                            l0 += p0;
                            l1 += p1;
                            l2 += p2;
                            l3 += p3;
                            sr[e] = cvt_f16x2(p0, p1);
                            sr[e + 1] = cvt_f16x2(p2, p3);
                        }
                    }
                };
                using std::integral_constant;
                // ... in two pieces: the first half of the packed probabilities is stored (tcgen05.st) after half of the
                // exponentials; the p_free spin in between also splits the basic block, so ptxas cannot sink all packs
                // behind the last ex2 (r1o: +10 %).  The wait for P(i-1) to be consumed by its P V product overlaps the
                // first half of the exponentials.
                auto two_pieces = [&](auto hp_tag, auto sh_tag) {
                    exp_block(hp_tag, sh_tag, integral_constant<int, 0>{}, integral_constant<int, CW / 4>{});
                    if (!waited) {
                        mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)(i - 1)) & 1u);
                        tc_fence_after();
                    }
                    tmem_st_packed<0, CW / 4>(p_t, sr);
                    exp_block(hp_tag, sh_tag, integral_constant<int, CW / 4>{}, integral_constant<int, CW / 2>{});
                    tmem_st_packed<CW / 4, CW / 4>(p_t + CW / 4, sr);
                };
                if (C::HPOLY_OF_8 > 0 && p.l_from_o) {
                    if (C::SHIFT_IN_MMA && shifted) two_pieces(std::true_type{}, std::true_type{});
                    else two_pieces(std::true_type{}, std::false_type{});
                } else {
                    if (C::SHIFT_IN_MMA && shifted) two_pieces(std::false_type{}, std::true_type{});
                    else two_pieces(std::false_type{}, std::false_type{});
                    l += (l0 + l1) + (l2 + l3);
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&sm.p_ready[t]));
            }
            // ---- end of source: acc += w_s * O / l  (the 16-column chunks of O are dealt to the RS threads of the row)
            const int ilast = s * nkt + nkt - 1;
            mbar_wait(smem_u32(&sm.p_free[t]), ((uint32_t)ilast) & 1u);
            tc_fence_after();
            if (p.l_from_o) {
                uint32_t lv[16];
                tmem_ld_32x32b_x16(o_t + (uint32_t)(D / 16 * 16), lv);  // the 16-column chunk that holds column D
                tc_wait_ld();
                l = __uint_as_float(lv[D % 16]);
            } else {
                l = row_sum(l, s & 1);
            }
            const float wl = p.weight[s] / l;
#pragma unroll
            for (int c = 0; c < C::NV / 16; ++c) {
                if (c % RS != ch) continue;
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
        // ---- store the row: D halves = D/8 x 16 B; every thread stores the chunks it accumulated itself
        const long long grow_ = (long long)b * p.Nq + (qt * NSLOT + t) * BM + row;
        uint4* dst = reinterpret_cast<uint4*>(p.out + grow_ * p.ld_out + head * D);
#pragma unroll
        for (int c = 0; c < (D + 15) / 16; ++c) {
            if (c % RS != ch) continue;
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
    attn_tc_kernel<D_><<<grid, attn_threads<D_>(), smem, stream>>>(tmQ, tmK, tmV, tmK2, tmV2, p);
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
