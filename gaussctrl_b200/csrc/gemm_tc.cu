// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (kind::f16) -> TMEM accumulator
//   -> software-pipelined tcgen05.ld epilogue (bias / per-image vector / residual / SiLU / GEGLU) -> 32x64 fp16 tiles
//   staged in the (by then free) operand ring -> cp.async.bulk.tensor store (TMA), clipped at the tensor bounds.
// Two schedules share one kernel body (template parameter PERSIST):
//   * one tile per CTA: two CTAs resident per SM (<= 110 KB of ring each, <= 256 TMEM columns each, 167 registers x 192
//     threads), so the epilogue of one tile overlaps the main loop of another.  Per layer shape: profiles/r1i_gemm_table.txt.
//   * persistent: one CTA per SM walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring runs across tile
//     boundaries (the operands of tile i+1 are in flight while tile i is multiplied), TWO accumulators live in TMEM so
//     the main loop of tile i+1 runs under the epilogue of tile i, and EIGHT epilogue warps (two per TMEM lane quarter,
//     each draining half of the 64-column groups) write through their own staging buffers.  This is the schedule for the
//     short-K layers (K = 320 .. 1280 at M = 98 304), where a one-tile CTA spends most of its life in setup, first-load
//     latency and drain instead of in the tensor pipe.
//
// y[M, Cout] = act( im2col(x)[M, taps*Cin] * w[Cout, taps*Cin]^T + bias + rowvec[b] ) + residual
//
// A operand: the NHWC activation tensor is described to TMA as a 4-D tensor (C, W, H, B).  One M tile = 128
// consecutive output pixels = a (bw x bh x bb) pixel box; filter tap (kh,kw) is the SAME box shifted by
// (kw-1, kh-1) – TMA's out-of-bounds zero fill implements the padding, so no im2col buffer ever exists.
// B operand: w is [Cout, taps*Cin] (K contiguous) -> 2-D TMA box (64 x BN).
// Warp roles (192 / 320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 (2..9 when persistent) = epilogue (each owns the 32 TMEM lanes (warp_id % 4) * 32 ...).
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

#include <type_traits>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 halves = 128 B = one swizzle-128B atom row
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmTcParams {
    int M, N;          // GEMM sizes (N = Cout)
    int nkb;           // total K blocks
    int kb_per_tap;    // Cin / 64 (conv) or nkb (linear)
    int mode;          // 0 = linear (2-D A map), 1 = conv3x3 (4-D A map)
    int H, W, HW;
    int bw, bh, bb;    // pixel box of one M tile
    int BN, stages;
    uint32_t tmem_cols, idesc;
    const __half* bias;
    const __half* rowvec;
    int rowvec_ld;
    const __half* residual;
    __half* y;
    int ldy;
    int act;
    int epi;           // 0 = each thread stores its own row (64 B pieces); 1 = 32x64 tiles staged in smem + TMA store
    int n_tiles, ntiles;      // tiles along N, tiles in total (tile index = m_tile * n_tiles + n_tile)
    uint32_t stage_off;       // persistent: byte offset of the epilogue staging area behind the operand ring
    int stage_bufs;           // persistent: staging buffers per epilogue warp (1 or 2)
};

__device__ __forceinline__ float apply_act(float v, int act) { return act == GCB_ACT_SILU ? silu_f(v) : v; }

// Epilogue staging (p.epi = 1): once tmem_full_bar has fired every MMA has consumed its operands, so the TMA ring is
// free: epilogue warp q owns two 4 KB buffers at ring offset q * 8 KB.  A buffer is one 32-row x 64-column fp16 tile in
// the 128B-swizzled layout of the output tensor map: row r at r * 128 B, 16-byte piece j at (j ^ (r & 7)) * 16 -
// conflict-free for "one row per thread" writes - and leaves through ONE cp.async.bulk.tensor store of full 128 B
// lines instead of 32 lanes x 4 scattered 16 B stores (32 L1 wavefronts per instruction).
__device__ __forceinline__ void stage_row_half(uint32_t buf, int lane, int half, const float (&v)[32]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int j = half * 4 + g;
        st_shared_v4(buf + (uint32_t)(lane * 128) + (uint32_t)((j ^ (lane & 7)) << 4), pack_half2(v[g * 8 + 0], v[g * 8 + 1]),
                     pack_half2(v[g * 8 + 2], v[g * 8 + 3]), pack_half2(v[g * 8 + 4], v[g * 8 + 5]),
                     pack_half2(v[g * 8 + 6], v[g * 8 + 7]));
    }
}

// Fused GEMM + all-gather (sharded reference pass): besides its own output tensor the epilogue stores every
// staged 32x64 tile into the SAME coordinates of up to 7 peers' arenas (TMA bulk stores over NVLink) - the transfer of
// tile i overlaps the main loop / epilogue of tile i+1 instead of following the GEMM as a separate push kernel.
constexpr int MAX_PEER_MAPS = 7;
struct alignas(64) PeerMaps {
    CUtensorMap m[MAX_PEER_MAPS];
    __half* y[MAX_PEER_MAPS];   // the same tensors as raw pointers: a tile whose last 32-column chunk has no partner (BN = 160)
    int n;                      // is stored per thread, and then to the peers as well
    int col_min;                // output columns below this stay local (the Q third of a fused q|k|v projection)
};
struct NoPeers {
    int n;
};

template <bool PEER, bool PERSIST>
__global__ void __launch_bounds__(PERSIST ? 320 : 192, PERSIST ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const GemmTcParams p,
               const __grid_constant__ typename std::conditional<PEER, PeerMaps, NoPeers>::type pm) {
    static_assert(!(PEER && PERSIST), "the fused all-gather epilogue runs on the one-tile schedule");
    constexpr int NH = PERSIST ? 2 : 1;          // epilogue warps per TMEM lane quarter (column halves)
    constexpr int NEPI = 4 * NH;                 // epilogue warps
    constexpr int NACC = PERSIST ? 2 : 1;        // accumulators in TMEM
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t acc_full_bar[NACC];    // MMA -> epilogue: accumulator complete
    __shared__ __align__(8) uint64_t acc_empty_bar[NACC];   // epilogue -> MMA: accumulator drained (persistent only)
    __shared__ uint32_t tmem_base_smem;

    const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_stage_bytes = (uint32_t)p.BN * BK * 2;
    const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    const int tile_step = PERSIST ? (int)gridDim.x : p.ntiles;   // one-tile schedule: the loops below run once

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(smem_u32(&acc_full_bar[a]), 1);
            mbar_init(smem_u32(&acc_empty_bar[a]), NEPI);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols * NACC);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ---------------- TMA producer: the whole warp walks the ring, one elected lane issues ----------------
        if (elect_one_sync()) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += tile_step) {
            const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
            int b0 = 0, h0 = 0, w0 = 0;
            if (p.mode == 1) {
                if (p.HW >= BM) {
                    const int tiles_per_img = p.HW / BM;
                    b0 = m_tile / tiles_per_img;
                    const int r = m_tile % tiles_per_img;
                    if (p.W >= BM) {
                        const int segs = p.W / BM;
                        h0 = r / segs;
                        w0 = (r % segs) * BM;
                    } else {
                        h0 = r * p.bh;
                    }
                } else {
                    b0 = m_tile * p.bb;
                }
            }
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                if (elect_one_sync()) {
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, stage_bytes);
                    const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    if (p.mode == 0) {
                        tma_load_2d(sa, &tmA, fb, kb * BK, m_tile * BM);
                    } else {
                        const int tap = kb / p.kb_per_tap, cb = kb - tap * p.kb_per_tap;
                        const int kh = tap / 3, kw = tap - kh * 3;
                        tma_load_4d(sa, &tmA, fb, cb * BK, w0 + kw - 1, h0 + kh - 1, b0);
                    }
                    tma_load_2d(sb, &tmB, fb, kb * BK, n_tile * p.BN);
                }
                __syncwarp();
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: whole warp waits, one elected lane issues tcgen05.mma / commit ----------------
        int s = 0;
        uint32_t ph = 0;
        int it = 0;   // tiles done by this CTA
        for (int tile = blockIdx.x; tile < p.ntiles; tile += tile_step, ++it) {
            const int acc = PERSIST ? (it & 1) : 0;
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * p.tmem_cols;
            if (PERSIST) {
                // the epilogue of the tile before last must have drained this accumulator (first two uses pass at once)
                mbar_wait(smem_u32(&acc_empty_bar[acc]), (((uint32_t)it >> 1) & 1u) ^ 1u);
                tc_fence_after();
            }
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(smem_u32(&full_bar[s]), ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t sb = sa + A_STAGE_BYTES;
                    const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
                    const uint64_t bdesc = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 halves = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
                        tc_mma_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc,
                                  (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(smem_u32(&empty_bar[s]));
                    if (kb == p.nkb - 1) tc_commit(smem_u32(&acc_full_bar[acc]));
                }
                __syncwarp();
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        // ---------------- epilogue: warps 2.., TMEM lane quarter = warp % 4 ----------------
        // Software-pipelined: the tcgen05.ld of the NEXT column chunk is in flight while the current chunk is
        // converted, so the TMEM read latency is exposed once per tile instead of once per chunk.
        // Persistent schedule: warps 2..5 take the even 64-column groups of a tile, warps 6..9 the odd ones.
        const int q = warp & 3;
        const int half = PERSIST ? ((warp - 2) >> 2) : 0;
        const int row = q * 32 + lane;
        const bool issuer = elect_one_sync() != 0;            // the one lane that issues / waits on this warp's TMA stores
        // staging: one-tile schedule = the (by then free) operand ring, 2 x 4 KB per warp; persistent = own area
        const uint32_t stage_buf = PERSIST ? smem_base + p.stage_off + (uint32_t)((warp - 2) * p.stage_bufs) * 4096u
                                           : smem_base + (uint32_t)q * 8192u;
        const int nbufs = PERSIST ? p.stage_bufs : 2;
        int n_pairs = 0;                                        // staged 64-column groups so far
        int cur_buf = 0;                                        // staging buffer of the current group
        auto add8 = [](float* v, const uint4& u) {
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_half2(w[t]);
                v[t * 2] += f.x;
                v[t * 2 + 1] += f.y;
            }
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += tile_step, ++it) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const int acc = PERSIST ? (it & 1) : 0;
        const uint32_t acc_par = PERSIST ? (((uint32_t)it >> 1) & 1u) : 0u;
        const long long m = (long long)m_tile * BM + row;
        const uint32_t taddr = tmem_base + (uint32_t)acc * p.tmem_cols + ((uint32_t)(q * 32) << 16);
        const bool row_ok = m < p.M;
        const int img = (p.rowvec != nullptr && row_ok) ? (int)(m / p.HW) : 0;
        const int y_row0 = m_tile * BM + q * 32;
        // a 64-column group is complete in the staging buffer: one bulk store, then switch buffers
        auto flush_group = [&](int col0) {
            fence_proxy_async_smem();
            __syncwarp();
            if (issuer) {
                tma_store_2d(&tmY, stage_buf + (uint32_t)(cur_buf * 4096), col0, y_row0);
                if constexpr (PEER) {
                    if (col0 + 64 > pm.col_min)   // any column of the group at or above col_min
                        for (int q_ = 0; q_ < pm.n; ++q_)
                            tma_store_2d(&pm.m[q_], stage_buf + (uint32_t)(cur_buf * 4096), col0, y_row0);
                }
                tma_store_commit();
            }
            ++n_pairs;
            cur_buf = (cur_buf + 1 == nbufs) ? 0 : cur_buf + 1;
        };
        // before the first write into a staging buffer: the bulk store that last used it must have read it
        auto acquire_buffer = [&]() {
            if (n_pairs >= nbufs) {
                if (issuer) {
                    if (nbufs == 2) tma_store_wait_read<1>();
                    else tma_store_wait_read<0>();
                }
                __syncwarp();
            }
        };
        // the accumulator is in registers / stored: hand it back to the MMA issuer
        auto release_acc = [&]() {
            if (PERSIST) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[acc]));
            }
        };
        if (p.act != GCB_ACT_GEGLU) {
            const int nchunks = p.BN / 32;
            const int ncol0 = n_tile * p.BN;
            const __half* res_row = p.residual ? p.residual + m * p.ldy : nullptr;
            const __half* rv_row = p.rowvec ? p.rowvec + (long long)img * p.rowvec_ld : nullptr;
            // the residual of chunk c+1 is in flight while chunk c is converted and stored
            uint4 res_cur[4], res_nxt[4];
            auto load_res = [&](int c, uint4 (&dst)[4]) {
                const int n0 = ncol0 + c * 32;
                if (res_row && row_ok && n0 + 32 <= p.N) {
                    const uint4* rp = reinterpret_cast<const uint4*>(res_row + n0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) dst[g] = rp[g];
                }
            };
            bool pair_staged = false;
            // c = this chunk, c_next = the chunk this warp converts after it (>= nchunks: none)
            auto chunk = [&](const uint32_t (&r)[32], int c, int c_next) {
                const int n0 = ncol0 + c * 32;
                if ((c & 1) == 0) {
                    pair_staged = p.epi && c + 1 < nchunks && n0 + 64 <= p.N;
                    if (pair_staged) acquire_buffer();
                }
                if (c_next < nchunks) load_res(c_next, res_nxt);
                const bool full = n0 + 32 <= p.N;
                if ((row_ok || pair_staged) && n0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    __half* yp = p.y + m * p.ldy + n0;
                    if (full) {
                        if (p.bias) {
                            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
                            for (int g = 0; g < 4; ++g) add8(v + g * 8, bp[g]);
                        }
                        if (rv_row) {
                            const uint4* rp = reinterpret_cast<const uint4*>(rv_row + n0);
#pragma unroll
                            for (int g = 0; g < 4; ++g) add8(v + g * 8, rp[g]);
                        }
                        if (p.act == GCB_ACT_SILU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                        }
                        if (res_row && row_ok) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) add8(v + g * 8, res_cur[g]);
                        }
                        if (pair_staged) {
                            stage_row_half(stage_buf + (uint32_t)(cur_buf * 4096), lane, c & 1, v);
                        } else {
                            uint4* op = reinterpret_cast<uint4*>(yp);
                            uint4 o[4];
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                o[g].x = pack_half2(v[g * 8 + 0], v[g * 8 + 1]);
                                o[g].y = pack_half2(v[g * 8 + 2], v[g * 8 + 3]);
                                o[g].z = pack_half2(v[g * 8 + 4], v[g * 8 + 5]);
                                o[g].w = pack_half2(v[g * 8 + 6], v[g * 8 + 7]);
                                op[g] = o[g];
                            }
                            if constexpr (PEER) {
                                for (int q_ = 0; q_ < (n0 + 32 > pm.col_min ? pm.n : 0); ++q_) {
                                    uint4* pq = reinterpret_cast<uint4*>(pm.y[q_] + m * p.ldy + n0);
#pragma unroll
                                    for (int g = 0; g < 4; ++g) pq[g] = o[g];
                                }
                            }
                        }
                    } else {
                        const int nvalid = p.N - n0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nvalid) {
                                float o = v[j];
                                if (p.bias) o += __half2float(p.bias[n0 + j]);
                                if (rv_row) o += __half2float(rv_row[n0 + j]);
                                if (p.act == GCB_ACT_SILU) o = silu_f(o);
                                if (res_row) o += __half2float(res_row[n0 + j]);
                                yp[j] = __float2half_rn(o);
                            }
                        }
                    }
                }
                if (pair_staged && (c & 1)) flush_group(n0 - 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) res_cur[g] = res_nxt[g];
            };
            int c = 2 * half;   // this warp's chunks: the pairs (c, c+1), c = 2*half, 2*half + 2*NH, ...
            if (c < nchunks) load_res(c, res_cur);
            mbar_wait(smem_u32(&acc_full_bar[acc]), acc_par);
            tc_fence_after();
            if (!PERSIST && p.epi) fence_proxy_async_smem();  // generic writes below follow the TMA (async proxy) fills of the ring
            if constexpr (PERSIST) {
                // two epilogue warps per scheduler hide each other's TMEM-load latency: one chunk in registers at a time
                // (the second register buffer of the one-tile schedule would not fit under the 168-register cap of a
                // 320-thread CTA)
                uint32_t ra[32];
                while (c < nchunks) {
                    const bool has2 = c + 1 < nchunks;
                    const int c2 = c + 2 * NH;   // first chunk of this warp's next pair
                    tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), ra);
                    tc_wait_ld();
                    if (!has2) release_acc();    // an unpaired last chunk (BN = 160): everything is in registers
                    chunk(ra, c, has2 ? c + 1 : c2);
                    if (has2) {
                        tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), ra);
                        tc_wait_ld();
                        if (c2 >= nchunks) release_acc();   // every column this warp owns has left TMEM
                        chunk(ra, c + 1, c2);
                    }
                    c = c2;
                }
            } else {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(taddr, ra);
                for (; c < nchunks; c += 2) {
                    tc_wait_ld();  // ra = chunk c
                    const bool has2 = c + 1 < nchunks;
                    if (has2) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), rb);
                    chunk(ra, c, c + 1);
                    if (has2) {
                        tc_wait_ld();  // rb = chunk c+1
                        if (c + 2 < nchunks) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 2) * 32), ra);
                        chunk(rb, c + 1, c + 2);
                    }
                }
            }
            if (PERSIST && 2 * half >= nchunks) release_acc();   // nothing to drain in this tile (BN = 64)
        } else {
            // GEGLU: tile columns [0, BN/2) = value, [BN/2, BN) = gate; output width N/2.  Steps of 16 output columns
            // (16 value + 16 gate accumulators), double-buffered; four steps fill one 64-column staging group.
            // BN is 128 or 256 and N % BN == 0 (gcb_geglu_tile_n), so every group is full.
            const int half_bn = p.BN / 2;
            const int nsteps = half_bn / 16;                    // 4 or 8
            const int t0 = n_tile * p.BN;                       // packed-row offset of this tile (bias index)
            const int ocol0 = n_tile * half_bn;                 // first output column of this tile
            auto step = [&](const uint32_t (&rv)[16], const uint32_t (&rg)[16], int s) {
                if (p.epi && (s & 3) == 0) acquire_buffer();
                float val[16], gate[16], o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    val[j] = __uint_as_float(rv[j]);
                    gate[j] = __uint_as_float(rg[j]);
                }
                if (p.bias) {
                    const uint4* bvp = reinterpret_cast<const uint4*>(p.bias + t0 + s * 16);
                    const uint4* bgp = reinterpret_cast<const uint4*>(p.bias + t0 + half_bn + s * 16);
                    add8(val, bvp[0]);
                    add8(val + 8, bvp[1]);
                    add8(gate, bgp[0]);
                    add8(gate + 8, bgp[1]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = (val[j] * gate[j]) * gelu_erf_phi(gate[j]);
                const uint32_t h0 = pack_half2(o[0], o[1]), h1 = pack_half2(o[2], o[3]), h2 = pack_half2(o[4], o[5]),
                               h3 = pack_half2(o[6], o[7]), h4 = pack_half2(o[8], o[9]), h5 = pack_half2(o[10], o[11]),
                               h6 = pack_half2(o[12], o[13]), h7 = pack_half2(o[14], o[15]);
                if (p.epi) {
                    const uint32_t buf = stage_buf + (uint32_t)(cur_buf * 4096) + (uint32_t)(lane * 128);
                    const int j0 = (s & 3) * 2;                 // 16-byte piece of the 128-byte row
                    st_shared_v4(buf + (uint32_t)(((j0) ^ (lane & 7)) << 4), h0, h1, h2, h3);
                    st_shared_v4(buf + (uint32_t)(((j0 + 1) ^ (lane & 7)) << 4), h4, h5, h6, h7);
                    if ((s & 3) == 3) flush_group(ocol0 + (s - 3) * 16);
                } else if (row_ok) {
                    uint4* op = reinterpret_cast<uint4*>(p.y + m * p.ldy + ocol0 + s * 16);
                    op[0] = make_uint4(h0, h1, h2, h3);
                    op[1] = make_uint4(h4, h5, h6, h7);
                }
            };
            mbar_wait(smem_u32(&acc_full_bar[acc]), acc_par);
            tc_fence_after();
            if (!PERSIST && p.epi) fence_proxy_async_smem();
            uint32_t va[16], ga[16], vb[16], gb[16];
            // this warp's steps: the groups of four (s .. s+3), s = 4*half, 4*half + 4*NH, ...
            auto next_step = [](int s) { return (s & 3) == 3 ? s + 1 + 4 * (NH - 1) : s + 1; };
            int s = 4 * half;
            if (s < nsteps) {
                tmem_ld_32x32b_x16(taddr + (uint32_t)(s * 16), va);
                tmem_ld_32x32b_x16(taddr + (uint32_t)(half_bn + s * 16), ga);
            } else {
                release_acc();
            }
            while (s < nsteps) {
                tc_wait_ld();  // va / ga = step s (even)
                const int s1 = s + 1, s2 = next_step(s1);
                tmem_ld_32x32b_x16(taddr + (uint32_t)(s1 * 16), vb);
                tmem_ld_32x32b_x16(taddr + (uint32_t)(half_bn + s1 * 16), gb);
                step(va, ga, s);
                tc_wait_ld();  // vb / gb = step s+1
                if (s2 < nsteps) {
                    tmem_ld_32x32b_x16(taddr + (uint32_t)(s2 * 16), va);
                    tmem_ld_32x32b_x16(taddr + (uint32_t)(half_bn + s2 * 16), ga);
                } else {
                    release_acc();
                }
                step(vb, gb, s1);
                s = s2;
            }
        }
        }  // tile loop
        if constexpr (PEER) {
            // the peers' copies (bulk stores and the per-thread stores of an unpaired chunk) must be complete and visible
            // system-wide before the kernel that follows in the stream raises this rank's flag
            if (p.epi && issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __threadfence_system();
        } else {
            if (p.epi && issuer) tma_store_wait_read<0>();  // smem must outlive the bulk stores that read it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols * NACC);
}

// Which schedule a shape gets when nobody forces one.  Measured on every layer shape of a denoising step
// (profiles/r3a_gemm_persist_table.txt: 92.2 ms of one-tile launches -> 84.0 ms all persistent -> 81.0 ms per-shape best):
// the persistent schedule wins wherever a CTA's life is dominated by set-up, first-load latency and drain (short K:
// K = 320 GEGLU 546 -> 810 TFLOP/s, qkv 528 -> 754) and loses where two co-resident one-tile CTAs fill the tensor pipe
// better than one persistent CTA walking two tiles (1 < waves < 2 with long K: the 8x8 / 16x16 levels, up to -40 %).
bool persistent_by_shape(int ntiles, int nkb, int bn) {
    const double waves = (double)ntiles / gcb_sm_count();
    if (bn == 256) return waves > 1.0;
    if (bn == 160) return (waves > 1.0 && nkb <= 20) || (waves >= 2.0 && nkb <= 90) || waves >= 6.0;
    return false;   // BN = 128 / 64 are only chosen for problems of about one wave
}

int choose_bn(int M, int N, int act) {
    if (act == GCB_ACT_GEGLU) return gcb_geglu_tile_n(N);
    const int cands[4] = {256, 160, 128, 64};
    int best = 64;
    long long best_pad = -1;
    static int max_bn = -1;
    if (max_bn < 0) {
        const char* e = getenv("GCB_GEMM_MAX_BN");
        max_bn = e ? atoi(e) : 256;
    }
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        if (bn > max_bn) continue;
        const long long pad = (long long)gcb_cdiv(N, bn) * bn;
        if (best_pad < 0 || pad < best_pad) {
            best_pad = pad;
            best = bn;
        }
    }
    // small problems: prefer more CTAs over bigger tiles
    const long long mt = gcb_cdiv(M, BM);
    while (best > 64 && mt * gcb_cdiv(N, best) < gcb_sm_count() && (N % (best / 2) == 0 || best == 160)) {
        best = best == 160 ? 64 : best / 2;
    }
    return best;
}

}  // namespace

extern "C" int gcb_geglu_tile_n(int Cout) {
    static int forced = -1;  // GCB_GEGLU_BN=128: tuning knob (the weight packing follows this function, so it is consistent)
    if (forced < 0) {
        const char* e = getenv("GCB_GEGLU_BN");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 128) return 128;
    return ((Cout / 2) % 128 == 0) ? 256 : 128;
}

// perm[r_packed] = source row in the diffusers [value(4C) ; gate(4C)] projection
extern "C" int gcb_geglu_pack_rows(int Cout, int32_t* h_perm) {
    GCB_CHECK_ARG(Cout > 0 && Cout % 128 == 0, "GEGLU Cout=%d must be a multiple of 128", Cout);
    const int bn = gcb_geglu_tile_n(Cout), half = bn / 2, n_out = Cout / 2;
    for (int t = 0; t < Cout / bn; ++t)
        for (int j = 0; j < half; ++j) {
            h_perm[t * bn + j] = t * half + j;
            h_perm[t * bn + half + j] = n_out + t * half + j;
        }
    return GCB_OK;
}

int gcb_gemm_tc_supported(int B, int H, int W, int Cin, int Cout, int ksize, int act) {
    if (Cout % 8 != 0 || Cin % 8 != 0) return 0;
    if (act == GCB_ACT_GEGLU && Cout % 128 != 0) return 0;
    if (ksize == 1) return 1;
    if (ksize != 3 || Cin % BK != 0) return 0;
    // pixel box: W must tile 128 (or be tiled by it), H must be a multiple of the box height
    if (W >= BM) return (W % BM == 0);
    if (BM % W != 0) return 0;
    const int bh = (H * W >= BM) ? BM / W : H;
    if (H % bh != 0) return 0;
    if (H * W < BM && BM % (H * W) != 0) return 0;
    return 1;
}

int gcb_gemm_tc_launch_peers(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                             const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                             int direct_epilogue, void* const* peer_y, int n_peer, int peer_col_min, cudaStream_t stream);

int gcb_gemm_tc_launch(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                       const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                       int direct_epilogue, cudaStream_t stream) {
    return gcb_gemm_tc_launch_peers(x, w, bias, rowvec, rowvec_ld, residual, y, B, H, W, Cin, Cout, ksize, act,
                                    direct_epilogue, nullptr, 0, 0, stream);
}

namespace {
// Schedule of the calling thread's next launches: -1 = by shape (product default), 0 = one tile per CTA, 1 = persistent
// wherever it is built (TMA-store epilogue, no peers).  Set by gcb_conv2d_nhwc_fwd from its impl argument right before
// it launches (A/B tools), or process-wide through GCB_GEMM_PERSIST.  Per thread: concurrent callers do not interfere.
thread_local int g_schedule = -1;
}  // namespace
void gcb_gemm_tc_set_schedule(int schedule) { g_schedule = schedule; }

// peer_y[0..n_peer): the same output tensor in the peers' memory (every staged tile is stored there as well).  Needs the
// TMA-store epilogue with complete 64-column groups: Cout % 64 == 0, no GEGLU, no direct epilogue.  Only output columns
// >= peer_col_min (a multiple of 64) travel to the peers.
int gcb_gemm_tc_launch_peers(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                             const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                             int direct_epilogue, void* const* peer_y, int n_peer, int peer_col_min, cudaStream_t stream) {
    GemmTcParams p;
    memset(&p, 0, sizeof(p));
    const long long M = (long long)B * H * W;
    const int taps = ksize * ksize;
    const int Ktot = taps * Cin;
    p.M = (int)M;
    p.N = Cout;
    p.mode = (ksize == 3) ? 1 : 0;
    p.H = H;
    p.W = W;
    p.HW = H * W;
    p.BN = choose_bn((int)M, Cout, act);
    p.n_tiles = gcb_cdiv(Cout, p.BN);
    p.ntiles = p.n_tiles * (int)gcb_cdiv(M, BM);
    p.kb_per_tap = (ksize == 3) ? Cin / BK : gcb_cdiv(Cin, BK);
    p.nkb = taps * p.kb_per_tap;
    const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
    p.ldy = (act == GCB_ACT_GEGLU) ? Cout / 2 : Cout;
    // default: 32x64 output tiles staged in shared memory + TMA store; GCB_GEMM_TCGEN05_DIRECT keeps the round-1
    // per-thread row stores for A/B measurements
    p.epi = (!direct_epilogue && p.ldy >= 64) ? 1 : 0;
    static int env_schedule = -2;
    if (env_schedule == -2) {
        const char* e = getenv("GCB_GEMM_PERSIST");
        env_schedule = e ? atoi(e) : -1;
    }
    const int schedule = g_schedule >= 0 ? g_schedule : env_schedule;
    bool persist = p.epi && n_peer == 0;
    if (persist && schedule == 0) persist = false;
    if (persist && schedule < 0) persist = persistent_by_shape(p.ntiles, p.nkb, p.BN);
    size_t smem;
    if (persist) {
        // one CTA per SM: the ring takes what 227 KB leave after the epilogue staging (8 warps x 1 or 2 x 4 KB); it is
        // NOT clipped to the K blocks of one tile - the producer runs ahead into the next tile
        const int avail = 227 * 1024 - 1024 - 512;   // alignment slack, static barriers
        p.stage_bufs = 2;
        p.stages = (avail - 8 * 2 * 4096) / stage_bytes;
        if (p.stages < 4) {
            p.stage_bufs = 1;
            p.stages = (avail - 8 * 4096) / stage_bytes;
        }
        if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
        p.stage_off = (uint32_t)(p.stages * stage_bytes);
        smem = (size_t)p.stages * stage_bytes + (size_t)8 * p.stage_bufs * 4096 + 1024;
    } else {
        int budget = 110 * 1024;  // two CTAs per SM: one tile's epilogue overlaps the other's main loop
        if (const char* e = getenv("GCB_GEMM_SMEM_KB")) budget = atoi(e) * 1024;
        p.stages = budget / stage_bytes;
        if (p.stages < 2) p.stages = 2;
        if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
        if (p.stages > p.nkb) p.stages = p.nkb < 2 ? 2 : p.nkb;
        smem = (size_t)p.stages * stage_bytes + 1024;
    }
    p.tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
    p.idesc = make_idesc_f16(BM, p.BN, 0, 0);
    p.bias = (const __half*)bias;
    p.rowvec = (const __half*)rowvec;
    p.rowvec_ld = rowvec_ld;
    p.residual = (const __half*)residual;
    p.y = (__half*)y;
    p.act = act;

    CUtensorMap tmA, tmB, tmY;
    int rc;
    if (p.mode == 0) {
        const uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)M};
        const uint64_t strides[1] = {(uint64_t)Cin * 2};
        const uint32_t box[2] = {BK, BM};
        rc = gcb_encode_tma(&tmA, x, 2, dims, strides, box, 1);
    } else {
        p.bw = W >= BM ? BM : W;
        p.bh = (H * W >= BM) ? (BM / p.bw) : H;
        p.bb = (H * W >= BM) ? 1 : BM / (H * W);
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
        const uint32_t box[4] = {BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bb};
        rc = gcb_encode_tma(&tmA, x, 4, dims, strides, box, 1);
    }
    if (rc != GCB_OK) return rc;
    {
        const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)Cout};
        const uint64_t strides[1] = {(uint64_t)Ktot * 2};
        const uint32_t box[2] = {BK, (uint32_t)p.BN};
        rc = gcb_encode_tma(&tmB, w, 2, dims, strides, box, 1);
        if (rc != GCB_OK) return rc;
    }
    {   // output: [M, ldy] fp16, 32-row x 64-column boxes, 128B swizzle; the box is clipped at M and ldy by the hardware
        const uint64_t dims[2] = {(uint64_t)p.ldy, (uint64_t)M};
        const uint64_t strides[1] = {(uint64_t)p.ldy * 2};
        const uint32_t box[2] = {64, 32};
        if (p.epi) {
            rc = gcb_encode_tma(&tmY, y, 2, dims, strides, box, 1);
            if (rc != GCB_OK) return rc;
        } else {
            tmY = tmB;
        }
    }
    static unsigned long long configured = 0;   // one bit per device ordinal
    if (gcb_first_use_on_device(configured)) {
        GCB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        GCB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        GCB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    }
    const dim3 grid(persist ? (unsigned)(p.ntiles < gcb_sm_count() ? p.ntiles : gcb_sm_count()) : (unsigned)p.ntiles);
    if (n_peer > 0) {
        GCB_CHECK_ARG(n_peer <= MAX_PEER_MAPS && peer_y, "at most %d peer outputs", MAX_PEER_MAPS);
        GCB_CHECK_ARG(p.epi && act != GCB_ACT_GEGLU && Cout % 64 == 0,
                      "fused all-gather needs the TMA-store epilogue with full 64-column groups (Cout=%d)", Cout);
        PeerMaps pm;
        memset(&pm, 0, sizeof(pm));
        pm.n = n_peer;
        pm.col_min = peer_col_min;
        GCB_CHECK_ARG(peer_col_min >= 0 && peer_col_min % 64 == 0, "peer_col_min must be a multiple of 64");
        const uint64_t dims[2] = {(uint64_t)p.ldy, (uint64_t)M};
        const uint64_t strides[1] = {(uint64_t)p.ldy * 2};
        const uint32_t box[2] = {64, 32};
        for (int q = 0; q < n_peer; ++q) {
            rc = gcb_encode_tma(&pm.m[q], peer_y[q], 2, dims, strides, box, 1);
            if (rc != GCB_OK) return rc;
            pm.y[q] = (__half*)peer_y[q];
        }
        gemm_tc_kernel<true, false><<<grid, 192, smem, stream>>>(tmA, tmB, tmY, p, pm);
    } else if (persist) {
        NoPeers np{0};
        gemm_tc_kernel<false, true><<<grid, 320, smem, stream>>>(tmA, tmB, tmY, p, np);
    } else {
        NoPeers np{0};
        gemm_tc_kernel<false, false><<<grid, 192, smem, stream>>>(tmA, tmB, tmY, p, np);
    }
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
