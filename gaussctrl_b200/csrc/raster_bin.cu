// Tile binning of the projected Gaussians - the "cumsum -> map_gaussian_to_intersects -> sort -> get_tile_bin_edges" half of
// gsplat 0.1.3's rasterize_gaussians (call sites gc_model.py:174-186, :191-202) - without a host round trip.
//
// NOT gsplat's "sort M 64-bit (tile|depth) keys":
//   1. the N Gaussians are ordered by depth once: 4 stable 8-bit radix passes over N (key, id) pairs;
//   2. intersections (M ~ 4-10 N) are emitted in that order, so inside every tile they are already depth-ordered;
//   3. ONE stable radix pass by tile id groups them.  The tile histogram that pass needs is accumulated while emitting,
//      and its exclusive scan IS the tile-bin table, so no "find bin edges" pass over M exists.
// The result equals a stable sort by (tile << 32 | depth bits) with ties by Gaussian id - the order the oracle defines.
//
// Every radix pass is ONE kernel ("onesweep"): block-local stable ranks from per-bit warp ballots + per-warp counters, block
// offsets from a decoupled look-back over per-block digit counts (one 32-bit word per (block, digit): 2 flag bits +
// 30 count bits), blocks ordered by an atomic ticket so a block only ever waits on blocks that are already running.
// The number of intersections M never visits the host: it is produced by a single-pass scan (same look-back scheme),
// lives in a device word, and M-dependent kernels are launched for the caller's capacity and exit early above M.
//
// Launches per view: 1 memset + depth histogram + digit scan + 4 passes + gather-scan + emit + tile offsets + 1 pass
// = 11 (round 1: 22 kernels + 2 host synchronisations for the same work).
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int BLOCK = 16;
constexpr uint32_t FLAG_AGG = 1u << 30, FLAG_INC = 2u << 30, VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive prefix of this block's `total` over all blocks with a smaller ticket; publishes the inclusive value.
// One thread per chain (the count and its flag share a word, so no fence is needed).  The walk fetches a window of
// four predecessors with independent loads per round trip instead of one dependent load per predecessor.
__device__ __forceinline__ uint32_t lookback(uint32_t* state, long long stride, int bid, uint32_t total) {
    if (bid == 0) {
        st_relaxed(state, total | FLAG_INC);
        return 0;
    }
    st_relaxed(state + (long long)bid * stride, total | FLAG_AGG);
    uint32_t excl = 0;
    int p = bid - 1;
    while (true) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = p - k >= 0 ? ld_relaxed(state + (long long)(p - k) * stride) : FLAG_INC;
        bool done = false;
        int used = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (done || used != k) continue;
            if ((v[k] >> 30) == 0) continue;       // not published yet: poll again from this predecessor
            excl += v[k] & VAL_MASK;
            used = k + 1;
            if (v[k] & FLAG_INC) done = true;
        }
        if (done) break;
        p -= used;
    }
    st_relaxed(state + (long long)bid * stride, (excl + total) | FLAG_INC);
    return excl;
}

// single chain walked by a whole warp: 32 predecessors per round trip (the scan kernel's chain is ~250 blocks long)
__device__ __forceinline__ uint32_t lookback_warp(uint32_t* state, int bid, uint32_t total) {
    const int lane = threadIdx.x & 31;
    if (bid == 0) {
        if (lane == 0) st_relaxed(state, total | FLAG_INC);
        return 0;
    }
    if (lane == 0) st_relaxed(state + bid, total | FLAG_AGG);
    uint32_t excl = 0;
    int p = bid - 1;
    while (true) {
        const int idx = p - lane;
        const uint32_t v = idx >= 0 ? ld_relaxed(state + idx) : FLAG_INC;
        const unsigned ready = __ballot_sync(0xffffffffu, (v >> 30) != 0);
        const unsigned inc = __ballot_sync(0xffffffffu, (v & FLAG_INC) != 0);
        const int first_inc = inc ? __ffs(inc) - 1 : 31;
        const unsigned need = first_inc == 31 ? 0xffffffffu : ((2u << first_inc) - 1);
        if ((ready & need) != need) continue;
        uint32_t c = lane <= first_inc ? (v & VAL_MASK) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        excl += c;
        if (inc) break;
        p -= 32;
    }
    if (lane == 0) st_relaxed(state + bid, (excl + total) | FLAG_INC);
    return excl;
}

__device__ __forceinline__ int f2i_sat(float v) { return __float2int_rz(v); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(hi, max(lo, v)); }
__device__ __forceinline__ void tile_bbox(float x, float y, float radius, int tbx, int tby, int& x0, int& x1, int& y0,
                                          int& y1) {
    const float blk = (float)BLOCK;
    const float tcx = __fdiv_rn(x, blk), tcy = __fdiv_rn(y, blk), tr = __fdiv_rn(radius, blk);
    x0 = clampi(f2i_sat(__fsub_rn(tcx, tr)), 0, tbx);
    x1 = clampi(f2i_sat(__fadd_rn(__fadd_rn(tcx, tr), 1.0f)), 0, tbx);
    y0 = clampi(f2i_sat(__fsub_rn(tcy, tr)), 0, tby);
    y1 = clampi(f2i_sat(__fadd_rn(__fadd_rn(tcy, tr), 1.0f)), 0, tby);
}

// ------------------------------------------------------------------------------------------ block scan helper
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    const int r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------ depth digit histograms
// depths are >= 0, so their bit patterns order like unsigned integers: four 8-bit digits, one read of the keys
__global__ void __launch_bounds__(256) depth_hist_kernel(const uint32_t* __restrict__ keys, int N,
                                                         uint32_t* __restrict__ hist /* [4][256] */) {
    __shared__ uint32_t s_h[4 * 256];
    for (int i = threadIdx.x; i < 1024; i += 256) s_h[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256) {
        const uint32_t k = keys[i];
        atomicAdd(&s_h[k & 255], 1u);
        atomicAdd(&s_h[256 + ((k >> 8) & 255)], 1u);
        atomicAdd(&s_h[512 + ((k >> 16) & 255)], 1u);
        atomicAdd(&s_h[768 + (k >> 24)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

// in-place exclusive scan of gridDim.x independent arrays of `bins` (<= 2048) counters
__global__ void __launch_bounds__(1024) digit_scan_kernel(uint32_t* __restrict__ hist, int bins) {
    __shared__ int s_warp[33];
    uint32_t* h = hist + (long long)blockIdx.x * bins;
    const int i0 = threadIdx.x * 2;
    const int a = i0 < bins ? (int)h[i0] : 0, b = i0 + 1 < bins ? (int)h[i0 + 1] : 0;
    int total;
    const int ex = block_exclusive_scan(a + b, s_warp, total);
    if (i0 < bins) h[i0] = (uint32_t)ex;
    if (i0 + 1 < bins) h[i0 + 1] = (uint32_t)(ex + a);
}

// ------------------------------------------------------------------------------------------ one-kernel stable radix pass
constexpr int RX_WARPS = 8;
constexpr int RX_MAX_BINS = 2048;
constexpr int RX_BATCHES_DEPTH = 8, RX_BATCHES_TILE = 16;   // keys per block = 8 warps x batches x 32 = 2048 / 4096

// keys/vals [n] -> keys_out/vals_out ordered by digit (keys >> shift) & (bins-1), stable.  vals == nullptr: the value is
// the element's index (first depth pass).  n_dev != nullptr: the element count is min(*n_dev, n_cap) (tile pass: M lives
// on the device).  digit_base [bins]: exclusive scan of the global digit histogram.  state [blocks][bins] zeroed, ticket 0.
// Ranks inside a 32-key batch come from one ballot per digit bit (peers = lanes whose digit agrees in every bit) -
// `match.any` costs one iteration per DISTINCT value in the warp, ~30 for random digits (measured: 2x the whole pass).
template <int RX_BATCHES>
__global__ void __launch_bounds__(256) radix_pass_kernel(const uint32_t* __restrict__ keys,
                                                         const int32_t* __restrict__ vals,
                                                         const int32_t* __restrict__ n_dev, long long n_cap, int shift,
                                                         int nbits, const uint32_t* __restrict__ digit_base,
                                                         uint32_t* __restrict__ state, uint32_t* __restrict__ ticket,
                                                         uint32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
    constexpr int RX_PER_WARP = RX_BATCHES * 32, RX_TILE = RX_WARPS * RX_PER_WARP;
    extern __shared__ int s_cnt[];  // [RX_WARPS][bins]: per-warp digit counts, then running write bases
    __shared__ int s_bid;
    const int bins = 1 << nbits;
    if (threadIdx.x == 0) s_bid = (int)atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RX_WARPS * bins; i += 256) s_cnt[i] = 0;
    __syncthreads();
    const int bid = s_bid;
    long long n = n_cap;
    if (n_dev) n = min((long long)*n_dev, n_cap);
    const long long b0 = (long long)bid * RX_TILE;
    if (b0 >= n) return;  // blocks past the device-side count (uniform: nobody waits on them)
    const uint32_t mask = (uint32_t)bins - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* cnt_w = s_cnt + warp * bins;
    const long long w0 = b0 + (long long)warp * RX_PER_WARP;
    uint32_t key[RX_BATCHES];
    int32_t val[RX_BATCHES];
    int lrank[RX_BATCHES];
#pragma unroll
    for (int it = 0; it < RX_BATCHES; ++it) {
        const long long i = w0 + it * 32 + lane;
        const bool act = i < n;
        key[it] = act ? keys[i] : 0u;
        val[it] = act ? (vals ? vals[i] : (int32_t)i) : 0;
    }
#pragma unroll
    for (int it = 0; it < RX_BATCHES; ++it) {
        const long long i = w0 + it * 32 + lane;
        const bool act = i < n;
        const unsigned am = __ballot_sync(0xffffffffu, act);
        lrank[it] = -1;
        const uint32_t bin = (key[it] >> shift) & mask;
        unsigned peers = am;
        for (int b = 0; b < nbits; ++b) {
            const bool bit = (bin >> b) & 1u;
            const unsigned vote = __ballot_sync(0xffffffffu, act && bit);
            peers &= bit ? vote : ~vote;
        }
        if (act) {
            const int rank = __popc(peers & ((1u << lane) - 1));
            const int prior = cnt_w[bin];
            __syncwarp(am);
            if (rank == 0) cnt_w[bin] = prior + __popc(peers);
            lrank[it] = prior + rank;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += 256) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) {
            const int c = s_cnt[w * bins + b];
            s_cnt[w * bins + b] = run;
            run += c;
        }
        const uint32_t base = digit_base[b] + lookback(state + b, bins, bid, (uint32_t)run);
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) s_cnt[w * bins + b] += (int)base;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RX_BATCHES; ++it) {
        if (lrank[it] >= 0) {
            const int pos = cnt_w[(key[it] >> shift) & mask] + lrank[it];
            if (keys_out) keys_out[pos] = key[it];
            vals_out[pos] = val[it];
        }
    }
}

// ------------------------------------------------------------------------------------------ single-pass (gather +) scan
constexpr int GS_T = 512, GS_IPT = 8, GS_TILE = GS_T * GS_IPT;  // 4096

// out[i] = inclusive sum over j <= i of in[idx ? idx[j] : j].  total_out (nullable): [0] = grand total (clamped to
// INT32_MAX-free range by the caller's sizes), [1] = 1 when the total exceeds `cap` (cap < 0: no check).
__global__ void __launch_bounds__(GS_T) gather_scan_kernel(const int32_t* __restrict__ in, const int32_t* __restrict__ idx,
                                                           int N, int32_t* __restrict__ out, uint32_t* __restrict__ state,
                                                           uint32_t* __restrict__ ticket, int32_t* __restrict__ total_out,
                                                           long long cap) {
    __shared__ int s_warp[33];
    __shared__ int s_bid;
    __shared__ uint32_t s_excl;
    if (threadIdx.x == 0) s_bid = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int bid = s_bid;
    const long long base = (long long)bid * GS_TILE + (long long)threadIdx.x * GS_IPT;
    int x[GS_IPT];
    int v = 0;
#pragma unroll
    for (int j = 0; j < GS_IPT; ++j) {
        x[j] = 0;
        if (base + j < N) x[j] = in[idx ? idx[base + j] : base + j];
        v += x[j];
    }
    int total;
    const int ex = block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x < 32) {
        const uint32_t e = lookback_warp(state, bid, (uint32_t)total);
        if (threadIdx.x == 0) s_excl = e;
    }
    __syncthreads();
    int run = (int)s_excl + ex;
#pragma unroll
    for (int j = 0; j < GS_IPT; ++j) {
        run += x[j];
        if (base + j < N) out[base + j] = run;
    }
    if (total_out && threadIdx.x == 0 && (long long)(bid + 1) * GS_TILE >= N) {
        const long long m = (long long)s_excl + total;
        total_out[0] = (int32_t)m;
        total_out[1] = (cap >= 0 && m > cap) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------ emit + tile histogram
constexpr int EM_T = 256, EM_CHUNKS = 4, EM_TILE = EM_T * EM_CHUNKS;  // 1024 Gaussians per block (6-7 blocks per SM)
constexpr int EM_SMEM_TILES = 8192;

// Intersections of the depth-ordered Gaussians: tile ids (the radix keys) and Gaussian ids, written at the positions the
// scan assigned.  Warp-cooperative: a warp takes 32 consecutive Gaussians and walks their concatenated output range 32
// positions at a time - each lane finds the Gaussian owning its position with a 5-step search over the lanes' end
// offsets (shuffles) - so the stores are fully coalesced and the work does not depend on how many tiles one Gaussian
// covers.  The tile histogram is aggregated per block in shared memory (ntiles <= 8192) before it reaches L2.
__global__ void __launch_bounds__(EM_T) emit_isects_kernel(const float* __restrict__ xys, const int32_t* __restrict__ radii,
                                                           const int32_t* __restrict__ sorted_ids,
                                                           const int32_t* __restrict__ cum_sorted, int N, int tbx, int tby,
                                                           long long cap, uint32_t* __restrict__ tile_of,
                                                           int32_t* __restrict__ gid_of, uint32_t* __restrict__ tile_hist) {
    extern __shared__ uint32_t s_hist[];
    const int ntiles = tbx * tby;
    const bool use_smem = ntiles <= EM_SMEM_TILES;
    if (use_smem) {
        for (int i = threadIdx.x; i < ntiles; i += EM_T) s_hist[i] = 0;
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < EM_CHUNKS; ++c) {
        const long long i0 = (long long)blockIdx.x * EM_TILE + ((long long)c * (EM_T / 32) + warp) * 32;
        if (i0 >= N) break;
        const long long i = i0 + lane;
        int g = 0, x0 = 0, y0 = 0, w = 0, end = 0;
        if (i < N) {
            g = sorted_ids[i];
            end = cum_sorted[i];
            const int rad = radii[g];
            if (rad > 0) {
                int x1, y1;
                const float2 xy = reinterpret_cast<const float2*>(xys)[g];
                tile_bbox(xy.x, xy.y, (float)rad, tbx, tby, x0, x1, y0, y1);
                w = x1 - x0;
            }
        }
        // offsets are non-decreasing along the warp; lanes past N repeat the last valid end
        const int last_valid = (int)min(31ll, (long long)N - 1 - i0);
        end = __shfl_sync(0xffffffffu, end, min(lane, last_valid));
        int start = __shfl_up_sync(0xffffffffu, end, 1);
        if (lane == 0) start = i0 > 0 ? cum_sorted[i0 - 1] : 0;
        const int warp_start = __shfl_sync(0xffffffffu, start, 0), warp_end = __shfl_sync(0xffffffffu, end, 31);
        for (int pos0 = warp_start; pos0 < warp_end; pos0 += 32) {
            const int pos = pos0 + lane;
            int lo = 0, hi = 31;    // first lane whose end offset exceeds pos
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5) {
                const int mid = (lo + hi) >> 1;
                const int e = __shfl_sync(0xffffffffu, end, mid);
                if (e > pos)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            const int o = min(lo, 31);
            const int so = __shfl_sync(0xffffffffu, start, o), wo = __shfl_sync(0xffffffffu, w, o);
            const int xo = __shfl_sync(0xffffffffu, x0, o), yo = __shfl_sync(0xffffffffu, y0, o);
            const int go = __shfl_sync(0xffffffffu, g, o);
            if (pos < warp_end && pos < cap && wo > 0) {
                const int local = pos - so;
                const int row = local / wo;
                const uint32_t t = (uint32_t)((yo + row) * tbx + xo + (local - row * wo));
                tile_of[pos] = t;
                gid_of[pos] = go;
                if (use_smem)
                    atomicAdd(&s_hist[t], 1u);
                else
                    atomicAdd(&tile_hist[t], 1u);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < ntiles; i += EM_T)
            if (s_hist[i]) atomicAdd(&tile_hist[i], s_hist[i]);
    }
}

// tile_hist [ntiles] -> tile_bins [ntiles][2] = (start, end) of every tile's run, plus the digit tables of the tile
// pass(es): one pass (lo_bits == 0): digit_base[t] = start[t]; two passes: histograms of the low / high digit.
__global__ void __launch_bounds__(1024) tile_offsets_kernel(const uint32_t* __restrict__ tile_hist, int ntiles,
                                                            int32_t* __restrict__ tile_bins, uint32_t* __restrict__ digit_lo,
                                                            uint32_t* __restrict__ digit_hi, int lo_bits) {
    __shared__ int s_warp[33];
    int carry = 0;
    for (int c0 = 0; c0 < ntiles; c0 += 1024) {
        const int t = c0 + threadIdx.x;
        const int v = t < ntiles ? (int)tile_hist[t] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, total);
        if (t < ntiles) {
            tile_bins[2 * t] = carry + ex;
            tile_bins[2 * t + 1] = carry + ex + v;
            if (lo_bits == 0) {
                digit_lo[t] = (uint32_t)(carry + ex);
            } else if (v) {
                atomicAdd(&digit_lo[t & ((1 << lo_bits) - 1)], (uint32_t)v);
                atomicAdd(&digit_hi[t >> lo_bits], (uint32_t)v);
            }
        }
        carry += total;
    }
}

// test/diagnostic: the 64-bit keys gsplat would have sorted, rebuilt from the result
__global__ void isect_keys_kernel(const int32_t* __restrict__ tile_bins, int ntiles, const int32_t* __restrict__ gids,
                                  const float* __restrict__ depths, int64_t* __restrict__ keys) {
    const int t = blockIdx.x;
    if (t >= ntiles) return;
    const int s = tile_bins[2 * t], e = tile_bins[2 * t + 1];
    for (int i = s + threadIdx.x; i < e; i += blockDim.x)
        keys[i] = ((int64_t)t << 32) | (int64_t)(uint32_t)__float_as_uint(depths[gids[i]]);
}

inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct BinLayout {
    // zeroed region
    size_t off_tickets, off_depth_hist, off_tile_hist, off_digits, off_scan_state, off_depth_state, off_tile_state, zero_bytes;
    // scratch
    size_t off_keyA, off_keyB, off_valA, off_valB, off_cum, off_tileA, off_gidA, off_tileB, off_gidB, total;
    int depth_blocks, scan_blocks, tile_blocks, tile_bits, lo_bits, hi_bits, tile_bins_pass;
};

BinLayout bin_layout(int N, long long cap, int ntiles) {
    BinLayout L;
    L.depth_blocks = (int)cdivll(N, RX_WARPS * 32 * RX_BATCHES_DEPTH);
    L.scan_blocks = (int)cdivll(N, GS_TILE);
    L.tile_blocks = (int)cdivll(cap, RX_WARPS * 32 * RX_BATCHES_TILE);
    int bits = 1;
    while ((1 << bits) < ntiles) ++bits;
    L.tile_bits = bits;
    if (bits <= 11) {
        L.lo_bits = 0;
        L.hi_bits = bits;
    } else {
        L.lo_bits = bits / 2;
        L.hi_bits = bits - L.lo_bits;
    }
    L.tile_bins_pass = 1 << (L.lo_bits ? (L.lo_bits > L.hi_bits ? L.lo_bits : L.hi_bits) : bits);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += align256(bytes);
        return at;
    };
    L.off_tickets = take(16 * sizeof(uint32_t));
    L.off_depth_hist = take(4 * 256 * sizeof(uint32_t));
    L.off_tile_hist = take((size_t)ntiles * sizeof(uint32_t));
    L.off_digits = take((size_t)2 * RX_MAX_BINS * sizeof(uint32_t));
    L.off_scan_state = take((size_t)L.scan_blocks * sizeof(uint32_t));
    L.off_depth_state = take((size_t)4 * L.depth_blocks * 256 * sizeof(uint32_t));
    L.off_tile_state = take((size_t)(L.lo_bits ? 2 : 1) * L.tile_blocks * L.tile_bins_pass * sizeof(uint32_t));
    L.zero_bytes = o;
    L.off_keyA = take((size_t)N * 4);
    L.off_keyB = take((size_t)N * 4);
    L.off_valA = take((size_t)N * 4);
    L.off_valB = take((size_t)N * 4);
    L.off_cum = take((size_t)N * 4);
    L.off_tileA = take((size_t)cap * 4);
    L.off_gidA = take((size_t)cap * 4);
    L.off_tileB = take(L.lo_bits ? (size_t)cap * 4 : 0);
    L.off_gidB = take(L.lo_bits ? (size_t)cap * 4 : 0);
    L.total = o;
    return L;
}

int configure_radix_smem() {
    static thread_local int configured_dev = -1;
    int dev = 0;
    GCB_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        GCB_CUDA(cudaFuncSetAttribute(radix_pass_kernel<RX_BATCHES_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      RX_WARPS * RX_MAX_BINS * (int)sizeof(int)));
        // 32-64 KB of counters per block: ask for the largest shared-memory carve-out so 6-7 blocks fit on an SM
        // (the default carve-out admitted 3: measured 36 % warp occupancy)
        GCB_CUDA(cudaFuncSetAttribute(radix_pass_kernel<RX_BATCHES_TILE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
        configured_dev = dev;
    }
    return GCB_OK;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" size_t gcb_bin_gaussians_workspace_bytes(int N, long long isect_capacity, int tile_bx, int tile_by) {
    if (N <= 0 || isect_capacity <= 0 || tile_bx <= 0 || tile_by <= 0) return 0;
    return bin_layout(N, isect_capacity, tile_bx * tile_by).total;
}

extern "C" int gcb_bin_gaussians(const float* xys, const float* depths, const int32_t* radii, const int32_t* num_tiles_hit,
                                 int N, int tile_bx, int tile_by, long long isect_capacity, int32_t* gaussian_ids,
                                 int32_t* tile_bins, int32_t* isect_count, int64_t* isect_keys, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    GCB_CHECK_ARG(xys && depths && radii && num_tiles_hit && gaussian_ids && tile_bins && isect_count && workspace,
                  "null pointer");
    GCB_CHECK_ARG(N > 0, "N must be positive");
    const long long ntiles_ll = (long long)tile_bx * tile_by;
    GCB_CHECK_ARG(tile_bx > 0 && tile_by > 0 && ntiles_ll <= (1 << 20), "tile grid %dx%d unsupported", tile_bx, tile_by);
    GCB_CHECK_ARG(isect_capacity > 0 && isect_capacity < (1ll << 30), "isect_capacity out of range (1 .. 2^30-1)");
    const int ntiles = (int)ntiles_ll;
    const BinLayout L = bin_layout(N, isect_capacity, ntiles);
    if (workspace_bytes < L.total) {
        gcb_set_error("bin workspace too small: %zu < %zu", workspace_bytes, L.total);
        return GCB_ERR_WORKSPACE;
    }
    int rc = configure_radix_smem();
    if (rc != GCB_OK) return rc;
    char* ws = (char*)workspace;
    GCB_CUDA(cudaMemsetAsync(ws, 0, L.zero_bytes, ST));
    uint32_t* tickets = (uint32_t*)(ws + L.off_tickets);
    uint32_t* depth_hist = (uint32_t*)(ws + L.off_depth_hist);
    uint32_t* tile_hist = (uint32_t*)(ws + L.off_tile_hist);
    uint32_t* digit_lo = (uint32_t*)(ws + L.off_digits);
    uint32_t* digit_hi = digit_lo + RX_MAX_BINS;
    uint32_t* scan_state = (uint32_t*)(ws + L.off_scan_state);
    uint32_t* depth_state = (uint32_t*)(ws + L.off_depth_state);
    uint32_t* tile_state = (uint32_t*)(ws + L.off_tile_state);
    uint32_t* keyA = (uint32_t*)(ws + L.off_keyA);
    uint32_t* keyB = (uint32_t*)(ws + L.off_keyB);
    int32_t* valA = (int32_t*)(ws + L.off_valA);
    int32_t* valB = (int32_t*)(ws + L.off_valB);
    int32_t* cum = (int32_t*)(ws + L.off_cum);
    uint32_t* tileA = (uint32_t*)(ws + L.off_tileA);
    int32_t* gidA = (int32_t*)(ws + L.off_gidA);

    // 1. depth order: four stable 8-bit passes over (depth bits, id)
    const uint32_t* dk = reinterpret_cast<const uint32_t*>(depths);
    const int sms = gcb_sm_count();
    depth_hist_kernel<<<min(2 * sms, gcb_cdiv(N, 256)), 256, 0, ST>>>(dk, N, depth_hist);
    digit_scan_kernel<<<4, 1024, 0, ST>>>(depth_hist, 256);
    const size_t smem8 = (size_t)RX_WARPS * 256 * sizeof(int);
    const size_t dstate = (size_t)L.depth_blocks * 256;
    radix_pass_kernel<RX_BATCHES_DEPTH><<<L.depth_blocks, 256, smem8, ST>>>(
        dk, nullptr, nullptr, N, 0, 8, depth_hist, depth_state, tickets + 0, keyA, valA);
    radix_pass_kernel<RX_BATCHES_DEPTH><<<L.depth_blocks, 256, smem8, ST>>>(
        keyA, valA, nullptr, N, 8, 8, depth_hist + 256, depth_state + dstate, tickets + 1, keyB, valB);
    radix_pass_kernel<RX_BATCHES_DEPTH><<<L.depth_blocks, 256, smem8, ST>>>(
        keyB, valB, nullptr, N, 16, 8, depth_hist + 512, depth_state + 2 * dstate, tickets + 2, keyA, valA);
    radix_pass_kernel<RX_BATCHES_DEPTH><<<L.depth_blocks, 256, smem8, ST>>>(
        keyA, valA, nullptr, N, 24, 8, depth_hist + 768, depth_state + 3 * dstate, tickets + 3, nullptr, valB);
    const int32_t* sorted_ids = valB;
    // 2. offsets of every depth-ordered Gaussian's intersections; M and the overflow flag stay on the device
    gather_scan_kernel<<<L.scan_blocks, GS_T, 0, ST>>>(num_tiles_hit, sorted_ids, N, cum, scan_state, tickets + 4, isect_count,
                                                       isect_capacity);
    // 3. emit (tile, id) pairs in depth order + tile histogram
    const size_t em_smem = ntiles <= EM_SMEM_TILES ? (size_t)ntiles * sizeof(uint32_t) : 0;
    emit_isects_kernel<<<gcb_cdiv(N, EM_TILE), EM_T, em_smem, ST>>>(xys, radii, sorted_ids, cum, N, tile_bx, tile_by,
                                                                    isect_capacity, tileA, gidA, tile_hist);
    // 4. tile bins = exclusive scan of the tile histogram (also the digit table of a single tile pass)
    tile_offsets_kernel<<<1, 1024, 0, ST>>>(tile_hist, ntiles, tile_bins, digit_lo, digit_hi, L.lo_bits);
    // 5. group by tile: one stable pass (<= 2048 tiles) or two
    if (L.lo_bits == 0) {
        const int bins = 1 << L.tile_bits;
        radix_pass_kernel<RX_BATCHES_TILE><<<L.tile_blocks, 256, (size_t)RX_WARPS * bins * sizeof(int), ST>>>(
            tileA, gidA, isect_count, isect_capacity, 0, L.tile_bits, digit_lo, tile_state, tickets + 5, nullptr,
            gaussian_ids);
    } else {
        uint32_t* tileB = (uint32_t*)(ws + L.off_tileB);
        int32_t* gidB = (int32_t*)(ws + L.off_gidB);
        const int bins_lo = 1 << L.lo_bits, bins_hi = 1 << L.hi_bits;
        digit_scan_kernel<<<1, 1024, 0, ST>>>(digit_lo, bins_lo);
        digit_scan_kernel<<<1, 1024, 0, ST>>>(digit_hi, bins_hi);
        radix_pass_kernel<RX_BATCHES_TILE><<<L.tile_blocks, 256, (size_t)RX_WARPS * bins_lo * sizeof(int), ST>>>(
            tileA, gidA, isect_count, isect_capacity, 0, L.lo_bits, digit_lo, tile_state, tickets + 5, tileB, gidB);
        radix_pass_kernel<RX_BATCHES_TILE><<<L.tile_blocks, 256, (size_t)RX_WARPS * bins_hi * sizeof(int), ST>>>(
            tileB, gidB, isect_count, isect_capacity, L.lo_bits, L.hi_bits, digit_hi,
            tile_state + (size_t)L.tile_blocks * L.tile_bins_pass, tickets + 6, nullptr, gaussian_ids);
    }
    if (isect_keys) isect_keys_kernel<<<ntiles, 128, 0, ST>>>(tile_bins, ntiles, gaussian_ids, depths, isect_keys);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" size_t gcb_scan_workspace_bytes(int N) { return align256((size_t)cdivll(N > 0 ? N : 1, GS_TILE) * 4) + 256; }

extern "C" int gcb_cumsum_i32(const int32_t* in, int32_t* out, int N, void* workspace, size_t workspace_bytes,
                              void* stream) {
    GCB_CHECK_ARG(in && out && workspace, "null pointer");
    if (N <= 0) return GCB_OK;
    const size_t need = gcb_scan_workspace_bytes(N);
    if (workspace_bytes < need) {
        gcb_set_error("scan workspace too small");
        return GCB_ERR_WORKSPACE;
    }
    GCB_CUDA(cudaMemsetAsync(workspace, 0, need, ST));
    uint32_t* ticket = (uint32_t*)workspace;
    uint32_t* state = ticket + 64;
    gather_scan_kernel<<<(unsigned)cdivll(N, GS_TILE), GS_T, 0, ST>>>(in, nullptr, N, out, state, ticket, nullptr, -1);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}
