// Reference-K/V exchange over NVLink peer memory ("ncclAllGather of ref K and V per self-attention layer per
// step" - the ONE collective of the path), as plain kernels so that the sharded reference pass can live inside a CUDA
// graph and no NCCL launch sits on the per-layer critical path.
//
// One process per GPU.  Every rank owns an ARENA (cudaMalloc) whose IPC handle the host side exchanges once
// (torch.distributed.all_gather_object); every rank maps all peers' arenas.  An all-gather of `bytes_per_rank` bytes at
// arena offset `off` is
//   push kernel   : rank r stores its block to arena[p] + off + r * bytes_per_rank on EVERY rank p (16-byte stores over
//                   NVLink; the local copy goes through the same loop), then fences at system scope;
//   signal kernel : (next in the stream, so the push has completed) thread p stores the slot's new epoch into
//                   flags[p][slot][r] of peer p and then spins until its own flags[r][slot][q] reached the epoch for every q.
// The epoch is a per-slot counter in the local arena, incremented by the signal kernel itself: a captured graph can be
// replayed any number of times.  Spins are bounded (~20 s of clock64): on timeout the kernel raises the handle's error
// flag and returns instead of hanging the GPU.
//
// Buffer reuse: the gathered block of step s is read by the view-batch kernels of step s; a peer may only overwrite it
// once every rank is past those reads.  The host side therefore opens each sharded reference pass with gcb_peer_barrier
// (same signal kernel on a dedicated slot).
#include <string.h>

#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_WORLD = 16;
constexpr int MAX_SLOTS = 128;
constexpr long long SPIN_TIMEOUT_CYCLES = 40ll * 1000 * 1000 * 1000;  // ~20 s at 1.9 GHz

struct Control {                       // at the head of every arena
    uint32_t flags[MAX_SLOTS][MAX_WORLD];   // flags[slot][q]: last epoch rank q signalled to me
    uint32_t epoch[MAX_SLOTS];              // my own epoch counter per slot
    uint32_t error;                         // != 0 after a spin timed out
    uint32_t pad[63];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct PeerPtrs {
    char* arena[MAX_WORLD];
};

// grid.y = destination rank, grid.x strides over the block
__global__ void __launch_bounds__(256) peer_push_kernel(PeerPtrs peers, size_t dst_off, const uint4* __restrict__ src,
                                                        size_t n16) {
    uint4* dst = reinterpret_cast<uint4*>(peers.arena[blockIdx.y] + dst_off);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) dst[i] = src[i];
    __threadfence_system();
}

__global__ void __launch_bounds__(32) peer_signal_wait_kernel(PeerPtrs peers, int world, int rank, int slot) {
    Control* me = reinterpret_cast<Control*>(peers.arena[rank]);
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = me->epoch[slot] + 1;
        me->epoch[slot] = s_epoch;
    }
    __syncwarp();
    const uint32_t e = s_epoch;
    const int q = threadIdx.x;
    if (q < world) {
        Control* other = reinterpret_cast<Control*>(peers.arena[q]);
        __threadfence_system();
        st_release_sys(&other->flags[slot][rank], e);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(&me->flags[slot][q]) - e) < 0) {
            if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
                me->error = 1u + (uint32_t)slot;
                break;
            }
        }
    }
    __threadfence_system();
}

}  // namespace

struct gcb_handle {
    int world, rank, device;
    size_t arena_bytes;
    char* arena[MAX_WORLD];   // [rank] = own allocation, others = IPC mappings
    bool opened[MAX_WORLD];
};

extern "C" size_t gcb_handle_control_bytes(void) { return (sizeof(Control) + 255) & ~(size_t)255; }

extern "C" int gcb_handle_create(int world, int rank, size_t arena_bytes, gcb_handle_t** out) {
    GCB_CHECK_ARG(out, "null out");
    GCB_CHECK_ARG(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "bad world=%d / rank=%d", world, rank);
    GCB_CHECK_ARG(arena_bytes >= gcb_handle_control_bytes(), "arena smaller than its control block");
    gcb_handle* h = new gcb_handle();
    memset(h, 0, sizeof(*h));
    h->world = world;
    h->rank = rank;
    h->arena_bytes = arena_bytes;
    GCB_CUDA(cudaGetDevice(&h->device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, arena_bytes);
    if (e != cudaSuccess) {
        delete h;
        gcb_set_error("cudaMalloc(%zu) for the peer arena failed: %s", arena_bytes, cudaGetErrorString(e));
        return GCB_ERR_CUDA;
    }
    h->arena[rank] = (char*)p;
    GCB_CUDA(cudaMemset(p, 0, gcb_handle_control_bytes()));
    *out = h;
    return GCB_OK;
}

extern "C" int gcb_handle_destroy(gcb_handle_t* h) {
    if (!h) return GCB_OK;
    for (int q = 0; q < h->world; ++q)
        if (q != h->rank && h->opened[q]) cudaIpcCloseMemHandle(h->arena[q]);
    if (h->arena[h->rank]) cudaFree(h->arena[h->rank]);
    delete h;
    return GCB_OK;
}

extern "C" void* gcb_handle_arena(gcb_handle_t* h) { return h ? (void*)h->arena[h->rank] : nullptr; }

extern "C" int gcb_handle_ipc_export(gcb_handle_t* h, unsigned char* out64) {
    GCB_CHECK_ARG(h && out64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t ipc;
    GCB_CUDA(cudaIpcGetMemHandle(&ipc, h->arena[h->rank]));
    memcpy(out64, &ipc, 64);
    return GCB_OK;
}

extern "C" int gcb_handle_ipc_open(gcb_handle_t* h, int peer, const unsigned char* in64) {
    GCB_CHECK_ARG(h && in64, "null pointer");
    GCB_CHECK_ARG(peer >= 0 && peer < h->world && peer != h->rank, "bad peer %d", peer);
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, in64, 64);
    void* p = nullptr;
    GCB_CUDA(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->arena[peer] = (char*)p;
    h->opened[peer] = true;
    return GCB_OK;
}

static int peers_ready(gcb_handle_t* h, PeerPtrs& P) {
    for (int q = 0; q < h->world; ++q) {
        if (!h->arena[q]) {
            gcb_set_error("peer arena %d not opened (gcb_handle_ipc_open)", q);
            return GCB_ERR_INVALID;
        }
        P.arena[q] = h->arena[q];
    }
    return GCB_OK;
}

extern "C" int gcb_peer_barrier(gcb_handle_t* h, int slot, void* stream) {
    GCB_CHECK_ARG(h, "null handle");
    GCB_CHECK_ARG(slot >= 0 && slot < MAX_SLOTS, "slot %d out of range", slot);
    PeerPtrs P;
    int rc = peers_ready(h, P);
    if (rc != GCB_OK) return rc;
    peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, h->world, h->rank, slot);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_allgather_ref_kv(gcb_handle_t* h, size_t arena_offset, const void* local_src, size_t bytes_per_rank,
                                    int slot, void* stream) {
    GCB_CHECK_ARG(h && local_src, "null pointer");
    GCB_CHECK_ARG(slot >= 0 && slot < MAX_SLOTS, "slot %d out of range", slot);
    GCB_CHECK_ARG(bytes_per_rank % 16 == 0 && arena_offset % 16 == 0 && ((uintptr_t)local_src) % 16 == 0,
                  "all-gather blocks must be 16-byte aligned");
    GCB_CHECK_ARG(arena_offset >= gcb_handle_control_bytes() &&
                      arena_offset + bytes_per_rank * (size_t)h->world <= h->arena_bytes,
                  "all-gather region [%zu, +%zu x %d) outside the arena (%zu bytes)", arena_offset, bytes_per_rank, h->world,
                  h->arena_bytes);
    PeerPtrs P;
    int rc = peers_ready(h, P);
    if (rc != GCB_OK) return rc;
    const size_t n16 = bytes_per_rank / 16;
    if (n16 > 0) {
        const int sms = gcb_sm_count();
        // W destinations share the grid: about two waves of 256-thread CTAs in total, at least one CTA per destination
        int bx = (int)((n16 + 255) / 256);
        const int cap = (4 * sms + h->world - 1) / h->world;
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        dim3 grid((unsigned)bx, (unsigned)h->world);
        peer_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, arena_offset + (size_t)h->rank * bytes_per_rank,
                                                                  (const uint4*)local_src, n16);
    }
    peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, h->world, h->rank, slot);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

int gcb_gemm_tc_supported(int B, int H, int W, int Cin, int Cout, int ksize, int act);
int gcb_gemm_tc_launch_peers(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                             const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                             int direct_epilogue, void* const* peer_y, int n_peer, int peer_col_min, cudaStream_t stream);

// y = x w^T (+ bias) for this rank's M rows, written by the GEMM's own epilogue into EVERY rank's arena at
// arena_offset + rank * M * Cout * 2 (the layout of gcb_allgather_ref_kv), then the flag exchange: one kernel computes and
// transfers, tile by tile, instead of GEMM -> push kernel.
extern "C" int gcb_linear_allgather_fwd(gcb_handle_t* h, const void* x, const void* w, const void* bias, int M, int Cin,
                                        int Cout, int peer_col_min, size_t arena_offset, int slot, void* stream) {
    GCB_CHECK_ARG(h && x && w, "null pointer");
    GCB_CHECK_ARG(slot >= 0 && slot < MAX_SLOTS, "slot %d out of range", slot);
    GCB_CHECK_ARG(M > 0 && Cin > 0 && Cout > 0 && Cout % 64 == 0, "bad GEMM shape M=%d Cin=%d Cout=%d (Cout %% 64 == 0)", M,
                  Cin, Cout);
    GCB_CHECK_ARG(peer_col_min >= 0 && peer_col_min < Cout && peer_col_min % 64 == 0, "peer_col_min=%d (multiple of 64, < Cout)",
                  peer_col_min);
    const size_t bytes_per_rank = (size_t)M * Cout * 2;
    GCB_CHECK_ARG(arena_offset % 128 == 0 && arena_offset >= gcb_handle_control_bytes() &&
                      arena_offset + bytes_per_rank * (size_t)h->world <= h->arena_bytes,
                  "all-gather region [%zu, +%zu x %d) outside the arena (%zu bytes)", arena_offset, bytes_per_rank, h->world,
                  h->arena_bytes);
    if (!gcb_gemm_tc_supported(1, 1, M, Cin, Cout, 1, GCB_ACT_NONE)) {
        gcb_set_error("fused linear + all-gather: shape M=%d Cin=%d Cout=%d not supported by the tcgen05 GEMM", M, Cin, Cout);
        return GCB_ERR_UNSUPPORTED;
    }
    PeerPtrs P;
    int rc = peers_ready(h, P);
    if (rc != GCB_OK) return rc;
    const size_t off = arena_offset + (size_t)h->rank * bytes_per_rank;
    void* peers[MAX_WORLD];
    int n = 0;
    for (int q = 0; q < h->world; ++q)
        if (q != h->rank) peers[n++] = P.arena[q] + off;
    GCB_CHECK_ARG(n <= 7, "fused all-gather supports up to 8 ranks");
    rc = gcb_gemm_tc_launch_peers(x, w, bias, nullptr, 0, nullptr, P.arena[h->rank] + off, 1, 1, M, Cin, Cout, 1, GCB_ACT_NONE,
                                  0, peers, n, peer_col_min, (cudaStream_t)stream);
    if (rc != GCB_OK) return rc;
    peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(P, h->world, h->rank, slot);
    GCB_LAUNCH_CHECK();
    return GCB_OK;
}

extern "C" int gcb_handle_error(gcb_handle_t* h, int* out) {
    GCB_CHECK_ARG(h && out, "null pointer");
    uint32_t v = 0;
    GCB_CUDA(cudaMemcpy(&v, h->arena[h->rank] + offsetof(Control, error), sizeof(v), cudaMemcpyDeviceToHost));
    *out = (int)v;
    return GCB_OK;
}
