"""Multi-GPU sharding of the edit stage: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box,
gloo in the CPU tests) for the plumbing.

What shards:
  * the V views: independent once the reference views' K/V is known - each rank edits `shard_views(...)`;
  * the reference pass: its 2R CFG rows (uncond x R | cond x R) are split across ranks; every self-attention layer
    all-gathers the rows' fused q|k|v projections so that each reference row attends to all references and every
    rank ends up holding the complete reference K/V for its own views.  This is the ONE exchange step of the path
    ("NCCL all-gather of reference K/V", BASELINE.json north_star); nothing else communicates.
The reference itself has no multi-GPU code (pipe_device is hard-coded to 'cuda:0', gc_pipeline.py:96)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(V: int, world: int, rank: int, ref_indices: Sequence[int]) -> List[int]:
    """Non-reference views of this rank (round-robin over the non-reference views so shards differ by <= 1 view).
    Reference views are produced by the reference pass and reported by the rank that owns their CFG rows' result
    (rank 0 gathers them)."""
    refs = set(ref_indices)
    non_ref = [v for v in range(V) if v not in refs]
    return non_ref[rank::world]


def ref_decode_owner(R: int, world: int, rank: int) -> List[int]:
    """Which of the R reference views (positions in `ref_indices`) this rank decodes and reports.  Every rank holds all R
    reference latents after the reference pass, so any rank can; they are dealt from the LAST rank downwards because
    `shard_views` gives the low ranks the extra non-reference views."""
    return [i for i in range(R) if (world - 1 - i) % world == rank]


def broadcast_shape(shape: Sequence[int], device, group: Optional[dist.ProcessGroup] = None) -> Tuple[int, ...]:
    """Element-wise maximum of a small integer tuple over the ranks (a rank without views does not know the image
    size; it still has to take part in the gathers with correctly shaped empty contributions)."""
    t = torch.tensor(list(shape), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return tuple(int(v) for v in t.tolist())


def make_kv_gather(device, group: Optional[dist.ProcessGroup] = None, arena_bytes: Optional[int] = None):
    """The reference-K/V exchange of this process group.  On CUDA: PeerKVAllGather - our own push + flag kernels over
    NVLink peer memory (CUDA-graph capturable, gcb_allgather_ref_kv); if the peers' memory cannot be mapped (no IPC
    between the processes) or GCB_KV_GATHER=nccl is set: KVAllGather (torch.distributed / NCCL, runs eagerly)."""
    import os
    dev = torch.device(device)
    if dev.type == "cuda" and os.environ.get("GCB_KV_GATHER", "peer") != "nccl":
        try:
            return PeerKVAllGather(dev, group, arena_bytes or int(os.environ.get("GCB_PEER_ARENA_BYTES", 6 << 30)))
        except Exception as exc:   # e.g. cudaIpcOpenMemHandle refused in a sandbox: keep running on NCCL, say so
            import warnings
            warnings.warn(f"peer-memory K/V exchange unavailable ({type(exc).__name__}: {exc}); using NCCL all-gather")
    return KVAllGather(group)


def ref_row_partition(R: int, world: int, rank: int) -> List[int]:
    """Global CFG-row ids (0..2R-1, layout [uncond x R | cond x R]) of the reference pass owned by `rank`.
    Rows are dealt contiguously; when world > 2R the extra ranks own no reference row (they still take part in the
    all-gather with an empty contribution padded to the common size)."""
    n = 2 * R
    per = (n + world - 1) // world
    return [g for g in range(rank * per, min(n, (rank + 1) * per))]


def padded_rows_per_rank(R: int, world: int) -> int:
    return (2 * R + world - 1) // world


def sharded_ref_src_index(R: int, world: int, rank: int, ref_frames: Sequence[int] = (0, 1, 2, 3)) -> List[List[int]]:
    """src_index rows for this rank's reference rows when ALL K/V come from the gathered buffer (negative ids):
    source 0 = the row itself, sources 1.. = frames `ref_frames` of the row's CFG half.  Gathered row id of global
    row g is g itself (ranks contribute contiguous, padded blocks: see `gathered_row`)."""
    per = padded_rows_per_rank(R, world)
    rows = []
    for g in ref_row_partition(R, world, rank):
        half = g // R
        rows.append([-(gathered_row(g, per) + 1)] + [-(gathered_row(half * R + r, per) + 1) for r in ref_frames])
    return rows


def gathered_row(g: int, per: int) -> int:
    """Row of global reference row g inside the all-gathered [world*per, ...] buffer (contiguous deal => identity)."""
    return (g // per) * per + (g % per)


def view_src_index(Bv: int, R: int, world: int, ref_frames: Sequence[int] = (0, 1, 2, 3)) -> List[List[int]]:
    """src_index for a views-only batch [uncond x Bv | cond x Bv] reading reference K/V from the gathered buffer."""
    per = padded_rows_per_rank(R, world)
    rows = []
    for half in range(2):
        for f in range(Bv):
            rows.append([half * Bv + f] + [-(gathered_row(half * R + r, per) + 1) for r in ref_frames])
    return rows


class KVAllGather:
    """All-gather of the reference rows' fused q|k|v projection of one self-attention layer (torch.distributed: NCCL on
    GPUs, gloo in the CPU tests).  Not captured in CUDA graphs (`graph_capturable = False`).

    local [per, N, 3C] (rows beyond the rank's real rows are padding) -> gathered [world*per, N, 3C].  Output buffers
    are cached per layer so that captured CUDA graphs (and the view graphs that read them) see stable addresses."""

    graph_capturable = False

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buffers: Dict[str, torch.Tensor] = {}
        self.bytes = 0

    def begin_pass(self) -> None:   # NCCL collectives are ordered by the communicator: nothing to do
        return None

    def check(self) -> None:
        return None

    def __call__(self, layer: str, local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return local
        out = self.buffers.get(layer)
        shape = (self.world * local.shape[0],) + tuple(local.shape[1:])
        if out is None or tuple(out.shape) != shape:
            out = torch.empty(shape, dtype=local.dtype, device=local.device)
            self.buffers[layer] = out
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        self.bytes += out.numel() * out.element_size()
        return out


class _ArenaView:
    """Zero-copy torch view of a region of the peer arena (raw cudaMalloc memory) via __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerKVAllGather:
    """Same call contract as KVAllGather - `gather(layer, local [per, ...]) -> gathered [world*per, ...]` - on the
    peer-memory kernels of csrc/peer.cu.  Regions of the arena are assigned per layer name on first use (every rank
    issues the same sequence of calls, so offsets and flag slots agree without communication); the returned tensor
    aliases the local arena, i.e. later kernels read the gathered K/V in place.  `begin_pass()` is the cross-rank
    barrier that must precede re-writing regions the peers may still be reading (engine: once per sharded reference pass).
    Everything is enqueued on the current stream and can be captured in a CUDA graph."""
    graph_capturable = True

    def __init__(self, device, group: Optional[dist.ProcessGroup] = None, arena_bytes: int = 6 << 30):
        import ctypes
        from ._lib import check, lib
        self._lib, self._check = lib, check
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = torch.device(device)
        self.arena_bytes = int(arena_bytes)
        with torch.cuda.device(self.dev):
            h = ctypes.c_void_p()
            check(lib.gcb_handle_create(self.world, self.rank, self.arena_bytes, ctypes.byref(h)))
            self.handle = h
            mine = ctypes.create_string_buffer(64)
            check(lib.gcb_handle_ipc_export(h, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            ok = True
            try:
                for q in range(self.world):
                    if q != self.rank:
                        check(lib.gcb_handle_ipc_open(h, q, handles[q]))
            except Exception:
                ok = False
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=group)     # all ranks take the same branch
            if not all(flags):
                lib.gcb_handle_destroy(h)
                self.handle = None
                raise RuntimeError("cudaIpcOpenMemHandle failed on at least one rank")
        self.base = int(lib.gcb_handle_arena(self.handle))
        self.cursor = int(lib.gcb_handle_control_bytes())
        self.regions: Dict[str, Tuple[int, int, torch.Tensor, int]] = {}   # layer -> (offset, slot, view, bytes_per_rank)
        self.bytes = 0
        dist.barrier(group=group)

    def _region(self, layer: str, local):
        """`local`: this rank's block (a tensor, or just its shape): [rows, ...] fp16."""
        shape_l = tuple(local.shape) if isinstance(local, torch.Tensor) else tuple(local)
        ent = self.regions.get(layer)
        per_bytes = 2
        for d_ in shape_l:
            per_bytes *= int(d_)
        if ent is None or ent[3] != per_bytes or tuple(ent[2].shape[1:]) != shape_l[1:]:
            if isinstance(local, torch.Tensor) and local.dtype != torch.float16:
                raise TypeError("PeerKVAllGather carries fp16 tensors")
            if per_bytes % 16:
                raise ValueError(f"{layer}: {per_bytes} bytes per rank is not a multiple of 16")
            off = (self.cursor + 255) // 256 * 256
            total = per_bytes * self.world
            if off + total > self.arena_bytes:
                raise MemoryError(f"peer arena of {self.arena_bytes} bytes exhausted at layer {layer} "
                                  f"(set GCB_PEER_ARENA_BYTES)")
            if ent is not None:
                slot = ent[1]                      # same layer at a new shape: its flag slot carries over
            else:
                slot = 1 + len(self.regions)       # slot 0 is the pass barrier
            if slot >= 128:
                raise RuntimeError("more than 127 gathered buffers")
            shape = (self.world * shape_l[0],) + shape_l[1:]
            view = torch.as_tensor(_ArenaView(self.base + off, shape, "<f2"), device=self.dev)
            self.cursor = off + total
            ent = (off, slot, view, per_bytes)
            self.regions[layer] = ent
        return ent

    def begin_pass(self) -> None:
        self._check(self._lib.gcb_peer_barrier(self.handle, 0, torch.cuda.current_stream().cuda_stream))

    def __call__(self, layer: str, local: torch.Tensor) -> torch.Tensor:
        off, slot, view, per_bytes = self._region(layer, local)
        src = local.contiguous()
        self._check(self._lib.gcb_allgather_ref_kv(self.handle, off, src.data_ptr(), per_bytes, slot,
                                                   torch.cuda.current_stream().cuda_stream))
        self.bytes += per_bytes * self.world
        return view

    def linear_gather(self, layer: str, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      local_cols: int = 0):
        """Projection + exchange in ONE kernel (gcb_linear_allgather_fwd): y = x w^T for this rank's rows x [per, N, Cin],
        every output tile stored by the GEMM epilogue into all ranks' arenas.  -> (gathered [world*per, N, Cout], the
        local rows' view into it).  The first `local_cols` output columns (a multiple of 64) are NOT sent to the peers -
        the Q third of a q|k|v projection is read by its own rank only; the peers' copies of those columns are undefined.
        Falls back to GEMM + push (None) when the shape is outside the fused kernel's limits."""
        if self.world > 8 or w.shape[0] % 64 != 0 or x.dtype != torch.float16 or not x.is_contiguous():
            return None
        if local_cols % 64:
            local_cols = 0
        per, N, Cin = x.shape
        Cout = w.shape[0]
        off, slot, view, per_bytes = self._region(layer, (per, N, Cout))
        self._check(self._lib.gcb_linear_allgather_fwd(self.handle, x.data_ptr(), w.data_ptr(),
                                                       None if bias is None else bias.data_ptr(), per * N, Cin, Cout,
                                                       int(local_cols), off, slot,
                                                       torch.cuda.current_stream().cuda_stream))
        self.bytes += per_bytes * self.world
        return view, view[self.rank * per:(self.rank + 1) * per]

    def check(self) -> None:
        """Synchronous: raises if any wait of this rank timed out."""
        import ctypes
        torch.cuda.synchronize(self.dev)      # the flag is read with a blocking copy that does not order with torch's streams
        err = ctypes.c_int(0)
        self._check(self._lib.gcb_handle_error(self.handle, ctypes.byref(err)))
        if err.value:
            raise RuntimeError(f"peer all-gather timed out waiting on flag slot {err.value - 1}: a rank fell out of step")

    def close(self) -> None:
        if getattr(self, "handle", None) is not None:
            self.regions.clear()
            self._lib.gcb_handle_destroy(self.handle)
            self.handle = None


def gather_view_results(local: torch.Tensor, local_ids: Sequence[int], V: int, world: int,
                        group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Collect per-rank results [n_local, ...] into [V, ...] on every rank (64 KB per view for latents).  Ranks may
    hold different numbers of views: contributions are padded to the maximum."""
    if world == 1:
        out = torch.zeros((V,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out[list(local_ids)] = local
        return out
    n_max = torch.tensor([len(local_ids)], device=local.device)
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=group)
    n_max = int(n_max.item())
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: len(local_ids)] = local
    ids = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
    ids[: len(local_ids)] = torch.tensor(list(local_ids), dtype=torch.int64, device=local.device)
    all_vals = [torch.empty_like(pad) for _ in range(world)]
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(all_vals, pad, group=group)
    dist.all_gather(all_ids, ids, group=group)
    out = torch.zeros((V,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for vals, idv in zip(all_vals, all_ids):
        keep = idv >= 0
        out[idv[keep]] = vals[keep]
    return out
