"""Denoising loops of the hot path on device tensors: DDIM inversion (render_reverse's `pipe(...)` call,
gaussctrl/gc_pipeline.py:141-145) and cross-view-attention editing (edit_images' `pipe(...)` call, :206-219).

Two edit schedules produce the same latents (reference-view rows never depend on chunk-view rows):
  * "reference"  - the reference's own: every chunk denoises [R refs + c views] together, refs recomputed per chunk;
  * "refs_once"  - B200-first: per DDIM step the R reference views are denoised once (their per-layer K/V recorded),
                   then every other view batch reads that K/V; 2.3x fewer network evaluations at R=4, c=3, and
                   view batches become independent (this is what shards across GPUs, see parallel.py).
Each step's network evaluation + CFG combine + DDIM update is captured in a CUDA graph and replayed per step
(timestep and scheduler coefficients live in device buffers)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops
from .diffusion import (AttnPlan, SD15Denoiser, cached_crossview_plan, literal_crossview_plan, vanilla_plan)
from .sd15_spec import DDIMTables


class _GraphedStep:
    """One denoising step for a fixed batch shape as a CUDA graph: x <- ddim(x, eps(x, t, cond))."""

    def __init__(self, den: SD15Denoiser, n_lat: int, cfg: bool, guidance: float, plan: AttnPlan, hw: int,
                 use_graph: bool, share: Optional["_GraphedStep"] = None):
        dev = den.dev
        self.den, self.cfg, self.guidance, self.plan, self.n_lat = den, cfg, guidance, plan, n_lat
        B = 2 * n_lat if cfg else n_lat
        if share is not None:
            # a second graph over the SAME state (double-buffered reference pass: only the recorded K/V differ)
            self.x, self.cond, self.t, self.coef = share.x, share.cond, share.t, share.coef
        else:
            self.x = torch.zeros((n_lat, hw, hw, 4), dtype=torch.float16, device=dev)
            self.cond = torch.zeros((B, hw, hw, den.ch[0]), dtype=torch.float16, device=dev)
            self.t = torch.zeros((B,), dtype=torch.float32, device=dev)
            self.coef = torch.zeros((4,), dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph
        self.launches = 0

    def _body(self) -> None:
        n = self.n_lat
        xin = torch.cat([self.x, self.x], dim=0) if self.cfg else self.x
        eps = self.den.eps(xin, self.t, self.cond, self.plan)
        if self.cfg:
            ops.cfg_ddim_step(eps[:n], eps[n:], self.x, self.guidance, self.coef, out=self.x)
        else:
            ops.cfg_ddim_step(eps, None, self.x, 0.0, self.coef, out=self.x)

    def set_cond(self, cond_emb: torch.Tensor) -> None:
        """cond_emb [n_lat, hw, hw, C0] (the conditioning image is the same for both CFG halves)."""
        n = self.n_lat
        self.cond[:n].copy_(cond_emb)
        if self.cfg:
            self.cond[n:].copy_(cond_emb)

    def run(self, t, coefs) -> None:
        """t / coefs: DEVICE tensors (1 and 4 floats, rows of EditEngine._schedule_tables) - copied device-to-device, so
        the host never waits for the stream (it enqueues the whole loop ahead; needed for the two-stream overlap of the
        reference pass with the view batches) - or plain Python numbers (host copy, synchronous)."""
        if isinstance(t, torch.Tensor):
            self.t.copy_(t.reshape(1).expand_as(self.t))
            self.coef.copy_(coefs)
        else:
            self.t.fill_(float(t))
            self.coef.copy_(torch.tensor(coefs, dtype=torch.float32), non_blocking=False)
        if not self.use_graph:
            self._body()
            return
        if self.graph is None:
            l0 = ops.LAUNCHES[0]
            x_save = self.x.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()  # warm-up outside capture (lazy packing of weights, cudaFuncSetAttribute calls)
            torch.cuda.current_stream().wait_stream(s)
            self.launches = ops.LAUNCHES[0] - l0
            self.x.copy_(x_save)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()
            self.x.copy_(x_save)
        else:
            ops.LAUNCHES[0] += self.launches
        self.graph.replay()


class _ShardedRefStep:
    """Reference pass of one DDIM step with its 2R CFG rows split across ranks (parallel.py).  Every rank holds the R
    reference latents; it evaluates the network on its own rows (per-layer K/V all-gather inside), all-gathers the
    eps rows (32 KB each) and applies the CFG combine + DDIM update to all R latents."""

    def __init__(self, den: SD15Denoiser, R: int, guidance: float, hw: int, world: int, rank: int, gather,
                 ref_frames: Sequence[int], use_graph: bool, share: Optional["_ShardedRefStep"] = None):
        from . import parallel as par
        dev = den.dev
        self.den, self.R, self.guidance, self.world, self.rank = den, R, guidance, world, rank
        self.per = par.padded_rows_per_rank(R, world)
        self.rows = par.ref_row_partition(R, world, rank)
        rows = self.rows + [self.rows[-1] if self.rows else 0] * (self.per - len(self.rows))  # pad with a real row
        self.lat_of_row = torch.tensor([g % R for g in rows], dtype=torch.long, device=dev)
        src = par.sharded_ref_src_index(R, world, rank, ref_frames)
        src = src + [src[-1] if src else [-1] * (1 + len(ref_frames))] * (self.per - len(src))
        K = len(ref_frames)
        self.rec: Dict[str, torch.Tensor] = {}
        self.plan = AttnPlan(torch.tensor(src, dtype=torch.int32, device=dev), [0.6] + [0.4 / K] * K,
                             [0.0] + [1.0 / K] * K, record_kv=self.rec,
                             text_index=torch.tensor([[g // R] for g in rows], dtype=torch.int32, device=dev),
                             gather=gather)
        if share is not None:
            self.x, self.cond_all, self.t, self.coef = share.x, share.cond_all, share.t, share.coef
        else:
            self.x = torch.zeros((R, hw, hw, 4), dtype=torch.float16, device=dev)
            self.cond_all = torch.zeros((R, hw, hw, den.ch[0]), dtype=torch.float16, device=dev)
            self.t = torch.zeros((self.per,), dtype=torch.float32, device=dev)
            self.coef = torch.zeros((4,), dtype=torch.float32, device=dev)
        self.gather = gather
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0

    def set_cond(self, cond_emb: torch.Tensor) -> None:
        self.cond_all.copy_(cond_emb)

    def _body(self) -> None:
        # the peers may still be reading last step's gathered K/V in their view batches: barrier before re-writing it
        self.gather.begin_pass()
        xin = self.x[self.lat_of_row]
        cond = self.cond_all[self.lat_of_row]
        eps = self.den.eps(xin, self.t, cond, self.plan)
        eps_all = self.gather("eps", eps.contiguous())          # [world*per, hw, hw, 4]: 32 KB per row
        R = self.R
        ops.cfg_ddim_step(eps_all[:R], eps_all[R:2 * R], self.x, self.guidance, self.coef, out=self.x)

    run = _GraphedStep.run


class _ParityGather:
    """The K/V exchange seen by one of the two double-buffered reference-pass graphs: same exchange object, region /
    flag-slot names tagged with the buffer parity."""

    def __init__(self, gather, parity: int):
        self.inner, self.tag = gather, f"#p{parity}"
        self.graph_capturable = bool(getattr(gather, "graph_capturable", False))

    def __call__(self, layer: str, local: torch.Tensor) -> torch.Tensor:
        return self.inner(layer + self.tag, local)

    def begin_pass(self) -> None:
        self.inner.begin_pass()

    def check(self) -> None:
        self.inner.check()

    def linear_gather(self, layer: str, x: torch.Tensor, w: torch.Tensor, bias=None, local_cols: int = 0):
        fn = getattr(self.inner, "linear_gather", None)
        return None if fn is None else fn(layer + self.tag, x, w, bias, local_cols)


class EditEngine:
    def __init__(self, denoiser: SD15Denoiser, tables: Optional[DDIMTables] = None, use_graphs: bool = True,
                 overlap_refs: Optional[bool] = None):
        self.den = denoiser
        self.dev = denoiser.dev
        self.tables = tables or DDIMTables()
        self.use_graphs = use_graphs
        # refs_once schedule: run the reference pass one step AHEAD of the view batches on its own stream (double-buffered
        # K/V).  The reference trajectory never depends on the other views, its pass is small (2R CFG rows, or 2R/N rows
        # per rank) and leaves most SMs idle - on several GPUs it also waits on 23 exchanges per step; overlapped, the view
        # batches fill those gaps.  GCB_REF_OVERLAP=0 turns it off.
        import os
        self.overlap_refs = (os.environ.get("GCB_REF_OVERLAP", "1") != "0") if overlap_refs is None else overlap_refs
        self._ref_stream: Optional[torch.cuda.Stream] = None
        self._steps: Dict[tuple, object] = {}  # captured graphs are reused across calls (static buffers inside)

    def _schedule_tables(self, ts: Sequence[int], coef_fn) -> tuple:
        """Device tables of the timesteps [n] and of the four scheduler coefficients per step [n, 4] (one upload)."""
        t_tab = torch.tensor([float(t) for t in ts], dtype=torch.float32).to(self.dev, non_blocking=True)
        c_tab = torch.tensor([list(coef_fn(t)) for t in ts], dtype=torch.float32).to(self.dev, non_blocking=True)
        return t_tab, c_tab

    # ------------------------------------------------------------------------------------------ helpers
    def _latents_in(self, z_nchw: torch.Tensor) -> torch.Tensor:
        return ops.nchw_to_nhwc(z_nchw.to(device=self.dev, dtype=torch.float16).contiguous())

    def _cond_emb(self, disparity_nchw: torch.Tensor, batch: int = 8) -> torch.Tensor:
        """[V,3,H,W] disparity images -> ControlNet conditioning embeddings [V,h,w,C0] (step-invariant)."""
        outs = []
        for i in range(0, disparity_nchw.shape[0], batch):
            d = ops.nchw_to_nhwc(disparity_nchw[i:i + batch].to(device=self.dev, dtype=torch.float16).contiguous())
            outs.append(self.den.controlnet_cond(d))
        return torch.cat(outs, dim=0)

    # ------------------------------------------------------------------------------------------ inversion
    @torch.no_grad()
    def invert(self, z0: torch.Tensor, disparity: torch.Tensor, prompt_embed: torch.Tensor, S: int,
               batch: int = 8) -> torch.Tensor:
        """DDIM inversion of V views (gc_pipeline.py:141-145: guidance_scale=0 => no CFG, vanilla attention).
        z0 [V,4,h,w], disparity [V,3,H,W], prompt_embed [1,77,768] -> z_T [V,4,h,w] fp16.
        Views are independent, so they are batched `batch` at a time instead of the reference's batch 1."""
        V, hw = z0.shape[0], z0.shape[-1]
        self.den.set_prompts(prompt_embed)
        cond = self._cond_emb(disparity)
        x_all = self._latents_in(z0)
        out = torch.empty_like(x_all)
        ts = self.tables.inverse_timesteps(S)
        t_tab, c_tab = self._schedule_tables(ts, lambda t: self.tables.inverse_step_coefs(t, S))
        for i in range(0, V, batch):
            n = min(batch, V - i)
            key = ("invert", n, hw)
            if key not in self._steps:
                self._steps[key] = _GraphedStep(self.den, n, False, 0.0, vanilla_plan(n, self.dev), hw, self.use_graphs)
            st = self._steps[key]
            st.x.copy_(x_all[i:i + n])
            st.set_cond(cond[i:i + n])
            for si in range(len(ts)):
                st.run(t_tab[si], c_tab[si])
            out[i:i + n].copy_(st.x)
        return ops.nhwc_to_nchw(out)

    # ------------------------------------------------------------------------------------------ editing
    @torch.no_grad()
    def edit_reference_schedule(self, latents: torch.Tensor, disparity: torch.Tensor, pos_embed: torch.Tensor,
                                neg_embed: torch.Tensor, S: int, guidance: float, num_ref: int,
                                ref_frames: Sequence[int] = (0, 1, 2, 3), stop_after: Optional[int] = None) -> torch.Tensor:
        """One `pipe(...)` call of edit_images exactly as the reference batches it (gc_pipeline.py:206-219):
        latents [F,4,h,w] = [refs | chunk], CFG rows [uncond x F | cond x F].  Returns the final latents of the
        chunk rows [F-num_ref,4,h,w] (decode is the VAE's job)."""
        assert guidance > 1.0, "the reference assumes CFG doubling (utils.py:94)"
        F, hw = latents.shape[0], latents.shape[-1]
        self.den.set_prompts(torch.cat([neg_embed, pos_embed], dim=0))
        key = ("reference", F, hw, float(guidance), tuple(ref_frames))
        if key not in self._steps:
            self._steps[key] = _GraphedStep(self.den, F, True, guidance, literal_crossview_plan(F, self.dev, ref_frames),
                                            hw, self.use_graphs)
        st = self._steps[key]
        st.x.copy_(self._latents_in(latents))
        st.set_cond(self._cond_emb(disparity))
        ts = list(self.tables.timesteps(S))[:stop_after]  # stop_after: parity tests truncate long schedules
        t_tab, c_tab = self._schedule_tables(ts, lambda t: self.tables.step_coefs(t, S))
        for si in range(len(ts)):
            st.run(t_tab[si], c_tab[si])
        return ops.nhwc_to_nchw(st.x[num_ref:].contiguous())

    @torch.no_grad()
    def edit_refs_once(self, latents: torch.Tensor, disparity: torch.Tensor, ref_indices: Sequence[int],
                       pos_embed: torch.Tensor, neg_embed: torch.Tensor, S: int, guidance: float, view_batch: int = 4,
                       ref_frames: Sequence[int] = (0, 1, 2, 3), view_ids: Optional[Sequence[int]] = None,
                       dist_ctx: Optional[dict] = None, stop_after: Optional[int] = None) -> torch.Tensor:
        """Edit all V views: latents [V,4,h,w] (z_T of every view), disparity [V,3,H,W], ref_indices = the R reference
        view ids.  Per DDIM step: (1) the R references are denoised once, recording every self-attention layer's
        K/V; (2) the remaining views are denoised in batches of `view_batch`, reading that K/V.
        Reference views' own edited latents are rows of pass (1) (SURVEY §8a gotcha 6).
        `view_ids`: the subset of views this rank edits (multi-GPU sharding); default all.
        `dist_ctx` = {"world", "rank", "gather": parallel.PeerKVAllGather (our NVLink peer-memory kernels, captured in the
        reference pass's CUDA graph) or parallel.KVAllGather (NCCL, eager), "graph_refs": optional override}: the
        reference pass is
        sharded over the ranks' CFG rows with a per-layer K/V all-gather; result rows of views this rank does not own
        stay zero (parallel.gather_view_results assembles them)."""
        assert guidance > 1.0
        V, hw = latents.shape[0], latents.shape[-1]
        R = len(ref_indices)
        ids = list(range(V)) if view_ids is None else list(view_ids)
        non_ref = [v for v in ids if v not in set(ref_indices)]
        self.den.set_prompts(torch.cat([neg_embed, pos_embed], dim=0))
        x_all = self._latents_in(latents)
        need = sorted(set(non_ref) | set(ref_indices))
        cond_map = {v: c for v, c in zip(need, self._cond_emb(disparity[need]))}
        # (1) reference pass
        # balanced batches: the fewest batches of at most `view_batch` views, all of (almost) the same size, so the
        # padding of the last batch is at most nb-1 views (39 views, view_batch 12 -> 4 batches of 10, not 12+12+12+3)
        n_batches = max(1, -(-len(non_ref) // max(1, view_batch)))
        bsz = max(1, -(-len(non_ref) // n_batches))
        world = dist_ctx["world"] if dist_ctx else 1
        # P = 2: reference pass on its own stream, one step ahead of the views, K/V double-buffered by step parity
        P = 2 if (self.overlap_refs and self.use_graphs and self.dev.type == "cuda") else 1
        key = ("refs_once", R, bsz, hw, float(guidance), tuple(ref_frames), world, P)
        if key not in self._steps:
            refs_p: List[object] = []
            recs_p: List[Dict[str, torch.Tensor]] = []
            for par_i in range(P):
                share = refs_p[0] if refs_p else None
                if world > 1:
                    # our peer-memory exchange is plain kernels and is captured with the rest of the pass; the NCCL
                    # all-gather stays eager (capturing it deadlocked on the 2-GPU box in round 1)
                    capturable = bool(getattr(dist_ctx["gather"], "graph_capturable", False))
                    gather = _ParityGather(dist_ctx["gather"], par_i) if P > 1 else dist_ctx["gather"]
                    rs = _ShardedRefStep(self.den, R, guidance, hw, world, dist_ctx["rank"], gather, ref_frames,
                                         self.use_graphs and dist_ctx.get("graph_refs", capturable), share=share)
                    refs_p.append(rs)
                    recs_p.append(rs.rec)
                else:
                    rec_i: Dict[str, torch.Tensor] = {}
                    refs_p.append(_GraphedStep(self.den, R, True, guidance,
                                               literal_crossview_plan(R, self.dev, ref_frames, record_kv=rec_i), hw,
                                               self.use_graphs, share=share))
                    recs_p.append(rec_i)
            self._steps[key] = [refs_p, recs_p, [None] * P]
        ref_steps, recs, view_steps = self._steps[key]
        ref_step = ref_steps[0]
        ref_step.x.copy_(x_all[list(ref_indices)])
        ref_step.set_cond(torch.stack([cond_map[v] for v in ref_indices]))
        # (2) view batches (the last one is padded by repeating its final view so one graph serves all)
        batches: List[List[int]] = [non_ref[i:i + bsz] for i in range(0, len(non_ref), bsz)]
        x_views: List[torch.Tensor] = []
        conds: List[torch.Tensor] = []
        for b in batches:
            padded = b + [b[-1]] * (bsz - len(b))
            x_views.append(x_all[padded].clone())
            conds.append(torch.stack([cond_map[v] for v in padded]))
        ts = list(self.tables.timesteps(S))[:stop_after]
        t_tab, c_tab = self._schedule_tables(ts, lambda t: self.tables.step_coefs(t, S))
        main = torch.cuda.current_stream() if self.dev.type == "cuda" else None
        if P > 1:
            if self._ref_stream is None:
                self._ref_stream = torch.cuda.Stream(self.dev)
            ref_stream = self._ref_stream
            ref_stream.wait_stream(main)          # the reference latents / conditioning set above
            ev_ref = [torch.cuda.Event() for _ in ts]
            ev_view = [torch.cuda.Event() for _ in ts]
        for si in range(len(ts)):
            par_i = si % P
            if P > 1:
                with torch.cuda.stream(ref_stream):
                    if si >= P:
                        ref_stream.wait_event(ev_view[si - P])   # the views of step si-2 have finished reading this buffer
                    ref_steps[par_i].run(t_tab[si], c_tab[si])
                    ev_ref[si].record(ref_stream)
                main.wait_event(ev_ref[si])
            else:
                ref_steps[0].run(t_tab[si], c_tab[si])
            if batches and view_steps[par_i] is None:
                vplan = cached_crossview_plan(bsz, R, self.dev, recs[par_i], ref_frames)
                if world > 1:
                    from . import parallel as par
                    vplan.src_index = torch.tensor(par.view_src_index(bsz, R, world, ref_frames), dtype=torch.int32,
                                                   device=self.dev)
                view_steps[par_i] = _GraphedStep(self.den, bsz, True, guidance, vplan, hw, self.use_graphs,
                                                 share=view_steps[0])
            view_step = view_steps[par_i]
            for bi in range(len(batches)):
                if len(batches) > 1 or si == 0:
                    view_step.x.copy_(x_views[bi])
                    view_step.set_cond(conds[bi])
                view_step.run(t_tab[si], c_tab[si])
                if len(batches) > 1 or si == len(ts) - 1:
                    x_views[bi].copy_(view_step.x)
            if P > 1:
                ev_view[si].record(main)
        if P > 1:
            main.wait_stream(ref_stream)
        if world > 1:
            dist_ctx["gather"].check()   # one synchronisation per edit: did every wait of the K/V exchange complete?
        out = torch.zeros_like(x_all)
        for ri, v in enumerate(ref_indices):
            out[v].copy_(ref_step.x[ri])
        for b, xv in zip(batches, x_views):
            for j, v in enumerate(b):
                out[v].copy_(xv[j])
        return ops.nhwc_to_nchw(out)
