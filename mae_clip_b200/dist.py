"""Global-batch contrastive loss across ranks (SURVEY.md section 8 e; not in the reference, which is
single-process - the semantics are "the reference loss, ``CLIP.py:34-43``, on the concatenated
batch").

One process per GPU.  Rank r owns rows [r*b, (r+1)*b) of the global (B, D) embeddings.

  forward   all-gather the local embeddings (NCCL over NVLink)            2 x B x D
            sweep 1 over the owned (b x B) strip  -> row/col LSE of S, row LSE of Z (owned rows)
            all-gather the three statistic vectors                         3 x B floats
            sweep 2 -> row_g, col_sum_p (owned rows) and the loss partial
            all-gather those + the partial                                  2 x B + W floats
  backward  sweep 3 -> dI_loc, dT_loc with EVERY term of the owned rows (strip of dS, transposed
            strip of dS, dZ + dZ^T by symmetry of Z): no gradient reduce-scatter is needed, only the
            small vectors above ever cross NVLink after the embedding all-gather.  (This is the
            OWN-ROWS form.  The peer transport's default from 4096 x 4096 logits per rank on is the
            STORED-WEIGHTS form: S once per rank, dT and the fp16 weight strip from the row half,
            this rank's contribution to every row of dI from the column half, then a peer-memory
            reduce-scatter of the partial dI - see ``PeerStep.backward`` and DESIGN.md section 5.)

The sweeps are the C-ABI phases ``mc_clip_stats / mc_clip_rowloss / mc_clip_bwd``.  The engine is
injectable so the collective choreography can be tested on CPU with gloo (tests supply a torch
engine; the product engine below is CUDA-only).

Two transports carry the exchange steps on GPUs:

  ``peer``  (default for the tcgen05 engines) our own kernels over mapped peer memory
            (``mae_clip_b200/peer.py``): the embedding all-gather is fused into the operand staging
            and the vectors are pushed straight into every peer's region - no collective-library
            call per step;
  ``nccl``  the all-gathers below through ``torch.distributed`` (any engine, any backend; also the
            path the CPU/gloo tests exercise).

``MAE_CLIP_TRANSPORT=nccl|peer`` overrides the choice.
"""
from __future__ import annotations

import os

import torch
from torch.autograd.function import once_differentiable
import torch.distributed as dist

from . import _lib
from ._lib import check, cur_stream, lib, ptr, require_cuda, workspace


class CudaStripEngine:
    """The product engine: the three sweeps through the C ABI on this rank's GPU."""

    def __init__(self, mode=None, group=None):
        from .functional import _mode
        self.mode = _mode(mode)
        self.group = group
        self.sparse = os.environ.get("MAE_CLIP_DENSE", "0") != "1"   # tile flags (see mc_clip_tile_flags_bytes)

    def _ws(self, b, B, D, dev):
        return workspace(lib().mc_clip_loss_workspace_bytes(b, B, D, self.mode), dev)

    def stats(self, I_all, T_all, b, row_offset, tau):
        require_cuda(I_all, T_all)
        B, D = I_all.shape
        dev = I_all.device
        out = torch.empty(4, b, device=dev, dtype=torch.float32)  # r, c, rz and sum_j P_ij S_ij (local only)
        planes = None
        with torch.cuda.device(dev):
            nb = lib().mc_clip_planes_bytes(B, D, self.mode)
            if nb:
                # every rank stages all B rows locally from the gathered fp32 embeddings
                planes = torch.empty(nb, device=dev, dtype=torch.uint8)
                check(lib().mc_clip_prepare(ptr(I_all), ptr(T_all), B, B, D, 0, self.mode, ptr(planes),
                                            cur_stream()), "mc_clip_prepare")
                ws, own_ws = self._ws(b, B, D, dev), None
            else:
                # The fp32 FMA engine (mode simt_fp32, or any D the tcgen05 engine does not cover) KEEPS its materialised
                # S / S^T / Z strips in the workspace from this sweep to the gradient sweep: that state must survive
                # whatever runs between forward and backward (another loss, a head, an MAE op all use the shared
                # scratch cache), so it lives in a tensor of its own that travels with the autograd context.
                own_ws = torch.empty(lib().mc_clip_loss_workspace_bytes(b, B, D, self.mode), device=dev, dtype=torch.uint8)
                ws = own_ws
            nf = lib().mc_clip_tile_flags_bytes(b, B, D, self.mode) if self.sparse else 0
            flags_raw = torch.empty(nf, device=dev, dtype=torch.uint8) if nf else None
            check(lib().mc_clip_stats(ptr(I_all), ptr(T_all), ptr(planes), b, B, D, row_offset,
                                      float(tau), self.mode, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                      ptr(out[3]), ptr(flags_raw), ptr(ws), ws.numel(), cur_stream()), "mc_clip_stats")
        return out[:3], {"planes": planes, "ps": out[3], "flags_raw": flags_raw, "flags": None, "flags_done": False,
                         "strips": own_ws}

    def _final_flags(self, ctx, b, B, row_offset):
        """Gather every rank's raw tile flags and OR in the transposed relation (first call after the sweep)."""
        flags_raw = ctx["flags_raw"]
        if flags_raw is None or ctx["flags_done"]:
            return ctx["flags"]
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            allf = torch.empty(world * flags_raw.numel(), device=flags_raw.device, dtype=torch.uint8)
            dist.all_gather_into_tensor(allf, flags_raw, group=self.group)
        else:
            allf = flags_raw
        final = torch.empty_like(flags_raw)
        check(lib().mc_clip_flags_finalize(ptr(allf), B, b, row_offset, ptr(final), cur_stream()), "mc_clip_flags_finalize")
        ctx["flags"], ctx["flags_done"], ctx["flags_raw"] = final, True, None
        return final

    def rowloss(self, I_all, T_all, planes, b, row_offset, tau, stats_all):
        ctx = planes
        planes, ps_loc = ctx["planes"], ctx["ps"]
        B, D = I_all.shape
        dev = I_all.device
        out = torch.empty(2, b, device=dev, dtype=torch.float32)
        part = torch.empty(1, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            flags = self._final_flags(ctx, b, B, row_offset)
            ws = ctx["strips"] if ctx["strips"] is not None else self._ws(b, B, D, dev)
            check(lib().mc_clip_rowloss(ptr(I_all), ptr(T_all), ptr(planes), b, B, D, row_offset,
                                        float(tau), self.mode, ptr(stats_all[0]), ptr(stats_all[1]),
                                        ptr(stats_all[2]), ptr(ps_loc), ptr(out[0]), ptr(out[1]), ptr(part), ptr(flags),
                                        ptr(ws), ws.numel(), cur_stream()), "mc_clip_rowloss")
        return out, part

    def bwd(self, I_all, T_all, planes, b, row_offset, tau, stats_all, gq_all, grad_loss):
        ctx = planes
        planes = ctx["planes"]
        flags = ctx["flags"]
        B, D = I_all.shape
        dev = I_all.device
        dI = torch.empty(b, D, device=dev, dtype=torch.float32)
        dT = torch.empty(b, D, device=dev, dtype=torch.float32)
        gl = grad_loss.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            # the fp32 FMA engine's gradient sweep overwrites the strips with the gradient weights: it works on a copy,
            # so that a second backward over the same graph (retain_graph=True) still finds S / S^T / Z
            ws = ctx["strips"].clone() if ctx["strips"] is not None else self._ws(b, B, D, dev)
            check(lib().mc_clip_bwd(ptr(I_all), ptr(T_all), ptr(planes), b, B, D, row_offset,
                                    float(tau), self.mode, ptr(stats_all[0]), ptr(stats_all[1]),
                                    ptr(stats_all[2]), ptr(gq_all[0]), ptr(gq_all[1]), ptr(gl), ptr(dI),
                                    ptr(dT), ptr(flags), ptr(ws), ws.numel(), cur_stream()), "mc_clip_bwd")
        return dI, dT


def _all_gather_rows(x: torch.Tensor, group) -> torch.Tensor:
    """(k, b) per rank -> (k, W*b): concatenates along the last dim in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return x
    k, b = x.shape
    buf = _all_gather(x, group)
    return buf.permute(1, 0, 2).reshape(k, world * b).contiguous()


def _all_gather(x: torch.Tensor, group) -> torch.Tensor:
    """x (any shape) per rank -> (world, *x.shape); flat 1-D buffers so NCCL and gloo agree."""
    world = dist.get_world_size(group)
    buf = torch.empty((world,) + tuple(x.shape), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(buf.view(-1), x.contiguous().view(-1), group=group)
    return buf


class _GlobalClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_emb, text_emb, temperature, engine, group):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        I = image_emb.detach().float().contiguous()
        T = text_emb.detach().float().contiguous()
        b, D = I.shape
        B = b * world
        if world > 1:
            both = torch.stack([I, T])  # one collective for both towers
            buf = _all_gather(both, group)
            I_all = buf[:, 0].reshape(B, D)
            T_all = buf[:, 1].reshape(B, D)
            I_all, T_all = I_all.contiguous(), T_all.contiguous()
        else:
            I_all, T_all = I, T
        row_offset = rank * b
        stats_loc, planes = engine.stats(I_all, T_all, b, row_offset, temperature)
        stats_all = _all_gather_rows(stats_loc, group) if world > 1 else stats_loc
        gq_loc, part = engine.rowloss(I_all, T_all, planes, b, row_offset, temperature, stats_all)
        if world > 1:
            # one collective for the two vectors AND the loss partial (appended as a third row's first entry)
            packed = torch.zeros(3, b, device=gq_loc.device, dtype=gq_loc.dtype)
            packed[:2] = gq_loc
            packed[2, 0] = part.reshape(())
            gathered = _all_gather(packed, group)                       # (world, 3, b)
            gq_all = gathered[:, :2].permute(1, 0, 2).reshape(2, world * b).contiguous()
            part = gathered[:, 2, 0].sum().reshape(1)
        else:
            gq_all = gq_loc
        ctx.engine, ctx.cfg = engine, (b, row_offset, float(temperature))
        ctx.save_for_backward(I_all, T_all, stats_all, gq_all)
        ctx.planes = planes
        ctx.in_dtypes = (image_emb.dtype, text_emb.dtype)
        return part.reshape(())

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_loss):
        I_all, T_all, stats_all, gq_all = ctx.saved_tensors
        b, row_offset, tau = ctx.cfg
        dI, dT = ctx.engine.bwd(I_all, T_all, ctx.planes, b, row_offset, tau, stats_all, gq_all,
                                grad_loss)
        return dI.to(ctx.in_dtypes[0]), dT.to(ctx.in_dtypes[1]), None, None, None


# ------------------------------------------------------------------------------------------------
# peer-memory transport
# ------------------------------------------------------------------------------------------------
class _Trace:
    """MAE_CLIP_PEER_TRACE=1: a CUDA event after every exchange / sweep call of a PeerStep, printed as a timeline (ms since
    the step's first mark) by rank 0 - the substitute for an nsys timeline of the multi-GPU step."""

    def __init__(self):
        self.on = os.environ.get("MAE_CLIP_PEER_TRACE", "0") == "1"
        self.marks = []

    def mark(self, label):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((label, e))

    def dump(self, rank):
        if not self.on or not self.marks:
            return
        torch.cuda.synchronize()
        if rank == 0:
            t0 = self.marks[0][1]
            prev = 0.0
            for label, e in self.marks:
                t = t0.elapsed_time(e)
                print(f"[peer trace] {t:8.3f} ms  (+{t - prev:6.3f})  {label}", flush=True)
                prev = t
        self.marks = []


class PeerStep:
    """One forward (+ backward) of the row-sharded loss over a ``PeerExchange``: every exchange step
    is a kernel of libmae_clip_b200.so on the caller's stream.  Shared by the autograd Function
    below and by ``bench.py`` (``events``: optional list filled with CUDA events between phases)."""

    def __init__(self, exchange, mode, exchange_mode=None):
        from .functional import _mode
        self.ex = exchange
        self.mode = _mode(mode)
        self.exchange_mode = exchange_mode or os.environ.get("MAE_CLIP_PEER_MODE", "push")
        # tile flags travel as 32-bit words: every rank's share of the bitmap must be a whole number of words
        nf = lib().mc_clip_tile_flags_bytes(exchange.b, exchange.B, exchange.D, self.mode)
        self.flag_bytes = nf if (os.environ.get("MAE_CLIP_DENSE", "0") != "1" and nf % 4 == 0 and nf > 0) else 0
        if self.exchange_mode not in ("push", "pull"):
            raise ValueError(f"unknown peer exchange mode {self.exchange_mode!r}")
        if self.mode == _lib.GEMM_SIMT_FP32 or lib().mc_clip_planes_bytes(exchange.B, exchange.D, self.mode) == 0:
            raise ValueError("the peer transport feeds the tcgen05 engines (D in {128, 256}); use transport='nccl'")
        if exchange.b % 128:
            raise ValueError(f"tcgen05 engine: rows per rank must be a multiple of 128 (got {exchange.b})")
        # column LSE of S from column partials (no transposed strip in the statistics sweep; the ranks exchange one
        # B-vector each and merge) - MAE_CLIP_COLPART=0 keeps the transposed strip
        self.colpart = os.environ.get("MAE_CLIP_COLPART", "1") != "0"
        # gradient form: "stored" computes S once per rank (row half -> dT and the fp16 weight strip, column half -> this
        # rank's contribution to EVERY row of dI) and reduces the ranks' partial dI over peer memory; "ownrows"
        # (MAE_CLIP_PEER_BWD=ownrows) recomputes the transposed strip instead and needs no exchange in backward
        self.bwd_form = os.environ.get("MAE_CLIP_PEER_BWD", "stored")
        if self.bwd_form not in ("stored", "ownrows"):
            raise ValueError(f"unknown MAE_CLIP_PEER_BWD {self.bwd_form!r}")
        self._stored = None   # (W strip, zero-padded dIz, column-half workspace): allocated on the first backward
        self.trace = _Trace()

    def forward(self, I_loc, T_loc, tau, events=None):
        ex, mode, L = self.ex, self.mode, lib()
        b, B, D, rank, world = ex.b, ex.B, ex.D, ex.rank, ex.world
        dev = I_loc.device
        f32 = dict(device=dev, dtype=torch.float32)

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)

        with torch.cuda.device(dev):
            st = cur_stream()
            mark()
            self.trace.mark("forward")
            pull = self.exchange_mode == "pull"
            shard = rank * b * D * 4
            if pull:
                # local amax + copy of the shards into our rows of the region (peers read them there);
                # amax to every rank's slot
                check(L.mc_clip_amax(ptr(I_loc), ptr(T_loc), b, D, ex.local(ex.off_emb_i + shard),
                                     ex.local(ex.off_emb_t + shard), ex.local(ex.OFF_AMAX_LOCAL), st), "mc_clip_amax")
                self.trace.mark("mc_clip_amax")
                ex.publish(ex.local(ex.OFF_AMAX_LOCAL), 1, 1, 0, ex.OFF_AMAX_SLOTS, 0, rank)
                self.trace.mark("ex.publish")
            else:
                # posted stores of our shards into every rank's image of the batch + amax, one kernel
                check(L.mc_clip_push_shards(ptr(I_loc), ptr(T_loc), b, D, rank, world, ex.table(ex.off_emb_i),
                                            ex.table(ex.off_emb_t), ex.table(ex.OFF_AMAX_SLOTS),
                                            ex.local(ex.OFF_PUSH_SCRATCH), st), "mc_clip_push_shards")
                self.trace.mark("mc_clip_push_shards")
            ex.barrier()
            self.trace.mark("barrier")
            planes = torch.empty(L.mc_clip_planes_bytes(B, D, mode), device=dev, dtype=torch.uint8)
            tab_i, tab_t = ex.row_tables(pull)
            check(L.mc_clip_prepare_peers(tab_i, tab_t, world, b, D, mode, ex.local(ex.OFF_AMAX_SLOTS), ptr(planes), st),
                  "mc_clip_prepare_peers")
            self.trace.mark("mc_clip_prepare_peers")
            mark()
            nws = max(L.mc_clip_loss_workspace_bytes(b, B, D, mode),
                      L.mc_clip_stats_colpart_workspace_bytes(b, B, D, mode) if self.colpart else 0)
            ws = workspace(nws, dev)
            loc = torch.empty(6, b, **f32)  # r, c, rz, sum_j P_ij S_ij, g, q of the owned rows
            nf = self.flag_bytes
            flags_raw = torch.empty(nf, device=dev, dtype=torch.uint8) if nf else None
            if self.colpart:
                cpart = torch.empty(B, **f32)   # LSE over OUR rows of every column of S
                check(L.mc_clip_stats_colpart(ptr(planes), b, B, D, rank * b, float(tau), mode, ptr(loc[0]), ptr(cpart),
                                              ptr(loc[2]), ptr(loc[3]), ptr(flags_raw), ptr(ws), ws.numel(), st),
                      "mc_clip_stats_colpart")
                self.trace.mark("mc_clip_stats_colpart")
                ex.publish(ptr(loc[0]), 2, b, 2 * b, ex.OFF_VECS, 2 * ex.vec_stride, rank * b)   # r -> vector 0, rz -> vector 2
                self.trace.mark("ex.publish")
                ex.publish(ptr(cpart), 1, B, 0, ex.off_cpart, 0, rank * ex.vec_stride)
                self.trace.mark("ex.publish")
            else:
                check(L.mc_clip_stats(None, None, ptr(planes), b, B, D, rank * b, float(tau), mode, ptr(loc[0]), ptr(loc[1]),
                                      ptr(loc[2]), ptr(loc[3]), ptr(flags_raw), ptr(ws), ws.numel(), st), "mc_clip_stats")
                self.trace.mark("mc_clip_stats")
                ex.publish(ptr(loc), 3, b, b, ex.OFF_VECS, ex.vec_stride, rank * b)
                self.trace.mark("ex.publish")
            if nf:  # this rank's rows of the tile-flag bitmap, to every rank
                ex.publish(ptr(flags_raw), 1, nf // 4, 0, ex.off_flags, 0, rank * (nf // 4))
                self.trace.mark("ex.publish")
            ex.barrier()
            self.trace.mark("barrier")
            if self.colpart:   # every rank's partial vector has landed: fold them into c (the region's local c vector)
                check(L.mc_clip_colpart_merge(ex.local(ex.off_cpart), world, ex.vec_stride, B, ex.vec(1), st),
                      "mc_clip_colpart_merge")
                self.trace.mark("mc_clip_colpart_merge")
            flags = None
            if nf:
                flags = torch.empty(nf, device=dev, dtype=torch.uint8)
                check(L.mc_clip_flags_finalize(ex.local(ex.off_flags), B, b, rank * b, ptr(flags), st),
                      "mc_clip_flags_finalize")
                self.trace.mark("mc_clip_flags_finalize")
            mark()
            check(L.mc_clip_rowloss(None, None, ptr(planes), b, B, D, rank * b, float(tau), mode, ex.vec(0), ex.vec(1),
                                    ex.vec(2), ptr(loc[3]), ptr(loc[4]), ptr(loc[5]), ex.local(ex.OFF_PART_LOCAL),
                                    ptr(flags), ptr(ws), ws.numel(), st), "mc_clip_rowloss")
            self.trace.mark("mc_clip_rowloss")
            ex.publish(ptr(loc[4]), 2, b, b, ex.OFF_VECS + 4 * 3 * ex.vec_stride, ex.vec_stride, rank * b)
            self.trace.mark("ex.publish")
            ex.publish(ex.local(ex.OFF_PART_LOCAL), 1, 1, 0, ex.OFF_PART_SLOTS, 0, rank)
            self.trace.mark("ex.publish")
            ex.barrier()
            self.trace.mark("barrier")
            # out of the region: backward reads its vectors from these copies, so a second forward before it is harmless
            # (the stored-weights backward reuses only the dead image of I and the latest flag bitmap, see backward)
            vecs = torch.empty(5, B, **f32)
            parts = torch.empty(18, **f32)   # 16 partial slots + the barrier's {epoch, error} words, one copy
            ex.copy_out(ex.OFF_VECS, 5, B, ex.vec_stride, vecs, B)
            self.trace.mark("ex.copy_out")
            ex.copy_out(ex.OFF_PART_SLOTS, 1, 18, 0, parts, 0)
            self.trace.mark("ex.copy_out")
            # a barrier that gave up on a peer (peer.py, "Skew between ranks") makes the step's loss NaN instead of
            # silently wrong; no host synchronisation here - PeerExchange.check() names the missing rank
            loss = torch.where(parts.view(torch.int32)[17] != 0, parts.new_full((), float("nan")), parts[:world].sum())
            mark()
        return loss, (planes, vecs, flags)

    def backward(self, saved, tau, grad_loss=None, events=None):
        ex, mode, L = self.ex, self.mode, lib()
        planes, vecs, flags = saved
        b, B, D = ex.b, ex.B, ex.D
        dev = planes.device
        dI = torch.empty(b, D, device=dev, dtype=torch.float32)
        dT = torch.empty(b, D, device=dev, dtype=torch.float32)
        gl = None if grad_loss is None else grad_loss.reshape(1).to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            ws = workspace(L.mc_clip_loss_workspace_bytes(b, B, D, mode), dev)
            if (self.bwd_form == "stored" and ex.world > 1 and flags is not None and mode != _lib.GEMM_SIMT_FP32
                    and b * B >= 4096 * 4096):   # smaller strips are launch-bound: one own-rows sweep
                st = cur_stream()
                if self._stored is None:
                    Bp = (B + 127) // 128 * 128
                    self._stored = (torch.empty(L.mc_clip_stored_weights_bytes(b, B), device=dev, dtype=torch.uint8),
                                    torch.zeros(Bp, D, device=dev, dtype=torch.float32),   # only OUR rows are ever written
                                    torch.empty(L.mc_clip_bwd_cols_workspace_bytes(B, D), device=dev, dtype=torch.uint8),
                                    torch.zeros(1, device=dev, dtype=torch.int32))
                W, diz, wsc, gate = self._stored
                row0 = ex.rank * b
                # stored form or own-rows sweep: decided on the device from the flag density of the WHOLE bitmap (every
                # rank holds it after the forward exchange), so all ranks take the same branch.  The bitmap in the region
                # is the LATEST forward's; after two forwards the older backward may take the other (equally exact) form.
                nt = (B + 127) // 128
                check(L.mc_clip_bwd_gate(ex.local(ex.off_flags), nt * nt, ptr(gate), st), "mc_clip_bwd_gate")
                self.trace.mark("mc_clip_bwd_gate")
                gp = gate
                check(L.mc_clip_bwd_rows(ptr(planes), b, B, D, row0, float(tau), mode, ptr(vecs[0]), ptr(vecs[1]),
                                         ptr(vecs[2]), ptr(vecs[3]), ptr(vecs[4]), ptr(gl), ptr(dT), ptr(diz[row0:]), ptr(W),
                                         ptr(flags), ptr(gp), ptr(dI), ptr(ws), ws.numel(), st), "mc_clip_bwd_rows")
                self.trace.mark("mc_clip_bwd_rows")
                # our contribution to every row of dI goes where the peers can read it: the (B, D) image of I in the
                # exchange region is dead once the planes are staged
                check(L.mc_clip_bwd_cols(ptr(planes), B, D, float(tau), mode, ptr(vecs[0]), ptr(vecs[1]), ptr(vecs[2]),
                                         ptr(vecs[4]), ptr(gl), ptr(W), b, row0, 0, B, ptr(diz), ex.local(ex.off_emb_i),
                                         ptr(gp), ptr(wsc), wsc.numel(), st), "mc_clip_bwd_cols")
                self.trace.mark("mc_clip_bwd_cols")
                ex.barrier()
                self.trace.mark("barrier")
                ex.reduce_rows(ex.off_emb_i + row0 * D * 4, b * D, dI, gp)
                self.trace.mark("ex.reduce_rows")
                ex.barrier()   # nobody pushes the next step's shards into an image a peer is still reading
                self.trace.mark("barrier")
            else:
                check(L.mc_clip_bwd(None, None, ptr(planes), b, B, D, ex.rank * b, float(tau), mode, ptr(vecs[0]),
                                    ptr(vecs[1]), ptr(vecs[2]), ptr(vecs[3]), ptr(vecs[4]), ptr(gl), ptr(dI), ptr(dT),
                                    ptr(flags), ptr(ws), ws.numel(), cur_stream()), "mc_clip_bwd")
                self.trace.mark("mc_clip_bwd")
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)
            self.trace.dump(ex.rank)
        return dI, dT


class _PeerGlobalClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_emb, text_emb, temperature, step):
        I = image_emb.detach().float().contiguous()
        T = text_emb.detach().float().contiguous()
        loss, saved = step.forward(I, T, temperature)
        ctx.step, ctx.saved, ctx.tau = step, saved, float(temperature)
        ctx.in_dtypes = (image_emb.dtype, text_emb.dtype)
        return loss

    @staticmethod
    @once_differentiable   # the C library's gradients are not themselves differentiable
    def backward(ctx, grad_loss):
        dI, dT = ctx.step.backward(ctx.saved, ctx.tau, grad_loss)
        return dI.to(ctx.in_dtypes[0]), dT.to(ctx.in_dtypes[1]), None, None


def _pick_transport(transport, image_emb, mode, engine, world):
    """'peer' when every precondition of the peer-memory path holds, else 'nccl'."""
    want = transport or os.environ.get("MAE_CLIP_TRANSPORT", "auto")
    if want not in ("auto", "peer", "nccl"):
        raise ValueError(f"unknown transport {want!r}")
    if want == "nccl" or engine is not None or world == 1 or not image_emb.is_cuda:
        if want == "peer" and (engine is not None or not image_emb.is_cuda):
            raise ValueError("transport='peer' needs CUDA tensors and the built-in engine")
        return "nccl"
    from .functional import _mode
    b, D = image_emb.shape
    ok = (_mode(mode) != _lib.GEMM_SIMT_FP32 and lib().mc_clip_planes_bytes(b * world, D, _mode(mode)) > 0
          and b % 128 == 0 and world <= 16)
    if want == "peer" and not ok:
        raise ValueError("transport='peer' needs a tcgen05 engine (D in {128, 256}) and rows per rank % 128 == 0")
    return "peer" if ok else "nccl"


def global_clip_loss(image_emb_local, text_emb_local, temperature: float = 1.0, mode=None,
                     group=None, engine=None, transport=None):
    """Loss of the reference on the concatenation of every rank's (b, D) embeddings; the value is
    identical on all ranks and ``backward`` yields d loss_global / d (local embeddings) in full.
    (Under DDP, which averages parameter gradients over ranks, multiply the loss by the world
    size to obtain the single-process gradient.)"""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if _pick_transport(transport, image_emb_local, mode, engine, world) == "peer":
        from .peer import PeerUnavailable, get_exchange
        b, D = image_emb_local.shape
        try:
            ex = get_exchange(b, D, group)
            # one PeerStep per (exchange, engine): it owns the stored-weights buffers of the backward (b x B fp16)
            cache = ex.__dict__.setdefault("_peer_steps", {})
            step = cache.get(mode)
            if step is None:
                step = cache[mode] = PeerStep(ex, mode)
        except PeerUnavailable:
            # raised on every rank together: without an explicit request the NCCL transport takes over
            if (transport or os.environ.get("MAE_CLIP_TRANSPORT", "auto")) == "peer":
                raise
            step = None
        if step is not None:
            return _PeerGlobalClipLoss.apply(image_emb_local, text_emb_local, float(temperature), step)
    engine = engine if engine is not None else CudaStripEngine(mode, group)
    return _GlobalClipLoss.apply(image_emb_local, text_emb_local, float(temperature), engine, group)
