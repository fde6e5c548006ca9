"""Peer-memory exchange for the row-sharded contrastive loss (SURVEY.md section 8 e).

One process per GPU.  Every rank owns one exchange region (``mc_peer_alloc``), ships its CUDA IPC
handle to the other local ranks through ``torch.distributed`` (plumbing) and maps theirs; from then
on the data path is OUR kernels loading / storing peer memory over NVLink / NVSwitch - no
collective-library call per step:

  * the embedding all-gather is fused into the operand staging (``mc_clip_prepare_peers`` pulls
    each fp32 row from its owner while it writes the local fp16 planes);
  * the row statistics, (g, q) vectors and loss partials are pushed into every peer's region by
    ``mc_peer_publish`` and fenced by ``mc_peer_barrier`` (flag exchange, system-scope
    release / acquire).

The reference is single-process (nothing under /root/reference to cite); the result is, by
definition, the reference loss (``CLIP.py:34-43``) on the concatenated batch.

Region layout (bytes): [0, 64) barrier flags | [256, 320) amax slots | [512, 576) loss-partial slots | 576 private
barrier epoch, 580 barrier error word (0 = fine, 1 + q = "rank q never arrived within the timeout"; next to the
partials so that one copy brings both out) | 768 local amax, 772 local loss partial, 784 push scratch (2 words) | 1024.. five length-B vectors
(r, c, rz, g, q) | the [B/128][B/128]-byte tile-flag bitmap | [world] column-LSE partial vectors of B floats | then two (B, D) fp32 images of the global batch (image / text embeddings): in
"pull" mode each rank fills only its own rows and peers read them from there; in "push" mode every
rank stores its rows into all ranks' images and the staging reads locally.

Why the region can be reused every step without extra fences (B_k = k-th barrier of a step):
  shards + amax slots written before B_1, read by peers between B_1 and B_2;  r, c, rz pushed
  between B_1 and B_2, read after B_2;  g, q, partials pushed between B_2 and B_3, read after B_3
  and copied out of the region right there.  A rank can only write step n+1's data after passing
  B_1(n+1) / B_2(n+1), which every peer enters only after its own stream finished reading step n's.
  The own-rows backward never touches peer memory.  The stored-weights backward (``dist.PeerStep``)
  does: each rank writes its partial dI for ALL B rows into its OWN (B, D) image of the image
  embeddings - dead since the operand staging after B_1 - then B_4, every rank sums its rows of the
  peers' images (``mc_peer_reduce``), then B_5 so that no rank pushes the next step's shards into an
  image a peer is still reading.  Like a collective, that backward must run on every rank of the group.

Skew between ranks.  The barrier spins ON THE GPU until every peer's stream reaches the same barrier, at most
``MAE_CLIP_PEER_TIMEOUT_S`` seconds (default 600, the order of a process-group timeout: rank-0-only validation or
checkpointing, a data-loader stall or a first-step autotune on one rank are routine).  Rank-asymmetric GPU work that
itself waits for a peer (an NCCL collective issued on the same stream by only some ranks) would still deadlock the
step, as it would with NCCL.  On a timeout the kernel does NOT trap (a trap is a sticky context error): it records
the missing rank, the step's loss comes out NaN, ``PeerExchange.check()`` raises ``PeerTimeout`` - after which the
exchange must be closed (``close_all()``) and the job can continue over ``transport="nccl"``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from ._lib import check, cur_stream, lib

_HANDLE_BYTES = 64
MAX_WORLD = 16


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class PeerUnavailable(RuntimeError):
    """Peer memory could not be set up on at least one rank (no P2P / IPC between the devices, or a
    container without a shared IPC namespace).  Raised on EVERY rank, so callers can fall back to the
    NCCL transport together."""


class PeerTimeout(RuntimeError):
    """A peer did not reach a barrier of the exchange within the timeout (see the module docstring)."""


def default_timeout_s() -> float:
    try:
        return float(os.environ.get("MAE_CLIP_PEER_TIMEOUT_S", "600"))
    except ValueError:
        return 600.0


class PeerExchange:
    """The mapped exchange regions of every rank of ``group`` for a (b, D) shard shape."""

    OFF_FLAGS, OFF_AMAX_SLOTS, OFF_PART_SLOTS, OFF_EPOCH, OFF_AMAX_LOCAL, OFF_PART_LOCAL, OFF_VECS = \
        0, 256, 512, 576, 768, 772, 1024
    OFF_ERROR = OFF_EPOCH + 4  # second word of the barrier's private block: 1 + rank that never arrived, 0 = fine
    OFF_PUSH_SCRATCH = 784  # {amax accumulator, blocks-done counter} of mc_clip_push_shards

    def __init__(self, b: int, D: int, group=None, device=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > MAX_WORLD:
            raise ValueError(f"peer exchange supports up to {MAX_WORLD} ranks on one NVLink domain (got {self.world})")
        self.b, self.D, self.B = b, D, b * self.world
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.vec_stride = _round_up(self.B, 64)                         # floats
        # tile-flag bitmap of the whole batch ([B/128][B/128] bytes; every rank pushes its row blocks)
        self.off_flags = _round_up(self.OFF_VECS + 5 * self.vec_stride * 4, 256)
        nt = (self.B + 127) // 128
        # column-LSE partials of every rank: [world][vec_stride] floats (rank q's vector over ITS rows, all B columns)
        self.off_cpart = _round_up(self.off_flags + nt * nt, 256)
        # (B, D) fp32 images of the global batch: pull mode uses only the owner's rows of each
        self.off_emb_i = _round_up(self.off_cpart + self.world * self.vec_stride * 4, 256)
        self.off_emb_t = self.off_emb_i + _round_up(self.B * D * 4, 256)
        self.nbytes = self.off_emb_t + _round_up(self.B * D * 4, 256)
        self.ptrs = [0] * self.world
        self._tables = {}
        self._local = C.c_void_p()
        self._open = []
        # Every rank takes part in both object all-gathers whatever happens locally, so a failure on one rank
        # surfaces as PeerUnavailable on all of them instead of a hang.
        err = None
        raw = None
        with torch.cuda.device(self.device):
            try:
                if os.environ.get("MAE_CLIP_PEER_INJECT_FAILURE") == str(self.rank):  # test hook: the agreement path
                    raise RuntimeError("injected set-up failure")
                handle = C.create_string_buffer(_HANDLE_BYTES)
                check(lib().mc_peer_alloc(self.nbytes, C.byref(self._local), handle), "mc_peer_alloc")
                raw = bytes(handle.raw)
            except Exception as e:  # noqa: BLE001
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, raw, self.nbytes), group=group)
            if err is None:
                try:
                    for q, (rq, hq, nq) in enumerate(handles):
                        if hq is None:
                            raise PeerUnavailable(f"rank {q} could not allocate its exchange region")
                        if rq != q or nq != self.nbytes:
                            raise RuntimeError("peer exchange: ranks disagree on the region size (different b / D per rank?)")
                        if q == self.rank:
                            self.ptrs[q] = self._local.value
                            continue
                        p = C.c_void_p()
                        check(lib().mc_peer_open(C.create_string_buffer(hq, _HANDLE_BYTES), C.byref(p)), "mc_peer_open")
                        self._open.append(p)
                        self.ptrs[q] = p.value
                except Exception as e:  # noqa: BLE001
                    err = e
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
            if not all(oks):
                for p in self._open:
                    lib().mc_peer_close(p)
                if self._local.value:
                    lib().mc_peer_free(self._local)
                self._open, self._local = [], None
                bad = [q for q, ok in enumerate(oks) if not ok]
                raise PeerUnavailable(f"peer memory set-up failed on rank(s) {bad}" + (f": {err}" if err else ""))
        dist.barrier(group=group)  # every region is mapped (and zero-filled) before the first flag is written

    # ---- pointer helpers ----------------------------------------------------------------------
    def local(self, offset: int = 0) -> C.c_void_p:
        return C.c_void_p(self.ptrs[self.rank] + offset)

    def table(self, offset: int = 0):
        """Host array of ``world`` device pointers: every rank's region + offset (C-ABI ``void* const*``)."""
        t = self._tables.get(offset)
        if t is None:
            t = self._tables[offset] = (C.c_void_p * self.world)(*[p + offset for p in self.ptrs])
        return t

    def row_tables(self, pull: bool):
        """Per-owner source pointers for mc_clip_prepare_peers: rank q's rows in q's region (pull) or in ours (push)."""
        key = ("rows", pull)
        t = self._tables.get(key)
        if t is None:
            shard = self.b * self.D * 4
            base = self.ptrs if pull else [self.ptrs[self.rank]] * self.world
            t = self._tables[key] = tuple(
                (C.c_void_p * self.world)(*[base[q] + off + q * shard for q in range(self.world)])
                for off in (self.off_emb_i, self.off_emb_t))
        return t

    def vec(self, k: int) -> C.c_void_p:
        """Local copy of the k-th length-B vector (0 r, 1 c, 2 rz, 3 g, 4 q)."""
        return self.local(self.OFF_VECS + 4 * k * self.vec_stride)

    # ---- primitives ---------------------------------------------------------------------------
    def barrier(self, timeout_s: float | None = None):
        check(lib().mc_peer_barrier(self.table(self.OFF_FLAGS), self.rank, self.world, self.local(self.OFF_EPOCH),
                                    float(default_timeout_s() if timeout_s is None else timeout_s), cur_stream()),
              "mc_peer_barrier")

    def check(self):
        """Synchronise and raise ``PeerTimeout`` if a barrier of this exchange gave up on a peer."""
        w = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self.copy_out(self.OFF_ERROR, 1, 1, 0, w, 0)
        v = int(w.item())
        if v:
            raise PeerTimeout(f"rank {self.rank}: peer rank {v - 1} did not reach a barrier within the timeout; close the "
                              "exchange (peer.close_all()) and continue with transport='nccl'")

    def publish(self, src, k: int, n: int, src_stride: int, dst_offset_bytes: int, dst_stride: int, dst_index: int):
        """Push k vectors of n words from local ``src`` to [dst_offset + kk*dst_stride + dst_index + i] of every rank."""
        check(lib().mc_peer_publish(src, k, n, src_stride, self.table(dst_offset_bytes), dst_stride, dst_index,
                                    self.world, cur_stream()), "mc_peer_publish")

    def reduce_rows(self, src_offset_bytes: int, n_floats: int, out: torch.Tensor, gate=None):
        """out[i] = sum over ranks q of (q's region + src_offset_bytes)[i]: every rank pulls its own rows of the peers'
        partials (``gate``: optional device word, the kernel returns unless it is 1)."""
        check(lib().mc_peer_reduce(self.table(src_offset_bytes), self.world, n_floats, C.c_void_p(out.data_ptr()),
                                   None if gate is None else C.c_void_p(gate.data_ptr()), cur_stream()), "mc_peer_reduce")

    def copy_out(self, src_offset_bytes: int, k: int, n: int, src_stride: int, dst: torch.Tensor, dst_stride: int):
        """Copy k vectors of n words from the local region into an ordinary tensor (same kernel, one target)."""
        tab = (C.c_void_p * 1)(dst.data_ptr())
        check(lib().mc_peer_publish(self.local(src_offset_bytes), k, n, src_stride, tab, dst_stride, 0, 1,
                                    cur_stream()), "mc_peer_publish(copy out)")

    def close(self):
        if self._local is None:
            return
        torch.cuda.synchronize(self.device)
        try:
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still be reading
        except Exception:
            pass
        with torch.cuda.device(self.device):
            for p in self._open:
                lib().mc_peer_close(p)
            lib().mc_peer_free(self._local)
        self._open, self._local = [], None


_cache = {}


def get_exchange(b: int, D: int, group=None) -> PeerExchange:
    """One exchange per (group, device, b, D): set-up costs a device synchronisation and two object
    all-gathers, so it is created on first use and kept.  A failed set-up is remembered too
    (``PeerUnavailable`` is raised again without another collective round)."""
    key = (id(group) if group is not None else 0, torch.cuda.current_device(), b, D)
    ex = _cache.get(key)
    if ex is None:
        try:
            ex = PeerExchange(b, D, group)
        except PeerUnavailable as e:
            ex = e
        _cache[key] = ex
    if isinstance(ex, PeerUnavailable):
        raise ex
    return ex


def close_all():
    for ex in list(_cache.values()):
        if isinstance(ex, PeerExchange):
            ex.close()
    _cache.clear()
