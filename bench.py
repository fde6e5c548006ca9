#!/usr/bin/env python
"""Benchmark of the mae_clip training-loss hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload c4|c2] [--mode tc_f16x3|tc_f16|simt_fp32] [--batch B]

Workload (BASELINE.json): the metric "contrastive+MAE loss fwd/bwd samples/s at 1/2/4/8 B200; % TC
/HBM roofline" is quoted on config 4, the global-batch contrastive loss at B = 32768, D = 256: it
is the only config that names 1..8 GPUs, it fits one B200 because no B x B tensor is ever
materialised, and the tensor-core roofline is only meaningful at that size (C2's B = 1024 is a
3.8 GFLOP, launch-bound problem).  The global batch is FIXED as N grows ("strong" scaling): rank r
owns rows [r B/N, (r+1) B/N).  A step = one forward + backward of the loss over the whole global
batch from fp32 embeddings already in HBM: stage operands, row/col statistics sweep, row-loss
sweep, gradient sweep.  When N > 1 the exchange steps run over mapped peer memory by default
(--transport peer: the embedding all-gather is fused into the operand staging, statistics are pushed
by our own kernels; --transport nccl keeps the torch.distributed all-gathers).  At N = 1 the line also carries the C2
latency and C5 MAE numbers under "extra".

e2e is the same step through the C-ABI entry point that takes HOST buffers
(mc_clip_loss_fwd_bwd_host at N = 1; pinned-host shards -> H2D -> dist loss -> D2H at N > 1).

--impl reference times the reference's own CPU implementation of the path (the op-for-op oracle
port of CLIP.py:34-43 + autograd; the Python reference cannot travel to the GPU box) with all host
threads on a bounded sample; see cpu_baseline.sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_EMB = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2"])
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--batch", type=int, default=0, help="global batch (default: 32768 for c4, 1024 for c2)")
    ap.add_argument("--cpu-sample-batch", type=int, default=8192,
                    help="rows of the bounded CPU sample (the reference materialises ~20 B x B fp32 tensors: 5 GiB at 8192)")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: peer = our kernels over mapped peer memory (default), nccl = torch.distributed all-gathers")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the global batch stays 32768 as N grows; weak: 4096 rows per GPU, B = 4096 N "
                         "(config 4 read literally; the loss is O(B^2), so samples/s per GPU then FALLS with N by construction)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            # nvidia-smi needs a few hundred ms before its first sample: wait for it, or the whole (sub-second) run
            # is over before the sampler sees anything
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def make_shard(b, seed, scale=1.0):
    """Embeddings with the distribution ProjectionHead emits (LayerNorm'd gaussian rows)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, D_EMB, generator=g)
    return torch.nn.functional.layer_norm(x, (D_EMB,)) * scale


def cpu_reference_leg(args, B_work, steps, warmup):
    """The oracle port of the reference loss (CLIP.py:34-43 + autograd) on host cores."""
    import torch

    from oracle import loss_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(args.cpu_sample_batch, B_work)
    I = make_shard(Bs, 0)
    T = make_shard(Bs, 1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss_ref.clip_loss_fwd_bwd_ref(I, T, 1.0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    raw = Bs / t
    # cost per sample is linear in B (the loss is O(B^2 D)): project the sample to the workload's B
    value = raw * (Bs / B_work)
    sample = (f"full reference loss fwd+bwd (oracle port, torch CPU fp32, {cores} threads) at B={Bs}: "
              f"{t * 1e3:.1f} ms/step = {raw:.0f} samples/s at that B; projected to the workload's B={B_work} "
              f"by the O(B^2) cost (x {Bs}/{B_work})") if Bs != B_work else \
             (f"full reference loss fwd+bwd (oracle port, torch CPU fp32, {cores} threads) at B={Bs}: "
              f"{t * 1e3:.1f} ms/step")
    return {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_sample_step": t * 1e3, "sample_batch": Bs}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch or (32768 if args.workload == "c4" else 1024)
    steps, warmup = max(1, args.steps), max(0, args.warmup)  # exactly K timed steps after W warm-ups
    cb = cpu_reference_leg(args, B, steps, warmup)
    line = {"impl": "reference", "metric": "contrastive loss fwd+bwd samples/s", "value": cb["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["ms_per_sample_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, B, "cpu"),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, B, mode):
    name = ("c4: global-batch contrastive soft-target loss fwd+bwd, B=%d D=%d, rows sharded over ranks" % (B, D_EMB)
            if args.workload == "c4" else
            "c2: CLIP loss-only microbench fwd+bwd, B=%d D=%d fp32" % (B, D_EMB))
    return {"workload": name, "global_batch": B, "dim": D_EMB, "temperature": 1.0, "engine": mode,
            "parallelism": f"row-sharded x{args.gpus}", "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------------------
def pick_mode(args):
    from mae_clip_b200 import _lib
    if args.mode != "auto":
        return args.mode
    return "tc_f16x3" if _lib.lib().mc_clip_planes_bytes(128, D_EMB, _lib.GEMM_MODES["tc_f16x3"]) > 256 \
        else "simt_fp32"


class Phases:
    """One step on one rank through the C ABI, with CUDA events between the phases so the dominant
    kernel's duration is measured live on the launching stream."""

    def __init__(self, B, b, row_offset, mode, device):
        import torch

        from mae_clip_b200 import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.lib()
        self.B, self.b, self.off, self.mode = B, b, row_offset, _lib.GEMM_MODES[mode]
        f32 = dict(device=device, dtype=torch.float32)
        self.stats_loc = torch.empty(4, b, **f32)  # r, c, rz, sum_j P_ij S_ij
        self.gq_loc = torch.empty(2, b, **f32)
        self.part = torch.empty(1, **f32)
        self.part_buf = self.part
        self.pack = torch.zeros(3, b, **f32)
        self.dI = torch.empty(b, D_EMB, **f32)
        self.dT = torch.empty(b, D_EMB, **f32)
        nb = self.lib.mc_clip_planes_bytes(B, D_EMB, self.mode)
        self.planes = torch.empty(max(nb, 1), device=device, dtype=torch.uint8)
        nf = 0 if os.environ.get("MAE_CLIP_DENSE", "0") == "1" else self.lib.mc_clip_tile_flags_bytes(b, B, D_EMB, self.mode)
        self.flags_raw = torch.empty(nf, device=device, dtype=torch.uint8) if nf else None
        self.flags = torch.empty(nf, device=device, dtype=torch.uint8) if nf else None
        nws = self.lib.mc_clip_loss_workspace_bytes(b, B, D_EMB, self.mode)
        self.ws = torch.empty(max(nws, 1), device=device, dtype=torch.uint8)
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step(self, I_all, T_all, gather_vec=None, record=False):
        lib, ck, p = self.lib, self._lib.check, self._lib.ptr
        st = self._lib.cur_stream()
        B, b, off, mode = self.B, self.b, self.off, self.mode
        ev = self.ev
        if record: ev[0].record()
        ck(lib.mc_clip_prepare(p(I_all), p(T_all), B, B, D_EMB, 0, mode, p(self.planes), st), "prepare")
        if record: ev[1].record()
        ck(lib.mc_clip_stats(p(I_all), p(T_all), p(self.planes), b, B, D_EMB, off, 1.0, mode, p(self.stats_loc[0]),
                             p(self.stats_loc[1]), p(self.stats_loc[2]), p(self.stats_loc[3]), p(self.flags_raw), p(self.ws),
                             self.ws.numel(), st), "stats")
        stats_all = gather_vec(self.stats_loc[:3]) if gather_vec else self.stats_loc
        if self.flags_raw is not None:   # tile flags: gather the raw bitmap rows of every rank, OR in the transpose
            flags_all = self.flags_raw
            if gather_vec:
                import torch.distributed as dist
                flags_all = self.torch.empty(self.flags_raw.numel() * (B // b), device=self.flags_raw.device, dtype=self.torch.uint8)
                dist.all_gather_into_tensor(flags_all, self.flags_raw)
            ck(lib.mc_clip_flags_finalize(p(flags_all), B, b, off, p(self.flags), st), "flags_finalize")
        if record: ev[2].record()
        ck(lib.mc_clip_rowloss(p(I_all), p(T_all), p(self.planes), b, B, D_EMB, off, 1.0, mode, p(stats_all[0]),
                               p(stats_all[1]), p(stats_all[2]), p(self.stats_loc[3]), p(self.gq_loc[0]), p(self.gq_loc[1]), p(self.part_buf),
                               p(self.flags), p(self.ws), self.ws.numel(), st), "rowloss")
        if gather_vec:  # the loss partial rides along with the two vectors: one collective instead of two
            self.pack[:2] = self.gq_loc
            self.pack[2, 0:1] = self.part_buf
            g3 = gather_vec(self.pack)
            gq_all = g3[:2]
            self.part = g3[2].reshape(-1, self.b)[:, 0].sum().reshape(1)
        else:
            gq_all = self.gq_loc
        if record: ev[3].record()
        ck(lib.mc_clip_bwd(p(I_all), p(T_all), p(self.planes), b, B, D_EMB, off, 1.0, mode, p(stats_all[0]),
                           p(stats_all[1]), p(stats_all[2]), p(gq_all[0]), p(gq_all[1]), None, p(self.dI), p(self.dT),
                           p(self.flags), p(self.ws), self.ws.numel(), st), "bwd")
        if record: ev[4].record()
        return self.part

    def phase_ms(self):
        e = self.ev
        return [e[i].elapsed_time(e[i + 1]) for i in range(4)]


class PeerPhases:
    """Same step over the peer-memory transport (mae_clip_b200/dist.py PeerStep): every exchange is a
    kernel of libmae_clip_b200.so on this stream; the embedding all-gather is fused into the staging."""

    def __init__(self, B, b, rank, mode, device):
        from mae_clip_b200.dist import PeerStep
        from mae_clip_b200.peer import get_exchange
        self.step_impl = PeerStep(get_exchange(b, D_EMB), mode)
        self.part = self.dI = self.dT = None
        self.ev = []

    def step(self, I_loc, T_loc, record=False):
        ev = [] if record else None
        loss, saved = self.step_impl.forward(I_loc, T_loc, 1.0, ev)
        self.last_flags = saved[2]
        self.dI, self.dT = self.step_impl.backward(saved, 1.0, None, ev)
        self.part = loss.reshape(1)
        if record:
            self.ev = ev
        return self.part

    def phase_ms(self):
        e = self.ev  # start, staged, stats exchanged, row loss exchanged, gradients
        return [e[i].elapsed_time(e[i + 1]) for i in range(4)]


def run_b200(args):
    import torch
    import torch.distributed as dist

    from mae_clip_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    B = args.batch or (32768 if args.workload == "c4" else 1024)
    if args.scaling == "weak":
        B = (args.batch or 4096) * world
    assert B % world == 0
    b = B // world
    mode = pick_mode(args)
    lib = _lib.lib()
    peaks = load_peaks()

    # this rank's shard; C4 seeds 1000 + r per 4096-row block so the global batch is the same for every N
    blk = 4096 if B % 4096 == 0 and b % 4096 == 0 else b
    shard_I = torch.cat([make_shard(blk, 1000 + (rank * b) // blk + i) for i in range(b // blk)])
    shard_T = torch.cat([make_shard(blk, 5000 + (rank * b) // blk + i) for i in range(b // blk)])
    host_I, host_T = shard_I.pin_memory(), shard_T.pin_memory()
    I_loc, T_loc = host_I.to(dev), host_T.to(dev)

    def gather_rows(x):  # (b, D) per rank -> (B, D)
        if world == 1:
            return x
        out = torch.empty(world * x.shape[0], x.shape[1], device=dev, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x)
        return out

    def gather_vec(x):  # (k, b) per rank -> (k, B)
        k = x.shape[0]
        buf = torch.empty(world, k, b, device=dev, dtype=x.dtype)
        dist.all_gather_into_tensor(buf.view(-1), x.contiguous().view(-1))
        return buf.permute(1, 0, 2).reshape(k, B).contiguous()

    transport = args.transport
    if transport == "auto":
        transport = "peer" if (world > 1 and mode != "simt_fp32" and b % 128 == 0) else "nccl"
    if world == 1:
        transport = "none"
    ph = None
    if transport == "peer":
        from mae_clip_b200.peer import PeerUnavailable
        try:
            ph = PeerPhases(B, b, rank, mode, dev)
        except PeerUnavailable as e:  # raised on every rank together (no P2P / IPC on this box)
            if args.transport == "peer":
                raise
            if rank == 0:
                print(f"bench: peer memory unavailable ({e}); using the NCCL transport", file=sys.stderr)
            transport = "nccl"
    if ph is None:
        ph = Phases(B, b, rank * b, mode, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one_step(record=False, I=None, T=None):
        I, T = (I_loc if I is None else I), (T_loc if T is None else T)
        if transport == "peer":
            return ph.step(I, T, record=record)
        I_all, T_all = gather_rows(I), gather_rows(T)
        return ph.step(I_all, T_all, gather_vec if world > 1 else None, record=record)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the warm-up on (same workload): the timed region of a sharded run is only tens of ms
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.mc_kernel_launch_count()
    total_ms, phase_ms = 0.0, [0.0] * 4
    for _ in range(args.steps):
        flush.fill_(1)  # evict L2 between timed iterations (outside the timed region)
        barrier()
        t_start.record()
        one_step(record=True)
        t_end.record()
        barrier()
        total_ms += t_start.elapsed_time(t_end)
        for i, v in enumerate(ph.phase_ms()):
            phase_ms[i] += v
    launches = lib.mc_kernel_launch_count() - launches0
    loss_val = float(ph.part.item())
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = tmax.item()
    ms_per_step = total_ms / args.steps
    value = B / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, loss + gradients out, copies inside the timed region
    e2e_ms = 0.0
    h2d = 2 * b * D_EMB * 4
    d2h = 2 * b * D_EMB * 4 + 4
    if world == 1:
        nws = lib.mc_clip_loss_host_workspace_bytes(B, D_EMB, _lib.GEMM_MODES[mode])
        dws = torch.empty(nws, dtype=torch.uint8, device=dev)
        out_dI, out_dT = torch.empty_like(host_I).pin_memory(), torch.empty_like(host_T).pin_memory()
        out_loss = torch.zeros(1).pin_memory()
        st = torch.cuda.current_stream().cuda_stream

        def e2e_step():
            _lib.check(lib.mc_clip_loss_fwd_bwd_host(host_I.data_ptr(), host_T.data_ptr(), B, D_EMB, 1.0,
                                                     _lib.GEMM_MODES[mode], out_loss.data_ptr(), out_dI.data_ptr(),
                                                     out_dT.data_ptr(), dws.data_ptr(), nws, ctypes.c_void_p(st)),
                       "mc_clip_loss_fwd_bwd_host")
    else:
        out_dI, out_dT = torch.empty_like(host_I).pin_memory(), torch.empty_like(host_T).pin_memory()
        out_loss = torch.zeros(1).pin_memory()

        def e2e_step():
            Il = host_I.to(dev, non_blocking=True)
            Tl = host_T.to(dev, non_blocking=True)
            part = one_step(False, Il, Tl)
            out_dI.copy_(ph.dI, non_blocking=True)
            out_dT.copy_(ph.dT, non_blocking=True)
            out_loss.copy_(part, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_step()
    e_steps = max(1, min(args.steps, 5))
    for _ in range(e_steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    emax = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(emax, op=dist.ReduceOp.MAX)
    e2e_value = B / (emax.item() / e_steps * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        flops_step = 14.0 * B * B * D_EMB / world          # algorithmic, this rank's strip
        bwd_ms = phase_ms[3] / args.steps
        flops_bwd = 8.0 * B * B * D_EMB / world            # the gradient sweep: 4 GEMMs
        achieved = flops_bwd / (bwd_ms * 1e-3) / 1e12
        peak = peaks["tc_sustained"] or peaks["tc_burst"]
        # executed tensor-core work of the gradient sweep, in GEMM units of 2 B^2 D / world FLOPs: S and S^T recomputed
        # (2 units x passes) + their two gradient GEMMs, and - only on the tiles flagged as carrying soft-target mass -
        # Z (2 units x passes, K = 2D) + the two dZ GEMMs.  The algorithmic count is the 4 gradient GEMMs.
        passes = 3 if mode == "tc_f16x3" else 1
        flags_t = getattr(ph, "flags", None)
        if flags_t is None and hasattr(ph, "step_impl"):
            flags_t = getattr(ph, "last_flags", None)
        density = float(flags_t.float().mean().item()) if flags_t is not None else 1.0
        exec_units = ((2 * passes + 2) + density * (2 * passes + 2)) if mode != "simt_fp32" else 8
        executed = exec_units * 2.0 * B * B * D_EMB / world / (bwd_ms * 1e-3) / 1e12
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_roofline_ncu.json")))
        except Exception:
            pass
        line = {
            "metric": "contrastive loss fwd+bwd samples/s", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"simt_fp32": "fp32", "tc_f16x3": "fp32 (fp16 hi+lo split operands, 3 tcgen05 passes, fp32 accumulate)",
                      "tc_f16": "fp16 operands, fp32 accumulate"}[mode],
            "data": "synthetic", "config": dict(workload_config(args, B, mode), transport=(
                transport + "-" + ph.step_impl.exchange_mode if transport == "peer" else transport),
                soft_target_tiles=("every tile (dense)" if flags_t is None else
                                   "flagged tiles only: %.4f of the %d x %d tiles can hold P_ij >= 2^-44" % (
                                       density, flags_t.numel() // ((B + 127) // 128), (B + 127) // 128))), "loss": loss_val,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": emax.item() / e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms": {"prepare": phase_ms[0] / args.steps, "stats": phase_ms[1] / args.steps,
                          "rowloss": phase_ms[2] / args.steps, "bwd": bwd_ms},
            "roofline": {"bound": "tensor", "kernel": "gradient sweep (mc_clip_bwd)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peaks["source"] + ", bf16 sustained (fp16 runs at the same tensor rate)",
                         "traffic": ncu.get("traffic_bytes_per_launch") if (mode == "tc_f16x3" and world == 1 and B == 32768) else None,
                         "executed_tflops": executed, "executed_frac": executed / peak,
                         "tile_flag_density": density,
                         "ncu_tensor_pipe_active_pct": ncu.get("tensor_pipe_active_pct"), "ncu_source": ncu.get("source"),
                         "algorithmic_flops_per_launch": flops_bwd,
                         "step_achieved": flops_step / (ms_per_step * 1e-3) / 1e12,
                         "step_frac": flops_step / (ms_per_step * 1e-3) / 1e12 / peak},
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_reference_leg(args, B, 2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_extra and world == 1:
            line["extra"] = extras(dev, mode, peaks)
            if mode == "tc_f16x3":
                line["extra"].update(single_pass_context(B, I_loc, T_loc, dev, flush, peaks))
        print(json.dumps(line))
    if world > 1:
        if transport == "peer":
            from mae_clip_b200 import peer
            peer.close_all()
        dist.destroy_process_group()


def single_pass_context(B, I_loc, T_loc, dev, flush, peaks):
    """The same workload on the single-pass engine (`tc_f16`: fp16 operands, fp32 accumulate; stated tolerance
    5e-4 loss / 5e-3 gradients instead of the headline's fp32-class 1e-4 / 1e-3): what the sweeps reach against the
    ALGORITHMIC flop count when the logits do not have to be fp32-exact.  Context only - never the headline."""
    import torch
    ph = Phases(B, B, 0, "tc_f16", dev)
    for _ in range(3):
        ph.step(I_loc, T_loc)
    torch.cuda.synchronize()
    n, tot, phs = 5, 0.0, [0.0] * 4
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(n):
        flush.fill_(1)
        torch.cuda.synchronize()
        a.record()
        ph.step(I_loc, T_loc, record=True)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
        for i, v in enumerate(ph.phase_ms()):
            phs[i] += v
    ms, stats_ms, bwd_ms = tot / n, phs[1] / n, phs[3] / n
    peak = peaks["tc_sustained"] or peaks["tc_burst"]
    unit = 2.0 * B * B * D_EMB / 1e9   # GFLOP of one B x B x D GEMM
    return {"c4_single_pass_tc_f16_B%d" % B: {
        "ms": ms, "samples_per_s": B / (ms * 1e-3), "loss": float(ph.part.item()),
        "phases_ms": {"prepare": phs[0] / n, "stats": stats_ms, "rowloss": phs[2] / n, "bwd": bwd_ms},
        "stats_sweep_algorithmic_TFLOPs": 3 * unit / stats_ms, "stats_sweep_frac": 3 * unit / stats_ms / peak,
        "gradient_sweep_algorithmic_TFLOPs": 4 * unit / bwd_ms, "gradient_sweep_frac": 4 * unit / bwd_ms / peak,
        "step_frac": 7 * unit / ms / peak, "tolerance": "loss 5e-4, gradients 5e-3 relative (tests/test_gpu_loss.py)"}}


def extras(dev, mode, peaks):
    """C2 latency and C5 MAE numbers (not the headline; same process, CUDA events, L2 not flushed
    for the latency figure, inputs > L2 for the MAE sweeps)."""
    import torch

    import mae_clip_b200 as m
    out = {}

    def timeit(fn, iters=20, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    I = make_shard(1024, 0).to(dev).requires_grad_(True)
    T = make_shard(1024, 1).to(dev).requires_grad_(True)

    def c2():
        I.grad = T.grad = None
        m.clip_contrastive_loss(I, T, 1.0, mode=mode).backward()
    ms = timeit(c2)
    out["c2_loss_fwd_bwd_B1024"] = {"ms": ms, "samples_per_s": 1024 / (ms * 1e-3)}

    # C2 again with the whole fwd+bwd captured in one CUDA graph (launch-bound regime)
    from mae_clip_b200.train import GraphedStep
    Ig, Tg = I.detach().clone().requires_grad_(True), T.detach().clone().requires_grad_(True)
    gs = GraphedStep(lambda a, b: m.clip_contrastive_loss(a, b, 1.0, mode=mode), [Ig, Tg], [])
    ms = timeit(lambda: gs.graph.replay())
    out["c2_loss_fwd_bwd_B1024_cuda_graph"] = {"ms": ms, "samples_per_s": 1024 / (ms * 1e-3)}

    # C5 (N = 1024, 196 patches x 768, ratio 0.75) through the C ABI with pre-allocated outputs; algorithmic bytes of
    # SURVEY 8(d); every working set is several times the 126 MB L2, so nothing is flushed
    from mae_clip_b200._lib import check as ck, cur_stream as cs, lib as L_, ptr as p_
    lib = L_()
    N, L, P = 1024, 196, 768
    keep = int(L * 0.25)
    x = torch.randn(N, L, P, device=dev)
    noise = torch.rand(N, L, device=dev)
    imgs = torch.randn(N, 3, 224, 224, device=dev)
    pred = torch.randn(N, L, P, device=dev)
    dpred = torch.empty_like(pred)
    xm = torch.empty(N, keep, P, device=dev)
    mask = torch.empty(N, L, device=dev)
    restore = torch.empty(N, L, device=dev, dtype=torch.int64)
    ids_keep = torch.empty(N, keep, device=dev, dtype=torch.int64)
    loss, msum = torch.empty((), device=dev), torch.empty((), device=dev)
    ws = torch.empty(lib.mc_masked_mse_workspace_bytes(N, L), dtype=torch.uint8, device=dev)

    def hbm(name, fn, nbytes, **kw):
        ms = timeit(fn, iters=10)
        out[name] = dict(ms=ms, GBps=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peaks["hbm"], **kw)
        return ms

    hbm("c5_random_masking_N1024_r075",
        lambda: ck(lib.mc_random_masking(p_(x), 4, p_(noise), N, L, P, keep, p_(xm), p_(mask), p_(restore), p_(ids_keep), cs())),
        16 * N * L + 2 * N * keep * P * 4)
    r_eff = (L - keep) / L
    bytes_f = r_eff * N * L * P * 8 + 4 * N * L
    ms_f = hbm("c5_masked_mse_fwd_N1024_r075",
               lambda: ck(lib.mc_masked_mse_fwd(p_(pred), 4, p_(imgs), p_(mask), N, 224, 224, 16, 1, p_(loss), p_(msum), p_(ws),
                                                ws.numel(), cs())), bytes_f)
    ms_b = hbm("c5_masked_mse_bwd_N1024_r075",
               lambda: ck(lib.mc_masked_mse_bwd(p_(pred), 4, p_(imgs), p_(mask), N, 224, 224, 16, 1, p_(msum), None, p_(dpred),
                                                cs())), bytes_f + N * L * P * 4)
    out["c5_masked_mse_fwd_bwd_N1024_r075"] = {"ms": ms_f + ms_b, "samples_per_s": N / ((ms_f + ms_b) * 1e-3)}
    del x, xm, pred, dpred, imgs

    # L1-L2: one ProjectionHead forward + backward (dropout on), image head E = 2048 (dx needed) and text head E = 768
    for E, need_dx in ((2048, True), (768, False)):
        Bh = 32768
        h = m.ProjectionHead(E, gemm_mode=mode).to(dev).train()
        xx = torch.randn(Bh, E, device=dev, requires_grad=need_dx)
        keepm = (torch.rand(Bh, 256, device=dev) > 0.1).to(torch.uint8)
        go = torch.randn(Bh, 256, device=dev)

        def head_step():
            for q in h.parameters():
                q.grad = None
            xx.grad = None
            h(xx, keep_mask=keepm).backward(go)
        ms = timeit(head_step, iters=5, warm=3)
        flops = 2.0 * Bh * 256 * (E + 256) * 3 - (0 if need_dx else 2.0 * Bh * E * 256)
        out[f"proj_head_fwd_bwd_B{Bh}_E{E}"] = {"ms": ms, "algorithmic_TFLOPs": flops / ms / 1e9,
                                                "samples_per_s": Bh / (ms * 1e-3)}
        del h, xx, keepm, go
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
