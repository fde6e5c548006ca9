"""Global-batch contrastive loss choreography (mae_clip_b200/dist.py) at world_size 2 on CPU/gloo.

The sharded loss and gradients must equal the single-process reference loss (CLIP.py:34-43) on the
concatenated batch - that is the definition of the global-batch loss (SURVEY.md section 8 e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err
from oracle import loss_ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, b, tau, scale, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _torch_engine import TorchStripEngine
        from mae_clip_b200.dist import global_clip_loss
        I = loss_ref.make_embeddings(b, 256, seed=1000 + rank, scale=scale).requires_grad_(True)
        T = loss_ref.make_embeddings(b, 256, seed=2000 + rank, scale=scale).requires_grad_(True)
        loss = global_clip_loss(I, T, tau, engine=TorchStripEngine())
        (loss * 3.0).backward()  # non-unit upstream gradient
        ret[rank] = (loss.detach(), I.grad.clone(), T.grad.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("b,tau,scale", [(24, 1.0, 0.1), (17, 0.5, 1.0)])
def test_global_loss_world2_matches_single_process(b, tau, scale):
    world = 2
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), b, tau, scale, ret), nprocs=world, join=True)
    I = torch.cat([loss_ref.make_embeddings(b, 256, seed=1000 + r, scale=scale) for r in range(world)])
    T = torch.cat([loss_ref.make_embeddings(b, 256, seed=2000 + r, scale=scale) for r in range(world)])
    l64, dI64, dT64 = loss_ref.clip_loss_fwd_bwd_ref(I, T, tau, dtype=torch.float64)
    for r in range(world):
        loss, dI, dT = ret[r]
        assert abs(loss.item() - l64.item()) < 1e-5 * abs(l64.item())
        assert rel_err(dI, 3.0 * dI64[r * b:(r + 1) * b]) < 1e-4  # statistics cross ranks as fp32
        assert rel_err(dT, 3.0 * dT64[r * b:(r + 1) * b]) < 1e-4


def test_world1_path_without_process_group():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _torch_engine import TorchStripEngine
    from mae_clip_b200.dist import global_clip_loss
    I = loss_ref.make_embeddings(20, 256, seed=1, scale=0.1).requires_grad_(True)
    T = loss_ref.make_embeddings(20, 256, seed=2, scale=0.1).requires_grad_(True)
    loss = global_clip_loss(I, T, 2.0, engine=TorchStripEngine())
    loss.backward()
    l64, dI64, dT64 = loss_ref.clip_loss_fwd_bwd_ref(I.detach(), T.detach(), 2.0, dtype=torch.float64)
    assert abs(loss.item() - l64.item()) < 1e-5 * abs(l64.item())
    assert rel_err(I.grad, dI64) < 1e-5 and rel_err(T.grad, dT64) < 1e-5
