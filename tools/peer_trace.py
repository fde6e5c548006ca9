"""Timeline of one row-sharded loss step over the peer transport (MAE_CLIP_PEER_TRACE=1: a CUDA event after every
exchange / sweep call of dist.PeerStep, printed by rank 0).  Launch with torchrun, one rank per GPU:
  MAE_CLIP_PEER_TRACE=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_trace.py [B]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mae_clip_b200.dist import PeerStep  # noqa: E402
from mae_clip_b200.peer import get_exchange, close_all  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
b, D = B // world, 256
g = torch.Generator().manual_seed(1000 + rank)
I = torch.nn.functional.layer_norm(torch.randn(b, D, generator=g), (D,)).cuda()
T = torch.nn.functional.layer_norm(torch.randn(b, D, generator=g), (D,)).cuda()
step = PeerStep(get_exchange(b, D), "tc_f16x3")
step.trace.on = False
for _ in range(3):
    loss, saved = step.forward(I, T, 1.0)
    step.backward(saved, 1.0)
torch.cuda.synchronize()
dist.barrier()
step.trace.on = os.environ.get("MAE_CLIP_PEER_TRACE", "0") == "1"
step.trace.marks = []
loss, saved = step.forward(I, T, 1.0)
step.backward(saved, 1.0)
if rank == 0:
    print("loss", loss.item())
close_all()
dist.destroy_process_group()
