"""Full-size (60 blocks, D=3072, 24 heads, 8192+256 tokens) equality check of the two Ulysses forms under torchrun:
NCCL all-to-all vs fused peer-memory exchange, and both against rank 0's single-GPU forward.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29561 tools/check_fused_fullsize.py"""
import json, os, sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import qie_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = qie_b200.QwenImageDiTConfig()
model = qie_b200.B200QwenImageTransformer2DModel.from_random(cfg, seed=0, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(1, 8192, 64, generator=g, device=dev).bfloat16()
cond = (torch.randn(1, 219, cfg.joint_attention_dim, generator=g, device=dev) * 3).bfloat16()     # ragged text length
ts = torch.tensor([0.5], device=dev)
shapes = [[(1, 64, 64), (1, 64, 64)]]
single = model(x, cond, None, ts, shapes, [219], return_dict=False)[0]
nccl = qie_b200.UlyssesTransformer(model, None)(x, cond, None, ts, shapes, [219], return_dict=False)[0]
f = qie_b200.UlyssesTransformer(model, None, fused=True)
fused = f(x, cond, None, ts, shapes, [219], return_dict=False)[0]
torch.cuda.synchronize()
f.close()
rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
res = {"world": world, "rank": rank, "fused_vs_nccl": rel(fused, nccl), "fused_vs_single": rel(fused, single), "nccl_vs_single": rel(nccl, single),
       "barrier_timeouts": qie_b200.lib().qie_peer_barrier_timeouts()}
print(json.dumps(res), flush=True)
dist.destroy_process_group()
