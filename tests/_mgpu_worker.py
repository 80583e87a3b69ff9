"""Worker of tests/test_multigpu.py: one process per GPU (torchrun), NCCL.  Checks, across REAL ranks, that the sharded entry
points of SelftokPipeline reproduce the single-process result bit for bit (shard invariance)."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import SelftokPipeline, config as C, synth  # noqa: E402
from selftoktokenizer_b200 import dist as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = C.MID
sd = synth.synth_state_dict(d)
pipe = SelftokPipeline(cfg=None, ckpt_path=None, sd3_path=None, datasize=d.latent * 8, device=dev, state_dict=sd, dims=d, precision="fp16")
n = 5                                                   # ragged over 2 ranks (3 + 2)
x0 = synth.synth_tensor("mgpu.x0", (n, d.in_channels, d.latent, d.latent), "emb", 1.0)
tok = pipe.encode_latents_sharded(x0)
tok_single = pipe.encode_latents(x0)                    # the whole batch on this rank alone
assert torch.equal(tok, tok_single), "sharded token ids differ from the single-process ids"
lat = pipe.decode_latents_sharded(tok.cpu().numpy(), seed=7)
noise = D.host_noise(n, (d.in_channels, d.latent, d.latent), 7)
lat_single = pipe.decode_latents(tok.cpu().numpy(), noise=noise)
assert torch.equal(lat, lat_single), "sharded latents differ from the single-process latents"
# every rank holds the same global results
chk = torch.tensor([float(tok.double().sum()), float(lat.double().abs().sum())], device=dev, dtype=torch.float64)
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(torch.equal(a, allc[0]) for a in allc)
dist.barrier()
if rank == 0:
    print(f"MGPU_OK world={world} tokens={tuple(tok.shape)} latents={tuple(lat.shape)}")
dist.destroy_process_group()
