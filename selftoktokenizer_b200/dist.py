"""Data-parallel sharding of the path: images are independent, weights are replicated, and the only
exchange is one all-gather of token ids ([B/G, K] int64 per rank) after encode — NCCL over NVLink on GPUs,
gloo in the CPU tests.  No tensor/pipeline/sequence parallelism exists or is needed on this path."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_slice(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items owned by `rank`; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tokens(local_tokens: torch.Tensor, n_total: int) -> torch.Tensor:
    """all-gather ragged per-rank token blocks into the global [n_total, K] tensor (same on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_tokens
    world, rank = dist.get_world_size(), dist.get_rank()
    K = local_tokens.shape[1]
    per = (n_total + world - 1) // world
    pad = torch.zeros(per, K, dtype=local_tokens.dtype, device=local_tokens.device)
    pad[: local_tokens.shape[0]] = local_tokens
    out = torch.empty(world * per, K, dtype=local_tokens.dtype, device=local_tokens.device)
    dist.all_gather_into_tensor(out, pad)
    rows = []
    for r in range(world):
        lo, hi = shard_slice(n_total, r, world)
        rows.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(rows, dim=0)


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def host_noise(n_total: int, shape, seed: int) -> torch.Tensor:
    """The sampler's initial noise for the WHOLE batch, drawn once on the host exactly as the reference draws it
    (torch.randn on a CPU generator, SelftokPipeline.py:262-264) -- every rank evaluates the same deterministic draw and keeps
    its slice, so a sharded run starts from the noise a single process would have used (SURVEY 8e RNG-parity rule)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.randn(n_total, *shape, generator=g)


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all-gather of ragged per-rank row blocks of any trailing shape / dtype (latents, pixels) into [n_total, ...]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    w = dist.get_world_size()
    per = (n_total + w - 1) // w
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(w * per, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    rows = []
    for r in range(w):
        lo, hi = shard_slice(n_total, r, w)
        rows.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(rows, dim=0)
