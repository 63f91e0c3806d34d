"""Frame sharding across the GPUs of one box.

Scans are independent until the grid merge, so a batch of F frames is split into
contiguous blocks, one per rank; the only exchange is the exact integer reduction inside
gv_grid_finalize_multi (NCCL over NVLink).  torch.distributed is used for rendezvous and
to hand the NCCL unique id from rank 0 to the other ranks; nothing else.
"""
from __future__ import annotations

import os


def shard_frames(nframes: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) block of frames for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(nframes, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device="cpu") -> bytes:
    """Broadcast a small byte string from `src` to every rank over torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def init_context_comm(ctx, device="cuda"):
    """Give every rank's Context the same NCCL communicator (id from rank 0)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return
    uid = ctx.nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0, device)
    ctx.nccl_init(uid, rank, world)


def enable_p2p(ctx, device="cuda") -> bool:
    """Map every rank's grid planes into every rank (cudaIpc, NVLink P2P) so that
    gv_grid_finalize_multi runs fused over peer memory.  Call after grid_init on every rank.
    Returns False (and leaves the NCCL path active) if any rank cannot map its peers."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return False
    mine = torch.frombuffer(bytearray(ctx.ipc_export()), dtype=torch.uint8).to(device)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    blobs = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allb)
    ok = 1
    try:
        ctx.ipc_import(blobs, world, rank)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(f"[grid_vision_b200] peer-memory mapping unavailable, staying on NCCL collectives: {e}")
        ok = 0
    t = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 0 and ok:
        ctx.ipc_close()  # another rank could not map its peers: every rank must use the same path
    return bool(int(t.item()))
