"""Multi-GPU plumbing of the hot path: sequences (and frame pairs) are independent (viso.cpp:1208-1222 vs
:1317-1321: the pose never feeds back), so ranks own whole sequences and the ONLY collective is the gather of the
64-byte per-frame-pair records to rank 0, which then chains poses like viso.cpp:1313-1321.  NCCL on the GPU box,
gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist

from .api import RECORD_DTYPE


def shard_sequences(n_sequences, rank, world):
    """sequence s is owned by rank s % world"""
    return [s for s in range(n_sequences) if s % world == rank]


def gather_records(records_by_seq, n_sequences, n_frames, rank, world, device="cuda"):
    """records_by_seq: {sequence: RECORD_DTYPE[n_frames]} for the sequences this rank owns.
    Returns {sequence: records} on rank 0 (all sequences), None elsewhere."""
    per_rank = (n_sequences + world - 1) // world
    words = n_frames * RECORD_DTYPE.itemsize // 4
    buf = torch.zeros((per_rank, words), dtype=torch.int32)
    mine = shard_sequences(n_sequences, rank, world)
    for j, s in enumerate(mine):
        buf[j] = torch.from_numpy(np.ascontiguousarray(records_by_seq[s]).view(np.int32).reshape(-1).copy())
    buf = buf.to(device)
    if world == 1:
        gathered = [buf]
    else:
        gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    out = {}
    for r in range(world):
        g = gathered[r].cpu().numpy()
        for j, s in enumerate(shard_sequences(n_sequences, r, world)):
            out[s] = g[j].view(RECORD_DTYPE).reshape(-1).copy()
    return out
