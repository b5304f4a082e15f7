"""Multi-GPU plumbing of the hot path: sequences (and frame pairs) are independent (viso.cpp:1208-1222 vs
:1317-1321: the pose never feeds back), so ranks own whole sequences and the ONLY collective is the gather of the
64-byte per-frame-pair records to rank 0, which then chains poses like viso.cpp:1313-1321.  NCCL on the GPU box
(bench.py gathers straight from the device record buffers, viso_seq_records_device, on the library's stream), gloo in
the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist

from .api import RECORD_DTYPE

RECORD_WORDS = RECORD_DTYPE.itemsize // 4   # one record = 16 int32 words


def shard_sequences(n_sequences, rank, world):
    """sequence s is owned by rank s % world"""
    return [s for s in range(n_sequences) if s % world == rank]


def rows_per_rank(n_sequences, world):
    return (n_sequences + world - 1) // world


def gather_rows(buf, rank, world, gathered=None):
    """the collective: every rank contributes `buf` (rows_per_rank x words int32, device or host tensor); rank 0
    receives the list of all ranks' buffers (preallocated `gathered` is reused), the others None.  Enqueued on the
    current CUDA stream for device tensors."""
    if world == 1:
        return [buf]
    if rank == 0 and gathered is None:
        gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.gather(buf, gathered if rank == 0 else None, dst=0)
    return gathered if rank == 0 else None


def split_gathered(gathered, n_sequences, world):
    """rank 0: the gathered buffers -> {sequence: RECORD_DTYPE[n_frames]}"""
    out = {}
    for r in range(world):
        g = gathered[r].cpu().numpy()
        for j, s in enumerate(shard_sequences(n_sequences, r, world)):
            out[s] = g[j].view(RECORD_DTYPE).reshape(-1).copy()
    return out


def gather_records(records_by_seq, n_sequences, n_frames, rank, world, device="cuda"):
    """records_by_seq: {sequence: RECORD_DTYPE[n_frames]} (host arrays) for the sequences this rank owns.
    Returns {sequence: records} on rank 0 (all sequences), None elsewhere."""
    buf = torch.zeros((rows_per_rank(n_sequences, world), n_frames * RECORD_WORDS), dtype=torch.int32)
    for j, s in enumerate(shard_sequences(n_sequences, rank, world)):
        buf[j] = torch.from_numpy(np.ascontiguousarray(records_by_seq[s]).view(np.int32).reshape(-1).copy())
    gathered = gather_rows(buf.to(device), rank, world)
    return split_gathered(gathered, n_sequences, world) if rank == 0 else None
