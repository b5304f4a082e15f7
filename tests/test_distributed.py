"""The N>1 path on CPU: world_size-2 `gloo` run of bench.py's rank logic -- each rank owns one sequence's records,
rank 0 gathers the 64-byte records (the only collective of the path) and chains poses per sequence exactly like
viso.cpp:1313-1321.  The records come from the CPU oracle here (no GPU); on the box the same gather runs over NCCL."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_seeds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libviso_b200 import api, synth
    from libviso_b200.distributed import gather_records, shard_sequences
    from oracle import oracle
    # 3 sequences over 2 ranks: rank 0 gets {0, 2}, rank 1 gets {1}
    mine = shard_sequences(3, rank, world)
    P1, P2 = synth.kitti_calib()
    recs = {}
    for s in mine:
        frames, _ = synth.make_sequence(3, seed=1000 + s, n_features=300)
        seeds = make_seeds(3, 20, seed=s)
        recs[s] = oracle.sequence(frames, P1, P2, oracle.param_default(ransac_iter=20), seeds)["records"]
    gathered = gather_records(recs, 3, 3, rank, world, device="cpu")
    if rank == 0:
        assert sorted(gathered) == [0, 1, 2]
        np.save(os.path.join(out_dir, "gathered.npy"), np.stack([gathered[s] for s in range(3)]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_chain_world2(tmp_path, api, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    for s in range(3):
        frames, _ = synth.make_sequence(3, seed=1000 + s, n_features=300)
        want = oracle.sequence(frames, P1, P2, oracle.param_default(ransac_iter=20), make_seeds(3, 20, seed=s))
        assert got[s].tobytes() == want["records"].tobytes()
        poses = api.chain_poses(got[s].view(api.RECORD_DTYPE).reshape(-1))
        assert np.array_equal(poses, want["poses"])
