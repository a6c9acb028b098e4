"""Multi-GPU path on real devices: the batch is sharded contiguously over one process per GPU, every
rank runs the cascade on its own device, and the NCCL all-gather returns the feature matrix in input order — the same
bytes the single-GPU run produces.  Skipped on boxes with one GPU (the world_size-2 gloo test covers the host logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, x_path, out_dir):
    import wst_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    x = torch.load(x_path)
    B = x.shape[0]
    lo, hi = wst_b200.shard_range(B, rank, world)
    local = wst_b200.scattering_features(x[lo:hi].cuda(), 3, 8)                # [b_r, C*2*K] on this rank's GPU
    full = wst_b200.gather_features(local, B)
    torch.save(full.cpu(), os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [13, 8])
def test_nccl_sharded_features_equal_single_gpu(tmp_path, B):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import wst_b200
    rng = np.random.default_rng(B)
    x = torch.from_numpy((rng.integers(0, 256, (B, 3, 64, 64)) / 255.0).astype(np.float32))
    x_path = os.path.join(tmp_path, "x.pt")
    torch.save(x, x_path)
    ref = wst_b200.scattering_features(x.cuda(), 3, 8).cpu()
    mp.spawn(_worker, args=(world, _free_port(), x_path, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(tmp_path, "r%d.pt" % r)), ref)
