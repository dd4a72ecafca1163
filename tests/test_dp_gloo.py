"""World-size-2 gloo test of the data-parallel host logic (stmgcn_b200/dp.py): shard the batch, all-reduce the
flat gradient bucket, divide by world -> identical to the single-process full-batch gradient
(MSELoss(reduction='mean'), equal shards).  The model here is the CPU oracle (the CUDA
path needs a GPU); what is under test is the sharding / bucket / all-reduce plumbing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 1)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(repo, "st-mgcn_b200"))
    from stmgcn_b200 import dp
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    model = _Tiny()
    bucket = dp.GradBucket(model)
    gen = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=gen), torch.randn(8, 1, generator=gen)
    xs, ys = dp.shard_batch(x, rank, world), dp.shard_batch(y, rank, world)
    for _ in range(2):                       # second pass checks zero_() + in-place accumulation into the bucket
        bucket.zero_()
        torch.nn.functional.mse_loss(model(xs), ys).backward()
        bucket.all_reduce_mean_()
    # third pass the way the reference trainer does it (Model_Trainer.py:40-41): optimizer.zero_grad() defaults to
    # set_to_none=True and drops the views into the bucket; all_reduce_mean_ must notice and re-bind, not average zeros
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    opt.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    torch.nn.functional.mse_loss(model(xs), ys).backward()
    bucket.all_reduce_mean_()
    assert all(p.grad.data_ptr() == bucket.flat.data_ptr() + o * 4 for p, o in zip(bucket.params, bucket._offsets))
    np.save(os.path.join(out_dir, f"g{rank}.npy"), bucket.flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1)
    model = _Tiny()
    gen = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=gen), torch.randn(8, 1, generator=gen)
    torch.nn.functional.mse_loss(model(x), y).backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    assert np.allclose(g0, full, rtol=1e-5, atol=1e-7)


def test_shard_batch_requires_equal_shards():
    import pytest
    from stmgcn_b200 import dp
    with pytest.raises(ValueError):
        dp.shard_batch(torch.zeros(5, 2), 0, 2)
    assert dp.shard_batch(torch.arange(8).reshape(8, 1), 1, 4).flatten().tolist() == [2, 3]
