"""world_size-2 gloo tests (CPU): the data-parallel protocol of SURVEY 8e restated with the oracle --
each rank scales its shard's gradients by 1/n_global, a SUM all-reduce yields the global-batch gradient,
identical optimizer steps keep the weights bit-identical -- plus inference sharding without a collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from helpers import head_oracle, synthetic_features, synthetic_poses
    from hpose_b200 import keras_spec as K, train_96
    from hpose_b200.parallel import DataParallel, shard_bounds
    from oracle.keras_graph import apply_optimizer
    dp = DataParallel()
    assert (dp.rank, dp.world_size) == (rank, world)
    train_96.config.update(num_filters=16, dropout_rate=0.0, regularizer_rate=1e-4, optimizer="adam")
    K.reset_names(); K.set_seed(3)
    m = train_96.create_model()
    x = synthetic_features(130, 96, seed=1, sigma=0.55, p=0.31).reshape(130, 1, 1, 96)
    y = synthetic_poses(130, seed=2).reshape(130, 1, 1, 3)
    g, params = head_oracle(m, torch.float64, requires_grad=True)
    keys = list(params)
    state = {}
    n_global = 130
    for step in range(3):
        idx = np.arange(n_global)[rank::world]                 # the same slicing Model.fit uses
        for p in params.values():
            p.grad = None
        pred = g(torch.tensor(x[idx], dtype=torch.float64), training=True)
        data_loss = ((pred - torch.tensor(y[idx], dtype=torch.float64)) ** 2).sum() / (n_global * 3)
        data_loss.backward()
        flat = torch.cat([params[k].grad.reshape(-1) for k in keys])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        grads, off = {}, 0
        for k in keys:
            n = params[k].numel()
            lam = 1e-4
            grads[k] = flat[off:off + n].reshape(params[k].shape) + 2 * lam * params[k].detach()
            off += n
        apply_optimizer(params, grads, {"name": "adam", "learning_rate": 2.8e-4}, state)
    final = torch.cat([params[k].detach().reshape(-1) for k in keys])
    gathered = [torch.zeros_like(final) for _ in range(world)]
    dist.all_gather(gathered, final)
    if rank == 0:
        # single-process reference on the full batch
        K.reset_names(); K.set_seed(3)
        m1 = train_96.create_model()
        g1, p1 = head_oracle(m1, torch.float64, requires_grad=True)
        from oracle.keras_graph import keras_train_step
        st = {}
        for step in range(3):
            keras_train_step(g1, p1, x, y, {"name": "adam", "learning_rate": 2.8e-4}, st)
        ref = torch.cat([p1[k].detach().reshape(-1) for k in keys])
        out.put({"identical": bool(torch.equal(gathered[0], gathered[1])),
                 "max_dev": float((gathered[0] - ref).abs().max()),
                 "shards": [shard_bounds(4099, r, world) for r in range(world)]})
    dist.barrier()
    dist.destroy_process_group()


def test_dp_protocol_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["identical"], "ranks diverged"
    assert res["max_dev"] < 1e-12, res
    assert res["shards"] == [(0, 2050), (2050, 4099)]
