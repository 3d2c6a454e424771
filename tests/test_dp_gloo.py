"""CPU, world_size 2 over gloo: the host-side data-parallel logic (env sharding, the flat
gradient all-reduce with global-count normalisation, the 3-scalar moment all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "2048-ppo_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from g2048 import dp, ppo
    from g2048.policy import GameMLP, MLPConfig

    assert dp.world() == (rank, world)
    # global batch of 1000 samples, sharded like the envs
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1000, 48), generator=g)
    y = torch.randn((1000, 5), generator=g)
    G = torch.randn(1000, generator=g, dtype=torch.float64) * 30 + 5
    lo, hi = dp.shard_range(1000, rank, world)
    torch.manual_seed(7)
    model = GameMLP(MLPConfig(hidden_dim=32, num_layers=1, dropout=0.0))
    bucket = dp.FlatGradBucket(model.parameters())
    logits, v = model(x[lo:hi])
    # every rank normalises by the GLOBAL count; the all-reduced sum is then the global-mean gradient
    loss = ((torch.cat([logits, v], 1) - y[lo:hi]) ** 2).sum() / 1000
    loss.backward()
    bucket.allreduce()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    # moments: all-reduce {sum G, sum G^2, N}
    s = torch.tensor([G[lo:hi].sum(), (G[lo:hi] ** 2).sum(), float(hi - lo)], dtype=torch.float64)
    dp.allreduce_stats(s)
    mom = ppo.RtgMoments(mu=1.0, m2=50.0, step=4)
    mom.update(0.99, *s.tolist())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), grad=flat.numpy(), mom=np.array([mom.mu, mom.m2]),
             lohi=np.array([lo, hi]), nparam=bucket.numel())
    dist.destroy_process_group()


def test_dp_world2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    assert r[0]["lohi"].tolist() == [0, 500] and r[1]["lohi"].tolist() == [500, 1000]
    np.testing.assert_array_equal(r[0]["grad"], r[1]["grad"])          # replicas stay in lock-step
    np.testing.assert_array_equal(r[0]["mom"], r[1]["mom"])
    # single-process reference on the whole batch
    from g2048 import ppo
    from g2048.policy import GameMLP, MLPConfig
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1000, 48), generator=g)
    y = torch.randn((1000, 5), generator=g)
    G = torch.randn(1000, generator=g, dtype=torch.float64) * 30 + 5
    torch.manual_seed(7)
    model = GameMLP(MLPConfig(hidden_dim=32, num_layers=1, dropout=0.0))
    logits, v = model(x)
    (((torch.cat([logits, v], 1) - y) ** 2).sum() / 1000).backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert int(r[0]["nparam"]) == flat.numel()
    np.testing.assert_allclose(r[0]["grad"], flat.numpy(), rtol=1e-4, atol=1e-6)
    mom = ppo.RtgMoments(mu=1.0, m2=50.0, step=4)
    mom.update(0.99, float(G.sum()), float((G ** 2).sum()), 1000.0)
    np.testing.assert_allclose(r[0]["mom"], [mom.mu, mom.m2], rtol=1e-12)


def test_shard_range_partitions_everything():
    from g2048 import dp
    for total in (1, 7, 1 << 20, 1000003):
        for w in (1, 2, 3, 4, 8):
            parts = [dp.shard_range(total, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
