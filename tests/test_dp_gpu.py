"""Multi-GPU (NCCL) check, skipped on a single-GPU box: a 2-rank train step must reproduce the
1-GPU step on the same global env set (same Philox env ids => identical rollouts; the all-reduced
gradient equals the global-batch gradient up to fp32 summation order)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg():
    from g2048 import trainer as tr
    return tr.TrainConfig(hidden_dim=64, num_layers=2, envs=1024, horizon=24, chunk=8192, zero_heads=False, seed=5)


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "2048-ppo_b200"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from g2048 import trainer as tr
    t = tr.Trainer(_cfg(), torch.device("cuda", rank))
    stats = [t.train_step() for _ in range(2)]
    sd = {k: v.cpu().numpy() for k, v in t.model.state_dict().items()}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=np.array([s["loss"] for s in stats]),
             mom=np.array([t.moments.mu, t.moments.m2]), boards=t.boards.cpu().numpy(), **sd)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from g2048 import trainer as tr
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(2)]
    t = tr.Trainer(_cfg(), torch.device("cuda:0"))
    stats = [t.train_step() for _ in range(2)]
    # rollouts are bit-identical (shard invariance), so are the sharded final boards
    np.testing.assert_array_equal(np.concatenate([r[0]["boards"], r[1]["boards"]]), t.boards.cpu().numpy())
    np.testing.assert_allclose(r[0]["loss"], [s["loss"] for s in stats], rtol=1e-4)
    np.testing.assert_allclose(r[0]["mom"], [t.moments.mu, t.moments.m2], rtol=1e-9)
    np.testing.assert_array_equal(r[0]["mom"], r[1]["mom"])
    for k, v in t.model.state_dict().items():
        np.testing.assert_array_equal(r[0][k], r[1][k])                        # replicas in lock-step
        np.testing.assert_allclose(r[0][k], v.cpu().numpy(), rtol=2e-3, atol=2e-5, err_msg=k)
