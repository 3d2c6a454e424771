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
    return tr.TrainConfig(hidden_dim=64, num_layers=2, envs=1024, horizon=24, chunk=8192, zero_heads=False, seed=5, kl_stats=True)


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
             kl=np.array([[s["kl_total"], s["kl_average"], s["kl_max"]] for s in stats]),
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
    # the KL statistic is global (sum and count all-reduced, the maximum max-reduced) and the same on both ranks
    np.testing.assert_array_equal(r[0]["kl"], r[1]["kl"])
    np.testing.assert_allclose(r[0]["kl"], [[s["kl_total"], s["kl_average"], s["kl_max"]] for s in stats], rtol=2e-2, atol=1e-6)
    np.testing.assert_allclose(r[0]["mom"], [t.moments.mu, t.moments.m2], rtol=1e-9)
    np.testing.assert_array_equal(r[0]["mom"], r[1]["mom"])
    for k, v in t.model.state_dict().items():
        np.testing.assert_array_equal(r[0][k], r[1][k])                        # replicas in lock-step
        np.testing.assert_allclose(r[0][k], v.cpu().numpy(), rtol=2e-3, atol=2e-5, err_msg=k)


def _cfg_uneven():
    """Uneven shards (1001 envs over 2 ranks), several minibatches, symmetry augmentation with rank-local random counts and
    dropout: every rank must run the same number of optimizer steps with globally-normalised gradients (round-1 advisor)."""
    from g2048 import trainer as tr
    return tr.TrainConfig(hidden_dim=64, num_layers=2, envs=1001, horizon=16, chunk=4096, zero_heads=False, seed=9, minibatches=3,
                          epochs=2, upsample_ratio=0.25, dropout=0.1)


def _worker_uneven(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "2048-ppo_b200"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from g2048 import trainer as tr
    t = tr.Trainer(_cfg_uneven(), torch.device("cuda", rank))
    stats = [t.train_step() for _ in range(2)]
    sd = {k: v.cpu().numpy() for k, v in t.model.state_dict().items()}
    np.savez(os.path.join(out_dir, f"u{rank}.npz"), loss=np.array([s["loss"] for s in stats]), n=np.array([t.B, t.n_update_samples]), **sd)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_uneven_shards_minibatches_augmentation_and_dropout_stay_in_lock_step(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker_uneven, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)      # a mismatch of all-reduces would hang here
    r = [np.load(tmp_path / f"u{i}.npz") for i in range(2)]
    assert int(r[0]["n"][0]) + int(r[1]["n"][0]) == 1001 and int(r[0]["n"][0]) != int(r[1]["n"][0])
    assert int(r[0]["n"][1]) != int(r[1]["n"][1])                     # different local sample counts (shard size, augmentation)
    assert np.isfinite(r[0]["loss"]).all()
    np.testing.assert_array_equal(r[0]["loss"], r[1]["loss"])          # the all-reduced statistics
    for k in r[0].files:
        if k not in ("loss", "n"):
            np.testing.assert_array_equal(r[0][k], r[1][k], err_msg=k)  # replicas bit-identical after 2 x 2 x 3 optimizer steps
