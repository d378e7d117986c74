"""world_size-2 gloo tests (CPU) of the N>1 host-side logic: the PPO gradient / advantage-moment
all-reduces reproduce the single-process result on the concatenated batch, and the env-shard
configuration reproduces single-run seeds (rank offsets)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crowdnav_prediction_attngraph_b200.ppo import allreduce_gradients, global_advantage_normalize
    torch.manual_seed(0)
    full = torch.randn(6, 8, 1)                       # advantages of the whole job: [T, N_total, 1]
    shard = full[:, rank * 4:(rank + 1) * 4]
    norm = global_advantage_normalize(shard.clone())
    lin = torch.nn.Linear(5, 3)
    torch.manual_seed(100 + rank)
    x = torch.randn(7, 5)
    lin(x).pow(2).mean().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    allreduce_gradients(list(lin.parameters()))
    torch.save(dict(norm=norm, local=local, avg=[p.grad.clone() for p in lin.parameters()]),
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_ppo_collectives_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % k)) for k in range(2)]
    torch.manual_seed(0)
    full = torch.randn(6, 8, 1)
    ref = (full - full.mean()) / (full.std() + 1e-5)
    got = torch.cat([r[0]["norm"], r[1]["norm"]], dim=1)
    assert torch.allclose(got, ref, atol=1e-6)
    for k in range(2):
        for a, l0, l1 in zip(r[k]["avg"], r[0]["local"], r[1]["local"]):
            assert torch.allclose(a, (l0 + l1) / 2, atol=1e-7)


def test_env_shards_reproduce_single_run_seeds():
    """Oracle-level statement of the sharding rule the CUDA engine implements (rank_offset / nenv_total):
    shard r of a G-way job seeds env k with seed + r*N + k and advances case_counter by the TOTAL env count."""
    from oracle.crowd_env import EnvConfig, OracleVecEnv
    import rvo2
    rvo2.ONLY_AGENT0 = True
    cfg = EnvConfig(human_num=5)
    whole = OracleVecEnv(cfg, 4, seed=11)
    parts = [OracleVecEnv(cfg, 2, seed=11, rank_offset=r * 2, nenv_total=4) for r in range(2)]
    ow = whole.reset()
    op = [p.reset() for p in parts]
    for k in ow:
        assert np.array_equal(ow[k], np.concatenate([o[k] for o in op]))
    a = np.random.RandomState(0).uniform(-1, 1, (4, 2)).astype(np.float32)
    for _ in range(30):
        ow, rw, dw, _ = whole.step(a)
        res = [p.step(a[r * 2:(r + 1) * 2]) for r, p in enumerate(parts)]
        for k in ow:
            assert np.array_equal(ow[k], np.concatenate([x[0][k] for x in res]))
        assert np.array_equal(dw, np.concatenate([x[2] for x in res]))
    assert [e.case_counter["train"] for e in whole.envs] == [e.case_counter["train"] for p in parts for e in p.envs]
