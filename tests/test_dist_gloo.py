"""CPU, world_size 2 over gloo: the N>1 host logic (episode sharding + the per-step all-gather of episode metric
rows).  The per-rank simulation itself needs a GPU and is covered by the -m gpu tests; here each rank fabricates
the rows its shard would produce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddls_b200.dist import gather_episode_metrics, shard_range, shard_sizes


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_episodes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(n_episodes, rank, world)
    rows = torch.stack([torch.arange(lo, hi, dtype=torch.float64), torch.full((hi - lo,), float(rank), dtype=torch.float64)], dim=1)
    out = gather_episode_metrics(rows, n_episodes)
    if rank == 0:
        q.put(out.clone())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_episodes', [8, 7])
def test_all_gather_orders_episodes_globally(n_episodes):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_episodes, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (n_episodes, 2)
    assert out[:, 0].tolist() == list(range(n_episodes))
    sizes = shard_sizes(n_episodes, world)
    assert out[:, 1].tolist() == [0.0] * sizes[0] + [1.0] * sizes[1]


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 4096, 16384):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
