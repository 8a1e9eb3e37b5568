"""cProfile of BatchedRampJobPartitioningEnvironment.step on the bench workload (run on the GPU box)."""
import cProfile
import pstats
import sys

import numpy as np

sys.path.insert(0, '.')
from ddls_b200 import workload
from ddls_b200.batched import BatchedRampJobPartitioningEnvironment, DeviceRampJobPartitioningEnvironment

cfg = workload.CONFIGS['cfg3-resnet50-64w']
graphs = [workload.make_graph(k, **kw) for k, kw in cfg['graphs']]
device = len(sys.argv) > 1 and sys.argv[1] == "device"
env = (DeviceRampJobPartitioningEnvironment(tuple(cfg["shape"]), graphs, n_episodes=4096, jobs_per_episode=8, seed=1, prewarm=True) if device
       else BatchedRampJobPartitioningEnvironment(tuple(cfg["shape"]), graphs, n_episodes=4096, jobs_per_episode=8, seed=1))
degs = np.array([2, 4, 8, 16])
rng = np.random.default_rng(0)


def policy(obs):
    ok = obs['action_mask'][:, degs].astype(bool)
    r = rng.random(ok.shape) * ok
    return np.where(ok.any(axis=1), degs[r.argmax(axis=1)], 0)


def run(n):
    obs = env.reset()
    for s in range(n):
        if s % 8 == 0 and s:
            obs = env.reset()
        obs, _, _, _ = env.step(policy(obs))


run(8)
pr = cProfile.Profile()
pr.enable()
run(64)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
print(env.stats)
