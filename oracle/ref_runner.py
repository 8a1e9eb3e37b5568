"""BENCH INFRASTRUCTURE -- times the UNMODIFIED Python reference on the bench's workload (one process of bench.py's
``--impl reference`` arm; also usable by hand):

    python oracle/ref_runner.py --config cfg3-resnet50-64w --steps 4 --warmup 1 --budget 120 --seed 0

Builds ``RampJobPartitioningEnvironment`` (the reference's own class, imported through oracle/ref_shim.py from
the build container's checkout or the copy staged at oracle/_ref) with the reference's own heuristic agents (ramp first-fit op placer,
SRPT schedulers, first-fit dep placer: heuristic_config.yaml:191-197) on the config's topology and synthetic job graphs,
and takes env-steps whose action is a partition degree drawn uniformly from the config's degrees -- the same decision
rule bench.py scripts for the GPU arm (the role the PAC-ML policy plays in BASELINE.json config 3).  Prints one JSON line:
{"steps": n, "elapsed_s": s, ...}.  One env-step = one ``RampJobPartitioningEnvironment.step`` (RJPE:300-420)."""
import argparse
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='cfg3-resnet50-64w')
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--budget', type=float, default=120.0, help='stop taking timed steps after this many seconds')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--interarrival', type=float, default=1000.0)
    args = ap.parse_args()

    t_import = time.perf_counter()
    from oracle import ref_shim
    ref_shim.install()
    from ddls.environments.ramp_job_partitioning.ramp_job_partitioning_environment import RampJobPartitioningEnvironment
    from ddls.devices.processors.gpus.A100 import A100
    from ddls.distributions.fixed import Fixed
    from ddls.distributions.uniform import Uniform
    from ddls_b200 import workload
    t_import = time.perf_counter() - t_import

    cfg = workload.CONFIGS[args.config]
    c, r, s = cfg['shape']
    degrees = [d for d in cfg['degrees'] if d <= c * r * s]
    np.random.seed(args.seed)
    random.seed(args.seed)
    d = tempfile.mkdtemp(prefix='ref_runner_graphs_')
    for kind, kw in cfg['graphs']:
        workload.make_graph(kind, **kw).write(d)
    env = RampJobPartitioningEnvironment(
        topology_config={'type': 'ramp', 'kwargs': {'num_communication_groups': c, 'num_racks_per_communication_group': r,
                                                    'num_servers_per_rack': s, 'num_channels': 1, 'total_node_bandwidth': 1.6e12,
                                                    'intra_gpu_propagation_latency': 50e-9, 'worker_io_latency': 100e-9}},
        node_config={'type_1': {'num_nodes': c * r * s, 'workers_config': [{'num_workers': 1, 'worker': A100}]}},
        jobs_config={'path_to_files': d, 'job_interarrival_time_dist': Fixed(val=args.interarrival),
                     'max_acceptable_job_completion_time_frac_dist': Uniform(0.1, 1.0, decimals=2),
                     'replication_factor': 64, 'job_sampling_mode': 'remove', 'num_training_steps': 50, 'shuffle_files': True},
        max_partitions_per_op=max(degrees), min_op_run_time_quantum=0.01, reward_function='job_acceptance',
        reward_function_kwargs={'fail_reward': -1, 'success_reward': 1}, pad_obs_kwargs={'max_nodes': 1000},
        max_simulation_run_time=1e9, suppress_warnings=True)
    rng = np.random.default_rng(args.seed)
    obs = env.reset()

    def one_step(obs):
        valid = set(int(a) for a in obs['action_set'][obs['action_mask'].astype(bool)])
        cand = [dg for dg in degrees if dg in valid]
        a = int(rng.choice(cand)) if cand else 0
        obs, _, done, _ = env.step(a)
        if done:
            obs = env.reset()
        return obs, a

    for _ in range(args.warmup):
        obs, _ = one_step(obs)
    taken, acts = 0, []
    t0 = time.perf_counter()
    while taken < args.steps:
        obs, a = one_step(obs)
        acts.append(a)
        taken += 1
        if time.perf_counter() - t0 >= args.budget:
            break
    dt = time.perf_counter() - t0
    print(json.dumps({'steps': taken, 'elapsed_s': dt, 'actions': acts, 'import_s': t_import, 'config': args.config,
                      'reference_root': ref_shim.REFERENCE_ROOT}), flush=True)


if __name__ == '__main__':
    main()
