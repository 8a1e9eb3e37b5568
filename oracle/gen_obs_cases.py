"""TEST INFRASTRUCTURE -- records observations of the unmodified reference's RampJobPartitioningObservation in seeded
episodes together with everything the encoder read (job graph, job details, jobs_params, cluster scalars) as
tests/fixtures/obs_cases.npz, the fixture ddls_b200/observation.py is pinned against (tests/test_observation.py).
Build container only (needs the reference)."""
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden as G  # noqa: E402  (installs the import shim, imports the reference)
from ddls.environments.ramp_job_partitioning.observations import ramp_job_partitioning_observation as O  # noqa: E402

CASES = ('chain8_busy', 'mixed16', 'mixed64_busy', 'tfm32_acceptable')
PARAMS = ('job_total_num_ops', 'job_total_num_deps', 'job_sequential_completion_times', 'max_acceptable_job_completion_times',
          'max_acceptable_job_completion_time_fracs', 'job_total_op_memory_costs', 'job_total_dep_sizes', 'job_num_training_steps')


def main():
    out, n = {}, 0
    orig = O.RampJobPartitioningObservation._encode_obs

    def _encode_obs(self, job, env, flatten=True):
        nonlocal n
        obs = orig(self, job, env, flatten=flatten)
        cl = env.cluster
        dev = list(cl.topology.graph.graph['worker_types'])[0]
        g = job.computation_graph
        nodes = list(g.nodes)
        idx = {v: i for i, v in enumerate(nodes)}
        edges = list(g.edges)
        p = f'c{n}_'

        def which(x):      # the encoder compares the op id with this object with ==; anything but an op id never matches
            try:
                return idx[x] if x in idx else -1
            except TypeError:
                return -1
        out[p + 'op_compute'] = np.array([g.nodes[v]['compute_cost'][dev] for v in nodes], dtype=np.float64)
        out[p + 'op_memory'] = np.array([g.nodes[v]['memory_cost'] for v in nodes], dtype=np.float64)
        out[p + 'op_depth'] = np.array([job.details['node_to_depth'][v] for v in nodes], dtype=np.float64)
        out[p + 'edge_src'] = np.array([idx[e[0]] for e in edges], dtype=np.int64)
        out[p + 'edge_dst'] = np.array([idx[e[1]] for e in edges], dtype=np.int64)
        out[p + 'edge_size'] = np.array([g[e[0]][e[1]][e[2]]['size'] for e in edges], dtype=np.float64)
        jp = cl.jobs_generator.jobs_params
        out[p + 'params'] = np.array([[jp['min_' + k], jp['max_' + k]] for k in PARAMS], dtype=np.float64)
        out[p + 'scalars'] = np.array([
            job.details['max_compute_cost'][dev], which(job.details['max_compute_node']), job.details['max_memory_cost'],
            which(job.details['max_memory_node']), job.details['max_dep_size'], edges.index(job.details['max_dep_size_dep']),
            job.details['max_depth'], job.details['job_sequential_completion_time'][dev],
            job.details['max_acceptable_job_completion_time'][dev], job.max_acceptable_job_completion_time_frac,
            job.details['job_total_op_memory_cost'], job.details['job_total_dep_size'], job.num_training_steps,
            len(cl.mounted_workers), len(cl.jobs_running), cl.topology.graph.graph['num_workers'],
            cl.topology.num_communication_groups, cl.topology.num_racks_per_communication_group, cl.topology.num_servers_per_rack,
            env.max_partitions_per_op, self.pad_obs_kwargs['max_nodes'], self.machine_epsilon], dtype=np.float64)
        for k, v in obs.items():
            out[p + 'obs_' + k] = np.asarray(v)
        n += 1
        return obs
    O.RampJobPartitioningObservation._encode_obs = _encode_obs
    try:
        for name in CASES:
            spec = G.CASES[name]
            np.random.seed(spec['seed']); random.seed(spec['seed'])
            d = tempfile.mkdtemp(prefix='obs_')
            for g in spec['graphs']:
                g.write(d)
            env = G.make_env(d, spec['shape'], spec['n_jobs'], spec['max_partitions'], spec['interarrival'],
                             G.Uniform(spec['frac'][0], spec['frac'][1], decimals=2), max_sim_time=spec.get('max_sim_time', 1e6))
            np.random.seed(spec['seed']); random.seed(spec['seed'])
            n0 = n
            obs = env.reset()
            actor = {'random': G.Random(), 'sipml': G.SiPML(spec['max_partitions']), 'acceptable_jct': G.AcceptableJCT()}[spec['actor']]
            done = False
            while not done:
                job = list(env.cluster.job_queue.jobs.values())[0]
                obs, _, done, _ = env.step(int(actor.compute_action(obs, job_to_place=job)))
            print(name, n - n0, 'observations recorded')
    finally:
        O.RampJobPartitioningObservation._encode_obs = orig
    out['n_cases'] = np.array(n)
    path = os.path.join(ROOT, 'tests', 'fixtures', 'obs_cases.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB,', n, 'cases')


if __name__ == '__main__':
    main()
