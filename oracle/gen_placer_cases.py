"""TEST INFRASTRUCTURE -- records what the unmodified reference's RampFirstFitOpPlacer decided in seeded busy-cluster episodes
(inputs: cluster occupancy per server, forward graph, split counts; output: op -> server) as tests/fixtures/placer_cases.json,
the fixture ddls_b200/placer.py is pinned against (tests/test_placer.py).  Build container only (needs the reference)."""
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden as G  # noqa: E402  (installs the import shim, imports the reference)
from ddls.environments.ramp_cluster.agents.placers import ramp_first_fit_op_placer as P  # noqa: E402
from ddls.environments.ramp_cluster.agents.placers.utils import dummy_ramp  # noqa: E402
from ddls.utils import get_forward_graph  # noqa: E402

CASES = ('chain8_busy', 'mixed16', 'mixed64_busy', 'res16_flood', 'tfm32_acceptable', 'residual32_deg16')


def main():
    records = []
    orig_get = P.RampFirstFitOpPlacer.get

    def get(self, op_partition, cluster, verbose=False):
        shape = (cluster.topology.num_communication_groups, cluster.topology.num_racks_per_communication_group,
                 cluster.topology.num_servers_per_rack)
        before = dummy_ramp(shape, cluster)
        pre = {}
        for key in op_partition.action.keys():
            job = op_partition.partitioned_jobs[key]
            original = cluster.job_queue.jobs[job.job_id]
            fg = get_forward_graph(original.computation_graph)
            pre[job.job_id] = dict(
                job_idx=int(job.details['job_idx']),
                nodes=[str(n) for n in fg.nodes()], mem=[float(fg.nodes[n]['memory_cost']) for n in fg.nodes()],
                edges=[[str(u), str(v)] for (u, v) in fg.edges()],
                in_edges={str(n): [str(e[0]) for e in fg.in_edges(n)] for n in fg.nodes()},
                out_edges={str(n): [str(e[1]) for e in fg.out_edges(n)] for n in fg.nodes()},
                mp_split_ids=[str(x) for x in op_partition.job_id_to_mp_split_forward_op_ids[job.job_id]],
                mp_splits=[int(x) for x in op_partition.job_id_to_mp_splits[job.job_id]])
        out = orig_get(self, op_partition, cluster, verbose=verbose)
        for job_id, info in pre.items():
            placed = out.action.get(job_id) if hasattr(out.action, 'get') else None
            mapping = None
            if placed:
                w2n = cluster.topology.graph.graph['worker_to_node']
                mapping = {str(op): [int(x) for x in w2n[w].split('-')] for op, w in placed.items()}
            records.append(dict(shape=list(shape), servers=[[int(x) for x in n.split('-')] for n in cluster.topology.graph.nodes()],
                                ramp=[[list(k), float(v['mem']), sorted(int(j) for j in v['job_idxs'])] for k, v in before.items()],
                                placement=mapping, **info))
        return out
    P.RampFirstFitOpPlacer.get = get
    try:
        for name in CASES:
            spec = G.CASES[name]
            np.random.seed(spec['seed']); random.seed(spec['seed'])
            d = tempfile.mkdtemp(prefix='placer_')
            for g in spec['graphs']:
                g.write(d)
            env = G.make_env(d, spec['shape'], spec['n_jobs'], spec['max_partitions'], spec['interarrival'],
                             G.Uniform(spec['frac'][0], spec['frac'][1], decimals=2), max_sim_time=spec.get('max_sim_time', 1e6))
            np.random.seed(spec['seed']); random.seed(spec['seed'])
            obs = env.reset()
            actor = {'random': G.Random(), 'sipml': G.SiPML(spec['max_partitions']), 'acceptable_jct': G.AcceptableJCT()}[spec['actor']]
            done, n0 = False, len(records)
            while not done:
                job = list(env.cluster.job_queue.jobs.values())[0]
                obs, _, done, _ = env.step(int(actor.compute_action(obs, job_to_place=job)))
            print(name, len(records) - n0, 'placements recorded,', sum(1 for r in records[n0:] if r['placement'] is None), 'failed')
    finally:
        P.RampFirstFitOpPlacer.get = orig_get
    path = os.path.join(ROOT, 'tests', 'fixtures', 'placer_cases.json')
    json.dump(records, open(path, 'w'))
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB,', len(records), 'cases')


if __name__ == '__main__':
    main()
