"""State holders mirroring ddls/devices/processors/gpus/A100.py:7-56 and ddls/devices/channels/channel.py:7-41.
Agents read ``mounted_job_idx_to_ops`` / ``memory_occupied`` / ``mounted_job_idx_to_deps`` (placers/utils.py:250-255,
first_fit_dep_placer.py:148-150), so the drop-in environment keeps them live."""
import json
from collections import defaultdict


def gen_job_dep_str(job_idx, job_id, dep_id):          # ddls/utils.py:560-561
    return json.dumps(job_idx) + '_' + json.dumps(job_id) + '_' + json.dumps(dep_id)


def gen_channel_id(src, dst, channel_number):          # ddls/utils.py:550-555
    return f'src_{src}_dst_{dst}_channel_{channel_number}'


class A100:
    def __init__(self, processor_id=None):
        self.processor_id = id(self) if processor_id is None else processor_id
        self.device_type = 'A100'
        self.memory_capacity = int(80e9)
        self.reset()

    def __str__(self):
        return f'{self.device_type}_{self.processor_id}'

    def reset(self):
        self.memory_occupied = 0
        self.mounted_job_idx_to_ops = defaultdict(set)
        self.mounted_job_op_to_priority = dict()
        self.mounted_job_idx_to_job_id = {}

    def mount(self, job, op_id):
        g = job.computation_graph
        if op_id not in g.nodes:
            raise Exception(f'Op ID {op_id} not found in job {job}')
        if self.device_type not in g.nodes[op_id]['compute_cost']:
            raise Exception(f'Tried to mount op on device type {self.device_type} but only profile op compute cost for {g.nodes[op_id]["compute_cost"]}')
        if self.memory_occupied + g.nodes[op_id]['memory_cost'] > self.memory_capacity:
            raise Exception(f'Trying to allocate {g.nodes[op_id]["memory_cost"]} of memory for job {job} op {op_id} but have only {self.memory_capacity - self.memory_occupied} available on processor {self.processor_id}.')
        self.mounted_job_idx_to_ops[job.details['job_idx']].add(op_id)
        self.mounted_job_idx_to_job_id[job.details['job_idx']] = job.job_id
        self.memory_occupied += g.nodes[op_id]['memory_cost']

    def unmount(self, job, op_id):
        self.memory_occupied -= job.computation_graph.nodes[op_id]['memory_cost']
        self.mounted_job_idx_to_ops[job.details['job_idx']].remove(op_id)
        self.mounted_job_op_to_priority.pop(gen_job_dep_str(job.details['job_idx'], job.job_id, op_id), None)
        if len(self.mounted_job_idx_to_ops[job.details['job_idx']]) == 0:
            del self.mounted_job_idx_to_ops[job.details['job_idx']]
            del self.mounted_job_idx_to_job_id[job.details['job_idx']]


class Channel:
    def __init__(self, src, dst, channel_number, channel_bandwidth=int(1.25e9)):
        self.src, self.dst = src, dst
        self.channel_number = id(self) if channel_number is None else channel_number
        self.channel_id = gen_channel_id(self.src, self.dst, self.channel_number)
        self.channel_bandwidth = channel_bandwidth
        self.reset()

    def __str__(self):
        return f'Channel_{self.channel_id}'

    def reset(self):
        self.mounted_job_idx_to_deps = defaultdict(set)
        self.mounted_job_dep_to_priority = dict()

    def mount(self, job, dep):
        self.mounted_job_idx_to_deps[job.details['job_idx']].add(dep)

    def unmount(self, job, dep):
        self.mounted_job_idx_to_deps[job.details['job_idx']].remove(dep)
        self.mounted_job_dep_to_priority.pop(gen_job_dep_str(job.details['job_idx'], job.job_id, dep), None)
        if len(self.mounted_job_idx_to_deps[job.details['job_idx']]) == 0:
            del self.mounted_job_idx_to_deps[job.details['job_idx']]
