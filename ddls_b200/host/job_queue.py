"""ddls/environments/cluster/job_queue.py:8-38."""
from collections import OrderedDict


class JobQueue:
    def __init__(self, queue_capacity):
        self.jobs = OrderedDict()
        self.queue_capacity = queue_capacity

    def __len__(self):
        return len(self.jobs)

    def can_fit(self, jobs):
        if type(jobs) is not list:
            jobs = [jobs]
        return len(self) + len(jobs) <= self.queue_capacity

    def add(self, jobs):
        if type(jobs) is not list:
            jobs = [jobs]
        if not self.can_fit(jobs):
            raise Exception(f'Cannot fit all jobs, only have {self.queue_capacity - len(self)} of space remaining.')
        for job in jobs:
            self.jobs[job.job_id] = job

    def remove(self, jobs):
        if type(jobs) is not list:
            jobs = [jobs]
        for job in jobs:
            del self.jobs[job.job_id]
