"""The tiny state holders the agents read (worker, channel, job queue): inside the reference's process -- where the drop-in
cluster environment is meant to live (RJPE:199-206) -- the reference's own classes are used as they are; the attribute-
compatible mirrors in devices.py / job_queue.py serve stand-alone use (the GPU box has no reference)."""
try:
    from ddls.devices.processors.gpus.A100 import A100          # noqa: F401
    from ddls.devices.channels.channel import Channel           # noqa: F401
    from ddls.environments.cluster.job_queue import JobQueue    # noqa: F401
    USING_REFERENCE_CLASSES = True
except Exception:
    from .devices import A100, Channel          # noqa: F401
    from .job_queue import JobQueue             # noqa: F401
    USING_REFERENCE_CLASSES = False
