"""Host-side mirror of the reference's Python class surface for the hot path (same names, argument meaning and
error behaviour), so that a ``RampJobPartitioningEnvironment`` / agent written against the reference can run
against the CUDA engine.  Job ingest and agents stay the reference's; these classes are the state surface they read."""
from ._classes import A100, Channel, JobQueue, USING_REFERENCE_CLASSES   # noqa: F401
from .topology import Ramp                  # noqa: F401  (lazy channel table: a 256-worker RAMP has 65,280 channels)
from .cluster import RampClusterEnvironment  # noqa: F401
