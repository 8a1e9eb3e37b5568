"""RAMP topology mirror (ddls/topologies/ramp.py:11-67): nodes 'c-r-s', complete graph, two directed channels per
link per channel number, per-transceiver bandwidth = total / num communication groups.  Channel objects are created
lazily (a 256-worker RAMP has 65,280 of them), the id scheme and attributes are the reference's."""
import networkx as nx

from ._classes import Channel
from .devices import gen_channel_id


class _LazyChannels(dict):
    def __init__(self, topo):
        super().__init__()
        self._topo = topo

    def __missing__(self, channel_id):
        # 'src_{a}_dst_{b}_channel_{n}'
        try:
            rest = channel_id[len('src_'):]
            src, rest = rest.split('_dst_')
            dst, num = rest.split('_channel_')
            num = int(num)
        except Exception:
            raise KeyError(channel_id)
        if src not in self._topo.graph or dst not in self._topo.graph or src == dst or not (0 <= num < self._topo.num_channels):
            raise KeyError(channel_id)
        ch = Channel(src, dst, num, channel_bandwidth=self._topo.channel_bandwidth)
        self[channel_id] = ch
        return ch


class _AnyDirectionChannels:
    """``graph[u][v]['channels'][channel_id]`` of the reference (ramp.py:50-60): the two directed channels per channel number
    of a link, resolved lazily through the topology's channel table."""

    def __init__(self, topo):
        self._topo = topo

    def __getitem__(self, channel_id):
        return self._topo.channel_id_to_channel[channel_id]

    def __contains__(self, channel_id):
        try:
            self._topo.channel_id_to_channel[channel_id]
            return True
        except KeyError:
            return False


class _OneHopPaths(dict):
    """``graph.nodes[src]['target_to_shortest_paths'][dst]`` (ramp.py:62-67): RAMP is a complete graph, so the only shortest
    path is the direct hop."""

    def __init__(self, src):
        super().__init__()
        self._src = src

    def __missing__(self, dst):
        if dst == self._src:
            raise KeyError(dst)
        path = [[self._src, dst]]
        self[dst] = path
        return path


class Ramp:
    def __init__(self, num_communication_groups=4, num_racks_per_communication_group=2, num_servers_per_rack=4,
                 num_channels=1, total_node_bandwidth=int(1.6e12), intra_gpu_propagation_latency=1.25e-6,
                 worker_io_latency=100e-9):
        self.num_communication_groups = num_communication_groups
        self.num_racks_per_communication_group = num_racks_per_communication_group
        if num_racks_per_communication_group > num_communication_groups:
            raise Exception(f'num_racks_per_communication_group ({num_racks_per_communication_group}) must be <= num_communication_groups ({num_communication_groups})')
        self.num_servers_per_rack = num_servers_per_rack
        self.num_channels = num_channels
        self.total_node_bandwidth = total_node_bandwidth
        self.channel_bandwidth = total_node_bandwidth / num_communication_groups
        self.intra_gpu_propagation_latency = intra_gpu_propagation_latency
        self.worker_io_latency = worker_io_latency
        self.graph = nx.Graph()
        for c in range(num_communication_groups):
            for r in range(num_racks_per_communication_group):
                for s in range(num_servers_per_rack):
                    self.graph.add_node(f'{c}-{r}-{s}', workers=dict())
        self.channel_id_to_channel = _LazyChannels(self)
        # what the reference's agents read besides the channel table (first_fit_dep_placer.py:113, :145): the complete graph's
        # links with their channels, and the one-hop shortest paths
        channels = _AnyDirectionChannels(self)
        nodes = list(self.graph.nodes)
        self.graph.add_edges_from((u, v, {'channels': channels}) for i, u in enumerate(nodes) for v in nodes[i + 1:])
        for n in nodes:
            self.graph.nodes[n]['target_to_shortest_paths'] = _OneHopPaths(n)

    def channel_id(self, src, dst, num=0):
        return gen_channel_id(src, dst, num)
