"""Synthetic job-graph profiles in the PipeDream text format the reference ingests.

The reference does not ship the job-graph profiles its configs point at
(scripts/ramp_job_partitioning_configs/heuristic_config.yaml:85-87), only the
parser for the format (ddls/utils.py:278-340).  These writers emit files that
parser accepts:

    node{i} -- {Op}({args}) -- forward_compute_time=F, backward_compute_time=B, activation_size=A, parameter_size=P
    \tnode{u} -- node{v}

Node ids start at 1 and must be topologically increasing so that the mirrored
backward pass (ddls/utils.py:342-370, backward id = 2n-(i-1)) joins correctly.

The same generators also return the forward graph as plain python structures so
that ``ddls_b200.host`` can build jobs without going through a file.
"""
import os
import random
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class ForwardGraph:
    """Forward-pass profile: 1-indexed nodes, edges (u, v) with u < v."""
    name: str
    ops: List[str] = field(default_factory=list)          # op type per node
    fwd: List[float] = field(default_factory=list)        # forward compute time
    bwd: List[float] = field(default_factory=list)        # backward compute time
    act: List[float] = field(default_factory=list)        # activation size (bytes)
    par: List[float] = field(default_factory=list)        # parameter size (bytes)
    edges: List[Tuple[int, int]] = field(default_factory=list)

    @property
    def n(self):
        return len(self.ops)

    def add(self, op, fwd, bwd, act, par):
        self.ops.append(op)
        self.fwd.append(float(fwd))
        self.bwd.append(float(bwd))
        self.act.append(float(act))
        self.par.append(float(par))
        return self.n  # 1-indexed id of the node just added

    def write(self, directory):
        os.makedirs(directory, exist_ok=True)
        path = os.path.join(directory, f'{self.name}.txt')
        with open(path, 'w') as f:
            for i in range(self.n):
                f.write(f'node{i + 1} -- {self.ops[i]}() -- forward_compute_time={self.fwd[i]!r}, '
                        f'backward_compute_time={self.bwd[i]!r}, activation_size={self.act[i]!r}, '
                        f'parameter_size={self.par[i]!r}\n')
            for (u, v) in self.edges:
                f.write(f'\tnode{u} -- node{v}\n')
        return path


def chain_graph(n_fwd=6, name='chain'):
    """BASELINE.json config 1: chain DAG, costs 1.5*i / 3.1*i ms, activation 1e5*i bytes."""
    g = ForwardGraph(name=name)
    for i in range(1, n_fwd + 1):
        g.add('Linear', 1.5 * i, 3.1 * i, 1e5 * i, 4096.0 * i)
        if i > 1:
            g.edges.append((i - 1, i))
    return g


def resnet_like_graph(n_blocks=16, stem=3, name='resnet50_like', seed=1, body_per_block=9):
    """ResNet-50-shaped: stem, ``n_blocks`` bottleneck blocks each with a residual
    join (skip edge from block input to the add), head.  ~175 forward nodes by
    default (3 + 16*(9+1+...) ...).  fwd U(0.2,4) ms, bwd U(0.4,8) ms,
    activation in {1..40}*100352 B, params in {1..100}*4096 B."""
    rng = random.Random(seed)
    g = ForwardGraph(name=name)

    def node(op):
        return g.add(op, round(rng.uniform(0.2, 4.0), 6), round(rng.uniform(0.4, 8.0), 6),
                     float(rng.randint(1, 40) * 100352), float(rng.randint(1, 100) * 4096))

    prev = None
    for s in range(stem):
        cur = node(['Conv2d', 'BatchNorm2d', 'ReLU'][s % 3])
        if prev is not None:
            g.edges.append((prev, cur))
        prev = cur
    for b in range(n_blocks):
        block_in = prev
        for k in range(body_per_block):
            cur = node(['Conv2d', 'BatchNorm2d', 'ReLU'][k % 3])
            g.edges.append((prev, cur))
            prev = cur
        add = node('Add')
        g.edges.append((prev, add))
        g.edges.append((block_in, add))     # residual join
        prev = add
    for op in ('AvgPool2d', 'Linear'):
        cur = node(op)
        g.edges.append((prev, cur))
        prev = cur
    g.edges.sort()
    return g


def residual_small_graph(n_blocks=6, name='residual54', seed=3):
    """54-forward-node residual graph (the SURVEY.md measured anchor shape)."""
    return resnet_like_graph(n_blocks=n_blocks, stem=4, name=name, seed=seed, body_per_block=7)


def transformer_like_graph(n_layers=12, name='bert_base_like', seed=2, gpt=False):
    """BERT-base-shaped: embed, n_layers x (QKV, attn, proj, add+LN, FFN1, FFN2, add+LN), head.
    Residual edges around attention and FFN."""
    rng = random.Random(seed)
    g = ForwardGraph(name=name)

    def node(op, scale=1.0):
        return g.add(op, round(rng.uniform(0.3, 3.0) * scale, 6), round(rng.uniform(0.6, 6.0) * scale, 6),
                     float(rng.randint(4, 48) * 98304), float(rng.randint(1, 144) * 16384))

    prev = node('Embedding')
    for _ in range(n_layers):
        x = prev
        q = node('Linear'); k = node('Linear'); v = node('Linear')
        for t in (q, k, v):
            g.edges.append((x, t))
        att = node('Attention', 1.5)
        for t in (q, k, v):
            g.edges.append((t, att))
        proj = node('Linear'); g.edges.append((att, proj))
        ln1 = node('LayerNorm', 0.3); g.edges.append((proj, ln1)); g.edges.append((x, ln1))
        f1 = node('Linear', 2.0); g.edges.append((ln1, f1))
        f2 = node('Linear', 2.0); g.edges.append((f1, f2))
        ln2 = node('LayerNorm', 0.3); g.edges.append((f2, ln2)); g.edges.append((ln1, ln2))
        prev = ln2
    head = node('Linear'); g.edges.append((prev, head))
    if not gpt:
        pool = node('Tanh', 0.2); g.edges.append((head, pool))
    g.edges.sort()
    return g
