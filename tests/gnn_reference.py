"""TEST INFRASTRUCTURE -- a plain PyTorch fp32 restatement of the reference's GNN policy, module for module, with the reference's
parameter names so that its state_dict is what a GNNPolicy checkpoint holds:

  MeanPool   ml_models/models/mean_pool.py:38-150 (node / edge / reduce modules; per destination node the mean of reduce_module over
             [own (node | zeros) state, mailbox]); DGL is not installed here, so update_all is spelled out per node -- a node with no
             incoming edge gets zeros, which is what dgl fills in for nodes its reduce function is never called on
  GNN        ml_models/models/gnn.py:36-92
  GNNPolicy  ml_models/policies/gnn_policy.py:56-296 (graph module, RLlib FullyConnectedNetwork read-out with a separate value
             branch, log-mask on the logits)

The CUDA path (ddls_b200/csrc/ramp_policy.cu) is compared with this in tests/test_gpu_policy.py."""
import torch
from torch import nn

ACT = {'relu': nn.ReLU, 'leaky_relu': nn.LeakyReLU, 'tanh': nn.Tanh}


class MeanPool(nn.Module):
    def __init__(self, in_node, in_edge, msg, out, act):
        super().__init__()
        self.node_module = nn.Sequential(nn.LayerNorm(in_node), nn.Linear(in_node, msg // 2), ACT[act]())
        self.edge_module = nn.Sequential(nn.LayerNorm(in_edge), nn.Linear(in_edge, msg // 2), ACT[act]())
        self.reduce_module = nn.Sequential(nn.LayerNorm(msg), nn.Linear(msg, out), ACT[act]())
        self.msg = msg

    def forward(self, z, ef, src, dst):
        n = z.shape[0]
        m = torch.cat((self.node_module(z[src]), self.edge_module(ef)), -1)              # mp_func, one message per edge
        local = torch.cat((self.node_module(z), torch.zeros(n, self.msg // 2)), -1)      # reduce_func's own state
        out = []
        for v in range(n):
            inbox = m[dst == v]
            if inbox.shape[0] == 0:
                out.append(torch.zeros(self.reduce_module[1].out_features))
                continue
            states = torch.cat((local[v:v + 1], inbox), 0)
            out.append(torch.mean(self.reduce_module(states), dim=0))
        return torch.stack(out)


class GNN(nn.Module):
    def __init__(self, c):
        super().__init__()
        dims = [c['in_features_node']] + [c['out_features_hidden']] * (c['num_rounds'] - 1) + [c['out_features_node']]
        self.layers = nn.ModuleList([MeanPool(dims[r], c['in_features_edge'], c['out_features_msg'], dims[r + 1], c['aggregator_activation'])
                                     for r in range(c['num_rounds'])])

    def forward(self, z, ef, src, dst):
        for layer in self.layers:
            z = layer(z, ef, src, dst)
        return z


class SlimFC(nn.Module):
    def __init__(self, i, o, act=None):
        super().__init__()
        self._model = nn.Sequential(*([nn.Linear(i, o)] + ([ACT[act]()] if act else [])))

    def forward(self, x):
        return self._model(x)


class FullyConnectedNetwork(nn.Module):
    """ray.rllib.models.torch.fcnet.FullyConnectedNetwork with vf_share_layers False: hidden SlimFCs, a logits SlimFC, and the
    same stack again for the value."""

    def __init__(self, i, hiddens, n_out, act):
        super().__init__()
        dims = [i] + list(hiddens)
        self._hidden_layers = nn.Sequential(*[SlimFC(dims[k], dims[k + 1], act) for k in range(len(hiddens))])
        self._logits = SlimFC(dims[-1], n_out)
        self._value_branch_separate = nn.Sequential(*[SlimFC(dims[k], dims[k + 1], act) for k in range(len(hiddens))])
        self._value_branch = SlimFC(dims[-1], 1)

    def forward(self, x):
        return self._logits(self._hidden_layers(x)), self._value_branch(self._value_branch_separate(x)).squeeze(-1)


class GNNPolicy(nn.Module):
    def __init__(self, c, n_actions):
        super().__init__()
        self.c = c
        self.gnn_module = GNN(c)
        gin = c['in_features_graph'] + n_actions
        self.graph_module = nn.Sequential(nn.LayerNorm(gin), nn.Linear(gin, c['out_features_graph']))
        self.logit_module = FullyConnectedNetwork(c['out_features_graph'] + c['out_features_node'], c['fcnet_hiddens'], n_actions,
                                                  c['fcnet_activation'])

    def embed(self, node_features, edge_features, src, dst):
        return torch.mean(self.gnn_module(node_features, edge_features, src, dst), 0)

    def forward(self, emb_nodes, graph_features_and_mask, action_mask):
        """emb_nodes [n, out_node] (the node mean of each observation's job), graph_features_and_mask [n, in_graph + |A|] (the
        observation's 'graph_features'), action_mask [n, |A|]"""
        final = torch.cat((emb_nodes, self.graph_module(graph_features_and_mask)), dim=1)
        logits, value = self.logit_module(final)
        if self.c['apply_action_mask']:
            logits = logits + torch.maximum(torch.log(action_mask), torch.tensor(torch.finfo(torch.float32).min))
        return logits, value
