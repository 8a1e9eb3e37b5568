"""CPU: the weight blob of the device GNN policy (ddls_b200/policy.py) -- key order = the checkpoint order of the reference's GNNPolicy
module tree (restated in tests/gnn_reference.py with the reference's parameter names), sizes = what the C ABI expects
(ramp_policy_weight_count is host-only code), and malformed checkpoints are rejected with the offending key."""
import ctypes as C

import numpy as np
import pytest

from ddls_b200 import policy as P


@pytest.mark.parametrize('overrides,n_actions', [({}, 17), ({}, 9), (dict(num_rounds=3, out_features_msg=24, out_features_hidden=40,
                                                                       out_features_node=12, out_features_graph=6, fcnet_hiddens=(128,)), 5)])
def test_blob_layout_matches_the_module_tree_and_the_c_abi(overrides, n_actions):
    from gnn_reference import GNNPolicy
    cfg = dict(P.DEFAULT_CONFIG); cfg.update(overrides)
    ref = GNNPolicy(cfg, n_actions)
    sd = ref.state_dict()
    assert list(sd.keys()) == list(P.weight_keys(cfg))
    shapes = P.weight_shapes(cfg, n_actions)
    assert {k: tuple(v.shape) for k, v in sd.items()} == shapes
    blob = P.pack_weights(sd, cfg, n_actions)
    assert blob.dtype == np.float32 and blob.ndim == 1
    L = P._engine.load_library()
    P._bind(L)
    c = P._Config(cfg['in_features_node'], cfg['in_features_edge'], cfg['in_features_graph'], n_actions, cfg['out_features_msg'],
                  cfg['out_features_hidden'], cfg['out_features_node'], cfg['out_features_graph'], cfg['num_rounds'],
                  tuple(cfg['fcnet_hiddens'])[0], 0, 0, 1, 3)
    assert L.ramp_policy_weight_count(C.byref(c)) == len(blob) == sum(int(np.prod(s)) for s in shapes.values())
    # the blob is the parameters in key order, row-major
    off = 0
    for k in P.weight_keys(cfg):
        n = int(np.prod(shapes[k]))
        np.testing.assert_array_equal(blob[off:off + n], sd[k].detach().numpy().ravel())
        off += n


def test_malformed_checkpoints_are_rejected():
    cfg = dict(P.DEFAULT_CONFIG)
    sd = P.random_state_dict(cfg, 17, seed=1)
    missing = dict(sd); del missing['graph_module.1.bias']
    with pytest.raises(KeyError, match='graph_module.1.bias'):
        P.pack_weights(missing, cfg, 17)
    bad = dict(sd); bad['logit_module._logits._model.0.weight'] = np.zeros((9, 256), dtype=np.float32)
    with pytest.raises(ValueError, match='logit_module._logits'):
        P.pack_weights(bad, cfg, 17)
    L = P._engine.load_library()
    P._bind(L)
    c = P._Config(5, 2, 17, 17, 32, 64, 16, 8, 1, 256, 0, 0, 1, 3)            # num_rounds < 2 (gnn.py:40-41)
    assert L.ramp_policy_weight_count(C.byref(c)) == -1
