"""-m gpu: the GNN policy forward on the device (ramp_policy_*) against a plain PyTorch fp32 restatement of the
reference's GNNPolicy (tests/gnn_reference.py), on the observations recorded from the unmodified reference's encoder
(tests/fixtures/obs_cases.npz) and on live device rollouts.  Tolerance: 2e-5 absolute + 2e-5 relative on logits, values and
embeddings (fp32 both sides; only the summation order differs)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL = RTOL = 2e-5
D = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fixtures', 'obs_cases.npz'))


def _graphs():
    from ddls_b200 import synth
    return [synth.chain_graph(6, 'chain6'), synth.chain_graph(5, 'chain5'), synth.chain_graph(4, 'chain4'),
            synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3),
            synth.resnet_like_graph(n_blocks=1, stem=2, name='res1', seed=11, body_per_block=2),
            synth.transformer_like_graph(n_layers=1, name='tfm1', seed=4), synth.transformer_like_graph(n_layers=1, name='tfm1b', seed=6),
            synth.transformer_like_graph(n_layers=2, name='tfm2', seed=9), synth.resnet_like_graph()]


def _torch_policy(config, n_actions, sd):
    import torch
    from gnn_reference import GNNPolicy
    ref = GNNPolicy(config, n_actions)
    ref.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return ref.eval()


CONFIGS = {
    'gnn.yaml': {},
    'leaky_3_rounds_tanh': dict(num_rounds=3, aggregator_activation='leaky_relu', fcnet_activation='tanh', out_features_msg=24,
                                out_features_hidden=40, out_features_node=12, out_features_graph=6, fcnet_hiddens=(128,)),
    'unmasked_wide': dict(apply_action_mask=False, out_features_msg=64, out_features_hidden=128, out_features_node=32, fcnet_hiddens=(512,)),
}


@pytest.mark.parametrize('name,n_actions', [('gnn.yaml', 17), ('gnn.yaml', 9), ('leaky_3_rounds_tanh', 9), ('unmasked_wide', 9)])
def test_policy_forward_matches_torch_on_recorded_observations(name, n_actions):
    """The recorded episodes ran with max_partitions_per_op 8 (72 observations, |A| = 9) and 16 (9 observations, |A| = 17)."""
    import torch
    from ddls_b200 import policy as P
    from ddls_b200.observation import job_arrays
    graphs = _graphs()
    cfg = dict(P.DEFAULT_CONFIG); cfg.update(CONFIGS[name])
    sd = P.random_state_dict(cfg, n_actions, seed=3)
    assert list(sd) == list(P.weight_keys(cfg))
    pol = P.DeviceGNNPolicy(graphs, n_actions, cfg, sd)
    ref = _torch_policy(cfg, n_actions, sd)
    assert list(ref.state_dict().keys()) == list(P.weight_keys(cfg))        # blob order == checkpoint order
    # ---- message passing + node mean per job type ----
    emb = pol.embed()
    want = []
    with torch.no_grad():
        for st in pol.static:
            want.append(ref.embed(torch.from_numpy(st['node_features']), torch.from_numpy(st['edge_features']),
                                  torch.from_numpy(st['edges_src'].astype(np.int64)), torch.from_numpy(st['edges_dst'].astype(np.int64))).numpy())
    want = np.stack(want)
    np.testing.assert_allclose(emb, want, atol=ATOL, rtol=RTOL)
    assert np.abs(want).max() > 1e-3
    # ---- the read-out on every observation recorded from the reference's encoder ----
    arrs = [job_arrays(g) for g in graphs]
    model, gf, mask = [], [], []
    for i in range(int(D['n_cases'])):
        p = f'c{i}_'
        if len(D[p + 'obs_action_mask']) != n_actions:
            continue
        m = [k for k, a in enumerate(arrs) if len(a['op_compute']) == len(D[p + 'op_compute']) and np.array_equal(a['op_compute'], D[p + 'op_compute'])]
        model.append(m[0]); gf.append(D[p + 'obs_graph_features'][:17]); mask.append(D[p + 'obs_action_mask'])
    model, gf, mask = np.array(model), np.stack(gf).astype(np.float32), np.stack(mask).astype(np.uint8)
    logits, value = pol.forward(model, gf, mask)
    with torch.no_grad():
        full = torch.from_numpy(np.concatenate([gf, mask.astype(np.float32)], axis=1))
        wl, wv = ref(torch.from_numpy(want[model]), full, torch.from_numpy(mask.astype(np.float32)))
    wl, wv = wl.numpy(), wv.numpy()
    valid = mask.astype(bool) | (not cfg['apply_action_mask'])
    np.testing.assert_allclose(logits[valid], wl[valid], atol=ATOL, rtol=RTOL)
    if cfg['apply_action_mask']:
        assert (~valid).any()
        np.testing.assert_array_equal(logits[~valid], wl[~valid])            # logit + finfo.min, bit for bit
        assert (logits[~valid] < -3e38).all()
    np.testing.assert_allclose(value, wv, atol=ATOL, rtol=RTOL)
    assert np.ptp(wl[valid]) > 1e-2 and np.ptp(wv) > 1e-3
    pol.close()


def _env(B=256, J=6, seed=5, frac=(0.1, 1.0, 2)):
    from ddls_b200 import synth
    from ddls_b200.batched import DeviceRampJobPartitioningEnvironment
    graphs = [synth.resnet_like_graph(n_blocks=2, stem=2, name='res2', seed=7, body_per_block=3), synth.chain_graph(6, 'chain6'),
              synth.transformer_like_graph(n_layers=1, name='tfm1', seed=4)]
    env = DeviceRampJobPartitioningEnvironment((4, 4, 2), graphs, n_episodes=B, jobs_per_episode=J, max_partitions_per_op=16,
                                               min_op_run_time_quantum=2.0, interarrival=('exponential', 600.0), frac=frac, seed=seed)
    return env, graphs


def _want_logits(pol, ref, obs):
    import torch
    emb = pol.embed()
    m = np.clip(obs['model'], 0, None)
    stat = np.stack([st['graph_static'] for st in pol.static])[m]
    dyn = obs['graph_features_dynamic']
    gf = np.concatenate([dyn[:, :9], stat, dyn[:, 9:]], axis=1).astype(np.float32)
    mask = obs['action_mask'].astype(np.float32)
    with torch.no_grad():
        wl, wv = ref(torch.from_numpy(emb[m]), torch.from_numpy(np.concatenate([gf, mask], axis=1)), torch.from_numpy(mask))
    return wl.numpy(), wv.numpy()


def test_policy_drives_device_rollouts_without_the_host():
    """act() reads the environment's device buffers and writes its action buffer; every decision of a whole batch of rollouts is
    the first maximal logit of the torch restatement evaluated on the observation the host copy of the environment shows."""
    from ddls_b200 import policy as P
    env, graphs = _env()
    cfg = dict(P.DEFAULT_CONFIG)
    sd = P.random_state_dict(cfg, 17, seed=11)
    sd['logit_module._logits._model.0.bias'][[16, 8, 4]] += np.float32([4.5, 3.0, 1.5])   # prefers large blocks: the cluster fills up
    pol = P.DeviceGNNPolicy(graphs, 17, cfg, sd)
    ref = _torch_policy(cfg, 17, sd)
    obs = env.reset()
    n_steps, n_placed, seen = 0, 0, set()
    while not obs['done'].all() and n_steps < 40:
        live = ~obs['done']
        wl, wv = _want_logits(pol, ref, obs)
        pol.act(env)
        got = pol.read(env)
        np.testing.assert_allclose(got['logits'][live][obs['action_mask'][live].astype(bool)], wl[live][obs['action_mask'][live].astype(bool)],
                                   atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(got['value'][live], wv[live], atol=ATOL, rtol=RTOL)
        # greedy: the chosen logit is the maximum (ties to the first index); near-ties within the tolerance may legitimately differ
        chosen = wl[np.arange(env.B), got['actions']]
        assert (chosen[live] >= wl[live].max(axis=1) - 4 * ATOL).all()
        assert (obs['action_mask'][np.arange(env.B), got['actions']][live] == 1).all()
        assert (got['actions'][~live] == 0).all()
        lse = np.log(np.exp(wl[live] - wl[live].max(axis=1, keepdims=True)).sum(axis=1))
        np.testing.assert_allclose(got['logp'][live], chosen[live] - wl[live].max(axis=1) - lse, atol=1e-4)
        seen.update(got['actions'][live].tolist())
        obs, reward, done, info = env.step(None)
        n_placed += int((reward > 0).sum())
        n_steps += 1
    assert obs['done'].all() and n_placed > 0 and len(seen) >= 2
    env.close(); pol.close()


def test_sampled_actions_follow_the_softmax_and_respect_the_mask():
    from ddls_b200 import policy as P
    env, graphs = _env(B=4096, J=4, seed=9, frac=(0.5, 0.5, 2))
    pol = P.DeviceGNNPolicy(graphs, 17, None, P.random_state_dict(P.DEFAULT_CONFIG, 17, seed=2))
    obs = env.reset()
    pol.act(env, sample=True, seed=123)
    a = pol.read(env)
    assert (obs['action_mask'][np.arange(env.B), a['actions']] == 1).all()
    # episodes whose queued job is of the same type see the same observation at reset: their draws are i.i.d. from one softmax
    for m in range(len(graphs)):
        rows = np.flatnonzero(obs['model'] == m)
        assert len(rows) > 500
        p = np.exp(a['logits'][rows[0]].astype(np.float64) - a['logits'][rows[0]].max()); p /= p.sum()
        np.testing.assert_array_equal(a['logits'][rows], np.broadcast_to(a['logits'][rows[0]], (len(rows), 17)))
        freq = np.bincount(a['actions'][rows], minlength=17) / len(rows)
        assert np.abs(freq - p).max() < 5 * np.sqrt(0.25 / len(rows))
        np.testing.assert_allclose(a['logp'][rows], np.log(p[a['actions'][rows]]), atol=1e-4)
    # a different call draws differently, the same (seed, call index) is reproducible across policies
    pol.act(env, sample=True, seed=123)
    b = pol.read(env)
    assert (a['actions'] != b['actions']).any()
    pol2 = P.DeviceGNNPolicy(graphs, 17, None, P.random_state_dict(P.DEFAULT_CONFIG, 17, seed=2))
    pol2.act(env, sample=True, seed=123)
    np.testing.assert_array_equal(pol2.read(env)['actions'], a['actions'])
    env.close(); pol.close(); pol2.close()


def test_policy_rejects_bad_configurations():
    from ddls_b200 import policy as P
    g = _graphs()[:1]
    with pytest.raises(Exception, match='num_rounds'):
        P.DeviceGNNPolicy(g, 17, dict(num_rounds=1))
    with pytest.raises(Exception, match='multiple of 32'):
        P.DeviceGNNPolicy(g, 17, dict(fcnet_hiddens=(100,)))
    pol = P.DeviceGNNPolicy(g, 17)
    with pytest.raises(Exception, match='weights given'):
        pol.set_weights(np.zeros(10, dtype=np.float32))
    with pytest.raises(Exception, match='names node'):
        pol.set_model(0, np.zeros((3, 5)), np.zeros((1, 2)), [0], [7], np.zeros(6))
    pol.close()


def test_collected_trajectories_equal_step_by_step_reads():
    """DeviceGNNPolicy.collect keeps a whole segment on the device (decisions, log-probabilities, values, rewards, done flags recorded
    by device-to-device copies) and reads it back once: the same as acting and reading step by step on an identical environment."""
    from ddls_b200 import policy as P
    H = 6
    env, graphs = _env(B=192, J=H, seed=21)
    env2, _ = _env(B=192, J=H, seed=21)
    sd = P.random_state_dict(P.DEFAULT_CONFIG, 17, seed=4)
    pol, pol2 = P.DeviceGNNPolicy(graphs, 17, None, sd), P.DeviceGNNPolicy(graphs, 17, None, sd)
    traj = pol.collect(env, H, sample=True, seed=50)
    obs = env2.reset()
    for t in range(H):
        pol2.act(env2, sample=True, seed=50 + t)
        got = pol2.read(env2)
        np.testing.assert_array_equal(traj['model'][t], obs['model'])
        np.testing.assert_array_equal(traj['graph_features_dynamic'][t], obs['graph_features_dynamic'])
        np.testing.assert_array_equal(traj['action_mask'][t], obs['action_mask'].astype(np.uint8))
        np.testing.assert_array_equal(traj['action'][t], got['actions'])
        np.testing.assert_array_equal(traj['logp'][t], got['logp'])
        np.testing.assert_array_equal(traj['value'][t], got['value'])
        np.testing.assert_array_equal(traj['live'][t], ~obs['done'])
        obs, reward, done, _ = env2.step(None)
        np.testing.assert_array_equal(traj['reward'][t], reward)
        np.testing.assert_array_equal(traj['done'][t], done)
    assert traj['done'][-1].all() and (traj['reward'] > 0).any() and (traj['reward'] < 0).any()
    np.testing.assert_array_equal(env.decisions(), env2.decisions())
    env.close(); env2.close(); pol.close(); pol2.close()
