"""GPU tests of the pieces around the step kernel that config 5 (on-device actor-critic rollout) uses:
float32 policy input, discounted returns on the device, the batched actor-critic example itself."""
import numpy as np
import pytest
import torch

from examples.worlds import make_world
from oracle import campx_oracle as O

pytestmark = pytest.mark.gpu


def test_policy_input_f32_matches_reference_encoding():
    """actor_critic.py:147: state = board.layered_board.view(-1).float() -- 175 features for boat_race."""
    n = 96
    game = make_world("boat_race", num_envs=n)
    obs, _, _ = game.its_showtime()
    acts = game.native.fill_actions(12, seed=9)
    worlds = [O.World("boat_race") for _ in range(n)]
    a = acts.cpu().numpy()
    for t in range(12):
        obs, _, _ = game.play(acts[t])
        state = obs.layered_board_as(torch.float32).view(n, -1)
        assert state.shape == (n, 175) and state.dtype == torch.float32
        assert torch.equal(state, obs.layered_board.float().view(n, -1))
        s = state.cpu().numpy()
        for i in range(n):
            o, _, _ = worlds[i].step(int(a[t, i]))
            if i % 7 == 0:
                assert np.array_equal(s[i], np.asarray(o.layered_board, dtype=np.float32).reshape(-1))


def test_discounted_returns_kernel():
    """finish_episode (actor_critic.py:115-122) on the device; tolerance: exact in float32 when the host
    reference performs the same fused multiply-add order (it does: g = fma(gamma*c, g, r))."""
    n, T, gamma = 300, 57, 0.99
    game = make_world("hello", num_envs=n, max_episode_steps=20)
    game.its_showtime()
    acts = game.native.fill_actions(T, seed=4)
    boards, rewards, discounts, flags = game.rollout(acts)
    boot = torch.rand(n, device="cuda")
    got = game.native.discounted_returns(rewards, flags, gamma, discount=discounts, bootstrap=boot).cpu().numpy()
    r, d, f, b = rewards.cpu().numpy(), discounts.cpu().numpy(), flags.cpu().numpy(), boot.cpu().numpy()
    want = np.zeros((T, n), dtype=np.float32)
    g = b.astype(np.float64)
    for t in range(T - 1, -1, -1):
        ended = (f[t] & 3) != 0
        c = np.where(ended, 0.0, d[t]).astype(np.float32)
        gc = (np.float32(gamma) * c).astype(np.float32)
        g = (gc.astype(np.float64) * g + r[t].astype(np.float64)).astype(np.float32).astype(np.float64)   # fma
        want[t] = g
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert (f & 3).any()
    # without discounts / bootstrap: plain R = r + gamma R, cut at episode ends
    got2 = game.native.discounted_returns(rewards, flags, gamma).cpu().numpy()
    t_last = T - 1
    assert np.allclose(got2[t_last], r[t_last], rtol=0, atol=0)


def test_batched_actor_critic_example_runs_on_device():
    from examples.actor_critic_batched import run
    lines = []
    history, game = run(num_envs=512, steps=20, iterations=2, log=lines.append)
    assert len(history) == 2 and all(np.isfinite(h[0]) for h in history)
    assert -1.0 <= history[0][1] <= 2.0
    st = game.episode_stats()
    assert st["episodes"] == 2 * 512 and st["env_steps"] == 2 * 20 * 512


def test_vector_env_adapter_matches_oracle():
    """gym-style reset()/step() over the batched engine: observations in the reference's
    policy-input encoding, termination/truncation flags and masked resets, env for env against the oracle."""
    import numpy as np
    from campx_b200.vector_env import VectorEnv
    from oracle import campx_oracle as O
    n, T, limit = 64, 30, 12
    env = VectorEnv(make_world("boat_race", num_envs=n, max_episode_steps=limit), observation="features")
    obs = env.reset()
    assert obs.shape == (n, 175) and obs.dtype == torch.float32 and env.single_observation_shape == (175,)
    first = O.World("boat_race").first[0]
    # canonical channel order: sorted by code point; permute the oracle's layers accordingly
    chars = env.engine.characters
    enc = lambda o: np.stack([np.asarray(o.layers[c]) for c in chars]).reshape(-1).astype(np.float32)
    assert np.array_equal(obs[0].cpu().numpy(), enc(first))
    rng = np.random.Generator(np.random.PCG64(5))
    acts = rng.integers(0, 5, size=(T, n)).astype(np.uint8)
    oracles = [O.rollout("boat_race", acts[:, i], rebuild_on_done=True, max_episode_steps=limit) for i in range(n)]
    for t in range(T):
        a = torch.from_numpy(acts[t]).cuda()
        if t % 2:
            a = torch.nn.functional.one_hot(a.long(), 5).float()
        obs, reward, terminated, truncated, info = env.step(a)
        o_np, r_np = obs.cpu().numpy(), reward.cpu().numpy()
        tr_np, te_np = truncated.cpu().numpy(), terminated.cpu().numpy()
        for i in range(n):
            o, rew, dsc, term, trunc, eng = next(oracles[i])
            assert np.array_equal(o_np[i], enc(o)) and float(rew) == float(r_np[i])
            assert bool(tr_np[i]) == trunc and bool(te_np[i]) == term
        assert bool((info["discount"] == 1.0).all()) and not bool(info["reward_is_none"].any())
    # masked reset: only the selected envs go back to the first frame
    before = env._encode(env.engine._observation(env.engine._out_board)).clone()
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[::3] = True
    obs = env.reset(mask)
    assert np.array_equal(obs[0].cpu().numpy(), enc(first))
    assert torch.equal(obs[~mask], before[~mask])
    # board and layered kinds
    env2 = VectorEnv(make_world("hello", num_envs=8, max_episode_steps=5), observation="layered")
    assert env2.reset().shape == (8, 7, 13, 36)
    o2, r2, te2, tr2, info2 = env2.step(torch.full((8,), 4, dtype=torch.uint8, device="cuda"))   # quit
    assert bool(te2.all()) and bool(info2["reward_is_none"].all()) and bool((info2["discount"] == 0).all())


@pytest.mark.parametrize("world,n,kw", [
    ("boat_race", 4096, dict(max_episode_steps=10, track_returns=True)),     # fused kernel, time limit + stats
    ("boat_race", 64, dict()),                                               # fused kernel, no tracking
    ("demo1", 96, dict(max_episode_steps=7)),                                # agent walks over backdrop '#'
    ("demo3", 32, dict(max_episode_steps=9, auto_reset=False)),              # frozen envs after the limit
    ("demo4", 50, dict(max_episode_steps=10)),                               # n % 32 != 0: two-kernel route
    ("hello", 16, dict(max_episode_steps=6)),                                # generic path: two-kernel route
])
def test_rollout_observations_equals_rollout_plus_layers(world, n, kw):
    """cx_rollout_observations: boards, layered boards, rewards, flags and the state afterwards are identical to
    cx_rollout + cx_layers_from_board, and the layered boards equal the oracle's Observation.layered_board."""
    T = 23
    a = make_world(world, num_envs=n, **kw)
    b = make_world(world, num_envs=n, **kw)
    a.its_showtime()
    b.its_showtime()
    acts = a.native.fill_actions(T, seed=77)
    for chunk in (acts[:9], acts[9:]):                                       # two launches: state carries over
        boards, layered, rewards, discounts, flags = a.rollout_observations(chunk.contiguous())
        boards2, rewards2, discounts2, flags2 = b.rollout(chunk.contiguous())
        assert torch.equal(boards, boards2) and torch.equal(rewards, rewards2) and torch.equal(flags, flags2)
        assert (discounts is None and discounts2 is None) or torch.equal(discounts, discounts2)
        assert torch.equal(layered, b.native.layers_from_board(boards2))
    a.native.fold_stats(), b.native.fold_stats()
    assert torch.equal(a.native.state, b.native.state)
    # against the oracle (last chunk), channel by channel in canonical order
    chars = a.characters
    an, ln = acts.cpu().numpy(), layered.cpu().numpy()
    limit = kw.get("max_episode_steps", 0)
    if kw.get("auto_reset", True):
        for i in (0, n // 2, n - 1):
            # (oracle observations alias live renderer buffers, like the reference's: compare step by step)
            for t, frame in enumerate(O.rollout(world, an[:, i], rebuild_on_done=True, max_episode_steps=limit)):
                if t >= 9:
                    want = np.stack([np.asarray(frame[0].layers[c]) for c in chars]).astype(np.uint8)
                    assert np.array_equal(ln[t - 9, i], want), (world, i, t)


def test_cuda_graph_rollout_is_bit_exact_and_learns_on_device():
    """The T-step (encode -> policy -> sample -> play) loop captured as a CUDA graph: replaying the recorded
    actions through a fresh engine's fused rollout reproduces the graph's rewards, flags and states."""
    from examples.actor_critic_batched import run_graphed
    n, T = 512, 20
    history, game, roll = run_graphed(num_envs=n, steps=T, iterations=3, log=lambda *_: None)
    assert len(history) == 3 and all(np.isfinite(h[0]) for h in history)
    assert roll.graph is not None
    # iteration k continued from the state iteration k-1 left behind; replay all three from a fresh start
    fresh = make_world("boat_race", num_envs=n, max_episode_steps=T, track_returns=True)
    fresh.its_showtime()
    # the graph's last iteration: recorded actions -> same rewards/flags when re-played after the same prefix
    # (the prefix is unknown here, but episodes restart every T steps, so every iteration starts from its_showtime)
    boards, layered, rewards, _, flags = fresh.rollout_observations(roll.actions.clone())
    assert torch.equal(rewards, roll.rewards) and torch.equal(flags, roll.flags)
    # the states the policy saw at step t are the layered boards after step t-1 (first: the its_showtime frame)
    assert torch.equal(roll.states[1:], layered[:-1].view(T - 1, n, -1).float())
    first = fresh.reset().layered_board.view(n, -1).float()
    assert torch.equal(roll.states[0], first)
    assert game.episode_stats()["env_steps"] >= 3 * n * T


@pytest.mark.parametrize("n,n_hidden", [(4096, 32), (333, 17)])
def test_policy_sample_is_the_torch_policy_plus_sample_actions(n, n_hidden):
    """cx_policy_sample = relu(x W1^T + b1) W2^T + b2 -> softmax -> Categorical.sample() in one launch
    (examples/actor_critic.py:64-98): its logits equal the torch module's to rounding, and its actions / log-probs are
    exactly what cx_sample_actions draws from those logits (same Philox stream); a CUDA graph with a device-side step
    counter draws fresh numbers per replay."""
    from examples.actor_critic_batched import Policy
    torch.manual_seed(3)
    game = make_world("boat_race", num_envs=n, max_episode_steps=10)
    obs, _, _ = game.its_showtime()
    nat = game.native
    game.play(nat.fill_actions(1, seed=5)[0])
    x = game.play(nat.fill_actions(1, seed=6)[0])[0].layered_board_as(torch.float32).reshape(n, -1).contiguous()
    x = x + 0.25 * torch.randn_like(x)                       # not only 0/1 inputs
    pol = Policy(x.shape[1], n_hidden=n_hidden).cuda()
    logits = torch.empty((n, nat.n_actions), dtype=torch.float32, device="cuda")
    logp = torch.empty(n, dtype=torch.float32, device="cuda")
    step = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        acts = nat.policy_sample(x, pol.affine1.weight.t().contiguous(), pol.affine1.bias, pol.action_head.weight, pol.action_head.bias,
                                 seed=11, step=step, step_offset=3, logp=logp, logits_out=logits)
        want = pol.action_logits(x)
    assert torch.allclose(logits, want, rtol=1e-4, atol=1e-5), float((logits - want).abs().max())
    logp2 = torch.empty_like(logp)
    acts2 = nat.sample_actions(logits, 11, step=step, step_offset=3, logits=True, logp=logp2)
    assert torch.equal(acts, acts2) and torch.equal(logp, logp2)
    assert int(acts.max()) < nat.n_actions and len(torch.unique(acts)) > 1
    step += 1
    with torch.no_grad():
        acts3 = nat.policy_sample(x, pol.affine1.weight.t().contiguous(), pol.affine1.bias, pol.action_head.weight, pol.action_head.bias,
                                  seed=11, step=step, step_offset=3)
    assert not torch.equal(acts, acts3)
    with pytest.raises(Exception):
        big = Policy(x.shape[1], n_hidden=64).cuda()
        nat.policy_sample(x, big.affine1.weight.t().contiguous(), big.affine1.bias, big.action_head.weight, big.action_head.bias, seed=1)


@pytest.mark.parametrize("world,n,limit", [("boat_race", 4096, 100), ("boat_race", 96, 7), ("demo3", 64, 9)])
def test_persistent_policy_rollout_equals_the_two_kernel_loop(world, n, limit):
    """cx_rollout_policy (policy -> sample -> play, T steps in one launch, the policy input resident in shared memory)
    is bit-identical to T x (cx_policy_sample + cx_step_observations(float32)): states, actions, rewards, flags,
    log-probs, the env state and the statistics afterwards; and the env it drives obeys the oracle."""
    from examples.actor_critic_batched import Policy
    from campx_b200 import _native as N
    torch.manual_seed(5)
    T = 37
    a = make_world(world, num_envs=n, max_episode_steps=limit, track_returns=True)
    b = make_world(world, num_envs=n, max_episode_steps=limit, track_returns=True)
    oa, _, _ = a.its_showtime()
    ob, _, _ = b.its_showtime()
    na, nb = a.native, b.native
    feat = na.n_chars * na.cells
    pol = Policy(feat, n_actions=na.n_actions).cuda()
    w1t = pol.affine1.weight.detach().t().contiguous()
    b1, w2, b2 = pol.affine1.bias.detach(), pol.action_head.weight.detach(), pol.action_head.bias.detach()
    step = torch.full((1,), 11, dtype=torch.int64, device="cuda")
    # one launch
    states = torch.empty((T + 1, n, feat), dtype=torch.float32, device="cuda")
    actions = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    rewards = torch.empty((T, n), dtype=torch.float32, device="cuda")
    flags = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    logp = torch.empty((T, n), dtype=torch.float32, device="cuda")
    na.rollout_policy(T, w1t, b1, w2, b2, 77, states, actions, rewards, flags, step=step, logp=logp)
    # the two-kernel loop on a twin
    states2 = torch.empty_like(states)
    actions2, rewards2, flags2, logp2 = torch.empty_like(actions), torch.empty_like(rewards), torch.empty_like(flags), torch.empty_like(logp)
    board = torch.empty((n, nb.rows, nb.cols), dtype=torch.uint8, device="cuda")
    states2[0].copy_(ob.layered_board_as(torch.float32).reshape(n, -1))
    for t in range(T):
        nb.policy_sample(states2[t], w1t, b1, w2, b2, 77, step=step, step_offset=t, out=actions2[t], logp=logp2[t])
        nb.step_observations(actions2[t], board, states2[t + 1].view(n, nb.n_chars, nb.rows, nb.cols), rewards2[t], flags2[t])
    assert torch.equal(actions, actions2) and torch.equal(rewards, rewards2) and torch.equal(flags, flags2)
    assert torch.equal(states, states2) and torch.equal(logp, logp2)
    na.fold_stats(), nb.fold_stats()
    assert torch.equal(na.state, nb.state)
    assert len(torch.unique(actions)) > 1 and int(((flags & N.CX_FLAG_TRUNCATED) != 0).sum()) == n * (T // limit)
    # and the env side against the oracle on the sampled actions
    an, rn, fn = actions.cpu().numpy(), rewards.cpu().numpy(), flags.cpu().numpy()
    sn = states.cpu().numpy().reshape(T + 1, n, na.n_chars, na.rows, na.cols)
    chars = a.characters
    for i in (0, n // 2, n - 1):
        for t, frame in enumerate(O.rollout(world, an[:, i], rebuild_on_done=True, max_episode_steps=limit)):
            obs, rew, dsc, term, trunc, eng = frame
            assert (0.0 if rew is None else float(rew)) == float(rn[t, i]), (i, t)
            assert trunc == bool(fn[t, i] & N.CX_FLAG_TRUNCATED) and term == bool(fn[t, i] & N.CX_FLAG_TERMINATED), (i, t)
            if not (term or trunc):   # (after an episode end states[t + 1] is the first frame of the next episode)
                want = np.stack([np.asarray(obs.layers[c]) for c in chars]).astype(np.float32)
                assert np.array_equal(sn[t + 1, i], want), (i, t)
