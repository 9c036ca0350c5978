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
