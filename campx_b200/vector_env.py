"""Vector-environment adapter: gym-style `reset()` / `step()` over a batched `Engine`.

The reference ships no environment wrapper; its RL scripts drive `Engine` directly -- a fresh
`make_game()` per episode, `game.play(action)` per step, `board.layered_board.view(-1).float()` as the
policy input (examples/actor_critic.py:146-173) -- and its README points at the `safe-grid-agents`
agents, which expect the usual `reset() -> obs`, `step(action) -> obs, reward, done, info` loop
(README.md:10, .gitmodules:1-4).  This class is that loop for `num_envs`
environments at once.  It is host-side glue only: every call is one `Engine.play()` (one kernel launch)
plus, for the layered observation kinds, one `cx_layers_from_board[_f32]` launch.

Episode boundaries follow SURVEY H4: the step that ends an episode returns the terminal observation,
its reward and `terminated` / `truncated`; the environment's next `step()` starts from the
post-`its_showtime()` state.  Nothing is copied to the host.
"""
import torch

from . import _native as N

OBSERVATIONS = ("board", "layered", "features")


class VectorEnv(object):
    """num_envs copies of one game.

    engine       a batched `Engine` (from `ascii_art_to_game(..., num_envs=N)`), started or not; it must
                 have been created with auto_reset (the default for batched engines)
    observation  'board'    uint8   [N, rows, cols]            ASCII codes (Observation.board)
                 'layered'  uint8   [N, chars, rows, cols]     Observation.layered_board, canonical channel order
                 'features' float32 [N, chars * rows * cols]   layered_board.view(-1).float(), the reference's
                                                               policy input (actor_critic.py:147,173)
    """

    def __init__(self, engine, observation="board"):
        if observation not in OBSERVATIONS:
            raise ValueError("observation must be one of %r" % (OBSERVATIONS,))
        if not getattr(engine, "_batched", False):
            raise ValueError("VectorEnv needs a batched Engine (create the game with num_envs=...)")
        if not engine._auto_reset:
            raise ValueError("VectorEnv needs an Engine created with auto_reset=True")
        self.engine = engine
        self.observation = observation
        self._started = False

    # ---- shapes -------------------------------------------------------------------------------------
    @property
    def num_envs(self):
        return self.engine.num_envs

    @property
    def num_actions(self):
        return self.engine._num_actions

    @property
    def single_observation_shape(self):
        e = self.engine
        if self.observation == "board":
            return (e.rows, e.cols)
        if not self._started:
            raise RuntimeError("the number of characters is known after reset()")
        chars = len(e.characters)
        return (chars, e.rows, e.cols) if self.observation == "layered" else (chars * e.rows * e.cols,)

    # ---- the loop -----------------------------------------------------------------------------------
    def reset(self, mask=None):
        """Start (first call: compile and upload the game) or restart environments; returns observations.

        mask: optional bool/uint8 [N]; nonzero entries are restarted, the others keep their state."""
        if not self._started:
            if self.engine._showtime:
                obs = self.engine.reset(None)
            else:
                obs, _, _ = self.engine.its_showtime()
            self._started = True
            if mask is not None:
                raise ValueError("the first reset() starts every environment; mask must be None")
        else:
            obs = self.engine.reset(mask)
        return self._encode(obs)

    def step(self, actions):
        """actions: uint8/int64 [N] indices or one-hot float [N, num_actions] on the engine's device.

        Returns (obs, reward float32 [N], terminated bool [N], truncated bool [N], info) where info holds
        'discount' (float32 [N]), 'reward_is_none' (bool [N]: the reference would have returned None) and
        'flags' (uint8 [N], CX_FLAG_* bits).  All tensors are valid until the next step()."""
        if not self._started:
            raise RuntimeError("call reset() before step()")
        e = self.engine
        obs, reward, discount = e.play(actions)
        flags = e.flags
        if reward is None:                                   # a game that never pays: zeros, like gym expects
            reward = torch.zeros(e.num_envs, dtype=torch.float32, device=flags.device)
        terminated = (flags & N.CX_FLAG_TERMINATED) != 0
        truncated = (flags & N.CX_FLAG_TRUNCATED) != 0
        info = {"discount": discount, "reward_is_none": (flags & N.CX_FLAG_REWARD_NONE) != 0, "flags": flags}
        return self._encode(obs), reward, terminated, truncated, info

    def episode_stats(self):
        """Episode-return statistics since the last full reset (needs track_returns=True)."""
        return self.engine.episode_stats()

    def _encode(self, obs):
        if self.observation == "board":
            return obs.board
        if self.observation == "layered":
            return obs.layered_board
        lb = obs.layered_board_as(torch.float32)
        return lb.view(lb.shape[0], -1)
