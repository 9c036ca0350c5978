"""`occlusion_in_layers=False` (campx/engine.py:31,528): layers follow the intent of the reference's
BaseUnoccludedObservationRenderer (campx/rendering.py:227-353) -- layers[ch] is the whole curtain of drape ch, the cell
of sprite ch or the backdrop cells holding ch, occluded or not -- while the board is unchanged.

PARITY UNPINNED: the reference's implementation of that renderer cannot run (numpy calls on torch tensors, a two-field
Observation at rendering.py:348), so there is no recording to compare with.  The kernels are compared with the
restatement of its intent in oracle/campx_oracle.py (`UnoccludedRenderer`), and the board / reward / discount of every
step additionally with the ordinary occluded game (they must not change)."""
import numpy as np
import pytest
import torch

from campx_b200 import _native as N
from campx_b200.compiler import CompileError
from examples.worlds import make_world
from oracle import campx_oracle as O

pytestmark = pytest.mark.gpu

WORLDS = ["boat_race", "demo1", "demo2", "demo4"]


def layered_of(obs, chars):
    return np.stack([np.asarray(obs.layers[c]) for c in chars]).astype(np.uint8)


@pytest.mark.parametrize("world", WORLDS)
def test_play_and_fused_rollouts_write_unoccluded_layers(world):
    n, T, limit = 80, 45, 13
    game = make_world(world, num_envs=n, max_episode_steps=limit, track_returns=True, occlusion_in_layers=False)
    twin = make_world(world, num_envs=n, max_episode_steps=limit, track_returns=True)          # occluded
    fused = make_world(world, num_envs=n, max_episode_steps=limit, track_returns=True, occlusion_in_layers=False)
    obs, _, _ = game.its_showtime()
    twin.its_showtime()
    fused.its_showtime()
    chars = game.characters
    first = O.World(world, occlusion_in_layers=False).first[0]
    assert np.array_equal(obs.layered_board[0].cpu().numpy(), layered_of(first, chars))
    assert np.array_equal(obs.board[0].cpu().numpy(), np.asarray(first.board).astype(np.uint8))
    acts = game.native.fill_actions(T, seed=17)
    frames = []
    for t in range(T):                                                  # play(): the single-step composer
        obs, rew, dsc = game.play(acts[t])
        o2, r2, d2 = twin.play(acts[t])
        assert torch.equal(obs.board, o2.board) and torch.equal(rew, r2) and torch.equal(dsc, d2)
        assert torch.equal(game.flags, twin.flags)
        frames.append((obs.layered_board.clone(), obs.board.clone()))
        assert torch.equal(obs.layers["A"], obs.layered_board[:, chars.index("A")])
        assert torch.equal(obs.layered_board_as(torch.float32), obs.layered_board.float())
    boards, layered, rewards, discounts, flags = fused.rollout_observations(acts)               # lane-per-env kernel
    for t in range(T):
        assert torch.equal(layered[t], frames[t][0]) and torch.equal(boards[t], frames[t][1]), (world, t)
    a, lay = acts.cpu().numpy(), layered.cpu().numpy()
    for i in (0, 1, 33, n - 1):
        for t, (o, rew, dsc, term, trunc, eng) in enumerate(
                O.rollout(world, a[:, i], rebuild_on_done=True, max_episode_steps=limit, occlusion_in_layers=False)):
            assert np.array_equal(lay[t, i], layered_of(o, chars)), (world, i, t)
    # where something IS occluded the two kinds of layers differ: boat_race's agent stands on tiles, Demo 1's on '#'
    occluded = twin.native.layers_from_board(boards)
    if world in ("boat_race", "demo4", "demo1"):
        assert not torch.equal(occluded, layered)
    assert int(layered[:, :, chars.index("A")].sum()) == T * n             # the agent's curtain is always one cell


def test_single_env_mode_and_first_frame_after_reset():
    game = make_world("boat_race", occlusion_in_layers=False)
    obs, reward, discount = game.its_showtime()
    w = O.World("boat_race", occlusion_in_layers=False)
    chars = game.characters
    assert obs.layered_board.shape == (7, 5, 5) and reward is None
    assert np.array_equal(obs.layered_board.cpu().numpy(), layered_of(w.first[0], chars))
    for a in (1, 1, 3, 0, 4, 2):
        onehot = torch.zeros(5)
        onehot[a] = 1
        obs, reward, discount = game.play(onehot)
        o, r, d = w.step(a)
        assert float(r) == reward and np.array_equal(obs.layered_board.cpu().numpy(), layered_of(o, chars))
        assert np.array_equal(obs.layers[">"].cpu().numpy(), np.asarray(o.layers[">"]).astype(np.uint8))
    batch = make_world("demo2", num_envs=40, occlusion_in_layers=False)
    batch.its_showtime()
    batch.play(torch.ones(40, dtype=torch.uint8, device="cuda"))
    mask = torch.zeros(40, dtype=torch.bool, device="cuda")
    mask[::2] = True
    obs = batch.reset(mask)
    first = layered_of(O.World("demo2", occlusion_in_layers=False).first[0], batch.characters)
    assert np.array_equal(obs.layered_board[0].cpu().numpy(), first)
    assert not np.array_equal(obs.layered_board[1].cpu().numpy(), first)


def test_games_outside_the_definition_are_refused():
    # layers that are not a function of the board cannot come from cx_layers_from_board
    game = make_world("demo1", num_envs=8, occlusion_in_layers=False)
    obs, _, _ = game.its_showtime()
    with pytest.raises(NotImplementedError):
        game.native.layers_from_board(obs.board)
    # Demo 3 pays for FIRST entry onto '*' because the occluded '*' layer loses the agent's cell (quirk Q4); with
    # unoccluded layers its update() pays on every step spent there -- a different game, outside the primitives
    with pytest.raises(CompileError):
        make_world("demo3", num_envs=8, occlusion_in_layers=False).its_showtime()
    # likewise an agent that can walk underneath a sprite: its wall gate falls back to layers['A'], which the sprite
    # occludes in one game and does not in the other
    from examples.generality_worlds import make_generality_world
    with pytest.raises(CompileError):
        make_generality_world("goal2", num_envs=8, occlusion_in_layers=False).its_showtime()


GENERIC_WORLDS = ["hello", "zswap", "ghost", "scroll"]


@pytest.mark.parametrize("world", GENERIC_WORLDS)
def test_generic_games_read_unoccluded_layers_off_their_entity_state(world):
    """Games on the generic kernels (rolling drape + sprites + backdrop stamps; z-order directives; hiding sprites; a
    scrolling backdrop) write their unoccluded layers inside the step kernel, frame by frame
    -- terminal frames included -- and agree with the oracle's restatement; boards, rewards, discounts and flags are
    those of the occluded game."""
    from examples.generality_worlds import make_generality_world
    mk = make_world if world == "hello" else make_generality_world
    n, T, limit = 48, 40, 11
    game = mk(world, num_envs=n, max_episode_steps=limit, occlusion_in_layers=False)
    twin = mk(world, num_envs=n, max_episode_steps=limit)
    step = mk(world, num_envs=n, max_episode_steps=limit, occlusion_in_layers=False)
    obs, _, _ = game.its_showtime()
    twin.its_showtime()
    sobs, _, _ = step.its_showtime()
    chars = game.characters
    first = O.World(world, occlusion_in_layers=False).first[0]
    assert np.array_equal(obs.layered_board[0].cpu().numpy(), layered_of(first, chars))
    rng = np.random.Generator(np.random.PCG64(5))
    acts = rng.integers(0, 5, size=(T, n)).astype(np.uint8)
    if world == "hello":
        acts[(acts == 4) & (rng.random((T, n)) < 0.8)] = 2          # fewer quits: longer episodes
    dacts = torch.from_numpy(acts).cuda()
    boards, layered, rewards, discounts, flags = game.rollout_observations(dacts)
    b2, r2, d2, f2 = twin.rollout(dacts)
    assert torch.equal(boards, b2) and torch.equal(rewards, r2) and torch.equal(flags, f2)
    for t in range(T):                                              # play(): one launch per step, same layers
        sobs, _, _ = step.play(dacts[t])
        assert torch.equal(sobs.board, boards[t]) and torch.equal(sobs.layered_board, layered[t]), (world, t)
    lay = layered.cpu().numpy()
    differs = False
    for i in (0, 7, n - 1):
        for t, (o, rew, dsc, term, trunc, eng) in enumerate(
                O.rollout(world, acts[:, i], rebuild_on_done=True, max_episode_steps=limit, occlusion_in_layers=False)):
            want = layered_of(o, chars)
            assert np.array_equal(lay[t, i], want), (world, i, t)
            differs = differs or int(want.sum()) != game.rows * game.cols    # some cell lies in two layers
    assert differs or world == "scroll"
