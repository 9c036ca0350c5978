"""Host-side API mirror (no GPU): set-up validation and exception classes of campx/engine.py and
campx/ascii_art.py, Palette, Plot, Partial."""
import numpy as np
import pytest
import torch

from campx_b200 import things
from campx_b200.ascii_art import ascii_art_to_game, ascii_art_to_long_tensor, Partial
from campx_b200.engine import Engine, Palette
from campx_b200.plot import Plot


class Noop(things.Drape):
    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        pass


class Still(things.Sprite):
    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        pass


def test_ascii_art_to_long_tensor():
    t = ascii_art_to_long_tensor(["ab", "cd"])
    assert t.dtype == torch.int64 and t.tolist() == [[97, 98], [99, 100]]
    with pytest.raises(ValueError):
        ascii_art_to_long_tensor(["ab", "c"])
    with pytest.raises(ValueError):
        ascii_art_to_long_tensor(["aé"])
    with pytest.raises(TypeError):
        ascii_art_to_long_tensor([["a", "b"]])


def test_ascii_art_to_game_builds_masks_positions_backdrop():
    g = ascii_art_to_game(["#P.", "..#"], ".", sprites={"P": Still}, drapes={"#": Noop}, z_order="#P")
    assert list(g.things.keys()) == ["#", "P"]
    assert g.things["#"].curtain.tolist() == [[1, 0, 0], [0, 0, 1]]
    assert tuple(g.things["P"].position) == (0, 1)
    assert g.things["P"].corner == (2, 3)
    assert g.backdrop.curtain.tolist() == [[46, 46, 46], [46, 46, 46]]
    assert "." in g.backdrop.palette and "#" not in g.backdrop.palette
    assert (g.rows, g.cols) == (2, 3)


def test_ascii_art_to_game_validation():
    with pytest.raises(ValueError):          # schedule must list everything exactly once (ascii_art.py:196)
        ascii_art_to_game(["AB"], ".", drapes={"A": Noop, "B": Noop}, update_schedule="A")
    with pytest.raises(ValueError):          # z_order likewise (:205)
        ascii_art_to_game(["AB"], ".", drapes={"A": Noop, "B": Noop}, z_order="AA")
    with pytest.raises(ValueError):          # what_lies_beneath must not be an entity char (:229)
        ascii_art_to_game(["AB"], "A", drapes={"A": Noop, "B": Noop})
    with pytest.raises(ValueError):          # multi-char beneath string (:213)
        ascii_art_to_game(["AB"], "..", drapes={"A": Noop})
    with pytest.raises(ValueError):          # sprite twice in art (:282)
        ascii_art_to_game(["PP"], ".", sprites={"P": Still})
    with pytest.raises(TypeError):           # mixed flat / nested schedule (:190)
        ascii_art_to_game(["AB"], ".", drapes={"A": Noop, "B": Noop}, update_schedule=[["A"], 3])
    with pytest.raises(TypeError):
        Partial(int)
    g = ascii_art_to_game(["..."], ".", sprites={"P": Still})      # absent sprite -> (0, 0) (:287)
    assert tuple(g.things["P"].position) == (0, 0)
    g = ascii_art_to_game(["A.", ".."], ["xy", "zw"], drapes={"A": Noop})   # art-shaped what_lies_beneath
    assert g.backdrop.curtain.tolist() == [[ord("x"), 46], [46, 46]]


def test_engine_setup_errors():
    e = Engine(3, 3)
    e.update_group("g")
    e.add_prefilled_drape("A", np.zeros((3, 3)), Noop)
    with pytest.raises(RuntimeError):        # duplicate character (engine.py:343-350)
        e.add_prefilled_drape("A", np.zeros((3, 3)), Noop)
    with pytest.raises(ValueError):          # not a single character (:332-336)
        e.add_sprite("PP", (0, 0), Still)
    with pytest.raises(TypeError):           # wrong base class (:47-49)
        e.add_sprite("P", (0, 0), Noop)
    with pytest.raises(ValueError):          # off board (:50-53)
        e.add_sprite("P", (3, 0), Still)
    with pytest.raises(ValueError):          # bad z-order (:421-426)
        e.set_z_order("AB")
    with pytest.raises(TypeError):           # backdrop class (:470-472)
        e.set_prefilled_backdrop(".", np.full((3, 3), 46), Noop)
    with pytest.raises(RuntimeError):        # character claimed already (:465)
        e.set_prefilled_backdrop("A.", np.full((3, 3), 46), things.Backdrop)
    e.set_prefilled_backdrop(".", np.full((3, 3), 46), things.Backdrop)
    with pytest.raises(RuntimeError):        # second backdrop (:467-469)
        e.set_prefilled_backdrop(".", np.full((3, 3), 46), things.Backdrop)
    with pytest.raises(RuntimeError):        # play before showtime (:146-148)
        e.play(0)
    assert Engine(3, 3, occlusion_in_layers=False)._occlusion_in_layers is False   # engine.py:31: accepted, like the
    # reference; the unoccluded layers themselves are GPU work (tests/test_gpu_unoccluded.py)
    with pytest.raises(RuntimeError):        # no backdrop
        Engine(2, 2).compile()


def test_palette():
    p = Palette("#. a")
    assert p["#"] == 35 and p.hash == 35 and p.a == 97 and p.space == 32 and p.period == 46
    assert "#" in p and "x" not in p and sorted(p) == [" ", "#", ".", "a"]
    with pytest.raises(AttributeError):
        p.b
    with pytest.raises(IndexError):
        p["b"]
    with pytest.raises(ValueError):
        Palette(["ab"])


def test_plot_directives():
    p = Plot()
    d = p._get_engine_directives()
    assert (d.summed_reward, d.game_over, d.discount, d.z_updates) == (None, False, 1.0, [])
    p.add_reward(1)
    p.add_reward(2.5)
    assert p._get_engine_directives().summed_reward == 3.5           # plot.py:208-211
    p.change_default_discount(0.5)
    assert p.default_discount == 0.5
    p.terminate_episode()
    assert p._get_engine_directives().game_over and p.default_discount == 0.0
    with pytest.raises(ValueError):
        p.terminate_episode(1.5)
    with pytest.raises(ValueError):
        p.change_default_discount(-0.1)
    p.change_z_order("A", None)
    with pytest.raises(ValueError):
        p.change_z_order(3, None)
    p._clear_engine_directives()
    assert p._get_engine_directives().summed_reward is None and p.default_discount == 1.0   # quirk Q3
    assert p.frame == -1
    p.frame = 0
    with pytest.raises(AssertionError):
        p.frame = 5
    p["k"] = 1
    p.log("hello")
    assert p["log_messages"] == ["hello"] and p["k"] == 1


def test_shadow_matches_golden_boat_race(golden_dir):
    """The compile-time shadow executes user update() code with reference semantics: replay the golden
    preset lap through it (board, reward, discount)."""
    import json
    import os
    from campx_b200.compiler.fingerprint import encode_action
    from examples.worlds import make_world
    for world in ("boat_race", "demo3", "hello"):
        with open(os.path.join(golden_dir, world + ".json")) as f:
            fx = json.load(f)
        for ep in fx["episodes"][:3]:
            g = make_world(world)
            spec = g.compile()
            sh = g._shadow.clone()
            for t, a in enumerate(ep["actions"]):
                r, d = sh.play(encode_action(spec.action_format, a, 5))
                want = ep["frames"][t + 1]
                assert "".join(chr(int(v)) for v in sh.board.reshape(-1).tolist()) == want["board"]
                assert (r is None) == (want["reward"] is None)
                if r is not None:
                    assert float(r) == want["reward"]
                assert float(d) == want["discount"]


def test_shadow_clone_keeps_canvas_backdrop_aliasing(golden_dir):
    """Quirk Q1(ii) on the compile-time shadow and on its clones: in a sprite-only world the canvas
    aliases the backdrop storage (rendering.py:128) and clear() zeroes it (rendering.py:111)."""
    import json
    import os
    with open(os.path.join(golden_dir, "engine_semantics.json")) as f:
        want = json.load(f)["cases"]["sprite_only_world_boards"]

    class Still(things.Sprite):
        def update(self, actions, board, layers, backdrop, all_things, the_plot):
            if actions is None:
                return
            self._position = self.Position(self._position.row, (self._position.col + 1) % self.corner.col)

    g = ascii_art_to_game(["P..", "..."], ".", sprites={"P": Still}, action_format="index")
    spec = g.compile()
    assert spec.backdrop.reshape(-1).tolist() == want[0]          # zeros + the sprite, as after its_showtime
    sh = g._shadow.clone().clone()
    boards = []
    for _ in range(2):
        sh.play(0)
        boards.append(sh.board.reshape(-1).tolist())
    assert boards == want[1:]


def test_vector_env_argument_validation():
    """campx_b200.vector_env: host-side checks that need no GPU."""
    from campx_b200.vector_env import VectorEnv
    from examples.worlds import make_world
    with pytest.raises(ValueError):
        VectorEnv(make_world("boat_race", num_envs=8), observation="pixels")
    with pytest.raises(ValueError):
        VectorEnv(make_world("boat_race"))                                   # one unbatched env
    with pytest.raises(ValueError):
        VectorEnv(make_world("boat_race", num_envs=8, auto_reset=False))
    env = VectorEnv(make_world("boat_race", num_envs=8), observation="features")
    assert env.num_envs == 8 and env.num_actions == 5
    with pytest.raises(RuntimeError):
        env.step(None)                                                       # reset() first
    with pytest.raises(RuntimeError):
        env.single_observation_shape                                         # characters known after reset()
    assert VectorEnv(make_world("hello", num_envs=2)).single_observation_shape == (13, 36)
