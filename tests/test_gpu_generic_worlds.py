"""GPU parity of the generic kernels on worlds that stress what the six reference worlds do not:
several update groups (re-render between groups, engine.py:195-208), two moving one-cell drapes where
one blocks the other, an entity updated BEFORE the agent it watches, a sprite-only world (quirk Q1(ii):
canvas zeroed at every render), an invisible sprite.  Each world is written twice -- as user-level
campx_b200 classes (compiled by the front end) and as numpy oracle entities -- and compared frame by frame."""
import numpy as np
import pytest
import torch

from campx_b200 import things
from campx_b200.ascii_art import ascii_art_to_game, Partial
from examples.worlds import Walker, Arrow, Slider, RING_ART
from oracle import campx_oracle as O

pytestmark = pytest.mark.gpu


def onehot(a):
    v = np.zeros(5, dtype=np.float32)
    v[a] = 1
    return v


def compare(game, oracle_factory, encode, n=40, T=50, seed=3):
    obs, reward, discount = game.its_showtime()
    oracles = [oracle_factory() for _ in range(n)]
    firsts = [o.its_showtime() for o in oracles]
    assert np.array_equal(obs.board[0].cpu().numpy(), np.asarray(firsts[0][0].board).astype(np.uint8))
    rng = np.random.Generator(np.random.PCG64(seed))
    acts = rng.integers(0, 5, size=(T, n)).astype(np.uint8)
    boards, rewards, discounts, flags = game.rollout(torch.from_numpy(acts).cuda())
    b, r, f = boards.cpu().numpy(), rewards.cpu().numpy(), flags.cpu().numpy()
    for i in range(n):
        for t in range(T):
            o, rew, dsc = oracles[i].play(encode(int(acts[t, i])))
            ctx = "env %d t %d" % (i, t)
            assert np.array_equal(b[t, i], np.asarray(o.board).astype(np.uint8)), ctx
            assert (rew is None) == bool(f[t, i] & 4), ctx
            if rew is not None:
                assert float(rew) == float(r[t, i]), ctx
    return game


def _ring(schedule, **kw):
    def arrow(bonus):
        return Partial(Arrow, bonus=torch.FloatTensor(bonus), toll=-0.25)
    return ascii_art_to_game(
        RING_ART, ' ',
        drapes={'A': Partial(Walker, walls='#', strict=True), '#': things.FixedDrape,
                '^': arrow([0, 0, 3, 1, 0]), '>': arrow([1, 3, 0, 0, 0]),
                'v': arrow([0, 0, 1, 3, 0]), '<': arrow([3, 1, 0, 0, 0])},
        z_order='^>v<A#', update_schedule=schedule, **kw)


def _oracle_ring(schedule):
    hover = lambda d: (O.DirectionalHoverRewardDrape, (), dict(dctns=d, base_reward=-0.25))
    return O.ascii_art_to_game(
        O.BOAT_RACE_ART, ' ',
        drapes={'A': (O.AgentDrape, (), dict(variant='boat_race')), '#': O.FixedDrape,
                '^': hover([0, 0, 3, 1, 0]), '>': hover([1, 3, 0, 0, 0]),
                'v': hover([0, 0, 1, 3, 0]), '<': hover([3, 1, 0, 0, 0])},
        z_order='^>v<A#', update_schedule=schedule)


def test_two_update_groups_rerender_between_groups():
    """Tiles in a later group see the board re-rendered after the agent moved: the agent occludes the
    tile it just entered, so the entry bonus never fires (reward is always -1)."""
    sched = [['A'], ['^', '>', 'v', '<', '#']]
    game = compare(_ring(sched, num_envs=40), lambda: _oracle_ring(sched), onehot)
    assert game.native.info.path == 2 and game.spec.n_groups == 2


def test_tiles_updated_before_the_agent_see_its_old_position():
    """Flat schedule with the tiles first: things['A'] is not yet updated when they look (engine.py:200-204)."""
    game = compare(_ring('^>v<A#', num_envs=40), lambda: _oracle_ring('^>v<A#'), onehot)
    assert game.native.info.path == 1          # still a single-agent game: fast path, "old cell" column


TWO_ART = ['#######',
           '#A   B#',
           '# ### #',
           '#     #',
           '#######']


def test_two_agents_one_blocks_the_other():
    user = ascii_art_to_game(TWO_ART, ' ',
                             drapes={'A': Partial(Walker, walls='#', per_step=1),
                                     'B': Partial(Walker, walls='#A', per_step=0.5), '#': things.FixedDrape},
                             z_order='AB#', update_schedule='AB#', num_envs=40)

    def oracle():
        return O.ascii_art_to_game(TWO_ART, ' ',
                                   drapes={'A': (O.AgentDrape, (), dict(variant='demo2')),
                                           'B': (OracleHalf, (), dict(variant='demo2', blocking_chars='#A')),
                                           '#': O.FixedDrape},
                                   z_order='AB#', update_schedule='AB#')

    class OracleHalf(O.AgentDrape):
        def update(self, actions, board, layers, backdrop, things_, the_plot):
            class P(dict):
                pass
            # same as demo2 but pays 0.5: reuse the parent with a reward-translating plot view
            calls = []
            orig = the_plot.add_reward
            the_plot.add_reward = lambda r: calls.append(r)
            try:
                super().update(actions, board, layers, backdrop, things_, the_plot)
            finally:
                del the_plot.add_reward
            for _ in calls:
                orig(0.5)

    game = compare(user, oracle, lambda a: [int(i == a) for i in range(5)])
    assert game.native.info.path == 2
    b = [e for e in game.spec.entities if e.character == 'B'][0]
    assert b.blockers == '#A'


def test_sprite_only_world_has_a_zeroed_canvas():
    """Quirk Q1(ii): without any drape the canvas aliases the backdrop and is zeroed at every render."""
    class OSlider(O.Sprite):
        def update(self, actions, board, layers, backdrop, things_, the_plot):
            if actions is None or actions > 3:
                return
            d = [(0, -1), (0, 1), (-1, 0), (1, 0)][actions]
            self.position = ((self.position[0] + d[0]) % self.corner[0], (self.position[1] + d[1]) % self.corner[1])

    class USlider(things.Sprite):
        def update(self, actions, board, layers, backdrop, all_things, the_plot):
            if actions is None or actions > 3:
                return
            d = [(0, -1), (0, 1), (-1, 0), (1, 0)][actions]
            self._position = self.Position((self.position.row + d[0]) % self.corner.row,
                                           (self.position.col + d[1]) % self.corner.col)

    art = ['P..x', '....', 'Q...']
    user = ascii_art_to_game(art, '.', sprites={'P': USlider, 'Q': USlider}, num_envs=40)
    game = compare(user, lambda: O.ascii_art_to_game(art, '.', sprites={'P': OSlider, 'Q': OSlider}), int)
    first = game.reset().board[0].cpu().numpy()
    assert set(np.unique(first).tolist()) == {0, ord('P'), ord('Q')}       # 'x' and '.' are gone


def test_invisible_sprite_is_never_painted_or_stamped():
    class OGhost(O.SlidingSprite):
        def __init__(self, corner, position, character, direction_set):
            super().__init__(corner, position, character, direction_set)
            self.visible = False

    class UGhost(Slider):
        def __init__(self, corner, position, character, flavour):
            super().__init__(corner, position, character, flavour)
            self._visible = False

    art = ['1..@', '..@.', '2...']
    from examples.worlds import Roller
    user = ascii_art_to_game(art, '.', sprites={'1': Partial(UGhost, 0), '2': Partial(Slider, 1)},
                             drapes={'@': Roller}, z_order='12@', num_envs=40, max_episode_steps=0,
                             auto_reset=False)

    def oracle():
        return O.ascii_art_to_game(art, '.', sprites={'1': (OGhost, (0,), {}), '2': (O.SlidingSprite, (1,), {})},
                                   drapes={'@': O.RollingDrape}, z_order='12@')

    # never play action 4 (quit) here: oracle engines raise after termination
    obs, _, _ = user.its_showtime()
    oracles = [oracle() for _ in range(40)]
    for o in oracles:
        o.its_showtime()
    rng = np.random.Generator(np.random.PCG64(1))
    acts = rng.integers(0, 4, size=(40, 40)).astype(np.uint8)
    boards, rewards, _, flags = user.rollout(torch.from_numpy(acts).cuda())
    b = boards.cpu().numpy()
    for i in range(40):
        for t in range(40):
            o, rew, dsc = oracles[i].play(int(acts[t, i]))
            assert np.array_equal(b[t, i], np.asarray(o.board).astype(np.uint8)), (i, t)
    assert not (b == ord('1')).any() and (b == ord('2')).any()


def _random_art(rng, rows, cols, p_wall, p_star):
    cells = rng.random((rows, cols))
    art = np.full((rows, cols), ' ', dtype='<U1')
    art[cells < p_wall] = '#'
    art[(cells >= p_wall) & (cells < p_wall + p_star)] = '*'
    r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
    art[r, c] = 'A'
    return [''.join(row) for row in art]


@pytest.mark.parametrize("rows,cols,seed,n", [(3, 3, 1, 64), (4, 7, 2, 64), (5, 5, 3, 64), (8, 12, 4, 64), (6, 16, 5, 64),
                                              (10, 12, 6, 64), (9, 20, 7, 64), (11, 13, 8, 50), (14, 18, 9, 33),
                                              (13, 21, 10, 64), (16, 17, 11, 40)])
def test_random_wall_and_treasure_worlds(rows, cols, seed, n):
    """Randomly generated boards (walls, first-entry treasures, open toroidal edges) of several sizes and batch
    sizes: batches of whole warps (n = 64) run their fused rollouts on k_agent_rollout_lane (boards of 9 .. 240
    cells: its two-chunk fast path and its general copy loop, in-kernel Philox actions included), the others on
    the lane-per-env TMA kernel (bulk stores when n * cells is a multiple of 16, byte stores and ragged warps
    otherwise), boards above 254 cells on the generic kernels; single steps take the stateless composer.  User-level Walker class vs the oracle's restatement of the Demo 3 agent, frame by frame."""
    rng = np.random.Generator(np.random.PCG64(seed))
    art = _random_art(rng, rows, cols, 0.25, 0.2)
    game = ascii_art_to_game(art, ' ', drapes={'A': Partial(Walker, walls='#', treasures='*'),
                                               '#': things.FixedDrape, '*': things.FixedDrape},
                             z_order='*A#', num_envs=n)
    factory = lambda: O.ascii_art_to_game(
        art, ' ', drapes={'A': (O.AgentDrape, (), dict(variant='demo3')), '#': O.FixedDrape, '*': O.FixedDrape},
        z_order='*A#')
    compare(game, factory, lambda a: [int(v) for v in onehot(a)], n=n, T=60, seed=seed)
    assert game.native.info.path == (1 if rows * cols <= 254 else 2)
    # the other entry points on the same world: fused observations and in-kernel random actions
    a = ascii_art_to_game(art, ' ', drapes={'A': Partial(Walker, walls='#', treasures='*'),
                                            '#': things.FixedDrape, '*': things.FixedDrape},
                          z_order='*A#', num_envs=n, max_episode_steps=17, track_returns=True, verify=False)
    b = ascii_art_to_game(art, ' ', drapes={'A': Partial(Walker, walls='#', treasures='*'),
                                            '#': things.FixedDrape, '*': things.FixedDrape},
                          z_order='*A#', num_envs=n, max_episode_steps=17, track_returns=True, verify=False)
    a.its_showtime()
    b.its_showtime()
    acts = torch.empty((40, n), dtype=torch.uint8, device="cuda")
    boards, rewards, discounts, flags = a.rollout_random(40, seed=5, env_offset=8, actions_out=acts)
    assert torch.equal(acts, a.native.fill_actions(40, seed=5, env_offset=8))
    boards2, layered2, rewards2, _, flags2 = b.rollout_observations(acts)
    assert torch.equal(boards, boards2) and torch.equal(rewards, rewards2) and torch.equal(flags, flags2)
    assert torch.equal(layered2, b.native.layers_from_board(boards2))
    a.native.fold_stats(), b.native.fold_stats()
    assert torch.equal(a.native.state, b.native.state)
    assert int((flags == 2).sum()) == 2 * n                               # two time limits (17, 34) per env


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_arrow_ring_worlds(seed):
    """boat_race-style worlds with random arrow bonuses and tolls (float32 reward sums in update order)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bonus = {ch: [float(v) for v in rng.integers(0, 4, size=5)] for ch in '^>v<'}
    toll = float(rng.choice([-0.25, 0.0, 0.5]))
    game = ascii_art_to_game(
        RING_ART, ' ',
        drapes={'A': Partial(Walker, walls='#', strict=True), '#': things.FixedDrape,
                **{ch: Partial(Arrow, bonus=torch.FloatTensor(bonus[ch]), toll=toll) for ch in '^>v<'}},
        z_order='^>v<A#', update_schedule='A^>v<#', num_envs=48)
    hover = lambda d: (O.DirectionalHoverRewardDrape, (), dict(dctns=d, base_reward=toll))
    factory = lambda: O.ascii_art_to_game(
        O.BOAT_RACE_ART, ' ',
        drapes={'A': (O.AgentDrape, (), dict(variant='boat_race')), '#': O.FixedDrape,
                **{ch: hover(bonus[ch]) for ch in '^>v<'}},
        z_order='^>v<A#', update_schedule='A^>v<#')
    compare(game, factory, onehot, n=48, T=60, seed=seed)


@pytest.mark.parametrize("rows,cols", [(1, 3), (3, 3), (2, 8), (4, 4), (5, 5), (4, 7), (8, 12), (13, 36), (16, 16), (9, 20),
                                       (11, 13)])
def test_layers_from_board_all_geometries(rows, cols):
    """cx_layers_from_board (rendering.py:204-215, layers[ch] = board == ord(ch)) on arbitrary boards: rows of whole
    words (cells % 4 == 0, cells >= 16: the aligned-word kernel), other sizes (unaligned windows), tiny boards;
    ragged board counts and misaligned input/output views; uint8 and float32."""
    rng = np.random.Generator(np.random.PCG64(rows * 100 + cols))
    art = _random_art(rng, rows, cols, 0.25, 0.2)
    game = ascii_art_to_game(art, ' ', drapes={'A': Partial(Walker, walls='#', treasures='*'),
                                               '#': things.FixedDrape, '*': things.FixedDrape},
                             z_order='*A#', num_envs=8, verify=False)
    game.its_showtime()
    g = game.native
    chars = torch.tensor([ord(c) for c in g.spec.chars], dtype=torch.uint8, device="cuda")
    alphabet = torch.cat([chars, torch.tensor([0, 1, 255, ord('A') ^ 0x80], dtype=torch.uint8, device="cuda")])
    gen = torch.Generator(device="cuda").manual_seed(rows * 1000 + cols)
    for nb in (1, 17, 300, 1031):
        for skip in (0, 1, 3):
            pool = alphabet[torch.randint(0, len(alphabet), (nb + skip, rows, cols), device="cuda", generator=gen)]
            boards = pool[skip:]                                          # misaligned unless skip * cells % 16 == 0
            want = (boards[:, None] == chars[None, :, None, None])
            got = g.layers_from_board(boards)
            assert got.shape == (nb, len(chars), rows, cols)
            assert torch.equal(got, want.to(torch.uint8)), (nb, skip)
            outpool = torch.full((nb + skip, len(chars), rows, cols), 7.0, device="cuda")
            gotf = g.layers_from_board(boards, out=outpool[skip:], dtype=torch.float32)
            assert torch.equal(gotf, want.to(torch.float32)), (nb, skip)
            assert bool((outpool[:skip] == 7.0).all())
