"""Three small worlds that use the engine features none of CampX's own worlds touch:

    zswap    `the_plot.change_z_order(...)` directives          campx/plot.py:121-159, engine.py:242-281,163
    ghost    sprites that hide and show themselves              campx/things.py:320,390-392, engine.py:315
    scroll   a `Backdrop` subclass whose `update()` scrolls it  campx/things.py:103-148

written, like examples/worlds.py, as ordinary single-environment `Sprite` / `Drape` / `Backdrop`
subclasses.  The same three worlds exist against the reference API (oracle/gen_golden.py, executed on the
unmodified reference to record tests/golden/generality_*.json) and as numpy oracle entities
(oracle/campx_oracle.py); the GPU tests compare all of them frame by frame.  They also cover the interplay
with quirk Q1: which sprites stamp their character into the backdrop depends on the CURRENT z-order and
visibility.
"""
import numpy as np
import torch

from campx_b200 import things
from campx_b200.ascii_art import ascii_art_to_game, Partial
from examples.worlds import Walker

ZSWAP_ART = ['S.....',
             '.XX.Y.',
             '.XXYY.',
             '...YY.',
             '......']

GHOST_ART = ['G..@.',
             '..@..',
             'H....']

SCROLL_ART = ['A.~~.^',
              '.~..^.',
              '~..^..',
              '..^..~']


def _rolled(curtain, shift, axis):
    return torch.from_numpy(np.roll(curtain.numpy(), shift, axis).copy())


class Shuffler(things.Drape):
    """Slides right on action 0; actions 1-3 reorder the z-order."""

    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        if actions is None:
            return
        if actions == 0:
            self.curtain.set_(_rolled(self.curtain, 1, 1))
            the_plot.add_reward(1)
        elif actions == 1:
            the_plot.change_z_order('X', 'Y')
        elif actions == 2:
            the_plot.change_z_order('Y', None)
        elif actions == 3:
            the_plot.change_z_order('X', None)
            the_plot.change_z_order('Y', 'X')


class Climber(things.Sprite):
    """Moves right / down; action 4 brings it to the front, action 2 sends it to the back."""

    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        if actions is None:
            return
        row, col = self.position
        if actions == 0:
            col = (col + 1) % self.corner.col
        elif actions == 1:
            row = (row + 1) % self.corner.row
        self._position = self.Position(row, col)
        if actions == 4:
            the_plot.change_z_order('S', 'Y')
        elif actions == 2:
            the_plot.change_z_order('S', None)


class Phantom(things.Sprite):
    """A sprite that can hide: action 4 toggles; `switchable` ones also hide on 2 and show on 3."""

    def __init__(self, corner, position, character, start_visible, d_row, d_col, switchable):
        super(Phantom, self).__init__(corner, position, character)
        self._visible = start_visible
        self._step = (d_row, d_col)
        self._switchable = switchable

    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        if actions is None:
            return
        row, col = self.position
        if actions == 0:
            col = (col + self._step[1]) % self.corner.col
        elif actions == 1:
            row = (row + self._step[0]) % self.corner.row
        self._position = self.Position(row, col)
        if self._switchable and actions in (2, 3):
            self._visible = actions == 3
        if actions == 4:
            self._visible = not self._visible


class Drizzle(things.Drape):
    """Falls one row on action 1; pays 0.5 every step."""

    def update(self, actions, board, layers, backdrop, all_things, the_plot):
        if actions is None:
            return
        if actions == 1:
            self.curtain.set_(_rolled(self.curtain, 1, 0))
        the_plot.add_reward(0.5)


class Panorama(things.Backdrop):
    """Scenery that scrolls one cell in the direction of the (one-hot) action."""
    SCROLL = ((-1, 1), (1, 1), (-1, 0), (1, 0))      # (shift, axis) for left, right, up, down

    def update(self, actions, board, layers, all_things, the_plot):
        if actions is None:
            return
        a = int(torch.as_tensor(actions).argmax())
        if a < 4:
            shift, axis = self.SCROLL[a]
            self.curtain.set_(_rolled(self.curtain, shift, axis))


def make_generality_world(name, **engine_kwargs):
    """Un-started Engine for one of: zswap, ghost, scroll."""
    if name == 'zswap':
        return ascii_art_to_game(ZSWAP_ART, '.', sprites={'S': Climber},
                                 drapes={'X': Shuffler, 'Y': things.FixedDrape},
                                 update_schedule='SXY', z_order='SXY', **engine_kwargs)
    if name == 'ghost':
        return ascii_art_to_game(GHOST_ART, '.',
                                 sprites={'G': Partial(Phantom, True, 1, 1, True),
                                          'H': Partial(Phantom, False, 1, -1, False)},
                                 drapes={'@': Drizzle}, update_schedule='GH@', z_order='GH@', **engine_kwargs)
    if name == 'scroll':
        return ascii_art_to_game(SCROLL_ART, '.',
                                 drapes={'A': Partial(Walker, walls='^', per_step=1, strict=True)},
                                 backdrop=Panorama, z_order='A', **engine_kwargs)
    raise KeyError(name)
