"""Game entities: the plugin interface of the engine.

Mirrors the reference's `campx/things.py` public surface (same class names, constructor
signatures, `update()` signatures and read-only properties) so that world files written for CampX
keep working after `s/campx/campx_b200/`:

    Backdrop(curtain, palette).update(actions, board, layers, things, the_plot)   things.py:59-158
    Drape(curtain, character).update(actions, board, layers, backdrop, things, the_plot)   :161-262
    Sprite(corner, position, character).update(... same ...)                      :265-392
    FixedDrape                                                                    :395-398

How `update()` is used here differs from the reference: user `update()` code is never run on the
step path.  At `Engine.its_showtime()` the game compiler (campx_b200/compiler) executes it on a
single-environment CPU shadow of the game to fingerprint each entity as one of the kernel
primitives; from then on the CUDA kernels advance all environments.
"""
import abc
import collections


class Backdrop(object):
    """Background scenery: an [rows, cols] tensor of character codes painted first."""

    def __init__(self, curtain, palette):
        self._curtain = curtain
        self._palette = palette

    def update(self, actions, board, layers, things, the_plot):
        """Default scenery never changes."""

    @property
    def curtain(self):
        return self._curtain

    @property
    def palette(self):
        return self._palette


class Drape(abc.ABC):
    """A binary mask painted with one character."""

    def __init__(self, curtain, character):
        self._curtain = curtain
        self._character = character

    @abc.abstractmethod
    def update(self, actions, board, layers, backdrop, things, the_plot):
        """Change `self.curtain` in response to `actions` and the last rendered `board`/`layers`."""

    @property
    def character(self):
        return self._character

    @property
    def curtain(self):
        return self._curtain


class Sprite(abc.ABC):
    """A single cell painted with one character."""

    Position = collections.namedtuple('Position', ['row', 'col'])

    def __init__(self, corner, position, character):
        self._corner = corner
        self._character = character
        self._position = position
        self._visible = True

    @abc.abstractmethod
    def update(self, actions, board, layers, backdrop, things, the_plot):
        """Replace `self._position` in response to `actions` and the last rendered board."""

    @property
    def character(self):
        return self._character

    @property
    def corner(self):
        return self._corner

    @property
    def position(self):
        return self._position

    @property
    def visible(self):
        return self._visible


class FixedDrape(Drape):
    """A drape that never moves (walls, reward tiles)."""

    def update(self, actions, board, layers, backdrop, things, the_plot):
        pass
