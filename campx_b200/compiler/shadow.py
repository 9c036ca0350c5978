"""Single-environment CPU shadow of a game, used ONLY by the game compiler.

The shadow runs the user's own entity objects (the ones `Engine.add_sprite` /
`add_prefilled_drape` / `set_prefilled_backdrop` constructed) through the reference's step
semantics so that their `update()` methods observe exactly what they would observe under CampX:

    step order, update groups, re-render after each group      campx/engine.py:168-208
    plot directives applied after the step                     campx/engine.py:211-293
    painter's algorithm with the reference's storage aliasing  campx/engine.py:295-324,
                                                               campx/rendering.py:104-219

`board` is an int64 [R,C] tensor, `layers[ch]` uint8 [R,C] tensors whose identity is stable across
renders (user code keeps aliases of them in the Plot, examples/boat_race.py:59,91).

torch >= 1.x no longer allows `Tensor.set_()` across dtypes, which CampX-era entity code relies on
(`self.curtain.set_(b)` with an int64 `b`, boat_race.py:57).  Curtains are therefore handed to user
code as `Curtain` tensors whose `set_` casts first.
"""
import collections
import copy

import torch

from .. import things
from ..plot import Plot


class Curtain(torch.Tensor):
    """uint8/int64 tensor whose `set_(src)` tolerates a dtype mismatch (torch 0.3.1 behaviour)."""

    @staticmethod
    def wrap(t):
        return t.as_subclass(Curtain)

    def set_(self, source=None, *args, **kwargs):
        if isinstance(source, torch.Tensor) and not args and not kwargs:
            src = source.as_subclass(torch.Tensor)
            if src.dtype != self.dtype:
                src = src.to(self.dtype)
            with torch._C.DisableTorchFunctionSubclass():
                torch.Tensor.set_(self, src)
            return self
        with torch._C.DisableTorchFunctionSubclass():
            return torch.Tensor.set_(self, source, *args, **kwargs)

    def __deepcopy__(self, memo):
        # Go through the plain-tensor deepcopy so that torch's storage memo keeps aliases alive: the
        # canvas shares storage with the backdrop curtain (rendering.py:128) and a clone of the shadow
        # must preserve that (quirk Q1).
        if id(self) in memo:
            return memo[id(self)]
        out = Curtain.wrap(copy.deepcopy(self.as_subclass(torch.Tensor), memo))
        memo[id(self)] = out
        return out


class RecordingPlot(Plot):
    """Plot that also remembers which entity issued which directive during the current step."""

    def __init__(self):
        super(RecordingPlot, self).__init__()
        self.current_entity = None
        self.events = []          # (entity_char, kind, payload)

    def add_reward(self, reward):
        self.events.append((self.current_entity, 'reward', reward))
        super(RecordingPlot, self).add_reward(reward)

    def terminate_episode(self, discount=0.0):
        super(RecordingPlot, self).terminate_episode(discount)
        self.events.append((self.current_entity, 'terminate', float(discount)))

    def change_default_discount(self, discount):
        super(RecordingPlot, self).change_default_discount(discount)
        self.events.append((self.current_entity, 'discount', float(discount)))

    def change_z_order(self, move_this, in_front_of_that):
        super(RecordingPlot, self).change_z_order(move_this, in_front_of_that)
        self.events.append((self.current_entity, 'z_order', (move_this, in_front_of_that)))


class ShadowRenderer(object):
    """The reference canvas (rendering.py:86-219) including where it aliases and where it copies."""

    def __init__(self, rows, cols, characters, occlusion_in_layers=True):
        self.board = torch.zeros((rows, cols), dtype=torch.int64)
        self.layers = {ch: torch.zeros((rows, cols), dtype=torch.uint8) for ch in characters}
        # False: layers follow BaseUnoccludedObservationRenderer (rendering.py:227-353) -- a layer holds its
        # entity's whole mask / position and the backdrop's own cells whether or not something is painted over them
        self.occluded = bool(occlusion_in_layers)

    def clear(self):
        self.board.mul_(0)                       # in place, on whatever storage the canvas aliases
        self._sprites, self._drapes, self._curtain = [], {}, None

    def paint_all_of(self, curtain):
        self.board = curtain.as_subclass(torch.Tensor)    # alias of the backdrop storage (set_)
        self._curtain = self.board                        # the backdrop's own storage (unoccluded layers, render())

    def paint_sprite(self, character, position):
        if character not in self.layers:
            raise ValueError('character {} does not seem to be a valid character for '
                             'this game'.format(str(character)))
        self.board[position[0], position[1]] = ord(character)
        self._sprites.append((character, (int(position[0]), int(position[1]))))

    def paint_drape(self, character, curtain):
        if character not in self.layers:
            raise ValueError('character {} does not seem to be a valid character for '
                             'this game'.format(str(character)))
        m = curtain.as_subclass(torch.Tensor).long()
        self.board = self.board - m * self.board + m * ord(character)   # fresh storage
        self._drapes[character] = (m != 0)

    def render(self):
        if not self.occluded:
            # read off the finished frame (rendering.py:283-286,309,333; same definition as the oracle's
            # UnoccludedRenderer): a drape's layer is its curtain; every other character's layer is where the
            # backdrop curtain holds it once the frame is complete, plus a sprite's own cell
            for ch, layer in self.layers.items():
                if ch in self._drapes:
                    layer.copy_(self._drapes[ch])
                else:
                    layer.copy_(self._curtain == ord(ch))
            for ch, (r, c) in self._sprites:
                self.layers[ch][r, c] = 1
            return
        for ch, layer in self.layers.items():
            layer.copy_(self.board == ord(ch))   # identity of layers[ch] is preserved


class ShadowEngine(object):
    """Reference step semantics over the user's entity objects (one environment, CPU)."""

    def __init__(self, rows, cols, backdrop, things_in_z_order, update_groups, occlusion_in_layers=True):
        self.rows, self.cols = rows, cols
        self.backdrop = backdrop
        self.things = collections.OrderedDict(things_in_z_order)
        self.update_groups = update_groups            # [(group_name, [entity, ...]), ...] sorted
        self.the_plot = RecordingPlot()
        chars = set(self.things.keys()).union(backdrop.palette)
        self.renderer = ShadowRenderer(rows, cols, chars, occlusion_in_layers)
        self.game_over = False
        self.showtime = False

    def clone(self):
        return copy.deepcopy(self)

    @property
    def board(self):
        return self.renderer.board

    @property
    def layers(self):
        return self.renderer.layers

    def its_showtime(self):
        self.showtime = True
        self.render()                                 # engine.py:541
        return self.play(None)                        # engine.py:544

    def play(self, actions):
        plot = self.the_plot
        plot.events = []
        plot.frame += 1
        plot.update_group = None
        plot.current_entity = None
        self.backdrop.update(actions, self.board, self.layers, self.things, plot)
        for name, entities in self.update_groups:
            plot.update_group = name
            for ent in entities:
                plot.current_entity = ent.character
                ent.update(actions, self.board, self.layers, self.backdrop, self.things, plot)
            plot.current_entity = None
            self.render()                             # engine.py:208
        d = plot._get_engine_directives()
        for move_this, in_front_of_that in d.z_updates:      # engine.py:242-281
            for ch in (move_this,) if in_front_of_that is None else (move_this, in_front_of_that):
                if ch not in self.things:
                    raise RuntimeError('A z-order change directive names character {}, but no such Sprite or '
                                       'Drape exists'.format(repr(ch)))
            mover = self.things[move_this]
            reordered = collections.OrderedDict()
            if in_front_of_that is None:
                reordered[move_this] = mover
            for ch, ent in self.things.items():
                if ch == move_this:
                    continue
                reordered[ch] = ent
                if ch == in_front_of_that:
                    reordered[move_this] = mover
            self.things = reordered
        if d.z_updates:
            self.render()                             # engine.py:163
        self.game_over = d.game_over
        reward, discount = d.summed_reward, d.discount
        plot._clear_engine_directives()
        return reward, discount

    def render(self):
        r = self.renderer
        r.clear()
        r.paint_all_of(self.backdrop.curtain)
        for ch, ent in self.things.items():
            if isinstance(ent, things.Sprite):
                if ent.visible:
                    r.paint_sprite(ch, ent.position)
            elif isinstance(ent, things.Drape):
                r.paint_drape(ch, ent.curtain)
        r.render()
