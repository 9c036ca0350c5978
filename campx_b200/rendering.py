"""Observations returned by `Engine.its_showtime()` / `Engine.play()`.

The reference's `Observation` (campx/rendering.py:29) is a namedtuple
`(board, layers, layered_board)` built on the CPU by `BaseObservationRenderer.render()`
(rendering.py:181-219).  Here the board is composed inside the step kernel; `layers` and
`layered_board` are derived from the finished board exactly as the reference derives them
(`layers[ch] = board == ord(ch)`, rendering.py:204-209) by `cx_layers_from_board`, lazily, the first
time either is touched -- so a caller that only needs the board never pays for the other 7x bytes.

Differences that are deliberate and documented:
  * tensors live on the GPU, carry a leading batch dimension `[num_envs, ...]` (absent for an
    Engine created without `num_envs`) and are uint8 (the reference's are int64 / uint8); values are
    identical;
  * `layered_board` channels are in canonical order (sorted by code point); the reference's order
    follows `list(set(chars))`, i.e. PYTHONHASHSEED (rendering.py:198).
    `Observation.characters` names the channels.
As in the reference (rendering.py:56-64) the tensors are only valid until the next `play()`.
"""
import collections.abc

import torch


class LayersView(collections.abc.Mapping):
    """Read-only `{character: mask}` mapping over the channels of a layered board."""

    def __init__(self, obs):
        self._obs = obs

    def __getitem__(self, ch):
        k = self._obs.characters.find(ch) if isinstance(ch, str) and len(ch) == 1 else -1
        if k < 0:
            raise KeyError(ch)
        lb = self._obs.layered_board
        return lb[..., k, :, :]

    def __iter__(self):
        return iter(self._obs.characters)

    def __len__(self):
        return len(self._obs.characters)


class Observation(object):
    """`(board, layers, layered_board)`; unpacks and indexes like the reference's namedtuple."""

    _fields = ('board', 'layers', 'layered_board')

    def __init__(self, board, characters, layer_fn, layered_board=None):
        self.board = board
        self.characters = characters
        self._layer_fn = layer_fn
        self._layered = layered_board
        self._layers = None

    @property
    def layered_board(self):
        if self._layered is None:
            self._layered = self._layer_fn(self.board)
        return self._layered

    @property
    def layers(self):
        if self._layers is None:
            self._layers = LayersView(self)
        return self._layers

    def layered_board_as(self, dtype=torch.float32):
        """layered_board in another dtype straight from the board (float32: policy input)."""
        if dtype == torch.uint8:
            return self.layered_board
        return self._layer_fn(self.board, dtype)

    def __iter__(self):
        yield self.board
        yield self.layers
        yield self.layered_board

    def __len__(self):
        return 3

    def __getitem__(self, i):
        return (self.board, self.layers, self.layered_board)[i]

    def __repr__(self):
        return 'Observation(board=%s%s, characters=%r)' % (
            tuple(self.board.shape), ' on %s' % self.board.device, self.characters)
