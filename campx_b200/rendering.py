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
        self._planes = {}               # dtype -> layered board in that dtype (pre-filled by the fused step kernel)
        self.layers_were_read = False   # Engine.play() emits whole Observations for callers that read the layers
        self.read_dtype = None          # dtype of the layered board the caller asked for last

    @property
    def layered_board(self):
        self.layers_were_read = True
        self.read_dtype = torch.uint8
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
        self.read_dtype = dtype
        if dtype not in self._planes:
            self._planes[dtype] = self._layer_fn(self.board, dtype)
        return self._planes[dtype]

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


_TORCH_DTYPES = {
    'bool': torch.bool, 'uint8': torch.uint8, 'int8': torch.int8, 'int16': torch.int16, 'int32': torch.int32,
    'int64': torch.int64, 'float16': torch.float16, 'float32': torch.float32, 'float64': torch.float64,
}


def _batched_board(observation):
    """-> (contiguous uint8 board [..., rows, cols], number of leading batch dimensions)."""
    board = observation.board if hasattr(observation, 'board') else observation
    if not isinstance(board, torch.Tensor) or board.dtype != torch.uint8 or board.dim() < 2:
        raise TypeError('expected an Observation (or a uint8 board tensor [..., rows, cols]) of a campx_b200 Engine')
    return board.contiguous(), board.dim() - 2


class _BoardConverter(object):
    """Shared by the two converters below: a `[256, depth]` value table on the device and one launch per call."""

    def _setup(self, values, known, permute):
        self._values, self._known, self._permute = values, known, permute
        self._mapper = None
        self._unknown = None
        self._array = None

    def _convert(self, board, lead):
        from . import runtime
        if self._mapper is None or self._mapper.device != board.device:
            self._mapper = runtime.BoardMapper(self._values, self._known, device=board.device)
            self._unknown = torch.zeros(1, dtype=torch.int32, device=board.device)
            self._array = None
        depth = self._mapper.depth
        rows, cols = board.shape[-2], board.shape[-1]
        inner = (depth, rows, cols)
        if self._permute is not None and len(self._permute) == 3:
            inner = tuple(inner[p] for p in self._permute)
        shape = tuple(board.shape[:lead]) + inner
        # like the reference (rendering.py:560-564) the output buffer is reused from call to call
        if self._array is None or tuple(self._array.shape) != shape:
            self._array = torch.empty(shape, dtype=_TORCH_DTYPES[self._values.dtype.name], device=board.device)
        return shape


class ObservationToArray(_BoardConverter):
    """Convert `Observation`s to 2-D or 3-D arrays of mapped values (RGB images, repainted boards...).

    Same constructor, call signature and errors as the reference's `ObservationToArray`
    (campx/rendering.py:461-594); the conversion is one CUDA launch (`cx_board_mapper_apply`) over the whole
    environment batch and the result is a torch tensor on the board's device, with the batch dimensions of the
    board in front: `[num_envs, depth, rows, cols]` for vector values, `[num_envs, rows, cols]` for scalars,
    both reordered by `permute` exactly as the reference reorders its `(depth, rows, cols)` / `(rows, cols)`.

    `dtype` is a numpy dtype (or anything `np.dtype` accepts); when omitted it is inferred from the first value
    of the mapping like the reference does (rendering.py:506-507).  As in the reference the returned tensor is
    reused by the next call.  `check=False` skips the host synchronisation that turns an unmapped character
    into the reference's `RuntimeError` (for use inside CUDA graphs / asynchronous pipelines).
    """

    def __init__(self, value_mapping, dtype=None, permute=None, check=True):
        import numpy as np
        self._value_mapping = value_mapping
        first = next(iter(value_mapping.values()))
        np_dtype = np.dtype(dtype) if dtype is not None else np.array(first).dtype
        if np_dtype.name not in _TORCH_DTYPES:
            raise TypeError('ObservationToArray: values of dtype %s have no device representation' % np_dtype)
        try:
            depth = len(first)
            self._is_3d = True
        except TypeError:
            depth = 1
            self._is_3d = False
        permute = tuple(permute) if permute is not None else None
        if permute is not None:
            if self._is_3d and set(permute) != {0, 1, 2}:
                raise ValueError('When the value mapping contains 1-D vectors, the permute argument to the '
                                 'ObservationToArray constructor must be a list or tuple containing some '
                                 'permutation of the integers 0, 1, and 2.')
            elif not self._is_3d and set(permute) != {0, 1}:
                raise ValueError('When the value mapping contains scalars, the permute argument to the '
                                 'ObservationToArray constructor must be a list or tuple containing some '
                                 'permutation of the integers 0 and 1.')
        values = np.zeros((256, depth), dtype=np_dtype)
        known = np.zeros(256, dtype=np.uint8)
        for ch, value in value_mapping.items():
            code = ord(ch) if isinstance(ch, str) else int(ch)
            if not 0 <= code < 256:
                raise ValueError('ObservationToArray: character %r is not a single byte' % (ch,))
            values[code, :] = np.asarray(value, dtype=np_dtype).reshape(depth)
            known[code] = 1
        self._check_unknown = bool(check)
        # scalar values: the device permutation works on (vector, row, col) with a vector axis of size 1 in front
        device_permute = permute
        if permute is not None and not self._is_3d:
            device_permute = (0,) + tuple(p + 1 for p in permute)
        self._user_permute = permute
        self._setup(values, known, device_permute)

    def __call__(self, observation):
        board, lead = _batched_board(observation)
        shape = self._convert(board, lead)
        self._unknown.zero_()
        self._mapper.apply(board, self._array, permute=self._permute, unknown=self._unknown)
        if self._check_unknown and int(self._unknown.item()) != 0:
            raise RuntimeError(
                'This ObservationToArray only knows array values for the characters {}, but it received an '
                'observation with a character not in that set'.format(str(''.join(self._value_mapping.keys()))))
        if self._is_3d:
            return self._array
        # scalar mapping: drop the size-1 vector axis (it is the first inner axis for every 2-D permutation)
        return self._array.view(tuple(shape[:lead]) + tuple(shape[lead + 1:]))


class ObservationToFeatureArray(_BoardConverter):
    """Convert `Observation`s to float32 0/1 feature planes of chosen characters.

    Same constructor, call signature and errors as the reference's `ObservationToFeatureArray`
    (campx/rendering.py:597-712): plane k is `layers[layers_arg[k]]` cast to float32, all zeros for a character
    the game does not have; `permute` reorders (feature, row, col).  Since `layers[ch] == (board == ord(ch))`
    (rendering.py:204-209) the planes come straight from the board in one launch, for the whole batch:
    `[num_envs, len(layers), rows, cols]` (or permuted).
    """

    def __init__(self, layers, permute=None):
        import numpy as np
        self._layers = layers
        self._depth = len(layers)
        permute = tuple(permute) if permute is not None else None
        if permute is not None and sorted(permute) != [0, 1, 2]:
            raise ValueError('The permute argument to the ObservationToFeatureArray constructor must be a list or '
                             'tuple containing some permutation of the integers 0, 1, and 2.')
        self._np = np
        self._for_characters = ()
        self._setup(None, None, permute)

    def __call__(self, observation):
        characters = getattr(observation, 'characters', None)
        if characters is not None and not any(l in characters for l in self._layers):
            raise RuntimeError(
                'The layers argument to this ObservationToFeatureArray, {}, has no entry that refers to an actual '
                'feature in the input observation. Actual features in the observation are {}.'.format(
                    repr(self._layers), repr(''.join(sorted(characters)))))
        if self._values is None or characters != self._for_characters:
            # plane k is 1 where the board shows layers[k]; a character that is not a layer of this game gives
            # a zero plane (rendering.py:703-707) whatever the board holds
            np = self._np
            values = np.zeros((256, self._depth), dtype=np.float32)
            for k, ch in enumerate(self._layers):
                if characters is None or ch in characters:
                    values[ord(ch), k] = 1.0
            self._setup(values, np.ones(256, dtype=np.uint8), self._permute)
            self._for_characters = characters
        board, lead = _batched_board(observation)
        self._convert(board, lead)
        self._mapper.apply(board, self._array, permute=self._permute)
        return self._array
