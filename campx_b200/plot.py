"""The Plot: a blackboard shared by all entities plus the engine directives they may issue.

Same public methods as the reference's `campx/plot.py` (`add_reward` :186-211, `terminate_episode`
:161-184, `change_z_order` :121-159, `change_default_discount` :232-257, `log` :213-230, `frame` /
`update_group` :259-285).  It is used on the host by the game compiler's single-env shadow; on the
device the same directives are per-action tables (see `cx_entity_desc`).
"""


class _Directives(object):
    __slots__ = ('z_updates', 'summed_reward', 'game_over', 'discount')

    def __init__(self):
        self.z_updates = []
        self.summed_reward = None
        self.game_over = False
        self.discount = 1.0


class Plot(dict):

    def __init__(self):
        super(Plot, self).__init__()
        self._frame = -1
        self._update_group = None
        self._clear_engine_directives()

    # -- directives --------------------------------------------------------------------------------
    def add_reward(self, reward):
        d = self._engine_directives
        d.summed_reward = reward if d.summed_reward is None else reward + d.summed_reward

    def terminate_episode(self, discount=0.0):
        if not 0.0 <= discount <= 1.0:
            raise ValueError('Pcontinue must be in range [0,1]')
        self._engine_directives.game_over = True
        self._engine_directives.discount = discount

    def change_default_discount(self, discount):
        if not 0.0 <= discount <= 1.0:
            raise ValueError('Pcontinue must be in range [0,1]')
        self._engine_directives.discount = discount

    def change_z_order(self, move_this, in_front_of_that):
        for ch in (move_this,) if in_front_of_that is None else (move_this, in_front_of_that):
            try:
                ord(ch)
            except TypeError:
                raise ValueError('{} was used as an argument in a call to change_z_order, but only '
                                 'single ASCII characters are valid arguments'.format(repr(ch)))
        self._engine_directives.z_updates.append((move_this, in_front_of_that))

    def log(self, message):
        self.setdefault('log_messages', []).append(message)

    # -- bookkeeping ----------------------------------------------------------------------------------
    @property
    def frame(self):
        return self._frame

    @frame.setter
    def frame(self, val):
        assert val == self._frame + 1
        self._frame = val

    @property
    def update_group(self):
        return self._update_group

    @update_group.setter
    def update_group(self, group):
        self._update_group = group

    @property
    def default_discount(self):
        return self._engine_directives.discount

    def _clear_engine_directives(self):
        self._engine_directives = _Directives()

    def _get_engine_directives(self):
        return self._engine_directives
