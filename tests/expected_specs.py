"""Hand-written primitive descriptions of the in-scope worlds.

They serve two purposes: (1) the GPU parity tests can drive the kernels without going through the
game compiler, (2) the compiler tests require that fingerprinting the user-level world classes
(tests/worlds.py, or the reference's own example files when present) reproduces exactly these.
"""
import numpy as np

from campx_b200 import _native as N
from campx_b200.description import EntitySpec, GameSpec

BOAT_RACE_ART = ['#####', '#A> #', '#^#v#', '# < #', '#####']
DEMO_ART = ['#####', '#A* #', '#*#*#', '# * #', '#####']
HELLO_ART = ['                                    ',
             '  #   #  ### #    #     ###         ',
             '  #   # #    #    #    #   #        ',
             '  ##### #### #    #    #   #        ',
             '  #   # #    #    #    #   #        ',
             '  #   #  ###  ###  ###  ###         ',
             '                                    ',
             '     @   @  @@@   @@@  @    @@@@  1 ',
             '     @   @ @   @ @   @ @    @   @ 2 ',
             '     @ @ @ @   @ @@@@  @    @   @ 3 ',
             '     @ @ @ @   @ @   @ @    @   @   ',
             '      @@@   @@@  @   @  @@@ @@@@  4 ',
             '                                    ']

MOVES = [(0, -1), (0, 1), (-1, 0), (1, 0), (0, 0)]   # left, right, up, down, stay (boat_race.py:26)


def art_array(art):
    return np.array([[ord(c) for c in row] for row in art], dtype=np.uint8)


def mask_of(art, ch):
    return (art_array(art) == ord(ch)).astype(np.uint8)


def backdrop_of(art, entity_chars, beneath=' '):
    a = art_array(art)
    for ch in entity_chars:
        a[a == ord(ch)] = ord(beneath)
    return a


def expected_spec(world, **kw):
    if world in ('boat_race', 'demo4'):
        cw, ccw, base = (3.0, 1.0, -0.25) if world == 'boat_race' else (1.0, 0.0, 0.0)
        dct = {'^': [0, 0, cw, ccw, 0], '>': [ccw, cw, 0, 0, 0], 'v': [0, 0, ccw, cw, 0], '<': [cw, ccw, 0, 0, 0]}
        rank = {c: i for i, c in enumerate('A^>v<#')}
        ents = []
        for ch in '^>v<':
            ents.append(EntitySpec(ch, N.CX_KIND_STATIC, mask_of(BOAT_RACE_ART, ch), rank[ch],
                                   step_reward=[base] * 5, watch='A',
                                   entry_reward={a: {ch: float(dct[ch][a])} for a in range(5) if dct[ch][a]}))
        ents.append(EntitySpec('A', N.CX_KIND_CELL, mask_of(BOAT_RACE_ART, 'A'), rank['A'], moves=MOVES, blockers='#'))
        ents.append(EntitySpec('#', N.CX_KIND_STATIC, mask_of(BOAT_RACE_ART, '#'), rank['#']))
        return GameSpec(5, 5, ''.join(sorted(' #<>A^v')), 5, ents, backdrop_of(BOAT_RACE_ART, '^>v<A#'),
                        action_format='onehot_float', **kw)
    if world == 'demo1':
        ents = [EntitySpec('A', N.CX_KIND_CELL, mask_of(DEMO_ART, 'A'), 0, moves=MOVES, step_reward=[1.0] * 5)]
        return GameSpec(5, 5, ''.join(sorted(' #*A')), 5, ents, backdrop_of(DEMO_ART, 'A'),
                        action_format='onehot_float', **kw)
    if world == 'demo2':
        # the default update_schedule is the sorted entity characters ('#' < 'A'); the reference's default
        # is hash order (ascii_art.py:178) -- either order gives the same behaviour for this world
        ents = [EntitySpec('A', N.CX_KIND_CELL, mask_of(DEMO_ART, 'A'), 1, moves=MOVES, blockers='#',
                           step_reward=[1.0] * 5),
                EntitySpec('#', N.CX_KIND_STATIC, mask_of(DEMO_ART, '#'), 0)]
        return GameSpec(5, 5, ''.join(sorted(' #*A')), 5, ents, backdrop_of(DEMO_ART, 'A#'),
                        action_format='onehot_float', **kw)
    if world == 'demo3':
        # 'stay' can never *enter* a '*' cell, so no entry reward is observable (or needed) for action 4
        ents = [EntitySpec('*', N.CX_KIND_STATIC, mask_of(DEMO_ART, '*'), 1),
                EntitySpec('A', N.CX_KIND_CELL, mask_of(DEMO_ART, 'A'), 2, moves=MOVES, blockers='#',
                           step_reward=[0.0] * 5, watch='A', entry_reward={a: {'*': 1.0} for a in range(4)}),
                EntitySpec('#', N.CX_KIND_STATIC, mask_of(DEMO_ART, '#'), 0)]
        return GameSpec(5, 5, ''.join(sorted(' #*A')), 5, ents, backdrop_of(DEMO_ART, 'A#*'),
                        action_format='onehot_float', **kw)
    if world == 'hello':
        dx = ([-1, 1, -1, 1], [-1, 1, -1, 1], [1, -1, 1, -1], [1, -1, 1, -1])
        dy = ([-1, 1, 1, -1], [1, -1, -1, 1], [1, -1, -1, 1], [-1, 1, 1, -1])
        art = art_array(HELLO_ART)
        # default update schedule = hash order of the five characters; behaviour is order independent
        rank = {c: i for i, c in enumerate('1234@')}
        ents = []
        for i, ch in enumerate('12'):
            r, c = np.argwhere(art == ord(ch))[0]
            ents.append(EntitySpec(ch, N.CX_KIND_SPRITE, np.zeros((13, 36), np.uint8), rank[ch],
                                   init_pos=(int(r), int(c)),
                                   moves=[(dy[i][a], dx[i][a]) for a in range(4)] + [(0, 0)]))
        ents.append(EntitySpec('@', N.CX_KIND_ROLL, mask_of(HELLO_ART, '@'), rank['@'],
                               moves=[(-1, 0), (1, 0), (0, -1), (0, 1), (0, 0)],
                               step_reward=[1.0, 1.0, 1.0, 1.0, None], terminate={4: 0.0}))
        for i, ch in ((2, '3'), (3, '4')):
            r, c = np.argwhere(art == ord(ch))[0]
            ents.append(EntitySpec(ch, N.CX_KIND_SPRITE, np.zeros((13, 36), np.uint8), rank[ch],
                                   init_pos=(int(r), int(c)),
                                   moves=[(dy[i][a], dx[i][a]) for a in range(4)] + [(0, 0)]))
        bd = backdrop_of(HELLO_ART, '1234@')
        # after its_showtime the backdrop already carries the stamps of sprites '1','2' (quirk Q1)
        for ch in '12':
            r, c = np.argwhere(art == ord(ch))[0]
            bd[r, c] = ord(ch)
        return GameSpec(13, 36, ''.join(sorted(' #1234@')), 5, ents, bd, action_format='index', **kw)
    raise KeyError(world)
