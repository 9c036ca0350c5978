"""campx_b200: a B200-native batched grid-world engine with CampX's PyColab-style API.

    from campx_b200 import things
    from campx_b200.ascii_art import ascii_art_to_game, Partial
    game = ascii_art_to_game(ART, ' ', drapes={...}, z_order=..., update_schedule=..., num_envs=1 << 20)
    obs, reward, discount = game.its_showtime()
    obs, reward, discount = game.play(actions)        # one CUDA kernel launch for all environments

The step path is hand-written sm_100a CUDA behind a C ABI (include/campx_b200.h); there is no CPU
fallback.  See DESIGN.md.
"""
__version__ = "0.1.0"

from . import things  # noqa: F401
from . import plot  # noqa: F401
from . import rendering  # noqa: F401
