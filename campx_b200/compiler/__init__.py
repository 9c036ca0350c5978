"""Game compiler front end: turns user Sprite/Drape/Backdrop classes into kernel primitives.

`compile_game(engine)` runs at `Engine.its_showtime()`.  It executes the user's unmodified
`update()` methods on a single-environment CPU *shadow* of the game (shadow.py) for a set of
probes (every discrete action from every reachable agent cell, plus teleported positions/offsets
for sprites and rolling drapes), fits each entity to one of the four kernel primitives and checks
the fit on every probe (fingerprint.py).  The result is a `GameSpec` (description.py) which the
native back end lowers to device tables.  The shadow is compile-time only: it is never on the step
path and there is no CPU fallback for stepping.
"""
from .fingerprint import compile_game, CompileError  # noqa: F401
