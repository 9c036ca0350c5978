"""The C-ABI library loads and exports every symbol include/campx_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from campx_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "campx_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"^CX_API\s+[\w\s\*]+?\b(cx_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("cx_game_create", "cx_game_destroy", "cx_reset", "cx_render", "cx_step", "cx_rollout",
                 "cx_layers_from_board", "cx_last_error", "cx_fill_actions", "cx_step_perf"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    assert os.path.exists(N.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "libcampx_b200.so does not export %s" % name


def test_binding_covers_every_declared_symbol_and_struct_layouts_match():
    assert sorted(N.PROTOTYPES) == declared_symbols()
    lib = N.load()                      # also checks ABI version and sizeof() of the three structs
    assert lib.cx_abi_version() == N.CX_ABI_VERSION
    assert lib.cx_abi_sizeof(0) == ctypes.sizeof(N.EntityDesc)
    assert lib.cx_abi_sizeof(1) == ctypes.sizeof(N.GameDesc)
    assert lib.cx_abi_sizeof(2) == ctypes.sizeof(N.GameInfo)


def test_header_constants_match_binding():
    with open(os.path.join(ROOT, "include", "campx_b200.h")) as f:
        text = f.read()
    for name in ("CX_ABI_VERSION", "CX_MAX_ENTITIES", "CX_MAX_ACTIONS", "CX_MAX_CHARS", "CX_MAX_CELLS",
                 "CX_MAX_GROUPS", "CX_STATS_DOUBLES"):
        m = re.search(r"#define\s+%s\s+(\d+)" % name, text)
        assert m and int(m.group(1)) == getattr(N, name), name
    for name in ("CX_FLAG_TERMINATED", "CX_FLAG_TRUNCATED", "CX_FLAG_REWARD_NONE", "CX_FLAG_ALREADY_OVER",
                 "CX_FLAG_BAD_ACTION"):
        m = re.search(r"#define\s+%s\s+(0x[0-9A-Fa-f]+)" % name, text)
        assert m and int(m.group(1), 16) == getattr(N, name), name


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under campx_b200/ or examples/ may reference it."""
    for base in ("campx_b200", "examples"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    with open(os.path.join(dirpath, fn)) as f:
                        src = f.read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, fn)


def test_no_cpu_fallback_without_a_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    from examples.worlds import make_world
    game = make_world("boat_race", num_envs=4)
    with pytest.raises(N.NativeLibraryError):
        game.its_showtime()
