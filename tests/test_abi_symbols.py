"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/snake_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from tests.util import ROOT, graft, pkg

HEADER = os.path.join(ROOT, "include", "snake_b200.h")


@pytest.fixture(scope="module")
def built():
    graft.build()
    return pkg()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SNK_API[^;(]*?\b(snk_\w+)\s*\(", src)))


def test_header_declares_the_whole_boundary():
    names = declared_symbols()
    for must in ("snk_create", "snk_destroy", "snk_reset", "snk_set_food_list_host", "snk_available_actions",
                 "snk_step", "snk_step_abs", "snk_step_fused", "snk_step_fused_host", "snk_state",
                 "snk_losing_mask", "snk_select_action", "snk_masked_target", "snk_get_score", "snk_sync",
                 "snk_last_error", "snk_center_columns"):
        assert must in names
    # every declaration cites the reference interface it replaces
    src = open(HEADER).read()
    for cite in ("structs.jl:33-99", "utils.jl:7-10", "utils.jl:100-109", "utils.jl:112-132", "utils.jl:135-149",
                 "utils.jl:153-172", "utils.jl:448-451", "compute_D.jl"):
        assert cite in src, cite


def test_library_exports_every_declared_symbol(built):
    S = built
    out = subprocess.check_output(["nm", "-D", "--defined-only", S.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [n for n in declared_symbols() if n not in exported]
    assert not missing, missing
    # nothing but the ABI is exported
    assert all(n.startswith("snk_") for n in exported), sorted(n for n in exported if not n.startswith("snk_"))
    L = S.lib()
    for n in declared_symbols():
        assert hasattr(L, n)
    assert L.snk_version() == 200


def test_library_is_sm100a_native_and_has_no_oracle_dependency(built):
    S = built
    sass = subprocess.run(["cuobjdump", "-lelf", S.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    needed = subprocess.check_output(["readelf", "-d", S.LIB_PATH], text=True)
    assert "oracle" not in needed


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_means_loud_failure_not_fallback(built):
    S = built
    h = C.c_void_p()
    rc = S.lib().snk_create(C.byref(h), 16, 0, 0)
    assert rc == -3 and not h.value                       # SNK_ERR_NODEVICE
    assert b"no CPU fallback" in S.lib().snk_last_error()
    with pytest.raises(S.SnakeB200Error):
        S.SnakeGame(16)


def test_argument_validation_without_touching_the_gpu(built):
    S = built
    L = S.lib()
    assert L.snk_masked_target(None, None, None, None, 0.97, -100.0, None, None, 4, None) == -1
    assert b"null input" in L.snk_last_error()
    assert L.snk_center_columns(None, 10, 10, None, None, None) == -1
    h = C.c_void_p()
    assert L.snk_create(C.byref(h), 0, 0, 0) == -1
    assert L.snk_create(None, 8, 0, 0) == -1
    assert L.snk_step(None, None, None, None) == -1
    assert L.snk_destroy(None) == 0
    with pytest.raises(ValueError):
        S.SnakeGame(4, board_size=12)
    with pytest.raises(ValueError):
        S.SnakeGame(4, n_frames=1)


def test_default_food_list_is_the_pinned_xoshiro42_list(built):
    from oracle import oracle_lib as O
    assert built.default_food_list() == O.DEFAULT_FOOD_RC


def test_product_sources_never_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the product package (host mirror, CUDA sources, Julia
    wrapper) may import, include, link or execute it."""
    pkg_dir = os.path.join(ROOT, "laplace-dqn-snake-game_b200")
    offenders = []
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".jl", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle_lib|libsnake_oracle|oracle/", txt):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders


def test_ctypes_binding_matches_the_header(built):
    """every prototype the Python binding declares has the header's parameter count and compatible ctypes (the Julia wrapper is
    checked the same way in test_julia_wrapper_static.py)"""
    from tests.test_julia_wrapper_static import c_declarations
    decls, handles = c_declarations()
    L = built.lib()
    table = {"int": {C.c_int}, "int32_t": {C.c_int, C.c_int32}, "int64_t": {C.c_int64, C.c_longlong}, "uint32_t": {C.c_uint32, C.c_uint},
             "uint64_t": {C.c_uint64, C.c_ulonglong}, "float": {C.c_float}, "double": {C.c_double}, "size_t": {C.c_size_t}}
    problems, checked = [], 0
    for name, params in decls.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue                                     # bound ad hoc by a tool (debug hooks)
        checked += 1
        if len(fn.argtypes) != len(params):
            problems.append("%s: %d ctypes arguments for %d C parameters" % (name, len(fn.argtypes), len(params)))
            continue
        for k, (cp, at) in enumerate(zip(params, fn.argtypes)):
            c = cp.replace("const ", "").strip()
            base = c.rsplit(" ", 1)[0].strip() if " " in c else c
            is_ptr = "*" in c or base in handles
            ok = (at is C.c_void_p or at is C.c_char_p or hasattr(at, "contents")) if is_ptr else at in table.get(base, set())
            if not ok:
                problems.append("%s: argument %d is `%s` in C but %s in the binding" % (name, k + 1, cp, at))
    assert checked >= 60 and not problems, "\n".join(problems)


def test_gram_tile_plan_arithmetic_without_a_gpu(built):
    """snk_gram_block_flops is the library's own tile plan (no launch): a symmetric block computes only the tiles that meet the
    upper triangle, 3 products per k-step for the hi/lo split, on the padded problem"""
    L = built.lib()
    v = C.c_double(0)
    P, Ppad = 181395, 181440
    tile = 2.0 * 256 * 256 * Ppad
    assert L.snk_gram_block_flops(1000, 1000, P, 3, 1, C.byref(v)) == 0 and v.value == 10 * 3 * tile        # 4 x 4 tiles -> 10
    assert L.snk_gram_block_flops(1000, 1000, P, 3, 0, C.byref(v)) == 0 and v.value == 16 * 3 * tile
    assert L.snk_gram_block_flops(6250, 6250, P, 3, 1, C.byref(v)) == 0 and v.value == 325 * 3 * tile       # 25 x 25 -> 325
    assert L.snk_gram_block_flops(3125, 6250, P, 3, 0, C.byref(v)) == 0 and v.value == 13 * 25 * 3 * tile
    assert L.snk_gram_block_flops(6250, 6250, P, 1, 1, C.byref(v)) == 0 and v.value == 325 * tile
    assert L.snk_gram_block_flops(10, 20, P, 3, 1, C.byref(v)) == -1                                         # symmetric needs a square block
    assert L.snk_gram_block_flops(10, 10, P, 2, 0, C.byref(v)) == -1
