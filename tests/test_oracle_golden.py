"""Pins the CPU oracle to the reference's own artefacts.

The fixtures under tests/golden/ were decoded from /root/reference by
tests/golden/make_golden.py (BSON checkpoint + GIFs); nothing here reads /root/reference.
"""
import json
import os

import numpy as np

from oracle import oracle_lib as O
from oracle.xoshiro_food import default_food_list
from tests.util import G2_ACTIONS

G = os.path.join(os.path.dirname(__file__), "golden")


def test_g1_food_list_from_xoshiro42_matches_bson():
    gold = json.load(open(os.path.join(G, "g1_food_list.json")))
    cells, state = default_food_list(42, 50)
    assert [list(c) for c in cells] == gold["food_list_rc_1based"]
    # generator state after the 100 draws is stored in the checkpoint too (s0..s3)
    assert ["0x%016x" % w for w in state] == gold["xoshiro_state_after_draws_hex"][:4]
    # and the table the oracle / product use is that list
    assert [list(c) for c in O.DEFAULT_FOOD_RC] == gold["food_list_rc_1based"]
    assert len(set(O.DEFAULT_FOOD_RC)) == 35


def test_g4_initial_board_matches_gif_frame0_and_bson():
    boards = np.load(os.path.join(G, "g2_boards_double3.npy"))
    g = O.OracleGame()
    assert np.array_equal(g.board, boards[0]) and np.array_equal(g.board, boards[1])
    s = g.next_state()
    assert np.array_equal(s[:100].reshape(10, 10).T, boards[0])
    assert np.array_equal(s[100:].reshape(10, 10).T, boards[0])


def test_g4_bson_game_is_one_left_move_into_the_wall():
    """The checkpointed tr.game (older struct, same step rules) is a game lost on its first
    move: L from (8,2) into the wall at (8,1).  Pins R4 (wall), R6 (wall cell overwritten
    with 1, tail already popped) and the initial snake."""
    bson = json.load(open(os.path.join(G, "g4_bson_game.json")))
    g = O.OracleGame()
    g.step(2)
    assert g.lost and g.score == 0 and np.float32(g.reward) == np.float32(-1.0)
    assert np.array_equal(g.board, np.array(bson["board_rc"]))
    assert bson["snake_rc_1based"] == [[8, 1], [8, 2]]
    assert [4, 0] in bson["scalar_fields"] and [11, True] in bson["scalar_fields"]   # score 0, lost


def _replay(boards, n_steps):
    """Replays frames: every frame must be produced by exactly one available action."""
    g = O.OracleGame()
    actions, foods = [], []
    for t in range(1, n_steps + 1):
        target = boards[1 + t]
        av = g.available_actions()
        hits = []
        for a in av:
            h = O.OracleGame()
            for b in actions:
                h.step(b)
            h.step(a)
            if np.array_equal(h.board, target):
                hits.append(int(a))
        assert len(hits) == 1, (t, hits)
        score_before = g.score
        g.step(hits[0])
        actions.append(hits[0])
        if g.score > score_before:
            fr, fc = np.argwhere(g.board == 2)[0]
            foods.append((int(fr) + 1, int(fc) + 1))
        assert np.array_equal(g.board, target)
    return g, actions, foods


def test_g2_score33_trajectory_replays_exactly():
    boards = np.load(os.path.join(G, "g2_boards_double3.npy"))
    assert boards.shape == (240, 10, 10)
    g, actions, foods = _replay(boards, 237)
    assert g.lost and g.score == 33 and g.error == 0
    # README.md:54-58 best game = 33 apples; death by wall at (10,5), wall cell drawn as 1 (R6)
    assert g.board[9, 4] == 1 and boards[238][9, 4] == 1
    # frame 239 is the padding copy (utils.jl:223)
    assert np.array_equal(boards[239], boards[238])
    acts = "".join("UDLR"[a] for a in actions)
    assert acts == G2_ACTIONS
    # skip-and-keep food rule R5: (7,7) is skipped twice while occupied and used later
    assert foods == [(7, 5), (5, 7), (7, 3), (6, 7), (5, 4), (4, 4), (6, 2), (7, 7), (5, 5), (4, 3), (2, 6),
                     (3, 6), (5, 8), (7, 4), (5, 4), (4, 3), (2, 2), (4, 6), (7, 3), (7, 5), (6, 5), (4, 9),
                     (4, 7), (8, 6), (8, 3), (4, 4), (6, 6), (9, 3), (4, 2), (2, 8), (7, 4), (9, 5), (3, 9)]
    # rewards on that path: 33 eats, 203 plain moves, one death
    h = O.OracleGame()
    rs = []
    for a in actions:
        h.step(a)
        rs.append(np.float32(h.reward).view(np.uint32))
    assert rs.count(0x3F800000) == 33 and rs.count(0xBF800000) == 1 and rs.count(0xBC23D70A) == 203
    assert rs[-1] == 0xBF800000


def test_g5_old_gif_first_food_cells_match_list():
    boards = np.load(os.path.join(G, "g5_boards_training1.npy"))
    foods = []
    last = None
    for b in boards:
        pos = np.argwhere(b == 2)
        if len(pos) and (last is None or tuple(pos[0]) != last):
            last = tuple(pos[0])
            foods.append((int(last[0]) + 1, int(last[1]) + 1))
    assert foods[0] == (4, 5)                       # structs.jl:43
    assert foods[1:9] == O.DEFAULT_FOOD_RC[:8]


def test_history_modes_agree():
    rng = np.random.default_rng(0)
    a = O.OracleBatch(64, auto_reset=True, keep_history=True)
    b = O.OracleBatch(64, auto_reset=True, keep_history=False)
    for t in range(300):
        act = rng.integers(0, 3, 64).astype(np.uint8)
        ra, rb = a.step(act), b.step(act)
        for k in ("reward", "done", "mask", "obs_f32", "ep_return", "ep_score"):
            assert np.array_equal(ra[k], rb[k]), (t, k)


def test_config1_trace_is_stable():
    """BASELINE config 1 (1 env, 10,000 random-action steps): the oracle's trace is pinned by hash so that the
    oracle cannot drift silently; tools/trace_config1.py --cuda and bench_ref/trace_config1.jl write the same file."""
    import hashlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import trace_config1
    gold = json.load(open(os.path.join(G, "g7_config1_trace.json")))
    lines = trace_config1.trace_lines(10000)
    text = "\n".join(lines) + "\n"
    assert len(lines) == gold["n_lines"] and lines[:3] == gold["first"] and lines[-1] == gold["last"]
    assert hashlib.sha256(text.encode()).hexdigest() == gold["sha256"]
