"""Edge rules of the reference restated in SURVEY §8-SPEC, checked on the oracle itself with explicit
expected values (the GPU tests then check the CUDA path against the oracle on the same cases)."""
import numpy as np

from oracle import oracle_lib as O


def _cycle(t):
    return [3, 1, 2, 0][t % 4]          # R, D, L, U: a 2x2 loop that never eats


def test_r2_available_actions_order():
    g = O.OracleGame()
    assert g.available_actions().tolist() == [0, 2, 3]          # prev U -> [U, L, R]
    g.step(3)
    assert g.available_actions().tolist() == [0, 1, 3]          # prev R -> [U, D, R]
    g.step(1)
    assert g.available_actions().tolist() == [1, 2, 3]          # prev D -> [D, L, R]
    g.step(2)
    assert g.available_actions().tolist() == [0, 1, 2]          # prev L -> [U, D, L]


def test_r7_cap_and_mask_one_step_early():
    g = O.OracleGame(keep_history=True)
    for t in range(498):
        g.step(_cycle(t))
        assert not g.lost
        assert not g.virtual_step()[1].all()
    g.step(_cycle(498))
    assert not g.lost and g.n_hist == 501
    assert g.virtual_step()[1].tolist() == [1, 1, 1]            # all-true after 499 real steps
    g.step(_cycle(499))
    assert g.lost and np.float32(g.reward) == np.float32(-1.0)
    av, lost = g.virtual_step()
    assert lost.tolist() == [1, 1, 1] and av.tolist() == [255, 255, 255]   # placeholder for a lost game


def test_r3_eat_then_die_keeps_the_score():
    # list with one cell right above the first food so the snake eats twice going up, then hits the wall
    g = O.OracleGame(food_rc=[(3, 5), (2, 5)])
    path = [3, 3, 3, 0, 0, 0, 0]        # R R R to column 5, then U x4: (7,5) (6,5) (5,5) (4,5)=eat
    for a in path:
        g.step(a)
    assert g.score == 1 and np.float32(g.reward) == np.float32(1.0) and not g.lost
    g.step(0)                            # (3,5) eat
    assert g.score == 2
    g.step(0)                            # (2,5) eat, list now empty -> food search fails -> error flag
    assert g.score == 3 and g.error == 1 and not g.lost
    assert (g.board == 2).sum() == 0
    g.step(0)                            # (1,5) wall
    assert g.lost and g.board[0, 4] == 1 and np.float32(g.reward) == np.float32(-1.0)


def test_r5_food_skips_occupied_entries_and_keeps_them():
    # second list entry is under the snake when the first apple is eaten -> skipped, kept for later
    g = O.OracleGame(food_rc=[(5, 5), (6, 5), (2, 2)])
    for a in [3, 3, 3, 0, 0, 0, 0]:
        g.step(a)                        # eats (4,5): list entry (5,5) is the body right below the head
    assert g.score == 1
    assert tuple(np.argwhere(g.board == 2)[0] + 1) == (2, 2) or tuple(np.argwhere(g.board == 2)[0] + 1) == (6, 5)
    # body is (4,5),(5,5),(6,5): both (5,5) and (6,5) are occupied -> (2,2) chosen, two entries remain
    assert tuple(np.argwhere(g.board == 2)[0] + 1) == (2, 2) and g.n_food == 2


def test_r8_mask_equals_closed_form_on_random_play():
    """virtual_step by deep copy == closed form (wall | body after the conditional tail pop | t >= 499).
    The snake is tracked here independently from the action stream and the pre-step boards."""
    rng = np.random.default_rng(3)
    n_checked = 0
    for ep in range(300):
        g = O.OracleGame()
        snake = [(7, 1), (8, 1)]                      # 0-based (row, col), head first
        t = 0
        while True:
            before = g.board
            av = g.available_actions()
            d = int(av[rng.integers(0, 3)])
            nh = (snake[0][0] + O_DIRS[d][0], snake[0][1] + O_DIRS[d][1])
            ate = before[nh] == 2
            snake.insert(0, nh)
            if not ate:
                snake.pop()
            g.step(d)
            t += 1
            if g.lost:
                break
            b = g.board
            assert sorted(map(tuple, np.argwhere(b == 1))) == sorted(snake)
            av2, lost = g.virtual_step()
            for k, d2 in enumerate(av2):
                r, c = snake[0][0] + O_DIRS[d2][0], snake[0][1] + O_DIRS[d2][1]
                if b[r, c] == -1:
                    want = 1
                elif b[r, c] == 2:
                    want = int(t >= 499)
                else:
                    want = int(t >= 499 or (r, c) in snake[:-1])
                assert lost[k] == want, (ep, t, k)
                n_checked += 1
    assert n_checked > 3000


O_DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1)]


def test_r10_state_layout_column_major_two_frames():
    g = O.OracleGame()
    g.step(0)
    s = g.next_state()
    older, newer = s[:100].reshape(10, 10).T, s[100:].reshape(10, 10).T
    assert older[7, 1] == 1 and older[8, 1] == 1 and older[6, 1] == 0
    assert newer[6, 1] == 1 and newer[7, 1] == 1 and newer[8, 1] == 0
    assert newer[3, 4] == 2 and (newer[0] == -1).all() and (newer[:, 9] == -1).all()
    # element (r,c,f) 1-based at (r-1)+10(c-1)+100(f-1)
    assert s[(7 - 1) + 10 * (2 - 1) + 100] == 1
    # assemble_state! of a live game is the same pair (utils.jl:136, not-lost branch)
    assert np.array_equal(g.assemble_state(), s)


def test_r11_reward_bit_patterns():
    g = O.OracleGame()
    g.step(0)
    assert np.float32(g.reward).view(np.uint32) == 0xBC23D70A
    g2 = O.OracleGame()
    g2.step(2)
    assert np.float32(g2.reward).view(np.uint32) == 0xBF800000


def test_r12_argmax_semantics():
    f = O.epsilon_greedy_idx
    assert f([1, 3, 2], 0.05, 0.9, 2) == 1
    assert f([3, 3, 2], 0.05, 0.9, 2) == 0                 # first maximum
    assert f([1, 3, 2], 0.05, 0.01, 2) == 2                # u < eps -> injected random index
    assert f([1, 3, 2], 0.05, 0.05, 2) == 1                # strict <
    assert f([1, np.nan, 2], 0.0, 0.5, 0) == 1             # NaN wins (Julia isless)
    assert f([-0.0, 0.0, -0.0], 0.0, 0.5, 0) == 1          # -0.0 < 0.0 under isless
    assert f([0.0, -0.0, 0.0], 0.0, 0.5, 0) == 0


def test_r13_target_float64_promotion():
    q = np.array([[0.5, 2.0, -1.0], [0.5, 2.0, -1.0], [3.0, 1.0, 2.0]], np.float32)
    mask = np.array([[0, 1, 0], [1, 1, 1], [0, 0, 0]], np.uint8)
    r = np.array([-0.01, 1.0, -1.0], np.float32)
    done = np.array([0, 0, 1], np.uint8)
    y = O.masked_target(q, mask, r, done)
    assert y.dtype == np.float64
    assert y[0] == np.float64(np.float32(-0.01)) + 0.97 * 0.5
    assert y[1] == 1.0 + 0.97 * -100.0                     # all-masked, non-terminal: r - 97
    assert y[2] == -1.0                                    # terminal
    assert y[0] != np.float64(np.float32(np.float32(-0.01) + np.float32(0.97) * np.float32(0.5)))


def test_welford_matches_numpy_and_is_centred():
    rng = np.random.default_rng(0)
    K, P = 58, 300
    D = rng.normal(0, 1, (K, P))
    flat = D.copy().reshape(-1)
    mean, var = O.center_columns(flat, P, K)
    assert np.allclose(mean, D.mean(0), rtol=0, atol=1e-14)
    assert np.allclose(var, D.var(0, ddof=1), rtol=1e-12)
    assert np.allclose(flat.reshape(K, P), D - D.mean(0), atol=1e-14)
    # K = 1: var = m2 / max(n-1, 1) = 0
    one = rng.normal(0, 1, (1, 7)).reshape(-1)
    m1, v1 = O.center_columns(one, 7, 1)
    assert (v1 == 0).all() and (one == 0).all()


def test_board_invariants_under_arbitrary_action_streams():
    """hypothesis-driven: for any stream of action indices the oracle keeps the board invariants of the reference
    (snake cells = score + 2 on live boards, at most one food, walls intact until a wall death overwrites one cell,
    rewards in {+1, -0.01, -1}, a lost game reports trues(3))."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=2), min_size=1, max_size=120), st.integers(0, 2 ** 16))
    def run(actions, seed):
        g = O.OracleGame()
        rng = np.random.default_rng(seed)
        for a in actions:
            av = g.available_actions()
            g.step(int(av[a]))
            b = g.board
            r = np.float32(g.reward)
            assert r in (np.float32(1.0), np.float32(-0.01), np.float32(-1.0))
            assert (b == 2).sum() <= 1
            border = np.ones((10, 10), bool); border[1:9, 1:9] = False
            if g.lost:
                assert r == np.float32(-1.0)
                assert (b[border] != -1).sum() <= 1
                assert g.virtual_step()[1].tolist() == [1, 1, 1]
                break
            assert (b[border] == -1).all()
            assert (b == 1).sum() == g.score + 2
            if rng.random() < 0.2:
                av2, lost = g.virtual_step()
                assert set(av2.tolist()) <= {0, 1, 2, 3} and len(set(av2.tolist())) == 3
    run()
