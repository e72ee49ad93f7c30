"""GPU parity tests for chess (BASELINE.json configs[4]): device move generation in the reference's
order, legality (incl. the literal isSquareAttacked quirk C5), castling / en passant / promotion, terminal rules (mate,
stalemate, material, fifty-move, placement-only threefold), the 18 feature planes, and the batched search with the hash
evaluator — against the oracle's chess restatement, which is pinned to the reference's own chess (oracle/_ref, recursion cut by
shim 6) position by position and search by search in tests/test_ref_chess_dataset.py.  Bar: bit-exact."""
import numpy as np
import pytest

import _orc
from _orc import CHESS

pytestmark = pytest.mark.gpu


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def code(frm, to, promo=0):
    return (promo << 12) | (frm << 6) | to


def chess_engine(n_slots, sims=100, **kw):
    from _eng import E
    cfg = dict(game=E.CHESS, board_size=8, n_slots=n_slots, num_simulations=sims, evaluator=E.EVAL_HASH, deterministic=1,
               auto_restart=0, max_nodes_per_tree=(sims + 2) * 256 + 1, n_streams=1)
    cfg.update(kw)
    return E.Engine(**cfg)


def _random_games(O, n_games, max_plies, seed):
    rng = np.random.default_rng(seed)
    games = []
    for g in range(n_games):
        s = O.new_state(CHESS, 8)
        mv = []
        for _ in range(max_plies):
            lg = O.legal(s)
            if len(lg) == 0 or O.state_is_terminal(s):
                break
            caps = [a for a in lg if O.chess_piece(s, int(a) & 63) != 0]          # prefer captures now and then: shorter, sharper games
            a = int(rng.choice(caps)) if (caps and rng.random() < 0.3) else int(rng.choice(lg))
            assert O.state_make_move(s, a) == 0
            mv.append(a)
        games.append(mv)
    return games


SCRIPTED = [
    [code(53, 45), code(12, 28), code(54, 38), code(3, 39)],                                             # fool's mate
    [code(52, 36), code(12, 28), code(62, 45), code(1, 18), code(61, 34), code(5, 26), code(60, 62), code(6, 21), code(59, 52), code(4, 6)],   # both sides castle short
    [code(52, 36), code(8, 16), code(36, 28), code(11, 27), code(28, 19)],                               # e4 a6 e5 d5 exd6 e.p.
    [code(62, 45), code(6, 21), code(45, 62), code(21, 6)] * 2,                                          # threefold by placement
    [code(48, 32), code(9, 25), code(32, 25), code(8, 16), code(25, 16), code(1, 18), code(16, 8), code(0, 1), code(8, 0, 1)],     # the a-pawn walks to a8 and promotes to a queen
]


def test_chess_rules_kernels_match_oracle():
    O = _orc.oracle()
    eng = chess_engine(2, sims=4)
    games, expect = [], []
    for mv in _random_games(O, 10, 160, seed=3) + SCRIPTED:
        s = O.new_state(CHESS, 8)
        for ply in range(len(mv) + 1):
            games.append(mv[:ply])
            expect.append((O.legal(s), O.state_is_terminal(s), O.state_result(s), O.state_current_player(s), O.tensor(s)))
            if ply < len(mv):
                assert O.state_make_move(s, mv[ply]) == 0, (mv, ply)
    assert len(games) > 500
    r = eng.rules_replay(games)
    n_term = 0
    for i, (legal, term, res, pl, planes) in enumerate(expect):
        assert r["n_legal"][i] >= 0, (i, games[i])
        assert np.array_equal(r["legal"][i], legal), (i, len(games[i]))
        assert r["terminal"][i] == term and r["result"][i] == res and r["player"][i] == pl, (i, len(games[i]))
        assert np.array_equal(r["planes"][i], planes), (i, len(games[i]))
        n_term += int(term)
    assert n_term >= 2
    # illegal moves are reported, not applied (the reference's makeMove throws): pawn moving backwards, moving into check, out of range
    bad = eng.rules_replay([[code(48, 56)], [code(52, 36), code(12, 28), code(60, 52), code(3, 39), code(53, 45)], [20480], [-1]])
    exp = []
    for c in ([code(48, 56)], [code(52, 36), code(12, 28), code(60, 52), code(3, 39), code(53, 45)], [20480], [-1]):
        s = O.new_state(CHESS, 8)
        exp.append(any(O.state_make_move(s, a) != 0 for a in c))
    assert [int(n) == -1 for n in bad["n_legal"]] == exp and exp[0] and exp[2] and exp[3]
    eng.close()


def test_chess_search_matches_oracle():
    """Batched search on chess positions (start position, an open middlegame, a position with castling / e.p. rights) against
    the oracle's serial search with the same hash evaluator over the 20480-entry action space: child order, visit counts,
    valueSum / prior bits, then play on for a few moves (subtree reuse)."""
    O = _orc.oracle()
    sims = 120
    # SCRIPTED[1][:6], SCRIPTED[4][:6] and the seed-11 game contain QUIRK C8 events within these searches (a leaf whose piece placement was
    # evaluated before with other castling rights / side to move gets that position's cached evaluation from the reference's
    # TranspositionTable, tests/test_ref_chess_dataset.py): the engine's table model (tree.cuh EvalTT) must reproduce them
    roots = [[], SCRIPTED[1][:6], SCRIPTED[2][:4], SCRIPTED[4][:6], _random_games(O, 1, 40, seed=9)[0], _random_games(O, 1, 40, seed=11)[0]]
    eng = chess_engine(len(roots), sims=sims)
    searches = []
    for t, mv in enumerate(roots):
        s = O.new_state(CHESS, 8)
        for a in mv:
            assert O.state_make_move(s, a) == 0
        if O.state_is_terminal(s):
            s = O.new_state(CHESS, 8); mv = []
        eng.set_root(t, mv)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 0, None, None))
    for move in range(3):
        eng.search()
        acts = []
        for t in range(len(roots)):
            O.mcts_search(searches[t])
            a, b = eng.root_stats(t), O.root_stats(searches[t])
            assert np.array_equal(a["actions"], b["actions"]), (move, t)
            assert np.array_equal(a["N"], b["N"]), (move, t)
            assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (move, t)
            assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0]
            act = O.mcts_select_action(searches[t], 1, 1.0)
            acts.append(act)
            O.mcts_update_with_move(searches[t], act)
        eng.advance(acts)
    assert eng.stats()["pool_overflows"] == 0
    eng.close()


def test_chess_selfplay_auto_restart_and_noise_smoke():
    """Throughput-mode loop on chess (Dirichlet noise, temperature sampling, auto-restart): invariants only.  Samples carry
    (action, count) pairs in child order."""
    sims, T = 24, 32
    eng = chess_engine(T, sims=sims, deterministic=0, auto_restart=1, seed=5, sample_ring_capacity=T * 600)
    for _ in range(40):
        eng.play(4)
    smp = eng.drain_samples(cap=T * 600)
    st = eng.stats()
    assert st["moves"] == 160 * T and st["pool_overflows"] == 0 and st["samples_dropped"] == 0
    if len(smp):
        pairs = smp["visits"].reshape(len(smp), -1, 2)
        assert np.all(pairs[:, :, 1].sum(1) >= sims - 1)
        assert np.all(np.abs(smp["z"]) <= 1) and np.all(smp["result"] >= 1)
        played = [(pairs[i, :, 0] == smp["action"][i]) & (pairs[i, :, 1] >= 1) for i in range(len(smp))]
        assert all(p.any() for p in played)
    eng.close()


def test_chess_selfplay_with_resnet_evaluator():
    """Chess self-play with the tcgen05 ResNet evaluator in the loop (18 planes -> 32-channel stem, 20480-wide policy head,
    random-init 1-block net): the wave pipeline runs, simulations are accounted for, the root priors match the fp32 network."""
    import torch
    from _eng import E, N
    O = _orc.oracle()
    sims, T = 16, 16
    model = N.make_random_model(seed=6, blocks=1, in_planes=18, board=8, actions=20480)
    with torch.no_grad():
        model.p_fc.weight *= 0.2; model.v_fc1.weight *= 0.2
    eng = E.Engine(game=E.CHESS, board_size=8, n_slots=T, evaluator=E.EVAL_RESNET, net_blocks=1, num_simulations=sims, deterministic=1,
                   auto_restart=1, max_nodes_per_tree=(sims + 2) * 256 + 1)
    eng.load_weights(N.export_weights(model))
    eng.search()
    st = eng.root_stats(3)
    s0 = O.new_state(CHESS, 8)
    legal = O.legal(s0)
    assert st["actions"].tolist() == legal.tolist() and int(st["N"].sum()) == sims
    with torch.no_grad():
        p32 = torch.softmax(model(torch.tensor(O.tensor(s0)[None]))[0], 1).numpy()[0]
    ref = p32[legal] / p32[legal].sum()
    assert np.abs(st["P"] - ref).max() <= 2e-3
    eng.play(2)
    s = eng.stats()
    assert s["moves"] == 2 * T and s["simulations"] == 3 * sims * T and s["pool_overflows"] == 0
    eng.close()


def test_chess_search_without_table_model_matches_tt_free_oracle():
    """tt_entries = -1 switches the TranspositionTable model off: every leaf is evaluated on its own input.  On a root with a QUIRK C8
    event the engine then equals the oracle's TT-free search and differs from the reference-faithful one exactly where the oracle says."""
    O = _orc.oracle()
    sims, mv = 120, SCRIPTED[4][:6]
    eng = chess_engine(1, sims=sims, tt_entries=-1)
    s = O.new_state(CHESS, 8)
    for a in mv:
        assert O.state_make_move(s, a) == 0
    eng.set_root(0, mv)
    on, off = O.mcts_new(s, sims, 1.5, 3, 0, None, None), O.mcts_new(s, sims, 1.5, 3, 0, None, None)
    O.mcts_set_tt(off, 0)
    differed = False
    for move in range(3):
        eng.search(); O.mcts_search(on); O.mcts_search(off)
        a, b, c = eng.root_stats(0), O.root_stats(off), O.root_stats(on)
        assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]) and np.array_equal(bits(a["W"]), bits(b["W"])), move
        differed |= not (np.array_equal(b["N"], c["N"]) and np.array_equal(bits(b["W"]), bits(c["W"])))
        act = O.mcts_select_action(off, 1, 1.0)
        if O.mcts_select_action(on, 1, 1.0) != act:
            differed = True
            break
        O.mcts_update_with_move(off, act); O.mcts_update_with_move(on, act); eng.advance([act])
    assert differed
    eng.close()


def test_chess_legal_only_policy_equals_dense_softmax():
    """Inside the waves the chess network computes the logits of the leaf's LEGAL moves only (one 2048-long dot product per legal move
    instead of the 20480 x 2048 policy FC + a 20480-wide softmax): softmax over all A followed by the expansion's renormalisation over
    the legal moves is the softmax over the legal logits.  Root priors of the legal-only engine vs (a) the dense engine
    (dense_policy = 1: all logits, full softmax, renormalised at expansion) and (b) the fp32 PyTorch network: KL <= 1e-3 (north-star
    tolerance), in practice ~1e-6."""
    import torch
    from _eng import E, N
    O = _orc.oracle()
    m = N.make_random_model(seed=2, randomize_bn=True, blocks=3, in_planes=18, board=8, actions=20480)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        m.p_fc.weight *= 0.2; m.v_fc1.weight *= 0.05; m.v_fc2.weight *= 0.2
        m.p_fc.bias.copy_(torch.rand(m.p_fc.bias.shape, generator=gen) - 0.5)
    blob = N.export_weights(m)
    games = _random_games(O, 6, 60, seed=21)
    roots = [[]] + [g[:k] for g, k in zip(games, (5, 12, 20, 33, 41, 58))]
    stats = []
    for dense in (0, 1):
        eng = E.Engine(game=E.CHESS, board_size=8, n_slots=len(roots), evaluator=E.EVAL_RESNET, net_blocks=3, num_simulations=4, deterministic=1,
                       auto_restart=0, max_nodes_per_tree=4096, dense_policy=dense)
        eng.load_weights(blob)
        for t, mv in enumerate(roots):
            eng.set_root(t, mv)
        eng.search(4)
        stats.append([eng.root_stats(t) for t in range(len(roots))])
        eng.close()
    for t, mv in enumerate(roots):
        a, b = stats[0][t], stats[1][t]
        assert np.array_equal(a["actions"], b["actions"]) and len(a["actions"]) > 0
        assert np.allclose(a["P"], b["P"], rtol=2e-3, atol=1e-6), (t, np.abs(a["P"] - b["P"]).max())
        s = O.new_state(CHESS, 8)
        for x in mv:
            assert O.state_make_move(s, x) == 0
        with torch.no_grad():
            lg, _ = m(torch.tensor(O.tensor(s)[None]))
        ref = torch.softmax(lg[0, torch.tensor(a["actions"].astype(np.int64))], 0).numpy()
        kl = float((ref * (np.log(ref + 1e-30) - np.log(a["P"] + 1e-30))).sum())
        assert kl <= 1e-3, (t, kl)


def test_chess_full_size_1024_slots_identical_and_oracle():
    """BASELINE.json configs[4] at full size (chess, 1024 slots, 800 simulations), hash evaluator + the reference-table model, deterministic mode:
    the slots cycle through four roots (start position, castling rights on both sides, an e.p. square, a random middlegame); every slot of a root
    must hold the same tree — wherever it sits in the 1024-wide waves — and that tree must equal the oracle's serial 800-simulation search
    bit for bit (child order, visit counts, valueSum / prior bits), over two moves with subtree reuse."""
    O = _orc.oracle()
    sims, T = 800, 1024
    roots = [[], SCRIPTED[1][:8], SCRIPTED[2][:4], _random_games(O, 1, 40, seed=9)[0]]
    R = len(roots)
    eng = chess_engine(T, sims=sims)
    searches = []
    for r, mv in enumerate(roots):
        s = O.new_state(CHESS, 8)
        for a in mv:
            assert O.state_make_move(s, a) == 0
        assert not O.state_is_terminal(s)
        searches.append(O.mcts_new(s, sims, 1.5, 3, 0, None, None))
    for t in range(T):
        eng.set_root(t, roots[t % R])
    for move in range(2):
        eng.search()
        acts = []
        for r in range(R):
            O.mcts_search(searches[r])
            b = O.root_stats(searches[r])
            for slot in (r, r + R * 1, r + R * 127, r + R * 128, r + R * 255):
                a = eng.root_stats(slot)
                assert np.array_equal(a["actions"], b["actions"]) and np.array_equal(a["N"], b["N"]), (move, r, slot)
                assert np.array_equal(bits(a["W"]), bits(b["W"])) and np.array_equal(bits(a["P"]), bits(b["P"])), (move, r, slot)
                assert a["rootN"] == b["rootN"] and bits([a["rootW"]])[0] == bits([b["rootW"]])[0], (move, r, slot)
            acts.append(O.mcts_select_action(searches[r], 1, 1.0))
            O.mcts_update_with_move(searches[r], acts[-1])
        first = [eng.root_stats(r) for r in range(R)]
        for t in range(R, T):                      # every slot, not only the sampled ones: same visit counts and valueSum bits as its root's first slot
            a = eng.root_stats(t)
            assert np.array_equal(a["N"], first[t % R]["N"]) and np.array_equal(bits(a["W"]), bits(first[t % R]["W"])), (move, t)
        eng.advance([acts[t % R] for t in range(T)])
    st = eng.stats()
    assert st["simulations"] == 2 * sims * T and st["pool_overflows"] == 0, st
    eng.close()
