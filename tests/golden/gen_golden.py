"""Generate tests/golden/*.json from the patched REFERENCE itself (oracle/_ref/libaz_ref.so).

Run in the authoring container only (needs /root/reference to have built oracle/_ref):
    make -C oracle ref && python tests/golden/gen_golden.py
The fixtures pin (a) the serial ParallelMCTS root statistics with the stateless HashEvaluator and
(b) the state API (legal-move order, terminal flag, result, 11/8-plane tensors) along seeded random
playouts.  Floats are stored as their IEEE-754 bit patterns (uint32) so equality is bit-exact.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _orc  # noqa: E402


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32).tolist()


def search_golden(R, game, n, sims, moves):
    s = R.new_state(game, n)
    m = R.mcts_new(s, sims, 1.5, 3, 0, None, None)
    out = []
    for mv in range(moves):
        R.mcts_search(m)
        st = R.root_stats(m)
        T = 1.0
        act = R.mcts_select_action(m, 1, T)
        out.append(dict(actions=st["actions"].tolist(), N=st["N"].tolist(), W=bits(st["W"]), P=bits(st["P"]),
                        rootN=int(st["rootN"]), rootW=bits([st["rootW"]])[0],
                        root_value=bits([R.mcts_root_value(m)])[0], action=int(act)))
        R.mcts_update_with_move(m, act)
        R.state_make_move(s, act)
        if R.state_is_terminal(s):
            break
    return dict(game=game, board=n, sims=sims, cpuct=1.5, virtual_loss=3, evals=int(R.mcts_eval_calls(m)), moves=out)


def playout_golden(R, game, n, seed, max_ply, pass_prob=0.03):
    rng = np.random.default_rng(seed)
    s = R.new_state(game, n)
    moves, digest = [], hashlib.sha256()
    per_ply = []
    ply = 0
    while True:
        legal = R.legal(s)
        term = R.state_is_terminal(s)
        res = R.state_result(s)
        t = R.tensor(s)
        h = hashlib.sha256()
        h.update(legal.astype(np.int32).tobytes()); h.update(bytes([term, res, R.state_current_player(s)]))
        h.update(t.astype(np.float32).tobytes())
        per_ply.append(h.hexdigest()[:16])
        if term or ply >= max_ply:
            break
        cand = legal
        if game == _orc.GO and len(legal) > 1 and rng.random() >= pass_prob:
            cand = legal[1:]
        a = int(rng.choice(cand))
        assert R.state_make_move(s, a) == 0
        moves.append(a); ply += 1
    return dict(game=game, board=n, seed=seed, moves=moves, ply_digest=per_ply,
                final_terminal=int(R.state_is_terminal(s)), final_result=int(R.state_result(s)))


def chess_golden(R):
    """Reference chess (recursion cut, oracle/build_ref.sh shim 6): digests of (legal order, terminal, result, player, isInCheck,
    18 planes) on every position of seeded random games and on the C5 / C7 FEN cases; hash-evaluator searches."""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_ref_chess_dataset as T
    plays = []
    for seed in range(8):
        rng = np.random.default_rng(1000 + seed)
        s = R.new_state(_orc.CHESS, 8)
        moves, dig = [], []
        for ply in range(450):
            dig.append(T._digest(R, s))
            if R.state_is_terminal(s):
                break
            lg = R.legal(s)
            caps = [a for a in lg if R.chess_piece(s, int(a) & 63) != 0]
            a = int(rng.choice(caps)) if (caps and rng.random() < 0.35) else int(rng.choice(lg))
            assert R.state_make_move(s, a) == 0
            moves.append(a)
        plays.append(dict(seed=1000 + seed, moves=moves, ply_digest=dig, final_result=int(R.state_result(s))))
    fens = {fen: T._digest(R, R.chess_from_fen(fen)) for fen in T.C5_FENS + [f for f, _ in T.C7_FENS]}
    searches = []
    for sims, n_moves, opening in [(150, 3, []), (100, 3, plays[0]["moves"][:16])]:
        s = R.new_state(_orc.CHESS, 8)
        for a in opening:
            assert R.state_make_move(s, a) == 0
        m = R.mcts_new(s, sims, 1.5, 3, 0, None, None)
        out = []
        for mv in range(n_moves):
            R.mcts_search(m)
            st = R.root_stats(m)
            act = R.mcts_select_action(m, 1, 1.0)
            out.append(dict(actions=st["actions"].tolist(), N=st["N"].tolist(), W=bits(st["W"]), P=bits(st["P"]), rootN=int(st["rootN"]), action=int(act)))
            R.mcts_update_with_move(m, act)
        searches.append(dict(sims=sims, opening=[int(a) for a in opening], moves=out))
    return dict(playouts=plays, fens=fens, search=searches)


def dataset_golden(R):
    """Reference Dataset::augmentExample / extractExamples: digests per image / per example list."""
    import test_ref_chess_dataset as T
    O = _orc.oracle()
    aug = []
    for seed, (c, n, p, game) in enumerate([(11, 15, 225, 0), (8, 9, 82, 2), (8, 19, 362, 2), (11, 15, 97, 0), (8, 9, 30, 2)]):
        pl, po = T._random_example(np.random.default_rng(seed), c, n, p)
        r_pl, r_po = T._ref_augment(R, pl, po, game)
        aug.append(dict(seed=seed, c=c, n=n, p=p, images=[hashlib.sha256(r_pl[k].tobytes() + r_po[k].tobytes()).hexdigest()[:16] for k in range(7)]))
    ext = []
    for i, (game, board, plen) in enumerate([(0, 9, 81), (2, 9, 82), (0, 15, 40), (1, 8, 37)]):
        rng = np.random.default_rng(50 + i)
        mv, res = T._random_game(O, rng, game, board, 12)
        res = res if res != 0 else 3
        pols = np.random.default_rng(70 + i).random((len(mv), plen)).astype(np.float32)
        for augm in (1, 0):
            k = 8 if (augm and game != 1) else 1
            c = O.state_tensor(O.new_state(game, board), None)
            n_ex = len(mv) * k
            r_pl = np.zeros((n_ex, c, board, board), np.float32); r_po = np.zeros((n_ex, plen), np.float32); r_va = np.zeros(n_ex, np.float32)
            assert R.dataset_extract(game, board, np.asarray(mv, np.int32).ctypes.data, len(mv), pols.ctypes.data, plen, res, augm, r_pl.ctypes.data, r_po.ctypes.data, r_va.ctypes.data) == n_ex
            ext.append(dict(game=game, board=board, plen=plen, seed=70 + i, moves=[int(a) for a in mv], result=int(res), augment=augm,
                            digest=hashlib.sha256("".join(T._canon(r_pl, r_po, r_va)).encode()).hexdigest()[:16]))
    return dict(augment=aug, extract=ext)


def main():
    R = _orc.reference()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    json.dump(chess_golden(R), open(os.path.join(HERE, "chess_reference.json"), "w"))
    json.dump(dataset_golden(R), open(os.path.join(HERE, "dataset_reference.json"), "w"))
    if "--chess-dataset-only" in sys.argv:
        return
    searches = [search_golden(R, _orc.GOMOKU, 15, 800, 6), search_golden(R, _orc.GO, 9, 400, 4),
                search_golden(R, _orc.GOMOKU, 9, 200, 30)]
    json.dump(searches, open(os.path.join(HERE, "search_hash_eval.json"), "w"))
    plays = [playout_golden(R, _orc.GOMOKU, 15, sd, 400) for sd in range(6)]
    plays += [playout_golden(R, _orc.GOMOKU, 9, 100 + sd, 400) for sd in range(4)]
    plays += [playout_golden(R, _orc.GO, 9, 200 + sd, 300) for sd in range(6)]
    plays += [playout_golden(R, _orc.GO, 19, 300, 250)]
    json.dump(plays, open(os.path.join(HERE, "state_playouts.json"), "w"))
    # first-fill legal order of fresh Gomoku boards (QUIRK G2)
    ff = {}
    for n in (9, 15, 19):
        s = R.new_state(_orc.GOMOKU, n)
        ff[str(n)] = R.legal(s).tolist()
    json.dump(ff, open(os.path.join(HERE, "gomoku_first_fill_order.json"), "w"))
    print("golden written:", [len(x["moves"]) for x in searches], len(plays))


if __name__ == "__main__":
    main()
