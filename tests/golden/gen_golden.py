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


def main():
    R = _orc.reference()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
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
