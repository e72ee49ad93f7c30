"""CPU suite: the reference's OWN unit-test sources for the state classes — tests/games/{gomoku,go,chess}/*_state_test.cpp,
tests/core/{igamestate,game_factory}_test.cpp, 48 cases — compiled UNMODIFIED (read in place from /root/reference, never copied) twice:
against the reference build (oracle/build_ref.sh -> oracle/_ref/ref_state_tests) and against the B200 host mirror's classes of the same names
(oracle/run_mirror_unit_tests.sh -> oracle/_ref/mirror_state_tests; forwarding headers in oracle/mirror_shim_include, stand-in gtest in
oracle/mini_gtest — GoogleTest is not in this image).  Source-level drop-in check for boundary #2: every case the reference passes
on its own code passes on the mirror; the only cases the mirror does not pass are the variant-rule ones (Renju / Omok / Chess960), which crash or
fail on the reference itself."""
import os
import subprocess

import pytest

import _orc

ROOT = _orc.ROOT
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_state_tests")
MIR_BIN = os.path.join(ROOT, "oracle", "_ref", "mirror_state_tests")
HAVE_REF_TREE = os.path.isdir("/root/reference/tests/games")

VARIANT_CASES = {"GomokuStateTest.RenjuRules", "ChessStateTest.Chess960Mode", "GameFactoryTest.CreateGomokuState", "GameFactoryTest.CreateChessState",
                 "GameFactoryTest.CreateGameState"}


def _verdicts(binary, script):
    if not os.path.exists(binary):
        if not HAVE_REF_TREE:
            pytest.skip("reference test sources not present and no prebuilt binary")
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", script)], stdout=subprocess.DEVNULL)
    out = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    v = {}
    for line in out.stdout.splitlines():
        name, verdict = line.split()[:2]
        if name != "summary":
            v[name] = verdict
    return v


def test_reference_unit_tests_on_the_reference_and_on_the_mirror():
    ref = _verdicts(REF_BIN, "build_ref.sh")
    mir = _verdicts(MIR_BIN, "run_mirror_unit_tests.sh")
    assert len(ref) == 48 and set(ref) == set(mir)
    # what the reference does with its own tests at HEAD: 33 pass; the Renju and Chess960 cases kill the process (stack overflow / failed
    # assert); the IGameState / GameFactory suites fail because no game is ever registered in its GameRegistry
    assert sorted(n for n, x in ref.items() if x == "CRASH") == ["ChessStateTest.Chess960Mode", "GomokuStateTest.RenjuRules"]
    assert sum(x == "PASS" for x in ref.values()) == 33
    # drop-in: nothing the reference passes is lost
    lost = [n for n, x in ref.items() if x == "PASS" and mir[n] != "PASS"]
    assert lost == [], lost
    # and the mirror's only non-passing cases are the variant ones, none of which passes (or survives) on the reference
    not_passing = {n for n, x in mir.items() if x != "PASS"}
    assert not_passing == VARIANT_CASES, not_passing
    assert all(ref[n] != "PASS" for n in VARIANT_CASES)
    assert "CRASH" not in mir.values()
    assert sum(x == "PASS" for x in mir.values()) == 43
