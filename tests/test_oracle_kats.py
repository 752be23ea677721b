"""The CPU oracle against the reference's own known-answer tests (Voxelyze2 gtest suite)."""
import pytest

import kats
import oracle_lib as O


@pytest.mark.parametrize("name,case", kats.all_kats(), ids=[n for n, _ in kats.all_kats()])
def test_oracle_reproduces_reference_kat(name, case):
    model, steps, check = case
    w = O.OracleWorld(model)
    assert w.run(steps) == steps
    check(w.state())


SCRIPTED = kats.scripted_kats()


@pytest.mark.parametrize("name,script", SCRIPTED, ids=[n for n, _ in SCRIPTED])
def test_oracle_reproduces_scripted_reference_kat(name, script):
    """KATs whose reference test reads the state every step or changes the load between steps (tVoxelyze.h freqency,
    largeDeformation, impulse, staticFriction)."""
    build, drive = script
    drive(O.OracleWorld(build()))
