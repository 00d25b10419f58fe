"""The oracle reproduces the committed golden fixtures (tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from tests.common import EXACT_KEYS, GOLDEN, SOLVE_KEYS, STATE_KEYS, rel_err, run_sequence
from tests.golden.make_golden import CASES
from oracle.pyoracle import Oracle


def test_fixture_set_is_complete():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    # the HB fixture and the two on the reference's own meshes have their own generators (make_golden_hb.py, make_golden_tutorials.py)
    assert have == set(CASES) | {"hb_box_roe_3instants", "forwardstep_c2", "vki_c5"}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    case = CASES[name]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    o = case.apply(Oracle())
    out = run_sequence(o, case)
    for k in EXACT_KEYS:
        assert rel_err(out[k], gold[k]) <= 1e-13, k
    for k in SOLVE_KEYS + STATE_KEYS:
        assert rel_err(out[k], gold[k]) <= 1e-10, k
    assert out["restarts"][0] == gold["restarts"][0]
    assert np.array_equal(out["history"][:, -1], gold["history"][:, -1])
    assert rel_err(out["history"][:, :5], gold["history"][:, :5]) <= 1e-9
    o.close()
