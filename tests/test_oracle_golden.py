"""CPU: the numpy oracle (oracle/pdas_oracle.py) against the golden vectors produced by the real reference."""
import numpy as np
import pytest

from oracle import pdas_oracle as orc
from tests.helpers import RTOL, assert_same_support, golden_names, load_golden, rel_err


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    seq = np.arange(1, g["smax"] + 1)
    out = orc.bess_cpp(g["x"], g["y"], g["data_type"], g["weight"], True, g["model_type"], 20, g["path_type"], True,
                       g["ic_type"], g["is_cv"], g["K"], seq, 1, g["smax"], g["scr"] > 0, max(g["scr"], 1),
                       fold_of_row=g["fold_of_row"], lambda_seq=g["lambda_seq"])
    assert_same_support(out["beta"], g["beta"])
    assert rel_err(out["beta"], g["beta"]) < RTOL
    assert abs(out["coef0"] - g["coef0"]) <= RTOL * max(1.0, abs(g["coef0"]))
    assert abs(out["train_loss"] - g["train_loss"]) <= RTOL * abs(g["train_loss"])
    assert abs(out["ic"] - g["ic"]) <= RTOL * abs(g["ic"])
    if "screening_A" in g:
        assert out["screening_A"].tolist() == g["screening_A"].tolist()
    if "beta_all" in g:  # per-level trace of the sequential path
        for lvl in range(len(seq)):
            assert_same_support(out["beta_all"][lvl], g["beta_all"][lvl])
        assert rel_err(out["beta_all"], g["beta_all"]) < RTOL
        assert rel_err(out["ic_all"], g["ic_all"]) < RTOL
        assert rel_err(out["loss_all"], g["loss_all"]) < RTOL
        assert out["l_all"].tolist() == g["l_all"].tolist()
    assert out["min_gap"] > 1e-9, "a top-k decision sits inside rounding noise"
