"""CPU: oracle/variates.py against known-answer vectors generated with the reference's own DistributionGenerator
(engine/utils.py) by oracle/gen_golden.py, and the SURVEY Appendix F.3 / 8(c) values."""
import json

import numpy as np

from conftest import GOLDEN
from oracle.variates import ReferenceVariates, derived_params, run_seeds


def test_scalar_and_batch_kat():
    kat = json.loads((GOLDEN / "variates_kat.json").read_text())
    gens = {}
    for row in kat["scalar"]:
        g = gens.setdefault(("s", row["seed"]), ReferenceVariates(row["seed"]))
        v = g.scalar(row["dist"], row["params"])
        assert float(v) == row["value"]
        assert isinstance(v, int) == row["is_int"]
    for row in kat["batch"]:
        g = gens.setdefault(("b", row["seed"]), ReferenceVariates(row["seed"]))
        assert g.batch_sum(row["dist"], row["params"], row["count"]) == row["value"]
    for seed, vals in kat["run_seeds"].items():
        assert run_seeds(int(seed), len(vals)) == vals


def test_survey_known_answers():
    assert run_seeds(123, 5) == [54907, 280679, 91421, 806309, 427023]
    g = ReferenceVariates(42)
    got = [g.scalar("erlang", [10, 5.77, 3]), g.scalar("uniform", [10, 5.77, 0]), g.scalar("expo", [4635]),
           g.scalar("normal", [5.24, 2.97]), g.scalar("gamma", [0.78, 0.45]), g.scalar("constant", [63])]
    want = [6.4846463, 14.726555, 5233.717, 4.8299646, 0.441518, 63.0]
    assert all(isinstance(v, np.float32) for v in got)
    np.testing.assert_allclose(np.array(got, dtype=np.float64), want, rtol=1e-7)
    assert derived_params("uniform", [10, 5.77]) == (0.00606684032757876, 19.993933159672423)
    assert derived_params("gamma", [10, 5.77]) == (3.003643419467815, 3.3292899999999994)
    assert derived_params("gamma", [0.78, 0.45]) == (3.0044444444444447, 0.25961538461538464)


def test_float32_round6_identity():
    """round(np.float32 x, 6) == rint(x*1e6f)/1e6f in float32 -- what the kernel's round6 does."""
    rng = np.random.default_rng(5)
    x = rng.gamma(2.0, 3.0, size=20000).astype(np.float32)
    a = np.array([round(v, 6) for v in x], dtype=np.float32)
    b = (np.rint(x * np.float32(1e6)) / np.float32(1e6)).astype(np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
