"""GPU parity: the CUDA engine (through the C ABI) must reproduce, bit for bit, the record stream, the SimPy step
count and the per-ring reductions of the golden vectors produced by the unmodified reference, when fed the
reference's own variate tapes (replay mode)."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu

POLICY = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}
BASE_NAMES = [n for n in GOLDEN_NAMES if not n.startswith("dbr")]


def _engine(scn, **kw):
    from manusim_b200 import BatchEngine, compile_scenario

    t = compile_scenario(scn["cfg"], scn["resources"], scn["products"], policy=POLICY[scn["policy"]],
                         dbr_repair="P3" in scn["patches"], **kw)
    return BatchEngine(t), t


@pytest.mark.parametrize("name", BASE_NAMES)
def test_replay_trace_bit_exact(name):
    import torch

    from manusim_b200.engine import decode_log

    g = load_golden(name)
    eng, t = _engine(g["scenario"])
    n = 5
    tapes = [{k: g["tape_" + k] for k in "ABCD"}] * n
    cap = len(g["ring"]) + 64
    res = eng.run(n, tapes=tapes, n_log_reps=n, log_cap=cap)
    torch.cuda.synchronize()
    status = res.status.cpu().numpy()
    assert (status == 0).all(), status
    assert (res.n_events.cpu().numpy() == g["steps"]).all(), (res.n_events.cpu().numpy(), g["steps"])
    counts = res.log_count.cpu().numpy()
    logs = res.log.cpu().numpy()
    for i in range(n):
        assert counts[i] == len(g["ring"])
        ring, tm, val, seq = decode_log(logs[i], int(counts[i]))
        assert np.array_equal(ring, g["ring"].astype(np.uint32))
        assert np.array_equal(tm.view(np.uint32), g["time"].view(np.uint32))
        assert np.array_equal(val.view(np.uint32), g["value"].view(np.uint32))
        assert np.array_equal(seq, np.arange(len(seq), dtype=np.uint32))
    # fused reductions == plain float64 sums / counts of the golden records
    K = eng.K
    s = np.zeros(K)
    c = np.zeros(K, dtype=np.int64)
    np.add.at(s, g["ring"].astype(np.int64), g["value"].astype(np.float64))
    np.add.at(c, g["ring"].astype(np.int64), 1)
    acc = res.acc_sum.cpu().numpy()
    cnt = res.acc_cnt.cpu().numpy()
    for i in range(n):
        np.testing.assert_allclose(acc[i], s, rtol=1e-12, atol=1e-9)
        assert np.array_equal(cnt[i].astype(np.int64), c)
