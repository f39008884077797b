import sys, numpy as np, torch
sys.path.insert(0, '.')
from manusim_b200 import BatchEngine, compile_scenario, scenarios
from manusim_b200.engine import decode_log
from oracle.factory_oracle import run_oracle, metrics_from_trace
sim, res, prod, _ = scenarios.load_example("toc_dbr")
ru, wu = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(sim, run_until=ru, warmup=wu); cfg.pop("print_mode", None)
t = compile_scenario(cfg, res, prod, policy="dbr", dbr_repair=True)
eng = BatchEngine(t)
n = 2
r = eng.run(n, seed=5, n_log_reps=n, log_cap=1 << 21, record_tapes=1 << 20)
torch.cuda.synchronize()
cnt = r.rec["count"].cpu().numpy()
print("status", r.status.cpu().numpy(), "counts", cnt, "events", r.n_events.cpu().numpy())
for i in range(n):
    tp = {k: r.rec[k][i, : cnt[i, j]].cpu().numpy() for j, k in enumerate("ABCD")}
    o = run_oracle(cfg, res, prod, None, "dbr", tapes=tp, dbr_repair=True)
    ring, tm, val, _ = decode_log(r.log[i].cpu().numpy(), int(r.log_count[i]))
    m = min(len(ring), len(o["ring"]))
    same = (ring[:m] == o["ring"][:m].astype(np.uint32)) & (tm[:m].view(np.uint32) == o["time"][:m].view(np.uint32)) & (val[:m].view(np.uint32) == o["value"][:m].view(np.uint32))
    first = int(np.argmin(same)) if not same.all() else -1
    print(i, "gpu events", int(r.n_events[i]), "oracle steps", o["steps"], "records", len(ring), len(o["ring"]), "first mismatch", first, o["error"])
    if first >= 0:
        print(" gpu   ", ring[first-2:first+3], tm[first-2:first+3], val[first-2:first+3])
        print(" oracle", o["ring"][first-2:first+3], o["time"][first-2:first+3], o["value"][first-2:first+3])
    s, c, v = metrics_from_trace(o, 10, 13, ru, wu)
    print("  oracle util", np.round(v[110:123], 3), "lost", v[20:30].sum())
