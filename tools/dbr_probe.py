import sys, numpy as np, torch
sys.path.insert(0, '.')
from manusim_b200 import BatchEngine, compile_scenario, scenarios
from manusim_b200.metrics import metric_values
sim, res, prod, _ = scenarios.load_example("toc_dbr")
for ru, wu in ((20001, 2000), (60001, 30000), (120001, 60000), (200001, 100000), (200001, 1000)):
    cfg = dict(sim, run_until=ru, warmup=wu); cfg.pop("print_mode", None)
    t = compile_scenario(cfg, res, prod, policy="dbr", dbr_repair=True)
    eng = BatchEngine(t)
    r = eng.run(512, seed=5)
    torch.cuda.synchronize()
    v = metric_values(r.acc_sum.cpu().numpy(), r.acc_cnt.cpu().numpy(), t, float(ru - wu))
    print(ru, wu, "status", np.bincount(r.status.cpu().numpy(), minlength=9)[:6], "events/rep", float(r.n_events.double().mean()),
          "ev/time", float(r.n_events.double().mean()) / ru, "util", np.round(np.nanmean(v[:, 110:123], axis=0), 3),
          "lost", float(np.nanmean(v[:, 20:30].sum(axis=1))), "ontime", float(np.nanmean(v[:, 0:10].sum(axis=1))),
          "bd", np.round(np.nanmean(v[:, 123:136], axis=0), 4)[:4])
