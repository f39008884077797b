import sys, numpy as np, torch
sys.path.insert(0, '.')
from manusim_b200 import BatchEngine, compile_scenario, scenarios
for mode, mdo, mpo in (("asReady", 96, 96), ("onDue", 160, 96), ("onDue", 250, 250)):
    cfg, res, prod = scenarios.jobshop20(mode=mode)
    eng = BatchEngine(compile_scenario(cfg, res, prod, max_demand_orders=mdo, max_production_orders=mpo))
    r = eng.run(8192, seed=20260104)
    torch.cuda.synchronize()
    st = r.status.cpu().numpy()
    print(mode, mdo, mpo, eng.info, np.bincount(st, minlength=9), float(r.n_events.double().mean()))
