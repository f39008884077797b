"""factory_oracle -- CPU restatement of the reference factory model (TEST INFRASTRUCTURE ONLY).

Nothing under ``manusim_b200/`` imports this module.  It is the *checker* for the CUDA engine:
``tests/`` compare the device record stream against it, ``__graft_entry__.smoke()`` checks one small
run against it and ``bench.py`` times it as the ``cpu_baseline`` ("port") / ``--impl reference`` arm.

It restates, process by process, the SimPy model of /root/reference/src/manusim/factory_sim.py on
top of ``oracle/minisimpy.py`` (third-party SimPy 4.1.1 restated -- PARITY UNPINNED at that layer, see
its header).  Entities are integer-indexed (config order) and records go to one ordered stream
``(ring, float32 time, float32 value)`` instead of the reference's per-key ``Logger`` arrays; ring ids
follow ``oracle.run_reference.ring_index``.

The restatement itself IS pinned: ``tests/test_oracle_golden.py`` checks it record-for-record, tape-for-tape
and step-count-for-step-count against golden vectors produced by executing the unmodified reference
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).

Reference anchors (file:line under /root/reference/src/manusim unless noted):
  construction order / initial eids ........ factory_sim.py:61-89 (+ _start_resources :244-265, _start_products :139-142)
  _warmup_period ........................... factory_sim.py:94-101
  _generate_demand_orders .................. factory_sim.py:144-175
  scheduler (base) ......................... factory_sim.py:180-193
  _release_order ........................... factory_sim.py:195-242
  _breakdowns / _get_breakdown_time ........ factory_sim.py:267-295
  _transportation .......................... factory_sim.py:297-358
  _production_system ....................... factory_sim.py:372-461
  order_selection (FIFO) ................... factory_sim.py:463-465
  order_selection (max priority) ........... examples/simple_scheduler/simulation.py:31-44
  _delivery_orders + 3 modes ............... factory_sim.py:467-549   (onDue with oracle patch P1)
  _log_inventory_metrics ................... factory_sim.py:551-601
  _log_delivery_performance / _log_lead_time factory_sim.py:606-654
  DBR scheduler / process_order / dispatch . examples/toc_dbr/simulation.py:200-337 (patch P2; optional repair P3)
"""
from __future__ import annotations

import numpy as np

from oracle.simpy_select import pick as _pick_simpy

simpy, SIMPY_ENGINE = _pick_simpy()      # minisimpy unless ORACLE_SIMPY=real|auto finds the real package
from oracle.variates import ReferenceVariates, TapeSource

PM = ["deliveredOntime", "deliveredLate", "lostSales", "flowTime", "leadTime", "tardiness",
      "earliness", "wip", "finishedGoods", "totalInventory", "released"]
RM = ["utilization", "breakdown", "setup", "queue"]
GM = ["wip_general", "finishedGoods_general", "totalInventory_general"]


class _PO:
    __slots__ = ("prod", "qty", "schedule", "released", "due", "prio", "total", "done", "uid")
    _uid = 0

    def __init__(self, prod, qty, prio=None):
        self.prod = prod
        self.qty = qty
        self.schedule = None
        self.released = None
        self.due = None
        self.prio = prio
        self.total = None
        self.done = None
        _PO._uid += 1
        self.uid = _PO._uid


class _DO:
    __slots__ = ("prod", "qty", "due", "arrived", "delivered")

    def __init__(self, prod, qty, due, arrived):
        self.prod = prod
        self.qty = qty
        self.due = due
        self.arrived = arrived
        self.delivered = None


class FactoryOracle:
    """One replication.  ``policy`` in {"fifo", "max_priority", "dbr", "edd"}; ``variates`` is None (reference
    generators seeded with ``seed``) or a dict of ``TapeSource`` objects {"A","B","D"} (B carries tape C too)."""

    def __init__(self, config, resources, products, seed=None, policy="fifo", variates=None,
                 dbr_repair=False):
        self.cfg = config
        self.res_cfg = resources
        self.prod_cfg = products
        self.seed = seed
        self.policy = policy
        self.dbr_repair = dbr_repair
        self.rnames = list(resources.keys())
        self.pnames = list(products.keys())
        self.R = len(self.rnames)
        self.P = len(self.pnames)
        self.run_until = config.get("run_until", None)
        if self.run_until is None:
            raise ValueError("run_until must be specified")
        self.warmup = config.get("warmup", 0)
        self.mode = config.get("delivery_mode", "asReady")
        self.log_queues = config.get("log_queues", False)
        self.logging = config.get("enable_event_logging", True)
        # routing tables (stores.py:22-30): insertion order of `processes`
        self.route = []
        for p in self.pnames:
            steps = []
            for st in products[p]["processes"].values():
                pt = st["processing_time"]
                steps.append((self.rnames.index(st["resource"]), pt.get("dist"), pt.get("params")))
            self.route.append(steps)
        self._build(variates)

    # ------------------------------------------------------------------ construction (factory_sim.py:61-89)
    def _build(self, variates):
        env = self.env = simpy.Environment()
        R, P = self.R, self.P
        self.q_in = [simpy.FilterStore(env) for _ in range(R)]
        self.q_out = [simpy.FilterStore(env) for _ in range(R)]
        self.q_proc = [simpy.Store(env) for _ in range(R)]
        self.q_trans = [simpy.Store(env) for _ in range(R)]
        self.q_fin = [simpy.Store(env) for _ in range(R)]
        self.inbound = simpy.FilterStore(env)
        self.fg = [simpy.Container(env) for _ in range(P)]
        self.outbound = [simpy.Store(env) for _ in range(P)]
        self.wip = [simpy.Container(env) for _ in range(P)]
        self.tot_wip = 0
        self.fg_cache = [0] * P
        self.qq = [0] * R
        self.last_queue = [None] * R
        self.trace = []
        self.tapes = {"A": [], "B": [], "C": [], "D": []}
        if variates is None:
            self.gA = ReferenceVariates(self.seed, self.tapes["A"], None)
            self.gB = ReferenceVariates(self.seed, self.tapes["B"], self.tapes["C"])
            self.gD = ReferenceVariates(self.seed, self.tapes["D"], None)
        else:
            self.gA, self.gB, self.gD = variates["A"], variates["B"], variates["D"]
        self.warm = False
        env.process(self._p_warmup())
        self.tbf = [0] * R
        self.utf = [0] * R
        self.machine = [simpy.Resource(env, resources_quantity(self.res_cfg[r])) for r in self.rnames]
        for r in range(R):
            tb = self.res_cfg[self.rnames[r]].get("tbf", None)
            self.tbf[r] = self._draw_breakdown(tb) if tb else float("inf")
            env.process(self._p_transport(r))
            env.process(self._p_machine(r))
        for p in range(P):
            env.process(self._p_demand(p))
            env.process(self._p_delivery(p))
        if self.policy == "dbr":
            self._dbr_setup()
            env.process(self._p_dbr_rope())
        else:
            env.process(self._p_scheduler())

    # ------------------------------------------------------------------ records
    def _rec(self, ring, value):
        if self.logging:
            self.trace.append((ring, np.float32(value[0]), np.float32(value[1])))

    def _pm(self, name, p, value):
        self._rec(PM.index(name) * self.P + p, value)

    def _rmx(self, name, r, value):
        self._rec(len(PM) * self.P + RM.index(name) * self.R + r, value)
        if name == "queue":
            self.last_queue[r] = (np.float32(value[0]), np.float32(value[1]))

    def _gm(self, name, value):
        self._rec(len(PM) * self.P + len(RM) * self.R + GM.index(name), value)

    def _inventory(self, p, with_wip, with_fg):                       # factory_sim.py:551-601
        if not self.warm:
            return
        now = self.env.now
        if with_wip:
            self._pm("wip", p, (now, self.wip[p].level))
            self._gm("wip_general", (now, self.tot_wip))
        if with_fg:
            lvl = self.fg[p].level
            self.fg_cache[p] = lvl
            self._pm("finishedGoods", p, (now, lvl))
            self._gm("finishedGoods_general", (now, sum(self.fg_cache)))
        inv = self.wip[p].level + self.fg[p].level
        inv_all = self.tot_wip + sum(self.fg_cache)
        self._pm("totalInventory", p, (now, inv))
        self._gm("totalInventory_general", (now, inv_all))

    def _delivery_records(self, d):                                   # factory_sim.py:606-654
        now = self.env.now
        p = d.prod
        if not d.delivered:
            self._pm("lostSales", p, (now, d.qty))
        elif d.delivered <= d.due:
            self._pm("deliveredOntime", p, (now, d.qty))
            self._pm("earliness", p, (now, d.due - now))
        elif d.delivered > d.due:
            self._pm("deliveredLate", p, (now, d.qty))
            self._pm("tardiness", p, (now, now - d.due))
        self._pm("leadTime", p, (now, now - d.arrived))

    # ------------------------------------------------------------------ processes
    def _p_warmup(self):
        yield self.env.timeout(self.warmup)
        self.warm = True

    def _draw_breakdown(self, cfg):                                   # factory_sim.py:291-295
        return self.gD.scalar(cfg.get("dist", "constant"), cfg.get("params", [0]))

    def _p_demand(self, p):
        dm = self.prod_cfg[self.pnames[p]]["demand"]
        f, q, d = dm["freq"], dm["quantity"], dm["duedate"]
        while True:
            gap = np.float32(self.gA.scalar(f.get("dist"), f.get("params")))
            qty = round(self.gA.scalar(q.get("dist"), q.get("params")), 0)
            due = np.float32(self.gA.scalar(d.get("dist"), d.get("params")))
            yield self.env.timeout(gap)
            due += self.env.now
            yield self.inbound.put(_DO(p, qty, due, self.env.now))

    def _p_scheduler(self):
        while True:
            d = yield self.inbound.get()
            po = _PO(d.prod, d.qty)
            po.schedule = self.env.now
            po.due = d.due
            po.prio = 0
            self.env.process(self._p_release(po))
            yield self.outbound[d.prod].put(d)

    def _p_release(self, po):
        p = po.prod
        first = self.route[p][0][0]
        po.total = len(self.route[p])
        po.done = 0
        po.released = self.env.now
        yield self.q_in[first].put(po)
        yield self.wip[p].put(po.qty)
        self.tot_wip += po.qty
        self.qq[first] += po.qty
        if self.warm:
            now = self.env.now
            self._pm("released", p, (now, po.qty))
            self._inventory(p, True, False)
            if self.log_queues:
                last = self.last_queue[first]
                if last is None:
                    val = self.qq[first]
                else:
                    val = last[1] + po.qty
                    self.qq[first] = val
                self._rmx("queue", first, (now, val))

    def _p_breakdown(self, r):
        began = self.env.now
        yield self.env.timeout(self._draw_breakdown(self.res_cfg[self.rnames[r]]["ttr"]))
        ended = self.env.now
        if self.warm:
            self._rmx("breakdown", r, (began, round(ended - began, 6)))
        self.tbf[r] = self._draw_breakdown(self.res_cfg[self.rnames[r]]["tbf"])
        self.utf[r] = 0

    def _select(self, r):
        items = self.q_in[r].items
        if self.policy == "fifo" or not items:
            po = yield self.q_in[r].get()
            return po
        if self.policy == "max_priority":
            pick = max(items, key=lambda o: o.prio)
        elif self.policy == "edd":     # extension (not in the reference): earliest due date, FIFO among equals
            pick = min(items, key=lambda o: o.due)
        else:  # dbr buffer-penetration dispatch, toc_dbr/simulation.py:291-328
            for o in items:
                ahead = []
                for k in range(self.R):
                    ahead.extend(self.q_in[k].items)
                    ahead.extend(self.q_out[k].items)
                    ahead.extend(self.q_trans[k].items)
                    ahead.extend(self.q_proc[k].items)
                qs = [a.qty for a in ahead if a.released < o.released and a.prod == o.prod]
                o.prio = (sum(qs) + self.fg[o.prod].level) / self.ship_buf[o.prod]
            pick = min(items, key=lambda o: o.prio)
        po = yield self.q_in[r].get(lambda o: o.uid == pick.uid)
        return po

    def _p_machine(self, r):
        setup_cfg = self.res_cfg[self.rnames[r]]["setup"]
        prev_prod = None
        prev_step = None
        while True:
            if self.utf[r] >= self.tbf[r]:
                yield from self._p_breakdown(r)
            po = yield from self._select(r)
            self.qq[r] -= po.qty
            if self.warm and self.log_queues:
                self._rmx("queue", r, (self.env.now, self.qq[r]))
            yield self.q_proc[r].put(po)
            p, k = po.prod, po.done
            if prev_prod == p and prev_step == k:       # dead branch (E1): prev_prod never updated
                setup = 0
            else:
                setup = self.gB.scalar(setup_cfg.get("dist", "constant"), setup_cfg.get("params", [0]))
                if self.warm:
                    self._rmx("setup", r, (self.env.now, setup))
            prev_step = k
            with self.machine[r].request() as req:
                yield req
                yield self.env.timeout(setup)
                _, dist, params = self.route[p][k]
                t0 = self.env.now
                total = self.gB.batch_sum(dist, params, int(po.qty))
                yield self.env.timeout(total)
                t1 = self.env.now
                busy = round(t1 - t0, 6)
                self.utf[r] += busy
                po.done += 1
                yield self.q_proc[r].get()
                yield self.q_out[r].put(po)
                if self.policy == "dbr" and self.dbr_repair and r == self.ccr:      # patch P3
                    self.q_fin[r].put(po)
                if self.warm:
                    self._rmx("utilization", r, (self.env.now, busy))

    def _p_transport(self, r):
        while True:
            po = yield self.q_out[r].get()
            yield self.q_trans[r].put(po)
            p = po.prod
            if po.total == po.done:
                yield self.q_trans[r].get()
                yield self.fg[p].put(po.qty)
                yield self.wip[p].get(po.qty)
                self.tot_wip -= po.qty
                self.fg_cache[p] = self.fg[p].level
                if self.warm:
                    now = self.env.now
                    self._pm("flowTime", p, (now, now - po.released))
                    self._inventory(p, True, True)
            else:
                nxt = self.route[p][po.done][0]
                yield self.q_trans[r].get()
                yield self.q_in[nxt].put(po)
                self.qq[nxt] += po.qty
                if self.warm and self.log_queues:
                    last = self.last_queue[nxt]
                    if last is None:
                        val = self.qq[nxt]
                    else:
                        val = last[1] + po.qty
                        self.qq[nxt] = val
                    self._rmx("queue", nxt, (self.env.now, val))

    def _p_delivery(self, p):
        while True:
            d = yield self.outbound[p].get()
            if self.mode == "asReady":
                self.env.process(self._p_ship_ready(d))
            elif self.mode == "onDue":
                self.env.process(self._p_ship_due(d))       # patch P1: inner process only
            elif self.mode == "instantly":
                self.env.process(self._p_ship_now(d))

    def _p_ship_now(self, d):
        if self.fg[d.prod].level >= d.qty:
            yield self.fg[d.prod].get(d.qty)
            d.delivered = self.env.now
        if self.warm:
            self._inventory(d.prod, False, True)
            self._delivery_records(d)

    def _p_ship_ready(self, d):
        yield self.fg[d.prod].get(d.qty)
        d.delivered = self.env.now
        if self.warm:
            self._inventory(d.prod, False, True)
            self._delivery_records(d)

    def _p_ship_due(self, d):
        wait = max(0, d.due - self.env.now)
        if wait > 0:
            yield self.env.timeout(wait)
        yield self.fg[d.prod].get(d.qty)
        d.delivered = self.env.now
        if self.warm:
            self._inventory(d.prod, False, True)
            self._delivery_records(d)

    # ------------------------------------------------------------------ DBR (examples/toc_dbr/simulation.py)
    def _dbr_setup(self):
        cfg = self.cfg
        self.release_limit = cfg.get("order_release_limit", float("inf"))
        self.ccr_limit = cfg.get("ccr_release_limit", False)
        self.tick = cfg.get("scheduler_interval", 72)
        self.cb_target = cfg.get("cb_target_level", float("inf"))
        self.cb_level = 0                                             # E19: fresh-object semantics
        self.ship_buf = [self.prod_cfg[p].get("shipping_buffer", 0) for p in self.pnames]
        # constraint identification, toc_dbr/simulation.py:146-171 (float32 accumulation table)
        load = np.zeros((self.P, self.R), dtype=np.float32)
        for p, name in enumerate(self.pnames):
            dm = self.prod_cfg[name]["demand"]
            gap = dm["freq"]["params"][0]
            qty = dm["quantity"]["params"][0]
            for r, _, params in self.route[p]:
                load[p, r] += params[0]
            load[p, :] = load[p, :] * (1 / gap) * qty
        tot = load.sum(axis=0)
        self.ccr = int(np.argsort(-tot, kind="stable")[0])
        self.ccr_setup = self.res_cfg[self.rnames[self.ccr]].get("setup", {"params": None}).get("params", [0])[0]
        self.ccr_pt = [sum([prm[0] for r, _, prm in self.route[p] if r == self.ccr]) for p in range(self.P)]
        env = self.env
        env.process(self._p_dbr_drain())
        env.process(self._p_dbr_forward())
        for p in range(self.P):
            env.process(self._p_dbr_seed_fg(p))

    def _p_dbr_seed_fg(self, p):
        yield self.fg[p].put(self.ship_buf[p])

    def _p_dbr_forward(self):
        while True:
            d = yield self.inbound.get()
            yield self.outbound[d.prod].put(d)

    def _p_dbr_drain(self):
        while True:
            po = yield self.q_fin[self.ccr].get()
            mean_pt = self.route[po.prod][po.done - 1][2][0]
            self.cb_level -= mean_pt * po.qty + self.ccr_setup

    def _p_dbr_rope(self):
        min_lot = self.cfg.get("min_lotsize", 0)
        while True:
            cand = []
            for p in range(self.P):
                fg = self.fg[p].level
                target = self.ship_buf[p]
                pen = target - fg
                repl = max(target - (self.wip[p].level + self.fg[p].level), 0)
                po = _PO(p, max(repl, min_lot), prio=round(pen / target, 3))
                cand.append((po, self.ccr_pt[p], round(repl / target, 3)))
            cand = list(sorted(cand, key=lambda c: c[-1], reverse=True))
            if self.ccr_limit:
                room = self.tick * self.ccr_limit
            else:
                room = self.cb_target - self.cb_level
            if room > 0:
                n_rel = 0
                for po, ccr_t, _ in cand:
                    go = False
                    if ccr_t > 0:
                        if room > 0 and n_rel < self.release_limit:
                            ccr_t = (po.qty * ccr_t) + self.ccr_setup
                            po.schedule = self.env.now + ccr_t
                            go = True
                            n_rel += 1
                    else:
                        ccr_t = 0
                        po.schedule = self.env.now
                        go = True
                    if po.qty > 0 and go:
                        self.env.process(self._p_dbr_order(po, ccr_t))
                        room -= ccr_t
            yield self.env.timeout(self.tick)

    def _p_dbr_order(self, po, ccr_add):
        if po.schedule is not None and po.schedule > self.env.now:
            yield self.env.timeout(po.schedule - self.env.now)
        self.cb_level += ccr_add
        self.env.process(self._p_release(po))

    # ------------------------------------------------------------------ run
    def run(self):
        error = None
        try:
            self.env.run(until=self.run_until)
        except ValueError as exc:
            error = repr(exc)
        env = self.env
        tr = self.trace
        return {
            "ring": np.array([t[0] for t in tr], dtype=np.uint16),
            "time": np.array([t[1] for t in tr], dtype=np.float32),
            "value": np.array([t[2] for t in tr], dtype=np.float32),
            "tape_A": np.array(self.tapes["A"], dtype=np.float32),
            "tape_B": np.array(self.tapes["B"], dtype=np.float32),
            "tape_C": np.array(self.tapes["C"], dtype=np.float64),
            "tape_D": np.array(self.tapes["D"], dtype=np.float32),
            "steps": int(next(env._eid) - len(env._queue)),
            "now": env.now,
            "error": error,
        }


def resources_quantity(cfg):
    return cfg.get("quantity", 1)


def run_oracle(config, resources, products, seed, policy="fifo", tapes=None, dbr_repair=False):
    """Convenience: one replication; ``tapes`` = dict(A,B,C,D arrays) switches to replay."""
    variates = None
    if tapes is not None:
        variates = {"A": TapeSource(tapes["A"]), "B": TapeSource(tapes["B"], tapes["C"]),
                    "D": TapeSource(tapes["D"])}
    return FactoryOracle(config, resources, products, seed, policy, variates, dbr_repair).run()


def metrics_from_trace(res, P, R, run_until, warmup):
    """Per-ring (sum, count) in float64 + the reference's metric value (factory_sim.py:734-789,
    metrics.py:112-127): sums for sum-metrics (utilization/breakdown/setup divided by now-warmup),
    event-weighted means otherwise; NaN for empty mean-metrics, 0 for empty sum-metrics."""
    K = len(PM) * P + len(RM) * R + len(GM)
    s = np.zeros(K, dtype=np.float64)
    c = np.zeros(K, dtype=np.int64)
    np.add.at(s, res["ring"].astype(np.int64), res["value"].astype(np.float64))
    np.add.at(c, res["ring"].astype(np.int64), 1)
    val = np.full(K, np.nan)
    for ring in range(K):
        if ring < len(PM) * P:
            name = PM[ring // P]
        elif ring < len(PM) * P + len(RM) * R:
            name = RM[(ring - len(PM) * P) // R]
        else:
            name = GM[ring - len(PM) * P - len(RM) * R]
        if name in ("deliveredOntime", "deliveredLate", "lostSales"):
            val[ring] = s[ring]
        elif name in ("utilization", "breakdown", "setup"):
            val[ring] = s[ring] / (run_until - warmup)
        else:
            val[ring] = s[ring] / c[ring] if c[ring] else np.nan
    return s, c, val
