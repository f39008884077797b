"""CPU: oracle/minisimpy.py against the example outputs printed in SimPy's own documentation.

SimPy 4.1.1 cannot be installed here (no network, no wheel), so the event-kernel restatement stays formally unpinned
(DESIGN.md section 2).  What can be done without the package: its documentation prints the output of several small
programs ("SimPy in 10 minutes", front page, topical guide on resources).  Those programs and outputs are restated
below from memory of the 4.x documentation; each one pins a rule the GPU kernel's micro-event order depends on:
  * tie-breaking at equal times by scheduling order (front-page clock example: `slow 1` before `fast 1.0`);
  * a process is an event other processes can wait for (car/charge example);
  * Resource: release -> next request succeeds -> its process resumes at the same instant, after the releaser printed;
  * Store: the put's `_trigger_get` callback runs before the putting process resumes, one getter is served per put,
    and `run(until=t)` stops before NORMAL events scheduled at exactly t.
"""
from oracle import minisimpy as simpy


def test_front_page_clock_example_orders_ties_by_scheduling_order():
    out = []

    def clock(env, name, tick):
        while True:
            out.append((name, env.now))
            yield env.timeout(tick)

    env = simpy.Environment()
    env.process(clock(env, "fast", 0.5))
    env.process(clock(env, "slow", 1))
    env.run(until=2)
    assert out == [("fast", 0), ("slow", 0), ("fast", 0.5), ("slow", 1), ("fast", 1.0), ("fast", 1.5)]


def test_car_example_basic_concepts():
    out = []

    def car(env):
        while True:
            out.append(f"Start parking at {env.now}")
            yield env.timeout(5)
            out.append(f"Start driving at {env.now}")
            yield env.timeout(2)

    env = simpy.Environment()
    env.process(car(env))
    env.run(until=15)
    assert out == ["Start parking at 0", "Start driving at 5", "Start parking at 7", "Start driving at 12",
                   "Start parking at 14"]


def test_waiting_for_a_process():
    out = []

    class Car:
        def __init__(self, env):
            self.env = env
            self.action = env.process(self.run())

        def run(self):
            while True:
                out.append(f"Start parking and charging at {self.env.now}")
                yield self.env.process(self.charge(5))
                out.append(f"Start driving at {self.env.now}")
                yield self.env.timeout(2)

        def charge(self, duration):
            yield self.env.timeout(duration)

    env = simpy.Environment()
    Car(env)
    env.run(until=15)
    assert out == ["Start parking and charging at 0", "Start driving at 5", "Start parking and charging at 7",
                   "Start driving at 12", "Start parking and charging at 14"]


def test_battery_charging_station_resource_example():
    out = []

    def car(env, name, bcs, driving_time, charge_duration):
        yield env.timeout(driving_time)
        out.append("%s arriving at %d" % (name, env.now))
        with bcs.request() as req:
            yield req
            out.append("%s starting to charge at %s" % (name, env.now))
            yield env.timeout(charge_duration)
            out.append("%s leaving the bcs at %s" % (name, env.now))

    env = simpy.Environment()
    bcs = simpy.Resource(env, capacity=2)
    for i in range(4):
        env.process(car(env, "Car %d" % i, bcs, i * 2, 5))
    env.run()
    assert out == [
        "Car 0 arriving at 0", "Car 0 starting to charge at 0",
        "Car 1 arriving at 2", "Car 1 starting to charge at 2",
        "Car 2 arriving at 4",
        "Car 0 leaving the bcs at 5", "Car 2 starting to charge at 5",
        "Car 3 arriving at 6",
        "Car 1 leaving the bcs at 7", "Car 3 starting to charge at 7",
        "Car 2 leaving the bcs at 10", "Car 3 leaving the bcs at 12"]


def test_store_producer_consumer_example():
    out = []

    def producer(env, store):
        for i in range(100):
            yield env.timeout(2)
            yield store.put(f"spam {i}")
            out.append(f"Produced spam at {env.now}")

    def consumer(name, env, store):
        while True:
            yield env.timeout(1)
            out.append(f"{name} requesting spam at {env.now}")
            item = yield store.get()
            out.append(f"{name} got {item} at {env.now}")

    env = simpy.Environment()
    store = simpy.Store(env, capacity=2)
    env.process(producer(env, store))
    [env.process(consumer(i, env, store)) for i in range(2)]
    env.run(until=5)
    assert out == ["0 requesting spam at 1", "1 requesting spam at 1", "Produced spam at 2", "0 got spam 0 at 2",
                   "0 requesting spam at 3", "Produced spam at 4", "1 got spam 1 at 4"]


def test_run_until_semantics():
    """core.Environment.run: an int `until` keeps `now` an int; until <= now raises; events AT `until` are not run."""
    env = simpy.Environment()
    fired = []

    def p(env):
        yield env.timeout(3)
        fired.append(env.now)
        yield env.timeout(2)
        fired.append(env.now)

    env.process(p(env))
    env.run(until=5)
    assert fired == [3] and env.now == 5 and isinstance(env.now, int)
    env.run(until=5.5)
    assert fired == [3, 5] and env.now == 5.5
    try:
        env.run(until=5.5)
    except ValueError as e:
        assert "must be > the current simulation time" in str(e)
    else:
        raise AssertionError("until <= now must raise")


def test_filter_store_machine_shop_example():
    """Topical guide on stores: FilterStore.get(filter) -- the reference's resource_input queues are FilterStores
    (src/manusim/engine/stores.py:37-41) and simple_scheduler picks orders through an identity filter."""
    from collections import namedtuple

    out = []
    Machine = namedtuple("Machine", "size, duration")
    m1 = Machine(1, 2)
    m2 = Machine(2, 1)
    env = simpy.Environment()
    machine_shop = simpy.FilterStore(env, capacity=2)
    machine_shop.items = [m1, m2]

    def user(name, env, ms, size):
        machine = yield ms.get(lambda machine: machine.size == size)
        out.append(f"{name} got {machine} at {env.now}")
        yield env.timeout(machine.duration)
        yield ms.put(machine)
        out.append(f"{name} released {machine} at {env.now}")

    [env.process(user(i, env, machine_shop, (i % 2) + 1)) for i in range(3)]
    env.run()
    assert out == ["0 got Machine(size=1, duration=2) at 0", "1 got Machine(size=2, duration=1) at 0",
                   "1 released Machine(size=2, duration=1) at 1", "0 released Machine(size=1, duration=2) at 2",
                   "2 got Machine(size=1, duration=2) at 2", "2 released Machine(size=1, duration=2) at 4"]


def test_container_rules():
    """resources/container.py: amount <= 0 raises at construction of the request; a get larger than the level blocks
    and blocks the getters behind it (head-of-line), a put wakes them in order."""
    env = simpy.Environment()
    c = simpy.Container(env, init=1)
    for bad in (0, -1):
        try:
            c.put(bad)
        except ValueError as e:
            assert "must be > 0" in str(e)
        else:
            raise AssertionError
    got = []

    def getter(name, amount):
        yield c.get(amount)
        got.append((name, env.now, c.level))

    def putter():
        yield env.timeout(1)
        yield c.put(2)

    env.process(getter("big", 3))        # cannot be served from level 1
    env.process(getter("small", 1))      # could be, but waits behind "big"
    env.process(putter())
    env.run()
    assert got == [("big", 1, 0)] and c.level == 0 and len(c.get_queue) == 1
