"""minisimpy -- CPU restatement of the SimPy 4.1.1 discrete-event kernel.

TEST INFRASTRUCTURE ONLY.  Nothing under ``manusim_b200/`` may import this file;
it exists so that (a) the *unmodified* reference model under
``/root/reference/src/manusim`` can be executed in a container that has no
SimPy wheel (``oracle/run_reference.py``), and (b) the traveling oracle model
``oracle/factory_oracle.py`` has an event kernel to run on.

PARITY UNPINNED: ``simpy==4.1.1`` (reference ``requirements.txt:10``) is a
third-party dependency that is neither vendored in ``/root/reference`` nor
installable here (no network, no wheel in /opt/wheelhouse).  This module
restates SimPy's published algorithm from its documentation / source as
remembered; it has never been diffed against the
real package.  The reference ships no tests or golden vectors for this layer.
The only external anchor available offline: the example programs and printed
outputs of SimPy's documentation, restated from memory in
``tests/test_minisimpy_documented_behaviour.py`` (tie-breaking, Resource and
Store hand-over order, ``run(until)``).

Semantics restated (SimPy 4.1.1 module in brackets):
  * [core.Environment] heap of ``(time, priority, eid, event)``; URGENT=0,
    NORMAL=1; ``step`` pops, sets ``now``, detaches the callback list and runs
    it; un-defused failures propagate; ``run(until=number)`` schedules an
    URGENT stop event at ``until - now`` (int stays int, else float()).
  * [events.Event/Timeout/Initialize/Process] ``succeed`` schedules NORMAL at
    ``now``; a process keeps running synchronously while the yielded event is
    already *processed*, otherwise it appends ``_resume`` and sleeps; process
    termination schedules the process event itself (NORMAL).
  * [resources.base] ``Put.__init__`` appends to ``put_queue``, registers
    ``_trigger_get`` as first callback and calls ``_trigger_put(None)``;
    ``Get.__init__`` is the mirror image; ``_trigger_*`` walk the queue from
    index 0 and stop as soon as ``_do_*`` returns a falsy value.
  * [resources.store] Store (FIFO, ``_do_*`` return None), FilterStore
    (first item matching the filter, ``_do_get`` returns True).
  * [resources.container] Container (``amount <= 0`` -> ValueError, ``_do_*``
    return True on success, None otherwise => head-of-line blocking).
  * [resources.resource] Resource/Request/Release (``with`` exit releases).
"""
from __future__ import annotations

from heapq import heappop, heappush
from itertools import count

Infinity = float("inf")
URGENT = 0
NORMAL = 1
PENDING = object()


class EmptySchedule(Exception):
    pass


class StopSimulation(Exception):
    @classmethod
    def callback(cls, event):
        if event._ok:
            raise cls(event._value)
        raise event._value


class Interrupt(Exception):
    pass


class Event:
    def __init__(self, env):
        self.env = env
        self.callbacks = []
        self._value = PENDING

    @property
    def triggered(self):
        return self._value is not PENDING

    @property
    def processed(self):
        return self.callbacks is None

    @property
    def ok(self):
        return self._ok

    @property
    def value(self):
        if self._value is PENDING:
            raise AttributeError("value not yet available")
        return self._value

    def succeed(self, value=None):
        if self._value is not PENDING:
            raise RuntimeError(f"{self} has already been triggered")
        self._ok = True
        self._value = value
        self.env.schedule(self)
        return self

    def fail(self, exception):
        if self._value is not PENDING:
            raise RuntimeError(f"{self} has already been triggered")
        if not isinstance(exception, BaseException):
            raise TypeError(f"{exception} is not an exception.")
        self._ok = False
        self._value = exception
        self.env.schedule(self)
        return self


class Timeout(Event):
    def __init__(self, env, delay, value=None):
        if delay < 0:
            raise ValueError(f"Negative delay {delay}")
        self.env = env
        self.callbacks = []
        self._value = value
        self._delay = delay
        self._ok = True
        env.schedule(self, NORMAL, delay)


class Initialize(Event):
    def __init__(self, env, process):
        self.env = env
        self.callbacks = [process._resume]
        self._value = None
        self._ok = True
        env.schedule(self, URGENT)


class Process(Event):
    def __init__(self, env, generator):
        if not hasattr(generator, "throw"):
            raise ValueError(f"{generator} is not a generator.")
        self.env = env
        self.callbacks = []
        self._value = PENDING
        self._generator = generator
        self._target = Initialize(env, self)

    @property
    def is_alive(self):
        return self._value is PENDING

    def _resume(self, event):
        env = self.env
        env._active_proc = self
        while True:
            try:
                if event._ok:
                    event = self._generator.send(event._value)
                else:
                    event._defused = True
                    event = self._generator.throw(event._value)
            except StopIteration as stop:
                event = None
                self._ok = True
                self._value = stop.args[0] if len(stop.args) else None
                env.schedule(self)
                break
            except BaseException as exc:
                event = None
                self._ok = False
                self._value = exc
                env.schedule(self)
                break
            try:
                if event.callbacks is not None:
                    event.callbacks.append(self._resume)
                    break
            except AttributeError:
                raise RuntimeError(f'Invalid yield value "{event}"') from None
        self._target = event
        env._active_proc = None


class Environment:
    def __init__(self, initial_time=0):
        self._now = initial_time
        self._queue = []
        self._eid = count()
        self._active_proc = None
        self.steps = 0  # oracle instrumentation: number of step() calls completed

    @property
    def now(self):
        return self._now

    @property
    def active_process(self):
        return self._active_proc

    def process(self, generator):
        return Process(self, generator)

    def timeout(self, delay=0, value=None):
        return Timeout(self, delay, value)

    def event(self):
        return Event(self)

    def schedule(self, event, priority=NORMAL, delay=0):
        heappush(self._queue, (self._now + delay, priority, next(self._eid), event))

    def peek(self):
        try:
            return self._queue[0][0]
        except IndexError:
            return Infinity

    def step(self):
        try:
            self._now, _, _, event = heappop(self._queue)
        except IndexError:
            raise EmptySchedule() from None
        self.steps += 1
        callbacks, event.callbacks = event.callbacks, None
        for callback in callbacks:
            callback(event)
        if not event._ok and not hasattr(event, "_defused"):
            exc = type(event._value)(*event._value.args)
            exc.__cause__ = event._value
            raise exc

    def run(self, until=None):
        if until is not None:
            if not isinstance(until, Event):
                at = until if isinstance(until, int) else float(until)
                if at <= self.now:
                    raise ValueError(f"until(={at}) must be > the current simulation time.")
                until = Event(self)
                until._ok = True
                until._value = None
                self.schedule(until, URGENT, at - self.now)
            elif until.callbacks is None:
                return until.value
            until.callbacks.append(StopSimulation.callback)
        try:
            while True:
                self.step()
        except StopSimulation as exc:
            return exc.args[0]
        except EmptySchedule:
            if until is not None:
                assert not until.triggered
                raise RuntimeError(
                    f'No scheduled events left but "until" event was not triggered: {until}'
                ) from None
        return None


# ----------------------------------------------------------------- resources
class Put(Event):
    def __init__(self, resource):
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.put_queue.append(self)
        self.callbacks.append(resource._trigger_get)
        resource._trigger_put(None)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.cancel()
        return None

    def cancel(self):
        if not self.triggered:
            self.resource.put_queue.remove(self)


class Get(Event):
    def __init__(self, resource):
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.get_queue.append(self)
        self.callbacks.append(resource._trigger_put)
        resource._trigger_get(None)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.cancel()
        return None

    def cancel(self):
        if not self.triggered:
            self.resource.get_queue.remove(self)


class BaseResource:
    def __init__(self, env, capacity):
        self._env = env
        self._capacity = capacity
        self.put_queue = []
        self.get_queue = []

    @property
    def capacity(self):
        return self._capacity

    def _do_put(self, event):
        raise NotImplementedError

    def _do_get(self, event):
        raise NotImplementedError

    def _trigger_put(self, get_event):
        idx = 0
        while idx < len(self.put_queue):
            put_event = self.put_queue[idx]
            proceed = self._do_put(put_event)
            if not put_event.triggered:
                idx += 1
            elif self.put_queue.pop(idx) != put_event:
                raise RuntimeError("Put queue invariant violated")
            if not proceed:
                break

    def _trigger_get(self, put_event):
        idx = 0
        while idx < len(self.get_queue):
            get_event = self.get_queue[idx]
            proceed = self._do_get(get_event)
            if not get_event.triggered:
                idx += 1
            elif self.get_queue.pop(idx) != get_event:
                raise RuntimeError("Get queue invariant violated")
            if not proceed:
                break


class StorePut(Put):
    def __init__(self, store, item):
        self.item = item
        super().__init__(store)


class StoreGet(Get):
    pass


class FilterStoreGet(StoreGet):
    def __init__(self, resource, filter=lambda item: True):
        self.filter = filter
        super().__init__(resource)


class Store(BaseResource):
    def __init__(self, env, capacity=Infinity):
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        super().__init__(env, capacity)
        self.items = []

    def put(self, item):
        return StorePut(self, item)

    def get(self):
        return StoreGet(self)

    def _do_put(self, event):
        if len(self.items) < self._capacity:
            self.items.append(event.item)
            event.succeed()
        return None

    def _do_get(self, event):
        if self.items:
            event.succeed(self.items.pop(0))
        return None


class FilterStore(Store):
    def get(self, filter=lambda item: True):
        return FilterStoreGet(self, filter)

    def _do_get(self, event):
        for item in self.items:
            if event.filter(item):
                self.items.remove(item)
                event.succeed(item)
                break
        return True


class ContainerPut(Put):
    def __init__(self, container, amount):
        if amount <= 0:
            raise ValueError(f"amount(={amount}) must be > 0.")
        self.amount = amount
        super().__init__(container)


class ContainerGet(Get):
    def __init__(self, container, amount):
        if amount <= 0:
            raise ValueError(f"amount(={amount}) must be > 0.")
        self.amount = amount
        super().__init__(container)


class Container(BaseResource):
    def __init__(self, env, capacity=Infinity, init=0):
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        if init < 0:
            raise ValueError('"init" must be >= 0.')
        if init > capacity:
            raise ValueError('"init" must be <= "capacity".')
        super().__init__(env, capacity)
        self._level = init

    @property
    def level(self):
        return self._level

    def put(self, amount):
        return ContainerPut(self, amount)

    def get(self, amount):
        return ContainerGet(self, amount)

    def _do_put(self, event):
        if self._capacity - self._level >= event.amount:
            self._level += event.amount
            event.succeed()
            return True
        return None

    def _do_get(self, event):
        if self._level >= event.amount:
            self._level -= event.amount
            event.succeed()
            return True
        return None


class Request(Put):
    def __exit__(self, exc_type, exc_value, traceback):
        super().__exit__(exc_type, exc_value, traceback)
        if exc_type is not GeneratorExit:
            self.resource.release(self)
        return None


class Release(Get):
    def __init__(self, resource, request):
        self.request = request
        super().__init__(resource)


class Resource(BaseResource):
    def __init__(self, env, capacity=1):
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        super().__init__(env, capacity)
        self.users = []
        self.queue = self.put_queue

    @property
    def count(self):
        return len(self.users)

    def request(self):
        return Request(self)

    def release(self, request):
        return Release(self, request)

    def _do_put(self, event):
        if len(self.users) < self.capacity:
            self.users.append(event)
            event.usage_since = self._env.now
            event.succeed()

    def _do_get(self, event):
        try:
            self.users.remove(event.request)
        except ValueError:
            pass
        event.succeed()
