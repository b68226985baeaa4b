"""stillleben.profiling.Timer: a nesting wall-clock timer used as context manager or decorator, printed as an indented
tree when the outermost timer ends; a no-op unless `Timer.enabled` is set (interface of python/stillleben/profiling.py)."""
import time
from contextlib import ContextDecorator


class Timer(ContextDecorator):
    active_timers = []
    enabled = False

    def __init__(self, name):
        self.name, self.children, self.duration = name, [], 0.0

    def __enter__(self):
        if Timer.enabled:
            self._t0 = time.time()
            self.children = []
            Timer.active_timers.append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        if not Timer.enabled:
            return
        self.duration = time.time() - self._t0
        top = Timer.active_timers.pop()
        assert top is self
        if Timer.active_timers:
            Timer.active_timers[-1].children.append(self)
        else:
            print("Timings:")
            self._report(0)

    def _report(self, indent):
        print("%s%-*s%8.3fs" % (" " * indent, 30 - indent, self.name, self.duration))
        for c in self.children:
            c._report(indent + 2)
