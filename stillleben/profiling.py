"""stillleben.profiling — `Timer`, the nesting wall-clock timer the reference's scripts wrap around their stages
(interface of python/stillleben/profiling.py: usable as `with Timer("name"):` and as `@Timer("name")`, silent unless
`Timer.enabled` is set, report printed when the outermost timer closes).

Implementation: a module-level stack of open spans; every span records (depth, name, seconds) into the report of the
outermost span in the order the spans were OPENED, so the printed tree lists stages chronologically."""
import time
from contextlib import ContextDecorator

_open_spans = []        # the spans currently running, outermost first
_report = []            # [depth, name, seconds] rows of the running outermost span


class Timer(ContextDecorator):
    enabled = False
    active_timers = _open_spans     # (name kept from the reference's class attribute)

    def __init__(self, name):
        self.name = str(name)
        self.duration = 0.0
        self._row = None
        self._t0 = 0.0

    def __enter__(self):
        if Timer.enabled:
            self._row = [len(_open_spans), self.name, 0.0]
            _report.append(self._row)
            _open_spans.append(self)
            self._t0 = time.perf_counter()
        return self

    def __exit__(self, exc_type, exc, tb):
        if self._row is None:
            return False
        self.duration = self._row[2] = time.perf_counter() - self._t0
        self._row = None
        closed = _open_spans.pop()
        if closed is not self:
            raise RuntimeError("Timer spans must close in the order they were opened")
        if not _open_spans:
            print("Timings:")
            for depth, name, seconds in _report:
                pad = 2 * depth
                print(f"{'':{pad}}{name:<{30 - pad}}{seconds:8.3f}s")
            del _report[:]
        return False
