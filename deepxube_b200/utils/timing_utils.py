"""Cumulative named timers (mirror of deepxube/utils/timing_utils.py:33-112: same keys / string format, so
`solve`'s "Times - ..." lines stay comparable with the reference's output.txt)."""
from collections import OrderedDict
from typing import Dict, List, Optional


class Times:
    def __init__(self, time_names: Optional[List[str]] = None) -> None:
        self.times: "OrderedDict[str, float]" = OrderedDict((n, 0.0) for n in (time_names or []))
        self.counts: "OrderedDict[str, int]" = OrderedDict((n, 0) for n in (time_names or []))
        self.sub_times: Dict[str, "Times"] = {}
        self.sub_counts: Dict[str, int] = {}

    def _sub(self, name: str) -> "Times":
        if name not in self.sub_times:
            self.sub_times[name] = Times()
            self.sub_counts[name] = 0
        return self.sub_times[name]

    def record_time(self, time_name: str, time_elapsed: float, path: Optional[List[str]] = None) -> None:
        if path:
            head = path.pop(0)
            self._sub(head).record_time(time_name, time_elapsed, path=path)
            self.sub_counts[head] += 1
            return
        self.times[time_name] = self.times.get(time_name, 0.0) + time_elapsed
        self.counts[time_name] = self.counts.get(time_name, 0) + 1

    def add_times(self, time: "Times", path: Optional[List[str]] = None) -> None:
        if path:
            head = path.pop(0)
            self._sub(head).add_times(time, path=path)
            self.sub_counts[head] += sum(time.counts.values())
            return
        for k, v in time.times.items():
            self.times[k] = self.times.get(k, 0.0) + v
        for k, c in time.counts.items():
            self.counts[k] = self.counts.get(k, 0) + c
        for name, sub in time.sub_times.items():
            self._sub(name).add_times(sub)
            self.sub_counts[name] += sum(sub.counts.values())

    def reset_times(self) -> None:
        for k in self.times:
            self.times[k] = 0.0
            self.counts[k] = 0
        for sub in self.sub_times.values():
            sub.reset_times()

    def get_total_time(self) -> float:
        return sum(self.times.values()) + sum(s.get_total_time() for s in self.sub_times.values())

    def get_time_str(self, prefix: str = "", decplace: int = 2) -> str:
        parts = [f"{k}: {v:.{decplace}f}" for k, v in self.times.items()]
        parts += [f"->{k}: {s.get_total_time():.{decplace}f}" for k, s in self.sub_times.items()]
        out = ", ".join(parts + [f"Tot: {self.get_total_time():.{decplace}f}"])
        inner = f"\t{prefix}"
        for k, s in self.sub_times.items():
            out = f"{out}\n{inner}({k}): {s.get_time_str(prefix=inner)}"
        return out

    def __str__(self) -> str:
        return self.get_time_str()

    __repr__ = __str__
