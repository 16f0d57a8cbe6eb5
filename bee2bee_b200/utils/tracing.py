"""Tracing / profiling hooks (the reference has none beyond two wall-clock deltas,
SURVEY section 5.1).  ``BEE2BEE_TRACE=1`` turns on

* NVTX ranges around engine phases (visible in Nsight Systems / ncu --nvtx),
* CUDA-event timing of the same ranges on the launching stream (device time, not wall clock),
* a per-range summary (count / total / mean / p50 / max in ms) exposed through ``/metrics``.

Disabled tracing costs one attribute check per range.
"""
from __future__ import annotations

import contextlib
import os
import statistics
import threading
import time
from collections import defaultdict
from typing import Dict, List, Optional


class Tracer:
    def __init__(self, enabled: Optional[bool] = None):
        self.enabled = (os.environ.get("BEE2BEE_TRACE", "0") not in ("0", "", "false")) if enabled is None else enabled
        self._lock = threading.Lock()
        self._pending: List[tuple] = []           # (name, start_event, end_event)
        self._samples: Dict[str, List[float]] = defaultdict(list)

    @contextlib.contextmanager
    def range(self, name: str, stream=None, device_timed: bool = True):
        if not self.enabled:
            yield
            return
        import torch

        cuda = torch.cuda.is_available() and device_timed
        if cuda:
            torch.cuda.nvtx.range_push(name)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        t0 = time.perf_counter()
        try:
            yield
        finally:
            if cuda:
                e1.record(stream)
                torch.cuda.nvtx.range_pop()
                with self._lock:
                    self._pending.append((name, e0, e1))
            else:
                with self._lock:
                    self._samples[name].append((time.perf_counter() - t0) * 1e3)

    def _drain(self) -> None:
        keep = []
        for name, e0, e1 in self._pending:
            if e1.query():
                self._samples[name].append(e0.elapsed_time(e1))
            else:
                keep.append((name, e0, e1))
        self._pending = keep

    def summary(self) -> Dict[str, Dict[str, float]]:
        with self._lock:
            self._drain()
            out = {}
            for name, xs in self._samples.items():
                if xs:
                    out[name] = {"count": len(xs), "total_ms": sum(xs), "mean_ms": sum(xs) / len(xs),
                                 "p50_ms": statistics.median(xs), "max_ms": max(xs)}
            return out

    def reset(self) -> None:
        with self._lock:
            self._pending.clear()
            self._samples.clear()


TRACER = Tracer()
