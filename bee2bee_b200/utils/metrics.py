"""Node metrics with the reference's JSON keys (``throughput``, ``memory_percent``,
``gpu_percent``, ``trust_score``; /root/reference/bee2bee/utils.py:102-135) -- but
``throughput`` is the engine's *measured* decode tokens/s when an engine is attached
(the reference reports ``cpu% x 0.85``), and multi-GPU ``nvidia-smi`` output is averaged
instead of failing to parse."""
from __future__ import annotations

import shutil
import subprocess
from typing import Callable, Dict, List, Optional

_throughput_source: Optional[Callable[[], float]] = None


def set_throughput_source(fn: Optional[Callable[[], float]]) -> None:
    """Attach a callable returning measured tokens/s (e.g. ``lambda: engine.metrics()['tokens_per_s']``)."""
    global _throughput_source
    _throughput_source = fn


def get_gpu_usages() -> List[float]:
    if not shutil.which("nvidia-smi"):
        return []
    try:
        out = subprocess.check_output(
            ["nvidia-smi", "--query-gpu=utilization.gpu", "--format=csv,noheader,nounits"],
            stderr=subprocess.DEVNULL, timeout=5).decode()
        return [float(x) for x in out.split() if x.strip().replace(".", "", 1).isdigit()]
    except Exception:
        return []


def get_gpu_usage() -> float:
    vals = get_gpu_usages()
    return sum(vals) / len(vals) if vals else 0.0


def get_system_metrics() -> Dict[str, float]:
    try:
        import psutil

        gpu = get_gpu_usage()
        cpu = psutil.cpu_percent(interval=None)
        mem = psutil.virtual_memory().percent
        if _throughput_source is not None:
            try:
                tput = round(float(_throughput_source()), 1)
            except Exception:
                tput = 0.0
        else:
            tput = round(cpu * 0.85, 1)          # reference placeholder when no engine is attached
        return {"throughput": tput, "memory_percent": mem, "gpu_percent": gpu, "trust_score": 0.98 + gpu * 1e-4}
    except Exception:
        return {"throughput": 0.0, "memory_percent": 0.0, "gpu_percent": 0.0, "trust_score": 1.0}
