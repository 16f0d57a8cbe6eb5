"""Home directory and atomic JSON persistence (parity: /root/reference/bee2bee/utils.py:11-40)."""
from __future__ import annotations

import json
import os
import tempfile
from pathlib import Path
from typing import Any


def bee2bee_home() -> Path:
    """``$BEE2BEE_HOME`` or ``~/.bee2bee`` (created on demand)."""
    root = Path(os.environ.get("BEE2BEE_HOME") or (Path.home() / ".bee2bee"))
    root.mkdir(parents=True, exist_ok=True)
    return root


def data_file(name: str) -> Path:
    target = bee2bee_home() / name
    target.parent.mkdir(parents=True, exist_ok=True)
    return target


def load_json(path: Path, default: Any) -> Any:
    try:
        with open(path, "r", encoding="utf-8") as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return default


def save_json(path: Path, obj: Any) -> None:
    """Write-then-rename so readers never observe a torn file."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    fd, tmp = tempfile.mkstemp(prefix=path.name + ".", suffix=".tmp", dir=str(path.parent))
    try:
        with os.fdopen(fd, "w", encoding="utf-8") as fh:
            json.dump(obj, fh, indent=2, ensure_ascii=False)
        os.replace(tmp, path)
    except BaseException:
        try:
            os.unlink(tmp)
        except OSError:
            pass
        raise
