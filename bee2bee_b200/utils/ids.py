"""Identifiers, clocks and hashing helpers (parity: /root/reference/bee2bee/utils.py:43-64)."""
from __future__ import annotations

import hashlib
import platform
import secrets
import time


def new_id(prefix: str) -> str:
    return f"{prefix}-{secrets.token_hex(4)}"


def now_ms() -> int:
    return time.time_ns() // 1_000_000


def os_name() -> str:
    return platform.system()


def sha256_hex(text: str) -> str:
    return hashlib.sha256(text.encode("utf-8")).hexdigest()


def hash_password(password: str, salt: str) -> str:
    return sha256_hex(f"{password}:{salt}")


def gen_salt() -> str:
    return secrets.token_hex(16)
