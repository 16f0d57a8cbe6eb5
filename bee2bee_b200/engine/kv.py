"""Paged KV-cache bookkeeping (host side).  Pages hold 64 tokens (``ops.PAGE``); page 0 is a
scratch page that inactive batch rows point at, so it is never handed out.  The reference
has no KV management at all (HF ``DynamicCache`` grows by concatenation per request and is
dropped afterwards, /root/reference/bee2bee/hf.py:42-43)."""
from __future__ import annotations

from typing import Dict, List

PAGE = 64


class OutOfPages(RuntimeError):
    pass


class PageAllocator:
    def __init__(self, num_pages: int):
        self.num_pages = num_pages
        self._free: List[int] = list(range(num_pages - 1, 0, -1))     # page 0 reserved
        self._owned: Dict[int, List[int]] = {}

    @property
    def free_pages(self) -> int:
        return len(self._free)

    @staticmethod
    def pages_for(tokens: int) -> int:
        return (tokens + PAGE - 1) // PAGE

    def can_allocate(self, tokens: int) -> bool:
        return self.pages_for(tokens) <= len(self._free)

    def allocate(self, owner: int, tokens: int) -> List[int]:
        n = self.pages_for(tokens)
        if n > len(self._free):
            raise OutOfPages(f"need {n} pages, {len(self._free)} free")
        pages = [self._free.pop() for _ in range(n)]
        self._owned.setdefault(owner, []).extend(pages)
        return pages

    def release(self, owner: int) -> None:
        for p in self._owned.pop(owner, []):
            self._free.append(p)

    def owned(self, owner: int) -> List[int]:
        return list(self._owned.get(owner, []))

    def utilization(self) -> float:
        return 1.0 - len(self._free) / max(1, self.num_pages - 1)
