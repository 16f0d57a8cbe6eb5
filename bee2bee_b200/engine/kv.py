"""Paged KV-cache bookkeeping (host side).  Pages hold 64 tokens (``ops.PAGE``); page 0 is a
scratch page that inactive batch rows point at, so it is never handed out.  The reference
has no KV management at all (HF ``DynamicCache`` grows by concatenation per request and is
dropped afterwards, /root/reference/bee2bee/hf.py:42-43) and re-sends the whole chat transcript
every turn with no reuse.

**Prefix cache** (round 2): full prompt pages are content-addressed (hash chain over 64-token blocks).  A new request
whose prompt starts with blocks that are still resident shares those physical pages (reference-counted) and only its
suffix is prefilled -- every piece of the mesh keeps the same page ids, so the decision is replicated on all ranks.
Sequences that share a prefix therefore read the SAME KV pages in attention (one copy in HBM / L2 for the whole batch).
Released pages that carry a hash stay resident (LRU) until the pool needs them.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

PAGE = 64


class OutOfPages(RuntimeError):
    pass


def _block_keys(tokens: Sequence[int], n_blocks: int) -> List[bytes]:
    """hash chain: key_i identifies tokens[0 : (i + 1) * PAGE]"""
    keys, h = [], b""
    for i in range(n_blocks):
        blk = tokens[i * PAGE:(i + 1) * PAGE]
        h = hashlib.blake2b(h + b"".join(int(t).to_bytes(4, "little", signed=True) for t in blk), digest_size=16).digest()
        keys.append(h)
    return keys


class PageAllocator:
    def __init__(self, num_pages: int, prefix_cache: bool = True):
        self.num_pages = num_pages
        self.prefix_cache = prefix_cache
        self._free: List[int] = list(range(num_pages - 1, 0, -1))     # page 0 reserved
        self._owned: Dict[int, List[int]] = {}
        self._ref: Dict[int, int] = {}                                # page -> number of owners
        self._key_of: Dict[int, bytes] = {}                           # cached page -> prefix key
        self._page_of: Dict[bytes, int] = {}                          # prefix key -> page
        self._lru: "OrderedDict[int, None]" = OrderedDict()           # cached pages nobody owns (evictable), oldest first
        self._cached: Dict[int, int] = {}                             # owner -> prompt tokens served from the cache
        self.hit_tokens = 0
        self.lookup_tokens = 0

    # ------------------------------------------------------------------ capacity
    @property
    def free_pages(self) -> int:
        """pages available to a new sequence: never-used / recycled pages plus evictable cached ones"""
        return len(self._free) + len(self._lru)

    @staticmethod
    def pages_for(tokens: int) -> int:
        return (tokens + PAGE - 1) // PAGE

    def _match(self, prompt: Optional[Sequence[int]]) -> List[int]:
        if not self.prefix_cache or not prompt:
            return []
        # at least one prompt token must be computed (its logits produce the first generated token)
        n = (len(prompt) - 1) // PAGE
        pages = []
        for key in _block_keys(prompt, n):
            p = self._page_of.get(key)
            if p is None:
                break
            pages.append(p)
        return pages

    def can_allocate(self, tokens: int, prompt: Optional[Sequence[int]] = None) -> bool:
        matched = self._match(prompt)
        evictable = len(self._lru) - sum(1 for p in matched if p in self._lru)
        return self.pages_for(tokens) - len(matched) <= len(self._free) + evictable

    # ------------------------------------------------------------------ allocate / release
    def _take(self) -> int:
        if self._free:
            return self._free.pop()
        if self._lru:
            p, _ = self._lru.popitem(last=False)                      # evict the least recently used cached page
            self._page_of.pop(self._key_of.pop(p), None)
            return p
        raise OutOfPages("no free or evictable page")

    def allocate(self, owner: int, tokens: int, prompt: Optional[Sequence[int]] = None) -> List[int]:
        """Pages for ``tokens`` tokens (prompt + generation budget).  With ``prompt`` the longest resident prefix is
        shared; ``cached_tokens(owner)`` tells how many prompt tokens need no prefill."""
        matched = self._match(prompt)
        n = self.pages_for(tokens) - len(matched)
        evictable = len(self._lru) - sum(1 for p in matched if p in self._lru)
        if n > len(self._free) + evictable:
            raise OutOfPages(f"need {n} pages, {len(self._free) + evictable} free")
        for p in matched:                                             # pin the shared pages first: they must not be evicted below
            self._lru.pop(p, None)
            self._ref[p] = self._ref.get(p, 0) + 1
        fresh = [self._take() for _ in range(max(0, n))]
        for p in fresh:
            self._ref[p] = 1
        pages = matched + fresh
        self._owned.setdefault(owner, []).extend(pages)
        self._cached[owner] = len(matched) * PAGE
        if prompt is not None:
            self.lookup_tokens += len(prompt)
            self.hit_tokens += len(matched) * PAGE
        return pages

    def cached_tokens(self, owner: int) -> int:
        return self._cached.get(owner, 0)

    def commit(self, owner: int, prompt: Sequence[int]) -> None:
        """Register the full prompt pages of ``owner`` as shareable (their KV is written by the prefill that is enqueued
        before any later request's; they are never modified afterwards: decode appends behind the prompt)."""
        if not self.prefix_cache:
            return
        pages = self._owned.get(owner, [])
        n = min(len(prompt) // PAGE, len(pages))
        for key, p in zip(_block_keys(prompt, n), pages):
            if key not in self._page_of and p not in self._key_of:
                self._page_of[key] = p
                self._key_of[p] = key

    def release(self, owner: int) -> None:
        for p in self._owned.pop(owner, []):
            r = self._ref.get(p, 1) - 1
            if r > 0:
                self._ref[p] = r
                continue
            self._ref.pop(p, None)
            if p in self._key_of:
                self._lru[p] = None                                   # content stays valid: evictable cache entry
                self._lru.move_to_end(p)
            else:
                self._free.append(p)
        self._cached.pop(owner, None)

    def invalidate(self, owner: int) -> None:
        """The KV content of ``owner``'s pages cannot be trusted (its prefill failed): forget their prefix keys."""
        for p in self._owned.get(owner, []):
            if self._ref.get(p, 0) <= 1:
                key = self._key_of.pop(p, None)
                if key is not None:
                    self._page_of.pop(key, None)

    def owned(self, owner: int) -> List[int]:
        return list(self._owned.get(owner, []))

    def utilization(self) -> float:
        return 1.0 - self.free_pages / max(1, self.num_pages - 1)

    def cache_stats(self) -> Dict[str, float]:
        return {"cached_pages": len(self._key_of), "evictable_pages": len(self._lru), "hit_tokens": self.hit_tokens,
                "lookup_tokens": self.lookup_tokens,
                "hit_rate": self.hit_tokens / self.lookup_tokens if self.lookup_tokens else 0.0}
