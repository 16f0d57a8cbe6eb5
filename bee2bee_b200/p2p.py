"""Join links and content hashing (parity: /root/reference/bee2bee/p2p.py:8-52).

``coithub.org://join?network=<n>&model=<m>&hash=<sha256>&bootstrap=<urlsafe-b64, unpadded>...``
The parser also accepts ``coithub://`` and ``p2pnet://`` (the reference's runtime routes
``p2pnet://`` links to a parser that rejects them -- SURVEY.md section 8)."""
from __future__ import annotations

import base64
import hashlib
from typing import Any, Dict, Iterable, List
from urllib.parse import parse_qs, quote, urlparse

SCHEME = "coithub.org"
ACCEPTED_SCHEMES = ("coithub.org", "coithub", "p2pnet")


def _b64e(text: str) -> str:
    return base64.urlsafe_b64encode(text.encode()).decode().rstrip("=")


def _b64d(text: str) -> str:
    return base64.urlsafe_b64decode(text + "=" * (-len(text) % 4)).decode()


def generate_join_link(network: str, model: str, hash_hex: str, bootstrap: Iterable[str]) -> str:
    fields = [f"network={network}", f"model={model}", f"hash={hash_hex}"]
    fields += [f"bootstrap={_b64e(addr)}" for addr in bootstrap]
    return f"{SCHEME}://join?" + "&".join(fields)


def parse_join_link(link: str) -> Dict[str, Any]:
    scheme, sep, rest = link.partition("://")
    if not sep or scheme not in ACCEPTED_SCHEMES:
        raise ValueError("invalid_link")
    u = urlparse("x://" + rest)             # urlparse chokes on dots in custom schemes on some versions
    if u.netloc != "join":
        raise ValueError("invalid_link")
    q = parse_qs(u.query)
    first = lambda k: (q.get(k) or [None])[0]
    return {"network": first("network"), "model": first("model"), "hash": first("hash"),
            "bootstrap": [_b64d(b) for b in q.get("bootstrap", []) if b]}


def registration_url(join_link: str, region: str, tag: str, api_port: int) -> str:
    """Web hand-off URL printed at start-up (parity: p2p_runtime.py:924-926)."""
    return (f"https://coithub.org/register?link={quote(join_link, safe='')}&region={quote(str(region))}"
            f"&tag={quote(str(tag))}&api_port={api_port}")


def sha256_hex_bytes(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()


def chunk_bytes(data: bytes, piece_size: int) -> List[bytes]:
    if piece_size <= 0:
        raise ValueError("piece_size must be positive")
    view = memoryview(data)
    return [bytes(view[i:i + piece_size]) for i in range(0, len(data), piece_size)]


def bitfield_from_pieces(total_pieces: int, have_indices: Iterable[int]) -> List[int]:
    bits = [0] * total_pieces
    for i in have_indices:
        if 0 <= i < total_pieces:
            bits[i] = 1
    return bits
