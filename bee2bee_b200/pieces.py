"""Pieces, in both senses.

1. *Byte pieces* -- fixed-size chunks of a blob with SHA-256 hashes, verify + reassemble,
   ``<hash>_<idx:08d>.part`` files (parity: /root/reference/bee2bee/pieces.py:7-32; the
   reference's piece_request/piece_data handlers are empty stubs, p2p_runtime.py:675-683 --
   here they are served, see ``p2p_runtime``).
2. *Layer pieces* -- the north-star meaning: a contiguous layer range of a model resident
   on one GPU.  ``LayerPiece`` describes one, ``plan_pieces`` splits a model over N peers, and
   ``save_piece_checkpoint`` / ``load_piece_checkpoint`` persist exactly one piece's tensors
   as hash-verified byte pieces (the framework's checkpoint/resume story: a peer restarts by
   re-reading only its own shard).
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass
from typing import Dict, List, Optional, Sequence

from .p2p import chunk_bytes, sha256_hex_bytes

DEFAULT_PIECE_SIZE = 4 << 20


# ------------------------------------------------------------------ byte pieces
def split_pieces(data: bytes, piece_size: int) -> List[bytes]:
    return chunk_bytes(data, piece_size)


def piece_hashes(pieces: Sequence[bytes]) -> List[str]:
    return [sha256_hex_bytes(p) for p in pieces]


def verify_and_reassemble(pieces: Sequence[bytes], hashes: Sequence[str]) -> bytes:
    if len(pieces) != len(hashes):
        raise ValueError("length_mismatch")
    for i, (chunk, want) in enumerate(zip(pieces, hashes)):
        if sha256_hex_bytes(chunk) != want:
            raise ValueError(f"hash_mismatch_at_{i}")
    return b"".join(pieces)


def save_pieces(folder: str, content_hash: str, pieces: Sequence[bytes]) -> List[str]:
    os.makedirs(folder, exist_ok=True)
    out = []
    for i, chunk in enumerate(pieces):
        path = os.path.join(folder, f"{content_hash}_{i:08d}.part")
        with open(path, "wb") as fh:
            fh.write(chunk)
        out.append(path)
    return out


def load_pieces(folder: str, content_hash: str) -> List[bytes]:
    """Inverse of ``save_pieces`` (the reference writes .part files but never reads them)."""
    names = sorted(n for n in os.listdir(folder) if n.startswith(content_hash + "_") and n.endswith(".part"))
    out = []
    for n in names:
        with open(os.path.join(folder, n), "rb") as fh:
            out.append(fh.read())
    return out


# ----------------------------------------------------------------- layer pieces
@dataclass
class LayerPiece:
    """Layers [start, end) of ``model`` hosted by peer ``rank`` on CUDA device ``device``."""
    model: str
    index: int
    start: int
    end: int
    first: bool
    last: bool
    rank: int = 0
    device: str = "cpu"
    peer_id: str = ""

    @property
    def layers(self) -> range:
        return range(self.start, self.end)

    def describe(self) -> Dict:
        return asdict(self)


def plan_pieces(model: str, n_layers: int, n_peers: int, devices: Optional[Sequence[str]] = None) -> List[LayerPiece]:
    from .models.config import split_layers

    ranges = split_layers(n_layers, n_peers)
    out = []
    for i, r in enumerate(ranges):
        dev = devices[i] if devices and i < len(devices) else f"cuda:{i}"
        out.append(LayerPiece(model, i, r.start, r.stop, i == 0, i == len(ranges) - 1, rank=i, device=dev))
    return out


def save_piece_checkpoint(folder: str, piece: LayerPiece, tensors: Dict, piece_size: int = DEFAULT_PIECE_SIZE) -> str:
    """Serialise one piece's tensors (safetensors bytes) into hash-addressed .part files +
    a manifest. Returns the manifest path."""
    from safetensors.torch import save as st_save

    blob = st_save({k: v.contiguous().cpu() for k, v in tensors.items()})
    chunks = split_pieces(blob, piece_size)
    content_hash = sha256_hex_bytes(blob)
    save_pieces(folder, content_hash, chunks)
    manifest = {"piece": piece.describe(), "content_hash": content_hash, "piece_size": piece_size,
                "hashes": piece_hashes(chunks), "bytes": len(blob), "tensors": sorted(tensors)}
    path = os.path.join(folder, f"piece_{piece.index:03d}.manifest.json")
    with open(path, "w") as fh:
        json.dump(manifest, fh, indent=1)
    return path


def load_piece_checkpoint(manifest_path: str) -> Dict:
    from safetensors.torch import load as st_load

    with open(manifest_path) as fh:
        manifest = json.load(fh)
    folder = os.path.dirname(manifest_path)
    chunks = load_pieces(folder, manifest["content_hash"])
    blob = verify_and_reassemble(chunks, manifest["hashes"])
    if sha256_hex_bytes(blob) != manifest["content_hash"]:
        raise ValueError("content_hash_mismatch")
    return st_load(blob)
