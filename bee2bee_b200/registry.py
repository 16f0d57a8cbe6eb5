"""Node registry client (parity: /root/reference/bee2bee/registry.py:7-69).

Same env contract (``SUPABASE_URL|VITE_SUPABASE_URL`` + ``SUPABASE_ANON_KEY|VITE_...`` or
``BEE2BEE_ENTRYPOINT``), same upsert payload and headers, same offline mode when no
credentials exist.  In offline mode rows are additionally mirrored to
``$BEE2BEE_HOME/registry.json`` so ``bee2bee register`` is observable on a box without network."""
from __future__ import annotations

import os
from datetime import datetime, timezone
from typing import Any, Dict, List, Optional

from .utils import data_file, load_json, save_json

try:  # loguru is what the reference logs with; fall back to stdlib if absent
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("bee2bee")


class RegistryClient:
    def __init__(self, entrypoint_url: Optional[str] = None):
        env = os.environ.get
        self.supabase_url = env("VITE_SUPABASE_URL") or env("SUPABASE_URL")
        self.supabase_key = env("VITE_SUPABASE_ANON_KEY") or env("SUPABASE_ANON_KEY")
        self.entrypoint_url = entrypoint_url or env("BEE2BEE_ENTRYPOINT")
        self.api_url: Optional[str] = None
        self.headers: Dict[str, str] = {}
        if self.supabase_url and self.supabase_key:
            self.api_url = self.supabase_url.rstrip("/") + "/rest/v1/active_nodes"
            self.headers = {"apikey": self.supabase_key, "Authorization": f"Bearer {self.supabase_key}",
                            "Content-Type": "application/json", "Prefer": "resolution=merge-duplicates"}
        elif self.entrypoint_url:
            self.api_url = self.entrypoint_url.rstrip("/") + "/api/nodes/register"
            self.headers = {"Content-Type": "application/json"}
        self.enabled = self.api_url is not None
        if not self.enabled:
            logger.debug("No registry credentials: node runs in private/offline mode")

    @staticmethod
    def build_payload(peer_id: str, address: str, models: List[str], latency: float = 0.0, tag: str = "global",
                      region: str = "Auto", metrics: Optional[dict] = None) -> Dict[str, Any]:
        return {"peer_id": peer_id, "addr": address, "models": models, "latency_ms": latency, "region": region,
                "tag": tag, "metrics": metrics, "last_seen": datetime.now(timezone.utc).isoformat()}

    def _mirror_local(self, payload: Dict[str, Any]) -> None:
        path = data_file("registry.json")
        rows = load_json(path, {})
        rows[payload["peer_id"]] = payload
        save_json(path, rows)

    async def sync_node(self, peer_id: str, address: str, models: List[str], latency: float = 0.0,
                        tag: str = "global", region: str = "Auto", metrics: Optional[dict] = None) -> bool:
        payload = self.build_payload(peer_id, address, models, latency, tag, region, metrics)
        try:
            self._mirror_local(payload)
        except Exception:
            pass
        if not self.enabled:
            return False
        try:
            import httpx

            async with httpx.AsyncClient() as client:
                resp = await client.post(self.api_url, json=payload, headers=self.headers, timeout=5.0)
            if resp.status_code in (200, 201):
                return True
            logger.error(f"Registry sync failed: {resp.status_code} - {resp.text[:200]}")
        except Exception as exc:
            logger.error(f"Registry connection error: {exc}")
        return False

    @staticmethod
    def local_rows() -> Dict[str, Any]:
        return load_json(data_file("registry.json"), {})
