"""RFC 5389 STUN binding client (parity: /root/reference/bee2bee/stun_client.py:10-180 -- dead
code there, wired into ``nat`` here).  Pure sockets, asyncio datagram endpoints; every
network call is skipped when ``BEE2BEE_OFFLINE`` is set."""
from __future__ import annotations

import asyncio
import os
import socket
import struct
from typing import Any, Dict, List, Optional, Tuple

from .utils import offline

MAGIC_COOKIE = 0x2112A442
BINDING_REQUEST, BINDING_SUCCESS = 0x0001, 0x0101
ATTR_MAPPED_ADDRESS, ATTR_XOR_MAPPED_ADDRESS = 0x0001, 0x0020


class STUNClient:
    DEFAULT_SERVERS: List[Tuple[str, int]] = [
        ("stun.l.google.com", 19302), ("stun1.l.google.com", 19302), ("stun2.l.google.com", 19302),
        ("stun3.l.google.com", 19302), ("stun4.l.google.com", 19302), ("stun.cloudflare.com", 3478),
        ("stun.nextcloud.com", 3478),
    ]

    def __init__(self, local_port: int = 0, local_ip: str = "0.0.0.0", servers: Optional[List[Tuple[str, int]]] = None):
        self.local_port, self.local_ip = local_port, local_ip
        self.servers = list(servers or self.DEFAULT_SERVERS)
        self._txid = b""

    def _generate_transaction_id(self) -> bytes:
        return os.urandom(12)

    def create_binding_request(self) -> bytes:
        self._txid = self._generate_transaction_id()
        return struct.pack("!HHI", BINDING_REQUEST, 0, MAGIC_COOKIE) + self._txid

    def parse_binding_response(self, data: bytes) -> Optional[Dict[str, Any]]:
        if len(data) < 20:
            return None
        mtype, mlen, cookie = struct.unpack("!HHI", data[:8])
        if mtype != BINDING_SUCCESS or cookie != MAGIC_COOKIE or (self._txid and data[8:20] != self._txid):
            return None
        off, end = 20, min(len(data), 20 + mlen)
        mapped = None
        while off + 4 <= end:
            atype, alen = struct.unpack("!HH", data[off:off + 4])
            val = data[off + 4:off + 4 + alen]
            if atype in (ATTR_XOR_MAPPED_ADDRESS, ATTR_MAPPED_ADDRESS) and len(val) >= 8 and val[1] == 0x01:
                port = struct.unpack("!H", val[2:4])[0]
                addr = struct.unpack("!I", val[4:8])[0]
                if atype == ATTR_XOR_MAPPED_ADDRESS:
                    port ^= MAGIC_COOKIE >> 16
                    addr ^= MAGIC_COOKIE
                cand = {"ip": socket.inet_ntoa(struct.pack("!I", addr)), "port": port,
                        "xor": atype == ATTR_XOR_MAPPED_ADDRESS}
                if mapped is None or cand["xor"]:
                    mapped = cand
            off += 4 + alen + (-alen % 4)
        return mapped

    async def query_server(self, server: str, port: int, timeout: float = 3.0) -> Optional[Dict[str, Any]]:
        if offline():
            return None
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()
        client = self

        class Proto(asyncio.DatagramProtocol):
            def datagram_received(self, data, addr):
                res = client.parse_binding_response(data)
                if res and not fut.done():
                    fut.set_result(res)

            def error_received(self, exc):
                if not fut.done():
                    fut.set_result(None)

        try:
            transport, _ = await loop.create_datagram_endpoint(Proto, local_addr=(self.local_ip, self.local_port),
                                                               remote_addr=(server, port))
        except Exception:
            return None
        try:
            transport.sendto(self.create_binding_request())
            res = await asyncio.wait_for(fut, timeout)
            if res:
                res["server"] = f"{server}:{port}"
                res["local_port"] = transport.get_extra_info("sockname")[1]
            return res
        except Exception:
            return None
        finally:
            transport.close()

    async def get_public_info(self, timeout_per_server: float = 2.0) -> Optional[Dict[str, Any]]:
        """Query all servers concurrently, first valid answer wins."""
        if offline():
            return None
        tasks = [asyncio.create_task(self.query_server(h, p, timeout_per_server)) for h, p in self.servers]
        try:
            for t in asyncio.as_completed(tasks):
                res = await t
                if res:
                    return res
        finally:
            for t in tasks:
                t.cancel()
        return None

    async def detect_nat_type(self) -> Dict[str, Any]:
        """Cone vs symmetric: compare the mapping seen by two different servers from one socket."""
        if offline():
            return {"type": "Unknown", "detail": "offline"}
        seen = []
        for host, port in self.servers[:3]:
            r = await self.query_server(host, port, 2.0)
            if r:
                seen.append(r)
            if len(seen) == 2:
                break
        if not seen:
            return {"type": "Blocked", "detail": "no STUN response (UDP filtered?)"}
        if len(seen) == 1:
            return {"type": "Unknown", "public_ip": seen[0]["ip"], "public_port": seen[0]["port"]}
        same = seen[0]["ip"] == seen[1]["ip"] and (self.local_port == 0 or seen[0]["port"] == seen[1]["port"])
        return {"type": "Cone" if same else "Symmetric", "public_ip": seen[0]["ip"], "public_port": seen[0]["port"],
                "mappings": seen}
