"""Concurrent multi-client load against the FastAPI sidecar (`POST /generate`), BASELINE config 5:
N clients x R requests each, a mix of buffered and streamed generations; reports request latency percentiles,
time to first streamed chunk, aggregate generated tokens/s (from the engine's own counters, `GET /metrics`) and
the KV-cache page utilisation sampled while the load runs.

    python -m bee2bee_b200 serve-hf --random-weights --model zephyr-7b-beta --api-port 8000 &
    python tools/load_test.py --url http://127.0.0.1:8000 --clients 16 --requests 4 --max-new-tokens 64

`--spawn MODEL` starts (and stops) the server itself.  Parity: the reference has no load tool; its sidecar serves one
blocking `model.generate` per request (api.py:190-245), which is what `bench.py --impl reference` measures.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import signal
import socket
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pct(xs, p):
    if not xs:
        return None
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(p / 100.0 * (len(xs) - 1))))]


async def client(idx: int, args, lat, ttft, errors):
    import httpx

    async with httpx.AsyncClient(timeout=httpx.Timeout(args.timeout, connect=5.0)) as cli:
        for r in range(args.requests):
            prompt = f"user: client {idx} request {r}: " + " ".join(f"w{(idx * 131 + r * 17 + k) % 997}" for k in range(args.prompt_words))
            body = {"prompt": prompt, "max_new_tokens": args.max_new_tokens, "temperature": args.temperature}
            stream = (idx + r) % 2 == 1 and not args.no_stream
            t0 = time.perf_counter()
            try:
                if stream:
                    first = None
                    async with cli.stream("POST", args.url + "/generate", json=dict(body, stream=True)) as resp:
                        async for line in resp.aiter_lines():
                            if line and first is None:
                                first = time.perf_counter() - t0
                    if first is not None:
                        ttft.append(first)
                else:
                    d = (await cli.post(args.url + "/generate", json=body)).json()
                    if d.get("status") != "ok":
                        errors.append(str(d)[:200])
                lat.append(time.perf_counter() - t0)
            except Exception as e:          # keep the other clients going
                errors.append(repr(e)[:200])


async def sampler(args, kv, stop):
    import httpx

    async with httpx.AsyncClient(timeout=5.0) as cli:
        while not stop.is_set():
            try:
                m = (await cli.get(args.url + "/metrics")).json()
                for svc in m.values():
                    kv.append((svc.get("kv_utilization", 0.0), svc.get("running", 0), svc.get("waiting", 0)))
            except Exception:
                pass
            await asyncio.sleep(0.05)


async def run(args):
    import httpx

    async with httpx.AsyncClient(timeout=10.0) as cli:
        m0 = (await cli.get(args.url + "/metrics")).json()
    lat, ttft, errors, kv = [], [], [], []
    stop = asyncio.Event()
    samp = asyncio.create_task(sampler(args, kv, stop))
    t0 = time.perf_counter()
    await asyncio.gather(*[client(i, args, lat, ttft, errors) for i in range(args.clients)])
    wall = time.perf_counter() - t0
    stop.set()
    await samp
    async with httpx.AsyncClient(timeout=10.0) as cli:
        m1 = (await cli.get(args.url + "/metrics")).json()
    tok = sum(v.get("tokens_generated", 0) for v in m1.values()) - sum(v.get("tokens_generated", 0) for v in m0.values())
    steps = sum(v.get("decode_steps", 0) for v in m1.values()) - sum(v.get("decode_steps", 0) for v in m0.values())
    out = {"clients": args.clients, "requests": len(lat), "errors": len(errors), "wall_s": round(wall, 3),
           "generated_tokens": tok, "tokens_per_s": round(tok / wall, 1) if wall > 0 else None,
           "decode_steps": steps, "mean_batch_per_step": round(tok / steps, 2) if steps else None,
           "latency_s": {"p50": pct(lat, 50), "p90": pct(lat, 90), "p99": pct(lat, 99)},
           "stream_ttft_s": {"p50": pct(ttft, 50), "p90": pct(ttft, 90)},
           "kv_utilization": {"max": max((k[0] for k in kv), default=None),
                              "mean": statistics.fmean(k[0] for k in kv) if kv else None},
           "max_running": max((k[1] for k in kv), default=None), "max_waiting": max((k[2] for k in kv), default=None)}
    for name, v in m1.items():
        h0 = (m0.get(name) or {}).get("host_ms", {})
        out.setdefault("engine_host_ms", {})[name] = {k: round(x - h0.get(k, 0.0), 1) for k, x in (v.get("host_ms") or {}).items()}
        out.setdefault("engine_device_ms", {})[name] = {k: {"count": t["count"], "total_ms": round(t["total_ms"], 1)}
                                                         for k, t in (v.get("trace") or {}).items()}
        out.setdefault("engine_ttft_ms", {})[name] = v.get("ttft_ms")
        out.setdefault("engine_busy_s", {})[name] = round(v.get("uptime_s", 0) and v.get("tokens_generated", 0) / max(v.get("tokens_per_s", 1e-9), 1e-9), 3)
    if errors:
        out["first_errors"] = errors[:3]
    print(json.dumps(out))
    return out


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--url", default="")
    ap.add_argument("--spawn", default="", help="model to serve with `python -m bee2bee_b200 serve-hf` for the duration of the test")
    ap.add_argument("--pieces", type=int, default=1)
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--max-seq-len", type=int, default=1024)
    ap.add_argument("--clients", type=int, default=16)
    ap.add_argument("--requests", type=int, default=4)
    ap.add_argument("--max-new-tokens", type=int, default=64)
    ap.add_argument("--prompt-words", type=int, default=24)
    ap.add_argument("--temperature", type=float, default=0.7)
    ap.add_argument("--timeout", type=float, default=300.0)
    ap.add_argument("--no-stream", action="store_true")
    args = ap.parse_args(argv)
    proc = None
    if args.spawn:
        import httpx

        port = free_port()
        args.url = f"http://127.0.0.1:{port}"
        env = dict(os.environ, BEE2BEE_OFFLINE="1")
        proc = subprocess.Popen([sys.executable, "-m", "bee2bee_b200", "serve-hf", "--random-weights", "--model", args.spawn, "--pieces", str(args.pieces),
                                 "--api-port", str(port), "--max-batch", str(args.max_batch), "--max-seq-len", str(args.max_seq_len)],
                                env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT, start_new_session=True)
        t0 = time.time()
        while True:
            try:
                if httpx.get(args.url + "/", timeout=2).json().get("models"):
                    break
            except Exception:
                pass
            if proc.poll() is not None or time.time() - t0 > 600:
                raise SystemExit("server did not come up")
            time.sleep(1)
    try:
        return asyncio.run(run(args))
    finally:
        if proc is not None:
            os.killpg(proc.pid, signal.SIGTERM)      # exactly the process group this script started
            try:
                proc.wait(20)
            except subprocess.TimeoutExpired:
                os.killpg(proc.pid, signal.SIGKILL)


if __name__ == "__main__":
    main()
