"""Client-side demo (parity: /root/reference/examples/p2p_request_demo.py): join a mesh, list
providers, pick the cheapest one for a model and stream a generation from it.

    python -m bee2bee_b200 serve-hf --model tiny-llama --api-port 8000      # terminal 1 (prints its ws:// address)
    python examples/p2p_request_demo.py ws://127.0.0.1:<port> tiny-llama     # terminal 2
"""
import os
os.environ.setdefault("B2B_ALLOW_RANDOM_WEIGHTS", "1")     # no checkpoints offline: random-init weights
import asyncio
import sys

from bee2bee_b200.p2p_runtime import P2PNode


class Bee2BeeP2PClient:
    def __init__(self, bootstrap: str):
        self.bootstrap, self.node = bootstrap, P2PNode(host="127.0.0.1", port=0)

    async def __aenter__(self):
        await self.node.start()
        await self.node.connect_bootstrap(self.bootstrap)
        for _ in range(100):
            if self.node.providers:
                break
            await asyncio.sleep(0.05)
        return self

    async def __aexit__(self, *exc):
        await self.node.stop()

    async def generate(self, model: str, prompt: str, max_new_tokens: int = 64, stream: bool = True) -> str:
        picked = self.node.pick_provider(model)
        if picked is None:
            raise RuntimeError(f"no provider serves {model!r}; known: {self.node.list_providers()}")
        pid, meta = picked
        print(f"provider {pid} price={meta.get('price_per_token')} svc={meta['_svc_name']}")
        on_chunk = (lambda t: print(t, end="", flush=True)) if stream else None
        res = await self.node.request_generation(pid, prompt, max_new_tokens, model, on_chunk=on_chunk)
        return res.get("text", "")


async def main():
    bootstrap, model = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "distilgpt2")
    async with Bee2BeeP2PClient(bootstrap) as c:
        await c.generate(model, "user: tell me about NVLink\nassistant:")
        print()


if __name__ == "__main__":
    asyncio.run(main())
