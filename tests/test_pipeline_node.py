"""BASELINE.json config 1 -- distilgpt2 split into 2 layer pieces on CPU, hidden states hop over
the loopback p2p_runtime, /generate on a 16-token synthetic prompt -- plus the legacy
coordinator/worker protocol (node.py) and the numpy MLP (intent of tests/test_model.py)."""
import asyncio
import json

import numpy as np
import pytest
import torch

from bee2bee_b200 import api as api_mod
from bee2bee_b200 import model as mlp
from bee2bee_b200 import protocol as P
from bee2bee_b200.models.config import resolve_config
from bee2bee_b200.models.torch_ref import TorchPiece
from bee2bee_b200.models.weights import init_random
from bee2bee_b200.node import Coordinator, TaskExecutor, gather_resources, node_client
from bee2bee_b200.p2p_runtime import P2PNode
from bee2bee_b200.parallel.cpu_pipeline import (MeshPipelineService, PieceHost, decode_tensor, encode_tensor,
                                                 piece_key)
from bee2bee_b200.pieces import plan_pieces


def test_tensor_payload_is_binary_and_exact():
    for dt in (torch.float32, torch.bfloat16, torch.int64):
        t = (torch.randn(2, 3, 5) * 10).to(dt)
        p = encode_tensor(t)
        assert set(p) == {"shape", "dtype", "b64"} and torch.equal(decode_tensor(p), t)
    big = encode_tensor(torch.randn(1, 16, 768))
    assert len(big["b64"]) < 16 * 768 * 4 * 1.4          # ~1.33 B/byte, vs >10 chars per value as JSON lists


@pytest.mark.parametrize("model,n_pieces", [("distilgpt2", 2), ("tiny-llama", 3)])
def test_two_piece_pipeline_over_loopback_mesh(model, n_pieces):
    async def go():
        nodes = [P2PNode(host="127.0.0.1", port=0, transport="ws") for _ in range(n_pieces)]   # real loopback sockets
        for n in nodes:
            await n.start()
        try:
            cfg = resolve_config(model)
            plan = plan_pieces(model, cfg.n_layers, n_pieces, devices=["cpu"] * n_pieces)
            for i in range(1, n_pieces):
                nodes[i].piece_hosts[piece_key(model, i)] = PieceHost(model, i, n_pieces)
                nodes[i].add_layer_piece(plan[i])
                await nodes[0].connect_bootstrap(nodes[i].addr)
            head = nodes[0]
            head.add_layer_piece(plan[0])
            while not all(nodes[i].peer_id in head.peers and head.peers[nodes[i].peer_id].get("hello_seen")
                          for i in range(1, n_pieces)):
                await asyncio.sleep(0.01)
            svc = MeshPipelineService(head, model, n_pieces, [n.peer_id for n in nodes[1:]])
            svc.bind_loop(asyncio.get_running_loop())
            await head.add_service(svc)
            topo = head.mesh_topology()
            assert topo[nodes[1].peer_id]["pieces"][0]["start"] == plan[1].start          # piece table gossiped in hello
            prompt = [(37 * i + 11) % cfg.vocab_size for i in range(16)]                  # 16-token synthetic prompt
            out = await svc.agenerate(prompt, 6, temperature=0.0)
            # oracle: the unsplit model, greedy
            t = init_random(cfg, range(cfg.n_layers), True, True)
            whole = TorchPiece(cfg, range(cfg.n_layers), True, True, t)
            cache, ids, ref = whole.new_cache(), list(prompt), []
            x = torch.tensor([ids]); pos = torch.arange(len(ids))[None]
            for step in range(6):
                logits = whole.forward(x, pos, cache, logits_last_only=True)[0, -1]
                tok = int(logits.argmax()); ref.append(tok)
                x = torch.tensor([[tok]]); pos = torch.tensor([[len(prompt) + step]])
            assert out == ref
            assert svc.hops == (n_pieces - 1) * 6 and svc.hop_bytes > 0
            assert all(not h.sessions for n in nodes[1:] for h in n.piece_hosts.values())   # KV released on every peer
            # the same pipeline behind the HTTP sidecar
            api_mod.node = head
            from httpx import ASGITransport, AsyncClient
            async with AsyncClient(transport=ASGITransport(app=api_mod.app), base_url="http://t") as c:
                r = (await c.post("/generate", json={"prompt": "hello mesh", "model": model, "max_new_tokens": 4,
                                                     "temperature": 0})).json()
                assert r["status"] == "ok" and r["metadata"]["tokens"] == 4 and r["text"].startswith("hello mesh")
                resp = await c.post("/generate", json={"prompt": "user: hi\nassistant:", "max_new_tokens": 4, "stream": True})
                lines = [json.loads(l) for l in resp.text.splitlines() if l]
                assert lines[-1] == {"done": True}
            # a remote requester reaches the pipeline through gen_request like any other provider
            res = await nodes[1].request_generation(head.peer_id, "abc", 3, model, timeout=30)
            assert res["tokens"] == 3
        finally:
            api_mod.node = None
            for n in nodes:
                await n.stop()

    asyncio.run(go())


def test_numpy_mlp_and_backward_matches_numerical_gradient():
    layers = mlp.random_mlp(8, 16, 4, 2)
    assert [l.W.shape for l in layers] == [(8, 16), (16, 4)] and layers[0].activation == "relu"
    x = np.random.default_rng(0).normal(size=(3, 8)).astype(np.float32)
    assert mlp.layer_forward(layers[0], x).shape == (3, 16)
    assert mlp.act_derivative(x, "gelu").shape == x.shape and mlp.act_derivative(x, "none").max() == 1.0
    rt = mlp.deserialize_layer(json.loads(json.dumps(mlp.serialize_layer(layers[1]))))
    assert np.allclose(rt.W, layers[1].W) and rt.activation == "none"
    for kind in ("relu", "gelu", "none"):
        layer = mlp.Layer(W=np.random.default_rng(1).normal(0, 0.5, (8, 5)).astype(np.float64),
                          b=np.zeros(5), activation=kind)
        xd = x.astype(np.float64)
        y, z = mlp.layer_forward_train(layer, xd)
        g = np.ones_like(y)
        dX, gW, gb = mlp.layer_backward(layer, xd, z, g)
        eps = 1e-6
        W2 = layer.W.copy(); W2[2, 3] += eps
        num = (mlp.act(xd @ W2 + layer.b, kind).sum() - y.sum()) / eps
        assert abs(num - gW[2, 3]) < 1e-3, kind
        x2 = xd.copy(); x2[1, 4] += eps
        num = (mlp.act(x2 @ layer.W + layer.b, kind).sum() - y.sum()) / eps
        assert abs(num - dX[1, 4]) < 1e-3, kind
    yt, zt = mlp.dense_forward_device(layers[0].W, layers[0].b, "relu", x, device="cpu")
    assert np.allclose(yt.numpy(), mlp.layer_forward(layers[0], x), atol=1e-6)


def test_task_executor_all_kinds():
    ex = TaskExecutor(device="cpu")
    layer = mlp.random_mlp(4, 6, 2, 2)[0]
    x = np.ones((2, 4), dtype=np.float32)
    out = ex.execute({"kind": P.TASK_LAYER_FORWARD, "layer": mlp.serialize_layer(layer), "x": x.tolist()})
    assert np.allclose(out["output"], mlp.layer_forward(layer, x))
    tr = ex.execute({"kind": P.TASK_LAYER_FORWARD_TRAIN, "layer": mlp.serialize_layer(layer), "x": x.tolist(), "cache_id": "c1"})
    bw = ex.execute({"kind": P.TASK_LAYER_BACKWARD, "cache_id": "c1", "upstream_grad": np.ones((2, 6)).tolist()})
    assert set(bw) == {"dX", "gW", "gb"} and np.array(bw["gW"]).shape == (4, 6)
    with pytest.raises(Exception, match="cache_missing"):
        ex.execute({"kind": P.TASK_LAYER_BACKWARD, "cache_id": "c1", "upstream_grad": [[0.0] * 6] * 2})
    with pytest.raises(Exception, match="unknown_task"):
        ex.execute({"kind": "nope"})
    with pytest.raises(Exception, match="onnx_support_missing|No module"):
        ex.execute({"kind": P.ONNX_LOAD, "path": "/nonexistent.onnx"})
    a = ex.execute({"kind": P.HF_PART_LOAD, "model_name": "tiny-gpt2", "start": 0, "end": 2})
    b = ex.execute({"kind": P.HF_PART_LOAD, "model_name": "tiny-gpt2", "start": 2, "end": 4})
    h = ex.execute({"kind": P.HF_PART_FORWARD, "model_id": a["model_id"], "text": "hello"})       # JSON list (wire compat)
    assert np.array(h["hidden"]).shape[-1] == 128
    lg = ex.execute({"kind": P.HF_PART_FORWARD, "model_id": b["model_id"], "hidden": h["hidden"], "binary": True})
    assert decode_tensor(lg["hidden_b64"]).shape[-1] == resolve_config("tiny-gpt2").vocab_size
    assert ex.execute({"kind": P.HF_UNLOAD, "model_id": a["model_id"]}) == {"ok": True}
    m = ex.execute({"kind": P.HF_LOAD, "model_name": "tiny-llama"})
    txt = ex.execute({"kind": P.HF_INFER, "model_id": m["model_id"], "prompt": "hi", "max_new_tokens": 3, "temperature": 0})
    assert txt["text"].startswith("hi")
    assert gather_resources()["cpu_count"] >= 1


def test_coordinator_drives_workers_pipeline_train_and_hf_split():
    async def go():
        coord = Coordinator(transport="inproc", name="coord-test")
        addr = await coord.start()
        workers = [asyncio.create_task(node_client(addr, f"w{i}", price=0.1 * i)) for i in range(2)]
        try:
            await coord.wait_for_workers(2)
            assert len(coord.list_nodes()) == 2 and coord.list_nodes()[0]["resources"]["cpu_count"] >= 1
            layers = mlp.random_mlp(6, 12, 3, 3, seed=1)
            x = np.random.default_rng(0).normal(size=(5, 6)).astype(np.float32)
            y = await coord.run_pipeline(layers, x)                     # RUN_PIPELINE
            ref = x
            for l in layers:
                ref = mlp.layer_forward(l, ref)
            assert np.allclose(y, ref, atol=1e-5)
            target = np.random.default_rng(1).normal(size=(5, 3)).astype(np.float32)
            losses = [await coord.run_train_step(layers, x, target, lr=0.5) for _ in range(25)]   # RUN_TRAIN_STEP
            assert losses[-1] < losses[0] * 0.9
            logits = await coord.run_hf_pipeline("tiny-gpt2", "split me", n_parts=2)              # RUN_HF_PIPELINE
            cfg = resolve_config("tiny-gpt2")
            t = init_random(cfg, range(cfg.n_layers), True, True)
            from bee2bee_b200.engine.tokenizer import load_tokenizer
            ids = load_tokenizer("tiny-gpt2", cfg.vocab_size, cfg.eos_token_id, cfg.bos_token_id).encode("split me")
            full = TorchPiece(cfg, range(cfg.n_layers), True, True, t).forward(
                torch.tensor([ids]), torch.arange(len(ids))[None], None)[0, -1]
            assert np.allclose(logits, full.numpy(), atol=1e-4)
        finally:
            for w in workers:
                w.cancel()
            await coord.stop()

    asyncio.run(go())
