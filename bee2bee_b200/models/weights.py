"""Weight containers: deterministic random init (no network -> no checkpoints), Hugging
Face directory load/save (safetensors, HF tensor names) and the per-piece slicing that
gives the reference's "piece" its north-star meaning: a contiguous layer range resident on
one GPU (the reference prototype re-loads the *whole* model on every worker and runs a
slice, /root/reference/bee2bee/hf.py:180-205 — here only the slice is ever materialised).
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import Dict, Iterable, List, Optional

import torch

from .config import ModelConfig

Tensors = Dict[str, torch.Tensor]


def _seed_for(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:7], "little")


def _randn(name: str, shape, std: float, seed: int, device, dtype) -> torch.Tensor:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(_seed_for(name, seed))
    n = 1
    for s in shape:
        n *= s
    if dev.type == "cuda" and n > (1 << 26):
        # generate big matrices row-block-wise straight in the target dtype to bound scratch memory
        out = torch.empty(shape, device=dev, dtype=dtype)
        rows = max(1, (1 << 26) // shape[-1])
        for r0 in range(0, shape[0], rows):
            r1 = min(shape[0], r0 + rows)
            out[r0:r1] = (torch.randn((r1 - r0, *shape[1:]), generator=g, device=dev, dtype=torch.float32) * std).to(dtype)
        return out
    return (torch.randn(shape, generator=g, device=dev, dtype=torch.float32) * std).to(dtype)


def layer_tensor_specs(cfg: ModelConfig):
    """(name, shape, kind) for one layer; kind in {"w", "norm", "bias"}."""
    h, f, q, kv = cfg.hidden_size, cfg.ffn_size, cfg.q_dim, cfg.kv_dim
    specs = [("ln1_w", (h,), "norm"), ("wq", (q, h), "w"), ("wk", (kv, h), "w"), ("wv", (kv, h), "w"),
             ("wo", (h, q), "w"), ("ln2_w", (h,), "norm")]
    if cfg.norm == "ln":
        specs += [("ln1_b", (h,), "bias"), ("ln2_b", (h,), "bias")]
    if cfg.bias:
        specs += [("bq", (q,), "bias"), ("bk", (kv,), "bias"), ("bv", (kv,), "bias"), ("bo", (h,), "bias")]
    if cfg.post_norms:
        specs += [("post_attn_w", (h,), "norm"), ("post_ffn_w", (h,), "norm")]
    if cfg.glu:
        specs += [("w_gate", (f, h), "w"), ("w_up", (f, h), "w"), ("w_down", (h, f), "w")]
    else:
        specs += [("w_up", (f, h), "w"), ("w_down", (h, f), "w")]
        if cfg.bias:
            specs += [("b_up", (f,), "bias"), ("b_down", (h,), "bias")]
    return specs


def init_random(cfg: ModelConfig, layers: Iterable[int], first: bool, last: bool, device="cpu",
                dtype=torch.float32, seed: int = 0, std: float = 0.02) -> Tensors:
    """Random weights; every tensor is seeded by its own name so any piece split of the same
    (cfg, seed) materialises bit-identical tensors on the same device type."""
    t: Tensors = {}
    norm_fill = 0.0 if cfg.gemma_norm else 1.0

    def make(name, shape, kind):
        if kind == "w":
            return _randn(name, shape, std, seed, device, dtype)
        if kind == "norm":
            # slightly perturbed so gamma-folding bugs are visible in tests
            return (norm_fill + _randn(name, shape, 0.05, seed, device, torch.float32)).to(dtype)
        return _randn(name, shape, 0.01, seed, device, dtype)

    if first or (last and cfg.tie_embeddings):
        t["embed"] = make("embed", (cfg.vocab_size, cfg.hidden_size), "w")
        if cfg.rope_theta <= 0:
            t["pos_embed"] = make("pos_embed", (cfg.max_position, cfg.hidden_size), "w")
    for l in layers:
        for name, shape, kind in layer_tensor_specs(cfg):
            t[f"l{l}.{name}"] = make(f"l{l}.{name}", shape, kind)
    if last:
        t["final_norm_w"] = make("final_norm_w", (cfg.hidden_size,), "norm")
        if cfg.norm == "ln":
            t["final_norm_b"] = make("final_norm_b", (cfg.hidden_size,), "bias")
        if not cfg.tie_embeddings:
            t["lm_head"] = make("lm_head", (cfg.vocab_size, cfg.hidden_size), "w")
    return t


# ------------------------------------------------------------------- HF names
def hf_name_map(cfg: ModelConfig, layers: Iterable[int], first: bool, last: bool) -> Dict[str, str]:
    """our tensor name -> HF checkpoint tensor name (GPT-2 fused/transposed handled by the loader)."""
    m: Dict[str, str] = {}
    if cfg.family == "gpt2":
        if first or last:
            m["embed"] = "transformer.wte.weight"
        if first:
            m["pos_embed"] = "transformer.wpe.weight"
        for l in layers:
            p = f"transformer.h.{l}."
            m.update({f"l{l}.ln1_w": p + "ln_1.weight", f"l{l}.ln1_b": p + "ln_1.bias",
                      f"l{l}.ln2_w": p + "ln_2.weight", f"l{l}.ln2_b": p + "ln_2.bias",
                      f"l{l}.wo": p + "attn.c_proj.weight", f"l{l}.bo": p + "attn.c_proj.bias",
                      f"l{l}.w_up": p + "mlp.c_fc.weight", f"l{l}.b_up": p + "mlp.c_fc.bias",
                      f"l{l}.w_down": p + "mlp.c_proj.weight", f"l{l}.b_down": p + "mlp.c_proj.bias",
                      f"l{l}.__c_attn_w": p + "attn.c_attn.weight", f"l{l}.__c_attn_b": p + "attn.c_attn.bias"})
        if last:
            m["final_norm_w"] = "transformer.ln_f.weight"
            m["final_norm_b"] = "transformer.ln_f.bias"
        return m
    if first or (last and cfg.tie_embeddings):
        m["embed"] = "model.embed_tokens.weight"
    for l in layers:
        p = f"model.layers.{l}."
        m.update({f"l{l}.ln1_w": p + "input_layernorm.weight", f"l{l}.wq": p + "self_attn.q_proj.weight",
                  f"l{l}.wk": p + "self_attn.k_proj.weight", f"l{l}.wv": p + "self_attn.v_proj.weight",
                  f"l{l}.wo": p + "self_attn.o_proj.weight", f"l{l}.w_gate": p + "mlp.gate_proj.weight",
                  f"l{l}.w_up": p + "mlp.up_proj.weight", f"l{l}.w_down": p + "mlp.down_proj.weight"})
        if cfg.post_norms:
            m.update({f"l{l}.post_attn_w": p + "post_attention_layernorm.weight",
                      f"l{l}.ln2_w": p + "pre_feedforward_layernorm.weight",
                      f"l{l}.post_ffn_w": p + "post_feedforward_layernorm.weight"})
        else:
            m[f"l{l}.ln2_w"] = p + "post_attention_layernorm.weight"
    if last:
        m["final_norm_w"] = "model.norm.weight"
        if not cfg.tie_embeddings:
            m["lm_head"] = "lm_head.weight"
    return m


def save_hf_dir(path: str, cfg: ModelConfig, tensors: Tensors, dtype=torch.bfloat16, shard_gb: float = 4.0) -> None:
    """Write ``config.json`` + safetensors shards with HF names (whole model: all layers)."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    hf_cfg = cfg.to_hf_dict()
    hf_cfg["torch_dtype"] = str(dtype).replace("torch.", "")
    hf_cfg["dtype"] = hf_cfg["torch_dtype"]
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_cfg, f, indent=1)
    names = hf_name_map(cfg, range(cfg.n_layers), True, True)
    out: Tensors = {}
    for ours, hf in names.items():
        if ours.endswith("__c_attn_w"):
            l = ours.split(".")[0]
            w = torch.cat([tensors[f"{l}.wq"], tensors[f"{l}.wk"], tensors[f"{l}.wv"]], 0)
            out[hf] = w.t().contiguous()
        elif ours.endswith("__c_attn_b"):
            l = ours.split(".")[0]
            out[hf] = torch.cat([tensors[f"{l}.bq"], tensors[f"{l}.bk"], tensors[f"{l}.bv"]], 0)
        elif cfg.family == "gpt2" and ours.split(".")[-1] in ("wo", "w_up", "w_down"):
            out[hf] = tensors[ours].t().contiguous()      # Conv1D stores [in, out]
        else:
            out[hf] = tensors[ours]
    shards: List[Tensors] = [{}]
    size, limit = 0, shard_gb * (1 << 30)
    for k, v in out.items():
        v = v.to(dtype).contiguous().cpu()
        nbytes = v.numel() * v.element_size()
        if size + nbytes > limit and shards[-1]:
            shards.append({})
            size = 0
        shards[-1][k] = v
        size += nbytes
    if len(shards) == 1:
        save_file(shards[0], os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
    else:
        index = {"metadata": {}, "weight_map": {}}
        for i, sh in enumerate(shards):
            fn = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
            save_file(sh, os.path.join(path, fn), metadata={"format": "pt"})
            for k in sh:
                index["weight_map"][k] = fn
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump(index, f)


def load_hf_dir(path: str, cfg: ModelConfig, layers: Iterable[int], first: bool, last: bool, device="cpu",
                dtype=torch.float32) -> Optional[Tensors]:
    """Load only the tensors of one piece from a local HF directory. None if no weights there."""
    from safetensors import safe_open

    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        return None
    where: Dict[str, str] = {}
    for fn in files:
        with safe_open(os.path.join(path, fn), framework="pt") as f:
            for k in f.keys():
                where[k] = fn
    layers = list(layers)
    names = hf_name_map(cfg, layers, first, last)
    handles = {}

    def get(hf: str) -> torch.Tensor:
        if hf not in where and hf.startswith("transformer."):
            hf = hf[len("transformer."):]           # some GPT-2 checkpoints drop the prefix
        fn = where[hf]
        if fn not in handles:
            handles[fn] = safe_open(os.path.join(path, fn), framework="pt")
        return handles[fn].get_tensor(hf)

    out: Tensors = {}
    for ours, hf in names.items():
        if ours.endswith("__c_attn_w"):
            l = ours.split(".")[0]
            w = get(hf).t().contiguous()
            q, k, v = w.split([cfg.q_dim, cfg.kv_dim, cfg.kv_dim], 0)
            out[f"{l}.wq"], out[f"{l}.wk"], out[f"{l}.wv"] = q, k, v
        elif ours.endswith("__c_attn_b"):
            l = ours.split(".")[0]
            q, k, v = get(hf).split([cfg.q_dim, cfg.kv_dim, cfg.kv_dim], 0)
            out[f"{l}.bq"], out[f"{l}.bk"], out[f"{l}.bv"] = q, k, v
        elif cfg.family == "gpt2" and ours.split(".")[-1] in ("wo", "w_up", "w_down"):
            out[ours] = get(hf).t().contiguous()
        else:
            out[ours] = get(hf)
    return {k: v.to(device=device, dtype=dtype).contiguous() for k, v in out.items()}


class WeightsUnavailable(RuntimeError):
    """No checkpoint for ``model`` and random initialisation was not asked for."""


#: model name -> "checkpoint" | "random-init" (what the last load of that name actually used; service metadata)
WEIGHT_SOURCE: Dict[str, str] = {}


def random_weights_allowed(model: str) -> bool:
    """Random initialisation is OPT-IN: the synthetic test presets, or ``B2B_ALLOW_RANDOM_WEIGHTS=1``
    (benchmarks, smoke runs, CI -- there is no network in the build environment, hence no checkpoints).  Without it a
    node must not announce ``llama-3-8b`` to the mesh and serve noise; the reference would fail to load instead
    (/root/reference/bee2bee/hf.py:23-32)."""
    name = os.path.basename(os.path.normpath(model)).lower() if model else ""
    return name.startswith(("tiny-", "mini-")) or os.environ.get("B2B_ALLOW_RANDOM_WEIGHTS", "0") == "1"


def load_or_init(model: str, cfg: ModelConfig, layers: Iterable[int], first: bool, last: bool, device="cpu",
                 dtype=torch.float32, seed: int = 0) -> Tensors:
    layers = list(layers)
    if model and os.path.isdir(model):
        t = load_hf_dir(model, cfg, layers, first, last, device, dtype)
        if t is not None:
            WEIGHT_SOURCE[model] = "checkpoint"
            return t
    if model and not random_weights_allowed(model) and not random_weights_allowed(cfg.name):
        raise WeightsUnavailable(
            f"no safetensors checkpoint for '{model}' (not a local Hugging Face directory). Point --model at a local "
            f"directory, or opt into random-init weights with --random-weights / B2B_ALLOW_RANDOM_WEIGHTS=1 "
            f"(benchmarks and smoke tests only: the node would serve noise).")
    if model:
        if WEIGHT_SOURCE.get(model) != "random-init" and not os.path.basename(model).lower().startswith(("tiny-", "mini-")):
            import logging
            logging.getLogger("bee2bee_b200").warning(
                "model '%s': RANDOM-INIT weights (no checkpoint found; allowed by B2B_ALLOW_RANDOM_WEIGHTS)", model)
        WEIGHT_SOURCE[model] = "random-init"
    return init_random(cfg, layers, first, last, device, dtype, seed)
