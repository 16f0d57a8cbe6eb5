"""Builds the local Hugging Face directory the *unmodified reference* loads for the baseline
arm (``bench.py --impl reference``): config + random-init bf16 weights + a synthetic
word-level tokenizer, produced with ``transformers``/``tokenizers`` only — none of this
repo's model or kernel code is involved.  (No network: there is no real checkpoint.)"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HF_CONFIGS = {
    "llama-3-8b": dict(model_type="llama", vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                       num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8, head_dim=128,
                       max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0,
                       tie_word_embeddings=False, bos_token_id=128000, eos_token_id=128001),
    "tiny-llama": dict(model_type="llama", vocab_size=512, hidden_size=256, intermediate_size=512,
                       num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=128,
                       max_position_embeddings=8192, rms_norm_eps=1e-5, rope_theta=500000.0,
                       tie_word_embeddings=False, bos_token_id=0, eos_token_id=1),
    "distilgpt2": dict(model_type="gpt2", vocab_size=50257, n_embd=768, n_layer=6, n_head=12, n_positions=1024,
                       bos_token_id=50256, eos_token_id=50256),
}


def build_reference_checkpoint(model: str, rank: int = 0, root: str = "") -> str:
    import torch
    from transformers import AutoConfig, AutoModelForCausalLM, PreTrainedTokenizerFast

    root = root or os.environ.get("B2B_REF_DIR", "/tmp/b2b_ref")
    path = os.path.join(root, model.replace("/", "_"))
    done = os.path.join(path, ".complete")
    if rank == 0 and not os.path.exists(done):
        os.makedirs(path, exist_ok=True)
        kw = dict(HF_CONFIGS[model])
        mt = kw.pop("model_type")
        cfg = AutoConfig.for_model(mt, **kw)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        torch.manual_seed(0)
        with torch.device(dev):
            m = AutoModelForCausalLM.from_config(cfg, dtype=torch.bfloat16)
        # an EOS that random weights will practically never sample keeps every request at K new tokens
        m.generation_config.eos_token_id = None
        m.generation_config.pad_token_id = 0
        m.save_pretrained(path, safe_serialization=True)
        del m
        from tokenizers import Tokenizer
        from tokenizers.models import WordLevel
        from tokenizers.pre_tokenizers import WhitespaceSplit

        V = kw["vocab_size"]
        vocab = {f"t{i}": i for i in range(V)}
        tok = Tokenizer(WordLevel(vocab, unk_token="t3"))
        tok.pre_tokenizer = WhitespaceSplit()
        fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="t3", pad_token="t0")
        fast.save_pretrained(path)
        with open(done, "w") as f:
            f.write(str(time.time()))
    else:
        t0 = time.time()
        while not os.path.exists(done):
            if time.time() - t0 > 1800:
                raise TimeoutError("rank 0 never finished writing the reference checkpoint")
            time.sleep(1.0)
    return path


if __name__ == "__main__":
    print(build_reference_checkpoint(sys.argv[1] if len(sys.argv) > 1 else "tiny-llama"))
