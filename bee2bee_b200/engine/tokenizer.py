"""Tokenizers.  There is no network, hence no vocab files for the named checkpoints: the
default is a deterministic byte-level tokenizer that works with any vocab >= 260.  A real
Hugging Face tokenizer is used when the model directory ships one (same call the
reference makes, /root/reference/bee2bee/hf.py:24).

Also hosts the prompt helpers of the reference's streaming path: parsing
``user: / assistant:`` transcripts into chat messages and cutting generated text at stop
words (/root/reference/bee2bee/hf.py:54-71,112-133).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

STOP_WORDS = ("user:", "assistant:", "<|im_end|>", "<s>", "</s>")


class ByteTokenizer:
    """ids 0..3 = <pad>, <bos>, <eos>, <unk>; byte b -> 4 + b.  Ids beyond 259 decode to ''."""

    pad_id, bos_id, eos_id, unk_id = 0, 1, 2, 3
    OFFSET = 4

    def __init__(self, vocab_size: int = 260, eos_id: Optional[int] = None, bos_id: Optional[int] = None):
        if vocab_size < 260:
            # tiny test vocabularies: fold bytes into the available range
            self.fold = max(1, vocab_size - self.OFFSET)
        else:
            self.fold = 256
        self.vocab_size = vocab_size
        if eos_id is not None and 0 <= eos_id < vocab_size:
            self.eos_id = eos_id
        if bos_id is not None and 0 <= bos_id < vocab_size:
            self.bos_id = bos_id

    def encode(self, text: str, add_bos: bool = True) -> List[int]:
        ids = [self.OFFSET + (b % self.fold) for b in text.encode("utf-8")]
        return ([self.bos_id] if add_bos else []) + ids

    def decode(self, ids: Iterable[int], skip_special: bool = True) -> str:
        bs = bytearray()
        for i in ids:
            i = int(i)
            if i in (self.bos_id, self.eos_id, self.pad_id) and (i < self.OFFSET or i >= self.OFFSET + 256):
                if not skip_special:
                    bs.extend({self.bos_id: b"<bos>", self.eos_id: b"<eos>"}.get(i, b"<pad>"))
                continue
            if self.OFFSET <= i < self.OFFSET + 256 and self.fold == 256:
                bs.append(i - self.OFFSET)
            elif self.OFFSET <= i < self.OFFSET + self.fold:
                bs.append(32 + (i - self.OFFSET) % 95)          # printable stand-in for folded vocabularies
            elif not skip_special and i < self.OFFSET:
                bs.extend(("<pad>", "<bos>", "<eos>", "<unk>")[i].encode())
            elif i >= self.OFFSET + 256:
                # random-init models emit arbitrary ids; map them onto printable bytes so streams are visible
                bs.append(32 + (i % 95))
        return bs.decode("utf-8", errors="replace")

    def apply_chat_template(self, messages: Sequence[Dict[str, str]], add_generation_prompt: bool = True) -> str:
        out = "".join(f"<|{m['role']}|>\n{m['content']}\n" for m in messages)
        return out + ("<|assistant|>\n" if add_generation_prompt else "")


class HFTokenizerAdapter:
    def __init__(self, tok):
        self.tok = tok
        self.vocab_size = len(tok)
        self.eos_id = tok.eos_token_id if tok.eos_token_id is not None else -1
        self.bos_id = tok.bos_token_id if tok.bos_token_id is not None else -1

    def encode(self, text: str, add_bos: bool = True) -> List[int]:
        return list(self.tok.encode(text, add_special_tokens=add_bos))

    def decode(self, ids: Iterable[int], skip_special: bool = True) -> str:
        return self.tok.decode(list(ids), skip_special_tokens=skip_special)

    def apply_chat_template(self, messages, add_generation_prompt: bool = True) -> str:
        try:
            return self.tok.apply_chat_template(list(messages), tokenize=False,
                                                add_generation_prompt=add_generation_prompt)
        except Exception:
            return ByteTokenizer.apply_chat_template(self, messages, add_generation_prompt)  # type: ignore[arg-type]


def load_tokenizer(model: str, vocab_size: int, eos_id: int = -1, bos_id: int = -1):
    if os.path.isdir(model) and any(os.path.exists(os.path.join(model, f))
                                    for f in ("tokenizer.json", "tokenizer.model", "vocab.json")):
        try:
            from transformers import AutoTokenizer
            return HFTokenizerAdapter(AutoTokenizer.from_pretrained(model))
        except Exception:
            pass
    return ByteTokenizer(vocab_size, eos_id if eos_id >= 0 else None, bos_id if bos_id >= 0 else None)


def parse_transcript(prompt: str) -> List[Dict[str, str]]:
    """``user: hi\\nassistant: hello\\nuser: ...`` -> chat messages.  Lines without a role prefix
    continue the previous message; a prompt with no roles at all is a single user turn."""
    messages: List[Dict[str, str]] = []
    for line in prompt.split("\n"):
        stripped = line.strip()
        low = stripped.lower()
        role = None
        for r in ("user", "assistant", "system"):
            if low.startswith(r + ":"):
                role, stripped = r, stripped[len(r) + 1:].strip()
                break
        if role is not None:
            messages.append({"role": role, "content": stripped})
        elif messages:
            messages[-1]["content"] += ("\n" if messages[-1]["content"] else "") + line
        elif stripped:
            messages.append({"role": "user", "content": line})
    if messages and messages[-1]["role"] == "assistant" and not messages[-1]["content"].strip():
        messages.pop()          # trailing "assistant:" is just the generation cue
    return messages or [{"role": "user", "content": prompt}]


def cut_at_stop_words(text: str, stop_words: Sequence[str] = STOP_WORDS) -> Tuple[str, bool]:
    """Returns (text up to the first stop word, whether one was hit)."""
    low = text.lower()
    cut = min((i for i in (low.find(w.lower()) for w in stop_words) if i >= 0), default=-1)
    return (text[:cut], True) if cut >= 0 else (text, False)
