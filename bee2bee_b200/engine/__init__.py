"""Scheduler, paged KV, tokenizers and the per-GPU runner."""
from .core import Engine, Request, SamplingParams, TorchRunner  # noqa: F401
from .kv import PAGE, PageAllocator  # noqa: F401
from .tokenizer import ByteTokenizer, load_tokenizer, parse_transcript  # noqa: F401
