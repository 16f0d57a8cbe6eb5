"""Model families (GPT-2, Llama, Mistral, Gemma-2): configs, weights, torch oracle, native pieces."""
from .config import PRESETS, ModelConfig, resolve_config, split_layers  # noqa: F401
