"""Same import surface as the reference's `llama/__init__.py:4-6`."""
from .model import ModelArgs, Transformer
from .tokenizer import SyntheticTokenizer, Tokenizer

__all__ = ["ModelArgs", "Transformer", "Tokenizer", "SyntheticTokenizer"]
