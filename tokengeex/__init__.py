"""Drop-in module name of the reference's Python binding (`import tokengeex`), served by
tokengeex_b200 (/root/reference/bindings/python/src/lib.rs:226-233 registers the same two names)."""
from tokengeex_b200.tokenizer import TokenGeeXError, Tokenizer  # noqa: F401

__all__ = ["Tokenizer", "TokenGeeXError"]
