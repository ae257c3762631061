from .attention import TransformerNet  # noqa: F401
