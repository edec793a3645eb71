"""Drop-in for lib/transformer.py: same class names; see transformer_wk.py for the implementation notes."""
from .transformer_wk import TransformerDecoderLayer, TransformerEncoderLayer, transformer_wk as transformer  # noqa: F401
