"""Import-path shim for the reference's ``utils/alter_attention.py``."""
from paella_b200.alter_attention import CustomMultiheadAttention, replace_attention_layers  # noqa: F401
