"""``replace_attention_layers`` (ref/utils/alter_attention.py:45-53).

In the reference this swaps every ``nn.MultiheadAttention`` for ``CustomMultiheadAttention`` so that
``attn_weights`` can scale the post-softmax weights of the last key columns.  The CUDA attention kernel
honours ``attn_weights`` natively (paella_b200/csrc/attention.cu), so here the call is a no-op that keeps
the notebook line ``replace_attention_layers(model)`` valid; the parameter holders are left in place so the
state-dict keys do not change.
"""
from torch import nn


class CustomMultiheadAttention(nn.MultiheadAttention):
    """Kept for import compatibility; never executed (attention runs in the CUDA library)."""


def replace_attention_layers(model):
    return model
