"""trax_b200 — B200-native (sm_100a) implementation of ONE hot path of google/trax:
`trax.layers.research.efficient_attention.LSHSelfAttention` forward + backward.

Layout: csrc/ (CUDA kernels + C ABI, built to liblsh_attn_b200.so), _lib.py (ctypes binding),
ops.py (stage-level wrappers), lsh_attention.py (the layer with the reference's interface),
pure_lsh_attention.py (its weight-less core, `PureLSHSelfAttention`, and `PureLSHSelfAttentionWrapper` around it),
self_attention.py (`SelfAttention`, both `share_qk` settings, on the same core), predict.py (fast inference, `mode='predict'`, of
all four), reversible.py (`ReversibleHalfResidual` around the layer), dp.py (multi-GPU drivers).
Importing the package does not need a GPU; calling anything does, and raises otherwise.
"""
from trax_b200.lsh_attention import (LSHSelfAttention, ShapeDtype, host_io_bytes,  # noqa: F401
                                     set_async_host_io, set_reuse_forward_upload, set_weight_grad_allreduce, synchronize)
from trax_b200.pure_lsh_attention import PureLSHSelfAttention, PureLSHSelfAttentionWrapper  # noqa: F401
from trax_b200.reversible import ReversibleHalfResidual  # noqa: F401
from trax_b200.self_attention import SelfAttention  # noqa: F401
from trax_b200.dp import HeadShardedLSHSelfAttention  # noqa: F401

__all__ = ['HeadShardedLSHSelfAttention', 'LSHSelfAttention', 'PureLSHSelfAttention', 'PureLSHSelfAttentionWrapper', 'ReversibleHalfResidual', 'SelfAttention', 'ShapeDtype', 'set_async_host_io', 'set_reuse_forward_upload', 'set_weight_grad_allreduce', 'synchronize', 'host_io_bytes']
