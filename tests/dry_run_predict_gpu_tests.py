"""Helper (not a test, not collected): dry run of tests/test_zgpu_predict.py's LAYER-level tests on the CPU — torch.cuda
patched to no-ops, the CUDA entry points (`predict._step`, `predict._pure_step`, the layers' training-path forward) replaced by
the oracle — so that the test logic and `trax_b200/predict.py`'s plumbing can be exercised where there is no GPU.  It proves
nothing about the kernels (the GPU tests do); it is how the layer-level tests were debugged while no GPU was at hand — they
then passed on a B200 at the first attempt (profiles/r2_gputest_predict_layers.log).        python tests/dry_run_predict_gpu_tests.py"""
import contextlib, sys
import numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.cuda.is_available = lambda: True
torch.cuda.current_device = lambda: 0
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.device = lambda dev: contextlib.nullcontext()
torch.cuda.synchronize = lambda *a: None
_orig_to = torch.Tensor.to
def _to(self, *a, **k):
  a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == 'cuda'))
  if not a and not k: return self
  return _orig_to(self, *a, **k)
torch.Tensor.to = _to
_orig_zeros, _orig_full = torch.zeros, torch.zeros_like
import trax_b200
from trax_b200 import predict, lsh_attention, self_attention
from oracle import lsh_oracle as O, predict_oracle as P, self_attention_oracle as SA

def fake_step(layer, mem, weights, q_start, buckets, rotations, causal):
  B, M, D = mem.shape; H = layer._n_heads
  w = tuple(a.double().numpy() for a in weights)
  out = np.zeros((B, 1, D))
  if rotations is None:
    cfg = SA.SelfAttentionConfig(n_heads=H, d_qk=64, d_v=64, share_qk=layer._share_qk, causal=True, chunk_len=layer._chunk_len, n_chunks_before=layer._n_chunks_before)
    for u in range(B * H):
      out[u // H] += P.self_attention_incremental_unit(cfg, mem[u // H].double().numpy(), q_start, 1, tuple(a[u % H] for a in w))
  else:
    cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=layer._chunk_len, n_chunks_before=layer._n_chunks_before, n_hashes=layer._n_hashes, n_buckets=layer._n_buckets)
    pcfg = P.PredictConfig(layer._predict_mem_len, layer._predict_drop_len)
    for u in range(B * H):
      o, nb, _ = P.incremental_forward_unit(cfg, pcfg, mem[u // H].double().numpy(), q_start, 1, w[0][u % H], w[1][u % H], w[2][u % H],
                                            buckets[u].numpy(), q_start, lambda n, _u=u: rotations[_u].numpy())
      out[u // H] += o; buckets[u] = torch.from_numpy(nb)
  return torch.from_numpy(out).to(mem.dtype)
predict._step = fake_step

def fake_pure_step(layer, qk_mem, v_mem, q_start, buckets, rotations):
  BH = qk_mem.shape[0]
  cfg = O.LSHConfig(n_heads=layer._n_heads, d_qk=64, d_v=64, causal=True, chunk_len=layer._chunk_len, n_chunks_before=layer._n_chunks_before, n_hashes=layer._n_hashes, n_buckets=layer._n_buckets)
  pcfg = P.PredictConfig(layer._predict_mem_len, layer._predict_drop_len)
  from tests import util
  w_q, w_v, w_o = util.core_identity_weights()
  x = np.concatenate([qk_mem.double().numpy(), v_mem.double().numpy()], -1)
  out = np.zeros((BH, 1, 64))
  for u in range(BH):
    out[u], nb, _ = P.incremental_forward_unit(cfg, pcfg, x[u], q_start, 1, w_q, w_v, w_o, buckets[u].numpy(), q_start, lambda n, _u=u: rotations[_u].numpy())
    buckets[u] = torch.from_numpy(nb)
  return torch.from_numpy(out).to(qk_mem.dtype)
predict._pure_step = fake_pure_step
from trax_b200 import pure_lsh_attention
_orig_pure_faob = pure_lsh_attention.PureLSHSelfAttention.forward_and_or_backward
def pure_faob(self, inputs, state, rng, output_grad=None, compute_output=True, update_state=True, _raw=False):
  if self._incremental and not _raw:
    return predict.pure_forward_and_or_backward(self, inputs, state, rng, output_grad, compute_output, update_state)
  from tests import util
  cfg = O.LSHConfig(n_heads=self._n_heads, d_qk=64, d_v=64, causal=self._causal, chunk_len=self._chunk_len, n_chunks_before=self._n_chunks_before, n_hashes=self._n_hashes, n_buckets=self._n_buckets)
  w_q, w_v, w_o = util.core_identity_weights()
  x = np.concatenate([inputs[0].double().numpy(), inputs[1].double().numpy()], -1)
  rot = self._rotations_override.numpy()
  res = [O.forward_unit(cfg, x[u], w_q, w_v, w_o, rotations=rot[u]) for u in range(x.shape[0])]
  return torch.from_numpy(np.stack([r.out for r in res])).to(inputs[0].dtype), (torch.from_numpy(np.stack([r.buckets for r in res])), state[1]), None
pure_lsh_attention.PureLSHSelfAttention.forward_and_or_backward = pure_faob
torch.Tensor.is_cuda = property(lambda self: True)
_oip = pure_lsh_attention.PureLSHSelfAttention.init_weights_and_state
pure_lsh_attention.PureLSHSelfAttention.init_weights_and_state = lambda self, sig, device=None: _oip(self, sig, device='cpu')
_oiw = pure_lsh_attention.PureLSHSelfAttentionWrapper.init_weights_and_state
pure_lsh_attention.PureLSHSelfAttentionWrapper.init_weights_and_state = lambda self, sig, device=None: _oiw(self, sig, device='cpu')


def raw_lsh(self, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True, _stash=None, _residual=None, _io_dtype=None, _raw=False):
  if self._incremental and not _raw:
    return predict.forward_and_or_backward(self, inputs, weights, state, rng, output_grad, compute_output, update_state)
  x = inputs
  w = tuple(a.double().numpy() for a in weights)
  if isinstance(self, self_attention.SelfAttention):
    cfg = SA.SelfAttentionConfig(n_heads=self._n_heads, d_qk=64, d_v=64, share_qk=self._share_qk, causal=self._causal, chunk_len=self._chunk_len, n_chunks_before=self._n_chunks_before)
    return torch.from_numpy(SA.forward_and_or_backward(cfg, x.double().numpy(), w)[0]).to(x.dtype), state, None, None
  cfg = O.LSHConfig(n_heads=self._n_heads, d_qk=64, d_v=64, causal=self._causal, chunk_len=self._chunk_len, n_chunks_before=self._n_chunks_before, n_hashes=self._n_hashes, n_buckets=self._n_buckets)
  out, nb, _, _ = O.forward_and_or_backward(cfg, x.double().numpy(), w, rotations=self._rotations_override.numpy())
  return torch.from_numpy(out).to(x.dtype), (torch.from_numpy(nb), state[1]), None, None
lsh_attention.LSHSelfAttention._forward_and_or_backward = raw_lsh
self_attention.SelfAttention._forward_and_or_backward = raw_lsh

_oi = lsh_attention.LSHSelfAttention.init_weights_and_state
lsh_attention.LSHSelfAttention.init_weights_and_state = lambda self, sig, device=None: _oi(self, sig, device='cpu')
_os = self_attention.SelfAttention.init_weights_and_state
self_attention.SelfAttention.init_weights_and_state = lambda self, sig, device=None: _os(self, sig, device='cpu')
from tests import test_zgpu_predict as T
for sq in (False, True):
  T.test_self_attention_token_by_token_equals_the_full_forward(sq); print('token-by-token ok', sq)
for pl, dt in ((128, torch.float32), (100, torch.bfloat16), (0, torch.float32)):
  T.test_lsh_predict_layer_prefix_then_tokens(pl, dt); print('layer ok', pl, dt)
for nm in ('lsh', 'self'):
  T.test_predict_layers_against_the_reference_own_outputs(nm); print('fixture ok', nm)

T.test_pure_core_predict_layer_prefix_then_tokens(); print('pure layer ok')
T.test_pure_lsh_wrapper_predict_token_by_token(); print('wrapper ok')
