"""ORACLE PIN — TEST INFRASTRUCTURE ONLY: runs the reference's OWN code on CPU so that the oracle can be checked against it.

The reference (`/root/reference`, google/trax) is pure Python.  Its array maths goes through `trax.fastmath`, which has a
NumPy backend of its own (`trax/fastmath/numpy.py`), and `LSHSelfAttention(use_reference_code=True).forward` is a Python
loop over (example, head) units around `forward_unbatched` (EA:2127-2170) — no jit / vmap / vjp.  What stops a plain
`import trax` here is only that third-party packages (jax, tensorflow, gin, ...) are not installed and that
`trax/__init__.py` pulls in the data / training stack.  `load()` therefore

  1. registers import stubs for exactly those third-party packages that are absent (attribute access and decorator use
     succeed; nothing in them is ever computed with);
  2. registers an empty `trax` package whose `__path__` points INTO `/root/reference/trax`, so that submodules
     (`trax.fastmath`, `trax.layers.research.efficient_attention`, ...) are the reference's files, executed where they
     lie — nothing is copied — while the heavyweight `trax/__init__.py` is skipped;
  3. switches `trax.fastmath` to the reference's NumPy backend and adds the primitives that backend lacks and the
     path calls, each a one-line NumPy statement of the `jax` primitive the JAX backend binds
     (`trax/fastmath/jax.py:192-199, 213, 214`, `ops.py:300-315`):
       sort_key_val(keys, values, dimension)  -> stable argsort of keys applied to both
       stop_gradient(x)                       -> x
       lt(a, b)                               -> a < b
       custom_grad(f_vjp, f)                  -> f          (forward semantics of a custom-VJP function)
       index_update(x, idx, y) / index_add    -> copy of x with x[idx] = y / x[idx] += y   (`x.at[idx].set/add`)
       jax.numpy.index_exp                    -> numpy.index_exp
     and replaces that backend's `random_split` (it returns `None` keys, `fastmath/numpy.py:72`, which the layers'
     shape inference cannot take) by one returning zero-valued uint32 keys — the NumPy backend never reads a key.
     With these the batched drivers run too when asked for their Python loop (`use_python_loop=True,
     n_parallel_heads=1`; EA:2261-2561, 3052-3265), and so do `PureLSHSelfAttentionWrapper` and `ReversibleHalfResidual`.

Everything else that runs — hashing (EA:60-119), `look_adjacent`, `mask_self_attention`, `attend` (EA:164-281),
`LSHSelfAttention.forward_unbatched` (EA:1918-1997), `PureLSHSelfAttention.forward_unbatched` (EA:2739-2826), the
`use_reference_code` driver loop and the batched drivers' Python loop, `core.Dense`, `LayerNorm`, rotary embedding, head
split / merge, the Serial / Branch / reversible combinators — is the reference's code.
Random numbers under that backend come from NumPy's global generator (`fastmath/numpy.py:37-40` ignores the key), so the
caller seeds `numpy.random` to make the hash rotations reproducible.

Because the stubs are process-global, call `load()` only in a dedicated process (`tests/golden/make_reference_golden.py`
is run as a script; the tests start it with `subprocess`).  `/root/reference` exists only in the build container, never
on the GPU box: what travels is the fixture the script writes.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = '/root/reference'
_THIRD_PARTY = ('jax', 'jaxlib', 'tensorflow', 'tensorflow_datasets', 'tensorflow_text', 'gin', 'absl', 'tensor2tensor',
                'gym', 'funcsigs', 't5', 'seqio', 'matplotlib')


class _Anything:
  """Attribute of a stub package: any attribute chain exists, a call whose first argument is a function or class hands it
  back (bare decorator, `gin.external_configurable(cls, module=...)`), any other call returns another stub (so it also
  works as a decorator factory or a base class)."""

  def __init__(self, name):
    self._name = name

  def __getattr__(self, k):
    if k.startswith('__'):
      raise AttributeError(k)
    return _Anything(self._name + '.' + k)

  def __call__(self, *a, **kw):
    if a and callable(a[0]) and not isinstance(a[0], _Anything):     # decorator / gin.external_configurable(cls, ...)
      return a[0]
    return _Anything(self._name + '()')

  def __mro_entries__(self, bases):
    return (object,)

  def __iter__(self):
    return iter(())


class _StubPackage(types.ModuleType):
  def __init__(self, name):
    super().__init__(name)
    self.__path__ = []
    self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

  def __getattr__(self, k):
    if k.startswith('__'):
      raise AttributeError(k)
    v = _Anything(self.__name__ + '.' + k)
    setattr(self, k, v)
    return v


class _StubFinder:
  def __init__(self, roots):
    self.roots = roots

  def find_spec(self, name, path=None, target=None):
    if name.split('.')[0] in self.roots:
      return importlib.machinery.ModuleSpec(name, self, is_package=True)
    return None

  def create_module(self, spec):
    return _StubPackage(spec.name)

  def exec_module(self, module):
    pass


def available(root=REFERENCE_ROOT):
  return os.path.isfile(os.path.join(root, 'trax', 'layers', 'research', 'efficient_attention.py'))


def load(root=REFERENCE_ROOT):
  """Returns a namespace with the reference's modules: `.EA` (efficient_attention), `.fastmath`, `.shapes`, `.layers`,
  `.reversible`, `.normalization`, and `.stubbed` (the absent third-party packages that were stubbed)."""
  import numpy as np
  if not available(root):
    raise FileNotFoundError('the reference checkout is not present at %s' % root)
  warnings.simplefilter('ignore')
  sys.dont_write_bytecode = True                     # never write __pycache__ into the read-only reference tree
  absent = tuple(m for m in _THIRD_PARTY if m not in sys.modules and importlib.util.find_spec(m) is None)
  sys.meta_path.insert(0, _StubFinder(absent))
  # isinstance() targets that the reference (trax/shapes.py:42-48) and scipy's array-API probe (it looks at
  # sys.modules['jax'].Array) need to be real types
  if 'tensorflow' in absent:
    import tensorflow as tf
    tf.TensorShape = type('TensorShape', (), {})
    tf.DType = type('DType', (), {})
  if 'jax' in absent:
    import jax
    jax.Array = type('Array', (), {})
  pkg = types.ModuleType('trax')
  pkg.__path__ = [os.path.join(root, 'trax')]
  sys.modules['trax'] = pkg
  from trax import fastmath, shapes
  from trax.fastmath.numpy import NUMPY_BACKEND
  from trax import layers                            # imported under the default backend name: module-level pmap etc.
  from trax.layers import normalization, reversible
  from trax.layers.research import efficient_attention as EA

  def sort_key_val(keys, values, dimension=-1):
    order = np.argsort(keys, axis=dimension, kind='stable')
    return np.take_along_axis(keys, order, axis=dimension), np.take_along_axis(values, order, axis=dimension)

  def index_update(x, idx, y):
    x = np.array(x, copy=True)
    x[idx] = y
    return x

  def index_add(x, idx, y):
    x = np.array(x, copy=True)
    x[idx] += y
    return x
  NUMPY_BACKEND.update(sort_key_val=sort_key_val, stop_gradient=lambda x: x, lt=np.less,
                       custom_grad=lambda f_vjp, f: f, index_update=index_update, index_add=index_add,
                       random_split=lambda prng, num=2: np.zeros((num, 2), np.uint32))
  if 'jax' in absent:
    jax.numpy.index_exp = np.index_exp
  fastmath.ops.set_backend('numpy')
  assert fastmath.backend_name() == 'numpy'
  return types.SimpleNamespace(EA=EA, fastmath=fastmath, shapes=shapes, layers=layers, reversible=reversible, normalization=normalization,
                               stubbed=absent, root=root)
