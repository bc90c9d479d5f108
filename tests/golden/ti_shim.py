"""
TEST INFRASTRUCTURE.  A minimal pure-Python stand-in for the parts of `taichi` / `taichi.math`
that the reference's kernels use, so that the reference's OWN kernel bodies (3D/GSR.py, 2D/GSR.py
— Python syntax, normally JIT-compiled by Taichi, which is not installed here) can be executed
as ordinary Python on tiny inputs.  Used only by tests/golden/make_golden.py, in the build
container where /root/reference exists, to produce the committed golden vectors.

Semantics implemented (the [Taichi-sem] assumptions of SURVEY.md §8):
  * one scalar dtype per run (`set_dtype(np.float32 | np.float64)`); Python float literals adopt it;
  * `//` on floats is floor(a / b);  int(x) truncates;
  * vec/mat `@` is the matrix product (vec @ mat = row-vector product, vec @ vec = dot);
    products are accumulated left to right, unfused;
  * kernel arguments annotated `ti.f32` are rounded to float32 at the call;
  * `ti.atomic_add(field[idx], v)` returns the old value.
"""
import sys
import types

import numpy as np

_DT = np.float32


def set_dtype(dt):
	global _DT
	_DT = dt


def _s(x):
	"""to scalar of the run dtype"""
	if isinstance(x, (Vec, Mat)):
		return x
	return _DT(x)


class Vec:
	__array_ufunc__ = None
	__array_priority__ = 1000

	def __init__(self, a):
		self.a = np.array([_DT(v) for v in a], dtype=_DT)

	def __len__(self): return len(self.a)
	def __getitem__(self, i): return self.a[i]
	def __setitem__(self, i, v): self.a[i] = _DT(v)
	def _bin(self, o, f):
		if isinstance(o, Vec):
			return Vec([f(x, y) for x, y in zip(self.a, o.a)])
		o = _DT(o)
		return Vec([f(x, o) for x in self.a])
	def __add__(self, o): return self._bin(o, lambda x, y: x + y)
	__radd__ = __add__
	def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
	def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
	def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
	__rmul__ = __mul__
	def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
	def __neg__(self): return Vec([-x for x in self.a])
	def __matmul__(self, o):
		if isinstance(o, Vec):
			return self.dot(o)
		if isinstance(o, Mat):	# row vector times matrix
			n, m = o.a.shape
			out = []
			for j in range(m):
				s = self.a[0] * o.a[0, j]
				for k in range(1, n):
					s = s + self.a[k] * o.a[k, j]
				out.append(s)
			return Vec(out)
		return NotImplemented
	def dot(self, o):
		s = self.a[0] * o.a[0]
		for k in range(1, len(self.a)):
			s = s + self.a[k] * o.a[k]
		return s
	def outer_product(self, o):
		return Mat([[x * y for y in o.a] for x in self.a])


class Mat:
	__array_ufunc__ = None
	__array_priority__ = 1000

	def __init__(self, rows):
		self.a = np.array([[_DT(v) for v in r] for r in rows], dtype=_DT)

	def __getitem__(self, ij):
		if isinstance(ij, tuple) and isinstance(ij[0], slice):
			return Vec(self.a[:, ij[1]])
		if isinstance(ij, tuple) and isinstance(ij[1], slice):
			return Vec(self.a[ij[0], :])
		return self.a[ij]
	def __setitem__(self, ij, v): self.a[ij] = _DT(v)
	def _bin(self, o, f):
		n, m = self.a.shape
		if isinstance(o, Mat):
			return Mat([[f(self.a[i, j], o.a[i, j]) for j in range(m)] for i in range(n)])
		o = _DT(o)
		return Mat([[f(self.a[i, j], o) for j in range(m)] for i in range(n)])
	def __add__(self, o): return self._bin(o, lambda x, y: x + y)
	__radd__ = __add__
	def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
	def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
	def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
	__rmul__ = __mul__
	def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
	def __neg__(self): return self * -1.
	def __matmul__(self, o):
		n, m = self.a.shape
		if isinstance(o, Mat):
			p = o.a.shape[1]
			rows = []
			for i in range(n):
				row = []
				for j in range(p):
					s = self.a[i, 0] * o.a[0, j]
					for k in range(1, m):
						s = s + self.a[i, k] * o.a[k, j]
					row.append(s)
				rows.append(row)
			return Mat(rows)
		if isinstance(o, Vec):
			out = []
			for i in range(n):
				s = self.a[i, 0] * o.a[0]
				for k in range(1, m):
					s = s + self.a[i, k] * o.a[k]
				out.append(s)
			return Vec(out)
		return NotImplemented
	def transpose(self):
		return Mat(self.a.T.tolist())
	def trace(self):
		s = self.a[0, 0]
		for k in range(1, self.a.shape[0]):
			s = s + self.a[k, k]
		return s


def _vecn(n):
	def ctor(*args):
		if len(args) == 1 and isinstance(args[0], (list, tuple, np.ndarray, Vec)):
			vals = list(args[0].a if isinstance(args[0], Vec) else args[0])
		elif len(args) == 1:
			vals = [args[0]] * n
		else:
			vals = list(args)
		assert len(vals) == n
		return Vec(vals)
	return ctor


def _matn(n):
	def ctor(arg):
		if isinstance(arg, (list, tuple)):
			return Mat(arg)
		return Mat([[arg] * n for _ in range(n)])
	return ctor


def _map(f):
	def g(x):
		if isinstance(x, Vec):
			return Vec([f(v) for v in x.a])
		return f(_DT(x))
	return g


def _sign(v):
	return _DT(int(v > 0) - int(v < 0))


class Ref(int):
	"""value read from a field that remembers where it came from (for ti.atomic_add)"""
	def __new__(cls, value, arr, idx):
		o = int.__new__(cls, int(value))
		o.arr, o.idx = arr, idx
		return o


class Field:
	def __init__(self, dtype=None, shape=None):
		self.arr = np.zeros(shape, dtype=np.int32)
	def __getitem__(self, idx): return Ref(self.arr[idx], self.arr, idx)
	def __setitem__(self, idx, v): self.arr[idx] = int(v)


def _atomic_add(ref, v):
	old = int(ref)
	ref.arr[ref.idx] += v
	return old


class _F32Tag:
	"""ti.f32: as an annotation it marks kernel args to round; as a call it casts."""
	def __call__(self, x): return np.float32(x)


class _I32Tag:
	def __call__(self, x): return int(x)


F32, I32 = _F32Tag(), _I32Tag()


def _kernel(fn):
	ann = getattr(fn, '__annotations__', {})
	names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
	round_idx = [k for k, nm in enumerate(names) if ann.get(nm) is F32]
	if not round_idx:
		return fn
	def wrapped(*args):
		args = list(args)
		for k in round_idx:
			args[k] = float(np.float32(args[k]))
		return fn(*args)
	wrapped.__wrapped__ = fn
	return wrapped


_RANDOM = []


def set_random_sequence(seq):
	"""the values successive ti.random() calls return (a replayable stand-in for Taichi's generator)"""
	global _RANDOM
	_RANDOM = [float(v) for v in seq][::-1]


def _random():
	if not _RANDOM:
		raise RuntimeError('ti.random(): sequence exhausted (set_random_sequence)')
	return _DT(_RANDOM.pop())


def _cross(a, b):
	return Vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def install():
	"""Put stub `taichi`, `taichi.math`, `vtk`, `matplotlib` modules into sys.modules."""
	ti = types.ModuleType('taichi')
	tm = types.ModuleType('taichi.math')
	ti.kernel = _kernel
	ti.func = lambda f: f
	ti.data_oriented = lambda c: c
	ti.init = lambda **kw: None
	ti.cpu, ti.cuda = 'cpu', 'cuda'
	ti.f32, ti.i32 = F32, I32
	ti.types = types.SimpleNamespace(ndarray=lambda: None)
	ti.field = Field
	ti.atomic_add = _atomic_add
	ti.random = _random
	ti.min = min
	ti.math = tm
	tm.vec2, tm.vec3, tm.vec4 = _vecn(2), _vecn(3), _vecn(4)
	tm.mat2, tm.mat3 = _matn(2), _matn(3)
	tm.eye = lambda n: Mat([[1. if i == j else 0. for j in range(n)] for i in range(n)])
	tm.exp, tm.sin, tm.cos, tm.sqrt = _map(np.exp), _map(np.sin), _map(np.cos), _map(np.sqrt)
	tm.sign = _map(_sign)
	tm.length = lambda v: np.sqrt(v.dot(v))
	tm.cross = _cross
	tm.dot = lambda a, b: a.dot(b)
	tm.floor = lambda x, dtype=None: int(np.floor(x)) if dtype is I32 else _DT(np.floor(x))
	tm.normalize = lambda v: v / np.sqrt(v.dot(v))
	sys.modules['taichi'] = ti
	sys.modules['taichi.math'] = tm
	for name in ('vtk', 'vtk.util', 'vtk.util.numpy_support', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
		m = types.ModuleType(name)
		sys.modules[name] = m
	sys.modules['matplotlib.patches'].Ellipse = object
	sys.modules['vtk'].util = sys.modules['vtk.util']
	sys.modules['vtk.util'].numpy_support = sys.modules['vtk.util.numpy_support']
	sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
	sys.modules['matplotlib'].patches = sys.modules['matplotlib.patches']
	return ti, tm


class GArr(np.ndarray):
	"""ndarray with a `.grad` attribute, standing in for a torch parameter inside the kernels."""
	def __new__(cls, a, dtype=None):
		o = np.array(a, dtype=dtype).view(cls)
		o.grad = None
		return o
	def __array_finalize__(self, obj):
		self.grad = getattr(obj, 'grad', None)
	def __getitem__(self, idx):
		r = super().__getitem__(idx)
		return r if not isinstance(r, np.ndarray) else r


class IdxArr(np.ndarray):
	"""ndarray whose iteration yields index tuples, like a Taichi struct-for over an ndarray argument (`for i, j, k in field:`)"""
	def __new__(cls, a, dtype=None):
		return np.array(a, dtype=dtype).view(cls)

	def __iter__(self):
		return iter(np.ndindex(*self.shape))


def record_steps(gv, rec, set_method=None):
	"""TEST INFRASTRUCTURE: wrap gv.step (and the loss method that fills the vor_* / div_* gradient sets) so that, at every
	optimiser step of the reference's own loop, the total .grad Adam consumes, the scheduler metric, the learning rates in use and
	the raw gradient sets (before the loop's PCGrad projection) are copied into `rec`"""
	names = ('positions', 'scalings', 'rotations', 'values')
	orig_step = gv.step

	def step(metrics):
		rec.setdefault('grads', []).append({nm: getattr(gv, nm).grad.detach().numpy().copy() for nm in names})
		rec.setdefault('metric', []).append(float(metrics))
		rec.setdefault('lr', []).append([o.param_groups[0]['lr'] for o in gv.optimizers])
		return orig_step(metrics)
	gv.step = step
	if set_method:
		orig = getattr(gv, set_method)

		def losses(x, *a, **kw):
			res = orig(x, *a, **kw)
			if kw.get('vor_positions_grad') is not None:
				rec.setdefault('sets', []).append({f'{t}_{nm}': kw[f'{t}_{nm}_grad'].detach().numpy().copy() for t in ('vor', 'div') for nm in names})
			return res
		setattr(gv, set_method, losses)


def store_steps(out, rec):
	for k, grads in enumerate(rec.get('grads', [])):
		for nm, v in grads.items():
			out[f'it{k + 1}_total_{nm}_grad'] = v
		out[f'it{k + 1}_metric'] = np.float64(rec['metric'][k])
		out[f'it{k + 1}_lr_used'] = np.array(rec['lr'][k])
	for k, sets in enumerate(rec.get('sets', [])):
		for key, v in sets.items():
			out[f'it{k + 1}_{key}_grad'] = v
