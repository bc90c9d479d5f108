"""
Reseeding of over-stretched Gaussians on the device — the split of clone_velocity_field (3D/advance.py:59-90, 2D/advance.py:66-88;
SURVEY 8a row a8 / 8f row N4) on csrc/split.cu: one launch flags and counts the Gaussians whose axis ratio reaches the threshold,
one launch compacts the kept ones and appends the two children of every split one (mean + chol(Sigma) z, shortened longest axis),
with the stop_gradient bookkeeping.  The only host read is the count (the reference's `need_split.any()`).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream

NAMES = ('positions', 'scalings', 'rotations', 'values')
RULES = {3: dict(threshold=2., log_axis=float(np.log(2.)), log_all=float(np.log(2.) / 3.)),	# 3D/advance.py:66, :76-77
		 2: dict(threshold=1.5, log_axis=float(np.log(1.5)), log_all=0.)}			# 2D/advance.py:68, :75-77


def split_once(field, D, clamp_box=None, normals=None, seed=0):
	"""
	One round of splitting on `field` (a GaussianSplatting(3D)Fast): returns (n_split, stop_gradient).  With n_split == 0 nothing is
	touched and stop_gradient is None; otherwise the four parameter tensors are REPLACED by new leaf tensors of N + n_split rows
	(kept Gaussians first, in order, then the children) and stop_gradient (bool, N + n_split) is True for the kept ones.
	normals: optional (2, n_split, D) standard-normal draws (the tests replay the reference's); default: Philox from `seed`.
	"""
	lib = _lib.lib()
	rule = RULES[D]
	pos, scal, rot, val = [getattr(field, nm).detach().contiguous() for nm in NAMES]
	N = pos.shape[0]
	dev = pos.device
	flags = torch.empty(N, dtype=torch.int32, device=dev)
	count = torch.empty(1, dtype=torch.int32, device=dev)
	check(lib.gsr_split_flags(C.c_int(D), ptr(scal, name='scalings'), C.c_int64(N), C.c_float(rule['threshold']), ptr(flags, torch.int32), ptr(count, torch.int32), stream()),
		  'gsr_split_flags')
	n_split = int(count.item())
	if n_split == 0:
		return 0, None
	prefix = (torch.cumsum(flags, 0) - flags).to(torch.int32)
	M = N + n_split
	out = [torch.empty((M,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev) for t in (pos, scal, rot, val)]
	stop = torch.empty(M, dtype=torch.int32, device=dev)
	if normals is not None:
		normals = normals.detach().to(dev, torch.float32).contiguous()
		if tuple(normals.shape) != (2, n_split, D):
			raise _lib.GsrError(f'normals must have shape (2, {n_split}, {D})')
	box = (C.c_float * (2 * D))(*[float(v) for v in clamp_box]) if clamp_box is not None else None
	check(lib.gsr_split_apply(C.c_int(D), ptr(pos), ptr(scal), ptr(rot), ptr(val), C.c_int64(N), ptr(flags, torch.int32), ptr(prefix, torch.int32), C.c_int64(n_split),
							  ptr(normals, allow_none=True), C.c_uint64(int(seed)), box, C.c_float(rule['log_axis']), C.c_float(rule['log_all']),
							  ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), ptr(stop, torch.int32), stream()), 'gsr_split_apply')
	for nm, t in zip(NAMES, out):
		setattr(field, nm, t.requires_grad_())
	field.N = M
	return n_split, stop.bool()


def split_all(field, D, clamp_box=None, normals=None, seed=0, rounds=None, verbose=0):
	"""
	Split until no Gaussian reaches the threshold (3D: the reference loops, 3D/advance.py:62-90; 2D: one round, 2D/advance.py:66-88 —
	pass rounds=1).  Returns stop_gradient (bool, N after the splits; all True when nothing was split).  `normals`: a list with one
	(2, n_split, D) tensor per round, or None.
	"""
	dev = field.positions.device
	stop = torch.ones((field.N,), dtype=torch.bool, device=dev)
	k = 0
	while rounds is None or k < rounds:
		n, st = split_once(field, D, clamp_box, normals[k] if normals is not None else None, seed + k)
		if verbose:
			print(f'Add {n} particles.')
		if n == 0:
			break
		# a child of an earlier round that is kept now stays trainable: carry the old flags of the kept rows over
		kept_old = stop[: st.shape[0] - 2 * n] if k == 0 else None
		if k > 0:
			# rows of this round's input that were kept, in order: recover them from the new layout's front block
			raise_if = st[: st.shape[0] - 2 * n].all()
			assert bool(raise_if)
			prev = split_all._prev_keep_mask
			kept_old = stop[prev]
		st = st.clone()
		st[: kept_old.shape[0]] &= kept_old
		stop = st
		k += 1
	return stop
