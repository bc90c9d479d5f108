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
	One round of splitting on `field` (a GaussianSplatting(3D)Fast): returns (n_split, flags).  With n_split == 0 nothing is touched;
	otherwise the four parameter tensors are REPLACED by new leaf tensors of N + n_split rows — the kept Gaussians first, in their
	old order, then the children as [first samples | second samples] — and flags (int32, old N) tells which rows were split.
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
	n_split = int(count.item())	# the one host read (the reference's `need_split.any()`)
	if n_split == 0:
		return 0, flags
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
	return n_split, flags


def split_all(field, D, clamp_box=None, normals=None, seed=0, rounds=None, verbose=0):
	"""
	Split until no Gaussian reaches the threshold (3D: the reference loops, 3D/advance.py:62-90; 2D: one round, 2D/advance.py:66-88 —
	pass rounds=1).  Returns (stop_gradient, total): stop_gradient (bool, the final N) is True for the Gaussians that were never
	touched and False for every child, `total` the number of splits.  `normals`: a list with one (2, n_split, D) tensor per round.
	"""
	dev = field.positions.device
	stop = torch.ones((field.N,), dtype=torch.bool, device=dev)
	k = total = 0
	while rounds is None or k < rounds:
		n, flags = split_once(field, D, clamp_box, normals[k] if normals is not None else None, seed + k)
		if verbose:
			print(f'Add {n} particles.')
		if n == 0:
			break
		stop = torch.cat([stop[flags == 0], torch.zeros((2 * n,), dtype=torch.bool, device=dev)])	# kept rows carry their flag; children train
		total += n
		k += 1
	return stop, total


class EarlyStop:
	"""the stopping rule shared by the reference's optimisation loops (3D/advance.py:289-314, :141-160; 2D likewise): at every test,
	each watched loss must improve by its relative threshold or its stale counter grows by check_iter; stop when all are stale"""

	def __init__(self, names, patience, check_iter, thresholds=None):
		self.names, self.patience, self.check_iter = tuple(names), patience, check_iter
		self.thr = dict.fromkeys(self.names, 1e-3)
		self.thr.update(thresholds or {})
		self.best = dict.fromkeys(self.names, np.inf)
		self.stale = dict.fromkeys(self.names, 0)

	def update(self, cur):
		for k in self.names:
			if cur[k] < self.best[k] * (1. - self.thr[k]):
				self.best[k], self.stale[k] = cur[k], 0
			else:
				self.stale[k] += self.check_iter
		return all(self.stale[k] >= self.patience for k in self.names)


def refit(res, losses, data_generator, test_data_generator, trainable, batch_size, max_epoch, patience, verbose, check_iter=100):
	"""
	Train the trainable (new) Gaussians of `res` until the test losses stall: `losses(data, backward)` -> (total, value loss,
	gradient loss, ...) accumulates the gradients when backward is set; every check_iter iterations the value and gradient losses
	on the test points feed the stopping rule.  Returns the number of iterations.
	"""
	import time
	rule = EarlyStop(('loss', 'loss_grad'), patience, check_iter)
	t0 = time.time()
	done = 0
	while done < max_epoch:
		res.step(losses(data_generator(batch_size, res, trainable), True)[0])
		done += 1
		if done % check_iter:
			continue
		with torch.no_grad():
			out = losses(test_data_generator(res), False)
		cur = {'loss': float(out[1]), 'loss_grad': float(out[2])}
		if verbose:
			print(f"[clone] loss: {cur['loss']}, loss_grad: {cur['loss_grad']}, time: {time.time() - t0}")
			t0 = time.time()
		if rule.update(cur):
			break
	if verbose:
		print('[clone] Total epoch:', done, '' if done < max_epoch else '(Reached maximum iteration number)')
	return done
