"""
Device-side plumbing shared by the 2D and 3D drop-in classes: owns the hash buffers and scratch (torch tensors),
turns tensors into raw pointers and calls the C ABI (include/gsr_b200.h).  No arithmetic happens here.
"""
import ctypes as C
import os

import torch

from . import _lib, host
from ._lib import GSR_NSETS, LossCfg, check, ptr, stream


class _Scratch:
	"""grow-only byte buffers, one per purpose, so steady-state calls do not allocate"""

	def __init__(self, device):
		self.device = device
		self.bufs = {}
		self.retired = []

	def get(self, name, nbytes):
		b = self.bufs.get(name)
		if b is None or b.numel() < nbytes:
			if b is not None:
				# A captured CUDA graph may still hold this buffer's address (the iteration graph is captured once and replayed for
				# the rest of the run, while a later, larger batch — the test lattice — outgrows the buffer).  Never hand the old
				# block back to the allocator: a replay would scribble over whoever received it.
				self.retired.append(b)
			b = torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, device=self.device)
			self.bufs[name] = b
		return b

	def typed(self, name, shape, dtype):
		n = 1
		for s in shape:
			n *= int(s)
		itemsize = torch.empty((), dtype=dtype).element_size()
		return self.get(name, max(n, 1) * itemsize)[:n * itemsize].view(dtype).view(*shape)


class Bins:
	"""
	The engine's ordering of a batch of query points: perm (cell-sorted sample indices), scs (sample_cell_start on the
	padded grid, or None) and tiles (the tile -> row table of the large-Q kernels, or None).  Unpacks as (perm, scs).
	"""
	__slots__ = ('perm', 'scs', 'tiles')

	def __init__(self, perm, scs=None, tiles=None):
		self.perm, self.scs, self.tiles = perm, scs, tiles

	def __iter__(self):
		yield self.perm
		yield self.scs


def _unbin(b):
	if b is None or isinstance(b, Bins):
		return b
	return Bins(b)


class HashEngine:
	"""
	State: cell_start (ncell+1) int32, sorted_id (N) int32, packed (N, 12|8) f32 — see include/gsr_b200.h.
	`params` is a callable returning the CURRENT (positions, scalings, rotations, values) tensors of the owner,
	because the reference's callers replace those tensor objects outright (3D/advance.py:54-58, :178-179).
	"""

	def __init__(self, D, device):
		if not torch.cuda.is_available():
			raise _lib.GsrError('the B200 engine needs a CUDA device; there is no CPU fallback')
		self.D = D
		self.device = torch.device(device)
		self.lib = _lib.lib()
		self.scratch = _Scratch(self.device)
		self.desc = None
		self.N = 0
		self.cell_start = self.sorted_id = self.packed = self.cull = None
		self._packed_key = None
		self._bin_cache = {}
		# grid_scale lives in a persistent device scalar per field: the kernels always read it from there (desc.grid_scale_dev), so a
		# CUDA graph captured while this field had another grid_scale — e.g. the pull-back of the previous field inside the
		# iteration graph of the other field — bins with the current value on every replay
		self.gs_dev = torch.zeros(1, dtype=torch.float32, device=self.device)

	# ---- hash -------------------------------------------------------------------------------------
	def set_grid(self, ext_bounds, dims, grid_scale, tau):
		self.ext_bounds, self.dims = list(ext_bounds), list(dims)
		self.desc = host.make_desc(self.D, ext_bounds, dims, grid_scale, tau)
		self.gs_dev.fill_(float(grid_scale))	# stream-ordered; rounds to f32 exactly as the descriptor's host copy does
		self.desc.grid_scale_dev = self.gs_dev.data_ptr()
		self.ncell = host.n_cells(self.D, dims)

	def desc_at(self, gs_dev):
		"""the hash descriptor with ANOTHER device-resident grid scale: sample batches may be binned on a grid one iteration ahead of
		the hash's (gsr_step_cfg.sample_gs_slots), as long as the gather is told (backward_gather(sample_gs=...))"""
		if gs_dev is None:
			return self.desc
		d = _lib.GridDesc.from_buffer_copy(self.desc)
		d.grid_scale_dev = gs_dev.data_ptr()
		return d

	def build(self, positions, want_ref_format=False, params=None):
		"""gsr_build_grid: radix sort of the Gaussian cell keys; with params = (positions, scalings, rotations, values) the packed
		records are produced by the same call (one launch for small N)"""
		N = positions.shape[0]
		dev = self.device
		if self.cell_start is None or self.cell_start.numel() != self.ncell + 1:
			self.cell_start = torch.empty(self.ncell + 1, dtype=torch.int32, device=dev)
		if self.sorted_id is None or self.sorted_id.numel() != N:
			self.sorted_id = torch.empty(N, dtype=torch.int32, device=dev)
			self.packed = torch.empty((N, 12 if self.D == 3 else 8), dtype=torch.float32, device=dev)
			self.cull = torch.empty(N, dtype=torch.float32, device=dev)
		self.N = N
		nbytes = self.lib.gsr_build_grid_ws_bytes(C.byref(self.desc), C.c_int64(N))
		ws = self.scratch.get('sort', nbytes)
		cnt = off = None
		if want_ref_format:
			cnt = torch.empty(self.ncell, dtype=torch.int32, device=dev)
			off = torch.empty(self.ncell, dtype=torch.int32, device=dev)
		check(self.lib.gsr_build_grid(C.byref(self.desc), ptr(positions, name='positions'), C.c_int64(N),
									  ptr(self.cell_start, torch.int32), ptr(self.sorted_id, torch.int32),
									  ptr(cnt, torch.int32, True), ptr(off, torch.int32, True),
									  ptr(params[1], name='scalings') if params else None, ptr(params[2], name='rotations', align16=True) if params else None,
									  ptr(params[3], name='values') if params else None, ptr(self.packed, align16=True) if params else None,
									  ptr(self.cull) if params else None,
									  ptr(ws, torch.uint8), C.c_size_t(ws.numel()), stream()), 'gsr_build_grid')
		self._packed_key = self._key(params) if params else None
		return cnt, off

	@staticmethod
	def _key(tensors):
		return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)

	def ensure_packed(self, params):
		"""(re)pack {mu, Sigma^-1, v} in cell order when any parameter tensor changed since the last pack"""
		key = self._key(params)
		if key == self._packed_key:
			return
		p, s, r, v = params
		if p.shape[0] != self.N:
			raise _lib.GsrError('the number of Gaussians changed since the last reinitialize_grid()')
		check(self.lib.gsr_pack_gaussians(C.byref(self.desc), ptr(p, name='positions'), ptr(s, name='scalings'),
										  ptr(r, name='rotations', align16=True), ptr(v, name='values'), C.c_int64(self.N),
										  ptr(self.cell_start, torch.int32), ptr(self.sorted_id, torch.int32),
										  ptr(self.packed, align16=True), ptr(self.cull), stream()), 'gsr_pack_gaussians')
		self._packed_key = key

	def min_scaling(self, scalings):
		out = torch.empty(1, dtype=torch.float32, device=self.device)
		check(self.lib.gsr_min_scaling(ptr(scalings, name='scalings'), C.c_int64(scalings.numel()), ptr(out), stream()), 'gsr_min_scaling')
		return out

	# ---- samples ----------------------------------------------------------------------------------
	# smallest Q for which the tiled shared-memory kernels are used (mirrors GSR_TUNE_TILED_MIN_Q)
	TILED_MIN_Q = int(os.environ.get('GSR_TILED_MIN_Q', 1 << 17))
	TILED_MIN_SPC = float(os.environ.get('GSR_TILED_MIN_SPC', 48))	# samples per hash cell

	@classmethod
	def set_tiled_min_q(cls, q, min_spc=None):
		cls.TILED_MIN_Q = int(q)
		cls.TILED_MIN_SPC = (0. if int(q) <= 1 else 48.) if min_spc is None else float(min_spc)	# forcing the tiled path (tests) lifts the density rule
		check(_lib.lib().gsr_set_tuning(C.c_int(1), C.c_int(int(q))), 'gsr_set_tuning')

	BIN_CACHE_AGE = 16	# uses of a cached ordering before it is refreshed while grid_scale lives on the device

	def bin_samples(self, x, need_cells, tag='x', gs_dev=None):
		"""
		Order a batch of query points for the kernels.  Large forward-only batches (a static test / output lattice evaluated
		again and again) keep their ordering: the tiled kernels only use it for locality — every point recomputes its own cell
		and stencil, so an ordering made for a slightly older grid_scale costs speed, never correctness (tests/test_gpu_tiled.py
		evaluates with orderings of a different grid).  Batches that feed the backward gather are always ordered afresh.
		"""
		Q = x.shape[0]
		# tiled kernels: large batches with enough samples per cell for compact warps (measured: below ~48 samples per cell the
		# one-point-per-thread kernels win: profiles/README.md)
		need_tiles = self.D == 3 and Q >= self.TILED_MIN_Q and Q >= self.TILED_MIN_SPC * self.ncell
		key = None
		if need_tiles and not need_cells:
			key = (x.data_ptr(), x._version, Q, tuple(self.dims))
			gs_now = None	# grid_scale is device-resident: orderings are refreshed by age
			ent = self._bin_cache.get(key)
			# reuse while young; when both grid_scales are known on the host, also require them to be close (the ordering
			# only matters for locality, see the docstring)
			if ent is not None and ent['uses'] < self.BIN_CACHE_AGE and (gs_now is None or ent['gs'] is None or abs(ent['gs'] / gs_now - 1.) < .05):
				ent['uses'] += 1
				return ent['bins']
		alloc = (lambda name, shape: torch.empty(shape, dtype=torch.int32, device=self.device)) if key else (lambda name, shape: self.scratch.typed(name + tag, shape, torch.int32))
		perm = alloc('perm_', (Q,))
		scs = None
		if True:	# the single-launch and counting hash paths produce the cell table anyway, and it is small
			pcell = self.lib.gsr_padded_cells(C.byref(self.desc))
			scs = alloc('scs_', (pcell + 1,))
		nbytes = self.lib.gsr_bin_samples_ws_bytes(C.byref(self.desc), C.c_int64(Q))
		ws = self.scratch.get('sort_' + tag, nbytes)	# per tag: batches of different tags may be in flight on different streams
		check(self.lib.gsr_bin_samples(C.byref(self.desc_at(gs_dev)), ptr(x, name='x'), C.c_int64(Q), ptr(perm, torch.int32), ptr(scs, torch.int32, True), C.c_int(1 if need_tiles else 0),
									   ptr(ws, torch.uint8), C.c_size_t(ws.numel()), stream()), 'gsr_bin_samples')
		tiles = None
		if need_tiles:
			tiles = alloc('tiles_', (self.lib.gsr_tile_slots(C.byref(self.desc), C.c_int64(Q)),))
			check(self.lib.gsr_build_tiles(C.byref(self.desc), ptr(scs, torch.int32), C.c_int64(Q), ptr(tiles, torch.int32), stream()), 'gsr_build_tiles')
		bins = Bins(perm, scs, tiles)
		if key:
			if len(self._bin_cache) >= 4:
				self._bin_cache.pop(next(iter(self._bin_cache)))
			self._bin_cache[key] = {'bins': bins, 'gs': gs_now, 'uses': 1}
		return bins

	def _x(self, x):
		if x.dim() != 2 or x.shape[1] != self.D:
			raise _lib.GsrError(f'sample points must have shape (Q, {self.D})')
		return x.detach()

	# ---- kernels ----------------------------------------------------------------------------------
	def forward(self, x, val, grad, accumulate, perm=None):
		x = self._x(x)
		b = _unbin(perm) or self.bin_samples(x, False)
		check(self.lib.gsr_forward(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True), ptr(self.cull),
								   ptr(x, name='x'), C.c_int64(x.shape[0]), ptr(b.perm, torch.int32), ptr(b.scs, torch.int32, True), ptr(b.tiles, torch.int32, True),
								   ptr(val, allow_none=True, name='val'), ptr(grad, allow_none=True, name='grad'), C.c_int(1 if accumulate else 0), stream()), 'gsr_forward')

	def rk4(self, start, dt, goal_pos, deformation=None, goal_val=None, goal_grad=None):
		start = self._x(start)
		b = self.bin_samples(start, False)
		check(self.lib.gsr_rk4(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True), ptr(self.cull),
							   ptr(start, name='start_pos'), C.c_int64(start.shape[0]), ptr(b.perm, torch.int32), ptr(b.scs, torch.int32, True), ptr(b.tiles, torch.int32, True), C.c_float(dt),
							   ptr(goal_pos), ptr(deformation, allow_none=True), ptr(goal_val, allow_none=True), ptr(goal_grad, allow_none=True), stream()), 'gsr_rk4')

	def advected_vorticity(self, x, dt, ref_vor, ref_hel=None, domain=None, perm=None):
		x = self._x(x)
		b = _unbin(perm) or self.bin_samples(x, False)
		dom = (C.c_float * 4)(*domain) if domain is not None else None
		check(self.lib.gsr_advected_vorticity(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True), ptr(self.cull),
											  ptr(x, name='x'), C.c_int64(x.shape[0]), ptr(b.perm, torch.int32), ptr(b.scs, torch.int32, True), ptr(b.tiles, torch.int32, True),
											  C.c_float(dt), dom,
											  ptr(ref_vor), ptr(ref_hel, allow_none=True), stream()), 'gsr_advected_vorticity')

	def count_pairs(self, x, counts, evals=1, with_accepted=False):
		"""work census: counts (device int64[2]) += evals * (C(x), P(x)); P only when with_accepted"""
		x = self._x(x)
		one = torch.zeros(2, dtype=torch.int64, device=self.device)
		check(self.lib.gsr_count_pairs(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True) if with_accepted else None,
									   ptr(x, name='x'), C.c_int64(x.shape[0]), ptr(one, torch.int64), stream()), 'gsr_count_pairs')
		counts.add_(one, alpha=int(evals))

	def mark_neighbors(self, x, mark):
		x = self._x(x)
		check(self.lib.gsr_mark_neighbors(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.sorted_id, torch.int32),
										  ptr(self.packed, align16=True), ptr(x, name='x'), C.c_int64(x.shape[0]), ptr(mark, torch.int32), stream()), 'gsr_mark_neighbors')

	def backward_gather(self, x, perm, scs, val, grad, weights, refs, stop_gradient, Q_norm=None, tag='acc', want_losses=False, acc=None, loss_partials=None,
						sample_gs=None):
		"""returns (acc, sets_mask); acc is (3, N, 12|7) in original Gaussian order; sample_gs: the device grid scale `scs` was binned
		with when that is not the hash's (bin_samples(gs_dev=...))"""
		x = self._x(x)
		Q = x.shape[0]
		cfg = LossCfg()
		cfg.w_val, cfg.w_boundary, cfg.w_grad, cfg.w_vor, cfg.w_hel, cfg.w_div = [float(w) for w in weights]
		cfg.Q_norm = int(Q_norm if Q_norm is not None else Q)
		keep = []
		for name in ('ref_val', 'normals', 'normal_ref', 'ref_grad', 'ref_vor', 'ref_hel'):
			t = refs.get(name)
			if t is not None:
				t = t.detach()
				keep.append(t)
				setattr(cfg, name, ptr(t, name=name).value)
		if stop_gradient is not None:
			if stop_gradient.dtype != torch.int32:
				stop_gradient = stop_gradient.to(torch.int32)
			keep.append(stop_gradient)
			cfg.stop_gradient = ptr(stop_gradient, torch.int32, name='stop_gradient').value
		if sample_gs is not None:
			cfg.sample_grid_scale_dev = sample_gs.data_ptr()
		self.last_loss_partials = None
		if want_losses or loss_partials is not None:
			nblk = self.lib.gsr_loss_blocks(C.c_int64(Q))
			lp = loss_partials if loss_partials is not None else self.scratch.typed('lp_' + tag, (nblk, 8), torch.float32)
			if lp.numel() != nblk * 8:
				raise _lib.GsrError('loss_partials must hold gsr_loss_blocks(Q) * 8 floats')
			cfg.loss_partials = ptr(lp).value
			self.last_loss_partials = (lp, nblk)
		AF = 12 if self.D == 3 else 7
		if acc is None:
			acc = self.scratch.typed(tag, (GSR_NSETS, self.N, AF), torch.float32)
		elif acc.numel() != GSR_NSETS * self.N * AF:
			raise _lib.GsrError('acc must hold 3 * N * GSR_ACC_FLOATS floats')
		nbytes = self.lib.gsr_backward_ws_bytes(C.byref(self.desc), C.c_int64(self.N), C.c_int64(Q))
		ws = self.scratch.get('adjoint_' + tag, nbytes)
		mask = C.c_int(0)
		check(self.lib.gsr_backward_gather(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.sorted_id, torch.int32),
										   ptr(self.packed, align16=True), C.c_int64(self.N), ptr(x, name='x'), C.c_int64(Q),
										   ptr(perm, torch.int32), ptr(scs, torch.int32), ptr(val, allow_none=True, name='val'), ptr(grad, allow_none=True, name='grad'),
										   C.byref(cfg), ptr(acc, align16=True), C.byref(mask), ptr(ws, torch.uint8, align16=True), C.c_size_t(ws.numel()), stream()),
			  'gsr_backward_gather')
		return acc, mask.value

	def backward_epilogue(self, scalings, rotations, acc, mask, outs):
		"""outs: 3 lists (direct, vor, div) of 4 gradient tensors (positions, scalings, rotations, values), accumulated into"""
		if mask == 0:
			return
		arr = ((C.c_void_p * 4) * GSR_NSETS)()
		for s in range(GSR_NSETS):
			for k in range(4):
				t = outs[s][k] if outs[s] is not None else None
				arr[s][k] = ptr(t, name='gradient buffer').value if t is not None else None
		check(self.lib.gsr_backward_epilogue(C.byref(self.desc), ptr(scalings.detach(), name='scalings'), ptr(rotations.detach(), name='rotations'),
											 C.c_int64(self.N), ptr(acc, align16=True), C.c_int(mask), arr, stream()), 'gsr_backward_epilogue')

	def sample_box(self, box, out, seed, stream_id, iteration=None):
		"""gsr_sample_box: out (n,3) uniform in box = (x_min, x_max, y_min, y_max, z_min, z_max); iteration: device float scalar or None"""
		b = (C.c_float * 6)(*[float(v) for v in box])
		check(self.lib.gsr_sample_box(b, C.c_int64(out.shape[0]), C.c_uint64(seed), C.c_uint32(stream_id), ptr(iteration, allow_none=True), ptr(out), stream()), 'gsr_sample_box')
		return out

	def sample_box_surface_binned(self, box, data, normal, seed, stream_id, iteration=None, tag='pb', gs_dev=None):
		"""gsr_sample_box_surface_binned: draw the boundary samples AND order them for the kernels (one launch for the per-iteration
		batch sizes); returns Bins — exactly what sample_box_surface + bin_samples(data, True, tag) give"""
		b = (C.c_float * 6)(*[float(v) for v in box])
		Q = data.shape[0]
		perm = self.scratch.typed('perm_' + tag, (Q,), torch.int32)
		scs = self.scratch.typed('scs_' + tag, (self.lib.gsr_padded_cells(C.byref(self.desc)) + 1,), torch.int32)
		ws = self.scratch.get('sort_' + tag, self.lib.gsr_bin_samples_ws_bytes(C.byref(self.desc), C.c_int64(Q)))
		check(self.lib.gsr_sample_box_surface_binned(b, C.c_int64(Q), C.c_uint64(seed), C.c_uint32(stream_id), ptr(iteration, allow_none=True), ptr(data), ptr(normal),
													 C.byref(self.desc_at(gs_dev)), ptr(perm, torch.int32), ptr(scs, torch.int32), ptr(ws, torch.uint8), C.c_size_t(ws.numel()), stream()),
			  'gsr_sample_box_surface_binned')
		return Bins(perm, scs, None)

	def sample_box_surface(self, box, data, normal, seed, stream_id, iteration=None):
		b = (C.c_float * 6)(*[float(v) for v in box])
		check(self.lib.gsr_sample_box_surface(b, C.c_int64(data.shape[0]), C.c_uint64(seed), C.c_uint32(stream_id), ptr(iteration, allow_none=True),
											  ptr(data), ptr(normal), stream()), 'gsr_sample_box_surface')
		return data, normal

	def advect_density(self, axes, domain, dt, density_a, out_a, density_b=None, out_b=None, x_range=None, executed=None):
		"""gsr_advect_density(_slab): semi-Lagrangian step of one or two density fields on the lattice spanned by `axes` = (xs, ys, zs);
		x_range = (begin, end): only those x planes of the outputs are computed (a process's slab); executed (1 int64, device): run the
		census instantiation, which also adds the number of pair tests that survive the culling (measurement only)"""
		xs, ys, zs = axes
		dom = (C.c_float * 6)(*[float(v) for v in domain])
		x0, x1 = x_range if x_range is not None else (0, xs.numel())
		if executed is not None:
			check(self.lib.gsr_advect_density_census(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True), ptr(self.cull),
													 ptr(xs), ptr(ys), ptr(zs), C.c_int(xs.numel()), C.c_int(ys.numel()), C.c_int(zs.numel()), C.c_int(int(x0)), C.c_int(int(x1)),
													 dom, C.c_float(dt), ptr(density_a, name='density'), ptr(density_b, allow_none=True), ptr(out_a), ptr(out_b, allow_none=True),
													 ptr(executed, torch.int64), stream()), 'gsr_advect_density_census')
			return
		check(self.lib.gsr_advect_density_slab(C.byref(self.desc), ptr(self.cell_start, torch.int32), ptr(self.packed, align16=True), ptr(self.cull),
											   ptr(xs), ptr(ys), ptr(zs), C.c_int(xs.numel()), C.c_int(ys.numel()), C.c_int(zs.numel()), C.c_int(int(x0)), C.c_int(int(x1)),
											   dom, C.c_float(dt), ptr(density_a, name='density'), ptr(density_b, allow_none=True), ptr(out_a), ptr(out_b, allow_none=True),
											   stream()), 'gsr_advect_density')

	def sample_losses(self, val, grad, refs, Q):
		"""gsr_sample_losses: the 8 loss slots of include/gsr_b200.h summed over the samples (device tensor, no sync)"""
		cfg = LossCfg()
		keep = []
		for name in ('ref_val', 'normals', 'normal_ref', 'ref_grad', 'ref_vor', 'ref_hel'):
			t = refs.get(name)
			if t is not None:
				keep.append(t)
				setattr(cfg, name, ptr(t.detach(), name=name).value)
		sums = torch.empty(8, dtype=torch.float32, device=self.device)
		nblk = self.lib.gsr_loss_blocks(C.c_int64(Q))
		ws = self.scratch.get('loss_ws', nblk * 8 * 4)
		check(self.lib.gsr_sample_losses(C.byref(self.desc), C.c_int64(Q), ptr(val, allow_none=True), ptr(grad, allow_none=True), C.byref(cfg),
										 ptr(sums), ptr(ws, torch.uint8), C.c_size_t(ws.numel()), stream()), 'gsr_sample_losses')
		return sums


class FusedStepper:
	"""
	Device-resident optimiser of one `project` / fit phase: Adam moments, lrs, ReduceLROnPlateau state, the loss metric
	and the next grid_scale all live in one float32 state tensor (layout: include/gsr_b200.h); gsr_step advances it
	with 4 launches and no host synchronisation.
	"""

	def __init__(self, engine, lrs, patience, w_aniso, w_vol, w_valreg=0., w_dpos=0., pcgrad=True, factor=.9,
				 tau=None, min_grid_scale=None, ext_bounds=None, sample_grid_ahead=False):
		import numpy as np
		self.e = engine
		D = engine.D
		cfg = _lib.StepCfg()
		cfg.D = D
		for k in range(4):
			cfg.lr[k] = float(lrs[k])
		cfg.beta1, cfg.beta2, cfg.eps = .9, .999, 1e-8
		cfg.sched_factor, cfg.sched_threshold, cfg.sched_eps, cfg.sched_min_lr = float(factor), 1e-4, 1e-8, 0.
		cfg.sched_patience = int(patience)
		cfg.w_aniso, cfg.w_vol, cfg.w_valreg, cfg.w_dpos = float(w_aniso), float(w_vol), float(w_valreg), float(w_dpos)
		cfg.aniso_ratio = 1.5
		cfg.pcgrad = 1 if pcgrad else 0
		cfg.grid_coef = float(np.sqrt(-2. * np.log(tau))) if tau else 0.
		cfg.min_grid_scale = float(min_grid_scale)
		cfg.grid_scale_tau0 = float(max(ext_bounds[2 * k + 1] - ext_bounds[2 * k] for k in range(D)))
		cfg.grid_scale_out = engine.gs_dev.data_ptr()	# every step also leaves the next grid_scale in the field's persistent scalar
		self.sample_gs = None
		if sample_grid_ahead:	# sample grid scales one iteration ahead of the hash (gsr_step_cfg.sample_gs_slots): slot [iteration & 1]
			self.sample_gs = torch.zeros(2, dtype=torch.float32, device=engine.device)
			cfg.sample_gs_slots = self.sample_gs.data_ptr()
			cfg.sample_gs_margin = float(np.exp(4. * float(lrs[1])) * (1. + 1e-6))	# Adam moves a log-radius by < 3.17 lr per step
		self.cfg = cfg
		self.N = None
		self.state = None

	def init(self, scalings, keep_clock=True):
		"""start an optimisation phase; keep_clock: the sample clock (state[ST_CLOCK], the iteration number the sample generators
		read) runs on from the previous phase on the same state instead of restarting at 0"""
		N = scalings.shape[0]
		self.N = N
		nfl = self.e.lib.gsr_step_state_floats(C.c_int(self.e.D), C.c_int64(N))
		fresh = self.state is None or self.state.numel() != nfl
		if fresh:	# a restart on the same N keeps the buffers (and captured graphs) valid
			self.state = torch.empty(nfl, dtype=torch.float32, device=self.e.device)
			self.ws = torch.empty(self.e.lib.gsr_step_ws_bytes(C.c_int(self.e.D), C.c_int64(N)), dtype=torch.uint8, device=self.e.device)
		self.cfg.keep_clock = 0 if (fresh or not keep_clock) else 1
		check(self.e.lib.gsr_step_init(C.byref(self.cfg), C.c_int64(N), ptr(scalings.detach()), ptr(self.state), stream()), 'gsr_step_init')
		# (gsr_step_init and every gsr_step write the grid_scale into the state AND into engine.gs_dev, which the kernels read)

	def step(self, params, acc, mask, extra=(), loss_srcs=(), positions_org=None, rebuild=False):
		"""one optimiser iteration (gsr_step); rebuild=True also rebuilds the engine's hash and packed records from the updated
		parameters (gsr_step_rebuild: one launch for small N)"""
		p, s, r, v = params
		ex = (C.c_void_p * 2)()
		for k in range(2):
			ex[k] = ptr(extra[k], align16=True).value if k < len(extra) and extra[k] is not None else None
		srcs = (_lib.LossSrc * 3)()
		for k, (partials, nblk, w) in enumerate(loss_srcs):
			srcs[k].partials = ptr(partials).value
			srcs[k].nblocks = int(nblk)
			for q in range(8):
				srcs[k].w[q] = float(w[q])
		if rebuild:
			e = self.e
			if e.N != self.N or e.packed is None:
				raise _lib.GsrError('rebuild needs an engine that already holds a hash of these Gaussians')
			hws = e.scratch.get('sort', e.lib.gsr_build_grid_ws_bytes(C.byref(e.desc), C.c_int64(self.N)))
			check(e.lib.gsr_step_rebuild(C.byref(self.cfg), C.c_int64(self.N), ptr(p), ptr(s), ptr(r, align16=True), ptr(v),
										 ptr(acc, allow_none=True, align16=True), C.c_int(mask), ex, srcs, C.c_int(len(loss_srcs)),
										 ptr(positions_org, allow_none=True), ptr(self.state), ptr(self.ws, torch.uint8), C.c_size_t(self.ws.numel()),
										 C.byref(e.desc), ptr(e.cell_start, torch.int32), ptr(e.sorted_id, torch.int32), ptr(e.packed, align16=True), ptr(e.cull),
										 ptr(hws, torch.uint8), C.c_size_t(hws.numel()), stream()), 'gsr_step_rebuild')
			e._packed_key = None
			return
		check(self.e.lib.gsr_step(C.byref(self.cfg), C.c_int64(self.N), ptr(p), ptr(s), ptr(r), ptr(v),
								  ptr(acc, allow_none=True, align16=True), C.c_int(mask), ex, srcs, C.c_int(len(loss_srcs)),
								  ptr(positions_org, allow_none=True), ptr(self.state), ptr(self.ws, torch.uint8), C.c_size_t(self.ws.numel()), stream()), 'gsr_step')

	@property
	def clock(self):
		"""device scalar (1,) holding the running iteration number for the sample generators"""
		return self.state[_lib.ST_CLOCK:_lib.ST_CLOCK + 1]

	def scalars(self):
		"""host copy of the scalar block (synchronises)"""
		return self.state[:_lib.STATE_SCALARS].tolist()

	def detach(self):
		"""the final grid_scale as a host number (synchronises once); the kernels keep reading the field's device scalar"""
		gs = float(self.state[_lib.ST_GRID_SCALE].item())
		self.e.desc.grid_scale = gs
		return gs
