"""
GPU tests of the pieces added around the per-iteration step: the Philox sample kernels
(3D/advance.py:339-340, 3D/init_cond.py:227-249) and gsr_step_rebuild (step() -> zero_grad() -> reinitialize_grid()).
"""
import numpy as np
import pytest
import torch

from helpers import NAMES, rel_err

pytestmark = pytest.mark.gpu


def engine_field(n=8):
	from gaussian_fluids_code_b200 import gsr3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	P, S, R, V, mgs, gen = synthetic_field(n)
	return make_fast3d(P, S, R, V, 5e-3, mgs), gen


def test_sample_box_uniform_and_counter_based():
	o, _ = engine_field()
	e = o._engine
	box = (-.5, 1.5, 0., 2., 1., 4.)
	it = torch.zeros(1, device='cuda')
	a = e.sample_box(box, torch.empty((200000, 3), device='cuda'), 42, 0, it).clone()
	b = e.sample_box(box, torch.empty((200000, 3), device='cuda'), 42, 0, it).clone()
	assert torch.equal(a, b)					# counter-based: same (seed, stream, iteration) -> same samples
	it.fill_(1.)
	c = e.sample_box(box, torch.empty((200000, 3), device='cuda'), 42, 0, it).clone()
	d = e.sample_box(box, torch.empty((200000, 3), device='cuda'), 42, 1, it).clone()
	assert not torch.equal(a, c) and not torch.equal(c, d)	# the iteration (read from device memory) and the stream id both advance it
	lo, hi = torch.tensor(box[0::2], device='cuda'), torch.tensor(box[1::2], device='cuda')
	assert (a >= lo).all() and (a < hi).all()
	u = ((a - lo) / (hi - lo)).cpu().numpy()
	assert np.abs(u.mean(0) - .5).max() < 5e-3 and np.abs(u.var(0) - 1. / 12.).max() < 2e-3
	assert np.abs(np.corrcoef(u.T) - np.eye(3)).max() < 1e-2
	# replaying a captured graph draws fresh samples: the iteration counter lives on the device
	out = torch.empty((1000, 3), device='cuda')
	g = torch.cuda.CUDAGraph()
	s = torch.cuda.Stream()
	s.wait_stream(torch.cuda.current_stream())
	with torch.cuda.stream(s):
		e.sample_box(box, out, 7, 0, it)
	torch.cuda.current_stream().wait_stream(s)
	with torch.cuda.graph(g):
		e.sample_box(box, out, 7, 0, it)
	it.fill_(5.); g.replay(); r5 = out.clone()
	it.fill_(6.); g.replay(); r6 = out.clone()
	assert not torch.equal(r5, r6)


def test_sample_box_surface_matches_reference_semantics():
	"""3D/init_cond.py:227-249: area-weighted faces in the order x_min, x_max, y_min, y_max, z_min, z_max; inward normals"""
	o, _ = engine_field()
	e = o._engine
	box = (0., 1., 0., 2., 0., 3.)	# face areas: yz 6, zx 3, xy 2 (each twice)
	n = 400000
	data, normal = torch.empty((n, 3), device='cuda'), torch.empty((n, 3), device='cuda')
	e.sample_box_surface(box, data, normal, 1, 3, None)
	d, nm = data.cpu().numpy(), normal.cpu().numpy()
	assert np.allclose(np.abs(nm).sum(1), 1.) and set(np.unique(nm)) <= {-1., 0., 1.}
	axis = np.abs(nm).argmax(1)
	sign = nm[np.arange(n), axis]
	lo, hi = np.array(box[0::2]), np.array(box[1::2])
	on_face = d[np.arange(n), axis]
	assert np.all(np.where(sign > 0, on_face == lo[axis], on_face == hi[axis]))	# inward normal: +1 on the min face, -1 on the max face
	assert (d >= lo - 1e-6).all() and (d <= hi + 1e-6).all()
	freq = np.array([(axis == k).mean() for k in range(3)])
	assert np.abs(freq - np.array([6., 3., 2.]) / 11.).max() < 5e-3
	for k in range(3):		# the two faces of an axis are equally likely, and points are uniform on a face
		m = axis == k
		assert abs((sign[m] > 0).mean() - .5) < 1e-2
		others = [j for j in range(3) if j != k]
		u = (d[m][:, others] - lo[others]) / (hi[others] - lo[others])
		assert np.abs(u.mean(0) - .5).max() < 1e-2


@pytest.mark.parametrize('n,outside', [(8, 0), (6, 0), (10, 0), (10, 7), (12, 0)])
def test_step_rebuild_equals_step_then_build(n, outside):
	"""gsr_step_rebuild == gsr_step followed by gsr_build_grid + gsr_pack_gaussians on the updated parameters.  Up to 1024 Gaussians
	the rebuild happens inside the step's own cluster launch (step_cluster4_kernel<D, true>: counting sort over distributed shared
	memory): cell table, ids and packed records must be those of the separate hash build bit for bit; `outside` Gaussians sit
	beyond the extended domain (the tail bucket, whose internal order no kernel reads)"""
	from gaussian_fluids_code_b200.engine import FusedStepper
	res = []
	for rebuild in (False, True):
		o, gen = engine_field(n)
		if outside:
			with torch.no_grad():
				o.positions[torch.arange(outside) * 37 % o.N] += 5.
			o.reinitialize_grid()
		e = o._engine
		x = torch.rand((o.N, 3), generator=torch.Generator().manual_seed(3)).cuda()
		e.ensure_packed(o._params())
		bins = e.bin_samples(x, True)
		val, grad = torch.empty((o.N, 3), device='cuda'), torch.empty((o.N, 3, 3), device='cuda')
		e.forward(x, val, grad, False, perm=bins)
		ref_vor = torch.randn((o.N, 3), generator=torch.Generator().manual_seed(4)).cuda() * .1
		st = FusedStepper(e, [3e-4, 1e-5, 3e-4, 1e-5], 50, 10., 10., tau=o.clamp_threshold, min_grid_scale=o.min_grid_scale, ext_bounds=o._ext())
		st.init(o.scalings)
		acc, mask = e.backward_gather(x, bins.perm, bins.scs, val, grad, (0., 0., 0., 1., 0., 1.), {'ref_vor': ref_vor}, None, want_losses=True)
		lp, nblk = e.last_loss_partials
		params = [p.detach() for p in o._params()]
		st.step(params, acc, mask, loss_srcs=[(lp, nblk, [1. / o.N, 0., 1. / o.N, 0., 0., 0., 0., 0.])], rebuild=rebuild)
		if not rebuild:
			e.build(o.positions.detach(), params=params)
		torch.cuda.synchronize()
		res.append([p.cpu().numpy().copy() for p in params] + [e.cell_start.cpu().numpy().copy(), e.sorted_id.cpu().numpy().copy(), e.packed.cpu().numpy().copy(),
																   e.cull.cpu().numpy().copy(), np.array(st.scalars()[:14])])
	n_in = int(res[0][4][-1])	# cell_start[ncell]: Gaussians inside the hash
	assert n_in == o.N - outside
	for k, (a, b) in enumerate(zip(*res)):
		if k in (5, 6, 7) and outside:	# sorted_id, packed, cull: the tail bucket holds the same Gaussians in any order
			a, b = a.reshape(o.N, -1), b.reshape(o.N, -1)
			np.testing.assert_array_equal(a[:n_in], b[:n_in])
			np.testing.assert_array_equal(np.sort(a[n_in:], axis=0), np.sort(b[n_in:], axis=0))
			continue
		np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('n', [4, 8, 12])
def test_step_cluster_launch_equals_four_launch(n):
	"""gsr_step for small N (one launch of one 8-CTA cluster: partial sums and min(s) through distributed shared memory) against
	the four-launch form used for large N: same formulas, only the grouping of the float block sums differs.  Three consecutive
	steps, so that the Adam moments, the scheduler state and the re-armed min accumulator carry over."""
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	from gaussian_fluids_code_b200.engine import FusedStepper
	lib = _lib.lib()
	res = []
	for small_n in (2048, 0):
		assert lib.gsr_set_tuning(C.c_int(6), C.c_int(small_n)) == 0
		try:
			o, gen = engine_field(n)
			e = o._engine
			x = torch.rand((o.N, 3), generator=torch.Generator().manual_seed(3)).cuda()
			ref_vor = torch.randn((o.N, 3), generator=torch.Generator().manual_seed(4)).cuda() * .1
			st = FusedStepper(e, [3e-4, 1e-5, 3e-4, 1e-5], 50, 10., 10., tau=o.clamp_threshold, min_grid_scale=o.min_grid_scale, ext_bounds=o._ext())
			st.init(o.scalings)
			params = [p.detach() for p in o._params()]
			for _ in range(3):
				e.build(o.positions.detach(), params=params)
				bins = e.bin_samples(x, True)
				val, grad = torch.empty((o.N, 3), device='cuda'), torch.empty((o.N, 3, 3), device='cuda')
				e.forward(x, val, grad, False, perm=bins)
				acc, mask = e.backward_gather(x, bins.perm, bins.scs, val, grad, (0., 0., 0., 1., 0., 1.), {'ref_vor': ref_vor}, None, want_losses=True)
				lp, nblk = e.last_loss_partials
				st.step(params, acc, mask, loss_srcs=[(lp, nblk, [1. / o.N, 0., 1. / o.N, 0., 0., 0., 0., 0.])])
			torch.cuda.synchronize()
			res.append([p.cpu().numpy().copy() for p in params] + [np.array(st.scalars()[:14])])
		finally:
			lib.gsr_set_tuning(C.c_int(6), C.c_int(2048))
	for a, b in zip(res[0][:-1], res[1][:-1]):	# block partials are float sums over 32 or 128 Gaussians: ~1e-7 relative in the coefficients
		np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-8)
	np.testing.assert_allclose(res[0][-1], res[1][-1], rtol=1e-4, atol=1e-7)
	assert res[0][-1][0] == 3.	# three Adam steps counted


def test_xrank_sum_single_rank():
	"""gsr_xrank_sum with world = 1 (a single GPU cannot host kernels that wait for one another): the handshake with itself, the
	epoch arithmetic over two parities, and the copy-out.  The 2-rank behaviour is checked by tests/test_gpu_multirank.py on a two-GPU box
	(bit-identical replicas, equality with NCCL, lost peer) and the summation identity by tests/test_multirank_cpu.py."""
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	n = 4096
	bufs = [torch.randn(n, device='cuda') for _ in range(2)]
	pad = torch.zeros(64, dtype=torch.int32, device='cuda')
	out = torch.zeros(n, device='cuda')
	err = torch.zeros(1, dtype=torch.int32, device='cuda')
	base = torch.full((1,), 1000, dtype=torch.int32, device='cuda')
	it = torch.zeros(1, device='cuda')
	sig = (C.c_void_p * 1)(pad.data_ptr())
	for k in range(4):
		it.fill_(float(k))
		ptrs = (C.c_void_p * 1)(bufs[k % 2].data_ptr())
		rc = lib.gsr_xrank_sum(ptrs, sig, C.c_int(0), C.c_int(1), C.c_int64(n), _lib.ptr(it), _lib.ptr(base, torch.int32), _lib.ptr(out), _lib.ptr(err, torch.int32), _lib.stream())
		assert rc == 0
		torch.cuda.synchronize()
		assert torch.equal(out, bufs[k % 2]) and int(err.item()) == 0 and int(pad[0].item()) == 1000 + k + 1
	assert lib.gsr_xrank_sum(ptrs, sig, C.c_int(0), C.c_int(1), C.c_int64(n + 1), _lib.ptr(it), _lib.ptr(base, torch.int32), _lib.ptr(out), _lib.ptr(err, torch.int32), _lib.stream()) == -1


@pytest.mark.parametrize('n,Qb', [(10, 8192), (10, 1000), (30, 8192), (10, 20000)])
def test_surface_samples_drawn_and_binned_in_one_launch(n, Qb):
	"""gsr_sample_box_surface_binned == gsr_sample_box_surface followed by gsr_bin_samples: the same draws, the same order, the
	same cell table (one cluster launch when the batch fits it; n = 30 has too many cells and Qb = 20000 too many samples, which
	exercises the internal two-call form)"""
	o, gen = engine_field(n)
	e = o._engine
	it = torch.full((1,), 5., device='cuda')
	box = (0., 1.) * 3
	d1, n1 = torch.empty((Qb, 3), device='cuda'), torch.empty((Qb, 3), device='cuda')
	e.sample_box_surface(box, d1, n1, 42, 3, it)
	b1 = e.bin_samples(d1, True, tag='sep')
	d2, n2 = torch.empty((Qb, 3), device='cuda'), torch.empty((Qb, 3), device='cuda')
	b2 = e.sample_box_surface_binned(box, d2, n2, 42, 3, it, tag='fused')
	torch.cuda.synchronize()
	np.testing.assert_array_equal(d1.cpu().numpy(), d2.cpu().numpy())
	np.testing.assert_array_equal(n1.cpu().numpy(), n2.cpu().numpy())
	np.testing.assert_array_equal(b1.scs.cpu().numpy(), b2.scs.cpu().numpy())
	np.testing.assert_array_equal(b1.perm.cpu().numpy(), b2.perm.cpu().numpy())
