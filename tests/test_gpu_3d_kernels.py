"""
GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the drop-in class and hence the
C ABI, against (i) the golden vectors produced from the reference's own kernel bodies and (ii) the CPU oracle on
seeded inputs.  Tolerances (BASELINE.json north_star): integer outputs bit-exact; fields and gradients 1e-5
relative, measured per tensor as max|a-b| / max|b| against the float64 oracle.
"""
import numpy as np
import pytest
import torch

from helpers import NAMES, load_golden, oracle_from_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def make_fast3d(pos, scal, rot, vals, tau, mgs):
	from gaussian_fluids_code_b200 import gsr3d
	o = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., np.asarray(pos, np.float32), min_grid_scale=mgs, clamp_threshold=tau, dim=3)
	dev = gsr3d.device
	with torch.no_grad():
		o.scalings.copy_(torch.tensor(np.asarray(scal, np.float32), device=dev))
		o.rotations.copy_(torch.tensor(np.asarray(rot, np.float32), device=dev))
		o.values.copy_(torch.tensor(np.asarray(vals, np.float32), device=dev))
	o.zero_grad()
	return o


def from_golden(g):
	return make_fast3d(g['in_positions'], g['in_scalings'], g['in_rotations'], g['in_values'], float(g['in_tau']), float(g['in_min_grid_scale']))


def T(a, dtype=torch.float32):
	return torch.tensor(np.asarray(a), dtype=dtype, device='cuda')


def test_grid_matches_reference_golden():
	g = load_golden('ref3d_kernels_f32.npz')
	o = from_golden(g)
	assert list(g['grid_size']) == o.grid_size
	assert np.float32(o.grid_scale) == np.float32(g['grid_scale'])
	cnt, off, sid = o.grid_arrays()
	np.testing.assert_array_equal(cnt.cpu().numpy().ravel(), g['grid_cnt'])
	np.testing.assert_array_equal(off.cpu().numpy().ravel(), g['grid_offset'])
	np.testing.assert_array_equal(sid.cpu().numpy(), g['sorted_id'])


@pytest.mark.parametrize('case', ['project', 'fit', 'boundary', 'all'])
def test_losses_match_reference_golden(case):
	"""forward + backward of get_losses against the reference kernels run in float64 (tests/golden)"""
	g = load_golden('ref3d_kernels_f64.npz')
	o = from_golden(g)
	wv, wb, wg, wo, wh, wd = [float(w) for w in g[f'{case}_weights']]
	separate = f'{case}_vor_positions' in g
	x = T(g['in_x'])
	kw = {}
	if separate:
		for tag in ('vor', 'div'):
			for nm in NAMES:
				kw[f'{tag}_{nm}_grad'] = torch.zeros_like(getattr(o, nm))
	val, grad = o.get_losses(x, ref_val=T(g['in_ref_val']), weight_val=wv, normals=T(g['in_normals']), weight_boundary=wb,
							 ref_grad=T(g['in_ref_grad']), weight_grad=wg, ref_vor=T(g['in_ref_vor']), weight_vor=wo,
							 ref_hel=T(g['in_ref_hel']), weight_hel=wh, weight_div=wd,
							 stop_gradient=T(g['in_stop_gradient'], torch.int32) if case == 'all' else None, **kw)
	assert rel_err(val.cpu().numpy(), g[f'{case}_val']) < TOL
	assert rel_err(grad.cpu().numpy(), g[f'{case}_grad']) < TOL
	for nm in NAMES:
		assert rel_err(getattr(o, nm).grad.cpu().numpy(), g[f'{case}_direct_{nm}']) < TOL, ('direct', nm)
		if separate:
			for tag in ('vor', 'div'):
				assert rel_err(kw[f'{tag}_{nm}_grad'].cpu().numpy(), g[f'{case}_{tag}_{nm}']) < TOL, (tag, nm)


def test_rk4_and_neighbors_match_reference_golden():
	g = load_golden('ref3d_kernels_f64.npz')
	o = from_golden(g)
	x = T(g['in_x'])
	pos, deform, val, grad = o.advection_rk4(x, float(g['rk4_dt']), pos_only=False)
	for a, k in ((pos, 'rk4_pos'), (deform, 'rk4_deformation'), (val, 'rk4_val'), (grad, 'rk4_grad')):
		assert rel_err(a.cpu().numpy(), g[k]) < TOL, k
	assert rel_err(o.advection_rk4(x, float(g['rk4_dt'])).cpu().numpy(), g['rk4_pos']) < TOL
	np.testing.assert_array_equal(o.get_all_neighbors(x[:3].contiguous()).cpu().numpy(), g['neighbors_mark'])


def synthetic(n, seed=42, tau=5e-3):
	"""BASELINE.md §3 synthetic field: jittered n^3 lattice, s0 + N(0, .1^2) (clipped), random quaternions and values"""
	gen = torch.Generator().manual_seed(seed)
	N = n ** 3
	ax = torch.linspace(0., 1., n)
	P = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
	h = 1. / (n - 1)
	P = (P + (torch.rand(P.shape, generator=gen) - .5) * .5 * h).clamp(0., 1.)
	mgs = 2. * N ** (-1. / 3.)
	s0 = .5 * np.log(-2. * np.log(tau)) - np.log(mgs)
	S = s0 + (torch.randn((N, 3), generator=gen) * .1).clamp(-.2, .2)
	R = torch.randn((N, 4), generator=gen)
	V = torch.randn((N, 3), generator=gen) * .1
	return P.numpy(), S.numpy(), R.numpy(), V.numpy(), mgs, gen


@pytest.mark.parametrize('n,Q', [(10, 1000), (20, 8000), (32, 4096)])
def test_against_oracle_seeded(n, Q):
	"""seeded synthetic fields at sizes the oracle finishes in seconds: hash bit-exact, fields/gradients 1e-5 vs f64"""
	from oracle.oracle import OracleGSR, extended_bounds
	tau = 5e-3
	P, S, R, V, mgs, gen = synthetic(n)
	o = make_fast3d(P, S, R, V, tau, mgs)
	ext = extended_bounds(3, (0., 1.) * 3, mgs)
	orc = OracleGSR(3, ext, P, S, R, V, tau, mgs, precision='f64', nthreads=8)
	cnt, off, sid = o.grid_arrays()
	assert np.float32(o.grid_scale) == np.float32(orc.grid_scale)
	np.testing.assert_array_equal(cnt.cpu().numpy().ravel(), orc.cnt)
	np.testing.assert_array_equal(off.cpu().numpy().ravel(), orc.offset)
	np.testing.assert_array_equal(sid.cpu().numpy(), orc.sorted_id[:orc.n_in])
	X = torch.rand((Q, 3), generator=gen)
	x = X.cuda()
	# samples with a pair inside the borderline band |q - q_max| <= 1e-4 q_max are compared separately
	n_acc, n_band = orc.classify_pairs(X.numpy())
	clean = n_band == 0
	assert clean.mean() > .95
	val, grad = o.get_losses(x)
	oval, ograd = orc.forward(X.numpy())
	assert rel_err(val.cpu().numpy()[clean], oval[clean]) < TOL
	assert rel_err(grad.cpu().numpy()[clean], ograd[clean]) < TOL
	# in-band samples: val is continuous across the cut, grad jumps by O(tau |Sigma^-1 d| |v|)
	assert rel_err(val.cpu().numpy(), oval) < 1e-4
	# RK4 with all outputs
	res = o.advection_rk4(x, -.02, pos_only=False)
	ores = orc.rk4(X.numpy(), -.02, pos_only=False)
	rk_clean = clean.copy()
	for pts in orc.rk4_eval_points(X.numpy(), -.02)[1:]:	# every stage / end point must be free of borderline pairs too
		rk_clean &= orc.classify_pairs(pts)[1] == 0
	assert rk_clean.mean() > .9
	for a, b, nm in zip(res, ores, ('pos', 'deformation', 'val', 'grad')):
		assert rel_err(a.cpu().numpy()[rk_clean], b[rk_clean]) < TOL, nm
		assert rel_err(a.cpu().numpy(), b) < 2e-2, nm
	# backward (project weights) fed with the oracle's own forward totals so that no sign() can flip
	ref_vor = torch.randn((Q, 3), generator=gen) * .1
	ref_hel = torch.randn((Q,), generator=gen) * .1
	acc = {f'{tag}_{nm}_grad': torch.zeros_like(getattr(o, nm)) for tag in ('vor', 'div') for nm in NAMES}
	o.get_losses(x, ref_vor=ref_vor.cuda(), weight_vor=1., ref_hel=ref_hel.cuda(), weight_hel=1., weight_div=1., **acc)
	direct, vor, div = orc.zero_grads(), orc.zero_grads(), orc.zero_grads()
	orc.backward3d(X.numpy(), oval, ograd, ref_vor=ref_vor.numpy(), weight_vor=1., ref_hel=ref_hel.numpy(), weight_hel=1., weight_div=1.,
				   direct=direct, vor=vor, div=div)
	for tag, grp in (('vor', vor), ('div', div)):
		for nm, b in zip(NAMES, grp):
			a = acc[f'{tag}_{nm}_grad'].cpu().numpy()
			assert rel_err(a, b) < 5e-4, (tag, nm, rel_err(a, b))


@pytest.mark.parametrize('n,Q', [(10, 1000), (10, 6000), (20, 8000), (40, 30000)])
def test_backward_at_benchmark_scale_is_1e5_on_the_clean_set(n, Q):
	"""
	Why the test above needs 5e-4 for the backward pass, and what holds at the north-star tolerance.  The L1 losses differentiate to
	sign(residual): a sample whose vorticity / helicity residual is within the forward pass's f32 rounding of zero can take the
	other sign on the GPU than in the float64 oracle, and then a whole +-w term of that sample flips (likewise a pair inside the
	borderline band of the truncation test appears or disappears).  Those samples are identified with the oracle — residuals
	below 1e-4 of the residual scale, pairs within 1e-4 q_max of the cut — and left out: on the remaining samples the GPU
	gradients of every parameter tensor, both gradient sets, agree with the float64 oracle to 1e-5 (f32 accumulation of ~30-60
	signed terms per Gaussian).  With all samples the error is a handful of flipped +-w terms (each 1 / Q of the loss): loosely bounded.
	"""
	from oracle.oracle import OracleGSR, extended_bounds
	tau = 5e-3
	P, S, R, V, mgs, gen = synthetic(n)
	o = make_fast3d(P, S, R, V, tau, mgs)
	orc = OracleGSR(3, extended_bounds(3, (0., 1.) * 3, mgs), P, S, R, V, tau, mgs, precision='f64', nthreads=8)
	X = torch.rand((Q, 3), generator=gen)
	ref_vor = torch.randn((Q, 3), generator=gen) * .1
	ref_hel = torch.randn((Q,), generator=gen) * .1
	oval, ograd = orc.forward(X.numpy())
	om = np.stack((ograd[:, 2, 1] - ograd[:, 1, 2], ograd[:, 0, 2] - ograd[:, 2, 0], ograd[:, 1, 0] - ograd[:, 0, 1]), -1)
	r_vor = np.abs(om - ref_vor.numpy().astype(np.float64))
	r_hel = np.abs((oval * om).sum(-1) - ref_hel.numpy().astype(np.float64))
	risky = (r_vor.min(axis=1) < 1e-4 * np.abs(om).max()) | (r_hel < 1e-4 * np.abs((oval * om).sum(-1)).max()) | (orc.classify_pairs(X.numpy())[1] > 0)
	assert risky.mean() < .05

	def both(sel):
		"""(GPU sets, oracle sets) on the samples `sel`, normalised by the same count"""
		Xs, rv, rh = X[sel].contiguous(), ref_vor[sel].contiguous(), ref_hel[sel].contiguous()
		acc = {f'{tag}_{nm}_grad': torch.zeros_like(getattr(o, nm)) for tag in ('vor', 'div') for nm in NAMES}
		o.get_losses(Xs.cuda(), ref_vor=rv.cuda(), weight_vor=1., ref_hel=rh.cuda(), weight_hel=1., weight_div=1., **acc)
		direct, vor, div = orc.zero_grads(), orc.zero_grads(), orc.zero_grads()
		v, g_ = orc.forward(Xs.numpy())
		orc.backward3d(Xs.numpy(), v, g_, ref_vor=rv.numpy(), weight_vor=1., ref_hel=rh.numpy(), weight_hel=1., weight_div=1., direct=direct, vor=vor, div=div)
		return acc, {'vor': vor, 'div': div}
	keep = torch.from_numpy(~risky)
	acc, orc_sets = both(keep)
	for tag in ('vor', 'div'):
		for nm, b in zip(NAMES, orc_sets[tag]):
			a = acc[f'{tag}_{nm}_grad'].cpu().numpy()
			assert rel_err(a, b) < TOL, (tag, nm, rel_err(a, b))
	# all samples: at most the risky samples' terms differ — each is one sample's share 1 / Q of the loss
	acc, orc_sets = both(torch.ones(Q, dtype=torch.bool))
	for tag in ('vor', 'div'):
		for nm, b in zip(NAMES, orc_sets[tag]):
			a = acc[f'{tag}_{nm}_grad'].cpu().numpy()
			assert rel_err(a, b) < (5e-2 if risky.any() else TOL), (tag, nm, rel_err(a, b), int(risky.sum()))


@pytest.mark.parametrize('n,Q', [(10, 3000), (20, 12000), (30, 40000), (34, 70000), (10, 8192), (10, 16384), (12, 33), (10, 1025)])
def test_hash_paths_agree(n, Q):
	"""the single-launch hash (n <= 16384), the single-CTA stable counting sort of small sample batches (Q <= 16384 on <= 3200
	padded cells, crowded or not), the counting sort (larger, uncrowded cells) and the stable radix sort must produce the
	same cell tables, the same canonical Gaussian order and the same sample order; outputs downstream are then bit-identical"""
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	P, S, R, V, mgs, gen = synthetic(n)
	X = (torch.rand((Q, 3), generator=gen) * 1.3 - .15).cuda()	# some samples outside the domain / the padded grid
	ref_vor = torch.randn((Q, 3), generator=gen).cuda() * .1
	out = []
	for force_radix in (1, 0):
		assert lib.gsr_set_tuning(C.c_int(5), C.c_int(force_radix)) == 0
		try:
			o = make_fast3d(P, S, R, V, 5e-3, mgs)
			e = o._engine
			e.ensure_packed(o._params())
			bins = e.bin_samples(X, True)
			in_grid = int(bins.scs[-1].item())
			val, grad = torch.empty((Q, 3), device='cuda'), torch.empty((Q, 3, 3), device='cuda')
			e.forward(X, val, grad, False, perm=bins)
			acc, mask = e.backward_gather(X, bins.perm, bins.scs, val, grad, (0., 0., 0., 1., 0., 1.), {'ref_vor': ref_vor}, None)
			torch.cuda.synchronize()
			out.append([e.cell_start.cpu().numpy().copy(), e.sorted_id.cpu().numpy().copy(), e.packed.cpu().numpy().copy(), bins.scs.cpu().numpy().copy(),
						bins.perm[:in_grid].cpu().numpy().copy(), np.sort(bins.perm[in_grid:].cpu().numpy()), val.cpu().numpy().copy(), acc[1:].cpu().numpy().copy()])	# set 0 (direct) is not written by this call
		finally:
			lib.gsr_set_tuning(C.c_int(5), C.c_int(0))
	for a, b in zip(*out):
		np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('n,Q', [(10, 600_000), (10, 2_097_152), (34, 2_097_152)])
def test_radix_sort_full_size(n, Q):
	"""the stable radix sort at the benchmark's batch sizes (more than 256 key blocks per digit row, which the small cases of
	test_hash_paths_agree never reach): the sample order must equal a stable sort of keys recomputed here with the kernel's own
	IEEE f32 formula (cell_coord of csrc/common.cuh, sample_key of csrc/hash_small.cuh), and the cell table the key histogram."""
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	P, S, R, V, mgs, gen = synthetic(n)
	X = (torch.rand((Q, 3), generator=gen) * 1.3 - .15).cuda()
	assert lib.gsr_set_tuning(C.c_int(5), C.c_int(1)) == 0
	try:
		o = make_fast3d(P, S, R, V, 5e-3, mgs)
		e = o._engine
		e.ensure_packed(o._params())
		bins = e.bin_samples(X, True)
		torch.cuda.synchronize()
	finally:
		lib.gsr_set_tuning(C.c_int(5), C.c_int(0))
	d = e.desc
	dims = torch.tensor([d.dims[0], d.dims[1], d.dims[2]], device='cuda')
	lo = torch.tensor([d.lo[0], d.lo[1], d.lo[2]], dtype=torch.float32, device='cuda')
	gs = torch.tensor(d.grid_scale, dtype=torch.float32, device='cuda')
	q = (X - lo) / gs
	c = torch.floor(q).long()
	ok = ((c >= -1) & (c <= dims)).all(dim=1)
	pd = dims + 2
	pcell = int(pd.prod().item())
	key = torch.where(ok, ((c[:, 0] + 1) * pd[1] + (c[:, 1] + 1)) * pd[2] + (c[:, 2] + 1), torch.full_like(c[:, 0], pcell))
	scs = bins.scs.long()
	assert scs.numel() == pcell + 1
	hist = torch.bincount(key, minlength=pcell + 1)
	np.testing.assert_array_equal(scs.cpu().numpy(), (torch.cumsum(hist, 0) - hist)[:pcell + 1].cpu().numpy())
	if bins.tiles is not None:	# fine keys: 2 more bits per axis order the samples inside a cell
		sub = torch.clamp(((q - c.float()) * 4.).int(), 0, 3).long()
		key = key * 64 + torch.where(ok, (sub[:, 0] * 4 + sub[:, 1]) * 4 + sub[:, 2], torch.zeros_like(key))
	expect = torch.sort(key, stable=True).indices
	np.testing.assert_array_equal(bins.perm.long().cpu().numpy(), expect.cpu().numpy())


@pytest.mark.parametrize('weights,refs', [((0., 1., 0., 0., 0., 0.), 'normals'), ((0., 0., 0., 1., 0., 1.), 'ref_vor'), ((1., 0., 1., 0., 0., 0.), 'val_grad')])
def test_gather_cta_per_gaussian_equals_warp_per_gaussian(weights, refs):
	"""few Gaussians, many samples each (the 8192 boundary samples on 1000 Gaussians): the backward gather gives a whole CTA to a
	Gaussian.  Same visits as the warp-per-Gaussian kernel, summed in a different grouping: 1e-5 of the largest entry per set."""
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	P, S, R, V, mgs, gen = synthetic(10)
	Q = 8192
	X = torch.rand((Q, 3), generator=gen).cuda()
	X[: Q // 2, 0] = 0.	# half of them on a face, like the boundary batch
	aux = {'normals': {'normals': torch.nn.functional.normalize(torch.randn((Q, 3), generator=gen), dim=1).cuda()},
		   'ref_vor': {'ref_vor': torch.randn((Q, 3), generator=gen).cuda() * .1},
		   'val_grad': {'ref_val': torch.randn((Q, 3), generator=gen).cuda() * .1, 'ref_grad': torch.randn((Q, 3, 3), generator=gen).cuda() * .1}}[refs]
	out = []
	for max_n in (4096, 0):
		assert lib.gsr_set_tuning(C.c_int(7), C.c_int(max_n)) == 0
		try:
			o = make_fast3d(P, S, R, V, 5e-3, mgs)
			e = o._engine
			e.ensure_packed(o._params())
			bins = e.bin_samples(X, True)
			val, grad = torch.empty((Q, 3), device='cuda'), torch.empty((Q, 3, 3), device='cuda')
			e.forward(X, val, grad, False, perm=bins)
			acc, mask = e.backward_gather(X, bins.perm, bins.scs, val, grad, weights, aux, None)
			torch.cuda.synchronize()
			out.append((acc.cpu().numpy().copy(), mask))
		finally:
			lib.gsr_set_tuning(C.c_int(7), C.c_int(4096))
	assert out[0][1] == out[1][1]
	for s in range(3):
		if out[0][1] & (1 << s):
			a, b = out[0][0][s], out[1][0][s]
			assert np.abs(b).max() > 0.
			assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), s
