"""
GPU tests of clone_velocity_field's split / reseed path (3D/advance.py:51-165, 2D/advance.py:58-158):
Gaussians whose axis ratio passes the threshold are replaced by two samples of their own distribution, the untouched
Gaussians that are not neighbours of a new one are frozen (stop_gradient), and the new ones are refitted to the old field.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_clone_splits_and_refits_3d():
	from gaussian_fluids_code_b200 import advance3d, gsr3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	torch.manual_seed(1)
	P, S, R, V, mgs, gen = synthetic_field(8)
	S = S.copy()
	stretched = np.arange(0, 512, 37)	# 14 Gaussians with axis ratio e^0.8 > 2
	S[stretched, 0] += .5
	S[stretched, 1] -= .3
	old = make_fast3d(P, S, R, V, 5e-3, mgs)
	new = make_fast3d(P, S, R, V, 5e-3, mgs)
	x = torch.rand((8000, 3), generator=gen).cuda()
	before = old(x).clone()
	frozen_probe = new.positions.detach().clone()
	advance3d.clone_velocity_field(new, old, 0., 1., 0., 1., 0., 1., lambda n, gs, restrict=None: torch.rand((4096, 3), device='cuda'), lambda gs: x,
								   max_epoch=60, verbose=0)
	# the 3D reference splits repeatedly while any ratio is >= 2 (3D/advance.py:63-87), so a very stretched Gaussian splits twice
	assert 512 + len(stretched) <= new.N <= 512 + 3 * len(stretched) and new.positions.shape == (new.N, 3) and new.rotations.shape == (new.N, 4)
	ratio = torch.exp(new.scalings.max(dim=-1).values - new.scalings.min(dim=-1).values)
	assert float(ratio.max()) < 2.2		# the split halves the stretched axis; a short refit cannot undo that
	after = new(x)
	assert torch.isfinite(after).all()
	err = float((after - before).abs().mean() / before.abs().mean())
	assert err < .35, err			# the cloned field still represents the old one
	# Gaussians far from every new one are frozen: bit-identical parameters
	kept = np.setdiff1d(np.arange(512), stretched)
	same = (new.positions.detach()[:len(kept)] == frozen_probe[kept]).all(dim=1)
	assert 0 < int(same.sum()) < len(kept)	# some frozen, the neighbours of the new Gaussians were optimised
	# the old field is untouched and both objects keep working through the generic API
	assert torch.equal(old(x), before)


def test_clone_splits_and_refits_2d():
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	from gaussian_fluids_code_b200.init_cond2d import Scene2D
	gsr2d.device = torch.device('cuda', 0)
	torch.manual_seed(2)
	sc = Scene2D('taylor_green')
	x0, x1, y0, y1 = sc.scaled(sc.initialize_domain)
	pts = gsr2d.get_grid_points(x0, x1, y0, y1, 16, 16).cpu().numpy()
	fields = []
	for _ in range(2):
		o = gsr2d.GaussianSplattingFast(x0, x1, y0, y1, pts, dim=2)
		with torch.no_grad():
			o.values.copy_(sc.target_velocity(o.positions.detach()) * .3)
			o.scalings[::29, 0] += .45	# axis ratio e^0.45 > 1.5
		o.zero_grad()
		fields.append(o)
	old, new = fields
	n_split = len(range(0, 256, 29))
	x = sc.test_generator()
	before = old(x).clone()
	advance2d.clone_velocity_field(new, old, lambda n, gs, restrict=None: sc.data_generator(gs), lambda gs: x, max_epoch=60, verbose=0)
	assert new.N == 256 + n_split and new.rotations.shape == (new.N,)
	after = new(x)
	assert torch.isfinite(after).all()
	assert float((after - before).abs().mean() / before.abs().mean()) < .35


def test_clone2d_matches_reference_golden():
	"""the reference's OWN 2D clone_velocity_field (2D/advance.py:58-158, run through the Taichi shim with its MultivariateNormal
	draws recorded: tests/golden/make_golden_clone2d.py) against the device split + refit: the field right after the split, the
	stop_gradient mask after the neighbours of the new Gaussians are unfrozen (integer-exact), the total .grad / metric / lrs at
	every refit step and the parameters after 1 and 3 steps"""
	from helpers import NAMES, load_golden, rel_err
	from test_gpu_gradients_golden import T, check_steps, record_steps
	from gaussian_fluids_code_b200 import advance2d, gsr2d
	gsr2d.device = torch.device('cuda', 0)
	g = load_golden('ref2d_clone.npz')
	dom = [float(v) for v in g['domain']]

	def field():
		gv = gsr2d.GaussianSplattingFast(*dom, g['positions'], dim=2)
		with torch.no_grad():
			gv.scalings.copy_(T(g['scalings'])); gv.rotations.copy_(T(g['rotations']).reshape(gv.rotations.shape)); gv.values.copy_(T(g['values']))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv
	for epochs in (1, 3):
		src, res = field(), field()
		snap, masks, rec = {}, [], {}
		orig_unfreeze, orig_get_losses = res.unfreeze, res.get_losses

		def unfreeze():
			for nm in NAMES:
				snap[nm] = getattr(res, nm).detach().cpu().numpy().copy()
			return orig_unfreeze()

		def get_losses(x, *a, **kw):
			if kw.get('stop_gradient') is not None:
				masks.append(kw['stop_gradient'].detach().cpu().numpy().copy())
			return orig_get_losses(x, *a, **kw)
		res.unfreeze, res.get_losses = unfreeze, get_losses
		record_steps(res, rec)
		datas = iter([T(x) for x in g['samples']])
		advance2d.clone_velocity_field(res, src, lambda n, gs, restrict=None: next(datas), lambda gs: T(g['test_points']), batch_size=g['samples'].shape[1],
									   max_epoch=epochs, patience=500, verbose=0, normals=T(g['normals']))
		assert res.N == g['split_positions'].shape[0]
		for nm in NAMES:
			want = g[f'split_{nm}']
			assert rel_err(snap[nm].reshape(want.shape), want) < 2e-6, nm
		np.testing.assert_array_equal(masks[0].astype(bool), g['stop_gradient'].astype(bool))
		check_steps(g, rec, epochs, sets=False)
		for nm in NAMES:
			want = g[f'after{epochs}_{nm}']
			got = getattr(res, nm).detach().cpu().numpy().reshape(want.shape)
			assert rel_err(got, want) < 1e-5, nm
			d_ref, d_got = want - g[f'split_{nm}'], got - g[f'split_{nm}'].reshape(want.shape)
			assert rel_err(d_got, d_ref) < 2e-2, (nm, rel_err(d_got, d_ref))


@pytest.mark.parametrize('D', [3, 2])
def test_device_split_matches_oracle_and_draws_the_right_distribution(D):
	"""gsr_split_flags + gsr_split_apply against oracle.split_gaussians for given normal draws (3D rule: ratio >= 2, children
	clamped, repeated rounds handled by the caller; 2D rule: ratio >= 1.5), and — with the built-in Philox draws — the children's
	sample mean and covariance against mu and Sigma of their parents"""
	from helpers import rel_err
	import oracle.oracle as orc
	from gaussian_fluids_code_b200 import gsr2d, gsr3d, reseed
	gsr3d.device = gsr2d.device = torch.device('cuda', 0)
	gen = torch.Generator().manual_seed(5)
	N = 4000
	P = torch.rand((N, D), generator=gen) * .6 + .2
	S = torch.randn((N, D), generator=gen) * .25 + 3.
	R = torch.randn((N, 4), generator=gen) if D == 3 else (torch.rand((N,), generator=gen) * 6.28 - 3.14)
	V = torch.randn((N, D), generator=gen)

	class F_:
		pass
	f = F_()
	f.positions, f.scalings, f.rotations, f.values, f.N = P.cuda(), S.cuda(), R.cuda(), V.cuda(), N
	ratio = torch.exp(S.max(-1).values - S.min(-1).values)
	ns = int((ratio >= (2. if D == 3 else 1.5)).sum())
	assert 100 < ns < N
	z = torch.randn((2, ns, D), generator=gen)
	box = (0., 1.) * D if D == 3 else None
	n, flags = reseed.split_once(f, D, clamp_box=box, normals=z)
	assert n == ns and int(flags.sum()) == ns
	want = orc.split_gaussians(D, P.numpy(), S.numpy(), R.numpy(), V.numpy(), z.numpy(), clamp_box=box)
	for got, w in zip((f.positions, f.scalings, f.rotations, f.values), want[:4]):
		assert rel_err(got.detach().cpu().numpy().reshape(w.shape), w) < 5e-6
	# Philox draws: many children of ONE parent distribution -> mean and covariance
	M = 20000
	one = F_()
	s1 = torch.tensor([3.0, 3.9, 3.3][:D]) if D == 3 else torch.tensor([3.0, 3.6])
	one.positions = torch.full((M, D), .5).cuda()
	one.scalings = s1.repeat(M, 1).cuda()
	one.rotations = (torch.tensor([.3, -.5, .7, .2]).repeat(M, 1) if D == 3 else torch.full((M,), .4)).cuda()
	one.values = torch.zeros((M, D)).cuda()
	one.N = M
	n, _ = reseed.split_once(one, D, seed=11)
	assert n == M
	kids = one.positions.detach()[-2 * M:].double().cpu().numpy()
	# Sigma of the parent from the oracle's own construction
	q = one.rotations.detach()[:1].double().cpu().numpy()
	if D == 3:
		q = q / np.sqrt((q ** 2).sum())
		r, a, b, c = q[0]
		Rm = np.array([[1 - 2 * (b * b + c * c), 2 * (a * b - r * c), 2 * (a * c + r * b)], [2 * (a * b + r * c), 1 - 2 * (a * a + c * c), 2 * (b * c - r * a)],
					   [2 * (a * c - r * b), 2 * (b * c + r * a), 1 - 2 * (a * a + b * b)]])
	else:
		Rm = np.array([[np.cos(.4), -np.sin(.4)], [np.sin(.4), np.cos(.4)]])
	Sigma = Rm @ np.diag(np.exp(-2. * s1.double().numpy())) @ Rm.T
	assert np.abs(kids.mean(0) - .5).max() < 4. * np.sqrt(Sigma.max() / (2 * M))
	assert np.abs(np.cov(kids.T) - Sigma).max() < .05 * np.abs(Sigma).max()
