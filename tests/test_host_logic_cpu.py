"""
Host-side logic that needs no GPU: the stopping rule of the optimisation loops, the exchange buffer layout, the lattice shards of
both scaling modes, the boundary sampler tables — each against an independent statement of the reference's rule.
"""
import numpy as np
import pytest


def reference_stop_epoch(series, patience, check_iter, thresholds):
	"""the bookkeeping of 2D/advance.py:131-155 / 3D/advance.py:289-314, written out: returns the epoch at which the loop breaks"""
	best = [np.inf] * len(thresholds)
	stale = [0] * len(thresholds)
	for k, row in enumerate(series):
		for j, (v, thr) in enumerate(zip(row, thresholds)):
			if v < best[j] * (1. - thr):
				best[j], stale[j] = v, 0
			else:
				stale[j] += check_iter
		if all(s >= patience for s in stale):
			return (k + 1) * check_iter
	return None


@pytest.mark.parametrize('seed', range(6))
def test_early_stop_rule_is_the_references(seed):
	from gaussian_fluids_code_b200 import reseed
	rng = np.random.default_rng(seed)
	n = 60
	series = np.abs(np.cumsum(rng.normal(-.01, .02, (n, 2)), axis=0) + 1.)	# noisy, slowly improving, sometimes stalling
	for patience, check_iter, thr in ((500, 100, (1e-3, 1e-3)), (300, 100, (1e-3, 1e-2)), (50, 20, (1e-3, 1e-2))):
		rule = reseed.EarlyStop(('a', 'b'), patience, check_iter, thresholds={'a': thr[0], 'b': thr[1]})
		got = None
		for k, (a, b) in enumerate(series):
			if rule.update({'a': float(a), 'b': float(b)}):
				got = (k + 1) * check_iter
				break
		assert got == reference_stop_epoch(series, patience, check_iter, thr)


def test_domain_boundary_sampler_tables_cover_the_four_edges():
	"""the table form of sample_on_domain_boundary_2 (2D/init_cond.py:306-325): every point on the edge its normal names, arc-length
	uniform (perimeter-weighted edges), target 0"""
	import torch
	from gaussian_fluids_code_b200 import gsr2d, init_cond2d
	gsr2d.device = torch.device('cpu')
	sc = init_cond2d.Scene2D('leapfrog')
	torch.manual_seed(0)
	data, nrm, val = sc._on_domain_boundary_2(20000)
	x_min, x_max, y_min, y_max = sc.advance_domain
	d, n = data.numpy(), nrm.numpy()
	assert (val == 0).all()
	on = {(0., -1.): np.isclose(d[:, 1], y_min), (1., 0.): np.isclose(d[:, 0], x_max), (0., 1.): np.isclose(d[:, 1], y_max), (-1., 0.): np.isclose(d[:, 0], x_min)}
	seen = np.zeros(len(d), bool)
	for (nx, ny), mask in on.items():
		sel = (n[:, 0] == nx) & (n[:, 1] == ny)
		assert sel.any() and mask[sel].all()
		seen |= sel
	assert seen.all()
	assert (d[:, 0] >= x_min - 1e-6).all() and (d[:, 0] <= x_max + 1e-6).all() and (d[:, 1] >= y_min - 1e-6).all() and (d[:, 1] <= y_max + 1e-6).all()
	per = 2. * ((x_max - x_min) + (y_max - y_min))
	frac = np.array([((n[:, 0] == a) & (n[:, 1] == b)).mean() for a, b in ((0., -1.), (1., 0.), (0., 1.), (-1., 0.))])
	want = np.array([x_max - x_min, y_max - y_min, x_max - x_min, y_max - y_min]) / per
	assert np.abs(frac - want).max() < 1.5e-2


def test_test_pass_reference_is_reused_only_for_fixed_points_within_a_phase(monkeypatch):
	"""FusedProjector.evaluate (host logic, engines stubbed): the pull-back reference of the test points is evaluated again unless the
	caller vouches that the points are fixed, the phase is the same (restart() drops the key), the previous field and time step
	are the same objects / values, and GSR_HOIST_TEST_REFERENCE is on"""
	import types
	import torch
	from gaussian_fluids_code_b200 import advance3d

	calls = []

	class Engine:
		def bin_samples(self, data, need_cells):
			return types.SimpleNamespace(perm=None)

		def advected_vorticity(self, data, dt, ref_vor, ref_hel, perm=None):
			calls.append(float(dt))

		def forward(self, data, val, grad, accumulate=False, perm=None):
			pass

		def sample_losses(self, val, grad, refs, Q):
			return torch.ones(3) * Q

	bufs = {}
	fp = types.SimpleNamespace(gv=types.SimpleNamespace(_engine=Engine()), ref=types.SimpleNamespace(velocity_field=types.SimpleNamespace(_engine=Engine()), time_step=.02),
							   _tmp=lambda name, shape: bufs.setdefault((name, shape), torch.zeros(shape)))
	ev = lambda data, **kw: advance3d.FusedProjector.evaluate(fp, data, **kw)
	lattice, other = torch.zeros((8, 3)), torch.zeros((8, 3))
	monkeypatch.setattr(advance3d, 'HOIST_TEST_REFERENCE', True)
	ev(lattice, fixed=True); ev(lattice, fixed=True); ev(lattice, fixed=True)
	assert len(calls) == 1 and fp.reference_reused
	ev(lattice)	# a plain callable's points: never reused, and the next fixed call may not trust the buffer either
	assert len(calls) == 2 and not fp.reference_reused
	ev(lattice, fixed=True); ev(lattice, fixed=True)
	assert len(calls) == 3
	ev(other, fixed=True)	# other points
	assert len(calls) == 4
	fp._test_ref_key = None	# what ShardedProjector.restart() does at the start of a phase
	ev(other, fixed=True)
	assert len(calls) == 5
	fp.ref.time_step = .01	# another time step
	ev(other, fixed=True)
	assert len(calls) == 6 and calls[-1] == -.01
	monkeypatch.setattr(advance3d, 'HOIST_TEST_REFERENCE', False)
	ev(other, fixed=True); ev(other, fixed=True)
	assert len(calls) == 8 and not fp.reference_reused
	assert advance3d.LatticeGenerator.fixed_points is True
