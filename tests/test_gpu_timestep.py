"""
GPU tests of the measured fixed-work time step (timestep3d.LeapfrogTimestep = advance3d.advance_frame with the stock generators,
SURVEY 8d): the captured-graph execution with its forked streams must produce exactly what the plain eager, single-order execution
produces (the samples come from counter-based Philox kernels keyed by the device-resident sample clock, every sum is a fixed
tree) — also when a graph captured in one frame is replayed in later frames, after the previous field's grid_scale has moved —
and a step must be reproducible from a reset.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(use_graph, steps=2, iters=24):
	from gaussian_fluids_code_b200 import gsr3d, timestep3d
	gsr3d.device = torch.device('cuda', 0)
	ts = timestep3d.LeapfrogTimestep(n=8, iters=iters, test_res=24, check_iter=8, use_graph=use_graph)
	outs = []
	for _ in range(steps):
		vor, div = ts.step()
		torch.cuda.synchronize()
		outs.append([p.detach().cpu().numpy().copy() for p in ts.cur._params()] + [vor.cpu().numpy().copy(), div.cpu().numpy().copy(), ts.last_test.cpu().numpy().copy()])
	return ts, outs


@pytest.mark.parametrize('steps', [2, 5])
def test_graph_replay_equals_eager_execution(steps):
	"""5 frames without a reset: each orientation's graph is captured in its first frame and REPLAYED in its later ones, while the
	previous field's hash (grid_scale, cell table, packed records) has been rebuilt in between — the kernels read grid_scale from
	the field's persistent device scalar, so a replay bins with the current value"""
	_, a = run(True, steps=steps)
	_, b = run(False, steps=steps)
	for sa, sb in zip(a, b):
		for x, y in zip(sa, sb):
			np.testing.assert_array_equal(x, y)
	assert np.isfinite(a[-1][0]).all() and np.abs(a[-1][0] - a[0][0]).max() > 0	# the field moved between the steps


def test_hoisted_test_reference_equals_the_recomputed_one(monkeypatch):
	"""the pull-back reference on the fixed test lattice is evaluated by the first test pass of a frame and reused by the later passes
	of that frame (the reference, 3D/advance.py:290-291, recomputes it): bitwise the same test losses, parameters and output fields
	over several frames — which also shows that a new frame (new previous field, same objects and pointers) evaluates it anew"""
	from gaussian_fluids_code_b200 import advance3d
	monkeypatch.setattr(advance3d, 'HOIST_TEST_REFERENCE', True)
	ts, a = run(True, steps=4)
	fp = next(iter(ts.cur.__dict__.get('_pipelines', {}).values()), None) or next(iter(ts.new._pipelines.values()))
	assert fp.reference_reused	# the last test pass of the last frame did not launch the pull-back
	monkeypatch.setattr(advance3d, 'HOIST_TEST_REFERENCE', False)
	ts, b = run(True, steps=4)
	fp = next(iter(ts.cur.__dict__.get('_pipelines', {}).values()), None) or next(iter(ts.new._pipelines.values()))
	assert not fp.reference_reused
	for sa, sb in zip(a, b):
		for x, y in zip(sa, sb):
			np.testing.assert_array_equal(x, y)
	assert np.abs(a[-1][-1] - a[0][-1]).max() > 0	# the test losses differ between the frames


def test_step_is_reproducible_after_reset():
	ts, first = run(True, steps=1)
	ts.reset()
	vor, div = ts.step()
	torch.cuda.synchronize()
	again = [p.detach().cpu().numpy() for p in ts.cur._params()]
	for x, y in zip(first[0][:4], again):
		np.testing.assert_array_equal(x, y)


def project_phase(pipelined, ordered=False, iters=6):
	"""the project phase of one time step, iteration by iteration, in the plain order (samples, hashes, pull-back, forward, gathers,
	step + rebuild) or in the pipelined order of ShardedProjector (next samples and pull-back prepared behind the step)"""
	from gaussian_fluids_code_b200 import gsr3d, timestep3d
	gsr3d.device = torch.device('cuda', 0)
	ts = timestep3d.LeapfrogTimestep(n=8, iters=iters, test_res=16, check_iter=iters, use_graph=False)
	cur, new = ts.cur, ts.new
	with torch.no_grad():
		pos = cur.advection_rk4(new.positions.detach(), ts.dt)
		pos.clamp_(torch.tensor([new.x_min, new.y_min, new.z_min], device='cuda'), torch.tensor([new.x_max, new.y_max, new.z_max], device='cuda'))
		new.positions.copy_(pos)
	new.zero_grad()
	from gaussian_fluids_code_b200 import advance3d
	ref = advance3d.AdvectedCovectorField(cur, cur, ts.dt, 0., 1., 0., 1., 0., 1.)
	fp = timestep3d.ShardedProjector(new, ref, ts.boundary_lambda, ts.N, ts.Qb)
	if ordered:
		fp.ORDERED_REF_MIN_Q = 1	# the pull-back walks the samples in cell order (the large-batch form)
	if pipelined:
		fp.set_samplers(fp.draw_samples, fp.draw_boundary)
		fp.prime()
		for k in range(iters):
			fp.iterate(None, join_all=(k % 3 == 2))
	else:
		for k in range(iters):
			fp.iterate(fp.draw_samples(), fp.draw_boundary())
	fp.finish()
	torch.cuda.synchronize()
	return [p.detach().cpu().numpy().copy() for p in new._params()] + [np.array(fp.stepper.scalars()[:14])]


@pytest.mark.parametrize('ordered', [False, True])
def test_pipelined_iterations_equal_plain_order(ordered):
	"""same kernels on the same inputs in both orders: parameters and optimiser state bit for bit"""
	a, b = project_phase(True, ordered), project_phase(False)
	for x, y in zip(a, b):
		np.testing.assert_array_equal(x, y)
	assert a[-1][0] == 6.


def test_project_api_takes_the_pipelined_path_and_matches_the_eager_api():
	"""advance3d.project with the stock generator objects runs on the captured pipeline; handing it the same samples through plain
	callables (the eager fused path) gives the same parameters bit for bit"""
	from gaussian_fluids_code_b200 import advance3d, gsr3d, timestep3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	P, S, R, V, mgs, _ = synthetic_field(8)
	box = (0., 1.) * 3
	test_gen = advance3d.LatticeGenerator(*box, 16, 16, 16)
	res = []
	for mode in ('pipelined', 'eager'):
		cur, new = make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs)
		advance3d.advect_covector_field(new, cur, .02, new.x_min, new.x_max, new.y_min, new.y_max, new.z_min, new.z_max)
		ref = advance3d.AdvectedCovectorField(cur, cur, .02, *box)
		if mode == 'pipelined':
			ep = advance3d.project(new, ref, *box, advance3d.BoxSampler(*box), test_gen, boundary_generator=advance3d.BoxSurfaceSampler(*box), boundary_lambda=10.,
								   batch_size=2048, max_epoch=40, patience=10 ** 9, verbose=0, check_iter=20)
			fp = next(iter(new._pipelines.values()))
			assert isinstance(fp, timestep3d.ShardedProjector) and fp.graph is not None
		else:
			# replay the device sampler's draws through ordinary generator callables
			probe = timestep3d.ShardedProjector(make_fast3d(P, S, R, V, 5e-3, mgs), ref, 10., new.N, 2048)
			clock = torch.zeros(1, device='cuda')
			e = probe.gv._engine

			def data_gen(n, gs):
				return e.sample_box(box, torch.empty((new.N, 3), device='cuda'), 42, 0, clock)

			def boundary_gen(n):
				d, nm = torch.empty((2048, 3), device='cuda'), torch.empty((2048, 3), device='cuda')
				e.sample_box_surface(box, d, nm, 42, 1, clock)
				clock.add_(1.)
				return d, nm
			ep = advance3d.project(new, ref, *box, data_gen, test_gen, boundary_generator=boundary_gen, boundary_lambda=10., batch_size=2048, max_epoch=40,
								   patience=10 ** 9, verbose=0, check_iter=20)
		assert ep == 40
		torch.cuda.synchronize()
		res.append([p.detach().cpu().numpy().copy() for p in new._params()])
	for a, b in zip(*res):
		np.testing.assert_array_equal(a, b)


def test_advance_loop_writes_the_reference_files(tmp_path):
	"""advance(): two frames of 3D/advance.py:381-393 with the files the reference writes per frame"""
	import os
	from gaussian_fluids_code_b200 import advance3d, gsr3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	P, S, R, V, mgs, _ = synthetic_field(6)
	cur, new = make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs)
	seen = []
	a, b = advance3d.advance(cur, new, 0., 1., 0., 1., 0., 1., .02, .04 - 1e-9, boundary_generator=advance3d.BoxSurfaceSampler(0., 1., 0., 1., 0., 1.), boundary_lambda=10.,
							 visualize_res=(12, 12, 12), out_dir=str(tmp_path), max_epoch=20, patience=10 ** 9, verbose=0, check_iter=10, batch_size=1024,
							 on_frame=lambda k, f, vor, div: seen.append((k, float(vor.mean()), float(div.abs().mean()))))
	assert [k for k, _, _ in seen] == [1, 2] and a is cur and b is new	# two swaps
	for k in (1, 2):
		for name in (f'vorticity_{k}.vti', f'divergence_{k}.vti', f'gaussian_velocity_{k}.pt'):
			assert os.path.getsize(os.path.join(str(tmp_path), name)) > 0
	d = torch.load(os.path.join(str(tmp_path), 'gaussian_velocity_2.pt'))
	assert set(d) >= {'positions', 'scalings', 'rotations', 'values', 'clamp_threshold', 'min_grid_scale', 'domain_range'}
	np.testing.assert_array_equal(d['positions'].detach().cpu().numpy(), a.positions.detach().cpu().numpy())


def test_project_with_graph_safe_callables_replays_from_graphs_and_equals_eager():
	"""advance3d.project with generators that are NOT the stock sampler objects — the scene's own boundary sampler
	(init_cond3d.make_boundary_sampler: box, or box + obstacle mesh) and a plain callable — but carry `graph_safe = True`: the eager
	fused iteration is replayed from CUDA graphs (graphloop.py), bit for bit what the same call computes eagerly"""
	from gaussian_fluids_code_b200 import advance3d, graphloop, gsr3d, init_cond3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	P, S, R, V, mgs, _ = synthetic_field(8)
	box = (0., 1.) * 3
	test_gen = advance3d.LatticeGenerator(*box, 16, 16, 16)
	bgen = init_cond3d.make_boundary_sampler('leapfrog')
	assert bgen.graph_safe
	res = []
	for use_graph in (False, True):
		torch.manual_seed(21)
		cur, new = make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs)
		advance3d.advect_covector_field(new, cur, .02, new.x_min, new.x_max, new.y_min, new.y_max, new.z_min, new.z_max)
		ref = advance3d.AdvectedCovectorField(cur, cur, .02, *box)
		dgen = lambda n, gv: torch.rand_like(gv.positions.detach())
		dgen.graph_safe = True
		g0 = graphloop.GRAPH_LAUNCHES
		ep = advance3d.project(new, ref, *box, dgen, test_gen, boundary_generator=bgen, boundary_lambda=10., batch_size=2048, max_epoch=60, patience=10 ** 9,
							   verbose=0, check_iter=20, use_graph=use_graph)
		assert ep == 60 and (graphloop.GRAPH_LAUNCHES > g0) == use_graph
		assert not hasattr(new, '_pipelines') or not new._pipelines	# not the stock-sampler pipeline
		res.append([p.detach().cpu().numpy().copy() for p in new._params()] + [np.float64(new.grid_scale)])
	for a, b in zip(*res):
		np.testing.assert_array_equal(a, b)
