"""
GPU tests of the measured fixed-work time step (timestep3d.LeapfrogTimestep): the captured-graph execution with
its three forked streams must produce exactly what the plain eager, single-order execution produces (the samples come from
counter-based Philox kernels keyed by the device-resident iteration number, every sum is a fixed tree), and a step must be
reproducible from a reset.
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


def test_graph_replay_equals_eager_execution():
	_, a = run(True)
	_, b = run(False)
	for sa, sb in zip(a, b):
		for x, y in zip(sa, sb):
			np.testing.assert_array_equal(x, y)
	assert np.isfinite(a[-1][0]).all() and np.abs(a[-1][0] - a[0][0]).max() > 0	# the field moved between the steps


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
		pos.clamp_(ts._lo, ts._hi)
		new.positions.copy_(pos)
	new.zero_grad()
	fp = ts._projector(new, cur)['fp']
	if ordered:
		fp.ORDERED_REF_MIN_Q = 1	# the pull-back walks the samples in cell order (the large-batch form)
	if pipelined:
		fp.set_samplers(lambda: ts._samples(fp), lambda: ts._boundary(fp))
		fp.prime()
		for k in range(iters):
			fp.iterate(None, join_all=(k % 3 == 2))
	else:
		for k in range(iters):
			fp.iterate(ts._samples(fp), ts._boundary(fp))
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
