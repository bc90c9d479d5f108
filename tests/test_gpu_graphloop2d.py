"""
The 2D optimisation loops replayed from CUDA graphs (graphloop.py) against the same loops run eagerly: identical kernels, identical
random streams (torch's CUDA generator advances the same way in both), so the parameters must agree BITWISE — for a projection
phase with both boundary samplers (karman: obstacle values + wall / inlet / outlet normals) and for the initial fit.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
NAMES = ('positions', 'scalings', 'rotations', 'values')


def small_scene(name):
	from gaussian_fluids_code_b200 import gsr2d, init_cond2d
	gsr2d.device = torch.device('cuda', 0)
	scene = init_cond2d.Scene2D(name)
	scene.particle_count = (60, 12) if name == 'karman' else (24, 24)
	scene.visualize_res = (64, 32)
	return scene


@pytest.mark.parametrize('name', ['karman', 'leapfrog'])
def test_graphed_project_equals_eager_bitwise(name):
	from gaussian_fluids_code_b200 import advance2d, graphloop
	out = {}
	for use_graph in (False, True):
		scene = small_scene(name)
		torch.manual_seed(11)
		gv = advance2d.simulation_initialize(scene, max_epoch=40, verbose=0, project_epochs=0, use_graph=use_graph)
		src = advance2d.simulation_initialize(scene, max_epoch=0, verbose=0, project_epochs=0)
		with torch.no_grad():
			for nm in NAMES:
				getattr(src, nm).copy_(getattr(gv, nm))
		src.zero_grad()
		ref = advance2d.AdvectedCovectorField(src, src, .05, domain=scene.scaled(scene.advance_domain))
		gen = lambda n, gs, restrict=None: scene.data_generator(gs)
		gen.graph_safe = True
		b1, b2 = scene.boundary_samplers
		g0 = graphloop.GRAPH_LAUNCHES
		epochs = advance2d.project(gv, ref, gen, lambda gs: scene.test_generator(), boundary_generator_1=b1, boundary_generator_2=b2, boundary_lambda=1.,
								   max_epoch=130, patience=10 ** 9, verbose=0, use_graph=use_graph)
		assert epochs == 130
		assert (graphloop.GRAPH_LAUNCHES > g0) == use_graph
		out[use_graph] = [getattr(gv, nm).detach().clone() for nm in NAMES] + [torch.tensor(gv.grid_scale)]
	for a, b in zip(out[False], out[True]):
		assert torch.isfinite(a).all() and torch.equal(a.cpu(), b.cpu())


def test_default_turns_the_graph_on_only_for_graph_safe_generators():
	from gaussian_fluids_code_b200 import advance2d, graphloop
	scene = small_scene('leapfrog')
	torch.manual_seed(5)
	gv = advance2d.simulation_initialize(scene, max_epoch=20, verbose=0)
	src = advance2d.simulation_initialize(scene, max_epoch=0, verbose=0)
	ref = advance2d.AdvectedCovectorField(src, src, .025, domain=scene.scaled(scene.advance_domain))
	plain = lambda n, gs, restrict=None: scene.data_generator(gs)	# unmarked: may be stateful for all project() knows
	g0 = graphloop.GRAPH_LAUNCHES
	advance2d.project(gv, ref, plain, lambda gs: scene.test_generator(), max_epoch=40, patience=10 ** 9, verbose=0)
	assert graphloop.GRAPH_LAUNCHES == g0
	marked = lambda n, gs, restrict=None: scene.data_generator(gs)
	marked.graph_safe = True
	advance2d.project(gv, ref, marked, lambda gs: scene.test_generator(), max_epoch=40, patience=10 ** 9, verbose=0)
	assert graphloop.GRAPH_LAUNCHES > g0
