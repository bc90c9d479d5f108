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


def test_cached_projector_reuses_its_graph_across_frames_and_changes_nothing():
	"""three frames clone -> advect -> project on a small scene: with the projector (and its captured graph) kept on the field object
	across frames the parameters are bitwise those of projectors built afresh every frame, and the second frame on an
	orientation really is a cache hit (same GraphedLoop object, no new capture)"""
	from gaussian_fluids_code_b200 import advance2d
	out = {}
	for cache in (False, True):
		import bench_more
		scene = small_scene('leapfrog')
		torch.manual_seed(3)
		a, b = bench_more.synthetic_field2d(scene), bench_more.synthetic_field2d(scene)	# near-isotropic Gaussians: no frame splits or drops any
		n0 = a.N
		gen = lambda n, gs, restrict=None: scene.data_generator(gs)
		gen.graph_safe = True
		test = lambda gs: scene.test_generator()
		b1, b2 = scene.boundary_samplers
		loops = []
		cur, new = a, b
		for frame in range(4):
			advance2d.clone_velocity_field(new, cur, gen, test, max_epoch=50, verbose=0)
			advance2d.advect_covector_field(new, cur, .025)
			ref = advance2d.AdvectedCovectorField(cur, cur, .025, domain=scene.scaled(scene.advance_domain))
			advance2d.project(new, ref, gen, test, boundary_generator_1=b1, boundary_generator_2=b2, boundary_lambda=1., max_epoch=60, patience=10 ** 9, verbose=0,
							  check_iter=20, cache=cache)
			store = new.__dict__.get('_pipelines2d', {})
			loops.append(id(next(iter(store.values()))[1]) if store else None)
			cur, new = new, cur
		assert cur.N == new.N == n0
		if cache:
			assert loops[0] is not None and loops[2] == loops[0] and loops[3] == loops[1] and loops[0] != loops[1]
		else:
			assert loops == [None] * 4
		out[cache] = [getattr(cur, nm).detach().clone() for nm in NAMES] + [getattr(new, nm).detach().clone() for nm in NAMES]
	for x, y in zip(out[False], out[True]):
		assert torch.isfinite(x).all() and torch.equal(x, y)
