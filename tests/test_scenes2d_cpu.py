"""
Host logic of the 2D scenarios (init_cond2d.Scene2D) against tests/golden/ref2d_scenes.npz, which the reference's own
2D/init_cond.py produced scene by scene (tests/golden/make_golden_scenes2d.py): scale factors, domains, analytic fields and
Jacobians, and the boundary samplers — the reference draws with torch.rand, so the same CPU seed must reproduce its samples.
"""
import numpy as np
import pytest
import torch

from helpers import load_golden

SCENES = ('taylor_green', 'taylor_vortex', 'leapfrog', 'vortices_pass', 'vortices_pass_narrow', 'vortices_pass_noslip', 'karman')
N_SAMPLES, SEED = 64, 1234


@pytest.fixture
def cpu_scene():
	from gaussian_fluids_code_b200 import gsr2d
	from gaussian_fluids_code_b200.init_cond2d import Scene2D
	old = gsr2d.device
	gsr2d.device = torch.device('cpu')
	yield Scene2D
	gsr2d.device = old


@pytest.mark.parametrize('name', SCENES)
def test_tables_and_fields(cpu_scene, name):
	g = load_golden('ref2d_scenes.npz')
	sc = cpu_scene(name)
	assert sc.scaling_factor == pytest.approx(float(g[f'{name}_scaling_factor']), rel=1e-15)
	for dom in ('initialize_domain', 'advance_domain', 'visualize_domain'):
		np.testing.assert_allclose(np.asarray(getattr(sc, dom), np.float64), g[f'{name}_{dom}'], rtol=0., atol=0.)
	x = torch.tensor(g[f'{name}_x'])
	for got, key, tol in ((sc.velocity(x), 'val', 2e-6), (sc.gradient(x), 'grad', 2e-6),
						  (sc.target_velocity(x * sc.scaling_factor), 'target_val', 2e-6), (sc.target_gradient(x * sc.scaling_factor), 'target_grad', 2e-6)):
		ref = g[f'{name}_{key}']
		assert got.shape == ref.shape
		assert np.abs(got.numpy() - ref).max() <= tol * max(np.abs(ref).max(), 1e-30), (name, key)


@pytest.mark.parametrize('name', SCENES)
def test_boundary_samplers_reproduce_the_reference_draws(cpu_scene, name):
	g = load_golden('ref2d_scenes.npz')
	sc = cpu_scene(name)
	b1, b2 = sc.boundary_samplers
	assert (b1 is None) == (f'{name}_sampler1_0' not in g)
	for k, sampler in ((1, b1), (2, b2)):
		if sampler is None:
			continue
		torch.manual_seed(SEED)
		out = sampler(N_SAMPLES)
		for j, t in enumerate(out):
			ref = g[f'{name}_sampler{k}_{j}']
			assert t.shape == ref.shape, (name, k, j)
			np.testing.assert_allclose(t.numpy(), ref, rtol=0., atol=2e-6 * max(1., np.abs(ref).max()), err_msg=f'{name} sampler {k} output {j}')


def test_karman_inlet_moves_with_the_flow(cpu_scene):
	g = load_golden('ref2d_scenes.npz')
	sc = cpu_scene('karman')
	for step in range(3):
		sc.extra_advector(.5)
		np.testing.assert_allclose(np.asarray(sc.advance_domain, np.float64), g[f'karman_advance_domain_after{step + 1}'], rtol=0., atol=1e-15)
	torch.manual_seed(SEED)
	for j, t in enumerate(sc.boundary_sampler_2(N_SAMPLES)):
		ref = g[f'karman_sampler2_moved_{j}']
		np.testing.assert_allclose(t.numpy(), ref, rtol=0., atol=2e-6 * max(1., np.abs(ref).max()))
	sc.extra_loader(start_frame=1000, dt=.01)	# 2D/init_cond.py:285-289: never past the visualised window
	assert sc.advance_domain[0] == sc.visualize_domain[0]
	sc.extra_loader(start_frame=10, dt=.01)
	assert sc.advance_domain[0] == pytest.approx(sc.initialize_domain[0] + .1 * .5)


def test_particle_scene_field_and_jacobian(cpu_scene, tmp_path):
	"""vortices_pass_particles: the asset is not shipped; a small particle list pins the closed-form Jacobian against autograd of the
	reference's single-point formula (2D/init_cond.py:226-234)"""
	path = tmp_path / 'particles.obj'
	rng = np.random.default_rng(3)
	P, W = rng.uniform(-2., 2., (20, 2)), rng.normal(size=20)
	path.write_text(''.join(f'v {p[0]:.7f} 0.0 {p[1]:.7f} {w:.7f}\n' for p, w in zip(P, W)))
	sc = cpu_scene('vortices_pass_particles', particles_obj=str(path))
	x = torch.tensor(rng.uniform(-2., 2., (16, 2)), dtype=torch.float32)
	pos, strength = torch.tensor(P, dtype=torch.float32), torch.tensor(W, dtype=torch.float32)

	def single(xi):
		d = pos - xi[None, :]
		s = (strength[:, None] * d / ((d ** 2).sum(dim=-1)[:, None] + .1)).sum(dim=0)
		return torch.stack([-s[1], s[0]])
	ref_val = torch.stack([single(xi) for xi in x])
	ref_jac = torch.stack([torch.autograd.functional.jacobian(single, xi) for xi in x])
	np.testing.assert_allclose(sc.velocity(x).numpy(), ref_val.numpy(), rtol=0., atol=2e-6 * float(ref_val.abs().max()))
	np.testing.assert_allclose(sc.gradient(x).numpy(), ref_jac.numpy(), rtol=0., atol=2e-6 * float(ref_jac.abs().max()))
	b1, b2 = sc.boundary_samplers
	assert b1 is None and b2(10)[0].shape == (20, 2)
	with pytest.raises(FileNotFoundError):
		cpu_scene('vortices_pass_particles')
