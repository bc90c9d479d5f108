"""
The Karman initial state (2D/initialize.py:162-185 with its own copy of project, :44-160): fit with the scene's learning rates, then
a projection against the field's own vorticity (dt = 0) with divergence weight 10, no position anchor, lrs 1e-4 / 1e-5 / 1.2e-5 / 1e-4,
obstacle + channel-wall + inlet/outlet boundary samples at weight 10.  The fused device path against the statement-by-statement
formulation on the drop-in class (the parity partner of the other 2D project tests, itself pinned by the reference's goldens).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
NAMES = ('positions', 'scalings', 'rotations', 'values')


def run(fused, fit_epochs=3, project_epochs=3):
	from gaussian_fluids_code_b200 import advance2d, gsr2d, init_cond2d
	gsr2d.device = torch.device('cuda', 0)
	scene = init_cond2d.Scene2D('karman')
	torch.manual_seed(7)
	gv = advance2d.simulation_initialize(scene, max_epoch=fit_epochs, verbose=0, fused=fused, project_epochs=project_epochs)
	return scene, gv


def test_karman_initial_state_constants_and_paths_agree():
	from gaussian_fluids_code_b200 import advance2d
	scene, a = run(True)
	_, b = run(False)
	assert a.N == b.N == 400 * 60
	# the second project copy's constants (2D/initialize.py:55, :125-126)
	assert advance2d.INIT_PROJECT_WEIGHTS == dict(vor=1., div=10., aniso=10., vol=10., delta_pos=0.)
	assert [a.positions_lr, a.scalings_lr, a.rotations_lr, a.values_lr] == pytest.approx([1e-4, 1e-5, 1.201956e-5, 1e-4], rel=1e-12)
	for nm in NAMES:
		x, y = getattr(a, nm).detach().cpu().numpy(), getattr(b, nm).detach().cpu().numpy()
		assert np.isfinite(x).all()
		# Adam's first steps move every entry by ~ lr * sign(g): an entry whose gradient is rounding noise may flip (a few lr), the rest agrees
		d = np.abs(x - y) / np.abs(y).max()
		assert (d > 1e-5).mean() < .01 and d.max() < 2e-2, (nm, (d > 1e-5).mean(), d.max())
	assert a.grid_scale == pytest.approx(b.grid_scale, rel=1e-5)


def test_karman_projection_reduces_divergence_and_boundary_flux():
	"""40 projection iterations on the fitted inflow: the divergence and the flux through the obstacle's samples must go down"""
	from gaussian_fluids_code_b200 import init_cond2d
	scene, before = run(True, fit_epochs=30, project_epochs=0)
	_, after = run(True, fit_epochs=30, project_epochs=40)
	pts = scene.test_generator()

	def div2(gv):
		g = gv.gradient(pts)
		return float(((g[:, 0, 0] + g[:, 1, 1]) ** 2).mean())
	assert div2(after) < div2(before)
	torch.manual_seed(3)
	data, value = scene.boundary_samplers[0](4096)	# u = 0 on the cylinder
	err = lambda gv: float((gv(data) - value).abs().mean())
	assert err(after) < err(before)
