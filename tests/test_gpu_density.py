"""
GPU tests of the fused density advection (3D/advance_density.py): one kernel (lattice coordinates, RK4
back-trace, clamp, trilinear resampling) against the reference's own composition of advection_rk4 + trilinear taps.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_advect_density_matches_reference_composition():
	from gaussian_fluids_code_b200 import advance_density, gsr3d, init_cond3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = torch.device('cuda', 0)
	P, S, R, V, mgs, gen = synthetic_field(12)
	gv = make_fast3d(P, S, R, V * 8., 5e-3, mgs)		# strong field: back-traced points move by ~1 voxel
	dom = (0., 1., 0., 1., 0., 1.)
	adv = advance_density.DensityAdvector(*dom, res=(50, 41, 67))	# not multiples of the 8^3 CTA block
	info = init_cond3d.other_info['ring_collide']
	ring = dict(info['ring1'], radius=.25, thickness=.08)
	d1 = adv.set_ring(ring)
	assert 0 < float(d1.mean()) < .5
	d2 = torch.rand(adv.res, generator=torch.Generator().manual_seed(1)).cuda()
	o1, o2 = adv.advect(gv, .1, d1, d2)
	r1 = advance_density.advected_density_reference(d1, gv, .1, dom)
	r2 = advance_density.advected_density_reference(d2, gv, .1, dom)
	assert float((o2 - d2).abs().max()) > 1e-2			# the field really moved the density
	# the smooth field agrees to rounding; the {0,1} indicator amplifies position rounding by 1 / voxel at its edges
	assert float((o2 - r2).abs().max()) < 2e-4
	assert float((o1 - r1).abs().max()) < 2e-3 and float((o1 - r1).abs().mean()) < 1e-5
	single = adv.advect(gv, .1, d2)
	assert torch.equal(single, o2)
	# zero velocity: the resampling reproduces the field
	with torch.no_grad():
		gv.values.zero_()
	gv.zero_grad()
	still = adv.advect(gv, .1, d2)
	assert float((still - d2).abs().max()) < 1e-5


def test_advect_density_matches_reference_golden():
	"""the fused kernel against the reference's own kernels run through the shim (tests/golden/make_golden_density.py): ti_set_ring,
	and advected_density = advection_rk4_ti by -dt, clamp, ti_get_interp_val — float64 truth, float32 arithmetic here"""
	from helpers import load_golden
	from gaussian_fluids_code_b200 import advance_density, gsr3d
	from gaussian_fluids_code_b200.synth import make_fast3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_density.npz')
	gv = make_fast3d(g['in_positions'], g['in_scalings'], g['in_rotations'], g['in_values'], float(g['in_tau']), float(g['in_min_grid_scale']))
	adv = advance_density.DensityAdvector(*[float(v) for v in g['domain']], res=tuple(int(v) for v in g['res']))
	ring = dict(center=g['ring_center'].tolist(), normal=g['ring_normal'].tolist(), radius=float(g['ring_radius']), thickness=float(g['ring_thickness']))
	d = adv.set_ring(ring).cpu().numpy()
	assert int(g['ring_density_f32'].sum()) > 20
	assert int((d != g['ring_density_f32']).sum()) <= 1	# a voxel exactly on the torus surface may round either way
	ring0 = torch.tensor(g['ring_density_f32'], device='cuda')
	smooth0 = torch.tensor(g['smooth_density_f32'], device='cuda')
	o_ring, o_smooth = adv.advect(gv, float(g['dt']), ring0, smooth0)
	assert float(np.abs(g['smooth_next_f64'] - g['smooth_density_f64']).max()) > .1	# the field moved the density
	assert np.abs(o_smooth.double().cpu().numpy() - g['smooth_next_f64']).max() < 2e-5
	assert np.abs(o_ring.double().cpu().numpy() - g['ring_next_f64']).max() < 2e-5
