"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref3d_density.npz by executing the REFERENCE'S OWN passive density advection
(/root/reference/3D/advance_density.py: ti_set_ring, ti_get_coord, ti_get_interp_val; /root/reference/3D/GSR.py:
get_grid_points, advection_rk4_ti) as plain Python through tests/golden/ti_shim.py on a small lattice — the composition of
advected_density (:52-58): lattice -> RK4 back-trace by -dt -> clamp to the domain -> trilinear resampling — in float32 (the
reference's arithmetic) and float64.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_density.py
Nothing here is copied from the reference: the script imports it.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ti_shim  # noqa: E402
from ti_shim import IdxArr  # noqa: E402
import make_golden as mg  # noqa: E402  (scene(), new_fast(): the Gaussian field and the reference's fast class around it)

REF = '/root/reference/3D'
RES = (11, 9, 10)
DOMAIN = (0., 1., 0., 1., 0., 1.)
DT = .35


def load():
	ti_shim.install()
	for name in ('GSR', 'init_cond', 'mesh_sampler', 'advance_density'):
		sys.modules.pop(name, None)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp(), '--init_cond', 'ring_collide']
	try:
		spec = importlib.util.spec_from_file_location('ref_advance_density', os.path.join(REF, 'advance_density.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	return mod


if __name__ == '__main__':
	mod = load()
	out = {'res': np.array(RES), 'domain': np.array(DOMAIN), 'dt': np.float64(DT)}
	rng = np.random.default_rng(515)
	sc = mg.scene(3, rng, (4, 4, 3), 4, 5e-3)
	sc['values'] = mg.f32r(sc['values'] * 1.5)	# a field strong enough to move the density by about a voxel
	out.update({'in_' + k: v for k, v in sc.items()})
	ring = dict(center=[.45, .5, .55], normal=[1., 0., 0.], radius=.25, thickness=.12)
	out.update(ring_center=np.array(ring['center']), ring_normal=np.array(ring['normal']), ring_radius=np.float64(ring['radius']), ring_thickness=np.float64(ring['thickness']))
	x_min, x_max, y_min, y_max, z_min, z_max = DOMAIN
	for dt, tag in ((np.float32, 'f32'), (np.float64, 'f64')):
		ti_shim.set_dtype(dt)
		torch.set_default_dtype(torch.float32 if dt == np.float32 else torch.float64)
		o, P = mg.new_fast(mod, 3, sc, dt)
		# ti_set_ring (:13-22)
		density = IdxArr(np.zeros(RES), dtype=dt)
		vec3 = sys.modules['taichi.math'].vec3
		mod.ti_set_ring(density, vec3(ring['center']), vec3(ring['normal']), ring['radius'], ring['thickness'], *DOMAIN)
		# advected_density (:52-58), step by step with the module's own pieces
		x = mod.get_grid_points(x_min, x_max, y_min, y_max, z_min, z_max, *RES).numpy().astype(dt)
		Q = x.shape[0]
		Z = lambda *s: np.zeros(s, dt)
		bk, deform, gval, ggrad = Z(Q, 3), Z(Q, 3, 3), Z(Q, 3), Z(Q, 3, 3)
		o.advection_rk4_ti(P[0], P[1], P[2], P[3], o.grid_scale, x, -DT, bk, deform, gval, ggrad)
		bk = np.clip(bk, np.array([x_min, y_min, z_min], dt), np.array([x_max, y_max, z_max], dt)).reshape(*RES, 3)
		smooth = np.asarray(np.random.default_rng(9).uniform(size=RES), np.float32).astype(dt)	# a second, smooth-valued field
		for name, field in (('ring', np.asarray(density)), ('smooth', smooth)):
			nxt = IdxArr(np.zeros(RES), dtype=dt)
			mod.ti_get_interp_val(np.asarray(field, dt), bk, nxt, *DOMAIN)
			out[f'{name}_density_{tag}'] = np.asarray(field).copy()
			out[f'{name}_next_{tag}'] = np.asarray(nxt).copy()
		out[f'backtraced_{tag}'] = bk.copy()
		print(tag, 'ring voxels', int(np.asarray(density).sum()), 'moved', float(np.abs(out[f'smooth_next_{tag}'] - smooth).max()), flush=True)
	torch.set_default_dtype(torch.float32)
	np.savez_compressed(os.path.join(HERE, 'ref3d_density.npz'), **out)
