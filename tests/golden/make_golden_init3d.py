"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref3d_init_fields.npz by executing the REFERENCE'S OWN analytic initial fields
(/root/reference/3D/init_cond.py: vortex_particle, vortex_particle_gradient and the four scene functions built on them) as
plain Python through tests/golden/ti_shim.py — once in float32 (the reference's arithmetic) and once in float64.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_init3d.py
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

REF = '/root/reference/3D'
SCENES = ('leapfrog', 'single_vortex_ring', 'ring_collide', 'ring_with_obstacle')


def load_ref():
	ti_shim.install()
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp()]
	try:
		spec = importlib.util.spec_from_file_location('ref_init_cond3d', os.path.join(REF, 'init_cond.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	return mod


def points(rng):
	"""random points of the unit cube, plus points close to (not on) the core of the first leapfrog ring"""
	x = rng.uniform(0., 1., (40, 3))
	th = rng.uniform(0., 2. * np.pi, 8)
	core = np.stack([np.full(8, .75), .5 + np.cos(th) / 6., .5 + np.sin(th) / 6.], 1) + rng.normal(scale=.01, size=(8, 3))
	return np.concatenate([x, core]).astype(np.float32)


if __name__ == '__main__':
	mod = load_ref()
	x = points(np.random.default_rng(2024))
	out = {'x': x}
	for dt, tag in ((np.float32, 'f32'), (np.float64, 'f64')):
		ti_shim.set_dtype(dt)
		torch.set_default_dtype(torch.float32 if dt == np.float32 else torch.float64)
		xt = torch.tensor(x.astype(dt))
		for name in SCENES:
			f = getattr(mod, name)
			out[f'{name}_val_{tag}'] = f(xt).numpy().astype(dt)
			out[f'{name}_grad_{tag}'] = f.gradient(xt).numpy().astype(dt)
			print(name, tag, float(np.abs(out[f'{name}_val_{tag}']).max()), float(np.abs(out[f'{name}_grad_{tag}']).max()), flush=True)
	torch.set_default_dtype(torch.float32)
	np.savez_compressed(os.path.join(HERE, 'ref3d_init_fields.npz'), **out)
