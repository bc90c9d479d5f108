"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref2d_scenes.npz by executing the REFERENCE'S OWN 2D scenario module
(/root/reference/2D/init_cond.py) once per scene (it reads the scene from the command line at import): scale factor, domains,
the analytic field and Jacobian at fixed points (original and GSR-space versions), and the boundary samplers with a fixed
torch seed on the CPU (the reference draws with torch.rand, so the same seed reproduces its draws).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_scenes2d.py
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

REF = '/root/reference/2D'
SCENES = ('taylor_green', 'taylor_vortex', 'leapfrog', 'vortices_pass', 'vortices_pass_narrow', 'vortices_pass_noslip', 'karman')
N_SAMPLES, SEED = 64, 1234


def load(scene):
	ti_shim.install()
	for name in ('GSR', 'init_cond'):
		sys.modules.pop(name, None)
	if REF not in sys.path:
		sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp(), '--init_cond', scene]
	try:
		spec = importlib.util.spec_from_file_location('ref_init_cond2d_' + scene, os.path.join(REF, 'init_cond.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	return mod


if __name__ == '__main__':
	out = {}
	for scene in SCENES:
		mod = load(scene)
		out[f'{scene}_scaling_factor'] = np.float64(mod.scaling_factor)
		for dom in ('initialize_domain', 'advance_domain', 'visualize_domain'):
			out[f'{scene}_{dom}'] = np.asarray(getattr(mod, dom)[scene], np.float64)
		x_min, x_max, y_min, y_max = mod.initialize_domain[scene]
		g = torch.Generator().manual_seed(99)
		x = torch.rand((32, 2), generator=g) * torch.tensor([x_max - x_min, y_max - y_min]) + torch.tensor([x_min, y_min])
		out[f'{scene}_x'] = x.numpy()
		f = getattr(mod, scene)
		out[f'{scene}_val'] = f(x, False).numpy()
		out[f'{scene}_grad'] = f(x, True).numpy()
		xt = x * mod.scaling_factor
		out[f'{scene}_target_val'] = mod.target_field(lambda y: f(y, False))(xt).numpy()
		out[f'{scene}_target_grad'] = mod.target_gradient(lambda y: f(y, True))(xt).numpy()
		for k, sampler in enumerate(mod.boundary_sampler[scene]):
			if sampler is None:
				continue
			torch.manual_seed(SEED)
			for j, t in enumerate(sampler(N_SAMPLES)):
				out[f'{scene}_sampler{k + 1}_{j}'] = t.numpy()
		if mod.extra_advector[scene]:	# karman: the inlet moves with the flow until it reaches the visualised window
			for step in range(3):
				mod.extra_advector[scene](.5)
				out[f'{scene}_advance_domain_after{step + 1}'] = np.asarray(mod.advance_domain[scene], np.float64)
			torch.manual_seed(SEED)
			for j, t in enumerate(mod.boundary_sampler[scene][1](N_SAMPLES)):
				out[f'{scene}_sampler2_moved_{j}'] = t.numpy()
		print(scene, float(mod.scaling_factor), [k for k in out if k.startswith(scene + '_sampler')], flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref2d_scenes.npz'), **out)
