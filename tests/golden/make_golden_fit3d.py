"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref3d_fit.npz by executing the REFERENCE'S OWN initial fit
(/root/reference/3D/initialize.py: fit_velocity_with_gradient — value + gradient L1 losses through get_losses, autograd
regularisers, 4 x Adam, 4 x ReduceLROnPlateau with the class's default learning rates, grid rebuild) on the reference's
GaussianSplatting3DFast through tests/golden/ti_shim.py (float32) for a few iterations with recorded batches and targets.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fit3d.py
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
from ti_shim import GArr  # noqa: E402

REF = '/root/reference/3D'
EPOCHS = (1, 3)
Q = 40


def load():
	ti_shim.install()
	ti_shim.set_dtype(np.float32)
	for name in ('GSR', 'init_cond', 'mesh_sampler'):
		sys.modules.pop(name, None)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp()]
	try:
		spec = importlib.util.spec_from_file_location('ref_initialize3d', os.path.join(REF, 'initialize.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv

	def view(t):
		if isinstance(t, torch.Tensor):
			a = t.detach().numpy().view(GArr)
			a.grad = t.grad.numpy() if getattr(t, 'grad', None) is not None else None
			return a
		return t
	cls = mod.GaussianSplatting3DFast
	for name in ('reinitialize_grid_ti', 'get_losses_ti', 'advection_rk4_ti', 'get_all_neighbors_ti'):
		def adapt(orig):
			return lambda self, *a: orig(self, *[view(t) for t in a])
		setattr(cls, name, adapt(getattr(cls, name)))
	return mod


def target(x):
	"""a smooth analytic field and its Jacobian (the scene's ring field would be 500 shim-evaluated particles per point)"""
	s, c = torch.sin(2. * x), torch.cos(2. * x)
	val = torch.stack([s[:, 1] * c[:, 2], s[:, 2] * c[:, 0], s[:, 0] * c[:, 1]], dim=1) * .3
	jac = torch.zeros((x.shape[0], 3, 3))
	jac[:, 0, 1], jac[:, 0, 2] = 2. * c[:, 1] * c[:, 2] * .3, -2. * s[:, 1] * s[:, 2] * .3
	jac[:, 1, 2], jac[:, 1, 0] = 2. * c[:, 2] * c[:, 0] * .3, -2. * s[:, 2] * s[:, 0] * .3
	jac[:, 2, 0], jac[:, 2, 1] = 2. * c[:, 0] * c[:, 1] * .3, -2. * s[:, 0] * s[:, 1] * .3
	return val, jac


if __name__ == '__main__':
	mod = load()
	rng = np.random.default_rng(63)
	n = 3
	P = (np.stack(np.meshgrid(*[np.linspace(.12, .88, n)] * 3, indexing='ij'), -1).reshape(-1, 3) + rng.uniform(-.06, .06, (n ** 3, 3))).astype(np.float32)
	N = P.shape[0]
	probe = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)
	S = probe.scalings.detach().numpy() + rng.uniform(-.15, .15, (N, 3)).astype(np.float32)
	R = rng.normal(size=(N, 4)).astype(np.float32)
	V = rng.normal(scale=.1, size=(N, 3)).astype(np.float32)
	E = max(EPOCHS)
	samples = rng.uniform(0., 1., (E, Q, 3)).astype(np.float32)
	tv, tj = zip(*[target(torch.tensor(x)) for x in samples])
	out = dict(positions=P, scalings=S, rotations=R, values=V, samples=samples, ref_val=np.stack([t.numpy() for t in tv]), ref_grad=np.stack([t.numpy() for t in tj]),
			   lrs=np.array([probe.positions_lr, probe.scalings_lr, probe.rotations_lr, probe.values_lr]))
	for epochs in EPOCHS:
		gv = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(S)); gv.rotations.copy_(torch.tensor(R)); gv.values.copy_(torch.tensor(V))
		gv.reinitialize_grid()
		gv.zero_grad()
		it = {'k': 0}

		def data_gen(batch):
			x = torch.tensor(samples[it['k']]); it['k'] += 1
			return x
		rec = {}
		ti_shim.record_steps(gv, rec)	# total .grad, metric and lrs at every step()
		mod.fit_velocity_with_gradient(gv, lambda x: target(x)[0], lambda x: target(x)[1], data_gen, batch_size=Q, max_epoch=epochs, verbose=0)
		if epochs == max(EPOCHS):
			assert len(rec['grads']) == epochs
			ti_shim.store_steps(out, rec)
		assert it['k'] == epochs
		for name in ('positions', 'scalings', 'rotations', 'values'):
			out[f'after{epochs}_{name}'] = getattr(gv, name).detach().numpy().copy()
		out[f'after{epochs}_grid_scale'] = np.float64(gv.grid_scale)
		print('epochs', epochs, {nm: float(np.abs(out[f'after{epochs}_{nm}'] - out[nm]).max()) for nm in ('positions', 'scalings', 'rotations', 'values')}, flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref3d_fit.npz'), **out)
