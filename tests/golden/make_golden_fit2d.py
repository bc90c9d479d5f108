"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref2d_fit.npz by executing the REFERENCE'S OWN 2D initial fit
(/root/reference/2D/initialize.py: fit_velocity_with_gradient — value loss through get_losses, gradient loss through
get_grad_losses, autograd regularisers, 4 x Adam, 4 x ReduceLROnPlateau at the learning rates SimulationInitialize sets) on the
reference's GaussianSplattingFast through tests/golden/ti_shim.py (float32), a few iterations with recorded batches and targets.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fit2d.py
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

REF = '/root/reference/2D'
EPOCHS = (1, 3)
Q = 40
DOM = (-5., 5., -5., 5.)
LRS = dict(positions_lr=1.6e-3, scalings_lr=5e-2, rotations_lr=5e-2, values_lr=5e-3)


def load():
	ti_shim.install()
	ti_shim.set_dtype(np.float32)
	for name in ('GSR', 'init_cond', 'advance'):
		sys.modules.pop(name, None)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp(), '--init_cond', 'taylor_vortex']
	try:
		spec = importlib.util.spec_from_file_location('ref_initialize2d', os.path.join(REF, 'initialize.py'))
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
	cls = mod.GaussianSplattingFast
	for name in ('reinitialize_grid_ti', 'get_losses_ti', 'get_grad_losses_ti', 'advection_rk4_ti', 'get_all_neighbors_ti'):
		def adapt(orig):
			return lambda self, *a: orig(self, *[view(t) for t in a])
		setattr(cls, name, adapt(getattr(cls, name)))
	return mod


def target(x):
	s, c = torch.sin(.7 * x), torch.cos(.7 * x)
	val = torch.stack([s[:, 0] * c[:, 1], -c[:, 0] * s[:, 1]], dim=1)
	jac = torch.stack([.7 * c[:, 0] * c[:, 1], -.7 * s[:, 0] * s[:, 1], .7 * s[:, 0] * s[:, 1], -.7 * c[:, 0] * c[:, 1]], dim=1).reshape(-1, 2, 2)
	return val, jac


if __name__ == '__main__':
	mod = load()
	rng = np.random.default_rng(71)
	n = 5
	P = (np.stack(np.meshgrid(*[np.linspace(-4., 4., n)] * 2, indexing='ij'), -1).reshape(-1, 2) + rng.uniform(-.4, .4, (n * n, 2))).astype(np.float32)
	N = P.shape[0]
	probe = mod.GaussianSplattingFast(*DOM, P, dim=2)
	S = probe.scalings.detach().numpy() + rng.uniform(-.15, .15, (N, 2)).astype(np.float32)
	R = rng.uniform(-np.pi, np.pi, probe.rotations.shape).astype(np.float32)
	V = rng.normal(scale=.3, size=(N, 2)).astype(np.float32)
	E = max(EPOCHS)
	samples = rng.uniform(-5., 5., (E, Q, 2)).astype(np.float32)
	tv, tj = zip(*[target(torch.tensor(x)) for x in samples])
	out = dict(positions=P, scalings=S, rotations=R, values=V, samples=samples, ref_val=np.stack([t.numpy() for t in tv]), ref_grad=np.stack([t.numpy() for t in tj]),
			   lrs=np.array([LRS['positions_lr'], LRS['scalings_lr'], LRS['rotations_lr'], LRS['values_lr']]), domain=np.array(DOM),
			   min_grid_scale=np.float64(probe.min_grid_scale), tau=np.float64(probe.clamp_threshold))
	for epochs in EPOCHS:
		gv = mod.GaussianSplattingFast(*DOM, P, dim=2)
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(S)); gv.rotations.copy_(torch.tensor(R)); gv.values.copy_(torch.tensor(V))
		gv.set_lr(**LRS)
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
		for name in ('positions', 'scalings', 'rotations', 'values'):
			out[f'after{epochs}_{name}'] = getattr(gv, name).detach().numpy().copy()
		out[f'after{epochs}_grid_scale'] = np.float64(gv.grid_scale)
		print('epochs', epochs, {nm: float(np.abs(out[f'after{epochs}_{nm}'] - out[nm]).max()) for nm in ('positions', 'scalings', 'rotations', 'values')}, flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref2d_fit.npz'), **out)
