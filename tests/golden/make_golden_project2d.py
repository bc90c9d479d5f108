"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref2d_project.npz by executing the REFERENCE'S OWN 2D per-timestep optimisation
(/root/reference/2D/advance.py: AdvectedCovectorField.vorticity and project() — value samples on obstacles (boundary_generator_1),
normal samples (boundary_generator_2), PCGrad projection, autograd regularisers incl. the position-drift term, 4 x Adam,
4 x ReduceLROnPlateau, grid rebuild — on the reference's GaussianSplattingFast, whose Taichi kernels run as plain Python through
tests/golden/ti_shim.py, float32) for a few iterations on a tiny field with recorded sample batches.  Scene: taylor_vortex
(domain [-5, 5]^2, scale factor 1 — the reference's vorticity() reads the domain of the scene given on the command line).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_project2d.py
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
DT, BOUNDARY_LAMBDA, QB = .05, 1., 24
DOM = (-5., 5., -5., 5.)


def load():
	ti_shim.install()
	ti_shim.set_dtype(np.float32)
	for name in ('GSR', 'init_cond'):
		sys.modules.pop(name, None)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp(), '--init_cond', 'taylor_vortex']
	try:
		spec = importlib.util.spec_from_file_location('ref_advance2d', os.path.join(REF, 'advance.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv

	def view(t):	# numpy views of the torch tensors' memory for the kernel bodies (parameters carry a .grad view)
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
	assert mod.scaling_factor == 1.
	return mod


if __name__ == '__main__':
	mod = load()
	rng = np.random.default_rng(47)
	n = 5
	P = np.stack(np.meshgrid(*[np.linspace(-4., 4., n)] * 2, indexing='ij'), -1).reshape(-1, 2) + rng.uniform(-.4, .4, (n * n, 2))
	N = P.shape[0]
	probe = mod.GaussianSplattingFast(*DOM, P.astype(np.float32), dim=2)
	S = probe.scalings.detach().numpy() + rng.uniform(-.15, .15, (N, 2)).astype(np.float32)
	R = rng.uniform(-np.pi, np.pi, probe.rotations.shape).astype(np.float32)
	V = rng.normal(scale=.5, size=(N, 2)).astype(np.float32)
	init = dict(cur_positions=P.astype(np.float32), new_positions=(P + rng.normal(scale=.05, size=P.shape)).astype(np.float32), scalings=S, rotations=R, values=V)
	E = max(EPOCHS)
	samples = rng.uniform(-5., 5., (E, N, 2)).astype(np.float32)
	th = rng.uniform(0., 2. * np.pi, (E, QB))
	b1_data = np.stack([1.5 * np.cos(th), 1.5 * np.sin(th)], -1).astype(np.float32)	# value samples on a circle: u = (0.1, 0) there
	b1_val = np.broadcast_to(np.array([.1, 0.], np.float32), b1_data.shape).copy()
	t = rng.uniform(-5., 5., (E, QB)).astype(np.float32)
	b2_data = np.stack([t, np.full_like(t, -5.)], -1)	# normal samples on the bottom edge: u.n = 0.05 there
	b2_nrm = np.broadcast_to(np.array([0., -1.], np.float32), b2_data.shape).copy()
	b2_ref = np.full((E, QB), .05, np.float32)
	out = dict(init, samples=samples, b1_data=b1_data, b1_val=b1_val, b2_data=b2_data, b2_normal=b2_nrm, b2_ref=b2_ref, domain=np.array(DOM),
			   dt=np.float64(DT), boundary_lambda=np.float64(BOUNDARY_LAMBDA), min_grid_scale=np.float64(probe.min_grid_scale), tau=np.float64(probe.clamp_threshold))

	def field(Pos):
		gv = mod.GaussianSplattingFast(*DOM, Pos, dim=2)
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(S)); gv.rotations.copy_(torch.tensor(R)); gv.values.copy_(torch.tensor(V))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv
	for epochs in EPOCHS:
		cur, new = field(init['cur_positions']), field(init['new_positions'])
		ref = mod.AdvectedCovectorField(cur, cur, DT)
		it = {'k': 0, 'b1': 0, 'b2': 0}

		def data_gen(batch, gv):
			x = torch.tensor(samples[it['k']]); it['k'] += 1
			return x

		def gen1(batch):
			k = it['b1']; it['b1'] += 1
			return torch.tensor(b1_data[k]), torch.tensor(b1_val[k])

		def gen2(batch):
			k = it['b2']; it['b2'] += 1
			return torch.tensor(b2_data[k]), torch.tensor(b2_nrm[k]), torch.tensor(b2_ref[k])
		rec = {}
		ti_shim.record_steps(new, rec, 'get_grad_losses')	# raw vor / div sets, total .grad, metric and lrs at every step()
		mod.project(new, ref, data_gen, lambda gv: None, boundary_generator_1=gen1, boundary_generator_2=gen2, boundary_lambda=BOUNDARY_LAMBDA,
					batch_size=QB, max_epoch=epochs, patience=500, verbose=0)
		assert it['k'] == epochs and it['b1'] == epochs and it['b2'] == epochs and len(rec['grads']) == epochs and len(rec['sets']) == epochs
		if epochs == max(EPOCHS):
			ti_shim.store_steps(out, rec)
		for name in ('positions', 'scalings', 'rotations', 'values'):
			out[f'after{epochs}_{name}'] = getattr(new, name).detach().numpy().copy()
		out[f'after{epochs}_grid_scale'] = np.float64(new.grid_scale)
		print('epochs', epochs, {nm: float(np.abs(out[f'after{epochs}_{nm}'] - (init['new_positions'] if nm == 'positions' else init[nm])).max()) for nm in ('positions', 'scalings', 'rotations', 'values')}, flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref2d_project.npz'), **out)
