"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref3d_project.npz by executing the REFERENCE'S OWN per-timestep optimisation
(/root/reference/3D/advance.py: AdvectedCovectorField.vorticity and project() — PCGrad projection, regularisers through
autograd, 4 x Adam, 4 x ReduceLROnPlateau, grid rebuild — on the reference's GaussianSplatting3DFast, whose Taichi kernels run as
plain Python through tests/golden/ti_shim.py, float32) for a few iterations on a tiny field with recorded sample batches.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_project3d.py
Nothing here is copied from the reference: the script imports it.
"""
import importlib.util
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ti_shim  # noqa: E402
from ti_shim import GArr  # noqa: E402

REF = '/root/reference/3D'
EPOCHS = (1, 3)
DT, BOUNDARY_LAMBDA, QB = .05, 10., 48


def load():
	ti_shim.install()
	ti_shim.set_dtype(np.float32)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp()]
	try:
		spec = importlib.util.spec_from_file_location('ref_advance3d', os.path.join(REF, 'advance.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	mod.plt = mock.MagicMock()	# project() plots its loss curves when frame_id is given (and needs frame_id for its loss lists)
	mod.plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())

	# argument adapter: the kernel bodies index and assign into their array arguments element by element; hand them numpy views
	# of the torch tensors' memory (with .grad views for the parameters, which the kernels reach through `positions.grad`)
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


def make_fields(mod, init):
	def field(P, S, R, V):
		gv = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(S)); gv.rotations.copy_(torch.tensor(R)); gv.values.copy_(torch.tensor(V))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv
	cur = field(init['cur_positions'], init['scalings'], init['rotations'], init['values'])
	new = field(init['new_positions'], init['scalings'], init['rotations'], init['values'])
	return cur, new


if __name__ == '__main__':
	mod = load()
	rng = np.random.default_rng(31)
	n = 3
	P = np.stack(np.meshgrid(*[np.linspace(.12, .88, n)] * 3, indexing='ij'), -1).reshape(-1, 3) + rng.uniform(-.06, .06, (n ** 3, 3))
	N = P.shape[0]
	probe = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P.astype(np.float32), dim=3)	# the constructor's own initial scaling
	S = probe.scalings.detach().numpy() + rng.uniform(-.15, .15, (N, 3)).astype(np.float32)
	R = rng.normal(size=(N, 4)).astype(np.float32)
	V = rng.normal(scale=.3, size=(N, 3)).astype(np.float32)
	init = dict(cur_positions=P.astype(np.float32), new_positions=(P + rng.normal(scale=.01, size=P.shape)).astype(np.float32), scalings=S, rotations=R, values=V)
	samples = rng.uniform(0., 1., (max(EPOCHS), N, 3)).astype(np.float32)
	torch.manual_seed(5)
	bnd = [mod.sample_on_box(QB, 0., 1., 0., 1., 0., 1.) for _ in range(max(EPOCHS))]
	out = dict(init, samples=samples, boundary_data=np.stack([b[0].numpy() for b in bnd]), boundary_normal=np.stack([b[1].numpy() for b in bnd]),
			   dt=np.float64(DT), boundary_lambda=np.float64(BOUNDARY_LAMBDA), min_grid_scale=np.float64(probe.min_grid_scale), tau=np.float64(probe.clamp_threshold))
	for epochs in EPOCHS:
		cur, new = make_fields(mod, init)
		ref = mod.AdvectedCovectorField(cur, cur, DT, 0., 1., 0., 1., 0., 1.)
		it = {'k': 0, 'b': 0}

		def data_gen(batch, gv):
			x = torch.tensor(samples[it['k']]); it['k'] += 1
			return x

		def bnd_gen(batch):
			d, nrm = bnd[it['b']]; it['b'] += 1
			return d.clone(), nrm.clone()
		# recorders: the raw vorticity / divergence gradient sets as get_losses_ti leaves them (before project()'s PCGrad projection),
		# and, at every step(), the total .grad (projected sets + autograd regularisers + boundary pass) with the scheduler metric
		rec = {'sets': [], 'grads': [], 'metric': [], 'lr': []}
		NAMES = ('positions', 'scalings', 'rotations', 'values')
		orig_get_losses, orig_step = new.get_losses, new.step

		def get_losses(x, *a, **kw):
			res = orig_get_losses(x, *a, **kw)
			if kw.get('vor_positions_grad') is not None:
				rec['sets'].append({f'{t}_{nm}': kw[f'{t}_{nm}_grad'].detach().numpy().copy() for t in ('vor', 'div') for nm in NAMES})
			return res

		def step(metrics):
			rec['grads'].append({nm: getattr(new, nm).grad.detach().numpy().copy() for nm in NAMES})
			rec['metric'].append(float(metrics))
			rec['lr'].append([o.param_groups[0]['lr'] for o in new.optimizers])
			return orig_step(metrics)
		new.get_losses, new.step = get_losses, step
		mod.project(new, ref, 0., 1., 0., 1., 0., 1., data_gen, lambda gv: None, boundary_generator=bnd_gen, boundary_lambda=BOUNDARY_LAMBDA,
					batch_size=QB, max_epoch=epochs, patience=500, verbose=0, frame_id=0)
		assert it['k'] == epochs and it['b'] == epochs and len(rec['sets']) == epochs and len(rec['grads']) == epochs
		if epochs == max(EPOCHS):
			for k in range(epochs):
				for key, v in rec['sets'][k].items():
					out[f'it{k + 1}_{key}_grad'] = v	# raw set, before PCGrad
				for nm in NAMES:
					out[f'it{k + 1}_total_{nm}_grad'] = rec['grads'][k][nm]	# .grad at step(): what Adam consumes
				out[f'it{k + 1}_metric'] = np.float64(rec['metric'][k])
				out[f'it{k + 1}_lr_used'] = np.array(rec['lr'][k])
		for name in ('positions', 'scalings', 'rotations', 'values'):
			out[f'after{epochs}_{name}'] = getattr(new, name).detach().numpy().copy()
		out[f'after{epochs}_grid_scale'] = np.float64(new.grid_scale)
		out[f'after{epochs}_lr'] = np.array([o.param_groups[0]['lr'] for o in (new.positions_optimizer, new.scalings_optimizer, new.rotations_optimizer, new.values_optimizer)])
		print('epochs', epochs, 'max |d positions|', float(np.abs(out[f'after{epochs}_positions'] - init['new_positions']).max()),
			  'max |d values|', float(np.abs(out[f'after{epochs}_values'] - V).max()), flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref3d_project.npz'), **out)
