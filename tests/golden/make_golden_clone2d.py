"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref2d_clone.npz by executing the REFERENCE'S OWN reseeding step
(/root/reference/2D/advance.py:58-158, clone_velocity_field: split of the Gaussians with axis ratio >= 1.5 into two samples of their
own distribution, stop_gradient bookkeeping with the neighbours of the new Gaussians unfrozen, refit of the trainable ones with the
value + gradient losses) on the reference's GaussianSplattingFast through tests/golden/ti_shim.py, float32.

Recorded: the standard-normal draws MultivariateNormal consumed (so that the map draws -> child positions can be replayed), the
field right after the split, the final stop_gradient mask, the total .grad / metric / lrs at every step() and the parameters
after 1 and 3 refit iterations.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_clone2d.py
Nothing here is copied from the reference: the script imports it.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ti_shim  # noqa: E402
from make_golden_project2d import DOM, load  # noqa: E402

EPOCHS = (1, 3)
Q = 40

if __name__ == '__main__':
	mod = load()
	rng = np.random.default_rng(83)
	n = 8
	P = (np.stack(np.meshgrid(*[np.linspace(-4., 4., n)] * 2, indexing='ij'), -1).reshape(-1, 2) + rng.uniform(-.3, .3, (n * n, 2))).astype(np.float32)
	N = P.shape[0]
	probe = mod.GaussianSplattingFast(*DOM, P, dim=2)
	S = probe.scalings.detach().numpy() + rng.uniform(-.1, .1, (N, 2)).astype(np.float32)
	corner = np.flatnonzero((P[:, 0] < -1.7) & (P[:, 1] < -1.7))	# the over-stretched ones sit in one corner, so that far Gaussians stay frozen
	stretched = rng.choice(corner, 5, replace=False)
	S[stretched, rng.integers(0, 2, 5)] -= rng.uniform(.45, .7, 5).astype(np.float32)	# axis ratio 1.6 .. 2.2 for five of them
	R = rng.uniform(-np.pi, np.pi, probe.rotations.shape).astype(np.float32)
	V = rng.normal(scale=.3, size=(N, 2)).astype(np.float32)
	E = max(EPOCHS)

	# res starts as a copy of src, so away from the split Gaussians val - ref is rounding noise and the L1 losses' sign(val - ref)
	# is a coin flip in ANY implementation (the reference's included): the recorded batches lie inside the support of a split parent
	# (gaussian >= 0.1), where the parent's missing term makes the difference a real signal
	def parent_weight(x):
		w = np.zeros(x.shape[0])
		for j in stretched:
			c, s_ = np.cos(R.reshape(-1)[j]), np.sin(R.reshape(-1)[j])
			rot = np.array([[c, -s_], [s_, c]], np.float64)
			A = rot @ np.diag(np.exp(2. * S[j].astype(np.float64))) @ rot.T
			d = x - P[j]
			w = np.maximum(w, np.exp(-.5 * np.einsum('qi,ij,qj->q', d, A, d)))
		return w
	cand = rng.uniform(-5., 0., (400 * E * Q, 2))
	cand = cand[parent_weight(cand) >= .1]
	assert cand.shape[0] >= E * Q, cand.shape
	samples = cand[:E * Q].reshape(E, Q, 2).astype(np.float32)
	test_pts = rng.uniform(-5., 5., (Q, 2)).astype(np.float32)
	out = dict(positions=P, scalings=S, rotations=R, values=V, samples=samples, test_points=test_pts, domain=np.array(DOM),
			   min_grid_scale=np.float64(probe.min_grid_scale), tau=np.float64(probe.clamp_threshold))

	def field():
		gv = mod.GaussianSplattingFast(*DOM, P, dim=2)
		with torch.no_grad():
			gv.scalings.copy_(torch.tensor(S)); gv.rotations.copy_(torch.tensor(R)); gv.values.copy_(torch.tensor(V))
		gv.reinitialize_grid()
		gv.zero_grad()
		return gv

	import torch.distributions.multivariate_normal as mvn
	orig_normal = mvn._standard_normal
	for epochs in EPOCHS:
		src, res = field(), field()
		draws = []

		def recording_normal(shape, dtype, device):
			z = orig_normal(shape, dtype, device)
			draws.append(z.detach().numpy().copy())
			return z
		mvn._standard_normal = recording_normal
		torch.manual_seed(17)
		snap, masks, rec = {}, [], {}
		orig_unfreeze = res.unfreeze

		def unfreeze():	# called right after the split, before anything is trained (2D/advance.py:88)
			for nm in ('positions', 'scalings', 'rotations', 'values'):
				snap[nm] = getattr(res, nm).detach().numpy().copy()
			return orig_unfreeze()
		res.unfreeze = unfreeze
		orig_get_losses = res.get_losses

		def get_losses(x, *a, **kw):
			if kw.get('stop_gradient') is not None:
				masks.append(kw['stop_gradient'].detach().numpy().copy())
			return orig_get_losses(x, *a, **kw)
		res.get_losses = get_losses
		ti_shim.record_steps(res, rec)
		it = {'k': 0}

		def data_gen(batch, gv, restrict=None):
			x = torch.tensor(samples[it['k']]); it['k'] += 1
			return x
		mod.clone_velocity_field(res, src, data_gen, lambda gv: torch.tensor(test_pts), batch_size=Q, max_epoch=epochs, patience=500, verbose=0)
		mvn._standard_normal = orig_normal
		assert it['k'] == epochs and len(draws) == 1 and len(rec['grads']) == epochs
		if epochs == max(EPOCHS):
			out['normals'] = draws[0]	# (2, n_split, 2)
			for nm, v in snap.items():
				out[f'split_{nm}'] = v
			out['stop_gradient'] = masks[0]
			ti_shim.store_steps(out, rec)
		for nm in ('positions', 'scalings', 'rotations', 'values'):
			out[f'after{epochs}_{nm}'] = getattr(res, nm).detach().numpy().copy()
		print('epochs', epochs, 'N', N, '->', res.N, 'trainable', int((masks[0] == 0).sum()), 'draws', draws[0].shape, flush=True)
	np.savez_compressed(os.path.join(HERE, 'ref2d_clone.npz'), **out)
