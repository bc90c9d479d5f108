"""
On-disk formats of the drop-in classes: checkpoints WRITTEN BY THE REFERENCE'S OWN classes
(tests/golden/ref3d_checkpoint.pt, ref2d_checkpoint.pt: GaussianSplatting3DFast.save / GaussianSplattingFast.save run through the
Taichi shim, tests/golden/make_golden_checkpoint.py) load into the CUDA classes and give the fields the reference's kernels gave;
what the CUDA classes save is the reference's dict key by key; write_vti / write_obj round trips.
"""
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu


def same_dict(ours, ref):
	assert list(ours.keys()) == list(ref.keys())
	for k, v in ref.items():
		if isinstance(v, torch.Tensor):
			assert isinstance(ours[k], torch.Tensor) and ours[k].dtype == v.dtype and tuple(ours[k].shape) == tuple(v.shape) and ours[k].requires_grad == v.requires_grad, k
			np.testing.assert_array_equal(ours[k].detach().cpu().numpy(), v.detach().cpu().numpy())
		else:
			assert type(ours[k]) is type(v) and ours[k] == v, k


def test_reference_checkpoint_3d(tmp_path):
	from gaussian_fluids_code_b200 import gsr3d
	gsr3d.device = torch.device('cuda', 0)
	e = load_golden('ref_checkpoint_expect.npz')
	fn = os.path.join(GOLDEN, 'ref3d_checkpoint.pt')
	gv = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., np.zeros((1, 3), np.float32), dim=3, load_file=fn)	# 3D/advance.py:365
	assert gv.N == 27 and list(gv.grid_size) == [int(v) for v in e['grid_size3']] and gv.positions.requires_grad
	assert gv.grid_scale == pytest.approx(float(e['grid_scale3']), rel=2e-6)
	grad, val = gv.gradient(torch.tensor(e['x3'], device='cuda'), need_val=True)
	assert rel_err(val.detach().cpu().numpy(), e['val3']) < 1e-5 and rel_err(grad.detach().cpu().numpy(), e['grad3']) < 1e-5
	out = str(tmp_path / 'again.pt')
	gv.save(out)
	same_dict(torch.load(out, map_location='cpu'), torch.load(fn, map_location='cpu'))
	gv.load(out, first_time=False)	# the re-load of the time loop (3D/advance_density.py:104)
	assert rel_err(gv(torch.tensor(e['x3'], device='cuda')).detach().cpu().numpy(), e['val3']) < 1e-5


def test_reference_checkpoint_2d(tmp_path):
	from gaussian_fluids_code_b200 import gsr2d
	gsr2d.device = torch.device('cuda', 0)
	e = load_golden('ref_checkpoint_expect.npz')
	fn = os.path.join(GOLDEN, 'ref2d_checkpoint.pt')
	gv = gsr2d.GaussianSplattingFast(-5., 5., -5., 5., np.zeros((1, 2), np.float32), dim=2, load_file=fn)
	assert gv.N == 25 and list(gv.grid_size) == [int(v) for v in e['grid_size2']]
	assert gv.grid_scale == pytest.approx(float(e['grid_scale2']), rel=2e-6)
	grad, val = gv.gradient(torch.tensor(e['x2'], device='cuda'), need_val=True)
	assert rel_err(val.detach().cpu().numpy(), e['val2']) < 1e-5 and rel_err(grad.detach().cpu().numpy(), e['grad2']) < 1e-5
	out = str(tmp_path / 'again.pt')
	gv.save(out)
	same_dict(torch.load(out, map_location='cpu'), torch.load(fn, map_location='cpu'))


def test_write_vti_and_obj_round_trip(tmp_path):
	"""write_vti (3D/GSR.py:728-742): ImageData of the field sampled on get_grid_points, x fastest; write_obj (:744-747): one `v` line per Gaussian"""
	from gaussian_fluids_code_b200 import gsr3d
	gsr3d.device = torch.device('cuda', 0)
	gv = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., np.zeros((1, 3), np.float32), dim=3, load_file=os.path.join(GOLDEN, 'ref3d_checkpoint.pt'))
	res = (5, 4, 3)
	f = lambda x: gv(x)[:, 0]
	fn = str(tmp_path / 'f.vti')
	gsr3d.write_vti(f, 0., 1., 0., 1., 0., 1., fn, x_N=res[0], y_N=res[1], z_N=res[2])
	txt = open(fn).read()
	assert f'WholeExtent="0 {res[0] - 1} 0 {res[1] - 1} 0 {res[2] - 1}"' in txt and 'type="Float32"' in txt
	vals = np.array(re.search(r'<DataArray[^>]*>\s*(.*?)\s*</DataArray>', txt, re.S).group(1).split(), np.float64)
	want = f(gsr3d.get_grid_points(0., 1., 0., 1., 0., 1., *res)).reshape(res).detach().cpu().numpy().ravel(order='F')	# VTK order: x fastest
	np.testing.assert_allclose(vals, want, rtol=1e-5, atol=1e-8)
	fo = str(tmp_path / 'g.obj')
	gsr3d.write_obj(gv, fo)
	pts = np.array([[float(t) for t in line.split()[1:]] for line in open(fo) if line.startswith('v ')])
	np.testing.assert_allclose(pts, gv.positions.detach().cpu().numpy(), rtol=1e-6)
