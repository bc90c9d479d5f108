"""
N3: the mesh boundary sampler of 3D/mesh_sampler.py on the GPU (gsr_mesh_tri_areas, gsr_sample_mesh), against
(i) the golden made by the reference's own kernel bodies for a recorded sequence of uniforms, (ii) the oracle on a seeded
mesh, and (iii) the statistics the sampler must have (area weighting, points on the surface, unit normals).
"""
import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


def sampler_from_golden(g):
	from gaussian_fluids_code_b200 import gsr3d
	from gaussian_fluids_code_b200.mesh_sampler import MeshSampler
	gsr3d.device = torch.device('cuda', 0)
	return MeshSampler.from_text(bytes(g['obj']).decode(), float(g['scale']), g['rotate'], g['translate'])


def test_load_and_map_match_reference_golden():
	g = load_golden('ref3d_mesh_sampler.npz')
	s = sampler_from_golden(g)
	np.testing.assert_allclose(s.vertices.cpu().numpy(), g['vertices'], rtol=0., atol=1e-7)
	np.testing.assert_allclose(s.normals.cpu().numpy(), g['normals'], rtol=0., atol=1e-7)
	np.testing.assert_array_equal(s.faces.cpu().numpy(), g['faces'])
	np.testing.assert_array_equal(s.facenormals.cpu().numpy(), g['facenormals'])
	np.testing.assert_allclose(s.area_presum.cpu().numpy(), g['area_presum'], rtol=2e-6)
	# the map uniforms -> (point, normal) with the reference's own prefix sums, so that every sample picks the reference's face
	s.area_presum = torch.tensor(g['area_presum'], device='cuda')
	s.vertices, s.normals = torch.tensor(g['vertices'], device='cuda'), torch.tensor(g['normals'], device='cuda')
	data, normal = s.sample(g['uniforms'].shape[0], uniforms=torch.tensor(g['uniforms'], device='cuda'))
	np.testing.assert_allclose(data.cpu().numpy(), g['data'], rtol=0., atol=3e-7)
	np.testing.assert_allclose(normal.cpu().numpy(), g['normal'], rtol=0., atol=1e-6)


def test_map_matches_oracle_seeded():
	import oracle.oracle as orc
	g = load_golden('ref3d_mesh_sampler.npz')
	s = sampler_from_golden(g)
	n = 100000
	u = torch.rand((n, 3), generator=torch.Generator().manual_seed(5))
	data, normal = s.sample(n, uniforms=u.cuda())
	od, on = orc.mesh_sample(u.numpy(), s.vertices.cpu().numpy(), s.normals.cpu().numpy(), s.faces.cpu().numpy(), s.facenormals.cpu().numpy(), s.area_presum.cpu().numpy())
	np.testing.assert_allclose(data.cpu().numpy(), od, rtol=0., atol=3e-7)
	np.testing.assert_allclose(normal.cpu().numpy(), on, rtol=0., atol=1e-6)


def test_statistics_and_streams():
	g = load_golden('ref3d_mesh_sampler.npz')
	s = sampler_from_golden(g)
	n = 400000
	d1, n1 = s.sample(n)
	d2, n2 = s.sample(n)
	assert not torch.equal(d1, d2)	# the call counter advances the stream
	np.testing.assert_allclose(n1.norm(dim=1).cpu().numpy(), 1., atol=1e-5)
	# the mesh is a sphere of radius `scale` around `translate` (chordal faces: points lie at most the sagitta inside it)
	c = torch.tensor(g['translate'], device='cuda')
	r = (d1 - c).norm(dim=1)
	Vv, Ff = s.vertices, s.faces.long()
	fn = torch.linalg.cross(Vv[Ff[:, 1]] - Vv[Ff[:, 0]], Vv[Ff[:, 2]] - Vv[Ff[:, 0]])
	plane = ((Vv[Ff[:, 0]] - c) * fn / fn.norm(dim=1, keepdim=True)).sum(dim=1).abs()	# distance of every face plane from the centre
	assert float(r.max()) <= float(g['scale']) * (1. + 1e-5) and float(r.min()) >= float(plane.min()) - 1e-6
	# normals are the interpolated vertex normals = radial directions, up to the interpolation error of a coarse mesh
	cosang = ((d1 - c) / r[:, None] * n1).sum(dim=1)
	assert float(cosang.min()) > .95
	# area weighting: the count of samples per face is multinomial with p = area / total (4 sigma)
	V, F = s.vertices, s.faces.long()
	area = torch.linalg.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]).norm(dim=1) * .5
	p = (area / area.sum()).double()
	# the face of a sample, found independently: the triangle whose plane holds the point with barycentric coordinates in [0, 1]
	m = 50000
	P = d1[:m].double()
	A, E1, E2 = V[F[:, 0]].double(), (V[F[:, 1]] - V[F[:, 0]]).double(), (V[F[:, 2]] - V[F[:, 0]]).double()
	nrm = torch.linalg.cross(E1, E2)
	nrm = nrm / nrm.norm(dim=1, keepdim=True)
	D = P[:, None, :] - A[None]	# (m, F, 3)
	dist = (D * nrm[None]).sum(-1).abs()
	d11, d12, d22 = (E1 * E1).sum(-1), (E1 * E2).sum(-1), (E2 * E2).sum(-1)
	b1, b2 = (D * E1[None]).sum(-1), (D * E2[None]).sum(-1)
	det = d11 * d22 - d12 * d12
	sc, tc = (b1 * d22 - b2 * d12) / det, (b2 * d11 - b1 * d12) / det
	inside = (dist < 1e-5) & (sc > -1e-4) & (tc > -1e-4) & (sc + tc < 1. + 1e-4)
	assert bool(inside.any(dim=1).all())	# every sample lies on some triangle
	face = inside.double().argmax(dim=1)
	cnt = torch.bincount(face, minlength=F.shape[0]).double()
	z = (cnt - m * p) / torch.sqrt(m * p * (1. - p))
	assert float(z.abs().max()) < 5., float(z.abs().max())


def test_scene_sampler_concatenates_box_and_mesh(tmp_path):
	from gaussian_fluids_code_b200 import gsr3d, init_cond3d
	gsr3d.device = torch.device('cuda', 0)
	g = load_golden('ref3d_mesh_sampler.npz')
	path = tmp_path / 'obstacle.obj'
	path.write_text(bytes(g['obj']).decode())
	sampler = init_cond3d.make_boundary_sampler('ring_with_obstacle', obj_file=str(path))
	data, normal = sampler(1000)
	assert data.shape == (2000, 3) and normal.shape == (2000, 3)	# 3D/init_cond.py:253-256: n on the box, then n on the mesh
	on_box = ((data[:1000] == 0.) | (data[:1000] == 1.)).any(dim=1)
	assert bool(on_box.all())
	lo, hi = sampler.mesh.bounding_box()[0::2], sampler.mesh.bounding_box()[1::2]
	inside = ((data[1000:] >= torch.tensor(lo, device='cuda') - 1e-6) & (data[1000:] <= torch.tensor(hi, device='cuda') + 1e-6)).all()
	assert bool(inside)
	out = tmp_path / 'saved.obj'
	sampler.mesh.save_obj(str(out))
	with pytest.raises(FileNotFoundError):	# the scene's own asset (assets/bunny.obj) is not shipped with the reference
		init_cond3d.make_boundary_sampler('ring_with_obstacle')
	assert out.read_text().count('\nf ') + out.read_text().startswith('f ') == sampler.mesh.faces.shape[0]
