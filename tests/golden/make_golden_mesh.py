"""
TEST INFRASTRUCTURE.  Generates tests/golden/ref3d_mesh_sampler.npz by executing the REFERENCE'S OWN mesh boundary sampler
(/root/reference/3D/mesh_sampler.py: load_obj, ti_get_tri_area, ti_lower_bound, ti_sample) as plain Python through
tests/golden/ti_shim.py on a small OBJ written here (an octahedron-based sphere: the reference's bunny.obj asset is not
shipped), with ti.random() replaced by a recorded sequence of uniforms so that the map uniforms -> (point, normal) is pinned.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_mesh.py
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


def sphere_obj(level=2):
	"""a subdivided octahedron projected on the unit sphere, as OBJ text with per-vertex normals (v, vn, f a//a b//b c//c)"""
	V = [(1., 0., 0.), (-1., 0., 0.), (0., 1., 0.), (0., -1., 0.), (0., 0., 1.), (0., 0., -1.)]
	F = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
	for _ in range(level):
		mid, F2 = {}, []

		def m(a, b):
			key = (min(a, b), max(a, b))
			if key not in mid:
				p = np.add(V[a], V[b])
				V.append(tuple(p / np.linalg.norm(p)))
				mid[key] = len(V) - 1
			return mid[key]
		for a, b, c in F:
			ab, bc, ca = m(a, b), m(b, c), m(c, a)
			F2 += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]
		F = F2
	lines = [f'v {x:.9g} {y:.9g} {z:.9g}' for x, y, z in V] + [f'vn {x:.9g} {y:.9g} {z:.9g}' for x, y, z in V]
	lines += [f'f {a + 1}//{a + 1} {b + 1}//{b + 1} {c + 1}//{c + 1}' for a, b, c in F]
	return '\n'.join(lines) + '\n'


if __name__ == '__main__':
	ti_shim.install()
	ti_shim.set_dtype(np.float32)
	sys.path.insert(0, REF)
	argv = sys.argv
	sys.argv = ['x', '--device', 'cpu', '--dir', tempfile.mkdtemp()]
	try:
		spec = importlib.util.spec_from_file_location('ref_mesh_sampler', os.path.join(REF, 'mesh_sampler.py'))
		mod = importlib.util.module_from_spec(spec)
		spec.loader.exec_module(mod)
	finally:
		sys.argv = argv
	# argument adapter: the kernel bodies assign numpy scalars into their array arguments, which a torch tensor refuses;
	# hand them numpy views of the same memory instead (the bodies themselves are the reference's, unchanged)
	for name in ('ti_get_tri_area', 'ti_sample'):
		def adapt(orig):
			return lambda self, *a: orig(self, *[t.numpy() if isinstance(t, torch.Tensor) else t for t in a])
		setattr(mod.MeshSampler, name, adapt(getattr(mod.MeshSampler, name)))
	text = sphere_obj(2)
	path = os.path.join(tempfile.mkdtemp(), 'sphere.obj')
	open(path, 'w').write(text)
	scale = .25
	ang = .3
	rotate = torch.tensor([[np.cos(ang), -np.sin(ang), 0.], [np.sin(ang), np.cos(ang), 0.], [0., 0., 1.]], dtype=torch.float32)
	translate = torch.tensor([.5, .4, .6])
	s = mod.MeshSampler(path, scale, rotate, translate)
	n = 256
	rng = np.random.default_rng(77)
	u = rng.uniform(0., 1., (n, 3)).astype(np.float32)
	u[0] = (0., 0., 0.)	# edge cases of the three draws
	u[1] = (np.float32(1.) - np.float32(2.) ** -24, np.float32(1.) - np.float32(2.) ** -24, np.float32(1.) - np.float32(2.) ** -24)
	ti_shim.set_random_sequence(u.reshape(-1))
	data, normal = s.sample(n)
	np.savez_compressed(os.path.join(HERE, 'ref3d_mesh_sampler.npz'), obj=np.frombuffer(text.encode(), np.uint8), scale=np.float32(scale), rotate=rotate.numpy(),
						translate=translate.numpy(), uniforms=u, vertices=s.vertices.numpy(), normals=s.normals.numpy(), faces=s.faces.numpy(),
						facenormals=s.facenormals.numpy(), area_presum=s.area_presum.numpy(), data=data.numpy(), normal=normal.numpy())
	print('faces', s.faces.shape[0], 'area', float(s.area_presum[-1]), 'sphere area', 4 * np.pi * scale ** 2)
