"""
Mesh boundary sampler — the drop-in for the reference's 3D/mesh_sampler.py: OBJ reader / writer on the
host, triangle areas and the per-iteration sampling on the GPU (csrc/sampling.cu: gsr_mesh_tri_areas, gsr_sample_mesh).
Same constructor and `sample(n)` as the reference; samples come from the engine's counter-based Philox generator
(seed, stream id, iteration), not from Taichi's.
"""
import ctypes as C

import torch

from . import _lib, gsr3d
from ._lib import check, ptr, stream


def parse_obj(text):
	"""v / vn / f records of an OBJ text (3D/mesh_sampler.py:23-36): faces index vertices by the first field of `a/b/c`, normals by the last"""
	vertices, normals, faces, facenormals = [], [], [], []
	for line in text.splitlines():
		if line.startswith('v '):
			vertices.append([float(t) for t in line.split()[1:4]])
		elif line.startswith('vn '):
			normals.append([float(t) for t in line.split()[1:4]])
		elif line.startswith('f '):
			items = line.split()[1:]
			if len(items) != 3:
				raise _lib.GsrError('MeshSampler: only triangle faces are supported (as in the reference)')
			faces.append([int(t.split('/')[0]) - 1 for t in items])
			facenormals.append([int(t.split('/')[-1]) - 1 for t in items])
	if not vertices or not faces or not normals:
		raise _lib.GsrError('MeshSampler: the OBJ needs v, vn and f records')
	return vertices, normals, faces, facenormals


class MeshSampler:
	def __init__(self, obj_file, scale, rotate, translate, seed=42, stream_id=7):
		with open(obj_file, 'r') as fd:
			self._load(fd.read(), scale, rotate, translate)
		self.seed, self.stream_id = int(seed), int(stream_id)
		self._draws = torch.zeros(1, dtype=torch.float32, device=gsr3d.device)	# advanced by every sample() call, read on the device

	@classmethod
	def from_text(cls, text, scale, rotate, translate, seed=42, stream_id=7):
		self = cls.__new__(cls)
		self._load(text, scale, rotate, translate)
		self.seed, self.stream_id = int(seed), int(stream_id)
		self._draws = torch.zeros(1, dtype=torch.float32, device=gsr3d.device)
		return self

	def _load(self, text, scale, rotate, translate):
		dev = gsr3d.device
		vertices, normals, faces, facenormals = parse_obj(text)
		rotate = torch.as_tensor(rotate, dtype=torch.float32, device=dev)
		translate = torch.as_tensor(translate, dtype=torch.float32, device=dev)
		# 3D/mesh_sampler.py:37-41
		self.vertices = ((rotate[None] @ (scale * torch.tensor(vertices, dtype=torch.float32, device=dev)).unsqueeze(-1)).squeeze(-1) + translate).contiguous()
		nrm = (rotate[None] @ torch.tensor(normals, dtype=torch.float32, device=dev).unsqueeze(-1)).squeeze(-1)
		self.normals = (nrm / ((nrm ** 2.).sum(dim=-1) ** .5)[:, None]).contiguous()
		self.faces = torch.tensor(faces, dtype=torch.int32, device=dev)
		self.facenormals = torch.tensor(facenormals, dtype=torch.int32, device=dev)
		if int(self.faces.max()) >= self.vertices.shape[0] or int(self.facenormals.max()) >= self.normals.shape[0] or int(self.faces.min()) < 0 or int(self.facenormals.min()) < 0:
			raise _lib.GsrError('MeshSampler: face index out of range')
		area = torch.empty(self.faces.shape[0], dtype=torch.float32, device=dev)
		check(_lib.lib().gsr_mesh_tri_areas(ptr(self.vertices), ptr(self.faces, torch.int32), C.c_int64(self.faces.shape[0]), ptr(area), stream()), 'gsr_mesh_tri_areas')
		self.area_presum = torch.cumsum(area, 0).contiguous()	# the reference's serial prefix sum (:19-21)

	def bounding_box(self):
		lo, hi = self.vertices.min(dim=0).values, self.vertices.max(dim=0).values
		return tuple(float(v) for pair in zip(lo.tolist(), hi.tolist()) for v in pair)

	def save_obj(self, obj_file):
		"""3D/mesh_sampler.py:48-55"""
		v, n, f, fn = self.vertices.cpu(), self.normals.cpu(), self.faces.cpu(), self.facenormals.cpu()
		with open(obj_file, 'w') as fd:
			for i in range(v.shape[0]):
				fd.write(f'v {v[i, 0].item()} {v[i, 1].item()} {v[i, 2].item()}\n')
			for i in range(n.shape[0]):
				fd.write(f'vn {n[i, 0].item()} {n[i, 1].item()} {n[i, 2].item()}\n')
			for i in range(f.shape[0]):
				fd.write(f'f {f[i, 0].item() + 1}//{fn[i, 0].item() + 1} {f[i, 1].item() + 1}//{fn[i, 1].item() + 1} {f[i, 2].item() + 1}//{fn[i, 2].item() + 1}\n')

	def sample(self, n, uniforms=None, data=None, normal=None, iteration=None):
		"""n points on the surface with interpolated unit normals (3D/mesh_sampler.py:90-94).  uniforms (n,3): use exactly these
		draws (parity tests); iteration: device float scalar that indexes the Philox stream (default: an internal call counter)"""
		dev = self.vertices.device
		data = torch.empty((n, 3), dtype=torch.float32, device=dev) if data is None else data
		normal = torch.empty((n, 3), dtype=torch.float32, device=dev) if normal is None else normal
		it = iteration if iteration is not None else self._draws
		check(_lib.lib().gsr_sample_mesh(C.c_int64(n), ptr(self.vertices), ptr(self.normals), ptr(self.faces, torch.int32), ptr(self.facenormals, torch.int32),
										 ptr(self.area_presum), C.c_int64(self.faces.shape[0]), C.c_uint64(self.seed), C.c_uint32(self.stream_id),
										 ptr(it), ptr(uniforms, allow_none=True), ptr(data), ptr(normal), stream()), 'gsr_sample_mesh')
		if iteration is None and uniforms is None:
			self._draws += 1.
		return data, normal
