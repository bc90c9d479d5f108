// sampling.cu — sample-point generation of the per-timestep optimisation ("sample generation").
//
// The reference draws its training samples with ~10 small torch ops per iteration (3D/advance.py:339-340:
// rand_like(positions) * extent + min; 3D/init_cond.py:227-249: area-weighted points on the six faces of the domain box
// with inward normals, through boolean-mask indexing that synchronises the host).  Here each sample set is ONE kernel:
// a counter-based Philox4x32-10 generator keyed by (seed, stream id) and indexed by (sample index, iteration), where
// the iteration number is read from DEVICE memory (the optimiser's step count), so that a captured CUDA graph draws
// fresh samples on every replay without any host involvement.
#include "sampling.cuh"

namespace gsr {

__global__ void sample_box_kernel(Box b, int n, uint64_t seed, uint32_t stream_id, const float *__restrict__ iteration, float *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t it = iteration ? (uint32_t)__ldg(iteration) : 0u;
	const Philox r(seed, stream_id, (uint32_t)i, it);
#pragma unroll
	for (int k = 0; k < 3; k++) out[3 * (size_t)i + k] = fmaf(r.u(k), b.ext[k], b.lo[k]);
}

__global__ void sample_box_surface_kernel(Box b, int n, uint64_t seed, uint32_t stream_id, const float *__restrict__ iteration,
					  float *__restrict__ data, float *__restrict__ normal)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t it = iteration ? (uint32_t)__ldg(iteration) : 0u;
	const Philox r(seed, stream_id, (uint32_t)i, it);
	float p[3], nm[3];
	box_surface_point(b, r, p, nm);
#pragma unroll
	for (int k = 0; k < 3; k++) {
		data[3 * (size_t)i + k] = p[k];
		normal[3 * (size_t)i + k] = nm[k];
	}
}

// 3D/mesh_sampler.py:60-88 (ti_lower_bound, ti_sample): a triangle chosen with probability proportional to its area by a
// binary search in the inclusive prefix sums, a uniform point on it (u = 1 - sqrt(r1), v = r2 (1 - u)), the vertex normals
// interpolated with the same weights and normalised.  The three uniforms of a sample come from Philox, or — for parity tests
// against the reference's map — from a caller-supplied (n,3) array.
__global__ void sample_mesh_kernel(int n, const float *__restrict__ vertices, const float *__restrict__ normals, const int32_t *__restrict__ faces,
				   const int32_t *__restrict__ facenormals, const float *__restrict__ area_presum, int F,
				   uint64_t seed, uint32_t stream_id, const float *__restrict__ iteration, const float *__restrict__ uniforms,
				   float *__restrict__ data, float *__restrict__ normal)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float r0, r1, r2;
	if (uniforms) {
		r0 = uniforms[3 * (size_t)i]; r1 = uniforms[3 * (size_t)i + 1]; r2 = uniforms[3 * (size_t)i + 2];
	} else {
		const uint32_t it = iteration ? (uint32_t)__ldg(iteration) : 0u;
		const Philox r(seed, stream_id, (uint32_t)i, it);
		r0 = r.u(0); r1 = r.u(1); r2 = r.u(2);
	}
	const float t = __fmul_rn(r0, __ldg(area_presum + F - 1));
	int l = 0, h = F;
	while (l < h) {	// first index with area_presum[m] >= t
		const int m = (l + h) >> 1;
		if (__ldg(area_presum + m) < t) l = m + 1;
		else h = m;
	}
	const int f = min(l, F - 1);
	const float u = 1.f - sqrtf(r1), v = r2 * (1.f - u), w = 1.f - u - v;
	float p[3], nn[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const float a = vertices[3 * (size_t)faces[3 * (size_t)f] + k], b = vertices[3 * (size_t)faces[3 * (size_t)f + 1] + k], c = vertices[3 * (size_t)faces[3 * (size_t)f + 2] + k];
		p[k] = u * a + v * b + w * c;
		const float na = normals[3 * (size_t)facenormals[3 * (size_t)f] + k], nb = normals[3 * (size_t)facenormals[3 * (size_t)f + 1] + k],
			    nc = normals[3 * (size_t)facenormals[3 * (size_t)f + 2] + k];
		nn[k] = u * na + v * nb + w * nc;
	}
	const float len = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
#pragma unroll
	for (int k = 0; k < 3; k++) {
		data[3 * (size_t)i + k] = p[k];
		normal[3 * (size_t)i + k] = nn[k] / len;
	}
}

// 3D/mesh_sampler.py:12-21 (ti_get_tri_area), the per-face part: area[i] = |(b - a) x (c - a)| / 2
__global__ void tri_area_kernel(const float *__restrict__ vertices, const int32_t *__restrict__ faces, int F, float *__restrict__ area)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= F) return;
	float a[3], e1[3], e2[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		a[k] = vertices[3 * (size_t)faces[3 * (size_t)i] + k];
		e1[k] = vertices[3 * (size_t)faces[3 * (size_t)i + 1] + k] - a[k];
		e2[k] = vertices[3 * (size_t)faces[3 * (size_t)i + 2] + k] - a[k];
	}
	const float c0 = e1[1] * e2[2] - e1[2] * e2[1], c1 = e1[2] * e2[0] - e1[0] * e2[2], c2 = e1[0] * e2[1] - e1[1] * e2[0];
	area[i] = sqrtf(c0 * c0 + c1 * c1 + c2 * c2) * .5f;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_sample_box(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev, float *out, void *stream)
{
	if (!box || n < 0 || n >= ((int64_t)1 << 31) || !out) return GSR_EINVAL;
	if (n == 0) return GSR_OK;
	g_launches += 1;
	sample_box_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_box(box), (int)n, seed, stream_id, iteration_dev, out);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_sample_box_surface(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev, float *data, float *normal,
				      void *stream)
{
	if (!box || n < 0 || n >= ((int64_t)1 << 31) || !data || !normal) return GSR_EINVAL;
	if (n == 0) return GSR_OK;
	g_launches += 1;
	sample_box_surface_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_box(box), (int)n, seed, stream_id, iteration_dev, data, normal);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_mesh_tri_areas(const float *vertices, const int32_t *faces, int64_t F, float *area, void *stream)
{
	if (F < 0 || F >= ((int64_t)1 << 31) || (F > 0 && (!vertices || !faces || !area))) return GSR_EINVAL;
	if (F == 0) return GSR_OK;
	g_launches += 1;
	tri_area_kernel<<<(int)((F + 255) / 256), 256, 0, (cudaStream_t)stream>>>(vertices, faces, (int)F, area);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_sample_mesh(int64_t n, const float *vertices, const float *normals, const int32_t *faces, const int32_t *facenormals,
			       const float *area_presum, int64_t F, uint64_t seed, uint32_t stream_id, const float *iteration_dev,
			       const float *uniforms, float *data, float *normal, void *stream)
{
	if (n < 0 || n >= ((int64_t)1 << 31) || F <= 0 || F >= ((int64_t)1 << 31) || !vertices || !normals || !faces || !facenormals || !area_presum || !data || !normal)
		return GSR_EINVAL;
	if (n == 0) return GSR_OK;
	g_launches += 1;
	sample_mesh_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)n, vertices, normals, faces, facenormals, area_presum, (int)F, seed, stream_id,
											iteration_dev, uniforms, data, normal);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
