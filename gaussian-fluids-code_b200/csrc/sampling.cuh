// sampling.cuh — the counter-based generator and the per-sample maps of sampling.cu, shared with the kernels that generate
// and consume samples in one launch (grid.cu: cell_bin_kernel<D, true>).
#pragma once
#include "common.cuh"

namespace gsr {

struct Philox {
	uint32_t c[4];
	__device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
	{
		const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
		const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
		c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
	}
	__device__ __forceinline__ Philox(uint64_t seed, uint32_t stream_id, uint32_t index, uint32_t iteration)
	{
		c[0] = index; c[1] = iteration; c[2] = stream_id; c[3] = 0x9E3779B9u;
		uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
		for (int r = 0; r < 10; r++) {
			round(c, k0, k1);
			k0 += 0x9E3779B9u;
			k1 += 0xBB67AE85u;
		}
	}
	// uniform in [0, 1): 24 random bits, like torch.rand for float32
	__device__ __forceinline__ float u(int k) const { return (float)(c[k] >> 8) * (1.f / 16777216.f); }
};

struct Box {
	float lo[3], ext[3];
};

// 3D/init_cond.py:227-249: face chosen with probability proportional to its area (order x_min, x_max, y_min, y_max,
// z_min, z_max), uniform position on the face, inward unit normal.
__device__ __forceinline__ void box_surface_point(const Box &b, const Philox &r, float (&p)[3], float (&nm)[3])
{
	const float ayz = b.ext[1] * b.ext[2], azx = b.ext[2] * b.ext[0], axy = b.ext[0] * b.ext[1];
	const float t = r.u(3) * (ayz + azx + axy) * 2.f;
	int face;
	if (t < ayz) face = 0;
	else if (t < 2.f * ayz) face = 1;
	else if (t < 2.f * ayz + azx) face = 2;
	else if (t < 2.f * (ayz + azx)) face = 3;
	else if (t < 2.f * (ayz + azx) + axy) face = 4;
	else face = 5;
	const int axis = face >> 1, upper = face & 1;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		p[k] = fmaf(r.u(k), b.ext[k], b.lo[k]);
		nm[k] = 0.f;
	}
#pragma unroll
	for (int k = 0; k < 3; k++) {
		if (k == axis) {
			p[k] = upper ? b.lo[k] + b.ext[k] : b.lo[k];
			nm[k] = upper ? -1.f : 1.f;
		}
	}
}

inline Box make_box(const float *box)
{
	Box b;
	for (int k = 0; k < 3; k++) {
		b.lo[k] = box[2 * k];
		b.ext[k] = box[2 * k + 1] - box[2 * k];
	}
	return b;
}

}  // namespace gsr
