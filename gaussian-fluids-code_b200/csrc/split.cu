// split.cu — reseeding of over-stretched Gaussians on the device (reference
// 3D/advance.py:51-94, 2D/advance.py:58-93).
//
// clone_velocity_field replaces every Gaussian whose axis ratio exp(max s - min s) reaches a threshold (2 in 3D, 1.5 in 2D) by TWO
// samples of its own distribution N(mu, Sigma) with the longest axis shortened.  The reference does this with ~25 torch ops, a
// batched Cholesky inside torch.distributions.MultivariateNormal and boolean-mask indexing (host syncs); here:
//   gsr_split_flags      flag[i] = ratio_i >= threshold, and the number of flagged Gaussians (one device int);
//   gsr_split_apply      stream compaction + append in one launch: the kept Gaussians move to the front in their old order, the
//                        children go behind them as [first samples of all parents | second samples of all parents] — the layout
//                        of the reference's `.sample((2,)).flatten(0, 1)` / `.repeat(2, 1)` — with the split bookkeeping
//                        (stop_gradient = 1 for kept, 0 for children).  Child position = mu + chol(Sigma) z, clamped to the
//                        extended domain (3D; the 2D reference does not clamp): Sigma^-1 = R diag(e^{2s}) R^T is symmetrised as the
//                        reference does ((P + P^T) / 2), inverted in closed form, factorised by an unrolled Cholesky.  z: standard
//                        normals from a caller buffer (tests replay the reference's draws) or from Philox + Box-Muller.
// The exclusive prefix of the flags comes from the caller (one scan of N ints).
#include "common.cuh"
#include "sampling.cuh"
#include <math.h>

namespace gsr {

__global__ void split_flags_kernel(const float *__restrict__ scal, int N, int D, float thr, int32_t *__restrict__ flags, int32_t *__restrict__ count)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	int f = 0;
	if (i < N) {
		float mn = scal[(size_t)D * i], mx = mn;
		for (int k = 1; k < D; k++) {
			const float v = scal[(size_t)D * i + k];
			mn = fminf(mn, v);
			mx = fmaxf(mx, v);
		}
		// exp(max - min) >= thr, decided as the reference decides it: on the exponential, in float32
		f = expf(mx - mn) >= thr ? 1 : 0;
		flags[i] = f;
	}
	const unsigned b = __ballot_sync(0xffffffffu, f);
	if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

struct SplitArgs {
	const float *pos, *scal, *rot, *val;
	const int32_t *flags, *prefix;	// prefix: exclusive scan of flags
	const float *normals;		// (2, n_split, D) or NULL
	float *opos, *oscal, *orot, *oval;
	int32_t *ostop;
	int N, n_split;
	float lo[3], hi[3];
	int clamp;
	float log_axis, log_all;	// children's log inverse radii: the split axis += log_axis, every axis -= log_all
	unsigned long long seed;
};

__device__ __forceinline__ void normal_pair(const Philox &r, int k, float &a, float &b)
{
	// Box-Muller on two of the four 24-bit uniforms (u in (0, 1])
	const float u1 = 1.f - r.u(k), u2 = r.u(k + 1);
	const float m = sqrtf(-2.f * logf(u1));
	a = m * cosf(6.28318530717958647692f * u2);
	b = m * sinf(6.28318530717958647692f * u2);
}

template <int D>
__global__ void split_apply_kernel(SplitArgs a)
{
	constexpr int RD = (D == 3) ? 4 : 1;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.N) return;
	const int before = a.prefix[i];	// flagged Gaussians before i
	const int n_keep = a.N - a.n_split;
	if (!a.flags[i]) {
		const int o = i - before;
#pragma unroll
		for (int k = 0; k < D; k++) {
			a.opos[(size_t)D * o + k] = a.pos[(size_t)D * i + k];
			a.oscal[(size_t)D * o + k] = a.scal[(size_t)D * i + k];
			a.oval[(size_t)D * o + k] = a.val[(size_t)D * i + k];
		}
#pragma unroll
		for (int k = 0; k < RD; k++) a.orot[(size_t)RD * o + k] = a.rot[(size_t)RD * i + k];
		a.ostop[o] = 1;
		return;
	}
	const int j = before;	// index among the split Gaussians
	float s[D], mu[D];
#pragma unroll
	for (int k = 0; k < D; k++) { s[k] = a.scal[(size_t)D * i + k]; mu[k] = a.pos[(size_t)D * i + k]; }
	// precision P = R diag(e^{2 s}) R^T (get_variances), symmetrised, and its Cholesky-factorised inverse
	float L[D][D];
	if (D == 3) {
		float q[4];
		float nq = 0.f;
#pragma unroll
		for (int k = 0; k < 4; k++) { q[k] = a.rot[(size_t)4 * i + k]; nq += q[k] * q[k]; }
		nq = sqrtf(nq);
		const float r = q[0] / nq, x = q[1] / nq, y = q[2] / nq, z = q[3] / nq;
		const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
				       {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
				       {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
		// covariance C = P^-1 = R diag(e^{-2 s}) R^T: the inverse of an orthogonal similarity is taken on the diagonal
		float C[3][3];
#pragma unroll
		for (int p = 0; p < 3; p++)
#pragma unroll
			for (int q_ = 0; q_ < 3; q_++) {
				float c = 0.f;
#pragma unroll
				for (int k = 0; k < 3; k++) c += R[p][k] * expf(-2.f * s[k]) * R[q_][k];
				C[p][q_] = c;
			}
#pragma unroll
		for (int p = 0; p < 3; p++)
#pragma unroll
			for (int q_ = 0; q_ < p; q_++) C[p][q_] = C[q_][p] = .5f * (C[p][q_] + C[q_][p]);
		L[0][0] = sqrtf(C[0][0]);
		L[1][0] = C[1][0] / L[0][0];
		L[2][0] = C[2][0] / L[0][0];
		L[1][1] = sqrtf(fmaxf(C[1][1] - L[1][0] * L[1][0], 0.f));
		L[2][1] = (C[2][1] - L[2][0] * L[1][0]) / L[1][1];
		L[2][2] = sqrtf(fmaxf(C[2][2] - L[2][0] * L[2][0] - L[2][1] * L[2][1], 0.f));
		L[0][1] = L[0][2] = L[1][2] = 0.f;
	} else {
		const float th = a.rot[i];
		const float c = cosf(th), sn = sinf(th);
		const float e0 = expf(-2.f * s[0]), e1 = expf(-2.f * s[1]);
		const float C00 = c * c * e0 + sn * sn * e1, C01 = c * sn * (e0 - e1), C11 = sn * sn * e0 + c * c * e1;
		L[0][0] = sqrtf(C00);
		L[1][0] = C01 / L[0][0];
		L[1][1] = sqrtf(fmaxf(C11 - L[1][0] * L[1][0], 0.f));
		L[0][1] = 0.f;
	}
	// the children's scalings: the longest axis (smallest log inverse radius) is shortened
	int ax = 0;
#pragma unroll
	for (int k = 1; k < D; k++)
		if (s[k] < s[ax]) ax = k;	// first minimum, like torch.min
	if (D == 2) ax = (s[1] < s[0]) ? 1 : 0;	// 2D/advance.py:75-77
	float sc[D];
#pragma unroll
	for (int k = 0; k < D; k++) sc[k] = s[k] + (k == ax ? a.log_axis : 0.f) - a.log_all;
	float zz[2][D];
	if (a.normals) {
#pragma unroll
		for (int c = 0; c < 2; c++)
#pragma unroll
			for (int k = 0; k < D; k++) zz[c][k] = a.normals[((size_t)c * a.n_split + j) * D + k];
	} else {
		Philox r0(a.seed, 0x51u, (uint32_t)i, 0u), r1(a.seed, 0x51u, (uint32_t)i, 1u);
		float n0, n1, n2, n3, n4, n5, n6, n7;
		normal_pair(r0, 0, n0, n1);
		normal_pair(r0, 2, n2, n3);
		normal_pair(r1, 0, n4, n5);
		normal_pair(r1, 2, n6, n7);
		const float all[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
#pragma unroll
		for (int c = 0; c < 2; c++)
#pragma unroll
			for (int k = 0; k < D; k++) zz[c][k] = all[c * 4 + k];
	}
#pragma unroll
	for (int c = 0; c < 2; c++) {
		const int o = n_keep + c * a.n_split + j;
#pragma unroll
		for (int p = 0; p < D; p++) {
			float v = mu[p];
#pragma unroll
			for (int k = 0; k <= p; k++) v += L[p][k] * zz[c][k];
			if (a.clamp) v = fminf(fmaxf(v, a.lo[p]), a.hi[p]);
			a.opos[(size_t)D * o + p] = v;
			a.oscal[(size_t)D * o + p] = sc[p];
			a.oval[(size_t)D * o + p] = a.val[(size_t)D * i + p];
		}
#pragma unroll
		for (int k = 0; k < RD; k++) a.orot[(size_t)RD * o + k] = a.rot[(size_t)RD * i + k];
		a.ostop[o] = 0;
	}
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_split_flags(int D, const float *scalings, int64_t N, float ratio_threshold, int32_t *flags, int32_t *count, void *stream)
{
	if ((D != 2 && D != 3) || N < 0 || N >= ((int64_t)1 << 30) || !scalings || !flags || !count || !(ratio_threshold > 0.f)) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
	if (e != cudaSuccess) return (int)e;
	if (N == 0) return GSR_OK;
	g_launches += 1;
	split_flags_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(scalings, (int)N, D, ratio_threshold, flags, count);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_split_apply(int D, const float *positions, const float *scalings, const float *rotations, const float *values, int64_t N,
			       const int32_t *flags, const int32_t *flag_prefix, int64_t n_split, const float *normals, uint64_t seed,
			       const float *clamp_box, float log_axis, float log_all,
			       float *out_positions, float *out_scalings, float *out_rotations, float *out_values, int32_t *out_stop_gradient, void *stream)
{
	if ((D != 2 && D != 3) || N <= 0 || N >= ((int64_t)1 << 30) || n_split < 0 || n_split > N) return GSR_EINVAL;
	if (!positions || !scalings || !rotations || !values || !flags || !flag_prefix || !out_positions || !out_scalings || !out_rotations || !out_values || !out_stop_gradient)
		return GSR_EINVAL;
	SplitArgs a;
	a.pos = positions; a.scal = scalings; a.rot = rotations; a.val = values; a.flags = flags; a.prefix = flag_prefix; a.normals = normals;
	a.opos = out_positions; a.oscal = out_scalings; a.orot = out_rotations; a.oval = out_values; a.ostop = out_stop_gradient;
	a.N = (int)N; a.n_split = (int)n_split;
	a.clamp = clamp_box ? 1 : 0;
	for (int k = 0; k < 3; k++) {
		a.lo[k] = (clamp_box && k < D) ? clamp_box[2 * k] : 0.f;
		a.hi[k] = (clamp_box && k < D) ? clamp_box[2 * k + 1] : 0.f;
	}
	a.log_axis = log_axis; a.log_all = log_all; a.seed = seed;
	cudaStream_t st = (cudaStream_t)stream;
	g_launches += 1;
	if (D == 3) split_apply_kernel<3><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
	else split_apply_kernel<2><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(a);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
