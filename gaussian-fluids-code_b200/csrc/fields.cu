// fields.cu — analytic initial fields of the 3D scenes: the regularised Biot-Savart sum over the vortex
// particles of a ring, velocity and Jacobian in one pass (3D/init_cond.py:122-145, the Taichi kernels vortex_particle and
// vortex_particle_gradient).  Dense all-pairs: Q points x M particles (M = 500 per ring), called for every batch of the
// initial fit (3D/initialize.py:23-24).  FP32-issue bound; HBM traffic is the points in and 12 floats out.
//
// One thread per point; the particles are staged in shared memory as two float4 {x0, w.x} {w.y, w.z, -, -} and read by all
// lanes at one address (broadcast).  expf is the full-accuracy libm form, as the reference's f32 exp.
#include "common.cuh"

namespace gsr {

constexpr int BS_THREADS = 128;
constexpr int BS_CHUNK = 512;	// particles per shared-memory stage

template <bool VAL, bool GRAD>
__global__ void __launch_bounds__(BS_THREADS) vortex_particles_kernel(const float *__restrict__ x, int Q, const float *__restrict__ x0, const float *__restrict__ w, int M,
								       float U, float a, float *__restrict__ val, float *__restrict__ grad)
{
	__shared__ float4 sp[2 * BS_CHUNK];
	const int i = blockIdx.x * BS_THREADS + threadIdx.x;
	const bool on = i < Q;
	const float px = on ? x[3 * (size_t)i] : 0.f, py = on ? x[3 * (size_t)i + 1] : 0.f, pz = on ? x[3 * (size_t)i + 2] : 0.f;
	float u[3] = {0.f, 0.f, 0.f}, J[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	const float ia3 = 3.f / (a * a * a);
	for (int m0 = 0; m0 < M; m0 += BS_CHUNK) {
		const int mc = min(BS_CHUNK, M - m0);
		__syncthreads();
		for (int k = threadIdx.x; k < mc; k += BS_THREADS) {
			const size_t j = (size_t)(m0 + k);
			sp[2 * k] = make_float4(x0[3 * j], x0[3 * j + 1], x0[3 * j + 2], w[3 * j]);
			sp[2 * k + 1] = make_float4(w[3 * j + 1], w[3 * j + 2], 0.f, 0.f);
		}
		__syncthreads();
		if (!on) continue;
		for (int k = 0; k < mc; k++) {
			const float4 A = sp[2 * k], B = sp[2 * k + 1];
			const float d0 = px - A.x, d1 = py - A.y, d2 = pz - A.z;
			const float w0 = A.w, w1 = B.x, w2 = B.y;
			const float r2 = d0 * d0 + d1 * d1 + d2 * d2, r = sqrtf(r2);
			const float q = r / a, e = expf(-(q * q * q));
			const float ir = 1.f / r, ir3 = ir * ir * ir;
			const float fr = ir3 * (1.f - e);
			const float c0 = w1 * d2 - w2 * d1, c1 = w2 * d0 - w0 * d2, c2 = w0 * d1 - w1 * d0;	// w x d
			if (VAL) {
				u[0] += fr * c0; u[1] += fr * c1; u[2] += fr * c2;
			}
			if (GRAD) {
				const float g = (-3.f * ir3 * ir * (1.f - e) + ia3 * ir * e) * ir;	// f'(r) / r
				J[0] += g * c0 * d0;            J[1] += g * c0 * d1 - fr * w2;  J[2] += g * c0 * d2 + fr * w1;
				J[3] += g * c1 * d0 + fr * w2;  J[4] += g * c1 * d1;            J[5] += g * c1 * d2 - fr * w0;
				J[6] += g * c2 * d0 - fr * w1;  J[7] += g * c2 * d1 + fr * w0;  J[8] += g * c2 * d2;
			}
		}
	}
	if (!on) return;
	if (VAL) {
#pragma unroll
		for (int k = 0; k < 3; k++) val[3 * (size_t)i + k] += U * u[k];
	}
	if (GRAD) {
#pragma unroll
		for (int k = 0; k < 9; k++) grad[9 * (size_t)i + k] += U * J[k];
	}
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_vortex_particles(const float *x, int64_t Q, const float *x0, const float *w, int64_t M, float U, float a,
				    float *val, float *grad, void *stream)
{
	if (Q < 0 || M < 0 || Q >= ((int64_t)1 << 30) || M >= ((int64_t)1 << 30) || (!val && !grad) || !(a > 0.f)) return GSR_EINVAL;
	if (Q == 0 || M == 0) return GSR_OK;
	if (!x || !x0 || !w) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	const int blocks = (int)((Q + BS_THREADS - 1) / BS_THREADS);
	g_launches += 1;
	if (val && grad) vortex_particles_kernel<true, true><<<blocks, BS_THREADS, 0, st>>>(x, (int)Q, x0, w, (int)M, U, a, val, grad);
	else if (val) vortex_particles_kernel<true, false><<<blocks, BS_THREADS, 0, st>>>(x, (int)Q, x0, w, (int)M, U, a, val, grad);
	else vortex_particles_kernel<false, true><<<blocks, BS_THREADS, 0, st>>>(x, (int)Q, x0, w, (int)M, U, a, val, grad);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
