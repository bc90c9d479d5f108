// misc.cu — version string and the FP32-FMA / MUFU.EX2 peak micro-benchmarks that provide the roofline
// denominators for the evaluation kernels ("the builder must measure FMA and ex2.approx peaks").
#include "common.cuh"
#include "f32x2.cuh"

namespace gsr {

std::atomic<unsigned long long> g_launches{0};

constexpr int PK_CHAINS = 8;

__global__ void __launch_bounds__(1024) peak_fma_kernel(int iters, float a, float b, float *out)
{
	float x[PK_CHAINS];
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) x[k] = threadIdx.x * 1e-3f + k;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < PK_CHAINS; k++) x[k] = fmaf(x[k], a, b);
	}
	float s = 0.f;
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) s += x[k];
	if (s == 123.456f) out[0] = s;	// never true; keeps the loop alive
}

__global__ void __launch_bounds__(1024) peak_mufu_kernel(int iters, float *out)
{
	float x[PK_CHAINS];
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) x[k] = threadIdx.x * 1e-3f + k * .1f;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < PK_CHAINS; k++) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[k]) : "f"(-x[k]));
	}
	float s = 0.f;
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) s += x[k];
	if (s == 123.456f) out[0] = s;
}

// ---- pipe probes (tools/pipe_probe.py): what the FP32 pipe sustains for the operand shapes of the evaluation kernels ----
// which: 0 FFMA with three distinct per-thread registers   1 FFMA reg*reg+same reg (x = x*y + x)
//        2 FFMA reg * uniform + reg (operand from a kernel parameter)   3 FADD reg+reg   4 FMUL reg*reg
//        5 the candidate test of the forward kernels (3 FADD + 3 FMUL + 8 FFMA + FSETP, register operands)
//        6 LDS.128 with a warp-uniform address (3 per iteration, no math)
//        7 FFMA2 three distinct register pairs   8 FFMA2 pair * broadcast scalar + pair   9 probe 5 written with f32x2 pairs
template <int WHICH>
__global__ void __launch_bounds__(256) pipe_probe_kernel(int iters, const float *__restrict__ in, float ua, float *out)
{
	__shared__ float4 sm[96];
	float x[PK_CHAINS], y[PK_CHAINS], z[PK_CHAINS];
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) {
		x[k] = in[threadIdx.x + k];
		y[k] = in[threadIdx.x + 8 + k];
		z[k] = in[threadIdx.x + 16 + k];
	}
	if (threadIdx.x < 96) sm[threadIdx.x] = make_float4(x[0], y[0], z[0], x[1]);
	__syncthreads();
	float acc = 0.f;
	for (int i = 0; i < iters; i++) {
		if (WHICH == 0) {
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = fmaf(x[k], y[k], z[k]);
		} else if (WHICH == 1) {
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = fmaf(x[k], y[k], x[k]);
		} else if (WHICH == 2) {
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = fmaf(x[k], ua, z[k]);
		} else if (WHICH == 3) {
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = x[k] + y[k];
		} else if (WHICH == 4) {
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = x[k] * y[k];
		} else if (WHICH == 5) {
			// 4 points (x[0..2], x[3..5], y[0..2], y[3..5]) against one "candidate" held in z[0..7], x[6]
#pragma unroll
			for (int p = 0; p < 4; p++) {
				const float *pt = (p < 2) ? &x[3 * p] : &y[3 * (p - 2)];
				const float dx = pt[0] - z[0], dy = pt[1] - z[1], dz = pt[2] - z[2];
				const float wx = fmaf(z[5], dz, fmaf(z[4], dy, __fmul_rn(z[3], dx)));
				const float wy = fmaf(z[7], dz, fmaf(z[6], dy, __fmul_rn(z[4], dx)));
				const float wz = fmaf(x[6], dz, fmaf(z[7], dy, __fmul_rn(z[5], dx)));
				const float q = fmaf(dz, wz, fmaf(dy, wy, __fmul_rn(dx, wx)));
				if (q <= ua) acc += q;
			}
			z[0] += 1e-3f;	// a new candidate each iteration
		} else if (WHICH == 6) {
			const float4 a = sm[(i * 3) % 96], b = sm[(i * 3 + 1) % 96], c = sm[(i * 3 + 2) % 96];
			acc += a.x + b.y + c.z;
		} else if (WHICH == 7) {	// FFMA2, three distinct register pairs (4 chains of 2)
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k += 2) {
				f2 r = fma2(pack2(x[k], x[k + 1]), pack2(y[k], y[k + 1]), pack2(z[k], z[k + 1]));
				unpack2(r, x[k], x[k + 1]);
			}
		} else if (WHICH == 8) {	// FFMA2, pair * broadcast scalar + pair
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k += 2) {
				f2 r = fma2(pack2(x[k], x[k + 1]), bc(y[k]), pack2(z[k], z[k + 1]));
				unpack2(r, x[k], x[k + 1]);
			}
		} else {	// WHICH == 9: the candidate test on 4 points as 2 packed pairs
			const f2 X[2] = {pack2(x[0], x[3]), pack2(y[0], y[3])}, Y[2] = {pack2(x[1], x[4]), pack2(y[1], y[4])}, Z[2] = {pack2(x[2], x[5]), pack2(y[2], y[5])};
#pragma unroll
			for (int p = 0; p < 2; p++) {
				const f2 dx = add2(X[p], bc(-z[0])), dy = add2(Y[p], bc(-z[1])), dz = add2(Z[p], bc(-z[2]));
				const f2 wx = fma2(bc(z[5]), dz, fma2(bc(z[4]), dy, mul2(bc(z[3]), dx)));
				const f2 wy = fma2(bc(z[7]), dz, fma2(bc(z[6]), dy, mul2(bc(z[4]), dx)));
				const f2 wz = fma2(bc(x[6]), dz, fma2(bc(z[7]), dy, mul2(bc(z[5]), dx)));
				const f2 q = fma2(dz, wz, fma2(dy, wy, mul2(dx, wx)));
				float q0, q1;
				unpack2(q, q0, q1);
				if (q0 <= ua) acc += q0;
				if (q1 <= ua) acc += q1;
			}
			z[0] += 1e-3f;
		}
	}
	float s = acc;
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) s += x[k] + y[k] + z[k];
	if (s == 123.456f) out[0] = s;
}

template <typename F>
static int time_kernel(F launch, double *ms_out, cudaStream_t st)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	launch();	// warm-up
	cudaEventRecord(e0, st);
	launch();
	cudaEventRecord(e1, st);
	cudaError_t err = cudaEventSynchronize(e1);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*ms_out = ms;
	return (int)err;
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_peak_fma(int iters, double *tflops, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	float *out = nullptr;
	if (cudaMalloc(&out, 4) != cudaSuccess) return (int)cudaGetLastError();
	const int blocks = kSMs * 2, threads = 1024;
	double ms = 0.;
	int rc = time_kernel([&] { peak_fma_kernel<<<blocks, threads, 0, st>>>(iters, 1.0000001f, 1e-7f, out); }, &ms, st);
	cudaFree(out);
	if (tflops) *tflops = 2.0 * (double)blocks * threads * (double)iters * PK_CHAINS / (ms * 1e-3) / 1e12;
	return rc;
}

extern "C" int gsr_peak_mufu(int iters, double *tops, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	float *out = nullptr;
	if (cudaMalloc(&out, 4) != cudaSuccess) return (int)cudaGetLastError();
	const int blocks = kSMs * 2, threads = 1024;
	double ms = 0.;
	int rc = time_kernel([&] { peak_mufu_kernel<<<blocks, threads, 0, st>>>(iters, out); }, &ms, st);
	cudaFree(out);
	if (tops) *tops = (double)blocks * threads * (double)iters * PK_CHAINS / (ms * 1e-3) / 1e12;
	return rc;
}

// ops[0] = warp-level operations of the probed kind per second per SM-sub-partition cycle is derived by the caller;
// returns thread-level operations per second (FFMA/FADD/FMUL count 1 each; probe 5: candidate tests; probe 6: LDS.128)
extern "C" int gsr_pipe_probe(int which, int iters, double *ops_per_s, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	float *buf = nullptr;
	if (cudaMalloc(&buf, 4096 * 4) != cudaSuccess) return (int)cudaGetLastError();
	cudaMemsetAsync(buf, 0, 4096 * 4, st);
	const int blocks = kSMs * 8, threads = 256;
	double ms = 0.;
	int rc = 0;
	float ua = 1.0000001f;
	switch (which) {
	case 0: rc = time_kernel([&] { pipe_probe_kernel<0><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 1: rc = time_kernel([&] { pipe_probe_kernel<1><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 2: rc = time_kernel([&] { pipe_probe_kernel<2><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 3: rc = time_kernel([&] { pipe_probe_kernel<3><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 4: rc = time_kernel([&] { pipe_probe_kernel<4><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 5: rc = time_kernel([&] { pipe_probe_kernel<5><<<blocks, threads, 0, st>>>(iters, buf, -1.f, buf + 2048); }, &ms, st); break;
	case 6: rc = time_kernel([&] { pipe_probe_kernel<6><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 7: rc = time_kernel([&] { pipe_probe_kernel<7><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 8: rc = time_kernel([&] { pipe_probe_kernel<8><<<blocks, threads, 0, st>>>(iters, buf, ua, buf + 2048); }, &ms, st); break;
	case 9: rc = time_kernel([&] { pipe_probe_kernel<9><<<blocks, threads, 0, st>>>(iters, buf, -1.f, buf + 2048); }, &ms, st); break;
	default: cudaFree(buf); return GSR_EINVAL;
	}
	cudaFree(buf);
	const double per_iter = (which == 5 || which == 9) ? 4. : (which == 6 ? 3. : (double)PK_CHAINS);	// probes 7, 8 count scalar FMAs (2 per FFMA2)
	if (ops_per_s) *ops_per_s = (double)blocks * threads * (double)iters * per_iter / (ms * 1e-3);
	return rc;
}

extern "C" uint64_t gsr_launch_count(void) { return g_launches.load(); }

extern "C" const char *gsr_version(void) { return "gsr_b200 0.2.0 (sm_100a)"; }
