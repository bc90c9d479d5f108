// misc.cu — version string and the FP32-FMA / MUFU.EX2 peak micro-benchmarks that provide the roofline
// denominators for the evaluation kernels ("the builder must measure FMA and ex2.approx peaks").
#include "common.cuh"

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

extern "C" uint64_t gsr_launch_count(void) { return g_launches.load(); }

extern "C" const char *gsr_version(void) { return "gsr_b200 0.1.0 (sm_100a)"; }
