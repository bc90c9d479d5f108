// Probe: what does programmatic dependent launch (griddepcontrol) save per dependent edge inside a captured CUDA graph,
// and does stream capture accept it (a) on one stream, (b) with fork / join edges from other streams beside the programmatic one,
// (c) on a cluster launch?   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/pdl_probe tools/pdl_probe.cu && /tmp/pdl_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NB = 8, NT = 128, NEL = NB * NT;

// link k reads what link k-1 wrote (buffer k & 1) and writes buffer (k + 1) & 1; it also READS the line it is going to hand on
// two links later, so that a stale L1 line would be seen.  NC: the read goes through ld.global.nc (__ldg)
template <bool PDL, bool NC = false>
__global__ void link_kernel(float *buf2, int expect, int work, int *err)
{
	if (PDL) {
		asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
		asm volatile("griddepcontrol.wait;" ::: "memory");
	}
	const int i = blockIdx.x * NT + threadIdx.x;
	const float *src = buf2 + (expect & 1) * NEL;
	float *dst = buf2 + ((expect + 1) & 1) * NEL;
	const int j = (i + NT) % NEL;	// written by ANOTHER CTA of the previous link
	const float a = NC ? __ldg(src + j) : src[j];
	const float old = NC ? __ldg(dst + j) : dst[j];	// pulls the line the NEXT link reads into this SM's L1 (value expect - 1 or 0)
	if (a != (float)expect || old > (float)(expect + 1)) atomicAdd(err, 1);	// (old races with this link's own writers: k - 1 or k + 1)
	float x = a;
	for (int k = 0; k < work; k++) x = fmaf(x, 1.0000001f, 1e-9f);
	dst[i] = (x > -1.f) ? (float)(expect + 1) : x;
}

template <bool PDL, bool NC = false>
static void launch_link(cudaStream_t st, float *buf, int expect, int work, int *err, int cluster)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(NB);
	cfg.blockDim = dim3(NT);
	cfg.stream = st;
	cudaLaunchAttribute at[2];
	int na = 0;
	if (PDL) {
		at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[na].val.programmaticStreamSerializationAllowed = 1;
		na++;
	}
	if (cluster) {
		at[na].id = cudaLaunchAttributeClusterDimension;
		at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
		na++;
	}
	cfg.attrs = at;
	cfg.numAttrs = na;
	CK(cudaLaunchKernelEx(&cfg, link_kernel<PDL, NC>, buf, expect, work, err));
}

// one stream, `links` dependent kernels
template <bool PDL, bool NC = false>
static float run_chain(int links, int work, int cluster, int *errs_out)
{
	float *buf; int *err;
	CK(cudaMalloc(&buf, 2 * NEL * sizeof(float))); CK(cudaMalloc(&err, sizeof(int)));
	cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	cudaGraph_t g; cudaGraphExec_t ge;
	CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
	CK(cudaMemsetAsync(buf, 0, 2 * NEL * sizeof(float), st));
	for (int k = 0; k < links; k++) launch_link<PDL, NC>(st, buf, k, work, err, cluster);
	CK(cudaStreamEndCapture(st, &g));
	CK(cudaGraphInstantiate(&ge, g, 0));
	CK(cudaMemset(err, 0, sizeof(int)));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int r = 0; r < 20; r++) CK(cudaGraphLaunch(ge, st));
	CK(cudaEventRecord(e0, st));
	const int reps = 200;
	for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(ge, st));
	CK(cudaEventRecord(e1, st));
	CK(cudaStreamSynchronize(st));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	CK(cudaMemcpy(errs_out, err, sizeof(int), cudaMemcpyDeviceToHost));
	cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaFree(buf); cudaFree(err); cudaStreamDestroy(st);
	return ms * 1000.f / reps / links;
}

// the iteration's shape: main: A -> B -> [join side] -> C -> (next); side (forked after C of the previous round): D -> E
template <bool PDL>
static float run_forkjoin(int rounds, int work, int *errs_out, bool *ok)
{
	float *bm, *bs; int *err;
	CK(cudaMalloc(&bm, 2 * NEL * sizeof(float))); CK(cudaMalloc(&bs, 2 * NEL * sizeof(float))); CK(cudaMalloc(&err, sizeof(int)));
	cudaStream_t m, s; CK(cudaStreamCreateWithFlags(&m, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	std::vector<cudaEvent_t> evs;
	auto ev = [&]() { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evs.push_back(e); return e; };
	cudaGraph_t g; cudaGraphExec_t ge;
	*ok = true;
	CK(cudaStreamBeginCapture(m, cudaStreamCaptureModeThreadLocal));
	CK(cudaMemsetAsync(bm, 0, 2 * NEL * sizeof(float), m));
	CK(cudaMemsetAsync(bs, 0, 2 * NEL * sizeof(float), m));
	int km = 0, ks = 0;
	for (int r = 0; r < rounds; r++) {
		cudaEvent_t f = ev(); CK(cudaEventRecord(f, m)); CK(cudaStreamWaitEvent(s, f, 0));
		launch_link<PDL>(s, bs, ks++, work, err, 0);	// D (first kernel of the side stream after a cross-stream wait)
		launch_link<PDL>(s, bs, ks++, work, err, 0);	// E
		cudaEvent_t j = ev(); CK(cudaEventRecord(j, s));
		launch_link<PDL>(m, bm, km++, work, err, 0);	// A
		launch_link<PDL>(m, bm, km++, work, err, 0);	// B
		CK(cudaStreamWaitEvent(m, j, 0));
		launch_link<PDL>(m, bm, km++, work, err, 0);	// C: programmatic edge from B beside a full edge from E
	}
	cudaError_t e = cudaStreamEndCapture(m, &g);
	if (e != cudaSuccess) { printf("  capture failed: %s\n", cudaGetErrorString(e)); *ok = false; cudaGetLastError(); return 0.f; }
	e = cudaGraphInstantiate(&ge, g, 0);
	if (e != cudaSuccess) { printf("  instantiate failed: %s\n", cudaGetErrorString(e)); *ok = false; cudaGetLastError(); return 0.f; }
	CK(cudaMemset(err, 0, sizeof(int)));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int r = 0; r < 20; r++) CK(cudaGraphLaunch(ge, m));
	CK(cudaEventRecord(e0, m));
	const int reps = 200;
	for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(ge, m));
	CK(cudaEventRecord(e1, m));
	CK(cudaStreamSynchronize(m));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	CK(cudaMemcpy(errs_out, err, sizeof(int), cudaMemcpyDeviceToHost));
	return ms * 1000.f / reps / rounds;
}

int main()
{
	int errs; bool ok;
	for (int work : {0, 2000, 6000}) {
		const float a = run_chain<false>(60, work, 0, &errs); const int ea = errs;
		const float b = run_chain<true>(60, work, 0, &errs);
		printf("chain   work %5d: plain %.2f us/link (err %d)   pdl %.2f us/link (err %d)\n", work, a, ea, b, errs);
	}
	for (int work : {0, 500, 2000}) {
		const float a = run_chain<false, true>(60, work, 0, &errs); const int ea = errs;
		const float b = run_chain<true, true>(60, work, 0, &errs);
		printf("chain nc work %5d: plain %.2f us/link (err %d)   pdl %.2f us/link (err %d)\n", work, a, ea, b, errs);
	}
	{
		const float a = run_chain<false>(60, 2000, 8, &errs); const int ea = errs;
		const float b = run_chain<true>(60, 2000, 8, &errs);
		printf("cluster work  2000: plain %.2f us/link (err %d)   pdl %.2f us/link (err %d)\n", a, ea, b, errs);
	}
	for (int work : {2000, 6000}) {
		const float a = run_forkjoin<false>(20, work, &errs, &ok); const int ea = errs;
		const float b = run_forkjoin<true>(20, work, &errs, &ok);
		printf("forkjoin work %5d: plain %.2f us/round (err %d)   pdl %.2f us/round (err %d, ok %d)\n", work, a, ea, b, errs, (int)ok);
	}
	return 0;
}
