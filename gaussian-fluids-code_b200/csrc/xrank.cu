// xrank.cu — the one exchange step of the sample-sharded optimisation, fused into a single kernel over
// NVLink peer memory: every rank sums, in rank order, the per-Gaussian gradient accumulators (and loss partial sums) that
// all ranks' gather kernels left in their symmetric-memory buffers.
//
// Instead of a library all-reduce (NCCL: ~20 us of launch + protocol latency per iteration at this 150 KB payload, on an
// iteration that is ~75 us long), the kernel (1) publishes "my buffer of epoch e is complete" with a system-scope release
// store into every peer's signal pad, (2) waits until all peers have published epoch e, (3) reads every peer's buffer
// directly through NVLink (P2P loads, NVSwitch gives each pair full bandwidth) and adds them in rank order — so all ranks
// compute bit-identical sums and the replicas never diverge — and writes the result where the fused step reads it.
// Buffers alternate by iteration parity (the caller passes the parity's pointers), so a rank may start the next
// iteration's gather while a slower peer is still summing this one.  The wait is bounded: on a timeout the kernel raises an
// error flag instead of hanging the GPU.
#include "common.cuh"
#include <stdlib.h>

namespace gsr {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4 *p)
{
	float4 v;
	asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
	return v;
}

constexpr int XR_THREADS = 256;
constexpr int XR_MAX_WORLD = 16;
constexpr unsigned XR_SPIN_LIMIT = 1u << 23;	// default polls of the local signal pad (a few seconds: first-use module loads and graph
						// captures on a peer are legitimate skew) before giving up; once the flag is up nobody waits again

struct XrankArgs {
	const float4 *peer[XR_MAX_WORLD];	// every rank's buffer of this parity (own rank included), n4 float4 each
	uint32_t *sig[XR_MAX_WORLD];		// every rank's signal pad: slot [sender rank] holds the sender's last complete epoch
	int rank, world;
	unsigned spin_limit;
};

template <bool CONC>
__global__ void __launch_bounds__(XR_THREADS) xrank_sum_kernel(XrankArgs a, size_t n4, const float *__restrict__ iter_dev, const int32_t *__restrict__ base_dev,
							      float4 *__restrict__ out, int32_t *__restrict__ err)
{
	const uint32_t epoch = (uint32_t)(*base_dev) + (uint32_t)(*iter_dev) + 1u;
	if (blockIdx.x == 0 && threadIdx.x < a.world) {
		__threadfence_system();	// the gather kernels of this stream finished before this launch: make their stores visible system-wide
		st_release_sys(a.sig[threadIdx.x] + a.rank, epoch);
	}
	if (threadIdx.x < a.world && *((volatile int32_t *)err) == 0) {
		const uint32_t *mine = a.sig[a.rank] + threadIdx.x;
		unsigned spins = 0;
		while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
			if (++spins > a.spin_limit) {
				atomicExch(err, 1);
				break;
			}
		}
	}
	__syncthreads();
	// a peer that never published (or an earlier timeout): its buffer is not this epoch's — hand the step ZERO sample gradients
	// instead of a sum over stale data (the step then only applies the regularisers; the caller sees the sticky flag at its next
	// check and raises), so the parameters stay finite and the replicas do not silently diverge on garbage
	const bool bad = *((volatile int32_t *)err) != 0;
	for (size_t i = (size_t)blockIdx.x * XR_THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * XR_THREADS) {
		if (bad) {
			out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
			continue;
		}
		float4 s;
		if (CONC) {
			// every peer's value is requested before the first is used (pays at 8 ranks, costs at 2: see the launch code)
			float4 v[XR_MAX_WORLD];
#pragma unroll
			for (int r = 0; r < XR_MAX_WORLD; r++)
				if (r < a.world) v[r] = ld_peer4(a.peer[r] + i);
			s = v[0];
#pragma unroll
			for (int r = 1; r < XR_MAX_WORLD; r++)
				if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
		} else {
			s = ld_peer4(a.peer[0] + i);
			for (int r = 1; r < a.world; r++) {
				const float4 v = ld_peer4(a.peer[r] + i);
				s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
			}
		}
		out[i] = s;
	}
}

}  // namespace gsr

using namespace gsr;

extern "C" int gsr_xrank_sum(const void *const *peer_bufs, void *const *peer_signal_pads, int rank, int world, int64_t n_floats, const float *iteration_dev,
			     const int32_t *epoch_base_dev, float *out, int32_t *err_flag, void *stream)
{
	if (!peer_bufs || !peer_signal_pads || world < 1 || world > XR_MAX_WORLD || rank < 0 || rank >= world || n_floats < 0 || (n_floats & 3) || !iteration_dev ||
	    !epoch_base_dev || !out || !err_flag)
		return GSR_EINVAL;
	if (n_floats == 0) return GSR_OK;
	XrankArgs a;
	for (int r = 0; r < world; r++) {
		if (!peer_bufs[r] || !peer_signal_pads[r] || ((uintptr_t)peer_bufs[r] & 15)) return GSR_EINVAL;
		a.peer[r] = (const float4 *)peer_bufs[r];
		a.sig[r] = (uint32_t *)peer_signal_pads[r];
	}
	a.rank = rank;
	a.world = world;
	static const unsigned spin = []() {
		const char *e = getenv("GSR_XRANK_SPIN_LIMIT");	// polls before a missing peer is declared lost
		return e ? (unsigned)strtoul(e, nullptr, 10) : XR_SPIN_LIMIT;
	}();
	a.spin_limit = spin;
	const size_t n4 = (size_t)n_floats / 4;
	int blocks = (int)((n4 + XR_THREADS - 1) / XR_THREADS);
	if (blocks > kSMs) blocks = kSMs;	// all CTAs resident: every one of them polls the signal pad
	g_launches += 1;
	// measured (S1, ms per step): 2 ranks 40.8 with the {load, add} loop / 43.1 with all loads first; 8 ranks 44.3 / 43.7
	static const int conc_env = getenv("GSR_XRANK_CONC") ? atoi(getenv("GSR_XRANK_CONC")) : -1;
	const bool conc = conc_env >= 0 ? conc_env != 0 : world >= 8;
	if (conc) xrank_sum_kernel<true><<<blocks, XR_THREADS, 0, (cudaStream_t)stream>>>(a, n4, iteration_dev, epoch_base_dev, (float4 *)out, err_flag);
	else xrank_sum_kernel<false><<<blocks, XR_THREADS, 0, (cudaStream_t)stream>>>(a, n4, iteration_dev, epoch_base_dev, (float4 *)out, err_flag);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
