// backward.cu — analytic back-propagation of the six losses to every Gaussian.
//
// The reference (loop 2 of get_losses_ti, 3D/GSR.py:299-540; 2D/GSR.py:282-339, :396-476) walks samples,
// recomputes ~2 kflop of chain-rule matrices per accepted pair and issues 13-39 global float atomics per pair.
// Here the same gradient is produced in three atomics-free stages:
//   1. adjoint kernel   (per sample)   a = dL/du(x_j) in R^D,  A = dL/d grad u(x_j) in R^{DxD}, per accumulator
//                                      set (direct / vorticity+helicity / divergence), from the forward totals;
//   2. gather kernel    (per Gaussian) thread i walks the SAMPLES binned in the 27 (9) cells around its own
//                                      cell and sums, in registers, the compact per-pair terms
//                                          dL/dv  = a (g - tau) - g A w                     w = Sigma^-1 d
//                                          dL/dmu = g (c w + Sigma^-1 t)                    t = v^T A,  c = a.v - t.w
//                                          dL/dSigma^-1 = -g/2 (c d d^T + t d^T + d t^T)
//                                      so every Gaussian's sum is owned by one thread: deterministic, no atomics;
//   3. epilogue kernel  (per Gaussian) chain rule Sigma^-1 -> (scalings, rotation) and += into the caller's buffers.
// The stage-2 output (NSETS x N x 12|7 floats, original id order) is what a multi-GPU run all-reduces.
#include "eval.cuh"
#include "chain.cuh"
#include <math.h>

namespace gsr {

__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

struct LossW {
	float val, bnd, grad, vor, hel, div;	// already divided by their normalisers
	float tscale;				// 0 when the reference leaves d_grad_gaussian_* at zero (3D/GSR.py:385), else 1
};

struct AdjIn {
	const float *x, *val, *grad, *ref_val, *normals, *normal_ref, *ref_grad, *ref_vor, *ref_hel;
};

// record layout per sorted sample (float4 units):
//   [0]               {x, y, z|0, kappa}
//   VOR (3D):  [1]    {a.x, a.y, a.z, m.x}   [2] {m.y, m.z, 0, 0}          (A = sum_k m_k E_k, antisymmetric)
//   VOR (2D):  [1]    {s, 0, 0, 0}                                          (A = s [[0,-1],[1,0]])
//   DIRECT(3D): +3    {a.x, a.y, a.z, A00} {A01, A02, A10, A11} {A12, A20, A21, A22}
//   DIRECT(2D): +2    {a.x, a.y, A00, A01} {A10, A11, 0, 0}
template <int D> struct Rec {
	static constexpr int VOR = (D == 3) ? 2 : 1;
	static constexpr int DIR = (D == 3) ? 3 : 2;
};

constexpr int ADJ_THREADS = 128;

// Per-sample losses of one sample (the quantities advance.py forms with torch from val/grad:
// 3D/advance.py:228-235, :253; 2D/advance.py:236-237) — slots as documented in gsr_b200.h.
template <int D>
__device__ __forceinline__ void sample_losses(const AdjIn &in, size_t j, float L[6])
{
	float u[D], G[D * D];
#pragma unroll
	for (int k = 0; k < D; k++) u[k] = in.val ? in.val[D * j + k] : 0.f;
#pragma unroll
	for (int k = 0; k < D * D; k++) G[k] = in.grad ? in.grad[D * D * j + k] : 0.f;
	if (D == 3) {
		const float om[3] = {G[7] - G[5], G[2] - G[6], G[3] - G[1]};
		float lv = 0.f;
#pragma unroll
		for (int k = 0; k < 3; k++) lv += fabsf(om[k] - (in.ref_vor ? in.ref_vor[3 * j + k] : 0.f));
		L[0] = lv * (1.f / 3.f);
		L[1] = fabsf(u[0] * om[0] + u[1] * om[1] + u[2] * om[2] - (in.ref_hel ? in.ref_hel[j] : 0.f));
		const float dv = G[0] + G[4] + G[8];
		L[2] = dv * dv;
	} else {
		L[0] = fabsf((G[2] - G[1]) - (in.ref_vor ? in.ref_vor[j] : 0.f));
		L[1] = 0.f;
		const float dv = G[0] + G[3];
		L[2] = dv * dv;
	}
	float un = 0.f, lval = 0.f, lgrad = 0.f;
#pragma unroll
	for (int k = 0; k < D; k++) {
		un += u[k] * (in.normals ? in.normals[D * j + k] : 0.f);
		lval += fabsf(u[k] - (in.ref_val ? in.ref_val[D * j + k] : 0.f));
	}
#pragma unroll
	for (int k = 0; k < D * D; k++) lgrad += fabsf(G[k] - (in.ref_grad ? in.ref_grad[D * D * j + k] : 0.f));
	L[3] = fabsf(D == 3 ? un : un - (in.normal_ref ? in.normal_ref[j] : 0.f));
	L[4] = lval * (1.f / D);
	L[5] = lgrad * (1.f / (D * D));
}

// block partial sums of the sample losses -> partials[blockIdx][8]  (deterministic: fixed tree per block)
template <int D>
__global__ void __launch_bounds__(ADJ_THREADS) loss_partials_kernel(AdjIn in, int Q, float *__restrict__ partials)
{
	__shared__ float sm[ADJ_THREADS / 32][6];
	int t = blockIdx.x * ADJ_THREADS + threadIdx.x;
	float L[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	if (t < Q) sample_losses<D>(in, (size_t)t, L);
#pragma unroll
	for (int k = 0; k < 6; k++) {
#pragma unroll
		for (int o = 16; o; o >>= 1) L[k] += __shfl_xor_sync(0xffffffffu, L[k], o);
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int k = 0; k < 6; k++) sm[threadIdx.x >> 5][k] = L[k];
	}
	__syncthreads();
	if (threadIdx.x < 8) {
		float s = 0.f;
		if (threadIdx.x < 6) {
#pragma unroll
			for (int w = 0; w < ADJ_THREADS / 32; w++) s += sm[w][threadIdx.x];
		}
		partials[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
	}
}

// sums[k] = sum over blocks of partials[b][k], one block, fixed order
__global__ void __launch_bounds__(256) loss_final_kernel(const float *__restrict__ partials, int nblocks, float *__restrict__ sums)
{
	__shared__ double sm[8][32];
	int k = threadIdx.x & 7, lane = threadIdx.x >> 3;	// 32 strided lanes per slot
	double s = 0.;
	for (int b = lane; b < nblocks; b += 32) s += (double)partials[(size_t)b * 8 + k];
	sm[k][lane] = s;
	__syncthreads();
	if (threadIdx.x < 8) {
		double tot = 0.;
		for (int l = 0; l < 32; l++) tot += sm[threadIdx.x][l];
		sums[threadIdx.x] = (float)tot;
	}
}

// DIR_MODE: 0 no direct set; 1 direct set without a gradient loss (A = 0: value / boundary losses only — the boundary pass of
// every project iteration) -> one float4 {a, 0}; 2 full direct set
template <int D, int DIR_MODE, bool HAS_VOR>
__global__ void __launch_bounds__(ADJ_THREADS) adjoint_kernel(AdjIn in, int Q, const int32_t *__restrict__ perm, LossW w, float4 *__restrict__ rec, AdjIn lin,
							      float *__restrict__ partials)
{
	constexpr bool HAS_DIR = DIR_MODE != 0;
	constexpr int STRIDE = 1 + (HAS_VOR ? Rec<D>::VOR : 0) + (DIR_MODE == 2 ? Rec<D>::DIR : (DIR_MODE == 1 ? 1 : 0));
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (partials) {	// the sample losses of this block's samples (what loss_partials_kernel computes), fused into this pass
		__shared__ float sm[ADJ_THREADS / 32][6];
		float L[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
		if (t < Q) sample_losses<D>(lin, (size_t)(perm ? perm[t] : t), L);
#pragma unroll
		for (int k = 0; k < 6; k++) {
#pragma unroll
			for (int o = 16; o; o >>= 1) L[k] += __shfl_xor_sync(0xffffffffu, L[k], o);
		}
		if ((threadIdx.x & 31) == 0) {
#pragma unroll
			for (int k = 0; k < 6; k++) sm[threadIdx.x >> 5][k] = L[k];
		}
		__syncthreads();
		if (threadIdx.x < 8) {
			float s = 0.f;
			if (threadIdx.x < 6) {
#pragma unroll
				for (int wv = 0; wv < ADJ_THREADS / 32; wv++) s += sm[wv][threadIdx.x];
			}
			partials[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
		}
	}
	if (t >= Q) return;
	size_t j = perm ? perm[t] : t;
	float4 *r = rec + (size_t)STRIDE * t;
	float u[D], G[D * D];
#pragma unroll
	for (int k = 0; k < D; k++) u[k] = in.val ? in.val[D * j + k] : 0.f;
#pragma unroll
	for (int k = 0; k < D * D; k++) G[k] = in.grad ? in.grad[D * D * j + k] : 0.f;
	float div = (D == 3) ? (G[0] + G[4] + G[8]) : (G[0] + G[3]);
	float kappa = w.div * 2.f * div;
	r[0] = make_float4(in.x[D * j], in.x[D * j + 1], D == 3 ? in.x[D * j + 2] : 0.f, kappa);
	int o = 1;
	if (HAS_VOR) {
		if (D == 3) {
			float om[3] = {G[7] - G[5], G[2] - G[6], G[3] - G[1]};
			float m[3], a[3];
			float hel = u[0] * om[0] + u[1] * om[1] + u[2] * om[2];
			float sh = w.hel * sgnf(hel - (in.ref_hel ? in.ref_hel[j] : 0.f));
#pragma unroll
			for (int k = 0; k < 3; k++) {
				m[k] = w.vor * sgnf(om[k] - (in.ref_vor ? in.ref_vor[3 * j + k] : 0.f)) + sh * u[k];
				a[k] = sh * om[k];
			}
			r[o] = make_float4(a[0], a[1], a[2], m[0]);
			r[o + 1] = make_float4(m[1], m[2], 0.f, 0.f);
		} else {
			float s = w.vor * sgnf((G[2] - G[1]) - (in.ref_vor ? in.ref_vor[j] : 0.f));
			r[o] = make_float4(s, 0.f, 0.f, 0.f);
		}
		o += Rec<D>::VOR;
	}
	if (HAS_DIR) {
		float a[D], A[D * D];
		float un = 0.f;
#pragma unroll
		for (int k = 0; k < D; k++) un += u[k] * (in.normals ? in.normals[D * j + k] : 0.f);
		float sb = w.bnd * sgnf(D == 3 ? un : un - (in.normal_ref ? in.normal_ref[j] : 0.f));
#pragma unroll
		for (int k = 0; k < D; k++)
			a[k] = w.val * sgnf(u[k] - (in.ref_val ? in.ref_val[D * j + k] : 0.f)) + sb * (in.normals ? in.normals[D * j + k] : 0.f);
#pragma unroll
		for (int k = 0; k < D * D; k++) A[k] = w.grad * sgnf(G[k] - (in.ref_grad ? in.ref_grad[D * D * j + k] : 0.f));
		if (DIR_MODE == 1) {
			r[o] = make_float4(a[0], a[1], D == 3 ? a[D - 1] : 0.f, 0.f);
		} else if (D == 3) {
			r[o] = make_float4(a[0], a[1], a[2], A[0]);
			r[o + 1] = make_float4(A[1], A[2], A[3], A[4]);
			r[o + 2] = make_float4(A[5], A[6], A[7], A[8]);
		} else {
			r[o] = make_float4(a[0], a[1], A[0], A[1]);
			r[o + 1] = make_float4(A[2], A[3], 0.f, 0.f);
		}
	}
}

// ---- 3D gather ------------------------------------------------------------------------------------

struct Acc12 {
	float dv[3], dm[3], G[6];	// G: 00 01 02 11 12 22
	__device__ __forceinline__ void zero() {
#pragma unroll
		for (int k = 0; k < 3; k++) dv[k] = dm[k] = 0.f;
#pragma unroll
		for (int k = 0; k < 6; k++) G[k] = 0.f;
	}
};

// one accepted pair, generic form: a (3), t = v^T A (3), Aw = A w (3)
__device__ __forceinline__ void pair_accum3(Acc12 &acc, const float a[3], const float t_in[3], const float Aw[3], float tscale,
					    const float v[3], const float d[3], const float w[3], const float Am[6], const float dd[6], float g, float gm)
{
	const float s = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
	const float t[3] = {t_in[0] * tscale, t_in[1] * tscale, t_in[2] * tscale};
	const float c = s - (t[0] * w[0] + t[1] * w[1] + t[2] * w[2]);
#pragma unroll
	for (int k = 0; k < 3; k++) acc.dv[k] += a[k] * gm - g * Aw[k];
	const float St0 = Am[0] * t[0] + Am[1] * t[1] + Am[2] * t[2];
	const float St1 = Am[1] * t[0] + Am[3] * t[1] + Am[4] * t[2];
	const float St2 = Am[2] * t[0] + Am[4] * t[1] + Am[5] * t[2];
	acc.dm[0] += g * (c * w[0] + St0);
	acc.dm[1] += g * (c * w[1] + St1);
	acc.dm[2] += g * (c * w[2] + St2);
	const float hg = -.5f * g, hc = hg * c;
	acc.G[0] += hc * dd[0] + hg * (2.f * t[0] * d[0]);
	acc.G[1] += hc * dd[1] + hg * (t[0] * d[1] + d[0] * t[1]);
	acc.G[2] += hc * dd[2] + hg * (t[0] * d[2] + d[0] * t[2]);
	acc.G[3] += hc * dd[3] + hg * (2.f * t[1] * d[1]);
	acc.G[4] += hc * dd[4] + hg * (t[1] * d[2] + d[1] * t[2]);
	acc.G[5] += hc * dd[5] + hg * (2.f * t[2] * d[2]);
}

constexpr int GA_THREADS = 128;
int g_gather_cta_max_n = 4096;	// GSR_TUNE_GATHER_CTA_MAX_N

template <int LPG>
__device__ __forceinline__ void acc12_reduce(Acc12 &a)
{
#pragma unroll
	for (int k = 0; k < 3; k++) { a.dv[k] = lane_sum<LPG>(a.dv[k]); a.dm[k] = lane_sum<LPG>(a.dm[k]); }
#pragma unroll
	for (int k = 0; k < 6; k++) a.G[k] = lane_sum<LPG>(a.G[k]);
}

// LPG lanes cooperate on one Gaussian (1: throughput shape for large N; 8 / 32: latency shape for small N or Q >> N)
// a pair of the direct set when A = 0 (t = 0, c = a.v): ~40 flop instead of 115
__device__ __forceinline__ void pair_accum3_value(Acc12 &acc, const float a[3], const float v[3], const float w[3], const float dd[6], float g, float gm)
{
	const float c = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		acc.dv[k] += a[k] * gm;
		acc.dm[k] += g * (c * w[k]);
	}
	const float hc = -.5f * g * c;
#pragma unroll
	for (int k = 0; k < 6; k++) acc.G[k] += hc * dd[k];
}

template <int DIR_MODE, bool HAS_VOR, bool HAS_DIV, int LPG>
__global__ void __launch_bounds__(GA_THREADS) gather3d_kernel(EvalParams P, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_id,
							      const float4 *__restrict__ packed, int N, const int32_t *__restrict__ scs /* sample_cell_start */,
							      const float4 *__restrict__ rec, const int32_t *__restrict__ stop_gradient, float tscale,
							      float *__restrict__ acc_out)
{
	constexpr bool HAS_DIR = DIR_MODE != 0;
	constexpr int STRIDE = 1 + (HAS_VOR ? 2 : 0) + (DIR_MODE == 2 ? 3 : (DIR_MODE == 1 ? 1 : 0));
	const int gt = blockIdx.x * GA_THREADS + threadIdx.x;
	const int tq = gt / LPG, lane = gt % LPG;
	if (LPG == 1 && tq >= N) return;
	const bool valid = tq < N;
	const int t = valid ? tq : N - 1;
	const Grid &g = P.g;
	const int id = sorted_id[t];
	const int n_in = cell_start[g.ncell];
	Acc12 aD, aV, aX;
	aD.zero(); aV.zero(); aX.zero();
	const bool active = valid && t < n_in && !(stop_gradient && stop_gradient[id]);
	if (active) {
		const float4 p0 = packed[3 * (size_t)t], p1 = packed[3 * (size_t)t + 1], p2 = packed[3 * (size_t)t + 2];
		const float v[3] = {p0.w, p1.w, p2.w};
		const float Am[6] = {p1.x, p1.y, p1.z, p2.x, p2.y, p2.z};
		// Sigma^-1 v: constant per Gaussian, used by the divergence set
		const float Av[3] = {Am[0] * v[0] + Am[1] * v[1] + Am[2] * v[2], Am[1] * v[0] + Am[3] * v[1] + Am[4] * v[2], Am[2] * v[0] + Am[4] * v[1] + Am[5] * v[2]};
		const float gs = grid_gs(g);
		const int cx = cell_coord(p0.x, g.lo[0], gs), cy = cell_coord(p0.y, g.lo[1], gs), cz = cell_coord(p0.z, g.lo[2], gs);
		const float tau = g.tau, q_thr = P.q_thr;
		// one accepted-or-not visit of sorted sample k
		constexpr bool R1_EARLY = DIR_MODE == 1 && !HAS_VOR;	// tiny record: its second float4 is fetched with the first, not after the test
		auto visit = [&](int k, const float4 r0, const float4 r1v) {
			const float d[3] = {r0.x - p0.x, r0.y - p0.y, r0.z - p0.z};
			const float w[3] = {Am[0] * d[0] + Am[1] * d[1] + Am[2] * d[2], Am[1] * d[0] + Am[3] * d[1] + Am[4] * d[2], Am[2] * d[0] + Am[4] * d[1] + Am[5] * d[2]};
			const float q = d[0] * w[0] + d[1] * w[1] + d[2] * w[2];
			if (q <= q_thr) {
				const float gg = ex2_approx(q * kNegHalfLog2e), gm = gg - tau;
				const float dd[6] = {d[0] * d[0], d[0] * d[1], d[0] * d[2], d[1] * d[1], d[1] * d[2], d[2] * d[2]};
				int o = 1;
				if (HAS_VOR) {
					const float4 r1 = __ldg(rec + (size_t)STRIDE * k + o), r2 = __ldg(rec + (size_t)STRIDE * k + o + 1);
					const float a[3] = {r1.x, r1.y, r1.z}, m[3] = {r1.w, r2.x, r2.y};
					const float tt[3] = {v[1] * m[2] - v[2] * m[1], v[2] * m[0] - v[0] * m[2], v[0] * m[1] - v[1] * m[0]};	// v x m
					const float Aw[3] = {m[1] * w[2] - m[2] * w[1], m[2] * w[0] - m[0] * w[2], m[0] * w[1] - m[1] * w[0]};	// m x w
					pair_accum3(aV, a, tt, Aw, tscale, v, d, w, Am, dd, gg, gm);
					o += 2;
				}
				if (HAS_DIV) {
					const float gk = gg * r0.w, vw = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
					const float hk = -.5f * gk;
#pragma unroll
					for (int q = 0; q < 3; q++) {
						aX.dv[q] -= gk * w[q];
						aX.dm[q] += gk * (Av[q] - vw * w[q]);
					}
					aX.G[0] += hk * (2.f * v[0] * d[0] - vw * dd[0]);
					aX.G[1] += hk * (v[0] * d[1] + d[0] * v[1] - vw * dd[1]);
					aX.G[2] += hk * (v[0] * d[2] + d[0] * v[2] - vw * dd[2]);
					aX.G[3] += hk * (2.f * v[1] * d[1] - vw * dd[3]);
					aX.G[4] += hk * (v[1] * d[2] + d[1] * v[2] - vw * dd[4]);
					aX.G[5] += hk * (2.f * v[2] * d[2] - vw * dd[5]);
				}
				if (DIR_MODE == 1) {
					const float4 r1 = HAS_VOR ? __ldg(rec + (size_t)STRIDE * k + o) : r1v;
					const float a[3] = {r1.x, r1.y, r1.z};
					pair_accum3_value(aD, a, v, w, dd, gg, gm);
				}
				if (DIR_MODE == 2) {
					const float4 r1 = __ldg(rec + (size_t)STRIDE * k + o), r2 = __ldg(rec + (size_t)STRIDE * k + o + 1), r3 = __ldg(rec + (size_t)STRIDE * k + o + 2);
					const float a[3] = {r1.x, r1.y, r1.z};
					const float A[9] = {r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
					const float tt[3] = {v[0] * A[0] + v[1] * A[3] + v[2] * A[6], v[0] * A[1] + v[1] * A[4] + v[2] * A[7], v[0] * A[2] + v[1] * A[5] + v[2] * A[8]};
					const float Aw[3] = {A[0] * w[0] + A[1] * w[1] + A[2] * w[2], A[3] * w[0] + A[4] * w[1] + A[5] * w[2], A[6] * w[0] + A[7] * w[1] + A[8] * w[2]};
					pair_accum3(aD, a, tt, Aw, tscale, v, d, w, Am, dd, gg, gm);
				}
			}
		};
		if (LPG >= 32) {
			// a warp (or a whole CTA of 4 warps, each doing the lookup for itself) on one Gaussian: the 9 sample runs are looked up by the lanes at once, concatenated by a scan
			// inside the lane group, and the lanes stride through the flat list (one round trip instead of nine dependent ones)
			int pre[9], off[9], total;
			flat_runs3<(LPG >= 32 ? 32 : 8)>(lane & 31, true, [&](int r, int &s, int &n) {
				const int base = ((cx + r / 3) * g.pdims[1] + (cy + r % 3)) * g.pdims[2] + cz;
				s = __ldg(scs + base);
				n = __ldg(scs + base + 3) - s;
			}, pre, off, total);
			// 4 visits per trip with all their record loads issued first: the kernel is bound by the latency of these loads
			for (int f0 = lane; f0 < total; f0 += 4 * LPG) {
				int kk[4];
				bool ok[4];
				float4 R0[4], R1[4];
#pragma unroll
				for (int b = 0; b < 4; b++) {
					const int f = f0 + b * LPG;
					ok[b] = f < total;
					const int ff = ok[b] ? f : f0;	// a valid index to load from when this slot is past the end
					int dlt = off[0];
#pragma unroll
					for (int r = 1; r < 9; r++) dlt = (ff >= pre[r]) ? off[r] : dlt;
					kk[b] = ff + dlt;
				}
#pragma unroll
				for (int b = 0; b < 4; b++) {
					R0[b] = __ldg(rec + (size_t)STRIDE * kk[b]);
					R1[b] = R1_EARLY ? __ldg(rec + (size_t)STRIDE * kk[b] + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
				}
#pragma unroll
				for (int b = 0; b < 4; b++)
					if (ok[b]) visit(kk[b], R0[b], R1[b]);
			}
		} else {
			for (int pi = cx; pi <= cx + 2; pi++) {
				for (int pj = cy; pj <= cy + 2; pj++) {
					const int base = (pi * g.pdims[1] + pj) * g.pdims[2] + cz;
					const int s = __ldg(scs + base), e = __ldg(scs + base + 3);
					for (int k = s + lane; k < e; k += LPG)
						visit(k, __ldg(rec + (size_t)STRIDE * k), R1_EARLY ? __ldg(rec + (size_t)STRIDE * k + 1) : make_float4(0.f, 0.f, 0.f, 0.f));
				}
			}
		}
	}
	const Acc12 *sets[3] = {&aD, &aV, &aX};
	const bool has[3] = {HAS_DIR, HAS_VOR, HAS_DIV};
	if (LPG == 128) {	// one CTA per Gaussian (many samples per Gaussian, few Gaussians): warp sums, then the 4 warps in order
		__shared__ float red[3][GA_THREADS / 32][12];
		if (HAS_DIR) acc12_reduce<32>(aD);
		if (HAS_VOR) acc12_reduce<32>(aV);
		if (HAS_DIV) acc12_reduce<32>(aX);
		if ((threadIdx.x & 31) == 0) {
#pragma unroll
			for (int s = 0; s < 3; s++) {
				if (!has[s]) continue;
				float *o = red[s][threadIdx.x >> 5];
#pragma unroll
				for (int k = 0; k < 3; k++) { o[k] = sets[s]->dv[k]; o[3 + k] = sets[s]->dm[k]; }
#pragma unroll
				for (int k = 0; k < 6; k++) o[6 + k] = sets[s]->G[k];
			}
		}
		__syncthreads();
		if (threadIdx.x < 36 && has[threadIdx.x / 12]) {
			const int s = threadIdx.x / 12, k = threadIdx.x % 12;
			float t = 0.f;
#pragma unroll
			for (int ww = 0; ww < GA_THREADS / 32; ww++) t += red[s][ww][k];
			acc_out[((size_t)s * N + id) * 12 + k] = t;
		}
		return;
	}
	if (LPG > 1) {
		if (HAS_DIR) acc12_reduce<LPG>(aD);
		if (HAS_VOR) acc12_reduce<LPG>(aV);
		if (HAS_DIV) acc12_reduce<LPG>(aX);
		if (!valid || lane != 0) return;
	}
	// original-id order, set-major
#pragma unroll
	for (int s = 0; s < 3; s++) {
		if (!has[s]) continue;
		float4 *o = reinterpret_cast<float4 *>(acc_out + ((size_t)s * N + id) * 12);
		o[0] = make_float4(sets[s]->dv[0], sets[s]->dv[1], sets[s]->dv[2], sets[s]->dm[0]);
		o[1] = make_float4(sets[s]->dm[1], sets[s]->dm[2], sets[s]->G[0], sets[s]->G[1]);
		o[2] = make_float4(sets[s]->G[2], sets[s]->G[3], sets[s]->G[4], sets[s]->G[5]);
	}
}

// ---- 2D gather ------------------------------------------------------------------------------------

struct Acc7 {
	float dv[2], dm[2], G[3];	// G: 00 01 11
	__device__ __forceinline__ void zero() { dv[0] = dv[1] = dm[0] = dm[1] = G[0] = G[1] = G[2] = 0.f; }
};

__device__ __forceinline__ void pair_accum2(Acc7 &acc, const float a[2], const float t[2], const float Aw[2],
					    const float v[2], const float d[2], const float w[2], const float Am[3], float g, float gm)
{
	const float s = a[0] * v[0] + a[1] * v[1];
	const float c = s - (t[0] * w[0] + t[1] * w[1]);
	acc.dv[0] += a[0] * gm - g * Aw[0];
	acc.dv[1] += a[1] * gm - g * Aw[1];
	acc.dm[0] += g * (c * w[0] + Am[0] * t[0] + Am[1] * t[1]);
	acc.dm[1] += g * (c * w[1] + Am[1] * t[0] + Am[2] * t[1]);
	const float hg = -.5f * g, hc = hg * c;
	acc.G[0] += hc * d[0] * d[0] + hg * (2.f * t[0] * d[0]);
	acc.G[1] += hc * d[0] * d[1] + hg * (t[0] * d[1] + d[0] * t[1]);
	acc.G[2] += hc * d[1] * d[1] + hg * (2.f * t[1] * d[1]);
}

template <int LPG>
__device__ __forceinline__ void acc7_reduce(Acc7 &a)
{
	a.dv[0] = lane_sum<LPG>(a.dv[0]); a.dv[1] = lane_sum<LPG>(a.dv[1]);
	a.dm[0] = lane_sum<LPG>(a.dm[0]); a.dm[1] = lane_sum<LPG>(a.dm[1]);
	a.G[0] = lane_sum<LPG>(a.G[0]); a.G[1] = lane_sum<LPG>(a.G[1]); a.G[2] = lane_sum<LPG>(a.G[2]);
}

template <int DIR_MODE, bool HAS_VOR, bool HAS_DIV, int LPG>
__global__ void __launch_bounds__(GA_THREADS) gather2d_kernel(EvalParams P, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_id,
							      const float4 *__restrict__ packed, int N, const int32_t *__restrict__ scs,
							      const float4 *__restrict__ rec, const int32_t *__restrict__ stop_gradient, float *__restrict__ acc_out)
{
	constexpr bool HAS_DIR = DIR_MODE != 0;
	constexpr int STRIDE = 1 + (HAS_VOR ? 1 : 0) + (DIR_MODE == 2 ? 2 : (DIR_MODE == 1 ? 1 : 0));
	const int gt = blockIdx.x * GA_THREADS + threadIdx.x;
	const int tq = gt / LPG, lane = gt % LPG;
	if (LPG == 1 && tq >= N) return;
	const bool valid = tq < N;
	const int t = valid ? tq : N - 1;
	const Grid &g = P.g;
	const int id = sorted_id[t];
	const int n_in = cell_start[g.ncell];
	Acc7 aD, aV, aX;
	aD.zero(); aV.zero(); aX.zero();
	const bool active = valid && t < n_in && !(stop_gradient && stop_gradient[id]);
	if (active) {
		const float4 p0 = packed[2 * (size_t)t], p1 = packed[2 * (size_t)t + 1];
		const float v[2] = {p0.z, p0.w};
		const float Am[3] = {p1.x, p1.y, p1.z};
		const float gs = grid_gs(g);
		const int cx = cell_coord(p0.x, g.lo[0], gs), cy = cell_coord(p0.y, g.lo[1], gs);
		const float tau = g.tau, q_thr = P.q_thr;
		for (int pi = cx; pi <= cx + 2; pi++) {
			const int base = pi * g.pdims[1] + cy;
			const int s = __ldg(scs + base), e = __ldg(scs + base + 3);
			for (int k = s + lane; k < e; k += LPG) {
				const float4 r0 = __ldg(rec + (size_t)STRIDE * k);
				const float d[2] = {r0.x - p0.x, r0.y - p0.y};
				const float w[2] = {Am[0] * d[0] + Am[1] * d[1], Am[1] * d[0] + Am[2] * d[1]};
				const float q = d[0] * w[0] + d[1] * w[1];
				if (q <= q_thr) {
					const float gg = ex2_approx(q * kNegHalfLog2e), gm = gg - tau;
					int o = 1;
					if (HAS_VOR) {
						const float s_ = __ldg(rec + (size_t)STRIDE * k + o).x;
						const float a[2] = {0.f, 0.f};
						const float tt[2] = {s_ * v[1], -s_ * v[0]};	// v^T (s [[0,-1],[1,0]])
						const float Aw[2] = {-s_ * w[1], s_ * w[0]};
						pair_accum2(aV, a, tt, Aw, v, d, w, Am, gg, gm);
						o += 1;
					}
					if (HAS_DIV) {
						const float kap = r0.w;
						const float a[2] = {0.f, 0.f};
						const float tt[2] = {kap * v[0], kap * v[1]};
						const float Aw[2] = {kap * w[0], kap * w[1]};
						pair_accum2(aX, a, tt, Aw, v, d, w, Am, gg, gm);
					}
					if (DIR_MODE == 1) {
						const float4 r1 = __ldg(rec + (size_t)STRIDE * k + o);
						const float a[2] = {r1.x, r1.y}, zero2[2] = {0.f, 0.f};
						pair_accum2(aD, a, zero2, zero2, v, d, w, Am, gg, gm);	// t = 0, A w = 0: the compiler folds the dead terms
					}
					if (DIR_MODE == 2) {
						const float4 r1 = __ldg(rec + (size_t)STRIDE * k + o), r2 = __ldg(rec + (size_t)STRIDE * k + o + 1);
						const float a[2] = {r1.x, r1.y};
						const float A[4] = {r1.z, r1.w, r2.x, r2.y};
						const float tt[2] = {v[0] * A[0] + v[1] * A[2], v[0] * A[1] + v[1] * A[3]};
						const float Aw[2] = {A[0] * w[0] + A[1] * w[1], A[2] * w[0] + A[3] * w[1]};
						pair_accum2(aD, a, tt, Aw, v, d, w, Am, gg, gm);
					}
				}
			}
		}
	}
	if (LPG > 1) {
		if (HAS_DIR) acc7_reduce<LPG>(aD);
		if (HAS_VOR) acc7_reduce<LPG>(aV);
		if (HAS_DIV) acc7_reduce<LPG>(aX);
		if (!valid || lane != 0) return;
	}
	const Acc7 *sets[3] = {&aD, &aV, &aX};
	const bool has[3] = {HAS_DIR, HAS_VOR, HAS_DIV};
#pragma unroll
	for (int s = 0; s < 3; s++) {
		if (!has[s]) continue;
		float *o = acc_out + ((size_t)s * N + id) * 7;
		o[0] = sets[s]->dv[0]; o[1] = sets[s]->dv[1]; o[2] = sets[s]->dm[0]; o[3] = sets[s]->dm[1];
		o[4] = sets[s]->G[0]; o[5] = sets[s]->G[1]; o[6] = sets[s]->G[2];
	}
}

struct OutPtrs {
	float *p[GSR_NSETS][4];
};

template <int D>
__global__ void epilogue_kernel(const float *__restrict__ scal, const float *__restrict__ rot, int N, const float *__restrict__ acc, int sets_mask, OutPtrs out)
{
	constexpr int AF = (D == 3) ? 12 : 7;
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	for (int s = 0; s < GSR_NSETS; s++) {
		if (!((sets_mask >> s) & 1)) continue;
		const float *a = acc + ((size_t)s * N + i) * AF;
		float *gp = out.p[s][0], *gs = out.p[s][1], *gr = out.p[s][2], *gv = out.p[s][3];
		if (D == 3) {
			float ds[3], dr[4];
			const float sc[3] = {scal[3 * (size_t)i], scal[3 * (size_t)i + 1], scal[3 * (size_t)i + 2]};
			const float r[4] = {rot[4 * (size_t)i], rot[4 * (size_t)i + 1], rot[4 * (size_t)i + 2], rot[4 * (size_t)i + 3]};
			chain3d(a + 6, sc, r, ds, dr);
			// sequential read-modify-write: buffers of different sets may alias (3D/GSR.py:564-579)
			for (int k = 0; k < 3; k++) {
				gv[3 * (size_t)i + k] += a[k];
				gp[3 * (size_t)i + k] += a[3 + k];
				gs[3 * (size_t)i + k] += ds[k];
			}
			for (int k = 0; k < 4; k++) gr[4 * (size_t)i + k] += dr[k];
		} else {
			float ds[2], dth;
			const float sc[2] = {scal[2 * (size_t)i], scal[2 * (size_t)i + 1]};
			chain2d(a + 4, sc, rot[i], ds, &dth);
			for (int k = 0; k < 2; k++) {
				gv[2 * (size_t)i + k] += a[k];
				gp[2 * (size_t)i + k] += a[2 + k];
				gs[2 * (size_t)i + k] += ds[k];
			}
			gr[i] += dth;
		}
	}
}

}  // namespace gsr

using namespace gsr;

static size_t align256b(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t gsr_backward_ws_bytes(const gsr_grid_desc *d, int64_t, int64_t Q)
{
	int D = d ? d->D : 3;
	int stride = 1 + (D == 3 ? 5 : 3);
	return align256b(sizeof(float4) * (size_t)stride * (size_t)(Q > 0 ? Q : 1));
}

template <int D>
static int launch_adjoint(int dir, bool vor, const AdjIn &in, int Q, const int32_t *perm, const LossW &w, float4 *rec, const AdjIn &lin, float *partials,
			  cudaStream_t st)
{
	int blocks = (Q + ADJ_THREADS - 1) / ADJ_THREADS;
	if (dir == 2 && vor) adjoint_kernel<D, 2, true><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	else if (dir == 2) adjoint_kernel<D, 2, false><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	else if (dir == 1 && vor) adjoint_kernel<D, 1, true><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	else if (dir == 1) adjoint_kernel<D, 1, false><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	else if (vor) adjoint_kernel<D, 0, true><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	else adjoint_kernel<D, 0, false><<<blocks, ADJ_THREADS, 0, st>>>(in, Q, perm, w, rec, lin, partials);
	GSR_CHECK_LAUNCH();
	return 0;
}

#define GATHER_CASE(KERNEL, A, B, C_, LN, ...) KERNEL<A, B, C_, LN><<<blocks, GA_THREADS, 0, st>>>(__VA_ARGS__)
#define GATHER_DISPATCH_L(KERNEL, LN, ...)                                          \
	do {                                                                        \
		int sel = dirmode * 4 + (vor ? 2 : 0) + (dv ? 1 : 0);              \
		switch (sel) {                                                      \
		case 1: GATHER_CASE(KERNEL, 0, false, true, LN, __VA_ARGS__); break;       \
		case 2: GATHER_CASE(KERNEL, 0, true, false, LN, __VA_ARGS__); break;       \
		case 3: GATHER_CASE(KERNEL, 0, true, true, LN, __VA_ARGS__); break;        \
		case 4: GATHER_CASE(KERNEL, 1, false, false, LN, __VA_ARGS__); break;      \
		case 5: GATHER_CASE(KERNEL, 1, false, true, LN, __VA_ARGS__); break;       \
		case 6: GATHER_CASE(KERNEL, 1, true, false, LN, __VA_ARGS__); break;       \
		case 7: GATHER_CASE(KERNEL, 1, true, true, LN, __VA_ARGS__); break;        \
		case 8: GATHER_CASE(KERNEL, 2, false, false, LN, __VA_ARGS__); break;      \
		case 9: GATHER_CASE(KERNEL, 2, false, true, LN, __VA_ARGS__); break;       \
		case 10: GATHER_CASE(KERNEL, 2, true, false, LN, __VA_ARGS__); break;      \
		case 11: GATHER_CASE(KERNEL, 2, true, true, LN, __VA_ARGS__); break;       \
		default: break;                                                     \
		}                                                                   \
	} while (0)
#define GATHER_DISPATCH(KERNEL, ...)                                                \
	do {                                                                        \
		if (lpg == 1) GATHER_DISPATCH_L(KERNEL, 1, __VA_ARGS__);            \
		else if (lpg == 8) GATHER_DISPATCH_L(KERNEL, 8, __VA_ARGS__);       \
		else GATHER_DISPATCH_L(KERNEL, 32, __VA_ARGS__);                    \
	} while (0)

extern "C" int gsr_backward_gather(const gsr_grid_desc *d, const int32_t *cell_start, const int32_t *sorted_id, const float *packed, int64_t N,
				   const float *x, int64_t Q, const int32_t *perm, const int32_t *sample_cell_start,
				   const float *val, const float *grad, const gsr_loss_cfg *cfg,
				   float *acc, int *sets_mask, void *ws, size_t ws_bytes, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || !cfg || !acc || !sets_mask || N < 0 || Q < 0 || !perm || !sample_cell_start || !cell_start || !sorted_id || !packed)
		return GSR_EINVAL;
	if (N >= ((int64_t)1 << 30) || Q >= ((int64_t)1 << 30)) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	const int D = g.D;
	bool dir, vor, dv;
	float tscale = 1.f;
	*sets_mask = 0;
	if (D == 3) {
		// 3D/GSR.py:299 — the whole loop is skipped unless one of these is non-zero (weight_hel is not in the test)
		if (cfg->w_val == 0.f && cfg->w_boundary == 0.f && cfg->w_grad == 0.f && cfg->w_vor == 0.f && cfg->w_div == 0.f) return GSR_OK;
		dir = cfg->w_val != 0.f || cfg->w_boundary != 0.f || cfg->w_grad != 0.f;
		vor = (cfg->w_vor + cfg->w_hel) != 0.f;	// :454
		dv = cfg->w_div != 0.f;			// :523
		tscale = (cfg->w_grad != 0.f || cfg->w_vor != 0.f || cfg->w_div != 0.f) ? 1.f : 0.f;	// :385
	} else {
		dir = cfg->w_val != 0.f || cfg->w_boundary != 0.f || cfg->w_grad != 0.f;
		vor = cfg->w_vor != 0.f;
		dv = cfg->w_div != 0.f;
		if (!dir && !vor && !dv) return GSR_OK;
	}
	// NULL references are read as zeros (the reference substitutes zero tensors, 3D/GSR.py:554-563)
	*sets_mask = (dir ? 1 : 0) | (vor ? 2 : 0) | (dv ? 4 : 0);
	if (N == 0) return GSR_OK;
	if (ws_bytes < gsr_backward_ws_bytes(d, N, Q)) return GSR_EWS;
	const double Qn = (double)(cfg->Q_norm > 0 ? cfg->Q_norm : Q);
	LossW w;
	if (D == 3) {
		w.val = (float)(cfg->w_val / (3.0 * Qn));
		w.bnd = (float)(cfg->w_boundary / Qn);
		w.grad = (float)(cfg->w_grad / (9.0 * Qn));
		w.vor = (float)(cfg->w_vor / (3.0 * Qn));
		w.hel = (float)(cfg->w_hel / Qn);
		w.div = (float)(cfg->w_div / Qn);
	} else {
		w.val = (float)(cfg->w_val / (2.0 * Qn));	// 2D/GSR.py:306  weight / (2 m)
		w.bnd = (float)(cfg->w_boundary / Qn);
		w.grad = (float)(cfg->w_grad / (4.0 * Qn));	// :424
		w.vor = (float)(cfg->w_vor / Qn);
		w.hel = 0.f;
		w.div = (float)(cfg->w_div / Qn);
	}
	w.tscale = tscale;
	AdjIn in = {x, val, grad, cfg->w_val != 0.f ? cfg->ref_val : nullptr, cfg->w_boundary != 0.f ? cfg->normals : nullptr,
		    cfg->w_boundary != 0.f ? cfg->normal_ref : nullptr, cfg->w_grad != 0.f ? cfg->ref_grad : nullptr,
		    cfg->w_vor != 0.f ? cfg->ref_vor : nullptr, cfg->w_hel != 0.f ? cfg->ref_hel : nullptr};
	float4 *rec = (float4 *)ws;
	const AdjIn lin = {x, val, grad, cfg->ref_val, cfg->normals, cfg->normal_ref, cfg->ref_grad, cfg->ref_vor, cfg->ref_hel};
	if (cfg->loss_partials && Q == 0) cudaMemsetAsync(cfg->loss_partials, 0, 8 * sizeof(float), st);
	EvalParams P = make_params(g);
	if (cfg->sample_grid_scale_dev) P.g.gs_dev = cfg->sample_grid_scale_dev;	// the gather uses the grid scale only to find a Gaussian's sample cells
	// lanes per Gaussian: by the amount of parallelism N offers, and more when there are many samples per Gaussian
	int lpg = pick_lanes(N);
	if (lpg == 1 && Q >= 8 * N) lpg = 8;
	if (D == 3 && N <= g_gather_cta_max_n && Q >= 4 * N) lpg = 128;	// few Gaussians, many samples each (the boundary batch): a CTA per Gaussian
	int blocks = (int)((N * lpg + GA_THREADS - 1) / GA_THREADS);
	g_launches += Q > 0 ? 2 : 1;
	const int dirmode = !dir ? 0 : (cfg->w_grad != 0.f ? 2 : 1);	// no gradient loss: A = 0 in the direct set
	if (Q > 0) {
		int rc = (D == 3) ? launch_adjoint<3>(dirmode, vor, in, (int)Q, perm, w, rec, lin, cfg->loss_partials, st)
				  : launch_adjoint<2>(dirmode, vor, in, (int)Q, perm, w, rec, lin, cfg->loss_partials, st);
		if (rc) return rc;
	}
	if (D == 3 && lpg == 128)
		GATHER_DISPATCH_L(gather3d_kernel, 128, P, cell_start, sorted_id, (const float4 *)packed, (int)N, sample_cell_start, rec, cfg->stop_gradient, tscale, acc);
	else if (D == 3)
		GATHER_DISPATCH(gather3d_kernel, P, cell_start, sorted_id, (const float4 *)packed, (int)N, sample_cell_start, rec, cfg->stop_gradient, tscale, acc);
	else
		GATHER_DISPATCH(gather2d_kernel, P, cell_start, sorted_id, (const float4 *)packed, (int)N, sample_cell_start, rec, cfg->stop_gradient, acc);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int64_t gsr_loss_blocks(int64_t Q) { return Q > 0 ? (Q + ADJ_THREADS - 1) / ADJ_THREADS : 1; }

extern "C" int gsr_sample_losses(const gsr_grid_desc *d, int64_t Q, const float *val, const float *grad, const gsr_loss_cfg *cfg,
				 float *sums, void *ws, size_t ws_bytes, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || !cfg || !sums || Q < 0 || Q >= ((int64_t)1 << 30)) return GSR_EINVAL;
	int lb = (int)gsr_loss_blocks(Q);
	if (ws_bytes < sizeof(float) * 8 * (size_t)lb) return GSR_EWS;
	cudaStream_t st = (cudaStream_t)stream;
	float *partials = (float *)ws;
	AdjIn lin = {nullptr, val, grad, cfg->ref_val, cfg->normals, cfg->normal_ref, cfg->ref_grad, cfg->ref_vor, cfg->ref_hel};
	if (g.D == 3) loss_partials_kernel<3><<<lb, ADJ_THREADS, 0, st>>>(lin, (int)Q, partials);
	else loss_partials_kernel<2><<<lb, ADJ_THREADS, 0, st>>>(lin, (int)Q, partials);
	loss_final_kernel<<<1, 256, 0, st>>>(partials, lb, sums);
	g_launches += 2;
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_backward_epilogue(const gsr_grid_desc *d, const float *scalings, const float *rotations, int64_t N,
				     const float *acc, int sets_mask, float *const out[GSR_NSETS][4], void *stream)
{
	Grid g;
	if (!make_grid(d, g) || N < 0 || !acc || !out) return GSR_EINVAL;
	if (N == 0 || sets_mask == 0) return GSR_OK;
	OutPtrs o;
	for (int s = 0; s < GSR_NSETS; s++)
		for (int k = 0; k < 4; k++) {
			o.p[s][k] = out[s][k];
			if (((sets_mask >> s) & 1) && !out[s][k]) return GSR_EINVAL;
		}
	cudaStream_t st = (cudaStream_t)stream;
	int blocks = (int)((N + 127) / 128);
	g_launches += 1;
	if (g.D == 3) epilogue_kernel<3><<<blocks, 128, 0, st>>>(scalings, rotations, (int)N, acc, sets_mask, o);
	else epilogue_kernel<2><<<blocks, 128, 0, st>>>(scalings, rotations, (int)N, acc, sets_mask, o);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
