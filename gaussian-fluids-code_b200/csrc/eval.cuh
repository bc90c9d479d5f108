// eval.cuh — device functions that evaluate the truncated-Gaussian field at one point.
//
// u(x) = sum_i v_i (g_i - tau)+,  grad u[d][l] = sum_i v_i[d] * (-g_i (Sigma_i^-1 (x - mu_i))[l]),
// g_i = exp(-1/2 (x-mu_i)^T Sigma_i^-1 (x-mu_i)), accepted iff g_i >= tau   (3D/GSR.py:599-632, 2D/GSR.py:527-547)
//
// Candidates come from the 27 (9) cells around the point's cell.  Because the hash is cell-sorted and the
// cell key is row-major with z (y in 2D) fastest, the stencil is 9 (3) CONTIGUOUS runs of packed records.
// Per candidate: 3 FADD + 9 FFMA/FMUL (w = A d) + 3 (q = d.w) + compare  = 24 flop;
// per accepted pair: 1 FMUL + 1 MUFU.EX2 + 28 flop.
//
// LANES lanes cooperate on one point (LANES = 1: one thread per point — the throughput shape for large Q;
// LANES = 8 / 32: the lanes stride through each run and the 12 partial sums are combined with xor-shuffles —
// the latency shape for the reference's own sizes, where Q is a few thousand points).
#pragma once
#include "common.cuh"

namespace gsr {

// Acceptance is decided on q = d^T Sigma^-1 d against q_thr = -2 h_thr, h_thr = min{h : expf(h) >= tau}
// (bisection with libm's expf on the host).  Scaling by -1/2 is exact, so `q <= q_thr` is the reference's
// `exp(-q/2) >= tau` for a monotone expf, and the MUFU stays off the rejected ~85 % of the candidates.
struct EvalParams {
	Grid g;
	float h_thr;
	float q_thr;
};

constexpr float kNegHalfLog2e = -0.72134752044448170368f;	// exp(-q/2) = 2^(q * kNegHalfLog2e)

__device__ __forceinline__ float ex2_approx(float x)
{
	float y;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

template <int LANES>
__device__ __forceinline__ float lane_sum(float v)
{
#pragma unroll
	for (int o = LANES / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Flatten 9 runs [s_r, s_r + n_r) into one list for a group of LANES (8 or 32) lanes: run r is looked up by lane r % LANES
// (`lookup(r, s, n)` leaves s = n = 0 for an absent run), a scan inside the group orders them; every lane receives
// pre[r] = first flat index of run r and off[r] = s_r - pre[r], so that flat index f of run r is sorted index f + off[r].
template <int LANES, typename F>
__device__ __forceinline__ void flat_runs3(int lane, bool any, F lookup, int (&pre)[9], int (&off)[9], int &total)
{
	static_assert(LANES == 8 || LANES == 32, "lane groups of 8 or 32");
	int sa = 0, na = 0, sb = 0, nb = 0;
	if (any && lane < 9) lookup(lane, sa, na);
	if (any && LANES == 8 && lane == 0) lookup(8, sb, nb);
	int incl = na;
#pragma unroll
	for (int o = 1; o < (LANES == 8 ? 8 : 16); o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, incl, o, LANES);
		if (lane >= o) incl += t;
	}
	const int tot_a = __shfl_sync(0xffffffffu, incl, LANES == 8 ? 7 : 8, LANES);
#pragma unroll
	for (int r = 0; r < (LANES == 8 ? 8 : 9); r++) {
		pre[r] = __shfl_sync(0xffffffffu, incl - na, r, LANES);
		off[r] = __shfl_sync(0xffffffffu, sa, r, LANES) - pre[r];
	}
	total = tot_a;
	if (LANES == 8) {
		pre[8] = tot_a;
		off[8] = __shfl_sync(0xffffffffu, sb, 0, LANES) - tot_a;
		total = tot_a + __shfl_sync(0xffffffffu, nb, 0, LANES);
	}
}

// `lane` in [0, LANES).  With LANES > 1 every lane of the warp must call this (shuffles), valid or not.
template <bool NEED_GRAD, int LANES>
__device__ __forceinline__ void eval_point3(const EvalParams &P, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
					    float x, float y, float z, int lane, float u[3], float G[9])
{
	const Grid &g = P.g;
	u[0] = u[1] = u[2] = 0.f;
	if (NEED_GRAD) {
#pragma unroll
		for (int k = 0; k < 9; k++) G[k] = 0.f;
	}
	const float gs = grid_gs(g);
	const int cx = cell_coord(x, g.lo[0], gs), cy = cell_coord(y, g.lo[1], gs), cz = cell_coord(z, g.lo[2], gs);
	const int zlo = max(cz - 1, 0), zhi = min(cz + 1, g.dims[2] - 1);
	const float tau = g.tau, q_thr = P.q_thr;
	if (LANES == 32) {	// (measured: for 8-lane groups the run lookup per candidate costs more than the dependent loads it saves)
		// Latency shape: LANES lanes (a warp, or 8 lanes) on ONE point (the reference's own sizes: up to ~10^5 points with
		// ~200-300 candidates each).  The 9 runs are looked up by the lanes at once (one round trip instead of nine dependent
		// ones), concatenated by a scan inside the lane group, and the lanes then stride through the FLAT candidate list —
		// every lane busy, all record loads independent.
		int pre[9], off[9], total;
		flat_runs3<(LANES >= 8 ? LANES : 8)>(lane, zlo <= zhi, [&](int r, int &s, int &n) {
			const int gi = cx - 1 + r / 3, gj = cy - 1 + r % 3;
			if (gi >= 0 && gi < g.dims[0] && gj >= 0 && gj < g.dims[1]) {
				const int base = (gi * g.dims[1] + gj) * g.dims[2];
				s = __ldg(cell_start + base + zlo);
				n = __ldg(cell_start + base + zhi + 1) - s;
			}
		}, pre, off, total);
		// two candidates per trip, all six record loads issued first (the shape is bound by the latency of these loads)
		for (int f0 = lane; f0 < total; f0 += 2 * LANES) {
			float4 Pa[2], Pb[2], Pc[2];
			bool okc[2];
#pragma unroll
			for (int b = 0; b < 2; b++) {
				const int f = f0 + b * LANES;
				okc[b] = f < total;
				const int ff = okc[b] ? f : f0;
				int d = off[0];
#pragma unroll
				for (int r = 1; r < 9; r++) d = (ff >= pre[r]) ? off[r] : d;
				const int t = ff + d;
				Pa[b] = __ldg(packed + 3 * t); Pb[b] = __ldg(packed + 3 * t + 1); Pc[b] = __ldg(packed + 3 * t + 2);
			}
#pragma unroll
			for (int b = 0; b < 2; b++) {
				if (!okc[b]) continue;
				const float4 p0 = Pa[b], p1 = Pb[b], p2 = Pc[b];
				const float dx = x - p0.x, dy = y - p0.y, dz = z - p0.z;
				const float wx = p1.x * dx + p1.y * dy + p1.z * dz;
				const float wy = p1.y * dx + p2.x * dy + p2.y * dz;
				const float wz = p1.z * dx + p2.y * dy + p2.z * dz;
				const float q = dx * wx + dy * wy + dz * wz;
				if (q <= q_thr) {
					const float gs_ = ex2_approx(q * kNegHalfLog2e);
					const float gm = gs_ - tau;
					u[0] = fmaf(p0.w, gm, u[0]);
					u[1] = fmaf(p1.w, gm, u[1]);
					u[2] = fmaf(p2.w, gm, u[2]);
					if (NEED_GRAD) {
						const float ax = -gs_ * wx, ay = -gs_ * wy, az = -gs_ * wz;
						G[0] = fmaf(p0.w, ax, G[0]); G[1] = fmaf(p0.w, ay, G[1]); G[2] = fmaf(p0.w, az, G[2]);
						G[3] = fmaf(p1.w, ax, G[3]); G[4] = fmaf(p1.w, ay, G[4]); G[5] = fmaf(p1.w, az, G[5]);
						G[6] = fmaf(p2.w, ax, G[6]); G[7] = fmaf(p2.w, ay, G[7]); G[8] = fmaf(p2.w, az, G[8]);
					}
				}
			}
		}
	} else if (zlo <= zhi) {
		for (int gi = max(cx - 1, 0); gi <= min(cx + 1, g.dims[0] - 1); gi++) {
			for (int gj = max(cy - 1, 0); gj <= min(cy + 1, g.dims[1] - 1); gj++) {
				const int base = (gi * g.dims[1] + gj) * g.dims[2];
				const int s = __ldg(cell_start + base + zlo), e = __ldg(cell_start + base + zhi + 1);
				for (int t = s + lane; t < e; t += LANES) {
					const float4 p0 = __ldg(packed + 3 * t), p1 = __ldg(packed + 3 * t + 1), p2 = __ldg(packed + 3 * t + 2);
					const float dx = x - p0.x, dy = y - p0.y, dz = z - p0.z;
					const float wx = p1.x * dx + p1.y * dy + p1.z * dz;
					const float wy = p1.y * dx + p2.x * dy + p2.y * dz;
					const float wz = p1.z * dx + p2.y * dy + p2.z * dz;
					const float q = dx * wx + dy * wy + dz * wz;
					if (q <= q_thr) {
						const float gs_ = ex2_approx(q * kNegHalfLog2e);
						const float gm = gs_ - tau;
						u[0] = fmaf(p0.w, gm, u[0]);
						u[1] = fmaf(p1.w, gm, u[1]);
						u[2] = fmaf(p2.w, gm, u[2]);
						if (NEED_GRAD) {
							const float ax = -gs_ * wx, ay = -gs_ * wy, az = -gs_ * wz;
							G[0] = fmaf(p0.w, ax, G[0]); G[1] = fmaf(p0.w, ay, G[1]); G[2] = fmaf(p0.w, az, G[2]);
							G[3] = fmaf(p1.w, ax, G[3]); G[4] = fmaf(p1.w, ay, G[4]); G[5] = fmaf(p1.w, az, G[5]);
							G[6] = fmaf(p2.w, ax, G[6]); G[7] = fmaf(p2.w, ay, G[7]); G[8] = fmaf(p2.w, az, G[8]);
						}
					}
				}
			}
		}
	}
	if (LANES > 1) {
#pragma unroll
		for (int k = 0; k < 3; k++) u[k] = lane_sum<LANES>(u[k]);
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 9; k++) G[k] = lane_sum<LANES>(G[k]);
		}
	}
}

template <bool NEED_GRAD, int LANES>
__device__ __forceinline__ void eval_point2(const EvalParams &P, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
					    float x, float y, int lane, float u[2], float G[4])
{
	const Grid &g = P.g;
	u[0] = u[1] = 0.f;
	if (NEED_GRAD) G[0] = G[1] = G[2] = G[3] = 0.f;
	const float gs = grid_gs(g);
	const int cx = cell_coord(x, g.lo[0], gs), cy = cell_coord(y, g.lo[1], gs);
	const int ylo = max(cy - 1, 0), yhi = min(cy + 1, g.dims[1] - 1);
	const float tau = g.tau, q_thr = P.q_thr;
	if (ylo <= yhi) {
		for (int gi = max(cx - 1, 0); gi <= min(cx + 1, g.dims[0] - 1); gi++) {
			const int base = gi * g.dims[1];
			const int s = __ldg(cell_start + base + ylo), e = __ldg(cell_start + base + yhi + 1);
			for (int t = s + lane; t < e; t += LANES) {
				const float4 p0 = __ldg(packed + 2 * t), p1 = __ldg(packed + 2 * t + 1);
				const float dx = x - p0.x, dy = y - p0.y;
				const float wx = p1.x * dx + p1.y * dy, wy = p1.y * dx + p1.z * dy;
				const float q = dx * wx + dy * wy;
				if (q <= q_thr) {
					const float gs_ = ex2_approx(q * kNegHalfLog2e);
					const float gm = gs_ - tau;
					u[0] = fmaf(p0.z, gm, u[0]);
					u[1] = fmaf(p0.w, gm, u[1]);
					if (NEED_GRAD) {
						const float ax = -gs_ * wx, ay = -gs_ * wy;
						G[0] = fmaf(p0.z, ax, G[0]); G[1] = fmaf(p0.z, ay, G[1]);
						G[2] = fmaf(p0.w, ax, G[2]); G[3] = fmaf(p0.w, ay, G[3]);
					}
				}
			}
		}
	}
	if (LANES > 1) {
		u[0] = lane_sum<LANES>(u[0]);
		u[1] = lane_sum<LANES>(u[1]);
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 4; k++) G[k] = lane_sum<LANES>(G[k]);
		}
	}
}

// C = A * B (row-major 3x3)
__device__ __forceinline__ void mm3(const float *A, const float *B, float *C)
{
#pragma unroll
	for (int i = 0; i < 3; i++)
#pragma unroll
		for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ void mm2(const float *A, const float *B, float *C)
{
	C[0] = A[0] * B[0] + A[1] * B[2]; C[1] = A[0] * B[1] + A[1] * B[3];
	C[2] = A[2] * B[0] + A[3] * B[2]; C[3] = A[2] * B[1] + A[3] * B[3];
}

float host_h_threshold(float tau);	// defined in eval.cu

inline EvalParams make_params(const Grid &g)
{
	EvalParams P;
	P.g = g;
	P.h_thr = host_h_threshold(g.tau);
	P.q_thr = -2.f * P.h_thr;	// exact
	return P;
}

// lanes per point (or per Gaussian) for a problem of n items: keep >= ~250k threads in flight when n is small
extern int g_lanes8_min_n;	// eval.cu, GSR_TUNE_LANES8_MIN_N
inline int pick_lanes(int64_t n) { return n >= (1 << 18) ? 1 : (n >= g_lanes8_min_n ? 8 : 32); }

}  // namespace gsr
