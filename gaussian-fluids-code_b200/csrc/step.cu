// step.cu — the per-iteration optimiser step of project(), fused and device-resident.
//
// Reference (3D/advance.py:183-287 + 3D/GSR.py:148-152, :704-716; 2D/advance.py:187-259): PCGrad-style mutual
// projection of the vorticity and divergence gradients per parameter group (4 x 3 global dot products, each a
// host sync), torch-autograd regularisers, 4 x torch.optim.Adam, 4 x ReduceLROnPlateau (float(loss) sync),
// scalings.min().item() (sync) — about 200 small launches and >= 8 host syncs per iteration.
// Here: 4 launches, no sync.  HBM-bound: per Gaussian reads params 52 B + Adam m,v 104 B + accumulators
// 48 B/set (twice: kernels A and B), writes params 52 B + m,v 104 B.
#include "chain.cuh"
#include "hash_small.cuh"
#include <math.h>
#include <cooperative_groups.h>

namespace gsr {

int g_step_small_n = 2048;	// GSR_TUNE_STEP_SMALL_N: up to this many Gaussians gsr_step is ONE cluster launch (else four launches)

template <int D> struct Dim {
	static constexpr int P = (D == 3) ? 13 : 7;	// parameter floats per Gaussian
	static constexpr int AF = (D == 3) ? 12 : 7;	// accumulator floats per set
	static constexpr int NR = (D == 3) ? 4 : 1;	// rotation parameters
};

// parameter-space gradient (positions D, scalings D, rotations NR, values D) of one compact accumulator record
template <int D>
__device__ __forceinline__ void param_grad(const float *a, const float *sc, const float *rot, float *gp, float *gs, float *gr, float *gv)
{
	if (D == 3) {
		chain3d(a + 6, sc, rot, gs, gr);
#pragma unroll
		for (int k = 0; k < 3; k++) { gv[k] = a[k]; gp[k] = a[3 + k]; }
	} else {
		chain2d(a + 4, sc, rot[0], gs, gr);
#pragma unroll
		for (int k = 0; k < 2; k++) { gv[k] = a[k]; gp[k] = a[2 + k]; }
	}
}

// partial-sum slots of kernel A
enum { S_DOT = 0 /* [g]: <g_vor, g_div> */, S_N1 = 4 /* |g_vor|^2 */, S_N2 = 8 /* |g_div|^2 */, S_V = 12, S_V2 = 13, S_ANISO = 14, S_ABSV = 15, S_DPOS = 16, S_COUNT = 20 };

constexpr int ST_THREADS = 128;

// per-Gaussian contributions to the partial sums of kernel A (added into S)
template <int D>
__device__ __forceinline__ void step_moments_one(int i, int N, const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot,
						 const float *__restrict__ vals, const float *__restrict__ acc, int sets_mask, const float *__restrict__ pos_org,
						 float aniso_ratio, float (&S)[S_COUNT])
{
	constexpr int AF = Dim<D>::AF, NR = Dim<D>::NR;
	float sc[D], r[NR];
#pragma unroll
	for (int k = 0; k < D; k++) sc[k] = scal[(size_t)D * i + k];
#pragma unroll
	for (int k = 0; k < NR; k++) r[k] = rot[(size_t)NR * i + k];
	if ((sets_mask & 6) == 6) {
		float p1[D], s1[D], r1[NR], v1[D], p2[D], s2[D], r2[NR], v2[D];
		param_grad<D>(acc + ((size_t)1 * N + i) * AF, sc, r, p1, s1, r1, v1);
		param_grad<D>(acc + ((size_t)2 * N + i) * AF, sc, r, p2, s2, r2, v2);
#pragma unroll
		for (int k = 0; k < D; k++) {
			S[S_DOT + 0] += p1[k] * p2[k]; S[S_N1 + 0] += p1[k] * p1[k]; S[S_N2 + 0] += p2[k] * p2[k];
			S[S_DOT + 1] += s1[k] * s2[k]; S[S_N1 + 1] += s1[k] * s1[k]; S[S_N2 + 1] += s2[k] * s2[k];
			S[S_DOT + 3] += v1[k] * v2[k]; S[S_N1 + 3] += v1[k] * v1[k]; S[S_N2 + 3] += v2[k] * v2[k];
		}
#pragma unroll
		for (int k = 0; k < NR; k++) { S[S_DOT + 2] += r1[k] * r2[k]; S[S_N1 + 2] += r1[k] * r1[k]; S[S_N2 + 2] += r2[k] * r2[k]; }
	}
	// regulariser moments (3D/advance.py:237-242): V = exp(-sum s), rho = exp(max s - min s)
	float ssum = 0.f, smin = sc[0], smax = sc[0];
#pragma unroll
	for (int k = 0; k < D; k++) { ssum += sc[k]; smin = fminf(smin, sc[k]); smax = fmaxf(smax, sc[k]); }
	const float V = expf(-ssum);
	S[S_V] += V;
	S[S_V2] += V * V;
	const float rho = expf(smax - smin);
	S[S_ANISO] += (rho >= aniso_ratio ? rho : aniso_ratio) - aniso_ratio;
#pragma unroll
	for (int k = 0; k < D; k++) S[S_ABSV] += fabsf(vals[(size_t)D * i + k]);
	if (pos_org) {
#pragma unroll
		for (int k = 0; k < D; k++) {
			const float dlt = pos[(size_t)D * i + k] - pos_org[(size_t)D * i + k];
			S[S_DPOS] += dlt * dlt;
		}
	}
}

template <int D>
__global__ void __launch_bounds__(ST_THREADS) stepA_kernel(int N, const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot,
							   const float *__restrict__ vals, const float *__restrict__ acc, int sets_mask,
							   const float *__restrict__ pos_org, float aniso_ratio, float *__restrict__ partials)
{
	__shared__ float sm[ST_THREADS / 32][S_COUNT];
	float S[S_COUNT];
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) S[k] = 0.f;
	int i = blockIdx.x * ST_THREADS + threadIdx.x;
	if (i < N) step_moments_one<D>(i, N, pos, scal, rot, vals, acc, sets_mask, pos_org, aniso_ratio, S);
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) {
#pragma unroll
		for (int o = 16; o; o >>= 1) S[k] += __shfl_xor_sync(0xffffffffu, S[k], o);
	}
	if ((threadIdx.x & 31) == 0) {
#pragma unroll
		for (int k = 0; k < S_COUNT; k++) sm[threadIdx.x >> 5][k] = S[k];
	}
	__syncthreads();
	if (threadIdx.x < S_COUNT) {
		float s = 0.f;
#pragma unroll
		for (int w = 0; w < ST_THREADS / 32; w++) s += sm[w][threadIdx.x];
		partials[(size_t)blockIdx.x * S_COUNT + threadIdx.x] = s;
	}
}

// coefficient block written by kernel R for kernel B (inside the state scalars)
enum { C_A1 = 24 /* [4] */, C_A2 = 28 /* [4] */, C_STEP = 32 /* [4] lr / bias_correction1 */, C_BC2 = 36 /* 1/sqrt(bias_correction2) */,
       C_MEANV = 37, C_MEANR2 = 38, C_LRUSED = 40 /* [4] */ };

struct LossSrcs {
	const float *partials[3];
	int nblocks[3];
	float w[3][8];
	int n;
};

// everything kernel R does once the global sums T are known: PCGrad coefficients, losses, Adam bias corrections, scheduler
template <int D>
__device__ __forceinline__ void step_reduce_tail(const gsr_step_cfg &cfg, int N, const double *T, float *__restrict__ st)
{
	// PCGrad coefficients (3D/advance.py:202-225): g1 -= <g1,n2> n2, g2 -= <g2,n1> n1 when <g1,g2> < 0
	for (int g = 0; g < 4; g++) {
		float a1 = 1.f, a2 = 1.f;
		if (cfg.pcgrad && T[S_DOT + g] < 0.) {
			a1 = (float)(1. - T[S_DOT + g] / T[S_N1 + g]);
			a2 = (float)(1. - T[S_DOT + g] / T[S_N2 + g]);
		}
		st[C_A1 + g] = a1;
		st[C_A2 + g] = a2;
	}
	const double n = (double)N;
	const double meanV = T[S_V] / n, meanR2 = T[S_V2] / n / (meanV * meanV);
	st[C_MEANV] = (float)meanV;
	st[C_MEANR2] = (float)meanR2;
	const double L_aniso = T[S_ANISO] / n, L_vol = meanR2 - 1., L_valreg = T[S_ABSV] / (n * D), L_dpos = T[S_DPOS] / (n * D);
	double loss_src = 0.;
	for (int k = 0; k < 8; k++) loss_src += T[S_COUNT + k];
	const double loss_tot = loss_src + cfg.w_aniso * L_aniso + cfg.w_vol * L_vol + cfg.w_valreg * L_valreg + cfg.w_dpos * L_dpos;
	st[GSR_ST_LOSS_TOT] = (float)loss_tot;
	st[GSR_ST_L_ANISO] = (float)L_aniso;
	st[GSR_ST_L_VOL] = (float)L_vol;
	st[GSR_ST_L_VALREG] = (float)L_valreg;
	st[GSR_ST_L_DPOS] = (float)L_dpos;
	// Adam bias corrections for step t (torch: step_size = lr / (1 - beta1^t), denom = sqrt(v)/sqrt(1 - beta2^t) + eps)
	const double t = (double)st[GSR_ST_T] + 1.;
	st[GSR_ST_T] = (float)t;
	st[GSR_ST_CLOCK] += 1.f;
	const double p1 = pow((double)cfg.beta1, t), p2 = pow((double)cfg.beta2, t);
	const double bc1 = 1. - p1, bc2 = 1. - p2;
	reinterpret_cast<double *>(st + GSR_ST_BPOW)[0] = p1;	// (the four-lane step advances these by multiplication instead of calling pow)
	reinterpret_cast<double *>(st + GSR_ST_BPOW)[1] = p2;
	for (int g = 0; g < 4; g++) {
		st[C_LRUSED + g] = st[GSR_ST_LR + g];
		st[C_STEP + g] = (float)((double)st[GSR_ST_LR + g] / bc1);
	}
	st[C_BC2] = (float)(1. / sqrt(bc2));
	// ReduceLROnPlateau.step(loss_tot): mode min, rel threshold, cooldown 0
	const float cur = (float)loss_tot;
	float best = st[GSR_ST_BEST], bad = st[GSR_ST_BAD];
	if (cur < best * (1.f - cfg.sched_threshold)) { best = cur; bad = 0.f; } else bad += 1.f;
	if (bad > (float)cfg.sched_patience) {
		for (int g = 0; g < 4; g++) {
			const float old_lr = st[GSR_ST_LR + g], new_lr = fmaxf(old_lr * cfg.sched_factor, cfg.sched_min_lr);
			if (old_lr - new_lr > cfg.sched_eps) st[GSR_ST_LR + g] = new_lr;
		}
		bad = 0.f;
	}
	st[GSR_ST_BEST] = best;
	st[GSR_ST_BAD] = bad;
}

template <int D>
__global__ void __launch_bounds__(256) stepR_kernel(gsr_step_cfg cfg, int N, int nblkA, const float *__restrict__ partials, LossSrcs ls, float *__restrict__ st)
{
	__shared__ double red[S_COUNT + 8][8];
	// deterministic reduction: 8 strided lanes per slot, then a fixed-order sum
	const int slot = threadIdx.x >> 3, lane = threadIdx.x & 7;
	if (slot < S_COUNT) {
		double s = 0.;
		for (int b = lane; b < nblkA; b += 8) s += (double)partials[(size_t)b * S_COUNT + slot];
		red[slot][lane] = s;
	} else if (slot < S_COUNT + 8) {
		const int k = slot - S_COUNT;
		double s = 0.;
		for (int src = 0; src < ls.n; src++) {
			double t = 0.;
			for (int b = lane; b < ls.nblocks[src]; b += 8) t += (double)ls.partials[src][(size_t)b * 8 + k];
			s += (double)ls.w[src][k] * t;
		}
		red[slot][lane] = s;
	}
	__syncthreads();
	if (threadIdx.x != 0) return;
	double T[S_COUNT + 8];
	for (int k = 0; k < S_COUNT + 8; k++) {
		double s = 0.;
		for (int l = 0; l < 8; l++) s += red[k][l];
		T[k] = s;
	}
	step_reduce_tail<D>(cfg, N, T, st);
}

__device__ __forceinline__ void atomic_min_f(float *addr, float v)
{
	if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
	else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// projected total gradient + regulariser gradients + Adam for Gaussian i; returns min over its new scalings
template <int D>
__device__ __forceinline__ float step_update_one(const gsr_step_cfg &cfg, int N, int i, float *__restrict__ pos, float *__restrict__ scal, float *__restrict__ rot,
						 float *__restrict__ vals, const float *__restrict__ acc, int sets_mask, const float *__restrict__ ex0,
						 const float *__restrict__ ex1, const float *__restrict__ pos_org, float *__restrict__ st)
{
	constexpr int AF = Dim<D>::AF, NR = Dim<D>::NR, P = Dim<D>::P;
	float smin = __int_as_float(0x7f800000);
	float sc[D], r[NR], p[D], v[D];
#pragma unroll
	for (int k = 0; k < D; k++) { sc[k] = scal[(size_t)D * i + k]; p[k] = pos[(size_t)D * i + k]; v[k] = vals[(size_t)D * i + k]; }
#pragma unroll
	for (int k = 0; k < NR; k++) r[k] = rot[(size_t)NR * i + k];
	float g[P];	// total gradient: [pos D][scal D][rot NR][val D]
#pragma unroll
	for (int k = 0; k < P; k++) g[k] = 0.f;
	float gp[D], gs[D], gr[NR], gv[D];
	auto add = [&](const float *a, float cp, float cs, float cr, float cv) {
		param_grad<D>(a, sc, r, gp, gs, gr, gv);
#pragma unroll
		for (int k = 0; k < D; k++) { g[k] += cp * gp[k]; g[D + k] += cs * gs[k]; g[2 * D + NR + k] += cv * gv[k]; }
#pragma unroll
		for (int k = 0; k < NR; k++) g[2 * D + k] += cr * gr[k];
	};
	if (sets_mask & 2) add(acc + ((size_t)1 * N + i) * AF, st[C_A1 + 0], st[C_A1 + 1], st[C_A1 + 2], st[C_A1 + 3]);
	if (sets_mask & 4) add(acc + ((size_t)2 * N + i) * AF, st[C_A2 + 0], st[C_A2 + 1], st[C_A2 + 2], st[C_A2 + 3]);
	if (sets_mask & 1) add(acc + (size_t)i * AF, 1.f, 1.f, 1.f, 1.f);
	if (ex0) add(ex0 + (size_t)i * AF, 1.f, 1.f, 1.f, 1.f);
	if (ex1) add(ex1 + (size_t)i * AF, 1.f, 1.f, 1.f, 1.f);
	// closed-form regulariser gradients (autograd in the reference, 3D/advance.py:237-244)
	{
		float ssum = 0.f;
		int kmin = 0, kmax = 0;
#pragma unroll
		for (int k = 0; k < D; k++) {
			ssum += sc[k];
			if (sc[k] < sc[kmin]) kmin = k;	// first index on ties, like torch.min / torch.max
			if (sc[k] > sc[kmax]) kmax = k;
		}
		const float rho = expf(sc[kmax] - sc[kmin]);
		if (rho >= cfg.aniso_ratio && kmin != kmax) {
			const float c = cfg.w_aniso * rho / (float)N;
#pragma unroll
			for (int k = 0; k < D; k++) g[D + k] += (k == kmax ? c : 0.f) - (k == kmin ? c : 0.f);
		}
		const float rV = expf(-ssum) / st[C_MEANV];
		const float cv = -cfg.w_vol * 2.f / (float)N * rV * (rV - st[C_MEANR2]);
#pragma unroll
		for (int k = 0; k < D; k++) g[D + k] += cv;
		if (cfg.w_valreg != 0.f) {
			const float c = cfg.w_valreg / (float)(N * D);
#pragma unroll
			for (int k = 0; k < D; k++) g[2 * D + NR + k] += c * (float)((v[k] > 0.f) - (v[k] < 0.f));
		}
		if (pos_org && cfg.w_dpos != 0.f) {
			const float c = cfg.w_dpos * 2.f / (float)(N * D);
#pragma unroll
			for (int k = 0; k < D; k++) g[k] += c * (p[k] - pos_org[(size_t)D * i + k]);
		}
	}
	// Adam (torch.optim.Adam defaults; lr of the group as it was BEFORE this step's scheduler update)
	float *m = st + GSR_STATE_SCALARS + (size_t)i * P, *vv = st + GSR_STATE_SCALARS + (size_t)N * P + (size_t)i * P;
	float prm[P];
#pragma unroll
	for (int k = 0; k < D; k++) { prm[k] = p[k]; prm[D + k] = sc[k]; prm[2 * D + NR + k] = v[k]; }
#pragma unroll
	for (int k = 0; k < NR; k++) prm[2 * D + k] = r[k];
	const float bc2 = st[C_BC2];
#pragma unroll
	for (int k = 0; k < P; k++) {
		const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
		const float mk = m[k] + (g[k] - m[k]) * (1.f - cfg.beta1);		// exp_avg.lerp_(grad, 1 - beta1)
		const float vk = vv[k] * cfg.beta2 + (1.f - cfg.beta2) * g[k] * g[k];	// exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
		m[k] = mk;
		vv[k] = vk;
		const float denom = sqrtf(vk) * bc2 + cfg.eps;
		prm[k] -= st[C_STEP + grp] * (mk / denom);
	}
#pragma unroll
	for (int k = 0; k < D; k++) {
		pos[(size_t)D * i + k] = prm[k];
		scal[(size_t)D * i + k] = prm[D + k];
		vals[(size_t)D * i + k] = prm[2 * D + NR + k];
		smin = fminf(smin, prm[D + k]);
	}
#pragma unroll
	for (int k = 0; k < NR; k++) rot[(size_t)NR * i + k] = prm[2 * D + k];
	return smin;
}

template <int D>
__global__ void __launch_bounds__(ST_THREADS) stepB_kernel(gsr_step_cfg cfg, int N, float *__restrict__ pos, float *__restrict__ scal, float *__restrict__ rot, float *__restrict__ vals,
							   const float *__restrict__ acc, int sets_mask, const float *__restrict__ ex0, const float *__restrict__ ex1,
							   const float *__restrict__ pos_org, float *__restrict__ st)
{
	int i = blockIdx.x * ST_THREADS + threadIdx.x;
	float smin = __int_as_float(0x7f800000);
	if (i < N) smin = step_update_one<D>(cfg, N, i, pos, scal, rot, vals, acc, sets_mask, ex0, ex1, pos_org, st);
#pragma unroll
	for (int o = 16; o; o >>= 1) smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
	if ((threadIdx.x & 31) == 0) atomic_min_f(st + GSR_ST_MIN_S, smin);
}

// next grid_scale (3D/GSR.py:248-251), formed in double like the reference's host code, then rounded to f32;
// re-arms the min accumulator for the next iteration.
__device__ __forceinline__ void step_grid_scale(const gsr_step_cfg &cfg, float *st, float min_s, int t_new = -1)
{
	double gs = cfg.grid_scale_tau0;
	if (cfg.grid_coef > 0.) gs = fmax(cfg.grid_coef * exp(-(double)min_s), cfg.min_grid_scale);
	st[GSR_ST_GRID_SCALE] = (float)gs;
	if (cfg.grid_scale_out) *cfg.grid_scale_out = (float)gs;
	if (cfg.sample_gs_slots) {	// the sample grid scales, one iteration ahead (include/gsr_b200.h)
		const float ahead = (float)(gs * (double)cfg.sample_gs_margin);
		if (t_new <= 0) {	// gsr_step_init
			cfg.sample_gs_slots[0] = cfg.sample_gs_slots[1] = ahead;
			st[GSR_ST_SGS_ERR] = 0.f;
		} else {
			if ((float)gs > cfg.sample_gs_slots[t_new & 1]) st[GSR_ST_SGS_ERR] = 1.f;
			cfg.sample_gs_slots[(t_new + 1) & 1] = ahead;
		}
	}
	st[GSR_ST_MIN_S] = __int_as_float(0x7f800000);
}

__global__ void stepS_kernel(gsr_step_cfg cfg, float *st, float *min_out)
{
	const float min_s = st[GSR_ST_MIN_S];
	if (min_out) *min_out = min_s;
	step_grid_scale(cfg, st, min_s, (int)st[GSR_ST_T]);	// T was advanced by kernel R (0 at gsr_step_init)
}

// ---- small N (the reference's own sizes): the whole step as ONE launch of ONE 8-CTA cluster ---------------------------------
// The four-launch form is a chain of dependent memory round trips at these sizes (ncu: long-scoreboard stalls, 18 us inside the
// captured iteration for 1000 Gaussians): kernel boundaries, the re-read of parameters and accumulators by kernel B, Adam
// moments loaded one by one between stores through the same pointer.  Here every Gaussian's record is loaded once, up front;
// the parameter-space gradients of both loss sets are formed once and stay in registers; the two grid-wide steps (PCGrad dots /
// regulariser moments, and min(s)) are cluster barriers with the per-CTA partials read through distributed shared memory; the
// reduction tail (PCGrad coefficients, scheduler, bias corrections) runs redundantly in every CTA on a shared-memory copy of
// the state scalars, which CTA 0 writes back.
constexpr int SC_CTAS = 8;
constexpr int SC_MAX_THREADS = 256;
constexpr int SC_MAX_N = SC_CTAS * SC_MAX_THREADS;

// step_reduce_tail with its three independent strands on three warps (in one thread they are ~4 us of dependent double
// divisions and two pow() while every other thread of the cluster waits at the barrier — ncu: barrier stalls dominated):
// PCGrad coefficients (4 lanes), losses + scheduler decision (1 lane), Adam bias corrections (2 lanes); then the per-group
// step sizes and learning rates.  Same formulas, same results.  All threads of the CTA must call it (needs >= 3 warps).
template <int D>
__device__ __forceinline__ void step_reduce_tail_split(const gsr_step_cfg &cfg, int N, const double *T, float *st, double *bc_sm, int *decay_sm)
{
	const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	if (w == 0 && lane < 4) {
		const int g = lane;
		float a1 = 1.f, a2 = 1.f;
		if (cfg.pcgrad && T[S_DOT + g] < 0.) {
			a1 = (float)(1. - T[S_DOT + g] / T[S_N1 + g]);
			a2 = (float)(1. - T[S_DOT + g] / T[S_N2 + g]);
		}
		st[C_A1 + g] = a1;
		st[C_A2 + g] = a2;
	} else if (w == 1 && lane == 0) {
		const double n = (double)N;
		const double meanV = T[S_V] / n, meanR2 = T[S_V2] / n / (meanV * meanV);
		st[C_MEANV] = (float)meanV;
		st[C_MEANR2] = (float)meanR2;
		const double L_aniso = T[S_ANISO] / n, L_vol = meanR2 - 1., L_valreg = T[S_ABSV] / (n * D), L_dpos = T[S_DPOS] / (n * D);
		double loss_src = 0.;
		for (int k = 0; k < 8; k++) loss_src += T[S_COUNT + k];
		const double loss_tot = loss_src + cfg.w_aniso * L_aniso + cfg.w_vol * L_vol + cfg.w_valreg * L_valreg + cfg.w_dpos * L_dpos;
		st[GSR_ST_LOSS_TOT] = (float)loss_tot;
		st[GSR_ST_L_ANISO] = (float)L_aniso;
		st[GSR_ST_L_VOL] = (float)L_vol;
		st[GSR_ST_L_VALREG] = (float)L_valreg;
		st[GSR_ST_L_DPOS] = (float)L_dpos;
		const float cur = (float)loss_tot;
		float best = st[GSR_ST_BEST], bad = st[GSR_ST_BAD];
		if (cur < best * (1.f - cfg.sched_threshold)) { best = cur; bad = 0.f; } else bad += 1.f;
		const int decay = bad > (float)cfg.sched_patience;
		if (decay) bad = 0.f;
		st[GSR_ST_BEST] = best;
		st[GSR_ST_BAD] = bad;
		*decay_sm = decay;
	} else if (w == 2 && lane < 2) {
		const double t = (double)st[GSR_ST_T] + 1.;
		const double pw = pow((double)(lane ? cfg.beta2 : cfg.beta1), t);
		bc_sm[lane] = 1. - pw;
		reinterpret_cast<double *>(st + GSR_ST_BPOW)[lane] = pw;
	}
	__syncthreads();
	if (tid < 4) {
		const int g = tid;
		const float lr = st[GSR_ST_LR + g];
		st[C_LRUSED + g] = lr;
		st[C_STEP + g] = (float)((double)lr / bc_sm[0]);
		if (*decay_sm) {
			const float new_lr = fmaxf(lr * cfg.sched_factor, cfg.sched_min_lr);
			if (lr - new_lr > cfg.sched_eps) st[GSR_ST_LR + g] = new_lr;
		}
	} else if (tid == 4) {
		st[C_BC2] = (float)(1. / sqrt(bc_sm[1]));
	} else if (tid == 5) {
		st[GSR_ST_T] = (float)((double)st[GSR_ST_T] + 1.);
		st[GSR_ST_CLOCK] += 1.f;
	}
}

template <int D>
__global__ void __cluster_dims__(SC_CTAS, 1, 1) __launch_bounds__(SC_MAX_THREADS, 1)
step_cluster_kernel(gsr_step_cfg cfg, int N, float *__restrict__ pos, float *__restrict__ scal, float *__restrict__ rot, float *__restrict__ vals,
		    const float *__restrict__ acc, int sets_mask, const float *__restrict__ ex0, const float *__restrict__ ex1,
		    const float *__restrict__ pos_org, LossSrcs ls, float *st)
{
	namespace cg = cooperative_groups;
	constexpr int AF = Dim<D>::AF, NR = Dim<D>::NR, P = Dim<D>::P;
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ __align__(8) float cst[GSR_STATE_SCALARS];	// the state scalars: read once, advanced by the tail, written back by CTA 0
	__shared__ float wpart[SC_MAX_THREADS / 32][S_COUNT];
	__shared__ float part[S_COUNT];	// this CTA's partial sums (read by every CTA of the cluster)
	__shared__ double Tsm[S_COUNT + 8];
	__shared__ double bc_sm[2];
	__shared__ int decay_sm;
	__shared__ float wmin[SC_MAX_THREADS / 32];
	__shared__ float bmin;	// this CTA's min over the new scalings (read by CTA 0)
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
	const int rank = (int)cluster.block_rank();
	const int i = rank * blockDim.x + tid;
	const bool on = i < N;
	// ---- every load of the step, issued together --------------------------------------------------------------------------
	for (int k = tid; k < GSR_STATE_SCALARS; k += blockDim.x) cst[k] = st[k];
	float sc[D], r[NR], p[D], v[D], po[D], mo[P], vo[P], a0[AF], a1[AF], a2[AF], b0[AF], b1[AF];
	float *m = st + GSR_STATE_SCALARS + (size_t)i * P, *vv = st + GSR_STATE_SCALARS + (size_t)N * P + (size_t)i * P;
	if (on) {
#pragma unroll
		for (int k = 0; k < D; k++) {
			sc[k] = scal[(size_t)D * i + k]; p[k] = pos[(size_t)D * i + k]; v[k] = vals[(size_t)D * i + k];
			po[k] = pos_org ? pos_org[(size_t)D * i + k] : 0.f;
		}
#pragma unroll
		for (int k = 0; k < NR; k++) r[k] = rot[(size_t)NR * i + k];
#pragma unroll
		for (int k = 0; k < P; k++) { mo[k] = m[k]; vo[k] = vv[k]; }
#pragma unroll
		for (int k = 0; k < AF; k++) {
			a0[k] = (sets_mask & 1) ? acc[(size_t)i * AF + k] : 0.f;
			a1[k] = (sets_mask & 2) ? acc[((size_t)1 * N + i) * AF + k] : 0.f;
			a2[k] = (sets_mask & 4) ? acc[((size_t)2 * N + i) * AF + k] : 0.f;
			b0[k] = ex0 ? ex0[(size_t)i * AF + k] : 0.f;
			b1[k] = ex1 ? ex1[(size_t)i * AF + k] : 0.f;
		}
	}
	if (tid >= S_COUNT && tid < S_COUNT + 8) {	// weighted sums of the sample-loss partials (blocks in order, double)
		const int k = tid - S_COUNT;
		double s = 0.;
		for (int src = 0; src < ls.n; src++) {
			double t = 0.;
			for (int b = 0; b < ls.nblocks[src]; b++) t += (double)ls.partials[src][(size_t)b * 8 + k];
			s += (double)ls.w[src][k] * t;
		}
		Tsm[tid] = s;
	}
	// ---- parameter-space gradients [pos D][scal D][rot NR][val D] of the sets, once -------------------------------------------
	float g1[P], g2[P], gd[P];	// vorticity set, divergence set, direct sets (boundary / value / gradient losses)
	float S[S_COUNT];
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) S[k] = 0.f;
	float ssum = 0.f;
	int kmin = 0, kmax = 0;
	if (on) {
		auto to_param = [&](const float *a, float *g) {
			float gp[D], gs[D], gr[NR], gv[D];
			param_grad<D>(a, sc, r, gp, gs, gr, gv);
#pragma unroll
			for (int k = 0; k < D; k++) { g[k] = gp[k]; g[D + k] = gs[k]; g[2 * D + NR + k] = gv[k]; }
#pragma unroll
			for (int k = 0; k < NR; k++) g[2 * D + k] = gr[k];
		};
#pragma unroll
		for (int k = 0; k < P; k++) g1[k] = g2[k] = gd[k] = 0.f;
		if (sets_mask & 2) to_param(a1, g1);
		if (sets_mask & 4) to_param(a2, g2);
		float t[P];
		if (sets_mask & 1) {
			to_param(a0, t);
#pragma unroll
			for (int k = 0; k < P; k++) gd[k] += t[k];
		}
		if (ex0) {
			to_param(b0, t);
#pragma unroll
			for (int k = 0; k < P; k++) gd[k] += t[k];
		}
		if (ex1) {
			to_param(b1, t);
#pragma unroll
			for (int k = 0; k < P; k++) gd[k] += t[k];
		}
		if ((sets_mask & 6) == 6) {
#pragma unroll
			for (int k = 0; k < P; k++) {
				const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
				S[S_DOT + grp] += g1[k] * g2[k]; S[S_N1 + grp] += g1[k] * g1[k]; S[S_N2 + grp] += g2[k] * g2[k];
			}
		}
		// regulariser moments (3D/advance.py:237-242): V = exp(-sum s), rho = exp(max s - min s)
#pragma unroll
		for (int k = 0; k < D; k++) {
			ssum += sc[k];
			if (sc[k] < sc[kmin]) kmin = k;	// first index on ties, like torch.min / torch.max
			if (sc[k] > sc[kmax]) kmax = k;
		}
	}
	float smin_k = 0.f, smax_k = 0.f;
#pragma unroll
	for (int k = 0; k < D; k++) { if (k == kmin) smin_k = on ? sc[k] : 0.f; if (k == kmax) smax_k = on ? sc[k] : 0.f; }
	const float V = expf(-ssum), rho = expf(smax_k - smin_k);
	if (on) {
		S[S_V] += V;
		S[S_V2] += V * V;
		S[S_ANISO] += (rho >= cfg.aniso_ratio ? rho : cfg.aniso_ratio) - cfg.aniso_ratio;
#pragma unroll
		for (int k = 0; k < D; k++) S[S_ABSV] += fabsf(v[k]);
		if (pos_org) {
#pragma unroll
			for (int k = 0; k < D; k++) { const float dlt = p[k] - po[k]; S[S_DPOS] += dlt * dlt; }
		}
	}
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) {
#pragma unroll
		for (int o = 16; o; o >>= 1) S[k] += __shfl_xor_sync(0xffffffffu, S[k], o);
	}
	if (lane == 0) {
#pragma unroll
		for (int k = 0; k < S_COUNT; k++) wpart[w][k] = S[k];
	}
	__syncthreads();
	if (tid < S_COUNT) {
		float s = 0.f;
		for (int ww = 0; ww < nw; ww++) s += wpart[ww][tid];
		part[tid] = s;
	}
	cluster.sync();
	if (tid < S_COUNT) {	// CTAs in order, double: the same T in every CTA
		double s = 0.;
#pragma unroll
		for (int q = 0; q < SC_CTAS; q++) s += (double)*cluster.map_shared_rank(&part[tid], q);
		Tsm[tid] = s;
	}
	__syncthreads();
	if (nw >= 3) {
		step_reduce_tail_split<D>(cfg, N, Tsm, cst, bc_sm, &decay_sm);
	} else if (tid == 0) {
		double T[S_COUNT + 8];
#pragma unroll
		for (int k = 0; k < S_COUNT + 8; k++) T[k] = Tsm[k];
		step_reduce_tail<D>(cfg, N, T, cst);
	}
	__syncthreads();
	// ---- projected gradient + regulariser gradients + Adam (step_update_one, on the registers loaded above) -------------------
	float smin = __int_as_float(0x7f800000);
	if (on) {
		float g[P];
#pragma unroll
		for (int k = 0; k < P; k++) {
			const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
			float t = 0.f;
			if (sets_mask & 2) t += cst[C_A1 + grp] * g1[k];
			if (sets_mask & 4) t += cst[C_A2 + grp] * g2[k];
			g[k] = t + gd[k];
		}
		if (rho >= cfg.aniso_ratio && kmin != kmax) {
			const float c = cfg.w_aniso * rho / (float)N;
#pragma unroll
			for (int k = 0; k < D; k++) g[D + k] += (k == kmax ? c : 0.f) - (k == kmin ? c : 0.f);
		}
		const float rV = V / cst[C_MEANV];
		const float cv = -cfg.w_vol * 2.f / (float)N * rV * (rV - cst[C_MEANR2]);
#pragma unroll
		for (int k = 0; k < D; k++) g[D + k] += cv;
		if (cfg.w_valreg != 0.f) {
			const float c = cfg.w_valreg / (float)(N * D);
#pragma unroll
			for (int k = 0; k < D; k++) g[2 * D + NR + k] += c * (float)((v[k] > 0.f) - (v[k] < 0.f));
		}
		if (pos_org && cfg.w_dpos != 0.f) {
			const float c = cfg.w_dpos * 2.f / (float)(N * D);
#pragma unroll
			for (int k = 0; k < D; k++) g[k] += c * (p[k] - po[k]);
		}
		float prm[P];
#pragma unroll
		for (int k = 0; k < D; k++) { prm[k] = p[k]; prm[D + k] = sc[k]; prm[2 * D + NR + k] = v[k]; }
#pragma unroll
		for (int k = 0; k < NR; k++) prm[2 * D + k] = r[k];
		const float bc2 = cst[C_BC2];
#pragma unroll
		for (int k = 0; k < P; k++) {
			const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
			const float mk = mo[k] + (g[k] - mo[k]) * (1.f - cfg.beta1);
			const float vk = vo[k] * cfg.beta2 + (1.f - cfg.beta2) * g[k] * g[k];
			mo[k] = mk;
			vo[k] = vk;
			prm[k] -= cst[C_STEP + grp] * (mk / (sqrtf(vk) * bc2 + cfg.eps));
		}
#pragma unroll
		for (int k = 0; k < P; k++) { m[k] = mo[k]; vv[k] = vo[k]; }
#pragma unroll
		for (int k = 0; k < D; k++) {
			pos[(size_t)D * i + k] = prm[k];
			scal[(size_t)D * i + k] = prm[D + k];
			vals[(size_t)D * i + k] = prm[2 * D + NR + k];
			smin = fminf(smin, prm[D + k]);
		}
#pragma unroll
		for (int k = 0; k < NR; k++) rot[(size_t)NR * i + k] = prm[2 * D + k];
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
	if (lane == 0) wmin[w] = smin;
	__syncthreads();
	if (tid == 0) {
		float mn = wmin[0];
		for (int ww = 1; ww < nw; ww++) mn = fminf(mn, wmin[ww]);
		bmin = mn;
	}
	cluster.sync();
	if (rank == 0) {
		for (int k = tid; k < GSR_STATE_SCALARS; k += blockDim.x)
			if (k != GSR_ST_GRID_SCALE && k != GSR_ST_MIN_S && k != GSR_ST_SGS_ERR) st[k] = cst[k];
		if (tid == 0) {
			float mn = fminf(bmin, cst[GSR_ST_MIN_S]);
#pragma unroll
			for (int q = 1; q < SC_CTAS; q++) mn = fminf(mn, *cluster.map_shared_rank(&bmin, q));
			step_grid_scale(cfg, st, mn, (int)cst[GSR_ST_T]);
		}
	}
	cluster.sync();	// no CTA leaves while CTA 0 may still read its shared memory
}

__global__ void init_state_kernel(float *st, size_t n_moments, gsr_step_cfg cfg)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_moments) st[GSR_STATE_SCALARS + i] = 0.f;
	if (i < GSR_STATE_SCALARS) {
		float v = 0.f;
		if (i == GSR_ST_BEST || i == GSR_ST_MIN_S) v = __int_as_float(0x7f800000);
		if (i >= GSR_ST_LR && i < GSR_ST_LR + 4) v = cfg.lr[i - GSR_ST_LR];
		if (i == GSR_ST_CLOCK && cfg.keep_clock) return;	// the sample clock runs on across optimisation phases
		if (i >= GSR_ST_BPOW && i < GSR_ST_BPOW + 4) {	// beta1^0 = beta2^0 = 1 (doubles)
			if (!(i & 1)) reinterpret_cast<double *>(st + i)[0] = 1.;
			return;
		}
		st[i] = v;
	}
}

// ---- the same step with FOUR lanes per Gaussian (N <= 1024: the reference's 3D scenes have 1000 Gaussians) --------------------
// ncu of step_cluster_kernel at N = 1000 (profiles/README.md, round 2): 14.8 us for 2100 dependent instructions per thread at one
// warp per scheduler — the chain rule of up to five accumulator sets, then 13 IEEE sqrt + divide of Adam, one Gaussian per thread.
// Here a Gaussian is a group of four consecutive lanes: lane s of the group runs the chain rule of ONE set (vorticity, divergence,
// direct, boundary) and leaves its 13 (7) parameter gradients in shared memory; then lane s owns parameters s, s + 4, s + 8 (, 12)
// for the PCGrad dots, the regulariser gradients and Adam.  Same formulas; the float block sums are grouped differently.
#ifdef GSR_STEP_TIMING
__device__ unsigned long long g_step_stamps[16];
__device__ __forceinline__ unsigned long long gtime()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
#define STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_step_stamps[k] = gtime(); } while (0)
#else
#define STAMP(k) do { } while (0)
#endif

// The reduction tail of the four-lane step in two pieces.  EARLY (needs only the step count and the learning rates, so it runs while
// the loads of the step are in flight): Adam's bias corrections (two double pow) and the per-group step sizes lr / bc1, 1 / sqrt(bc2).
// LATE (once the global sums T are known): PCGrad coefficients and the volume moments — all Adam waits for — beside the loss total,
// the scheduler decision and the learning-rate update, which only the next step needs.  Same formulas as step_reduce_tail.
// early_sm: [0..3] lr / bc1 per group, [4..7] the learning rates in use, [8] 1 / sqrt(bc2)
__device__ __forceinline__ void step_tail_early(const gsr_step_cfg &cfg, const float *__restrict__ st_global, int lane, float *early_sm)
{
	double pw = 1.;	// beta^t by one multiplication per step (pow() in double is ~2 us of one thread's latency, which every warp would wait for)
	if (lane < 2) pw = reinterpret_cast<const double *>(st_global + GSR_ST_BPOW)[lane] * (double)(lane ? cfg.beta2 : cfg.beta1);
	const double p1 = __shfl_sync(0xffffffffu, pw, 0), p2 = __shfl_sync(0xffffffffu, pw, 1);
	const double bc1 = 1. - p1, bc2 = 1. - p2;
	if (lane == 0) { reinterpret_cast<double *>(early_sm + 12)[0] = p1; reinterpret_cast<double *>(early_sm + 12)[1] = p2; }
	if (lane < 4) {
		const float lr = st_global[GSR_ST_LR + lane];
		early_sm[lane] = (float)((double)lr / bc1);
		early_sm[4 + lane] = lr;
	} else if (lane == 4) {
		early_sm[8] = (float)(1. / sqrt(bc2));
	}
}

template <int D>
__device__ __forceinline__ void step_tail_late(const gsr_step_cfg &cfg, int N, const double *T, float *st, const float *early_sm)
{
	const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	if (w == 0 && lane < 4) {
		const int g = lane;
		float a1 = 1.f, a2 = 1.f;
		if (cfg.pcgrad && T[S_DOT + g] < 0.) {
			a1 = (float)(1. - T[S_DOT + g] / T[S_N1 + g]);
			a2 = (float)(1. - T[S_DOT + g] / T[S_N2 + g]);
		}
		st[C_A1 + g] = a1;
		st[C_A2 + g] = a2;
	} else if (w == 1 && lane == 0) {
		const double n = (double)N;
		const double meanV = T[S_V] / n, meanR2 = T[S_V2] / n / (meanV * meanV);
		st[C_MEANV] = (float)meanV;
		st[C_MEANR2] = (float)meanR2;
		const double L_aniso = T[S_ANISO] / n, L_vol = meanR2 - 1., L_valreg = T[S_ABSV] / (n * D), L_dpos = T[S_DPOS] / (n * D);
		double loss_src = 0.;
		for (int k = 0; k < 8; k++) loss_src += T[S_COUNT + k];
		const double loss_tot = loss_src + cfg.w_aniso * L_aniso + cfg.w_vol * L_vol + cfg.w_valreg * L_valreg + cfg.w_dpos * L_dpos;
		st[GSR_ST_LOSS_TOT] = (float)loss_tot;
		st[GSR_ST_L_ANISO] = (float)L_aniso;
		st[GSR_ST_L_VOL] = (float)L_vol;
		st[GSR_ST_L_VALREG] = (float)L_valreg;
		st[GSR_ST_L_DPOS] = (float)L_dpos;
		const float cur = (float)loss_tot;
		float best = st[GSR_ST_BEST], bad = st[GSR_ST_BAD];
		if (cur < best * (1.f - cfg.sched_threshold)) { best = cur; bad = 0.f; } else bad += 1.f;
		const int decay = bad > (float)cfg.sched_patience;
		if (decay) bad = 0.f;
		st[GSR_ST_BEST] = best;
		st[GSR_ST_BAD] = bad;
		for (int g = 0; g < 4; g++) {	// the step sizes of THIS step were formed from the rates before this update (early_sm)
			const float lr = early_sm[4 + g];
			const float new_lr = fmaxf(lr * cfg.sched_factor, cfg.sched_min_lr);
			if (decay && lr - new_lr > cfg.sched_eps) st[GSR_ST_LR + g] = new_lr;
		}
	} else if (w == 2) {
		if (lane < 4) { st[C_STEP + lane] = early_sm[lane]; st[C_LRUSED + lane] = early_sm[4 + lane]; }
		else if (lane == 4) st[C_BC2] = early_sm[8];
		else if (lane == 5) { st[GSR_ST_T] = (float)((double)st[GSR_ST_T] + 1.); st[GSR_ST_CLOCK] += 1.f; }
		else if (lane == 6) {
			reinterpret_cast<double *>(st + GSR_ST_BPOW)[0] = reinterpret_cast<const double *>(early_sm + 12)[0];
			reinterpret_cast<double *>(st + GSR_ST_BPOW)[1] = reinterpret_cast<const double *>(early_sm + 12)[1];
		}
	}
}

constexpr int SC4_G = 4;
constexpr int SC4_MAX_THREADS = 512;
constexpr int SC4_MAX_N = SC_CTAS * SC4_MAX_THREADS / SC4_G;
int g_step_lanes4 = 1;	// GSR_TUNE_STEP_LANES4
int g_step_fused_hash = 1;	// GSR_TUNE_STEP_FUSED_HASH

constexpr int SC4_MAX_CELLS = 1024;	// hash cells the fused rebuild keeps in shared memory (+ 1 tail bucket)

struct HashOut {	// HASH: the rebuilt hash of gsr_build_grid, written by the step kernel itself
	Grid g;
	int32_t *cell_start, *sorted_id;
	float4 *packed;
	float *cull;
};

// HASH = true: the kernel goes on to rebuild the cell hash and the packed records from the parameters it has just updated (what
// gsr_step_rebuild otherwise appends as a second launch): a stable counting sort over the cluster — per-CTA histograms in shared
// memory, read by every CTA through distributed shared memory; a Gaussian's slot is cell_start[key] + the counts of the lower-ranked
// CTAs in its cell + the number of lower ids with the same key in its own CTA (ids ascend with CTA rank and thread), which is the
// canonical ascending-id order of every other hash path.
template <int D, bool HASH>
__global__ void __cluster_dims__(SC_CTAS, 1, 1) __launch_bounds__(SC4_MAX_THREADS, 1)
step_cluster4_kernel(gsr_step_cfg cfg, int N, float *__restrict__ pos, float *__restrict__ scal, float *__restrict__ rot, float *__restrict__ vals,
		     const float *__restrict__ acc, int sets_mask, const float *__restrict__ ex0, const float *__restrict__ ex1,
		     const float *__restrict__ pos_org, LossSrcs ls, float *st, HashOut H)
{
	namespace cg = cooperative_groups;
	constexpr int AF = Dim<D>::AF, NR = Dim<D>::NR, P = Dim<D>::P, G = SC4_G, KPL = (P + G - 1) / G;
	cg::cluster_group cluster = cg::this_cluster();
	__shared__ __align__(8) float cst[GSR_STATE_SCALARS];
	__shared__ float wpart[SC4_MAX_THREADS / 32][S_COUNT];
	__shared__ float part[S_COUNT];
	__shared__ double Tsm[S_COUNT + 8];
	__shared__ float wmin[SC4_MAX_THREADS / 32];
	__shared__ float bmin;
	__shared__ float Tg[SC4_MAX_THREADS / G][G][P + 1];	// parameter-space gradients of the four roles of every Gaussian of this CTA
	__shared__ uint32_t hist[HASH ? SC4_MAX_CELLS + 1 : 1];	// this CTA's Gaussians per cell (read by the other CTAs)
	__shared__ uint32_t hstart[HASH ? SC4_MAX_CELLS + 1 : 1];	// cell totals over the cluster, then their exclusive prefix (= cell_start)
	__shared__ uint32_t hbase[HASH ? SC4_MAX_CELLS + 1 : 1];	// Gaussians of lower-ranked CTAs per cell
	__shared__ uint32_t hkey[HASH ? SC4_MAX_THREADS / G : 1];	// key of every Gaussian of this CTA
	__shared__ uint32_t hwarp[32];
	__shared__ float gs_sm;
	__shared__ __align__(8) float early_sm[16];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
	const int rank = (int)cluster.block_rank();
	const int gl = tid / G, sub = tid % G;
	const int i = rank * (blockDim.x / G) + gl;
	const bool on = i < N;
	STAMP(0);
	for (int k = tid; k < GSR_STATE_SCALARS; k += blockDim.x) cst[k] = st[k];
	if (HASH) {
		for (int c = tid; c <= H.g.ncell; c += blockDim.x) hist[c] = 0;
	}
	// ---- every load of this lane, issued together ------------------------------------------------------------------------
	float sc[D], r[NR], a[AF], a2nd[AF];
	// role of this lane: 0 vorticity set, 1 divergence set, 2 direct set, 3 the extra direct sets (boundary passes)
	const float *src = nullptr, *src2 = nullptr;
	if (on) {
		if (sub == 0 && (sets_mask & 2)) src = acc + ((size_t)1 * N + i) * AF;
		if (sub == 1 && (sets_mask & 4)) src = acc + ((size_t)2 * N + i) * AF;
		if (sub == 2 && (sets_mask & 1)) src = acc + (size_t)i * AF;
		if (sub == 3) {
			src = ex0 ? ex0 + (size_t)i * AF : (ex1 ? ex1 + (size_t)i * AF : nullptr);
			src2 = (ex0 && ex1) ? ex1 + (size_t)i * AF : nullptr;
		}
#pragma unroll
		for (int k = 0; k < D; k++) sc[k] = scal[(size_t)D * i + k];
#pragma unroll
		for (int k = 0; k < NR; k++) r[k] = rot[(size_t)NR * i + k];
	}
#pragma unroll
	for (int k = 0; k < AF; k++) { a[k] = src ? src[k] : 0.f; a2nd[k] = src2 ? src2[k] : 0.f; }
	// the parameters this lane owns: k = sub + 4 j
	float prm[KPL], mo[KPL], vo[KPL], po[KPL];
	float *pp[KPL];
	float *m = st + GSR_STATE_SCALARS + (size_t)i * P, *vv = st + GSR_STATE_SCALARS + (size_t)N * P + (size_t)i * P;
#pragma unroll
	for (int j = 0; j < KPL; j++) {
		const int k = sub + G * j;
		pp[j] = nullptr;
		prm[j] = mo[j] = vo[j] = po[j] = 0.f;
		if (on && k < P) {
			pp[j] = k < D ? pos + (size_t)D * i + k : (k < 2 * D ? scal + (size_t)D * i + (k - D) : (k < 2 * D + NR ? rot + (size_t)NR * i + (k - 2 * D) : vals + (size_t)D * i + (k - 2 * D - NR)));
			prm[j] = *pp[j];
			mo[j] = m[k];
			vo[j] = vv[k];
			if (k < D && pos_org) po[j] = pos_org[(size_t)D * i + k];
		}
	}
	if (w == nw - 1) {	// weighted sums of the sample-loss partials: lanes 8 q + k sum slot k of source q (blocks in order, double,
		const int k = lane & 7, q = lane >> 3;	// loads eight deep); then the sources in order, as step_reduce_tail's callers do
		double t = 0.;
		if (q < ls.n) {
			const float *pt = ls.partials[q] + k;
			const int nb = ls.nblocks[q];
			int b = 0;
			for (; b + 8 <= nb; b += 8) {
				float v[8];
#pragma unroll
				for (int u = 0; u < 8; u++) v[u] = pt[(size_t)(b + u) * 8];
#pragma unroll
				for (int u = 0; u < 8; u++) t += (double)v[u];
			}
			for (; b < nb; b++) t += (double)pt[(size_t)b * 8];
		}
		double s = 0.;
#pragma unroll
		for (int qq = 0; qq < 3; qq++) {
			const double tq = __shfl_sync(0xffffffffu, t, 8 * qq + k);
			if (qq < ls.n) s += (double)ls.w[qq][k] * tq;
		}
		if (lane < 8) Tsm[S_COUNT + k] = s;
	}
	if (w == (nw >= 2 ? nw - 2 : 0)) step_tail_early(cfg, st, lane, early_sm);
	// ---- chain rule of this lane's set ---------------------------------------------------------------------------------------
	{
		float g[P];
#pragma unroll
		for (int k = 0; k < P; k++) g[k] = 0.f;
		auto to_param = [&](const float *aa, float *gg) {
			float gp[D], gs[D], gr[NR], gv[D];
			param_grad<D>(aa, sc, r, gp, gs, gr, gv);
#pragma unroll
			for (int k = 0; k < D; k++) { gg[k] = gp[k]; gg[D + k] = gs[k]; gg[2 * D + NR + k] = gv[k]; }
#pragma unroll
			for (int k = 0; k < NR; k++) gg[2 * D + k] = gr[k];
		};
		if (src) to_param(a, g);
		if (src2) {
			float t[P];
			to_param(a2nd, t);
#pragma unroll
			for (int k = 0; k < P; k++) g[k] += t[k];
		}
#pragma unroll
		for (int k = 0; k < P; k++) Tg[gl][sub][k] = g[k];
	}
	STAMP(1);
	__syncwarp();
	// ---- this lane's parameters: PCGrad dots, regulariser moments ----------------------------------------------------------------
	float S[S_COUNT];
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) S[k] = 0.f;
	float g1[KPL], g2[KPL], gd[KPL];
	float ssum = 0.f;
	int kmin = 0, kmax = 0;
#pragma unroll
	for (int k = 0; k < D; k++) {
		if (on) {
			ssum += sc[k];
			if (sc[k] < sc[kmin]) kmin = k;	// first index on ties, like torch.min / torch.max
			if (sc[k] > sc[kmax]) kmax = k;
		}
	}
	float smin_k = 0.f, smax_k = 0.f;
#pragma unroll
	for (int k = 0; k < D; k++) { if (k == kmin) smin_k = on ? sc[k] : 0.f; if (k == kmax) smax_k = on ? sc[k] : 0.f; }
	const float V = expf(-ssum), rho = expf(smax_k - smin_k);
#pragma unroll
	for (int j = 0; j < KPL; j++) {
		const int k = sub + G * j;
		g1[j] = g2[j] = gd[j] = 0.f;
		if (on && k < P) {
			const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
			g1[j] = Tg[gl][0][k];
			g2[j] = Tg[gl][1][k];
			gd[j] = Tg[gl][2][k] + Tg[gl][3][k];
			if ((sets_mask & 6) == 6) {
				// one slot per group and lane: which slot is a run-time index, so add through a select over the four groups
#pragma unroll
				for (int q = 0; q < 4; q++) {
					S[S_DOT + q] += (q == grp) ? g1[j] * g2[j] : 0.f;
					S[S_N1 + q] += (q == grp) ? g1[j] * g1[j] : 0.f;
					S[S_N2 + q] += (q == grp) ? g2[j] * g2[j] : 0.f;
				}
			}
			if (grp == 3) S[S_ABSV] += fabsf(prm[j]);
			if (grp == 0 && pos_org) { const float dlt = prm[j] - po[j]; S[S_DPOS] += dlt * dlt; }
		}
	}
	if (on && sub == 0) {
		S[S_V] += V;
		S[S_V2] += V * V;
		S[S_ANISO] += (rho >= cfg.aniso_ratio ? rho : cfg.aniso_ratio) - cfg.aniso_ratio;
	}
#pragma unroll
	for (int k = 0; k < S_COUNT; k++) {
#pragma unroll
		for (int o = 16; o; o >>= 1) S[k] += __shfl_xor_sync(0xffffffffu, S[k], o);
	}
	if (lane == 0) {
#pragma unroll
		for (int k = 0; k < S_COUNT; k++) wpart[w][k] = S[k];
	}
	__syncthreads();
	if (tid < S_COUNT) {
		float t = 0.f;
		for (int ww = 0; ww < nw; ww++) t += wpart[ww][tid];
		part[tid] = t;
	}
	STAMP(2);
	cluster.sync();
	STAMP(3);
	if (tid < S_COUNT) {	// CTAs in order, double: the same T in every CTA
		double t = 0.;
#pragma unroll
		for (int q = 0; q < SC_CTAS; q++) t += (double)*cluster.map_shared_rank(&part[tid], q);
		Tsm[tid] = t;
	}
	__syncthreads();
	if (nw >= 3) {
		step_tail_late<D>(cfg, N, Tsm, cst, early_sm);
	} else if (tid == 0) {
		double T[S_COUNT + 8];
#pragma unroll
		for (int k = 0; k < S_COUNT + 8; k++) T[k] = Tsm[k];
		step_reduce_tail<D>(cfg, N, T, cst);
	}
	__syncthreads();
	STAMP(4);
	// ---- projected gradient + regulariser gradients + Adam on this lane's parameters ----------------------------------------
	float smin = __int_as_float(0x7f800000);
	const float bc2 = early_sm[8];
	const float rV = V / cst[C_MEANV];
	const float cv = -cfg.w_vol * 2.f / (float)N * rV * (rV - cst[C_MEANR2]);
	const float ca = (rho >= cfg.aniso_ratio && kmin != kmax) ? cfg.w_aniso * rho / (float)N : 0.f;
#pragma unroll
	for (int j = 0; j < KPL; j++) {
		const int k = sub + G * j;
		if (on && k < P) {
			const int grp = k < D ? 0 : (k < 2 * D ? 1 : (k < 2 * D + NR ? 2 : 3));
			float t = 0.f;
			if (sets_mask & 2) t += cst[C_A1 + grp] * g1[j];
			if (sets_mask & 4) t += cst[C_A2 + grp] * g2[j];
			float g = t + gd[j];
			if (grp == 1) {
				const int ks = k - D;
				if (ca != 0.f) g += (ks == kmax ? ca : 0.f) - (ks == kmin ? ca : 0.f);
				g += cv;
			}
			if (grp == 3 && cfg.w_valreg != 0.f) g += cfg.w_valreg / (float)(N * D) * (float)((prm[j] > 0.f) - (prm[j] < 0.f));
			if (grp == 0 && pos_org && cfg.w_dpos != 0.f) g += cfg.w_dpos * 2.f / (float)(N * D) * (prm[j] - po[j]);
			const float mk = mo[j] + (g - mo[j]) * (1.f - cfg.beta1);
			const float vk = vo[j] * cfg.beta2 + (1.f - cfg.beta2) * g * g;
			const float pk = prm[j] - early_sm[grp] * (mk / (sqrtf(vk) * bc2 + cfg.eps));
			m[k] = mk;
			vv[k] = vk;
			*pp[j] = pk;
			if (HASH) Tg[gl][0][k] = pk;	// (the gradients were consumed above) the updated record, for the key and the packed record
			if (grp == 1) smin = fminf(smin, pk);
		}
	}
	STAMP(5);
#pragma unroll
	for (int o = 16; o; o >>= 1) smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
	if (lane == 0) wmin[w] = smin;
	__syncthreads();
	if (tid == 0) {
		float mn = wmin[0];
		for (int ww = 1; ww < nw; ww++) mn = fminf(mn, wmin[ww]);
		bmin = mn;
	}
	cluster.sync();
	STAMP(6);
	if (rank == 0) {
		for (int k = tid; k < GSR_STATE_SCALARS; k += blockDim.x)
			if (k != GSR_ST_GRID_SCALE && k != GSR_ST_MIN_S && k != GSR_ST_SGS_ERR) st[k] = cst[k];
	}
	if (tid == 0 && (rank == 0 || HASH)) {	// the next grid_scale: every CTA needs it for the keys, CTA 0 publishes it
		float mn = cst[GSR_ST_MIN_S];
#pragma unroll
		for (int q = 0; q < SC_CTAS; q++) mn = fminf(mn, *cluster.map_shared_rank(&bmin, q));
		double gs = cfg.grid_scale_tau0;
		if (cfg.grid_coef > 0.) gs = fmax(cfg.grid_coef * exp(-(double)mn), cfg.min_grid_scale);
		gs_sm = (float)gs;
		if (rank == 0) step_grid_scale(cfg, st, mn, (int)cst[GSR_ST_T]);
	}
	if (HASH) {
		const Grid &g = H.g;
		const int ncell = g.ncell;
		__syncthreads();	// gs_sm (the Tg rows of the updated parameters were written by this warp's own lanes)
		uint32_t key = 0;
		if (sub == 0) {
			if (on) {
				float pt[3] = {0.f, 0.f, 0.f};
#pragma unroll
				for (int k = 0; k < D; k++) pt[k] = Tg[gl][0][k];
				key = gauss_key_vals<D>(pt, g, gs_sm);
				atomicAdd(&hist[key], 1u);
			}
			hkey[gl] = on ? key : 0xffffffffu;	// (not a Gaussian: matches no key)
		}
		STAMP(7);
		cluster.sync();	// every CTA's histogram is complete
		STAMP(8);
		for (int c = tid; c <= ncell; c += blockDim.x) {
			uint32_t tot = 0, base = 0;
#pragma unroll
			for (int q = 0; q < SC_CTAS; q++) {
				const uint32_t h = *cluster.map_shared_rank(&hist[c], q);
				base += (q < rank) ? h : 0u;
				tot += h;
			}
			hstart[c] = tot;
			hbase[c] = base;
		}
		__syncthreads();
		{	// exclusive prefix of hstart[0..ncell]: a chunk per thread, warp scan, scan of the warp sums
			const int mm = ncell + 1, chunk = (mm + (int)blockDim.x - 1) / (int)blockDim.x;
			const int b = min(tid * chunk, mm), e = min(b + chunk, mm);
			uint32_t sum = 0;
			for (int c = b; c < e; c++) sum += hstart[c];
			uint32_t v = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
				if (lane >= o) v += t;
			}
			if (lane == 31) hwarp[w] = v;
			__syncthreads();
			if (w == 0) {
				const uint32_t ws = lane < nw ? hwarp[lane] : 0u;
				uint32_t t2 = ws;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t t = __shfl_up_sync(0xffffffffu, t2, o);
					if (lane >= o) t2 += t;
				}
				hwarp[lane] = t2 - ws;
			}
			__syncthreads();
			uint32_t run = hwarp[w] + (v - sum);
			for (int c = b; c < e; c++) {
				const uint32_t t = hstart[c];
				hstart[c] = run;
				if (rank == 0) H.cell_start[c] = (int32_t)run;
				run += t;
			}
		}
		__syncthreads();
		STAMP(9);
		// slot of this Gaussian: its four lanes count the lower ids of the CTA with the same key, a quarter of the range each
		const int l0 = lane & ~(G - 1);
		key = __shfl_sync(0xffffffffu, key, l0);
		uint32_t lower = 0;
		for (int q = sub; q < gl; q += G) lower += (hkey[q] == key) ? 1u : 0u;
		lower += __shfl_xor_sync(0xffffffffu, lower, 1);
		lower += __shfl_xor_sync(0xffffffffu, lower, 2);
		// the double-precision exponentials (3D) / exponentials, cosine and sine (2D) of the record, one per lane
		float piece = 0.f;
		if (on) {
			if (D == 3) {
				if (sub < 3) piece = exp2s(Tg[gl][0][D + sub]);
			} else {
				const double th = (double)Tg[gl][0][2 * D];
				piece = sub < 2 ? exp2s(Tg[gl][0][D + sub]) : (sub == 2 ? (float)cos(th) : (float)sin(th));
			}
		}
		const float e0 = __shfl_sync(0xffffffffu, piece, l0), e1 = __shfl_sync(0xffffffffu, piece, l0 + 1), e2 = __shfl_sync(0xffffffffu, piece, l0 + 2),
			    e3 = __shfl_sync(0xffffffffu, piece, l0 + 3);
		if (on && sub == 0) {
			const int slot = (int)(hstart[key] + hbase[key] + lower);
			H.sorted_id[slot] = i;
			const float *q = Tg[gl][0];
			if (D == 3) {
				const float S3[3] = {e0, e1, e2};
				pack3d_core(q, q + D, S3, make_float4(q[2 * D], q[2 * D + 1], q[2 * D + 2], q[2 * D + 3]), q + 2 * D + NR, slot, H.packed, H.cull);
			} else {
				pack2d_core(q, q + D, e0, e1, e2, e3, q + 2 * D + NR, slot, H.packed, H.cull);
				(void)e3;
			}
		}
	}
	STAMP(10);
	cluster.sync();	// no CTA may exit while another still reads its shared memory
	STAMP(11);
}

__global__ void min_into_kernel(const float *__restrict__ a, size_t n, float *out)
{
	float m = __int_as_float(0x7f800000);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fminf(m, a[i]);
#pragma unroll
	for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) atomic_min_f(out, m);
}

}  // namespace gsr

using namespace gsr;

#ifdef GSR_STEP_TIMING
extern "C" int gsr_debug_step_stamps(unsigned long long *out)
{
	return (int)cudaMemcpyFromSymbol(out, g_step_stamps, sizeof(unsigned long long) * 16);
}
#endif

extern "C" size_t gsr_step_state_floats(int D, int64_t N) { return GSR_STATE_SCALARS + 2 * (size_t)(D == 3 ? 13 : 7) * (size_t)N; }

extern "C" size_t gsr_step_ws_bytes(int, int64_t N)
{
	size_t nblk = (size_t)((N + ST_THREADS - 1) / ST_THREADS);
	return sizeof(float) * S_COUNT * (nblk ? nblk : 1);
}

extern "C" int gsr_step_init(const gsr_step_cfg *cfg, int64_t N, const float *scalings, float *state, void *stream)
{
	if (!cfg || (cfg->D != 2 && cfg->D != 3) || N < 0 || !state) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	size_t nm = 2 * (size_t)(cfg->D == 3 ? 13 : 7) * (size_t)N;
	size_t total = nm > GSR_STATE_SCALARS ? nm : GSR_STATE_SCALARS;
	init_state_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(state, nm, *cfg);
	if (scalings && N > 0) {
		size_t n = (size_t)N * cfg->D;
		int blocks = (int)((n + 1023) / 1024);
		if (blocks > kSMs * 8) blocks = kSMs * 8;
		min_into_kernel<<<blocks, 256, 0, st>>>(scalings, n, state + GSR_ST_MIN_S);
		stepS_kernel<<<1, 1, 0, st>>>(*cfg, state, nullptr);
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

static int step_impl(const gsr_step_cfg *cfg, int64_t N, float *positions, float *scalings, float *rotations, float *values,
		     const float *acc, int sets_mask, const float *const extra_direct[2], const gsr_loss_src *loss_src, int n_loss_src,
		     const float *positions_org, float *state, void *ws, size_t ws_bytes,
		     const gsr_grid_desc *gd, int32_t *cell_start, int32_t *sorted_id, float *packed, float *cull, void *hash_ws, size_t hash_ws_bytes, void *stream)
{
	if (!cfg || (cfg->D != 2 && cfg->D != 3) || N <= 0 || N >= ((int64_t)1 << 30) || !state || !positions || !scalings || !rotations || !values) return GSR_EINVAL;
	if ((sets_mask & 7) && !acc) return GSR_EINVAL;
	if (n_loss_src < 0 || n_loss_src > 3) return GSR_EINVAL;
	if (ws_bytes < gsr_step_ws_bytes(cfg->D, N)) return GSR_EWS;
	Grid g;
	if (gd) {
		if (!make_grid(gd, g) || g.D != cfg->D || !cell_start || !sorted_id || !packed) return GSR_EINVAL;
		if (hash_ws_bytes < gsr_build_grid_ws_bytes(gd, N)) return GSR_EWS;
	}
	cudaStream_t st = (cudaStream_t)stream;
	const int n = (int)N, nblk = (n + ST_THREADS - 1) / ST_THREADS;
	float *partials = (float *)ws;
	LossSrcs ls;
	ls.n = n_loss_src;
	for (int s = 0; s < 3; s++) {
		ls.partials[s] = (s < n_loss_src) ? loss_src[s].partials : nullptr;
		ls.nblocks[s] = (s < n_loss_src) ? loss_src[s].nblocks : 0;
		for (int k = 0; k < 8; k++) ls.w[s][k] = (s < n_loss_src) ? loss_src[s].w[k] : 0.f;
	}
	const float *ex0 = extra_direct ? extra_direct[0] : nullptr, *ex1 = extra_direct ? extra_direct[1] : nullptr;
	if (N <= g_step_small_n && N <= SC4_MAX_N && g_step_lanes4) {
		const int bt = ((((n + SC_CTAS - 1) / SC_CTAS) * SC4_G) + 31) & ~31;	// Gaussians per CTA x 4 lanes, whole warps
		HashOut H;
		H.cell_start = nullptr;
		if (gd && packed && g_step_fused_hash && g.ncell <= SC4_MAX_CELLS) {	// step + hash rebuild + packed records: one launch
			H.g = g; H.cell_start = cell_start; H.sorted_id = sorted_id; H.packed = (float4 *)packed; H.cull = cull;
			if (cfg->D == 3) step_cluster4_kernel<3, true><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state, H);
			else step_cluster4_kernel<2, true><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state, H);
			g_launches += 1;
			GSR_CHECK_LAUNCH();
			return GSR_OK;
		}
		if (cfg->D == 3) step_cluster4_kernel<3, false><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state, H);
		else step_cluster4_kernel<2, false><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state, H);
		g_launches += 1;
	} else if (N <= g_step_small_n && N <= SC_MAX_N) {
		const int bt = (((n + SC_CTAS - 1) / SC_CTAS) + 31) & ~31;
		if (cfg->D == 3) step_cluster_kernel<3><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state);
		else step_cluster_kernel<2><<<SC_CTAS, bt, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, ls, state);
		g_launches += 1;
	} else {
		if (cfg->D == 3) {
			stepA_kernel<3><<<nblk, ST_THREADS, 0, st>>>(n, positions, scalings, rotations, values, acc, sets_mask, positions_org, cfg->aniso_ratio, partials);
			stepR_kernel<3><<<1, 256, 0, st>>>(*cfg, n, nblk, partials, ls, state);
			stepB_kernel<3><<<nblk, ST_THREADS, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, state);
		} else {
			stepA_kernel<2><<<nblk, ST_THREADS, 0, st>>>(n, positions, scalings, rotations, values, acc, sets_mask, positions_org, cfg->aniso_ratio, partials);
			stepR_kernel<2><<<1, 256, 0, st>>>(*cfg, n, nblk, partials, ls, state);
			stepB_kernel<2><<<nblk, ST_THREADS, 0, st>>>(*cfg, n, positions, scalings, rotations, values, acc, sets_mask, ex0, ex1, positions_org, state);
		}
		stepS_kernel<<<1, 1, 0, st>>>(*cfg, state, nullptr);
		g_launches += 4;
	}
	GSR_CHECK_LAUNCH();
	if (gd) return gsr_build_grid(gd, positions, N, cell_start, sorted_id, nullptr, nullptr, scalings, rotations, values, packed, cull, hash_ws, hash_ws_bytes, stream);
	return GSR_OK;
}

extern "C" int gsr_step(const gsr_step_cfg *cfg, int64_t N, float *positions, float *scalings, float *rotations, float *values,
			const float *acc, int sets_mask, const float *const extra_direct[2], const gsr_loss_src *loss_src, int n_loss_src,
			const float *positions_org, float *state, void *ws, size_t ws_bytes, void *stream)
{
	return step_impl(cfg, N, positions, scalings, rotations, values, acc, sets_mask, extra_direct, loss_src, n_loss_src, positions_org, state, ws, ws_bytes,
			 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int gsr_step_rebuild(const gsr_step_cfg *cfg, int64_t N, float *positions, float *scalings, float *rotations, float *values,
				const float *acc, int sets_mask, const float *const extra_direct[2], const gsr_loss_src *loss_src, int n_loss_src,
				const float *positions_org, float *state, void *ws, size_t ws_bytes,
				const gsr_grid_desc *g, int32_t *cell_start, int32_t *sorted_id, float *packed, float *cull, void *hash_ws, size_t hash_ws_bytes,
				void *stream)
{
	if (!g) return GSR_EINVAL;
	return step_impl(cfg, N, positions, scalings, rotations, values, acc, sets_mask, extra_direct, loss_src, n_loss_src, positions_org, state, ws, ws_bytes,
			 g, cell_start, sorted_id, packed, cull, hash_ws, hash_ws_bytes, stream);
}
