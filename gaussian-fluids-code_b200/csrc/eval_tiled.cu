// eval_tiled.cu — the throughput shape of the forward / RK4 kernels for large Q.
//
// Measured on B200 (profiles/): with one thread per point and one 48-byte record load per (point, candidate) pair
// the kernels of eval.cu are bound by the L1/shared-memory return path, not by FP32 issue — a warp-wide 128-bit load
// occupies the SM's load pipe for 4 cycles even when all lanes read the same address, i.e. 12 cycles per candidate per
// warp against ~5 cycles of arithmetic.  This file removes that bound and the divergence on the truncation test:
//
//   * REGISTER TILING: every thread owns P points (P = 4 forward, 2 in the RK4 pull-back), so one candidate record,
//     loaded once per warp, is tested against 32 P points;
//   * WARP-UNIFORM CANDIDATE STREAM: a warp walks the hull of its points' (clamped) 27-cell stencils cell by cell; a
//     point that does not have the current cell in its own stencil gets the threshold -1 for that cell (q >= 0 is never
//     accepted), so each point still sees exactly the reference's candidate list, in the reference's order;
//   * COHERENT WARPS: gsr_bin_samples sorts the samples by (cell, 4x4x4 sub-cell), so the 32 P points of a warp occupy
//     a small part of one cell and most candidates are rejected by the whole warp — the accepted branch
//     (ex2 + 28 flop) is then skipped by a uniform branch instead of running under a mostly idle warp;
//   * TMA STAGING: a CTA owns a TILE of up to 512 sorted samples of one (x, y) row of cells.  The Gaussians of the
//     3 x 3 neighbouring rows over the tile's z range (+-1 cell) are 9 CONTIGUOUS runs of packed records (the hash is
//     cell-sorted, z fastest): one lane per run issues a TMA bulk copy (cp.async.bulk global -> shared, completion on
//     an mbarrier).  The staged box serves every point of the tile and, in the RK4 kernels, all 4-5 evaluations.
//     Cells outside the staged box (an RK4 stage point that drifted into another row, a box above the staging
//     capacity) are read from global memory through the same loop.
#include "eval.cuh"
#include "f32x2.cuh"
#include <limits.h>

namespace gsr {

constexpr int TL_TILE = GSR_TILE_SAMPLES;	// samples per tile (CTA)

// ---- mbarrier / TMA bulk copy (PTX) ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
		     "r"(smem_u32(bar))
		     : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}

struct TiledArgs {
	EvalParams P;
	const int32_t *cell_start;
	const float4 *packed;
	const float *cull;	// per Gaussian (cell order): (1 + margin) / lambda_min(Sigma^-1), or NULL (no culling)
	const float *x;
	int Q;
	const int32_t *perm;
	const int32_t *scs;	// sample_cell_start on the padded grid
	const int32_t *tile_row;
	int cap;		// staging capacity in Gaussians
};

struct TileSh {
	int soff[12];	// first shared-memory slot of each staged run
	int gstart[12];	// first global (cell-sorted) index of each staged run
	int tcx, tcy, zlo, zhi;
	int staged, total;
};

// 128-bit generic load (shared or global)
__device__ __forceinline__ float4 ld4(const float4 *p)
{
	float4 v;
	asm("ld.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
	return v;
}

// Tile lookup: the row and the range [t0, t1) of sorted samples of this CTA.  Returns false for an unused tile slot.
__device__ __forceinline__ bool tile_locate3(const TiledArgs &a, int &r, int &t0, int &t1)
{
	const Grid &g = a.P.g;
	r = __ldg(a.tile_row + blockIdx.x);
	if (r < 0) return false;
	const int rowlen = g.pdims[2];
	const int nrows = g.pdims[0] * g.pdims[1];
	const int s_r = __ldg(a.scs + (r < nrows ? r * rowlen : g.pcell));
	const int e_r = (r < nrows) ? __ldg(a.scs + (r + 1) * rowlen) : a.Q;
	const int base = s_r / TL_TILE + r;
	t0 = s_r + ((int)blockIdx.x - base) * TL_TILE;
	t1 = min(t0 + TL_TILE, e_r);
	return true;
}

// Staging box and TMA bulk copies of the tile (the threads load their points before calling this, so that those
// loads are in flight while warp 0 walks the dependent chain perm -> x -> cell_start -> bulk copy).
__device__ __forceinline__ void tile_stage3(const TiledArgs &a, int r, int t0, int t1, TileSh &sh, float4 *srec, uint64_t *mbar)
{
	const Grid &g = a.P.g;
	const int nrows = g.pdims[0] * g.pdims[1];
	if (threadIdx.x < 32) {
		const int lane = threadIdx.x;
		const bool stage = r < nrows;	// the tail "row" holds the samples outside the padded grid: nothing to stage
		const int tcx = r / g.pdims[1] - 1, tcy = r % g.pdims[1] - 1;
		const float gs = grid_gs(g);
		const size_t j0 = (size_t)__ldg(a.perm + t0), j1 = (size_t)__ldg(a.perm + t1 - 1);
		const int c0 = cell_coord(__ldg(a.x + 3 * j0 + 2), g.lo[2], gs), c1 = cell_coord(__ldg(a.x + 3 * j1 + 2), g.lo[2], gs);
		const int Zlo = max(c0 - 1, 0), Zhi = min(c1 + 1, g.dims[2] - 1);
		const int gi = tcx - 1 + lane / 3, gj = tcy - 1 + lane % 3;
		int cnt = 0, g0 = 0;
		if (lane < 9 && stage && Zlo <= Zhi && gi >= 0 && gi < g.dims[0] && gj >= 0 && gj < g.dims[1]) {
			const int cb = (gi * g.dims[1] + gj) * g.dims[2];
			g0 = __ldg(a.cell_start + cb + Zlo);
			cnt = __ldg(a.cell_start + cb + Zhi + 1) - g0;
		}
		int incl = cnt;
#pragma unroll
		for (int o = 1; o < 16; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		const int total = __shfl_sync(0xffffffffu, incl, 8);
		const bool staged = stage && total > 0 && total <= a.cap;
		if (lane < 9) {
			sh.soff[lane] = incl - cnt;
			sh.gstart[lane] = g0;
		}
		if (lane == 0) {
			sh.tcx = tcx; sh.tcy = tcy; sh.zlo = Zlo; sh.zhi = Zhi;
			sh.staged = staged ? 1 : 0;
			sh.total = total;
			if (staged) {
				mbar_init(mbar, 1);
				fence_mbar_init();
				mbar_expect_tx(mbar, (uint32_t)total * 48u);
			}
		}
		__syncwarp();
		if (staged && lane < 9 && cnt > 0) bulk_g2s(srec + 3 * (incl - cnt), a.packed + 3 * (size_t)g0, (uint32_t)cnt * 48u, mbar);
	}
	__syncthreads();
	if (sh.staged) mbar_wait(mbar, 0);
}

// order-preserving float <-> int map (for redux.sync min / max on floats)
__device__ __forceinline__ int f2ord(float f)
{
	const int i = __float_as_int(f);
	return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// u (and grad u) at the P points of every lane of a warp.  WARP-COLLECTIVE: all 32 lanes call it; ok[p] = false
// for a slot without a point.  Each point sees exactly the occupants of its own clamped 27-cell stencil, in
// cell-sorted order (the order of eval_point3), and every sum is formed exactly as eval_point3 forms it.
//   * warp-level culling: a candidate whose truncation ellipsoid (bounded by the sphere of radius^2 q_thr * cull[i],
//     cull = (1 + margin) / lambda_min(Sigma^-1), from gsr_pack_gaussians) misses the bounding box of the warp's 32 P
//     points is skipped by a uniform branch before its covariance is even loaded;
//   * the survivors are tested on point PAIRS with packed FP32 (f32x2.cuh): 14 FFMA2/FMUL2/FADD2 per two points.
template <int P, bool NEED_GRAD, bool PREFETCH>
__device__ __forceinline__ void warp_eval3(const TiledArgs &a, const TileSh &sh, const float4 *srec, const float (&x)[P], const float (&y)[P],
					   const float (&z)[P], const bool (&ok)[P], float (&u)[P][3], float (&G)[P][9])
{
	static_assert(P % 2 == 0, "points are processed in packed pairs");
	constexpr int H = P / 2;
	const Grid &g = a.P.g;
	const unsigned FULL = 0xffffffffu;
	const float gs = grid_gs(g);
	const float q_thr = a.P.q_thr, tau = g.tau;
	int cx[P], cy[P], cz[P];
	int hx0 = INT_MAX, hx1 = INT_MIN, hy0 = INT_MAX, hy1 = INT_MIN, hz0 = INT_MAX, hz1 = INT_MIN;
	const float INF = __int_as_float(0x7f800000);
	float bx0 = INF, bx1 = -INF, by0 = INF, by1 = -INF, bz0 = INF, bz1 = -INF;	// bounding box of the active points
#pragma unroll
	for (int p = 0; p < P; p++) {
		cx[p] = cell_coord(x[p], g.lo[0], gs);
		cy[p] = cell_coord(y[p], g.lo[1], gs);
		cz[p] = cell_coord(z[p], g.lo[2], gs);
		const int x0 = max(cx[p] - 1, 0), x1 = min(cx[p] + 1, g.dims[0] - 1);
		const int y0 = max(cy[p] - 1, 0), y1 = min(cy[p] + 1, g.dims[1] - 1);
		const int z0 = max(cz[p] - 1, 0), z1 = min(cz[p] + 1, g.dims[2] - 1);
		if (ok[p] && x0 <= x1 && y0 <= y1 && z0 <= z1) {
			hx0 = min(hx0, x0); hx1 = max(hx1, x1);
			hy0 = min(hy0, y0); hy1 = max(hy1, y1);
			hz0 = min(hz0, z0); hz1 = max(hz1, z1);
			bx0 = fminf(bx0, x[p]); bx1 = fmaxf(bx1, x[p]);
			by0 = fminf(by0, y[p]); by1 = fmaxf(by1, y[p]);
			bz0 = fminf(bz0, z[p]); bz1 = fmaxf(bz1, z[p]);
		} else {
			cx[p] = -(1 << 24);	// never within one cell of a grid row: the point sees no candidate
		}
	}
	f2 U[H][3], GG[H][9];
	const f2 zero2 = pack2(0.f, 0.f);
#pragma unroll
	for (int h = 0; h < H; h++) {
		U[h][0] = U[h][1] = U[h][2] = zero2;
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 9; k++) GG[h][k] = zero2;
		}
	}
	hx0 = __reduce_min_sync(FULL, hx0); hx1 = __reduce_max_sync(FULL, hx1);
	hy0 = __reduce_min_sync(FULL, hy0); hy1 = __reduce_max_sync(FULL, hy1);
	hz0 = __reduce_min_sync(FULL, hz0); hz1 = __reduce_max_sync(FULL, hz1);
	if (hx0 <= hx1) {	// some point of this warp has a non-empty stencil (warp-uniform)
		bx0 = ord2f(__reduce_min_sync(FULL, f2ord(bx0))); bx1 = ord2f(__reduce_max_sync(FULL, f2ord(bx1)));
		by0 = ord2f(__reduce_min_sync(FULL, f2ord(by0))); by1 = ord2f(__reduce_max_sync(FULL, f2ord(by1)));
		bz0 = ord2f(__reduce_min_sync(FULL, f2ord(bz0))); bz1 = ord2f(__reduce_max_sync(FULL, f2ord(bz1)));
		f2 X2[H], Y2[H], Z2[H];
#pragma unroll
		for (int h = 0; h < H; h++) {
			X2[h] = pack2(x[2 * h], x[2 * h + 1]);
			Y2[h] = pack2(y[2 * h], y[2 * h + 1]);
			Z2[h] = pack2(z[2 * h], z[2 * h + 1]);
		}
		const f2 ntau2 = bc(-tau);
		const int lane = threadIdx.x & 31;
		// do all active points of the warp sit in ONE cell?  (then every hull cell is in every point's stencil)
		int k0[3] = {INT_MAX, INT_MAX, INT_MAX}, k1[3] = {INT_MIN, INT_MIN, INT_MIN};
		float thr[P];
#pragma unroll
		for (int p = 0; p < P; p++) {
			const bool act = cx[p] != -(1 << 24);
			thr[p] = act ? q_thr : -1.f;
			if (act) {
				k0[0] = min(k0[0], cx[p]); k1[0] = max(k1[0], cx[p]);
				k0[1] = min(k0[1], cy[p]); k1[1] = max(k1[1], cy[p]);
				k0[2] = min(k0[2], cz[p]); k1[2] = max(k1[2], cz[p]);
			}
		}
		bool mixed = false;
#pragma unroll
		for (int k = 0; k < 3; k++) mixed |= __reduce_min_sync(FULL, k0[k]) != __reduce_max_sync(FULL, k1[k]);
		for (int gi = hx0; gi <= hx1; gi++) {
			for (int gj = hy0; gj <= hy1; gj++) {
				bool rowok[P];
				if (mixed) {
					bool any_row = false;
#pragma unroll
					for (int p = 0; p < P; p++) {
						rowok[p] = abs(gi - cx[p]) <= 1 && abs(gj - cy[p]) <= 1;
						any_row |= rowok[p];
					}
					if (!__any_sync(FULL, any_row)) continue;
				}
				const int cb = (gi * g.dims[1] + gj) * g.dims[2];
				const int dx_ = gi - sh.tcx, dy_ = gj - sh.tcy;
				const int rr = (sh.staged && abs(dx_) <= 1 && abs(dy_) <= 1) ? (dx_ + 1) * 3 + (dy_ + 1) : -1;
				const int soff = rr >= 0 ? sh.soff[rr] - sh.gstart[rr] : 0;	// staged slot of sorted index c: soff + c
				for (int zg = hz0; zg <= hz1; zg += 4) {	// groups of up to 4 z cells: one contiguous run of records
					const int nz = min(4, hz1 - zg + 1);
					const int b0 = __ldg(a.cell_start + cb + zg), b1 = __ldg(a.cell_start + cb + zg + min(1, nz)),
						  b2 = __ldg(a.cell_start + cb + zg + min(2, nz)), b3 = __ldg(a.cell_start + cb + zg + min(3, nz)),
						  b4 = __ldg(a.cell_start + cb + zg + nz);
					unsigned cellmask = 0xfu;	// cells of the group that are in some active point's stencil
					if (mixed) {
						cellmask = 0;
#pragma unroll
						for (int k = 0; k < 4; k++) {
							bool act = false;
#pragma unroll
							for (int p = 0; p < P; p++) act |= rowok[p] && abs(zg + k - cz[p]) <= 1;
							cellmask |= __any_sync(FULL, act) ? (1u << k) : 0u;
						}
					}
					for (int c0 = b0; c0 < b4; c0 += 32) {
						// lane-parallel culling: lane l looks at candidate c0 + l
						const int c = c0 + lane;
						const int kc = (c >= b1) + (c >= b2) + (c >= b3);
						const bool st_c = rr >= 0 && zg + kc >= sh.zlo && zg + kc <= sh.zhi;
						bool keep = c < b4 && ((cellmask >> kc) & 1u);
						if (keep && a.cull) {
							const float4 m = ld4(st_c ? srec + 3 * (soff + c) : a.packed + 3 * (size_t)c);
							const float ex = fmaxf(fmaxf(bx0 - m.x, m.x - bx1), 0.f);
							const float ey = fmaxf(fmaxf(by0 - m.y, m.y - by1), 0.f);
							const float ez = fmaxf(fmaxf(bz0 - m.z, m.z - bz1), 0.f);
							keep = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) <= q_thr * __ldg(a.cull + c);
						}
						unsigned mask = __ballot_sync(FULL, keep);
						const unsigned smask = __ballot_sync(FULL, st_c);
						// survivors, in cell-sorted order; everything below is warp-uniform.  The record of the NEXT survivor
						// is requested before the current one is evaluated (the loads are the only long-latency step left).
						int ci = 0;
						float4 n0, n1, n2;
						if (mask) {
							const int l = __ffs(mask) - 1;
							ci = c0 + l;
							const float4 *ptr = ((smask >> l) & 1u) ? srec + 3 * (soff + ci) : a.packed + 3 * (size_t)ci;
							n0 = ld4(ptr); n1 = ld4(ptr + 1); n2 = ld4(ptr + 2);
						}
						while (mask) {
							const float4 p0 = n0, p1 = n1, p2 = n2;
							const int ci_cur = ci;
							mask &= mask - 1;
							if (PREFETCH && mask) {
								const int l = __ffs(mask) - 1;
								ci = c0 + l;
								const float4 *ptr = ((smask >> l) & 1u) ? srec + 3 * (soff + ci) : a.packed + 3 * (size_t)ci;
								n0 = ld4(ptr); n1 = ld4(ptr + 1); n2 = ld4(ptr + 2);
							}
							if (mixed) {
								const int zci = zg + (ci_cur >= b1) + (ci_cur >= b2) + (ci_cur >= b3);
#pragma unroll
								for (int p = 0; p < P; p++) thr[p] = (rowok[p] && abs(zci - cz[p]) <= 1) ? q_thr : -1.f;
							}
							// test all pairs first, then ONE branch for the whole survivor: its accepted block runs the pairs'
							// chains interleaved (a pair or a half that is not accepted contributes exact zeros)
							f2 wx[H], wy[H], wz[H];
							float g0[H], g1[H];
							bool any = false;
#pragma unroll
							for (int h = 0; h < H; h++) {
								const f2 dx = add2(X2[h], bc(-p0.x)), dy = add2(Y2[h], bc(-p0.y)), dz = add2(Z2[h], bc(-p0.z));
								wx[h] = fma2(bc(p1.z), dz, fma2(bc(p1.y), dy, mul2(bc(p1.x), dx)));
								wy[h] = fma2(bc(p2.y), dz, fma2(bc(p2.x), dy, mul2(bc(p1.y), dx)));
								wz[h] = fma2(bc(p2.z), dz, fma2(bc(p2.y), dy, mul2(bc(p1.z), dx)));
								const f2 q2 = fma2(dz, wz[h], fma2(dy, wy[h], mul2(dx, wx[h])));
								unpack2(q2, g0[h], g1[h]);
								any |= g0[h] <= thr[2 * h] || g1[h] <= thr[2 * h + 1];
							}
							if (any) {
#pragma unroll
								for (int h = 0; h < H; h++) {
									const bool a0 = g0[h] <= thr[2 * h], a1 = g1[h] <= thr[2 * h + 1];
									const float e0 = ex2_approx(g0[h] * kNegHalfLog2e), e1 = ex2_approx(g1[h] * kNegHalfLog2e);
									const float ga = a0 ? e0 : 0.f, gb = a1 ? e1 : 0.f;
									f2 gm = add2(pack2(ga, gb), ntau2);
									float m0, m1;
									unpack2(gm, m0, m1);
									gm = pack2(a0 ? m0 : 0.f, a1 ? m1 : 0.f);
									U[h][0] = fma2(bc(p0.w), gm, U[h][0]);
									U[h][1] = fma2(bc(p1.w), gm, U[h][1]);
									U[h][2] = fma2(bc(p2.w), gm, U[h][2]);
									if (NEED_GRAD) {
										const f2 ng = pack2(-ga, -gb);
										const f2 ax = mul2(ng, wx[h]), ay = mul2(ng, wy[h]), az = mul2(ng, wz[h]);
										GG[h][0] = fma2(bc(p0.w), ax, GG[h][0]); GG[h][1] = fma2(bc(p0.w), ay, GG[h][1]); GG[h][2] = fma2(bc(p0.w), az, GG[h][2]);
										GG[h][3] = fma2(bc(p1.w), ax, GG[h][3]); GG[h][4] = fma2(bc(p1.w), ay, GG[h][4]); GG[h][5] = fma2(bc(p1.w), az, GG[h][5]);
										GG[h][6] = fma2(bc(p2.w), ax, GG[h][6]); GG[h][7] = fma2(bc(p2.w), ay, GG[h][7]); GG[h][8] = fma2(bc(p2.w), az, GG[h][8]);
									}
								}
							}
							if (!PREFETCH && mask) {
								const int l = __ffs(mask) - 1;
								ci = c0 + l;
								const float4 *ptr = ((smask >> l) & 1u) ? srec + 3 * (soff + ci) : a.packed + 3 * (size_t)ci;
								n0 = ld4(ptr); n1 = ld4(ptr + 1); n2 = ld4(ptr + 2);
							}
						}
					}
				}
			}
		}
	}
#pragma unroll
	for (int h = 0; h < H; h++) {
#pragma unroll
		for (int k = 0; k < 3; k++) unpack2(U[h][k], u[2 * h][k], u[2 * h + 1][k]);
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 9; k++) unpack2(GG[h][k], G[2 * h][k], G[2 * h + 1][k]);
		}
	}
}

extern __shared__ __align__(16) unsigned char tl_smem[];

// sorted position of slot p of this thread: a warp's slot p is a run of 32 consecutive sorted samples
template <int P>
__device__ __forceinline__ int slot_pos(int t0, int p)
{
	return t0 + (int)(threadIdx.x >> 5) * (32 * P) + p * 32 + (int)(threadIdx.x & 31);
}

template <int P, bool NEED_VAL, bool NEED_GRAD, bool ACCUM>
__global__ void __launch_bounds__(TL_TILE / P, P == 4 ? (NEED_GRAD ? 4 : 5) : 2) forward_tiled3_kernel(TiledArgs a, float *__restrict__ val, float *__restrict__ grad)
{
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float x[P], y[P], z[P], u[P][3], G[P][9];
	bool ok[P];
	size_t j[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		j[p] = ok[p] ? (size_t)a.perm[t] : 0;
		x[p] = ok[p] ? a.x[3 * j[p]] : 0.f;
		y[p] = ok[p] ? a.x[3 * j[p] + 1] : 0.f;
		z[p] = ok[p] ? a.x[3 * j[p] + 2] : 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	warp_eval3<P, NEED_GRAD, NEED_GRAD>(a, sh, srec, x, y, z, ok, u, G);
#pragma unroll
	for (int p = 0; p < P; p++) {
		if (!ok[p]) continue;
		if (NEED_VAL) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float *o = val + 3 * j[p] + k;
				*o = ACCUM ? *o + u[p][k] : u[p][k];
			}
		}
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 9; k++) {
				float *o = grad + 9 * j[p] + k;
				*o = ACCUM ? *o + G[p][k] : G[p][k];
			}
		}
	}
}

// RK4 (modes as rk4_3d_kernel in eval.cu): one staged box serves all 4-5 evaluations of the tile's points.
template <int MODE, int P>
__global__ void __launch_bounds__(TL_TILE / P) rk4_tiled3_kernel(TiledArgs a, float dt, float *__restrict__ goal_pos, float *__restrict__ deformation,
								 float *__restrict__ goal_val, float *__restrict__ goal_grad, float *__restrict__ ref_vor,
								 float *__restrict__ ref_hel)
{
	constexpr bool FULLM = MODE != 0;
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float x0[P], x1[P], x2[P], px[P], py[P], pz[P], v[P][3], dv[P][9], vs[P][3], A[P][9], S[P][9];
	bool ok[P];
	size_t j[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		j[p] = ok[p] ? (size_t)a.perm[t] : 0;
		px[p] = x0[p] = ok[p] ? a.x[3 * j[p]] : 0.f;
		py[p] = x1[p] = ok[p] ? a.x[3 * j[p] + 1] : 0.f;
		pz[p] = x2[p] = ok[p] ? a.x[3 * j[p] + 2] : 0.f;
		vs[p][0] = vs[p][1] = vs[p][2] = 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	const float hdt = dt * .5f, dt6 = dt / 6.f;
#pragma unroll 1
	for (int st = 0; st < 4; st++) {
		warp_eval3<P, FULLM, false>(a, sh, srec, px, py, pz, ok, v, dv);
		const float wgt = (st == 0 || st == 3) ? 1.f : 2.f;	// RK4 weights 1 2 2 1
		const float step = (st < 2) ? hdt : dt;			// the NEXT stage point: x + step * v
#pragma unroll
		for (int p = 0; p < P; p++) {
			vs[p][0] += wgt * v[p][0]; vs[p][1] += wgt * v[p][1]; vs[p][2] += wgt * v[p][2];
			if (FULLM) {
				float B[9];
				if (st == 0) {
#pragma unroll
					for (int k = 0; k < 9; k++) B[k] = dv[p][k];
				} else {
					mm3(dv[p], A[p], B);	// dv_st @ dphi_st
				}
#pragma unroll
				for (int k = 0; k < 9; k++) {
					S[p][k] = (st == 0) ? B[k] : S[p][k] + wgt * B[k];
					A[p][k] = ((k % 4 == 0) ? 1.f : 0.f) + step * B[k];	// dphi_{st+1}
				}
			}
			px[p] = x0[p] + step * v[p][0]; py[p] = x1[p] + step * v[p][1]; pz[p] = x2[p] + step * v[p][2];
		}
	}
#pragma unroll
	for (int p = 0; p < P; p++) {
		px[p] = x0[p] + dt6 * vs[p][0]; py[p] = x1[p] + dt6 * vs[p][1]; pz[p] = x2[p] + dt6 * vs[p][2];
		if (MODE != 2 && ok[p]) {
			goal_pos[3 * j[p]] = px[p]; goal_pos[3 * j[p] + 1] = py[p]; goal_pos[3 * j[p] + 2] = pz[p];
		}
	}
	if (FULLM) {
		warp_eval3<P, true, false>(a, sh, srec, px, py, pz, ok, v, dv);
#pragma unroll
		for (int p = 0; p < P; p++) {
			if (!ok[p]) continue;
			float D[9];
#pragma unroll
			for (int k = 0; k < 9; k++) D[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt6 * S[p][k];	// dphi
			const size_t jj = j[p];
			if (MODE == 1) {
#pragma unroll
				for (int k = 0; k < 9; k++) { deformation[9 * jj + k] = D[k]; goal_grad[9 * jj + k] = dv[p][k]; }
				goal_val[3 * jj] = v[p][0]; goal_val[3 * jj + 1] = v[p][1]; goal_val[3 * jj + 2] = v[p][2];
			} else {
				const float *d = dv[p];
				const float w0 = d[7] - d[5], w1 = d[2] - d[6], w2 = d[3] - d[1];
				if (ref_hel) ref_hel[jj] = v[p][0] * w0 + v[p][1] * w1 + v[p][2] * w2;
				const float c00 = D[4] * D[8] - D[5] * D[7], c01 = D[2] * D[7] - D[1] * D[8], c02 = D[1] * D[5] - D[2] * D[4];
				const float c10 = D[5] * D[6] - D[3] * D[8], c11 = D[0] * D[8] - D[2] * D[6], c12 = D[2] * D[3] - D[0] * D[5];
				const float c20 = D[3] * D[7] - D[4] * D[6], c21 = D[1] * D[6] - D[0] * D[7], c22 = D[0] * D[4] - D[1] * D[3];
				const float det = D[0] * c00 + D[1] * c10 + D[2] * c20;
				const float inv = 1.f / det;
				ref_vor[3 * jj] = (c00 * w0 + c01 * w1 + c02 * w2) * inv;
				ref_vor[3 * jj + 1] = (c10 * w0 + c11 * w1 + c12 * w2) * inv;
				ref_vor[3 * jj + 2] = (c20 * w0 + c21 * w1 + c22 * w2) * inv;
			}
		}
	}
}

// RK4 with the deformation chain, 4 points per thread: the per-point integrator state (start point, velocity sum, S and
// dphi: 24 floats) lives in SHARED memory between the evaluations, [slot][thread] so that lanes touch consecutive banks;
// only the evaluation itself (the forward kernel's register footprint) is in registers.
constexpr int RKS_P = 4;
constexpr int RKS_STATE = 24;	// x0 3 | vs 3 | S 9 | A 9

template <int MODE>
__global__ void __launch_bounds__(TL_TILE / RKS_P, 3) rk4_tiled3s_kernel(TiledArgs a, float dt, float *__restrict__ goal_pos, float *__restrict__ deformation,
									  float *__restrict__ goal_val, float *__restrict__ goal_grad, float *__restrict__ ref_vor,
									  float *__restrict__ ref_hel)
{
	constexpr int P = RKS_P, T = TL_TILE / RKS_P;
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	float *state = reinterpret_cast<float *>(tl_smem + (size_t)a.cap * 48) + threadIdx.x;	// + (p * RKS_STATE + k) * T
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float px[P], py[P], pz[P], v[P][3], dv[P][9];
	bool ok[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		const size_t j = ok[p] ? (size_t)a.perm[t] : 0;
		px[p] = ok[p] ? a.x[3 * j] : 0.f;
		py[p] = ok[p] ? a.x[3 * j + 1] : 0.f;
		pz[p] = ok[p] ? a.x[3 * j + 2] : 0.f;
		float *sp = state + (p * RKS_STATE) * T;
		sp[0] = px[p]; sp[T] = py[p]; sp[2 * T] = pz[p];
		sp[3 * T] = 0.f; sp[4 * T] = 0.f; sp[5 * T] = 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	const float hdt = dt * .5f, dt6 = dt / 6.f;
#pragma unroll 1
	for (int st = 0; st < 4; st++) {
		warp_eval3<P, true, true>(a, sh, srec, px, py, pz, ok, v, dv);
		const float wgt = (st == 0 || st == 3) ? 1.f : 2.f;
		const float step = (st < 2) ? hdt : dt;
#pragma unroll
		for (int p = 0; p < P; p++) {
			float *sp = state + (p * RKS_STATE) * T;
#pragma unroll
			for (int k = 0; k < 3; k++) sp[(3 + k) * T] += wgt * v[p][k];
			float B[9];
			if (st == 0) {
#pragma unroll
				for (int k = 0; k < 9; k++) B[k] = dv[p][k];
			} else {
				float A[9];
#pragma unroll
				for (int k = 0; k < 9; k++) A[k] = sp[(15 + k) * T];
				mm3(dv[p], A, B);	// dv_st @ dphi_st
			}
#pragma unroll
			for (int k = 0; k < 9; k++) {
				sp[(6 + k) * T] = (st == 0) ? B[k] : sp[(6 + k) * T] + wgt * B[k];
				sp[(15 + k) * T] = ((k % 4 == 0) ? 1.f : 0.f) + step * B[k];	// dphi_{st+1}
			}
			px[p] = sp[0] + step * v[p][0]; py[p] = sp[T] + step * v[p][1]; pz[p] = sp[2 * T] + step * v[p][2];
		}
	}
#pragma unroll
	for (int p = 0; p < P; p++) {
		const float *sp = state + (p * RKS_STATE) * T;
		px[p] = sp[0] + dt6 * sp[3 * T]; py[p] = sp[T] + dt6 * sp[4 * T]; pz[p] = sp[2 * T] + dt6 * sp[5 * T];
	}
	warp_eval3<P, true, true>(a, sh, srec, px, py, pz, ok, v, dv);
#pragma unroll
	for (int p = 0; p < P; p++) {
		if (!ok[p]) continue;
		const float *sp = state + (p * RKS_STATE) * T;
		const size_t jj = (size_t)a.perm[slot_pos<P>(t0, p)];
		float D[9];
#pragma unroll
		for (int k = 0; k < 9; k++) D[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt6 * sp[(6 + k) * T];	// dphi
		if (MODE == 1) {
			goal_pos[3 * jj] = px[p]; goal_pos[3 * jj + 1] = py[p]; goal_pos[3 * jj + 2] = pz[p];
#pragma unroll
			for (int k = 0; k < 9; k++) { deformation[9 * jj + k] = D[k]; goal_grad[9 * jj + k] = dv[p][k]; }
			goal_val[3 * jj] = v[p][0]; goal_val[3 * jj + 1] = v[p][1]; goal_val[3 * jj + 2] = v[p][2];
		} else {
			const float *d = dv[p];
			const float w0 = d[7] - d[5], w1 = d[2] - d[6], w2 = d[3] - d[1];
			if (ref_hel) ref_hel[jj] = v[p][0] * w0 + v[p][1] * w1 + v[p][2] * w2;
			const float c00 = D[4] * D[8] - D[5] * D[7], c01 = D[2] * D[7] - D[1] * D[8], c02 = D[1] * D[5] - D[2] * D[4];
			const float c10 = D[5] * D[6] - D[3] * D[8], c11 = D[0] * D[8] - D[2] * D[6], c12 = D[2] * D[3] - D[0] * D[5];
			const float c20 = D[3] * D[7] - D[4] * D[6], c21 = D[1] * D[6] - D[0] * D[7], c22 = D[0] * D[4] - D[1] * D[3];
			const float det = D[0] * c00 + D[1] * c10 + D[2] * c20;
			const float inv = 1.f / det;
			ref_vor[3 * jj] = (c00 * w0 + c01 * w1 + c02 * w2) * inv;
			ref_vor[3 * jj + 1] = (c10 * w0 + c11 * w1 + c12 * w2) * inv;
			ref_vor[3 * jj + 2] = (c20 * w0 + c21 * w1 + c22 * w2) * inv;
		}
	}
}

// ---- tile -> row table ----------------------------------------------------------------------------
// Row r (a fixed (x, y) pair of the padded sample grid; r == nrows: the samples outside it) owns the tile ids
// [s_r / T + r, s_r / T + r + ceil(n_r / T)), s_r = first sorted sample of the row.  The ranges of different rows
// never overlap, so no prefix sum over rows is needed; unused ids stay -1.
__global__ void tile_rows_kernel(const int32_t *__restrict__ scs, int Q, int rowlen, int nrows, int pcell, int32_t *__restrict__ tile_row)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > nrows) return;
	const int s = scs[r < nrows ? r * rowlen : pcell];
	const int e = (r < nrows) ? scs[(r + 1) * rowlen] : Q;
	const int base = s / TL_TILE + r;
	const int n = (e - s + TL_TILE - 1) / TL_TILE;
	for (int c = 0; c < n; c++) tile_row[base + c] = r;
}

static int64_t tile_slots(const Grid &g, int64_t Q)
{
	const int rowlen = g.pdims[g.D - 1];
	return Q / TL_TILE + g.pcell / rowlen + 2;
}

// tunables (gsr_set_tuning)
int g_tiled_min_q = 1 << 17;
int g_tiled_cap = 512;
int g_fw_p4_min_spc = 0;	// samples per cell above which the forward kernel takes 4 points per thread (else 2)
int g_rk4_smem_state = 1;	// 1: RK4 with the deformation chain keeps its state in shared memory (4 points / thread)

static size_t tiled_smem(int cap) { return (size_t)cap * 48; }

template <typename K>
static int prep(K kernel, int cap, bool max_shared)
{
	cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem(cap));
	// forward: 4-5 CTAs per SM are register-feasible, so ask for the large shared-memory carve-out; the RK4 kernels are
	// register-limited to 2 CTAs and keep the default split (their few spilled values live in L1)
	if (e == cudaSuccess && max_shared) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace gsr

using namespace gsr;

extern "C" int64_t gsr_tile_slots(const gsr_grid_desc *d, int64_t Q)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0) return GSR_EINVAL;
	return tile_slots(g, Q);
}

extern "C" int gsr_build_tiles(const gsr_grid_desc *d, const int32_t *sample_cell_start, int64_t Q, int32_t *tile_row, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 30) || !sample_cell_start || !tile_row) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	const int rowlen = g.pdims[g.D - 1], nrows = g.pcell / rowlen;
	cudaError_t e = cudaMemsetAsync(tile_row, 0xff, sizeof(int32_t) * (size_t)tile_slots(g, Q), st);
	if (e != cudaSuccess) return (int)e;
	g_launches += 1;
	tile_rows_kernel<<<(nrows + 1 + 127) / 128, 128, 0, st>>>(sample_cell_start, (int)Q, rowlen, nrows, g.pcell, tile_row);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_set_tuning(int key, int value)
{
	switch (key) {
	case GSR_TUNE_TILED_MIN_Q: g_tiled_min_q = value; return GSR_OK;
	case GSR_TUNE_RK4_SMEM_STATE: g_rk4_smem_state = value; return GSR_OK;
	case GSR_TUNE_FW_P4_MIN_SPC: g_fw_p4_min_spc = value; return GSR_OK;
	case GSR_TUNE_TILED_CAP:
		if (value < 0 || tiled_smem(value) > 200 * 1024) return GSR_EINVAL;
		g_tiled_cap = value;
		return GSR_OK;
	default: return GSR_EINVAL;
	}
}

namespace gsr {

bool use_tiled(const Grid &g, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row)
{
	return g.D == 3 && perm && scs && tile_row && Q >= g_tiled_min_q;
}

static TiledArgs make_targs(const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
			    const int32_t *scs, const int32_t *tile_row)
{
	TiledArgs a;
	a.P = P; a.cell_start = cell_start; a.packed = (const float4 *)packed; a.cull = cull; a.x = x; a.Q = (int)Q; a.perm = perm; a.scs = scs; a.tile_row = tile_row;
	a.cap = g_tiled_cap;
	return a;
}

#define TL_LAUNCH(PP, KERNEL, ...)                                                              \
	do {                                                                                    \
		int rc__ = prep(KERNEL, a.cap, (PP) == FW_P);                                   \
		if (rc__) return rc__;                                                          \
		KERNEL<<<blocks, TL_TILE / (PP), tiled_smem(a.cap), st>>>(__VA_ARGS__);         \
	} while (0)

constexpr int FW_P = 4;	// points per thread, forward
constexpr int RK_P = 2;	// points per thread, RK4 with the deformation chain (register budget)

int launch_forward_tiled3(const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
			  const int32_t *scs, const int32_t *tile_row, float *val, float *grad, bool accumulate, cudaStream_t st)
{
	TiledArgs a = make_targs(P, cell_start, packed, cull, x, Q, perm, scs, tile_row);
	const int blocks = (int)tile_slots(P.g, Q);
	const bool dense = (double)Q >= (double)g_fw_p4_min_spc * (double)P.g.ncell;	// samples per cell: see launch_rk4_tiled3
#define FW_DISPATCH(PP)                                                                                                 \
	do {                                                                                                            \
		if (accumulate) {                                                                                       \
			if (val && grad) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, true, true>), a, val, grad);    \
			else if (val) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, false, true>), a, val, grad);      \
			else TL_LAUNCH(PP, (forward_tiled3_kernel<PP, false, true, true>), a, val, grad);               \
		} else {                                                                                                \
			if (val && grad) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, true, false>), a, val, grad);   \
			else if (val) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, false, false>), a, val, grad);     \
			else TL_LAUNCH(PP, (forward_tiled3_kernel<PP, false, true, false>), a, val, grad);              \
		}                                                                                                       \
	} while (0)
	if (dense) FW_DISPATCH(4);
	else FW_DISPATCH(2);
#undef FW_DISPATCH
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

int launch_rk4_tiled3(int mode, const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
		      const int32_t *scs, const int32_t *tile_row, float dt, float *goal_pos, float *deformation, float *goal_val, float *goal_grad,
		      float *ref_vor, float *ref_hel, cudaStream_t st)
{
	TiledArgs a = make_targs(P, cell_start, packed, cull, x, Q, perm, scs, tile_row);
	const int blocks = (int)tile_slots(P.g, Q);
	// many samples per cell: warps of 128 points are still compact, and 4 points per thread halve the per-candidate overhead;
	// fewer: 64-point warps cull better
	if (mode != 0 && g_rk4_smem_state && (double)Q >= 1024. * (double)P.g.ncell) {
		const size_t sm = tiled_smem(a.cap) + sizeof(float) * RKS_STATE * RKS_P * (TL_TILE / RKS_P);
		cudaError_t e = (mode == 1) ? cudaFuncSetAttribute(rk4_tiled3s_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
					    : cudaFuncSetAttribute(rk4_tiled3s_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e == cudaSuccess)
			e = (mode == 1) ? cudaFuncSetAttribute(rk4_tiled3s_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
					: cudaFuncSetAttribute(rk4_tiled3s_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		if (e != cudaSuccess) return (int)e;
		if (mode == 1) rk4_tiled3s_kernel<1><<<blocks, TL_TILE / RKS_P, sm, st>>>(a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
		else rk4_tiled3s_kernel<2><<<blocks, TL_TILE / RKS_P, sm, st>>>(a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
		GSR_CHECK_LAUNCH();
		return GSR_OK;
	}
	if (mode == 0) TL_LAUNCH(FW_P, (rk4_tiled3_kernel<0, FW_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	else if (mode == 1) TL_LAUNCH(RK_P, (rk4_tiled3_kernel<1, RK_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	else TL_LAUNCH(RK_P, (rk4_tiled3_kernel<2, RK_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

}  // namespace gsr
