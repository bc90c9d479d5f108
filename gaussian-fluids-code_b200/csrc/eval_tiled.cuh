// eval_tiled.cuh — device code of the large-Q evaluation kernels (see eval_tiled.cu for the design): tile staging with TMA
// bulk copies and the warp-collective evaluation of 32 P points.  Shared by eval_tiled.cu and density.cu.
#pragma once
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
	unsigned long long *exec_count;	// census instantiations only (COUNT): + 32 P per candidate that survives the warp's culling
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
template <int P, bool NEED_GRAD, bool PREFETCH, bool COUNT = false>
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
						if (COUNT && lane == 0 && mask) atomicAdd(a.exec_count, (unsigned long long)__popc(mask) * 32ull * P);
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

}  // namespace gsr
