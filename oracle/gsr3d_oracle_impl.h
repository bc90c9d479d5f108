/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's 3D Gaussian-Spatial-Representation kernels
 * (reference: 3D/GSR.py).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path
 * (gaussian-fluids-code_b200/csrc) never links or calls it.
 *
 * Parity status: pinned.  The restatement is checked against golden vectors that
 * were produced by executing the reference's own kernel bodies (3D/GSR.py, run as
 * plain Python through tests/golden/ti_shim.py) and its dense torch class; see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py.
 *
 * This header is included twice by gsr_oracle.c: REAL=float (suffix _f32, the
 * reference's arithmetic) and REAL=double (suffix _f64, "truth" for tolerance
 * accounting).  Inputs are always float arrays; binning is always done in f32
 * exactly as the reference does, so both variants see the same neighbour lists.
 *
 * The loops, the per-pair recomputation of R, S^2 and Sigma^-1, and the operation
 * order follow the reference so that timing this code is a fair stand-in for the
 * reference's Taichi-CPU path (which cannot run here: taichi is not installed).
 */

#ifndef REAL
#error "define REAL and SFX before including"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct { REAL m[3][3]; } FN(M3);
typedef struct { REAL v[3]; } FN(V3);
#define M3T FN(M3)
#define V3T FN(V3)

static inline REAL FN(r_exp)(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL FN(r_sqrt)(REAL x) { return sizeof(REAL) == 4 ? (REAL)sqrtf((float)x) : (REAL)sqrt((double)x); }
static inline REAL FN(r_sign)(REAL x) { return (REAL)((x > 0) - (x < 0)); }

static inline M3T FN(m3_zero)(void) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = 0; return r; }
static inline M3T FN(m3_eye)(void) { M3T r = FN(m3_zero)(); r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
static inline M3T FN(m3_mul)(M3T a, M3T b) {
	M3T r;
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
		REAL s = 0;
		for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
		r.m[i][j] = s;
	}
	return r;
}
static inline M3T FN(m3_t)(M3T a) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
static inline M3T FN(m3_add)(M3T a, M3T b) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
static inline M3T FN(m3_sub)(M3T a, M3T b) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
static inline M3T FN(m3_scale)(REAL s, M3T a) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j]; return r; }
static inline M3T FN(m3_divs)(M3T a, REAL s) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] / s; return r; }
static inline V3T FN(m3_mulv)(M3T a, V3T x) { V3T r; for (int i = 0; i < 3; i++) { REAL s = 0; for (int k = 0; k < 3; k++) s += a.m[i][k] * x.v[k]; r.v[i] = s; } return r; }
/* row-vector times matrix (Taichi `vec @ mat`) */
static inline V3T FN(v_mulm)(V3T x, M3T a) { V3T r; for (int j = 0; j < 3; j++) { REAL s = 0; for (int k = 0; k < 3; k++) s += x.v[k] * a.m[k][j]; r.v[j] = s; } return r; }
static inline REAL FN(v_dot)(V3T a, V3T b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
static inline V3T FN(v_scale)(REAL s, V3T a) { V3T r = {{s * a.v[0], s * a.v[1], s * a.v[2]}}; return r; }
static inline V3T FN(v_add)(V3T a, V3T b) { V3T r = {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; return r; }
static inline V3T FN(v_sub)(V3T a, V3T b) { V3T r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static inline M3T FN(v_outer)(V3T a, V3T b) { M3T r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.v[i] * b.v[j]; return r; }
static inline V3T FN(m3_col)(M3T a, int c) { V3T r = {{a.m[0][c], a.m[1][c], a.m[2][c]}}; return r; }
static inline V3T FN(m3_row)(M3T a, int c) { V3T r = {{a.m[c][0], a.m[c][1], a.m[c][2]}}; return r; }

/* R(q), S2, cov_inv exactly as 3D/GSR.py:278-289 (and :311-323, :612-623). */
static inline void FN(gauss_geom)(const float *rot, const float *scal, REAL q[4], M3T *R, M3T *S2, M3T *cov_inv)
{
	REAL r0 = rot[0], r1 = rot[1], r2 = rot[2], r3 = rot[3];
	REAL len = FN(r_sqrt)(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
	q[0] = r0 / len; q[1] = r1 / len; q[2] = r2 / len; q[3] = r3 / len;
	R->m[0][0] = (REAL)1 - (REAL)2 * (q[2] * q[2] + q[3] * q[3]);
	R->m[0][1] = (REAL)2 * (q[1] * q[2] - q[0] * q[3]);
	R->m[0][2] = (REAL)2 * (q[1] * q[3] + q[0] * q[2]);
	R->m[1][0] = (REAL)2 * (q[1] * q[2] + q[0] * q[3]);
	R->m[1][1] = (REAL)1 - (REAL)2 * (q[1] * q[1] + q[3] * q[3]);
	R->m[1][2] = (REAL)2 * (q[2] * q[3] - q[0] * q[1]);
	R->m[2][0] = (REAL)2 * (q[1] * q[3] - q[0] * q[2]);
	R->m[2][1] = (REAL)2 * (q[2] * q[3] + q[0] * q[1]);
	R->m[2][2] = (REAL)1 - (REAL)2 * (q[1] * q[1] + q[2] * q[2]);
	*S2 = FN(m3_zero)();
	S2->m[0][0] = FN(r_exp)((REAL)2 * (REAL)scal[0]);
	S2->m[1][1] = FN(r_exp)((REAL)2 * (REAL)scal[1]);
	S2->m[2][2] = FN(r_exp)((REAL)2 * (REAL)scal[2]);
	*cov_inv = FN(m3_mul)(FN(m3_mul)(*R, *S2), FN(m3_t)(*R));
}

/* Forward u, grad u at one point: 3D/GSR.py:599-632 (== loop 1, :270-298). */
static void FN(o3_point)(const gsr_grid3 *g, const float *pos, const float *scal, const float *rot, const float *vals,
			 REAL tau, int dim, const REAL x[3], REAL *val, REAL *grad /* dim*3 or NULL */)
{
	int c[3];
	gsr_cell_of_f32(g, (float)x[0], (float)x[1], (float)x[2], c);
	for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
	for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++)
	for (int gk = imax(c[2] - 1, 0); gk <= imin(c[2] + 1, g->dims[2] - 1); gk++) {
		long cell = ((long)gi * g->dims[1] + gj) * g->dims[2] + gk;
		for (int i_id = g->offset[cell]; i_id < g->offset[cell] + g->cnt[cell]; i_id++) {
			int i = g->sorted_id[i_id];
			V3T d = {{x[0] - (REAL)pos[3 * i], x[1] - (REAL)pos[3 * i + 1], x[2] - (REAL)pos[3 * i + 2]}};
			REAL q[4]; M3T R, S2, C;
			FN(gauss_geom)(rot + 4 * i, scal + 3 * i, q, &R, &S2, &C);
			REAL gaussian = FN(r_exp)((REAL)-.5 * FN(v_dot)(FN(v_mulm)(d, C), d));
			if (gaussian >= tau) {
				V3T gg = FN(v_scale)(-gaussian, FN(m3_mulv)(C, d));
				for (int dd = 0; dd < dim; dd++) {
					REAL v = vals[dim * i + dd];
					val[dd] += v * (gaussian - tau);
					if (grad) {
						grad[3 * dd + 0] += v * gg.v[0];
						grad[3 * dd + 1] += v * gg.v[1];
						grad[3 * dd + 2] += v * gg.v[2];
					}
				}
			}
		}
	}
}

/* Loop 1 of get_losses_ti (3D/GSR.py:270-298).  val/grad are accumulated into (caller zero-fills). */
void FN(o3_forward)(const gsr_grid3 *g, const float *pos, const float *scal, const float *rot, const float *vals,
		    double tau_d, int dim, const float *x, long Q, REAL *val, REAL *grad, int nthreads)
{
	REAL tau = (REAL)(float)tau_d; /* clamp_threshold is an f32 constant in the kernel */
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		REAL xx[3] = {x[3 * j], x[3 * j + 1], x[3 * j + 2]};
		FN(o3_point)(g, pos, scal, rot, vals, tau, dim, xx, val + dim * j, grad ? grad + 3 * dim * j : NULL);
	}
}

#define ACC(ptr, inc) do { REAL inc__ = (inc); _Pragma("omp atomic") ptr += inc__; } while (0)

/*
 * Loop 2 of get_losses_ti (3D/GSR.py:299-540): analytic backward of the six losses.
 * val/grad are the totals of loop 1 (inputs).  Accumulates into the three gradient sets
 * (direct = the parameters' own .grad; vor; div) which may alias each other.
 * weights[6] = {val, boundary, grad, vor, hel, div}.
 */
void FN(o3_backward)(const gsr_grid3 *g, const float *pos, const float *scal, const float *rot, const float *vals,
		     double tau_d, int dim, const float *x, long Q,
		     const REAL *val, const REAL *grad,
		     const float *ref_val, const float *normals, const float *ref_grad, const float *ref_vor, const float *ref_hel,
		     const double *weights, const int *stop_gradient,
		     REAL *g_pos, REAL *g_scal, REAL *g_rot, REAL *g_val,
		     REAL *vor_pos, REAL *vor_scal, REAL *vor_rot, REAL *vor_val,
		     REAL *div_pos, REAL *div_scal, REAL *div_rot, REAL *div_val, int nthreads)
{
	const REAL tau = (REAL)(float)tau_d;
	const REAL weight_val = (REAL)(float)weights[0], weight_boundary = (REAL)(float)weights[1], weight_grad = (REAL)(float)weights[2];
	const REAL weight_vor = (REAL)(float)weights[3], weight_hel = (REAL)(float)weights[4], weight_div = (REAL)(float)weights[5];
	/* :299-300 — note weight_hel is not part of this test */
	if (weight_val == 0 && weight_boundary == 0 && weight_grad == 0 && weight_vor == 0 && weight_div == 0) return;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		int c[3];
		gsr_cell_of_f32(g, x[3 * j], x[3 * j + 1], x[3 * j + 2], c);
		const REAL *vj = val + dim * j;
		const REAL *Gj = grad + 3 * dim * j;
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++)
		for (int gk = imax(c[2] - 1, 0); gk <= imin(c[2] + 1, g->dims[2] - 1); gk++) {
			long cell = ((long)gi * g->dims[1] + gj) * g->dims[2] + gk;
			for (int i_id = g->offset[cell]; i_id < g->offset[cell] + g->cnt[cell]; i_id++) {
				int i = g->sorted_id[i_id];
				if (stop_gradient && stop_gradient[i]) continue;
				V3T d = {{(REAL)x[3 * j] - (REAL)pos[3 * i], (REAL)x[3 * j + 1] - (REAL)pos[3 * i + 1], (REAL)x[3 * j + 2] - (REAL)pos[3 * i + 2]}};
				REAL r[4] = {rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3]};
				REAL q[4]; M3T R, S2, C;
				FN(gauss_geom)(rot + 4 * i, scal + 3 * i, q, &R, &S2, &C);
				REAL gaussian = FN(r_exp)((REAL)-.5 * FN(v_dot)(FN(v_mulm)(d, C), d));
				if (!(gaussian >= tau)) continue;

				/* :327-352 */
				V3T Cd = FN(m3_mulv)(C, d);
				V3T grad_gaussian = FN(v_scale)(-gaussian, Cd);
				M3T dRq[4];
				{
					REAL a0[3][3] = {{0, -2 * q[3], 2 * q[2]}, {2 * q[3], 0, -2 * q[1]}, {-2 * q[2], 2 * q[1], 0}};
					REAL a1[3][3] = {{0, 2 * q[2], 2 * q[3]}, {2 * q[2], -4 * q[1], -2 * q[0]}, {2 * q[3], 2 * q[0], -4 * q[1]}};
					REAL a2[3][3] = {{-4 * q[2], 2 * q[1], 2 * q[0]}, {2 * q[1], 0, 2 * q[3]}, {-2 * q[0], 2 * q[3], -4 * q[2]}};
					REAL a3[3][3] = {{-4 * q[3], -2 * q[0], 2 * q[1]}, {2 * q[0], -4 * q[3], 2 * q[2]}, {2 * q[1], 2 * q[2], 0}};
					memcpy(dRq[0].m, a0, sizeof a0); memcpy(dRq[1].m, a1, sizeof a1);
					memcpy(dRq[2].m, a2, sizeof a2); memcpy(dRq[3].m, a3, sizeof a3);
				}
				REAL r_length = FN(r_sqrt)(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
				REAL r_len3 = r_length * r_length * r_length;
				M3T rsum = FN(m3_add)(FN(m3_add)(FN(m3_add)(FN(m3_scale)(r[0], dRq[0]), FN(m3_scale)(r[1], dRq[1])),
								 FN(m3_scale)(r[2], dRq[2])), FN(m3_scale)(r[3], dRq[3]));
				M3T dRr[4];
				for (int m = 0; m < 4; m++)
					dRr[m] = FN(m3_add)(FN(m3_scale)(-r[m] / r_len3, rsum), FN(m3_divs)(dRq[m], r_length));
				/* :353-364 */
				M3T dCr[4];
				M3T RS2 = FN(m3_mul)(R, S2);
				for (int m = 0; m < 4; m++) {
					dCr[m] = FN(m3_zero)();
					for (int kk = 0; kk < 3; kk++) for (int ll = 0; ll < 3; ll++) for (int ii = 0; ii < 3; ii++)
						dCr[m].m[kk][ll] += RS2.m[kk][ii] * dRr[m].m[ll][ii] + RS2.m[ll][ii] * dRr[m].m[kk][ii];
				}
				/* :365-368 */
				M3T dCs[3];
				for (int k = 0; k < 3; k++) {
					V3T col = FN(m3_col)(R, k);
					dCs[k] = FN(m3_scale)((REAL)2 * FN(r_exp)((REAL)2 * (REAL)scal[3 * i + k]), FN(v_outer)(col, col));
				}
				/* :369-375 */
				REAL dg_r[4], dg_s[3];
				for (int m = 0; m < 4; m++) dg_r[m] = (REAL)-.5 * gaussian * FN(v_dot)(FN(v_mulm)(d, dCr[m]), d);
				for (int k = 0; k < 3; k++) dg_s[k] = (REAL)-.5 * gaussian * FN(v_dot)(FN(v_mulm)(d, dCs[k]), d);
				/* :376-393 */
				M3T dgg_pos = FN(m3_zero)();
				V3T dgg_s[3], dgg_r[4];
				for (int k = 0; k < 3; k++) dgg_s[k] = (V3T){{0, 0, 0}};
				for (int m = 0; m < 4; m++) dgg_r[m] = (V3T){{0, 0, 0}};
				if (weight_grad != 0 || weight_vor != 0 || weight_div != 0) {
					dgg_pos = FN(m3_scale)(gaussian, FN(m3_mul)(C, FN(m3_sub)(FN(m3_eye)(), FN(m3_mul)(FN(v_outer)(d, d), C))));
					for (int k = 0; k < 3; k++)
						dgg_s[k] = FN(v_sub)(FN(v_scale)(-dg_s[k], Cd), FN(v_scale)(gaussian, FN(m3_mulv)(dCs[k], d)));
					for (int m = 0; m < 4; m++)
						dgg_r[m] = FN(v_sub)(FN(v_scale)(-dg_r[m], Cd), FN(v_scale)(gaussian, FN(m3_mulv)(dCr[m], d)));
				}
				V3T value = {{0, 0, 0}};
				for (int dd = 0; dd < dim; dd++) value.v[dd] = vals[dim * i + dd];

				/* value loss :396-411 */
				if (weight_val != 0) {
					REAL w = weight_val / (REAL)(dim * Q);
					REAL value_dot_sign = 0;
					for (int dd = 0; dd < dim; dd++) {
						REAL sg = FN(r_sign)(vj[dd] - (REAL)ref_val[dim * j + dd]);
						value_dot_sign += value.v[dd] * sg;
						ACC(g_val[dim * i + dd], w * (gaussian - tau) * sg);
					}
					for (int k = 0; k < 3; k++) ACC(g_pos[3 * i + k], w * gaussian * value_dot_sign * Cd.v[k]);
					for (int k = 0; k < 3; k++) ACC(g_scal[3 * i + k], w * value_dot_sign * dg_s[k]);
					for (int m = 0; m < 4; m++) ACC(g_rot[4 * i + m], w * value_dot_sign * dg_r[m]);
				}
				/* boundary :414-433 */
				if (weight_boundary != 0) {
					REAL w = weight_boundary / (REAL)Q;
					REAL svn = 0, value_dot_normal = 0;
					for (int dd = 0; dd < dim; dd++) {
						svn += vj[dd] * (REAL)normals[dim * j + dd];
						value_dot_normal += value.v[dd] * (REAL)normals[dim * j + dd];
					}
					svn = FN(r_sign)(svn);
					for (int dd = 0; dd < dim; dd++) ACC(g_val[dim * i + dd], w * svn * (gaussian - tau) * (REAL)normals[dim * j + dd]);
					for (int k = 0; k < 3; k++) ACC(g_pos[3 * i + k], w * svn * value_dot_normal * gaussian * Cd.v[k]);
					for (int k = 0; k < 3; k++) ACC(g_scal[3 * i + k], w * svn * value_dot_normal * dg_s[k]);
					for (int m = 0; m < 4; m++) ACC(g_rot[4 * i + m], w * svn * value_dot_normal * dg_r[m]);
				}
				/* gradient loss :436-451 */
				if (weight_grad != 0) {
					REAL w = weight_grad / (REAL)(3 * dim * Q);
					V3T vts = {{0, 0, 0}};
					for (int dd = 0; dd < dim; dd++) {
						V3T sg = {{FN(r_sign)(Gj[3 * dd] - (REAL)ref_grad[3 * dim * j + 3 * dd]),
							   FN(r_sign)(Gj[3 * dd + 1] - (REAL)ref_grad[3 * dim * j + 3 * dd + 1]),
							   FN(r_sign)(Gj[3 * dd + 2] - (REAL)ref_grad[3 * dim * j + 3 * dd + 2])}};
						vts = FN(v_add)(vts, FN(v_scale)(value.v[dd], sg));
						ACC(g_val[dim * i + dd], w * FN(v_dot)(sg, grad_gaussian));
					}
					for (int k = 0; k < 3; k++) ACC(g_pos[3 * i + k], w * FN(v_dot)(vts, FN(m3_col)(dgg_pos, k)));
					for (int k = 0; k < 3; k++) ACC(g_scal[3 * i + k], w * FN(v_dot)(vts, dgg_s[k]));
					for (int m = 0; m < 4; m++) ACC(g_rot[4 * i + m], w * FN(v_dot)(vts, dgg_r[m]));
				}
				/* vorticity + helicity :454-520 */
				if (weight_vor + weight_hel != 0 && dim == 3) {
					REAL w = weight_vor / (REAL)(3 * Q);
					V3T vor = {{Gj[3 * 2 + 1] - Gj[3 * 1 + 2], Gj[3 * 0 + 2] - Gj[3 * 2 + 0], Gj[3 * 1 + 0] - Gj[3 * 0 + 1]}};
					M3T E0 = FN(m3_zero)(), E1 = FN(m3_zero)(), E2 = FN(m3_zero)();
					E0.m[1][2] = -1; E0.m[2][1] = 1;
					E1.m[0][2] = 1; E1.m[2][0] = -1;
					E2.m[0][1] = -1; E2.m[1][0] = 1;
					V3T svd = {{FN(r_sign)(vor.v[0] - (REAL)ref_vor[3 * j]), FN(r_sign)(vor.v[1] - (REAL)ref_vor[3 * j + 1]), FN(r_sign)(vor.v[2] - (REAL)ref_vor[3 * j + 2])}};
					M3T M_vor = FN(m3_add)(FN(m3_add)(FN(m3_scale)(svd.v[0], E0), FN(m3_scale)(svd.v[1], E1)), FN(m3_scale)(svd.v[2], E2));
					V3T vM = FN(v_mulm)(value, M_vor);
					for (int dd = 0; dd < 3; dd++) ACC(vor_val[3 * i + dd], w * FN(v_dot)(FN(m3_row)(M_vor, dd), grad_gaussian));
					for (int k = 0; k < 3; k++) ACC(vor_pos[3 * i + k], w * FN(v_dot)(vM, FN(m3_col)(dgg_pos, k)));
					for (int k = 0; k < 3; k++) ACC(vor_scal[3 * i + k], w * FN(v_dot)(vM, dgg_s[k]));
					for (int m = 0; m < 4; m++) ACC(vor_rot[4 * i + m], w * FN(v_dot)(vM, dgg_r[m]));
					/* helicity :489-520 */
					w = weight_hel / (REAL)Q;
					V3T vv = {{vj[0], vj[1], vj[2]}};
					REAL shd = FN(r_sign)(FN(v_dot)(vv, vor) - (REAL)ref_hel[j]);
					M3T M_hel = FN(m3_add)(FN(m3_add)(FN(m3_scale)(vj[0], E0), FN(m3_scale)(vj[1], E1)), FN(m3_scale)(vj[2], E2));
					V3T vMh = FN(v_mulm)(value, M_hel);
					REAL d_hel_value[3];
					for (int dd = 0; dd < 3; dd++)
						d_hel_value[dd] = (gaussian - tau) * vor.v[dd] + FN(v_dot)(FN(m3_row)(M_hel, dd), grad_gaussian);
					REAL value_dot_vor = FN(v_dot)(value, vor);
					for (int dd = 0; dd < 3; dd++) ACC(vor_val[3 * i + dd], w * shd * d_hel_value[dd]);
					/* NB :498-500 uses ROWS of d_grad_gaussian_position here (columns elsewhere); the matrix is symmetric */
					for (int k = 0; k < 3; k++)
						ACC(vor_pos[3 * i + k], w * shd * (gaussian * Cd.v[k] * value_dot_vor + FN(v_dot)(vMh, FN(m3_row)(dgg_pos, k))));
					for (int k = 0; k < 3; k++) ACC(vor_scal[3 * i + k], w * shd * (dg_s[k] * value_dot_vor + FN(v_dot)(vMh, dgg_s[k])));
					for (int m = 0; m < 4; m++) ACC(vor_rot[4 * i + m], w * shd * (dg_r[m] * value_dot_vor + FN(v_dot)(vMh, dgg_r[m])));
				}
				/* divergence :523-540 */
				if (weight_div != 0 && dim == 3) {
					REAL w = weight_div / (REAL)Q;
					REAL M_div = (REAL)2 * (Gj[0] + Gj[4] + Gj[8]);
					for (int dd = 0; dd < 3; dd++) ACC(div_val[3 * i + dd], w * M_div * grad_gaussian.v[dd]);
					for (int k = 0; k < 3; k++) ACC(div_pos[3 * i + k], w * M_div * FN(v_dot)(value, FN(m3_col)(dgg_pos, k)));
					for (int k = 0; k < 3; k++) ACC(div_scal[3 * i + k], w * M_div * FN(v_dot)(value, dgg_s[k]));
					for (int m = 0; m < 4; m++) ACC(div_rot[4 * i + m], w * M_div * FN(v_dot)(value, dgg_r[m]));
				}
			}
		}
	}
}

/* advection_rk4_ti, 3D/GSR.py:634-665.  Any of deformation/goal_val/goal_grad may be NULL. */
void FN(o3_rk4)(const gsr_grid3 *g, const float *pos, const float *scal, const float *rot, const float *vals,
		double tau_d, const float *start, long Q, double dt_d,
		REAL *goal_pos, REAL *deformation, REAL *goal_val, REAL *goal_grad, int nthreads)
{
	const REAL tau = (REAL)(float)tau_d, dt = (REAL)(float)dt_d;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		REAL x[3] = {start[3 * j], start[3 * j + 1], start[3 * j + 2]};
		REAL v[4][3] = {{0}}, p[3];
		M3T dv[4];
		for (int s = 0; s < 4; s++) dv[s] = FN(m3_zero)();
		FN(o3_point)(g, pos, scal, rot, vals, tau, 3, x, v[0], &dv[0].m[0][0]);
		for (int k = 0; k < 3; k++) p[k] = x[k] + dt * (REAL).5 * v[0][k];
		FN(o3_point)(g, pos, scal, rot, vals, tau, 3, p, v[1], &dv[1].m[0][0]);
		for (int k = 0; k < 3; k++) p[k] = x[k] + dt * (REAL).5 * v[1][k];
		FN(o3_point)(g, pos, scal, rot, vals, tau, 3, p, v[2], &dv[2].m[0][0]);
		for (int k = 0; k < 3; k++) p[k] = x[k] + dt * v[2][k];
		FN(o3_point)(g, pos, scal, rot, vals, tau, 3, p, v[3], &dv[3].m[0][0]);
		REAL phi[3];
		for (int k = 0; k < 3; k++) {
			phi[k] = x[k] + dt / (REAL)6 * (v[0][k] + (REAL)2 * v[1][k] + (REAL)2 * v[2][k] + v[3][k]);
			goal_pos[3 * j + k] = phi[k];
		}
		if (deformation) {
			M3T I = FN(m3_eye)();
			M3T dphi1 = FN(m3_add)(I, FN(m3_scale)(dt * (REAL).5, dv[0]));
			M3T a1 = FN(m3_mul)(dv[1], dphi1);
			M3T dphi2 = FN(m3_add)(I, FN(m3_scale)(dt * (REAL).5, a1));
			M3T a2 = FN(m3_mul)(dv[2], dphi2);
			M3T dphi3 = FN(m3_add)(I, FN(m3_scale)(dt, a2));
			M3T sum = FN(m3_add)(FN(m3_add)(FN(m3_add)(dv[0], FN(m3_scale)((REAL)2, a1)), FN(m3_scale)((REAL)2, a2)), FN(m3_mul)(dv[3], dphi3));
			M3T dphi = FN(m3_add)(I, FN(m3_scale)(dt / (REAL)6, sum));
			memcpy(deformation + 9 * j, dphi.m, sizeof dphi.m);
		}
		if (goal_val && goal_grad) {
			REAL vp[3] = {0, 0, 0};
			M3T dvp = FN(m3_zero)();
			FN(o3_point)(g, pos, scal, rot, vals, tau, 3, phi, vp, &dvp.m[0][0]);
			for (int k = 0; k < 3; k++) goal_val[3 * j + k] = vp[k];
			memcpy(goal_grad + 9 * j, dvp.m, sizeof dvp.m);
		}
	}
}

/* ---- analytic initial field: regularised Biot-Savart sum over vortex particles (3D/init_cond.py:122-145) -------------
 * res[i] += U f(r) (w_j x d),  f = (1 - exp(-(r/a)^3)) / r^3,  d = x_i - x0_j, r = |d|          (vortex_particle, :122-131)
 * jac[i] += U (f'/r) [w_j]x d d^T + U f [w_j]x,  f' = -3/r^4 (1 - e) + 3/(a^3 r) e,  e = exp(-(r/a)^3)  (vortex_particle_gradient, :132-145)
 * U and a are f32 kernel arguments in the reference; accumulation into res/jac (+=), as there. */
void FN(o3_vortex_particles)(const float *x, long Q, const REAL *x0, const REAL *w, long M, float U_, float a_, REAL *res, REAL *jac, int nthreads)
{
	const REAL U = (REAL)U_, a = (REAL)a_;
#pragma omp parallel for num_threads(nthreads) schedule(static)
	for (long i = 0; i < Q; i++) {
		for (long j = 0; j < M; j++) {
			const REAL d[3] = {(REAL)x[3 * i] - x0[3 * j], (REAL)x[3 * i + 1] - x0[3 * j + 1], (REAL)x[3 * i + 2] - x0[3 * j + 2]};
			const REAL r = FN(r_sqrt)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
			const REAL q = r / a, e = FN(r_exp)(-(q * q * q));
			const REAL fr = (REAL)1 / (r * r * r) * ((REAL)1 - e);
			const REAL wj[3] = {w[3 * j], w[3 * j + 1], w[3 * j + 2]};
			const REAL c[3] = {wj[1] * d[2] - wj[2] * d[1], wj[2] * d[0] - wj[0] * d[2], wj[0] * d[1] - wj[1] * d[0]};	/* w x d */
			if (res)
				for (int k = 0; k < 3; k++) res[3 * i + k] += U * fr * c[k];
			if (jac) {
				const REAL frp = (REAL)-3 / (r * r * r * r) * ((REAL)1 - e) + (REAL)3 / (a * a * a * r) * e;
				const REAL W[3][3] = {{0, -wj[2], wj[1]}, {wj[2], 0, -wj[0]}, {-wj[1], wj[0], 0}};
				for (int k = 0; k < 3; k++)
					for (int l = 0; l < 3; l++) jac[9 * i + 3 * k + l] += U * (frp / r) * (c[k] * d[l]) + U * fr * W[k][l];
			}
		}
	}
}

/* ---- N1: trilinear resampling of a lattice field at given positions (ti_get_interp_val, 3D/advance_density.py:24-50) ----
 * field (nx,ny,nz), positions (Q,3) inside the domain, result (Q).  The domain bounds are f32 kernel arguments in the reference. */
void FN(o3_interp_val)(const REAL *field, const int *dims, const REAL *positions, long Q, const float *domain, REAL *result, int nthreads)
{
	const int nx = dims[0], ny = dims[1], nz = dims[2];
	const REAL lo[3] = {(REAL)domain[0], (REAL)domain[2], (REAL)domain[4]}, hi[3] = {(REAL)domain[1], (REAL)domain[3], (REAL)domain[5]};
	const REAL d[3] = {(hi[0] - lo[0]) / (REAL)(nx - 1), (hi[1] - lo[1]) / (REAL)(ny - 1), (hi[2] - lo[2]) / (REAL)(nz - 1)};
	const int n[3] = {nx, ny, nz};
#define FLD(i, j, k) field[((size_t)(i) * ny + (j)) * nz + (k)]
#pragma omp parallel for num_threads(nthreads) schedule(static)
	for (long q = 0; q < Q; q++) {
		int i0[3], i1[3];
		REAL w[3];
		for (int a = 0; a < 3; a++) {
			const REAL p = positions[3 * q + a] - lo[a];
			i0[a] = (int)(sizeof(REAL) == 4 ? floorf((float)(p / d[a])) : floor((double)(p / d[a])));
			i1[a] = i0[a] + 1 < n[a] - 1 ? i0[a] + 1 : n[a] - 1;
			/* corner_min = ti_get_coord(pi, ...) - zero_p = (lo + (hi - lo) / (n - 1) * pi) - lo */
			const REAL corner = (lo[a] + (hi[a] - lo[a]) / (REAL)(n[a] - 1) * (REAL)i0[a]) - lo[a];
			w[a] = (p - corner) / d[a];
		}
		const REAL one = (REAL)1;
		result[q] = FLD(i0[0], i0[1], i0[2]) * (one - w[0]) * (one - w[1]) * (one - w[2]) + FLD(i1[0], i0[1], i0[2]) * w[0] * (one - w[1]) * (one - w[2])
			  + FLD(i0[0], i1[1], i0[2]) * (one - w[0]) * w[1] * (one - w[2]) + FLD(i1[0], i1[1], i0[2]) * w[0] * w[1] * (one - w[2])
			  + FLD(i0[0], i0[1], i1[2]) * (one - w[0]) * (one - w[1]) * w[2] + FLD(i1[0], i0[1], i1[2]) * w[0] * (one - w[1]) * w[2]
			  + FLD(i0[0], i1[1], i1[2]) * (one - w[0]) * w[1] * w[2] + FLD(i1[0], i1[1], i1[2]) * w[0] * w[1] * w[2];
	}
#undef FLD
}

#undef ACC
#undef M3T
#undef V3T
#undef FN
#undef CAT
#undef CAT_
