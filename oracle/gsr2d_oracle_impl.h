/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See gsr3d_oracle_impl.h for the rules.
 *
 * CPU restatement of the reference's 2D kernels (reference: 2D/GSR.py).  Included
 * twice (REAL=float → _f32, REAL=double → _f64).  Parity status: pinned by golden
 * vectors produced from the reference's own kernel bodies (tests/golden/make_golden.py).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

static inline REAL FN(q_exp)(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL FN(q_sin)(REAL x) { return sizeof(REAL) == 4 ? (REAL)sinf((float)x) : (REAL)sin((double)x); }
static inline REAL FN(q_cos)(REAL x) { return sizeof(REAL) == 4 ? (REAL)cosf((float)x) : (REAL)cos((double)x); }
static inline REAL FN(q_sign)(REAL x) { return (REAL)((x > 0) - (x < 0)); }

/* cov_inv = R S2 R^T with R = [[c,-s],[s,c]] — 2D/GSR.py:275-277 */
static inline void FN(geom2)(float theta, const float *scal, REAL C[2][2])
{
	REAL c = FN(q_cos)((REAL)theta), s = FN(q_sin)((REAL)theta);
	REAL R[2][2] = {{c, -s}, {s, c}};
	REAL S0 = FN(q_exp)((REAL)2 * (REAL)scal[0]), S1 = FN(q_exp)((REAL)2 * (REAL)scal[1]);
	REAL RS[2][2] = {{R[0][0] * S0 + R[0][1] * (REAL)0, R[0][0] * (REAL)0 + R[0][1] * S1},
			 {R[1][0] * S0 + R[1][1] * (REAL)0, R[1][0] * (REAL)0 + R[1][1] * S1}};
	for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++)
		C[i][j] = RS[i][0] * R[j][0] + RS[i][1] * R[j][1];
}

static inline REAL FN(gauss2)(const REAL d[2], REAL C[2][2])
{
	REAL t0 = d[0] * C[0][0] + d[1] * C[1][0], t1 = d[0] * C[0][1] + d[1] * C[1][1];
	return FN(q_exp)((REAL)-.5 * (t0 * d[0] + t1 * d[1]));
}

/* get_2d_val_grad_ti, 2D/GSR.py:527-547 (== loop 1 of :266-281 and :378-395) */
static void FN(o2_point)(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot, const float *vals,
			 REAL tau, int dim, const REAL x[2], REAL *val, REAL *grad)
{
	int c[2];
	gsr_cell_of2_f32(g, (float)x[0], (float)x[1], c);
	for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
	for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++) {
		long cell = (long)gi * g->dims[1] + gj;
		for (int jj = 0; jj < g->cnt[cell]; jj++) {
			int i = g->sorted_id[g->offset[cell] + jj];
			REAL d[2] = {x[0] - (REAL)pos[2 * i], x[1] - (REAL)pos[2 * i + 1]};
			REAL C[2][2];
			FN(geom2)(rot[i], scal + 2 * i, C);
			REAL gaussian = FN(gauss2)(d, C);
			if (gaussian >= tau) {
				REAL gg0 = -gaussian * (C[0][0] * d[0] + C[0][1] * d[1]);
				REAL gg1 = -gaussian * (C[1][0] * d[0] + C[1][1] * d[1]);
				for (int dd = 0; dd < dim; dd++) {
					REAL v = vals[dim * i + dd];
					if (val) val[dd] += v * (gaussian - tau);
					if (grad) { grad[2 * dd] += v * gg0; grad[2 * dd + 1] += v * gg1; }
				}
			}
		}
	}
}

void FN(o2_forward)(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot, const float *vals,
		    double tau_d, int dim, const float *x, long Q, REAL *val, REAL *grad, int nthreads)
{
	REAL tau = (REAL)(float)tau_d;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		REAL xx[2] = {x[2 * j], x[2 * j + 1]};
		/* the 2D kernels zero their output rows themselves (2D/GSR.py:267-268, :379-380) */
		if (val) for (int dd = 0; dd < dim; dd++) val[dim * j + dd] = 0;
		if (grad) for (int dd = 0; dd < 2 * dim; dd++) grad[2 * dim * j + dd] = 0;
		FN(o2_point)(g, pos, scal, rot, vals, tau, dim, xx, val ? val + dim * j : NULL, grad ? grad + 2 * dim * j : NULL);
	}
}

#define ACC(ptr, inc) do { REAL inc__ = (inc); _Pragma("omp atomic") ptr += inc__; } while (0)

/* Loop 2 of the 2D value kernel get_losses_ti, 2D/GSR.py:282-339. weights = {weight, weight_boundary}. */
void FN(o2_backward_val)(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot, const float *vals,
			 double tau_d, int dim, const float *x, long Q, const REAL *val,
			 const float *ref, const float *normals, const float *normal_ref, const double *weights,
			 const int *stop_gradient, REAL *g_pos, REAL *g_scal, REAL *g_rot, REAL *g_val, int nthreads)
{
	const REAL tau = (REAL)(float)tau_d, weight = (REAL)(float)weights[0], weight_boundary = (REAL)(float)weights[1];
	if (weight == 0 && weight_boundary == 0) return;
	const REAL m = (REAL)Q;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		int c[2];
		gsr_cell_of2_f32(g, x[2 * j], x[2 * j + 1], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++) {
			long cell = (long)gi * g->dims[1] + gj;
			for (int jj = 0; jj < g->cnt[cell]; jj++) {
				int i = g->sorted_id[g->offset[cell] + jj];
				if (stop_gradient && stop_gradient[i]) continue;
				REAL d[2] = {(REAL)x[2 * j] - (REAL)pos[2 * i], (REAL)x[2 * j + 1] - (REAL)pos[2 * i + 1]};
				REAL C[2][2];
				FN(geom2)(rot[i], scal + 2 * i, C);
				REAL gaussian = FN(gauss2)(d, C);
				if (!(gaussian >= tau)) continue;
				REAL val_dot_normal = 0;
				for (int dd = 0; dd < dim; dd++) val_dot_normal += val[dim * j + dd] * (REAL)normals[dim * j + dd];
				REAL svn = FN(q_sign)(val_dot_normal - (REAL)normal_ref[j]);
				for (int dd = 0; dd < dim; dd++) {
					ACC(g_val[dim * i + dd], weight / ((REAL)2 * m) * (gaussian - tau) * FN(q_sign)(val[dim * j + dd] - (REAL)ref[dim * j + dd]));
					ACC(g_val[dim * i + dd], weight_boundary / m * svn * (gaussian - tau) * (REAL)normals[dim * j + dd]);
				}
				REAL value_dot_sign = 0, value_dot_normal = 0;
				for (int dd = 0; dd < dim; dd++) {
					value_dot_sign += (REAL)vals[dim * i + dd] * FN(q_sign)(val[dim * j + dd] - (REAL)ref[dim * j + dd]);
					value_dot_normal += (REAL)vals[dim * i + dd] * (REAL)normals[dim * j + dd];
				}
				REAL dgp[2] = {gaussian * (C[0][0] * d[0] + C[0][1] * d[1]), gaussian * (C[1][0] * d[0] + C[1][1] * d[1])};
				for (int k = 0; k < 2; k++) {
					ACC(g_pos[2 * i + k], weight / ((REAL)2 * m) * value_dot_sign * dgp[k]);
					ACC(g_pos[2 * i + k], weight_boundary / m * svn * value_dot_normal * dgp[k]);
				}
				REAL ct = FN(q_cos)((REAL)rot[i]), st = FN(q_sin)((REAL)rot[i]);
				REAL e0 = FN(q_exp)((REAL)2 * (REAL)scal[2 * i]), e1 = FN(q_exp)((REAL)2 * (REAL)scal[2 * i + 1]);
				REAL p0 = ct * d[0] + st * d[1], p1 = -st * d[0] + ct * d[1];
				REAL dgs[2] = {-gaussian * e0 * (p0 * p0), -gaussian * e1 * (p1 * p1)};
				for (int k = 0; k < 2; k++) {
					ACC(g_scal[2 * i + k], weight / ((REAL)2 * m) * value_dot_sign * dgs[k]);
					ACC(g_scal[2 * i + k], weight_boundary / m * svn * value_dot_normal * dgs[k]);
				}
				REAL s2 = FN(q_sin)((REAL)2 * (REAL)rot[i]), c2 = FN(q_cos)((REAL)2 * (REAL)rot[i]);
				/* trace( (d d^T) @ [[-s2, c2],[c2, s2]] ) */
				REAL tr = (d[0] * d[0] * -s2 + d[0] * d[1] * c2) + (d[1] * d[0] * c2 + d[1] * d[1] * s2);
				REAL dgr = (REAL)-.5 * gaussian * (e0 - e1) * tr;
				ACC(g_rot[i], weight / ((REAL)2 * m) * value_dot_sign * dgr);
				ACC(g_rot[i], weight_boundary / m * svn * value_dot_normal * dgr);
			}
		}
	}
}

/* Loop 2 of get_grad_losses_ti, 2D/GSR.py:396-476. weights = {grad, vor, div}. */
void FN(o2_backward_grad)(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot, const float *vals,
			  double tau_d, int dim, const float *x, long Q, const REAL *grad,
			  const float *ref_grad, const float *ref_vor, const double *weights, const int *stop_gradient,
			  REAL *g_pos, REAL *g_scal, REAL *g_rot, REAL *g_val,
			  REAL *vor_pos, REAL *vor_scal, REAL *vor_rot, REAL *vor_val,
			  REAL *div_pos, REAL *div_scal, REAL *div_rot, REAL *div_val, int nthreads)
{
	const REAL tau = (REAL)(float)tau_d;
	const REAL weight_grad = (REAL)(float)weights[0], weight_vor = (REAL)(float)weights[1], weight_div = (REAL)(float)weights[2];
	if (weight_grad == 0 && weight_vor == 0 && weight_div == 0) return;
	const REAL m = (REAL)Q;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		int c[2];
		gsr_cell_of2_f32(g, x[2 * j], x[2 * j + 1], c);
		const REAL *G = grad + 2 * dim * j;
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++) {
			long cell = (long)gi * g->dims[1] + gj;
			for (int jj = 0; jj < g->cnt[cell]; jj++) {
				int i = g->sorted_id[g->offset[cell] + jj];
				if (stop_gradient && stop_gradient[i]) continue;
				REAL d[2] = {(REAL)x[2 * j] - (REAL)pos[2 * i], (REAL)x[2 * j + 1] - (REAL)pos[2 * i + 1]};
				REAL C[2][2];
				FN(geom2)(rot[i], scal + 2 * i, C);
				REAL gaussian = FN(gauss2)(d, C);
				if (!(gaussian >= tau)) continue;
				REAL Cd[2] = {C[0][0] * d[0] + C[0][1] * d[1], C[1][0] * d[0] + C[1][1] * d[1]};
				REAL gg[2] = {-gaussian * Cd[0], -gaussian * Cd[1]};
				REAL sign_vor_diff = 0, div2 = 0, value[2] = {0, 0};
				if (dim == 2) {
					sign_vor_diff = FN(q_sign)((G[2] - G[1]) - (REAL)ref_vor[j]);
					div2 = (REAL)2 * (G[0] + G[3]);
					value[0] = vals[2 * i]; value[1] = vals[2 * i + 1];
				}
				REAL stv[2] = {0, 0}; /* sign_times_value */
				for (int dd = 0; dd < dim; dd++) {
					REAL s0 = FN(q_sign)(G[2 * dd] - (REAL)ref_grad[2 * dim * j + 2 * dd]);
					REAL s1 = FN(q_sign)(G[2 * dd + 1] - (REAL)ref_grad[2 * dim * j + 2 * dd + 1]);
					ACC(g_val[dim * i + dd], weight_grad / ((REAL)4 * m) * (s0 * gg[0] + s1 * gg[1]));
					stv[0] += s0 * (REAL)vals[dim * i + dd];
					stv[1] += s1 * (REAL)vals[dim * i + dd];
				}
				if (dim == 2) {
					ACC(vor_val[2 * i], weight_vor / m * sign_vor_diff * -gg[1]);
					ACC(vor_val[2 * i + 1], weight_vor / m * sign_vor_diff * gg[0]);
					ACC(div_val[2 * i], weight_div / m * div2 * gg[0]);
					ACC(div_val[2 * i + 1], weight_div / m * div2 * gg[1]);
				}
				/* d_grad_gaussian_position = gaussian * C @ (I - d d^T C)   (:434) */
				REAL ddC[2][2] = {{d[0] * Cd[0], d[0] * Cd[1]}, {d[1] * Cd[0], d[1] * Cd[1]}}; /* (d d^T) C, using symmetry as C^T d = C d */
				REAL ImD[2][2] = {{(REAL)1 - (d[0] * d[0] * C[0][0] + d[0] * d[1] * C[1][0]), -(d[0] * d[0] * C[0][1] + d[0] * d[1] * C[1][1])},
						  {-(d[1] * d[0] * C[0][0] + d[1] * d[1] * C[1][0]), (REAL)1 - (d[1] * d[0] * C[0][1] + d[1] * d[1] * C[1][1])}};
				(void)ddC;
				REAL P[2][2];
				for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++)
					P[a][b] = (gaussian * C[a][0]) * ImD[0][b] + (gaussian * C[a][1]) * ImD[1][b];
				REAL sdv[2] = {sign_vor_diff * ((REAL)0 * value[0] + (REAL)1 * value[1]), sign_vor_diff * ((REAL)-1 * value[0] + (REAL)0 * value[1])};
				REAL s2v[2] = {div2 * value[0], div2 * value[1]};
				for (int k = 0; k < 2; k++) {
					ACC(g_pos[2 * i + k], weight_grad / ((REAL)4 * m) * (P[0][k] * stv[0] + P[1][k] * stv[1]));
					if (dim == 2) {
						ACC(vor_pos[2 * i + k], weight_vor / m * (P[0][k] * sdv[0] + P[1][k] * sdv[1]));
						ACC(div_pos[2 * i + k], weight_div / m * (P[0][k] * s2v[0] + P[1][k] * s2v[1]));
					}
				}
				/* scalings :452-465 */
				REAL ct = FN(q_cos)((REAL)rot[i]), st = FN(q_sin)((REAL)rot[i]);
				REAL e[2] = {FN(q_exp)((REAL)2 * (REAL)scal[2 * i]), FN(q_exp)((REAL)2 * (REAL)scal[2 * i + 1])};
				REAL ax[2][2] = {{ct, st}, {-st, ct}};
				for (int k = 0; k < 2; k++) {
					REAL pk = ax[k][0] * d[0] + ax[k][1] * d[1];
					REAL dgs = -gaussian * e[k] * (pk * pk);
					REAL o0 = ((REAL)2 * e[k] * ax[k][0] * ax[k][0]) * d[0] + ((REAL)2 * e[k] * ax[k][0] * ax[k][1]) * d[1];
					REAL o1 = ((REAL)2 * e[k] * ax[k][1] * ax[k][0]) * d[0] + ((REAL)2 * e[k] * ax[k][1] * ax[k][1]) * d[1];
					REAL dggs[2] = {-dgs * Cd[0] - gaussian * o0, -dgs * Cd[1] - gaussian * o1};
					ACC(g_scal[2 * i + k], weight_grad / ((REAL)4 * m) * (dggs[0] * stv[0] + dggs[1] * stv[1]));
					if (dim == 2) {
						ACC(vor_scal[2 * i + k], weight_vor / m * (dggs[0] * sdv[0] + dggs[1] * sdv[1]));
						ACC(div_scal[2 * i + k], weight_div / m * (dggs[0] * s2v[0] + dggs[1] * s2v[1]));
					}
				}
				/* rotations :468-476 */
				REAL s2 = FN(q_sin)((REAL)2 * (REAL)rot[i]), c2 = FN(q_cos)((REAL)2 * (REAL)rot[i]);
				REAL dC[2][2] = {{(e[0] - e[1]) * -s2, (e[0] - e[1]) * c2}, {(e[0] - e[1]) * c2, (e[0] - e[1]) * s2}};
				REAL tr = (d[0] * d[0] * dC[0][0] + d[0] * d[1] * dC[1][0]) + (d[1] * d[0] * dC[0][1] + d[1] * d[1] * dC[1][1]);
				REAL dgr = (REAL)-.5 * gaussian * tr;
				REAL dggr[2] = {-dgr * Cd[0] - gaussian * (dC[0][0] * d[0] + dC[0][1] * d[1]),
						-dgr * Cd[1] - gaussian * (dC[1][0] * d[0] + dC[1][1] * d[1])};
				ACC(g_rot[i], weight_grad / ((REAL)4 * m) * (dggr[0] * stv[0] + dggr[1] * stv[1]));
				if (dim == 2) {
					ACC(vor_rot[i], weight_vor / m * (dggr[0] * sdv[0] + dggr[1] * sdv[1]));
					ACC(div_rot[i], weight_div / m * (dggr[0] * s2v[0] + dggr[1] * s2v[1]));
				}
			}
		}
	}
}

/* advection_rk4_ti, 2D/GSR.py:549-580 */
void FN(o2_rk4)(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot, const float *vals,
		double tau_d, const float *start, long Q, double dt_d,
		REAL *goal_pos, REAL *deformation, REAL *goal_val, REAL *goal_grad, int nthreads)
{
	const REAL tau = (REAL)(float)tau_d, dt = (REAL)(float)dt_d;
	#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
	for (long j = 0; j < Q; j++) {
		REAL x[2] = {start[2 * j], start[2 * j + 1]};
		REAL v[4][2] = {{0}}, dv[4][4] = {{0}}, p[2];
		FN(o2_point)(g, pos, scal, rot, vals, tau, 2, x, v[0], dv[0]);
		for (int k = 0; k < 2; k++) p[k] = x[k] + dt * (REAL).5 * v[0][k];
		FN(o2_point)(g, pos, scal, rot, vals, tau, 2, p, v[1], dv[1]);
		for (int k = 0; k < 2; k++) p[k] = x[k] + dt * (REAL).5 * v[1][k];
		FN(o2_point)(g, pos, scal, rot, vals, tau, 2, p, v[2], dv[2]);
		for (int k = 0; k < 2; k++) p[k] = x[k] + dt * v[2][k];
		FN(o2_point)(g, pos, scal, rot, vals, tau, 2, p, v[3], dv[3]);
		REAL phi[2];
		for (int k = 0; k < 2; k++) {
			phi[k] = x[k] + dt / (REAL)6 * (v[0][k] + (REAL)2 * v[1][k] + (REAL)2 * v[2][k] + v[3][k]);
			goal_pos[2 * j + k] = phi[k];
		}
		if (deformation) {
#define MM(r, a, b) do { (r)[0] = (a)[0] * (b)[0] + (a)[1] * (b)[2]; (r)[1] = (a)[0] * (b)[1] + (a)[1] * (b)[3]; \
			 (r)[2] = (a)[2] * (b)[0] + (a)[3] * (b)[2]; (r)[3] = (a)[2] * (b)[1] + (a)[3] * (b)[3]; } while (0)
			const REAL I[4] = {1, 0, 0, 1};
			REAL dphi1[4], a1[4], dphi2[4], a2[4], dphi3[4], a3[4];
			for (int k = 0; k < 4; k++) dphi1[k] = I[k] + dt * (REAL).5 * dv[0][k];
			MM(a1, dv[1], dphi1);
			for (int k = 0; k < 4; k++) dphi2[k] = I[k] + dt * (REAL).5 * a1[k];
			MM(a2, dv[2], dphi2);
			for (int k = 0; k < 4; k++) dphi3[k] = I[k] + dt * a2[k];
			MM(a3, dv[3], dphi3);
			for (int k = 0; k < 4; k++)
				deformation[4 * j + k] = I[k] + dt / (REAL)6 * (dv[0][k] + (REAL)2 * a1[k] + (REAL)2 * a2[k] + a3[k]);
#undef MM
		}
		if (goal_val && goal_grad) {
			REAL vp[2] = {0, 0}, dvp[4] = {0, 0, 0, 0};
			FN(o2_point)(g, pos, scal, rot, vals, tau, 2, phi, vp, dvp);
			goal_val[2 * j] = vp[0]; goal_val[2 * j + 1] = vp[1];
			for (int k = 0; k < 4; k++) goal_grad[4 * j + k] = dvp[k];
		}
	}
}

#undef ACC
#undef FN
#undef CAT
#undef CAT_
