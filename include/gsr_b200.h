/*
 * gsr_b200.h — C ABI of the B200-native Gaussian-Spatial-Representation engine.
 *
 * This is the drop-in boundary for the hot path of DvvCz/Gaussian-Fluids-Code: one entry point per
 * Taichi kernel of the reference (3D/GSR.py, 2D/GSR.py) plus the fused per-iteration optimisation
 * step of advance.py.  Plain pointers and sizes only — no torch types.  All pointers are DEVICE
 * pointers unless the parameter is a `const gsr_grid_desc*` or `const gsr_*_cfg*` (host structs).
 * Every function enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises
 * the host, never allocates (scratch comes from the caller's `ws` buffer, sized by
 * gsr_*_ws_bytes), and is CUDA-graph capturable.  Return value: 0 = ok, otherwise a negative
 * GSR_E* code or a positive cudaError_t.
 *
 * Layouts (all float32, C-contiguous, as the reference's torch tensors):
 *   3D: positions (N,3) scalings (N,3) rotations (N,4 quaternion [w,x,y,z]) values (N,3)
 *   2D: positions (N,2) scalings (N,2) rotations (N,)  angle                values (N,2)
 *   val (Q,D)   grad (Q,D,D) with grad[j,d,l] = d u_d / d x_l.
 *
 * The engine's own hash representation (built by gsr_build_grid + gsr_pack_gaussians):
 *   cell_start (ncell+1) int32 : exclusive prefix of the per-cell counts in row-major cell order
 *                                (== the reference's grid_offset, plus the total at [ncell]);
 *   sorted_id  (N) int32       : Gaussian ids ordered by cell, ascending id inside a cell
 *                                (== the reference's sorted_id, canonical intra-cell order);
 *   packed     (N * GSR_PACK_FLOATS(D)) float32 : per-Gaussian {mu, Sigma^-1, v} in cell order.
 */
#ifndef GSR_B200_H
#define GSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_OK 0
#define GSR_EINVAL (-1)	/* bad argument (D, dims, NULL pointer, ...) */
#define GSR_EWS (-2)	/* workspace too small */

#define GSR_PACK_FLOATS(D) ((D) == 3 ? 12 : 8)

/* Spatial-hash descriptor.  Mirrors the constants the reference bakes into its kernels
 * (3D/GSR.py:160-177, 2D/GSR.py:177-192): the extended domain rounded to f32, grid_size, tau,
 * and the run-time grid_scale (3D/GSR.py:247-252). */
typedef struct {
	int32_t D;		/* 2 or 3 */
	int32_t dims[3];	/* grid_size; dims[2] ignored for D == 2 */
	float lo[3], hi[3];	/* x_min.. / x_max.. of the EXTENDED domain */
	float grid_scale;
	float tau;		/* clamp_threshold */
	const float *grid_scale_dev;	/* optional DEVICE scalar: when non-NULL the kernels read grid_scale from it
					 * (kept up to date by gsr_step, so the optimisation loop never syncs the host) */
} gsr_grid_desc;

/* ---- a1: spatial hash  (reinitialize_grid_ti: 3D/GSR.py:205-245, 2D/GSR.py:194-222) ------------ */
size_t gsr_build_grid_ws_bytes(const gsr_grid_desc *g, int64_t N);
/* Outputs: cell_start (ncell+1), sorted_id (N; entries past cell_start[ncell] list the Gaussians outside the hash),
 * optionally the reference-format grid_cnt (ncell) and grid_offset (ncell) (may be NULL), and — when `packed` is given
 * together with scalings / rotations / values — the packed records and cull coefficients of gsr_pack_gaussians (for
 * small N the whole hash and the packing are ONE single-CTA launch). */
int gsr_build_grid(const gsr_grid_desc *g, const float *positions, int64_t N,
		   int32_t *cell_start, int32_t *sorted_id, int32_t *grid_cnt, int32_t *grid_offset,
		   const float *scalings, const float *rotations, const float *values, float *packed, float *cull,
		   void *ws, size_t ws_bytes, void *stream);

/* Per-Gaussian precompute {mu, Sigma^-1 = R diag(e^{2s}) R^T, v}, gathered into cell order
 * (hoists the per-pair recomputation of 3D/GSR.py:277-289, 2D/GSR.py:274-277 out of the pair loop).
 * cull (N floats, cell order, may be NULL): (1 + margin) / lambda_min(Sigma^-1) — the bounding-sphere coefficient the
 * large-Q kernels use to skip, per warp, candidates that no point of the warp can accept. */
int gsr_pack_gaussians(const gsr_grid_desc *g, const float *positions, const float *scalings, const float *rotations,
		       const float *values, int64_t N, const int32_t *cell_start, const int32_t *sorted_id,
		       float *packed, float *cull, void *stream);

/* min over all entries of scalings -> *out_min (device float); the reduction behind
 * `self.scalings.min().item()` of reinitialize_grid (3D/GSR.py:249). */
int gsr_min_scaling(const float *scalings, int64_t count, float *out_min, void *stream);

/* ---- sample binning (engine-internal ordering of the query points) ------------------------------ */
size_t gsr_bin_samples_ws_bytes(const gsr_grid_desc *g, int64_t Q);
/* perm (Q): sample indices ordered by (padded) cell key, stable;  sample_cell_start: (pcell+1)
 * with pcell = prod(dims+2), may be NULL when only the ordering is needed.  fine != 0: samples of one cell are further
 * ordered by their position on a 4^D sub-cell raster (spatially compact warps for the tiled kernels). */
int gsr_bin_samples(const gsr_grid_desc *g, const float *x, int64_t Q, int32_t *perm, int32_t *sample_cell_start, int fine,
		    void *ws, size_t ws_bytes, void *stream);
int64_t gsr_padded_cells(const gsr_grid_desc *g);

/* Tiles of the large-Q evaluation kernels: runs of up to GSR_TILE_SAMPLES cell-sorted samples that share one (x, y)
 * row of cells, so that their candidate Gaussians are 9 contiguous runs of packed records that one CTA stages in
 * shared memory with TMA bulk copies (3D; the 2D kernels keep the one-thread-per-point shape).  tile_row (gsr_tile_slots(g, Q) int32) maps a
 * tile id to its row (-1: unused id); built from sample_cell_start. */
#define GSR_TILE_SAMPLES 512
int64_t gsr_tile_slots(const gsr_grid_desc *g, int64_t Q);
int gsr_build_tiles(const gsr_grid_desc *g, const int32_t *sample_cell_start, int64_t Q, int32_t *tile_row, void *stream);

/* Engine tunables (process-wide; tests use them to force a code path at small sizes). */
#define GSR_TUNE_TILED_MIN_Q 1	/* smallest Q evaluated by the tiled kernels when a tile table is supplied */
#define GSR_TUNE_TILED_CAP 2	/* shared-memory staging capacity of a tile, in Gaussians */
#define GSR_TUNE_FW_P4_MIN_SPC 4	/* samples per cell from which the tiled forward takes 4 points per thread (default 0: always), else 2 */
#define GSR_TUNE_FORCE_RADIX 5	/* 1: always build hashes with the radix sort (tests compare it with the single-launch and counting paths) */
#define GSR_TUNE_STEP_SMALL_N 6	/* up to this many Gaussians (at most 2048, the default) gsr_step is ONE launch of one 8-CTA cluster; 0: always the four-launch form */
#define GSR_TUNE_STEP_LANES4 9	/* 1 (default): the cluster step gives four lanes to a Gaussian when N <= 1024; 0: one thread per Gaussian */
#define GSR_TUNE_STEP_FUSED_HASH 10	/* 1 (default): gsr_step_rebuild's four-lane cluster step also rebuilds the hash and the packed records (<= 1024 cells); 0: second launch */
#define GSR_TUNE_RK4S_CAP 11	/* staging capacity (Gaussians) of the RK4 pull-back kernel that keeps its integrator state in shared memory (default 128: four CTAs per SM) */
#define GSR_TUNE_RK4S_MIN_SPC 12	/* samples per hash cell from which the RK4 pull-back uses the shared-memory-state kernel (4 points per thread), else 2 points per thread in registers */
#define GSR_TUNE_GATHER_CTA_MAX_N 7	/* 3D backward gather: up to this many Gaussians, with >= 4 samples per Gaussian, one CTA per Gaussian (default 4096; 0: never) */
#define GSR_TUNE_LANES8_MIN_N 8	/* items (points / Gaussians) from which the latency-shape kernels give 8 lanes to an item instead of a warp (default 16384) */
#define GSR_TUNE_RK4_SMEM_STATE 3	/* 1 (default): tiled RK4 keeps the integrator state in shared memory, 4 points per thread; 0: registers, 2 per thread */
/* gsr_sample_box_surface followed by gsr_bin_samples(fine = 0) on the drawn points, as ONE launch when the batch fits the
 * cluster sort (n <= 16384 on <= 8191 padded cells; otherwise the two calls are made internally): data, normal (n,3), perm (n),
 * sample_cell_start (padded cells + 1); ws as for gsr_bin_samples.  Same draws, same order as the two separate calls. */
int gsr_sample_box_surface_binned(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev,
				  float *data, float *normal, const gsr_grid_desc *g, int32_t *perm, int32_t *sample_cell_start,
				  void *ws, size_t ws_bytes, void *stream);

/* ---- N2: analytic initial fields (3D/init_cond.py:122-145, Taichi kernels vortex_particle / vortex_particle_gradient) ----
 * Regularised Biot-Savart sum over M vortex particles (x0 (M,3), w (M,3) strength-scaled tangents, U = radius / (2 n),
 * a = thickness): val (Q,3) += U f(r) (w x d), grad (Q,3,3) += its Jacobian; either output may be NULL.  Accumulates, like
 * the reference's kernels (callers zero the outputs, 3D/init_cond.py:156, :169).  One launch. */
int gsr_vortex_particles(const float *x, int64_t Q, const float *x0, const float *w, int64_t M, float U, float a,
			 float *val, float *grad, void *stream);

/* ---- N3: mesh boundary sampler (3D/mesh_sampler.py:12-21, :60-88) --------------------------------------------------------
 * gsr_mesh_tri_areas: area (F) of every triangle (faces (F,3) int32 into vertices (V,3)); the caller forms the inclusive
 * prefix sums (ti_get_tri_area's second loop).
 * gsr_sample_mesh: n points on the mesh, triangles weighted by area (binary search in area_presum), uniform on the triangle,
 * vertex normals (facenormals (F,3) into normals) interpolated and normalised.  Uniforms: Philox keyed like gsr_sample_box, or,
 * when `uniforms` (n,3) is given, exactly those three draws per sample in the reference's order (face, u, v).  One launch. */
int gsr_mesh_tri_areas(const float *vertices, const int32_t *faces, int64_t F, float *area, void *stream);
int gsr_sample_mesh(int64_t n, const float *vertices, const float *normals, const int32_t *faces, const int32_t *facenormals,
		    const float *area_presum, int64_t F, uint64_t seed, uint32_t stream_id, const float *iteration_dev,
		    const float *uniforms, float *data, float *normal, void *stream);

int gsr_set_tuning(int key, int value);

/* ---- a2: forward  (loop 1 of get_losses_ti 3D/GSR.py:270-298; 2D/GSR.py:266-281, :378-395) ------ */
/* perm may be NULL (process samples in the given order).  val and/or grad may be NULL.
 * accumulate != 0: outputs are += (3D reference semantics); 0: overwritten (2D semantics).
 * sample_cell_start + tile_row (both from gsr_bin_samples / gsr_build_tiles for this x and this grid) may be NULL;
 * when given and Q is large, the tiled shared-memory kernels are used (same results); cull (from gsr_pack_gaussians, may be
 * NULL) enables their warp-level candidate culling. */
int gsr_forward(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
		const float *x, int64_t Q, const int32_t *perm, const int32_t *sample_cell_start, const int32_t *tile_row,
		float *val, float *grad, int accumulate, void *stream);

/* ---- a4: RK4 advection + pull-back  (advection_rk4_ti 3D/GSR.py:634-665; 2D/GSR.py:549-580) ----- */
/* deformation / goal_val / goal_grad may be NULL (the reference's size-0 outputs). */
int gsr_rk4(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
	    const float *start, int64_t Q, const int32_t *perm, const int32_t *sample_cell_start, const int32_t *tile_row, float dt,
	    float *goal_pos, float *deformation, float *goal_val, float *goal_grad, void *stream);

/* ---- a5: advected-covector reference  (AdvectedCovectorField.vorticity 3D/advance.py:24-47,
 *          2D/advance.py:46-54): RK4 back-trace fused with curl, 3x3 inverse and helicity -------- */
/* 3D: ref_vor (Q,3), ref_hel (Q) | NULL.  2D: ref_vor (Q), zeroed where the back-traced point leaves
 * domain[4] = {x_min,x_max,y_min,y_max} (host pointer, may be NULL); ref_hel ignored. */
int gsr_advected_vorticity(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
			   const float *x, int64_t Q, const int32_t *perm, const int32_t *sample_cell_start, const int32_t *tile_row,
			   float dt, const float *domain,
			   float *ref_vor, float *ref_hel, void *stream);

/* ---- a6: neighbour marking  (get_all_neighbors_ti 3D/GSR.py:679-690; 2D/GSR.py:620-630) --------- */
int gsr_mark_neighbors(const gsr_grid_desc *g, const int32_t *cell_start, const int32_t *sorted_id, const float *packed,
		       const float *x, int64_t Q, int32_t *mark, void *stream);

/* ---- a3: backward  (loop 2 of get_losses_ti 3D/GSR.py:299-540; 2D/GSR.py:282-339, :396-476) ----- */
typedef struct {
	/* 3D: weight_val, weight_boundary, weight_grad, weight_vor, weight_hel, weight_div.
	 * 2D: `weight`, weight_boundary, weight_grad, weight_vor, (unused), weight_div — the value kernel
	 *     and the gradient kernel of 2D/GSR.py share this struct (set the other kernel's weights to 0). */
	float w_val, w_boundary, w_grad, w_vor, w_hel, w_div;
	int64_t Q_norm;		/* the Q of the loss normalisers (global sample count when sharded) */
	/* reference inputs, each (Q, ...) or NULL when its weight is 0 */
	const float *ref_val;	/* (Q,D) */
	const float *normals;	/* (Q,D) */
	const float *normal_ref;	/* 2D only: (Q) target of u.n (2D/GSR.py:302) */
	const float *ref_grad;	/* (Q,D,D) */
	const float *ref_vor;	/* 3D (Q,3); 2D (Q) */
	const float *ref_hel;	/* 3D (Q) */
	const int32_t *stop_gradient;	/* (N) int32 or NULL */
	const float *sample_grid_scale_dev;	/* optional DEVICE scalar: the grid scale `sample_cell_start` was built with (gsr_bin_samples with a
				 * descriptor pointing at it) when that is not the hash's own — it must be >= the hash's grid_scale: a Gaussian then
				 * still finds every sample of its support in the 3^D sample cells around its own */
	float *loss_partials;	/* optional out, device (gsr_loss_blocks(Q), 8): per-block partial sums of the sample losses,
				 * reduced deterministically by gsr_step / gsr_sample_losses.  Slots (sums over the samples):
				 * 0 mean_k|omega-omega_ref| (3D) or |omega-omega_ref| (2D)   1 |u.omega - hel_ref|   2 (div u)^2
				 * 3 |u.n| (3D) or |u.n - normal_ref| (2D)   4 mean_d|u-ref_val|   5 mean|grad u - ref_grad|   6,7 unused */
} gsr_loss_cfg;

int64_t gsr_loss_blocks(int64_t Q);
/* The sample losses alone (no gradient): sums[8] (device) = slots above summed over the Q samples.
 * ws: gsr_loss_blocks(Q)*8 floats. */
int gsr_sample_losses(const gsr_grid_desc *g, int64_t Q, const float *val, const float *grad, const gsr_loss_cfg *cfg,
		      float *sums, void *ws, size_t ws_bytes, void *stream);

/* number of compact accumulator sets the gather produces: direct, vor(+hel), div */
#define GSR_NSETS 3
/* floats per Gaussian per set: d/dv (D), d/dmu (D), d/dSigma^-1 (D(D+1)/2) */
#define GSR_ACC_FLOATS(D) ((D) == 3 ? 12 : 7)

size_t gsr_backward_ws_bytes(const gsr_grid_desc *g, int64_t N, int64_t Q);
/*
 * Stage 1+2: per-sample adjoints, then the atomics-free Gaussian-centric gather.
 * acc: (GSR_NSETS, N, GSR_ACC_FLOATS(D)) in ORIGINAL Gaussian id order, overwritten (zero for
 * out-of-domain or stop_gradient Gaussians).  `sets_mask` (host, out): bit s set when set s is active.
 * This buffer is what a multi-GPU run all-reduces (sum) across sample shards.
 */
int gsr_backward_gather(const gsr_grid_desc *g, const int32_t *cell_start, const int32_t *sorted_id, const float *packed, int64_t N,
			const float *x, int64_t Q, const int32_t *perm, const int32_t *sample_cell_start,
			const float *val, const float *grad, const gsr_loss_cfg *cfg,
			float *acc, int *sets_mask, void *ws, size_t ws_bytes, void *stream);
/*
 * Stage 3: per-Gaussian chain rule Sigma^-1 -> (scalings, rotations) and accumulation (+=) into the
 * reference's gradient buffers.  out[s] = {positions, scalings, rotations, values} grads of set s
 * (s = 0 direct, 1 vor, 2 div); buffers of different sets MAY alias (3D/GSR.py:564-579).
 */
int gsr_backward_epilogue(const gsr_grid_desc *g, const float *scalings, const float *rotations, int64_t N,
			  const float *acc, int sets_mask, float *const out[GSR_NSETS][4], void *stream);

/* ---- a7: fused per-iteration optimiser step of project()  (3D/advance.py:183-287, GSR.py:148-152,
 *          :704-716; 2D/advance.py:187-259) -------------------------------------------------------- */
typedef struct {
	int32_t D;
	float lr[4];		/* initial lrs: positions, scalings, rotations, values */
	float beta1, beta2, eps;	/* torch.optim.Adam defaults .9 .999 1e-8 */
	/* torch ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0): factor .9, threshold 1e-4, eps 1e-8 */
	float sched_factor, sched_threshold, sched_eps, sched_min_lr;
	int32_t sched_patience;
	float w_aniso, w_vol, w_valreg, w_dpos;	/* regulariser weights (3D: 10,10,0,-; 2D: 10,10,-,.5) */
	float aniso_ratio;	/* 1.5 */
	int32_t pcgrad;		/* 1: mutual projection of the vor and div sets (3D/advance.py:202-225) */
	double grid_coef;	/* sqrt(-2 ln tau) in host double (3D/GSR.py:249); 0 when tau == 0 */
	double min_grid_scale;
	double grid_scale_tau0;	/* the constant grid_scale used when tau == 0 (3D/GSR.py:251) */
	int32_t keep_clock;	/* gsr_step_init: leave GSR_ST_CLOCK as it is (the state has been initialised before) */
	float *sample_gs_slots;	/* optional DEVICE float[2]: a SAMPLE grid scale per iteration parity, one iteration ahead of the hash.  Step k of a
				 * phase (k = 0, 1, ...) writes slot k & 1 = sample_gs_margin x the grid_scale it leaves — the scale the caller bins
				 * the samples of iteration k + 2 with while iteration k + 1 still runs (gsr_loss_cfg.sample_grid_scale_dev tells the
				 * gather) — after checking that slot (k + 1) & 1, which the samples of iteration k + 1 were binned with, still covers
				 * that grid_scale (else state[GSR_ST_SGS_ERR] = 1, sticky).  gsr_step_init sets both slots. */
	float sample_gs_margin;	/* > 1: bound on the growth of grid_scale over one step (Adam moves a log-radius by < 3.2 lr) */
	float *grid_scale_out;	/* optional DEVICE scalar that also receives the next grid_scale: the hash descriptor's grid_scale_dev
				 * of the field being optimised, so that every kernel launched on that field — also from a CUDA
				 * graph captured earlier — bins with the current value */
} gsr_step_cfg;

/* one source of sample-loss partial sums and its weights in the scheduler metric:
 * loss_tot += sum_k w[k] * (sum over blocks of partials[.,k]) */
typedef struct {
	const float *partials;
	int32_t nblocks;
	float w[8];
} gsr_loss_src;

/* Device-resident optimiser state (float32): [GSR_STATE_SCALARS scalars][Adam m: P*N][Adam v: P*N], P = 13 (3D) / 7 (2D),
 * parameter order positions, scalings, rotations, values.  Scalars: */
#define GSR_STATE_SCALARS 64
#define GSR_ST_T 0		/* Adam step count */
#define GSR_ST_BEST 1		/* scheduler: best metric */
#define GSR_ST_BAD 2		/* scheduler: num_bad_epochs */
#define GSR_ST_LR 3		/* [3..6] current lr per group */
#define GSR_ST_GRID_SCALE 7	/* grid_scale for the NEXT hash build (point gsr_grid_desc.grid_scale_dev here) */
#define GSR_ST_MIN_S 8		/* min over scalings after the update */
#define GSR_ST_LOSS_TOT 9	/* the scheduler metric of this iteration */
#define GSR_ST_L_ANISO 10
#define GSR_ST_L_VOL 11
#define GSR_ST_L_VALREG 12
#define GSR_ST_L_DPOS 13
#define GSR_ST_LOSS_SRC 14	/* [14..21] sum over sources of the raw loss slots divided by nothing (plain sums) */
#define GSR_ST_BPOW 48		/* [48..51] two DOUBLES (8-byte aligned): beta1^t, beta2^t of Adam's bias corrections, advanced by one multiplication per step */
#define GSR_ST_CLOCK 22
#define GSR_ST_SGS_ERR 23	/* 1: grid_scale outgrew the sample grid scale of a pre-binned batch (sample_gs_slots) */		/* steps taken since the state was created: like GSR_ST_T, but gsr_step_init keeps it when
				 * cfg->keep_clock != 0 — the iteration number for the sample generators, so that successive
				 * optimisation phases (time steps) draw fresh samples (the reference draws from one running RNG) */
size_t gsr_step_state_floats(int D, int64_t N);
size_t gsr_step_ws_bytes(int D, int64_t N);
/* zero the moments, t = 0, best = +inf, lrs = cfg->lr, grid_scale from the current scalings */
int gsr_step_init(const gsr_step_cfg *cfg, int64_t N, const float *scalings, float *state, void *stream);
/*
 * One optimiser iteration on device with no host sync (4 launches):
 *   A  per Gaussian: chain rule of the vor and div sets, block partial sums of the 12 PCGrad dot products,
 *      the volume moments and the regulariser losses;
 *   R  one block: deterministic reduction, PCGrad coefficients, loss_tot, Adam bias corrections,
 *      ReduceLROnPlateau update (the lr used by this step is the one before the update, as in torch);
 *   B  per Gaussian: total gradient = projected vor + div sets + extra direct sets (boundary passes)
 *      + closed-form regulariser gradients; Adam; parameters updated in place; min over the new scalings;
 *   S  next grid_scale = max(grid_coef * exp(-min s), min_grid_scale) in double, rounded to f32.
 * acc/sets_mask: from gsr_backward_gather (already all-reduced when sharded).  extra_direct[k]: accumulator buffers
 * whose set 0 (direct) is added (boundary passes), or NULL.  positions_org: 2D position-drift anchor or NULL.
 */
int gsr_step(const gsr_step_cfg *cfg, int64_t N, float *positions, float *scalings, float *rotations, float *values,
	     const float *acc, int sets_mask, const float *const extra_direct[2], const gsr_loss_src *loss_src, int n_loss_src,
	     const float *positions_org, float *state, void *ws, size_t ws_bytes, void *stream);

/* gsr_step followed by the hash rebuild and the packed records of the UPDATED Gaussians (the reference's step() ->
 * zero_grad() -> reinitialize_grid(), 3D/GSR.py:704-716) in one call.  (A single-CTA fusion of the five kernels was measured
 * SLOWER at N = 1000 — 50 us against 35 us: the per-Gaussian dependency chains, not the launches, set the latency.)
 * g->grid_scale_dev should point at state + GSR_ST_GRID_SCALE so that the rebuilt hash uses the new grid_scale.
 * hash_ws: gsr_build_grid_ws_bytes(g, N). */
int gsr_step_rebuild(const gsr_step_cfg *cfg, int64_t N, float *positions, float *scalings, float *rotations, float *values,
		     const float *acc, int sets_mask, const float *const extra_direct[2], const gsr_loss_src *loss_src, int n_loss_src,
		     const float *positions_org, float *state, void *ws, size_t ws_bytes,
		     const gsr_grid_desc *g, int32_t *cell_start, int32_t *sorted_id, float *packed, float *cull, void *hash_ws, size_t hash_ws_bytes,
		     void *stream);

/* ---- multi-GPU exchange: out[i] = sum over ranks r = 0..world-1, in rank order, of peer_bufs[r][i] — ONE kernel over
 *          NVLink peer memory instead of a library all-reduce.  peer_bufs / peer_signal_pads: HOST arrays of `world` device pointers
 *          (this rank's own included) into symmetric memory; a signal pad holds >= world uint32, zero-initialised.  The epoch of a call
 *          is *epoch_base_dev + (uint32)*iteration_dev + 1 and must grow by one per call on every rank (buffers alternate between two
 *          parities in the caller).  n_floats a multiple of 4, buffers 16-byte aligned.  *err_flag (device) is set to 1 when a peer
 *          did not arrive within ~1 s (the kernel never hangs). */
int gsr_xrank_sum(const void *const *peer_bufs, void *const *peer_signal_pads, int rank, int world, int64_t n_floats, const float *iteration_dev,
		  const int32_t *epoch_base_dev, float *out, int32_t *err_flag, void *stream);

/* ---- a7: sample generation  (3D/advance.py:339-340 rand_like(positions) * extent + min; 3D/init_cond.py:227-249
 *          sample_on_box) — one kernel per sample set, counter-based Philox keyed by (seed, stream_id) and indexed by
 *          (sample, iteration); the iteration number is read from DEVICE memory (e.g. state + GSR_ST_T, may be NULL = 0)
 *          so a captured CUDA graph draws fresh samples on every replay. box = {x_min, x_max, y_min, y_max, z_min, z_max} (host). */
int gsr_sample_box(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev, float *out, void *stream);
/* area-weighted points on the six faces with inward unit normals: data (n,3), normal (n,3) */
int gsr_sample_box_surface(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev, float *data, float *normal,
			   void *stream);

/* ---- a8 / N4: reseeding of over-stretched Gaussians (clone_velocity_field, 3D/advance.py:51-94, 2D/advance.py:58-93) ---------------
 * gsr_split_flags: flags[i] = exp(max_k s_ik - min_k s_ik) >= ratio_threshold (2 in 3D, 1.5 in 2D); *count (device) = how many.
 * gsr_split_apply: the N + n_split Gaussians after the split — kept ones first in their old order, then the children as
 *   [first samples of all parents | second samples of all parents] (the reference's `.sample((2,)).flatten(0, 1)` / `.repeat(2, 1)`):
 *   child position = mu + chol(Sigma) z, Sigma^-1 = R diag(e^{2s}) R^T, clamped to clamp_box (host, {x_min, x_max, ...}; NULL: no
 *   clamp, as in 2D); child scalings = the parent's with log_axis added on the longest axis and log_all subtracted everywhere
 *   (3D: ln 2, ln 2 / 3; 2D: ln 1.5, 0); rotations and values copied; out_stop_gradient = 1 for kept, 0 for children.
 *   flag_prefix: exclusive scan of flags.  normals: (2, n_split, D) standard normals (device) or NULL = Philox + Box-Muller from `seed`. */
int gsr_split_flags(int D, const float *scalings, int64_t N, float ratio_threshold, int32_t *flags, int32_t *count, void *stream);
int gsr_split_apply(int D, const float *positions, const float *scalings, const float *rotations, const float *values, int64_t N,
		    const int32_t *flags, const int32_t *flag_prefix, int64_t n_split, const float *normals, uint64_t seed,
		    const float *clamp_box, float log_axis, float log_all,
		    float *out_positions, float *out_scalings, float *out_rotations, float *out_values, int32_t *out_stop_gradient, void *stream);

/* ---- N1: passive density advection on a regular lattice  (advected_density + ti_get_interp_val, 3D/advance_density.py:25-63):
 *          per voxel, RK4 back-trace by `dt` (callers pass -dt) through the field, clamp to `domain`, trilinear resampling of the
 *          old density — one kernel, no lattice or back-traced coordinates in memory.  xs/ys/zs: the lattice's axis coordinates
 *          (device, nx / ny / nz floats, the reference's torch.linspace); domain = {x_min, x_max, y_min, y_max, z_min, z_max} (host);
 *          density_* / out_* (nx, ny, nz) float32, out must not alias in; the second field is optional (both NULL). */
int gsr_advect_density(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
		       const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, const float *domain, float dt,
		       const float *density_a, const float *density_b, float *out_a, float *out_b, void *stream);
/* the same for the x planes [x_begin, x_end) only — one process's slab of a lattice shared between GPUs (the fields keep their full
 * (nx, ny, nz) shape: the back-trace of a slab voxel may read the old density a few planes outside the slab) */
int gsr_advect_density_slab(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
			    const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, int x_begin, int x_end, const float *domain, float dt,
			    const float *density_a, const float *density_b, float *out_a, float *out_b, void *stream);

/* ---- measurement helpers (bench.py work census and roofline denominators) ------------------------ */
/* counts[0] (device uint64) += candidate visits C for one evaluation of the Q points (occupancy of each point's 27 (9)-cell
 * stencil under the reference binning: the unit of work of SURVEY 8d); when packed != NULL also counts[1] += accepted pairs P. */
int gsr_count_pairs(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *x, int64_t Q, uint64_t *counts, void *stream);
/* gsr_advect_density_slab, and *executed_pair_tests (device uint64) += the (voxel, candidate) pair tests the kernel really runs: a
 * candidate that survives the warp-level culling is tested on all 32 x 4 voxels of the warp, four times (RK4).  bench.py forms the
 * density kernel's roofline fraction from this count; gsr_count_pairs over the same voxels gives the un-culled stencil occupancy. */
int gsr_advect_density_census(const gsr_grid_desc *g, const int32_t *cell_start, const float *packed, const float *cull,
			      const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, int x_begin, int x_end, const float *domain, float dt,
			      const float *density_a, const float *density_b, float *out_a, float *out_b, unsigned long long *executed_pair_tests, void *stream);
/* runs an FFMA-only / ex2.approx-only loop on every SM; returns elapsed ms via *ms (host sync inside) */
int gsr_peak_fma(int iters, double *tflops, void *stream);
int gsr_peak_mufu(int iters, double *tops, void *stream);
/* pipe probes (misc.cu): thread-level operations per second of one operand shape; which = 0..6, see misc.cu */
int gsr_pipe_probe(int which, int iters, double *ops_per_s, void *stream);

/* number of kernels this library has launched so far (host-side counter; bench.py's gpu_launches) */
uint64_t gsr_launch_count(void);

const char *gsr_version(void);

#ifdef __cplusplus
}
#endif
#endif
