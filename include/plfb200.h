/*
 * plfb200.h -- C ABI of libplfb200.so: the B200 (sm_100a) implementation of pyLabFEA's per-element
 * elastic-plastic stress update and global stiffness assembly.
 *
 * The reference (pyLabFEA v4.4.2, pure Python) has no FFI of its own; its boundary for this path is the Python
 * class API.  Every entry point below names the reference call site it replaces (paths relative to
 * /root/reference/src/pylabfea/).  INTEGRATION.md shows the ctypes binding a pyLabFEA maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point data is IEEE float64, indices are int32 (int64 counts).
 *   - Voigt order (11,22,33,23,13,12) with engineering shear, 6x6 matrices row-major, as in the reference.
 *   - host arrays are borrowed for the duration of the call; device buffers are owned by the context.
 *   - every function returns PLF_OK (0) or a negative error code; plf_last_error() gives the message.
 *     The Python wrapper maps PLF_EINVAL -> ValueError, PLF_ESTATE -> AttributeError (solve before mesh,
 *     model.py:1025), PLF_EUNSUPPORTED -> NotImplementedError / ValueError as the reference raises them.
 *   - one context = one CUDA device + one stream; calls on a context are serialised by the caller.
 *   - there is NO CPU fallback: without a CUDA device plf_create() fails with PLF_ECUDA.
 */
#ifndef PLFB200_H
#define PLFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLF_OK 0
#define PLF_EINVAL (-1)       /* bad argument (shape, id, NULL)                                  */
#define PLF_ECUDA (-2)        /* CUDA runtime error or no device                                  */
#define PLF_ESTATE (-3)       /* call order: mesh / materials / K not set up yet                  */
#define PLF_EUNSUPPORTED (-4) /* feature the reference raises for (e.g. Tresca gradient)          */
#define PLF_ENOCONV (-5)      /* linear solver did not reach the requested tolerance              */

#define PLF_MAX_MATERIALS 16
#define PLF_MAX_GEOM 16

/* Material.response thresholds, named as in the reference */
#define PLF_YF_TOLERANCE 5.e-3 /* basic.py:26 yf_tolerance                     */
#define PLF_SPLIT_FY0 (-0.15)  /* material.py:260 elastic/plastic split         */
#define PLF_MAXIT 50           /* material.py:207 maxit (sub-steps)             */
#define PLF_DSTIFF_TOL 1.e-3   /* model.py:1348 Frobenius change of the tangent */

typedef enum plf_mat_kind {
    PLF_MAT_ELASTIC = 0, /* Material.sy is None                                                     */
    PLF_MAT_HILL = 1,    /* sdim=6 Hill-6p equivalent stress; J2 is hill = 1 (material.py:653-661)  */
    PLF_MAT_SVC = 2,     /* ML_yf: RBF support-vector yield function, sdim=6 (material.py:398-405)  */
    PLF_MAT_BARLAT = 3,  /* Yld2004-18p equivalent stress (material.py:678-702)                     */
    PLF_MAT_HILL3 = 4,   /* sdim=3: 3-parameter Hill / J2 on ordered principal stresses (663-673)   */
    PLF_MAT_SVC3 = 5     /* ML_yf, sdim=3: SVC on (seq_J2/scale-1, polar angle/pi) of the principal stresses
                            (material.py:2331-2334, gradient 780-805); svc_sv rows: 6 columns, 2..5 zero */
} plf_mat_kind;

/* What Material.elasticity / Material.plasticity / train_SVC leave on a Material object
 * (material.py:2401-2594, 1109-1278), flattened. */
typedef struct plf_material {
    int32_t kind;     /* plf_mat_kind */
    int32_t nsv;      /* SVC: number of support vectors */
    int32_t dev_only; /* SVC: features use the deviatoric stress (material.py:2337-2338) */
    int32_t tresca;   /* Tresca equivalent stress max - min principal (material.py:630-632; no gradient, 824-825) and the
                         ML_full_yf start value (470-473) */
    int32_t lhs;      /* d[] holds LHS parameters (material.py:646-647, 829-830) instead of drucker*[1,1,1] */
    int32_t pad_;
    double E, nu;     /* isotropic constants; used for the plane-stress CV and B row (model.py:274-283, 498-501) */
    double C44;       /* Material.C44 (plane-stress CV[5][5]) */
    double CV[36];    /* Material.CV, the 3-d Voigt stiffness */
    double sy, khard;
    double hill[6];
    double d[3];
    double barlat_m1[36], barlat_m2[36], barlat_exp; /* Material.Bar_m1 / Bar_m2 / barlat_exp */
    double svc_gamma, svc_rho, svc_scale_seq;        /* gam_yf, svm_yf.intercept_[0], scale_seq */
    const double *svc_sv;   /* host, [nsv][svc_nfeat]: svm_yf.support_vectors_ */
    const double *svc_dual; /* host, [nsv]:    svm_yf.dual_coef_[0]    */
    /* work-hardening SVC (Material.whdat, material.py:2342-2346, 2667-2669): svc_nfeat = 15 = 6 stress + 6 plastic strain /
     * scale_wh + accumulated strain + maximum stress + flag (the last three are 0 on this path, the defaults of calc_yf);
     * 0 or 6: stress features only.  khard then is what calc_fgrad leaves on the material (808-814), per element. */
    int32_t svc_nfeat;
    int32_t pad2_;
    double svc_scale_wh;
} plf_material;

typedef struct plf_ctx plf_ctx;

/* ---- context --------------------------------------------------------------------------------------------- */
int plf_create(int device, plf_ctx **out);
void plf_destroy(plf_ctx *ctx);
const char *plf_last_error(const plf_ctx *ctx); /* ctx may be NULL: message of the last failed plf_create */
const char *plf_version(void);
size_t plf_sizeof_material(void); /* sizeof(plf_material), for binding self-checks */
int plf_device_info(const plf_ctx *ctx, char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor,
                    int64_t *mem_bytes);
/* CUDA-event timer on the context's stream (the stream every kernel of this library is launched on). */
int plf_timer_start(plf_ctx *ctx);
int plf_timer_stop(plf_ctx *ctx, float *ms); /* records, synchronises, returns elapsed ms */
/* event pool for phase timing inside a step: record event `idx` (0..4095) on the stream; elapsed ms between two */
int plf_event_record(plf_ctx *ctx, int idx);
int plf_event_elapsed(plf_ctx *ctx, int idx0, int idx1, float *ms); /* synchronises on idx1 */
int plf_synchronize(plf_ctx *ctx);
int plf_flush_l2(plf_ctx *ctx);              /* overwrites a 256 MiB scratch buffer (L2 is 126 MB) */
int64_t plf_launch_count(const plf_ctx *ctx); /* kernels launched by this context so far */
/* Names of the return-mapping / assembly kernel variants launched since the previous call, ';'-separated, spelled as ncu
 * prints them ("response_passA_kernel<1, 4, 0>;..."); the first call only switches the recording on.  bench.py matches its
 * roofline.traffic figures to exactly these names. */
int plf_last_kernels(plf_ctx *ctx, char *buf, int len);
/* page-locked host memory for callers that want asynchronous H2D/D2H copies of their numpy buffers */
int plf_pinned_alloc(size_t bytes, void **out);
int plf_pinned_free(void *p);
/* tuning / A-B switches; unknown keys return PLF_EINVAL:
 *   "compact"      1 (default): two-pass return mapping, the 50-sub-step elements are compacted into a dense second
 *                  launch instead of diverging inside the first one; 0: one unsplit kernel
 *   "max_fm"       cap of the fast mode of the Hill / J2 materials of the set (hill_fast.cuh; default 4, 0 = general Hill kernels;
 *                  batch kernels and plf_sweep alike; elastic materials in the set do not switch it off)
 *   "passb_recur"  1 (default): recurrence form of the planar isotropic sub-step loop of pass B; 0: direct form
 *   "passa_direct" 1: pass A writes its records with per-thread 256-bit stores; 0 (default): staged coalesced stores
 *   "assembly_tiled" 1 (default): staged coalesced CSR stores; 0: direct stores
 *   "solver_mg"    1 (default): multigrid V-cycle preconditioner for plf_solve_linear; 0: Jacobi
 *   "solver_warm"  1 (default): CG starts from the previous du; 0: from zero
 *   "mg_nu", "mg_coarse_sweeps", "mg_omega_permille": smoothing sweeps per level / on the coarsest level / Jacobi damping
 *   "k1_only"      diagnostics: 1 = launch pass A only, 2 = pass A without the constitutive math
 *   "passa_bulk"   1: pass A with cp.async.bulk / mbarrier (TMA engine) instead of the cp.async ring (default 0: measured slower)
 *   "mg_fuse_tail" 2 (default): coarse levels of the V-cycle in one cooperative kernel + one fused block; 1: fused block only; 0: per level
 *   "skip_unchanged" 1 (default): an unchanged K is neither re-assembled nor re-coarsened
 *   "solver_hybrid" replicated multi-GPU solve: the CG vectors and the finest levels of the V-cycle (level 0, and every further
 *                   level of >= 1 M dofs) split over the ranks, one ghost column swapped per stencil step: 0 never, 1 always,
 *                   2 (default) when the exchanges run through peer mailboxes (plf_comm_p2p) */
int plf_set_option(plf_ctx *ctx, const char *key, int value);

/* ---- materials ------------------------------------------------------------------------------------------- */
/* Model.assign(mats) (model.py:554-577).  planestress selects Element.CV (model.py:270-304). */
int plf_set_materials(plf_ctx *ctx, int nmat, const plf_material *mats, int planestress);
/* Element.CV of material `id` as the kernels use it (6x6 row-major) */
int plf_get_element_CV(const plf_ctx *ctx, int id, double *CV);

/* ---- stateless batched constitutive calls on HOST arrays (H2D + kernel + D2H inside the call) ------------- */
/* Material.response (material.py:207-346) for n independent elements.
 * mat_id[n] may be NULL (all material 0).  Outputs: fy[n], sig_out[n][6], depl[n][6], ctan[n][36], nsteps[n]
 * (= Material.msg['nsteps'], material.py:345).  Any output pointer may be NULL. */
int plf_response_batch(plf_ctx *ctx, int64_t n, const int32_t *mat_id, const double *sig, const double *epl,
                       const double *deps, double *fy, double *sig_out, double *depl, double *ctan, int32_t *nsteps);
/* Material.calc_seq (576-676), calc_yf (348-412), calc_fgrad (704-858), ML_full_yf (414-516),
 * epl_dot (1009-1055), C_tan (1057-1086) on (n,6) arrays, one material. */
int plf_calc_seq_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, double *seq);
int plf_calc_yf_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *yf);
int plf_calc_fgrad_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *fgrad);
int plf_ml_full_yf_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, const double *ld,
                         double *yf);
int plf_epl_dot_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, const double *deps,
                      double *pdot);
int plf_ctan_batch(plf_ctx *ctx, int mat, int64_t n, const double *sig, const double *epl, double *ctan);

/* ---- device-resident batch (bench / pipelines that keep state in HBM) -------------------------------------- */
typedef struct plf_batch plf_batch; /* n elements: mat_id, sig, epl, deps in; fy, sig_out, depl, ctan, nsteps out */
int plf_batch_create(plf_ctx *ctx, int64_t n, plf_batch **out);
void plf_batch_destroy(plf_ctx *ctx, plf_batch *b);
int plf_batch_upload(plf_ctx *ctx, plf_batch *b, const int32_t *mat_id, const double *sig, const double *epl,
                     const double *deps);
int plf_batch_run(plf_ctx *ctx, plf_batch *b); /* asynchronous on the context stream */
int plf_batch_download(plf_ctx *ctx, plf_batch *b, double *fy, double *sig_out, double *depl, double *ctan,
                       int32_t *nsteps);

/* ---- mesh, element stiffness, global stiffness -------------------------------------------------------------- */
/* Model.mesh (model.py:758-952) for the structured Q4 mesh: NX x NY elements, element id = j*NY + k, node id =
 * j*(NY+1) + k, element nodes [n1, n1+1, n1+NY+1, n1+NY+2] (932-949).  dx[NX]: element width per column
 * (laminate sections may differ, 839), dy: element height, thick: Model.thick.  elem_mat[NX*NY]: material index
 * per element (0-based).  Resets the element state (sig = eps = epl = 0, elstiff = CV) like the first
 * Model.solve does (1214-1227). */
int plf_mesh_structured(plf_ctx *ctx, int NX, int NY, const double *dx, double dy, double thick,
                        const int32_t *elem_mat);
int plf_mesh_sizes(const plf_ctx *ctx, int64_t *Nel, int64_t *Nnode, int64_t *Ndof, int64_t *nnz);
/* Element.Bmat (model.py:439-512) of element e: 4 x (6x8) row-major */
int plf_get_bmat(const plf_ctx *ctx, int64_t e, double *B4);
/* Element.elstiff for all elements, (Nel,36).  set: NULL resets to Element.CV (model.py:1223). */
int plf_set_tangent(plf_ctx *ctx, const double *elstiff);
int plf_get_tangent(plf_ctx *ctx, double *elstiff);
/* Element.calc_Kel (model.py:365-370) for all elements -> host (Nel,64) */
int plf_get_kel(plf_ctx *ctx, double *Kel);
/* Model.setupK (model.py:954-977): integrate B^T C B per element and assemble the global stiffness as CSR on the
 * device.  Per entry the element contributions are summed in ascending element order (bit-exact with the
 * reference's dense K). */
int plf_setupK(plf_ctx *ctx);
int plf_get_csr(plf_ctx *ctx, int32_t *rowptr, int32_t *col, double *val); /* sizes from plf_mesh_sizes */
/* Same assembly, element tangents taken from the `ctan` output of a device-resident batch of Nel elements (the
 * grad_stiff that Model.solve would store into Element.elstiff at model.py:1351) without a copy in between. */
int plf_setupK_from_batch(plf_ctx *ctx, const plf_batch *b);
/* One inner pass of Model.solve's non-linear loop with HOST buffers (model.py:1338-1361 followed by setupK, 954-977):
 * (Nel,6) element states in; Material.response results (fy, sig, depl, grad_stiff, nsteps) and the nnz CSR values of
 * the K assembled from the new tangents out.  Upload, update + assembly and download are pipelined over chunks of
 * element columns on three streams (full-duplex PCIe); pass pinned buffers (plf_pinned_alloc).  Output pointers may
 * be NULL.  `b` must hold Nel elements (material ids are taken from the mesh). */
int plf_step_host(plf_ctx *ctx, plf_batch *b, const double *sig, const double *epl, const double *deps, double *fy,
                  double *sig_out, double *depl, double *ctan, int32_t *nsteps, double *kval);

/* plf_step_host with compact results.  grad_stiff is symmetric and equals Element.CV bit for bit on the elastic branch
 * (material.py:253-256), so: elastic_words[(Nel+31)/32] has bit (e % 32) of word e / 32 set for elements that took the
 * elastic branch (their grad_stiff is CV, their depl is 0), and ct21 receives, in element order, the 21 upper-triangle
 * entries (row-major, i <= j) of grad_stiff of all OTHER elements; *n_packed is their number (ct21 must hold Nel * 21
 * doubles).  kval == NULL leaves the assembled K on the device for plf_solve_linear (the Newton loop of Model.solve never
 * needs it on the host); fy, sig_out, depl, nsteps may be NULL.  Mixed-material meshes are supported (ids from the mesh). */
int plf_step_host_packed(plf_ctx *ctx, plf_batch *b, const double *sig, const double *epl, const double *deps, double *fy,
                         double *sig_out, double *depl, int32_t *nsteps, double *ct21, uint32_t *elastic_words,
                         int64_t *n_packed, double *kval);

/* ---- element state --------------------------------------------------------------------------------------------- */
typedef enum plf_field {
    PLF_SIG = 0, PLF_EPS = 1, PLF_EPL = 2,         /* Element.sig / eps / epl          (Nel,6) */
    PLF_RES_SIG = 3, PLF_RES_DEPL = 4,             /* Element.res_sig / res_depl       (Nel,6) */
    PLF_U = 5, PLF_DU = 6, PLF_F = 7,              /* Model.u / du / f                 (Ndof)  */
    PLF_FYLD = 8,                                  /* f list of model.py:1345          (Nel)   */
    PLF_NSTEPS = 9                                 /* msg['nsteps'] per element, int32 (Nel)   */
} plf_field;
int plf_get_field(plf_ctx *ctx, int field, void *host);
int plf_set_field(plf_ctx *ctx, int field, const void *host);

/* ---- one pass of the non-linear loop (model.py:1338-1361) on the device ---------------------------------------- */
/* For every plastic element: deps = sum_g B_g du_e (387-398); response(sig, epl, deps, CV) (1342-1343);
 * f = fyld / get_sflow(eps_eq(epl)) (1345); hh = |elstiff - grad_stiff|_F; if hh > 1e-3 replace the element tangent
 * (average instead when nit >= 15) (1346-1355).  Outputs: change (any tangent replaced), conv (all f <= 5.0005e-3),
 * max_hh, and counts of elements per branch {elastic, plastic 1 step, plastic 50 sub-steps, radial corrections}. */
int plf_sweep(plf_ctx *ctx, int nit, int32_t *change, int32_t *conv, double *max_hh, int64_t counts[4]);
/* state commit (model.py:1383-1392): u += du; f += K du; epl += res_depl; sig = res_sig (or += dsig / depl for
 * elements that never went through response); eps = B u. */
int plf_commit(plf_ctx *ctx);
/* calc_global element part (model.py:1500-1511): volume-weighted means of sig, eps, epl -> 3 x 6 */
int plf_reduce_global(plf_ctx *ctx, double *sig6, double *eps6, double *epl6);
/* calc_scf (model.py:1036-1067) -> scaling factor of the load increment; sld = loading direction (1245-1258) */
int plf_calc_scf(plf_ctx *ctx, const double *sld, double *scf);

/* ---- linear solve (model.py:1028-1033, 1070-1206, 1291, 1335) -------------------------------------------------- */
/* Solve K[ind][:,ind] du[ind] = df[ind] on the device with preconditioned CG (geometric multigrid V-cycle; Jacobi for
 * meshes that cannot be coarsened).
 *   fixed[Ndof]  : 1 where a displacement BC removed the dof from `ind`
 *   du_bc[Ndof]  : prescribed du on fixed dofs (0 elsewhere)                      -- calc_BC's du
 *   w[Ndof]      : sum of all BC values applied to a dof (what calc_BC multiplies K's columns with)
 *   f_ext[Ndof]  : force-type BC contributions (1145-1151, 1173-1179, 1203-1205)
 * df = f_ext - K w on the free dofs.  Result: Model.du on the device (PLF_DU).  rtol is on |r|/|b|.
 * Passing NULL for all four vectors re-uses the ones of the previous call (same boundary conditions, new K). */
int plf_solve_linear(plf_ctx *ctx, const uint8_t *fixed, const double *du_bc, const double *w, const double *f_ext,
                     double rtol, int max_iter, int *iters, double *relres);

/* ---- load-increment controller (model.py:1208-1450) ------------------------------------------------------------- */
/* Model.solve's loop over load increments on the library side of the boundary, so that an increment -- elastic
 * predictor, calc_scf, the Newton-like passes (setupK, reduced solve, response sweep with tangent replacement),
 * the commit and calc_global -- runs without per-pass host round trips through the binding.  The boundary-condition
 * vectors of calc_BC (model.py:1070-1206) are built here from the plain description below (O(boundary) host work,
 * scattered into device vectors).  Behaviour restated, not re-designed: same increment sizes, same halving rule for
 * il < 6 and nit > 1 (1308-1330), same termination test (1395-1411), same nsteps / niter / co_nconv. */
typedef struct plf_bc {
    double left[2], bot[2];                 /* Model.bcl, bcb: static values                                   */
    double right[2], top[2], node[2];       /* Model.bcr, bct, bcn: targets of the incremented conditions      */
    int32_t left_disp[2], bot_disp[2], right_disp[2], top_disp[2], node_disp[2]; /* ubc*: 1 displacement, 0 force */
    int32_t n_noset;                        /* Model.noset (bcnode): number of nodes, 0 = none                 */
    int32_t pad_;
    const int32_t *noset;                   /* node ids j*(NY+1)+k of the GLOBAL mesh                          */
} plf_bc;

typedef struct plf_load_state {
    double reached_right[2], reached_top[2], reached_node[2];   /* bcr0, bct0, bcn0 (in/out: Model.bcr_mem ...) */
    int32_t il;       /* increments completed in this Model.solve call                                          */
    int32_t nconv;    /* passes without plastic convergence so far (co_nconv)                                   */
    int32_t more;     /* out: 1 while a target is not reached (bc_inc, 1411)                                    */
    int32_t pad_;
} plf_load_state;

#define PLF_MAX_PASSES 17   /* the inner loop runs while nit <= 15 (1305): at most 16 passes */
typedef struct plf_increment_report {
    int32_t passes;                 /* nit when the inner loop ended (niter gets passes - 1; linear models: 0)   */
    int32_t solves;                 /* linear solves of this increment                                           */
    int64_t cg_iterations;
    double max_relres;              /* largest relative residual any of these solves stopped at                  */
    double scale_bc;                /* calc_scf result (1296)                                                    */
    double step_right[2], step_top[2], step_node[2];   /* increment finally applied: dbcr, dbct, dbcn           */
    double glob_sig[6], glob_eps[6], glob_epl[6];      /* calc_global element part (1500-1511)                  */
    double edge[4][4];              /* bcval of left, right, bottom, top nodes: mean ux, mean uy, sum fx, sum fy */
    /* per pass (verbose output, statistics) */
    int32_t pass_change[PLF_MAX_PASSES], pass_conv[PLF_MAX_PASSES];
    double pass_max_hh[PLF_MAX_PASSES];
    int64_t pass_counts[PLF_MAX_PASSES][4];   /* elastic, 1 step, 50 sub-steps, radial corrections              */
    int32_t pass_cg_iters[PLF_MAX_PASSES + 1];/* [0] predictor solve, [1 + nit] the solve of pass nit          */
    int32_t n_warnings;             /* inconsistent displacement BCs met while building the vectors             */
    char warning[200];              /* text of the first one (the reference warns, model.py:1119, 1137, 1167)   */
} plf_increment_report;

/* First part of Model.solve (1214-1258): `reset` != 0 zeroes u, f and the element state and sets elstiff = CV (first
 * call after meshing); assembles K (1242); stores the boundary description and the loading direction sld. */
int plf_solve_begin(plf_ctx *ctx, const plf_bc *bc, int reset);
/* One pass of the `while bc_inc` loop (1270-1418).  min_step <= 0: None.  rtol / max_iter: linear solver. */
int plf_load_increment(plf_ctx *ctx, plf_load_state *st, int min_step, double rtol, int max_iter,
                       plf_increment_report *rep);
/* Model.bcval (1452-1471) for the four edges from the device-resident u and f: edge[left,right,bottom,top][mean ux,
 * mean uy, sum fx, sum fy]. */
int plf_boundary_values(plf_ctx *ctx, double edge[4][4]);
size_t plf_sizeof_increment_report(void);

/* ---- strip-partitioned solve over several GPUs (one process per GPU) -------------------------------------------- */
/* The reference has no counterpart (it is a single-process NumPy code); SURVEY.md 8e: rank g holds the strip of
 * element columns it owns plus one halo element column on its left (pylabfea_b200/shard.py), so the rows of K of its
 * owned node columns are complete and bit-identical to the global K (model.py:954-977) without communication.  The
 * solve is the one exchange step: plf_solve_linear on a context with a communicator runs Jacobi-preconditioned CG on
 * the distributed system -- per iteration one halo node column (2 (NY+1) doubles) to and from each neighbour and two
 * all-reduces of the dot products, all enqueued on the context's stream (NCCL, loaded with dlopen on first use).
 * Every rank passes the boundary-condition vectors of ITS local node grid (slices of the global calc_BC vectors,
 * ghost columns included) and gets du on its local node grid back (ghost columns filled from the neighbours). */
#define PLF_COMM_ID_BYTES 128
int plf_comm_unique_id(unsigned char *id);                 /* rank 0: ncclGetUniqueId; hand the bytes to every rank */
int plf_comm_init(plf_ctx *ctx, const unsigned char *id, int rank, int world);   /* collective over all ranks */
/* owned node columns [own_first, own_end) of the local mesh: own_first = 1 on ranks with a left neighbour (column 0 is
 * the neighbour's), own_end = NX on ranks with a right neighbour (column NX is the neighbour's), else 0 / NX + 1 */
int plf_comm_set_strip(plf_ctx *ctx, int own_first, int own_end);
/* plf_comm_set_strip plus the position of the strip in the GLOBAL mesh: first_col_global = global index of the local
 * element column 0 (halo column included), NX_global element columns in all, x_offset = x coordinate of the local node
 * column 0, lenx_global = Model.lenx.  With this a context can run plf_solve_begin / plf_load_increment as one rank of a
 * distributed Model.solve: boundary conditions are given for the global model (node sets by global node id), every rank
 * builds the calc_BC vectors of its node grid, and sweep flags (model.py:1355-1361), calc_scf moments (1056-1062), the
 * calc_global sums (1500-1511) and the bcval sums (1452-1471) are combined over the ranks, so all ranks step alike. */
int plf_comm_set_partition(plf_ctx *ctx, int own_first, int own_end, int first_col_global, int NX_global, double x_offset,
                           double lenx_global);
/* Replicated solve.  The element work of a pass (stress updates, K rows) shards perfectly over the strips; the linear
 * solve does not have to: `global_ctx` is a second context of the same process holding the GLOBAL mesh (plf_set_materials +
 * plf_mesh_structured with the global NX).  With it, every solve of plf_load_increment gathers the owned CSR rows of all
 * ranks into global_ctx (one NCCL broadcast per rank; the rows of a strip are one contiguous range of the global value
 * array), solves the complete system there -- on every rank alike, with the single-GPU multigrid-preconditioned CG, so du
 * is bit-identical to the single-GPU solve -- and copies the strip's node columns of du back.  This is the north star's
 * "NCCL allreduce on the assembled system" variant: no communication inside the solve, the solve itself is not sharded.
 * Pass NULL to return to the strip-partitioned CG. */
int plf_comm_set_replica(plf_ctx *ctx, plf_ctx *global_ctx);
/* 1 when the halo columns and all-reduce scalars of this context travel through peer mailboxes -- every rank maps one device
 * buffer of every other rank with CUDA IPC at plf_comm_set_strip, and an exchange is ONE kernel that stores into the peers'
 * HBM over NVLink, releases a flag and spins on its own -- 0 when they go through NCCL (no IPC between the ranks, or the
 * environment variable PLF_P2P=0).  The results are the same bits either way (all-reduces sum in rank order). */
int plf_comm_p2p(plf_ctx *ctx);
/* sum of n <= 64 host doubles over all ranks, in place (sweep flags, calc_global sums, calc_scf moments) */
int plf_comm_allreduce(plf_ctx *ctx, double *vals, int n);
int plf_comm_destroy(plf_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PLFB200_H */
