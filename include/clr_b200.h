/*
 * clr_b200.h -- C ABI of libclr_b200.so: the B200 (sm_100a) hot path of ClosedLoopReachability.jl.
 *
 * This is the drop-in boundary.  The Julia host files (src/solve.jl, src/splitters.jl,
 * src/simulate.jl of the reference) `ccall` these entry points; the Python mirror in
 * closedloopreachability.jl_b200/ binds the same symbols with ctypes (INTEGRATION.md shows both).
 *
 * Conventions
 *  - every matrix is Float64 column-major (Julia native); a batch of vectors is stored
 *    "dim x N" column-major, i.e. unit (cell / trajectory) j owns the contiguous slice [j*dim, (j+1)*dim);
 *  - the caller owns every buffer it passes; the library owns all device memory inside clr_ctx;
 *  - every function returns 0 on success and a negative clr_status on error; the message is
 *    available from clr_last_error(); no C++ exception crosses the boundary;
 *  - a clr_ctx is bound to ONE CUDA device and is not thread-safe (one ctx per host task);
 *    calls block the calling thread (with CLR_FLAG_DEVICE_IO no bulk data is copied, see the flag);
 *  - there is no CPU fallback: without a CUDA device clr_create fails with CLR_ERR_CUDA.
 *
 * Reference interface replaced by each entry point (paths relative to the reference repo):
 *   clr_net_upload            FeedforwardNetwork / DenseLayerOp as used by src/solve.jl:212, src/simulate.jl:104
 *   clr_deepz_boxgrid         the cell loops of _solve, src/solve.jl:153-166 and :187-195, over
 *                             apply(BoxSplitter, X0) (src/splitters.jl:20-28), each cell going through
 *                             controller_forward (src/solve.jl:210-218) with algorithm_controller = DeepZ()
 *   clr_deepz_zonotopes       same loops for arbitrary zonotope cells (k > 1, ZonotopeSplitter cells,
 *                             models/VerticalCAS/VerticalCAS.jl:122 direct forward(Y, net, DeepZ()) calls)
 *   clr_deepz_fetch_zonotopes the output set U of controller_forward (src/solve.jl:213) when the caller
 *                             needs more than its box (TaylorModelReconstructor(box=false), src/setops.jl:84-107)
 *   clr_deepz_fetch_reduced   _reduce_order(Z0, 2; force_reduction=true) of the same reconstructor, src/setops.jl:84-86
 *   clr_project_oa            _project_oa(R::AbstractTaylorModelReachSet, vars, t), src/setops.jl:9-12, for all chains of a
 *                             control period at once (src/solve.jl:181-184)
 *   clr_forward_points        forward(x, network) + pre/post-processing, src/simulate.jl:101-106
 *   clr_rollout               the period loop of simulate (src/simulate.jl:98-140) incl. _solve_ensemble
 *                             (ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:28-45: Tsit5 ensemble)
 */
#ifndef CLR_B200_H
#define CLR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLR_MAX_LAYERS 32

typedef enum {
    CLR_OK = 0,
    CLR_ERR_ARG = -1,        /* invalid argument (Julia side: ArgumentError) */
    CLR_ERR_CAPACITY = -2,   /* keep_cap / log_cap too small for a result */
    CLR_ERR_DIM = -3,        /* dimension mismatch (Julia side: ArgumentError, src/solve.jl:130-134) */
    CLR_ERR_CUDA = -4,       /* CUDA runtime error or no device */
    CLR_ERR_ALLOC = -5,
    CLR_ERR_STATE = -6,      /* call order (e.g. fetch before a keep run) */
    CLR_ERR_UNSTABLE = -7    /* a rollout integration gave up: OrdinaryDiffEq's retcodes Unstable / DtLessThanMin / MaxIters */
} clr_status;

/* activations (ControllerFormats names) */
enum { CLR_ACT_ID = 0, CLR_ACT_RELU = 1, CLR_ACT_SIGMOID = 2, CLR_ACT_TANH = 3 };
/* post-processing kinds (src/problem.jl:5-58) */
enum { CLR_POST_NONE = 0, CLR_POST_SHIFT = 1, CLR_POST_SCALE = 2, CLR_POST_MATRIX = 3, CLR_POST_PROJECT = 4 };
/* built-in plant right-hand sides on the extended state [x; w; u] */
enum { CLR_PLANT_UNICYCLE = 1, CLR_PLANT_TORA = 2, CLR_PLANT_ACC = 3, CLR_PLANT_INVPEND = 4 };
/* ODE algorithms */
enum { CLR_ALG_TSIT5 = 0, CLR_ALG_RK4 = 1 };
/* flags */
enum {
    CLR_FLAG_DEVICE_IO = 1,   /* the bulk data pointers (centers, radii, C, col_off, G, X, x0, Wd and every output
                                 array incl. the hull) are DEVICE pointers: no host<->device copies are made and the
                                 results are complete in stream order on clr_stream(ctx).  partition, clr_prepost and
                                 clr_stats stay host pointers.  The DeepZ calls still synchronise once internally
                                 (they read back how many cells need the large-cell pass). */
    CLR_FLAG_NO_STATS = 2,    /* skip the per-layer statistics (stats may then be NULL) */
    CLR_FLAG_MERGE_REDUNDANT = 4  /* clr_project_oa: LazySets' remove_redundant_generators on the full zonotope before the projection --
                                 the default of overapproximate(::Vector{TaylorModel1}, Zonotope), i.e. what src/setops.jl:10 runs */
};

typedef struct clr_ctx clr_ctx;
typedef struct clr_net clr_net;
typedef struct clr_plant clr_plant;

/* Pre/post-processing of controller_forward (src/solve.jl:211, :213).  Always HOST pointers. */
typedef struct {
    int32_t pre_rows;        /* 0 = identity; else x -> preA*x + preb with preA pre_rows x n (col-major) */
    const double* preA;
    const double* preb;
    int32_t post_kind;       /* CLR_POST_* */
    int32_t post_rows;       /* rows of postA (MATRIX) or length of post_dims (PROJECT) */
    const double* postA;     /* SCALE: postA[0]; MATRIX: post_rows x m col-major */
    const double* postb;     /* SHIFT: postb[0] */
    const int32_t* post_dims;/* PROJECT: 0-based output rows kept */
} clr_prepost;

/* Per-call statistics (same layout as the CPU oracle's). */
typedef struct {
    int64_t n_cells;
    int64_t sum_p_in[CLR_MAX_LAYERS];     /* generator columns entering network layer l, summed over cells
                                             (zero columns excluded, as after remove_zero_columns) */
    int64_t sum_unstable[CLR_MAX_LAYERS]; /* generators appended by the activation of layer l */
    int64_t near_threshold;               /* ReLU neurons with |l| or |u| < 1e-12 (branch could flip) */
    int64_t max_p_out;                    /* largest generator count of an output zonotope */
    double flops;                         /* sum_l 2 m_l n_l (p_l + 1) over all cells */
} clr_stats;

/* ---- context ------------------------------------------------------------------------------- */
int32_t clr_create(int32_t device, clr_ctx** out);
void clr_destroy(clr_ctx* ctx);
/* message of the last error on this ctx (or of the last failed clr_create when ctx == NULL) */
const char* clr_last_error(const clr_ctx* ctx);
/* the cudaStream_t all work of this ctx is enqueued on */
void* clr_stream(clr_ctx* ctx);
int32_t clr_synchronize(clr_ctx* ctx);
/* number of kernel launches issued by this ctx so far */
int64_t clr_launch_count(const clr_ctx* ctx);
/* device memory owned by the library: for callers that keep inputs/outputs resident (CLR_FLAG_DEVICE_IO) */
int32_t clr_device_alloc(clr_ctx* ctx, int64_t bytes, void** out);
int32_t clr_device_free(clr_ctx* ctx, void* p);
int32_t clr_memcpy_h2d(clr_ctx* ctx, void* dst, const void* src, int64_t bytes);
int32_t clr_memcpy_d2h(clr_ctx* ctx, void* dst, const void* src, int64_t bytes);
/* Host buffers.  Blocking calls accept any host memory; into PAGEABLE memory (a Julia Array, a NumPy array) the driver stages
   the copy (measured on B200: ~5-11 GB/s, and a chunked pinned staging pipeline inside the library was no better -- the first
   touch of fresh pages dominates).  Buffers that are page-locked -- cudaHostAlloc, or registered IN PLACE with clr_host_register
   (cudaHostRegister; once per buffer, ~0.1 ms per MB) -- are written by a single DMA at PCIe speed: register the arrays a loop of
   calls reuses.  clr_host_unregister before the memory is freed; both are idempotent. */
int32_t clr_host_register(clr_ctx* ctx, void* p, int64_t bytes);
int32_t clr_host_unregister(clr_ctx* ctx, void* p);
/* CLR_FLAG_DEVICE_IO calls return before their kernels have run: the device-side status word of the last rollout / order
   reduction / zonotope check (0, CLR_ERR_CAPACITY, CLR_ERR_UNSTABLE, CLR_ERR_ARG) after synchronising the stream */
int32_t clr_last_device_error(clr_ctx* ctx, int32_t* code);
/* tuning knob of the DeepZ tile scheduler: cells per tile (0 = automatic) */
int32_t clr_set_tile_cells(clr_ctx* ctx, int32_t cells);
/* Neuron reordering of large box-grid calls (default on): a pilot run over 2048 cells of the WHOLE split counts how often
   each hidden neuron stays non-zero; calls with >= 16384 cells and keep_cap == 0 then run an equivalent network whose
   inactive neurons come last, so that the layer products skip their all-zero rows.  Results agree with the unordered
   network to ~1e-16 relative (summation order); every shard of one split uses the same order. */
int32_t clr_set_reorder(clr_ctx* ctx, int32_t on);
/* Narrow networks (every layer, incl. folded pre/post-processing, at most 32 neurons wide -- ACC, VerticalCAS, InvertedPendulum:
   the reference's own CPU-runnable sizes) run a one-warp-per-cell kernel when neither statistics nor kept zonotopes are
   requested (default on; results agree with the tiled kernel to rounding, ~1e-16 relative, not bit for bit). */
int32_t clr_set_tiny(clr_ctx* ctx, int32_t on);
/* on: CUDA events bracket each pass of the following DeepZ calls (main, retry, big-cell) */
int32_t clr_set_profiling(clr_ctx* ctx, int32_t on);
/* device time (ms) of the three passes of the last DeepZ call (needs profiling on), the number of
   cells that were retried alone and the number that went through the global-store pass */
int32_t clr_last_pass_info(clr_ctx* ctx, double* pass_ms /*3*/, int64_t* retried, int64_t* big);
/* The roofline denominator of the layer products, measured on THIS device: a register-only loop of independent FP64
   mma.sync.m8n8k4 (SASS DMMA) chains on every SM for about `ms` milliseconds; *tflops = 512 FLOP per warp instruction /
   time (CUDA events).  MEASURED_PEAKS.json carries no FP64 figure, so bench.py measures it in the same run. */
int32_t clr_measure_fp64_peak(clr_ctx* ctx, double ms, double* tflops);

/* ---- networks ------------------------------------------------------------------------------ */
/* dims has n_layers+1 entries; W[l] is dims[l+1] x dims[l] column-major, b[l] has dims[l+1] entries.
   A clr_net belongs to the context it was uploaded to and is used by that context's host task only: the calls that take a
   `const clr_net*` never change the network, but they build and cache its device images inside it (one per pre / post-processing
   and input dimension, plus the pilot-ordered image of large box-grid calls) -- logically const, not safe to share between threads. */
int32_t clr_net_upload(clr_ctx* ctx, int32_t n_layers, const int32_t* dims, const int32_t* act,
                       const double* const* W, const double* const* b, clr_net** out);
void clr_net_free(clr_net* net);

/* ---- DeepZ over the cells of a box split --------------------------------------------------- */
/*
 * Cells are the elements of split(H, partition) in Iterators.product order (dimension 1 fastest).
 * centers / radii are the concatenated per-dimension tables (sum(partition) entries each): cell with
 * block index k_d in dimension d has centre centers[base_d + k_d] and radius radii[base_d + k_d]
 * (uniform splits repeat one radius per dimension; cut-point splits do not).  A zero radius drops
 * that dimension's generator (convert(Zonotope, H)).  Only cells [cell_begin, cell_begin+cell_count)
 * are processed: this is the shard of one GPU.
 *
 * out_lo / out_hi: m_out x cell_count, the box of U = controller_forward(cell)  (src/solve.jl:214-216,
 * src/setops.jl:76-83).  hull_lo / hull_hi (m_out, may be NULL): min / max over the processed cells.
 * keep_cap > 0 additionally keeps every output zonotope (at most keep_cap generators) on the device
 * for clr_deepz_fetch_zonotopes.
 */
int32_t clr_deepz_boxgrid(clr_ctx* ctx, const clr_net* net, int32_t n, const int32_t* partition,
                          const double* centers, const double* radii, int64_t cell_begin,
                          int64_t cell_count, const clr_prepost* pp, double* out_lo, double* out_hi,
                          double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap,
                          int32_t flags);
/* Block-cyclic ids (block > 0): local cell i of the call is cell  cell_begin + (i / block) * stride + i % block  of the split;
   results are in local order.  Device r of G takes cell_begin = first + r * block and stride = G * block: its share is then
   spread evenly over the whole range, so that shards take the same time without a cost model (sharding is this library's own:
   the reference's loop is serial, src/solve.jl:153-166).  block = 0: contiguous, = clr_deepz_boxgrid. */
int32_t clr_deepz_boxgrid_cyclic(clr_ctx* ctx, const clr_net* net, int32_t n, const int32_t* partition,
                                 const double* centers, const double* radii, int64_t cell_begin, int64_t cell_count,
                                 int64_t block, int64_t stride, const clr_prepost* pp, double* out_lo, double* out_hi,
                                 double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap, int32_t flags);

/* ---- DeepZ over explicit zonotopes ----------------------------------------------------------- */
/* C: n x N centres; generators of cell i are columns col_off[i] .. col_off[i+1]-1 of G (n x col_off[N]) */
int32_t clr_deepz_zonotopes(clr_ctx* ctx, const clr_net* net, int32_t n, int64_t N, const double* C,
                            const int64_t* col_off, const double* G, const clr_prepost* pp,
                            double* out_lo, double* out_hi, double* hull_lo, double* hull_hi,
                            clr_stats* stats, int32_t keep_cap, int32_t flags);

/* Same for zonotopes in the padded layout that clr_project_oa and clr_deepz_fetch_zonotopes produce: generators of cell i are
   the first ncols[i] columns of the n x cap block at G + i*cap*n.  With CLR_FLAG_DEVICE_IO the output arrays of
   clr_project_oa feed this call directly (the cells of control period k + 1 never visit the host, src/solve.jl:181-195). */
int32_t clr_deepz_zonotopes_padded(clr_ctx* ctx, const clr_net* net, int32_t n, int64_t N, const double* C, const double* G,
                                   const int32_t* ncols, int32_t cap, const clr_prepost* pp, double* out_lo, double* out_hi,
                                   double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap, int32_t flags);

/*
 * Output zonotopes of the last keep_cap > 0 run: out_ncols (N), out_c (m_out x N),
 * out_G (m_out x keep_cap x N: generator j of cell i at out_G + (i*keep_cap + j)*m_out).
 * Zero generators are removed (remove_zero_columns) and the reference's column order is kept:
 * propagated input generators first, then the generators appended by each activation layer in
 * ascending neuron order.  HOST pointers.
 */
int32_t clr_deepz_fetch_zonotopes(clr_ctx* ctx, int32_t* out_ncols, double* out_c, double* out_G);

/*
 * Order reduction of the kept output zonotopes (Girard's method, LazySets reduce_order(Z, r, GIR05())), for all
 * cells at once: replaces the per-cell  _reduce_order(Z0, 2; force_reduction=true)  of
 * TaylorModelReconstructor(box=false) (src/setops.jl:84-86).  The generators are ordered by decreasing
 * |g|_1 - |g|_inf (stable), the first m (order - 1) are kept, the others are replaced by their interval hull:
 * out_G (m x (m order) x N) = [kept generators | diag(d)], the layout src/setops.jl:96-101 asserts for order 2;
 * out_c (m x N).  order = 1 is the box.  A zonotope with fewer than m (order - 1) generators is an error
 * (CLR_ERR_DIM; the reference throws there).  flags: CLR_FLAG_DEVICE_IO makes out_c / out_G device pointers.
 */
int32_t clr_deepz_fetch_reduced(clr_ctx* ctx, int32_t order, double* out_c, double* out_G, int32_t flags);

/*
 * Batched _project_oa of Taylor-model reach-sets (src/setops.jl:9-12), the step that produces the cells of a control
 * period k > 1 (src/solve.jl:181-184):  project(set(overapproximate(R, Zonotope, t)), vars)  for N reach-sets at once.
 * Every reach-set has n components; component v is a polynomial in time (order_t + 1 coefficients) whose coefficients
 * are polynomials in the n normalised space variables (n_monomials monomials, exponents[m*n + j] = exponent of variable
 * j in monomial m -- TaylorSeries' coefficient order, i.e. coeff_table, flattened over the degrees), plus an interval
 * remainder.  coeffs: n_monomials x (order_t+1) x n x N (monomial index fastest), rem: 2 x n x N (lo, hi),
 * dt: N evaluation times relative to the expansion point of the time polynomial, vars: n_keep 0-based components kept.
 * Output zonotopes: out_c (n_keep x N), out_G (n_keep x 2n x N: generator j of reach-set i at out_G + (i*2n + j)*n_keep,
 * columns [linear terms | remainder radii]), out_ncols (N).  With remove_zero_generators != 0 all-zero columns are
 * dropped (the keyword of the reference); parallel columns are NOT merged (LazySets' remove_redundant_generators):
 * the result is the same set.  The arrays are device pointers with CLR_FLAG_DEVICE_IO (exponents / vars stay host).
 */
int32_t clr_project_oa(clr_ctx* ctx, int32_t n, int32_t order_t, int32_t n_monomials, const int32_t* exponents, int64_t N,
                       const double* coeffs, const double* rem, const double* dt, const int32_t* vars, int32_t n_keep,
                       int32_t remove_zero_generators, double* out_c, double* out_G, int32_t* out_ncols, int32_t flags);

/* ---- pointwise controller -------------------------------------------------------------------- */
/* X: nx x N, U: m_out x N */
int32_t clr_forward_points(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, int32_t nx,
                           int64_t N, const double* X, double* U, int32_t flags);

/* ---- simulate(): batched closed-loop rollouts ------------------------------------------------- */
/*
 * x0: n_state x N.  Wd: n_dist x N x iters (fresh disturbance per period, src/simulate.jl:110-115) or NULL.
 * Period it integrates [t0 + it*tau, t0 + (it+1)*tau], the last one ends at T (src/simulate.jl:126).
 * alg = CLR_ALG_TSIT5: adaptive Tsit5 with OrdinaryDiffEq's default PI controller (abstol, reltol;
 * pow_mode 1 = DiffEqBase fastpow, 0 = exact pow); alg = CLR_ALG_RK4: `substeps` fixed steps per period.
 * X_end: n_state x N x iters, U: n_ctrl x N x iters; naccept / nreject (N x iters, may be NULL).
 * log_cap > 0 records the accepted steps of every period (piece.t / piece.u of the reference's
 * ODESolution pieces): log_n (N x iters), log_t (log_cap x N x iters), log_u (n_ext x log_cap x N x iters).
 */
int32_t clr_rollout(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, int32_t plant,
                    int32_t n_state, int32_t n_dist, int32_t n_ctrl, int64_t N, int32_t iters,
                    const double* x0, const double* Wd, double t0, double tau, double T, int32_t alg,
                    double abstol, double reltol, int32_t substeps, int32_t pow_mode, double* X_end,
                    double* U, int32_t* naccept, int32_t* nreject, int32_t log_cap, int32_t* log_n,
                    double* log_t, double* log_u, int32_t flags);

/* ---- run-time compiled plants ------------------------------------------------------------------- */
/*
 * The reference's plant is an arbitrary user function f!(dx, x, p, t) on the extended state
 * (RA.inplace_field!(ivp), ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:33-37).  clr_plant_compile takes the BODY of
 *     __device__ void rhs(const double* x, const double* q, double* dx)
 * in CUDA C (x: the n_state dynamic states, q: the constant tail [w (n_dist); u (n_ctrl)], dx: n_state derivatives),
 * compiles the rollout kernels around it with NVRTC for sm_100a and returns a handle for clr_rollout_plant, which
 * otherwise behaves exactly like clr_rollout.  ctx == NULL compiles only (syntax check, no device needed).
 */
int32_t clr_plant_compile(clr_ctx* ctx, int32_t n_state, int32_t n_dist, int32_t n_ctrl, const char* rhs_body,
                          clr_plant** out);
void clr_plant_free(clr_plant* plant);
int32_t clr_rollout_plant(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, const clr_plant* plant, int64_t N,
                          int32_t iters, const double* x0, const double* Wd, double t0, double tau, double T, int32_t alg,
                          double abstol, double reltol, int32_t substeps, int32_t pow_mode, double* X_end, double* U,
                          int32_t* naccept, int32_t* nreject, int32_t log_cap, int32_t* log_n, double* log_t,
                          double* log_u, int32_t flags);

/* ---- the other splitters on the device (reference src/splitters.jl:30-43, :60-71) ---------------------------------
 * clr_split_zonotopes: apply(ZonotopeSplitter(gens, nparts), Z) = LazySets split(Z, gens, nparts) for a batch of zonotopes in
 * the padded layout (C n x N, G n x cap x N, ncols N; gens 0-based): generator gens[i] is halved nparts[i] times, each
 * halving doubles the list ((Z - g/2, Z + g/2) in place of Z), so every input has 2^H children, H = sum(nparts), written
 * child-major behind one another in the same padded layout (ready for clr_deepz_zonotopes_padded).
 * clr_sign_split: apply(SignSplitter(), U) for N control intervals: an interval with 0 in its interior becomes [lo, 0] and
 * [0, hi]; out_lo / out_hi / out_parent need room for 2 N entries, *out_count receives the number written (order kept). */
int32_t clr_split_zonotopes(clr_ctx* ctx, int32_t n, int64_t N, const double* C, const double* G, const int32_t* ncols, int32_t cap,
                            int32_t n_gens, const int32_t* gens, const int32_t* nparts, double* out_C, double* out_G,
                            int32_t* out_ncols, int32_t flags);
int32_t clr_sign_split(clr_ctx* ctx, int64_t N, const double* lo, const double* hi, double* out_lo, double* out_hi,
                       int64_t* out_parent, int64_t* out_count, int32_t flags);

/* ---- several GPUs of one box behind one handle --------------------------------------------------
 * Replaces, for a caller with more than one device, the same reference loops as clr_deepz_boxgrid / clr_rollout
 * (src/solve.jl:153-166, :187-195; src/simulate.jl:98-140): the units (cells, trajectories) are independent, so device i
 * takes a contiguous range of the unit ids with the network replicated and NO exchange step; one host thread and one
 * stream per device for the duration of a call.  The only traffic between devices is the final gather of the per-device
 * output hulls (2 m doubles each): peer copies (NVLink) to the first device, one min / max kernel there, one D2H copy.
 * All data pointers are HOST pointers; results are in unit order exactly as from a single device (bit for bit when the
 * neuron reordering is off; with it on every shard still uses the order derived from the whole split).
 *   balance = 0: contiguous ranges of equal counts.  balance = 1: contiguous ranges whose boundaries are weighted by the
 *   algorithmic cost (FLOPs per cell) that a statistics pilot over short ranges of 16 blocks per device measures; cached per
 *   (network, split, range).  balance = 2: block-cyclic shards (device i takes the blocks i, i + G, ... of 1024 cells, see
 *   clr_deepz_boxgrid_cyclic): equal time without a cost model -- the default to use when cell costs vary over the split.
 *   The results are in cell order in every mode.  shard_bounds (n_devices + 1; ids for modes 0 / 1, cumulative shard sizes
 *   for mode 2) and shard_ms (device time per shard) may be NULL. */
typedef struct clr_multi clr_multi;
typedef struct clr_multi_net clr_multi_net;
int32_t clr_create_multi(const int32_t* devices, int32_t n_devices, clr_multi** out);
void clr_destroy_multi(clr_multi* mc);
int32_t clr_multi_size(const clr_multi* mc);
clr_ctx* clr_multi_ctx(clr_multi* mc, int32_t i);           /* the per-device context (tuning knobs, launch counts) */
const char* clr_multi_last_error(const clr_multi* mc);
int32_t clr_multi_net_upload(clr_multi* mc, int32_t n_layers, const int32_t* dims, const int32_t* act,
                             const double* const* W, const double* const* b, clr_multi_net** out);
void clr_multi_net_free(clr_multi_net* net);
int32_t clr_multi_deepz_boxgrid(clr_multi* mc, const clr_multi_net* net, int32_t n, const int32_t* partition,
                                const double* centers, const double* radii, int64_t cell_begin, int64_t cell_count,
                                const clr_prepost* pp, double* out_lo, double* out_hi, double* hull_lo, double* hull_hi,
                                clr_stats* stats, int32_t balance, int64_t* shard_bounds, double* shard_ms);
int32_t clr_multi_rollout(clr_multi* mc, const clr_multi_net* net, const clr_prepost* pp, int32_t plant, int32_t n_state,
                          int32_t n_dist, int32_t n_ctrl, int64_t N, int32_t iters, const double* x0, const double* Wd, double t0,
                          double tau, double T, int32_t alg, double abstol, double reltol, int32_t substeps, int32_t pow_mode,
                          double* X_end, double* U, int32_t* naccept, int32_t* nreject);

#ifdef __cplusplus
}
#endif
#endif /* CLR_B200_H */
