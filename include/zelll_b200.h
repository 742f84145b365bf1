/* zelll_b200.h -- C ABI of the B200-native cell-list engine (libzelll_b200.so).
 *
 * Drop-in boundary for ONE hot path of microscopic-image-analysis/zelll v0.5.0:
 *   CellGrid::new / rebuild / rebuild_mut  ->  particle_pairs / par_particle_pairs
 *   ->  distance filter + pair list / Lennard-Jones energy.
 *
 * The reference is pure Rust and has no FFI of its own; the entry points below are what a
 * `zelll-b200-sys` crate (Rust `extern "C"`), the PyO3 module and the ctypes front end in
 * zelll_b200/ bind (INTEGRATION.md shows the stubs).  Each entry cites the reference item
 * (file:line under the reference tree) it replaces.
 *
 * Conventions
 *   - every call returns a zb_status (0 = ok); nothing throws or aborts across the boundary
 *     (the reference panics instead: src/cellgrid.rs:227-229, src/cellgrid/util.rs:229-232);
 *   - `xyz` is a packed [n][ndim] array of f32 or f64 (the layout of `[T; N]` particles,
 *     src/lib.rs:236-244); particle labels are the enumerate() order 0..n-1
 *     (src/lib.rs:225-234, benches/lj.rs:71-78);
 *   - input/output pointers may be HOST or DEVICE memory (detected); device pointers make the
 *     call asynchronous on the handle's stream, host pointers synchronise it;
 *   - the handle owns all device memory and reuses it across rebuilds (the rebuild_mut
 *     contract, src/cellgrid.rs:251-252); one handle = one stream, not re-entrant; distinct
 *     handles are independent;
 *   - there is NO CPU fallback: without a CUDA device zb_grid_create fails with ZB_ERR_CUDA.
 */
#ifndef ZELLL_B200_H
#define ZELLL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZB_ABI_VERSION 1

typedef struct zb_grid zb_grid; /* opaque: CellGrid<(usize,[T;N]),N,T>, src/cellgrid.rs:112-126 */

enum zb_dtype { ZB_F32 = 0, ZB_F64 = 1 };

/* distance filter applied to the candidate pairs particle_pairs() yields:
 * NONE = unfiltered candidates (src/cellgrid.rs:338-340), LT = `dsq < c*c` (benches/lj.rs:85),
 * LE = `dsq <= c*c` (benches/cellgrid.rs:86, benches/iters.rs:74) */
enum zb_cmp { ZB_CMP_NONE = 0, ZB_CMP_LT = 1, ZB_CMP_LE = 2 };

enum zb_status {
  ZB_OK = 0,
  ZB_ERR_BAD_ARG = 1,
  ZB_ERR_CUDA = 2,          /* CUDA runtime error; see zb_last_error */
  ZB_ERR_CAPACITY = 3,      /* caller's output buffer too small; *n_out holds the needed size */
  ZB_ERR_TOO_MANY = 4,      /* n > i32::MAX (src/cellgrid/flatindex.rs:87) */
  ZB_ERR_GRID_TOO_LARGE = 5,/* bounding box / cutoff needs more cells than the engine can index (2^62; boxes
                               beyond 2^31 cells, or with far more cells than particles, are built as compact
                               sorted non-empty cells in O(n) memory; slab-sharded grids: 2^31) */
  ZB_ERR_NOT_BUILT = 6,
  ZB_ERR_OUT_OF_WINDOW = 7  /* sharded rebuild: a particle lies outside the imposed box/slab */
};

/* GridInfo + Aabb (src/cellgrid/util.rs:19-27, 81-90) plus sizes. */
typedef struct zb_info {
  double inf[3];      /* Aabb::inf  -- GridInfo::origin() (util.rs:139-141) */
  double sup[3];      /* Aabb::sup */
  double cutoff;      /* GridInfo::cutoff() (util.rs:179-181) */
  int32_t shape[3];   /* GridInfo::shape()   (util.rs:144-146) */
  int32_t strides[3]; /* GridInfo::strides() (util.rs:149-151): 1, shape0+4, (shape0+4)(shape1+4) */
  uint64_t n;         /* particles in the grid (FlatIndex.index.len()) */
  uint64_t n_cells;   /* non-empty cells (cells.len(); CellGrid::iter().count(), iters.rs:261-266) */
  int32_t ndim;
  int32_t dtype;
  int32_t keys_changed; /* FlatIndex::rebuild_mut's return value (flatindex.rs:140-152) for the
                           last rebuild when key tracking is on, else -1 */
  int32_t reserved;
} zb_info;

/* -- lifetime ------------------------------------------------------------------------------ */

/* CellGrid::default() (src/cellgrid.rs:112): an empty grid with cutoff 1. ndim is 2 or 3. */
int zb_grid_create(int dtype, int ndim, int device, zb_grid** out);
void zb_grid_destroy(zb_grid* g);

/* Stream all later calls on this handle are enqueued on (a cudaStream_t; NULL = legacy default
 * stream).  A fresh handle owns a private non-blocking stream. */
int zb_grid_set_stream(zb_grid* g, void* cuda_stream);

/* Keep the previous per-particle keys so that zb_info.keys_changed reports what
 * FlatIndex::rebuild_mut returns (flatindex.rs:113-153).  Off by default (costs one extra pass). */
int zb_grid_track_key_changes(zb_grid* g, int enable);

/* Stable cell storage: later rebuilds keep the particles of a cell in label (= input) order, as
 * CellStorage::push does (src/cellgrid/storage.rs:77-81), at the cost of one extra pass; it also
 * makes pair order inside a cell and the energy's summation order reproducible.  Off by default. */
int zb_grid_set_stable(zb_grid* g, int enable);

const char* zb_last_error(const zb_grid* g);

/* -- construction -------------------------------------------------------------------------- */

/* CellGrid::new / rebuild / rebuild_mut (src/cellgrid.rs:166-172, 187-238, 264-312).
 * cutoff_or_null == NULL keeps the previous cutoff (Option<T>::None, flatindex.rs:118).
 * Stages: Aabb::from_particles (util.rs:35-52) -> GridInfo::new (util.rs:191-220) ->
 * flat_cell_index per particle (util.rs:291-297) -> counting sort into one contiguous,
 * cell-sorted buffer (cellgrid.rs:196-231, storage.rs:77-81, 106-111). */
int zb_grid_rebuild(zb_grid* g, const void* xyz, uint64_t n, const double* cutoff_or_null);

/* Double buffering for callers that stream frames from HOST memory (trajectory analysis): starts the
 * host-to-device copy of a LATER rebuild's input on a private copy stream and returns at once, so
 * the copy overlaps the build / pair / LJ kernels of the current frame (two staging slots: call it
 * for frame k+1 before zb_grid_rebuild of frame k).  A zb_grid_rebuild with the same (xyz, n) uses
 * the staged copy instead of copying again.  xyz should be pinned memory
 * (pageable memory makes the copy synchronous); the caller must neither modify nor free it until that
 * rebuild (slots are matched by address; only zb_grid_rebuild consumes them, and a slot two rebuilds in a
 * row did not ask for is dropped). */
int zb_grid_prefetch(zb_grid* g, const void* xyz_host, uint64_t n);
/* Orders the handle's stream behind every copy zb_grid_prefetch has started (no host stall). */
int zb_grid_prefetch_wait(zb_grid* g);

/* Slab-sharded rebuild for multi-GPU runs (no counterpart upstream):
 * the bounding box is IMPOSED (the all-reduced global Aabb) so every rank derives the same
 * GridInfo and hence the same keys as a single-GPU grid; `xyz` holds this rank's particles of
 * the z-layers [z_begin - 1, z_end) (home layers plus the lower halo layer, halo particles
 * included by the caller); only cells of layers [z_begin, z_end) act as home cells, so every
 * unordered pair is owned by exactly one rank.  labels_or_null gives the global label of each
 * local particle (NULL: local order). */
int zb_grid_rebuild_sharded(zb_grid* g, const void* xyz, uint64_t n, const uint32_t* labels_or_null,
                            const double* cutoff_or_null, const double* inf, const double* sup,
                            int64_t z_begin, int64_t z_end);

/* Aabb::from_particles alone (util.rs:35-52): out[0..ndim) = inf, out[3..3+ndim) = sup as f64.
 * Used by the sharded host to all-reduce the global box before zb_grid_rebuild_sharded; with a
 * DEVICE out6 the call is asynchronous and the box can be all-reduced without a host round trip. */
int zb_aabb(zb_grid* g, const void* xyz, uint64_t n, double* out6);

/* Cell coordinate of every particle along one axis, floor((x[axis] - inf_axis) / cutoff) as i32 in
 * the grid's dtype (the per-axis term of flat_cell_index, util.rs:294-296).  The sharded host uses
 * it to assign particles to slabs and to pick the halo layer.  out: n int32 (host or device). */
int zb_layer_of(zb_grid* g, const void* xyz, uint64_t n, double inf_axis, double cutoff, int axis,
                int32_t* out);

/* Slab-local input of the sharded host: checks that all n particles lie in layers [z_begin, z_end)
 * of the slab axis (*out_of_slab = 1 otherwise) and compacts the particles of the top layer
 * z_end - 1 -- the next rank's lower halo -- into rows {x, y, z, label_offset + index} of the grid's
 * dtype at halo_rows[1 .. 1 + *n_top) (DEVICE memory, capacity cap_rows + 1 rows of 4 values; row 0
 * receives the row count as a number, so the block can be sent over NVLink as it is).  The label
 * is stored as raw bits in the 4th value (reinterpret as uint32 / int64, not a number).
 * With n_top == NULL and out_of_slab == NULL the call is asynchronous: the slab check is then
 * reported by the next zb_grid_rebuild_sharded on this handle (ZB_ERR_OUT_OF_WINDOW). */
int zb_slab_top_layer(zb_grid* g, const void* xyz, uint64_t n, double inf_axis, double cutoff,
                      int64_t z_begin, int64_t z_end, uint32_t label_offset, void* halo_rows,
                      uint64_t cap_rows, uint64_t* n_top, int* out_of_slab);

/* -- native multi-GPU step (one process per GPU) ------------------------- *
 * The same sequence zelll_b200/sharded.py drives through torch.distributed, issued by the library
 * itself with NCCL on the handle's stream: local Aabb -> all-reduce(min/max) -> identical GridInfo
 * on every rank -> z-layers split evenly -> top layer to rank + 1 (ncclSend/ncclRecv) -> sharded
 * counting sort -> [pairs / LJ] -> all-reduce(sum).  No counterpart upstream (the reference's only
 * parallelism is rayon over cells, src/cellgrid/iters.rs:282-290).  NCCL is resolved at run time
 * (dlopen of nccl_lib_path, else the libnccl.so.2 the process already uses). */
typedef struct zb_slab_info {
  double inf[3], sup[3]; /* the all-reduced bounding box */
  int32_t shape[3];
  int32_t reserved;
  int64_t z_begin, z_end; /* layers of the slab axis this rank owns */
  uint64_t n_local;
  uint64_t n_halo;        /* UINT64_MAX from zb_grid_rebuild_slab_local: the halo rows are counted on the
                             device; zb_grid_info().n - n_local gives the number once the step's verdict
                             has been collected (any later call that synchronises) */
} zb_slab_info;

/* rank 0 creates the 128-byte NCCL unique id; the launcher distributes it to all ranks */
int zb_comm_unique_id(const char* nccl_lib_path, void* out128);
int zb_comm_init(zb_grid* g, const char* nccl_lib_path, const void* unique_id128, int world, int rank);

/* Slab-local rebuild: buf (DEVICE, cap_rows x ndim values of the grid's dtype) holds this rank's
 * n_local particles -- exactly those of its own layers -- in its first rows; the halo rows received
 * from rank - 1 are appended behind them.  Labels are label_offset + row for local particles.
 * The grid needs at least one layer per rank along the slab axis (ZB_ERR_BAD_ARG otherwise, on every rank).
 * The call returns WITHOUT a host round trip for the halo count or the build's verdict (a particle outside
 * its slab, a halo larger than halo_cap / the spare rows of buf): those are reported by the next call on the
 * handle that synchronises -- zb_grid_lj_energy_allreduce folds them into its all-reduce, so every rank
 * learns of any rank's failure.  A step whose all-reduced bounding box equals the previous step's also skips
 * the wait for the box (it is checked on the device; if it did change, the step is repeated transparently,
 * by every rank).  buf must stay unchanged until that next synchronising call. */
int zb_grid_rebuild_slab_local(zb_grid* g, void* buf, uint64_t n_local, uint64_t cap_rows,
                               const double* cutoff_or_null, uint32_t label_offset, uint64_t halo_cap,
                               zb_slab_info* out);

/* zb_grid_lj_energy followed by the all-reduce(sum) of energy and pair count over the communicator: the one
 * host round trip of a slab step.  Every rank must call it, also a rank whose zb_grid_rebuild_slab_local
 * failed (it then contributes an error flag instead of hanging its peers); all ranks return an error if
 * any rank's step failed. */
int zb_grid_lj_energy_allreduce(zb_grid* g, int cmp, double filter_cutoff, double* energy, uint64_t* n_pairs);

/* -- inspection ---------------------------------------------------------------------------- */

/* CellGrid::info() (src/cellgrid.rs:346-348) */
int zb_grid_info(zb_grid* g, zb_info* out);

/* FlatIndex.index (flatindex.rs:19): the reference's flat cell key of every particle, in input
 * order.  out: n int32 (host or device). */
int zb_grid_keys(zb_grid* g, int32_t* out);

/* FlatIndex::neighbor_indices (flatindex.rs:55-65): 3^ndim - 1 relative flat keys; the first
 * half is the Half space (iters.rs:58-63).  out: room for 26 int32 (host). */
int zb_grid_neighbor_indices(zb_grid* g, int32_t* out, int32_t* count);

/* CellGrid::iter() (iters.rs:261-266): the non-empty cells, ascending key order: reference
 * flat key, first slot in the cell-sorted buffer, particle count.  Host arrays of capacity cap. */
int zb_grid_cells(zb_grid* g, int32_t* keys, uint32_t* begin, uint32_t* count, uint64_t cap,
                  uint64_t* n_out);

/* cell_storage() (src/cellgrid.rs:412-414): the cell-sorted buffer; labels: n uint32,
 * xyz: packed [n][ndim] of the grid's dtype (host or device; either may be NULL). */
int zb_grid_cell_storage(zb_grid* g, uint32_t* labels, void* xyz);

/* -- pair enumeration and its consumers ---------------------------------------------------- */

/* particle_pairs().filter(cmp).count() (src/cellgrid.rs:338-340 + benches/cellgrid.rs:84-88);
 * par_particle_pairs (cellgrid.rs:447-451) is the same call: the enumeration is parallel over
 * cells on the device.  filter_cutoff is squared in the grid's dtype (cutoff.powi(2)). */
int zb_grid_pair_count(zb_grid* g, int cmp, double filter_cutoff, uint64_t* out);

/* Materialised particle_pairs(): `ij` receives n_out rows (label_home, label_neighbor) of
 * uint32 (interleaved; host or device).  Row order is unspecified, as upstream (iters.rs:251).
 * ij == NULL or cap == 0 is a sizing call: *n_out = rows needed (ZB_ERR_CAPACITY unless that is 0).
 * A DEVICE buffer is filled in one pass over the pairs, without a counting pass in front of it; if it turns
 * out too small: ZB_ERR_CAPACITY, *n_out = needed, and the buffer's contents are unspecified.  A host buffer
 * that is too small is left untouched.  Rows [*n_out, cap) of a device buffer are scratch in either case;
 * nothing is written at or beyond row cap. */
int zb_grid_pairs(zb_grid* g, int cmp, double filter_cutoff, uint32_t* ij, uint64_t cap,
                  uint64_t* n_out);

/* Fused consumer of benches/lj.rs:42-47, 81-92: sum over kept pairs of 4*t*(t-1),
 * t = (1/dsq)^3, accumulated hierarchically in f64.  energy: 1 double (host or device);
 * n_pairs (optional, host or device): pairs kept. */
int zb_grid_lj_energy(zb_grid* g, int cmp, double filter_cutoff, double* energy, uint64_t* n_pairs);

/* -- point queries ------------------------------------------------- */

/* CellGrid::query_neighbors for a batch (src/cellgrid.rs:360-401): for each of nq query points,
 * the labels of all particles in the point's cell and its Full neighbourhood; with cmp != NONE
 * filtered like the Python binding's `neighbors` (python/src/lib.rs:229-241).
 * offsets: nq+1 uint64 (CSR over `labels`); a query outside the grid's one-cell margin
 * (try_cell_index == None, util.rs:245-256) gets offsets[q+1]-offsets[q] == 0 and valid[q] = 0.
 * If cap < needed: ZB_ERR_CAPACITY with *n_out = needed (offsets/valid are still written). */
int zb_grid_query_neighbors(zb_grid* g, const void* queries, uint64_t nq, int cmp,
                            double filter_cutoff, uint64_t* offsets, uint8_t* valid,
                            uint32_t* labels, uint64_t cap, uint64_t* n_out);

/* -- introspection for benches -------------------------------------------------------------- */

/* Per-stage device time of the hot launches, measured with cudaEvent pairs on the handle's stream
 * (bench.py's roofline numbers).  zb_grid_profile(g, 1) clears the accumulators and turns recording
 * on; zb_grid_profile_read synchronises the stream and returns, per stage, the summed milliseconds
 * and the number of launches since then (arrays of ZB_NSTAGES).  `enable`: 0 = off, 1 = every stage,
 * any other value = a mask in which bit (s + 1) selects stage s (recording costs two event records
 * per launch, ~2 % of a 1.6 ms step with every stage on: time one stage when the total matters). */
enum zb_stage {
  ZB_STAGE_BBOX = 0,       /* K1 Aabb::from_particles */
  ZB_STAGE_COUNT = 1,      /* K2 cell keys + histogram */
  ZB_STAGE_SCAN = 2,       /* K3 counts -> CSR offsets */
  ZB_STAGE_SCATTER = 3,    /* K4 records into cell order */
  ZB_STAGE_PAIR_COUNT = 4, /* K5 filtered pair count */
  ZB_STAGE_PAIR_EMIT = 5,  /* K5 pair list */
  ZB_STAGE_PAIR_LJ = 6,    /* K6 fused Lennard-Jones energy */
  ZB_STAGE_OTHER = 7,
  ZB_NSTAGES = 8
};
int zb_grid_profile(zb_grid* g, int enable);
int zb_grid_profile_read(zb_grid* g, double* stage_ms, uint64_t* stage_launches);

/* Number of kernel launches this handle has issued since creation (bench.py's gpu_launches). */
uint64_t zb_grid_launch_count(const zb_grid* g);
int zb_abi_version(void);
/* SHA-256 (hex) of the sources this binary was compiled from (zelll_b200/csrc/ and this header, as
 * hashed by zelll_b200/build.py): lets a test prove that the shipped, git-ignored .so matches the tracked sources. */
const char* zb_build_id(void);

#ifdef __cplusplus
}
#endif
#endif /* ZELLL_B200_H */
