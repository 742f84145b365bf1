// zelll_b200.cu -- host side of the C ABI declared in include/zelll_b200.h (libzelll_b200.so).
//
// One translation unit: the kernels of build_kernels.cuh / pair_kernels.cuh / query_kernels.cuh
// plus the handle that owns device memory, the stream and the launch sequence.  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true
// so that every floating-point operation is the separately rounded IEEE operation the Rust
// reference performs ("bit-exact pair set").
//
// There is no CPU fallback anywhere in this file: without a CUDA device zb_grid_create fails.
#include "../../include/zelll_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "build_kernels.cuh"
#include "pair_kernels.cuh"
#include "pair_pf_kernels.cuh"
#include "p2p_kernels.cuh"
#include "query_kernels.cuh"
#include "sparse_kernels.cuh"

using namespace zb;

// ---------------------------------------------------------------------------------------------
// handle

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

constexpr uint32_t kMaxBboxBlocks = 148 * 4;
constexpr uint32_t kMaxPairCtasPerSm = 8;
constexpr uint32_t kMaxEmitCtasPerSm = 4;  // pair-list passes: bounds the partly filled output slots (emit_impl)
constexpr uint64_t kMaxDenseCells = (1ull << 31) - 16;  // uint32 cell ids, table of 4 B entries

// small device scratch block, zeroed at creation; re-armed by the kernels / rebuild
struct Misc {
  unsigned ticket;         // bbox last-block ticket
  uint32_t tile_counter;   // scan: dynamic tile ids
  uint32_t nonempty;       // scan: non-empty cells
  int flags;               // count: bit0 = particle outside the window
  int keys_changed;        // keys_changed_kernel
  uint32_t pair_next;      // pair kernels: dynamic tile claim (self re-arming, 0 between launches)
  uint32_t pair_done;      // pair kernels: CTAs finished
  int pad;
  double out6[6];          // bbox result (T-typed, stored in the leading bytes)
  double energy;           // finalize_kernel
  unsigned long long pair_total;
  uint32_t slab_count;     // slab_top_kernel: rows in the halo block
  uint32_t slab_flag;      // bit0 = a particle outside the slab, bit1 = halo overflow, bit2 = box changed (speculative step)
  uint32_t halo_n;         // slab-local step: halo rows received (counted on the device)
  unsigned halo_ticket;    // p2p_halo_push_kernel: blocks finished (self re-arming)
  EmitFix emit_fix;        // host copy only: verdict of the last pair-list pass
};

}  // namespace

struct PairPlan {
  uint32_t tile_cells, ntiles, stage_recs, blocks;
  uint32_t row_tiles = 0;  // wide grids: tiles are segments of one x-row, this many per row
  size_t smem;        // dynamic shared memory of the kernel that runs first
  bool prefilter;     // f64 grids: pf_pair_kernel first, then the exact kernel over the work items it declined
  uint32_t stage_recs_exact;
  size_t smem_exact;  // ... of that second launch
};

struct zb_grid {
  int device = 0;
  int dtype = ZB_F64;
  int ndim = 3;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;

  // grid state (reference: GridInfo + FlatIndex + cells + cell_lists)
  bool built = false;
  double cutoff = 1.0;          // value of T, widened
  double inf[3] = {0, 0, 0};
  double sup[3] = {0, 0, 0};
  int shape[3] = {1, 1, 1};
  int wlo[3] = {0, 0, 0};
  int wshape[3] = {1, 1, 1};
  uint64_t n = 0;
  uint64_t n_cells_nonempty = 0;
  uint32_t ncells = 1;          // stored (windowed) cells
  uint32_t home_lo = 0, home_hi = 1;
  bool sharded = false;
  bool track_keys = false;
  bool stable = false;  // zb_grid_set_stable: records of a cell in label (= input) order
  bool slab_check_pending = false;  // an asynchronous zb_slab_top_layer awaits its verdict
  int keys_changed = -1;
  uint64_t n_keys_old = 0;

  // device memory, grown on demand and reused across rebuilds (rebuild_mut contract)
  DevBuf in;        // staged input when the caller passes host memory
  // zb_grid_prefetch: two staging slots filled on a private copy stream; a rebuild whose (host
  // pointer, n) matches a filled slot consumes it without copying
  struct Prefetch {
    DevBuf buf;
    const void* host = nullptr;   // non-null: the slot holds (or is receiving) this array
    uint64_t n = 0;
    cudaEvent_t copied = nullptr; // recorded on the copy stream after the H2D
    cudaEvent_t released = nullptr; // recorded on the main stream after the build that read the slot
    int age = 0;                  // rebuilds that passed this armed slot by (dropped at 2: it was never meant for them)
  } pre[2];
  cudaStream_t copy_stream = nullptr;
  int pre_in_use = -1;            // slot the current / last build read from
  DevBuf labels_in; // staged labels (sharded, host labels)
  DevBuf table;     // uint32 [4 + ncells + pad]; csr = table + 3, cursor = table + 4
  DevBuf sorted;    // Rec<T>[n]
  DevBuf scan_state;
  DevBuf partials;  // bbox partials
  DevBuf keys_old, keys_new;
  DevBuf tile_counts, tile_offsets, block_energy, block_totals;
  DevBuf emit_ctl, emit_tab, emit_spill, emit_tmp;  // zb_grid_pairs (emit_impl)
  DevBuf out_stage; // staging for host-destination outputs
  // sparse grids (sparse_kernels.cuh): compact sorted cells instead of the dense table
  bool sparse = false;
  uint32_t nuniq = 0;
  DevBuf skeys[2], sidx[2], shist, sflags, ukeys, ubegin;
  DevBuf pf_list;   // prefiltered pass: [count, work items...] left to the exact kernel
  const void* pf_list_zeroed = nullptr;  // the allocation whose count has been cleared once
  DevBuf tile_list; // sparse boxes: [count, tile ids...] of the tiles with home particles
  uint64_t tile_list_build = ~0ull;  // build_id / tile_cells the list was made for
  uint32_t tile_list_cells = 0, tile_list_rows = 0;
  Misc* misc = nullptr;       // device
  Misc* h_misc = nullptr;     // pinned host mirror for small read-backs
  uint32_t pair_ntiles_cap = 0;
  // experiment knobs, read from the environment ONCE at zb_grid_create (never on the launch path)
  struct Tune {
    // ZB_PREFILTER=<mask>: which consumers of an f64 grid run through the f32-prefiltered kernel
    // (1 = count, 2 = LJ, 4 = pair list; 0 = exact-arithmetic kernel everywhere).  Default: count only --
    // measured on B200 at n = 10^7: count 0.72 -> 0.6 ms, LJ 1.16 -> 1.43 ms, list 1.36 -> 2.06 ms
    // (deciding every "maybe" in f64 costs what the cheaper tests save).
    uint32_t prefilter = 1;
    // ZB_SPLIT=1: exact kernel as two launches, staged tiles (half the code size) then global-memory tiles.
    // Measured neutral on B200 (LJ 1.159 unsplit vs 1.178 ms split at n = 10^7): the 5 k-instruction kernel
    // is not instruction-cache bound, so one launch stays the default.
    bool split = false;
    bool row_tiles = true;  // ZB_ROW_TILES=0: wide grids read records through L1/L2 instead of staging row segments
    bool slab_spec = true;  // ZB_SLAB_SPEC=0: native slab steps always wait for the box all-reduce
    bool p2p = true;        // ZB_P2P=0: the slab step's exchanges go through NCCL instead of mapped peer memory
    uint32_t p2p_halo_rows = 8192;  // ZB_P2P_HALO_ROWS: rows of the mapped halo blocks
    uint32_t stage_recs = 0;   // ZB_STAGE_RECS: records per shared-memory stage (0 = default)
    uint32_t tile_cells = 0;   // ZB_TILE_CELLS: home cells per tile (0 = derived from the load)
    // ZB_SPARSE: 0 = never use the compact-cell build (boxes beyond 2^31 cells are refused, as in round 1),
    // 2 = always use it (tests), default 1 = when the dense table would cost more than ~4x the records
    int sparse = 1;
  } tune;

  // native multi-GPU step (zb_comm_*): NCCL entry points resolved at run time from the NCCL the
  // process already uses (torch's), so the library has no link-time NCCL dependency
  struct Nccl {
    void* dl = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
  } nccl;
  // slab-local step without host round trips: the halo count stays on the device and the build's verdict
  // travels with the energy all-reduce; `slab_pending` = that verdict has not been looked at yet.  A step may
  // also run SPECULATIVELY with the previous step's global box (no wait for the box all-reduce); the box is
  // then checked on the device and, if it changed, every rank repeats the step (all see the same box).
  struct SlabStep {
    bool pending = false, spec = false, have_box = false;
    double box[6] = {0, 0, 0, 0, 0, 0};  // (-inf, sup) of the last completed step, as all-reduced
    int backoff = 0, penalty = 0;        // steps to run without speculation after a miss
    void* buf = nullptr;
    uint64_t n_local = 0, cap_rows = 0, halo_cap = 0;
    uint32_t label_offset = 0;
  } slab;
  // exchanges over NVLink peer memory instead of NCCL (p2p_kernels.cuh): every rank's mailbox is mapped into
  // every process of the node at zb_comm_init
  struct P2p {
    bool ok = false;
    void* local = nullptr;
    void* opened[kP2pMaxWorld] = {nullptr};
    P2pPeers peers{};
    uint64_t halo_rows = 0;    // capacity of the mapped halo blocks (rows)
    size_t block_bytes = 0;
    unsigned long long seq_box = 0, seq_energy = 0, seq_halo = 0;
  } p2p;
  DevBuf halo_send, halo_recv, halo_labels, red;  // halo blocks, halo labels, 8-double reduction scratch
  double* h_red = nullptr;                        // pinned mirror of `red`
  uint64_t n_local = 0, n_halo = 0;
  uint64_t build_id = 0;  // bumped by every rebuild
  // plain (unsharded, no key tracking) rebuilds return without a final host sync: the non-empty
  // cell count arrives through this event and is collected by the first call that needs it
  cudaEvent_t info_event = nullptr;
  bool info_pending = false;
  struct OccEntry { const void* kern; size_t smem; int occ; };
  std::vector<OccEntry> occ_cache;  // launch_pairs: occupancy per (kernel, dynamic smem)
  cudaEvent_t bbox_event = nullptr;  // fires when the bounding box has reached the host
  bool built_once = false;           // the table / scan state hold a previous build (sizes in ncells)
  uint32_t precleared = 0;           // cells whose table entries (+ scan state, counters) were cleared speculatively
  // zb_grid_pairs: the length a sizing call found (same build, same filter: still exact)
  bool pairs_sized = false;
  uint64_t pairs_sized_build = 0, pairs_sized_total = 0;
  int pairs_sized_cmp = 0;
  double pairs_sized_fc = 0.0;

  // optional per-stage device timing (zb_grid_profile): cudaEvent pairs around the hot launches
  uint32_t profile = 0;  // bit s: record stage s
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_free;
  double stage_ms[ZB_NSTAGES] = {0};
  uint64_t stage_launches[ZB_NSTAGES] = {0};
};

static int slab_collect(zb_grid* g, bool* redone);

namespace {

int fail(zb_grid* g, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (g) g->err = buf;
  return code;
}

// records an event pair around one launch when profiling is on
struct StageSpan {
  zb_grid* g;
  cudaEvent_t b = nullptr;
  StageSpan(zb_grid* g_, int stage) : g(g_) {
    if (!(g->profile >> stage & 1u)) return;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (auto& e : ev) {
      if (!g->ev_free.empty()) {
        e = g->ev_free.back();
        g->ev_free.pop_back();
      } else if (cudaEventCreate(&e) != cudaSuccess) {
        cudaGetLastError();
        return;
      }
    }
    cudaEventRecord(ev[0], g->stream);
    b = ev[1];
    g->spans.push_back({stage, ev[0], ev[1]});
  }
  ~StageSpan() {
    if (b) cudaEventRecord(b, g->stream);
  }
};

#define ZB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(g, ZB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                  __FILE__, __LINE__);                                                         \
  } while (0)

#define ZB_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != ZB_OK) return rc__; \
  } while (0)

int reserve(zb_grid* g, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return ZB_OK;
  if (b.p) {
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    ZB_CUDA(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = std::max<size_t>(bytes, 256);
  ZB_CUDA(cudaMalloc(&b.p, want));
  b.cap = want;
  return ZB_OK;
}

// 1 = device-accessible pointer, 0 = plain host memory
int is_device_ptr(const void* p) {
  if (!p) return 0;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

size_t elem_size(const zb_grid* g) { return g->dtype == ZB_F32 ? 4 : 8; }

uint32_t* csr_ptr(zb_grid* g) {
  return g->sparse ? static_cast<uint32_t*>(g->ubegin.p) : static_cast<uint32_t*>(g->table.p) + 3;
}
uint32_t* cursor_ptr(zb_grid* g) { return static_cast<uint32_t*>(g->table.p) + 4; }

template <class T>
GridParams<T> make_params(const zb_grid* g) {
  GridParams<T> p;
  for (int d = 0; d < 3; ++d) {
    p.inf[d] = (T)g->inf[d];
    p.shape[d] = g->shape[d];
    p.wlo[d] = g->wlo[d];
    p.wshape[d] = g->wshape[d];
  }
  p.cutoff = (T)g->cutoff;
  p.ndim = g->ndim;
  p.ncells = g->ncells;
  return p;
}

// copy `bytes` from device memory of the handle to a caller pointer (host or device)
int deliver(zb_grid* g, void* dst, const void* dev_src, size_t bytes) {
  if (!bytes) return ZB_OK;
  if (is_device_ptr(dst)) {
    ZB_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, g->stream));
  } else {
    ZB_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, g->stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
  }
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------
// K1 launch: bounding box of a device array -> g->misc->out6 (as T)

template <class T>
int launch_bbox(zb_grid* g, const T* xyz, uint64_t n) {
  const uint64_t total = n * (uint64_t)g->ndim;
  constexpr int V = Vec16<T>::n;
  const bool aligned = (reinterpret_cast<uintptr_t>(xyz) % 16) == 0;
  const uint64_t per_block = (uint64_t)kBboxThreads * (aligned ? V : 1) * 4;
  uint32_t blocks = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((total + per_block - 1) / per_block, 1),
                                                 std::min<uint64_t>((uint64_t)g->sm_count * 4, kMaxBboxBlocks));
  ZB_TRY(reserve(g, g->partials, (size_t)kMaxBboxBlocks * 6 * sizeof(double)));
  T* partials = static_cast<T*>(g->partials.p);
  T* out6 = reinterpret_cast<T*>(g->misc->out6);
  unsigned* ticket = &g->misc->ticket;
  StageSpan span(g, ZB_STAGE_BBOX);
  if (g->ndim == 3) {
    if (aligned) bbox_kernel<T, 3, V><<<blocks, kBboxThreads, 0, g->stream>>>(xyz, n, partials, ticket, out6);
    else bbox_kernel<T, 3, 1><<<blocks, kBboxThreads, 0, g->stream>>>(xyz, n, partials, ticket, out6);
  } else {
    if (aligned) bbox_kernel<T, 2, V><<<blocks, kBboxThreads, 0, g->stream>>>(xyz, n, partials, ticket, out6);
    else bbox_kernel<T, 2, 1><<<blocks, kBboxThreads, 0, g->stream>>>(xyz, n, partials, ticket, out6);
  }
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  return ZB_OK;
}

template <class T>
int fetch_bbox(zb_grid* g, double* out6, bool preclear) {
  ZB_CUDA(cudaMemcpyAsync(g->h_misc->out6, g->misc->out6, 6 * sizeof(T), cudaMemcpyDeviceToHost, g->stream));
  if (!g->bbox_event) ZB_CUDA(cudaEventCreateWithFlags(&g->bbox_event, cudaEventDisableTiming));
  ZB_CUDA(cudaEventRecord(g->bbox_event, g->stream));
  // While the box travels to the host the GPU would idle: clear the count table, scan state and
  // counters for the PREVIOUS build's cell count now (rebuild_mut of a similar cloud needs exactly
  // that); build_sorted clears again only if this box turns out to need more cells.
  g->precleared = 0;
  if (preclear && g->built_once && g->table.p && g->scan_state.p && g->ncells > 0) {
    const size_t nc = g->ncells;
    const size_t ntile = (nc + kScanTile - 1) / kScanTile;
    if ((4 + nc) * 4 <= g->table.cap && ntile * sizeof(unsigned long long) <= g->scan_state.cap) {
      ZB_CUDA(cudaMemsetAsync(g->table.p, 0, (4 + nc) * 4, g->stream));
      ZB_CUDA(cudaMemsetAsync(g->scan_state.p, 0, ntile * sizeof(unsigned long long), g->stream));
      ZB_CUDA(cudaMemsetAsync(&g->misc->tile_counter, 0, 3 * sizeof(uint32_t), g->stream));
      g->precleared = (uint32_t)nc;
    }
  }
  ZB_CUDA(cudaEventSynchronize(g->bbox_event));
  const T* t = reinterpret_cast<const T*>(g->h_misc->out6);
  for (int d = 0; d < 6; ++d) out6[d] = (double)t[d];
  return ZB_OK;
}

// GridInfo::new (util.rs:191-220), in the arithmetic of T
template <class T>
int derive_shape(zb_grid* g) {
  const T c = (T)g->cutoff;
  for (int d = 0; d < 3; ++d) g->shape[d] = 1;
  for (int d = 0; d < g->ndim; ++d) {
    const T q = std::floor(((T)g->sup[d] - (T)g->inf[d]) / c);
    // Rust `as i32`: saturating, NaN -> 0; then `+ 1` (wrapping in release builds)
    int s;
    if (q != q) s = 0;
    else if (q >= (T)2147483648.0) s = 2147483647;
    else if (q <= (T)-2147483649.0) s = (-2147483647 - 1);
    else s = (int)q;
    if (s < 0 || s >= 2147483647 - 8)
      return fail(g, ZB_ERR_GRID_TOO_LARGE, "axis %d needs %d cells (cutoff %g too small or not positive)", d, s,
                  g->cutoff);
    g->shape[d] = s + 1;
  }
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------
// K2..K4 launches: counting sort of `xyz` (device) into g->sorted / g->table

// Slab-local multi-GPU step: the count pass over a rank's own rows also extracts its top layer
// (TopLayerOut); `exchange` then trades halos with the neighbours, appends the received rows behind
// the local ones and returns their number -- they are counted by a second, tiny launch.
template <class T>
struct SlabHook {
  TopLayerOut<T> top;
  uint64_t n_cap;                              // upper bound of local + halo rows (buffer sizing)
  // trades halos; *halo_rows = how many rows the launch over the halo must cover (a capacity: the number
  // actually received stays on the device, in misc->halo_n)
  std::function<int(uint64_t* halo_rows)> exchange;
};

template <class T>
int build_sorted(zb_grid* g, const T* xyz, const LabelSrc& labels, uint64_t& n, SlabHook<T>* hook = nullptr) {
  // window -> stored cell count
  uint64_t nc = 1;
  for (int d = 0; d < 3; ++d) {
    nc *= (uint64_t)g->wshape[d];
    if (nc > kMaxDenseCells)
      return fail(g, ZB_ERR_GRID_TOO_LARGE,
                  "bounding box / cutoff needs more than 2^31 cells (%d x %d x %d): box too sparse for the "
                  "dense cell table",
                  g->wshape[0], g->wshape[1], g->wshape[2]);
  }
  g->ncells = (uint32_t)nc;
  const size_t table_elems = 4 + (size_t)nc + kScanTile;  // slack so full-tile vector stores stay in bounds
  size_t free_b = 0, total_b = 0;
  if (table_elems * 4 > g->table.cap) {
    ZB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (table_elems * 4 > free_b + g->table.cap)
      return fail(g, ZB_ERR_GRID_TOO_LARGE, "dense cell table of %zu cells does not fit device memory", (size_t)nc);
  }
  ZB_TRY(reserve(g, g->table, table_elems * 4));
  // + 16 records of slack: the packed / fused test loops of the pair kernels may read (and discard) up to 15 records
  // past a home cell
  ZB_TRY(reserve(g, g->sorted, ((size_t)(hook ? hook->n_cap : n) + 16) * sizeof(Rec<T>)));
  const uint32_t ntile = (uint32_t)((nc + kScanTile - 1) / kScanTile);
  ZB_TRY(reserve(g, g->scan_state, (size_t)ntile * sizeof(unsigned long long)));

  if (g->precleared == 0 || g->precleared < nc) {  // not already cleared while the box was fetched
    // one memset clears the 4 leading entries (csr[0] = 0) and all counts
    ZB_CUDA(cudaMemsetAsync(g->table.p, 0, (4 + (size_t)nc) * 4, g->stream));
    ZB_CUDA(cudaMemsetAsync(g->scan_state.p, 0, (size_t)ntile * sizeof(unsigned long long), g->stream));
    ZB_CUDA(cudaMemsetAsync(&g->misc->tile_counter, 0, 3 * sizeof(uint32_t), g->stream));  // tile_counter, nonempty, flags
  }
  g->precleared = 0;
  g->built_once = true;

  const GridParams<T> p = make_params<T>(g);
  uint32_t* cursor = cursor_ptr(g);
  auto count = [&](const T* rows, uint64_t m, const TopLayerOut<T>* top, const uint32_t* m_dev) {
    if (m == 0) return;
    StageSpan span(g, rows == xyz ? ZB_STAGE_COUNT : ZB_STAGE_OTHER);  // the halo rows' launch is not "the" K2
    const uint32_t blocks = (uint32_t)((m + kPointThreads * kPointIlp - 1) / (kPointThreads * kPointIlp));
    const TopLayerOut<T> tl = top ? *top : TopLayerOut<T>{};
    if (g->ndim == 3) {
      if (top) count_kernel<T, 3, true><<<blocks, kPointThreads, 0, g->stream>>>(rows, (uint32_t)m, p, cursor, &g->misc->flags, tl, m_dev);
      else count_kernel<T, 3, false><<<blocks, kPointThreads, 0, g->stream>>>(rows, (uint32_t)m, p, cursor, &g->misc->flags, tl, m_dev);
    } else {
      if (top) count_kernel<T, 2, true><<<blocks, kPointThreads, 0, g->stream>>>(rows, (uint32_t)m, p, cursor, &g->misc->flags, tl, m_dev);
      else count_kernel<T, 2, false><<<blocks, kPointThreads, 0, g->stream>>>(rows, (uint32_t)m, p, cursor, &g->misc->flags, tl, m_dev);
    }
    g->launches++;
  };
  count(xyz, n, hook ? &hook->top : nullptr, nullptr);
  const uint64_t n_own = n;
  uint64_t halo_rows = 0;
  if (hook) {
    ZB_TRY(hook->exchange(&halo_rows));
    count(xyz + n * (uint64_t)g->ndim, halo_rows, nullptr, &g->misc->halo_n);
  }
  {
    StageSpan span(g, ZB_STAGE_SCAN);
    scan_kernel<<<ntile, kScanThreads, 0, g->stream>>>(cursor, (uint32_t)nc,
                                                       static_cast<unsigned long long*>(g->scan_state.p),
                                                       &g->misc->tile_counter, &g->misc->nonempty);
  }
  g->launches++;
  if (n + halo_rows > 0) {
    StageSpan span(g, ZB_STAGE_SCATTER);
    const uint64_t m = n + halo_rows;  // launch capacity; the kernel stops at n_own + *halo_n
    const uint32_t blocks = (uint32_t)((m + kPointThreads * kPointIlp - 1) / (kPointThreads * kPointIlp));
    Rec<T>* sorted = static_cast<Rec<T>*>(g->sorted.p);
    const uint32_t* n_dev = hook ? &g->misc->halo_n : nullptr;
    if (g->ndim == 3)
      scatter_kernel<T, 3><<<blocks, kPointThreads, 0, g->stream>>>(xyz, labels, (uint32_t)m, p, cursor, sorted, (uint32_t)n_own, n_dev);
    else
      scatter_kernel<T, 2><<<blocks, kPointThreads, 0, g->stream>>>(xyz, labels, (uint32_t)m, p, cursor, sorted, (uint32_t)n_own, n_dev);
    g->launches++;
  }
  if (g->stable && n > 1) {
    cell_sort_kernel<T><<<(uint32_t)((nc + 127) / 128), 128, 0, g->stream>>>(csr_ptr(g), (uint32_t)nc,
                                                                            static_cast<Rec<T>*>(g->sorted.p));
    g->launches++;
  }
  ZB_CUDA(cudaGetLastError());
  return ZB_OK;
}

// exclusive in-place scan of a uint32 array with K3's kernel; *nonempty (misc) counts its non-zero entries
int scan_u32(zb_grid* g, uint32_t* a, uint32_t m) {
  const uint32_t ntile = (m + kScanTile - 1) / kScanTile;
  ZB_TRY(reserve(g, g->scan_state, (size_t)ntile * sizeof(unsigned long long)));
  ZB_CUDA(cudaMemsetAsync(g->scan_state.p, 0, (size_t)ntile * sizeof(unsigned long long), g->stream));
  ZB_CUDA(cudaMemsetAsync(&g->misc->tile_counter, 0, 2 * sizeof(uint32_t), g->stream));  // tile_counter, nonempty
  scan_kernel<<<ntile, kScanThreads, 0, g->stream>>>(a, m, static_cast<unsigned long long*>(g->scan_state.p),
                                                     &g->misc->tile_counter, &g->misc->nonempty);
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  return ZB_OK;
}

// Sparse build (sparse_kernels.cuh): 64-bit cell keys -> radix sort -> compact sorted cells + CSR -> records.
// O(n) memory whatever the box; `key_bits` = width of the largest key.  One host round trip (the number of
// non-empty cells sizes the pair pass).
template <class T>
int build_sorted_sparse(zb_grid* g, const T* xyz, const LabelSrc& labels, uint64_t n, int key_bits) {
  const uint32_t nn = (uint32_t)n;
  for (int k = 0; k < 2; ++k) {
    ZB_TRY(reserve(g, g->skeys[k], (size_t)n * 8));
    ZB_TRY(reserve(g, g->sidx[k], (size_t)n * 4));
  }
  const uint32_t nblk = (nn + kSortTile - 1) / kSortTile;
  // the scan kernel writes whole 16-element groups: pad to its tile
  ZB_TRY(reserve(g, g->shist, ((size_t)256 * nblk + kScanTile) * 4));
  ZB_TRY(reserve(g, g->sflags, ((size_t)n + 1 + kScanTile) * 4));
  ZB_TRY(reserve(g, g->ukeys, ((size_t)n + 1) * 8));
  ZB_TRY(reserve(g, g->ubegin, ((size_t)n + 2) * 4));
  ZB_TRY(reserve(g, g->sorted, ((size_t)n + 16) * sizeof(Rec<T>)));
  ZB_CUDA(cudaMemsetAsync(&g->misc->flags, 0, sizeof(int), g->stream));
  const GridParams<T> p = make_params<T>(g);
  const uint32_t pblocks = (nn + 255) / 256;
  auto* k0 = static_cast<unsigned long long*>(g->skeys[0].p);
  auto* k1 = static_cast<unsigned long long*>(g->skeys[1].p);
  auto* i0 = static_cast<uint32_t*>(g->sidx[0].p);
  auto* i1 = static_cast<uint32_t*>(g->sidx[1].p);
  if (g->ndim == 3) keys64_kernel<T, 3><<<pblocks, 256, 0, g->stream>>>(xyz, nn, p, k0, i0, &g->misc->flags);
  else keys64_kernel<T, 2><<<pblocks, 256, 0, g->stream>>>(xyz, nn, p, k0, i0, &g->misc->flags);
  g->launches++;
  uint32_t* hist = static_cast<uint32_t*>(g->shist.p);
  // (+1 bit: the all-ones key of an out-of-box particle must sort last -- it does in every pass that runs,
  // and passes above key_bits see only zeros for real keys)
  for (int shift = 0; shift < key_bits; shift += 8) {
    radix_hist_kernel<<<nblk, kSortThreads, 0, g->stream>>>(k0, nn, shift, nblk, hist);
    g->launches++;
    ZB_TRY(scan_u32(g, hist, 256u * nblk));
    radix_scatter_kernel<<<nblk, kSortThreads, 0, g->stream>>>(k0, i0, nn, shift, nblk, hist, k1, i1);
    g->launches++;
    std::swap(k0, k1);
    std::swap(i0, i1);
  }
  uint32_t* flags = static_cast<uint32_t*>(g->sflags.p);
  heads_kernel<<<(nn + 1 + 255) / 256, 256, 0, g->stream>>>(k0, nn, flags);
  g->launches++;
  ZB_TRY(scan_u32(g, flags, nn + 1));  // misc->nonempty = number of heads = non-empty cells
  uniq_compact_kernel<<<pblocks, 256, 0, g->stream>>>(k0, flags, nn, static_cast<unsigned long long*>(g->ukeys.p),
                                                     static_cast<uint32_t*>(g->ubegin.p));
  Rec<T>* sorted = static_cast<Rec<T>*>(g->sorted.p);
  if (g->ndim == 3) gather_kernel<T, 3><<<pblocks, 256, 0, g->stream>>>(xyz, i0, nn, labels, sorted);
  else gather_kernel<T, 2><<<pblocks, 256, 0, g->stream>>>(xyz, i0, nn, labels, sorted);
  g->launches += 2;
  ZB_CUDA(cudaGetLastError());
  ZB_CUDA(cudaMemcpyAsync(&g->h_misc->tile_counter, &g->misc->tile_counter, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  if (g->h_misc->flags & 1) return fail(g, ZB_ERR_BAD_ARG, "a particle has a non-finite coordinate");
  g->nuniq = g->h_misc->nonempty;
  return ZB_OK;
}

template <class T>
int track_keys(zb_grid* g) {
  // FlatIndex::rebuild_mut's return value (flatindex.rs:140-152)
  const uint64_t n = g->n;
  ZB_TRY(reserve(g, g->keys_new, std::max<uint64_t>(n, 1) * 4));
  ZB_CUDA(cudaMemsetAsync(&g->misc->keys_changed, 0, sizeof(int), g->stream));
  if (n) {
    const uint32_t blocks = (uint32_t)((n + 255) / 256);
    keys_kernel<T><<<blocks, 256, 0, g->stream>>>(static_cast<const Rec<T>*>(g->sorted.p), (uint32_t)n,
                                                   make_params<T>(g), static_cast<int32_t*>(g->keys_new.p));
    keys_changed_kernel<<<blocks, 256, 0, g->stream>>>(static_cast<const int32_t*>(g->keys_old.p),
                                                       (uint32_t)g->n_keys_old,
                                                       static_cast<const int32_t*>(g->keys_new.p), (uint32_t)n,
                                                       &g->misc->keys_changed);
    g->launches += 2;
  }
  ZB_CUDA(cudaGetLastError());
  std::swap(g->keys_old, g->keys_new);
  g->n_keys_old = n;
  return ZB_OK;
}

// stage caller input on the device if it is host memory.  Only a rebuild (`match_prefetch`) may
// consume a zb_grid_prefetch slot; zb_aabb / zb_layer_of / zb_slab_top_layer always copy.  A slot is
// matched by (host pointer, n): the caller must leave the buffer untouched -- and alive -- between the
// prefetch and the rebuild that consumes it (the contract of any asynchronous copy).  An armed slot
// that two rebuilds in a row did not ask for is dropped, so a stale frame cannot wait forever for an
// unrelated array that happens to reuse its address.
int stage_input(zb_grid* g, const void* xyz, uint64_t n, const void** dev, bool match_prefetch = false) {
  const size_t bytes = (size_t)n * g->ndim * elem_size(g);
  if (n == 0) {
    *dev = nullptr;
    return ZB_OK;
  }
  if (is_device_ptr(xyz)) {
    *dev = xyz;
    return ZB_OK;
  }
  if (match_prefetch) {
    int hit = -1;
    for (int k = 0; k < 2; ++k) {
      auto& sl = g->pre[k];
      if (hit < 0 && sl.host == xyz && sl.n == n && sl.buf.p) hit = k;
      else if (sl.host && ++sl.age >= 2) sl.host = nullptr;  // nobody came for it
    }
    if (hit >= 0) {
      auto& sl = g->pre[hit];
      // the copy was started by zb_grid_prefetch on the copy stream: wait for it on the device
      ZB_CUDA(cudaStreamWaitEvent(g->stream, sl.copied, 0));
      sl.host = nullptr;  // consumed (the slot stays reserved until `released` fires)
      g->pre_in_use = hit;
      *dev = sl.buf.p;
      return ZB_OK;
    }
  }
  ZB_TRY(reserve(g, g->in, bytes));
  ZB_CUDA(cudaMemcpyAsync(g->in.p, xyz, bytes, cudaMemcpyHostToDevice, g->stream));
  *dev = g->in.p;
  return ZB_OK;
}

template <class T>
int rebuild_impl(zb_grid* g, const void* xyz_any, uint64_t n, const uint32_t* labels_any, const double* cutoff,
                 const double* inf, const double* sup, int64_t z_begin, int64_t z_end, bool sharded,
                 const LabelSrc* label_src = nullptr, SlabHook<T>* hook = nullptr) {
  if (n > 2147483647ull) return fail(g, ZB_ERR_TOO_MANY, "n = %llu exceeds i32::MAX", (unsigned long long)n);
  if (n > 0 && !xyz_any) return fail(g, ZB_ERR_BAD_ARG, "xyz is NULL");
  if (cutoff) {
    const T c = (T)*cutoff;
    if (!(c > (T)0) || !std::isfinite((double)c)) return fail(g, ZB_ERR_BAD_ARG, "cutoff must be positive and finite");
    g->cutoff = (double)c;
  }
  g->built = false;
  g->build_id++;
  g->precleared = 0;
  if (g->info_pending) {  // the previous build's counters are about to be overwritten
    ZB_CUDA(cudaEventSynchronize(g->info_event));
    g->info_pending = false;
  }
  const void* dev = nullptr;
  ZB_TRY(stage_input(g, xyz_any, n, &dev, true));
  const T* xyz = static_cast<const T*>(dev);

  const uint32_t* labels = nullptr;
  if (labels_any && n) {
    if (is_device_ptr(labels_any)) labels = labels_any;
    else {
      ZB_TRY(reserve(g, g->labels_in, n * 4));
      ZB_CUDA(cudaMemcpyAsync(g->labels_in.p, labels_any, n * 4, cudaMemcpyHostToDevice, g->stream));
      labels = static_cast<const uint32_t*>(g->labels_in.p);
    }
  }

  if (!sharded) {
    if (n == 0) {
      // Aabb::from_particles on an empty iterator: zeros (util.rs:41)
      for (int d = 0; d < 3; ++d) g->inf[d] = g->sup[d] = 0.0;
    } else {
      ZB_TRY(launch_bbox<T>(g, xyz, n));
      double o[6];
      ZB_TRY(fetch_bbox<T>(g, o, true));
      for (int d = 0; d < 3; ++d) {
        g->inf[d] = d < g->ndim ? o[d] : 0.0;
        g->sup[d] = d < g->ndim ? o[3 + d] : 0.0;
      }
      for (int d = 0; d < g->ndim; ++d)
        if (!std::isfinite(g->inf[d]) || !std::isfinite(g->sup[d]))
          return fail(g, ZB_ERR_BAD_ARG, "non-finite coordinate on axis %d", d);
    }
  } else {
    for (int d = 0; d < 3; ++d) {
      g->inf[d] = d < g->ndim ? (double)(T)inf[d] : 0.0;
      g->sup[d] = d < g->ndim ? (double)(T)sup[d] : 0.0;
    }
  }
  ZB_TRY(derive_shape<T>(g));
  for (int d = 0; d < 3; ++d) {
    g->wlo[d] = 0;
    g->wshape[d] = g->shape[d];
  }
  g->sharded = sharded;
  if (sharded) {
    const int ax = g->ndim - 1;  // slab axis = slowest (largest stride) axis
    if (z_begin < 0 || z_end > g->shape[ax] || z_begin > z_end)
      return fail(g, ZB_ERR_BAD_ARG, "slab [%lld, %lld) outside the %d layers of the grid", (long long)z_begin,
                  (long long)z_end, g->shape[ax]);
    const int lo = (int)std::max<int64_t>(z_begin - 1, 0);
    g->wlo[ax] = lo;
    g->wshape[ax] = std::max<int>((int)z_end - lo, 1);
  }
  LabelSrc ls{labels, nullptr, 0u, 0xffffffffu};  // no array: label = position
  if (label_src) ls = *label_src;
  if (hook) {  // window-relative layers of the rank's own slab [z_begin, z_end - 1]
    const int ax = g->ndim - 1;
    hook->top.first = (int)(z_begin - g->wlo[ax]);
    hook->top.top = (int)(z_end - 1 - g->wlo[ax]);
  }
  // Dense count table (4 B per cell of the box) or compact sorted cells (O(n) memory)?  The compact build
  // takes over when the table would be out of proportion to the particles or beyond 2^31 cells -- the
  // reference's hash map pays only for non-empty cells (README.md:21-22, src/cellgrid.rs:120).
  g->sparse = false;
  bool use_sparse = false;
  int key_bits = 1;
  if (!sharded && n > 0 && g->tune.sparse) {
    unsigned __int128 nc = 1;
    for (int d = 0; d < 3; ++d) nc *= (unsigned __int128)(uint64_t)g->wshape[d];
    use_sparse = g->tune.sparse == 2 || nc > (unsigned __int128)kMaxDenseCells ||
                 nc > (unsigned __int128)(32ull * n + (1ull << 22));
    if (use_sparse) {
      if (nc >= ((unsigned __int128)1 << 62))
        return fail(g, ZB_ERR_GRID_TOO_LARGE, "bounding box / cutoff needs more than 2^62 cells (%d x %d x %d)", g->wshape[0],
                    g->wshape[1], g->wshape[2]);
      while (((unsigned __int128)1 << key_bits) <= nc) ++key_bits;  // real keys < 2^key_bits - 1
    }
  }
  if (use_sparse) {
    ZB_TRY(build_sorted_sparse<T>(g, xyz, ls, n, key_bits));
    g->sparse = true;
    g->ncells = g->nuniq;  // "cells" of the pair pass = the compact non-empty cells
  } else {
    ZB_TRY(build_sorted<T>(g, xyz, ls, n, hook));  // n grows by the halo rows the hook received
  }

  // home-cell range of the pair kernels
  {
    uint64_t plane = 1;
    for (int d = 0; d < g->ndim - 1; ++d) plane *= (uint64_t)g->wshape[d];
    if (sharded) {
      const int ax = g->ndim - 1;
      g->home_lo = (uint32_t)(plane * (uint64_t)(z_begin - g->wlo[ax]));
      g->home_hi = (uint32_t)(plane * (uint64_t)(z_end - g->wlo[ax]));
    } else {
      g->home_lo = 0;
      g->home_hi = g->ncells;
    }
  }
  g->n = n;

  // flags / non-empty count come back with the info read (one small sync copy)
  ZB_CUDA(cudaMemcpyAsync(&g->h_misc->tile_counter, &g->misc->tile_counter, 3 * sizeof(uint32_t),
                          cudaMemcpyDeviceToHost, g->stream));
  const bool slab_check = sharded && g->slab_check_pending;
  if (slab_check)
    ZB_CUDA(cudaMemcpyAsync(&g->h_misc->slab_count, &g->misc->slab_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            g->stream));
  if (g->track_keys && !sharded) ZB_TRY(track_keys<T>(g));
  if (g->track_keys && !sharded)
    ZB_CUDA(cudaMemcpyAsync(&g->h_misc->keys_changed, &g->misc->keys_changed, sizeof(int), cudaMemcpyDeviceToHost,
                            g->stream));
  if (g->pre_in_use >= 0) {  // the staging slot this build read is free again
    ZB_CUDA(cudaEventRecord(g->pre[g->pre_in_use].released, g->stream));
    g->pre_in_use = -1;
  }
  if (hook) {
    // Native slab step: no host round trip here.  The build's verdict (window flag, slab flag, halo count)
    // is copied to pinned memory behind the kernels and looked at by the first call that synchronises
    // anyway (slab_validate; zb_grid_lj_energy_allreduce folds it into its all-reduce on the device).
    ZB_CUDA(cudaMemcpyAsync(&g->h_misc->slab_count, &g->misc->slab_count, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            g->stream));
    if (!g->info_event) ZB_CUDA(cudaEventCreateWithFlags(&g->info_event, cudaEventDisableTiming));
    ZB_CUDA(cudaEventRecord(g->info_event, g->stream));
    g->info_pending = true;
    g->slab_check_pending = false;
    g->slab.pending = true;
    g->keys_changed = -1;
    g->built = true;
    return ZB_OK;
  }
  if (!sharded && !g->track_keys) {
    // Nothing here can fail any more: every particle lies inside its own bounding box (a NaN
    // coordinate maps to cell 0 like Rust's `as i32`; infinities were rejected with the box), so the
    // window flag cannot be set.  Skip the host round trip; n_cells arrives lazily (collect_info).
    if (!g->info_event) ZB_CUDA(cudaEventCreateWithFlags(&g->info_event, cudaEventDisableTiming));
    ZB_CUDA(cudaEventRecord(g->info_event, g->stream));
    g->info_pending = true;
    g->keys_changed = -1;
    g->built = true;
    return ZB_OK;
  }
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  if (slab_check) {
    g->slab_check_pending = false;
    if (g->h_misc->slab_flag & 1u)
      return fail(g, ZB_ERR_OUT_OF_WINDOW, "slab-local input held a particle outside its own layers");
  }
  if (g->h_misc->flags & 1)
    return fail(g, sharded ? ZB_ERR_OUT_OF_WINDOW : ZB_ERR_BAD_ARG,
                sharded ? "a particle lies outside the imposed box / slab window"
                        : "a particle has a non-finite coordinate");
  g->n_cells_nonempty = g->h_misc->nonempty;
  g->keys_changed = (g->track_keys && !sharded) ? (g->h_misc->keys_changed ? 1 : 0) : -1;
  g->built = true;
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------
// pair kernels

template <class T>
PairPlan plan_pairs(const zb_grid* g, size_t warp_smem, size_t pf_warp_smem, int cmp, double fc, uint32_t pf_bit) {
  PairPlan pl;
  const uint64_t plane = (g->ndim == 3) ? (uint64_t)g->wshape[0] * g->wshape[1] : 0;
  const uint64_t halo = plane + (uint64_t)g->wshape[0] + 1;
  const uint32_t nhome = g->home_hi - g->home_lo;
  const double ppc = g->ncells ? (double)g->n / (double)g->ncells : 0.0;
  pl.stage_recs_exact = sizeof(T) == 8 ? 1408u : 2816u;  // 44 KB of stage: 3-4 CTAs per SM
  if (g->tune.stage_recs) pl.stage_recs_exact = g->tune.stage_recs;
  // f64 grids with a distance filter run through the f32 prefilter kernel (pair_pf_kernels.cuh) when a
  // tile of the expected load fits its stage and the squared radius is an ordinary f32 number
  pl.prefilter = false;
  pl.stage_recs = pl.stage_recs_exact;
  if (sizeof(T) == 8 && cmp != ZB_CMP_NONE && (g->tune.prefilter & pf_bit) && !g->sparse) {
    const double c2 = fc * fc;
    const uint32_t sr = kPfStageRecs;
    if (c2 > 1e-30 && c2 < 1e30 && halo + 9 < (uint64_t)kStageCells &&
        (double)(halo + 9) * std::max(ppc, 0.25) <= 0.75 * sr) {
      pl.prefilter = true;
      pl.stage_recs = sr;
    }
  }
  uint32_t tc = 64;
  if (halo + 1 + 8 < (uint64_t)kStageCells) {
    // fill ~75 % of the stage with the expected load, bounded by the staged CSR window
    const double budget = 0.75 * pl.stage_recs / std::max(ppc, 0.25);
    const double room = std::min<double>(budget, (double)kStageCells - 1) - (double)halo;
    if (room >= 8) tc = (uint32_t)room;
  }
  // enough tiles to balance the persistent grid
  const uint32_t want_tiles = (uint32_t)g->sm_count * 4 * 4;
  if (nhome / std::max(tc, 1u) < want_tiles) tc = std::max<uint32_t>(8, (nhome + want_tiles - 1) / want_tiles);
  if (g->tune.tile_cells) tc = g->tune.tile_cells;
  tc = std::min<uint32_t>(std::max<uint32_t>(tc, 1), kMaxTileCells);
  pl.tile_cells = tc;
  pl.ntiles = nhome ? (nhome + tc - 1) / tc : 0;
  // Wide grids (cubes, slabs many cells across): a tile plus its lower halo -- a whole plane of cells -- does
  // not fit the stage as ONE range.  Tiles then become segments of one x-row and the stage takes the five row
  // segments of the half shell (three in the plane below, the row before, the home row): 5 (T + 2) cells.
  pl.row_tiles = 0;
  if (!g->sparse && !pl.prefilter && g->tune.row_tiles && halo + 1 + 8 >= (uint64_t)kStageCells && nhome && !g->tune.tile_cells) {
    const uint32_t w0 = (uint32_t)g->wshape[0];
    const double fit = 0.8 * pl.stage_recs_exact / (5.0 * std::max(ppc, 0.25)) - 2.0;
    if (fit >= 4.0) {
      uint32_t seg = std::min<uint32_t>({(uint32_t)fit, w0, (uint32_t)kMaxTileCells});
      const uint32_t per_row = (w0 + seg - 1) / seg;
      seg = (w0 + per_row - 1) / per_row;  // equal segments
      pl.row_tiles = per_row;
      pl.tile_cells = seg;
      pl.ntiles = (nhome / w0) * per_row;
    }
  }
  pl.blocks = (uint32_t)g->sm_count * kMaxPairCtasPerSm * 2;  // upper bound (buffers): both launches of a prefiltered pass
  pl.smem_exact = (size_t)pl.stage_recs_exact * sizeof(Rec<T>) + kMaxTileCells * sizeof(CellRuns) +
                  (kStageCells + 4) * sizeof(uint32_t) + kPairWarps * warp_smem;
  pl.smem = pl.prefilter ? pf_smem_bytes(pf_warp_smem) : pl.smem_exact;
  return pl;
}

// q = (t + ((n - t) >> sh1)) >> sh2 with t = umulhi(n, mul): exact for every 32-bit n and d >= 1
FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;  // ceil(log2 d)
  f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 0 ? l - 1 : 0;
  return f;
}

template <class T>
PairParams<T> pair_params(zb_grid* g, const PairPlan& pl, double filter_cutoff) {
  PairParams<T> p;
  p.sorted = static_cast<const Rec<T>*>(g->sorted.p);
  p.csr = csr_ptr(g);
  p.w0 = g->wshape[0];
  p.w1 = g->wshape[1];
  p.w2 = g->wshape[2];
  p.home_lo = g->home_lo;
  p.home_hi = g->home_hi;
  p.tile_cells = pl.tile_cells;
  p.ntiles = pl.ntiles;
  p.row_tiles = pl.row_tiles;
  p.stage_recs = pl.stage_recs;
  p.fb_list = nullptr;
  p.fb_count = nullptr;
  p.work_list = nullptr;
  p.work_list_n = nullptr;
  p.ukeys = g->sparse ? static_cast<const unsigned long long*>(g->ukeys.p) : nullptr;
  const T c = (T)filter_cutoff;
  p.c2 = c * c;  // cutoff.powi(2) in T (benches/lj.rs:85)
  p.fc = c;
  p.cell = (T)g->cutoff;
  p.tile_list = nullptr;
  p.tile_list_n = nullptr;
  p.tile_next = &g->misc->pair_next;
  p.tile_done = &g->misc->pair_done;
  p.div0 = make_fastdiv((uint32_t)g->wshape[0]);
  p.div1 = make_fastdiv((uint32_t)g->wshape[1]);
  return p;
}

// Sparse boxes (few of the stored cells hold particles): list the tiles that have home particles
// once per build so the persistent CTAs do not walk millions of empty tiles.
template <class T>
int sparse_tile_list(zb_grid* g, const PairPlan& pl, PairParams<T>& p) {
  // Sparse for certain when even n (an upper bound of the non-empty cells) is small against the
  // cell count; otherwise the tiles are walked directly (a clumped cloud then meets empty tiles,
  // which the kernel skips cheaply).  Needs no host-side count: rebuilds return without a sync.
  if (pl.ntiles < 4096 || g->n * 8 > (uint64_t)g->ncells) return ZB_OK;
  ZB_TRY(reserve(g, g->tile_list, ((size_t)pl.ntiles + 1) * 4));
  uint32_t* buf = static_cast<uint32_t*>(g->tile_list.p);
  if (g->tile_list_build != g->build_id || g->tile_list_cells != pl.tile_cells || g->tile_list_rows != pl.row_tiles) {
    ZB_CUDA(cudaMemsetAsync(buf, 0, 4, g->stream));
    tile_list_kernel<<<(pl.ntiles + 255) / 256, 256, 0, g->stream>>>(csr_ptr(g), g->home_lo, g->home_hi, pl.tile_cells,
                                                                 pl.row_tiles, (uint32_t)g->wshape[0], pl.ntiles, buf + 1, buf);
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    g->tile_list_build = g->build_id;
    g->tile_list_cells = pl.tile_cells;
    g->tile_list_rows = pl.row_tiles;
  }
  p.tile_list = buf + 1;
  p.tile_list_n = buf;
  return ZB_OK;
}

// per-block output slots of the second launch of a prefiltered pass sit behind the first launch's
template <class T>
typename CountConsumer<T>::Args shift_args(typename CountConsumer<T>::Args a, uint32_t blocks, CountConsumer<T>*) {
  a.block_totals += blocks;
  return a;
}
template <class T>
typename LjConsumer<T>::Args shift_args(typename LjConsumer<T>::Args a, uint32_t blocks, LjConsumer<T>*) {
  a.block_energy += blocks;
  a.block_totals += blocks;
  return a;
}
template <class T>
typename EmitConsumer<T>::Args shift_args(typename EmitConsumer<T>::Args a, uint32_t, EmitConsumer<T>*) {
  return a;
}

template <class T, class Consumer>
int launch_pairs(zb_grid* g, int cmp, PairPlan& pl, const PairParams<T>& p_in, typename Consumer::Args args) {
  PairParams<T> p = p_in;
  ZB_TRY(sparse_tile_list<T>(g, pl, p));
  auto go = [&](auto kern, size_t smem, const PairParams<T>& pp, typename Consumer::Args a, uint32_t* blocks_out) -> int {
    // persistent grid: one wave of resident CTAs.  The attribute / occupancy queries cost ~10 us of
    // host time, so their result is cached per (kernel, shared-memory size).
    int occ = 0;
    const void* key = reinterpret_cast<const void*>(kern);
    for (const auto& e : g->occ_cache)
      if (e.kern == key && e.smem == smem) occ = e.occ;
    if (occ == 0) {
      ZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPairThreads, smem));
      occ = std::max(1, std::min(occ, (int)(Consumer::kStage == 5 ? kMaxEmitCtasPerSm : kMaxPairCtasPerSm)));
      g->occ_cache.push_back({key, smem, occ});
    }
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(pl.ntiles, (uint32_t)g->sm_count * (uint32_t)occ));
    {
      StageSpan span(g, Consumer::kStage);
      kern<<<blocks, kPairThreads, smem, g->stream>>>(pp, a);
    }
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    *blocks_out = blocks;
    return ZB_OK;
  };
  // the exact kernel in one of its three shapes: 0 = staged and global tiles in one kernel, 1 = staged tiles
  // only (others go to the work list), 2 = global-memory tiles only
  auto exact = [&](int mode, const PairParams<T>& pp, size_t smem, typename Consumer::Args a, uint32_t* blocks_out) -> int {
#define ZB_GO(M)                                                                          \
  switch (cmp) {                                                                          \
    case ZB_CMP_NONE: return go(pair_kernel<T, 0, Consumer, M>, smem, pp, a, blocks_out); \
    case ZB_CMP_LT: return go(pair_kernel<T, 1, Consumer, M>, smem, pp, a, blocks_out);   \
    case ZB_CMP_LE: return go(pair_kernel<T, 2, Consumer, M>, smem, pp, a, blocks_out);   \
  }
    if (mode == 1) { ZB_GO(1) } else if (mode == 2) { ZB_GO(2) } else if (mode == 3) { ZB_GO(3) } else { ZB_GO(0) }
#undef ZB_GO
    return fail(g, ZB_ERR_BAD_ARG, "bad cmp %d", cmp);
  };
  // the work list shared by the two launches of a split pass
  auto work_list = [&](PairParams<T>& pp) -> int {
    ZB_TRY(reserve(g, g->pf_list, ((size_t)pl.ntiles + 1) * 4));
    if (g->pf_list_zeroed != g->pf_list.p) {
      ZB_CUDA(cudaMemsetAsync(g->pf_list.p, 0, 4, g->stream));  // afterwards the second launch re-arms the count
      g->pf_list_zeroed = g->pf_list.p;
    }
    uint32_t* fb = static_cast<uint32_t*>(g->pf_list.p);
    pp.fb_count = fb;
    pp.fb_list = fb + 1;
    return ZB_OK;
  };
  const size_t smem_global = kMaxTileCells * sizeof(CellRuns) + 8 * sizeof(uint32_t) + kPairWarps * Consumer::kWarpSmemBytes;
  if (g->sparse) {  // compact cells: global-memory tiles with searched descriptors
    uint32_t b = 0;
    ZB_TRY(exact(3, p, smem_global, args, &b));
    pl.blocks = b;
    return ZB_OK;
  }
  if constexpr (sizeof(T) == 8) {
    if (pl.prefilter && cmp != ZB_CMP_NONE) {
      // 1. the f32-prefiltered kernel; work items it declines (oversized tiles, guard band too wide,
      //    non-finite coordinates) go to a list ...
      ZB_TRY(work_list(p));
      uint32_t b1 = 0, b2 = 0;
      if (cmp == ZB_CMP_LT) ZB_TRY(go(pf_pair_kernel<1, Consumer>, pl.smem, p, args, &b1));
      else ZB_TRY(go(pf_pair_kernel<2, Consumer>, pl.smem, p, args, &b1));
      // 2. ... which the exact kernel walks; its per-block results sit behind the first launch's
      PairParams<T> pe = p;
      pe.stage_recs = pl.stage_recs_exact;
      pe.work_list = p.fb_list;
      pe.work_list_n = p.fb_count;
      ZB_TRY(exact(0, pe, pl.smem_exact, shift_args<T>(args, b1, static_cast<Consumer*>(nullptr)), &b2));
      pl.blocks = b1 + b2;
      return ZB_OK;
    }
  }
  if (g->tune.split && cmp != ZB_CMP_NONE) {
    // staged tiles first (the small, hot kernel), then the tiles that did not fit the stage from global memory
    ZB_TRY(work_list(p));
    uint32_t b1 = 0, b2 = 0;
    ZB_TRY(exact(1, p, pl.smem_exact, args, &b1));
    PairParams<T> pe = p;
    pe.work_list = p.fb_list;
    pe.work_list_n = p.fb_count;
    ZB_TRY(exact(2, pe, smem_global, shift_args<T>(args, b1, static_cast<Consumer*>(nullptr)), &b2));
    pl.blocks = b1 + b2;
    return ZB_OK;
  }
  uint32_t b = 0;
  ZB_TRY(exact(0, p, pl.smem_exact, args, &b));
  pl.blocks = b;
  return ZB_OK;
}

int finalize(zb_grid* g, bool with_energy, uint32_t nblocks) {
  finalize_kernel<<<1, 256, 0, g->stream>>>(with_energy ? static_cast<const double*>(g->block_energy.p) : nullptr,
                                            static_cast<const unsigned long long*>(g->block_totals.p), nblocks,
                                            &g->misc->energy, &g->misc->pair_total);
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  return ZB_OK;
}

template <class T>
int pair_count_impl(zb_grid* g, int cmp, double fc) {
  PairPlan pl = plan_pairs<T>(g, CountConsumer<T>::kWarpSmemBytes, CountConsumer<T>::kPfWarpSmemBytes, cmp, fc, 1u);
  ZB_TRY(reserve(g, g->block_totals, (size_t)pl.blocks * 8));
  typename CountConsumer<T>::Args a;
  a.tile_counts = nullptr;
  a.block_totals = static_cast<unsigned long long*>(g->block_totals.p);
  // every launched block writes its slot of block_totals (finish()): clear only when nothing runs
  if (pl.ntiles) ZB_TRY((launch_pairs<T, CountConsumer<T>>(g, cmp, pl, pair_params<T>(g, pl, fc), a)));
  else ZB_CUDA(cudaMemsetAsync(g->block_totals.p, 0, (size_t)pl.blocks * 8, g->stream));
  ZB_TRY(finalize(g, false, pl.blocks));
  return ZB_OK;
}

template <class T>
int lj_impl(zb_grid* g, int cmp, double fc, uint32_t* blocks_out = nullptr) {
  PairPlan pl = plan_pairs<T>(g, LjConsumer<T>::kWarpSmemBytes, LjConsumer<T>::kPfWarpSmemBytes, cmp, fc, 2u);
  ZB_TRY(reserve(g, g->block_totals, (size_t)pl.blocks * 8));
  ZB_TRY(reserve(g, g->block_energy, (size_t)pl.blocks * 8));
  typename LjConsumer<T>::Args a;
  a.block_energy = static_cast<double*>(g->block_energy.p);
  a.block_totals = static_cast<unsigned long long*>(g->block_totals.p);
  if (pl.ntiles) {
    ZB_TRY((launch_pairs<T, LjConsumer<T>>(g, cmp, pl, pair_params<T>(g, pl, fc), a)));
  } else {  // every launched block writes its slots (finish()): clear only when nothing runs
    ZB_CUDA(cudaMemsetAsync(g->block_totals.p, 0, (size_t)pl.blocks * 8, g->stream));
    ZB_CUDA(cudaMemsetAsync(g->block_energy.p, 0, (size_t)pl.blocks * 8, g->stream));
  }
  if (blocks_out) *blocks_out = pl.blocks;  // the caller folds the per-block partials itself (p2p_energy_kernel)
  else ZB_TRY(finalize(g, true, pl.blocks));
  return ZB_OK;
}

// One pass over the pairs into `out_dev` (room for cap_rows rows); the list's length arrives in h_fix.
// See EmitConsumer (pair_kernels.cuh) for the slot scheme and the emit_fix_* kernels that close its holes.
template <class T>
int emit_impl(zb_grid* g, int cmp, double fc, uint2* out_dev, uint64_t cap_rows, EmitFix* h_fix) {
  PairPlan pl = plan_pairs<T>(g, EmitConsumer<T>::kWarpSmemBytes, EmitConsumer<T>::kPfWarpSmemBytes, cmp, fc, 4u);
  // every warp of every launch leaves at most two partly filled slots (EmitConsumer::finish); launch_pairs runs
  // at most two launches of at most kMaxEmitCtasPerSm CTAs per SM
  const uint32_t pcap = 2u * kPairWarps * 2u * (uint32_t)g->sm_count * kMaxEmitCtasPerSm;
  // control block: cursor (8), npartial (4, padded to 8), EmitFix; tables: partial[pcap], moves[pcap], poff[pcap]
  ZB_TRY(reserve(g, g->emit_ctl, 16 + sizeof(EmitFix)));
  ZB_TRY(reserve(g, g->emit_tab, (size_t)pcap * (8 + 8 + 4)));
  ZB_TRY(reserve(g, g->emit_spill, (size_t)pcap * kEmitChunk * 8));
  ZB_TRY(reserve(g, g->emit_tmp, (size_t)pcap * kEmitChunk * 8));
  unsigned char* ctl = static_cast<unsigned char*>(g->emit_ctl.p);
  unsigned char* tab = static_cast<unsigned char*>(g->emit_tab.p);
  ZB_CUDA(cudaMemsetAsync(ctl, 0, 16, g->stream));
  EmitOut o;
  o.out = out_dev;
  o.spill = static_cast<uint2*>(g->emit_spill.p);
  o.out_chunks = (uint32_t)std::min<uint64_t>(cap_rows / kEmitChunk, 0x7fffffffull);
  o.spill_chunks = pcap;
  o.cursor = reinterpret_cast<unsigned long long*>(ctl);
  o.npartial = reinterpret_cast<uint32_t*>(ctl + 8);
  o.partial = reinterpret_cast<uint2*>(tab);
  o.partial_cap = pcap;
  uint2* moves = reinterpret_cast<uint2*>(tab + (size_t)pcap * 8);
  uint32_t* poff = reinterpret_cast<uint32_t*>(tab + (size_t)pcap * 16);
  EmitFix* fix = reinterpret_cast<EmitFix*>(ctl + 16);
  if (pl.ntiles) ZB_TRY((launch_pairs<T, EmitConsumer<T>>(g, cmp, pl, pair_params<T>(g, pl, fc), o)));
  {
    StageSpan span(g, EmitConsumer<T>::kStage);
    emit_fix_plan_kernel<<<1, 1024, (size_t)pcap, g->stream>>>(o, cap_rows, poff, moves, fix);
    emit_fix_save_kernel<<<g->sm_count * 2, 256, 0, g->stream>>>(o, poff, fix, static_cast<uint2*>(g->emit_tmp.p));
    emit_fix_move_kernel<<<g->sm_count * 2, 256, 0, g->stream>>>(o, moves, fix);
    emit_fix_pack_kernel<<<g->sm_count * 2, 256, 0, g->stream>>>(o, static_cast<const uint2*>(g->emit_tmp.p), fix);
  }
  g->launches += 4;
  ZB_CUDA(cudaGetLastError());
  ZB_CUDA(cudaMemcpyAsync(&g->h_misc->emit_fix, fix, sizeof(EmitFix), cudaMemcpyDeviceToHost, g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  *h_fix = g->h_misc->emit_fix;
  return ZB_OK;
}

// the non-empty cell count of a rebuild that returned without a host sync
int collect_info(zb_grid* g) {
  if (!g->info_pending) return ZB_OK;
  ZB_CUDA(cudaEventSynchronize(g->info_event));
  g->info_pending = false;
  g->n_cells_nonempty = g->h_misc->nonempty;
  return ZB_OK;
}

int check_built(zb_grid* g) {
  if (!g) return ZB_ERR_BAD_ARG;
  if (g->slab.pending) ZB_TRY(slab_collect(g, nullptr));  // a native slab step reports its verdict lazily
  if (!g->built) return fail(g, ZB_ERR_NOT_BUILT, "grid has not been (successfully) built");
  return ZB_OK;
}

int enter(zb_grid* g) {
  if (!g) return ZB_ERR_BAD_ARG;
  ZB_CUDA(cudaSetDevice(g->device));
  return ZB_OK;
}

void free_buf(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI

extern "C" {

int zb_abi_version(void) { return ZB_ABI_VERSION; }

#ifndef ZB_BUILD_ID
#define ZB_BUILD_ID "unknown"
#endif
const char* zb_build_id(void) { return ZB_BUILD_ID; }

int zb_grid_create(int dtype, int ndim, int device, zb_grid** out) {
  if (!out) return ZB_ERR_BAD_ARG;
  *out = nullptr;
  if ((dtype != ZB_F32 && dtype != ZB_F64) || (ndim != 2 && ndim != 3)) return ZB_ERR_BAD_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return ZB_ERR_CUDA;  // no CPU fallback
  }
  zb_grid* g = new (std::nothrow) zb_grid();
  if (!g) return ZB_ERR_CUDA;
  g->device = device;
  g->dtype = dtype;
  g->ndim = ndim;
  if (const char* e = getenv("ZB_PREFILTER")) g->tune.prefilter = (uint32_t)atoi(e);
  if (const char* e = getenv("ZB_SPLIT")) g->tune.split = atoi(e) != 0;
  if (const char* e = getenv("ZB_SPARSE")) g->tune.sparse = atoi(e);
  if (const char* e = getenv("ZB_SLAB_SPEC")) g->tune.slab_spec = atoi(e) != 0;
  if (const char* e = getenv("ZB_ROW_TILES")) g->tune.row_tiles = atoi(e) != 0;
  if (const char* e = getenv("ZB_P2P")) g->tune.p2p = atoi(e) != 0;
  if (const char* e = getenv("ZB_P2P_HALO_ROWS")) {
    const long v = atol(e);
    if (v >= 64 && v <= (1 << 24)) g->tune.p2p_halo_rows = (uint32_t)v;
  }
  if (const char* e = getenv("ZB_STAGE_RECS")) {
    const long v = atol(e);
    if (v >= 64 && v <= 6144) g->tune.stage_recs = (uint32_t)v;
  }
  if (const char* e = getenv("ZB_TILE_CELLS")) {
    const long v = atol(e);
    if (v >= 1) g->tune.tile_cells = (uint32_t)v;
  }
  auto bail = [&](int rc) {
    zb_grid_destroy(g);
    return rc;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail(ZB_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(ZB_ERR_CUDA);
  if (prop.major < 10) return bail(ZB_ERR_CUDA);  // sm_100a code only
  g->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(ZB_ERR_CUDA);
  g->own_stream = true;
  if (cudaMalloc(reinterpret_cast<void**>(&g->misc), sizeof(Misc)) != cudaSuccess) return bail(ZB_ERR_CUDA);
  if (cudaMemset(g->misc, 0, sizeof(Misc)) != cudaSuccess) return bail(ZB_ERR_CUDA);
  if (cudaMallocHost(reinterpret_cast<void**>(&g->h_misc), sizeof(Misc)) != cudaSuccess) return bail(ZB_ERR_CUDA);
  memset(g->h_misc, 0, sizeof(Misc));
  *out = g;
  return ZB_OK;
}

void zb_grid_destroy(zb_grid* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  DevBuf* bufs[] = {&g->in,        &g->labels_in,   &g->table,        &g->sorted,       &g->scan_state,
                    &g->partials,  &g->keys_old,    &g->keys_new,     &g->tile_counts,  &g->tile_offsets, &g->emit_ctl, &g->emit_tab, &g->emit_spill, &g->emit_tmp,
                    &g->block_energy, &g->block_totals, &g->out_stage, &g->tile_list, &g->pf_list, &g->skeys[0], &g->skeys[1], &g->sidx[0], &g->sidx[1], &g->shist, &g->sflags,
                    &g->ukeys, &g->ubegin, &g->halo_send, &g->halo_recv, &g->halo_labels,
                    &g->red};
  for (DevBuf* b : bufs) free_buf(*b);
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  for (auto& sl : g->pre) {
    if (sl.copied) cudaEventDestroy(sl.copied);
    if (sl.released) cudaEventDestroy(sl.released);
    free_buf(sl.buf);
  }
  if (g->info_event) cudaEventDestroy(g->info_event);
  if (g->bbox_event) cudaEventDestroy(g->bbox_event);
  for (void* o : g->p2p.opened)
    if (o) cudaIpcCloseMemHandle(o);
  if (g->p2p.local) cudaFree(g->p2p.local);
  if (g->nccl.comm && g->nccl.CommDestroy) g->nccl.CommDestroy(g->nccl.comm);
  if (g->nccl.dl) dlclose(g->nccl.dl);
  if (g->h_red) cudaFreeHost(g->h_red);
  for (auto& sp : g->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (auto e : g->ev_free) cudaEventDestroy(e);
  if (g->misc) cudaFree(g->misc);
  if (g->h_misc) cudaFreeHost(g->h_misc);
  if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
  cudaGetLastError();
  delete g;
}

int zb_grid_set_stream(zb_grid* g, void* cuda_stream) {
  ZB_TRY(enter(g));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
  g->stream = static_cast<cudaStream_t>(cuda_stream);
  g->own_stream = false;
  return ZB_OK;
}

int zb_grid_set_stable(zb_grid* g, int enable) {
  if (!g) return ZB_ERR_BAD_ARG;
  g->stable = enable != 0;
  return ZB_OK;
}

int zb_grid_track_key_changes(zb_grid* g, int enable) {
  if (!g) return ZB_ERR_BAD_ARG;
  g->track_keys = enable != 0;
  if (!g->track_keys) g->keys_changed = -1;
  return ZB_OK;
}

const char* zb_last_error(const zb_grid* g) { return g ? g->err.c_str() : "null handle"; }

int zb_grid_rebuild(zb_grid* g, const void* xyz, uint64_t n, const double* cutoff_or_null) {
  ZB_TRY(enter(g));
  return g->dtype == ZB_F32 ? rebuild_impl<float>(g, xyz, n, nullptr, cutoff_or_null, nullptr, nullptr, 0, 0, false)
                            : rebuild_impl<double>(g, xyz, n, nullptr, cutoff_or_null, nullptr, nullptr, 0, 0, false);
}

int zb_grid_prefetch(zb_grid* g, const void* xyz_host, uint64_t n) {
  ZB_TRY(enter(g));
  if (n == 0 || !xyz_host) return fail(g, ZB_ERR_BAD_ARG, "nothing to prefetch");
  if (n > 2147483647ull) return fail(g, ZB_ERR_TOO_MANY, "n = %llu exceeds i32::MAX", (unsigned long long)n);
  if (is_device_ptr(xyz_host)) return ZB_OK;  // already resident
  if (!g->copy_stream) {
    ZB_CUDA(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    for (auto& sl : g->pre) {
      ZB_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
      ZB_CUDA(cudaEventCreateWithFlags(&sl.released, cudaEventDisableTiming));
    }
  }
  // a slot that already holds this array is refreshed; else take a slot that holds nothing pending
  int k = -1;
  for (int j = 0; j < 2; ++j)
    if (g->pre[j].host == xyz_host && g->pre[j].n == n) k = j;
  for (int j = 0; j < 2 && k < 0; ++j)
    if (g->pre[j].host == nullptr && j != g->pre_in_use) k = j;
  if (k < 0) return fail(g, ZB_ERR_CAPACITY, "both prefetch slots hold frames that no rebuild has consumed yet");
  auto& sl = g->pre[k];
  const size_t bytes = (size_t)n * g->ndim * elem_size(g);
  if (bytes > sl.buf.cap || !sl.buf.p) {
    ZB_CUDA(cudaStreamSynchronize(g->copy_stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    if (sl.buf.p) ZB_CUDA(cudaFree(sl.buf.p));
    sl.buf.p = nullptr;
    sl.buf.cap = 0;
    ZB_CUDA(cudaMalloc(&sl.buf.p, bytes));
    sl.buf.cap = bytes;
  }
  ZB_CUDA(cudaStreamWaitEvent(g->copy_stream, sl.released, 0));  // the build that last read this slot is done
  ZB_CUDA(cudaMemcpyAsync(sl.buf.p, xyz_host, bytes, cudaMemcpyHostToDevice, g->copy_stream));
  ZB_CUDA(cudaEventRecord(sl.copied, g->copy_stream));
  sl.host = xyz_host;
  sl.n = n;
  sl.age = 0;
  return ZB_OK;
}

int zb_grid_prefetch_wait(zb_grid* g) {
  ZB_TRY(enter(g));
  for (auto& sl : g->pre)
    if (sl.host && sl.copied) ZB_CUDA(cudaStreamWaitEvent(g->stream, sl.copied, 0));
  return ZB_OK;
}

int zb_grid_rebuild_sharded(zb_grid* g, const void* xyz, uint64_t n, const uint32_t* labels_or_null,
                            const double* cutoff_or_null, const double* inf, const double* sup, int64_t z_begin,
                            int64_t z_end) {
  ZB_TRY(enter(g));
  if (!inf || !sup) return fail(g, ZB_ERR_BAD_ARG, "inf / sup must be given");
  return g->dtype == ZB_F32
             ? rebuild_impl<float>(g, xyz, n, labels_or_null, cutoff_or_null, inf, sup, z_begin, z_end, true)
             : rebuild_impl<double>(g, xyz, n, labels_or_null, cutoff_or_null, inf, sup, z_begin, z_end, true);
}

int zb_aabb(zb_grid* g, const void* xyz, uint64_t n, double* out6) {
  ZB_TRY(enter(g));
  if (!out6) return fail(g, ZB_ERR_BAD_ARG, "out6 is NULL");
  const bool dev_out = is_device_ptr(out6);
  if (n == 0) {
    if (dev_out) ZB_CUDA(cudaMemsetAsync(out6, 0, 6 * sizeof(double), g->stream));
    else for (int d = 0; d < 6; ++d) out6[d] = 0.0;
    return ZB_OK;
  }
  if (!xyz) return fail(g, ZB_ERR_BAD_ARG, "xyz is NULL");
  const void* dev = nullptr;
  ZB_TRY(stage_input(g, xyz, n, &dev));
  if (g->dtype == ZB_F32) ZB_TRY(launch_bbox<float>(g, static_cast<const float*>(dev), n));
  else ZB_TRY(launch_bbox<double>(g, static_cast<const double*>(dev), n));
  if (dev_out) {
    // stays on the device (asynchronous): the sharded host all-reduces it in place
    if (g->dtype == ZB_F32) widen6_kernel<float><<<1, 32, 0, g->stream>>>(reinterpret_cast<const float*>(g->misc->out6), out6, g->ndim, 0);
    else widen6_kernel<double><<<1, 32, 0, g->stream>>>(reinterpret_cast<const double*>(g->misc->out6), out6, g->ndim, 0);
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    return ZB_OK;
  }
  double o[6];
  if (g->dtype == ZB_F32) ZB_TRY(fetch_bbox<float>(g, o, false));
  else ZB_TRY(fetch_bbox<double>(g, o, false));
  for (int d = 0; d < 6; ++d) out6[d] = 0.0;
  for (int d = 0; d < g->ndim; ++d) {
    out6[d] = o[d];
    out6[3 + d] = o[3 + d];
  }
  return ZB_OK;
}

int zb_layer_of(zb_grid* g, const void* xyz, uint64_t n, double inf_axis, double cutoff, int axis, int32_t* out) {
  ZB_TRY(enter(g));
  if (axis < 0 || axis >= g->ndim) return fail(g, ZB_ERR_BAD_ARG, "bad axis %d", axis);
  if (n > 2147483647ull) return fail(g, ZB_ERR_TOO_MANY, "n = %llu exceeds i32::MAX", (unsigned long long)n);
  if (n == 0) return ZB_OK;
  if (!xyz || !out) return fail(g, ZB_ERR_BAD_ARG, "xyz / out is NULL");
  const void* dev = nullptr;
  ZB_TRY(stage_input(g, xyz, n, &dev));
  const bool dev_out = is_device_ptr(out);
  int32_t* dst = out;
  if (!dev_out) {
    ZB_TRY(reserve(g, g->out_stage, n * 4));
    dst = static_cast<int32_t*>(g->out_stage.p);
  }
  const uint32_t blocks = (uint32_t)((n + 255) / 256);
  if (g->dtype == ZB_F32)
    layer_kernel<float><<<blocks, 256, 0, g->stream>>>(static_cast<const float*>(dev), (uint32_t)n, g->ndim, axis,
                                                        (float)inf_axis, (float)cutoff, dst);
  else
    layer_kernel<double><<<blocks, 256, 0, g->stream>>>(static_cast<const double*>(dev), (uint32_t)n, g->ndim, axis,
                                                         inf_axis, cutoff, dst);
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  if (!dev_out) ZB_TRY(deliver(g, out, dst, n * 4));
  return ZB_OK;
}

int zb_slab_top_layer(zb_grid* g, const void* xyz, uint64_t n, double inf_axis, double cutoff, int64_t z_begin,
                      int64_t z_end, uint32_t label_offset, void* halo_rows, uint64_t cap_rows, uint64_t* n_top,
                      int* out_of_slab) {
  ZB_TRY(enter(g));
  const bool async = n_top == nullptr && out_of_slab == nullptr;
  if (!async && (!n_top || !out_of_slab)) return fail(g, ZB_ERR_BAD_ARG, "n_top and out_of_slab go together");
  if (!async) {
    *n_top = 0;
    *out_of_slab = 0;
  }
  if (n > 2147483647ull) return fail(g, ZB_ERR_TOO_MANY, "n = %llu exceeds i32::MAX", (unsigned long long)n);
  if (!halo_rows || !is_device_ptr(halo_rows)) return fail(g, ZB_ERR_BAD_ARG, "halo_rows must be device memory");
  if (n > 0 && !xyz) return fail(g, ZB_ERR_BAD_ARG, "xyz is NULL");
  const void* dev = nullptr;
  ZB_TRY(stage_input(g, xyz, n, &dev));
  // slab_count / slab_flag live apart from the rebuild's counters: the flag is read by the NEXT sharded rebuild
  ZB_CUDA(cudaMemsetAsync(&g->misc->slab_count, 0, 2 * sizeof(uint32_t), g->stream));
  const uint32_t blocks = (uint32_t)((n + 255) / 256);
  const uint32_t cap = (uint32_t)std::min<uint64_t>(cap_rows, 0xfffffff0ull);
  uint32_t* cnt = &g->misc->slab_count;
  int* bad = reinterpret_cast<int*>(&g->misc->slab_flag);
  if (n > 0) {
    if (g->dtype == ZB_F32) {
      auto* x = static_cast<const float*>(dev);
      auto* o = static_cast<float*>(halo_rows);
      if (g->ndim == 3)
        slab_top_kernel<float, 3><<<blocks, 256, 0, g->stream>>>(x, (uint32_t)n, (float)inf_axis, (float)cutoff, (int)z_begin,
                                                                 (int)z_end, label_offset, o, cap, cnt, bad);
      else
        slab_top_kernel<float, 2><<<blocks, 256, 0, g->stream>>>(x, (uint32_t)n, (float)inf_axis, (float)cutoff, (int)z_begin,
                                                                 (int)z_end, label_offset, o, cap, cnt, bad);
    } else {
      auto* x = static_cast<const double*>(dev);
      auto* o = static_cast<double*>(halo_rows);
      if (g->ndim == 3)
        slab_top_kernel<double, 3><<<blocks, 256, 0, g->stream>>>(x, (uint32_t)n, inf_axis, cutoff, (int)z_begin, (int)z_end,
                                                                  label_offset, o, cap, cnt, bad);
      else
        slab_top_kernel<double, 2><<<blocks, 256, 0, g->stream>>>(x, (uint32_t)n, inf_axis, cutoff, (int)z_begin, (int)z_end,
                                                                  label_offset, o, cap, cnt, bad);
    }
    g->launches++;
  }
  // the block header (row 0) carries the row count to the receiver
  if (g->dtype == ZB_F32) halo_header_kernel<float><<<1, 1, 0, g->stream>>>(cnt, cap, static_cast<float*>(halo_rows));
  else halo_header_kernel<double><<<1, 1, 0, g->stream>>>(cnt, cap, static_cast<double*>(halo_rows));
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  g->slab_check_pending = true;
  if (async) return ZB_OK;
  ZB_CUDA(cudaMemcpyAsync(&g->h_misc->slab_count, &g->misc->slab_count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  g->slab_check_pending = false;
  *n_top = g->h_misc->slab_count;
  *out_of_slab = (int)(g->h_misc->slab_flag & 1u);
  if (*n_top > cap_rows)
    return fail(g, ZB_ERR_CAPACITY, "top layer holds %llu particles, halo buffer %llu rows", (unsigned long long)*n_top,
                (unsigned long long)cap_rows);
  return ZB_OK;
}

int zb_grid_info(zb_grid* g, zb_info* out) {
  if (!g || !out) return ZB_ERR_BAD_ARG;
  if (g->slab.pending) {
    ZB_TRY(enter(g));
    ZB_TRY(slab_collect(g, nullptr));
  }
  if (g->info_pending) {
    ZB_TRY(enter(g));
    ZB_TRY(collect_info(g));
  }
  memset(out, 0, sizeof *out);
  for (int d = 0; d < 3; ++d) {
    out->inf[d] = g->inf[d];
    out->sup[d] = g->sup[d];
    out->shape[d] = d < g->ndim ? g->shape[d] : 0;
  }
  // GridInfo strides (util.rs:200-212): wrapping i32 products of (shape + 4)
  uint32_t s = 1;
  for (int d = 0; d < 3; ++d) {
    out->strides[d] = d < g->ndim ? (int32_t)s : 0;
    s *= (uint32_t)(g->shape[d] + 4);
  }
  out->cutoff = g->cutoff;
  out->n = g->n;
  out->n_cells = g->n_cells_nonempty;  // (collected above when the last rebuild returned without a sync)
  out->ndim = g->ndim;
  out->dtype = g->dtype;
  out->keys_changed = g->keys_changed;
  return ZB_OK;
}

int zb_grid_keys(zb_grid* g, int32_t* out) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  if (g->sharded) return fail(g, ZB_ERR_BAD_ARG, "zb_grid_keys is not defined for sharded grids");
  if (g->n == 0) return ZB_OK;
  if (!out) return fail(g, ZB_ERR_BAD_ARG, "out is NULL");
  const bool dev_out = is_device_ptr(out);
  int32_t* dst = out;
  if (!dev_out) {
    ZB_TRY(reserve(g, g->out_stage, g->n * 4));
    dst = static_cast<int32_t*>(g->out_stage.p);
  }
  const uint32_t blocks = (uint32_t)((g->n + 255) / 256);
  if (g->dtype == ZB_F32)
    keys_kernel<float><<<blocks, 256, 0, g->stream>>>(static_cast<const Rec<float>*>(g->sorted.p), (uint32_t)g->n,
                                                       make_params<float>(g), dst);
  else
    keys_kernel<double><<<blocks, 256, 0, g->stream>>>(static_cast<const Rec<double>*>(g->sorted.p), (uint32_t)g->n,
                                                        make_params<double>(g), dst);
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  if (!dev_out) ZB_TRY(deliver(g, out, dst, g->n * 4));
  return ZB_OK;
}

int zb_grid_neighbor_indices(zb_grid* g, int32_t* out, int32_t* count) {
  if (!g || !out || !count) return ZB_ERR_BAD_ARG;
  // lexicographic product {-1,0,1}^N, last axis fastest, centre removed (flatindex.rs:55-65)
  uint32_t strides[3];
  uint32_t s = 1;
  for (int d = 0; d < 3; ++d) {
    strides[d] = s;
    s *= (uint32_t)(g->shape[d] + 4);
  }
  int k = 0;
  const int nd = g->ndim;
  int total = 1;
  for (int d = 0; d < nd; ++d) total *= 3;
  for (int t = 0; t < total; ++t) {
    int rem = t, off[3] = {0, 0, 0};
    for (int d = nd - 1; d >= 0; --d) {
      off[d] = rem % 3 - 1;
      rem /= 3;
    }
    bool centre = true;
    for (int d = 0; d < nd; ++d) centre = centre && off[d] == 0;
    if (centre) continue;
    uint32_t key = 0;
    for (int d = 0; d < nd; ++d) key += (uint32_t)off[d] * strides[d];
    out[k++] = (int32_t)key;
  }
  *count = k;
  return ZB_OK;
}

int zb_grid_cells(zb_grid* g, int32_t* keys, uint32_t* begin, uint32_t* count, uint64_t cap, uint64_t* n_out) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  ZB_TRY(collect_info(g));
  if (!n_out) return fail(g, ZB_ERR_BAD_ARG, "n_out is NULL");
  *n_out = g->n_cells_nonempty;
  if (cap < g->n_cells_nonempty) return fail(g, ZB_ERR_CAPACITY, "need room for %llu cells", (unsigned long long)*n_out);
  if (g->n_cells_nonempty == 0) return ZB_OK;
  if (g->sparse) {  // the compact cells are already the answer
    const uint64_t nu = g->nuniq;
    ZB_TRY(reserve(g, g->out_stage, nu * 12 + 64));
    int32_t* dk = static_cast<int32_t*>(g->out_stage.p);
    uint32_t* db = reinterpret_cast<uint32_t*>(dk + nu);
    uint32_t* dc = db + nu;
    cells_sparse_kernel<<<(uint32_t)((nu + 255) / 256), 256, 0, g->stream>>>(
        static_cast<const unsigned long long*>(g->ukeys.p), static_cast<const uint32_t*>(g->ubegin.p), (uint32_t)nu, g->shape[0],
        g->shape[1], dk, db, dc);
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    if (keys) ZB_TRY(deliver(g, keys, dk, nu * 4));
    if (begin) ZB_TRY(deliver(g, begin, db, nu * 4));
    if (count) ZB_TRY(deliver(g, count, dc, nu * 4));
    return ZB_OK;
  }
  // flags -> exclusive scan (K3's scan kernel) -> ordered compaction, then one copy per array
  const uint64_t nc = g->n_cells_nonempty;
  const uint32_t ntile = (g->ncells + kScanTile - 1) / kScanTile;
  const size_t pos_elems = (size_t)ntile * kScanTile;
  ZB_TRY(reserve(g, g->out_stage, pos_elems * 4 + nc * 12 + 64));
  ZB_TRY(reserve(g, g->scan_state, (size_t)ntile * sizeof(unsigned long long)));
  uint32_t* dpos = static_cast<uint32_t*>(g->out_stage.p);
  int32_t* dk = reinterpret_cast<int32_t*>(dpos + pos_elems);
  uint32_t* db = reinterpret_cast<uint32_t*>(dk + nc);
  uint32_t* dc = db + nc;
  ZB_CUDA(cudaMemsetAsync(g->scan_state.p, 0, (size_t)ntile * sizeof(unsigned long long), g->stream));
  ZB_CUDA(cudaMemsetAsync(&g->misc->tile_counter, 0, 2 * sizeof(uint32_t), g->stream));
  const uint32_t blocks = (uint32_t)((g->ncells + 255) / 256);
  cells_flag_kernel<<<blocks, 256, 0, g->stream>>>(csr_ptr(g), g->ncells, dpos);
  scan_kernel<<<ntile, kScanThreads, 0, g->stream>>>(dpos, g->ncells, static_cast<unsigned long long*>(g->scan_state.p),
                                                     &g->misc->tile_counter, &g->misc->nonempty);
  cells_compact_kernel<<<blocks, 256, 0, g->stream>>>(csr_ptr(g), dpos, g->ncells, g->shape[0], g->shape[1], g->wlo[0],
                                                     g->wlo[1], g->wlo[2], g->wshape[0], g->wshape[1], dk, db, dc);
  g->launches += 3;
  ZB_CUDA(cudaGetLastError());
  if (keys) ZB_TRY(deliver(g, keys, dk, nc * 4));
  if (begin) ZB_TRY(deliver(g, begin, db, nc * 4));
  if (count) ZB_TRY(deliver(g, count, dc, nc * 4));
  return ZB_OK;
}

int zb_grid_cell_storage(zb_grid* g, uint32_t* labels, void* xyz) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  const uint64_t n = g->n;
  if (n == 0) return ZB_OK;
  const size_t es = elem_size(g);
  const bool ldev = labels && is_device_ptr(labels), xdev = xyz && is_device_ptr(xyz);
  const size_t lbytes = n * 4, xbytes = n * g->ndim * es;
  ZB_TRY(reserve(g, g->out_stage, lbytes + xbytes + 64));
  uint32_t* dl = labels ? (ldev ? labels : static_cast<uint32_t*>(g->out_stage.p)) : nullptr;
  void* dx = xyz ? (xdev ? xyz : static_cast<void*>(static_cast<char*>(g->out_stage.p) + ((lbytes + 15) / 16) * 16))
                 : nullptr;
  const uint32_t blocks = (uint32_t)((n + 255) / 256);
  if (g->dtype == ZB_F32) {
    auto* s = static_cast<const Rec<float>*>(g->sorted.p);
    if (g->ndim == 3) unpack_kernel<float, 3><<<blocks, 256, 0, g->stream>>>(s, (uint32_t)n, dl, static_cast<float*>(dx));
    else unpack_kernel<float, 2><<<blocks, 256, 0, g->stream>>>(s, (uint32_t)n, dl, static_cast<float*>(dx));
  } else {
    auto* s = static_cast<const Rec<double>*>(g->sorted.p);
    if (g->ndim == 3) unpack_kernel<double, 3><<<blocks, 256, 0, g->stream>>>(s, (uint32_t)n, dl, static_cast<double*>(dx));
    else unpack_kernel<double, 2><<<blocks, 256, 0, g->stream>>>(s, (uint32_t)n, dl, static_cast<double*>(dx));
  }
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  if (labels && !ldev) ZB_TRY(deliver(g, labels, dl, lbytes));
  if (xyz && !xdev) ZB_TRY(deliver(g, xyz, dx, xbytes));
  return ZB_OK;
}

int zb_grid_pair_count(zb_grid* g, int cmp, double filter_cutoff, uint64_t* out) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  if (!out) return fail(g, ZB_ERR_BAD_ARG, "out is NULL");
  if (cmp < 0 || cmp > 2) return fail(g, ZB_ERR_BAD_ARG, "bad cmp %d", cmp);
  if (g->dtype == ZB_F32) ZB_TRY(pair_count_impl<float>(g, cmp, filter_cutoff));
  else ZB_TRY(pair_count_impl<double>(g, cmp, filter_cutoff));
  return deliver(g, out, &g->misc->pair_total, 8);
}

int zb_grid_pairs(zb_grid* g, int cmp, double filter_cutoff, uint32_t* ij, uint64_t cap, uint64_t* n_out) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  if (!n_out) return fail(g, ZB_ERR_BAD_ARG, "n_out is NULL");
  if (cmp < 0 || cmp > 2) return fail(g, ZB_ERR_BAD_ARG, "bad cmp %d", cmp);
  // the length of the list, if an earlier call on the same build and filter already counted it
  bool known = g->pairs_sized && g->pairs_sized_build == g->build_id && g->pairs_sized_cmp == cmp &&
               g->pairs_sized_fc == filter_cutoff;
  uint64_t total = known ? g->pairs_sized_total : 0;
  auto size_it = [&]() -> int {
    if (g->dtype == ZB_F32) ZB_TRY(pair_count_impl<float>(g, cmp, filter_cutoff));
    else ZB_TRY(pair_count_impl<double>(g, cmp, filter_cutoff));
    ZB_CUDA(cudaMemcpyAsync(&g->h_misc->pair_total, &g->misc->pair_total, 8, cudaMemcpyDeviceToHost, g->stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    total = g->h_misc->pair_total;
    g->pairs_sized = known = true;
    g->pairs_sized_build = g->build_id;
    g->pairs_sized_cmp = cmp;
    g->pairs_sized_fc = filter_cutoff;
    g->pairs_sized_total = total;
    return ZB_OK;
  };
  const bool dev_out = ij && is_device_ptr(ij);
  // a sizing call (no buffer), or a host destination: count first (the staging buffer gets the exact size, and
  // the count pass is small change beside the copy to the host).  A device destination with room is written in
  // ONE pass, no count in front of it.
  if (!known && (!ij || cap == 0 || !dev_out)) ZB_TRY(size_it());
  if (known) {
    *n_out = total;
    if (total > cap || (total && !ij))
      return fail(g, ZB_ERR_CAPACITY, "pair list needs %llu rows, capacity is %llu", (unsigned long long)total,
                  (unsigned long long)cap);
    if (total == 0) return ZB_OK;
  }
  const uint64_t room = known ? total : cap;
  uint2* dst = reinterpret_cast<uint2*>(ij);
  if (!dev_out) {
    ZB_TRY(reserve(g, g->out_stage, room * 8));
    dst = static_cast<uint2*>(g->out_stage.p);
  }
  EmitFix fix;
  if (g->dtype == ZB_F32) ZB_TRY(emit_impl<float>(g, cmp, filter_cutoff, dst, room, &fix));
  else ZB_TRY(emit_impl<double>(g, cmp, filter_cutoff, dst, room, &fix));
  *n_out = fix.total;
  if (fix.ok == 2 || (known && fix.total != total))
    return fail(g, ZB_ERR_CUDA, "pair list: slot bookkeeping failed (%llu rows, %llu expected, %u partial slots)",
                (unsigned long long)fix.total, (unsigned long long)total, fix.npartial);
  if (fix.ok != 1)
    return fail(g, ZB_ERR_CAPACITY, "pair list needs %llu rows, capacity is %llu", (unsigned long long)fix.total,
                (unsigned long long)cap);
  if (!dev_out && fix.total) ZB_TRY(deliver(g, ij, dst, fix.total * 8));
  return ZB_OK;
}

int zb_grid_lj_energy(zb_grid* g, int cmp, double filter_cutoff, double* energy, uint64_t* n_pairs) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  if (!energy) return fail(g, ZB_ERR_BAD_ARG, "energy is NULL");
  if (cmp < 1 || cmp > 2) return fail(g, ZB_ERR_BAD_ARG, "lj energy needs a distance filter (cmp LT or LE)");
  if (g->dtype == ZB_F32) ZB_TRY(lj_impl<float>(g, cmp, filter_cutoff));
  else ZB_TRY(lj_impl<double>(g, cmp, filter_cutoff));
  const bool edev = is_device_ptr(energy);
  const bool pdev = n_pairs && is_device_ptr(n_pairs);
  if (edev) ZB_CUDA(cudaMemcpyAsync(energy, &g->misc->energy, 8, cudaMemcpyDeviceToDevice, g->stream));
  if (pdev) ZB_CUDA(cudaMemcpyAsync(n_pairs, &g->misc->pair_total, 8, cudaMemcpyDeviceToDevice, g->stream));
  if (!edev || (n_pairs && !pdev)) {
    ZB_CUDA(cudaMemcpyAsync(&g->h_misc->energy, &g->misc->energy, 16, cudaMemcpyDeviceToHost, g->stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    if (!edev) *energy = g->h_misc->energy;
    if (n_pairs && !pdev) *n_pairs = g->h_misc->pair_total;
  }
  return ZB_OK;
}

int zb_grid_query_neighbors(zb_grid* g, const void* queries, uint64_t nq, int cmp, double filter_cutoff,
                            uint64_t* offsets, uint8_t* valid, uint32_t* labels, uint64_t cap, uint64_t* n_out) {
  ZB_TRY(enter(g));
  ZB_TRY(check_built(g));
  if (!offsets || !n_out) return fail(g, ZB_ERR_BAD_ARG, "offsets / n_out is NULL");
  if (cmp < 0 || cmp > 2) return fail(g, ZB_ERR_BAD_ARG, "bad cmp %d", cmp);
  if (g->sharded) return fail(g, ZB_ERR_BAD_ARG, "point queries are not defined for sharded grids");
  if (nq > 0xffffffffull) return fail(g, ZB_ERR_TOO_MANY, "too many queries");
  *n_out = 0;
  if (nq == 0) {
    if (is_device_ptr(offsets)) ZB_CUDA(cudaMemsetAsync(offsets, 0, 8, g->stream));
    else offsets[0] = 0;
    return ZB_OK;
  }
  if (!queries) return fail(g, ZB_ERR_BAD_ARG, "queries is NULL");
  const size_t es = elem_size(g);
  const size_t qbytes = nq * g->ndim * es;
  // device scratch: [queries (if host)] [counts u64 nq+1] [valid u8 nq]
  const bool qdev = is_device_ptr(queries);
  ZB_TRY(reserve(g, g->in, qdev ? 256 : qbytes));  // reuse the input stage (grid data lives in `sorted`)
  const void* dq = queries;
  if (!qdev) {
    ZB_CUDA(cudaMemcpyAsync(g->in.p, queries, qbytes, cudaMemcpyHostToDevice, g->stream));
    dq = g->in.p;
  }
  ZB_TRY(reserve(g, g->tile_offsets, (nq + 1) * 8 + nq + 64));
  unsigned long long* doff = static_cast<unsigned long long*>(g->tile_offsets.p);
  uint8_t* dvalid = reinterpret_cast<uint8_t*>(doff + nq + 1);
  const uint32_t wblocks = (uint32_t)((nq * 32 + 255) / 256);
  auto run = [&](auto tag, bool emit, uint32_t* dl) -> int {
    using T = decltype(tag);
    QueryParams<T> qp;
    qp.g = make_params<T>(g);
    qp.sorted = static_cast<const Rec<T>*>(g->sorted.p);
    qp.csr = csr_ptr(g);
    const T c = (T)filter_cutoff;
    qp.c2 = c * c;
    qp.cmp = cmp;
    qp.ukeys = g->sparse ? static_cast<const unsigned long long*>(g->ukeys.p) : nullptr;
    qp.nuniq = g->nuniq;
    const T* q = static_cast<const T*>(dq);
    if (g->sparse) {
      if (!emit) query_kernel<T, false, true><<<wblocks, 256, 0, g->stream>>>(qp, q, (uint32_t)nq, doff, dvalid, nullptr);
      else query_kernel<T, true, true><<<wblocks, 256, 0, g->stream>>>(qp, q, (uint32_t)nq, doff, dvalid, dl);
    } else {
      if (!emit) query_kernel<T, false, false><<<wblocks, 256, 0, g->stream>>>(qp, q, (uint32_t)nq, doff, dvalid, nullptr);
      else query_kernel<T, true, false><<<wblocks, 256, 0, g->stream>>>(qp, q, (uint32_t)nq, doff, dvalid, dl);
    }
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    return ZB_OK;
  };
  // pass 1: counts -> exclusive scan -> offsets
  if (g->dtype == ZB_F32) ZB_TRY(run(float(), false, nullptr));
  else ZB_TRY(run(double(), false, nullptr));
  ZB_TRY(reserve(g, g->tile_counts, (nq + 1) * 8));
  ZB_CUDA(cudaMemcpyAsync(g->tile_counts.p, doff, nq * 8, cudaMemcpyDeviceToDevice, g->stream));
  tile_offsets_kernel<<<1, 1024, 0, g->stream>>>(static_cast<const unsigned long long*>(g->tile_counts.p), (uint32_t)nq,
                                                 nullptr, doff);
  g->launches++;
  ZB_CUDA(cudaGetLastError());
  ZB_CUDA(cudaMemcpyAsync(&g->h_misc->pair_total, doff + nq, 8, cudaMemcpyDeviceToHost, g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  const uint64_t total = g->h_misc->pair_total;
  *n_out = total;
  ZB_TRY(deliver(g, offsets, doff, (nq + 1) * 8));
  if (valid) ZB_TRY(deliver(g, valid, dvalid, nq));
  if (total > cap || (total && !labels))
    return fail(g, ZB_ERR_CAPACITY, "neighbour list needs %llu entries, capacity is %llu", (unsigned long long)total,
                (unsigned long long)cap);
  if (total == 0) return ZB_OK;
  const bool ldev = is_device_ptr(labels);
  uint32_t* dl = labels;
  if (!ldev) {
    ZB_TRY(reserve(g, g->out_stage, total * 4));
    dl = static_cast<uint32_t*>(g->out_stage.p);
  }
  if (g->dtype == ZB_F32) ZB_TRY(run(float(), true, dl));
  else ZB_TRY(run(double(), true, dl));
  if (!ldev) ZB_TRY(deliver(g, labels, dl, total * 4));
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------
// native multi-GPU step: the same sequence zelll_b200/sharded.py drives through torch.distributed,
// issued from here on the handle's stream with NCCL directly (3 host round trips per step)

#define ZB_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (expr);                                                                     \
    if (r__ != ncclSuccess)                                                                        \
      return fail(g, ZB_ERR_CUDA, "%s failed: %s", #expr,                                          \
                  g->nccl.GetErrorString ? g->nccl.GetErrorString(r__) : "nccl error");            \
  } while (0)

}  // extern "C"

static int nccl_load(zb_grid* g, const char* path) {
  if (g->nccl.dl) return ZB_OK;
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  void* dl = nullptr;
  for (const char* c : cands) {
    if (!c || !*c) continue;
    dl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (dl) break;
  }
  if (!dl) return fail(g, ZB_ERR_CUDA, "cannot load NCCL: %s", dlerror());
  auto sym = [&](const char* n) { return dlsym(dl, n); };
  auto& N = g->nccl;
  N.dl = dl;
  N.CommInitRank = reinterpret_cast<decltype(N.CommInitRank)>(sym("ncclCommInitRank"));
  N.CommDestroy = reinterpret_cast<decltype(N.CommDestroy)>(sym("ncclCommDestroy"));
  N.AllReduce = reinterpret_cast<decltype(N.AllReduce)>(sym("ncclAllReduce"));
  N.Send = reinterpret_cast<decltype(N.Send)>(sym("ncclSend"));
  N.Recv = reinterpret_cast<decltype(N.Recv)>(sym("ncclRecv"));
  N.AllGather = reinterpret_cast<decltype(N.AllGather)>(sym("ncclAllGather"));
  N.GroupStart = reinterpret_cast<decltype(N.GroupStart)>(sym("ncclGroupStart"));
  N.GroupEnd = reinterpret_cast<decltype(N.GroupEnd)>(sym("ncclGroupEnd"));
  N.GetErrorString = reinterpret_cast<decltype(N.GetErrorString)>(sym("ncclGetErrorString"));
  if (!N.CommInitRank || !N.AllReduce || !N.Send || !N.Recv || !N.GroupStart || !N.GroupEnd)
    return fail(g, ZB_ERR_CUDA, "NCCL library lacks a required symbol");
  return ZB_OK;
}

// Map every rank's mailbox into this process (p2p_kernels.cuh).  Collective over the new communicator; any
// failure on any rank (no peer access, IPC refused, more than kP2pMaxWorld ranks) leaves ALL ranks on NCCL.
static int p2p_setup(zb_grid* g) {
  auto& N = g->nccl;
  auto& P = g->p2p;
  P.ok = false;
  // a second zb_comm_init on the same handle: drop the previous communicator's mappings first
  for (void*& o : P.opened)
    if (o) {
      cudaIpcCloseMemHandle(o);
      o = nullptr;
    }
  if (P.local) {
    cudaFree(P.local);
    P.local = nullptr;
  }
  g->slab = zb_grid::SlabStep{};
  if (N.world < 2) return ZB_OK;
  int ok = (g->tune.p2p && N.world <= kP2pMaxWorld && N.AllGather) ? 1 : 0;
  P.halo_rows = g->tune.p2p_halo_rows;
  P.block_bytes = ((size_t)P.halo_rows + 1) * 4 * sizeof(double);
  const size_t bytes = P2pLayout::total(P.block_bytes);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok) {
    if (cudaMalloc(&P.local, bytes) != cudaSuccess || cudaMemset(P.local, 0, bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, P.local) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  }
  // all-gather the handles (and each rank's verdict so far) through the communicator
  constexpr size_t kRec = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<unsigned char> h_all(kRec * (size_t)N.world, 0), h_mine(kRec, 0);
  memcpy(h_mine.data(), &mine, sizeof mine);
  h_mine[sizeof mine] = (unsigned char)ok;
  DevBuf d_mine, d_all;
  ZB_TRY(reserve(g, d_mine, kRec));
  ZB_TRY(reserve(g, d_all, kRec * (size_t)N.world));
  ZB_CUDA(cudaMemcpyAsync(d_mine.p, h_mine.data(), kRec, cudaMemcpyHostToDevice, g->stream));
  ZB_NCCL(N.AllGather(d_mine.p, d_all.p, kRec, ncclChar, N.comm, g->stream));
  ZB_CUDA(cudaMemcpyAsync(h_all.data(), d_all.p, kRec * (size_t)N.world, cudaMemcpyDeviceToHost, g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  free_buf(d_mine);
  free_buf(d_all);
  for (int r = 0; r < N.world; ++r) ok = ok && h_all[kRec * r + sizeof mine];
  if (ok) {
    for (int r = 0; r < N.world && ok; ++r) {
      if (r == N.rank) {
        P.peers.base[r] = static_cast<unsigned char*>(P.local);
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, &h_all[kRec * r], sizeof h);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      P.opened[r] = ptr;
      P.peers.base[r] = static_cast<unsigned char*>(ptr);
    }
  }
  // second agreement: did every rank manage to open every handle?
  double* red = static_cast<double*>(g->red.p);
  const double v = ok ? 1.0 : 0.0;
  ZB_CUDA(cudaMemcpyAsync(red, &v, sizeof v, cudaMemcpyHostToDevice, g->stream));
  ZB_NCCL(N.AllReduce(red, red, 1, ncclDouble, ncclMin, N.comm, g->stream));
  ZB_CUDA(cudaMemcpyAsync(g->h_red, red, sizeof v, cudaMemcpyDeviceToHost, g->stream));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  P.ok = g->h_red[0] == 1.0;
  P.seq_box = P.seq_energy = P.seq_halo = 0;
  return ZB_OK;
}

extern "C" {

int zb_comm_unique_id(const char* nccl_lib_path, void* out128) {
  if (!out128) return ZB_ERR_BAD_ARG;
  const char* cands[] = {nccl_lib_path, "libnccl.so.2", "libnccl.so"};
  void* dl = nullptr;
  for (const char* c : cands) {
    if (!c || !*c) continue;
    dl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (dl) break;
  }
  if (!dl) return ZB_ERR_CUDA;
  auto get = reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(dlsym(dl, "ncclGetUniqueId"));
  if (!get) return ZB_ERR_CUDA;
  ncclUniqueId id;
  if (get(&id) != ncclSuccess) return ZB_ERR_CUDA;
  memcpy(out128, &id, sizeof id);
  return ZB_OK;
}

int zb_comm_init(zb_grid* g, const char* nccl_lib_path, const void* unique_id128, int world, int rank) {
  ZB_TRY(enter(g));
  if (!unique_id128 || world < 1 || rank < 0 || rank >= world) return fail(g, ZB_ERR_BAD_ARG, "bad communicator arguments");
  ZB_TRY(nccl_load(g, nccl_lib_path));
  if (g->nccl.comm) {
    g->nccl.CommDestroy(g->nccl.comm);
    g->nccl.comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof id);
  ZB_NCCL(g->nccl.CommInitRank(&g->nccl.comm, world, id, rank));
  g->nccl.world = world;
  g->nccl.rank = rank;
  ZB_TRY(reserve(g, g->red, 64 * sizeof(double)));
  if (!g->h_red) ZB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g->h_red), 8 * sizeof(double)));
  return p2p_setup(g);
}

}  // extern "C"

template <class T>
static int slab_step_impl(zb_grid* g, void* buf, uint64_t n_local, uint64_t cap_rows, const double* cutoff,
                          uint32_t label_offset, uint64_t halo_cap, zb_slab_info* out, bool allow_spec) {
  auto& N = g->nccl;
  if (!N.comm) return fail(g, ZB_ERR_NOT_BUILT, "zb_comm_init has not been called");
  if (n_local > 2147483647ull || cap_rows < n_local) return fail(g, ZB_ERR_BAD_ARG, "bad n_local / cap_rows");
  if (!buf || !is_device_ptr(buf)) return fail(g, ZB_ERR_BAD_ARG, "buf must be device memory");
  bool cutoff_changed = false;
  if (cutoff) {
    const T c = (T)*cutoff;
    if (!(c > (T)0) || !std::isfinite((double)c)) return fail(g, ZB_ERR_BAD_ARG, "cutoff must be positive and finite");
    cutoff_changed = g->cutoff != (double)c;
    g->cutoff = (double)c;
  }
  g->built = false;
  g->slab.pending = false;
  T* xyz = static_cast<T*>(buf);
  double* red = static_cast<double*>(g->red.p);
  const int nd = g->ndim;
  // Speculate that the global box is the one of the last completed step (trajectory frames of a closed
  // system, repeated analysis of one frame): no wait for the box all-reduce.  The box is still all-reduced
  // and compared ON THE DEVICE; a mismatch surfaces with the energy all-reduce and the step is repeated.
  const bool spec = allow_spec && g->tune.slab_spec && g->slab.have_box && g->slab.backoff == 0 && !cutoff_changed;
  if (!spec && g->slab.backoff > 0) g->slab.backoff--;

  // slab_count / slab_flag / halo_n live apart from the rebuild's counters
  ZB_CUDA(cudaMemsetAsync(&g->misc->slab_count, 0, 3 * sizeof(uint32_t), g->stream));
  // 1. global box: local K1 -> (-inf, sup) -> all-reduce(max) [-> compared with the assumed box]
  if (g->p2p.ok) {
    // over peer memory: widen + all-reduce + check in ONE launch behind K1
    if (n_local) ZB_TRY(launch_bbox<T>(g, xyz, n_local));
    Box6v expect;
    for (int k = 0; k < 6; ++k) expect.v[k] = g->slab.box[k];
    p2p_box_kernel<T><<<1, 32, 0, g->stream>>>(g->p2p.peers, N.world, N.rank, ++g->p2p.seq_box,
                                               reinterpret_cast<const T*>(g->misc->out6), nd, n_local ? 1 : 0, red, spec ? 1 : 0, expect,
                                               &g->misc->slab_flag);
    g->launches++;
    ZB_CUDA(cudaGetLastError());
  } else {
    if (n_local) {
      ZB_TRY(launch_bbox<T>(g, xyz, n_local));
      widen6_kernel<T><<<1, 32, 0, g->stream>>>(reinterpret_cast<const T*>(g->misc->out6), red, nd, 1);
      g->launches++;
    } else {
      const double empty[6] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
      ZB_CUDA(cudaMemcpyAsync(red, empty, sizeof empty, cudaMemcpyHostToDevice, g->stream));
    }
    ZB_NCCL(N.AllReduce(red, red, 6, ncclDouble, ncclMax, N.comm, g->stream));
    if (spec) {
      Box6 expect;
      for (int k = 0; k < 6; ++k) expect.v[k] = g->slab.box[k];
      spec_check_kernel<<<1, 32, 0, g->stream>>>(red, expect, &g->misc->slab_flag);
      g->launches++;
    }
  }
  if (!spec) {
    ZB_CUDA(cudaMemcpyAsync(g->h_red, red, 6 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    for (int k = 0; k < 6; ++k) g->slab.box[k] = g->h_red[k];
    g->slab.have_box = true;
  }
  g->slab.spec = spec;
  double inf[3] = {0, 0, 0}, sup[3] = {0, 0, 0};
  bool any = true;
  for (int d = 0; d < nd; ++d) {
    inf[d] = -g->slab.box[d];
    sup[d] = g->slab.box[3 + d];
    any = any && std::isfinite(inf[d]) && std::isfinite(sup[d]);
  }
  if (!any)  // no particle anywhere: Aabb of an empty set is zeros (util.rs:41)
    for (int d = 0; d < 3; ++d) inf[d] = sup[d] = 0.0;
  for (int d = 0; d < 3; ++d) {
    g->inf[d] = inf[d];
    g->sup[d] = sup[d];
  }
  ZB_TRY(derive_shape<T>(g));
  const int ax = nd - 1;
  const int64_t nz = g->shape[ax];
  // every rank must own at least one layer: the lower halo is taken from rank - 1's top layer only, a
  // rank with an empty slab would cut the chain.  nz comes from the all-reduced box, so every rank takes
  // this exit together and nobody is left waiting in a collective.
  if (nz < N.world)
    return fail(g, ZB_ERR_BAD_ARG, "the grid has %lld layers along the slab axis, fewer than the %d ranks", (long long)nz,
                N.world);
  const int64_t z_begin = (int64_t)N.rank * nz / N.world, z_end = (int64_t)(N.rank + 1) * nz / N.world;

  // 2 + 3. sharded rebuild with the imposed box.  K2 over the local rows also extracts this slab's top
  //    layer (no extra pass over the input); the hook then trades halos -- top layer -> rank + 1, top
  //    layer of rank - 1 -> behind the local rows -- and K2 counts the received rows; K3, K4 run over
  //    local + halo rows.  The number of halo rows never visits the host: the launches are sized for the
  //    capacity and stop at the count the unpack kernel left in misc->halo_n.
  const uint64_t room = std::min<uint64_t>(halo_cap, cap_rows - n_local);  // halo rows buf can take
  const size_t block = (halo_cap + 1) * 4 * sizeof(T);
  ZB_TRY(reserve(g, g->halo_send, block));
  ZB_TRY(reserve(g, g->halo_recv, block));
  ZB_TRY(reserve(g, g->halo_labels, std::max<uint64_t>(halo_cap, 1) * 4));
  SlabHook<T> hook;
  hook.top.out = static_cast<T*>(g->halo_send.p);
  hook.top.cap = (uint32_t)std::min<uint64_t>(halo_cap, 0xfffffff0ull);
  hook.top.count = &g->misc->slab_count;
  hook.top.bad = reinterpret_cast<int*>(&g->misc->slab_flag);
  hook.top.label_offset = label_offset;
  hook.n_cap = n_local + room;
  hook.exchange = [&](uint64_t* halo_rows) -> int {
    *halo_rows = 0;
    const bool up = N.rank + 1 < N.world, down = N.rank > 0;
    // over mapped peer memory when the halo block fits the mapped one (every rank passes the same halo_cap),
    // else NCCL send / recv of the fixed-size block
    const bool p2p = g->p2p.ok && halo_cap <= g->p2p.halo_rows;
    const T* recv_block = static_cast<const T*>(g->halo_recv.p);
    if (p2p) {
      const unsigned long long seq = ++g->p2p.seq_halo;
      const int parity = (int)(seq & 1ull);
      if (up) {
        unsigned char* peer = g->p2p.peers.base[N.rank + 1];
        p2p_halo_push_kernel<T><<<8, 256, 0, g->stream>>>(
            static_cast<const T*>(g->halo_send.p), &g->misc->slab_count, hook.top.cap,
            reinterpret_cast<T*>(peer + P2pLayout::halo_block_off(parity, g->p2p.block_bytes)),
            reinterpret_cast<unsigned long long*>(peer + P2pLayout::halo_flag_off(parity)), seq, &g->misc->halo_ticket);
        g->launches++;
      }
      if (down) {
        unsigned char* mine = g->p2p.peers.base[N.rank];
        p2p_halo_wait_kernel<<<1, 32, 0, g->stream>>>(reinterpret_cast<const unsigned long long*>(mine + P2pLayout::halo_flag_off(parity)), seq,
                                                      &g->misc->slab_flag);
        g->launches++;
        recv_block = reinterpret_cast<const T*>(mine + P2pLayout::halo_block_off(parity, g->p2p.block_bytes));
      }
      ZB_CUDA(cudaGetLastError());
    } else if (up || down) {
      // the block header (row 0) carries the row count to the receiver
      halo_header_kernel<T><<<1, 1, 0, g->stream>>>(&g->misc->slab_count, hook.top.cap, static_cast<T*>(g->halo_send.p));
      g->launches++;
      ZB_NCCL(N.GroupStart());
      if (up) ZB_NCCL(N.Send(g->halo_send.p, block, ncclChar, N.rank + 1, N.comm, g->stream));
      if (down) ZB_NCCL(N.Recv(g->halo_recv.p, block, ncclChar, N.rank - 1, N.comm, g->stream));
      ZB_NCCL(N.GroupEnd());
    }
    if (!down || room == 0) {
      if (down) {  // no room at all behind the local rows: any halo row is an overflow
        halo_unpack_kernel<T, 3><<<1, 32, 0, g->stream>>>(recv_block, 0u, xyz, nullptr, &g->misc->halo_n, &g->misc->slab_flag);
        g->launches++;
      }
      return ZB_OK;  // misc->halo_n stays 0
    }
    const uint32_t blocks = (uint32_t)((room + 255) / 256);
    if (nd == 3)
      halo_unpack_kernel<T, 3><<<blocks, 256, 0, g->stream>>>(recv_block, (uint32_t)room, xyz + n_local * 3,
                                                              static_cast<uint32_t*>(g->halo_labels.p), &g->misc->halo_n,
                                                              &g->misc->slab_flag);
    else
      halo_unpack_kernel<T, 2><<<blocks, 256, 0, g->stream>>>(recv_block, (uint32_t)room, xyz + n_local * 2,
                                                              static_cast<uint32_t*>(g->halo_labels.p), &g->misc->halo_n,
                                                              &g->misc->slab_flag);
    g->launches++;
    ZB_CUDA(cudaGetLastError());
    *halo_rows = room;
    return ZB_OK;
  };
  LabelSrc ls{nullptr, static_cast<const uint32_t*>(g->halo_labels.p), label_offset, (uint32_t)n_local};
  ZB_TRY(rebuild_impl<T>(g, xyz, n_local, nullptr, nullptr, inf, sup, z_begin, z_end, true, &ls, &hook));
  g->n_local = n_local;
  g->n_halo = 0;  // known once the step's verdict has been collected (slab_validate)
  g->slab.buf = buf;
  g->slab.n_local = n_local;
  g->slab.cap_rows = cap_rows;
  g->slab.halo_cap = halo_cap;
  g->slab.label_offset = label_offset;
  if (out) {
    memset(out, 0, sizeof *out);
    for (int d = 0; d < 3; ++d) {
      out->inf[d] = g->inf[d];
      out->sup[d] = g->sup[d];
      out->shape[d] = d < nd ? g->shape[d] : 0;
    }
    out->z_begin = z_begin;
    out->z_end = z_end;
    out->n_local = n_local;
    out->n_halo = ~0ull;  // counted on the device: zb_grid_info().n - n_local after the next synchronising call
  }
  return ZB_OK;
}

static int slab_step(zb_grid* g, void* buf, uint64_t n_local, uint64_t cap_rows, const double* cutoff, uint32_t label_offset,
                     uint64_t halo_cap, zb_slab_info* out, bool allow_spec) {
  return g->dtype == ZB_F32 ? slab_step_impl<float>(g, buf, n_local, cap_rows, cutoff, label_offset, halo_cap, out, allow_spec)
                            : slab_step_impl<double>(g, buf, n_local, cap_rows, cutoff, label_offset, halo_cap, out, allow_spec);
}

// The verdict of a slab step whose copies have reached the host (the caller has synchronised the stream or
// waits here).  A speculative step whose box changed is repeated without speculation -- by EVERY rank: all
// compare the same all-reduced box with the same remembered one.  *redone tells the caller to repeat what it
// had queued behind the step.
static int slab_collect(zb_grid* g, bool* redone) {
  if (redone) *redone = false;
  if (!g->slab.pending) return ZB_OK;
  ZB_CUDA(cudaEventSynchronize(g->info_event));
  g->slab.pending = false;
  const uint32_t flag = g->h_misc->slab_flag;
  if (g->slab.spec && (flag & 4u)) {
    g->slab.penalty = std::min(64, std::max(2, 2 * g->slab.penalty));
    g->slab.backoff = g->slab.penalty;
    ZB_TRY(slab_step(g, g->slab.buf, g->slab.n_local, g->slab.cap_rows, nullptr, g->slab.label_offset, g->slab.halo_cap, nullptr,
                     false));
    ZB_CUDA(cudaEventSynchronize(g->info_event));
    g->slab.pending = false;
    if (redone) *redone = true;
  } else if (g->slab.spec) {
    g->slab.penalty = 0;
  }
  g->info_pending = false;
  g->n_cells_nonempty = g->h_misc->nonempty;
  const uint32_t f2 = g->h_misc->slab_flag;
  g->n_halo = g->h_misc->halo_n;
  g->n = g->n_local + g->n_halo;
  if (f2 & 8u) {
    g->built = false;
    return fail(g, ZB_ERR_CUDA, "a peer did not arrive at an exchange of the slab step within %llu s",
                (unsigned long long)(kP2pTimeoutNs / 1000000000ull));
  }
  if (f2 & 2u) {
    g->built = false;
    return fail(g, ZB_ERR_CAPACITY, "the neighbour's top layer exceeds halo_cap = %llu rows (or the spare rows of buf)",
                (unsigned long long)g->slab.halo_cap);
  }
  if (f2 & 1u) {
    g->built = false;
    return fail(g, ZB_ERR_OUT_OF_WINDOW, "slab-local input held a particle outside its own layers");
  }
  if (g->h_misc->flags & 1) {
    g->built = false;
    return fail(g, ZB_ERR_OUT_OF_WINDOW, "a particle lies outside the imposed box / slab window");
  }
  return ZB_OK;
}

extern "C" {

int zb_grid_rebuild_slab_local(zb_grid* g, void* buf, uint64_t n_local, uint64_t cap_rows, const double* cutoff_or_null,
                               uint32_t label_offset, uint64_t halo_cap, zb_slab_info* out) {
  ZB_TRY(enter(g));
  return slab_step(g, buf, n_local, cap_rows, cutoff_or_null, label_offset, halo_cap, out, true);
}

int zb_grid_lj_energy_allreduce(zb_grid* g, int cmp, double filter_cutoff, double* energy, uint64_t* n_pairs) {
  ZB_TRY(enter(g));
  auto& N = g->nccl;
  if (!N.comm) return fail(g, ZB_ERR_NOT_BUILT, "zb_comm_init has not been called");
  if (!energy) return fail(g, ZB_ERR_BAD_ARG, "energy is NULL");
  if (cmp < 1 || cmp > 2) return fail(g, ZB_ERR_BAD_ARG, "lj energy needs a distance filter (cmp LT or LE)");
  // A rank whose slab step failed (halo overflow, particle outside its layers, ...) still enters the
  // collective -- with a raised error flag in the third slot -- so that its peers are not left hanging;
  // every rank then reports the failure.  The flag is raised on the DEVICE from the build's own flags, so
  // the whole step needs ONE host round trip: this read.
  double* red = static_cast<double*>(g->red.p);
  for (int attempt = 0; attempt < 2; ++attempt) {
    int local_rc = g->built ? ZB_OK : ZB_ERR_NOT_BUILT;
    if (g->p2p.ok) {
      // over peer memory: fold of the per-block partials + verdict + all-reduce in ONE launch behind the LJ kernel
      uint32_t nblocks = 0;
      if (local_rc == ZB_OK)
        local_rc = g->dtype == ZB_F32 ? lj_impl<float>(g, cmp, filter_cutoff, &nblocks) : lj_impl<double>(g, cmp, filter_cutoff, &nblocks);
      p2p_energy_kernel<<<1, 256, 0, g->stream>>>(g->p2p.peers, N.world, N.rank, ++g->p2p.seq_energy,
                                                  static_cast<const double*>(g->block_energy.p),
                                                  static_cast<const unsigned long long*>(g->block_totals.p),
                                                  local_rc == ZB_OK ? nblocks : 0u, local_rc == ZB_OK ? 0 : 1, &g->misc->flags,
                                                  &g->misc->slab_flag, &g->misc->energy, &g->misc->pair_total, red);
      g->launches++;
      ZB_CUDA(cudaGetLastError());
    } else {
      if (local_rc == ZB_OK)
        local_rc = g->dtype == ZB_F32 ? lj_impl<float>(g, cmp, filter_cutoff) : lj_impl<double>(g, cmp, filter_cutoff);
      if (local_rc == ZB_OK) {
        // (energy, pair count as f64: exact below 2^53, error flag) -> one all-reduce(sum) -> host
        pack_energy_count_kernel<<<1, 1, 0, g->stream>>>(&g->misc->energy, &g->misc->pair_total, &g->misc->flags,
                                                         &g->misc->slab_flag, red);
        g->launches++;
      } else {
        const double bad[3] = {0.0, 0.0, 1.0};
        ZB_CUDA(cudaMemcpyAsync(red, bad, sizeof bad, cudaMemcpyHostToDevice, g->stream));
      }
      ZB_NCCL(N.AllReduce(red, red, 3, ncclDouble, ncclSum, N.comm, g->stream));
    }
    ZB_CUDA(cudaMemcpyAsync(g->h_red, red, 3 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    ZB_CUDA(cudaStreamSynchronize(g->stream));
    if (local_rc != ZB_OK) return local_rc;  // g->err holds this rank's own reason
    bool redone = false;
    const int rc = slab_collect(g, &redone);  // no wait: the stream has just been synchronised
    if (redone && rc == ZB_OK) continue;      // the box had changed under a speculative step: every rank repeats
    if (rc != ZB_OK) return rc;
    if (g->h_red[2] != 0.0)
      return fail(g, ZB_ERR_NOT_BUILT, "%d rank(s) failed their slab step; the all-reduced energy is not valid", (int)g->h_red[2]);
    *energy = g->h_red[0];
    if (n_pairs) *n_pairs = (uint64_t)g->h_red[1];
    return ZB_OK;
  }
  return fail(g, ZB_ERR_NOT_BUILT, "slab step could not be completed");
}

int zb_grid_profile(zb_grid* g, int enable) {
  ZB_TRY(enter(g));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  for (auto& sp : g->spans) {
    g->ev_free.push_back(sp.a);
    g->ev_free.push_back(sp.b);
  }
  g->spans.clear();
  for (int i = 0; i < ZB_NSTAGES; ++i) {
    g->stage_ms[i] = 0.0;
    g->stage_launches[i] = 0;
  }
  // 0 = off, 1 = every stage, otherwise bit (s + 1) selects stage s
  g->profile = enable == 0 ? 0u : (enable == 1 ? 0xffffffffu : ((uint32_t)enable >> 1));
  return ZB_OK;
}

int zb_grid_profile_read(zb_grid* g, double* stage_ms, uint64_t* stage_launches) {
  ZB_TRY(enter(g));
  ZB_CUDA(cudaStreamSynchronize(g->stream));
  for (auto& sp : g->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.stage >= 0 && sp.stage < ZB_NSTAGES) {
      g->stage_ms[sp.stage] += (double)ms;
      g->stage_launches[sp.stage] += 1;
    } else {
      cudaGetLastError();
    }
    g->ev_free.push_back(sp.a);
    g->ev_free.push_back(sp.b);
  }
  g->spans.clear();
  for (int i = 0; i < ZB_NSTAGES; ++i) {
    if (stage_ms) stage_ms[i] = g->stage_ms[i];
    if (stage_launches) stage_launches[i] = g->stage_launches[i];
  }
  return ZB_OK;
}

uint64_t zb_grid_launch_count(const zb_grid* g) { return g ? g->launches : 0; }

}  // extern "C"
