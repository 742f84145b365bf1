//! Raw bindings to `include/zelll_b200.h` (ABI version 1): one `extern "C"` item per entry point, one
//! `#[repr(C)]` struct per C struct, field for field.  `tests/test_rust_shim.py` parses this file and
//! the header and fails when the two drift apart (symbol set, argument count, struct field order).
//!
//! Not compiled in the repository's CI image (no cargo/rustc there); kept compilable in principle.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const ZB_ABI_VERSION: c_int = 1;

/// opaque handle: `CellGrid<(usize, [T; N]), N, T>` (src/cellgrid.rs:112-126)
#[repr(C)]
pub struct zb_grid {
    _private: [u8; 0],
}

// enum zb_dtype
pub const ZB_F32: c_int = 0;
pub const ZB_F64: c_int = 1;

// enum zb_cmp
pub const ZB_CMP_NONE: c_int = 0;
pub const ZB_CMP_LT: c_int = 1;
pub const ZB_CMP_LE: c_int = 2;

// enum zb_status
pub const ZB_OK: c_int = 0;
pub const ZB_ERR_BAD_ARG: c_int = 1;
pub const ZB_ERR_CUDA: c_int = 2;
pub const ZB_ERR_CAPACITY: c_int = 3;
pub const ZB_ERR_TOO_MANY: c_int = 4;
pub const ZB_ERR_GRID_TOO_LARGE: c_int = 5;
pub const ZB_ERR_NOT_BUILT: c_int = 6;
pub const ZB_ERR_OUT_OF_WINDOW: c_int = 7;

// enum zb_stage
pub const ZB_STAGE_BBOX: c_int = 0;
pub const ZB_STAGE_COUNT: c_int = 1;
pub const ZB_STAGE_SCAN: c_int = 2;
pub const ZB_STAGE_SCATTER: c_int = 3;
pub const ZB_STAGE_PAIR_COUNT: c_int = 4;
pub const ZB_STAGE_PAIR_EMIT: c_int = 5;
pub const ZB_STAGE_PAIR_LJ: c_int = 6;
pub const ZB_STAGE_OTHER: c_int = 7;
pub const ZB_NSTAGES: usize = 8;

/// `GridInfo` + `Aabb` (src/cellgrid/util.rs:19-27, 81-90) plus sizes
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct zb_info {
    pub inf: [f64; 3],
    pub sup: [f64; 3],
    pub cutoff: f64,
    pub shape: [i32; 3],
    pub strides: [i32; 3],
    pub n: u64,
    pub n_cells: u64,
    pub ndim: i32,
    pub dtype: i32,
    pub keys_changed: i32,
    pub reserved: i32,
}

/// result of one slab-local multi-GPU rebuild
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct zb_slab_info {
    pub inf: [f64; 3],
    pub sup: [f64; 3],
    pub shape: [i32; 3],
    pub reserved: i32,
    pub z_begin: i64,
    pub z_end: i64,
    pub n_local: u64,
    pub n_halo: u64,
}

extern "C" {
    // -- lifetime
    pub fn zb_grid_create(dtype: c_int, ndim: c_int, device: c_int, out: *mut *mut zb_grid) -> c_int;
    pub fn zb_grid_destroy(g: *mut zb_grid);
    pub fn zb_grid_set_stream(g: *mut zb_grid, cuda_stream: *mut c_void) -> c_int;
    pub fn zb_grid_track_key_changes(g: *mut zb_grid, enable: c_int) -> c_int;
    pub fn zb_grid_set_stable(g: *mut zb_grid, enable: c_int) -> c_int;
    pub fn zb_last_error(g: *const zb_grid) -> *const c_char;

    // -- construction
    pub fn zb_grid_rebuild(g: *mut zb_grid, xyz: *const c_void, n: u64, cutoff_or_null: *const f64) -> c_int;
    pub fn zb_grid_prefetch(g: *mut zb_grid, xyz_host: *const c_void, n: u64) -> c_int;
    pub fn zb_grid_prefetch_wait(g: *mut zb_grid) -> c_int;
    pub fn zb_grid_rebuild_sharded(g: *mut zb_grid, xyz: *const c_void, n: u64, labels_or_null: *const u32, cutoff_or_null: *const f64, inf: *const f64, sup: *const f64, z_begin: i64, z_end: i64) -> c_int;
    pub fn zb_aabb(g: *mut zb_grid, xyz: *const c_void, n: u64, out6: *mut f64) -> c_int;
    pub fn zb_layer_of(g: *mut zb_grid, xyz: *const c_void, n: u64, inf_axis: f64, cutoff: f64, axis: c_int, out: *mut i32) -> c_int;
    pub fn zb_slab_top_layer(g: *mut zb_grid, xyz: *const c_void, n: u64, inf_axis: f64, cutoff: f64, z_begin: i64, z_end: i64, label_offset: u32, halo_rows: *mut c_void, cap_rows: u64, n_top: *mut u64, out_of_slab: *mut c_int) -> c_int;

    // -- native multi-GPU step
    pub fn zb_comm_unique_id(nccl_lib_path: *const c_char, out128: *mut c_void) -> c_int;
    pub fn zb_comm_init(g: *mut zb_grid, nccl_lib_path: *const c_char, unique_id128: *const c_void, world: c_int, rank: c_int) -> c_int;
    pub fn zb_grid_rebuild_slab_local(g: *mut zb_grid, buf: *mut c_void, n_local: u64, cap_rows: u64, cutoff_or_null: *const f64, label_offset: u32, halo_cap: u64, out: *mut zb_slab_info) -> c_int;
    pub fn zb_grid_lj_energy_allreduce(g: *mut zb_grid, cmp: c_int, filter_cutoff: f64, energy: *mut f64, n_pairs: *mut u64) -> c_int;

    // -- inspection
    pub fn zb_grid_info(g: *mut zb_grid, out: *mut zb_info) -> c_int;
    pub fn zb_grid_keys(g: *mut zb_grid, out: *mut i32) -> c_int;
    pub fn zb_grid_neighbor_indices(g: *mut zb_grid, out: *mut i32, count: *mut i32) -> c_int;
    pub fn zb_grid_cells(g: *mut zb_grid, keys: *mut i32, begin: *mut u32, count: *mut u32, cap: u64, n_out: *mut u64) -> c_int;
    pub fn zb_grid_cell_storage(g: *mut zb_grid, labels: *mut u32, xyz: *mut c_void) -> c_int;

    // -- pair enumeration and its consumers
    pub fn zb_grid_pair_count(g: *mut zb_grid, cmp: c_int, filter_cutoff: f64, out: *mut u64) -> c_int;
    pub fn zb_grid_pairs(g: *mut zb_grid, cmp: c_int, filter_cutoff: f64, ij: *mut u32, cap: u64, n_out: *mut u64) -> c_int;
    pub fn zb_grid_lj_energy(g: *mut zb_grid, cmp: c_int, filter_cutoff: f64, energy: *mut f64, n_pairs: *mut u64) -> c_int;

    // -- point queries
    pub fn zb_grid_query_neighbors(g: *mut zb_grid, queries: *const c_void, nq: u64, cmp: c_int, filter_cutoff: f64, offsets: *mut u64, valid: *mut u8, labels: *mut u32, cap: u64, n_out: *mut u64) -> c_int;

    // -- introspection for benches
    pub fn zb_grid_profile(g: *mut zb_grid, enable: c_int) -> c_int;
    pub fn zb_grid_profile_read(g: *mut zb_grid, stage_ms: *mut f64, stage_launches: *mut u64) -> c_int;
    pub fn zb_grid_launch_count(g: *const zb_grid) -> u64;
    pub fn zb_abi_version() -> c_int;
    pub fn zb_build_id() -> *const c_char;
}
