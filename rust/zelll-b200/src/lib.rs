//! `CellGrid` with zelll's API surface (src/cellgrid.rs:166-451 of zelll v0.5.0) on top of the C ABI of
//! libzelll_b200.so.  Construction, pair enumeration and the consumers the reference benches put behind it
//! run on the GPU; this crate only marshals particles in and labels out.
//!
//! Differences a caller can observe, all inside what upstream leaves unspecified:
//! * pair order and cell order are the device's (upstream: hash-map order, src/cellgrid/iters.rs:251, 262);
//! * `particle_pairs()` materialises the (label, label) list on first use after a rebuild instead of
//!   walking cells lazily;
//! * a non-zero status from the engine panics with its message, where upstream would `expect`
//!   (src/cellgrid.rs:227-229).
//!
//! Not compiled in the repository's CI image (no cargo/rustc there); kept compilable in principle.
use std::cell::{Ref, RefCell};
use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::c_int;
use std::ptr;

use zelll_b200_sys as sys;

/// zelll's `ParticleLike` (src/lib.rs:132-135): anything that yields coordinates.
pub trait ParticleLike<T = [f64; 3]>: Clone {
    fn coords(&self) -> T;
}
impl<T: Copy, const N: usize> ParticleLike<[T; N]> for [T; N] {
    fn coords(&self) -> [T; N] {
        *self
    }
}
/// enumerated particles `(usize, P)` (src/lib.rs:225-234): the label is carried along, the engine's own
/// labels are the positions in the input
impl<P: ParticleLike<C>, C> ParticleLike<C> for (usize, P) {
    fn coords(&self) -> C {
        self.1.coords()
    }
}
impl<P: ParticleLike<C>, C> ParticleLike<C> for &P {
    fn coords(&self) -> C {
        (*self).coords()
    }
}

/// coordinate types the engine computes in
pub trait Scalar: Copy + Default + PartialOrd + 'static {
    const DTYPE: c_int;
    fn to_f64(self) -> f64;
    fn from_f64(v: f64) -> Self;
}
impl Scalar for f32 {
    const DTYPE: c_int = sys::ZB_F32;
    fn to_f64(self) -> f64 {
        self as f64
    }
    fn from_f64(v: f64) -> f32 {
        v as f32
    }
}
impl Scalar for f64 {
    const DTYPE: c_int = sys::ZB_F64;
    fn to_f64(self) -> f64 {
        self
    }
    fn from_f64(v: f64) -> f64 {
        v
    }
}

/// distance filter applied to candidate pairs (`zb_cmp`)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Filter {
    /// unfiltered candidates, what `particle_pairs()` yields upstream (src/cellgrid.rs:338-340)
    None,
    /// `dsq < cutoff^2` (benches/lj.rs:85)
    Lt,
    /// `dsq <= cutoff^2` (benches/cellgrid.rs:86, benches/iters.rs:74)
    Le,
}
impl Filter {
    fn code(self) -> c_int {
        match self {
            Filter::None => sys::ZB_CMP_NONE,
            Filter::Lt => sys::ZB_CMP_LT,
            Filter::Le => sys::ZB_CMP_LE,
        }
    }
}

/// `GridInfo` (src/cellgrid/util.rs:81-181): shape, strides, origin, cutoff of the last rebuild
#[derive(Clone, Copy, Debug, Default)]
pub struct GridInfo<const N: usize, T> {
    raw: sys::zb_info,
    _t: PhantomData<T>,
}
impl<const N: usize, T: Scalar> GridInfo<N, T> {
    pub fn origin(&self) -> [T; N] {
        std::array::from_fn(|d| T::from_f64(self.raw.inf[d]))
    }
    pub fn bounding_box(&self) -> ([T; N], [T; N]) {
        (self.origin(), std::array::from_fn(|d| T::from_f64(self.raw.sup[d])))
    }
    pub fn shape(&self) -> [i32; N] {
        std::array::from_fn(|d| self.raw.shape[d])
    }
    pub fn strides(&self) -> [i32; N] {
        std::array::from_fn(|d| self.raw.strides[d])
    }
    pub fn cutoff(&self) -> T {
        T::from_f64(self.raw.cutoff)
    }
    /// number of non-empty cells (`CellGrid::iter().count()`)
    pub fn n_cells(&self) -> usize {
        self.raw.n_cells as usize
    }
}

/// The cell grid.  `P` is the caller's particle type; the engine sees only `coords()`.
pub struct CellGrid<P, const N: usize = 3, T: Scalar = f64> {
    handle: *mut sys::zb_grid,
    particles: Vec<P>,
    info: GridInfo<N, T>,
    /// unfiltered candidate pairs of the current build, materialised on first use
    pairs: RefCell<Option<Vec<[u32; 2]>>>,
}

// the handle owns device memory and one stream; it is not re-entrant but may move between threads
unsafe impl<P: Send, const N: usize, T: Scalar> Send for CellGrid<P, N, T> {}

fn check(handle: *const sys::zb_grid, status: c_int) {
    if status != sys::ZB_OK {
        let msg = unsafe { CStr::from_ptr(sys::zb_last_error(handle)) }.to_string_lossy().into_owned();
        panic!("zelll-b200: status {status}: {msg}");
    }
}

impl<P, const N: usize, T: Scalar> Default for CellGrid<P, N, T> {
    /// `CellGrid::default()` (src/cellgrid.rs:112): an empty grid with cutoff 1 on device 0
    fn default() -> Self {
        assert!(N == 2 || N == 3, "the engine supports N = 2 and N = 3");
        let mut handle = ptr::null_mut();
        let status = unsafe { sys::zb_grid_create(T::DTYPE, N as c_int, 0, &mut handle) };
        assert!(status == sys::ZB_OK && !handle.is_null(), "zelll-b200: no CUDA device (there is no CPU fallback)");
        let mut raw = sys::zb_info::default();
        raw.cutoff = 1.0;
        Self { handle, particles: Vec::new(), info: GridInfo { raw, _t: PhantomData }, pairs: RefCell::new(None) }
    }
}

impl<P, const N: usize, T: Scalar> Drop for CellGrid<P, N, T> {
    fn drop(&mut self) {
        unsafe { sys::zb_grid_destroy(self.handle) }
    }
}

impl<P: ParticleLike<[T; N]>, const N: usize, T: Scalar> CellGrid<P, N, T> {
    /// `CellGrid::new(particles, cutoff)` (src/cellgrid.rs:166-172)
    pub fn new<I>(particles: I, cutoff: T) -> Self
    where
        I: IntoIterator<Item = P> + Clone,
    {
        CellGrid::default().rebuild(particles, Some(cutoff))
    }

    /// `rebuild(self, particles, cutoff)` (src/cellgrid.rs:187-238): consumes and returns the grid
    #[must_use = "rebuild() consumes `self` and returns the rebuilt `CellGrid`"]
    pub fn rebuild<I>(mut self, particles: I, cutoff: Option<T>) -> Self
    where
        I: IntoIterator<Item = P> + Clone,
    {
        self.rebuild_mut(particles, cutoff);
        self
    }

    /// `rebuild_mut(&mut self, particles, cutoff)` (src/cellgrid.rs:264-312): device buffers are re-used.
    /// The input is walked ONCE (the reference iterates it three to four times).
    pub fn rebuild_mut<I>(&mut self, particles: I, cutoff: Option<T>)
    where
        I: IntoIterator<Item = P> + Clone,
    {
        self.particles.clear();
        self.particles.extend(particles);
        let xyz: Vec<[T; N]> = self.particles.iter().map(|p| p.coords()).collect();
        let cutoff = cutoff.map(Scalar::to_f64);
        let cutoff_ptr = cutoff.as_ref().map_or(ptr::null(), |c| c as *const f64);
        check(self.handle, unsafe { sys::zb_grid_rebuild(self.handle, xyz.as_ptr().cast(), xyz.len() as u64, cutoff_ptr) });
        check(self.handle, unsafe { sys::zb_grid_info(self.handle, &mut self.info.raw) });
        *self.pairs.borrow_mut() = None;
    }

    /// `info()` (src/cellgrid.rs:346-348)
    pub fn info(&self) -> &GridInfo<N, T> {
        &self.info
    }

    /// (label, label) rows of the pairs the filter keeps; `Filter::None` = upstream's candidate pairs
    pub fn pair_labels(&self, filter: Filter, cutoff: T) -> Vec<[u32; 2]> {
        let fc = cutoff.to_f64();
        let mut needed = 0u64;
        let status = unsafe { sys::zb_grid_pairs(self.handle, filter.code(), fc, ptr::null_mut(), 0, &mut needed) };
        if status != sys::ZB_ERR_CAPACITY {
            check(self.handle, status);
        }
        let mut rows = vec![[0u32; 2]; needed as usize];
        if needed > 0 {
            let mut written = 0u64;
            check(self.handle, unsafe {
                sys::zb_grid_pairs(self.handle, filter.code(), fc, rows.as_mut_ptr().cast(), needed, &mut written)
            });
            rows.truncate(written as usize);
        }
        rows
    }

    fn candidates(&self) -> Ref<'_, Vec<[u32; 2]>> {
        if self.pairs.borrow().is_none() {
            let rows = self.pair_labels(Filter::None, self.info.cutoff());
            *self.pairs.borrow_mut() = Some(rows);
        }
        Ref::map(self.pairs.borrow(), |p| p.as_ref().unwrap())
    }

    /// `particle_pairs()` (src/cellgrid.rs:338-340): every unordered pair of particles in the same or in
    /// neighbouring cells, once
    #[must_use = "iterators are lazy and do nothing unless consumed"]
    pub fn particle_pairs(&self) -> impl Iterator<Item = (&P, &P)> + Clone + '_ {
        let rows: Vec<[u32; 2]> = self.candidates().clone();
        rows.into_iter().map(move |[i, j]| (&self.particles[i as usize], &self.particles[j as usize]))
    }

    /// `par_particle_pairs()` (src/cellgrid.rs:447-451): the enumeration itself already ran in parallel on
    /// the device; this hands the materialised list to rayon
    #[cfg(feature = "rayon")]
    pub fn par_particle_pairs(&self) -> impl rayon::iter::ParallelIterator<Item = (&P, &P)> + '_
    where
        P: Send + Sync,
    {
        use rayon::prelude::*;
        let rows: Vec<[u32; 2]> = self.candidates().clone();
        rows.into_par_iter().map(move |[i, j]| (&self.particles[i as usize], &self.particles[j as usize]))
    }

    /// `query_neighbors(particle)` (src/cellgrid.rs:391-401): the particles of the query's cell and of its
    /// full neighbourhood, or `None` when the query lies outside the grid's one-cell margin
    #[must_use = "iterators are lazy and do nothing unless consumed"]
    pub fn query_neighbors<Q: ParticleLike<[T; N]>>(&self, particle: Q) -> Option<impl Iterator<Item = &P> + Clone + '_> {
        let q = particle.coords();
        let mut offsets = [0u64; 2];
        let mut valid = 0u8;
        let mut needed = 0u64;
        let status = unsafe {
            sys::zb_grid_query_neighbors(self.handle, q.as_ptr().cast(), 1, sys::ZB_CMP_NONE, 0.0, offsets.as_mut_ptr(), &mut valid, ptr::null_mut(), 0, &mut needed)
        };
        if status != sys::ZB_ERR_CAPACITY {
            check(self.handle, status);
        }
        if valid == 0 {
            return None;
        }
        let mut labels = vec![0u32; needed as usize];
        if needed > 0 {
            check(self.handle, unsafe {
                sys::zb_grid_query_neighbors(self.handle, q.as_ptr().cast(), 1, sys::ZB_CMP_NONE, 0.0, offsets.as_mut_ptr(), &mut valid, labels.as_mut_ptr(), needed, &mut needed)
            });
        }
        Some(labels.into_iter().map(move |l| &self.particles[l as usize]))
    }

    /// `cell_storage()` (src/cellgrid.rs:412-414): the particles in cell order
    pub fn cell_storage(&self) -> Vec<&P> {
        let mut labels = vec![0u32; self.particles.len()];
        check(self.handle, unsafe { sys::zb_grid_cell_storage(self.handle, labels.as_mut_ptr(), ptr::null_mut()) });
        labels.into_iter().map(|l| &self.particles[l as usize]).collect()
    }

    // -- fused consumers: what the reference benches do with the iterator, without a pair list ----------

    /// `particle_pairs().filter(dsq <cmp> cutoff^2).count()` (benches/cellgrid.rs:84-88)
    pub fn count_pairs(&self, filter: Filter, cutoff: T) -> u64 {
        let mut out = 0u64;
        check(self.handle, unsafe { sys::zb_grid_pair_count(self.handle, filter.code(), cutoff.to_f64(), &mut out) });
        out
    }

    /// `particle_pairs().filter(dsq < cutoff^2).map(lj).sum()` (benches/lj.rs:42-47, 81-92); also returns
    /// the number of pairs kept
    pub fn lj_energy(&self, filter: Filter, cutoff: T) -> (f64, u64) {
        let (mut energy, mut kept) = (0f64, 0u64);
        check(self.handle, unsafe { sys::zb_grid_lj_energy(self.handle, filter.code(), cutoff.to_f64(), &mut energy, &mut kept) });
        (energy, kept)
    }
}
