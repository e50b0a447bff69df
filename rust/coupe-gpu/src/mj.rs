//! Multi-Jagged, axis_sort and the cartesian RCB behind the reference's own types
//! (include/coupe_b200_mj.h).  NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in
//! the build image); the tested binding is coupe_b200/multi_jagged.py, argument for argument the same.
use std::ffi::c_void;
use std::os::raw::c_int;

use coupe::{Partition, PointND};

use crate::{Context, Ctx, GpuError};

extern "C" {
    /// coupe/src/algorithms/multi_jagged.rs:150-179, :354-366
    pub fn coupe_b200_multi_jagged_host(
        ctx: *mut Ctx, part: *mut usize, dim: usize, n: usize, points: *const f64,
        weights: *const f64, part_count: usize, max_iter: usize,
    ) -> c_int;
    pub fn coupe_b200_multi_jagged_device(
        ctx: *mut Ctx, stream: *mut c_void, part_dev: *mut u64, dim: usize, n: usize,
        points_dev: *const f64, weights_dev: *const f64, part_count: usize, max_iter: usize,
    ) -> c_int;
    /// coupe/src/algorithms/recursive_bisection.rs:815-827
    pub fn coupe_b200_axis_sort_device(
        ctx: *mut Ctx, stream: *mut c_void, dim: usize, n_points: usize, points_dev: *const f64,
        permutation_dev: *mut u64, len: usize, coord: usize,
    ) -> c_int;
    pub fn coupe_b200_mj_scheme(
        part_count: usize, max_iter: usize, leaves_out: *mut u64, levels_out: *mut u64,
    ) -> c_int;
    /// coupe/src/cartesian/mod.rs:119-181
    pub fn coupe_b200_grid_rcb_host(
        ctx: *mut Ctx, part: *mut usize, dim: usize, sizes: *const u64, wtype: c_int,
        weights: *const c_void, iter_count: usize, threads: usize,
    ) -> c_int;
}

/// `coupe::MultiJagged { part_count, max_iter }` on the GPU: same fields, same `Partition` impl
/// (multi_jagged.rs:347-366: `(&[PointND<D>], &[f64])`, `Metadata = ()`).
pub struct GpuMultiJagged<'c> {
    pub part_count: usize,
    pub max_iter: usize,
    pub context: &'c Context,
}

impl<'a, 'c, const D: usize> Partition<(&'a [PointND<D>], &'a [f64])> for GpuMultiJagged<'c> {
    type Metadata = ();
    type Error = GpuError;

    fn partition(
        &mut self, part_ids: &mut [usize], (points, weights): (&'a [PointND<D>], &'a [f64]),
    ) -> Result<(), GpuError> {
        // the reference indexes `weights` and `partition` by point and panics on a short slice
        assert!(weights.len() >= points.len() && part_ids.len() >= points.len());
        let err = unsafe {
            coupe_b200_multi_jagged_host(
                self.context.0, part_ids.as_mut_ptr(), D, points.len(),
                points.as_ptr() as *const f64, weights.as_ptr(), self.part_count, self.max_iter,
            )
        };
        GpuError::check(err)
    }
}

/// `coupe::Grid::rcb` (cartesian/mod.rs:119-181) for f64 weights; `threads` is what
/// `rayon::current_num_threads()` would return in the reference (its weighted median chunks by it).
pub fn grid_rcb_f64(
    ctx: &Context, sizes: &[usize], partition: &mut [usize], weights: &[f64], iter_count: usize,
) -> Result<(), GpuError> {
    let sz: Vec<u64> = sizes.iter().map(|s| *s as u64).collect();
    let cells: usize = sizes.iter().product();
    assert!(weights.len() == cells && partition.len() == cells && (sz.len() == 2 || sz.len() == 3));
    let err = unsafe {
        coupe_b200_grid_rcb_host(
            ctx.0, partition.as_mut_ptr(), sz.len(), sz.as_ptr(), 2 /* COUPE_DOUBLE */,
            weights.as_ptr() as *const c_void, iter_count, rayon::current_num_threads().max(2),
        )
    };
    GpuError::check(err)
}
