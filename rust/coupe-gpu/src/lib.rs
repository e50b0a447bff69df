//! GPU (B200, sm_100a) backend for coupe's recursive coordinate / inertial
//! bisection, behind the reference's own `Partition` trait.
//!
//! NOT COMPILED IN THIS REPOSITORY (the build image has no Rust toolchain);
//! the tested boundary is the C ABI of `include/coupe.h` / `include/coupe_b200.h`,
//! which this module binds one to one.
//!
//! `GpuRcb { iter_count, tolerance }` and `GpuRib { .. }` take the same
//! `(points, weights)` metadata as `coupe::Rcb` / `coupe::Rib`
//! (coupe/src/algorithms/recursive_bisection.rs:794-812, :911-930), overwrite
//! `part_ids` in place and return `coupe::Error::InputLenMismatch` on the same
//! conditions (:661-672).  There is no CPU fallback: a missing device or a
//! CUDA failure surfaces as `GpuError::Backend`.
use std::ffi::c_void;
use std::os::raw::{c_char, c_int};

use coupe::{Partition, PointND};

pub mod mj;
pub mod tools;

#[repr(C)]
pub struct Ctx {
    _private: [u8; 0],
}

/// include/coupe_b200.h — `enum coupe_b200_wtype`.
#[repr(i32)]
#[derive(Clone, Copy)]
pub enum WType {
    I32 = 0,
    I64 = 1,
    F64 = 2,
}

extern "C" {
    fn coupe_b200_ctx_create(out: *mut *mut Ctx, device: c_int) -> c_int;
    fn coupe_b200_ctx_destroy(ctx: *mut Ctx);
    fn coupe_b200_host_release(ctx: *mut Ctx);
    fn coupe_b200_rcb_host(
        ctx: *mut Ctx, partition: *mut usize, dim: usize, n: usize, points: *const f64,
        wtype: c_int, weights: *const c_void, wconst: *const c_void, iter_count: usize,
        tolerance: f64,
    ) -> c_int;
    fn coupe_b200_rib_host(
        ctx: *mut Ctx, partition: *mut usize, dim: usize, n: usize, points: *const f64,
        wtype: c_int, weights: *const c_void, wconst: *const c_void, iter_count: usize,
        tolerance: f64,
    ) -> c_int;
    fn coupe_strerror(err: c_int) -> *const c_char;
}

/// Errors of the GPU backend: the reference's own error for argument problems,
/// plus the `coupe_err` code of a device-side failure.
#[derive(Debug)]
pub enum GpuError {
    Coupe(coupe::Error),
    Backend { code: i32, message: String },
}

impl GpuError {
    /// `Ok(())` for COUPE_ERR_OK, the backend error otherwise.
    pub(crate) fn check(code: c_int) -> Result<(), GpuError> {
        if code == 0 {
            Ok(())
        } else {
            Err(backend(code))
        }
    }
}

fn backend(code: c_int) -> GpuError {
    let message = unsafe { std::ffi::CStr::from_ptr(coupe_strerror(code)) }
        .to_string_lossy()
        .into_owned();
    GpuError::Backend { code, message }
}

/// Weight element types the GPU path accepts (the ones coupe-ffi dispatches
/// on, coupe-ffi/src/data.rs:249-255).
pub trait GpuWeight: Copy {
    const TAG: WType;
}
impl GpuWeight for i32 {
    const TAG: WType = WType::I32;
}
impl GpuWeight for i64 {
    const TAG: WType = WType::I64;
}
impl GpuWeight for f64 {
    const TAG: WType = WType::F64;
}

/// One context per process and GPU (scratch buffers are kept between calls).
pub struct Context(pub(crate) *mut Ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, GpuError> {
        let mut p = std::ptr::null_mut();
        let err = unsafe { coupe_b200_ctx_create(&mut p, device) };
        if err != 0 {
            return Err(backend(err));
        }
        Ok(Context(p))
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        // the device and pinned buffers of the *_host entry points, then the context itself
        unsafe {
            coupe_b200_host_release(self.0);
            coupe_b200_ctx_destroy(self.0)
        }
    }
}

fn run<const D: usize, W: GpuWeight>(
    rib: bool, ctx: &Context, part_ids: &mut [usize], points: &[PointND<D>], weights: &[W],
    iter_count: usize, tolerance: f64,
) -> Result<(), GpuError> {
    // same order of checks as rcb(), recursive_bisection.rs:661-672
    if weights.len() != part_ids.len() {
        return Err(GpuError::Coupe(coupe::Error::InputLenMismatch {
            expected: part_ids.len(),
            actual: weights.len(),
        }));
    }
    if points.len() != part_ids.len() {
        return Err(GpuError::Coupe(coupe::Error::InputLenMismatch {
            expected: part_ids.len(),
            actual: points.len(),
        }));
    }
    // PointND<D> = SVector<f64, D> is D consecutive f64 (coupe/src/geometry.rs:14-16)
    let f = if rib { coupe_b200_rib_host } else { coupe_b200_rcb_host };
    let err = unsafe {
        f(
            ctx.0, part_ids.as_mut_ptr(), D, points.len(), points.as_ptr() as *const f64,
            W::TAG as c_int, weights.as_ptr() as *const c_void, std::ptr::null(), iter_count,
            tolerance,
        )
    };
    if err != 0 {
        return Err(backend(err));
    }
    Ok(())
}

/// Drop-in for `coupe::Rcb` (same fields, same meaning).
pub struct GpuRcb<'c> {
    pub iter_count: usize,
    pub tolerance: f64,
    pub context: &'c Context,
}

impl<'c, 'a, const D: usize, W: GpuWeight> Partition<(&'a [PointND<D>], &'a [W])> for GpuRcb<'c> {
    type Metadata = ();
    type Error = GpuError;

    fn partition(
        &mut self, part_ids: &mut [usize], (points, weights): (&'a [PointND<D>], &'a [W]),
    ) -> Result<(), GpuError> {
        run::<D, W>(false, self.context, part_ids, points, weights, self.iter_count, self.tolerance)
    }
}

/// Drop-in for `coupe::Rib`.
pub struct GpuRib<'c> {
    pub iter_count: usize,
    pub tolerance: f64,
    pub context: &'c Context,
}

impl<'c, 'a, const D: usize, W: GpuWeight> Partition<(&'a [PointND<D>], &'a [W])> for GpuRib<'c> {
    type Metadata = ();
    type Error = GpuError;

    fn partition(
        &mut self, part_ids: &mut [usize], (points, weights): (&'a [PointND<D>], &'a [W]),
    ) -> Result<(), GpuError> {
        run::<D, W>(true, self.context, part_ids, points, weights, self.iter_count, self.tolerance)
    }
}
