//! `extern "C"` declarations of include/coupe_b200_tools.h: the steps either side of the
//! RCB path in the reference's tool chain (barycentres, weight-gen, part loads / imbalance,
//! MeWe / MePe files, the `rcb,ITER[,TOL]` spec).  NOT COMPILED IN THIS REPOSITORY (no Rust
//! toolchain in the build image); the tested binding is coupe_b200/_lib.py, argument for
//! argument the same.
use std::ffi::c_void;
use std::os::raw::{c_char, c_int};

use crate::Ctx;

extern "C" {
    /// tools/lib/lib.rs:511-539
    pub fn coupe_b200_barycentres_device(
        ctx: *mut Ctx, stream: *mut c_void, dim: usize, n_elems: usize, nodes_per_elem: usize,
        elem_nodes_dev: *const u64, coords_dev: *const f64, n_nodes: usize, out_dev: *mut f64,
    ) -> c_int;
    /// tools/bins/weight-gen.rs:122-137
    pub fn coupe_b200_weight_linear_device(
        ctx: *mut Ctx, stream: *mut c_void, dim: usize, n: usize, points_dev: *const f64,
        axis: c_int, from: f64, to: f64, out_dev: *mut f64, min_out: *mut f64, max_out: *mut f64,
        alpha_out: *mut f64,
    ) -> c_int;
    pub fn coupe_b200_linear_alpha(from: f64, to: f64, min: f64, max: f64) -> f64;
    /// tools/bins/weight-gen.rs:138-151
    pub fn coupe_b200_weight_spike_device(
        ctx: *mut Ctx, stream: *mut c_void, dim: usize, n: usize, points_dev: *const f64,
        n_spikes: usize, heights: *const f64, positions: *const f64, out_dev: *mut f64,
    ) -> c_int;
    pub fn coupe_b200_weight_constant_device(
        ctx: *mut Ctx, stream: *mut c_void, n: usize, value: f64, out_dev: *mut f64,
    ) -> c_int;
    /// `criterion as i64`, tools/bins/weight-gen.rs:179-181
    pub fn coupe_b200_weight_to_i64_device(
        ctx: *mut Ctx, stream: *mut c_void, n: usize, in_dev: *const f64, out_dev: *mut i64,
    ) -> c_int;
    /// coupe/src/imbalance.rs:14-78
    pub fn coupe_b200_imbalance_device(
        ctx: *mut Ctx, stream: *mut c_void, n: usize, part_dev: *const u64, num_parts: usize,
        wtype: c_int, weights_dev: *const c_void, loads_out: *mut c_void, imbalance_out: *mut f64,
    ) -> c_int;
    /// mesh-io/src/weight.rs
    pub fn coupe_b200_mewe_write(
        path: *const c_char, is_integer: c_int, criterion_count: u16, count: u64,
        values: *const c_void,
    ) -> c_int;
    pub fn coupe_b200_mewe_read(
        path: *const c_char, is_integer: *mut c_int, criterion_count: *mut u16, count: *mut u64,
        values: *mut *mut c_void,
    ) -> c_int;
    /// mesh-io/src/partition.rs:45-86
    pub fn coupe_b200_mepe_write(path: *const c_char, count: u64, ids: *const u64) -> c_int;
    pub fn coupe_b200_mepe_read(path: *const c_char, count: *mut u64, ids: *mut *mut u64) -> c_int;
    pub fn coupe_b200_free(p: *mut c_void);
    /// tools/lib/lib.rs:418-421
    pub fn coupe_b200_parse_rcb_spec(
        spec: *const c_char, iter_count: *mut usize, tolerance: *mut f64,
    ) -> c_int;
}
