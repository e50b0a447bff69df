"""Host-side mirror of the reference tool chain either side of the RCB path
(SURVEY.md §8f N1–N3), on torch CUDA tensors:

  Mesh / barycentres   tools/lib/lib.rs:511-539, mesh-io/src/lib.rs:24-65,181-224
  weight_gen           tools/bins/weight-gen.rs:52-183 (distribution specs, `-i`)
  write/read_weights   mesh-io/src/weight.rs ("MeWe")
  write/read_partition mesh-io/src/partition.rs ("MePe")
  parse_algorithm      tools/lib/lib.rs:418-421 ("rcb,ITER[,TOL]")
  imbalance            coupe/src/imbalance.rs:42-78
  mesh_part            tools/bins/mesh-part.rs: mesh + weights -> partition, RCB only

Everything that touches per-element data runs in CUDA kernels
(coupe_b200/csrc/tools.cu, include/coupe_b200_tools.h); there is no CPU
fallback — without the library or a GPU the calls raise."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .api import BackendError, Context, Error, Rcb, default_context

# ElementType::node_count / dimension (mesh-io/src/lib.rs:35-55)
ELEMENT_NODES = {"vertex": 1, "edge": 2, "triangle": 3, "quadrangle": 4, "quadrilateral": 4,
                 "tetrahedron": 4, "hexahedron": 8}
ELEMENT_DIM = {"vertex": 0, "edge": 1, "triangle": 2, "quadrangle": 2, "quadrilateral": 2,
               "tetrahedron": 3, "hexahedron": 3}


@dataclass
class Mesh:
    """The arrays of mesh_io::Mesh (mesh-io/src/lib.rs:60-65): `coordinates` is an (n_nodes, D)
    float64 CUDA tensor, `topology` a list of (element type name, (n_elems, nodes_per_elem)
    int64/uint64 CUDA tensor of node indices)."""

    dimension: int
    coordinates: "object"
    topology: list = field(default_factory=list)

    def element_count(self) -> int:
        return sum(int(t.shape[0]) for _, t in self.topology)


def _ctx(ctx, tensor) -> Context:
    return ctx or default_context(tensor.device.index)


def _stream(t):
    import torch

    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(err: int):
    if err != 0:
        raise BackendError(err)


def barycentres(mesh: Mesh, ctx: Context | None = None):
    """coupe_tools::barycentres: centres of the elements of the highest dimension present
    (edges excluded), in topology order.  Returns an (n, D) float64 CUDA tensor."""
    import torch

    if not mesh.topology:
        return torch.empty((0, mesh.dimension), dtype=torch.float64, device=mesh.coordinates.device)
    top = max(ELEMENT_DIM[name] for name, _ in mesh.topology)
    blocks = [(name, t) for name, t in mesh.topology if ELEMENT_DIM[name] == top and name != "edge"]
    co = mesh.coordinates
    if co.dtype != torch.float64 or not co.is_contiguous() or co.shape[1] != mesh.dimension:
        raise Error("coordinates must be a contiguous (n_nodes, D) float64 CUDA tensor")
    n = sum(int(t.shape[0]) for _, t in blocks)
    out = torch.empty((n, mesh.dimension), dtype=torch.float64, device=co.device)
    L = _lib.lib()
    at = 0
    with torch.cuda.device(co.device):
        for name, t in blocks:
            if t.dtype not in (torch.int64, torch.uint64) or not t.is_contiguous() or t.shape[1] != ELEMENT_NODES[name]:
                raise Error(f"{name} block must be a contiguous (n, {ELEMENT_NODES[name]}) int64 tensor")
            m = int(t.shape[0])
            _check(L.coupe_b200_barycentres_device(
                _ctx(ctx, co)._h, _stream(co), mesh.dimension, m, int(t.shape[1]), C.c_void_p(t.data_ptr()),
                C.c_void_p(co.data_ptr()), int(co.shape[0]), C.c_void_p(out[at:].data_ptr())))
            at += m
    return out


# ---- weight-gen -------------------------------------------------------------------------------
_AXES = {"0": 0, "x": 0, "X": 0, "1": 1, "y": 1, "Y": 1, "2": 2, "z": 2, "Z": 2}


def _f64_arg(s: str) -> float:
    try:
        v = float(s)
    except ValueError:
        raise Error(f"arg {s!r} is not a valid float") from None
    if not np.isfinite(v):
        raise Error(f"arg {s!r} is not finite")
    return v


def parse_distribution(definition: str, dim: int):
    """weight-gen.rs:52-114: "constant,V" | "linear,AXIS,FROM,TO" | "spike,H,POS..,[H,POS..]"."""
    args = definition.split(",")
    name = args[0]
    if name == "constant":
        if len(args) < 2:
            raise Error("not enough arguments")
        return ("constant", _f64_arg(args[1]))
    if name == "linear":
        if len(args) < 4:
            raise Error("not enough arguments")
        if args[1] not in _AXES:
            raise Error(f"arg {args[1]!r} is not a valid axis")
        return ("linear", _AXES[args[1]], _f64_arg(args[2]), _f64_arg(args[3]))
    if name == "spike":
        spikes, rest = [], args[1:]
        while rest:
            h = _f64_arg(rest[0])
            if h <= 0.0:
                raise Error(f"expected 'spike' height to be strictly positive, found {h}")
            if len(rest) < 1 + dim:
                raise Error("not enough arguments")
            spikes.append((h, [_f64_arg(a) for a in rest[1:1 + dim]]))
            rest = rest[1 + dim:]
        return ("spike", spikes)
    raise Error(f"unknown distribution {name!r}")


def weight_gen(points, distribution, integers: bool = False, ctx: Context | None = None):
    """One criterion of weight-gen for the given (n, D) float64 CUDA points: `distribution` is a
    spec string or a parsed tuple.  Returns a float64 tensor, or int64 with `integers` (-i)."""
    import torch

    if points.dtype != torch.float64 or points.dim() != 2 or not points.is_contiguous() or not points.is_cuda:
        raise Error("points must be a contiguous (n, D) float64 CUDA tensor")
    n, dim = int(points.shape[0]), int(points.shape[1])
    d = parse_distribution(distribution, dim) if isinstance(distribution, str) else distribution
    out = torch.empty(n, dtype=torch.float64, device=points.device)
    L, h = _lib.lib(), _ctx(ctx, points)._h
    with torch.cuda.device(points.device):
        if d[0] == "constant":
            _check(L.coupe_b200_weight_constant_device(h, _stream(points), n, float(d[1]), C.c_void_p(out.data_ptr())))
        elif d[0] == "linear":
            if d[1] >= dim:
                raise BackendError(3)
            _check(L.coupe_b200_weight_linear_device(h, _stream(points), dim, n, C.c_void_p(points.data_ptr()),
                                                     int(d[1]), float(d[2]), float(d[3]),
                                                     C.c_void_p(out.data_ptr()), None, None, None))
        elif d[0] == "spike":
            hs = np.array([s[0] for s in d[1]], dtype=np.float64)
            ps = np.array([s[1] for s in d[1]], dtype=np.float64).reshape(-1)
            _check(L.coupe_b200_weight_spike_device(h, _stream(points), dim, n, C.c_void_p(points.data_ptr()),
                                                    len(hs), hs.ctypes.data, ps.ctypes.data,
                                                    C.c_void_p(out.data_ptr())))
        else:
            raise Error(f"unknown distribution {d[0]!r}")
        if integers:
            iout = torch.empty(n, dtype=torch.int64, device=points.device)
            _check(L.coupe_b200_weight_to_i64_device(h, _stream(points), n, C.c_void_p(out.data_ptr()),
                                                     C.c_void_p(iout.data_ptr())))
            return iout
    return out


# ---- file formats -----------------------------------------------------------------------------
def write_weights(path: str, weights) -> None:
    """MeWe file of one or more criteria: `weights` is (n,) or (n, criteria), int64 or float64
    (numpy or torch; device tensors are copied to the host)."""
    w = weights.cpu().numpy() if hasattr(weights, "cpu") else np.asarray(weights)
    if w.ndim == 1:
        w = w.reshape(-1, 1)
    if w.dtype not in (np.int64, np.float64):
        raise Error("weights must be int64 or float64")
    w = np.ascontiguousarray(w)
    _check(_lib.lib().coupe_b200_mewe_write(path.encode(), int(w.dtype == np.int64), int(w.shape[1]),
                                            int(w.shape[0]), w.ctypes.data))


def read_weights(path: str) -> np.ndarray:
    """Returns an (n, criteria) int64 or float64 array (mesh_io::weight::read)."""
    L = _lib.lib()
    is_int, cc, count, buf = C.c_int(0), C.c_uint16(0), C.c_uint64(0), C.c_void_p()
    _check(L.coupe_b200_mewe_read(path.encode(), C.byref(is_int), C.byref(cc), C.byref(count), C.byref(buf)))
    try:
        dt = np.int64 if is_int.value else np.float64
        n = count.value * cc.value
        arr = np.frombuffer(C.string_at(buf.value, n * 8), dtype=dt).copy() if n else np.zeros(0, dt)
    finally:
        L.coupe_b200_free(buf)
    return arr.reshape(count.value, cc.value) if cc.value else arr.reshape(0, 0)


def write_partition(path: str, part_ids) -> None:
    p = part_ids.cpu().numpy() if hasattr(part_ids, "cpu") else np.asarray(part_ids)
    p = np.ascontiguousarray(p.astype(np.uint64, copy=False))
    _check(_lib.lib().coupe_b200_mepe_write(path.encode(), int(p.shape[0]), p.ctypes.data))


def read_partition(path: str) -> np.ndarray:
    L = _lib.lib()
    count, buf = C.c_uint64(0), C.c_void_p()
    _check(L.coupe_b200_mepe_read(path.encode(), C.byref(count), C.byref(buf)))
    try:
        n = count.value
        arr = np.frombuffer(C.string_at(buf.value, n * 8), dtype=np.uint64).copy() if n else np.zeros(0, np.uint64)
    finally:
        L.coupe_b200_free(buf)
    return arr


def parse_algorithm(spec: str, ctx: Context | None = None) -> Rcb:
    """"rcb,ITER[,TOL]" -> Rcb (tools/lib/lib.rs:418-421); every other algorithm name is outside
    this repository's path and raises."""
    it, tol = C.c_size_t(0), C.c_double(0.0)
    err = _lib.lib().coupe_b200_parse_rcb_spec(spec.encode(), C.byref(it), C.byref(tol))
    if err != 0:
        raise Error(f"invalid algorithm {spec!r} (only rcb,ITER[,TOL] is available)")
    return Rcb(int(it.value), float(tol.value), ctx)


def imbalance(num_parts: int, part_ids, weights, ctx: Context | None = None, return_loads: bool = False):
    """coupe::imbalance::imbalance on CUDA tensors: part_ids int64/uint64 (n,), weights
    int32/int64/float64 (n,)."""
    import torch

    tag = {torch.int32: 0, torch.int64: 1, torch.float64: 2}.get(weights.dtype)
    if tag is None:
        raise BackendError(4)
    if part_ids.shape[0] != weights.shape[0]:
        raise Error("partition and weights have different lengths")
    loads = np.zeros(int(num_parts), dtype=np.float64 if tag == 2 else np.int64)
    imb = C.c_double(0.0)
    with torch.cuda.device(part_ids.device):
        _check(_lib.lib().coupe_b200_imbalance_device(
            _ctx(ctx, part_ids)._h, _stream(part_ids), int(part_ids.shape[0]), C.c_void_p(part_ids.data_ptr()),
            int(num_parts), tag, C.c_void_p(weights.data_ptr()), loads.ctypes.data if num_parts else None,
            C.byref(imb)))
    return (imb.value, loads) if return_loads else imb.value


def mesh_part(mesh: Mesh, weights, algorithm: str, ctx: Context | None = None):
    """mesh-part for the RCB path: barycentres -> Rcb::partition on criterion 0 (tools/lib/lib.rs:
    213-231).  `weights` is an (n,) or (n, criteria) CUDA tensor.  Returns the uint64 part ids."""
    import torch

    pts = barycentres(mesh, ctx)
    w = weights if weights.dim() == 1 else weights[:, 0].contiguous()
    part = torch.empty(pts.shape[0], dtype=torch.int64, device=pts.device)
    algo = parse_algorithm(algorithm, ctx)
    algo.partition(part, (pts, w))
    return part


def hex_grid(nx: int, ny: int, nz: int, device, spacing: float = 1.0) -> Mesh:
    """A structured nx*ny*nz hexahedral mesh (x fastest), built on the device: the generator of
    config C3 (mesh-io itself cannot refine hexahedra, mesh-io/src/lib.rs:394-397)."""
    import torch

    gx = torch.arange(nx + 1, device=device, dtype=torch.float64) * spacing
    gy = torch.arange(ny + 1, device=device, dtype=torch.float64) * spacing
    gz = torch.arange(nz + 1, device=device, dtype=torch.float64) * spacing
    zz, yy, xx = torch.meshgrid(gz, gy, gx, indexing="ij")
    coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
    del xx, yy, zz
    i = torch.arange(nx, device=device).view(1, 1, nx)
    j = torch.arange(ny, device=device).view(1, ny, 1)
    k = torch.arange(nz, device=device).view(nz, 1, 1)
    base = (k * (ny + 1) + j) * (nx + 1) + i
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    corners = [0, sx, sx + sy, sy, sz, sz + sx, sz + sx + sy, sz + sy]  # medit hexahedron order
    nodes = torch.stack([(base + c).reshape(-1) for c in corners], dim=1).contiguous()
    return Mesh(3, coords, [("hexahedron", nodes)])
