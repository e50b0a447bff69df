// engine.cu — host side of the B200 RCB/RIB engine and its device-level C ABI
// (include/coupe_b200.h).  One context per process and GPU; the level loop
// below replaces the recursion of coupe/src/algorithms/recursive_bisection.rs
// (rcb :644-705, rcb_recurse :575-642) with one dense sweep per tree level
// plus sparse refinement sweeps only while some node's bisection is undecided.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <string>
#include <vector>

#include "../../include/coupe.h"
#include "../../include/coupe_b200.h"
#include "engine_internal.h"
#include "rcb_kernels.cuh"

namespace {

using namespace cb;

// Node tables are dense arrays of 2^level entries replicated on every GPU: 2^24 parts is where they reach a few GB.
// Beyond that the call returns COUPE_ERR_ALLOC (documented in coupe.h); the reference itself recurses to any depth.
constexpr int MAX_LEVELS = 24;
constexpr uint64_t FLAG_SLOTS = 256;  // passes in flight are at most a handful  // 2^20 parts; node tables are replicated on every GPU

struct CudaFail {
  cudaError_t err;
  const char *what;
};
#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) throw CudaFail{e__, #call}; \
  } while (0)

// Programmatic dependent launch: the kernels of the level loop are launched with the
// "programmatic stream serialization" attribute, so the next kernel's blocks are scheduled while
// the previous kernel drains; every such kernel starts with griddepcontrol.launch_dependents /
// griddepcontrol.wait (rcb_kernels.cuh: pdl_enter) and touches global memory only after the wait.
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool off = [] { const char *e = getenv("COUPE_B200_NO_PDL"); return e && *e && *e != '0'; }();
  cfg.attrs = attr;
  cfg.numAttrs = off ? 0 : 1;
  CU(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  bool load() {
    if (handle) return true;
    // Resolve against the NCCL already mapped into the process (torch's) when there is one.
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
NcclApi g_nccl;
struct NcclFail {
  ncclResult_t err;
};
#define NC(call)                                  \
  do {                                            \
    ncclResult_t r__ = (call);                    \
    if (r__ != ncclSuccess) throw NcclFail{r__}; \
  } while (0)

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) CU(cudaFree(p));
    p = nullptr;
    cap = 0;
    const size_t want = ((bytes + (1u << 20) - 1) >> 20) << 20;
    CU(cudaMalloc(&p, want));
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return static_cast<T *>(p);
  }
};

// --- small dense symmetric eigen / Householder helpers for RIB (host, f64) ----
// geometry.rs:286-319.  The reference calls nalgebra's symmetric_eigen; a cyclic
// Jacobi iteration gives the same eigenvector up to rounding (RIB).
void principal_axis(int D, const double *m, double *v) {
  double a[3][3] = {{0}}, q[3][3] = {{0}};
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) {
      a[r][s] = 0.5 * (m[r * D + s] + m[s * D + r]);
      q[r][s] = r == s;
    }
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0;
    for (int r = 0; r < D; ++r)
      for (int s = r + 1; s < D; ++s) off += std::fabs(a[r][s]);
    if (off == 0.0) break;
    for (int i = 0; i < D; ++i)
      for (int j = i + 1; j < D; ++j) {
        if (a[i][j] == 0.0) continue;
        const double tau = (a[j][j] - a[i][i]) / (2.0 * a[i][j]);
        const double t = std::copysign(1.0, tau) / (std::fabs(tau) + std::hypot(1.0, tau));
        const double c = 1.0 / std::hypot(1.0, t), s = t * c;
        for (int r = 0; r < D; ++r) {  // A <- A J
          const double x = a[r][i], y = a[r][j];
          a[r][i] = c * x - s * y;
          a[r][j] = s * x + c * y;
        }
        for (int r = 0; r < D; ++r) {  // A <- J^T A
          const double x = a[i][r], y = a[j][r];
          a[i][r] = c * x - s * y;
          a[j][r] = s * x + c * y;
        }
        for (int r = 0; r < D; ++r) {  // Q <- Q J
          const double x = q[r][i], y = q[r][j];
          q[r][i] = c * x - s * y;
          q[r][j] = s * x + c * y;
        }
      }
  }
  int best = 0;
  for (int d = 1; d < D; ++d)
    if (a[d][d] > a[best][best]) best = d;
  for (int d = 0; d < D; ++d) v[d] = q[d][best];
}

bool ulps_close(double a, double b) {  // approx::Ulps::default() on f64: eps, then 4 ulps
  if (std::fabs(a - b) <= 2.220446049250313e-16) return true;
  if (std::signbit(a) != std::signbit(b)) return false;
  int64_t ia, ib;
  memcpy(&ia, &a, 8);
  memcpy(&ib, &b, 8);
  return (ia > ib ? ia - ib : ib - ia) <= 4;
}

// H = I - 2 w w^T / (w^T w), w = v + sign * |v| e0; identity if v is parallel to e0.
void reflection(int D, const double *v, double *h) {
  double norm = 0;
  for (int d = 0; d < D; ++d) norm += v[d] * v[d];
  norm = std::sqrt(norm);
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) h[r * D + s] = r == s;
  bool par = true;
  for (int d = 0; d < D; ++d) par = par && ulps_close(v[d] / norm, d == 0 ? 1.0 : 0.0);
  if (par) return;
  const double sign = v[0] > 0.0 ? -1.0 : 1.0;
  double w[3] = {0, 0, 0}, ww = 0;
  for (int d = 0; d < D; ++d) w[d] = v[d] + (d == 0 ? sign * norm : 0.0);
  for (int d = 0; d < D; ++d) ww += w[d] * w[d];
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) h[r * D + s] -= 2.0 * w[r] * w[s] / ww;
}

bool inverse(int D, const double *a, double *o) {  // cofactor inverse (try_inverse, geometry.rs:219)
  if (D == 2) {
    const double det = a[0] * a[3] - a[1] * a[2];
    if (det == 0) return false;
    o[0] = a[3] / det; o[1] = -a[1] / det; o[2] = -a[2] / det; o[3] = a[0] / det;
    return true;
  }
  const double c00 = a[4] * a[8] - a[7] * a[5], c01 = a[3] * a[8] - a[6] * a[5],
               c02 = a[3] * a[7] - a[6] * a[4];
  const double det = a[0] * c00 - a[1] * c01 + a[2] * c02;
  if (det == 0) return false;
  o[0] = c00 / det;
  o[1] = (a[2] * a[7] - a[8] * a[1]) / det;
  o[2] = (a[1] * a[5] - a[4] * a[2]) / det;
  o[3] = -c01 / det;
  o[4] = (a[0] * a[8] - a[6] * a[2]) / det;
  o[5] = (a[2] * a[3] - a[5] * a[0]) / det;
  o[6] = c02 / det;
  o[7] = (a[1] * a[6] - a[7] * a[0]) / det;
  o[8] = (a[0] * a[4] - a[3] * a[1]) / det;
  return true;
}

}  // namespace

struct coupe_b200_ctx {
  int device = 0;
  int num_sms = 148;
  int sm_khz = 0;  // SM clock (kHz): converts clock64() differences to time
  size_t max_smem = 0;
  std::mutex mu;
  // scratch
  Buf host_w, host_ids, host_pts;  // host path: device copies of the caller's weights / (RIB) points, compact ids
  Buf xcols, ids, w32, node_rt, rfast, part_w, part_min, hist_w, hist_min, nodes_a, nodes_b, table_a, table_b, thi_a, thi_b,
      tsp_a, tsp_b, nsh_a, nsh_b, target, rtable, gp,
      tr_visited, tr_split, tr_wl, tr_sum, tr_iters, mom_partial,
      def_rec, def_slot, def_count, target_first, rpart_w, rpart_min;  // deferred points (rcb_kernels.cuh)
  uint32_t *h_pinned = nullptr;  // pinned host scratch (64 words)
  volatile unsigned long long *h_flags = nullptr;  // mapped pinned: one word per pass, written by the GPU
  unsigned long long *d_flags = nullptr;           // device view of h_flags
  uint64_t flag_seq = 0;
  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange of the level histograms (rcb_kernels.cuh: Xchg); off -> NCCL all-reduces
  bool xchg_ok = false;
  bool xchg_local = false;  // the peers' buffers belong to contexts of this process (no IPC handles to close)
  int use_xchg_opt = 1;
  unsigned char *xchg_peer[XCHG_MAX_WORLD] = {nullptr};  // [rank] is this rank's own cudaMalloc'ed buffer
  unsigned int *xchg_aux = nullptr;                       // {ticket, error}
  // options
  int kmax_a = 8, nb_smem_log2 = 14, kmax_refine = 10, force_global = 0, trace_on = 1, time_sweeps = 0;
  int smem_pad = 0;      // experiments: bytes added to the dense sweeps' dynamic shared memory request
  int table_rep_max = 3; // experiments: cap on the bank-private copies of the per-parent table (log2)
  int carve_fit = 1;     // keep the dense sweeps under the 196 KB shared-memory carve-out when the tables allow it
  int sample_w_opt = 1;  // f64 weights: max |w| from a sample, verified by the root sweep
  int defer_opt = 1;     // undecided levels: the next dense sweep lists the points of the undecided bins (no rescan of idx);
                         // 2: the deferring variant of the sweep at every level (no prediction)
  std::vector<uint8_t> undecided_last;  // per level: the context's previous call left it undecided after its dense pass
  std::vector<cudaEvent_t> events;  // time_sweeps: start/stop pairs
  std::vector<int> event_kind;      // 0 dense, 1 refine
  std::vector<int> event_level;     // tree level of the sweep
  std::vector<double> sweep_ms;     // time_sweeps: device time of every timed sweep of the last call
  // last call
  coupe_b200_stats stats{};
  uint32_t trace_levels = 0;
  bool funcs_ready = false;
};

namespace {

void setup_xchg(coupe_b200_ctx *c);

// Shared memory of a dense sweep in shared-memory mode: the three histogram arrays at fixed
// offsets (rcb_kernels.cuh: HIST_BYTES), then the per-parent table replicated 2^rep_log2 times,
// the split positions and the bracket ends.
size_t sweep_smem_bytes(int level, int rep_log2, bool aux) {
  const size_t parents = (size_t)1 << (level > 0 ? level - 1 : 0);
  size_t b = HIST_BYTES + (parents << rep_log2) * sizeof(float4);
  // the rarely read values: split positions, bracket ends, per-node shifts (f64 weights, wide form)
  if (aux) b += parents * 2 * sizeof(float) + ((((size_t)2 << level) + 15) & ~(size_t)15);
  return b + 128;  // the per-warp counters of deferred points (last 128 bytes)
}
// Shared memory is carved out of the SM's 228 KB in steps (..., 164, 196, 228 KB, 1 KB of each reserved by the
// system); what is left is the L1 cache and the staging of global loads.  A sweep that needs a few bytes more
// than 195 KB gets the 228 KB carve-out and no L1 at all, and runs 20 % slower (levels 6-9 of round 1): the
// tables are therefore shrunk (fewer bank-private copies, rare values read from global memory) to stay below.
constexpr size_t SMEM_CARVE_196 = 196 * 1024 - 1024;

template <int WIN, bool ROOT>
void launch_sweep(bool smem, bool tsm, bool idx16, int grid, size_t bytes, cudaStream_t st,
                  const SweepArgs &a) {
  constexpr bool DEFERRABLE = !ROOT && (WIN == WIN_I32 || WIN == WIN_CONST);
  if (DEFERRABLE && smem && a.def_rec) {  // the variant that lists the points of undecided bins
    if (idx16) launch_pdl(sweep_kernel<WIN, true, ROOT, true, uint16_t, DEFERRABLE>, grid, SWEEP_THREADS, bytes, st, a);
    else launch_pdl(sweep_kernel<WIN, true, ROOT, true, uint32_t, DEFERRABLE>, grid, SWEEP_THREADS, bytes, st, a);
    return;
  }
  if (smem && idx16) launch_pdl(sweep_kernel<WIN, true, ROOT, true, uint16_t>, grid, SWEEP_THREADS, bytes, st, a);
  else if (smem) launch_pdl(sweep_kernel<WIN, true, ROOT, true, uint32_t>, grid, SWEEP_THREADS, bytes, st, a);
  else if (tsm) launch_pdl(sweep_kernel<WIN, false, ROOT, true, uint32_t>, grid, SWEEP_THREADS, bytes, st, a);
  else launch_pdl(sweep_kernel<WIN, false, ROOT, false, uint32_t>, grid, SWEEP_THREADS, bytes, st, a);
}

void launch_sweep_any(int win, bool root, bool smem, bool tsm, bool idx16, int grid, size_t bytes,
                      cudaStream_t st, const SweepArgs &a) {
  if (root) {
    switch (win) {
      case WIN_I32: launch_sweep<WIN_I32, true>(smem, tsm, idx16, grid, bytes, st, a); break;
      case WIN_I64: launch_sweep<WIN_I64, true>(smem, tsm, idx16, grid, bytes, st, a); break;
      case WIN_F64: launch_sweep<WIN_F64, true>(smem, tsm, idx16, grid, bytes, st, a); break;
      default: launch_sweep<WIN_CONST, true>(smem, tsm, idx16, grid, bytes, st, a); break;
    }
  } else {
    switch (win) {
      case WIN_I32: launch_sweep<WIN_I32, false>(smem, tsm, idx16, grid, bytes, st, a); break;
      case WIN_I64: launch_sweep<WIN_I64, false>(smem, tsm, idx16, grid, bytes, st, a); break;
      case WIN_F64: launch_sweep<WIN_F64, false>(smem, tsm, idx16, grid, bytes, st, a); break;
      default: launch_sweep<WIN_CONST, false>(smem, tsm, idx16, grid, bytes, st, a); break;
    }
  }
}

template <int WIN>
void launch_refine(bool idx16, bool rts, int grid, size_t bytes, cudaStream_t st, const RefineArgs &a) {
  if (idx16) launch_pdl(sweep_refine_kernel<WIN, uint16_t, true>, grid, SWEEP_THREADS, bytes, st, a);  // 2^16 idx values: always staged
  else if (rts) launch_pdl(sweep_refine_kernel<WIN, uint32_t, true>, grid, SWEEP_THREADS, bytes, st, a);
  else launch_pdl(sweep_refine_kernel<WIN, uint32_t, false>, grid, SWEEP_THREADS, bytes, st, a);
}

void prepare_funcs(coupe_b200_ctx *c) {
  if (c->funcs_ready) return;
  const int m = (int)c->max_smem;
#define SETATTR(fn) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, m))
#define SETALL(win, root)                                           \
  SETATTR((sweep_kernel<win, true, root, true, uint16_t>));         \
  SETATTR((sweep_kernel<win, true, root, true, uint32_t>));         \
  SETATTR((sweep_kernel<win, false, root, true, uint32_t>));        \
  SETATTR((sweep_kernel<win, false, root, false, uint32_t>))
  SETATTR((sweep_kernel<WIN_I32, true, false, true, uint16_t, true>));
  SETATTR((sweep_kernel<WIN_I32, true, false, true, uint32_t, true>));
  SETATTR((sweep_kernel<WIN_CONST, true, false, true, uint16_t, true>));
  SETATTR((sweep_kernel<WIN_CONST, true, false, true, uint32_t, true>));
  SETALL(WIN_I32, true);
  SETALL(WIN_I64, true);
  SETALL(WIN_F64, true);
  SETALL(WIN_CONST, true);
  SETALL(WIN_I32, false);
  SETALL(WIN_I64, false);
  SETALL(WIN_F64, false);
  SETALL(WIN_CONST, false);
#define SETREF(win)                                            \
  SETATTR((sweep_refine_kernel<win, uint16_t, true>));         \
  SETATTR((sweep_refine_kernel<win, uint32_t, true>));         \
  SETATTR((sweep_refine_kernel<win, uint32_t, false>))
  SETREF(WIN_I32);
  SETREF(WIN_I64);
  SETREF(WIN_F64);
  SETREF(WIN_CONST);
#undef SETREF
  SETATTR(defer_refine_kernel<WIN_I32>);
  SETATTR(defer_refine_kernel<WIN_CONST>);
#undef SETALL
#undef SETATTR
  c->funcs_ready = true;
}

// How the dense first pass of a level is run.
struct FirstPlan {
  int k;               // bisection iterations resolved by the pass (2^k bins per node)
  bool smem;           // block-private shared-memory histograms, else L2 atomics
  int copies_log2;     // privatised copies per block (smem mode)
  bool table_in_smem;  // per-parent table staged in shared memory
  bool aux_in_smem;    // ... and the rarely read per-parent values with it
  int table_rep_log2;  // ... replicated 2^this times (bank-private copies at the deep levels)
  size_t bytes;        // dynamic shared memory
  size_t def_off;      // offset of the deferred-point counter in it (smem mode)
};

FirstPlan plan_first(const coupe_b200_ctx *c, int level) {
  FirstPlan p{};
  p.k = std::min(c->kmax_a, c->nb_smem_log2 - level);
  p.smem = !c->force_global && p.k >= 1;
  p.table_in_smem = true;
  if (p.smem) {
    p.copies_log2 = std::min(5, c->nb_smem_log2 - (level + p.k));
    // first choice: under the 196 KB carve-out (L1 kept), with as many table copies as fit, the rare values
    // staged if there is room for them; else the whole 227 KB
    p.table_rep_log2 = -1;
    if (c->carve_fit)
      for (int aux = 1; aux >= 0 && p.table_rep_log2 < 0; --aux)
        for (int rep = c->table_rep_max; rep >= 0; --rep)
          if (sweep_smem_bytes(level, rep, aux != 0) <= SMEM_CARVE_196) {
            p.table_rep_log2 = rep;
            p.aux_in_smem = aux != 0;
            break;
          }
    if (p.table_rep_log2 < 0) {
      p.aux_in_smem = true;
      p.table_rep_log2 = c->table_rep_max;
      while (p.table_rep_log2 > 0 && sweep_smem_bytes(level, p.table_rep_log2, true) > c->max_smem) --p.table_rep_log2;
    }
    p.bytes = sweep_smem_bytes(level, p.table_rep_log2, p.aux_in_smem);
    p.def_off = p.bytes - 128;
    if (p.bytes + c->smem_pad <= c->max_smem) p.bytes += c->smem_pad;
    if (p.bytes > c->max_smem) p.smem = false;
  }
  if (!p.smem) {
    p.k = std::max(1, std::min(c->kmax_a, 17 - level));
    p.copies_log2 = 0;
    p.table_rep_log2 = 0;
    const size_t tb = ((size_t)1 << (level > 0 ? level - 1 : 0)) * (sizeof(float4) + 2 * sizeof(float)) +
                      ((((size_t)2 << level) + 15) & ~(size_t)15);
    p.table_in_smem = tb <= 64 * 1024;
    p.aux_in_smem = p.table_in_smem;
    p.bytes = p.table_in_smem ? tb : 0;
  }
  return p;
}

// NVTX ranges around the whole call and around every pass as it is enqueued: the counterpart of
// the reference's tracing spans (recursive_bisection.rs:467-468 "rcb_split", :590-595 "rcb_recurse").
// The kernels of a pass run asynchronously; a timeline tool projects the range onto the stream.
struct NvtxRange {
  template <class... A>
  explicit NvtxRange(const char *fmt, A... a) {
    char buf[96];
    snprintf(buf, sizeof buf, fmt, a...);
    nvtxRangePushA(buf);
  }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
};

struct Run {
  coupe_b200_ctx *c;
  cudaStream_t st;
  coupe_b200_stats &s;
  void launched(int n = 1) { s.kernel_launches += n; }
  void allreduce(void *buf, size_t count, ncclDataType_t dt, ncclRedOp_t op) {
    if (c->world <= 1) return;
    NC(g_nccl.AllReduce(buf, buf, count, dt, op, c->comm, st));
    s.collectives += 1;
  }
  void sync() {
    CU(cudaStreamSynchronize(st));
    s.host_syncs += 1;
  }
};

int run_impl(coupe_b200_ctx *c, bool rib, cudaStream_t st, uint64_t *part_dev, uintptr_t dim,
             uintptr_t n, const double *pts, int wtype, const void *w_dev, const void *wconst_host,
             uintptr_t iter_count, double tolerance, const cb_engine::Prefilled *pre = nullptr,
             int *compact_id_bytes = nullptr) {
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype < 0 || wtype > 2) return COUPE_ERR_BAD_TYPE;
  if (n > 0 && !w_dev && !wconst_host) return COUPE_ERR_CRASH;
  if (iter_count > (uintptr_t)MAX_LEVELS) return COUPE_ERR_ALLOC;  // tables of 2^iter_count nodes (see MAX_LEVELS)
  if (n >= ((uintptr_t)1 << 32) * 4) return COUPE_ERR_ALLOC;       // more points than one GPU's HBM holds
  CU(cudaSetDevice(c->device));
  prepare_funcs(c);
  NvtxRange call_range("coupe_b200 %s: %zu points, %d levels", rib ? "rib" : "rcb", (size_t)n, (int)iter_count);
  const int D = (int)dim, L = (int)iter_count;
  c->stats = coupe_b200_stats{};
  coupe_b200_stats &S = c->stats;
  S.n_local = n;
  S.levels = (uint32_t)L;
  Run R{c, st, S};

  // ---- scratch -------------------------------------------------------------
  const size_t npad = ((n + 3) / 4) * 4 + 4;
  c->xcols.ensure(npad * sizeof(float) * 3);
  c->ids.ensure(npad * sizeof(uint32_t));
  const size_t nb_smem = (size_t)1 << c->nb_smem_log2;
  c->part_w.ensure((size_t)c->num_sms * nb_smem * 8);
  c->part_min.ensure((size_t)c->num_sms * nb_smem * 4);
  size_t hist_entries = std::max<size_t>(nb_smem, (size_t)1 << 17);
  if (L > 0) hist_entries = std::max<size_t>(hist_entries, (size_t)2 << (L - 1));
  c->hist_w.ensure(hist_entries * 8);
  c->hist_min.ensure(hist_entries * 4);
  const size_t max_nodes = (size_t)1 << (L > 0 ? L - 1 : 0);
  c->nodes_a.ensure(max_nodes * sizeof(NodeState));
  c->nodes_b.ensure(max_nodes * sizeof(NodeState));
  c->table_a.ensure(max_nodes * sizeof(float4));
  c->table_b.ensure(max_nodes * sizeof(float4));
  c->thi_a.ensure(max_nodes * sizeof(float));
  c->thi_b.ensure(max_nodes * sizeof(float));
  c->rtable.ensure(max_nodes * sizeof(float4));
  c->tsp_a.ensure(max_nodes * sizeof(float));
  c->tsp_b.ensure(max_nodes * sizeof(float));
  c->nsh_a.ensure(max_nodes * 2 * sizeof(short));
  c->nsh_b.ensure(max_nodes * 2 * sizeof(short));
  c->target.ensure(max_nodes * sizeof(uint32_t));
  c->target_first.ensure(max_nodes * sizeof(uint32_t));
  c->node_rt.ensure(max_nodes * sizeof(uint2));
  c->rfast.ensure(max_nodes * sizeof(float2));
  c->gp.ensure(sizeof(GlobalParams));
  c->mom_partial.ensure((size_t)c->num_sms * 8 * 16 * sizeof(double));
  const size_t tr_n = L > 0 ? ((size_t)1 << L) - 1 : 1;
  Trace tr{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (c->trace_on) {
    c->tr_visited.ensure(tr_n);
    c->tr_split.ensure(tr_n * 4);
    c->tr_wl.ensure(tr_n * 8);
    c->tr_sum.ensure(tr_n * 8);
    c->tr_iters.ensure(tr_n * 4);
    CU(cudaMemsetAsync(c->tr_visited.p, 0, tr_n, st));
    tr = Trace{c->tr_visited.as<uint8_t>(), c->tr_split.as<float>(), c->tr_wl.as<double>(),
               c->tr_sum.as<double>(), c->tr_iters.as<uint32_t>()};
    c->trace_levels = (uint32_t)L;
  }
  GlobalParams *gp = c->gp.as<GlobalParams>();
  float *x[3] = {c->xcols.as<float>(), c->xcols.as<float>() + npad, c->xcols.as<float>() + 2 * npad};
  void *ids = c->ids.p;

  const size_t ngroups = (n + 3) / 4;
  const int sweep_grid =
      std::max(1, (int)std::min<size_t>((size_t)c->num_sms, (ngroups + SWEEP_THREADS - 1) / SWEEP_THREADS));
  // points one block of a dense sweep processes at most (grid-stride over groups of four, plus the tail)
  const size_t block_points = ((n / 4) / ((size_t)sweep_grid * SWEEP_THREADS) + 1) * SWEEP_THREADS * 4 + 4;
  // Deferred points (rcb_kernels.cuh): per-block lists with room for every point the block sweeps, so a list
  // cannot overflow whatever share of a node sits in the bin under refinement.  20 bytes per point of scratch;
  // without it (allocation failure, option "defer" 0) undecided levels are refined by scanning the idx words.
  const size_t warp_points = block_points / DEFER_LISTS + 4;  // ... and one of its warps (every thread runs the same number of groups, +-1)
  bool defer_call = c->defer_opt && L >= 2 && n < ((size_t)1 << 32);
  if (defer_call) {
    try {
      c->def_rec.ensure((size_t)sweep_grid * DEFER_LISTS * warp_points * sizeof(uint4));
      c->def_slot.ensure((size_t)sweep_grid * DEFER_LISTS * warp_points * sizeof(uint32_t));
      c->def_count.ensure((size_t)c->num_sms * DEFER_LISTS * sizeof(uint32_t));
      c->rpart_w.ensure((size_t)c->num_sms * nb_smem * 8);
      c->rpart_min.ensure((size_t)c->num_sms * nb_smem * 4);
    } catch (const CudaFail &f) {
      if (f.err != cudaErrorMemoryAllocation) throw;
      cudaGetLastError();
      c->def_rec.release();
      c->def_slot.release();
      defer_call = false;
    }
  }

  // ---- global point count ---------------------------------------------------
  unsigned long long n_global = n;
  bool any_rank_has_array_weights = w_dev != nullptr;
  if (c->world > 1) {  // a rank with an empty shard may not know whether weights are per point
    unsigned long long *d = reinterpret_cast<unsigned long long *>(c->hist_w.p);
    unsigned long long h[3] = {n_global, w_dev ? 1ull : 0ull, defer_call ? 1ull : 0ull};
    CU(cudaMemcpyAsync(d, h, 24, cudaMemcpyHostToDevice, st));
    R.allreduce(d, 3, ncclUint64, ncclSum);
    CU(cudaMemcpyAsync(c->h_pinned, d, 24, cudaMemcpyDeviceToHost, st));
    R.sync();
    memcpy(h, c->h_pinned, 24);
    n_global = h[0];
    any_rank_has_array_weights = h[1] != 0;
    defer_call = h[2] == (unsigned long long)c->world;  // every rank must run the same pass sequence
  }
  S.n_global = n_global;
  // i32 column read by the sweeps below the root: narrowed f64 / i64 weights, or an aligned copy
  // of caller's i32 weights that are not 16-byte aligned.  Decided on the GLOBAL "weights are an
  // array" fact: a rank with an empty shard must take the same collective steps as the others.
  const bool narrow_w = any_rank_has_array_weights &&
                        (wtype == WT_F64 || (wtype == WT_I64 && L > 1) ||
                         (wtype == WT_I32 && L > 1 && w_dev && ((uintptr_t)w_dev % 16) != 0));
  if (narrow_w) c->w32.ensure(npad * sizeof(int));
  if (n_global == 0) return COUPE_ERR_OK;  // BoundingBox::from_points -> None (:685-688)
  const int id_bytes = L <= 16 ? 2 : 4;  // compact ids of the host path
  if (compact_id_bytes) *compact_id_bytes = id_bytes;
  if (L == 0) {                            // iter_count == 0: every id is 0
    if (n && compact_id_bytes) CU(cudaMemsetAsync(c->host_ids.p, 0, n * id_bytes, st));
    else if (n) CU(cudaMemsetAsync(part_dev, 0, n * sizeof(uint64_t), st));
    R.sync();
    return COUPE_ERR_OK;
  }

  // ---- RIB: principal axis, obb_to_aabb matrix ------------------------------
  Mat3 rot{};
  if (rib) {
    const int mgrid = std::max(1, (int)std::min<size_t>((size_t)c->num_sms * 8, (n + 255) / 256));
    double *partial = c->mom_partial.as<double>();
    double *dsums = reinterpret_cast<double *>(c->hist_w.p);
    double hs[16];
    auto reduce_to_host = [&](int nv) {
      moments_final_kernel<<<1, 32, 0, st>>>(partial, mgrid, nv, dsums);
      R.launched();
      R.allreduce(dsums, nv, ncclFloat64, ncclSum);
      CU(cudaMemcpyAsync(c->h_pinned, dsums, nv * 8, cudaMemcpyDeviceToHost, st));
      R.sync();
      memcpy(hs, c->h_pinned, nv * 8);
    };
    if (D == 2) moments_partial_kernel<2, false><<<mgrid, 256, 0, st>>>(pts, n, partial, 0, 0, 0);
    else moments_partial_kernel<3, false><<<mgrid, 256, 0, st>>>(pts, n, partial, 0, 0, 0);
    R.launched();
    reduce_to_host(D);
    double cen[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) cen[d] = hs[d] / (double)n_global;  // geometry.rs:274-275
    if (D == 2)
      moments_partial_kernel<2, true><<<mgrid, 256, 0, st>>>(pts, n, partial, cen[0], cen[1], 0);
    else
      moments_partial_kernel<3, true><<<mgrid, 256, 0, st>>>(pts, n, partial, cen[0], cen[1], cen[2]);
    R.launched();
    reduce_to_host(D * D);
    double v[3], h[9], inv[9];
    principal_axis(D, hs, v);
    reflection(D, v, h);
    if (!inverse(D, h, inv)) return COUPE_ERR_CRASH;  // `.unwrap()` on try_inverse
    for (int k = 0; k < D * D; ++k) rot.m[k] = S.matrix[k] = inv[k];
  }

  // ---- prologue: narrow, bbox, max |w| --------------------------------------
  {
    GlobalParams init{};
    for (int k = 0; k < 8; ++k) init.bbox_keys[k] = KEY_EMPTY;
    init.leaf_min = 0xFFFFFFFFu;
    static_assert(sizeof(GlobalParams) % 4 == 0, "");
    CU(cudaMemcpyAsync(gp, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    const size_t ngroups = (n + 3) / 4;
    const int grid = std::max(1, (int)std::min<size_t>((size_t)c->num_sms * 8, (ngroups + 255) / 256));
    const double *wf = (wtype == WT_F64 && w_dev) ? static_cast<const double *>(w_dev) : nullptr;
    const int ws = c->sample_w_opt ? 1 : 0;
    const int pa = ((uintptr_t)pts % 16) == 0, wa = ((uintptr_t)w_dev % 16) == 0;
    if (pre) {
      // host path: the columns were filled from the host (narrowed there), the box came with them
      memcpy(c->h_pinned + 16, pre->bbox_keys, sizeof(pre->bbox_keys));
      CU(cudaMemcpyAsync(gp->bbox_keys, c->h_pinned + 16, sizeof(pre->bbox_keys), cudaMemcpyHostToDevice, st));
      if (wf && n) {
        const size_t items = ws ? ((ngroups + 16383) / 16384) * 256 : ngroups;
        wsample_kernel<<<std::max(1, (int)std::min<size_t>((size_t)c->num_sms * 8, (items + 255) / 256)), 256, 0, st>>>(wf, n, gp, ws);
        R.launched();
      }
    } else if (D == 2) {
      if (rib) narrow_kernel<2, true><<<grid, 256, 0, st>>>(pts, n, x[0], x[1], x[2], rot, gp, wf, pa, wa, ws);
      else narrow_kernel<2, false><<<grid, 256, 0, st>>>(pts, n, x[0], x[1], x[2], rot, gp, wf, pa, wa, ws);
    } else {
      if (rib) narrow_kernel<3, true><<<grid, 256, 0, st>>>(pts, n, x[0], x[1], x[2], rot, gp, wf, pa, wa, ws);
      else narrow_kernel<3, false><<<grid, 256, 0, st>>>(pts, n, x[0], x[1], x[2], rot, gp, wf, pa, wa, ws);
    }
    if (!pre) R.launched();
    if (c->world > 1) {
      R.allreduce(gp->bbox_keys, 8, ncclUint32, ncclMin);
      // every statistic is kept as a maximum; decided on the GLOBAL "weights are an array" fact (empty shards)
      if (wtype == WT_F64 && any_rank_has_array_weights) R.allreduce(gp->wstat_sample, WS_N, ncclUint64, ncclMax);
    }
  }
  long long wconst_i = 1;
  double wconst_f = 0.0;
  const int w_is_const = !any_rank_has_array_weights;
  if (w_is_const) {
    if (!wconst_host) return COUPE_ERR_CRASH;
    if (wtype == WT_I32) wconst_i = *static_cast<const int *>(wconst_host);
    else if (wtype == WT_I64) wconst_i = *static_cast<const long long *>(wconst_host);
    else wconst_f = *static_cast<const double *>(wconst_host);
  }
  NodeState *cur = c->nodes_a.as<NodeState>(), *nxt = c->nodes_b.as<NodeState>();
  float4 *tab_cur = c->table_a.as<float4>(), *tab_next = c->table_b.as<float4>();
  float *thi_cur = c->thi_a.as<float>(), *thi_next = c->thi_b.as<float>();
  float *tsp_cur = c->tsp_a.as<float>(), *tsp_next = c->tsp_b.as<float>();
  short *nsh_cur = c->nsh_a.as<short>(), *nsh_next = c->nsh_b.as<short>();
  float4 *rtable = c->rtable.as<float4>();
  uint32_t *target = c->target.as<uint32_t>();
  uint32_t *target_first = c->target_first.as<uint32_t>();
  // per-point f64 weights: the fixed-point form and scale come from statistics of a sample of the weights;
  // the root sweep computes the true ones and the walk of the root asks for another root pass when they differ
  const bool verify_form = wtype == WT_F64 && !w_is_const;
  auto enqueue_init_root = [&](int reuse) {
    init_root_kernel<<<1, 1, 0, st>>>(gp, cur, tab_cur, thi_cur, nsh_cur, plan_first(c, 0).k, D, wtype,
                                      w_is_const, wconst_i, wconst_f, n_global, reuse);
    R.launched();
  };
  enqueue_init_root(0);

  unsigned long long *hist_w = c->hist_w.as<unsigned long long>();
  uint32_t *hist_min = c->hist_min.as<uint32_t>();
  int *w32 = narrow_w ? c->w32.as<int>() : nullptr;

  // weight column read by the sweeps: the caller's array at the root level, the
  // narrowed i32 copy afterwards when there is one
  int win = WIN_CONST;
  const void *wp = nullptr;
  if (!w_is_const) {
    win = wtype == WT_I32 ? WIN_I32 : wtype == WT_I64 ? WIN_I64 : WIN_F64;
    wp = w_dev;
  }

  size_t ev_used = 0;
  int last_first_pass_event = -1;  // index (event_kind) of the dense sweep enqueued last, when it is being timed
  c->event_kind.clear();
  c->event_level.clear();
  c->sweep_ms.clear();
  bool timing_open = false;
  auto time_begin = [&](int kind, int level) {  // option time_sweeps: 1 = dense sweeps, 2 = refinement sweeps too, 3 = and print
    timing_open = c->time_sweeps > kind;
    if (!timing_open) return;
    while (c->events.size() < ev_used + 2) {
      cudaEvent_t e;
      CU(cudaEventCreate(&e));
      c->events.push_back(e);
    }
    c->event_kind.push_back(kind);
    c->event_level.push_back(level);
    CU(cudaEventRecord(c->events[ev_used], st));
  };
  auto time_end = [&]() {
    if (!timing_open) return;
    CU(cudaEventRecord(c->events[ev_used + 1], st));
    ev_used += 2;
  };
  uint2 *node_rt = c->node_rt.as<uint2>();
  // every level in shared-memory mode keeps (node << k) + bin below 2^16
  bool idx16 = true;
  for (int level = 0; level < L; ++level) {
    const FirstPlan pl = plan_first(c, level);
    idx16 = idx16 && pl.smem && level + pl.k <= 16;
  }
  float2 *rfast = c->rfast.as<float2>();
  // multi-GPU: histograms travel through the peer-memory exchange when every pass of the call fits
  // its slots (shared-memory mode at every level), else through NCCL all-reduces
  bool use_xchg = c->world > 1 && c->xchg_ok && c->use_xchg_opt;
  for (int level = 0; level < L; ++level) use_xchg = use_xchg && plan_first(c, level).smem;
  auto make_xchg = [&](uint64_t pass_seq) {
    Xchg x{};
    if (!use_xchg) return x;
    for (int r = 0; r < c->world; ++r) x.peer[r] = c->xchg_peer[r];
    x.world = c->world;
    x.rank = c->rank;
    x.slot = (uint32_t)(pass_seq % XCHG_DEPTH);
    x.seq = pass_seq + 1;
    x.ticket = c->xchg_aux;
    x.error = c->xchg_aux + 1;
    return x;
  };
  auto allreduce_hist = [&](uint32_t count) {
    if (use_xchg) return;  // pushed by reduce_partials_kernel, gathered by walk_kernel
    R.allreduce(hist_w, count, ncclUint64, ncclSum);
    R.allreduce(hist_min, count, ncclUint32, ncclMin);
  };
  // histogram slots a refinement pass may use at a level (shared memory next to the match queue
  // and, when it fits, the per-node target table)
  auto refine_cap = [&](int level, bool &rts, size_t &rt_bytes) {
    rt_bytes = ((size_t)1 << level) * sizeof(uint32_t);
    rts = rt_bytes <= 64 * 1024;
    const size_t fixed = (size_t)REFINE_QBYTES + (rts ? rt_bytes : 0) + 64;
    const size_t room = c->carve_fit ? SMEM_CARVE_196 : c->max_smem;  // stay under the 196 KB carve-out (see SMEM_CARVE_196)
    return (uint32_t)std::min<size_t>((size_t)1 << c->nb_smem_log2, (room - fixed) / 12);
  };
  // ---- level loop -----------------------------------------------------------------
  // No stream synchronisation inside: the rank kernel that ends every pass writes the
  // number of undecided nodes to mapped pinned host memory and the host polls that word.
  // The first pass of level l+1 (and the final emit) is enqueued BEFORE the host knows
  // whether level l needs refinement; its kernels return at once when it does
  // (gp->unresolved != 0) and the pass is enqueued again after the refinement.
  const uint32_t *guard_ptr = &gp->unresolved;
  uint32_t w_wide = 0, rescale = 0, f64_wide = 0;
  uint64_t seq = c->flag_seq;
  // whatever way the call ends (error returns and exceptions included) the next call must not reuse the
  // sequence numbers of passes already enqueued: host flags and exchange slots still carry them
  struct SeqGuard {
    uint64_t &dst;
    const uint64_t &src;
    ~SeqGuard() { dst = src; }
  } seq_guard{c->flag_seq, seq};
  auto flag_slot = [&](uint64_t s) { return s % FLAG_SLOTS; };
  auto wait_flag = [&](uint64_t s) -> uint32_t {
    volatile unsigned long long *f = c->h_flags + flag_slot(s);
    unsigned long long v;
    for (uint64_t spin = 0; (v = *f) == 0; ++spin) {
      if ((spin & 0xFFF) == 0xFFF) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q != cudaSuccess && q != cudaErrorNotReady) throw CudaFail{q, "cudaStreamQuery (flag wait)"};
        if (q == cudaSuccess && *f == 0) throw CudaFail{cudaErrorUnknown, "pass ended without its flag"};
      }
    }
    S.flag_waits += 1;
    if (v & FLAG_ABORTED) throw CudaFail{cudaErrorUnknown, "waited on an aborted pass"};
    w_wide = (uint32_t)(v >> 32) & 1u;
    rescale = (uint32_t)(v >> 33) & 1u;
    f64_wide = (uint32_t)(v >> 34) & 1u;
    return (uint32_t)v;
  };
  // walk + rank of one pass; returns the sequence number of its flag
  auto enqueue_walk = [&](int level, int k, int k0, int first, uint32_t rank_limit, const uint32_t *guard) {
    bool rts_;
    size_t rtb_;
    const uint64_t s = seq++;
    c->h_flags[flag_slot(s)] = 0;
    WalkArgs wa{cur, nxt, hist_w, hist_min, gp, tab_next, thi_next, tsp_next, nsh_next, target, target_first, node_rt, rtable,
                tr, tolerance, level, k, D, first, level == L - 1, w_is_const, k0, rank_limit, guard,
                plan_first(c, level + 1).k, rfast, refine_cap(level, rts_, rtb_), c->kmax_refine,
                verify_form ? 1 : 0, (unsigned long long)block_points,
                c->d_flags + flag_slot(s), make_xchg(s)};
    const size_t bytes = ((size_t)2 << k) * 12;
    const uint32_t nodes = 1u << level;
    // the last block to finish ranks the undecided nodes and reports to the host flag
    if (wtype == WT_I32) launch_pdl(walk_kernel<WT_I32>, nodes, WALK_THREADS, bytes, st, wa);
    else if (wtype == WT_I64) launch_pdl(walk_kernel<WT_I64>, nodes, WALK_THREADS, bytes, st, wa);
    else launch_pdl(walk_kernel<WT_F64>, nodes, WALK_THREADS, bytes, st, wa);
    R.launched(1);
    return s;
  };
  // the dense sweep of `level`; kprev = bins of the level before; defer: points whose parent is still
  // undecided are listed instead of binned (the sweep then needs no guard)
  auto enqueue_dense = [&](int level, int kprev, const uint32_t *guard, bool defer) {
    NvtxRange range("rcb level %d: dense sweep%s", level, guard ? " (optimistic)" : defer ? " (deferring)" : "");
    const int axis = level % D, prev_axis = (level + D - 1) % D;
    const FirstPlan plan = plan_first(c, level);
    const int k = plan.k;
    const uint32_t nb = 1u << (level + k);
    SweepArgs sa{};
    sa.n = n;
    sa.x = x[axis];
    sa.xp = x[prev_axis];
    sa.idx = ids;
    sa.w = wp;
    sa.w32_out = level == 0 ? w32 : nullptr;
    sa.gp = gp;
    sa.guard = guard;
    sa.table = tab_cur;
    sa.table_hi = thi_cur;
    sa.table_split = tsp_cur;
    sa.nshift = nsh_cur;
    sa.part_w = c->part_w.as<long long>();
    sa.part_min = c->part_min.as<uint32_t>();
    sa.hist_w = hist_w;
    sa.hist_min = hist_min;
    sa.level = level;
    sa.k = k;
    sa.kprev = kprev;
    sa.copies_log2 = plan.copies_log2;
    sa.w_vec = ((uintptr_t)wp % 16) == 0;
    sa.one = 1;
    sa.table_rep_log2 = plan.table_rep_log2;
    sa.aux_in_smem = plan.aux_in_smem ? 1 : 0;
    if (defer) {
      sa.def_rec = c->def_rec.as<uint4>();
      sa.def_slot = c->def_slot.as<uint32_t>();
      sa.def_count = c->def_count.as<uint32_t>();
      sa.def_seg = (uint32_t)warp_points;
    }
    sa.def_smem_off = (uint32_t)plan.def_off;
    if (!plan.smem) {
      launch_pdl(fill_hist_kernel, (nb + 255) / 256, 256, 0, st, hist_w, hist_min, nb, guard);
      R.launched();
    }
    time_begin(0, level);
    last_first_pass_event = timing_open ? (int)c->event_kind.size() - 1 : -1;
    launch_sweep_any(win, level == 0, plan.smem, plan.table_in_smem, idx16, sweep_grid, plan.bytes, st, sa);
    time_end();
    R.launched();
    S.dense_sweeps += 1;
  };
  // partial histograms of the dense sweep of `level` -> level histogram -> walk
  auto enqueue_reduce_walk = [&](int level, const uint32_t *guard) {
    const FirstPlan plan = plan_first(c, level);
    const uint32_t nb = 1u << (level + plan.k);
    if (plan.smem) {
      launch_pdl(reduce_partials_kernel, (nb + 31) / 32, 256, 0, st, (const long long *)c->part_w.as<long long>(),
                 (const uint32_t *)c->part_min.as<uint32_t>(), sweep_grid, nb, hist_w, hist_min,
                 guard, make_xchg(seq));
      R.launched();
    }
    allreduce_hist(nb);
    if (level == 0 && verify_form) R.allreduce(gp->wstat_true, WS_N, ncclUint64, ncclMax);
    return enqueue_walk(level, plan.k, plan.k, 1, 0, guard);
  };
  auto enqueue_first_pass = [&](int level, int kprev, const uint32_t *guard) {
    enqueue_dense(level, kprev, guard, false);
    return enqueue_reduce_walk(level, guard);
  };
  // --- deferred points: the level below an undecided one has listed the points of the undecided bins ---
  // The deferring variant of the sweep costs the levels that have nothing to defer ~6 % (rcb_kernels.cuh: DEFER), so
  // it is launched only below a level that is EXPECTED to stay undecided: one that resolves fewer candidates per
  // pass than the first levels do, or one that the previous call on this context left undecided (the calls of a
  // context are collective, so every rank predicts alike).  A level that stays undecided against the prediction is
  // refined by rescanning; the results do not depend on the choice.
  std::vector<uint8_t> undecided_now((size_t)L, 0);
  auto defer_possible = [&](int level_next) {
    return defer_call && level_next < L && plan_first(c, level_next).smem && (win == WIN_I32 || win == WIN_CONST);
  };
  auto can_defer = [&](int level_next) {  // ... and predicted to pay
    if (!defer_possible(level_next)) return false;
    if (c->defer_opt >= 2) return true;
    const int lv = level_next - 1;
    const bool few_bits = plan_first(c, lv).k < plan_first(c, 0).k;
    const bool last_time = (int)c->undecided_last.size() == L && c->undecided_last[(size_t)lv] != 0;
    return few_bits || last_time;
  };
  auto defer_args = [&](int level_next) {
    DeferArgs d{};
    d.rec = c->def_rec.as<uint4>();
    d.slot0 = c->def_slot.as<uint32_t>();
    d.count = c->def_count.as<uint32_t>();
    d.seg = (uint32_t)warp_points;
    d.kslot = plan_first(c, level_next).k;
    d.node_rt = node_rt;
    d.rtable = rtable;
    d.rfast = rfast;
    d.part_w = c->rpart_w.as<long long>();
    d.part_min = c->rpart_min.as<uint32_t>();
    d.one = 1;
    d.gp = gp;
    d.idx = ids;
    d.target_first = target_first;
    return d;
  };
  // one refinement pass of `level` over the list written by the dense sweep of level + 1
  auto enqueue_defer_refine_round = [&](int level, int k0, uint32_t unresolved) {
    NvtxRange range("rcb level %d: refinement pass over the deferred points, %u undecided nodes", level, unresolved);
    bool rts;
    size_t rt_bytes;
    const uint32_t cap = refine_cap(level, rts, rt_bytes);  // (the walk ranks with the same capacity)
    const int kr = refine_bits(unresolved, cap, c->kmax_refine);
    const uint32_t limit = std::min<uint32_t>(unresolved, cap >> kr);
    DeferArgs da = defer_args(level + 1);
    da.nslots = limit << kr;
    da.rank_limit = limit;
    da.k = kr;
    time_begin(1, level);
    if (win == WIN_I32) launch_pdl(defer_refine_kernel<WIN_I32>, sweep_grid, SWEEP_THREADS, (size_t)da.nslots * 12, st, da);
    else launch_pdl(defer_refine_kernel<WIN_CONST>, sweep_grid, SWEEP_THREADS, (size_t)da.nslots * 12, st, da);
    time_end();
    launch_pdl(reduce_partials_kernel, (da.nslots + 31) / 32, 256, 0, st, (const long long *)da.part_w,
               (const uint32_t *)da.part_min, sweep_grid, da.nslots, hist_w, hist_min, (const uint32_t *)nullptr,
               make_xchg(seq));
    R.launched(2);
    S.refine_sweeps += 1;
    S.list_refine_sweeps += 1;
    allreduce_hist(da.nslots);
    return enqueue_walk(level, kr, k0, 0, limit, nullptr);
  };
  // every split of the level above `level_next` is decided (tab_cur / tsp_cur are final): the deferred
  // points take their child and join level_next's partial histograms
  auto enqueue_fixup = [&](int level_next) {
    NvtxRange range("rcb level %d: deferred points take their child", level_next);
    const FirstPlan plan = plan_first(c, level_next);
    const uint32_t nb = 1u << (level_next + plan.k);
    DeferArgs da = defer_args(level_next);
    da.table = tab_cur;
    da.table_split = tsp_cur;
    da.row_w = c->part_w.as<unsigned long long>();
    da.row_min = c->part_min.as<uint32_t>();
    da.row_stride = nb;
    if (win == WIN_I32) {
      if (idx16) launch_pdl(defer_fixup_kernel<WIN_I32, uint16_t>, sweep_grid, SWEEP_THREADS, 0, st, da);
      else launch_pdl(defer_fixup_kernel<WIN_I32, uint32_t>, sweep_grid, SWEEP_THREADS, 0, st, da);
    } else {
      if (idx16) launch_pdl(defer_fixup_kernel<WIN_CONST, uint16_t>, sweep_grid, SWEEP_THREADS, 0, st, da);
      else launch_pdl(defer_fixup_kernel<WIN_CONST, uint32_t>, sweep_grid, SWEEP_THREADS, 0, st, da);
    }
    R.launched();
  };
  auto enqueue_refine_round = [&](int level, int k0, uint32_t unresolved) {
    NvtxRange range("rcb level %d: refinement pass, %u undecided nodes", level, unresolved);
    // ranked histograms of as many undecided nodes as shared memory holds, 2^kr bins each
    const int axis = level % D;
    bool rts;
    size_t rt_bytes;
    const uint32_t cap = refine_cap(level, rts, rt_bytes);
    const int kr = refine_bits(unresolved, cap, c->kmax_refine);
    const uint32_t limit = std::min<uint32_t>(unresolved, cap >> kr);
    const uint32_t nslots = limit << kr;
    const size_t rbytes = (size_t)REFINE_QBYTES + (size_t)nslots * 12 + (rts ? rt_bytes : 0);
    RefineArgs ra{n, x[axis], ids, wp, node_rt, rtable, rfast, c->part_w.as<long long>(),
                  c->part_min.as<uint32_t>(), nslots, limit, level, kr, k0, rts, 1u, gp, nsh_cur};
    time_begin(1, level);
    switch (win) {
      case WIN_I32: launch_refine<WIN_I32>(idx16, rts, sweep_grid, rbytes, st, ra); break;
      case WIN_I64: launch_refine<WIN_I64>(idx16, rts, sweep_grid, rbytes, st, ra); break;
      case WIN_F64: launch_refine<WIN_F64>(idx16, rts, sweep_grid, rbytes, st, ra); break;
      default: launch_refine<WIN_CONST>(idx16, rts, sweep_grid, rbytes, st, ra); break;
    }
    time_end();
    launch_pdl(reduce_partials_kernel, (nslots + 31) / 32, 256, 0, st, ra.part_w, ra.part_min, sweep_grid, nslots,
               hist_w, hist_min, (const uint32_t *)nullptr, make_xchg(seq));
    R.launched(2);
    S.refine_sweeps += 1;
    allreduce_hist(nslots);
    return enqueue_walk(level, kr, k0, 0, limit, nullptr);
  };
  auto advance_level = [&]() {
    std::swap(cur, nxt);
    std::swap(tab_cur, tab_next);
    std::swap(thi_cur, thi_next);
    std::swap(tsp_cur, tsp_next);
    std::swap(nsh_cur, nsh_next);
  };
  auto enqueue_emit = [&](int klast, const uint32_t *guard) {
    const int grid = std::max(1, (int)std::min<size_t>((size_t)c->num_sms * 4, (ngroups + 511) / 512));
    unsigned long long *out = reinterpret_cast<unsigned long long *>(part_dev);
    const int out_vec = ((uintptr_t)part_dev % 32) == 0 ? 2 : ((uintptr_t)part_dev % 16) == 0 ? 1 : 0;
    const float *xl = x[(L - 1) % D];
    if (compact_id_bytes && id_bytes == 2) {  // (L <= 16 implies 16-bit idx words)
      launch_pdl(emit_kernel<uint16_t, uint16_t>, grid, 512, 0, st, n, ids, xl, tab_cur, tsp_cur, klast, gp, c->host_ids.as<uint16_t>(), 0, guard);
    } else if (compact_id_bytes) {
      if (idx16) launch_pdl(emit_kernel<uint16_t, uint32_t>, grid, 512, 0, st, n, ids, xl, tab_cur, tsp_cur, klast, gp, c->host_ids.as<uint32_t>(), 0, guard);
      else launch_pdl(emit_kernel<uint32_t, uint32_t>, grid, 512, 0, st, n, ids, xl, tab_cur, tsp_cur, klast, gp, c->host_ids.as<uint32_t>(), 0, guard);
    } else if (idx16) {
      launch_pdl(emit_kernel<uint16_t, unsigned long long>, grid, 512, 0, st, n, ids, xl, tab_cur, tsp_cur, klast, gp, out, out_vec, guard);
    } else {
      launch_pdl(emit_kernel<uint32_t, unsigned long long>, grid, 512, 0, st, n, ids, xl, tab_cur, tsp_cur, klast, gp, out, out_vec, guard);
    }
    R.launched();
  };

  // cur/tab_cur always describe the level whose pass is enqueued next
  uint64_t pending = enqueue_first_pass(0, 0, nullptr);  // flag of the pass of `level`
  for (int level = 0; level < L; ++level) {
    const int k = plan_first(c, level).k;
    // i64 weights: which column the later sweeps read is only known after the root pass
    const bool can_speculate = !(level == 0 && ((w32 && wtype == WT_I64) || verify_form));
    advance_level();
    bool speculated = false;
    bool spec_deferred = false;  // the dense sweep of level + 1 is enqueued, and it defers
    uint64_t next_pending = 0;
    int win_root = win;
    const void *wp_root = wp;
    if (level == 0 && w32 && wtype != WT_I64) {  // f64 (and unaligned i32) weights: always narrowed
      win = WIN_I32;
      wp = w32;
    }
    if (can_speculate) {
      if (level + 1 < L) {
        spec_deferred = can_defer(level + 1);
        if (spec_deferred) {
          // the sweep runs whatever the walk of this level says; only its reduce + walk are optimistic
          enqueue_dense(level + 1, k, nullptr, true);
          next_pending = enqueue_reduce_walk(level + 1, guard_ptr);
        } else {
          next_pending = enqueue_first_pass(level + 1, k, guard_ptr);
        }
      } else {
        enqueue_emit(k, guard_ptr);
      }
      speculated = true;
    }
    uint32_t unresolved = wait_flag(pending);
    undecided_now[(size_t)level] = unresolved > 0 ? 1 : 0;
    for (int redo = 0; level == 0 && rescale; ++redo) {
      // the sample of the weights gave another fixed-point form or scale than all of them do (an
      // outlier, a negative or a tiny weight outside the sample), or the wide form wants a finer root
      // shift: the walk left the parameters to use in GlobalParams; redo the root pass with them.
      // Every rank sees the same flag.
      if (redo >= 3) return COUPE_ERR_CRASH;  // cannot happen: sample -> true statistics -> finer root shift
      advance_level();  // back to the root's tables
      enqueue_init_root(1);
      S.dense_sweeps -= 1;
      S.weight_rescales += 1;
      std::swap(win, win_root);  // the root sweep reads the caller's f64 column
      std::swap(wp, wp_root);
      pending = enqueue_first_pass(0, 0, nullptr);
      std::swap(win, win_root);
      std::swap(wp, wp_root);
      advance_level();
      unresolved = wait_flag(pending);
    }
    if (level == 0 && verify_form && f64_wide) {  // wide form: the sweeps keep reading the caller's f64 weights
      win = WIN_F64;
      wp = w_dev;
    }
    if (level == 0 && w32 && wtype == WT_I64) {
      if (c->world > 1) {  // every rank must take the same path
        R.allreduce(&gp->w_wide, 1, ncclUint32, ncclMax);
        CU(cudaMemcpyAsync(c->h_pinned, &gp->w_wide, 4, cudaMemcpyDeviceToHost, st));
        R.sync();
        w_wide = c->h_pinned[0];
      }
      if (!w_wide) {
        win = WIN_I32;
        wp = w32;
      }
    }
    if (unresolved > 0 && (speculated ? spec_deferred : defer_possible(level + 1))) {  // (not speculated: known undecided)
      // Deferred refinement: the dense sweep of level + 1 lists the points of the undecided bins; the
      // refinement passes of this level read that list, then the listed points take their child.
      if (!speculated) enqueue_dense(level + 1, k, nullptr, true);
      S.deferred_levels += 1;
      advance_level();  // back to this level's tables
      int rounds = 0;
      while (unresolved > 0) {
        if (++rounds > 100000) return COUPE_ERR_CRASH;  // cannot happen: f32 brackets shrink
        unresolved = wait_flag(enqueue_defer_refine_round(level, k, unresolved));
      }
      advance_level();
      enqueue_fixup(level + 1);
      next_pending = enqueue_reduce_walk(level + 1, nullptr);
      speculated = true;
    } else if (unresolved > 0) {
      // the optimistic pass (if any) returned at once on the device; refine this level, then redo it
      if (speculated && level + 1 < L) {
        S.dense_sweeps -= 1;
        if (last_first_pass_event >= 0) c->event_kind[last_first_pass_event] = 2;  // returned at once: not a sweep
      }
      advance_level();  // back to this level's tables
      int guard = 0;
      while (unresolved > 0) {
        if (++guard > 100000) return COUPE_ERR_CRASH;  // cannot happen: f32 brackets shrink
        unresolved = wait_flag(enqueue_refine_round(level, k, unresolved));
      }
      advance_level();
      speculated = false;
    }
    if (!speculated) {
      if (level + 1 < L) next_pending = enqueue_first_pass(level + 1, k, nullptr);
      else enqueue_emit(k, nullptr);
    }
    pending = next_pending;
  }
  c->undecided_last = undecided_now;
  CU(cudaMemcpyAsync(c->h_pinned + 2, &gp->refine_points, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_pinned, &gp->shift, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_pinned + 6, &gp->ec, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_pinned + 8, &gp->xchg_wait_cycles, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_pinned + 5, &gp->nocarry, 4, cudaMemcpyDeviceToHost, st));
  if (use_xchg) CU(cudaMemcpyAsync(c->h_pinned + 1, c->xchg_aux + 1, 4, cudaMemcpyDeviceToHost, st));
  R.sync();
  {
    int sh, ec;
    memcpy(&sh, c->h_pinned, 4);
    memcpy(&ec, c->h_pinned + 6, 4);
    S.weight_shift = wtype == WT_F64 ? sh - ec : 0;
    S.weight_wide = f64_wide;
    unsigned long long cyc;
    memcpy(&cyc, c->h_pinned + 8, 8);
    S.exchange_wait_ms = c->sm_khz > 0 ? (double)cyc / (double)c->sm_khz : 0.0;
  }
  S.peer_exchange = use_xchg ? 1 : 0;
  S.carry_free = c->h_pinned[5];
  memcpy(&S.refine_points, c->h_pinned + 2, 8);
  if (use_xchg && c->h_pinned[1] != 0) {
    fprintf(stderr, "coupe_b200: a rank did not deliver its histogram in time (peer-memory exchange)\n");
    return COUPE_ERR_CRASH;
  }
  for (size_t e = 0; e + 1 < ev_used; e += 2) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->events[e], c->events[e + 1]));
    const int kind = c->event_kind[e / 2];
    if (kind == 0) S.dense_sweep_ms += ms;
    if (kind == 1) S.refine_sweep_ms += ms;
    c->sweep_ms.push_back(ms);
    if (c->time_sweeps > 2)
      fprintf(stderr, "coupe_b200: level %d %s sweep %zu: %.1f us\n", c->event_level[e / 2],
              kind == 0 ? "dense" : kind == 1 ? "refine" : "(void optimistic)", e / 2, ms * 1e3);
  }
  CU(cudaGetLastError());
  return COUPE_ERR_OK;
}

int guarded_locked(coupe_b200_ctx *c, bool rib, void *stream, uint64_t *part_dev, uintptr_t dim, uintptr_t n,
                   const double *pts, int wtype, const void *w_dev, const void *wconst_host,
                   uintptr_t iter_count, double tolerance, const cb_engine::Prefilled *pre, int *compact_id_bytes) {
  try {
    return run_impl(c, rib, static_cast<cudaStream_t>(stream), part_dev, dim, n, pts, wtype, w_dev,
                    wconst_host, iter_count, tolerance, pre, compact_id_bytes);
  } catch (const CudaFail &f) {
    fprintf(stderr, "coupe_b200: CUDA error %s at %s\n", cudaGetErrorString(f.err), f.what);
    cudaGetLastError();
    return f.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  } catch (const NcclFail &f) {
    fprintf(stderr, "coupe_b200: NCCL error %d\n", (int)f.err);
    return COUPE_ERR_CRASH;
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

int guarded(coupe_b200_ctx *c, bool rib, void *stream, uint64_t *part_dev, uintptr_t dim, uintptr_t n,
            const double *pts, int wtype, const void *w_dev, const void *wconst_host,
            uintptr_t iter_count, double tolerance) {
  if (!c) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(c->mu);
  return guarded_locked(c, rib, stream, part_dev, dim, n, pts, wtype, w_dev, wconst_host, iter_count, tolerance,
                        nullptr, nullptr);
}

// Peer-memory exchange: allocate this rank's buffer, share its IPC handle with the other ranks
// of the box (all-gather through the NCCL communicator) and map theirs.  Any failure (ranks on
// different nodes, IPC disabled in a container, no peer access) leaves xchg_ok false.
void setup_xchg(coupe_b200_ctx *c) {
  c->xchg_ok = false;
  if (c->world < 2 || c->world > XCHG_MAX_WORLD || !g_nccl.AllGather) return;
  if (const char *e = getenv("COUPE_B200_NO_PEER_EXCHANGE"))
    if (*e && *e != '0') return;
  void *mine = nullptr, *hbuf = nullptr;
  unsigned int *aux = nullptr;
  int ok = 1;
  const size_t bytes = xchg_bytes(c->world);
  if (cudaMalloc(&mine, bytes) != cudaSuccess || cudaMemset(mine, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && (cudaMalloc(reinterpret_cast<void **>(&aux), 64) != cudaSuccess || cudaMemset(aux, 0, 64) != cudaSuccess)) ok = 0;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (ok && cudaIpcGetMemHandle(&h, mine) != cudaSuccess) ok = 0;
  // every rank takes part in the two collectives below whatever happened above
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  const size_t rec = 64 + 8;  // handle + "ok" word
  std::vector<unsigned char> all((size_t)c->world * rec);
  bool coll_ok = cudaMalloc(&hbuf, (size_t)(c->world + 1) * rec) == cudaSuccess;
  if (coll_ok) {
    unsigned char send[72];
    memcpy(send, &h, 64);
    const unsigned long long okw = (unsigned long long)ok;
    memcpy(send + 64, &okw, 8);
    unsigned char *d = static_cast<unsigned char *>(hbuf);
    coll_ok = cudaMemcpy(d, send, rec, cudaMemcpyHostToDevice) == cudaSuccess &&
              g_nccl.AllGather(d, d + rec, rec, ncclUint8, c->comm, nullptr) == ncclSuccess &&
              cudaStreamSynchronize(nullptr) == cudaSuccess &&
              cudaMemcpy(all.data(), d + rec, (size_t)c->world * rec, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  if (hbuf) cudaFree(hbuf);
  bool every = coll_ok;
  for (int r = 0; every && r < c->world; ++r) {
    unsigned long long okw;
    memcpy(&okw, all.data() + (size_t)r * rec + 64, 8);
    every = okw == 1;
  }
  if (every) {
    for (int r = 0; r < c->world && every; ++r) {
      if (r == c->rank) {
        c->xchg_peer[r] = static_cast<unsigned char *>(mine);
        continue;
      }
      cudaIpcMemHandle_t ph;
      memcpy(&ph, all.data() + (size_t)r * rec, 64);
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) every = false;
      c->xchg_peer[r] = static_cast<unsigned char *>(p);
    }
  }
  cudaGetLastError();
  // all ranks must agree: one more tiny all-reduce (min of "mapped everything")
  {
    unsigned int *d = nullptr;
    unsigned int v = every ? 1u : 0u;
    if (cudaMalloc(reinterpret_cast<void **>(&d), 4) == cudaSuccess &&
        cudaMemcpy(d, &v, 4, cudaMemcpyHostToDevice) == cudaSuccess &&
        g_nccl.AllReduce(d, d, 1, ncclUint32, ncclMin, c->comm, nullptr) == ncclSuccess &&
        cudaStreamSynchronize(nullptr) == cudaSuccess && cudaMemcpy(&v, d, 4, cudaMemcpyDeviceToHost) == cudaSuccess)
      every = every && v == 1;
    else
      every = false;
    if (d) cudaFree(d);
  }
  if (!every) {
    for (int r = 0; r < c->world; ++r) {
      if (r != c->rank && c->xchg_peer[r]) cudaIpcCloseMemHandle(c->xchg_peer[r]);
      c->xchg_peer[r] = nullptr;
    }
    if (mine) cudaFree(mine);
    if (aux) cudaFree(aux);
    cudaGetLastError();
    if (c->rank == 0) fprintf(stderr, "coupe_b200: peer-memory exchange unavailable, using NCCL all-reduces\n");
    return;
  }
  c->xchg_aux = aux;
  c->xchg_ok = true;
}

}  // namespace

namespace cb_engine {

void (*on_destroy)(coupe_b200_ctx *c) = nullptr;
void release_host_buffers(coupe_b200_ctx *c) {
  for (Buf *b : {&c->host_w, &c->host_ids, &c->host_pts}) b->release();
}
void lock(coupe_b200_ctx *c) { c->mu.lock(); }
void unlock(coupe_b200_ctx *c) { c->mu.unlock(); }
int device_of(const coupe_b200_ctx *c) { return c->device; }
int rank_of(const coupe_b200_ctx *c) { return c->rank; }
int world_of(const coupe_b200_ctx *c) { return c->world; }

int host_columns(coupe_b200_ctx *c, size_t n, size_t dim, size_t wbytes, bool raw_points, size_t iter_count,
                 HostColumns *out) {
  try {
    CU(cudaSetDevice(c->device));
    const size_t npad = ((n + 3) / 4) * 4 + 4;
    c->xcols.ensure(npad * sizeof(float) * 3);
    if (wbytes) c->host_w.ensure(npad * wbytes);
    c->host_ids.ensure(npad * (iter_count <= 16 ? 2 : 4));
    if (raw_points) c->host_pts.ensure(npad * dim * sizeof(double));
    out->npad = npad;
    for (int d = 0; d < 3; ++d) out->x[d] = c->xcols.as<float>() + (size_t)d * npad;
    out->w = wbytes ? c->host_w.p : nullptr;
    out->ids_compact = c->host_ids.p;
    out->pts_raw = raw_points ? c->host_pts.as<double>() : nullptr;
    return COUPE_ERR_OK;
  } catch (const CudaFail &f) {
    cudaGetLastError();
    return f.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  }
}

int run_locked(coupe_b200_ctx *c, bool rib, cudaStream_t st, const Prefilled *pre, int *id_bytes, uintptr_t dim,
               uintptr_t n, const double *points_dev, int wtype, const void *weights_dev, const void *wconst_host,
               uintptr_t iter_count, double tolerance) {
  return guarded_locked(c, rib, st, nullptr, dim, n, points_dev, wtype, weights_dev, wconst_host, iter_count,
                        tolerance, pre, id_bytes);
}

}  // namespace cb_engine

extern "C" {

int coupe_b200_ctx_create(coupe_b200_ctx **out, int device) {
  if (!out) return COUPE_ERR_CRASH;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return COUPE_ERR_CRASH;  // no CPU fallback
  }
  coupe_b200_ctx *c = new (std::nothrow) coupe_b200_ctx();
  if (!c) return COUPE_ERR_ALLOC;
  try {
    c->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->max_smem = prop.sharedMemPerBlockOptin;
    CU(cudaDeviceGetAttribute(&c->sm_khz, cudaDevAttrClockRate, device));
    CU(cudaHostAlloc(reinterpret_cast<void **>(&c->h_pinned), 64 * sizeof(uint32_t) * 4,
                     cudaHostAllocDefault));
    void *hf = nullptr;
    CU(cudaHostAlloc(&hf, FLAG_SLOTS * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(hf, 0, FLAG_SLOTS * sizeof(unsigned long long));
    c->h_flags = static_cast<volatile unsigned long long *>(hf);
    void *df = nullptr;
    CU(cudaHostGetDevicePointer(&df, hf, 0));
    c->d_flags = static_cast<unsigned long long *>(df);
  } catch (const CudaFail &f) {
    fprintf(stderr, "coupe_b200: CUDA error %s at %s\n", cudaGetErrorString(f.err), f.what);
    delete c;
    return f.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  }
  *out = c;
  return COUPE_ERR_OK;
}

void coupe_b200_ctx_destroy(coupe_b200_ctx *c) {
  if (!c) return;
  if (cb_engine::on_destroy) cb_engine::on_destroy(c);  // the host path's pinned buffers and streams (ffi.cu)
  cudaSetDevice(c->device);
  if (c->xchg_ok) {
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) {
      if (!c->xchg_peer[r]) continue;
      if (r == c->rank) cudaFree(c->xchg_peer[r]);
      else if (!c->xchg_local) cudaIpcCloseMemHandle(c->xchg_peer[r]);
    }
    if (c->xchg_aux) cudaFree(c->xchg_aux);
  }
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (Buf *b : {&c->host_w, &c->host_ids, &c->host_pts, &c->xcols, &c->ids, &c->w32, &c->node_rt, &c->rfast, &c->tsp_a, &c->tsp_b, &c->nsh_a, &c->nsh_b, &c->target, &c->part_w, &c->part_min, &c->hist_w, &c->hist_min, &c->nodes_a,
                 &c->nodes_b, &c->table_a, &c->table_b, &c->thi_a, &c->thi_b, &c->rtable, &c->gp, &c->tr_visited,
                 &c->tr_split, &c->tr_wl, &c->tr_sum, &c->tr_iters, &c->mom_partial})
    b->release();
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_flags) cudaFreeHost(const_cast<unsigned long long *>(c->h_flags));
  for (cudaEvent_t e : c->events) cudaEventDestroy(e);
  delete c;
}

int coupe_b200_nccl_unique_id(void *out128) {
  if (!out128 || !g_nccl.load()) return COUPE_ERR_CRASH;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return COUPE_ERR_CRASH;
  memcpy(out128, &id, 128);
  return COUPE_ERR_OK;
}

int coupe_b200_ctx_init_comm(coupe_b200_ctx *c, const void *unique_id128, int rank, int world) {
  if (!c || !unique_id128 || world < 1 || rank < 0 || rank >= world) return COUPE_ERR_CRASH;
  if (world == 1) {
    c->rank = 0;
    c->world = 1;
    return COUPE_ERR_OK;
  }
  if (!g_nccl.load()) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(c->mu);
  if (cudaSetDevice(c->device) != cudaSuccess) return COUPE_ERR_CRASH;
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  if (g_nccl.CommInitRank(&c->comm, world, id, rank) != ncclSuccess) return COUPE_ERR_CRASH;
  c->rank = rank;
  c->world = world;
  setup_xchg(c);  // optional: without it the histograms go through NCCL all-reduces
  return COUPE_ERR_OK;
}

// One process, several GPUs: one context per device, ranks of a communicator made with
// ncclCommInitAll; the exchange buffers are mapped with plain peer access.
int coupe_b200_group_create(coupe_b200_group **out, const int *devices, int ndev) {
  if (!out) return COUPE_ERR_CRASH;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    cudaGetLastError();
    return COUPE_ERR_CRASH;  // no CPU fallback
  }
  std::vector<int> devs;
  if (ndev <= 0 || !devices) {
    for (int d = 0; d < count; ++d) devs.push_back(d);
  } else {
    for (int i = 0; i < ndev; ++i) {
      if (devices[i] < 0 || devices[i] >= count) return COUPE_ERR_CRASH;
      for (int j = 0; j < i; ++j)
        if (devices[j] == devices[i]) return COUPE_ERR_CRASH;
      devs.push_back(devices[i]);
    }
  }
  const int world = (int)devs.size();
  if (world > XCHG_MAX_WORLD) return COUPE_ERR_CRASH;
  coupe_b200_group *g = new (std::nothrow) coupe_b200_group();
  if (!g) return COUPE_ERR_ALLOC;
  auto fail = [&](int err) {
    for (coupe_b200_ctx *c : g->ctx) coupe_b200_ctx_destroy(c);
    delete g;
    return err;
  };
  for (int d : devs) {
    coupe_b200_ctx *c = nullptr;
    const int err = coupe_b200_ctx_create(&c, d);
    if (err != COUPE_ERR_OK) return fail(err);
    g->ctx.push_back(c);
  }
  if (world > 1) {
    if (!g_nccl.load() || !g_nccl.CommInitAll) return fail(COUPE_ERR_CRASH);
    std::vector<ncclComm_t> comms(world);
    if (g_nccl.CommInitAll(comms.data(), world, devs.data()) != ncclSuccess) return fail(COUPE_ERR_CRASH);
    for (int r = 0; r < world; ++r) {
      g->ctx[r]->comm = comms[r];
      g->ctx[r]->rank = r;
      g->ctx[r]->world = world;
    }
    // peer-memory exchange: every device maps every other one's buffer directly
    bool ok = true;
    if (const char *e = getenv("COUPE_B200_NO_PEER_EXCHANGE"))
      if (*e && *e != '0') ok = false;
    for (int a = 0; a < world && ok; ++a)
      for (int b = 0; b < world && ok; ++b) {
        if (a == b) continue;
        int can = 0;
        ok = cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) == cudaSuccess && can;
        if (!ok) break;
        ok = cudaSetDevice(devs[a]) == cudaSuccess;
        const cudaError_t e = ok ? cudaDeviceEnablePeerAccess(devs[b], 0) : cudaErrorUnknown;
        ok = ok && (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled);
        cudaGetLastError();
      }
    std::vector<void *> bufs(world, nullptr), aux(world, nullptr);
    const size_t bytes = xchg_bytes(world);
    for (int r = 0; r < world && ok; ++r)
      ok = cudaSetDevice(devs[r]) == cudaSuccess && cudaMalloc(&bufs[r], bytes) == cudaSuccess &&
           cudaMemset(bufs[r], 0, bytes) == cudaSuccess && cudaMalloc(&aux[r], 64) == cudaSuccess &&
           cudaMemset(aux[r], 0, 64) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    if (ok) {
      for (int r = 0; r < world; ++r) {
        coupe_b200_ctx *c = g->ctx[r];
        for (int p = 0; p < world; ++p) c->xchg_peer[p] = static_cast<unsigned char *>(bufs[p]);
        c->xchg_aux = static_cast<unsigned int *>(aux[r]);
        c->xchg_ok = true;
        c->xchg_local = true;
      }
    } else {
      cudaGetLastError();
      for (int r = 0; r < world; ++r) {
        cudaSetDevice(devs[r]);
        if (bufs[r]) cudaFree(bufs[r]);
        if (aux[r]) cudaFree(aux[r]);
      }
      fprintf(stderr, "coupe_b200: peer-memory exchange unavailable in this process, using NCCL all-reduces\n");
    }
  }
  *out = g;
  return COUPE_ERR_OK;
}

void coupe_b200_group_destroy(coupe_b200_group *g) {
  if (!g) return;
  // every context frees its own exchange buffer; the peers' pointers are plain aliases
  for (coupe_b200_ctx *c : g->ctx) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
  }
  for (coupe_b200_ctx *c : g->ctx) coupe_b200_ctx_destroy(c);
  delete g;
}

int coupe_b200_group_size(const coupe_b200_group *g) { return g ? (int)g->ctx.size() : 0; }

coupe_b200_ctx *coupe_b200_group_ctx(coupe_b200_group *g, int i) {
  return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr;
}

int coupe_b200_rcb_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
                          uintptr_t n, const double *points_dev, int wtype, const void *weights_dev,
                          const void *wconst_host, uintptr_t iter_count, double tolerance) {
  return guarded(ctx, false, stream, part_dev, dim, n, points_dev, wtype, weights_dev, wconst_host,
                 iter_count, tolerance);
}

int coupe_b200_rib_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
                          uintptr_t n, const double *points_dev, int wtype, const void *weights_dev,
                          const void *wconst_host, uintptr_t iter_count, double tolerance) {
  return guarded(ctx, true, stream, part_dev, dim, n, points_dev, wtype, weights_dev, wconst_host,
                 iter_count, tolerance);
}

int coupe_b200_last_stats(const coupe_b200_ctx *ctx, coupe_b200_stats *out) {
  if (!ctx || !out) return COUPE_ERR_CRASH;
  *out = ctx->stats;
  return COUPE_ERR_OK;
}

uint32_t coupe_b200_last_sweep_times(const coupe_b200_ctx *c, double *ms, int32_t *level, int32_t *kind, uint32_t cap) {
  if (!c) return 0;
  const uint32_t m = (uint32_t)c->sweep_ms.size();
  for (uint32_t i = 0; i < m && i < cap; ++i) {
    if (ms) ms[i] = c->sweep_ms[i];
    if (level) level[i] = c->event_level[i];
    if (kind) kind[i] = c->event_kind[i];
  }
  return m;
}

int coupe_b200_last_trace(coupe_b200_ctx *c, uint8_t *visited, float *split_pos, double *weight_left,
                          double *sum, uint32_t *iters) {
  if (!c || !c->trace_on) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(c->mu);
  const size_t m = c->trace_levels > 0 ? ((size_t)1 << c->trace_levels) - 1 : 0;
  if (m == 0) return COUPE_ERR_OK;
  if (cudaSetDevice(c->device) != cudaSuccess) return COUPE_ERR_CRASH;
  bool ok = true;
  if (visited) ok = ok && cudaMemcpy(visited, c->tr_visited.p, m, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (split_pos) ok = ok && cudaMemcpy(split_pos, c->tr_split.p, m * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (weight_left) ok = ok && cudaMemcpy(weight_left, c->tr_wl.p, m * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (sum) ok = ok && cudaMemcpy(sum, c->tr_sum.p, m * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (iters) ok = ok && cudaMemcpy(iters, c->tr_iters.p, m * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
  return ok ? COUPE_ERR_OK : COUPE_ERR_CRASH;
}

int coupe_b200_reserve(coupe_b200_ctx *c, uintptr_t n, uintptr_t dim, uintptr_t iter_count) {
  if (!c || (dim != 2 && dim != 3)) return COUPE_ERR_CRASH;
  if (iter_count > (uintptr_t)MAX_LEVELS) return COUPE_ERR_ALLOC;
  std::lock_guard<std::mutex> lock(c->mu);
  try {
    CU(cudaSetDevice(c->device));
    const size_t npad = ((n + 3) / 4) * 4 + 4;
    c->xcols.ensure(npad * sizeof(float) * 3);
    c->ids.ensure(npad * sizeof(uint32_t));
  } catch (const CudaFail &f) {
    cudaGetLastError();
    return f.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  }
  return COUPE_ERR_OK;
}

int coupe_b200_set_option(coupe_b200_ctx *c, const char *name, int64_t value) {
  if (!c || !name) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(c->mu);
  const std::string s(name);
  if (s == "kmax_a") c->kmax_a = (int)std::max<int64_t>(1, std::min<int64_t>(10, value));
  else if (s == "nb_smem_log2") c->nb_smem_log2 = (int)std::max<int64_t>(6, std::min<int64_t>(14, value));
  else if (s == "kmax_refine") c->kmax_refine = (int)std::max<int64_t>(1, std::min<int64_t>(10, value));
  else if (s == "force_global") c->force_global = value != 0;
  else if (s == "trace") c->trace_on = value != 0;
  else if (s == "time_sweeps") c->time_sweeps = (int)value;
  else if (s == "peer_exchange") c->use_xchg_opt = value != 0;
  else if (s == "sample_weights") c->sample_w_opt = value != 0;
  else if (s == "carve_fit") c->carve_fit = value != 0;
  else if (s == "defer") c->defer_opt = (int)std::max<int64_t>(0, std::min<int64_t>(2, value));
  else if (s == "smem_pad") c->smem_pad = (int)std::max<int64_t>(0, std::min<int64_t>(32768, value));
  else if (s == "table_rep_max") c->table_rep_max = (int)std::max<int64_t>(0, std::min<int64_t>(3, value));
  else return COUPE_ERR_NOT_FOUND;
  return COUPE_ERR_OK;
}

int coupe_b200_ctx_device(const coupe_b200_ctx *c) { return c ? c->device : -1; }

const char *coupe_b200_version(void) { return "coupe_b200 0.1 (sm_100a, CUDA " __DATE__ ")"; }

}  // extern "C"
