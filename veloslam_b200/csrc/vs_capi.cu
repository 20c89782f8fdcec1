// vs_capi.cu -- C-ABI implementation (include/veloslam_b200.h) over the sm_100a kernels.
//
// One context = one parser + one pose snapshot on one GPU.  Each result slot owns a CUDA
// stream, its device buffers and pinned host mirrors of the small tables, so that with two
// slots the H2D copy of batch k+1 overlaps the kernels and D2H of batch k.
// No CPU fallback: without a device vs_create fails.
#include "../../include/veloslam_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "vs_kernels.cuh"
#include "vs_layout.cuh"

using namespace vsd;

namespace {

constexpr int kEagerFrames = kEagerRows;  // minimum frame-table rows copied back with the header

struct HostFrameRow {  // pinned mirror of the per-frame device tables
  std::vector<long long> first;
  std::vector<int> start;
};

// How the stream operations of a batch are issued.  A rotation-sized batch (online use: ~350
// packets) is 17 small copies, memsets and kernels whose launch latency, not their work, sets the
// submit -> index latency; such batches are recorded ONCE per result slot as a CUDA graph (stream
// capture of the very same code) and afterwards only the node parameters that differ between
// batches (sizes, grids, kernel arguments, host pointers) are refreshed before one
// cudaGraphLaunch.  UPDATE walks the captured nodes in issue order, so one code path serves the
// direct launch, the capture and the refresh.
struct GraphOps {
  enum Mode { DIRECT, CAPTURE, UPDATE };
  Mode mode = DIRECT;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<cudaGraphNode_t>* nodes = nullptr;
  // what each node was last set to (empty: unknown): a refresh skips the nodes whose operation
  // did not change, every cudaGraphExec*SetParams call costs about a microsecond
  std::vector<std::string>* sigs = nullptr;
  size_t cursor = 0;

  static std::string sig_of(const void* a, size_t na, const void* b = nullptr, size_t nb = 0) {
    std::string s(static_cast<const char*>(a), na);
    if (nb) s.append(static_cast<const char*>(b), nb);
    return s;
  }
  // UPDATE: true when node `cursor` already holds this operation (then it is skipped)
  bool unchanged(const std::string& sig) {
    if (cursor >= sigs->size()) return false;
    std::string& have = (*sigs)[cursor];
    if (!sig.empty() && have == sig) {
      ++cursor;
      return true;
    }
    have = sig;
    return false;
  }

  cudaError_t captured() {
    cudaStreamCaptureStatus st;
    unsigned long long id = 0;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t nd = 0;
    cudaError_t e = cudaStreamGetCaptureInfo(stream, &st, &id, &g, &deps, &nd);
    if (e != cudaSuccess) return e;
    if (st != cudaStreamCaptureStatusActive || nd != 1) return cudaErrorStreamCaptureInvalidated;
    nodes->push_back(deps[0]);
    sigs->push_back(pending_sig);
    return cudaSuccess;
  }
  std::string pending_sig;
  cudaGraphNode_t next() { return cursor < nodes->size() ? (*nodes)[cursor++] : nullptr; }

  cudaError_t copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    if (mode != DIRECT) {
      const void* key[3] = {dst, src, reinterpret_cast<const void*>(bytes)};
      pending_sig = sig_of(key, sizeof(key));
    }
    if (mode == UPDATE) {
      if (unchanged(pending_sig)) return cudaSuccess;
      cudaGraphNode_t nd = next();
      if (!nd) return cudaErrorInvalidValue;
      return cudaGraphExecMemcpyNodeSetParams1D(exec, nd, dst, src, bytes, kind);
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, stream);
    if (e == cudaSuccess && mode == CAPTURE) e = captured();
    return e;
  }
  // arg0_bytes: size of the kernel's one parameter struct (0: not known, the node is always refreshed)
  cudaError_t launch(const void* func, unsigned grid, unsigned block, size_t smem, void** args,
                     size_t arg0_bytes = 0) {
    if (mode != DIRECT) {
      pending_sig.clear();
      if (arg0_bytes) {
        const size_t key[4] = {reinterpret_cast<size_t>(func), grid, block, smem};
        pending_sig = sig_of(key, sizeof(key), args[0], arg0_bytes);
      }
    }
    if (mode == UPDATE) {
      if (unchanged(pending_sig)) return cudaSuccess;
      cudaGraphNode_t nd = next();
      if (!nd) return cudaErrorInvalidValue;
      cudaKernelNodeParams kp;
      std::memset(&kp, 0, sizeof(kp));
      kp.func = const_cast<void*>(func);
      kp.gridDim = dim3(grid, 1, 1);
      kp.blockDim = dim3(block, 1, 1);
      kp.sharedMemBytes = (unsigned)smem;
      kp.kernelParams = args;
      return cudaGraphExecKernelNodeSetParams(exec, nd, &kp);
    }
    cudaError_t e = cudaLaunchKernel(func, dim3(grid, 1, 1), dim3(block, 1, 1), args, smem, stream);
    if (e == cudaSuccess && mode == CAPTURE) e = captured();
    return e;
  }
  // timing events exist only on the direct path: an event recorded by a graph node cannot be
  // read back with cudaEventElapsedTime (run_batch brackets the whole graph instead)
  cudaError_t record(cudaEvent_t ev) {
    if (mode != DIRECT) return cudaSuccess;
    return cudaEventRecord(ev, stream);
  }
};

// what decides the SEQUENCE of operations (and the kernel functions) of a batch: a captured
// graph is re-used only for batches with the same key
struct GraphKey {
  int dev_in, pcap_t, fused, index_only, adj, crop, deskew, pose_valid, has_decode, eager_packed;
  bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};

struct Slot {
  cudaStream_t stream = nullptr;
  GraphOps ops;
  cudaGraphExec_t gexec = nullptr;
  cudaGraph_t ggraph = nullptr;  // kept: the node handles used for the refresh belong to it
  std::vector<cudaGraphNode_t> gnodes;
  std::vector<std::string> gsigs;
  GraphKey gkey;
  int graph_launches = 0;
  bool graph_issued = false;  // the batch in flight went out as a graph (no per-kernel events)
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr, ev_done = nullptr;
  // device
  uint8_t* d_in = nullptr;  // staged packets (host input path only, allocated lazily)
  size_t d_in_bytes = 0;
  long long* d_time = nullptr;
  const long long* d_time_used = nullptr;  // times of the batch in flight (may be caller-owned)
  PktSeg* d_seg = nullptr;
  BlkRec* d_recs = nullptr;               // 12 block records per packet (k_scan -> k_decode)
  unsigned long long* d_pkt_off = nullptr;  // first emitted point of each packet (k_pose)
  double* d_pose_mat = nullptr;
  float *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
  uint8_t *d_inten = nullptr, *d_laser = nullptr;
  uint16_t *d_az = nullptr, *d_dist = nullptr;
  uint32_t* d_t = nullptr;
  uint8_t* d_zero = nullptr;  // one allocation zeroed per batch: look-back words, counters,
  size_t zero_bytes = 0;      //   per-frame laser counts
  unsigned long long *d_st_map = nullptr, *d_st_wrap = nullptr, *d_st_cnt = nullptr;
  int* d_counters = nullptr;
  unsigned long long* d_grp_cnt = nullptr;  // per group of kGroupTiles scan tiles (atomics)
  unsigned *d_grp_wsum = nullptr, *d_grp_wmax = nullptr;
  unsigned* d_frame_counts = nullptr;
  uint8_t* d_ff = nullptr;  // one allocation set to 0xff per batch: first_point / start_block
  long long* d_frame_first = nullptr;
  int* d_frame_start = nullptr;
  int* d_frame_meta_pkt = nullptr;
  long long* d_frame_meta_time = nullptr;
  int* d_frame_skips = nullptr;
  BatchHeader* d_hdr = nullptr;
  EagerBlock* d_eager = nullptr;  // header + first rows of the frame tables, packed (k_frames)
  EagerBlock* h_eager = nullptr;
  BatchHeader* h_hdr_init = nullptr;  // initial header values of a batch issued with memsets  // pinned
  bool eager_packed = false;      // the batch in flight brings its index back through h_eager
  // HDLFrame layout of the batch (vs_layout_frames), allocated on first use
  uint8_t* d_lay_xyzi = nullptr;
  uint8_t* d_lay_meta = nullptr;
  size_t lay_xyzi_bytes = 0, lay_meta_bytes = 0;
  unsigned long long* d_lay_rows = nullptr;  // [frame_cap][64]
  uint8_t* d_lay_st = nullptr;               // [chunk counter | look-back words]
  size_t lay_st_bytes = 0;
  cudaEvent_t ev_l0 = nullptr, ev_l1 = nullptr;
  bool lay_valid = false;     // layout kernels of the current ticket were launched
  bool lay_inflight = false;  // ... and nobody has synchronised with them yet
  std::vector<vs_frame_rows> lay_rows;
  vs_layout layout;
  // pinned host
  BatchHeader* h_hdr = nullptr;
  long long* h_frame_first = nullptr;
  int* h_frame_start = nullptr;
  int* h_frame_meta_pkt = nullptr;
  long long* h_frame_meta_time = nullptr;
  int* h_frame_skips = nullptr;
  unsigned* h_frame_counts = nullptr;
  size_t h_frames_cap = 0;
  // bookkeeping of the batch in flight
  bool busy = false;
  bool done = false;
  uint64_t ticket = 0;
  int64_t n = 0, halo = 0;
  int mode = 0;
  uint32_t flags = 0;
  int64_t t_base = 0;
  vs_carry carry_in;
  int n_launches = 0;
  int64_t eager_rows = 0;
  bool index_only = false;
  std::vector<vs_frame> frames;
  bool d1_recorded = false;  // ev_d1 has been recorded on this slot's stream at least once
  int n_frames_total = 0;   // frames of the batch; frames.size() is 2 at most with VS_FLAG_NO_FRAME_LIST
  bool sparse_frames = false;
  vs_result result;
};

// Opt-in and occupancy of one kernel variant, cached per context (see launch_decode).
struct KernelCache {
  bool attr_set = false;
  size_t smem = 0;
  int per_sm = 0;
};

}  // namespace

struct vs_ctx {
  int device = 0;
  int sm_count = 0;
  int64_t max_packets = 0;
  int64_t max_poses = 0;
  int64_t frame_cap = 0;
  int n_slots = 1;
  Slot slots[2];
  uint64_t next_ticket = 1;
  bool calibrated = false;
  DevConfig h_cfg;
  DevConfig* d_cfg = nullptr;
  double *d_lut_sin = nullptr, *d_lut_cos = nullptr;
  long long* d_pose_t = nullptr;
  double* d_pose_trv = nullptr;
  std::vector<int64_t> pose_t;
  std::vector<double> pose_trv;
  cudaStream_t cfg_stream = nullptr;  // small table uploads (calibration, filters, poses)
  bool layout_attr_set = false;
  bool use_graph = false;  // rotation-sized contexts: batches are issued as a CUDA graph
  KernelCache dec_cache[3][2][2];  // [ADJ][DSK][FUSED]
  KernelCache scan_cache[3][2];    // [ADJ][CROP]
  bool chain_decode = true;  // VELOSLAM_DECODE_CHAIN=0: do not order the two slots' k_decode launches
  int reset_kernel = -1;  // VELOSLAM_RESET_KERNEL=0: memsets + a header copy for batches issued directly
  bool two_pass = true;  // false (VELOSLAM_SINGLE_PASS=1): k_pose_pre + k_decode<.., FUSED> where it applies
  std::string err;
};

namespace {

thread_local std::string g_create_err;

int fail(vs_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define VS_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(ctx, VS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
  } while (0)

inline double to_radians_h(double x) { return (x * M_PI) / 180.0; }  // HDLParser.cxx:59

// laser selection per return slot of a lower / upper block (see DevConfig::sel_lo)
void update_selection(DevConfig& c) {
  c.sel_lo = c.sel_hi = 0;
  for (int bank = 0; bank < 2; ++bank) {
    for (int lane = 0; lane < 32; ++lane) {
      int id = lane + 32 * bank;
      if (c.adj_mode == 2 && id >= 16) id -= 16;  // VLP-16 remap, HDLParser.cxx:935-943
      const bool ok = ((c.laser_mask >> id) & 1ull) && id < c.n_enabled;
      if (ok) (bank ? c.sel_hi : c.sel_lo) |= 1u << lane;
    }
  }
}

// TransformManager::interpolateTransform (TransformManager.cxx:149-177) over the sorted host
// snapshot; bracket = clamp(lower_bound, 1, N-1) (TimeLine.h:384-468 net semantics).
// *hint (optional): where the previous lookup landed.  Frames come in time order, so the next
// bracket is at or right behind the last one: a galloping search from there instead of a cold
// binary search over the whole timeline (36 000 frames x 360 000 poses for an hour of data).
void host_interpolate(const vs_ctx* c, int64_t t, double out[9], bool* found, bool* valid,
                      int64_t* hint = nullptr) {
  const int64_t n = (int64_t)c->pose_t.size();
  for (int i = 0; i < 9; ++i) out[i] = 0.0;
  *found = false;
  *valid = false;
  if (n == 0) return;
  *found = true;
  if (n == 1) {
    const double* f = c->pose_trv.data();
    const double sec = (float)(t - c->pose_t[0]) / 1e6f;  // long / float, TransformManager.cxx:161
    for (int i = 0; i < 3; ++i) {
      out[6 + i] = f[6 + i];
      out[3 + i] = f[3 + i];
      out[i] = f[i] + f[6 + i] * sec;
    }
    return;
  }
  const int64_t* pt = c->pose_t.data();
  int64_t lo = 0, hi = n;  // lower_bound lies in [lo, hi]
  if (hint && *hint >= 0 && *hint < n) {
    const int64_t h = *hint;
    if (pt[h] < t) {
      int64_t step = 1;
      lo = h + 1;
      while (lo + step < n && pt[lo + step] < t) {
        lo += step;
        step *= 2;
      }
      hi = std::min(n, lo + step + 1);
    } else {
      int64_t step = 1;
      hi = h;
      while (hi - step > 0 && pt[hi - step] >= t) {
        hi -= step;
        step *= 2;
      }
      lo = std::max<int64_t>(0, hi - step);
    }
  }
  int64_t i = std::lower_bound(pt + lo, pt + hi, t) - pt;
  if (hint) *hint = std::min(i, n - 1);
  i = std::min(std::max<int64_t>(i, 1), n - 1);
  const double* f = c->pose_trv.data() + (i - 1) * 9;
  const double* b = c->pose_trv.data() + i * 9;
  const double ratio = double(t - c->pose_t[i - 1]) / (c->pose_t[i] - c->pose_t[i - 1]);
  for (int k = 0; k < 9; ++k) out[k] = f[k] + ((b[k] - f[k]) * ratio);
  *valid = true;
}

void drop_graph(Slot& s) {
  if (s.gexec) cudaGraphExecDestroy(s.gexec);
  if (s.ggraph) cudaGraphDestroy(s.ggraph);
  s.gexec = nullptr;
  s.ggraph = nullptr;
  s.gnodes.clear();
  s.gsigs.clear();
}

void free_slot(Slot& s) {
  cudaFree(s.d_in);
  cudaFree(s.d_time);
  cudaFree(s.d_seg);
  cudaFree(s.d_recs);
  cudaFree(s.d_pkt_off);
  cudaFree(s.d_pose_mat);
  cudaFree(s.d_x);
  cudaFree(s.d_y);
  cudaFree(s.d_z);
  cudaFree(s.d_inten);
  cudaFree(s.d_laser);
  cudaFree(s.d_az);
  cudaFree(s.d_dist);
  cudaFree(s.d_t);
  cudaFree(s.d_zero);
  cudaFree(s.d_ff);
  cudaFree(s.d_frame_meta_pkt);
  cudaFree(s.d_frame_meta_time);
  cudaFree(s.d_frame_skips);
  cudaFree(s.d_hdr);
  cudaFree(s.d_eager);
  cudaFreeHost(s.h_eager);
  cudaFreeHost(s.h_hdr_init);
  drop_graph(s);
  cudaFree(s.d_lay_xyzi);
  cudaFree(s.d_lay_meta);
  cudaFree(s.d_lay_rows);
  cudaFree(s.d_lay_st);
  if (s.ev_l0) cudaEventDestroy(s.ev_l0);
  if (s.ev_l1) cudaEventDestroy(s.ev_l1);
  cudaFreeHost(s.h_hdr);
  cudaFreeHost(s.h_frame_first);
  cudaFreeHost(s.h_frame_start);
  cudaFreeHost(s.h_frame_meta_pkt);
  cudaFreeHost(s.h_frame_meta_time);
  cudaFreeHost(s.h_frame_skips);
  cudaFreeHost(s.h_frame_counts);
  if (s.ev_k0) cudaEventDestroy(s.ev_k0);
  if (s.ev_k1) cudaEventDestroy(s.ev_k1);
  if (s.ev_d0) cudaEventDestroy(s.ev_d0);
  if (s.ev_d1) cudaEventDestroy(s.ev_d1);
  if (s.ev_done) cudaEventDestroy(s.ev_done);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

// Copy a small host table (pageable memory) to the device so that it is complete and visible to
// every slot stream when the call returns.  cudaMemcpy from pageable memory may return once the
// data is staged, and the slot streams are non-blocking (not ordered against the legacy stream),
// so the copy goes through the context's configuration stream and is waited for.  Refused while a batch that may
// read the table is still running.
int upload_table(vs_ctx* ctx, void* dst, const void* src, size_t bytes, bool layout_reads_it,
                 const char* who) {
  for (int i = 0; i < ctx->n_slots; ++i) {
    const Slot& s = ctx->slots[i];
    if ((s.busy && !s.done) || (layout_reads_it && s.lay_inflight))
      return fail(ctx, VS_ERR_STATE, std::string(who) + ": a batch is still in flight (vs_wait / vs_sync it first)");
  }
  // a stream of its own: waiting on a slot's stream would also wait for that slot's
  // device -> host copies of an earlier batch and stall a pipelined caller
  cudaStream_t st = ctx->cfg_stream;
  VS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
  VS_CUDA(cudaStreamSynchronize(st));
  return VS_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int alloc_slot(vs_ctx* ctx, Slot& s) {
  const int64_t np = ctx->max_packets;
  const int64_t pts = np * 384;
  const int64_t fc = ctx->frame_cap;
  VS_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  VS_CUDA(cudaEventCreate(&s.ev_k0));
  VS_CUDA(cudaEventCreate(&s.ev_k1));
  VS_CUDA(cudaEventCreate(&s.ev_d0));
  VS_CUDA(cudaEventCreate(&s.ev_d1));
  VS_CUDA(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
  VS_CUDA(cudaMalloc(&s.d_time, np * sizeof(long long)));
  VS_CUDA(cudaMalloc(&s.d_seg, np * sizeof(PktSeg)));
  VS_CUDA(cudaMalloc(&s.d_recs, np * kBlocks * sizeof(BlkRec)));
  VS_CUDA(cudaMalloc(&s.d_pkt_off, (np + 2) * sizeof(unsigned long long)));  // k_decode copies even-aligned pairs
  VS_CUDA(cudaMalloc(&s.d_pose_mat, np * kDeskewRow * sizeof(double)));  // 12 used without deskew
  VS_CUDA(cudaMalloc(&s.d_x, pts * sizeof(float)));
  VS_CUDA(cudaMalloc(&s.d_y, pts * sizeof(float)));
  VS_CUDA(cudaMalloc(&s.d_z, pts * sizeof(float)));
  VS_CUDA(cudaMalloc(&s.d_inten, pts));
  VS_CUDA(cudaMalloc(&s.d_laser, pts));
  VS_CUDA(cudaMalloc(&s.d_az, pts * sizeof(uint16_t)));
  VS_CUDA(cudaMalloc(&s.d_dist, pts * sizeof(uint16_t)));
  VS_CUDA(cudaMalloc(&s.d_t, pts * sizeof(uint32_t)));
  // zeroed-per-batch block: [st_map | st_wrap | st_cnt | counters | group aggregates | frame_counts]
  const size_t scan_tiles = (size_t)((np + kTilePkts - 1) / kTilePkts);
  size_t off = 0;
  const size_t o_map = off;
  off += align_up(scan_tiles * 8, 256);
  const size_t o_wrap = off;
  off += align_up(scan_tiles * 8, 256);
  const size_t o_cnt = off;
  off += align_up(scan_tiles * 8, 256);
  const size_t o_ctr = off;
  off += 256;
  const size_t groups = scan_tiles / kGroupTiles + 1;
  const size_t o_gc = off;
  off += align_up(groups * 8, 256);
  const size_t o_gs = off;
  off += align_up(groups * 4, 256);
  const size_t o_gm = off;
  off += align_up(groups * 4, 256);
  const size_t o_fc = off;
  off += (size_t)fc * kMaxLasers * sizeof(unsigned);
  // single-pass pipeline: [tile counter | 32-byte look-back records per 8-packet tile], carved
  // per batch right behind the frame counts the batch can touch (one memset covers both)
  off += 256 + (size_t)((np + kDecTile - 1) / kDecTile) * 32;
  s.zero_bytes = off;
  VS_CUDA(cudaMalloc(&s.d_zero, off));
  s.d_st_map = reinterpret_cast<unsigned long long*>(s.d_zero + o_map);
  s.d_st_wrap = reinterpret_cast<unsigned long long*>(s.d_zero + o_wrap);
  s.d_st_cnt = reinterpret_cast<unsigned long long*>(s.d_zero + o_cnt);
  s.d_counters = reinterpret_cast<int*>(s.d_zero + o_ctr);
  s.d_grp_cnt = reinterpret_cast<unsigned long long*>(s.d_zero + o_gc);
  s.d_grp_wsum = reinterpret_cast<unsigned*>(s.d_zero + o_gs);
  s.d_grp_wmax = reinterpret_cast<unsigned*>(s.d_zero + o_gm);
  s.d_frame_counts = reinterpret_cast<unsigned*>(s.d_zero + o_fc);
  VS_CUDA(cudaMalloc(&s.d_ff, (size_t)fc * 12));
  s.d_frame_first = reinterpret_cast<long long*>(s.d_ff);
  s.d_frame_start = reinterpret_cast<int*>(s.d_ff + (size_t)fc * 8);
  VS_CUDA(cudaMalloc(&s.d_frame_meta_pkt, fc * sizeof(int)));
  VS_CUDA(cudaMalloc(&s.d_frame_meta_time, fc * sizeof(long long)));
  VS_CUDA(cudaMalloc(&s.d_frame_skips, fc * sizeof(int)));
  VS_CUDA(cudaMalloc(&s.d_hdr, sizeof(BatchHeader)));
  VS_CUDA(cudaMalloc(&s.d_eager, sizeof(EagerBlock)));
  VS_CUDA(cudaMallocHost(&s.h_eager, sizeof(EagerBlock)));
  VS_CUDA(cudaMallocHost(&s.h_hdr_init, sizeof(BatchHeader)));
  VS_CUDA(cudaMallocHost(&s.h_hdr, sizeof(BatchHeader)));
  return VS_OK;
}

int ensure_host_frames(vs_ctx* ctx, Slot& s, size_t need) {
  if (need <= s.h_frames_cap) return VS_OK;
  size_t cap = std::max<size_t>(need, std::max<size_t>(kEagerFrames, s.h_frames_cap * 2));
  cudaFreeHost(s.h_frame_first);
  cudaFreeHost(s.h_frame_start);
  cudaFreeHost(s.h_frame_meta_pkt);
  cudaFreeHost(s.h_frame_meta_time);
  cudaFreeHost(s.h_frame_skips);
  cudaFreeHost(s.h_frame_counts);
  s.h_frames_cap = 0;
  VS_CUDA(cudaMallocHost(&s.h_frame_first, cap * sizeof(long long)));
  VS_CUDA(cudaMallocHost(&s.h_frame_start, cap * sizeof(int)));
  VS_CUDA(cudaMallocHost(&s.h_frame_meta_pkt, cap * sizeof(int)));
  VS_CUDA(cudaMallocHost(&s.h_frame_meta_time, cap * sizeof(long long)));
  VS_CUDA(cudaMallocHost(&s.h_frame_skips, cap * sizeof(int)));
  VS_CUDA(cudaMallocHost(&s.h_frame_counts, cap * kMaxLasers * sizeof(unsigned)));
  s.h_frames_cap = cap;
  return VS_OK;
}

// rows [r0, r0 + n_rows) of the six frame tables into their pinned mirrors
int copy_frame_rows(vs_ctx* ctx, Slot& s, size_t n_rows, size_t r0 = 0) {
  GraphOps& g = s.ops;
  VS_CUDA(g.copy(s.h_frame_first + r0, s.d_frame_first + r0, n_rows * sizeof(long long), cudaMemcpyDeviceToHost));
  VS_CUDA(g.copy(s.h_frame_start + r0, s.d_frame_start + r0, n_rows * sizeof(int), cudaMemcpyDeviceToHost));
  VS_CUDA(g.copy(s.h_frame_meta_pkt + r0, s.d_frame_meta_pkt + r0, n_rows * sizeof(int), cudaMemcpyDeviceToHost));
  VS_CUDA(g.copy(s.h_frame_meta_time + r0, s.d_frame_meta_time + r0, n_rows * sizeof(long long),
                 cudaMemcpyDeviceToHost));
  VS_CUDA(g.copy(s.h_frame_skips + r0, s.d_frame_skips + r0, n_rows * sizeof(int), cudaMemcpyDeviceToHost));
  VS_CUDA(g.copy(s.h_frame_counts + r0 * kMaxLasers, s.d_frame_counts + r0 * kMaxLasers,
                 n_rows * kMaxLasers * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return VS_OK;
}

// Opt-in and occupancy of one kernel variant, cached per context.  The dynamic shared memory
// opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) is a per-device attribute of the function:
// every context raises it to the size the largest legal stride (1280) needs, never to its own
// batch's size, so contexts on other threads / other GPUs cannot lower it under a launch.
constexpr int64_t kMaxStride = 1280;

template <int ADJ, int DSK, int FUSED = 0>
int launch_decode(vs_ctx* ctx, Slot& s, DecParams dp, int64_t stride) {
  typedef DecLayout<ADJ, DSK, FUSED> L;
  constexpr int kThreads = FUSED ? kFusedThreads : kDecThreads;
  auto stage_bytes = [](int64_t st) {
    return align_up((size_t)L::kDPkts + (size_t)kDecTile * (size_t)st + 48, 128);
  };
  dp.stage_bytes = (int)stage_bytes(stride);
  const size_t smem = (size_t)L::kStages + L::kNumStages * (size_t)dp.stage_bytes;
  KernelCache& kc = ctx->dec_cache[ADJ][DSK][FUSED];
  if (!kc.attr_set) {
    const size_t smem_max = (size_t)L::kStages + L::kNumStages * stage_bytes(kMaxStride);
    VS_CUDA(cudaFuncSetAttribute(k_decode<ADJ, DSK, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem_max));
    kc.attr_set = true;
  }
  if (kc.smem != smem) {
    VS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kc.per_sm, k_decode<ADJ, DSK, FUSED>, kThreads, smem));
    kc.smem = smem;
  }
  if (kc.per_sm < 1) return fail(ctx, VS_ERR_CUDA, "k_decode does not fit on an SM");
  int grid = ctx->sm_count * kc.per_sm;
  if (grid > dp.n_tiles) grid = dp.n_tiles;
  void* args[] = {&dp};
  VS_CUDA(s.ops.launch(reinterpret_cast<const void*>(&k_decode<ADJ, DSK, FUSED>), (unsigned)grid, kThreads,
                       smem, args));
  return VS_OK;
}

inline size_t scan_stage_bytes(int64_t stride) {
  return align_up((size_t)(kTilePkts + 1) * (size_t)stride + kLead + 48, 128);
}

template <int ADJ, bool CROP>
int launch_scan(vs_ctx* ctx, Slot& s, const ScanParams& sp, size_t smem) {
  KernelCache& kc = ctx->scan_cache[ADJ][CROP ? 1 : 0];
  if (!kc.attr_set) {
    const size_t smem_max = align_up(sizeof(ScanShared), 128) + scan_stage_bytes(kMaxStride);
    VS_CUDA(cudaFuncSetAttribute(k_scan<ADJ, CROP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem_max));
    kc.attr_set = true;
  }
  if (kc.smem != smem) {
    VS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kc.per_sm, k_scan<ADJ, CROP>, kScanThreads,
                                                          smem));
    kc.smem = smem;
  }
  if (kc.per_sm < 1) return fail(ctx, VS_ERR_CUDA, "k_scan does not fit on an SM");
  int grid = ctx->sm_count * kc.per_sm;
  if (grid > sp.n_tiles) grid = sp.n_tiles;
  ScanParams spv = sp;
  void* args[] = {&spv};
  VS_CUDA(s.ops.launch(reinterpret_cast<const void*>(&k_scan<ADJ, CROP>), (unsigned)grid, kScanThreads, smem,
                       args));
  return VS_OK;
}

// rows of the frame tables copied back with the batch header, before the frame count is known
int64_t eager_frame_rows(const vs_ctx* ctx, int64_t n) {
  const int64_t frames_possible = std::min<int64_t>(ctx->frame_cap, 12 * n + 1);
  return std::min<int64_t>(frames_possible, std::max<int64_t>(kEagerFrames, n / 100 + 8));
}

// Everything a batch puts on its slot's stream, through s.ops (direct, captured or refreshed).
int enqueue_batch(vs_ctx* ctx, Slot& s, const uint8_t* pkts, int64_t stride, const int64_t* pkt_time,
                  int64_t n, int64_t halo, int mode, uint32_t flags, int64_t t_base,
                  const vs_carry* carry_in, bool index_only) {
  GraphOps& g = s.ops;
  const bool dev_in = (flags & VS_FLAG_DEVICE_INPUT) != 0;
  const bool pcap_t = (flags & VS_FLAG_PCAP_TIMES) != 0;
  const uint8_t* d_pkts = pkts;
  const long long* d_time = reinterpret_cast<const long long*>(pkt_time);
  const int64_t payload_bytes = (n - 1) * stride + kPacketBytes;
  s.n_launches = 0;

  // ---- per-batch resets ------------------------------------------------------------------
  const int64_t scan_tiles = (n + kTilePkts - 1) / kTilePkts;
  const int64_t pose_tiles = (n + kPoseThreads - 1) / kPoseThreads;
  const int64_t n_dec = n - halo;
  const int64_t dec_tiles = (n_dec + kDecTile - 1) / kDecTile;
  const int64_t frames_possible = std::min<int64_t>(ctx->frame_cap, 12 * n + 1);
  const int n_poses = (int)ctx->pose_t.size();
  const bool pose_valid = n_poses >= 2;
  const int adj = ctx->h_cfg.adj_mode;
  const bool crop = ctx->h_cfg.crop_returns != 0 && !index_only;
  const bool deskew = (flags & VS_FLAG_DESKEW_PER_POINT) != 0 && pose_valid && !index_only;
  // single-pass pipeline (k_decode<.., FUSED>): everything but the crop filter, the per-point
  // deskew extension and index-only passes
  const bool fused = !ctx->two_pass && !crop && !deskew && !index_only;
  const int64_t fused_tiles = (n + kDecTile - 1) / kDecTile;
  const size_t fc_bytes = (size_t)frames_possible * kMaxLasers * sizeof(unsigned);
  int* d_ctr2 = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s.d_frame_counts) + fc_bytes);
  ulonglong2* d_st = reinterpret_cast<ulonglong2*>(reinterpret_cast<uint8_t*>(d_ctr2) + 256);
  // (Measured: the resets and the packet times on a parallel graph branch next to the packet
  // copy make a rotation-sized batch 3 us slower, not faster -- 84 us against 81 -- so the
  // prologue stays one chain.)
  {
    // zero only what this batch can touch; first_point / start_block tables (adjacent in one
    // allocation): all rows "unset"; the header's initial values -- one launch (k_reset)
    ResetParams rp;
    size_t zero_bytes;
    if (fused) {
      rp.zero = reinterpret_cast<unsigned*>(s.d_frame_counts);
      zero_bytes = fc_bytes + 256 + (size_t)fused_tiles * 32;
    } else {
      rp.zero = reinterpret_cast<unsigned*>(s.d_zero);
      zero_bytes = (size_t)((uint8_t*)s.d_frame_counts - s.d_zero) + fc_bytes;
    }
    rp.zero_words = (long long)(zero_bytes / 4);
    rp.ones = reinterpret_cast<unsigned*>(s.d_ff);
    rp.ones_words = (long long)ctx->frame_cap * 3;
    rp.hdr = s.d_hdr;
    if (zero_bytes % 4 != 0 || (reinterpret_cast<uintptr_t>(rp.zero) & 3) != 0 ||
        (reinterpret_cast<uintptr_t>(rp.ones) & 3) != 0)
      return fail(ctx, VS_ERR_STATE, "reset block is not word aligned");
    const bool as_kernel = g.mode != GraphOps::DIRECT || ctx->reset_kernel != 0;
    if (as_kernel) {
      const long long words = std::max(rp.zero_words, rp.ones_words);
      const unsigned grid = (unsigned)std::min<long long>((words + 1023) / 1024 + 1, 148 * 8);
      void* args[] = {&rp};
      VS_CUDA(g.launch(reinterpret_cast<const void*>(&k_reset), grid, 256, 0, args, sizeof(rp)));
      ++s.n_launches;
    } else {
      // VELOSLAM_RESET_KERNEL=0 (A/B runs): the driver's memsets + a 100-byte copy
      VS_CUDA(cudaMemsetAsync(rp.zero, 0, zero_bytes, s.stream));
      VS_CUDA(cudaMemsetAsync(s.d_ff, 0xff, (size_t)ctx->frame_cap * 12, s.stream));
      BatchHeader& hi = *s.h_hdr_init;
      std::memset(&hi, 0, sizeof(hi));
      hi.first_upper_block = LLONG_MAX;
      hi.first_const_pkt = INT_MAX;
      hi.origin_at_halo = -1;
      hi.last_origin_packet = -1;
      VS_CUDA(cudaMemcpyAsync(s.d_hdr, s.h_hdr_init, sizeof(BatchHeader), cudaMemcpyHostToDevice, s.stream));
    }
  }
  if (!dev_in) {
    // stage host packets (and the pcap record headers in front of them when asked); s.d_in was
    // sized by run_batch
    const int64_t lead = pcap_t ? 64 : 0;  // keeps the payload 2-byte aligned, covers the 58 B
    const int64_t src_lead = pcap_t ? 58 : 0;
    VS_CUDA(g.copy(s.d_in + lead - src_lead, pkts - src_lead, (size_t)(payload_bytes + src_lead),
                   cudaMemcpyHostToDevice));
    d_pkts = s.d_in + lead;
    if (!pcap_t) {
      VS_CUDA(g.copy(s.d_time, pkt_time, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice));
      d_time = s.d_time;
    }
  }
  if (pcap_t) d_time = s.d_time;
  VS_CUDA(g.record(s.ev_k0));
  if (pcap_t) {
    const uint8_t* a0 = d_pkts;
    long long a1 = stride;
    int a2 = (int)n;
    long long* a3 = s.d_time;
    void* args[] = {&a0, &a1, &a2, &a3};
    VS_CUDA(g.launch(reinterpret_cast<const void*>(&k_pcap_times), (unsigned)((n + 255) / 256), 256, 0, args));
    ++s.n_launches;
  }

  vs_carry cin;
  if (halo > 0 || !carry_in)
    vs_carry_init(&cin);  // state is rebuilt from the halo
  else
    cin = *carry_in;

  if (fused) {
    if (pose_valid) {
      const long long* a0 = d_time;
      int a1 = (int)n, a2 = n_poses;
      const long long* a3 = ctx->d_pose_t;
      const double* a4 = ctx->d_pose_trv;
      double* a5 = s.d_pose_mat;
      void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5};
      VS_CUDA(g.launch(reinterpret_cast<const void*>(&k_pose_pre), (unsigned)((n + 255) / 256), 256, 0, args));
      ++s.n_launches;
    }
    DecParams dp;
    std::memset(&dp, 0, sizeof(dp));
    dp.pkts = d_pkts;
    dp.stride = stride;
    dp.total_bytes = payload_bytes;
    dp.pose_mat = s.d_pose_mat;
    dp.lut_sin = ctx->d_lut_sin;
    dp.lut_cos = ctx->d_lut_cos;
    dp.cfg = ctx->d_cfg;
    dp.n = (int)n;
    dp.halo = (int)halo;
    dp.mode = mode;
    dp.pose_valid = pose_valid ? 1 : 0;
    dp.n_tiles = (int)fused_tiles;
    dp.x = s.d_x;
    dp.y = s.d_y;
    dp.z = s.d_z;
    dp.intensity = s.d_inten;
    dp.laser = s.d_laser;
    dp.azimuth = s.d_az;
    dp.distance = s.d_dist;
    dp.t_us = s.d_t;
    dp.frame_laser_counts = s.d_frame_counts;
    dp.frame_cap = (int)ctx->frame_cap;
    dp.pkt_time = d_time;
    dp.t_base = t_base;
    dp.seg_out = s.d_seg;
    dp.st = d_st;
    dp.tile_counter = d_ctr2;
    dp.frame_first_point = s.d_frame_first;
    dp.frame_start_block = s.d_frame_start;
    dp.hdr = s.d_hdr;
    dp.carry_last_az = cin.last_azimuth;
    dp.carry_skip = cin.firing_skip;
    dp.carry_meta_inited = cin.frame_meta_inited;
    for (int k = 0; k < 3; ++k) dp.carry_origin_T[k] = cin.origin_T[k];
    VS_CUDA(g.record(s.ev_d0));
    int rc;
    if (adj == 0)
      rc = launch_decode<0, 0, 1>(ctx, s, dp, stride);
    else if (adj == 1)
      rc = launch_decode<1, 0, 1>(ctx, s, dp, stride);
    else
      rc = launch_decode<2, 0, 1>(ctx, s, dp, stride);
    if (rc != VS_OK) return rc;
    ++s.n_launches;
    VS_CUDA(g.record(s.ev_d1));
  } else {
  {
    ScanParams sp;
    sp.pkts = d_pkts;
    sp.stride = stride;
    sp.total_bytes = payload_bytes;
    sp.cfg = ctx->d_cfg;
    sp.lut_sin = ctx->d_lut_sin;
    sp.lut_cos = ctx->d_lut_cos;
    sp.n = (int)n;
    sp.mode = mode;
    sp.halo = (int)halo;
    sp.carry_last_az = cin.last_azimuth;
    sp.carry_skip = cin.firing_skip;
    sp.n_tiles = (int)scan_tiles;
    sp.stage_bytes = (int)scan_stage_bytes(stride);
    sp.pkt_seg = s.d_seg;
    sp.recs = s.d_recs;
    sp.st_map = s.d_st_map;
    sp.agg_wrap = s.d_st_wrap;
    sp.agg_cnt = s.d_st_cnt;
    sp.grp_cnt = s.d_grp_cnt;
    sp.grp_wsum = s.d_grp_wsum;
    sp.grp_wmax = s.d_grp_wmax;
    sp.tile_counter = s.d_counters + 0;
    sp.hdr = s.d_hdr;
    const size_t smem = align_up(sizeof(ScanShared), 128) + (size_t)sp.stage_bytes;
    int rc;
    if (adj == 0)
      rc = crop ? launch_scan<0, true>(ctx, s, sp, smem) : launch_scan<0, false>(ctx, s, sp, smem);
    else if (adj == 1)
      rc = crop ? launch_scan<1, true>(ctx, s, sp, smem) : launch_scan<1, false>(ctx, s, sp, smem);
    else
      rc = crop ? launch_scan<2, true>(ctx, s, sp, smem) : launch_scan<2, false>(ctx, s, sp, smem);
    if (rc != VS_OK) return rc;
    ++s.n_launches;
  }
  {
    PoseParams pp;
    pp.pkt_time = d_time;
    pp.pkt_seg = s.d_seg;
    pp.recs = s.d_recs;
    pp.pkt_off = s.d_pkt_off;
    pp.agg_wrap = s.d_st_wrap;
    pp.agg_cnt = s.d_st_cnt;
    pp.grp_cnt = s.d_grp_cnt;
    pp.grp_wsum = s.d_grp_wsum;
    pp.grp_wmax = s.d_grp_wmax;
    pp.n = (int)n;
    pp.halo = (int)halo;
    pp.mode = mode;
    pp.n_poses = index_only ? 0 : n_poses;
    pp.carry_meta_inited = cin.frame_meta_inited;
    pp.check_time = index_only ? 0 : 1;
    pp.t_base = t_base;
    pp.pose_t = ctx->d_pose_t;
    pp.pose_trv = ctx->d_pose_trv;
    for (int k = 0; k < 3; ++k) pp.carry_origin_T[k] = cin.origin_T[k];
    pp.deskew = deskew ? 1 : 0;
    pp.carry_origin_time = cin.frame_timestamp_us;
    pp.pose_mat = s.d_pose_mat;
    pp.frame_first_point = s.d_frame_first;
    pp.frame_start_block = s.d_frame_start;
    pp.frame_cap = (int)ctx->frame_cap;
    pp.hdr = s.d_hdr;
    void* args[] = {&pp};
    VS_CUDA(g.launch(reinterpret_cast<const void*>(&k_pose), (unsigned)pose_tiles, kPoseThreads, 0, args));
    ++s.n_launches;
  }
  if (!index_only) {
    DecParams dp;
    dp.pkts = d_pkts;
    dp.stride = stride;
    dp.total_bytes = payload_bytes;
    dp.pkt_seg = s.d_seg;
    dp.recs = s.d_recs;
    dp.pkt_off = s.d_pkt_off;
    dp.pose_mat = s.d_pose_mat;
    dp.lut_sin = ctx->d_lut_sin;
    dp.lut_cos = ctx->d_lut_cos;
    dp.cfg = ctx->d_cfg;
    dp.n = (int)n;
    dp.halo = (int)halo;
    dp.mode = mode;
    dp.pose_valid = pose_valid ? 1 : 0;
    dp.n_tiles = (int)dec_tiles;
    dp.stage_bytes = 0;  // set per kernel variant by launch_decode
    dp.x = s.d_x;
    dp.y = s.d_y;
    dp.z = s.d_z;
    dp.intensity = s.d_inten;
    dp.laser = s.d_laser;
    dp.azimuth = s.d_az;
    dp.distance = s.d_dist;
    dp.t_us = s.d_t;
    dp.frame_laser_counts = s.d_frame_counts;
    dp.frame_cap = (int)ctx->frame_cap;
    // Two result slots: k_decode fills every SM (2 persistent CTAs of 107 KB), so this batch's
    // cannot run before the other slot's has left -- but its event bracket could open while it
    // still queues behind it (whenever this batch's k_scan / k_pose slipped in early), and would
    // then time the wait, not the kernel.  Waiting for the other slot's k_decode HERE keeps
    // decode_ms the kernel's own duration and costs nothing the SMs would not have imposed.
    if (g.mode == GraphOps::DIRECT && ctx->n_slots == 2 && ctx->chain_decode) {
      Slot& other = ctx->slots[&s == &ctx->slots[0] ? 1 : 0];
      if (other.d1_recorded) VS_CUDA(cudaStreamWaitEvent(s.stream, other.ev_d1, 0));
    }
    VS_CUDA(g.record(s.ev_d0));
    if (dec_tiles > 0) {
      int rc;
      if (deskew) {
        if (adj == 0)
          rc = launch_decode<0, 1>(ctx, s, dp, stride);
        else if (adj == 1)
          rc = launch_decode<1, 1>(ctx, s, dp, stride);
        else
          rc = launch_decode<2, 1>(ctx, s, dp, stride);
      } else {
        if (adj == 0)
          rc = launch_decode<0, 0>(ctx, s, dp, stride);
        else if (adj == 1)
          rc = launch_decode<1, 0>(ctx, s, dp, stride);
        else
          rc = launch_decode<2, 0>(ctx, s, dp, stride);
      }
      if (rc != VS_OK) return rc;
      ++s.n_launches;
    }
    VS_CUDA(g.record(s.ev_d1));
    if (g.mode == GraphOps::DIRECT) s.d1_recorded = true;
  }
  }  // two-pass pipeline

  {
    // frame meta gather over every frame the batch can hold; rows beyond total_wraps are
    // ignored by the host
    FrameParams fp;
    std::memset(&fp, 0, sizeof(fp));  // padding included: the graph refresh compares the bytes
    fp.pkt_seg = s.d_seg;
    fp.pkt_time = d_time;
    fp.frame_start_block = s.d_frame_start;
    fp.n = (int)n;
    fp.n_frames = (int)frames_possible;
    fp.mode = mode;
    fp.frame_meta_packet = s.d_frame_meta_pkt;
    fp.frame_meta_time = s.d_frame_meta_time;
    fp.frame_skips = s.d_frame_skips;
    // one packed copy instead of seven when the eager rows are the minimum anyway
    s.eager_packed = eager_frame_rows(ctx, n) <= kEagerRows && !fused;
    fp.eager = s.eager_packed ? s.d_eager : nullptr;
    fp.hdr = s.d_hdr;
    fp.frame_first_point = s.d_frame_first;
    fp.frame_laser_counts = s.d_frame_counts;
    void* args[] = {&fp};
    VS_CUDA(g.launch(reinterpret_cast<const void*>(&k_frames), (unsigned)((frames_possible + 255) / 256), 256, 0,
                     args, sizeof(fp)));
    ++s.n_launches;
  }
  VS_CUDA(g.record(s.ev_k1));

  s.eager_rows = eager_frame_rows(ctx, n);
  if (s.eager_packed) {
    VS_CUDA(g.copy(s.h_eager, s.d_eager, sizeof(EagerBlock), cudaMemcpyDeviceToHost));
  } else {
    VS_CUDA(g.copy(s.h_hdr, s.d_hdr, sizeof(BatchHeader), cudaMemcpyDeviceToHost));
    // rows copied back right away: enough for ~3 frames per sensor rotation of the batch (the
    // pinned mirrors were sized by run_batch)
    const int rc = copy_frame_rows(ctx, s, (size_t)s.eager_rows);
    if (rc != VS_OK) return rc;
  }
  // the completion event is recorded by a host call, never by a graph node: cudaEventSynchronize
  // waits for the most recent cudaEventRecord CALL, and a node only records when it executes
  if (g.mode == GraphOps::DIRECT) VS_CUDA(cudaEventRecord(s.ev_done, s.stream));

  s.busy = true;
  s.done = false;
  s.lay_valid = false;
  s.d_time_used = d_time;
  s.n = n;
  s.halo = halo;
  s.mode = mode;
  s.flags = flags;
  s.t_base = t_base;
  s.carry_in = cin;
  s.index_only = index_only;
  return VS_OK;
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int run_batch(vs_ctx* ctx, Slot& s, const uint8_t* pkts, int64_t stride, const int64_t* pkt_time,
              int64_t n, int64_t halo, int mode, uint32_t flags, int64_t t_base,
              const vs_carry* carry_in, bool index_only) {
  const bool dev_in = (flags & VS_FLAG_DEVICE_INPUT) != 0;
  const bool pcap_t = (flags & VS_FLAG_PCAP_TIMES) != 0;
  // allocations never happen under a capture: the staging buffer and the pinned frame-row mirrors
  if (!dev_in) {
    const int64_t payload_bytes = (n - 1) * stride + kPacketBytes;
    const size_t need = (size_t)(payload_bytes + (pcap_t ? 64 : 0) + 64);
    if (need > s.d_in_bytes) {
      drop_graph(s);  // the captured copies point into the old buffer
      cudaFree(s.d_in);
      s.d_in = nullptr;
      s.d_in_bytes = 0;
      const size_t want = std::max(need, (size_t)(ctx->max_packets * stride + 128));
      VS_CUDA(cudaMalloc(&s.d_in, want));
      s.d_in_bytes = want;
    }
  }
  {
    const size_t had = s.h_frames_cap;
    const int rc = ensure_host_frames(ctx, s, (size_t)eager_frame_rows(ctx, n));
    if (rc != VS_OK) return rc;
    if (s.h_frames_cap != had) drop_graph(s);  // the captured copies point at the old mirrors
  }
  GraphOps& g = s.ops;
  g.stream = s.stream;
  g.mode = GraphOps::DIRECT;
  s.graph_issued = false;
  // graph path: small contexts, inputs the graph can copy from (device memory or page-locked host
  // memory; a pageable source cannot be captured)
  bool graph = ctx->use_graph;
  if (graph && !dev_in)
    graph = is_pinned_host(pkts - (pcap_t ? 58 : 0)) && (pcap_t || is_pinned_host(pkt_time));
  if (graph) {
    GraphKey key;
    std::memset(&key, 0, sizeof(key));
    const bool pose_valid = ctx->pose_t.size() >= 2;
    const bool crop = ctx->h_cfg.crop_returns != 0 && !index_only;
    const bool deskew = (flags & VS_FLAG_DESKEW_PER_POINT) != 0 && pose_valid && !index_only;
    key.dev_in = dev_in;
    key.pcap_t = pcap_t;
    key.fused = !ctx->two_pass && !crop && !deskew && !index_only;
    key.index_only = index_only;
    key.adj = ctx->h_cfg.adj_mode;
    key.crop = crop;
    key.deskew = deskew;
    key.pose_valid = pose_valid;
    key.has_decode = (n - halo) > 0;
    key.eager_packed = eager_frame_rows(ctx, n) <= kEagerRows && !key.fused;
    if (s.gexec && !(key == s.gkey)) drop_graph(s);
    g.nodes = &s.gnodes;
    g.sigs = &s.gsigs;
    int rc = VS_OK;
    cudaError_t ce = cudaSuccess;
    if (!s.gexec) {
      // record this batch's operations once
      s.gnodes.clear();
      s.gsigs.clear();
      s.gkey = key;
      ce = cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeRelaxed);
      if (ce == cudaSuccess) {
        g.mode = GraphOps::CAPTURE;
        rc = enqueue_batch(ctx, s, pkts, stride, pkt_time, n, halo, mode, flags, t_base, carry_in, index_only);
        ce = cudaStreamEndCapture(s.stream, &s.ggraph);
        if (rc == VS_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&s.gexec, s.ggraph, 0);
      }
    } else {
      // same sequence as the recorded one: refresh what differs (sizes, grids, arguments)
      g.mode = GraphOps::UPDATE;
      g.exec = s.gexec;
      g.cursor = 0;
      rc = enqueue_batch(ctx, s, pkts, stride, pkt_time, n, halo, mode, flags, t_base, carry_in, index_only);
      if (rc == VS_OK && g.cursor != s.gnodes.size()) ce = cudaErrorInvalidValue;
    }
    g.mode = GraphOps::DIRECT;
    if (rc == VS_OK && ce == cudaSuccess) ce = cudaEventRecord(s.ev_k0, s.stream);
    if (rc == VS_OK && ce == cudaSuccess) ce = cudaGraphLaunch(s.gexec, s.stream);
    if (rc == VS_OK && ce == cudaSuccess) ce = cudaEventRecord(s.ev_k1, s.stream);
    if (rc == VS_OK && ce == cudaSuccess) ce = cudaEventRecord(s.ev_done, s.stream);
    if (rc == VS_OK && ce == cudaSuccess) {
      ++s.graph_launches;
      s.graph_issued = true;
      return VS_OK;
    }
    // anything the graph path cannot do: drop it for this context and issue the batch directly
    if (std::getenv("VELOSLAM_GRAPH_DEBUG"))
      std::fprintf(stderr, "veloslam_b200: graph path dropped (rc=%d, %s, node %zu of %zu): %s\n", rc,
                   cudaGetErrorString(ce), g.cursor, s.gnodes.size(), ctx->err.c_str());
    cudaGetLastError();
    drop_graph(s);
    s.busy = false;
    ctx->use_graph = false;
  }
  return enqueue_batch(ctx, s, pkts, stride, pkt_time, n, halo, mode, flags, t_base, carry_in, index_only);
}

int finish_batch(vs_ctx* ctx, Slot& s) {
  VS_CUDA(cudaEventSynchronize(s.ev_done));
  if (s.eager_packed) {
    // unpack the one-copy index into the mirrors the rest of this function reads
    const EagerBlock& e = *s.h_eager;
    *s.h_hdr = e.hdr;
    std::memcpy(s.h_frame_first, e.first, sizeof(e.first));
    std::memcpy(s.h_frame_start, e.start, sizeof(e.start));
    std::memcpy(s.h_frame_meta_pkt, e.meta_pkt, sizeof(e.meta_pkt));
    std::memcpy(s.h_frame_meta_time, e.meta_time, sizeof(e.meta_time));
    std::memcpy(s.h_frame_skips, e.skips, sizeof(e.skips));
    std::memcpy(s.h_frame_counts, e.counts, sizeof(e.counts));
  }
  const BatchHeader& h = *s.h_hdr;
  if (h.frame_overflow || (int64_t)h.total_wraps + 1 > ctx->frame_cap) {
    s.busy = false;
    return fail(ctx, VS_ERR_CAPACITY, "frame table capacity exceeded (too many azimuth wraps)");
  }
  if (h.time_range_error && !s.index_only) {
    s.busy = false;
    return fail(ctx, VS_ERR_INVALID_ARG,
                "vs_submit: a packet time lies outside [t_base_us, t_base_us + 2^32 - 65536) us: "
                "the t_us column cannot hold it (split the recording or move t_base_us)");
  }
  const int W = h.total_wraps;
  const int64_t frames_possible = std::min<int64_t>(ctx->frame_cap, 12 * s.n + 1);
  if (W + 1 > s.eager_rows) {
    const size_t had = s.h_frames_cap;
    int rc = ensure_host_frames(ctx, s, (size_t)W + 1);
    if (rc != VS_OK) return rc;
    if (s.h_frames_cap != had) drop_graph(s);  // a recorded graph copies into the old mirrors
    // VS_FLAG_NO_FRAME_LIST reads the first two frames (among the rows copied with the batch) and
    // the last one only
    const int f_first = (s.halo > 0) ? h.frame_at_halo : 0;
    if ((s.flags & VS_FLAG_NO_FRAME_LIST) && W - f_first + 1 > 2 && f_first + 2 <= s.eager_rows &&
        W < frames_possible)
      rc = copy_frame_rows(ctx, s, 1, (size_t)W);
    else
      rc = copy_frame_rows(ctx, s, (size_t)std::min<int64_t>(W + 1, frames_possible));
    if (rc != VS_OK) return rc;
    VS_CUDA(cudaStreamSynchronize(s.stream));
  }
  if (s.halo > 0) {
    // the state entering the first decoded packet must not depend on what preceded the
    // halo: a constant skip map and, after it, a wrap
    const bool skip_ok = (s.mode == VS_MODE_OFFLINE) || h.first_const_pkt < s.halo;
    const int need_after = (s.mode == VS_MODE_OFFLINE) ? 0 : h.first_const_pkt + 1;
    const bool origin_ok = h.origin_at_halo > need_after;
    if (!skip_ok || !origin_ok) {
      s.busy = false;
      return fail(ctx, VS_ERR_HALO, "halo holds no azimuth wrap: cannot resolve the frame origin");
    }
  }

  const int f_lo = (s.halo > 0) ? h.frame_at_halo : 0;  // first frame with decoded points
  const int n_frames = W - f_lo + 1;
  // VS_FLAG_NO_FRAME_LIST: only the first and the last frame are assembled (vs_wait's carry-out
  // and first-frame patch read them); the table itself goes out through k_table_rows
  s.n_frames_total = n_frames;
  s.sparse_frames = (s.flags & VS_FLAG_NO_FRAME_LIST) != 0 && n_frames > 2;
  s.frames.resize(s.sparse_frames ? 2 : (size_t)n_frames);  // every entry is cleared below
  const int64_t total_points = s.index_only ? 0 : h.total_points;
  const bool any_upper_before = s.carry_in.is_hdl64 != 0;
  int64_t pose_hint = -1;
  for (int i = 0; i < n_frames; ++i) {
    if (s.sparse_frames && i == 1) i = n_frames - 1;
    const int f = f_lo + i;
    vs_frame& fr = s.frames[s.sparse_frames ? (size_t)(i != 0) : (size_t)i];
    std::memset(&fr, 0, sizeof(fr));
    long long first = s.h_frame_first[f];
    int sb = s.h_frame_start[f];
    if (i == 0) {
      // continues from the carry / the halo (or is the stream's very first frame)
      if (first < 0) first = 0;
      if (s.halo > 0 || sb < 0) sb = -1;
    }
    fr.first_point = first;
    long long next_first = total_points;
    if (i + 1 < n_frames) next_first = s.h_frame_first[f + 1];
    fr.n_points = s.index_only ? 0 : (next_first - first);
    fr.start_packet = sb < 0 ? -1 : sb / 12;
    fr.start_block = sb < 0 ? -1 : sb % 12;
    fr.closed = (i + 1 < n_frames) ? 1 : 0;
    if (fr.closed) {
      const long long close_block = s.h_frame_start[f + 1];
      fr.hdl64_order =
          (any_upper_before || (h.first_upper_block <= close_block)) ? 1 : 0;
    }
    for (int l = 0; l < kMaxLasers; ++l) fr.laser_counts[l] = s.h_frame_counts[(size_t)f * 64 + l];

    // meta: timestamp / skips / carpose
    int mp = -2;
    int64_t mt = VS_TIME_NONE;
    int sk = -1;
    if (i == 0 && (sb < 0)) {
      if (s.halo == 0 && s.carry_in.frame_meta_inited) {
        mp = -1;
        fr.timestamp_us = s.carry_in.frame_timestamp_us;
        fr.skips = s.carry_in.frame_skips;
        fr.carpose_valid = s.carry_in.frame_carpose_valid;
        std::memcpy(fr.carpose, s.carry_in.frame_carpose, sizeof(fr.carpose));
        fr.meta_packet = -1;
        continue;
      }
      // frame meta comes from the origin packet of the first decoded packet: packet 0 of a
      // fresh stream, or (halo) the packet that re-initialised the meta inside the halo
      mp = (s.halo > 0) ? h.origin_at_halo : 0;
      sk = -1;  // filled below from the device rows when available
      mt = VS_TIME_NONE;
      // timestamps of halo/first packets are fetched lazily below
      fr.meta_packet = mp;
      fr.skips = (s.halo > 0) ? -1 : s.carry_in.firing_skip;
      fr.timestamp_us = VS_TIME_NONE;  // patched by patch_first_frame()
      continue;
    }
    mp = s.h_frame_meta_pkt[f];
    mt = s.h_frame_meta_time[f];
    sk = s.h_frame_skips[f];
    if (mp >= 0) {
      fr.meta_packet = mp;
      fr.timestamp_us = mt;
      fr.skips = sk;
      bool found, valid;
      host_interpolate(ctx, mt, fr.carpose, &found, &valid, &pose_hint);
      fr.carpose_valid = valid ? 1 : 0;
    } else {
      fr.meta_packet = (mp == -3) ? -2 : mp;  // pending meta: initialised by the next batch
      fr.timestamp_us = VS_TIME_NONE;
      fr.skips = -1;
      fr.carpose_valid = 0;
    }
  }
  return VS_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* vs_version(void) { return "veloslam_b200 0.1.0 sm_100a"; }

void vs_carry_init(vs_carry* c) {
  std::memset(c, 0, sizeof(*c));
  c->last_azimuth = -1;   // HDLParser.cxx:157, 480
  c->firing_skip = 0;     // :156
  c->frame_meta_inited = 0;
  c->is_hdl64 = 0;
  c->frame_timestamp_us = VS_TIME_NONE;
  c->frame_skips = -1;
}

int vs_create(int device, int64_t max_batch_packets, int64_t max_poses, int n_slots, vs_ctx** out) {
  if (!out || max_batch_packets < 1 || max_batch_packets > (1ll << 26) || max_poses < 0 ||
      n_slots < 1 || n_slots > 2)
    return VS_ERR_INVALID_ARG;
  *out = nullptr;
  // Contexts are created one at a time: several threads bringing up contexts on one device at
  // once (HDLManager::setDevices with a device named more than once) made the runtime's own
  // first-use initialisation fail now and then.  Not a hot path.
  static std::mutex create_mutex;
  std::lock_guard<std::mutex> create_lock(create_mutex);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    g_create_err = "no such CUDA device";
    return VS_ERR_NO_DEVICE;
  }
  vs_ctx* ctx = new (std::nothrow) vs_ctx;
  if (!ctx) return VS_ERR_CAPACITY;
  auto bail = [&](int rc) {
    g_create_err = ctx->err;
    vs_destroy(ctx);
    return rc;
  };
  ctx->device = device;
  ctx->max_packets = max_batch_packets;
  ctx->max_poses = std::max<int64_t>(max_poses, 2);
  ctx->n_slots = n_slots;
  ctx->frame_cap = std::min<int64_t>(12 * max_batch_packets + 1,
                                     std::max<int64_t>(4096, max_batch_packets / 64));
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    ctx->err = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
    cudaGetLastError();
    return bail(VS_ERR_CUDA);
  }
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, device);
  if (ce != cudaSuccess) {
    ctx->err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(ce);
    cudaGetLastError();
    return bail(VS_ERR_CUDA);
  }
  if (prop.major < 10) {
    ctx->err = "veloslam_b200 is built for sm_100a only";
    return bail(VS_ERR_NO_DEVICE);
  }
  ctx->sm_count = prop.multiProcessorCount;
  {
    // the single-pass pipeline (k_decode<.., FUSED>) is opt-in: measured slower than the two-pass
    // one on B200
    const char* e = std::getenv("VELOSLAM_SINGLE_PASS");
    ctx->two_pass = !(e && e[0] == '1');
    const char* dc = std::getenv("VELOSLAM_DECODE_CHAIN");
    if (dc && dc[0] == '0') ctx->chain_decode = false;
    const char* rk = std::getenv("VELOSLAM_RESET_KERNEL");
    if (rk && (rk[0] == '0' || rk[0] == '1')) ctx->reset_kernel = rk[0] - '0';
    // rotation-sized contexts issue their batches as one CUDA graph (VELOSLAM_GRAPH=0 turns it
    // off, =1 forces it on for any size)
    const char* gr = std::getenv("VELOSLAM_GRAPH");
    ctx->use_graph = gr ? (gr[0] == '1') : (max_batch_packets <= 4096);
  }
  auto init = [&]() -> int {
    VS_CUDA(cudaMalloc(&ctx->d_cfg, sizeof(DevConfig)));
    VS_CUDA(cudaMalloc(&ctx->d_lut_sin, kLutSize * sizeof(double)));
    VS_CUDA(cudaMalloc(&ctx->d_lut_cos, kLutSize * sizeof(double)));
    VS_CUDA(cudaMalloc(&ctx->d_pose_t, ctx->max_poses * sizeof(long long)));
    VS_CUDA(cudaMalloc(&ctx->d_pose_trv, ctx->max_poses * 9 * sizeof(double)));
    // HDLParser::initLookUpTables (HDLParser.cxx:755-768), evaluated with the host libm so the
    // rotCorrection == 0 branch is bit-identical to the reference's table
    std::vector<double> ls(kLutSize), lc(kLutSize);
    for (int i = 0; i < kLutSize; ++i) {
      const double rad = to_radians_h(i / 100.0);
      lc[i] = std::cos(rad);
      ls[i] = std::sin(rad);
    }
    // Everything goes through the context's own stream and is waited for there: a device-wide
    // synchronisation (or the legacy default stream) here would collide with another context of
    // this process that is recording a CUDA graph on another thread at this moment.
    VS_CUDA(cudaStreamCreateWithFlags(&ctx->cfg_stream, cudaStreamNonBlocking));
    VS_CUDA(cudaMemcpyAsync(ctx->d_lut_sin, ls.data(), kLutSize * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->cfg_stream));
    VS_CUDA(cudaMemcpyAsync(ctx->d_lut_cos, lc.data(), kLutSize * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->cfg_stream));
    VS_CUDA(cudaMemcpyAsync(ctx->d_cfg, &ctx->h_cfg, sizeof(DevConfig), cudaMemcpyHostToDevice, ctx->cfg_stream));
    VS_CUDA(cudaStreamSynchronize(ctx->cfg_stream));  // ls / lc are pageable locals: landed before they go
    for (int i = 0; i < ctx->n_slots; ++i) {
      int rc = alloc_slot(ctx, ctx->slots[i]);
      if (rc != VS_OK) return rc;
    }
    return VS_OK;
  };
  std::memset(&ctx->h_cfg, 0, sizeof(ctx->h_cfg));
  ctx->h_cfg.laser_mask = ~0ull;  // HDLParser.cxx:172
  ctx->h_cfg.n_enabled = 64;      // :170
  update_selection(ctx->h_cfg);
  int rc = init();
  if (rc != VS_OK) return bail(rc);
  *out = ctx;
  return VS_OK;
}

void vs_destroy(vs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 2; ++i)
    if (ctx->slots[i].stream) {
      cudaStreamSynchronize(ctx->slots[i].stream);
      free_slot(ctx->slots[i]);
    }
  if (ctx->cfg_stream) cudaStreamDestroy(ctx->cfg_stream);
  cudaFree(ctx->d_cfg);
  cudaFree(ctx->d_lut_sin);
  cudaFree(ctx->d_lut_cos);
  cudaFree(ctx->d_pose_t);
  cudaFree(ctx->d_pose_trv);
  delete ctx;
}

const char* vs_last_error(vs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int vs_set_calibration(vs_ctx* ctx, const vs_laser_corr* corr, int n_rows, int n_lasers_enabled) {
  if (!ctx || !corr || n_rows < 0 || n_rows > kMaxLasers || n_lasers_enabled < 0 ||
      n_lasers_enabled > kMaxLasers)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_set_calibration: bad arguments");
  DevConfig& c = ctx->h_cfg;
  for (int i = 0; i < kMaxLasers; ++i) {
    // rows the file does not define: the reference leaves them indeterminate; match the
    // oracle's zero fill (azimuthCorrection 0 -> LUT branch, every other field 0)
    c.cal[0][i] = 1.0;
    c.cal[1][i] = 0.0;
    for (int k = 2; k < 7; ++k) c.cal[k][i] = 0.0;
  }
  for (int i = 0; i < n_rows; ++i) {
    // HDLParser.cxx:835-842
    c.cal[0][i] = std::cos(to_radians_h(corr[i].rot_correction_deg));
    c.cal[1][i] = std::sin(to_radians_h(corr[i].rot_correction_deg));
    c.cal[2][i] = corr[i].dist_correction_cm / 100.0;
    c.cal[3][i] = std::cos(to_radians_h(corr[i].vert_correction_deg));
    c.cal[4][i] = std::sin(to_radians_h(corr[i].vert_correction_deg));
    c.cal[5][i] = corr[i].vert_offset_correction_cm / 100.0;
    c.cal[6][i] = corr[i].horiz_offset_correction_cm / 100.0;
  }
  c.n_enabled = n_lasers_enabled;
  c.adj_mode = (n_lasers_enabled == 32) ? 1 : (n_lasers_enabled == 16 ? 2 : 0);
  for (int j = 0; j < kBlocks; ++j) {
    for (int dsr = 0; dsr < kReturns; ++dsr) {
      // HDLParser.cxx:946-962 with the reference's operation order (no FMA on the host:
      // this file is compiled with -ffp-contract=off)
      double ta = 0.0, b0 = 0.0, nb = 1.0;
      if (c.adj_mode == 1) {
        ta = (j * 46.08) + (dsr * 1.152);
        nb = ((j + 1) * 46.08) + (0 * 1.152);
        b0 = (j * 46.08) + (0 * 1.152);
      } else if (c.adj_mode == 2) {
        const int laser = dsr >= 16 ? dsr - 16 : dsr;
        const int fwb = dsr >= 16 ? 1 : 0;
        ta = (j * 110.592) + (laser * 2.304) + (fwb * 55.296);
        nb = ((j + 1) * 110.592) + (0 * 2.304) + (0 * 55.296);
        b0 = (j * 110.592) + (0 * 2.304) + (0 * 55.296);
      }
      c.az_ratio[j][dsr] = (ta - b0) / (nb - b0);
      c.tadj[j][dsr] = (uint16_t)std::round(ta);
    }
  }
  update_selection(c);
  cudaSetDevice(ctx->device);
  const int rc = upload_table(ctx, ctx->d_cfg, &ctx->h_cfg, sizeof(DevConfig), true, "vs_set_calibration");
  if (rc != VS_OK) return rc;
  ctx->calibrated = true;
  return VS_OK;
}

int vs_set_firing_offsets(vs_ctx* ctx, const uint16_t* off_us) {
  if (!ctx || !off_us) return fail(ctx, VS_ERR_INVALID_ARG, "vs_set_firing_offsets: bad arguments");
  if (!ctx->calibrated || ctx->h_cfg.adj_mode != 0)
    return fail(ctx, VS_ERR_STATE,
                "vs_set_firing_offsets: only for sensors without a built-in firing table "
                "(set the calibration first)");
  for (int j = 0; j < kBlocks; ++j)
    for (int d = 0; d < kReturns; ++d) ctx->h_cfg.tadj[j][d] = off_us[j * kReturns + d];
  cudaSetDevice(ctx->device);
  return upload_table(ctx, ctx->d_cfg, &ctx->h_cfg, sizeof(DevConfig), true, "vs_set_firing_offsets");
}

int vs_set_filters(vs_ctx* ctx, const vs_filters* f) {
  if (!ctx || !f || f->points_skip < 0)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_set_filters: bad arguments");
  DevConfig& c = ctx->h_cfg;
  c.laser_mask = f->laser_mask;
  c.points_skip = f->points_skip;
  c.crop_returns = f->crop_returns ? 1 : 0;
  c.crop_inside = f->crop_inside ? 1 : 0;
  for (int i = 0; i < 6; ++i) c.crop[i] = f->crop_region[i];
  update_selection(c);
  cudaSetDevice(ctx->device);
  return upload_table(ctx, ctx->d_cfg, &ctx->h_cfg, sizeof(DevConfig), true, "vs_set_filters");
}

int vs_set_poses(vs_ctx* ctx, const int64_t* t_us, const double* trv, int64_t n) {
  if (!ctx || n < 0 || (n > 0 && (!t_us || !trv)))
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_set_poses: bad arguments");
  if (n > ctx->max_poses) return fail(ctx, VS_ERR_CAPACITY, "vs_set_poses: more than max_poses");
  for (int64_t i = 1; i < n; ++i)
    if (t_us[i] <= t_us[i - 1])
      return fail(ctx, VS_ERR_INVALID_ARG, "vs_set_poses: times must be strictly increasing");
  cudaSetDevice(ctx->device);
  if (n > 0) {
    int rc = upload_table(ctx, ctx->d_pose_t, t_us, (size_t)n * sizeof(long long), false, "vs_set_poses");
    if (rc == VS_OK)
      rc = upload_table(ctx, ctx->d_pose_trv, trv, (size_t)n * 9 * sizeof(double), false, "vs_set_poses");
    if (rc != VS_OK) return rc;
  } else {
    for (int i = 0; i < ctx->n_slots; ++i)
      if (ctx->slots[i].busy && !ctx->slots[i].done)
        return fail(ctx, VS_ERR_STATE, "vs_set_poses: a batch is still in flight (vs_wait it first)");
  }
  ctx->pose_t.assign(t_us, t_us + n);
  ctx->pose_trv.assign(trv, trv + n * 9);
  return VS_OK;
}

int vs_interpolate(vs_ctx* ctx, int64_t t_us, double out_trv[9], int32_t* found, int32_t* valid) {
  if (!ctx || !out_trv) return fail(ctx, VS_ERR_INVALID_ARG, "vs_interpolate: bad arguments");
  bool f, v;
  host_interpolate(ctx, t_us, out_trv, &f, &v);
  if (found) *found = f ? 1 : 0;
  if (valid) *valid = v ? 1 : 0;
  return VS_OK;
}

static int submit_common(vs_ctx* ctx, const uint8_t* pkts, int64_t stride,
                         const int64_t* pkt_time_us, int64_t n, int64_t n_halo, int mode,
                         uint32_t flags, int64_t t_base_us, const vs_carry* carry_in,
                         uint64_t* ticket, bool index_only) {
  if (!ctx) return VS_ERR_INVALID_ARG;
  if (!index_only && !ctx->calibrated)
    return fail(ctx, VS_ERR_NOT_CALIBRATED, "corrections have not been set");
  const bool dev_in = (flags & VS_FLAG_DEVICE_INPUT) != 0;
  const bool pcap_t = (flags & VS_FLAG_PCAP_TIMES) != 0;
  if (!pkts || n < 1 || n_halo < 0 || n_halo >= n || (mode != 0 && mode != 1) || !ticket ||
      (!pcap_t && !pkt_time_us))
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_submit: bad arguments");
  if (stride < kPacketBytes || stride > 1280 || (stride & 1) || (reinterpret_cast<uintptr_t>(pkts) & 1))
    return fail(ctx, VS_ERR_INVALID_ARG,
                "vs_submit: stride must be even and in [1206, 1280], packets 2-byte aligned");
  if (n > ctx->max_packets) return fail(ctx, VS_ERR_CAPACITY, "vs_submit: batch exceeds max_batch_packets");
  if (t_base_us == VS_TIME_NONE) {
    if (dev_in) return fail(ctx, VS_ERR_INVALID_ARG, "vs_submit: device input needs an explicit t_base_us");
    if (pcap_t) {
      const uint8_t* h = pkts + n_halo * stride - 58;
      uint32_t sec, usec;
      std::memcpy(&sec, h, 4);
      std::memcpy(&usec, h + 4, 4);
      t_base_us = ((int64_t)sec + 8 * 3600) * 1000000ll + usec;
    } else {
      t_base_us = pkt_time_us[n_halo];
    }
  }
  if (!dev_in && !pcap_t && !index_only) {
    // the t_us column is u32 microseconds after t_base (+ a u16 firing offset), HDLParser.cxx:969-970
    for (int64_t i = n_halo; i < n; ++i)
      if ((uint64_t)(pkt_time_us[i] - t_base_us) >= kTimeSpanMax)
        return fail(ctx, VS_ERR_INVALID_ARG,
                    "vs_submit: a packet time lies outside [t_base_us, t_base_us + 2^32 - 65536) us: "
                    "the t_us column cannot hold it (split the recording or move t_base_us)");
  }
  const uint64_t tk = ctx->next_ticket;
  Slot& s = ctx->slots[tk % (uint64_t)ctx->n_slots];
  if (s.busy && !s.done) {
    // the slot still holds an unfinished batch: the caller must vs_wait it first
    return fail(ctx, VS_ERR_STATE, "vs_submit: result slot busy (vs_wait the previous ticket)");
  }
  cudaSetDevice(ctx->device);
  int rc = run_batch(ctx, s, pkts, stride, pkt_time_us, n, n_halo, mode, flags, t_base_us, carry_in,
                     index_only);
  if (rc != VS_OK) return rc;
  s.ticket = tk;
  ctx->next_ticket = tk + 1;
  *ticket = tk;
  return VS_OK;
}

int vs_submit(vs_ctx* ctx, const uint8_t* pkts, int64_t stride, const int64_t* pkt_time_us,
              int64_t n, int64_t n_halo, int mode, uint32_t flags, int64_t t_base_us,
              const vs_carry* carry_in, uint64_t* ticket) {
  return submit_common(ctx, pkts, stride, pkt_time_us, n, n_halo, mode, flags, t_base_us, carry_in,
                       ticket, false);
}

int vs_wait(vs_ctx* ctx, uint64_t ticket, vs_result* out) {
  if (!ctx || !out) return fail(ctx, VS_ERR_INVALID_ARG, "vs_wait: bad arguments");
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || s.ticket != ticket) return fail(ctx, VS_ERR_STATE, "vs_wait: unknown ticket");
  cudaSetDevice(ctx->device);
  if (!s.done) {
    int rc = finish_batch(ctx, s);
    if (rc != VS_OK) return rc;
    const BatchHeader& h = *s.h_hdr;
    vs_result& r = s.result;
    std::memset(&r, 0, sizeof(r));
    r.n_packets = s.n - s.halo;
    r.n_points = s.index_only ? 0 : h.total_points;
    r.n_frames = (int32_t)s.n_frames_total;
    r.n_closed = r.n_frames - 1;
    r.x = s.d_x;
    r.y = s.d_y;
    r.z = s.d_z;
    r.intensity = s.d_inten;
    r.laser = s.d_laser;
    r.azimuth = s.d_az;
    r.distance = s.d_dist;
    r.t_us = s.d_t;
    r.t_base_us = s.t_base;
    r.first_upper_block = (h.first_upper_block == LLONG_MAX) ? -1 : h.first_upper_block;
    r.n_kernel_launches = s.n_launches;
    r.reserved = s.graph_launches;  // batches of this slot issued as a CUDA graph so far
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1);
    r.gpu_ms = ms;
    ms = 0.f;
    // (a graph-issued batch is timed as a whole: gpu_ms then includes its copies, decode_ms is 0)
    if (!s.index_only && !s.graph_issued) cudaEventElapsedTime(&ms, s.ev_d0, s.ev_d1);
    r.decode_ms = ms;

    // first frame of a fresh stream / of a halo shard: meta from a packet of this batch
    vs_frame& f0 = s.frames[0];
    if (f0.meta_packet >= 0 && f0.timestamp_us == VS_TIME_NONE) {
      long long mt = VS_TIME_NONE;
      // (on the batch's own stream, like every other copy of this library: never the legacy
      // default stream, which other contexts of the process would have to agree with)
      VS_CUDA(cudaMemcpyAsync(&mt, s.d_time_used + f0.meta_packet, sizeof(mt), cudaMemcpyDeviceToHost, s.stream));
      VS_CUDA(cudaStreamSynchronize(s.stream));
      if (mt != VS_TIME_NONE) {
        f0.timestamp_us = mt;
        bool found, valid;
        host_interpolate(ctx, mt, f0.carpose, &found, &valid);
        f0.carpose_valid = valid ? 1 : 0;
      }
      if (s.halo > 0) {
        PktSeg sg;
        VS_CUDA(cudaMemcpyAsync(&sg, s.d_seg + f0.meta_packet, sizeof(sg), cudaMemcpyDeviceToHost, s.stream));
        VS_CUDA(cudaStreamSynchronize(s.stream));
        if (s.mode == VS_MODE_OFFLINE) {
          const int wm = (sg.x >> 4) & 0xfff;  // the frame starts at the packet's last wrap
          f0.skips = wm ? (31 - __builtin_clz((unsigned)wm)) : 0;
        } else {
          f0.skips = sg.x & 15;
        }
      }
    }

    // carry-out (HDLParser.cxx:196-216 state after the last packet)
    vs_carry& co = r.carry_out;
    co = s.carry_in;
    co.last_azimuth = h.last_azimuth;
    co.firing_skip = h.firing_skip_out;
    co.is_hdl64 = (s.carry_in.is_hdl64 || h.first_upper_block != LLONG_MAX) ? 1 : 0;
    const vs_frame& open = s.frames.back();
    const bool pending = (s.mode == VS_MODE_STREAMING) && h.last_has_wrap;
    co.frame_meta_inited = pending ? 0 : 1;
    if (!pending) {
      co.frame_timestamp_us = open.timestamp_us;
      co.frame_skips = open.skips;
      co.frame_carpose_valid = open.carpose_valid;
      std::memcpy(co.frame_carpose, open.carpose, sizeof(co.frame_carpose));
      if (ctx->pose_t.size() >= 2) {
        // k_pose ran: it wrote the frame origin the last packet used (or, offline, started)
        for (int k = 0; k < 3; ++k) co.origin_T[k] = h.carry_origin_T[k];
      } else if (h.last_origin_packet >= 0) {
        // fewer than two poses: carpose->T is what interpolateTransform wrote at the frame's
        // meta packet (zeros, or the one-sample extrapolation)
        double trv[9];
        bool found, valid;
        host_interpolate(ctx, h.last_origin_time, trv, &found, &valid);
        for (int k = 0; k < 3; ++k) co.origin_T[k] = trv[k];
      }
    } else {
      co.frame_timestamp_us = VS_TIME_NONE;
      co.frame_skips = -1;
      co.frame_carpose_valid = 0;
      std::memset(co.frame_carpose, 0, sizeof(co.frame_carpose));
    }
    co.frames_closed = s.carry_in.frames_closed + r.n_closed;
    co.points_emitted = s.carry_in.points_emitted + r.n_points;
    co.packets_seen = s.carry_in.packets_seen + r.n_packets;
    s.done = true;
  }
  *out = s.result;
  out->frames = (s.flags & VS_FLAG_NO_FRAME_LIST) ? nullptr : s.frames.data();
  return VS_OK;
}

int vs_fetch_points(vs_ctx* ctx, uint64_t ticket, int64_t first, int64_t count, float* x, float* y,
                    float* z, uint8_t* intensity, uint8_t* laser, uint16_t* azimuth,
                    uint16_t* distance, uint32_t* t_us) {
  if (!ctx) return VS_ERR_INVALID_ARG;
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || !s.done || s.ticket != ticket)
    return fail(ctx, VS_ERR_STATE, "vs_fetch_points: ticket not finished (vs_wait first)");
  if (first < 0 || count < 0 || first + count > s.result.n_points)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_fetch_points: range outside the batch");
  if (count == 0) return VS_OK;
  cudaSetDevice(ctx->device);
  const size_t c = (size_t)count;
  if (x) VS_CUDA(cudaMemcpyAsync(x, s.d_x + first, c * 4, cudaMemcpyDeviceToHost, s.stream));
  if (y) VS_CUDA(cudaMemcpyAsync(y, s.d_y + first, c * 4, cudaMemcpyDeviceToHost, s.stream));
  if (z) VS_CUDA(cudaMemcpyAsync(z, s.d_z + first, c * 4, cudaMemcpyDeviceToHost, s.stream));
  if (intensity)
    VS_CUDA(cudaMemcpyAsync(intensity, s.d_inten + first, c, cudaMemcpyDeviceToHost, s.stream));
  if (laser) VS_CUDA(cudaMemcpyAsync(laser, s.d_laser + first, c, cudaMemcpyDeviceToHost, s.stream));
  if (azimuth)
    VS_CUDA(cudaMemcpyAsync(azimuth, s.d_az + first, c * 2, cudaMemcpyDeviceToHost, s.stream));
  if (distance)
    VS_CUDA(cudaMemcpyAsync(distance, s.d_dist + first, c * 2, cudaMemcpyDeviceToHost, s.stream));
  if (t_us) VS_CUDA(cudaMemcpyAsync(t_us, s.d_t + first, c * 4, cudaMemcpyDeviceToHost, s.stream));
  VS_CUDA(cudaStreamSynchronize(s.stream));
  return VS_OK;
}

namespace {
const int kBeamLutHost[64] = {38, 39, 42, 43, 32, 33, 36, 37, 40, 41, 46, 47, 50, 51, 54, 55,
                              44, 45, 48, 49, 52, 53, 58, 59, 62, 63, 34, 35, 56, 57, 60, 61,
                              6,  7,  10, 11, 0,  1,  4,  5,  8,  9,  14, 15, 18, 19, 22, 23,
                              12, 13, 16, 17, 20, 21, 26, 27, 30, 31, 2,  3,  24, 25, 28, 29};

// Grows a device buffer.  cudaFree + cudaMalloc stall the whole device (measured: 5 to 640 ms
// under a live 10 Hz stream), so a buffer starts at `reserve` -- what a full batch of the context
// needs -- and, when a request still exceeds it, grows by half again instead of to the exact fit.
int ensure_device_bytes(vs_ctx* ctx, uint8_t** buf, size_t* have, size_t need, size_t reserve) {
  if (need <= *have) return VS_OK;
  const size_t want = std::max(need + (*have ? need / 2 : 0), reserve);
  cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  if (cudaMalloc(buf, want) != cudaSuccess) {
    cudaGetLastError();
    *buf = nullptr;
    VS_CUDA(cudaMalloc(buf, need));  // no room for the head-room: the exact fit
    *have = need;
    return VS_OK;
  }
  *have = want;
  return VS_OK;
}
}  // namespace

int vs_layout_frames(vs_ctx* ctx, uint64_t ticket, const uint32_t* carried_counts, int xyzi_stride,
                     int with_meta, vs_layout* out) {
  if (!ctx || !out || (xyzi_stride != 16 && xyzi_stride != 32))
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_layout_frames: bad arguments");
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || !s.done || s.ticket != ticket)
    return fail(ctx, VS_ERR_STATE, "vs_layout_frames: ticket not finished (vs_wait first)");
  if (s.index_only) return fail(ctx, VS_ERR_STATE, "vs_layout_frames: the batch decoded no points");
  if (s.flags & VS_FLAG_NO_FRAME_LIST)
    return fail(ctx, VS_ERR_STATE, "vs_layout_frames: the batch was submitted with VS_FLAG_NO_FRAME_LIST");
  cudaSetDevice(ctx->device);
  const BatchHeader& h = *s.h_hdr;
  const int n_frames = (int)s.frames.size();
  const int f_lo = (s.halo > 0) ? h.frame_at_halo : 0;
  uint64_t carried_total = 0;
  if (carried_counts)
    for (int l = 0; l < kMaxLasers; ++l) carried_total += carried_counts[l];
  const int64_t n_slots = s.result.n_points + (int64_t)carried_total;

  // host view of the rows (the same arithmetic as k_layout_rows)
  s.lay_rows.assign((size_t)n_frames, vs_frame_rows());
  for (int i = 0; i < n_frames; ++i) {
    const vs_frame& fr = s.frames[(size_t)i];
    vs_frame_rows& rw = s.lay_rows[(size_t)i];
    rw.first_slot = (i == 0) ? 0 : fr.first_point + (int64_t)carried_total;
    uint64_t acc = 0;
    for (int r = 0; r < kMaxLasers; ++r) {
      const int laser = fr.hdl64_order ? kBeamLutHost[r] : r;
      const uint32_t car = (i == 0 && carried_counts) ? carried_counts[laser] : 0u;
      rw.row_laser[r] = laser;
      rw.row_carried[r] = car;
      rw.row_start[r] = (uint32_t)acc;
      rw.row_count[r] = car + fr.laser_counts[laser];
      acc += (uint64_t)car + fr.laser_counts[laser];
    }
    if (acc >= (1ull << 32))
      return fail(ctx, VS_ERR_CAPACITY, "vs_layout_frames: a frame holds 2^32 points or more");
    rw.n_slots = (int64_t)acc;
  }

  // sized once for a full batch of this context plus the points an open rotation carries in
  // (one 5 Hz HDL-64E rotation), so that a live stream never re-allocates
  const size_t reserve_slots = (size_t)std::min<int64_t>(ctx->max_packets, 1 << 16) * 384 + 300000;
  int rc = ensure_device_bytes(ctx, &s.d_lay_xyzi, &s.lay_xyzi_bytes,
                               (size_t)std::max<int64_t>(n_slots, 1) * (size_t)xyzi_stride,
                               reserve_slots * (size_t)xyzi_stride);
  if (rc != VS_OK) return rc;
  if (with_meta) {
    rc = ensure_device_bytes(ctx, &s.d_lay_meta, &s.lay_meta_bytes, (size_t)std::max<int64_t>(n_slots, 1) * 12,
                             reserve_slots * 12);
    if (rc != VS_OK) return rc;
  }
  const int64_t n_dec = s.n - s.halo;
  const int n_chunks = (int)((n_dec + kLayChunk - 1) / kLayChunk);
  const size_t max_chunks = (size_t)((ctx->max_packets + kLayChunk - 1) / kLayChunk);
  rc = ensure_device_bytes(ctx, &s.d_lay_st, &s.lay_st_bytes, 256 + (size_t)n_chunks * 32 * 8,
                           256 + max_chunks * 32 * 8);
  if (rc != VS_OK) return rc;
  if (!s.d_lay_rows) VS_CUDA(cudaMalloc(&s.d_lay_rows, (size_t)ctx->frame_cap * kMaxLasers * 8));
  if (!s.ev_l0) {
    VS_CUDA(cudaEventCreate(&s.ev_l0));
    VS_CUDA(cudaEventCreate(&s.ev_l1));
  }

  VS_CUDA(cudaEventRecord(s.ev_l0, s.stream));
  int launches = 0;
  if (s.result.n_points > 0) {
    VS_CUDA(cudaMemsetAsync(s.d_lay_st, 0, 256 + (size_t)n_chunks * 32 * 8, s.stream));
    RowsParams rp;
    rp.frame_first = s.d_frame_first;
    rp.frame_start = s.d_frame_start;
    rp.frame_counts = s.d_frame_counts;
    rp.hdr = s.d_hdr;
    rp.f_lo = f_lo;
    rp.n_frames = n_frames;
    rp.carry_is_hdl64 = s.carry_in.is_hdl64 ? 1 : 0;
    for (int l = 0; l < kMaxLasers; ++l) rp.carried[l] = carried_counts ? carried_counts[l] : 0u;
    rp.carried_total = carried_total;
    rp.row_abs = s.d_lay_rows;
    k_layout_rows<<<(unsigned)((n_frames + 127) / 128), 128, 0, s.stream>>>(rp);
    VS_CUDA(cudaGetLastError());
    LayoutParams lp;
    lp.pkt_seg = s.d_seg;
    lp.recs = s.d_recs;
    lp.pkt_off = s.d_pkt_off;
    lp.x = s.d_x;
    lp.y = s.d_y;
    lp.z = s.d_z;
    lp.inten = s.d_inten;
    lp.az = s.d_az;
    lp.dist = s.d_dist;
    lp.cfg = ctx->d_cfg;
    lp.row_abs = s.d_lay_rows;
    lp.st = reinterpret_cast<unsigned long long*>(s.d_lay_st + 256);
    lp.chunk_counter = reinterpret_cast<int*>(s.d_lay_st);
    lp.n = (int)s.n;
    lp.halo = (int)s.halo;
    lp.n_chunks = n_chunks;
    lp.f_lo = f_lo;
    lp.adj = ctx->h_cfg.adj_mode;
    lp.xyzi = s.d_lay_xyzi;
    lp.xyzi_stride = xyzi_stride;
    lp.meta = with_meta ? s.d_lay_meta : nullptr;
    if (!ctx->layout_attr_set) {
      VS_CUDA(cudaFuncSetAttribute(k_layout, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(LayShared)));
      ctx->layout_attr_set = true;
    }
    const int lay_grid = std::min(n_chunks, 2 * ctx->sm_count);  // persistent CTAs, 2 per SM
    k_layout<<<(unsigned)lay_grid, kLayThreads, sizeof(LayShared), s.stream>>>(lp);
    VS_CUDA(cudaGetLastError());
    launches = 2;
  }
  VS_CUDA(cudaEventRecord(s.ev_l1, s.stream));
  s.lay_valid = true;
  s.lay_inflight = true;
  vs_layout& L = s.layout;
  std::memset(&L, 0, sizeof(L));
  L.xyzi = s.d_lay_xyzi;
  L.meta = with_meta ? s.d_lay_meta : nullptr;
  L.n_slots = n_slots;
  L.rows = s.lay_rows.data();
  L.n_frames = n_frames;
  L.xyzi_stride = xyzi_stride;
  L.n_kernel_launches = launches;
  *out = L;
  return VS_OK;
}

int vs_fetch_layout(vs_ctx* ctx, uint64_t ticket, int64_t first_slot, int64_t n_slots, void* xyzi_host,
                    void* meta_host) {
  if (!ctx) return VS_ERR_INVALID_ARG;
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || !s.done || s.ticket != ticket || !s.lay_valid)
    return fail(ctx, VS_ERR_STATE, "vs_fetch_layout: no layout for this ticket (vs_layout_frames first)");
  if (first_slot < 0 || n_slots < 0 || first_slot + n_slots > s.layout.n_slots)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_fetch_layout: range outside the layout");
  if (meta_host && !s.layout.meta)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_fetch_layout: the layout was built without PointMeta");
  if (n_slots == 0) return VS_OK;
  cudaSetDevice(ctx->device);
  const size_t st = (size_t)s.layout.xyzi_stride;
  if (xyzi_host)
    VS_CUDA(cudaMemcpyAsync(xyzi_host, s.d_lay_xyzi + (size_t)first_slot * st, (size_t)n_slots * st,
                            cudaMemcpyDeviceToHost, s.stream));
  if (meta_host)
    VS_CUDA(cudaMemcpyAsync(meta_host, s.d_lay_meta + (size_t)first_slot * 12, (size_t)n_slots * 12,
                            cudaMemcpyDeviceToHost, s.stream));
  s.lay_inflight = true;
  return VS_OK;
}

int vs_sync(vs_ctx* ctx, uint64_t ticket, float* layout_ms) {
  if (!ctx) return VS_ERR_INVALID_ARG;
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || s.ticket != ticket) return fail(ctx, VS_ERR_STATE, "vs_sync: unknown ticket");
  cudaSetDevice(ctx->device);
  VS_CUDA(cudaStreamSynchronize(s.stream));
  s.lay_inflight = false;
  if (layout_ms) {
    *layout_ms = 0.f;
    if (s.lay_valid) cudaEventElapsedTime(layout_ms, s.ev_l0, s.ev_l1);
  }
  return VS_OK;
}

int vs_read_frame_information(vs_ctx* ctx, const uint8_t* pkts, int64_t stride,
                              const int64_t* pkt_time_us, int64_t n, uint32_t flags,
                              int32_t* start_packet, int32_t* skips, int64_t* timestamp_us,
                              int32_t cap, int32_t* n_frames) {
  if (!ctx || !n_frames || cap < 0 || n < 1)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_read_frame_information: bad arguments");
  // arrays longer than one batch are indexed in chunks; the carry (lastAzimuth) links them
  vs_carry c;
  vs_carry_init(&c);
  int64_t total = 0;
  for (int64_t base = 0; base < n; base += ctx->max_packets) {
    const int64_t m = std::min<int64_t>(ctx->max_packets, n - base);
    uint64_t tk = 0;
    int rc = submit_common(ctx, pkts + base * stride, stride, pkt_time_us ? pkt_time_us + base : nullptr,
                           m, 0, VS_MODE_OFFLINE, flags, 0, &c, &tk, true);
    if (rc != VS_OK) return rc;
    vs_result r;
    rc = vs_wait(ctx, tk, &r);
    if (rc != VS_OK) return rc;
    // entry 0 of a later chunk continues the previous chunk's open frame
    for (int i = (base == 0 ? 0 : 1); i < r.n_frames; ++i, ++total) {
      if (total >= cap) continue;
      const vs_frame& f = r.frames[i];
      start_packet[total] = (total == 0) ? 0 : (int32_t)(base + f.start_packet);
      skips[total] = (total == 0) ? 0 : f.start_block;
      timestamp_us[total] = f.timestamp_us;
    }
    c = r.carry_out;
  }
  *n_frames = (int32_t)total;
  return VS_OK;
}

// ---- packet-range sharding (host arithmetic only) ----------------------------------------------
int vs_shard_range(int64_t n_packets, int32_t world, int32_t rank, int64_t halo, int64_t* first,
                   int64_t* n_halo, int64_t* end) {
  if (n_packets < 0 || world < 1 || rank < 0 || rank >= world || halo < 0 || !first || !n_halo || !end)
    return VS_ERR_INVALID_ARG;
  // 128-bit products: n_packets * rank does not overflow for any int64 n_packets
  const int64_t f = (int64_t)(((__int128)n_packets * rank) / world);
  const int64_t e = (int64_t)(((__int128)n_packets * (rank + 1)) / world);
  *first = f;
  *end = e;
  *n_halo = std::min(halo, f);
  return VS_OK;
}

int vs_frame_table_rows(const vs_frame* frames, int32_t n_frames, int32_t rank, int64_t first_packet,
                        int64_t n_halo, int64_t* rows) {
  if (n_frames < 0 || (n_frames > 0 && (!frames || !rows))) return VS_ERR_INVALID_ARG;
  const int64_t base = first_packet - n_halo;  // global index of packet 0 of the submitted array
  for (int32_t i = 0; i < n_frames; ++i) {
    const vs_frame& f = frames[i];
    int64_t* r = rows + (size_t)i * VS_FRAME_ROW_COLS;
    r[0] = f.n_points;
    r[1] = f.first_point;
    r[2] = f.start_packet >= 0 ? f.start_packet + base : -1;
    r[3] = f.start_block;
    r[4] = f.timestamp_us;
    r[5] = f.skips;
    r[6] = f.closed;
    r[7] = f.hdl64_order;
    r[8] = f.meta_packet >= 0 ? f.meta_packet + base : f.meta_packet;
    r[9] = rank;
  }
  return VS_OK;
}

int vs_frame_table_rows_device(vs_ctx* ctx, uint64_t ticket, int32_t rank, int64_t first_packet,
                               int64_t* d_rows, int64_t cap_rows) {
  if (!ctx || !d_rows || cap_rows < 1)
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_frame_table_rows_device: bad arguments");
  Slot& s = ctx->slots[ticket % (uint64_t)ctx->n_slots];
  if (!s.busy || s.ticket != ticket) return fail(ctx, VS_ERR_STATE, "vs_frame_table_rows_device: unknown ticket");
  cudaSetDevice(ctx->device);
  TableRowsParams tp;
  std::memset(&tp, 0, sizeof(tp));
  tp.hdr = s.d_hdr;
  tp.frame_first_point = s.d_frame_first;
  tp.frame_start_block = s.d_frame_start;
  tp.frame_meta_packet = s.d_frame_meta_pkt;
  tp.frame_meta_time = s.d_frame_meta_time;
  tp.frame_skips = s.d_frame_skips;
  tp.pkt_seg = s.d_seg;
  tp.pkt_time = s.d_time_used;
  tp.rows = reinterpret_cast<long long*>(d_rows);
  tp.cap_rows = cap_rows;
  tp.base = first_packet - s.halo;
  tp.carry_timestamp = s.carry_in.frame_timestamp_us;
  tp.carry_meta_inited = s.carry_in.frame_meta_inited;
  tp.carry_skips = s.carry_in.frame_skips;
  tp.carry_firing_skip = s.carry_in.firing_skip;
  tp.carry_is_hdl64 = s.carry_in.is_hdl64;
  tp.halo = (int)s.halo;
  tp.mode = s.mode;
  tp.index_only = s.index_only ? 1 : 0;
  tp.rank = rank;
  const int64_t frames_possible = std::min<int64_t>(ctx->frame_cap, 12 * s.n + 1);
  const unsigned grid = (unsigned)std::min<int64_t>((std::min(frames_possible, cap_rows) + 255) / 256 + 1, 1024);
  void* args[] = {&tp};
  VS_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(&k_table_rows), dim3(grid, 1, 1), dim3(256, 1, 1), args,
                           0, s.stream));
  return VS_OK;
}

int vs_stitch_frame_tables(const int64_t* rows, const int32_t* rows_per_rank, int32_t world,
                           int64_t rank_stride_rows, vs_global_frame* frames, int32_t frame_cap,
                           vs_frame_segment* segs, int32_t seg_cap, int32_t* n_frames, int32_t* n_segs) {
  if (!rows_per_rank || world < 1 || !n_frames || !n_segs || frame_cap < 0 || seg_cap < 0 ||
      rank_stride_rows < 0)
    return VS_ERR_INVALID_ARG;
  int64_t nf = 0, ns = 0;
  const int64_t* r = rows;
  for (int32_t g = 0; g < world; ++g) {
    if (rank_stride_rows > 0) {
      if (rows_per_rank[g] > rank_stride_rows) return VS_ERR_INVALID_ARG;
      r = rows + (size_t)g * (size_t)rank_stride_rows * VS_FRAME_ROW_COLS;
    }
    for (int32_t i = 0; i < rows_per_rank[g]; ++i, r += VS_FRAME_ROW_COLS) {
      if (!rows) return VS_ERR_INVALID_ARG;
      if (ns < seg_cap) {
        vs_frame_segment& sg = segs[ns];
        sg.rank = (int32_t)r[9];
        sg.reserved = 0;
        sg.first_point = r[1];
        sg.n_points = r[0];
      }
      const bool cont = i == 0 && g > 0 && nf > 0;  // continues the previous rank's open frame
      if (cont) {
        if (nf - 1 < frame_cap) {
          vs_global_frame& f = frames[nf - 1];
          f.n_segments += 1;
          f.n_points += r[0];
          f.closed = (int32_t)r[6];
          f.hdl64_order = (int32_t)r[7];
          // both sides rebuilt the frame's meta (the later one from its halo): they must agree
          if (r[4] != f.timestamp_us) f.timestamp_mismatch = 1;
        }
      } else {
        if (nf < frame_cap) {
          vs_global_frame& f = frames[nf];
          f.n_points = r[0];
          f.start_packet = r[2];
          f.timestamp_us = r[4];
          f.start_block = (int32_t)r[3];
          f.skips = (int32_t)r[5];
          f.closed = (int32_t)r[6];
          f.hdl64_order = (int32_t)r[7];
          f.first_segment = (int32_t)ns;
          f.n_segments = 1;
          f.timestamp_mismatch = 0;
          f.reserved = 0;
        }
        ++nf;
      }
      ++ns;
    }
  }
  *n_frames = (int32_t)nf;
  *n_segs = (int32_t)ns;
  return (nf > frame_cap || ns > seg_cap) ? VS_ERR_CAPACITY : VS_OK;
}

int vs_host_alloc(uint64_t bytes, void** out) {
  if (!out) return VS_ERR_INVALID_ARG;
  *out = nullptr;
  return cudaMallocHost(out, (size_t)bytes) == cudaSuccess ? VS_OK : VS_ERR_CUDA;
}
void vs_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int vs_device_alloc(vs_ctx* ctx, uint64_t bytes, void** out_dev) {
  if (!ctx || !out_dev || bytes == 0) return fail(ctx, VS_ERR_INVALID_ARG, "vs_device_alloc: bad arguments");
  *out_dev = nullptr;
  cudaSetDevice(ctx->device);
  VS_CUDA(cudaMalloc(out_dev, (size_t)bytes));
  return VS_OK;
}
void vs_device_free(vs_ctx* ctx, void* dev) {
  if (!ctx || !dev) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < ctx->n_slots; ++i)
    if (ctx->slots[i].stream) cudaStreamSynchronize(ctx->slots[i].stream);  // batches may read it
  cudaFree(dev);
}
int vs_device_upload(vs_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes) {
  if (!ctx || !dst_dev || !src_host) return fail(ctx, VS_ERR_INVALID_ARG, "vs_device_upload: bad arguments");
  cudaSetDevice(ctx->device);
  // complete (not merely staged) before any slot stream reads it
  cudaStream_t st = ctx->cfg_stream;
  VS_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, st));
  VS_CUDA(cudaStreamSynchronize(st));
  return VS_OK;
}

int vs_solve_packet_times(vs_ctx* ctx, const uint8_t* pkts, int64_t stride, int64_t n, uint32_t flags,
                          int64_t now_us, vs_time_solver* state, int64_t* out_time_us) {
  if (!ctx || !pkts || !state || !out_time_us || n < 1 || stride < kPacketBytes || (stride & 1))
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_solve_packet_times: bad arguments");
  const bool dev = (flags & VS_FLAG_DEVICE_INPUT) != 0;
  Slot& s = ctx->slots[0];
  if (s.busy && !s.done) return fail(ctx, VS_ERR_STATE, "vs_solve_packet_times: slot 0 is busy");
  cudaSetDevice(ctx->device);
  // scratch of slot 0: gps fields in the segment records' array, times in d_time
  uint32_t* d_gps = reinterpret_cast<uint32_t*>(s.d_seg);
  std::vector<uint32_t> h_gps;
  for (int64_t base = 0; base < n; base += ctx->max_packets) {
    const int64_t m = std::min<int64_t>(ctx->max_packets, n - base);
    const int tiles = (int)((m + kGpsThreads * kGpsItems - 1) / (kGpsThreads * kGpsItems));
    VS_CUDA(cudaMemsetAsync(s.d_st_map, 0, (size_t)tiles * 8, s.stream));
    VS_CUDA(cudaMemsetAsync(s.d_counters, 0, 64, s.stream));
    if (dev) {
      k_gps_gather<<<(unsigned)((m + 255) / 256), 256, 0, s.stream>>>(pkts + base * stride, stride, (int)m, d_gps);
      VS_CUDA(cudaGetLastError());
    } else {
      h_gps.resize((size_t)m);
      for (int64_t i = 0; i < m; ++i) std::memcpy(&h_gps[(size_t)i], pkts + (base + i) * stride + 1200, 4);
      VS_CUDA(cudaMemcpyAsync(d_gps, h_gps.data(), (size_t)m * 4, cudaMemcpyHostToDevice, s.stream));
    }
    uint32_t first_gps = 0;
    if (!state->inited) {
      // hdlOffset = now - hdlHourTime - gps of the first packet (TimeSolver.cxx:36-42)
      if (dev)
        VS_CUDA(cudaMemcpyAsync(&first_gps, d_gps, 4, cudaMemcpyDeviceToHost, s.stream));
      else
        first_gps = h_gps[0];
      VS_CUDA(cudaStreamSynchronize(s.stream));
      state->base_us = now_us - (int64_t)first_gps;
      state->last_report = 0;
      state->inited = 1;
    }
    GpsParams gp;
    gp.gps = d_gps;
    gp.n = (int)m;
    gp.last_report = state->last_report;
    gp.base_us = state->base_us;
    gp.t_out = dev ? reinterpret_cast<long long*>(out_time_us + base) : s.d_time;
    gp.st = s.d_st_map;
    gp.tile_counter = s.d_counters;
    gp.total_wraps = reinterpret_cast<unsigned long long*>(s.d_counters + 8);
    k_gps_times<<<tiles, kGpsThreads, 0, s.stream>>>(gp);
    VS_CUDA(cudaGetLastError());
    unsigned long long wraps = 0;
    uint32_t last = 0;
    VS_CUDA(cudaMemcpyAsync(&wraps, gp.total_wraps, 8, cudaMemcpyDeviceToHost, s.stream));
    VS_CUDA(cudaMemcpyAsync(&last, d_gps + (m - 1), 4, cudaMemcpyDeviceToHost, s.stream));
    if (!dev)
      VS_CUDA(cudaMemcpyAsync(out_time_us + base, s.d_time, (size_t)m * 8, cudaMemcpyDeviceToHost, s.stream));
    VS_CUDA(cudaStreamSynchronize(s.stream));
    state->base_us += (int64_t)wraps * 3600000000ll;  // hdlHourTime += 1 h per wrap
    state->last_report = last;
  }
  return VS_OK;
}

namespace {
// xyz2llh of the ENU origin (CoordiTran.cpp:82-150), evaluated once per call on the host
void origin_rotation(const double org[3], double R[9]) {
  const double pi = 3.141592653589793;
  const double x = org[0], y = org[1], z = org[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double a = 6378137.0000, b = 6356752.3142;
  const double e = std::sqrt(1 - (b / a) * (b / a));
  const double b2 = b * b, e2 = e * e, ep = e * (a / b);
  const double r = std::sqrt(x2 + y2), r2 = r * r;
  const double E2 = a * a - b * b;
  const double F = 54 * b2 * z2;
  const double G = r2 + (1 - e2) * z2 - e2 * E2;
  const double c = (e2 * e2 * F * r2) / (G * G * G);
  const double s = std::pow(double(1 + c + std::sqrt(c * c + 2 * c)), double(1.0 / 3.0));
  const double P = F / (3 * (s + 1 / s + 1) * (s + 1 / s + 1) * G * G);
  const double Q = std::sqrt(1 + 2 * e2 * e2 * P);
  const double ro = -(P * e2 * r) / (1 + Q) +
                    std::sqrt((a * a / 2) * (1 + 1 / Q) - (P * (1 - e2) * z2) / (Q * (1 + Q)) - P * r2 / 2);
  const double tmp = (r - e2 * ro) * (r - e2 * ro);
  const double V = std::sqrt(tmp + (1 - e2) * z2);
  const double zo = (b2 * z) / (a * V);
  const double lat = std::atan((z + ep * ep * zo) / r);
  const double t = std::atan(y / x);
  const double lon = (x >= 0) ? t : (((x < 0) & (y >= 0)) ? pi + t : t - pi);
  const double sinphi = std::sin(lat), cosphi = std::cos(lat);
  const double sinlam = std::sin(lon), coslam = std::cos(lon);
  const double M[9] = {-sinlam, coslam, 0, -sinphi * coslam, -sinphi * sinlam, cosphi,
                       cosphi * coslam, cosphi * sinlam, sinphi};  // CoordiTran.cpp:175
  std::memcpy(R, M, sizeof(M));
}
}  // namespace

int vs_poses_from_ins(vs_ctx* ctx, const vs_ins_pva* recs, int64_t n, const double origin_xyz[3],
                      const int64_t* arrival_us, int64_t* out_t_us, double* out_trv) {
  static_assert(sizeof(vs_ins_pva) == sizeof(InsPva) && sizeof(InsPva) == 104, "INSPVA layout");
  if (!ctx || !recs || !origin_xyz || !arrival_us || !out_t_us || !out_trv || n < 1 || n > (1ll << 28))
    return fail(ctx, VS_ERR_INVALID_ARG, "vs_poses_from_ins: bad arguments");
  cudaSetDevice(ctx->device);
  uint8_t* d = nullptr;
  const size_t o_arr = (size_t)n * sizeof(InsPva);
  const size_t o_t = o_arr + (size_t)n * 8;
  const size_t o_trv = o_t + (size_t)n * 8;
  VS_CUDA(cudaMalloc(&d, o_trv + (size_t)n * 72));
  InsParams ip;
  ip.recs = reinterpret_cast<const InsPva*>(d);
  ip.n = (int)n;
  for (int k = 0; k < 3; ++k) ip.org[k] = origin_xyz[k];
  origin_rotation(origin_xyz, ip.R);
  ip.arrival_us = reinterpret_cast<const long long*>(d + o_arr);
  ip.t_out = reinterpret_cast<long long*>(d + o_t);
  ip.trv_out = reinterpret_cast<double*>(d + o_trv);
  cudaStream_t st = ctx->slots[0].stream;
  cudaError_t e = cudaMemcpyAsync(d, recs, o_arr, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_arr, arrival_us, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    k_ins_pose<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ip);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_t_us, d + o_t, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_trv, d + o_trv, (size_t)n * 72, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, VS_ERR_CUDA, std::string("vs_poses_from_ins: ") + cudaGetErrorString(e));
  return VS_OK;
}

#ifdef VS_PROFILE_FUSED
__attribute__((visibility("default"))) int vs_debug_fused_prof(unsigned long long* out16, int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out16, g_fused_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_fused_prof, z, sizeof(z));
  }
  return 0;
}
#endif

void* vs_stream(vs_ctx* ctx) { return ctx ? (void*)ctx->slots[0].stream : nullptr; }
void* vs_slot_stream(vs_ctx* ctx, int slot) {
  return (ctx && slot >= 0 && slot < ctx->n_slots) ? (void*)ctx->slots[slot].stream : nullptr;
}

}  // extern "C"
