// vs_kernels.cuh -- the sm_100a kernels of the ingest hot path.
//
//   k_pcap_times  pcap record headers -> packet times           (vtkPacketFileReader.h:166-197)
//   k_segment     azimuth-wrap segmentation, firingSkip chain   (HDLParser.cxx:1013-1054)
//   k_pose        per-packet pose bracket + lerp + Ry.Rx.Rz     (TransformManager.cxx:149-177,
//                                                                 type_defs.h:134-146)
//   k_decode      decode + calibrate + transform + compaction   (HDLParser.cxx:587-752, 900-977)
//   k_frames      frame table gather (meta packet time / skips)  (HDLParser.cxx:993-1001)
//
// All FP64 arithmetic that feeds an output uses __dmul_rn/__dadd_rn/__dsub_rn so that nvcc
// cannot contract mul+add into FMA: the reference is x86-64 SSE2 code and the doubles must
// match before the final float cast.
#pragma once

#include "vs_device.cuh"

namespace vsd {

// =========================================================================================
// k_pcap_times: t = (ts_sec + 8 h) * 1e6 + ts_usec from the 16-byte pcap record header that
// sits 58 bytes before each payload (timevalToPtime adds 8 hours, type_defs.cxx:69-72).
// =========================================================================================
__global__ void k_pcap_times(const uint8_t* __restrict__ pkts, long long stride, int n,
                             long long* __restrict__ t_us) {
  const int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= n) return;
  const uint8_t* h = pkts + (long long)P * stride - 58;
  const uint32_t sec = h[0] | (h[1] << 8) | (h[2] << 16) | ((uint32_t)h[3] << 24);
  const uint32_t usec = h[4] | (h[5] << 8) | (h[6] << 16) | ((uint32_t)h[7] << 24);
  t_us[P] = ((long long)sec + 8 * 3600) * 1000000ll + (long long)usec;
}

// =========================================================================================
// k_segment
// =========================================================================================
struct SegParams {
  const uint8_t* pkts;
  long long stride;
  const long long* pkt_time;
  int n;     // packets including the halo
  int halo;  // index of the first decoded packet
  int mode;  // 0 streaming, 1 offline
  int carry_last_az;
  int carry_skip;
  int carry_meta_inited;
  PktSeg* pkt_seg;
  unsigned long long* st_map;  // look-back state, one word per tile
  unsigned long long* st_wrap;
  int* tile_counter;
  BatchHeader* hdr;
};

constexpr int kSegThreads = 512;  // one packet per thread, one tile per CTA

__global__ void __launch_bounds__(kSegThreads) k_segment(const SegParams p) {
  __shared__ unsigned long long s_incl[kSegThreads];
  __shared__ unsigned long long s_warp[kSegThreads / 32];
  __shared__ int s_tile;
  __shared__ int s_skip_in;
  __shared__ unsigned long long s_wrap_prefix;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  const int P = tile * kSegThreads + tid;
  const bool live = P < p.n;

  // ---- block headers: 12 x (id, azimuth) ------------------------------------------------
  int az11 = 0;
  unsigned wm = 0;  // bit j (1..11): az[j] < az[j-1]
  unsigned em = 0;  // bit s (0..11): az[s] < lastAzimuth entering the packet
  unsigned um = 0;  // bit j: block id != 0xeeff
  int azdiff = 0;
  if (live) {
    const uint8_t* pk = p.pkts + (long long)P * p.stride;
    int az[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const unsigned short id = __ldg(reinterpret_cast<const unsigned short*>(pk + 100 * j));
      az[j] = __ldg(reinterpret_cast<const unsigned short*>(pk + 100 * j + 2));
      if (id != 0xeeff) um |= 1u << j;
    }
    const int prev11 =
        (P > 0) ? (int)__ldg(reinterpret_cast<const unsigned short*>(pk - p.stride + 1102))
                : p.carry_last_az;
#pragma unroll
    for (int j = 1; j < 12; ++j)
      if (az[j] < az[j - 1]) wm |= 1u << j;
#pragma unroll
    for (int j = 0; j < 12; ++j)
      if (az[j] < prev11) em |= 1u << j;
    az11 = az[11];
    // azimuthDiff: element of rank 6 among the 11 modular deltas (nth_element, :1016-1026)
    int d[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) d[i] = (36000 + az[i + 1] - az[i]) % 36000;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
      int rank = 0;
#pragma unroll
      for (int k = 0; k < 11; ++k) rank += (d[k] < d[i]) || (d[k] == d[i] && k < i);
      if (rank == 6) azdiff = d[i];
    }
  }

  // ---- skip map of this packet (streaming only; offline never skips) --------------------
  unsigned long long m = kMapIdentity;
  if (live) {
    m = 0;
    if (p.mode == 0) {
#pragma unroll
      for (int s = 0; s < 12; ++s) {
        const unsigned hi = wm & ~((2u << s) - 1u);
        const int out = hi ? (31 - __clz(hi)) : (((em >> s) & 1u) ? s : 0);
        m |= (unsigned long long)out << (4 * s);
      }
    }
    if (P < p.halo && map_is_const(m)) atomicMin(&p.hdr->first_const_pkt, P);
  }

  // ---- inclusive scan of maps over the tile ------------------------------------------------
  unsigned long long inc = m;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = map_compose(prev, inc);
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < kSegThreads / 32) ? s_warp[lane] : kMapIdentity;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const unsigned long long prev = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = map_compose(prev, w);
    }
    if (lane < kSegThreads / 32) s_warp[lane] = w;
  }
  __syncthreads();
  if (warp > 0) inc = map_compose(s_warp[warp - 1], inc);
  s_incl[tid] = inc;
  __syncthreads();

  // ---- look-back #1: firingSkip entering the tile ------------------------------------------
  // Serial walk by one thread; it stops at the first constant composed map, which for real
  // sensor data is the previous tile's aggregate.
  if (tid == 0) {
    const unsigned long long agg = s_incl[kSegThreads - 1];
    int skip_in;
    if (tile == 0) {
      skip_in = p.carry_skip;
    } else {
      st_release_u64(&p.st_map[tile], kFlagAgg | agg);
      unsigned long long acc = kMapIdentity;  // maps of tiles (idx, tile) composed
      int idx = tile - 1;
      while (true) {
        unsigned long long v;
        do {
          v = ld_acquire_u64(&p.st_map[idx]);
        } while ((v >> 62) == 0);
        if ((v >> 62) == 2) {
          skip_in = map_apply(acc, (int)(v & 15ull));
          break;
        }
        acc = map_compose(v & kPayloadMask, acc);
        if (map_is_const(acc)) {
          skip_in = (int)(acc & 15ull);
          break;
        }
        if (--idx < 0) {
          skip_in = map_apply(acc, p.carry_skip);
          break;
        }
      }
    }
    st_release_u64(&p.st_map[tile], kFlagPrefix | (unsigned long long)map_apply(agg, skip_in));
    s_skip_in = skip_in;
  }
  __syncthreads();
  const int s = (p.mode == 0)
                    ? map_apply(tid == 0 ? kMapIdentity : s_incl[tid - 1], s_skip_in)
                    : 0;

  // ---- wraps over the iterated blocks -------------------------------------------------------
  unsigned wrapmask = 0;
  if (live) wrapmask = (wm & ~((2u << s) - 1u)) | (((em >> s) & 1u) << s);
  const int nw = __popc(wrapmask);
  // origin marker: streaming -> the packet after the wrap re-initialises the frame meta
  // (F4b); offline -> the wrap packet itself.  marker - 1 == origin packet index.
  const unsigned marker = nw ? (unsigned)(P + (p.mode == 0 ? 2 : 1)) : 0u;
  const unsigned long long v2 = ((unsigned long long)nw << 32) | marker;

  unsigned long long inc2 = v2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 = WrapTraits::combine(prev, inc2);
  }
  __syncthreads();  // s_warp reuse
  if (lane == 31) s_warp[warp] = inc2;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < kSegThreads / 32) ? s_warp[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const unsigned long long prev = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = WrapTraits::combine(prev, w);
    }
    if (lane < kSegThreads / 32) s_warp[lane] = w;
  }
  __syncthreads();
  if (warp > 0) inc2 = WrapTraits::combine(s_warp[warp - 1], inc2);
  // look-back #2 (warp 0): wraps and origin marker before this tile
  if (warp == 0) {
    const unsigned long long agg = s_warp[kSegThreads / 32 - 1];
    unsigned long long ex = lookback_exclusive<WrapTraits>(p.st_wrap, tile, agg);
    if (lane == 0) s_wrap_prefix = ex;
  }
  __syncthreads();
  unsigned long long excl = __shfl_up_sync(0xffffffffu, inc2, 1);
  if (lane == 0) excl = (warp > 0) ? s_warp[warp - 1] : 0ull;
  // (s_warp[warp-1] is the inclusive value of the previous warp's last thread)
  excl = WrapTraits::combine(s_wrap_prefix, excl);
  // seed: with no frame meta carried in, packet 0 is the origin until the first wrap
  if (!p.carry_meta_inited) excl = WrapTraits::combine(excl, 1ull);

  if (live) {
    const int frame_base = (int)(excl >> 32);
    const int origin = (int)(unsigned)excl - 1;
    PktSeg r;
    r.x = s | (int)(wrapmask << 4) | (int)(um << 16);
    r.y = frame_base;
    r.z = origin;
    r.w = azdiff;
    p.pkt_seg[P] = r;

    const unsigned ium = um & ~((1u << s) - 1u);
    if (ium) {
      const long long fu = (long long)P * 12 + (__ffs(ium) - 1);
      // monotone tile order: after the first tile almost every packet fails this test
      if (fu < *reinterpret_cast<volatile long long*>(&p.hdr->first_upper_block))
        atomicMin(reinterpret_cast<unsigned long long*>(&p.hdr->first_upper_block),
                  (unsigned long long)fu);
    }
    if (P == p.halo) {
      p.hdr->origin_at_halo = origin;
      p.hdr->frame_at_halo = frame_base;
    }
    if (P == p.n - 1) {
      p.hdr->total_wraps = frame_base + nw;
      p.hdr->last_azimuth = az11;
      p.hdr->firing_skip_out = (p.mode == 0) ? map_apply(m, s) : 0;
      p.hdr->last_has_wrap = nw > 0;
      int lo;
      if (p.mode == 0)
        lo = nw ? -2 : origin;  // -2: frame meta not initialised yet
      else
        lo = nw ? P : origin;
      p.hdr->last_origin_packet = lo;
      p.hdr->last_origin_time = (lo >= 0) ? p.pkt_time[lo] : 0;
    }
  }
}

// =========================================================================================
// k_pose: one thread per packet.  Output 12 doubles per packet, row-major [L | t] with
// t = T(packet) - T(origin packet).  Launched only when the snapshot holds >= 2 poses.
// =========================================================================================
struct PoseParams {
  const long long* pkt_time;
  const PktSeg* pkt_seg;
  int n;
  int mode;
  const long long* pose_t;
  const double* pose_trv;  // n_poses x 9
  int n_poses;
  int carry_meta_inited;
  double carry_origin_T[3];
  double* pose_mat;  // n x 12
  BatchHeader* hdr;
};

__device__ __forceinline__ double to_radians(double x) {
  return __ddiv_rn(__dmul_rn(x, 3.14159265358979323846), 180.0);
}

// TimeLine::getBoundaryData net semantics (TimeLine.h:384-468): i = clamp(lower_bound, 1, N-1),
// bracket (i-1, i); then TransformManager.cxx:168-175 fore + (back - fore) * ratio.
__device__ __forceinline__ void interp_pose(const long long* __restrict__ pt,
                                            const double* __restrict__ trv, int np, long long t,
                                            double T[3], double R[3], bool want_R) {
  int lo = 0, hi = np;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&pt[mid]) < t)
      lo = mid + 1;
    else
      hi = mid;
  }
  int i = lo < 1 ? 1 : lo;
  if (i > np - 1) i = np - 1;
  const long long tf = __ldg(&pt[i - 1]), tb = __ldg(&pt[i]);
  const double ratio = __ddiv_rn((double)(t - tf), (double)(tb - tf));
  const double* f = trv + (long long)(i - 1) * 9;
  const double* b = trv + (long long)i * 9;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double fk = __ldg(&f[k]);
    T[k] = __dadd_rn(fk, __dmul_rn(__dsub_rn(__ldg(&b[k]), fk), ratio));
  }
  if (want_R) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double fk = __ldg(&f[3 + k]);
      R[k] = __dadd_rn(fk, __dmul_rn(__dsub_rn(__ldg(&b[3 + k]), fk), ratio));
    }
  }
}

// Eigen AngleAxis::toRotationMatrix restated for a unit axis a (a[i] in {0,1}), then
// L <- L * Rm (Transform::rotate post-multiplies).  Same operation order as the oracle.
__device__ __forceinline__ void rotate_by(double L[3][3], double angle, int axis) {
  double s, c;
  sincos(angle, &s, &c);
  const double ax[3] = {axis == 0 ? 1.0 : 0.0, axis == 1 ? 1.0 : 0.0, axis == 2 ? 1.0 : 0.0};
  double sa[3], ca[3];
  const double omc = __dsub_rn(1.0, c);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sa[k] = __dmul_rn(s, ax[k]);
    ca[k] = __dmul_rn(omc, ax[k]);
  }
  double Rm[3][3];
  double tmp = __dmul_rn(ca[0], ax[1]);
  Rm[0][1] = __dsub_rn(tmp, sa[2]);
  Rm[1][0] = __dadd_rn(tmp, sa[2]);
  tmp = __dmul_rn(ca[0], ax[2]);
  Rm[0][2] = __dadd_rn(tmp, sa[1]);
  Rm[2][0] = __dsub_rn(tmp, sa[1]);
  tmp = __dmul_rn(ca[1], ax[2]);
  Rm[1][2] = __dsub_rn(tmp, sa[0]);
  Rm[2][1] = __dadd_rn(tmp, sa[0]);
  Rm[0][0] = __dadd_rn(__dmul_rn(ca[0], ax[0]), c);
  Rm[1][1] = __dadd_rn(__dmul_rn(ca[1], ax[1]), c);
  Rm[2][2] = __dadd_rn(__dmul_rn(ca[2], ax[2]), c);
  double out[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[i][j] = __dadd_rn(__dadd_rn(__dmul_rn(L[i][0], Rm[0][j]), __dmul_rn(L[i][1], Rm[1][j])),
                            __dmul_rn(L[i][2], Rm[2][j]));
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) L[i][j] = out[i][j];
}

__global__ void __launch_bounds__(256) k_pose(const PoseParams p) {
  const int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= p.n) return;
  const long long t = __ldg(&p.pkt_time[P]);
  const PktSeg seg = p.pkt_seg[P];
  double T[3], R[3];
  interp_pose(p.pose_t, p.pose_trv, p.n_poses, t, T, R, true);
  double To[3];
  const int origin = seg.z;
  if (origin < 0) {
    To[0] = p.carry_origin_T[0];
    To[1] = p.carry_origin_T[1];
    To[2] = p.carry_origin_T[2];
  } else if (origin == P) {
    To[0] = T[0];
    To[1] = T[1];
    To[2] = T[2];
  } else {
    double dummy[3];
    interp_pose(p.pose_t, p.pose_trv, p.n_poses, __ldg(&p.pkt_time[origin]), To, dummy, false);
  }
  double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  rotate_by(L, to_radians(R[0]), 1);  // type_defs.h:136 UnitY
  rotate_by(L, to_radians(R[1]), 0);  // :137 UnitX
  rotate_by(L, to_radians(R[2]), 2);  // :138 UnitZ
  double* o = p.pose_mat + (long long)P * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[4 * r + 0] = L[r][0];
    o[4 * r + 1] = L[r][1];
    o[4 * r + 2] = L[r][2];
    o[4 * r + 3] = __dsub_rn(T[r], To[r]);  // reprojectToFrameBeginning, HDLParser.cxx:1057
  }
  if (P == p.n - 1) {
    // frame origin inherited by the next batch
    const int nw = __popc((seg.x >> 4) & 0xfff);
    const bool self = (p.mode == 1) && nw;  // offline: the wrap packet is its frame's origin
#pragma unroll
    for (int k = 0; k < 3; ++k) p.hdr->carry_origin_T[k] = self ? T[k] : To[k];
  }
}

// =========================================================================================
// k_decode: persistent CTAs, tiles of kTilePkts packets staged into shared memory with TMA
// bulk copies (double-buffered), one warp per 100-byte firing block (lane == return slot),
// stream-order compaction through a decoupled look-back on the emitted-point count.
// =========================================================================================
struct DecParams {
  const uint8_t* pkts;  // first packet of the submitted array (halo included)
  long long stride;
  long long total_bytes;  // n * stride: bytes that may be read from pkts
  const long long* pkt_time;
  long long t_base;
  const PktSeg* pkt_seg;
  const double* pose_mat;
  const double* lut_sin;
  const double* lut_cos;
  const DevConfig* cfg;
  int n;  // packets including the halo
  int halo;
  int mode;
  int pose_valid;
  int n_tiles;
  int stage_bytes;  // bytes per shared-memory stage (multiple of 16)
  float* x;
  float* y;
  float* z;
  uint8_t* intensity;
  uint8_t* laser;
  uint16_t* azimuth;
  uint16_t* distance;
  uint32_t* t_us;
  unsigned long long* st_cnt;
  int* tile_counter;
  long long* frame_first_point;
  int* frame_start_block;
  unsigned* frame_laser_counts;  // frame_cap x 64
  int frame_cap;
  BatchHeader* hdr;
};

constexpr int kTilePkts = 32;
constexpr int kTileBlocks = kTilePkts * kBlocks;  // 384
constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;

struct DecShared {
  DevConfig cfg;
  uint64_t full[2];
  unsigned mask[kTileBlocks];  // ballot of emitted return slots per firing block
  unsigned offs[kTileBlocks];  // exclusive prefix of popc(mask) inside the tile
  PktSeg seg[kTilePkts];
  unsigned tpk[kTilePkts];     // packet time - t_base
  unsigned hist[2][kMaxLasers];
  unsigned long long tile_base;
  int tile_id[2];
};

__device__ __forceinline__ unsigned ld_smem_u16(const uint8_t* p) {
  // packets are only guaranteed 2-byte aligned (1206 = 2 * 603)
  return *reinterpret_cast<const unsigned short*>(p);
}

// Per-laser calibration row held in registers (reloaded only when the bank changes).
struct CalRow {
  double cC, sC, dc, cV, sV, vo, ho;
};
__device__ __forceinline__ void load_cal(const DevConfig& c, int row, CalRow& r) {
  r.cC = c.cal[0][row];
  r.sC = c.cal[1][row];
  r.dc = c.cal[2][row];
  r.cV = c.cal[3][row];
  r.sV = c.cal[4][row];
  r.vo = c.cal[5][row];
  r.ho = c.cal[6][row];
}

// Sensor-frame position of one return (HDLParser.cxx:597-623).  `az` is already adjusted and
// reduced mod 36000.  All arithmetic in separate IEEE mul/add, reference operation order.
__device__ __forceinline__ void sensor_point(const CalRow& c, double sA, double cA, unsigned dist,
                                             double& px, double& py, double& pz) {
  // sin/cos(rad(az/100) - rad(rotCorrection)); with rotCorrection == 0 (cC=1, sC=0) this is
  // exactly the reference's LUT branch (:602-606)
  const double sinAz = __dsub_rn(__dmul_rn(sA, c.cC), __dmul_rn(cA, c.sC));
  const double cosAz = __dadd_rn(__dmul_rn(cA, c.cC), __dmul_rn(sA, c.sC));
  const double dM = __dadd_rn(__dmul_rn((double)dist, 0.002), c.dc);  // :614
  const double xy = __dmul_rn(dM, c.cV);                              // :615
  px = __dsub_rn(__dmul_rn(xy, sinAz), __dmul_rn(c.ho, cosAz));       // :620
  py = __dadd_rn(__dmul_rn(xy, cosAz), __dmul_rn(c.ho, sinAz));       // :621
  pz = __dadd_rn(__dmul_rn(dM, c.sV), c.vo);                          // :622
}

template <int ADJ>
__device__ __forceinline__ unsigned adjusted_azimuth(const DevConfig& c, unsigned rot, int azdiff,
                                                     int j, int lane) {
  unsigned az = rot;
  if (ADJ != 0) {
    // HDLParser.cxx:961: std::round (half away from zero) of azimuthDiff * ratio
    const int adj = (int)round(__dmul_rn((double)azdiff, c.az_ratio[j][lane]));
    az = (unsigned)(unsigned short)(rot + adj);  // passed as unsigned short, :968
  }
  return az % 36000u;  // :597
}

// Byte span of a tile in the input array: [a0, a1) are the bytes the tile's packets occupy,
// [s0, s1) the 16-byte-granular span the TMA bulk copy moves.  When rounding a1 up would
// read past the bytes the caller owns, the copy stops at the last full granule and the
// (< 16) tail bytes are fetched with plain loads.
struct TileSpan {
  long long a0, a1, s0, s1;
  int npk;
};
__device__ __forceinline__ TileSpan tile_span(long long in_base, long long stride,
                                              long long total_bytes, int n, int halo, int tile) {
  TileSpan t;
  const long long first = (long long)halo + (long long)tile * kTilePkts;
  t.npk = n - (int)first;
  if (t.npk > kTilePkts) t.npk = kTilePkts;
  t.a0 = in_base + first * stride;
  t.a1 = t.a0 + (long long)(t.npk - 1) * stride + kPacketBytes;
  t.s0 = t.a0 & ~15ll;
  t.s1 = (t.a1 + 15) & ~15ll;
  if (t.s1 > in_base + total_bytes) t.s1 = t.a1 & ~15ll;
  if (t.s1 < t.s0) t.s1 = t.s0;
  return t;
}

// Work split inside a tile: warp w owns the packets {w/2 + 4k} and, inside them, the firing
// blocks of parity w&1.  On HDL-64 data block parity == laser bank (0xeeff / 0xddff), so a
// warp keeps one calibration bank in registers; the pose matrix is loaded once per packet.
template <int ADJ, bool CROP>
__global__ void __launch_bounds__(kDecThreads, 2) k_decode(const DecParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  DecShared& sh = *reinterpret_cast<DecShared*>(smem_raw);
  uint8_t* stage0 = smem_raw + ((sizeof(DecShared) + 127) & ~127);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp & 1, pk0 = warp >> 1;
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- one-time CTA setup -------------------------------------------------------------------
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.cfg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.cfg);
    for (int i = tid; i < (int)(sizeof(DevConfig) / 4); i += kDecThreads) dst[i] = __ldg(&src[i]);
  }
  if (tid == 0) {
    mbar_init(&sh.full[0], 1);
    mbar_init(&sh.full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long in_base = reinterpret_cast<long long>(p.pkts);

  // issue the staged load of tile `t` into buffer `b` (thread 0 only)
  auto issue = [&](int t, int b) {
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, p.halo, t);
    const uint32_t bytes = (uint32_t)(sp.s1 - sp.s0);
    fence_proxy_async();
    if (bytes) {
      mbar_expect_tx(&sh.full[b], bytes);
      bulk_g2s(stage0 + (size_t)b * p.stage_bytes, reinterpret_cast<const void*>(sp.s0), bytes,
               &sh.full[b]);
    } else {
      mbar_arrive(&sh.full[b]);
    }
  };

  if (tid == 0) {
    const int t = atomicAdd(p.tile_counter, 1);
    sh.tile_id[0] = t;
    if (t < p.n_tiles) issue(t, 0);
  }
  __syncthreads();

  uint32_t phase[2] = {0u, 0u};
  int cur = 0;
  const unsigned long long lmask = sh.cfg.laser_mask;
  const int pskip = sh.cfg.points_skip;
  const int n_enabled = sh.cfg.n_enabled;
  const bool pose_valid = p.pose_valid != 0;

  CalRow cal;
  int cal_bank = -1;

  while (true) {
    const int tile = sh.tile_id[cur];
    if (tile >= p.n_tiles) break;
    // prefetch the next tile into the other buffer
    if (tid == 0) {
      const int tn = atomicAdd(p.tile_counter, 1);
      sh.tile_id[cur ^ 1] = tn;
      if (tn < p.n_tiles) issue(tn, cur ^ 1);
    }
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, p.halo, tile);
    const long long first = (long long)p.halo + (long long)tile * kTilePkts;
    const int npk = sp.npk;
    uint8_t* stage = stage0 + (size_t)cur * p.stage_bytes;

    // per-packet records of this tile (overlaps the TMA wait)
    if (tid < kTilePkts) {
      PktSeg sg = make_int4(0, 0, 0, 0);
      unsigned tp = 0;
      if (tid < npk) {
        sg = p.pkt_seg[first + tid];
        tp = (unsigned)(__ldg(&p.pkt_time[first + tid]) - p.t_base);
      }
      sh.seg[tid] = sg;
      sh.tpk[tid] = tp;
    } else if (tid < kTilePkts + 2 * kMaxLasers) {
      (&sh.hist[0][0])[tid - kTilePkts] = 0;
    }

    mbar_wait(&sh.full[cur], phase[cur]);
    phase[cur] ^= 1u;
    if (sp.s1 < sp.a1) {
      // tail bytes the bulk copy could not cover (unaligned end of the caller's buffer)
      for (long long a = sp.s1 + tid; a < sp.a1; a += kDecThreads)
        stage[a - sp.s0] = *reinterpret_cast<const uint8_t*>(a);
    }
    __syncthreads();
    const uint8_t* tile_smem = stage + (sp.a0 - sp.s0);

    // ---- pass 1: which return slots of each firing block are emitted ---------------------
#pragma unroll 1
    for (int lp = pk0; lp < kTilePkts; lp += 4) {
      const PktSeg seg = sh.seg[lp];
      const int skip_in = seg.x & 15;
      const uint8_t* pk = tile_smem + (size_t)lp * p.stride;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int j = par + 2 * i;
        unsigned m = 0;
        if (lp < npk && j >= skip_in && (pskip == 0 || (j % (pskip + 1)) == 0)) {
          const uint8_t* blk = pk + 100 * j;
          const int off = (ld_smem_u16(blk) == 0xeeffu) ? 0 : 32;
          const unsigned dist = blk[4 + 3 * lane] | (blk[5 + 3 * lane] << 8);
          int laser_id = lane + off;
          if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
          bool valid = dist != 0 && ((lmask >> laser_id) & 1ull) && laser_id < n_enabled;
          if (CROP) {
            // the crop test is on the sensor-frame position, before the transform (:629-639)
            CalRow c;
            load_cal(sh.cfg, lane + off, c);
            const unsigned az = adjusted_azimuth<ADJ>(sh.cfg, ld_smem_u16(blk + 2), seg.w, j, lane);
            double px, py, pz;
            sensor_point(c, __ldg(&p.lut_sin[az]), __ldg(&p.lut_cos[az]), dist, px, py, pz);
            const bool in_box = px >= sh.cfg.crop[0] && px <= sh.cfg.crop[1] &&
                                py >= sh.cfg.crop[2] && py <= sh.cfg.crop[3] &&
                                pz >= sh.cfg.crop[4] && pz <= sh.cfg.crop[5];
            valid = valid && (in_box == (sh.cfg.crop_inside != 0));  // :634-638
          }
          m = __ballot_sync(0xffffffffu, valid);
        }
        if (lane == 0) sh.mask[lp * kBlocks + j] = m;
      }
    }
    __syncthreads();

    // ---- scan (warp 0: lane == packet) + tile base through the decoupled look-back ----------
    if (warp == 0) {
      unsigned c[kBlocks];
      unsigned tot = 0;
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) {
        c[j] = tot;
        tot += __popc(sh.mask[lane * kBlocks + j]);
      }
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned pk_excl = inc - tot;
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) sh.offs[lane * kBlocks + j] = pk_excl + c[j];
      const unsigned long long agg = __shfl_sync(0xffffffffu, inc, 31);
      const unsigned long long ex = lookback_exclusive<SumTraits>(p.st_cnt, tile, agg);
      if (lane == 0) {
        sh.tile_base = ex;
        if (tile == p.n_tiles - 1) p.hdr->total_points = (long long)(ex + agg);
      }
      // frame-start records: a wrap block opens frame f before it is decoded (:1035-1039)
      const PktSeg seg = sh.seg[lane];
      int wm = (seg.x >> 4) & 0xfff;
      while (wm) {
        const int j = __ffs(wm) - 1;
        wm &= wm - 1;
        const int f = seg.y + __popc(((seg.x >> 4) & 0xfff) & ((2 << j) - 1));
        if (f < p.frame_cap) {
          p.frame_first_point[f] = (long long)(ex + pk_excl + c[j]);
          p.frame_start_block[f] = (int)((first + lane) * 12 + j);
        } else {
          p.hdr->frame_overflow = 1;
        }
      }
    }
    __syncthreads();

    // ---- pass 2: decode, calibrate, transform, store ----------------------------------------
    const unsigned long long tb = sh.tile_base;
    float* const xt = p.x + tb;
    float* const yt = p.y + tb;
    float* const zt = p.z + tb;
    uint32_t* const tt = p.t_us + tb;
    uint16_t* const at = p.azimuth + tb;
    uint16_t* const dt = p.distance + tb;
    uint8_t* const it = p.intensity + tb;
    uint8_t* const lt = p.laser + tb;
    const int tile_f0 = sh.seg[0].y;

    unsigned cnt_lo = 0, cnt_hi = 0;  // per-lane emitted counts for the lower / upper laser bank
    int warp_frame = -1;
    auto flush_counts = [&](int f) {
      if (f < 0) return;
      const int d = f - tile_f0;
      int l_lo = lane, l_hi = lane + 32;
      if (ADJ == 2) {
        l_lo = lane >= 16 ? lane - 16 : lane;
        l_hi = lane + 16;
      }
      if (d == 0 || d == 1) {
        if (cnt_lo) atomicAdd(&sh.hist[d][l_lo], cnt_lo);
        if (cnt_hi) atomicAdd(&sh.hist[d][l_hi], cnt_hi);
      } else if (f < p.frame_cap) {
        if (cnt_lo) atomicAdd(&p.frame_laser_counts[(long long)f * kMaxLasers + l_lo], cnt_lo);
        if (cnt_hi) atomicAdd(&p.frame_laser_counts[(long long)f * kMaxLasers + l_hi], cnt_hi);
      }
      cnt_lo = cnt_hi = 0;
    };

#pragma unroll 1
    for (int lp = pk0; lp < npk; lp += 4) {
      const PktSeg seg = sh.seg[lp];
      const int wrapmask = (seg.x >> 4) & 0xfff;
      const int first_wrap = wrapmask ? (__ffs(wrapmask) - 1) : 12;
      const uint8_t* pk = tile_smem + (size_t)lp * p.stride;
      const unsigned tpk = sh.tpk[lp];
      double M[12];  // [L | t] of this packet, warp-uniform
      if (pose_valid) {
        const double* mp = p.pose_mat + (first + lp) * 12;
#pragma unroll
        for (int q = 0; q < 12; ++q) M[q] = __ldg(&mp[q]);
      }
      if (wrapmask == 0 && seg.y != warp_frame) {
        flush_counts(warp_frame);
        warp_frame = seg.y;
      }
#pragma unroll 2
      for (int i = 0; i < 6; ++i) {
        const int j = par + 2 * i;
        const int b = lp * kBlocks + j;
        const unsigned m = sh.mask[b];
        if (m == 0) continue;
        if (wrapmask) {
          const int f = seg.y + __popc(wrapmask & ((2 << j) - 1));
          if (f != warp_frame) {
            flush_counts(warp_frame);
            warp_frame = f;
          }
        }
        const unsigned o = sh.offs[b] + __popc(m & lt_mask);
        const bool valid = (m >> lane) & 1u;
        const uint8_t* blk = pk + 100 * j;
        const int off = (ld_smem_u16(blk) == 0xeeffu) ? 0 : 32;
        if (off != cal_bank) {
          load_cal(sh.cfg, lane + off, cal);
          cal_bank = off;
        }
        const unsigned rot = ld_smem_u16(blk + 2);
        const unsigned dist = blk[4 + 3 * lane] | (blk[5 + 3 * lane] << 8);
        const unsigned inten = blk[6 + 3 * lane];
        int laser_id = lane + off;
        if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
        const unsigned az = adjusted_azimuth<ADJ>(sh.cfg, rot, seg.w, j, lane);
        double px, py, pz;
        sensor_point(cal, __ldg(&p.lut_sin[az]), __ldg(&p.lut_cos[az]), dist, px, py, pz);
        if (pose_valid) {
          // offline: blocks at/after the packet's first wrap start a frame whose origin is
          // this very packet -> zero translation.  type_defs.h:160-166: row sums left to
          // right, translation last.
          const bool t_zero = (p.mode == 1) && (j >= first_wrap);
          const double tx = t_zero ? 0.0 : M[3], ty = t_zero ? 0.0 : M[7], tz = t_zero ? 0.0 : M[11];
          const double qx = __dadd_rn(
              __dadd_rn(__dadd_rn(__dmul_rn(M[0], px), __dmul_rn(M[1], py)), __dmul_rn(M[2], pz)),
              tx);
          const double qy = __dadd_rn(
              __dadd_rn(__dadd_rn(__dmul_rn(M[4], px), __dmul_rn(M[5], py)), __dmul_rn(M[6], pz)),
              ty);
          const double qz = __dadd_rn(
              __dadd_rn(__dadd_rn(__dmul_rn(M[8], px), __dmul_rn(M[9], py)), __dmul_rn(M[10], pz)),
              tz);
          px = qx;
          py = qy;
          pz = qz;
        }
        if (valid) {
          xt[o] = (float)px;
          yt[o] = (float)py;
          zt[o] = (float)pz;
          tt[o] = tpk + (ADJ != 0 ? (uint32_t)sh.cfg.tadj[j][lane] : 0u);
          at[o] = (uint16_t)az;
          dt[o] = (uint16_t)dist;
          it[o] = (uint8_t)inten;
          lt[o] = (uint8_t)laser_id;
          if (off)
            ++cnt_hi;
          else
            ++cnt_lo;
        }
      }
    }
    flush_counts(warp_frame);
    __syncthreads();
    if (tid < 2 * kMaxLasers) {
      const unsigned c = (&sh.hist[0][0])[tid];
      const int f = tile_f0 + (tid >> 6);
      if (c && f < p.frame_cap)
        atomicAdd(&p.frame_laser_counts[(long long)f * kMaxLasers + (tid & 63)], c);
    }
    __syncthreads();  // stage `cur`, masks and hist are free again
    cur ^= 1;
  }
}

// =========================================================================================
// k_frames: per frame started inside the batch, find the packet that initialises its meta
// (streaming: the packet after the one holding the frame's wrap, provided that wrap is the
// packet's last one, HDLParser.cxx:993-1001; offline: the wrap packet itself).
// =========================================================================================
struct FrameParams {
  const PktSeg* pkt_seg;
  const long long* pkt_time;
  const int* frame_start_block;
  int n;
  int n_frames;  // total_wraps + 1
  int mode;
  int* frame_meta_packet;   // out: -2 none, -3 pending (next batch)
  long long* frame_meta_time;
  int* frame_skips;
};

__global__ void k_frames(const FrameParams p) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < 1 || f >= p.n_frames) return;  // frame 0 comes from the carry / first packet
  const int sb = p.frame_start_block[f];
  if (sb < 0) {  // wrap inside the halo: not decoded
    p.frame_meta_packet[f] = -2;
    p.frame_meta_time[f] = 0;
    p.frame_skips[f] = -1;
    return;
  }
  const int P = sb / 12, j = sb % 12;
  int mp = -2, sk = -1;
  long long mt = 0;
  if (p.mode == 1) {
    mp = P;
    sk = j;
    mt = p.pkt_time[P];
  } else {
    const int wrapmask = (p.pkt_seg[P].x >> 4) & 0xfff;
    const bool last_wrap = (wrapmask >> (j + 1)) == 0;
    if (last_wrap) {
      if (P + 1 < p.n) {
        mp = P + 1;
        sk = p.pkt_seg[P + 1].x & 15;
        mt = p.pkt_time[P + 1];
      } else {
        mp = -3;
      }
    }
  }
  p.frame_meta_packet[f] = mp;
  p.frame_meta_time[f] = mt;
  p.frame_skips[f] = sk;
}

}  // namespace vsd
