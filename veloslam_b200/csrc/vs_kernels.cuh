// vs_kernels.cuh -- the sm_100a kernels of the ingest hot path.
//
//   k_pcap_times  pcap record headers -> packet times           (vtkPacketFileReader.h:166-197)
//   k_scan        azimuth-wrap segmentation, firingSkip chain,   (HDLParser.cxx:1013-1054,
//                 emission masks per firing block                 964, 629-639)
//   k_pose        frame ids / origins / point offsets (scans),   (TransformManager.cxx:149-177,
//                 per-packet pose bracket + lerp + Ry.Rx.Rz       type_defs.h:134-146)
//   k_decode      decode + calibrate + transform + SoA stores    (HDLParser.cxx:587-752, 900-977)
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
// Tiles: both streaming kernels (k_scan, k_decode) walk the packet array in tiles of 32
// packets (384 firing blocks, 12 288 return slots, 38.6 KB) staged into shared memory with
// one TMA bulk copy per tile, double buffered.
// =========================================================================================
constexpr int kTilePkts = 32;
constexpr int kTileBlocks = kTilePkts * kBlocks;  // 384
constexpr int kLead = 192;  // bytes in front of a tile that hold the previous packet's last
                            // azimuth (stride - 1102 <= 178 for stride <= 1280), 16-B multiple

// Byte span of a tile in the input array: [a0, a1) are the bytes the tile's packets occupy,
// [s0, s1) the 16-byte-granular span the bulk copy moves (optionally `lead` bytes earlier).
// When rounding a1 up would read past the bytes the caller owns, the copy stops at the last
// full granule and the (< 16) tail bytes are fetched with plain loads.
struct TileSpan {
  long long a0, a1, s0, s1;
  int npk;
};
__device__ __forceinline__ TileSpan tile_span(long long in_base, long long stride,
                                              long long total_bytes, int n, long long first,
                                              int lead) {
  TileSpan t;
  t.npk = n - (int)first;
  if (t.npk > kTilePkts) t.npk = kTilePkts;
  t.a0 = in_base + first * stride;
  t.a1 = t.a0 + (long long)(t.npk - 1) * stride + kPacketBytes;
  long long b = t.a0 - lead;
  if (b < in_base) b = in_base;
  t.s0 = b & ~15ll;
  t.s1 = (t.a1 + 15) & ~15ll;
  if (t.s1 > in_base + total_bytes) t.s1 = t.a1 & ~15ll;
  if (t.s1 < t.s0) t.s1 = t.s0;
  return t;
}

__device__ __forceinline__ unsigned ld_smem_u16(const uint8_t* p) {
  // packets are only guaranteed 2-byte aligned (1206 = 2 * 603)
  return *reinterpret_cast<const unsigned short*>(p);
}

// Loads through explicit 32-bit shared-window addresses (no generic->shared conversion in the
// block loop).
__device__ __forceinline__ unsigned lds_u8(uint32_t a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(uint32_t a) {
  unsigned v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u32(uint32_t a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(a));
  return v;
}
// Streaming stores: explicit global space (the column pointers are kept opaque to pin them in
// registers, which would otherwise degrade the stores to generic ST).
__device__ __forceinline__ void stg_f32(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_u16(uint16_t* p, unsigned v) {
  asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void stg_u8(uint8_t* p, unsigned v) {
  asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

// Per-laser calibration row held in registers.
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
// reduced mod 36000.  Separate IEEE mul/add in the reference's operation order.
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

// =========================================================================================
// k_scan: segmentation + emission masks in one streaming pass over the packets.
//   per packet : firingSkip entering it, wrap mask over the iterated blocks, azimuthDiff,
//                emitted-point count                                  -> PktSeg
//   per block  : 32-bit mask of the return slots the reference emits -> masks[n*12]
// The firingSkip recurrence is a scan of 12-entry maps; across tiles it is resolved with a
// look-back that stops at the first constant composed map.
// =========================================================================================
struct ScanParams {
  const uint8_t* pkts;
  long long stride;
  long long total_bytes;
  const DevConfig* cfg;
  const double* lut_sin;
  const double* lut_cos;
  int n;
  int mode;  // 0 streaming, 1 offline
  int halo;
  int carry_last_az;
  int carry_skip;
  int n_tiles;
  int stage_bytes;
  PktSeg* pkt_seg;
  unsigned* masks;
  unsigned long long* st_map;
  int* tile_counter;
  BatchHeader* hdr;
};

constexpr int kScanThreads = 256;

struct ScanShared {
  uint64_t full;
  unsigned nz[kTileBlocks];  // raw "distance != 0" (or crop-tested) bits per block
  int skip[kTilePkts];
  int tile_id;
};

// 12-entry skip map of one packet from its block azimuths and the azimuth that precedes it.
__device__ __forceinline__ unsigned long long packet_skip_map(const int az[12], int prev11,
                                                              unsigned& wm, unsigned& em) {
  wm = 0;
  em = 0;
#pragma unroll
  for (int j = 1; j < 12; ++j)
    if (az[j] < az[j - 1]) wm |= 1u << j;
#pragma unroll
  for (int j = 0; j < 12; ++j)
    if (az[j] < prev11) em |= 1u << j;
  unsigned long long m = 0;
#pragma unroll
  for (int s = 0; s < 12; ++s) {
    const unsigned hi = wm & ~((2u << s) - 1u);
    const int out = hi ? (31 - __clz(hi)) : (((em >> s) & 1u) ? s : 0);
    m |= (unsigned long long)out << (4 * s);
  }
  return m;
}

// One stage per CTA, 4-5 CTAs per SM: the TMA latency of one CTA is hidden by the others.
// The tile's lead bytes hold the whole previous packet, so the firingSkip entering the tile is
// known locally whenever that packet's skip map is constant (always, on sensor data); only
// otherwise does the tile fall back to the look-back over the published tile maps.
template <int ADJ, bool CROP>
__global__ void __launch_bounds__(kScanThreads) k_scan(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ScanShared& sh = *reinterpret_cast<ScanShared*>(smem_raw);
  uint8_t* stage = smem_raw + ((sizeof(ScanShared) + 127) & ~127);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DevConfig& cfg = *p.cfg;  // few uniform fields; the CROP variant reads the rows too

  if (tid == 0) {
    mbar_init(&sh.full, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long in_base = reinterpret_cast<long long>(p.pkts);
  const int lead = (int)p.stride + kLead;
  const unsigned sel_lo = cfg.sel_lo, sel_hi = cfg.sel_hi;
  const int pskip = cfg.points_skip;

  uint32_t phase = 0u;
  while (true) {
    if (tid == 0) {
      const int t = atomicAdd(p.tile_counter, 1);
      sh.tile_id = t;
      if (t < p.n_tiles) {
        const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, (long long)t * kTilePkts, lead);
        const uint32_t bytes = (uint32_t)(sp.s1 - sp.s0);
        fence_proxy_async();
        if (bytes) {
          mbar_expect_tx(&sh.full, bytes);
          bulk_g2s(stage, reinterpret_cast<const void*>(sp.s0), bytes, &sh.full);
        } else {
          mbar_arrive(&sh.full);
        }
      }
    }
    __syncthreads();
    const int tile = sh.tile_id;
    if (tile >= p.n_tiles) break;
    const long long first = (long long)tile * kTilePkts;
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, first, lead);
    const int npk = sp.npk;
    mbar_wait(&sh.full, phase);
    phase ^= 1u;
    if (sp.s1 < sp.a1) {
      for (long long a = sp.s1 + tid; a < sp.a1; a += kScanThreads)
        stage[a - sp.s0] = *reinterpret_cast<const uint8_t*>(a);
      __syncthreads();
    }
    const uint8_t* tile_smem = stage + (sp.a0 - sp.s0);

    // ---- phase A (warp 0, lane == packet): headers, skip maps, firingSkip per packet --------
    unsigned wm = 0, em = 0, um = 0, wrapmask = 0;
    int azdiff = 0, s_in = 0, az11 = 0;
    unsigned long long m = kMapIdentity;
    if (warp == 0) {
      const bool live = lane < npk;
      int az[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) az[j] = 0;
      if (live) {
        const uint8_t* pk = tile_smem + (size_t)lane * p.stride;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          if (ld_smem_u16(pk + 100 * j) != 0xeeffu) um |= 1u << j;
          az[j] = (int)ld_smem_u16(pk + 100 * j + 2);
        }
        az11 = az[11];
      }
      // lastAzimuth entering the packet: block 11 of the previous packet (always iterated)
      int prev11 = __shfl_up_sync(0xffffffffu, az11, 1);
      if (lane == 0)
        prev11 = (first > 0) ? (int)ld_smem_u16(tile_smem - p.stride + 1102) : p.carry_last_az;
      if (live) {
        m = packet_skip_map(az, prev11, wm, em);
        if (p.mode != 0) m = 0;
        if (ADJ != 0) {
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
        if (first + lane < p.halo && map_is_const(m))
          atomicMin(&p.hdr->first_const_pkt, (int)(first + lane));
      }
      // inclusive scan of the maps; trivial when no packet of the tile can skip
      unsigned long long inc = m;
      const bool all_zero = __all_sync(0xffffffffu, m == 0ull || !live);
      if (!all_zero) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc = map_compose(prev, inc);
        }
      } else if (!live) {
        inc = 0ull;  // dead lanes follow a constant-0 packet
      }
      const unsigned long long agg = __shfl_sync(0xffffffffu, inc, 31);
      int skip_tile = 0;
      if (lane == 0) {
        if (tile == 0) {
          skip_tile = p.carry_skip;
        } else {
          st_release_u64(&p.st_map[tile], kFlagAgg | agg);
          // the previous packet's own map, from the lead bytes
          bool known = false;
          if (p.mode != 0) {
            known = true;
            skip_tile = 0;
          } else {
            int paz[12];
            const uint8_t* pp = tile_smem - p.stride;
#pragma unroll
            for (int j = 0; j < 12; ++j) paz[j] = (int)ld_smem_u16(pp + 100 * j + 2);
            const int pprev = (first > 1) ? (int)ld_smem_u16(pp - p.stride + 1102) : p.carry_last_az;
            unsigned w2, e2;
            const unsigned long long pmap = packet_skip_map(paz, pprev, w2, e2);
            if (map_is_const(pmap)) {
              known = true;
              skip_tile = (int)(pmap & 15ull);
            }
          }
          if (!known) {
            unsigned long long acc = kMapIdentity;  // maps of tiles (idx, tile) composed
            int idx = tile - 1;
            while (true) {
              unsigned long long v;
              do {
                v = ld_acquire_u64(&p.st_map[idx]);
              } while ((v >> 62) == 0);
              if ((v >> 62) == 2) {
                skip_tile = map_apply(acc, (int)(v & 15ull));
                break;
              }
              acc = map_compose(v & kPayloadMask, acc);
              if (map_is_const(acc)) {
                skip_tile = (int)(acc & 15ull);
                break;
              }
              if (--idx < 0) {
                skip_tile = map_apply(acc, p.carry_skip);
                break;
              }
            }
          }
        }
        st_release_u64(&p.st_map[tile], kFlagPrefix | (unsigned long long)map_apply(agg, skip_tile));
      }
      skip_tile = __shfl_sync(0xffffffffu, skip_tile, 0);
      unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) excl = kMapIdentity;
      s_in = (p.mode == 0) ? map_apply(excl, skip_tile) : 0;
      if (live) wrapmask = (wm & ~((2u << s_in) - 1u)) | (((em >> s_in) & 1u) << s_in);
      sh.skip[lane] = s_in;
    }

    // ---- phase B1: which return slots of each block could be emitted ------------------------
    if (!CROP) {
      // lane per block: scan the 32 distance fields of a block with 16-bit loads (the 100-byte
      // block stride spreads the lanes of a warp over the banks)
      for (int b = tid; b < kTileBlocks; b += kScanThreads) {
        const int lp = b / kBlocks, j = b - lp * kBlocks;
        unsigned bits = 0;
        if (lp < npk) {
          const uint8_t* blk = tile_smem + (size_t)lp * p.stride + 100 * j;
          unsigned h[50];
#pragma unroll
          for (int k = 2; k < 50; ++k) h[k] = ld_smem_u16(blk + 2 * k);
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int o = 4 + 3 * r;  // byte offset of the distance field
            unsigned d;
            if ((o & 1) == 0)
              d = h[o >> 1];
            else
              d = (h[o >> 1] & 0xff00u) | (h[(o >> 1) + 1] & 0x00ffu);
            bits |= (d != 0u ? 1u : 0u) << r;
          }
        }
        sh.nz[b] = bits;
      }
    } else {
      // crop test needs the sensor-frame position (HDLParser.cxx:629-639): warp per block
      for (int b = warp; b < kTileBlocks; b += kScanThreads / 32) {
        const int lp = b / kBlocks, j = b - lp * kBlocks;
        unsigned bits = 0;
        if (lp < npk) {
          const uint8_t* pk = tile_smem + (size_t)lp * p.stride;
          const uint8_t* blk = pk + 100 * j;
          const int off = (ld_smem_u16(blk) == 0xeeffu) ? 0 : 32;
          const unsigned dist = blk[4 + 3 * lane] | (blk[5 + 3 * lane] << 8);
          int ad = 0;
          if (ADJ != 0) {
            int d[11];
#pragma unroll
            for (int i = 0; i < 11; ++i)
              d[i] = (36000 + (int)ld_smem_u16(pk + 100 * (i + 1) + 2) -
                      (int)ld_smem_u16(pk + 100 * i + 2)) % 36000;
#pragma unroll
            for (int i = 0; i < 11; ++i) {
              int rank = 0;
#pragma unroll
              for (int k = 0; k < 11; ++k) rank += (d[k] < d[i]) || (d[k] == d[i] && k < i);
              if (rank == 6) ad = d[i];
            }
          }
          CalRow c;
          c.cC = __ldg(&cfg.cal[0][lane + off]);
          c.sC = __ldg(&cfg.cal[1][lane + off]);
          c.dc = __ldg(&cfg.cal[2][lane + off]);
          c.cV = __ldg(&cfg.cal[3][lane + off]);
          c.sV = __ldg(&cfg.cal[4][lane + off]);
          c.vo = __ldg(&cfg.cal[5][lane + off]);
          c.ho = __ldg(&cfg.cal[6][lane + off]);
          const unsigned az = adjusted_azimuth<ADJ>(cfg, ld_smem_u16(blk + 2), ad, j, lane);
          double px, py, pz;
          sensor_point(c, __ldg(&p.lut_sin[az]), __ldg(&p.lut_cos[az]), dist, px, py, pz);
          const bool in_box = px >= cfg.crop[0] && px <= cfg.crop[1] && py >= cfg.crop[2] &&
                              py <= cfg.crop[3] && pz >= cfg.crop[4] && pz <= cfg.crop[5];
          bits = __ballot_sync(0xffffffffu, dist != 0 && (in_box == (cfg.crop_inside != 0)));
        }
        if (lane == 0) sh.nz[b] = bits;
      }
    }
    __syncthreads();

    // ---- phase B2: final masks (iterated, gated, selected lasers) ---------------------------
    for (int b = tid; b < kTileBlocks; b += kScanThreads) {
      const int lp = b / kBlocks, j = b - lp * kBlocks;
      if (lp < npk) {
        unsigned mk = 0;
        if (j >= sh.skip[lp] && (pskip == 0 || (j % (pskip + 1)) == 0)) {
          const bool upper = ld_smem_u16(tile_smem + (size_t)lp * p.stride + 100 * j) != 0xeeffu;
          mk = sh.nz[b] & (upper ? sel_hi : sel_lo);
        }
        sh.nz[b] = mk;
        p.masks[(first + lp) * kBlocks + j] = mk;
      }
    }
    __syncthreads();
    if (warp == 0 && lane < npk) {
      unsigned cnt = 0;
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) cnt += __popc(sh.nz[lane * kBlocks + j]);
      const long long P = first + lane;
      PktSeg r;
      r.x = s_in | (int)(wrapmask << 4) | (azdiff << 16);
      r.y = (int)cnt;
      r.z = 0;
      r.w = (int)um;
      p.pkt_seg[P] = r;
      const unsigned ium = um & ~((1u << s_in) - 1u);
      if (ium) {
        const long long fu = P * 12 + (__ffs(ium) - 1);
        // tiles run in order: after the first one almost every packet fails this test
        if (fu < *reinterpret_cast<volatile long long*>(&p.hdr->first_upper_block))
          atomicMin(reinterpret_cast<unsigned long long*>(&p.hdr->first_upper_block),
                    (unsigned long long)fu);
      }
      if (P == p.n - 1) {
        p.hdr->last_azimuth = az11;
        p.hdr->firing_skip_out = (p.mode == 0) ? map_apply(m, s_in) : 0;
      }
    }
    __syncthreads();  // stage, nz, skip and tile_id are free again
  }
}

// =========================================================================================
// k_pose: one thread per packet.
//   (1) scans over packets (decoupled look-back): wraps -> frame id, last wrap -> origin
//       packet, emitted counts -> point offset; frame-start records of the frame table;
//   (2) pose bracket + lerp + Ry.Rx.Rz and T(packet) - T(origin) when the snapshot has >= 2
//       poses: 12 doubles per packet, row-major [L | t].
// =========================================================================================
struct PoseParams {
  const long long* pkt_time;
  PktSeg* pkt_seg;           // in: x, y = count; out: y = frame id, z = time - t_base
  const unsigned* masks;
  unsigned long long* pkt_off;  // out: index of the packet's first emitted point
  unsigned long long* st_wrap;
  unsigned long long* st_cnt;
  int* tile_counter;
  int n;
  int halo;
  int mode;
  int n_poses;
  int carry_meta_inited;
  long long t_base;
  const long long* pose_t;
  const double* pose_trv;  // n_poses x 9
  double carry_origin_T[3];
  double* pose_mat;  // n x 12
  long long* frame_first_point;
  int* frame_start_block;
  int frame_cap;
  BatchHeader* hdr;
};

constexpr int kPoseThreads = 256;

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

__global__ void __launch_bounds__(kPoseThreads) k_pose(const PoseParams p) {
  __shared__ unsigned long long s_w[kPoseThreads / 32];
  __shared__ unsigned long long s_c[kPoseThreads / 32];
  __shared__ unsigned long long s_wrap_prefix, s_cnt_prefix;
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  const int P = tile * kPoseThreads + tid;
  const bool live = P < p.n;

  PktSeg seg = make_int4(0, 0, 0, 0);
  if (live) seg = p.pkt_seg[P];
  const unsigned wrapmask = (seg.x >> 4) & 0xfff;
  const int nw = __popc(wrapmask);
  // origin marker: streaming -> the packet after the wrap re-initialises the frame meta
  // (F4b); offline -> the wrap packet itself.  marker - 1 == origin packet index.
  const unsigned marker = nw ? (unsigned)(P + (p.mode == 0 ? 2 : 1)) : 0u;
  const unsigned long long v2 = ((unsigned long long)nw << 32) | marker;
  const unsigned long long cnt = (live && P >= p.halo) ? (unsigned long long)(unsigned)seg.y : 0ull;

  unsigned long long inc2 = v2, incc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, inc2, o);
    const unsigned long long b = __shfl_up_sync(0xffffffffu, incc, o);
    if (lane >= o) {
      inc2 = WrapTraits::combine(a, inc2);
      incc += b;
    }
  }
  if (lane == 31) {
    s_w[warp] = inc2;
    s_c[warp] = incc;
  }
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < kPoseThreads / 32) ? s_w[lane] : 0ull;
    unsigned long long c = (lane < kPoseThreads / 32) ? s_c[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < kPoseThreads / 32; o <<= 1) {
      const unsigned long long a = __shfl_up_sync(0xffffffffu, w, o);
      const unsigned long long b = __shfl_up_sync(0xffffffffu, c, o);
      if (lane >= o) {
        w = WrapTraits::combine(a, w);
        c += b;
      }
    }
    if (lane < kPoseThreads / 32) {
      s_w[lane] = w;
      s_c[lane] = c;
    }
  }
  __syncthreads();
  // two look-backs side by side: warp 0 wraps/origin, warp 1 point counts
  if (warp == 0) {
    const unsigned long long ex =
        lookback_exclusive<WrapTraits>(p.st_wrap, tile, s_w[kPoseThreads / 32 - 1]);
    if (lane == 0) s_wrap_prefix = ex;
  } else if (warp == 1) {
    const unsigned long long ex =
        lookback_exclusive<SumTraits>(p.st_cnt, tile, s_c[kPoseThreads / 32 - 1]);
    if (lane == 0) s_cnt_prefix = ex;
  }
  __syncthreads();
  unsigned long long ex2 = __shfl_up_sync(0xffffffffu, inc2, 1);
  unsigned long long exc = __shfl_up_sync(0xffffffffu, incc, 1);
  if (lane == 0) {
    ex2 = 0ull;
    exc = 0ull;
  }
  if (warp > 0) {
    ex2 = WrapTraits::combine(s_w[warp - 1], ex2);
    exc += s_c[warp - 1];
  }
  ex2 = WrapTraits::combine(s_wrap_prefix, ex2);
  exc += s_cnt_prefix;
  // seed: with no frame meta carried in, packet 0 is the origin until the first wrap
  if (!p.carry_meta_inited) ex2 = WrapTraits::combine(ex2, 1ull);
  if (!live) return;

  const int frame_base = (int)(ex2 >> 32);
  const int origin = (int)(unsigned)ex2 - 1;
  const long long t = __ldg(&p.pkt_time[P]);
  seg.y = frame_base;
  seg.z = (int)(unsigned)(t - p.t_base);
  p.pkt_seg[P] = seg;
  p.pkt_off[P] = exc;

  // frame table: a wrap block opens frame f before it is decoded (HDLParser.cxx:1035-1039)
  if (wrapmask && P >= p.halo) {
    unsigned wm = wrapmask;
    unsigned before = 0;
    int jprev = 0;
    int f = frame_base;
    while (wm) {
      const int j = __ffs(wm) - 1;
      wm &= wm - 1;
      for (int q = jprev; q < j; ++q) before += __popc(__ldg(&p.masks[(long long)P * kBlocks + q]));
      jprev = j;
      ++f;
      if (f < p.frame_cap) {
        p.frame_first_point[f] = (long long)(exc + before);
        p.frame_start_block[f] = P * 12 + j;
      } else {
        p.hdr->frame_overflow = 1;
      }
    }
  }
  if (P == p.halo) {
    p.hdr->origin_at_halo = origin;
    p.hdr->frame_at_halo = frame_base;
  }
  if (P == p.n - 1) {
    p.hdr->total_wraps = frame_base + nw;
    p.hdr->last_has_wrap = nw > 0;
    p.hdr->total_points = (long long)(exc + cnt);
    int lo;
    if (p.mode == 0)
      lo = nw ? -2 : origin;  // -2: frame meta not initialised yet
    else
      lo = nw ? P : origin;
    p.hdr->last_origin_packet = lo;
    p.hdr->last_origin_time = (lo >= 0) ? p.pkt_time[lo] : 0;
  }

  if (p.n_poses < 2) return;
  double T[3], R[3];
  interp_pose(p.pose_t, p.pose_trv, p.n_poses, t, T, R, true);
  double To[3];
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
    const bool self = (p.mode == 1) && nw;  // offline: the wrap packet is its frame's origin
#pragma unroll
    for (int k = 0; k < 3; ++k) p.hdr->carry_origin_T[k] = self ? T[k] : To[k];
  }
}

// =========================================================================================
// k_decode: decode + calibrate + transform + compacted SoA stores.  No inter-CTA dependency:
// emission masks and point offsets come from k_scan / k_pose.  Persistent CTAs (2 per SM),
// tiles staged by TMA bulk copies (packets, masks, segment records), one warp per 100-byte
// firing block with lane == return slot.
// Work split inside a tile: warp w owns the packets {w/2 + 4k} and, inside them, the firing
// blocks of parity w&1.  On HDL-64 data block parity == laser bank (0xeeff / 0xddff), so a
// warp keeps one calibration bank in registers; the pose row is loaded once per packet.
// =========================================================================================
struct DecParams {
  const uint8_t* pkts;  // first packet of the submitted array (halo included)
  long long stride;
  long long total_bytes;  // bytes that may be read from pkts
  const PktSeg* pkt_seg;
  const unsigned* masks;
  const unsigned long long* pkt_off;
  const double* pose_mat;
  const double* lut_sin;
  const double* lut_cos;
  const DevConfig* cfg;
  int n;  // packets including the halo
  int halo;
  int mode;
  int pose_valid;
  int n_tiles;
  int stage_bytes;  // bytes per shared-memory stage (multiple of 128)
  float* x;
  float* y;
  float* z;
  uint8_t* intensity;
  uint8_t* laser;
  uint16_t* azimuth;
  uint16_t* distance;
  uint32_t* t_us;
  int* tile_counter;
  unsigned* frame_laser_counts;  // frame_cap x 64
  int frame_cap;
};

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kMaskBytes = kTileBlocks * 4;           // 1536
constexpr int kSegBytes = kTilePkts * (int)sizeof(PktSeg);  // 512

constexpr int kPoseBytes = kTilePkts * 12 * 8;              // 3072

struct DecShared {
  DevConfig cfg;
  double sn[kTileBlocks];  // sin / cos of each firing block's azimuth (ADJ == 0), gathered
  double cs[kTileBlocks];  //   from the LUT once per tile so the block loop has no global loads
  uint64_t full[2];
  unsigned long long off[2][kTilePkts];
  unsigned hist[2][kMaxLasers];
  int tile_id[2];
};

template <int ADJ>
__global__ void __launch_bounds__(kDecThreads, 2) k_decode(const DecParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  DecShared& sh = *reinterpret_cast<DecShared*>(smem_raw);
  uint8_t* stage0 = smem_raw + ((sizeof(DecShared) + 127) & ~127);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp & 1, pk0 = warp >> 1;
  const unsigned lt_mask = (1u << lane) - 1u;

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
  // stage layout: [masks 1536 B | seg 512 B | pose rows 3072 B | packets]
  auto issue = [&](int t, int b) {
    const long long first = (long long)p.halo + (long long)t * kTilePkts;
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, first, 0);
    uint8_t* st = stage0 + (size_t)b * p.stage_bytes;
    const uint32_t bytes = (uint32_t)(sp.s1 - sp.s0);
    const uint32_t mbytes = (uint32_t)sp.npk * kBlocks * 4u;
    const uint32_t sbytes = (uint32_t)sp.npk * (uint32_t)sizeof(PktSeg);
    fence_proxy_async();
    const uint32_t pbytes = p.pose_valid ? (uint32_t)sp.npk * 96u : 0u;
    mbar_expect_tx(&sh.full[b], bytes + mbytes + sbytes + pbytes);
    bulk_g2s(st, p.masks + first * kBlocks, mbytes, &sh.full[b]);
    bulk_g2s(st + kMaskBytes, p.pkt_seg + first, sbytes, &sh.full[b]);
    if (pbytes) bulk_g2s(st + kMaskBytes + kSegBytes, p.pose_mat + first * 12, pbytes, &sh.full[b]);
    if (bytes)
      bulk_g2s(st + kMaskBytes + kSegBytes + kPoseBytes, reinterpret_cast<const void*>(sp.s0), bytes,
               &sh.full[b]);
  };

  unsigned long long next_off = 0;  // threads 0..31: point offset of packet `tid` of the next tile
  if (tid == 0) {
    const int t = atomicAdd(p.tile_counter, 1);
    sh.tile_id[0] = t;
    if (t < p.n_tiles) issue(t, 0);
  }
  __syncthreads();
  {
    const int t0 = sh.tile_id[0];
    if (tid < kTilePkts && t0 < p.n_tiles) {
      const long long P = (long long)p.halo + (long long)t0 * kTilePkts + tid;
      sh.off[0][tid] = __ldg(&p.pkt_off[P < p.n ? P : p.n - 1]);
    }
  }

  uint32_t phase[2] = {0u, 0u};
  int cur = 0;
  const bool pose_valid = p.pose_valid != 0;
  CalRow cal;
  int cal_bank = -1;

  while (true) {
    const int tile = sh.tile_id[cur];
    if (tile >= p.n_tiles) break;
    if (tid == 0) {
      const int tn = atomicAdd(p.tile_counter, 1);
      sh.tile_id[cur ^ 1] = tn;
      if (tn < p.n_tiles) issue(tn, cur ^ 1);
    }
    if (tid >= 128 && tid < 128 + 2 * kMaxLasers) (&sh.hist[0][0])[tid - 128] = 0;
    const long long first = (long long)p.halo + (long long)tile * kTilePkts;
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, first, 0);
    const int npk = sp.npk;
    uint8_t* stage = stage0 + (size_t)cur * p.stage_bytes;
    const uint32_t mask_a = smem_u32(stage);
    const uint32_t seg_a = mask_a + kMaskBytes;
    const uint32_t pose_a = seg_a + kSegBytes;
    const uint32_t off_a = smem_u32(&sh.off[cur][0]);
    uint8_t* s_pk = stage + kMaskBytes + kSegBytes + kPoseBytes;

    mbar_wait(&sh.full[cur], phase[cur]);
    phase[cur] ^= 1u;
    if (sp.s1 < sp.a1) {
      for (long long a = sp.s1 + tid; a < sp.a1; a += kDecThreads)
        s_pk[a - sp.s0] = *reinterpret_cast<const uint8_t*>(a);
    }
    __syncthreads();  // tile_id[cur^1], off[cur], hist and the tail bytes are visible
    if (ADJ == 0) {
      // sin/cos of every block azimuth of the tile: independent gathers, one round trip
      const uint8_t* base = s_pk + (sp.a0 - sp.s0);
      for (int b = tid; b < npk * kBlocks; b += kDecThreads) {
        const int lp = b / kBlocks, j = b - lp * kBlocks;
        const unsigned az = ld_smem_u16(base + (size_t)lp * p.stride + 100 * j + 2) % 36000u;
        sh.sn[b] = __ldg(&p.lut_sin[az]);
        sh.cs[b] = __ldg(&p.lut_cos[az]);
      }
    }
    // point offsets of the next tile: loaded now, parked in a register until the switch
    {
      const int tn = sh.tile_id[cur ^ 1];
      if (tid < kTilePkts && tn < p.n_tiles) {
        const long long P = (long long)p.halo + (long long)tn * kTilePkts + tid;
        next_off = __ldg(&p.pkt_off[P < p.n ? P : p.n - 1]);
      }
    }
    const uint8_t* tile_smem = s_pk + (sp.a0 - sp.s0);
    const unsigned long long tb = lds_u64(off_a);
    // per-tile column pointers, kept opaque so every store is base + 32-bit index
    float* xt = p.x + tb;
    float* yt = p.y + tb;
    float* zt = p.z + tb;
    uint32_t* tt = p.t_us + tb;
    uint16_t* at = p.azimuth + tb;
    uint16_t* dt = p.distance + tb;
    uint8_t* it = p.intensity + tb;
    uint8_t* lt = p.laser + tb;
    asm volatile("" : "+l"(xt), "+l"(yt), "+l"(zt), "+l"(tt));
    asm volatile("" : "+l"(at), "+l"(dt), "+l"(it), "+l"(lt));
    const int tile_f0 = lds_v4(seg_a).y;
    if (ADJ == 0) __syncthreads();  // sn / cs are complete

    unsigned cnt = 0;  // emitted points of (cnt_frame, cnt_bank) seen by this lane
    int cnt_frame = -1, cnt_bank = 0;
    auto flush_counts = [&]() {
      if (cnt) {
        int l = lane + cnt_bank;
        if (ADJ == 2 && l >= 16) l -= 16;
        const int d = cnt_frame - tile_f0;
        if (d == 0 || d == 1)
          atomicAdd(&sh.hist[d][l], cnt);
        else if (cnt_frame < p.frame_cap)
          atomicAdd(&p.frame_laser_counts[(long long)cnt_frame * kMaxLasers + l], cnt);
        cnt = 0;
      }
    };
    // 32-bit shared-window addresses, kept opaque so they stay in registers
    uint32_t tile_a = smem_u32(tile_smem) + 3u * (unsigned)lane;
    uint32_t sn_a = smem_u32(&sh.sn[0]);
    asm volatile("" : "+r"(tile_a), "+r"(sn_a));
    const unsigned pm = par ? 0xaaau : 0x555u;  // the blocks of this warp

    // one firing block: loads from shared memory, FP64 math, predicated SoA stores
    auto do_block = [&](const int lp, const int j, const unsigned m, const unsigned boff,
                        const int off, const int laser_id, const unsigned pkt_rel,
                        const unsigned tpk, const int azdiff, const uint32_t pk_a,
                        const double* M) {
      const uint32_t blk_a = pk_a + 100u * (unsigned)j;
      const unsigned rot = lds_u16(blk_a + 2u - 3u * (unsigned)lane);
      const unsigned dist = lds_u8(blk_a + 4u) | (lds_u8(blk_a + 5u) << 8);
      const unsigned inten = lds_u8(blk_a + 6u);
      const unsigned az = adjusted_azimuth<ADJ>(sh.cfg, rot, azdiff, j, lane);
      double sA, cA;
      if (ADJ == 0) {
        const uint32_t q = sn_a + 8u * (unsigned)(lp * kBlocks + j);
        sA = lds_f64(q);
        cA = lds_f64(q + (uint32_t)(kTileBlocks * 8));
      } else {
        sA = __ldg(&p.lut_sin[az]);
        cA = __ldg(&p.lut_cos[az]);
      }
      double px, py, pz;
      sensor_point(cal, sA, cA, dist, px, py, pz);
      if (pose_valid) {
        // type_defs.h:160-166: row sums left to right, translation last
        const double qx = __dadd_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(M[0], px), __dmul_rn(M[1], py)), __dmul_rn(M[2], pz)),
            M[3]);
        const double qy = __dadd_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(M[4], px), __dmul_rn(M[5], py)), __dmul_rn(M[6], pz)),
            M[7]);
        const double qz = __dadd_rn(
            __dadd_rn(__dadd_rn(__dmul_rn(M[8], px), __dmul_rn(M[9], py)), __dmul_rn(M[10], pz)),
            M[11]);
        px = qx;
        py = qy;
        pz = qz;
      }
      if ((m >> lane) & 1u) {
        const unsigned o = pkt_rel + boff + __popc(m & lt_mask);
        stg_f32(xt + o, (float)px);
        stg_f32(yt + o, (float)py);
        stg_f32(zt + o, (float)pz);
        stg_u32(tt + o, tpk + (ADJ != 0 ? (uint32_t)sh.cfg.tadj[j][lane] : 0u));
        stg_u16(at + o, az);
        stg_u16(dt + o, dist);
        stg_u8(it + o, inten);
        stg_u8(lt + o, (unsigned)laser_id);
        ++cnt;
      }
    };

#pragma unroll 1
    for (int lp = pk0; lp < npk; lp += 4) {
      const PktSeg seg = lds_v4(seg_a + 16u * (unsigned)lp);
      const int wrapmask = (seg.x >> 4) & 0xfff;
      const int azdiff = (seg.x >> 16) & 0xffff;
      const unsigned um = (unsigned)seg.w & 0xfffu;
      const uint32_t pk_a = tile_a + (unsigned)lp * (unsigned)p.stride;
      const unsigned tpk = (unsigned)seg.z;
      // exclusive prefix of the packet's 12 block counts (lanes 0..11)
      const unsigned mymask =
          (lane < kBlocks) ? lds_u32(mask_a + 4u * (unsigned)(lp * kBlocks + lane)) : 0u;
      unsigned pre = __popc(mymask);
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += v;
      }
      pre -= __popc(mymask);
      const unsigned pkt_rel = (unsigned)(lds_u64(off_a + 8u * (unsigned)lp) - tb);
      double M[12];  // [L | t] of this packet, warp-uniform
      if (pose_valid) {
#pragma unroll
        for (int q = 0; q < 12; ++q) M[q] = lds_f64(pose_a + 8u * (unsigned)(lp * 12 + q));
      }
      const unsigned ub = um & pm;
      if (wrapmask == 0 && (ub == 0u || ub == pm)) {
        // fast path (every packet of a real stream but the ~0.3 % that hold a wrap): one
        // frame, one laser bank for all blocks of this warp
        const int off = ub ? 32 : 0;
        if (seg.y != cnt_frame || off != cnt_bank) {
          flush_counts();
          cnt_frame = seg.y;
          cnt_bank = off;
        }
        if (off != cal_bank) {
          load_cal(sh.cfg, lane + off, cal);
          cal_bank = off;
        }
        int laser_id = lane + off;
        if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
#pragma unroll 2
        for (int i = 0; i < 6; ++i) {
          const int j = par + 2 * i;
          const unsigned m = __shfl_sync(0xffffffffu, mymask, j);
          const unsigned boff = __shfl_sync(0xffffffffu, pre, j);
          if (m != 0u) do_block(lp, j, m, boff, off, laser_id, pkt_rel, tpk, azdiff, pk_a, M);
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          const int j = par + 2 * i;
          const unsigned m = __shfl_sync(0xffffffffu, mymask, j);
          const unsigned boff = __shfl_sync(0xffffffffu, pre, j);
          if (m == 0u) continue;
          const int off = ((um >> j) & 1u) ? 32 : 0;
          const int frame = seg.y + __popc(wrapmask & ((2 << j) - 1));
          if (p.mode == 1 && pose_valid && ((wrapmask & ((2 << j) - 1)) != 0)) {
            // offline: blocks at/after the packet's first wrap start a frame whose origin is
            // this very packet -> zero translation
            M[3] = 0.0;
            M[7] = 0.0;
            M[11] = 0.0;
          }
          if (frame != cnt_frame || off != cnt_bank) {
            flush_counts();
            cnt_frame = frame;
            cnt_bank = off;
          }
          if (off != cal_bank) {
            load_cal(sh.cfg, lane + off, cal);
            cal_bank = off;
          }
          int laser_id = lane + off;
          if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
          do_block(lp, j, m, boff, off, laser_id, pkt_rel, tpk, azdiff, pk_a, M);
        }
      }
    }
    flush_counts();
    __syncthreads();
    if (tid < 2 * kMaxLasers) {
      const unsigned c = (&sh.hist[0][0])[tid];
      const int f = tile_f0 + (tid >> 6);
      if (c && f < p.frame_cap)
        atomicAdd(&p.frame_laser_counts[(long long)f * kMaxLasers + (tid & 63)], c);
    }
    if (tid < kTilePkts) sh.off[cur ^ 1][tid] = next_off;
    __syncthreads();  // stage `cur` and hist are free again; off[cur^1] is published
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
