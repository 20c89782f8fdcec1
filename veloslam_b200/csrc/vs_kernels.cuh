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
                                              int lead, int tile_pkts = kTilePkts) {
  TileSpan t;
  t.npk = n - (int)first;
  if (t.npk > tile_pkts) t.npk = tile_pkts;
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
//   per block  : BlkRec {mask of the return slots the reference emits, azimuth, points of the
//                packet in front of the block, laser bank, wraps up to the block} -> recs[n*12]
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
  BlkRec* recs;
  unsigned long long* st_map;
  unsigned long long* agg_wrap;  // per tile: (wraps << 32) | max origin marker
  unsigned long long* agg_cnt;   // per tile: emitted points of the decoded packets
  unsigned long long* grp_cnt;   // the same per group of kGroupTiles tiles (atomics)
  unsigned* grp_wsum;
  unsigned* grp_wmax;
  int* tile_counter;
  BatchHeader* hdr;
};

constexpr int kScanThreads = 256;
// Two-level aggregates for the packet scans of k_pose: per tile (32 packets) and per group of
// 256 tiles (8192 packets).  A k_pose CTA reduces the groups in front of its own and the tiles
// of its group in front of it: no scan kernel and no look-back chain in between.
constexpr int kGroupTiles = 256;

struct ScanShared {
  uint64_t full;
  unsigned nz[kTileBlocks];  // raw "distance != 0" (or crop-tested) bits per block
  int skip[kTilePkts];
  unsigned wrap[kTilePkts];  // wrap mask over the iterated blocks of each packet
  // block headers kept from the header pass, so that nothing reads the stage after the mid-tile
  // barrier and the next tile's copy can start there
  unsigned short hdr_az[kTilePkts][kBlocks];
  unsigned hdr_um[kTilePkts];  // bit j: block j is a 0xddff ("upper") block
  int tile_id;                 // first tile of the CTA
  int tile_next;               // tile whose copy was started at the mid-tile barrier
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
  // no wrap inside the packet and none against the azimuth in front of it (all but one packet
  // per rotation): every firingSkip leaves as 0
  if ((wm | em) == 0u) return 0ull;
  unsigned long long m = 0;
#pragma unroll
  for (int s = 0; s < 12; ++s) {
    const unsigned hi = wm & ~((2u << s) - 1u);
    const int out = hi ? (31 - __clz(hi)) : (((em >> s) & 1u) ? s : 0);
    m |= (unsigned long long)out << (4 * s);
  }
  return m;
}

// One stage per CTA, 4-5 CTAs per SM: the TMA latency of one CTA is hidden by the others, and
// inside a CTA the copy of the next tile starts at the mid-tile barrier (the block-record pass
// works from shared-memory copies of the headers), its tile id dealt one tile earlier still.
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
  unsigned gate_mask = 0;  // bit j: block j passes the pointsSkip gate (HDLParser.cxx:1042)
#pragma unroll
  for (int j = 0; j < kBlocks; ++j)
    if (pskip == 0 || (j % (pskip + 1)) == 0) gate_mask |= 1u << j;

  // start the copy of tile t into the stage (thread 0)
  auto start_copy = [&](int t) {
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
  };
  long long fub_seen = 0x7fffffffffffffffll;  // smallest first-upper-block candidate of this thread
  int next_t = 0;  // thread 0: id of the tile after the one being copied
  if (tid == 0) {
    const int t = atomicAdd(p.tile_counter, 1);
    sh.tile_id = t;
    start_copy(t);
    next_t = atomicAdd(p.tile_counter, 1);
  }
  __syncthreads();
  int tile = sh.tile_id;
  uint32_t phase = 0u;
  while (tile < p.n_tiles) {
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
#pragma unroll
        for (int j = 0; j < 12; j += 2)
          *reinterpret_cast<unsigned*>(&sh.hdr_az[lane][j]) = (unsigned)az[j] | ((unsigned)az[j + 1] << 16);
        sh.hdr_um[lane] = um;
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
      sh.wrap[lane] = wrapmask;
    }

    // ---- phase B1: which return slots of each block could be emitted ------------------------
    // (distance != 0 / crop, laser selection).  Independent of phase A: warps 1..7 run it while
    // warp 0 resolves the firingSkip chain.
    if (!CROP) {
      // lane per block: scan the 32 distance fields of a block with 16-bit loads (the 100-byte
      // block stride spreads the lanes of a warp over the banks)
      if (warp > 0) {
        for (int b = tid - 32; b < kTileBlocks; b += kScanThreads - 32) {
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
            bits &= (ld_smem_u16(blk) != 0xeeffu) ? sel_hi : sel_lo;
          }
          sh.nz[b] = bits;
        }
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
          bits &= off ? sel_hi : sel_lo;
        }
        if (lane == 0) sh.nz[b] = bits;
      }
    }
    __syncthreads();  // skip / wrap / headers (phase A) and nz (phase B1) are complete
    // nothing reads the stage from here on: the next tile's copy overlaps the record pass
    if (tid == 0) {
      sh.tile_next = next_t;
      start_copy(next_t);
      next_t = atomicAdd(p.tile_counter, 1);
    }

    // ---- phase B2: block records: the final mask applies the gates of HDLParser.cxx:1042-1051
    // (block iterated: j >= firingSkip; pointsSkip) to the slot bits -----------------------------
    // Half a warp per packet (lanes 0-11 of each half == its blocks): the points in front of a
    // block are a 16-lane prefix sum of the open blocks' popcounts.
    {
      static_assert(kTilePkts % (2 * (kScanThreads / 32)) == 0, "whole passes of two packets per warp");
      const int j = lane & 15;
#pragma unroll 1
      for (int lp = 2 * warp + (lane >> 4); lp < kTilePkts; lp += 2 * (kScanThreads / 32)) {
        const bool act = lp < npk && j < kBlocks;
        unsigned bits = 0, open = 0;
        if (act) {
          // blocks the parser iterates and emits: j >= firingSkip and the pointsSkip gate
          open = ((gate_mask & ~((1u << sh.skip[lp]) - 1u)) >> j) & 1u;
          bits = sh.nz[lp * kBlocks + j];
        }
        const unsigned v = open ? __popc(bits) : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, inc, o, 16);
          if (j >= o) inc += t;
        }
        if (act) {
          const unsigned pre = inc - v;
          const unsigned upper = (sh.hdr_um[lp] >> j) & 1u;
          const unsigned wb = __popc(sh.wrap[lp] & ((2u << j) - 1u));
          unsigned az = sh.hdr_az[lp][j];
          if (ADJ == 0) az %= 36000u;  // HDLParser.cxx:597 (ADJ != 0: adjusted per return later)
          BlkRec r;
          r.x = open ? bits : 0u;
          r.y = az | (pre << 16) | (upper << 25) | (wb << 26);
          p.recs[(first + lp) * kBlocks + j] = r;
        }
      }
    }
    if (warp == 0) {
      unsigned cnt = 0;
      if (lane < npk) {
        const unsigned open_mask = gate_mask & ~((1u << s_in) - 1u);
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          cnt += ((open_mask >> j) & 1u) ? __popc(sh.nz[lane * kBlocks + j]) : 0u;
        const long long P = first + lane;
        PktSeg r;
        r.x = s_in | (int)(wrapmask << 4) | (azdiff << 16);
        r.y = (int)cnt;
        r.z = 0;
        r.w = (int)(um | (cnt << 12));
        p.pkt_seg[P] = r;
        const unsigned ium = um & ~((1u << s_in) - 1u);
        if (ium) {
          const long long fu = P * 12 + (__ffs(ium) - 1);
          // a CTA's tile ids only grow: after a lane's first candidate every later one is larger,
          // so the header is touched once per lane and kernel
          if (fu < fub_seen) {
            const unsigned long long old = atomicMin(
                reinterpret_cast<unsigned long long*>(&p.hdr->first_upper_block), (unsigned long long)fu);
            fub_seen = min((long long)old, fu);
          }
        }
        if (P == p.n - 1) {
          p.hdr->last_azimuth = az11;
          p.hdr->firing_skip_out = (p.mode == 0) ? map_apply(m, s_in) : 0;
        }
      }
      // tile aggregates for the frame-id / origin / point-offset scans (k_tilescan, k_pose).
      // Origin marker: streaming -> the packet after the wrap re-initialises the frame meta
      // (F4b); offline -> the wrap packet itself.  marker - 1 == origin packet index.
      const long long P = first + lane;
      const unsigned nw = __popc(wrapmask);
      // per-tile sums fit 32 bits: one REDUX each (wraps, latest origin marker, points)
      const unsigned tw = __reduce_add_sync(0xffffffffu, nw);
      const unsigned tm = __reduce_max_sync(0xffffffffu, nw ? (unsigned)(P + (p.mode == 0 ? 2 : 1)) : 0u);
      const unsigned long long aw = ((unsigned long long)tw << 32) | tm;
      const unsigned long long ac =
          __reduce_add_sync(0xffffffffu, (lane < npk && P >= p.halo) ? cnt : 0u);
      if (lane == 0) {
        p.agg_wrap[tile] = aw;
        p.agg_cnt[tile] = ac;
        const int g = tile / kGroupTiles;
        if (ac) atomicAdd(&p.grp_cnt[g], ac);
        if (aw) {
          atomicAdd(&p.grp_wsum[g], (unsigned)(aw >> 32));
          atomicMax(&p.grp_wmax[g], (unsigned)aw);
        }
      }
    }
    __syncthreads();  // nz, skip, wrap and the headers are free again
    tile = sh.tile_next;
  }
}

// =========================================================================================
// k_pose: one thread per packet.
//   (1) scans over packets (block scan on top of k_scan's tile / group aggregates): wraps -> frame
//       id, last wrap -> origin packet, emitted counts -> point offset; frame-start records
//       of the frame table;
//   (2) pose bracket + lerp + Ry.Rx.Rz and T(packet) - T(origin) when the snapshot has >= 2
//       poses: 12 doubles per packet, row-major [L | t].
// =========================================================================================
struct PoseParams {
  const long long* pkt_time;
  PktSeg* pkt_seg;           // in: x, y = count; out: y = frame id, z = time - t_base
  const BlkRec* recs;
  unsigned long long* pkt_off;  // out: index of the packet's first emitted point
  const unsigned long long* agg_wrap;  // aggregates per 32-packet tile / per group (k_scan)
  const unsigned long long* agg_cnt;
  const unsigned long long* grp_cnt;
  const unsigned* grp_wsum;
  const unsigned* grp_wmax;
  int n;
  int halo;
  int mode;
  int n_poses;
  int carry_meta_inited;
  int check_time;  // flag packet times outside [t_base, t_base + kTimeSpanMax)
  long long t_base;
  const long long* pose_t;
  const double* pose_trv;  // n_poses x 9
  double carry_origin_T[3];
  int deskew;                   // per-point deskew rows (kDeskewRow doubles) instead of [L | t]
  long long carry_origin_time;  // time of the carried frame's origin packet (deskew, origin < 0)
  double* pose_mat;  // n x 12 (n x kDeskewRow with deskew)
  long long* frame_first_point;
  int* frame_start_block;
  int frame_cap;
  BatchHeader* hdr;
};

// -DVS_POSE_THREADS=128: at 76 registers a 128-thread k_pose CTA fits into what two resident
// k_decode CTAs leave of an SM's register file (co-residency experiment, see OutCols)
#ifndef VS_POSE_THREADS
#define VS_POSE_THREADS 256
#endif
// -DVS_POSE_FAST=1: k_pose requests the packet time and the pose table's ends before its scans
// and resolves the packet's and the frame origin's pose brackets side by side (four probe loads
// in flight instead of two dependent searches).  Bit-identical results; measured in same-box A/B
// runs at -0.3 % of a step with 3 CTAs per SM and nothing on top of 4 CTAs per SM (below), so the
// plain version stays the default.
#ifndef VS_POSE_FAST
#define VS_POSE_FAST 0
#endif
// -DVS_POSE_COALESCED=0: every thread stores its own 96-byte pose row (12 strided 8-byte stores)
#ifndef VS_POSE_COALESCED
#define VS_POSE_COALESCED 1
#endif
constexpr int kPoseThreads = VS_POSE_THREADS;

__device__ __forceinline__ double to_radians(double x) {
  return __ddiv_rn(__dmul_rn(x, 3.14159265358979323846), 180.0);
}

// TimeLine::getBoundaryData net semantics (TimeLine.h:384-468): i = clamp(lower_bound, 1, N-1),
// bracket (i-1, i); then TransformManager.cxx:168-175 fore + (back - fore) * ratio.
// The lower bound starts from an interpolated guess (INS poses are evenly sampled, so the guess
// is the answer or next to it: two loads instead of log2 N dependent ones) and widens the
// bracket geometrically when the guess is off; the result is the exact lower bound for any
// strictly increasing pose times.
__device__ __forceinline__ int pose_lower_bound(const long long* __restrict__ pt, int np, long long t) {
  const long long t_first = __ldg(&pt[0]), t_last = __ldg(&pt[np - 1]);
  if (t <= t_first) return 0;
  if (t > t_last) return np;
  int a = 0, b = np - 1;  // pt[a] < t <= pt[b]
  int g = (int)((double)(t - t_first) / (double)(t_last - t_first) * (double)(np - 1));
  g = g < 1 ? 1 : (g > np - 1 ? np - 1 : g);
  for (int r = 1;; r *= 4) {
    const int ph = min(g - 1 + r, np - 1), pl = max(g - r, 0);
    const bool hi_ok = __ldg(&pt[ph]) >= t, lo_ok = __ldg(&pt[pl]) < t;
    if (hi_ok) b = min(b, ph); else a = max(a, ph);
    if (lo_ok) a = max(a, pl); else b = min(b, pl);
    if (hi_ok && lo_ok) break;
  }
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    if (__ldg(&pt[mid]) < t)
      a = mid;
    else
      b = mid;
  }
  return b;
}
__device__ __forceinline__ void interp_pose(const long long* __restrict__ pt,
                                            const double* __restrict__ trv, int np, long long t,
                                            double T[3], double R[3], bool want_R) {
  const int lo = pose_lower_bound(pt, np, t);
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

// ---- per-point deskew extension (semantics: oracle/deskew_port.py) -------
struct Quat {
  double w, x, y, z;
};
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ Quat quat_conj(const Quat& q) {
  Quat r = {q.w, -q.x, -q.y, -q.z};
  return r;
}
// p + w t + v x t with t = 2 v x p
__device__ __forceinline__ void quat_rotate(const Quat& q, const double p[3], double o[3]) {
  const double tx = 2.0 * (q.y * p[2] - q.z * p[1]);
  const double ty = 2.0 * (q.z * p[0] - q.x * p[2]);
  const double tz = 2.0 * (q.x * p[1] - q.y * p[0]);
  o[0] = p[0] + q.w * tx + (q.y * tz - q.z * ty);
  o[1] = p[1] + q.w * ty + (q.z * tx - q.x * tz);
  o[2] = p[2] + q.w * tz + (q.x * ty - q.y * tx);
}
// Unit quaternion of PoseTransform::getMatrix's Ry(R0) Rx(R1) Rz(R2) (degrees), largest pivot.
__device__ __forceinline__ Quat quat_from_rpy(const double* __restrict__ R) {
  double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  rotate_by(m, to_radians(R[0]), 1);
  rotate_by(m, to_radians(R[1]), 0);
  rotate_by(m, to_radians(R[2]), 2);
  const double tr = m[0][0] + m[1][1] + m[2][2];
  Quat q;
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q.w = 0.25 * s;
    q.x = (m[2][1] - m[1][2]) / s;
    q.y = (m[0][2] - m[2][0]) / s;
    q.z = (m[1][0] - m[0][1]) / s;
  } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
    const double s = sqrt(1.0 + m[0][0] - m[1][1] - m[2][2]) * 2;
    q.w = (m[2][1] - m[1][2]) / s;
    q.x = 0.25 * s;
    q.y = (m[0][1] + m[1][0]) / s;
    q.z = (m[0][2] + m[2][0]) / s;
  } else if (m[1][1] > m[2][2]) {
    const double s = sqrt(1.0 + m[1][1] - m[0][0] - m[2][2]) * 2;
    q.w = (m[0][2] - m[2][0]) / s;
    q.x = (m[0][1] + m[1][0]) / s;
    q.y = 0.25 * s;
    q.z = (m[1][2] + m[2][1]) / s;
  } else {
    const double s = sqrt(1.0 + m[2][2] - m[0][0] - m[1][1]) * 2;
    q.w = (m[1][0] - m[0][1]) / s;
    q.x = (m[0][2] + m[2][0]) / s;
    q.y = (m[1][2] + m[2][1]) / s;
    q.z = 0.25 * s;
  }
  return q;
}
// bracket of time t: i = clamp(lower_bound, 1, N-1)
__device__ __forceinline__ int pose_bracket(const long long* __restrict__ pt, int np, long long t) {
  const int lo = pose_lower_bound(pt, np, t);
  const int i = lo < 1 ? 1 : lo;
  return i > np - 1 ? np - 1 : i;
}
// slerp weights along the shorter arc; theta = angle between the (sign-aligned) quaternions
__device__ __forceinline__ void slerp_weights(double theta, double r, double& w0, double& w1) {
  if (theta < 1e-8) {
    w0 = 1.0 - r;
    w1 = r;
  } else {
    const double inv = 1.0 / sin(theta);
    w0 = sin((1.0 - r) * theta) * inv;
    w1 = sin(r * theta) * inv;
  }
}
// the pose function of the extension at time t, evaluated in t's own bracket
__device__ __forceinline__ void pose_at(const long long* __restrict__ pt, const double* __restrict__ trv,
                                        int np, long long t, Quat& q, double T[3]) {
  const int i = pose_bracket(pt, np, t);
  const long long ta = __ldg(&pt[i - 1]), tb = __ldg(&pt[i]);
  const double r = (double)(t - ta) / (double)(tb - ta);
  const double* a = trv + (long long)(i - 1) * 9;
  const double* b = trv + (long long)i * 9;
  const Quat qa = quat_from_rpy(a + 3);
  Quat qb = quat_from_rpy(b + 3);
  double d = qa.w * qb.w + qa.x * qb.x + qa.y * qb.y + qa.z * qb.z;
  if (d < 0) {
    qb.w = -qb.w; qb.x = -qb.x; qb.y = -qb.y; qb.z = -qb.z;
    d = -d;
  }
  double w0, w1;
  slerp_weights(acos(fmin(1.0, d)), r, w0, w1);
  q.w = w0 * qa.w + w1 * qb.w;
  q.x = w0 * qa.x + w1 * qb.x;
  q.y = w0 * qa.y + w1 * qb.y;
  q.z = w0 * qa.z + w1 * qb.z;
#pragma unroll
  for (int k = 0; k < 3; ++k) T[k] = __ldg(&a[k]) + (__ldg(&b[k]) - __ldg(&a[k])) * r;
}
// doubles per packet: Q0[4] dQ[4] T0[3] dT[3] -- the blended quaternion / translation at the packet
// time and their derivatives per microsecond of firing offset (see k_pose)
constexpr int kDeskewRow = 14;

// k_pose is a chain of dependent loads per packet: what helps is more packets in flight.  4 CTAs
// of 256 threads per SM (64 registers, 60 bytes of spills) against the 3 the compiler's own 76-80
// registers allow: -0.8 % of a whole step (2.043 against 2.058-2.066 ms, same box, two rounds).
#ifndef VS_POSE_CTAS
#define VS_POSE_CTAS 4
#endif
__global__ void __launch_bounds__(kPoseThreads, VS_POSE_CTAS) k_pose(const PoseParams p) {
  __shared__ unsigned long long s_w[kPoseThreads / 32];
  __shared__ unsigned long long s_c[kPoseThreads / 32];
  __shared__ unsigned long long s_pw[kPoseThreads / 32], s_pc[kPoseThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int P = tile * kPoseThreads + tid;
  const bool live = P < p.n;
  PktSeg seg = make_int4(0, 0, 0, 0);
  if (live) seg = p.pkt_seg[P];
#if VS_POSE_FAST
  // everything that does not depend on the scans is requested before them: the kernel is a
  // chain of dependent loads, not work
  const long long t = live ? __ldg(&p.pkt_time[P]) : 0ll;
  long long pt_first = 0, pt_last = 0;
  if (p.n_poses >= 2) {
    pt_first = __ldg(&p.pose_t[0]);
    pt_last = __ldg(&p.pose_t[p.n_poses - 1]);
  }
#endif
  const unsigned wrapmask = (seg.x >> 4) & 0xfff;
  const int nw = __popc(wrapmask);
  // origin marker: streaming -> the packet after the wrap re-initialises the frame meta
  // (F4b); offline -> the wrap packet itself.  marker - 1 == origin packet index.
  const unsigned marker = nw ? (unsigned)(P + (p.mode == 0 ? 2 : 1)) : 0u;
  const unsigned long long v2 = ((unsigned long long)nw << 32) | marker;
  const unsigned long long cnt = (live && P >= p.halo) ? (unsigned long long)(unsigned)seg.y : 0ull;

  // what precedes this CTA's 256 packets: the groups in front of its group + the tiles of its
  // group in front of it (k_scan aggregates), reduced by the whole CTA
  unsigned long long pw = 0ull, pc = 0ull;
  {
    static_assert(kPoseThreads % kTilePkts == 0, "a pose tile is a whole number of scan tiles");
    const int my_tile = tile * (kPoseThreads / kTilePkts);  // first scan tile of this CTA
    const int g = my_tile / kGroupTiles;
    for (int q = tid; q < g; q += kPoseThreads) {
      pw = WrapTraits::combine(pw, ((unsigned long long)__ldg(&p.grp_wsum[q]) << 32) | __ldg(&p.grp_wmax[q]));
      pc += __ldg(&p.grp_cnt[q]);
    }
    for (int q = tid; q < kGroupTiles; q += kPoseThreads) {
      const int f = g * kGroupTiles + q;
      if (f < my_tile) {
        pw = WrapTraits::combine(pw, __ldg(&p.agg_wrap[f]));
        pc += __ldg(&p.agg_cnt[f]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pw = WrapTraits::combine(pw, __shfl_xor_sync(0xffffffffu, pw, o));
      pc += __shfl_xor_sync(0xffffffffu, pc, o);
    }
  }
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
  if (lane == 0) {
    s_pw[warp] = pw;
    s_pc[warp] = pc;
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
  unsigned long long s_wrap_prefix = 0ull, s_cnt_prefix = 0ull;
#pragma unroll
  for (int q = 0; q < kPoseThreads / 32; ++q) {
    s_wrap_prefix = WrapTraits::combine(s_wrap_prefix, s_pw[q]);
    s_cnt_prefix += s_pc[q];
  }
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
#if !VS_POSE_FAST
  const long long t = __ldg(&p.pkt_time[P]);
#endif
  seg.y = frame_base;
  seg.z = (int)(unsigned)(t - p.t_base);
  if (p.check_time && P >= p.halo && (unsigned long long)(t - p.t_base) >= kTimeSpanMax)
    p.hdr->time_range_error = 1;
  p.pkt_seg[P] = seg;
  p.pkt_off[P] = exc;

  // frame table: a wrap block opens frame f before it is decoded (HDLParser.cxx:1035-1039)
  if (wrapmask && P >= p.halo) {
    unsigned wm = wrapmask;
    int f = frame_base;
    while (wm) {
      const int j = __ffs(wm) - 1;
      wm &= wm - 1;
      const unsigned before = (__ldg(&p.recs[(long long)P * kBlocks + j].y) >> 16) & 0x1ffu;
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
  if (p.deskew) {
    // per-point deskew: the packet's bracket endpoints, re-based to the frame origin's pose
    Quat qo;
    double To[3];
    pose_at(p.pose_t, p.pose_trv, p.n_poses,
            origin < 0 ? p.carry_origin_time : __ldg(&p.pkt_time[origin]), qo, To);
    const Quat qoc = quat_conj(qo);
    const int i = pose_bracket(p.pose_t, p.n_poses, t);
    const long long ta = __ldg(&p.pose_t[i - 1]), tb = __ldg(&p.pose_t[i]);
    const double* a = p.pose_trv + (long long)(i - 1) * 9;
    const double* b = p.pose_trv + (long long)i * 9;
    const Quat qa = quat_mul(qoc, quat_from_rpy(a + 3));
    Quat qb = quat_mul(qoc, quat_from_rpy(b + 3));
    double d = qa.w * qb.w + qa.x * qb.x + qa.y * qb.y + qa.z * qb.z;
    if (d < 0) {
      qb.w = -qb.w; qb.x = -qb.x; qb.y = -qb.y; qb.z = -qb.z;
      d = -d;
    }
    const double theta = acos(fmin(1.0, d));
    const double da[3] = {__ldg(&a[0]) - To[0], __ldg(&a[1]) - To[1], __ldg(&a[2]) - To[2]};
    const double db[3] = {__ldg(&b[0]) - __ldg(&a[0]), __ldg(&b[1]) - __ldg(&a[1]), __ldg(&b[2]) - __ldg(&a[2])};
    double Ta[3], dT[3];
    quat_rotate(qoc, da, Ta);
    quat_rotate(qoc, db, dT);
    // Slerp weights at the packet time and their slope per microsecond.  Inside one packet the
    // firing offsets span < 1 ms, over which the weights are linear to (omega * 1 ms)^2 / 2 (1e-7
    // at 30 deg/s: < 2e-5 m at the sensor's maximum range, far inside the 1e-3 m bar), so
    // k_decode evaluates q(t_packet + f) = Q0 + f dQ with one fused multiply-add per component
    // instead of two sines per point.
    const double r0 = (double)(t - ta) / (double)(tb - ta);
    const double dr = 1.0 / (double)(tb - ta);
    double w0, w1, dw0, dw1;
    if (theta < 1e-8) {
      w0 = 1.0 - r0;
      w1 = r0;
      dw0 = -dr;
      dw1 = dr;
    } else {
      const double inv = 1.0 / sin(theta);
      double s0, c0, s1, c1;
      sincos((1.0 - r0) * theta, &s0, &c0);
      sincos(r0 * theta, &s1, &c1);
      w0 = s0 * inv;
      w1 = s1 * inv;
      dw0 = -theta * dr * c0 * inv;
      dw1 = theta * dr * c1 * inv;
    }
    double* o = p.pose_mat + (long long)P * kDeskewRow;
    o[0] = w0 * qa.w + w1 * qb.w;
    o[1] = w0 * qa.x + w1 * qb.x;
    o[2] = w0 * qa.y + w1 * qb.y;
    o[3] = w0 * qa.z + w1 * qb.z;
    o[4] = dw0 * qa.w + dw1 * qb.w;
    o[5] = dw0 * qa.x + dw1 * qb.x;
    o[6] = dw0 * qa.y + dw1 * qb.y;
    o[7] = dw0 * qa.z + dw1 * qb.z;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[8 + k] = Ta[k] + r0 * dT[k];
      o[11 + k] = dT[k] * dr;
    }
    if (P == p.n - 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) p.hdr->carry_origin_T[k] = To[k];
    }
    return;
  }
  double T[3], R[3];
  double To[3];
#if VS_POSE_FAST
  {
    // The packet's bracket and its frame origin's, resolved side by side: the interpolated guess
    // (INS poses are evenly sampled) is checked with two loads each, all four in flight at once;
    // a guess that is not the answer takes pose_lower_bound.  Same brackets, same arithmetic
    // as interp_pose.
    const long long* __restrict__ pt = p.pose_t;
    const int np = p.n_poses;
    const bool other = origin >= 0 && origin != P;
    const long long t_o = other ? __ldg(&p.pkt_time[origin]) : t;
    const double span = (double)(pt_last - pt_first), nm1 = (double)(np - 1);
    int i1 = (int)((double)(t - pt_first) / span * nm1);
    int i2 = (int)((double)(t_o - pt_first) / span * nm1);
    i1 = i1 < 1 ? 1 : (i1 > np - 1 ? np - 1 : i1);
    i2 = i2 < 1 ? 1 : (i2 > np - 1 ? np - 1 : i2);
    long long a1 = __ldg(&pt[i1 - 1]), b1 = __ldg(&pt[i1]);
    long long a2 = __ldg(&pt[i2 - 1]), b2 = __ldg(&pt[i2]);
    if (!(a1 < t && t <= b1)) {
      i1 = pose_bracket(pt, np, t);
      a1 = __ldg(&pt[i1 - 1]);
      b1 = __ldg(&pt[i1]);
    }
    if (other && !(a2 < t_o && t_o <= b2)) {
      i2 = pose_bracket(pt, np, t_o);
      a2 = __ldg(&pt[i2 - 1]);
      b2 = __ldg(&pt[i2]);
    }
    const double* f1 = p.pose_trv + (long long)(i1 - 1) * 9;
    const double* g1 = p.pose_trv + (long long)i1 * 9;
    const double* f2 = p.pose_trv + (long long)(i2 - 1) * 9;
    const double* g2 = p.pose_trv + (long long)i2 * 9;
    double fa[6], ga[6], fo[3], go[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      fa[k] = __ldg(&f1[k]);
      ga[k] = __ldg(&g1[k]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      fo[k] = other ? __ldg(&f2[k]) : 0.0;
      go[k] = other ? __ldg(&g2[k]) : 0.0;
    }
    const double r1 = __ddiv_rn((double)(t - a1), (double)(b1 - a1));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      T[k] = __dadd_rn(fa[k], __dmul_rn(__dsub_rn(ga[k], fa[k]), r1));
      R[k] = __dadd_rn(fa[3 + k], __dmul_rn(__dsub_rn(ga[3 + k], fa[3 + k]), r1));
    }
    if (origin < 0) {
      To[0] = p.carry_origin_T[0];
      To[1] = p.carry_origin_T[1];
      To[2] = p.carry_origin_T[2];
    } else if (!other) {
      To[0] = T[0];
      To[1] = T[1];
      To[2] = T[2];
    } else {
      const double r2 = __ddiv_rn((double)(t_o - a2), (double)(b2 - a2));
#pragma unroll
      for (int k = 0; k < 3; ++k) To[k] = __dadd_rn(fo[k], __dmul_rn(__dsub_rn(go[k], fo[k]), r2));
    }
  }
#else
  interp_pose(p.pose_t, p.pose_trv, p.n_poses, t, T, R, true);
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
#endif
  double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  rotate_by(L, to_radians(R[0]), 1);  // type_defs.h:136 UnitY
  rotate_by(L, to_radians(R[1]), 0);  // :137 UnitX
  rotate_by(L, to_radians(R[2]), 2);  // :138 UnitZ
#if VS_POSE_COALESCED
  {
    // The 32 rows of a warp are contiguous in pose_mat (3 KB): staged in shared memory (pitch of
    // 13 doubles: no bank conflicts) and written as 16-byte pieces, consecutive lanes to
    // consecutive addresses -- six whole-sector stores per lane instead of twelve 8-byte stores
    // that each touch 32 sectors (the LSU queue was what k_pose stalled on next to its loads).
    __shared__ double s_rows[kPoseThreads / 32][32 * 13];
    double* sw = s_rows[warp];
    double* sr = sw + lane * 13;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sr[4 * r + 0] = L[r][0];
      sr[4 * r + 1] = L[r][1];
      sr[4 * r + 2] = L[r][2];
      sr[4 * r + 3] = __dsub_rn(T[r], To[r]);  // reprojectToFrameBeginning, HDLParser.cxx:1057
    }
    // the lanes still here are the live ones: lanes [0, rows) of the warp
    const int left = p.n - (P - lane);
    const int rows = left < 32 ? left : 32;
    const unsigned mask = rows == 32 ? 0xffffffffu : ((1u << rows) - 1u);
    __syncwarp(mask);
    double* og = p.pose_mat + (long long)(P - lane) * 12;  // 16-byte aligned: (P - lane) % 32 == 0
    for (int e = lane; e < rows * 6; e += rows) {
      const int row = e / 6, col = (e % 6) * 2;
      const double2 v = make_double2(sw[row * 13 + col], sw[row * 13 + col + 1]);
      *reinterpret_cast<double2*>(og + 2 * e) = v;
    }
  }
#else
  double* o = p.pose_mat + (long long)P * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[4 * r + 0] = L[r][0];
    o[4 * r + 1] = L[r][1];
    o[4 * r + 2] = L[r][2];
    o[4 * r + 3] = __dsub_rn(T[r], To[r]);  // reprojectToFrameBeginning, HDLParser.cxx:1057
  }
#endif
  if (P == p.n - 1) {
    const bool self = (p.mode == 1) && nw;  // offline: the wrap packet is its frame's origin
#pragma unroll
    for (int k = 0; k < 3; ++k) p.hdr->carry_origin_T[k] = self ? T[k] : To[k];
  }
}

// =========================================================================================
// k_decode: decode + calibrate + transform + compacted SoA stores.  No inter-CTA dependency:
// block records (emission masks, azimuths, in-packet offsets) come from k_scan, frame ids,
// point offsets and pose rows from k_pose.
//
// Input side.  Persistent CTAs (2 per SM, 8 warps), tiles of 8 packets dealt round-robin to the
// CTAs.  A tile (block records, segment records, point offsets, pose rows and the packet bytes:
// five TMA bulk copies on one mbarrier) moves through a 3-stage shared-memory ring.  There is
// no block-wide barrier in the tile loop: every warp waits on the stage's "full" mbarrier by
// itself, and the warp that releases a stage last (shared-memory counter) re-arms it.
//
// Work split.  The warp pair (2q, 2q+1) owns the packets 2q, 2q+1 of a tile and warp w the
// firing blocks of parity w&1 inside them (lane == return slot).  On HDL-64 data block parity ==
// laser bank (0xeeff / 0xddff), so a warp keeps one calibration bank in registers.  Per tile a
// warp first expands the records of the 12 firing blocks it owns (lane == block) to 32 bytes:
// mask, staging position of the block's first point, azimuth, laser bank, frame id and sin/cos
// of the azimuth gathered from the LUT.  The block loop then reads everything warp-uniform
// with two broadcast LDS.128 and runs two blocks at a time as one straight-line body
// (independent FP64 chains, predicated stores, no branches).
//
// Output side.  Compaction makes the per-block output ranges start at arbitrary element
// offsets; written straight to HBM that is two partial 32-byte sectors per block and column,
// and partial-sector writes cost the B200 L2 about 3x (measured: 8 compacted columns reach
// 2.2-2.7 TB/s against 6.0-6.9 TB/s for sector-aligned ones).  So a pair stages the points of
// its two packets (<= 768 x 22 B) in shared memory at the same 16-element phase as their global
// position, and after a pair barrier every column leaves as ONE TMA bulk store of whole
// 16-element granules (cp.async.bulk.global.shared::cta); only the < 16 leading / trailing
// elements go out as scalar stores.  The bulk reads overlap the next tile's wait + record
// expansion; a second pair barrier in front of the next tile's first staging write closes the
// loop.
// =========================================================================================
struct DecParams {
  const uint8_t* pkts;  // first packet of the submitted array (halo included)
  long long stride;
  long long total_bytes;  // bytes that may be read from pkts
  const PktSeg* pkt_seg;
  const BlkRec* recs;
  const unsigned long long* pkt_off;  // n + 2 entries allocated
  const double* pose_mat;
  const double* lut_sin;
  const double* lut_cos;
  const DevConfig* cfg;
  int n;  // packets including the halo
  int halo;
  int mode;
  int pose_valid;
  int n_tiles;
  int stage_bytes;  // bytes per shared-memory input stage (multiple of 128)
  float* x;
  float* y;
  float* z;
  uint8_t* intensity;
  uint8_t* laser;
  uint16_t* azimuth;
  uint16_t* distance;
  uint32_t* t_us;
  unsigned* frame_laser_counts;  // frame_cap x 64
  int frame_cap;
  // ---- single-pass variant (FUSED): the segmentation scans run inside the kernel ------------
  const long long* pkt_time;
  long long t_base;
  PktSeg* seg_out;               // per-packet segment records (k_frames and the host read them)
  ulonglong2* st;                // look-back records per tile, 32 bytes: {flag | points, wraps << 32 |
                                 // origin marker} and {flag | firingSkip map (rare path), unused}
  int* tile_counter;             // dynamic tile ids
  long long* frame_first_point;
  int* frame_start_block;
  BatchHeader* hdr;
  int carry_last_az;
  int carry_skip;
  int carry_meta_inited;
  double carry_origin_T[3];
};

// geometry of a decode CTA (overridable for A/B builds): warps, input stages, CTAs per SM
#ifndef VS_DEC_WARPS
#define VS_DEC_WARPS 8
#endif
#ifndef VS_DEC_STAGES
#define VS_DEC_STAGES 3
#endif
#ifndef VS_DEC_CTAS
#define VS_DEC_CTAS 2
#endif
constexpr int kDecTile = VS_DEC_WARPS;  // packets per tile: two per warp pair
constexpr int kDecStagesMax = 3;
constexpr int kDecThreads = 32 * VS_DEC_WARPS;
constexpr int kDecWarps = kDecThreads / 32;
constexpr int kDecPairs = kDecWarps / 2;
constexpr int kDecPktsPerWarp = kDecTile / kDecPairs;  // 2
constexpr int kDecRecs = kDecPktsPerWarp * 6;          // firing blocks per warp and tile
#ifndef VS_DEC_ILP
#define VS_DEC_ILP 2
#endif
constexpr int kDecIlp = VS_DEC_ILP;  // firing blocks per straight-line body (divides 6)
// input stage layout (byte offsets, all multiples of 16)
constexpr int kDRec = 0;                            // 12 BlkRec per packet
constexpr int kDSeg = kDRec + kDecTile * 96;        // PktSeg per packet
constexpr int kDOff = kDSeg + kDecTile * 16;        // u64 point offset per packet (+2 pad)
constexpr int kDPose = kDOff + (kDecTile + 2) * 8;  // pose rows: 12 (18: per-point deskew) doubles per packet
static_assert(kDPose % 16 == 0, "stage sections must be 16-byte aligned");
static_assert(kDecRecs <= 32, "one lane per block record");
// output staging of one pair: its packets' points + 16 elements of phase, column after column.
// HAS_T == false (build with -DVS_DEC_T_DIRECT=1; measured, NOT the default): the t_us column is
// not staged -- without a per-return firing offset (HDL-64 data, no per-point deskew) it is the
// packet's time for every point of the packet and is written straight from registers, which
// takes 12.5 KB of shared memory off the CTA.  With -DVS_DEC_STAGES=2 on top the CTA drops to
// 83 KB and a k_scan CTA of the next batch becomes co-resident with two decode CTAs.  Round-2
// A/B on one B200: the register stores cost the latency-bound kernel more than
// the staging did (1.85 -> 1.96 ms), and the co-resident k_scan takes from k_decode almost what
// it hides (both live on the shared-memory pipe and the issue slots): 2.07 ms per step as is,
// 2.20 with direct t, 2.12 with direct t + co-residency.
#ifndef VS_DEC_T_DIRECT
#define VS_DEC_T_DIRECT 0
#endif
constexpr int kOutCap = kDecPktsPerWarp * 384 + 16;
template <bool HAS_T>
struct OutCols {
  static constexpr int kOX = 0, kOY = 4 * kOutCap, kOZ = 8 * kOutCap, kOT = 12 * kOutCap;
  static constexpr int kOAz = (HAS_T ? 16 : 12) * kOutCap, kODist = kOAz + 2 * kOutCap;
  static constexpr int kOInt = kODist + 2 * kOutCap, kOLas = kOInt + kOutCap;
  static constexpr int kOutBytes = ((HAS_T ? 22 : 18) * kOutCap + 127) & ~127;
};
static_assert(kOutCap % 16 == 0, "column bases must stay 16-byte aligned");

struct DecCtl {
  uint64_t full[kDecStagesMax];
  int released[kDecStagesMax];  // warps that are done with the stage
  // single-pass variant: the scan warp hands a stage over with `ready`, the decode warps hand it
  // back with `empty`; tile ids are dealt dynamically (-1: no tile left)
  int tile_id[kDecStagesMax];
  uint64_t ready[kDecStagesMax];
  uint64_t empty[kDecStagesMax];
  uint64_t masks[kDecStagesMax];  // the decode warps have written the slot bits of the stage's tile
  int grab_seq;                   // next tile sequence number that may deal itself a tile id
};

// dynamic shared memory: [DecCtl | block records | DevConfig (ADJ != 0) | staging | stages]
// DSK: per-point deskew extension (18-double pose rows; one input stage fewer when the DevConfig
// is in shared memory too, to stay at 2 CTAs per SM)
template <int ADJ, int DSK, int FUSED = 0>
struct DecLayout {
  static constexpr int kNumStages = (ADJ != 0 && DSK != 0) ? 2 : VS_DEC_STAGES;
  static constexpr bool kHasT = !(VS_DEC_T_DIRECT && ADJ == 0 && DSK == 0 && FUSED == 0);
  typedef OutCols<kHasT> Cols;
  static constexpr int kOutBytes = Cols::kOutBytes;
  static constexpr int kRowBytes = DSK ? kDeskewRow * 8 : 96;
  static constexpr int kDPkts = kDPose + kDecTile * kRowBytes;  // packet bytes (16-byte granular span)
  static constexpr int kRec = 128;
  static constexpr int kCfg = kRec + kDecWarps * kDecRecs * 32;
  static constexpr int kScr = kCfg + ((ADJ == 0) ? 0 : (((int)sizeof(DevConfig) + 127) & ~127));
  // single-pass variant: slot bits of the tile's blocks, per stage
  static constexpr int kOut = kScr + (FUSED ? ((kDecStagesMax * kDecTile * kBlocks * 4 + 127) & ~127) : 0);
  static constexpr int kStages = kOut + kDecPairs * Cols::kOutBytes;
  static_assert(kDPkts % 16 == 0, "stage sections must be 16-byte aligned");
};
static_assert(sizeof(DecCtl) <= 128, "DecCtl must fit its slot");

__device__ __forceinline__ void lds_v2f64(uint32_t a, double& v0, double& v1) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ void sts_v4(uint32_t a, unsigned x, unsigned y, unsigned z, unsigned w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
// TMA bulk copy shared -> global (16-byte granules), tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
// The same from a converged warp: one elected lane issues it (operands are warp-uniform), which
// lets ptxas feed UBLKCP from uniform registers without a per-value loop.
__device__ __forceinline__ void bulk_s2g_elect(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "@p cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
      "}\n" ::"l"(dst_gmem),
      "r"(src_smem), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void pair_barrier(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
}

// The eight column values of one point into the pair's staging buffer under one predicate
// (no branch in the block body).  a4 / a2 / a1: addresses of the point in the 4-, 2- and
// 1-byte column groups; the columns of a group sit at fixed distances.
template <bool HAS_T>
__device__ __forceinline__ void stage_point(unsigned pred, uint32_t a4, uint32_t a2, uint32_t a1,
                                            float vx, float vy, float vz, unsigned vt, unsigned va,
                                            unsigned vd, unsigned vi, unsigned vl) {
  typedef OutCols<HAS_T> C;
  if (HAS_T) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %11, 0;\n"
        "@p st.shared.f32 [%0], %3;\n"
        "@p st.shared.f32 [%0+%12], %4;\n"
        "@p st.shared.f32 [%0+%13], %5;\n"
        "@p st.shared.u32 [%0+%14], %6;\n"
        "@p st.shared.u16 [%1], %7;\n"
        "@p st.shared.u16 [%1+%15], %8;\n"
        "@p st.shared.u8 [%2], %9;\n"
        "@p st.shared.u8 [%2+%16], %10;\n"
        "}\n" ::"r"(a4),
        "r"(a2), "r"(a1), "f"(vx), "f"(vy), "f"(vz), "r"(vt), "r"(va), "r"(vd), "r"(vi), "r"(vl),
        "r"(pred), "n"(C::kOY - C::kOX), "n"(C::kOZ - C::kOX), "n"(C::kOT - C::kOX),
        "n"(C::kODist - C::kOAz), "n"(C::kOLas - C::kOInt)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %10, 0;\n"
        "@p st.shared.f32 [%0], %3;\n"
        "@p st.shared.f32 [%0+%11], %4;\n"
        "@p st.shared.f32 [%0+%12], %5;\n"
        "@p st.shared.u16 [%1], %6;\n"
        "@p st.shared.u16 [%1+%13], %7;\n"
        "@p st.shared.u8 [%2], %8;\n"
        "@p st.shared.u8 [%2+%14], %9;\n"
        "}\n" ::"r"(a4),
        "r"(a2), "r"(a1), "f"(vx), "f"(vy), "f"(vz), "r"(va), "r"(vd), "r"(vi), "r"(vl),
        "r"(pred), "n"(C::kOY - C::kOX), "n"(C::kOZ - C::kOX), "n"(C::kODist - C::kOAz),
        "n"(C::kOLas - C::kOInt)
        : "memory");
  }
}

// type_defs.h:160-166: row sums left to right, translation last.
__device__ __forceinline__ void rigid(const double* M, double& px, double& py, double& pz) {
  const double qx = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(M[0], px), __dmul_rn(M[1], py)), __dmul_rn(M[2], pz)), M[3]);
  const double qy = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(M[4], px), __dmul_rn(M[5], py)), __dmul_rn(M[6], pz)), M[7]);
  const double qz = __dadd_rn(
      __dadd_rn(__dadd_rn(__dmul_rn(M[8], px), __dmul_rn(M[9], py)), __dmul_rn(M[10], pz)), M[11]);
  px = qx;
  py = qy;
  pz = qz;
}

// sin(x) for |x| < 0.2: odd Taylor polynomial through x^9 (error < 2e-17 relative)
__device__ __forceinline__ double sin_small(double x) {
  const double x2 = x * x;
  double p = __fma_rn(x2, 1.0 / 362880, -1.0 / 5040);
  p = __fma_rn(x2, p, 1.0 / 120);
  p = __fma_rn(x2, p, -1.0 / 6);
  p = __fma_rn(x2, p, 1.0);
  return x * p;
}

}  // namespace vsd

#include "vs_single_pass.cuh"

namespace vsd {

template <int ADJ, int DSK, int FUSED>
__global__ void __launch_bounds__(FUSED ? kFusedThreads : kDecThreads, VS_DEC_CTAS) k_decode(const DecParams p) {
  static_assert(!(FUSED && DSK), "the per-point deskew extension takes the two-pass pipeline");
  typedef DecLayout<ADJ, DSK, FUSED> L;
  typedef typename L::Cols OC;
  constexpr bool kHasT = L::kHasT;
  constexpr int kOutBytes = L::kOutBytes;
  constexpr int kDecStages = L::kNumStages;
  constexpr int kDPkts = L::kDPkts;
  constexpr int kRowD = L::kRowBytes / 8;  // doubles per pose row
  extern __shared__ __align__(128) uint8_t smem_raw[];
  DecCtl& sh = *reinterpret_cast<DecCtl*>(smem_raw);
  const DevConfig& cfg =
      (ADJ == 0) ? *p.cfg : *reinterpret_cast<const DevConfig*>(smem_raw + L::kCfg);
  uint8_t* stage0 = smem_raw + L::kStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  const int par = warp & 1, pair = warp >> 1;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long in_base = reinterpret_cast<long long>(p.pkts);

  // start the copies of tile t into stage s (one thread); t >= n_tiles: nothing left
  auto produce = [&](int s, int t) {
    if (t >= p.n_tiles) {
      mbar_arrive(&sh.full[s]);
      return;
    }
    const long long first = (long long)p.halo + (long long)t * kDecTile;
    const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, first, 0, kDecTile);
    uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
    for (long long a = sp.s1; a < sp.a1; ++a)  // < 16 bytes, last tile of the array only
      st[kDPkts + (a - sp.s0)] = *reinterpret_cast<const uint8_t*>(a);
    const uint32_t bytes = (uint32_t)(sp.s1 - sp.s0);
    const uint32_t rbytes = (uint32_t)sp.npk * 96u;
    const uint32_t sbytes = (uint32_t)sp.npk * 16u;
    const int odd = (int)(first & 1);  // the offsets are copied from an even index
    const uint32_t obytes = (uint32_t)((sp.npk + odd + 1) & ~1) * 8u;
    const uint32_t pbytes = p.pose_valid ? (uint32_t)sp.npk * (uint32_t)L::kRowBytes : 0u;
    fence_proxy_async();
    mbar_expect_tx(&sh.full[s], bytes + rbytes + sbytes + obytes + pbytes);
    bulk_g2s(st + kDRec, p.recs + first * kBlocks, rbytes, &sh.full[s]);
    bulk_g2s(st + kDSeg, p.pkt_seg + first, sbytes, &sh.full[s]);
    bulk_g2s(st + kDOff, p.pkt_off + (first - odd), obytes, &sh.full[s]);
    if (pbytes) bulk_g2s(st + kDPose, p.pose_mat + first * kRowD, pbytes, &sh.full[s]);
    if (bytes) bulk_g2s(st + kDPkts, reinterpret_cast<const void*>(sp.s0), bytes, &sh.full[s]);
  };

  if (ADJ != 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.cfg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw + L::kCfg);
    for (int i = tid; i < (int)(sizeof(DevConfig) / 4); i += (int)blockDim.x) dst[i] = __ldg(&src[i]);
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kDecStages; ++s) {
      mbar_init(&sh.full[s], 1);
      sh.released[s] = 0;
      if (FUSED) {
        mbar_init(&sh.ready[s], 1);
        mbar_init(&sh.empty[s], kDecWarps);
        mbar_init(&sh.masks[s], kDecWarps);
        sh.tile_id[s] = -1;
      }
    }
    sh.grab_seq = 0;
    fence_mbar_init();
    if (!FUSED) {
#pragma unroll 1
      for (int s = 0; s < kDecStages; ++s) produce(s, (int)blockIdx.x + s * (int)gridDim.x);
    }
  }
  __syncthreads();
  if (FUSED && warp >= kDecWarps) {
    fused_scan_role<ADJ, L>(p, sh, smem_raw, lane, warp - kDecWarps);
    return;
  }

  const bool pose_valid = p.pose_valid != 0;
  CalRow cal;
  int cal_bank = -1;
  unsigned cnt = 0;  // emitted points of (cnt_frame, cnt_bank) seen by this lane
  int cnt_frame = -1, cnt_bank = 0;
  auto flush_counts = [&]() {
    if (cnt) {
      int l = lane + cnt_bank;
      if (ADJ == 2 && l >= 16) l -= 16;
      if (cnt_frame < p.frame_cap)
        atomicAdd(&p.frame_laser_counts[(long long)cnt_frame * kMaxLasers + l], cnt);
      cnt = 0;
    }
  };
  const uint32_t smem_a = smem_u32(smem_raw);
  const uint32_t rec_a = smem_a + L::kRec + (uint32_t)warp * (kDecRecs * 32);
  const uint32_t out_a = smem_a + L::kOut + (uint32_t)pair * kOutBytes;
  const uint32_t stage_a0 = smem_a + L::kStages;
  const unsigned stride = (unsigned)p.stride;

  // one firing block: everything after the shared-memory loads
  auto finish_block = [&](const int4 r, const double sn, const double cs, const unsigned d0,
                          const unsigned d1, const unsigned inten, const int j, const int laser_id,
                          const unsigned tpk, const int azdiff, const double* M) {
    const unsigned m = (unsigned)r.x;
    const unsigned dist = d0 | (d1 << 8);
    unsigned az;
    double sA, cA;
    if (ADJ == 0) {
      az = (unsigned)r.z & 0xffffu;
      sA = sn;
      cA = cs;
    } else {
      az = adjusted_azimuth<ADJ>(cfg, (unsigned)r.z & 0xffffu, azdiff, j, lane);
      sA = __ldg(&p.lut_sin[az]);
      cA = __ldg(&p.lut_cos[az]);
    }
    double px, py, pz;
    sensor_point(cal, sA, cA, dist, px, py, pz);
    // firing offset of this return (us): defines the t_us column and the per-point pose
    unsigned fire = 0;
    if (ADJ != 0)
      fire = cfg.tadj[j][lane];
    else if (DSK)
      fire = __ldg(&p.cfg->tadj[j][lane]);
    if (pose_valid) {
      if (DSK) {
        // pose at t_packet + fire inside the packet's bracket: slerp + lerp, already re-based to
        // the frame origin by k_pose (semantics: oracle/deskew_port.py)
        // (an extension with no reference arithmetic to reproduce: fused multiply-adds throughout)
        const double f = (double)fire;
        const double qw = __fma_rn(f, M[4], M[0]);
        const double qx = __fma_rn(f, M[5], M[1]);
        const double qy = __fma_rn(f, M[6], M[2]);
        const double qz = __fma_rn(f, M[7], M[3]);
        // p + w t + v x t with t = 2 v x p
        double tx = __fma_rn(qy, pz, -(qz * py));
        double ty = __fma_rn(qz, px, -(qx * pz));
        double tz = __fma_rn(qx, py, -(qy * px));
        tx += tx;
        ty += ty;
        tz += tz;
        const double ox = __fma_rn(qy, tz, __fma_rn(-qz, ty, __fma_rn(qw, tx, px)));
        const double oy = __fma_rn(qz, tx, __fma_rn(-qx, tz, __fma_rn(qw, ty, py)));
        const double oz = __fma_rn(qx, ty, __fma_rn(-qy, tx, __fma_rn(qw, tz, pz)));
        px = ox + __fma_rn(M[11], f, M[8]);
        py = oy + __fma_rn(M[12], f, M[9]);
        pz = oz + __fma_rn(M[13], f, M[10]);
      } else {
        rigid(M, px, py, pz);
      }
    }
    const unsigned pred = (m >> lane) & 1u;
    const unsigned o = (unsigned)r.y + __popc(m & lt_mask);  // position in the pair's staging
    stage_point<kHasT>(pred, out_a + OC::kOX + 4u * o, out_a + OC::kOAz + 2u * o, out_a + OC::kOInt + o,
                       (float)px, (float)py, (float)pz, tpk + fire, az, dist, inten, (unsigned)laser_id);
    cnt += pred;
  };

  const int lp0 = kDecPktsPerWarp * pair;  // first packet of the pair inside a tile
  // single-pass variant: slot bits "distance != 0 and laser selected" of this warp's 12 blocks of
  // the tile with sequence number `seq` (one ballot per block), for the scan warp's point counts
  const unsigned sel_lo = FUSED ? cfg.sel_lo : 0u, sel_hi = FUSED ? cfg.sel_hi : 0u;
  auto mask_pass = [&](int seq) {
    const int s1 = seq % kDecStages;
    mbar_wait(&sh.full[s1], (uint32_t)(seq / kDecStages) & 1u);
    const int t1 = *reinterpret_cast<volatile int*>(&sh.tile_id[s1]);
    if (t1 < 0) return;
    const int first1 = t1 * kDecTile;
    const int npk1 = min(kDecTile, p.n - first1);
    const uint32_t pkt1 = stage_a0 + (uint32_t)s1 * (uint32_t)p.stage_bytes + kDPkts +
                          (((unsigned)in_base + (unsigned)first1 * stride) & 15u);
    const uint32_t nz1 = smem_a + L::kScr + (uint32_t)s1 * (kDecTile * kBlocks * 4);
#pragma unroll
    for (int k = 0; k < kDecPktsPerWarp; ++k) {
      const int lp = lp0 + k;
      if (lp < npk1) {
        const uint32_t b0 = pkt1 + (unsigned)lp * stride + 100u * (unsigned)par;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const uint32_t a = b0 + 200u * (unsigned)i;
          const unsigned d = lds_u8(a + 4u + 3u * (unsigned)lane) | lds_u8(a + 5u + 3u * (unsigned)lane);
          const unsigned bits =
              __ballot_sync(0xffffffffu, d != 0u) & ((lds_u16(a) != 0xeeffu) ? sel_hi : sel_lo);
          if (lane == 0) sts_u32(nz1 + 4u * (unsigned)(lp * kBlocks + par + 2 * i), bits);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sh.masks[s1]);
  };
  VS_PROF_DECL
  if (FUSED) {
    mask_pass(0);
    mask_pass(1);
  }
  bool pf_ok = false;  // pf_sn / pf_cs hold the LUT values of this lane's block of the next tile
  double pf_sn = 0.0, pf_cs = 0.0;
#pragma unroll 1
  for (int it = 0;; ++it) {
    const int s = it % kDecStages;
    int tile;
    if (FUSED) {
      VS_PROF_T0();
      mbar_wait(&sh.ready[s], (uint32_t)(it / kDecStages) & 1u);
      VS_PROF_ADD(6);
      tile = *reinterpret_cast<volatile int*>(&sh.tile_id[s]);
      if (tile < 0) break;

    } else {
      tile = (int)blockIdx.x + it * (int)gridDim.x;
      if (tile >= p.n_tiles) break;
      mbar_wait(&sh.full[s], (uint32_t)(it / kDecStages) & 1u);
    }
    // consumer view of the tile: packet count and where its first byte sits in the 16-byte
    // granular span the producer copied (32-bit arithmetic; the producer did the full span)
    const int first = FUSED ? tile * kDecTile : p.halo + tile * kDecTile;
    const int npk = min(kDecTile, p.n - first);
    const uint32_t st_a = stage_a0 + (uint32_t)s * (uint32_t)p.stage_bytes;
    const uint32_t brec_a = st_a + kDRec;
    const uint32_t seg_a = st_a + kDSeg;
    const uint32_t off_a = st_a + kDOff + (FUSED ? 0u : 8u * (unsigned)(first & 1));
    const uint32_t pose_a = st_a + kDPose;
    const uint32_t pkt_a = st_a + kDPkts + (((unsigned)in_base + (unsigned)first * stride) & 15u);
    // global index of the pair's first point and the number of points of its packets
    unsigned long long P0 = 0;
    unsigned tot = 0;
    if (lp0 < npk) {
      P0 = lds_u64(off_a + 8u * (unsigned)lp0);
#pragma unroll
      for (int k = 0; k < kDecPktsPerWarp; ++k)
        if (lp0 + k < npk) tot += (lds_u32(seg_a + 16u * (unsigned)(lp0 + k) + 12u) >> 12) & 0x1ffu;
    }
    const unsigned ph = (unsigned)P0 & 15u;

    // ---- expand the block records of this warp (lane == block) -----------------------------
    if (lane < kDecRecs) {
      const int k = lane / 6, i = lane - 6 * k;
      const int lp = lp0 + k, j = par + 2 * i;
      unsigned m = 0, off = 0, azf = 0;
      int frame = 0;
      double sn = 0.0, cs = 0.0;
      if (lp < npk) {
        const unsigned long long br = lds_u64(brec_a + 8u * (unsigned)(lp * kBlocks + j));
        m = (unsigned)br;
        const unsigned info = (unsigned)(br >> 32);
        const unsigned wb = (info >> 26) & 15u;
        frame = (int)lds_u32(seg_a + 16u * (unsigned)lp + 4u) + (int)wb;
        off = ph + (lds_u32(off_a + 8u * (unsigned)lp) - (unsigned)P0) + ((info >> 16) & 0x1ffu);
        // bits 0-15 azimuth, 16 laser bank, 17 zero translation (offline, at/after a wrap)
        azf = (info & 0xffffu) | (((info >> 25) & 1u) << 16) |
              ((p.mode == 1 && wb != 0u) ? (1u << 17) : 0u);
        if (ADJ == 0) {
          if (pf_ok) {
            sn = pf_sn;
            cs = pf_cs;
          } else {
            sn = __ldg(&p.lut_sin[info & 0xffffu]);
            cs = __ldg(&p.lut_cos[info & 0xffffu]);
          }
        }
      }
      sts_v4(rec_a + 32u * (unsigned)lane, m, off, azf, (unsigned)frame);
      sts_v2f64(rec_a + 32u * (unsigned)lane + 16u, sn, cs);
    }
    __syncwarp();
    // LUT gather of the NEXT tile's blocks, issued now so that its L2 round trip hides behind
    // this tile's math; when that tile has not landed yet (a warp running ahead of its CTA), a
    // second attempt follows the block loops
    auto prefetch_lut = [&]() {
      if (!(ADJ == 0 && (FUSED || tile + (int)gridDim.x < p.n_tiles))) return;
      const int sn_ = (it + 1) % kDecStages;
      int first_n = first + (int)gridDim.x * kDecTile;
      if (FUSED) {
        pf_ok = __shfl_sync(0xffffffffu,
                            (int)mbar_test(&sh.ready[sn_], (uint32_t)((it + 1) / kDecStages) & 1u), 0) != 0;
        if (pf_ok) {
          const int tn = *reinterpret_cast<volatile int*>(&sh.tile_id[sn_]);
          first_n = tn * kDecTile;
          pf_ok = tn >= 0;
        }
      } else {
        pf_ok = __shfl_sync(0xffffffffu,
                            (int)mbar_test(&sh.full[sn_], (uint32_t)((it + 1) / kDecStages) & 1u), 0) != 0;
      }
      if (pf_ok && lane < kDecRecs) {
        const int k = lane / 6, i = lane - 6 * k;
        const int lp = lp0 + k, j = par + 2 * i;
        if (first_n + lp < p.n) {
          const unsigned info = lds_u32(stage_a0 + (uint32_t)sn_ * (uint32_t)p.stage_bytes + kDRec +
                                        8u * (unsigned)(lp * kBlocks + j) + 4u);
          pf_sn = __ldg(&p.lut_sin[info & 0xffffu]);
          pf_cs = __ldg(&p.lut_cos[info & 0xffffu]);
        }
      }
    };
    pf_ok = false;
    prefetch_lut();
    // staging is free once both warps' bulk stores of the previous tile have read it
    if (it > 0) {
      if (lane == 0) bulk_wait_read0();
      pair_barrier(pair);
    }

#pragma unroll 1
    for (int k = 0; k < kDecPktsPerWarp; ++k) {
      const int lp = lp0 + k;
      if (lp >= npk) break;
      const PktSeg seg = lds_v4(seg_a + 16u * (unsigned)lp);
      const unsigned wrapmask = ((unsigned)seg.x >> 4) & 0xfffu;
      const int azdiff = (seg.x >> 16) & 0xffff;
      const unsigned um = (unsigned)seg.w & 0xfffu;
      const unsigned tpk = (unsigned)seg.z;
      // lane's first return of block `par`: 3 bytes per return after the 4-byte block header
      const uint32_t blk_a = pkt_a + (unsigned)lp * stride + 100u * (unsigned)par + 4u + 3u * (unsigned)lane;
      const uint32_t rk_a = rec_a + 32u * 6u * (unsigned)k;
      double M[kRowD];  // pose row of this packet ([L | t], or the deskew row), warp-uniform
      if (pose_valid) {
#pragma unroll
        for (int q = 0; q < kRowD / 2; ++q)
          lds_v2f64(pose_a + (unsigned)L::kRowBytes * (unsigned)lp + 16u * (unsigned)q, M[2 * q], M[2 * q + 1]);
      }
      const unsigned pm = par ? 0xaaau : 0x555u;  // the blocks of this warp
      const unsigned ub = um & pm;
      if (wrapmask == 0u && (ub == 0u || ub == pm)) {
        // fast path (every packet of a real stream but the ~0.3 % that hold a wrap): one
        // frame, one laser bank for all blocks of this warp
        const int bank = ub ? 32 : 0;
        if (seg.y != cnt_frame || bank != cnt_bank) {
          flush_counts();
          cnt_frame = seg.y;
          cnt_bank = bank;
        }
        if (bank != cal_bank) {
          load_cal(cfg, lane + bank, cal);
          cal_bank = bank;
        }
        int laser_id = lane + bank;
        if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
#pragma unroll
        for (int i = 0; i < 6; i += kDecIlp) {
          // kDecIlp blocks as one straight-line body: loads first, then the FP64 chains
          int4 r[kDecIlp];
          double sn[kDecIlp], cs[kDecIlp];
          unsigned d0[kDecIlp], d1[kDecIlp], in[kDecIlp];
#pragma unroll
          for (int u = 0; u < kDecIlp; ++u) {
            r[u] = lds_v4(rk_a + 32u * (unsigned)(i + u));
            sn[u] = 0.0;
            cs[u] = 0.0;
            if (ADJ == 0) lds_v2f64(rk_a + 32u * (unsigned)(i + u) + 16u, sn[u], cs[u]);
          }
#pragma unroll
          for (int u = 0; u < kDecIlp; ++u) {
            const uint32_t a0 = blk_a + 200u * (unsigned)(i + u);
            d0[u] = lds_u8(a0);
            d1[u] = lds_u8(a0 + 1u);
            in[u] = lds_u8(a0 + 2u);
          }
#pragma unroll
          for (int u = 0; u < kDecIlp; ++u)
            finish_block(r[u], sn[u], cs[u], d0[u], d1[u], in[u], par + 2 * (i + u), laser_id, tpk,
                         azdiff, M);
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {
          const int4 r = lds_v4(rk_a + 32u * (unsigned)i);
          if (r.x == 0) continue;
          double sn = 0.0, cs = 0.0;
          if (ADJ == 0) lds_v2f64(rk_a + 32u * (unsigned)i + 16u, sn, cs);
          const int bank = (((unsigned)r.z >> 16) & 1u) ? 32 : 0;
          if (((unsigned)r.z >> 17) & 1u) {
            // offline: blocks at/after the packet's first wrap start a frame whose origin is
            // this very packet -> zero translation
            if (!DSK) {
              M[3] = 0.0;
              M[7] = 0.0;
              M[11] = 0.0;
            }
          }
          if (r.w != cnt_frame || bank != cnt_bank) {
            flush_counts();
            cnt_frame = r.w;
            cnt_bank = bank;
          }
          if (bank != cal_bank) {
            load_cal(cfg, lane + bank, cal);
            cal_bank = bank;
          }
          int laser_id = lane + bank;
          if (ADJ == 2 && laser_id >= 16) laser_id -= 16;
          const uint32_t a0 = blk_a + 200u * (unsigned)i;
          const unsigned d0 = lds_u8(a0), d1 = lds_u8(a0 + 1u), in = lds_u8(a0 + 2u);
          finish_block(r, sn, cs, d0, d1, in, par + 2 * i, laser_id, tpk, azdiff, M);
        }
      }
    }
    if (!kHasT) {
      // t_us of a packet without per-return firing offsets: one value for all its points, written
      // from registers in coalesced runs (warp `par` of the pair takes packet lp0 + par)
#pragma unroll
      for (int k = 0; k < kDecPktsPerWarp; ++k) {
        const int lp = lp0 + k;
        if ((k & 1) == par && lp < npk) {
          const unsigned cntp = (lds_u32(seg_a + 16u * (unsigned)lp + 12u) >> 12) & 0x1ffu;
          const unsigned tval = lds_u32(seg_a + 16u * (unsigned)lp + 8u);
          // every store instruction covers one 128-byte line of the column: whole sectors but
          // for the packet's first and last line
          const unsigned long long o64 = lds_u64(off_a + 8u * (unsigned)lp);
          uint32_t* dst = p.t_us + o64;
          const int skew = (int)((unsigned)o64 & 31u);
          for (int i = lane - skew; i < (int)cntp; i += 32)
            if (i >= 0) stg_u32(dst + i, tval);
        }
      }
    }
    if (!pf_ok) prefetch_lut();
    __syncwarp();  // every lane is done with the stage and with the records
    if (lane == 0) {
      if (FUSED) {
        mbar_arrive(&sh.empty[s]);
      } else {
        const int old = atomicAdd(&sh.released[s], 1);
        if (old == kDecWarps - 1) {
          sh.released[s] = 0;
          produce(s, tile + kDecStages * (int)gridDim.x);
        }
      }
    }
    if (FUSED) {
      // slot bits of the tile two sequence numbers ahead (landed by now): its scan gets a whole
      // tile of slack
      VS_PROF_ADD(8);
      mask_pass(it + 2);
      VS_PROF_ADD(7);
    }

    // ---- the pair's points leave: one bulk store per column + scalar head / tail -----------
    fence_proxy_async();  // this thread's staging writes -> async proxy
    pair_barrier(pair);
    {
      const unsigned long long Pu =
          ((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(P0 >> 32), 0) << 32) |
          __shfl_sync(0xffffffffu, (unsigned)P0, 0);
      const unsigned totu = __shfl_sync(0xffffffffu, tot, 0);
      if (totu != 0u) {
        const unsigned phu = (unsigned)Pu & 15u;
        const unsigned s1 = phu + totu;                   // staging range [phu, s1)
        const unsigned b0 = (phu + 15u) & ~15u, b1 = s1 & ~15u;  // whole granules [b0, b1)
        const unsigned long long g0 = Pu - phu;           // global index of staging element 0
        if (b1 > b0) {
          // warp `par` stores two 4-byte columns, one 2-byte and one 1-byte column (the whole
          // warp is converged here; one elected lane issues each copy)
          const unsigned nb = b1 - b0;
          const unsigned long long gb = g0 + b0;
          if (par == 0) {
            bulk_s2g_elect(p.x + gb, out_a + OC::kOX + 4u * b0, 4u * nb);
            bulk_s2g_elect(p.y + gb, out_a + OC::kOY + 4u * b0, 4u * nb);
            bulk_s2g_elect(p.azimuth + gb, out_a + OC::kOAz + 2u * b0, 2u * nb);
            bulk_s2g_elect(p.intensity + gb, out_a + OC::kOInt + b0, nb);
          } else {
            bulk_s2g_elect(p.z + gb, out_a + OC::kOZ + 4u * b0, 4u * nb);
            if (kHasT) bulk_s2g_elect(p.t_us + gb, out_a + OC::kOT + 4u * b0, 4u * nb);
            bulk_s2g_elect(p.distance + gb, out_a + OC::kODist + 2u * b0, 2u * nb);
            bulk_s2g_elect(p.laser + gb, out_a + OC::kOLas + b0, nb);
          }
          bulk_commit();
        }
        // head (warp 0) / tail (warp 1): the elements outside the whole granules, all inside
        // the first / last 16-element window of the range
        if (lane < 16) {
          const unsigned wf = phu & ~15u, wl = (s1 - 1u) & ~15u;
          const unsigned e = (par ? wl : wf) + (unsigned)lane;
          if (e >= phu && e < s1 && (e < b0 || e >= b1) && !(par && wl == wf)) {
            const unsigned long long g = g0 + e;
            p.x[g] = __uint_as_float(lds_u32(out_a + OC::kOX + 4u * e));
            p.y[g] = __uint_as_float(lds_u32(out_a + OC::kOY + 4u * e));
            p.z[g] = __uint_as_float(lds_u32(out_a + OC::kOZ + 4u * e));
            if (kHasT) p.t_us[g] = lds_u32(out_a + OC::kOT + 4u * e);
            p.azimuth[g] = (uint16_t)lds_u16(out_a + OC::kOAz + 2u * e);
            p.distance[g] = (uint16_t)lds_u16(out_a + OC::kODist + 2u * e);
            p.intensity[g] = (uint8_t)lds_u8(out_a + OC::kOInt + e);
            p.laser[g] = (uint8_t)lds_u8(out_a + OC::kOLas + e);
          }
        }
      }
    }
    VS_PROF_ADD(8);
  }
  flush_counts();
  VS_PROF_FLUSH(lane);
  if (lane == 0) bulk_wait0();  // shared memory must outlive the bulk stores that read it
}

// =========================================================================================
// SURVEY.md 8f row N3 -- the online front end's per-packet / per-record arithmetic in data-
// parallel form.
//
// k_gps_gather / k_gps_times: TimeSolver::calcTimestamp(uint32_t microsecToHour)
// (TimeSolver.cxx:34-49) for a packet array.  The reference adds one hour to hdlHourTime every
// time the sensor's microseconds-past-the-hour field steps backwards; over an array that is an
// inclusive scan of a 1-bit flag (decoupled look-back across 1024-packet tiles):
//   t[i] = base + 3600 s * #{k <= i : gps[k-1] > gps[k]} + gps[i]
// =========================================================================================
__global__ void k_gps_gather(const uint8_t* __restrict__ pkts, long long stride, int n,
                             uint32_t* __restrict__ gps) {
  const int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= n) return;
  // gpsTimestamp sits at byte 1200 of the payload (HDLSource.cxx:216); packets are 2-byte aligned
  const unsigned short* h = reinterpret_cast<const unsigned short*>(pkts + (long long)P * stride + 1200);
  gps[P] = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
}

struct GpsParams {
  const uint32_t* gps;
  int n;
  uint32_t last_report;  // lastHdlReport entering the array (0 before the first packet)
  long long base_us;     // hdlHourTime + hdlOffset
  long long* t_out;
  unsigned long long* st;  // look-back words, one per tile, zeroed
  int* tile_counter;       // zeroed
  unsigned long long* total_wraps;
};
constexpr int kGpsThreads = 256;
constexpr int kGpsItems = 4;

__global__ void __launch_bounds__(kGpsThreads) k_gps_times(const GpsParams p) {
  __shared__ unsigned s_warp[kGpsThreads / 32];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  const int i0 = (tile * kGpsThreads + tid) * kGpsItems;
  uint32_t g[kGpsItems];
  unsigned f[kGpsItems];
  uint32_t prev = (i0 == 0) ? p.last_report : ((i0 - 1 < p.n) ? __ldg(&p.gps[i0 - 1]) : 0u);
  unsigned mine = 0;
#pragma unroll
  for (int k = 0; k < kGpsItems; ++k) {
    g[k] = (i0 + k < p.n) ? __ldg(&p.gps[i0 + k]) : prev;
    f[k] = (i0 + k < p.n && prev > g[k]) ? 1u : 0u;  // TimeSolver.cxx:43
    mine += f[k];
    prev = g[k];
  }
  unsigned inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = (lane < kGpsThreads / 32) ? s_warp[lane] : 0u;
#pragma unroll
    for (int o = 1; o < kGpsThreads / 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    if (lane < kGpsThreads / 32) s_warp[lane] = w;
    const unsigned long long ex =
        lookback_exclusive<SumTraits>(p.st, tile, (unsigned long long)__shfl_sync(0xffffffffu, w, kGpsThreads / 32 - 1));
    if (lane == 0) s_prefix = ex;
  }
  __syncthreads();
  unsigned long long hours = s_prefix + (warp > 0 ? s_warp[warp - 1] : 0u) + (inc - mine);
#pragma unroll
  for (int k = 0; k < kGpsItems; ++k) {
    hours += f[k];
    if (i0 + k < p.n) {
      p.t_out[i0 + k] = p.base_us + (long long)hours * 3600000000ll + (long long)g[k];
      if (i0 + k == p.n - 1) *p.total_wraps = hours;
    }
  }
}

// =========================================================================================
// k_ins_pose: INSSource PacketConsumer::calcTransform (INSSource.cxx:300-326) and
// TimeSolver::calcTimestamp(InsPVA const*) (TimeSolver.cxx:20-33) for an array of NovAtel
// INSPVA records: geodetic -> ECEF (llh2xyz, CoordiTran.cpp:51-80) -> local ENU about the
// origin (xyz2enu, :152-187; the origin's rotation is evaluated once on the host).
// =========================================================================================
struct InsPva {  // type_defs.h:39-58, natural alignment (104 bytes)
  uint16_t message_id;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  uint32_t pad0;
  double seconds_pos;
  double LLH[3];
  double V[3];
  double Eulr[3];
  int32_t ins_status;
  int32_t pad1;
};
struct InsParams {
  const InsPva* recs;
  int n;
  double org[3];   // ECEF origin
  double R[9];     // ENU rotation at the origin (row-major)
  const long long* arrival_us;  // the local clock at reception of each record
  long long* t_out;
  double* trv_out;  // n x 9
};

__global__ void k_ins_pose(const InsParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const InsPva& d = p.recs[i];
  // TO_RADIUS(x) = x * M_PI / 180 (type_defs.h:25)
  const double phi = d.LLH[0] * 3.14159265358979323846 / 180;
  const double lambda = d.LLH[1] * 3.14159265358979323846 / 180;
  const double h = d.LLH[2];
  const double a = 6378137.0000, b = 6356752.3142;
  const double e = sqrt(1 - (b / a) * (b / a));
  const double sinphi = sin(phi), cosphi = cos(phi);
  const double coslam = cos(lambda), sinlam = sin(lambda);
  const double tan2phi = (tan(phi)) * (tan(phi));
  const double tmp = 1 - e * e;
  const double tmpden = sqrt(1 + tmp * tan2phi);
  const double x = (a * coslam) / tmpden + h * coslam * cosphi;
  const double y = (a * sinlam) / tmpden + h * sinlam * cosphi;
  const double tmp2 = sqrt(1 - e * e * sinphi * sinphi);
  const double z = (a * tmp * sinphi) / tmp2 + h * sinphi;
  const double dif[3] = {x - p.org[0], y - p.org[1], z - p.org[2]};
  double enu[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    enu[0] = enu[0] + p.R[0 + k] * dif[k];
    enu[1] = enu[1] + p.R[3 + k] * dif[k];
    enu[2] = enu[2] + p.R[6 + k] * dif[k];
  }
  double* o = p.trv_out + (long long)i * 9;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[k] = enu[k];
    o[3 + k] = d.Eulr[k];
    o[6 + k] = d.V[k];
  }
  // now + (time of pose - time of packet send); time_duration truncates the doubles to ticks
  const long long hour_us = 3600000000ll;
  const long long sent = (long long)((int)d.week_number * 168) * hour_us + (long long)((double)d.milliseconds * 1e3);
  const long long pose = (long long)((int)d.week_number_pos * 168) * hour_us + (long long)(d.seconds_pos * 1e6);
  p.t_out[i] = p.arrival_us[i] + (pose - sent);
}

// =========================================================================================
// k_reset: everything a batch starts from, in one launch -- the zeroed scan / frame-count state,
// the "unset" (all ones) rows of the first_point / start_block tables and the batch header's
// initial values (three stream operations before; a rotation-sized batch is a chain of
// dependent operations and pays for each link).
// =========================================================================================
struct ResetParams {
  unsigned* zero;        // 4-byte aligned
  long long zero_words;
  unsigned* ones;
  long long ones_words;
  BatchHeader* hdr;
};
__global__ void k_reset(const ResetParams p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < p.zero_words; i += stride) p.zero[i] = 0u;
  for (long long i = i0; i < p.ones_words; i += stride) p.ones[i] = 0xffffffffu;
  if (i0 == 0) {
    BatchHeader h;
    memset(&h, 0, sizeof(h));
    h.first_upper_block = LLONG_MAX;
    h.first_const_pkt = INT_MAX;
    h.origin_at_halo = -1;
    h.last_origin_packet = -1;
    *p.hdr = h;
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
  // rotation-sized batches: the batch header and the first kEagerRows rows of every frame table
  // packed into one block, so that ONE device -> host copy brings the whole index back
  EagerBlock* eager;        // null: the tables are copied one by one
  const BatchHeader* hdr;
  const long long* frame_first_point;
  const unsigned* frame_laser_counts;
};

__global__ void k_frames(const FrameParams p) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  int mp = -2, sk = -1;
  long long mt = 0;
  if (f >= 1 && f < p.n_frames) {  // frame 0 comes from the carry / first packet
    const int sb = p.frame_start_block[f];
    if (sb >= 0) {  // (< 0: wrap inside the halo, not decoded)
      const int P = sb / 12, j = sb % 12;
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
    }
    p.frame_meta_packet[f] = mp;
    p.frame_meta_time[f] = mt;
    p.frame_skips[f] = sk;
  }
  if (p.eager && blockIdx.x == 0) {
    // the first CTA packs the block: one row per thread, the header and the per-laser counts
    // word by word across the CTA (coalesced; a rotation-sized batch waits for this kernel)
    EagerBlock& e = *p.eager;
    static_assert(sizeof(BatchHeader) % 4 == 0, "header copied as 32-bit words");
    const unsigned* hs = reinterpret_cast<const unsigned*>(p.hdr);  // final: the batch's last kernel
    unsigned* hd = reinterpret_cast<unsigned*>(&e.hdr);
    for (int i = threadIdx.x; i < (int)(sizeof(BatchHeader) / 4); i += blockDim.x) hd[i] = hs[i];
    unsigned* cd = &e.counts[0][0];
    for (int i = threadIdx.x; i < kEagerRows * kMaxLasers; i += blockDim.x)
      cd[i] = (i / kMaxLasers) < p.n_frames ? p.frame_laser_counts[i] : 0u;
    if (f < kEagerRows) {
      const bool in = f < p.n_frames;
      e.first[f] = in ? p.frame_first_point[f] : -1ll;
      e.start[f] = in ? p.frame_start_block[f] : -1;
      e.meta_pkt[f] = mp;
      e.meta_time[f] = mt;
      e.skips[f] = sk;
    }
  }
}

// =========================================================================================
// k_table_rows: a batch's frame table as the exchange rows of the multi-GPU stitch
// (VS_FRAME_ROW_COLS int64 per frame, packet indices global), built on the device behind the
// batch's kernels -- what finish_batch + vs_frame_table_rows do on the host, so that the
// all-gather between ranks starts from HBM without a host pass.  Row 0 = {n_rows, 0, ...};
// more frames than cap_rows: row 0 = {-n_rows, ...} and no table.
// =========================================================================================
struct TableRowsParams {
  const BatchHeader* hdr;
  const long long* frame_first_point;
  const int* frame_start_block;
  const int* frame_meta_packet;
  const long long* frame_meta_time;
  const int* frame_skips;
  const PktSeg* pkt_seg;
  const long long* pkt_time;
  long long* rows;           // (cap_rows + 1) x kRowCols
  long long cap_rows;
  long long base;            // global index of packet 0 of the submitted array
  long long carry_timestamp; // carry-in: the open frame's meta
  int carry_meta_inited, carry_skips, carry_firing_skip, carry_is_hdl64;
  int halo, mode, index_only, rank;
};
constexpr int kRowCols = 10;  // VS_FRAME_ROW_COLS
constexpr long long kTimeNone = (long long)0x8000000000000000ull;  // VS_TIME_NONE

__global__ void k_table_rows(const TableRowsParams p) {
  const BatchHeader& h = *p.hdr;
  const int W = h.total_wraps;
  const int f_lo = p.halo > 0 ? h.frame_at_halo : 0;
  const int n_frames = W - f_lo + 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 == 0) {
    long long* r = p.rows;
    r[0] = n_frames <= p.cap_rows ? (long long)n_frames : -(long long)n_frames;
    for (int k = 1; k < kRowCols; ++k) r[k] = 0;
  }
  if (n_frames > p.cap_rows) return;
  const long long total_points = p.index_only ? 0ll : h.total_points;
  for (long long i = i0; i < n_frames; i += stride) {
    const int f = f_lo + (int)i;
    long long first = p.frame_first_point[f];
    int sb = p.frame_start_block[f];
    if (i == 0) {
      if (first < 0) first = 0;
      if (p.halo > 0 || sb < 0) sb = -1;
    }
    const bool closed = i + 1 < n_frames;
    const long long next_first = closed ? p.frame_first_point[f + 1] : total_points;
    int order = 0;
    if (closed) order = (p.carry_is_hdl64 || h.first_upper_block <= (long long)p.frame_start_block[f + 1]) ? 1 : 0;
    int mp, sk;
    long long ts;
    if (i == 0 && sb < 0) {
      if (p.halo == 0 && p.carry_meta_inited) {
        mp = -1;
        ts = p.carry_timestamp;
        sk = p.carry_skips;
      } else {
        // meta from the origin packet of the first decoded packet (vs_wait patches these in)
        mp = p.halo > 0 ? h.origin_at_halo : 0;
        sk = p.halo > 0 ? -1 : p.carry_firing_skip;
        ts = kTimeNone;
        if (mp >= 0) {
          ts = p.pkt_time[mp];
          if (p.halo > 0) {
            const int sx = p.pkt_seg[mp].x;
            if (p.mode == 1) {
              const unsigned wm = (unsigned)(sx >> 4) & 0xfffu;
              sk = wm ? 31 - __clz(wm) : 0;
            } else {
              sk = sx & 15;
            }
          }
        }
      }
    } else {
      mp = p.frame_meta_packet[f];
      if (mp >= 0) {
        ts = p.frame_meta_time[f];
        sk = p.frame_skips[f];
      } else {
        mp = (mp == -3) ? -2 : mp;
        ts = kTimeNone;
        sk = -1;
      }
    }
    long long* r = p.rows + (i + 1) * kRowCols;
    r[0] = p.index_only ? 0ll : next_first - first;
    r[1] = first;
    r[2] = sb < 0 ? -1ll : (long long)(sb / 12) + p.base;
    r[3] = sb < 0 ? -1ll : (long long)(sb % 12);
    r[4] = ts;
    r[5] = sk;
    r[6] = closed ? 1 : 0;
    r[7] = order;
    r[8] = mp >= 0 ? (long long)mp + p.base : (long long)mp;
    r[9] = p.rank;
  }
}

}  // namespace vsd
