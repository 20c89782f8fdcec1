// vs_layout.cuh -- the reference's HDLFrame layout, built on the device.
//
// The reference appends every emitted point to per-laser lists of the open frame
//     currentFrame->points[laserId]->points.push_back(p);       (HDLParser.cxx:733-743)
//     currentFrame->pointsMeta[laserId]->push_back(m);          (HDLParser.cxx:745-751)
// and permutes the rows by HDL64BeamLUT when a frame is closed on HDL-64 data (:880-893).
// k_decode emits the same points in the order the reference pushes them (packet, block, return
// slot); the two kernels here turn that stream order into the reference's layout so that a
// caller builds each points[row] from one contiguous run of ready-made PointXYZI / PointMeta
// records instead of scattering point by point on the CPU:
//
//   k_layout_rows  per frame: where each laser's row starts (rows in the order the closed frame
//                  holds them; the first frame leaves room for what earlier batches carried in)
//   k_layout       per point: its slot = row start + rank among the frame's points of the same
//                  laser.  The rank is a 64-counter prefix count over the frame's firing blocks
//                  (lane == return slot keeps the counters of its two laser banks in registers);
//                  across tiles it is a decoupled look-back over one 64-bit word per (tile, lane)
//                  that restarts at every frame boundary.  A tile (8 packets, one warp each) is
//                  transposed through shared memory so that every laser's run is written as one
//                  contiguous, coalesced copy.
//
// HBM-bound byte shuffling: per point 18 B read (x y z intensity azimuth distance) and 28 B
// written (16-B PointXYZI + 12-B PointMeta), plus 8 B per firing block of block records.
#pragma once

#include "vs_device.cuh"

namespace vsd {

// new[i] = old[HDL64BeamLUT[i]] (HDLParser.cxx:179-182, 888-889)
__constant__ unsigned char c_beam_lut[64] = {
    38, 39, 42, 43, 32, 33, 36, 37, 40, 41, 46, 47, 50, 51, 54, 55, 44, 45, 48, 49, 52, 53,
    58, 59, 62, 63, 34, 35, 56, 57, 60, 61, 6,  7,  10, 11, 0,  1,  4,  5,  8,  9,  14, 15,
    18, 19, 22, 23, 12, 13, 16, 17, 20, 21, 26, 27, 30, 31, 2,  3,  24, 25, 28, 29};

struct RowsParams {
  const long long* frame_first;  // [frame] first emitted point (stream order); < 0: carried in
  const int* frame_start;        // [frame] packet*12 + block of the frame's first block
  const unsigned* frame_counts;  // [frame][64] points per laser id as pushed
  const BatchHeader* hdr;
  int f_lo;                      // device frame id of the batch's first frame entry
  int n_frames;
  int carry_is_hdl64;
  unsigned carried[kMaxLasers];  // by laser id: points of the first frame held by earlier batches
  unsigned long long carried_total;
  unsigned long long* row_abs;   // [n_frames][64] out, by laser id: slot of the row's first NEW element
};

__global__ void k_layout_rows(const RowsParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_frames) return;
  const int f = p.f_lo + i;
  const bool closed = i + 1 < p.n_frames;
  // splitFrame re-orders when isHDL64Data is set at the closing wrap (sticky since the first
  // iterated 0xddff block, HDLParser.cxx:1032-1033)
  const bool order =
      closed && (p.carry_is_hdl64 || p.hdr->first_upper_block <= (long long)p.frame_start[f + 1]);
  unsigned long long base = 0;
  if (i > 0) base = (unsigned long long)p.frame_first[f] + p.carried_total;
  unsigned long long acc = 0;
  for (int r = 0; r < kMaxLasers; ++r) {
    const int laser = order ? (int)c_beam_lut[r] : r;
    const unsigned car = (i == 0) ? p.carried[laser] : 0u;
    p.row_abs[(long long)i * kMaxLasers + laser] = base + acc + car;
    acc += (unsigned long long)car + p.frame_counts[(long long)f * kMaxLasers + laser];
  }
}

constexpr int kLayTile = 8;     // packets per CTA tile
#ifndef VS_LAY_WARPS
#define VS_LAY_WARPS 16
#endif
constexpr int kLayWarps = VS_LAY_WARPS;  // each warp owns a run of consecutive firing blocks of the tile
constexpr int kLayThreads = 32 * kLayWarps;
constexpr int kLayChunk = kLayTile;  // look-back granularity (host sizing)
constexpr int kLayTileBlocks = kLayTile * kBlocks;           // 96
constexpr int kLayWarpBlocks = kLayTileBlocks / kLayWarps;   // blocks per warp
static_assert(kLayTileBlocks % kLayWarps == 0 && kBlocks % kLayWarpBlocks == 0,
              "a warp's blocks lie inside one packet");
constexpr int kLaySlots = kLayTile * kBlocks * kReturns + kMaxLasers;  // staging rows + 1 pad slot each
constexpr unsigned long long kLayM30 = (1ull << 30) - 1ull;
constexpr unsigned long long kLaySeg = 1ull << 60;  // a frame starts inside the tile

struct LayoutParams {
  const PktSeg* pkt_seg;
  const BlkRec* recs;
  const unsigned long long* pkt_off;
  const float* x;
  const float* y;
  const float* z;
  const uint8_t* inten;
  const uint16_t* az;
  const uint16_t* dist;
  const DevConfig* cfg;
  const unsigned long long* row_abs;
  unsigned long long* st;  // [n_tiles][32] look-back words, zeroed
  int* chunk_counter;      // zeroed
  int n;                   // packets including the halo
  int halo;
  int n_chunks;            // tiles
  int f_lo;
  int adj;                 // 2: VLP-16 (return slots l and l + 16 of a block are the same laser)
  uint8_t* xyzi;           // n_slots records of xyzi_stride bytes (16: x y z intensity;
  int xyzi_stride;         //   32: pcl::PointXYZI with its padding, data[3] = 1)
  uint8_t* meta;           // n_slots PointMeta records of 12 bytes, or null
};

__device__ __forceinline__ void stg_v4(void* p, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// Dynamic shared memory of a layout CTA.  The points of a tile are staged laser-major -- every
// laser's run contiguous, in stream order -- and leave as whole runs: the destination rows are
// then written in coalesced, (all but the first and last) complete 32-byte sectors.  Written
// point by point from the decode order instead (32 lanes -> 32 rows), every 16-byte PointXYZI and
// 12-byte PointMeta is a partial sector, which this part's ECC-protected HBM turns into a
// read-modify-write: measured 2.4x the DRAM traffic at a fifth of the bandwidth.
struct LayShared {
  uint4 xyzi[kLaySlots];          // staged PointXYZI
  unsigned meta[kLaySlots * 3];   // staged PointMeta, 3 words each
  uint2 rec[kLayTileBlocks];
  int4 seg[kLayTile];
  unsigned long long off[kLayTile];
  unsigned wcnt[kLayWarps][2][32];  // per warp, laser bank, return slot: points in the segment
  unsigned ltot[kMaxLasers];        // per laser id: points of the segment in this tile
  unsigned lofs[kMaxLasers];        // staging slot of the laser's run
  unsigned long long rdst[kMaxLasers];  // destination slot of the laser's run
  int seg_start[kLayTileBlocks + 2];    // tile-relative block index where each segment begins
  int n_seg;
  int tile;
};

// one tile (8 packets); called by all threads of the CTA
__device__ __forceinline__ void layout_tile(const LayoutParams& p, LayShared& sh, const int tile) {
  const int wb0 = (int)(threadIdx.x >> 5) * kLayWarpBlocks;  // this warp's first block of the tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned full = 0xffffffffu;
  const bool vlp = p.adj == 2;
  const int P0 = p.halo + tile * kLayTile;
  const int npk = min(kLayTile, p.n - P0);
  for (int i = tid; i < npk * kBlocks; i += kLayThreads) sh.rec[i] = __ldg(&p.recs[(long long)P0 * kBlocks + i]);
  if (tid < npk) {
    sh.seg[tid] = __ldg(&p.pkt_seg[P0 + tid]);
    sh.off[tid] = __ldg(&p.pkt_off[P0 + tid]);
  }
  __syncthreads();
  // segments: runs of blocks between frame starts (the split happens before the wrap block is
  // decoded, HDLParser.cxx:1035-1039).  42 of 43 tiles hold no frame start: one segment, no list.
  unsigned anywrap = 0;
  for (int lp = 0; lp < npk; ++lp) anywrap |= ((unsigned)sh.seg[lp].x >> 4) & 0xfffu;
  const bool simple = anywrap == 0u;
  const int n_blocks = npk * kBlocks;
  if (!simple) {
    if (tid == 0) {
      int ns = 0;
      sh.seg_start[ns++] = 0;
      for (int lp = 0; lp < npk; ++lp) {
        unsigned wm = ((unsigned)sh.seg[lp].x >> 4) & 0xfffu;
        while (wm) {
          const int j = __ffs(wm) - 1;
          wm &= wm - 1;
          const int g = lp * kBlocks + j;
          if (g == 0) continue;  // the tile's first block: segment 0 simply belongs to the new frame
          sh.seg_start[ns++] = g;
        }
      }
      sh.seg_start[ns] = n_blocks;
      sh.n_seg = ns;
    }
    __syncthreads();
  }
  const int n_seg = simple ? 1 : sh.n_seg;
  const bool starts_frame = (((unsigned)sh.seg[0].x >> 4) & 1u) != 0u;  // a frame starts at block 0

  // points of this warp's blocks per (bank, return slot) inside blocks [g0, g1) of the tile
  auto count_range = [&](int g0, int g1, unsigned& c0, unsigned& c1) {
    c0 = c1 = 0;
    const int j0 = max(g0, wb0), j1 = min(min(g1, wb0 + kLayWarpBlocks), n_blocks);
    for (int j = j0; j < j1; ++j) {
      const uint2 r = sh.rec[j];
      const unsigned bit = (r.x >> lane) & 1u;
      const unsigned bank = (r.y >> 25) & 1u;
      unsigned add = bit;
      if (vlp && !bank) add += __shfl_xor_sync(full, bit, 16);
      if (bank) c1 += add; else c0 += add;
    }
  };
  // what precedes this warp's blocks inside the counted range, per (bank, slot), and the totals
  unsigned b0 = 0, b1 = 0, t0 = 0, t1 = 0;
  auto prefix_over_warps = [&]() {
    b0 = b1 = t0 = t1 = 0;
#pragma unroll
    for (int w = 0; w < kLayWarps; ++w) {
      const unsigned v0 = sh.wcnt[w][0][lane], v1 = sh.wcnt[w][1][lane];
      if (w < warp) {
        b0 += v0;
        b1 += v1;
      }
      t0 += v0;
      t1 += v1;
    }
  };

  // ---- this tile's contribution to the open frame: counts of its last segment, published early ---
  {
    unsigned c0, c1;
    count_range(simple ? 0 : sh.seg_start[n_seg - 1], n_blocks, c0, c1);
    sh.wcnt[warp][0][lane] = c0;
    sh.wcnt[warp][1][lane] = c1;
    if (tid < kMaxLasers) sh.ltot[tid] = 0u;
  }
  __syncthreads();
  prefix_over_warps();
  const bool seg = n_seg > 1 || starts_frame;
  unsigned long long* my = p.st + (long long)tile * 32 + lane;
  const unsigned pub0 = t0, pub1 = t1;  // (warp 0 uses them)
  if (warp == 0) {
    const unsigned long long mine = (unsigned long long)pub0 | ((unsigned long long)pub1 << 30);
    // a tile that holds a frame start knows its inclusive prefix without looking back
    if (seg || tile == 0)
      st_release_u64(my, kFlagPrefix | (seg ? kLaySeg : 0ull) | mine);
    else
      st_release_u64(my, kFlagAgg | mine);
  }
  // what earlier tiles hold of the frame that is open at the tile's first block: one warp walks
  // back over the published words while the others stage (called by warp 0 in segment 0)
  auto look_back = [&](unsigned& e0, unsigned& e1) {
    e0 = e1 = 0;
    if (tile == 0 || starts_frame) return;
    // The lanes carry the 32 return slots, so the walk over the predecessors is serial; a
    // window of kWin words per lane is fetched at once so that a round trip covers kWin tiles
    // (a frame spans ~43 tiles and tiles retire faster than a prefix could hop tile by tile).
    constexpr int kWin = 8;
    bool done = false;
    int idx = tile - 1;
    while (true) {
      unsigned long long v[kWin];
#pragma unroll
      for (int w = 0; w < kWin; ++w)
        v[w] = (!done && idx - w >= 0) ? ld_acquire_u64(p.st + (long long)(idx - w) * 32 + lane) : 0ull;
      if (!done) {
#pragma unroll
        for (int w = 0; w < kWin; ++w) {
          const unsigned flag = (unsigned)(v[w] >> 62);
          if (flag == 0u) break;  // not published yet: poll again from here
          e0 += (unsigned)(v[w] & kLayM30);
          e1 += (unsigned)((v[w] >> 30) & kLayM30);
          if (flag == 2u || --idx < 0) {
            done = true;
            break;
          }
        }
      }
      if (__all_sync(full, done)) break;
    }
    if (!seg) st_release_u64(my, kFlagPrefix | (unsigned long long)(e0 + pub0) | ((unsigned long long)(e1 + pub1) << 30));
  };

  int id0 = lane, id1 = lane + 32;  // laser ids of this return slot in a 0xeeff / 0xddff block
  if (vlp) {                        // HDLParser.cxx:935-943
    if (id0 >= 16) id0 -= 16;
    id1 -= 16;
  }
  const double dc0 = __ldg(&p.cfg->cal[2][lane]), dc1 = __ldg(&p.cfg->cal[2][lane + 32]);
  const unsigned lt_mask = (1u << lane) - 1u;
  const int frame0 = sh.seg[0].y - p.f_lo + (starts_frame ? 1 : 0);  // frame of segment 0

  for (int s = 0; s < n_seg; ++s) {
    const int g0 = simple ? 0 : sh.seg_start[s], g1 = simple ? n_blocks : sh.seg_start[s + 1];
    const int fr = frame0 + s;
    if (n_seg > 1) {
      // ---- A: counts of the segment per warp (a one-segment tile counted them above) -------------
      __syncthreads();  // wcnt / ltot and, from the second segment on, the staging are free again
      unsigned c0, c1;
      count_range(g0, g1, c0, c1);
      sh.wcnt[warp][0][lane] = c0;
      sh.wcnt[warp][1][lane] = c1;
      if (tid < kMaxLasers) sh.ltot[tid] = 0u;
      __syncthreads();
      prefix_over_warps();
    }
    if (warp == 0) {
      // per laser id (VLP-16: slots l and l + 16 carry the same, combined count)
      if (t0) sh.ltot[id0] = t0;
      if (t1) sh.ltot[id1] = t1;
      __syncwarp();
      // staging slot of each laser's run: exclusive prefix over the laser ids, one pad slot per
      // laser so that equally long runs start in different shared-memory banks
      const unsigned a = sh.ltot[lane], b = sh.ltot[lane + 32];
      unsigned ia = a + 1u, ib = b + 1u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(full, ia, o), v = __shfl_up_sync(full, ib, o);
        if (lane >= o) {
          ia += u;
          ib += v;
        }
      }
      const unsigned tot_a = __shfl_sync(full, ia, 31);
      sh.lofs[lane] = ia - (a + 1u);
      sh.lofs[lane + 32] = tot_a + ib - (b + 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // destination of each run: the laser's row of the segment's frame, behind what earlier
      // tiles hold of that frame (first segment only; the walk back overlaps the staging)
      unsigned e0 = 0, e1 = 0;
      if (s == 0) look_back(e0, e1);
      sh.rdst[lane] = __ldg(&p.row_abs[(long long)fr * kMaxLasers + lane]);
      sh.rdst[lane + 32] = __ldg(&p.row_abs[(long long)fr * kMaxLasers + lane + 32]);
      __syncwarp();
      // counts by (bank, slot) -> by laser id
      if (e0 && (!vlp || lane < 16)) sh.rdst[id0] += e0;
      if (e1) sh.rdst[id1] += e1;
    }
    // ---- B: stage the segment's points of this warp's blocks laser-major ---------------------------
    {
      const int j0 = max(g0, wb0), j1 = min(min(g1, wb0 + kLayWarpBlocks), n_blocks);
      const unsigned long long poff = sh.off[min(wb0 / kBlocks, kLayTile - 1)];
      const unsigned so0 = sh.lofs[id0], so1 = sh.lofs[id1];
      unsigned c0 = b0, c1 = b1;
      // two steps per group of kLayIlp blocks: ranks and all the group's loads first, then the
      // shared-memory stores -- the loads of a group are in flight together instead of one
      // block's round trip after the other
      constexpr int kLayIlp = 3;
      static_assert(kLayWarpBlocks % kLayIlp == 0, "whole groups");
#pragma unroll 1
      for (int jg = wb0; jg < wb0 + kLayWarpBlocks; jg += kLayIlp) {
        unsigned slot[kLayIlp];
        bool put[kLayIlp];
        float vx[kLayIlp], vy[kLayIlp], vz[kLayIlp];
        unsigned vi[kLayIlp], va[kLayIlp], vd[kLayIlp], bk[kLayIlp];
#pragma unroll
        for (int u = 0; u < kLayIlp; ++u) {
          const int j = jg + u;
          put[u] = false;
          slot[u] = 0u;
          bk[u] = 0u;
          vx[u] = vy[u] = vz[u] = 0.f;
          vi[u] = va[u] = vd[u] = 0u;
          if (j < j0 || j >= j1) continue;  // warp-uniform
          const uint2 r = sh.rec[j];
          if (r.x == 0u) continue;
          const unsigned bit = (r.x >> lane) & 1u;
          const unsigned bank = (r.y >> 25) & 1u;
          unsigned rank = bank ? c1 : c0;
          unsigned add = bit;
          if (vlp && !bank) {
            const unsigned pb = __shfl_xor_sync(full, bit, 16);
            if (lane >= 16) rank += pb;  // slot l (first firing of the block) is pushed before l + 16
            add += pb;
          }
          if (bank) c1 += add; else c0 += add;
          if (bit) {
            const unsigned long long src = poff + ((r.y >> 16) & 0x1ffu) + __popc(r.x & lt_mask);
            put[u] = true;
            bk[u] = bank;
            slot[u] = (bank ? so1 : so0) + rank;
            vx[u] = __ldg(&p.x[src]);
            vy[u] = __ldg(&p.y[src]);
            vz[u] = __ldg(&p.z[src]);
            vi[u] = __ldg(&p.inten[src]);
            va[u] = __ldg(&p.az[src]);
            vd[u] = __ldg(&p.dist[src]);
          }
        }
#pragma unroll
        for (int u = 0; u < kLayIlp; ++u) {
          if (!put[u]) continue;
          // p.intensity = intensity (HDLParser.cxx:737)
          sh.xyzi[slot[u]] = make_uint4(__float_as_uint(vx[u]), __float_as_uint(vy[u]),
                                        __float_as_uint(vz[u]), __float_as_uint((float)vi[u]));
          // PointMeta{u16 azimuth; float distance = distanceM; 3 flag bytes} (type_defs.h:168-176,
          // HDLParser.cxx:614, 745-747); the flags the reference leaves indeterminate are zero
          const double dm = __dadd_rn(__dmul_rn((double)vd[u], 0.002), bk[u] ? dc1 : dc0);
          sh.meta[3u * slot[u]] = va[u];
          sh.meta[3u * slot[u] + 1u] = __float_as_uint((float)dm);
          sh.meta[3u * slot[u] + 2u] = 0u;
        }
      }
    }
    __syncthreads();
    // ---- C: every laser's run leaves as one contiguous, coalesced copy ------------------------------
    for (int L = warp; L < kMaxLasers; L += kLayWarps) {
      const unsigned nrun = sh.ltot[L];
      if (nrun == 0u) continue;
      const unsigned so = sh.lofs[L];
      const unsigned long long dst = sh.rdst[L];
      if (p.xyzi_stride == 16) {
        uint4* o = reinterpret_cast<uint4*>(p.xyzi) + dst;
        for (unsigned i = lane; i < nrun; i += 32) {
          const uint4 v = sh.xyzi[so + i];
          stg_v4(o + i, v.x, v.y, v.z, v.w);
        }
      } else {
        uint4* o = reinterpret_cast<uint4*>(p.xyzi) + 2ull * dst;
        for (unsigned i = lane; i < 2u * nrun; i += 32) {
          const uint4 v = sh.xyzi[so + (i >> 1)];
          if (i & 1u)
            stg_v4(o + i, v.w, 0u, 0u, 0u);
          else
            stg_v4(o + i, v.x, v.y, v.z, __float_as_uint(1.0f));
        }
      }
      if (p.meta) {
        unsigned* o = reinterpret_cast<unsigned*>(p.meta) + 3ull * dst;
        for (unsigned i = lane; i < 3u * nrun; i += 32) o[i] = sh.meta[3u * so + i];
      }
    }
  }
}

// Persistent CTAs (2 per SM); tile ids are dealt in execution order by an atomic counter, so a
// tile only ever waits (look-back) for tiles whose CTAs are already running.
__global__ void __launch_bounds__(kLayThreads, 2) k_layout(const LayoutParams p) {
  extern __shared__ __align__(16) uint8_t lay_smem[];
  LayShared& sh = *reinterpret_cast<LayShared*>(lay_smem);
  while (true) {
    __syncthreads();  // the previous tile's copies out of the staging are done
    if (threadIdx.x == 0) sh.tile = atomicAdd(p.chunk_counter, 1);
    __syncthreads();
    const int tile = sh.tile;
    if (tile >= p.n_chunks) break;
    layout_tile(p, sh, tile);
  }
}

}  // namespace vsd
