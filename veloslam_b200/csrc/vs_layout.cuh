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
//                  across warps it is a decoupled look-back over one 64-bit word per (chunk, lane)
//                  that restarts at every frame boundary.
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

constexpr int kLayChunk = 16;   // packets per warp
constexpr int kLayWarps = 8;
constexpr int kLayThreads = 32 * kLayWarps;
constexpr unsigned long long kLayM30 = (1ull << 30) - 1ull;
constexpr unsigned long long kLaySeg = 1ull << 60;  // a frame starts inside the chunk

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
  unsigned long long* st;  // [n_chunks][32] look-back words, zeroed
  int* chunk_counter;      // zeroed
  int n;                   // packets including the halo
  int halo;
  int n_chunks;
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

__global__ void __launch_bounds__(kLayThreads) k_layout(const LayoutParams p) {
  __shared__ uint2 s_rec[kLayWarps][kLayChunk * kBlocks];
  __shared__ int4 s_seg[kLayWarps][kLayChunk];
  __shared__ unsigned long long s_off[kLayWarps][kLayChunk];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // chunk ids are dealt in launch order, so a chunk only ever waits for chunks whose warps are
  // already running
  if (tid == 0) s_base = atomicAdd(p.chunk_counter, kLayWarps);
  __syncthreads();
  const int chunk = s_base + warp;
  if (chunk >= p.n_chunks) return;
  const int P0 = p.halo + chunk * kLayChunk;
  const int npk = min(kLayChunk, p.n - P0);
  for (int i = lane; i < npk * kBlocks; i += 32) s_rec[warp][i] = __ldg(&p.recs[(long long)P0 * kBlocks + i]);
  if (lane < npk) {
    s_seg[warp][lane] = __ldg(&p.pkt_seg[P0 + lane]);
    s_off[warp][lane] = __ldg(&p.pkt_off[P0 + lane]);
  }
  __syncwarp();
  const bool vlp = p.adj == 2;
  const unsigned full = 0xffffffffu;

  // ---- pass 1: points of this chunk per (laser bank, return slot) since the last frame start --
  unsigned c0 = 0, c1 = 0;
  bool seg = false;
  for (int lp = 0; lp < npk; ++lp) {
    const unsigned wrapmask = ((unsigned)s_seg[warp][lp].x >> 4) & 0xfffu;
#pragma unroll
    for (int j = 0; j < kBlocks; ++j) {
      const uint2 r = s_rec[warp][lp * kBlocks + j];
      if ((wrapmask >> j) & 1u) {  // the split happens before the block is decoded (:1035-1039)
        c0 = c1 = 0;
        seg = true;
      }
      const unsigned bit = (r.x >> lane) & 1u;
      const unsigned bank = (r.y >> 25) & 1u;
      unsigned add = bit;
      if (vlp && !bank) add += __shfl_xor_sync(full, bit, 16);
      if (bank) c1 += add; else c0 += add;
    }
  }
  unsigned long long* my = p.st + (long long)chunk * 32 + lane;
  const unsigned long long mine = (unsigned long long)c0 | ((unsigned long long)c1 << 30);
  // a chunk that holds a frame start knows its inclusive prefix without looking back
  if (seg || chunk == 0)
    st_release_u64(my, kFlagPrefix | (seg ? kLaySeg : 0ull) | mine);
  else
    st_release_u64(my, kFlagAgg | mine);
  // ---- look-back: counts between the open frame's start and this chunk --------------------------
  unsigned e0 = 0, e1 = 0;
  if (chunk > 0) {
    bool done = false;
    int idx = chunk - 1;
    while (true) {
      if (!done) {
        const unsigned long long v = ld_acquire_u64(p.st + (long long)idx * 32 + lane);
        const unsigned flag = (unsigned)(v >> 62);
        if (flag != 0u) {
          e0 += (unsigned)(v & kLayM30);
          e1 += (unsigned)((v >> 30) & kLayM30);
          if (flag == 2u || --idx < 0) done = true;
        }
      }
      if (__all_sync(full, done)) break;
    }
    if (!seg) st_release_u64(my, kFlagPrefix | (unsigned long long)(e0 + c0) | ((unsigned long long)(e1 + c1) << 30));
  }

  // ---- pass 2: move the points ------------------------------------------------------------------
  c0 = e0;
  c1 = e1;
  const unsigned lt_mask = (1u << lane) - 1u;
  int id0 = lane, id1 = lane + 32;  // laser ids of this return slot in a 0xeeff / 0xddff block
  if (vlp) {                        // HDLParser.cxx:935-943
    if (id0 >= 16) id0 -= 16;
    id1 -= 16;
  }
  const double dc0 = __ldg(&p.cfg->cal[2][lane]), dc1 = __ldg(&p.cfg->cal[2][lane + 32]);
  int f_cur = -1;
  unsigned long long ra0 = 0, ra1 = 0;
  for (int lp = 0; lp < npk; ++lp) {
    const int4 sg = s_seg[warp][lp];
    const unsigned wrapmask = ((unsigned)sg.x >> 4) & 0xfffu;
    const int fbase = sg.y - p.f_lo;
    const unsigned long long poff = s_off[warp][lp];
#pragma unroll 2
    for (int j = 0; j < kBlocks; ++j) {
      const uint2 r = s_rec[warp][lp * kBlocks + j];
      if ((wrapmask >> j) & 1u) c0 = c1 = 0;
      if (r.x == 0u) continue;
      const int fr = fbase + __popc(wrapmask & ((2u << j) - 1u));
      if (fr != f_cur) {
        f_cur = fr;
        ra0 = __ldg(&p.row_abs[(long long)fr * kMaxLasers + id0]);
        ra1 = __ldg(&p.row_abs[(long long)fr * kMaxLasers + id1]);
      }
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
        const unsigned long long dst = (bank ? ra1 : ra0) + rank;
        const float vx = __ldg(&p.x[src]), vy = __ldg(&p.y[src]), vz = __ldg(&p.z[src]);
        const unsigned vi = __ldg(&p.inten[src]);
        const unsigned va = __ldg(&p.az[src]), vd = __ldg(&p.dist[src]);
        const float fi = (float)vi;  // p.intensity = intensity (HDLParser.cxx:737)
        uint8_t* o = p.xyzi + dst * (unsigned long long)p.xyzi_stride;
        if (p.xyzi_stride == 16) {
          stg_v4(o, __float_as_uint(vx), __float_as_uint(vy), __float_as_uint(vz), __float_as_uint(fi));
        } else {
          stg_v4(o, __float_as_uint(vx), __float_as_uint(vy), __float_as_uint(vz), __float_as_uint(1.0f));
          stg_v4(o + 16, __float_as_uint(fi), 0u, 0u, 0u);
        }
        if (p.meta) {
          // PointMeta{u16 azimuth; float distance = distanceM; 3 flag bytes} (type_defs.h:168-176,
          // HDLParser.cxx:614, 745-747); the flags the reference leaves indeterminate are zero
          const double dm = __dadd_rn(__dmul_rn((double)vd, 0.002), bank ? dc1 : dc0);
          unsigned* m = reinterpret_cast<unsigned*>(p.meta + dst * 12ull);
          m[0] = va;
          m[1] = __float_as_uint((float)dm);
          m[2] = 0u;
        }
      }
    }
  }
}

}  // namespace vsd
