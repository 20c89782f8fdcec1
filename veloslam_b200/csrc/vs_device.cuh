// vs_device.cuh -- device-side structures and small PTX helpers shared by the kernels.
//
// sm_100a only.  Nothing here is a dense contraction, so no tensor-core path: the kernels
// are HBM-bound byte/integer + FP64 scalar work (DESIGN.md "Kernels").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vsd {

constexpr int kPacketBytes = 1206;
constexpr int kBlocks = 12;
constexpr int kReturns = 32;
constexpr int kMaxLasers = 64;
constexpr int kLutSize = 36001;  // type_defs.h:16 HDL_NUM_ROT_ANGLES

// Per-context decode configuration, resident in HBM and staged into shared memory by every
// decode CTA (7.5 KB).  Built on the host by vs_set_calibration / vs_set_filters.
struct DevConfig {
  // per-laser rows, SoA so lane l reads consecutive 8-byte words (HDLParser.cxx:89-100):
  //   0 cos(rad(rotCorrection)) 1 sin(rad(rotCorrection)) 2 distanceCorrection (m)
  //   3 cosVertCorrection       4 sinVertCorrection       5 verticalOffsetCorrection (m)
  //   6 horizontalOffsetCorrection (m)
  double cal[7][kMaxLasers];
  // (timestampadjustment - blockdsr0) / (nextblockdsr0 - blockdsr0) per (block, dsr), evaluated
  // on the host with the reference's exact operation order (HDLParser.cxx:946-961)
  double az_ratio[kBlocks][kReturns];
  // round(timestampadjustment) per (block, dsr), microseconds (HDLParser.cxx:962)
  uint16_t tadj[kBlocks][kReturns];
  double crop[6];
  unsigned long long laser_mask;
  // laser selection folded per return slot: bit l of sel_lo / sel_hi == "slot l of a 0xeeff /
  // 0xddff block is emitted" (laserSelections[laserId] && laserId < calibFileReportedNumLasers,
  // with the VLP-16 id remap of HDLParser.cxx:935-943)
  unsigned sel_lo;
  unsigned sel_hi;
  int points_skip;
  int crop_returns;
  int crop_inside;
  int n_enabled;   // calibFileReportedNumLasers
  int adj_mode;    // 0 none (HDL-64 / other), 1 HDL-32, 2 VLP-16
};

// Per-packet segmentation record written by k_segment, read by k_pose / k_decode.
//   x: bits 0-3 skip_in (firingSkip entering the packet), bits 4-15 wrap mask over the
//      iterated blocks, bits 16-31 azimuthDiff (7th smallest of the 11 azimuth deltas)
//   y: k_scan: emitted-point count; after k_pose: number of wraps before this packet
//      (frame id at the packet's first block)
//   z: after k_pose: packet time - t_base (microseconds, u32)
//   w: bits 0-11 mask of the 0xddff ("upper") blocks, bits 12-20 emitted points of the packet
typedef int4 PktSeg;

// Per-block record written by k_scan, read by k_pose / k_decode.
//   x: mask of the return slots the reference emits (iterated block, distance != 0, selected
//      laser, crop)
//   y: bits 0-15 block azimuth (mod 36000 when no per-return adjustment applies, else raw),
//      bits 16-24 emitted points of the packet in front of this block, bit 25 0xddff block,
//      bits 26-29 wraps of the packet up to and including this block
typedef uint2 BlkRec;

// Batch header: written by the kernels, copied to the host with the frame tables.
struct BatchHeader {
  long long total_points;
  long long first_upper_block;  // packet*12 + block of the first iterated 0xddff block
  long long last_origin_time;   // packet time of last_origin_packet
  double carry_origin_T[3];     // frame origin the next batch inherits (n_poses >= 2 only)
  int total_wraps;
  int last_azimuth;
  int firing_skip_out;
  int last_has_wrap;            // last packet closed a frame: next batch starts un-inited
  int first_const_pkt;          // first packet whose skip map is constant (halo check)
  int origin_at_halo;           // origin packet of the first decoded packet
  int frame_at_halo;            // frame id at the first decoded packet
  int last_origin_packet;       // origin packet the next batch's first packet inherits
  int frame_overflow;           // a frame id exceeded the table capacity
  int time_range_error;         // a decoded packet's time - t_base does not fit the t_us column
};

// Batch header + the first rows of the frame tables in one block (k_frames), for batches small
// enough that the number of copies, not their size, sets the latency.
constexpr int kEagerRows = 64;
struct EagerBlock {
  BatchHeader hdr;
  long long first[kEagerRows];
  long long meta_time[kEagerRows];
  int start[kEagerRows];
  int meta_pkt[kEagerRows];
  int skips[kEagerRows];
  unsigned counts[kEagerRows][kMaxLasers];
};

// The t_us column is u32 microseconds after t_base plus the return's firing offset (a u16), the
// same arithmetic as the reference's rawtime + round(timestampadjustment) (HDLParser.cxx:970):
// packet times must lie in [t_base, t_base + kTimeSpanMax).
constexpr unsigned long long kTimeSpanMax = 0xffff0000ull;

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Non-blocking probe of a phase (acquire): true when the phase with this parity has completed.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA bulk copy global -> shared (1-D, 16-byte granules), completion on an mbarrier.
// SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Look-back words carry flag and payload in one 64-bit value, so a relaxed (L2-coherent) load
// is enough; ld.acquire.gpu would make ptxas emit CCTL.IVALL after every poll and throw away
// the SM's L1 (LUT, pose rows) each time.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 16-byte look-back words (single-copy atomic: SASS LDG/STG.E.128.STRONG.GPU), so that two
// scanned quantities travel under one flag.
__device__ __forceinline__ void ld_relaxed_b128(const void* p, unsigned long long& lo,
                                                unsigned long long& hi) {
  asm volatile(
      "{\n"
      ".reg .b128 v;\n"
      "ld.relaxed.gpu.global.b128 v, [%2];\n"
      "mov.b128 {%0, %1}, v;\n"
      "}\n"
      : "=l"(lo), "=l"(hi)
      : "l"(p)
      : "memory");
}
__device__ __forceinline__ void st_relaxed_b128(void* p, unsigned long long lo,
                                                unsigned long long hi) {
  asm volatile(
      "{\n"
      ".reg .b128 v;\n"
      "mov.b128 v, {%0, %1};\n"
      "st.relaxed.gpu.global.b128 [%2], v;\n"
      "}\n" ::"l"(lo),
      "l"(hi), "l"(p)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// Decoupled look-back (single-pass chained scan across tiles).  One 64-bit word per tile:
// bits 63-62 flag (0 empty, 1 tile aggregate, 2 inclusive prefix), bits 61-0 payload.
// Traits: T, identity(), combine(a,b) (commutative + associative), pack(T)->62 bits, unpack.
// Called by all 32 lanes of one warp; returns the exclusive prefix in every lane.
// ---------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kPayloadMask = (1ull << 62) - 1;

template <class Tr>
__device__ __forceinline__ typename Tr::T lookback_exclusive(unsigned long long* state, int tile,
                                                             typename Tr::T agg) {
  typedef typename Tr::T T;
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) st_release_u64(&state[0], kFlagPrefix | Tr::pack(agg));
    return Tr::identity();
  }
  if (lane == 0) st_release_u64(&state[tile], kFlagAgg | Tr::pack(agg));
  T excl = Tr::identity();
  int base = tile - 1;
  while (true) {
    const int idx = base - lane;
    unsigned long long v;
    do {
      v = (idx >= 0) ? ld_acquire_u64(&state[idx]) : (kFlagPrefix | Tr::pack(Tr::identity()));
    } while (__any_sync(0xffffffffu, (v >> 62) == 0));
    const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    const int stop = pm ? (__ffs(pm) - 1) : 31;
    T val = (lane <= stop) ? Tr::unpack(v & kPayloadMask) : Tr::identity();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val = Tr::combine(val, Tr::shfl_xor(val, o));
    excl = Tr::combine(excl, val);
    if (pm) break;
    base -= 32;
  }
  if (lane == 0) st_release_u64(&state[tile], kFlagPrefix | Tr::pack(Tr::combine(excl, agg)));
  return excl;
}

// 62-bit sum (emitted point counts).
struct SumTraits {
  typedef unsigned long long T;
  __device__ static T identity() { return 0ull; }
  __device__ static T combine(T a, T b) { return a + b; }
  __device__ static unsigned long long pack(T v) { return v; }
  __device__ static T unpack(unsigned long long v) { return v; }
  __device__ static T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
};

// (30-bit sum of wraps, 32-bit max of origin markers).
struct WrapTraits {
  typedef unsigned long long T;  // high 32: sum, low 32: max
  __device__ static T identity() { return 0ull; }
  __device__ static T combine(T a, T b) {
    const unsigned long long s = (a >> 32) + (b >> 32);
    const unsigned ma = (unsigned)a, mb = (unsigned)b;
    return (s << 32) | (ma > mb ? ma : mb);
  }
  __device__ static unsigned long long pack(T v) { return ((v >> 32) << 32) | (v & 0xffffffffull); }
  __device__ static T unpack(unsigned long long v) { return v; }
  __device__ static T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
};

// 12-entry skip maps packed as 4-bit nibbles: nibble s = firingSkip leaving the packet when it
// was entered with firingSkip == s.
constexpr unsigned long long kMapIdentity = 0xBA9876543210ull;
__device__ __forceinline__ unsigned long long map_compose(unsigned long long first,
                                                          unsigned long long second) {
  // result[s] = second[first[s]]
  unsigned long long r = 0;
#pragma unroll
  for (int s = 0; s < 12; ++s) {
    const unsigned a = (unsigned)(first >> (4 * s)) & 15u;
    r |= ((second >> (4 * a)) & 15ull) << (4 * s);
  }
  return r;
}
__device__ __forceinline__ bool map_is_const(unsigned long long m) {
  return m == (m & 15ull) * 0x111111111111ull;
}
__device__ __forceinline__ int map_apply(unsigned long long m, int s) {
  return (int)((m >> (4 * s)) & 15ull);
}

}  // namespace vsd
