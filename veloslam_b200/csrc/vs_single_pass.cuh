// vs_single_pass.cuh -- the single-pass variant of the decode kernel: scan-warp role, look-back
// over 16-byte words, k_pose_pre.  Included by vs_kernels.cuh between the declarations it needs
// (DecParams, DecCtl, DecLayout, the shared-memory helpers) and k_decode, which calls the role.
// Opt-in at run time (VELOSLAM_SINGLE_PASS=1): measured slower than the two-pass pipeline on
// B200 (profiles/r1z_single_pass_experiment.txt).
#pragma once

namespace vsd {

// =========================================================================================
// Single-pass variant (FUSED): k_decode with the segmentation scans inside.
//
// Two extra warps per CTA (the scan warps, alternating over the CTA's tile sequence) do, for each
// 8-packet tile and up to two tiles ahead of the decode warps, what k_scan and the scan half of
// k_pose do in the two-pass pipeline:
//   * deal the CTA the next tile id (global atomic, in sequence order inside the CTA: forward
//     progress never depends on CTAs that are not resident), start the TMA copies of the packets
//     and of the pose rows [L | T];
//   * from the staged bytes: wrap masks, firingSkip chain (the previous packet's own map decides
//     it on sensor data; look-back over per-tile maps otherwise), azimuthDiff; the slot bits
//     "distance != 0 and laser selected" of the 96 firing blocks come from the decode warps (one
//     ballot per block, written two tiles ahead of their decode pass, `masks` mbarrier);
//   * one decoupled look-back over 16-byte words {flag | emitted points, wraps << 32 | origin
//     marker} gives the tile's first point, frame id and frame-origin packet;
//   * writes the block records, segment records and point offsets of the tile straight into the
//     shared-memory stage the decode warps read (the same layout the two-pass kernel receives by
//     TMA), re-bases the translation of the pose rows to the frame origin, fills the frame
//     table and the batch header, and hands the stage over through the `ready` mbarrier.
// The packets are read from HBM once; block records, point offsets and re-based pose rows never
// exist in HBM.  Not for the crop filter (needs the position of a return before its emission is
// known) and not for the per-point deskew extension: those batches take the two-pass pipeline.
// =========================================================================================
constexpr int kScanWarps = 2;
#ifdef VS_PROFILE_FUSED
// development aid: cycles spent per wait site, summed over warps (lane 0 clocks)
__device__ unsigned long long g_fused_prof[16];
#define VS_PROF_DECL long long prof_t0 = 0, prof_acc[12] = {0,0,0,0,0,0,0,0,0,0,0,0};
#define VS_PROF_T0() prof_t0 = clock64()
#define VS_PROF_ADD(i) do { const long long c_ = clock64(); prof_acc[i] += c_ - prof_t0; prof_t0 = c_; } while (0)
#define VS_PROF_FLUSH(lane) do { if ((lane) == 0) { for (int i_ = 0; i_ < 12; ++i_) if (prof_acc[i_]) atomicAdd(&g_fused_prof[i_], (unsigned long long)prof_acc[i_]); } } while (0)
#else
#define VS_PROF_DECL
#define VS_PROF_T0()
#define VS_PROF_ADD(i)
#define VS_PROF_FLUSH(lane)
#endif
constexpr int kFusedThreads = kDecThreads + 32 * kScanWarps;

__device__ __forceinline__ void sts_u32(uint32_t a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// Decoupled look-back over the 16-byte words, split in two: publish the tile's aggregate early,
// resolve the exclusive prefix of (points, wraps / origin marker) later.
__device__ __forceinline__ void lookback_publish(ulonglong2* st, int tile, unsigned long long agg_c,
                                                 unsigned long long agg_w, int lane) {
  if (lane == 0) st_relaxed_b128(&st[2 * tile], (tile == 0 ? kFlagPrefix : kFlagAgg) | agg_c, agg_w);
}
__device__ __forceinline__ void lookback_resolve(ulonglong2* st, int tile, unsigned long long agg_c,
                                                 unsigned long long agg_w, int lane,
                                                 unsigned long long& ex_c, unsigned long long& ex_w) {
  ex_c = 0ull;
  ex_w = 0ull;
  if (tile == 0) return;
  int base = tile - 1;
  while (true) {
    const int idx = base - lane;
    unsigned long long lo = kFlagPrefix, hi = 0ull;
    do {
      if (idx >= 0) ld_relaxed_b128(&st[2 * idx], lo, hi);
    } while (__any_sync(0xffffffffu, (lo >> 62) == 0ull));
    const unsigned pm = __ballot_sync(0xffffffffu, (lo >> 62) == 2ull);
    const int stop = pm ? (__ffs(pm) - 1) : 31;
    unsigned long long c = (lane <= stop) ? (lo & kPayloadMask) : 0ull;
    unsigned long long w = (lane <= stop) ? hi : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, o);
      w = WrapTraits::combine(w, __shfl_xor_sync(0xffffffffu, w, o));
    }
    ex_c += c;
    ex_w = WrapTraits::combine(ex_w, w);
    if (pm) break;
    base -= 32;
  }
  if (lane == 0)
    st_relaxed_b128(&st[2 * tile], kFlagPrefix | (ex_c + agg_c), WrapTraits::combine(ex_w, agg_w));
}

// Scan warp `sw` of kScanWarps: handles the CTA's tile sequence numbers k == sw (mod kScanWarps).
template <int ADJ, class L>
__device__ __forceinline__ void fused_scan_role(const DecParams& p, DecCtl& sh, uint8_t* smem_raw,
                                                const int lane, const int sw) {
  constexpr int S = L::kNumStages;
  constexpr int kDPkts = L::kDPkts;
  constexpr unsigned kFull = 0xffffffffu;
  const uint32_t smem_a = smem_u32(smem_raw);
  const uint32_t stage_a0 = smem_a + L::kStages;
  uint8_t* stage0 = smem_raw + L::kStages;
  const long long in_base = reinterpret_cast<long long>(p.pkts);
  const unsigned stride = (unsigned)p.stride;
  const int pskip = p.cfg->points_skip;
  const bool pose_valid = p.pose_valid != 0;
  unsigned long long* const st_map = reinterpret_cast<unsigned long long*>(p.st) + 2;  // stride 4 words
  long long fub_seen = 0x7fffffffffffffffll;  // smallest first-upper-block candidate seen so far
  unsigned gate_mask = 0;  // bit j: block j passes the pointsSkip gate (HDLParser.cxx:1042)
#pragma unroll
  for (int j = 0; j < kBlocks; ++j)
    if (pskip == 0 || (j % (pskip + 1)) == 0) gate_mask |= 1u << j;

  // fill stage k % S with the next tile; returns true when no tile is left.  Tile ids are dealt
  // in sequence order inside a CTA (grab_seq), so the first end marker is final.
  VS_PROF_DECL
  auto issue = [&](int k, bool ended) -> bool {
    const int s = k % S;
    VS_PROF_T0();
    if (k >= S) mbar_wait(&sh.empty[s], (uint32_t)(k / S - 1) & 1u);
    VS_PROF_ADD(3);
    fence_proxy_async();  // this lane's writes into the stage (records, pose rows) -> async proxy
    __syncwarp();
    int t = -1;
    if (lane == 0) {
      while (*reinterpret_cast<volatile int*>(&sh.grab_seq) != k) {
      }
      if (!ended) {
        t = atomicAdd(p.tile_counter, 1);
        if (t >= p.n_tiles) t = -1;
      }
      sh.tile_id[s] = t;
      if (t >= 0) {
        const long long first = (long long)t * kDecTile;
        const TileSpan sp = tile_span(in_base, p.stride, p.total_bytes, p.n, first, 0, kDecTile);
        uint8_t* st = stage0 + (size_t)s * p.stage_bytes;
        for (long long a = sp.s1; a < sp.a1; ++a)  // < 16 bytes, last tile of the array only
          st[kDPkts + (a - sp.s0)] = *reinterpret_cast<const uint8_t*>(a);
        const uint32_t bytes = (uint32_t)(sp.s1 - sp.s0);
        const uint32_t pbytes = pose_valid ? (uint32_t)sp.npk * 96u : 0u;
        mbar_expect_tx(&sh.full[s], bytes + pbytes);
        if (pbytes) bulk_g2s(st + kDPose, p.pose_mat + first * 12, pbytes, &sh.full[s]);
        if (bytes) bulk_g2s(st + kDPkts, reinterpret_cast<const void*>(sp.s0), bytes, &sh.full[s]);
      } else {
        mbar_arrive(&sh.full[s]);  // end marker: the decode warps' mask pass waits on `full`
      }
      __threadfence_block();
      *reinterpret_cast<volatile int*>(&sh.grab_seq) = k + 1;
    }
    t = __shfl_sync(kFull, t, 0);
    __syncwarp();
    VS_PROF_ADD(4);
    return t < 0;
  };

  // sequence numbers below S are issued up front
  bool ended = false;
  for (int k = sw; k < S; k += kScanWarps) ended = issue(k, ended);

#pragma unroll 1
  for (int it = sw;; it += kScanWarps) {
    const int s = it % S;
    const int tile = *reinterpret_cast<volatile int*>(&sh.tile_id[s]);
    if (tile < 0) {
      if (lane == 0) mbar_arrive(&sh.ready[s]);
      break;
    }
    const bool refill = it + kScanWarps >= S;

    const int first = tile * kDecTile;
    const int npk = min(kDecTile, p.n - first);
    const int P = first + lane;
    const bool live = lane < npk;
    const uint32_t st_a = stage_a0 + (uint32_t)s * (uint32_t)p.stage_bytes;
    const uint32_t nz_a = smem_a + L::kScr + (uint32_t)s * (kDecTile * kBlocks * 4);

    // what the tile needs from outside its own bytes, requested before the wait: the azimuths of
    // the packet in front of it (lanes 0-11), the last azimuth of the one before that (lane 12)
    // and the packet times
    int hv = p.carry_last_az;
    if (first > 0) {
      const uint8_t* pp = p.pkts + (long long)(first - 1) * p.stride;
      if (lane < 12)
        hv = (int)__ldg(reinterpret_cast<const unsigned short*>(pp + 100 * lane + 2));
      else if (lane == 12 && first > 1)
        hv = (int)__ldg(reinterpret_cast<const unsigned short*>(pp - p.stride + 1102));
    }
    long long tp = 0;
    if (live) tp = __ldg(&p.pkt_time[P]);

    VS_PROF_T0();
    mbar_wait(&sh.full[s], (uint32_t)(it / S) & 1u);
    VS_PROF_ADD(0);
    const uint32_t pkt_a = st_a + kDPkts + (((unsigned)in_base + (unsigned)first * stride) & 15u);

    // ---- lane == packet: headers, skip maps, firingSkip entering each packet ------------------
    unsigned wm = 0, em = 0, um = 0, wrapmask = 0;
    int azdiff = 0, s_in = 0, az11 = 0;
    unsigned long long m = kMapIdentity;
    int az[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) az[j] = 0;
    if (live) {
      const uint32_t pk_a = pkt_a + (unsigned)lane * stride;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (lds_u16(pk_a + 100u * j) != 0xeeffu) um |= 1u << j;
        az[j] = (int)lds_u16(pk_a + 100u * j + 2u);
      }
      az11 = az[11];
    }
    int prev11 = __shfl_up_sync(kFull, az11, 1);
    const int hv11 = __shfl_sync(kFull, hv, 11);
    if (lane == 0) prev11 = (first > 0) ? hv11 : p.carry_last_az;
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
      if (P < p.halo && map_is_const(m)) atomicMin(&p.hdr->first_const_pkt, P);
    }
    unsigned long long inc = m;
    const bool all_zero = __all_sync(kFull, m == 0ull || !live);
    if (!all_zero) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long prev = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc = map_compose(prev, inc);
      }
    } else if (!live) {
      inc = 0ull;  // dead lanes follow a constant-0 packet
    }
    const unsigned long long agg = __shfl_sync(kFull, inc, 31);
    // the map of the packet in front of the tile decides the firingSkip entering the tile whenever
    // it is constant (always, on sensor data); lane 0 evaluates it from the 13 azimuths the lanes
    // requested before the wait
    int paz[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) paz[j] = __shfl_sync(kFull, hv, j);
    const int pprev = __shfl_sync(kFull, hv, 12);
    int skip_tile = 0;
    if (lane == 0) {
      if (tile == 0) {
        skip_tile = p.carry_skip;
      } else {
        st_release_u64(st_map + 4 * tile, kFlagAgg | agg);
        bool known = false;
        if (p.mode != 0) {
          known = true;
        } else {
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
              v = ld_acquire_u64(st_map + 4 * idx);
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
      st_release_u64(st_map + 4 * tile, kFlagPrefix | (unsigned long long)map_apply(agg, skip_tile));
    }
    skip_tile = __shfl_sync(kFull, skip_tile, 0);
    unsigned long long excl = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) excl = kMapIdentity;
    s_in = (p.mode == 0) ? map_apply(excl, skip_tile) : 0;
    if (live) wrapmask = (wm & ~((2u << s_in) - 1u)) | (((em >> s_in) & 1u) << s_in);

    // ---- slot bits "distance != 0 and laser selected" of the tile's 96 blocks: written by the
    // decode warps (one ballot per block, a tile ahead of their decode pass) ---------------------
    VS_PROF_ADD(9);
    mbar_wait(&sh.masks[s], (uint32_t)(it / S) & 1u);
    VS_PROF_ADD(1);

    // ---- lane == packet: emitted points, tile aggregate published for the tiles behind ---------
    unsigned cnt = 0;
    unsigned nzv[kBlocks];
    // blocks the parser iterates and emits (HDLParser.cxx:1042): j >= firingSkip, pointsSkip gate
    const unsigned open_mask = gate_mask & ~((1u << s_in) - 1u);
#pragma unroll
    for (int j = 0; j < kBlocks; ++j) nzv[j] = 0u;
    if (live) {
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) nzv[j] = lds_u32(nz_a + 4u * (unsigned)(lane * kBlocks + j));
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) cnt += ((open_mask >> j) & 1u) ? __popc(nzv[j]) : 0u;
      if (P < p.halo) cnt = 0;  // halo packets rebuild state only
    }
    const unsigned nw = __popc(wrapmask);
    // origin marker: streaming -> the packet after the wrap re-initialises the frame meta
    // (F4b); offline -> the wrap packet itself.  marker - 1 == origin packet index.
    const unsigned long long v2 =
        nw ? (((unsigned long long)nw << 32) | (unsigned)(P + (p.mode == 0 ? 2 : 1))) : 0ull;
    unsigned long long inc2 = v2, incc = cnt;
#pragma unroll
    for (int o = 1; o < kDecTile; o <<= 1) {  // lanes >= kDecTile hold zeros
      const unsigned long long a = __shfl_up_sync(kFull, inc2, o);
      const unsigned long long b = __shfl_up_sync(kFull, incc, o);
      if (lane >= o) {
        inc2 = WrapTraits::combine(a, inc2);
        incc += b;
      }
    }
    const unsigned long long agg_c = __shfl_sync(kFull, incc, kDecTile - 1);
    const unsigned long long agg_w = __shfl_sync(kFull, inc2, kDecTile - 1);
    lookback_publish(p.st, tile, agg_c, agg_w, lane);

    // ---- block records into the stage (gates of HDLParser.cxx:1042-1051), lane == packet: the
    // azimuths and bank bits are still in registers from the header pass ------------------------
    if (live) {
      const bool decoded = P >= p.halo;
      unsigned pre = 0;  // emitted points of the packet in front of block j
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) {
        const bool open = ((open_mask >> j) & 1u) != 0u;
        const unsigned upper = (um >> j) & 1u;
        const unsigned wb = __popc(wrapmask & ((2u << j) - 1u));
        unsigned a = (unsigned)az[j];
        if (ADJ == 0) a %= 36000u;  // HDLParser.cxx:597 (ADJ != 0: adjusted per return later)
        const unsigned rx = (open && decoded) ? nzv[j] : 0u;
        const unsigned ry = a | (pre << 16) | (upper << 25) | (wb << 26);
        sts_u64(st_a + kDRec + 8u * (unsigned)(lane * kBlocks + j), ((unsigned long long)ry << 32) | rx);
        if (open) pre += __popc(nzv[j]);
      }
    }

    // ---- scans across tiles: first point, frame id, frame origin of every packet ---------------
    unsigned long long pre_c, pre_w;
    VS_PROF_ADD(5);
    lookback_resolve(p.st, tile, agg_c, agg_w, lane, pre_c, pre_w);
    VS_PROF_ADD(2);
    unsigned long long ex2 = __shfl_up_sync(kFull, inc2, 1);
    unsigned long long exc = __shfl_up_sync(kFull, incc, 1);
    if (lane == 0) {
      ex2 = 0ull;
      exc = 0ull;
    }
    ex2 = WrapTraits::combine(pre_w, ex2);
    exc += pre_c;
    // seed: with no frame meta carried in, packet 0 is the origin until the first wrap
    if (!p.carry_meta_inited) ex2 = WrapTraits::combine(ex2, 1ull);
    const int frame_base = (int)(ex2 >> 32);
    const int origin = (int)(unsigned)ex2 - 1;

    // pose rows: the origin's translation is requested now and used after the record writes
    const int pk3 = lane / 3, r3 = lane - 3 * pk3;
    double To = 0.0;
    int org3 = -1;
    unsigned nw3 = 0;
    if (pose_valid) {
      const int src = pk3 < kDecTile ? pk3 : 0;
      org3 = __shfl_sync(kFull, origin, src);
      nw3 = __shfl_sync(kFull, nw, src);
      if (pk3 < npk) {
        if (org3 < 0)
          To = (r3 == 0) ? p.carry_origin_T[0] : (r3 == 1 ? p.carry_origin_T[1] : p.carry_origin_T[2]);
        else if (org3 != first + pk3)
          To = __ldg(&p.pose_mat[(long long)org3 * 12 + 4 * r3 + 3]);
      }
    }

    if (live) {
      PktSeg seg;
      seg.x = s_in | (int)(wrapmask << 4) | (azdiff << 16);
      seg.y = frame_base;
      seg.z = (int)(unsigned)(tp - p.t_base);
      if ((unsigned long long)(tp - p.t_base) >= kTimeSpanMax && P >= p.halo) p.hdr->time_range_error = 1;
      seg.w = (int)(um | (cnt << 12));
      sts_v4(st_a + kDSeg + 16u * (unsigned)lane, (unsigned)seg.x, (unsigned)seg.y, (unsigned)seg.z,
             (unsigned)seg.w);
      sts_u64(st_a + kDOff + 8u * (unsigned)lane, exc);
      p.seg_out[P] = seg;
      const unsigned ium = um & ~((1u << s_in) - 1u);
      if (ium) {
        const long long fu = (long long)P * 12 + (__ffs(ium) - 1);
        if (fu < fub_seen) {
          // once per lane in practice: this CTA's tile ids only grow
          const unsigned long long old = atomicMin(
              reinterpret_cast<unsigned long long*>(&p.hdr->first_upper_block), (unsigned long long)fu);
          fub_seen = min((long long)old, fu);
        }
      }
      // frame table: a wrap block opens frame f before it is decoded (HDLParser.cxx:1035-1039)
      if (wrapmask && P >= p.halo) {
        unsigned w = wrapmask;
        int f = frame_base;
        while (w) {
          const int j = __ffs(w) - 1;
          w &= w - 1;
          unsigned before = 0;
          for (int q = s_in; q < j; ++q)
            if ((gate_mask >> q) & 1u) before += __popc(lds_u32(nz_a + 4u * (unsigned)(lane * kBlocks + q)));
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
        p.hdr->last_azimuth = az11;
        p.hdr->firing_skip_out = (p.mode == 0) ? map_apply(m, s_in) : 0;
        p.hdr->total_wraps = frame_base + (int)nw;
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
    }

    // ---- pose rows: translation re-based to the frame origin (reprojectToFrameBeginning,
    // HDLParser.cxx:1057); lane == (packet, axis) -----------------------------------------------
    if (pose_valid && pk3 < npk) {
      const int Pp = first + pk3;
      const uint32_t a = st_a + kDPose + 96u * (unsigned)pk3 + 8u * (unsigned)(4 * r3 + 3);
      const double T = lds_f64(a);
      if (org3 == Pp) To = T;
      sts_f64(a, __dsub_rn(T, To));
      if (Pp == p.n - 1) {
        const bool self = (p.mode == 1) && nw3;  // offline: the wrap packet is its frame's origin
        p.hdr->carry_origin_T[r3] = self ? T : To;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sh.ready[s]);
    VS_PROF_ADD(10);

    // refill the stage the decode warps free next
    if (refill) ended = issue(it + kScanWarps, ended);
  }
  VS_PROF_FLUSH(lane);
}

// Pose rows of the single-pass variant: [L | T] per packet from the packet time alone
// (TransformManager.cxx:149-177, type_defs.h:134-146); the translation is re-based to the frame
// origin by the scan warp once the origin packet is known.
__global__ void __launch_bounds__(256) k_pose_pre(const long long* __restrict__ pkt_time, int n,
                                                  int n_poses, const long long* __restrict__ pose_t,
                                                  const double* __restrict__ pose_trv,
                                                  double* __restrict__ pose_mat) {
  const int P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= n) return;
  double T[3], R[3];
  interp_pose(pose_t, pose_trv, n_poses, __ldg(&pkt_time[P]), T, R, true);
  double Lm[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  rotate_by(Lm, to_radians(R[0]), 1);  // type_defs.h:136 UnitY
  rotate_by(Lm, to_radians(R[1]), 0);  // :137 UnitX
  rotate_by(Lm, to_radians(R[2]), 2);  // :138 UnitZ
  double* o = pose_mat + (long long)P * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    o[4 * r + 0] = Lm[r][0];
    o[4 * r + 1] = Lm[r][1];
    o[4 * r + 2] = Lm[r][2];
    o[4 * r + 3] = T[r];
  }
}

}  // namespace vsd
