// HDLParser.cpp -- host side of the drop-in facade: batching into a pinned ring, GPU decode
// through the C ABI, and assembly of reference-shaped HDLFrames from the stream-order columns.
#include "HDLParser.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../include/veloslam_b200.h"
#include "CalibrationFile.h"
#include "vtkPacketFile.h"

namespace {
// HDL-64 beam re-order applied when a frame is closed on HDL-64 data: new[i] = old[LUT[i]]
// (sensor manual ordering, reference HDLParser.cxx:179-182)
const int kHDL64BeamLUT[64] = {38, 39, 42, 43, 32, 33, 36, 37, 40, 41, 46, 47, 50, 51, 54, 55,
                               44, 45, 48, 49, 52, 53, 58, 59, 62, 63, 34, 35, 56, 57, 60, 61,
                               6,  7,  10, 11, 0,  1,  4,  5,  8,  9,  14, 15, 18, 19, 22, 23,
                               12, 13, 16, 17, 20, 21, 26, 27, 30, 31, 2,  3,  24, 25, 28, 29};
}  // namespace

class HDLParser::vsInternal {
 public:
  // points of the open frame that earlier batches decoded, as they sit on the host: rows by
  // laser id (an open frame is not permuted yet), in an arena of its own
  struct Partial {
    std::shared_ptr<vs::Arena> arena;
    size_t metaOffset = 0;
    bool hasMeta = false;  // the arena holds PointMeta records behind the points
    uint32_t rowStart[HDL_MAX_NUM_LASERS];
    uint32_t rowCount[HDL_MAX_NUM_LASERS];
    uint64_t total = 0;
    Partial() {
      std::memset(rowStart, 0, sizeof(rowStart));
      std::memset(rowCount, 0, sizeof(rowCount));
    }
  };
  // one batch on its way through the GPU
  struct Batch {
    int stage = 0;  // 1: submitted (kernels running), 2: laid out, device -> host copies running
    uint64_t ticket = 0;
    int ring = 0;
    int64_t n = 0;
    int64_t recCursor = -1;
    int64_t packetBase = 0;
    int nLasers = 0;
    std::vector<vs_frame> frames;
    std::vector<vs_frame_rows> rows;
    std::vector<std::shared_ptr<vs::Arena> > arenas;
    std::vector<size_t> metaOffset;
    std::vector<std::vector<std::pair<ptime, std::string> > > packets;
  };

  vsInternal()
      : ctx(nullptr), device(0), batchPackets(4096), storePackets(true), fetchMeta(true),
        pipelined(false), fill(0), pending(0), pendingWrap(false), hostLastAz(-1),
        correctionsInitialized(false), calibFileReportedNumLasers(64), numberOfTrailingFrames(0),
        applyTransform(0), pointsSkip(0), shouldCropReturns(false), shouldCropInside(false),
        dualReturnFilter(0), configDirty(true), warnedNoCalib(false) {
    for (int i = 0; i < 6; ++i) cropRegion[i] = 0.0;
    for (int i = 0; i < HDL_MAX_NUM_LASERS; ++i) laserSelections[i] = 1;
    for (int i = 0; i < 2; ++i) {
      ringPkts[i] = nullptr;
      ringTimes[i] = nullptr;
    }
    std::memset(&calib, 0, sizeof(calib));
    std::memset(openCounts, 0, sizeof(openCounts));
    vs_carry_init(&carry);
  }
  ~vsInternal() {
    drain();
    unloadRecording();
    destroyContext();
  }

  void destroyContext() {
    if (ctx) vs_destroy(ctx);
    ctx = nullptr;
    for (int i = 0; i < 2; ++i) {
      vs_host_free(ringPkts[i]);
      vs_host_free(ringTimes[i]);
      ringPkts[i] = nullptr;
      ringTimes[i] = nullptr;
    }
  }

  // all or nothing: a context without its packet rings is torn down again
  bool ensureContext() {
    if (ctx) return true;
    const int64_t maxPoses = std::max<int64_t>(1 << 16, batchPackets / 4);
    const int rc = vs_create(device, batchPackets, maxPoses, pipelined ? 2 : 1, &ctx);
    if (rc != VS_OK) {
      error = std::string("vs_create failed: ") + vs_last_error(nullptr) +
              " (status " + std::to_string(rc) + "; there is no CPU fallback)";
      std::cerr << error << std::endl;
      ctx = nullptr;
      return false;
    }
    bool ok = true;
    for (int i = 0; i < (pipelined ? 2 : 1) && ok; ++i)
      ok = vs_host_alloc((uint64_t)batchPackets * VS_PACKET_BYTES, (void**)&ringPkts[i]) == VS_OK &&
           vs_host_alloc((uint64_t)batchPackets * sizeof(int64_t), (void**)&ringTimes[i]) == VS_OK;
    if (!ok) {
      error = "pinned host allocation failed for the packet ring";
      std::cerr << error << std::endl;
      destroyContext();
      return false;
    }
    maxPosesCtx = maxPoses;
    configDirty = true;
    return true;
  }

  // calibration / filters when they changed; the slice of the pose timeline the batch's packet
  // times [tmin, tmax] can touch (not the whole, ever-growing history)
  bool syncConfig(int64_t tmin, int64_t tmax) {
    if (configDirty) {
      if (correctionsInitialized) {
        if (vs_set_calibration(ctx, calib.rows, calib.n_rows, calibFileReportedNumLasers) != VS_OK) {
          error = vs_last_error(ctx);
          return false;
        }
      }
      vs_filters f;
      std::memset(&f, 0, sizeof(f));
      for (int i = 0; i < 64; ++i)
        if (laserSelections[i]) f.laser_mask |= 1ull << i;
      f.points_skip = pointsSkip;
      f.crop_returns = shouldCropReturns ? 1 : 0;
      f.crop_inside = shouldCropInside ? 1 : 0;
      for (int i = 0; i < 6; ++i) f.crop_region[i] = cropRegion[i];
      if (vs_set_filters(ctx, &f) != VS_OK) {
        error = vs_last_error(ctx);
        return false;
      }
      configDirty = false;
    }
    poseT.clear();
    poseTrv.clear();
    if (transMgr) transMgr->snapshotWindow(tmin, tmax, &poseT, &poseTrv);
    if ((int64_t)poseT.size() > maxPosesCtx) {
      error = "the batch spans more poses than the context holds (lower setBatchPackets)";
      return false;
    }
    if (vs_set_poses(ctx, poseT.data(), poseTrv.data(), (int64_t)poseT.size()) != VS_OK) {
      error = vs_last_error(ctx);
      return false;
    }
    return true;
  }

  // frame object without the reference's 128 reserve(2200) calls: its lists are built by
  // adoption when the frame closes
  std::shared_ptr<HDLFrame> newFrameShell() {
    std::shared_ptr<HDLFrame> f(new HDLFrame);
    f->isInMemory = true;
    return f;
  }
  std::shared_ptr<HDLFrame> createHDLFrame() {
    std::shared_ptr<HDLFrame> f(new HDLFrame);
    f->points.resize(calibFileReportedNumLasers);
    f->pointsMeta.resize(calibFileReportedNumLasers);
    for (int i = 0; i < calibFileReportedNumLasers; ++i) {
      f->points[i] = pcl::PointCloud<pcl::PointXYZI>::Ptr(new pcl::PointCloud<pcl::PointXYZI>);
      f->points[i]->points.reserve(HDL_MAX_PTS_PER_LASER);
      f->pointsMeta[i] = std::shared_ptr<PointMetaVector>(new PointMetaVector);
      f->pointsMeta[i]->reserve(HDL_MAX_PTS_PER_LASER);
    }
    f->isInMemory = true;
    return f;
  }

  // HDLFrame::points / pointsMeta of a frame on top of its arena: the rows the GPU laid out
  // (already in the order splitFrame leaves them, reference HDLParser.cxx:867-897) are adopted,
  // not copied.  nRows: 64 for a frame closed on HDL-64 data, else the calibrated laser count;
  // rows of lasers the calibration does not have stay empty.
  void adoptRows(HDLFrame& f, const std::shared_ptr<vs::Arena>& arena, size_t metaOff,
                 const uint32_t* rowStart, const uint32_t* rowCount, const int32_t* rowLaser,
                 bool hdl64Order, int nLasers, bool hasMeta) {
    const int nRows = hdl64Order ? 64 : nLasers;
    f.points.resize((size_t)nRows);
    f.pointsMeta.resize((size_t)nRows);
    pcl::PointXYZI* xyzi = arena ? reinterpret_cast<pcl::PointXYZI*>(arena->data()) : nullptr;
    PointMeta* meta = arena ? reinterpret_cast<PointMeta*>(arena->data() + metaOff) : nullptr;
    for (int r = 0; r < nRows; ++r) {
      pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>);
      std::shared_ptr<PointMetaVector> pm(new PointMetaVector);
      const int laser = rowLaser ? rowLaser[r] : r;
      if (arena && laser < nLasers && rowCount[r] > 0) {
        vs::adopt(cloud->points, arena, xyzi + rowStart[r], (size_t)rowCount[r]);
        if (hasMeta) vs::adopt(*pm, arena, meta + rowStart[r], (size_t)rowCount[r]);
      }
      cloud->width = (uint32_t)cloud->points.size();
      cloud->height = 1;
      f.points[(size_t)r] = cloud;
      f.pointsMeta[(size_t)r] = pm;
    }
  }

  // the open frame as the reference's forced split leaves it (getFrame at end of file,
  // HDLParser.cxx:540-543): rows by laser id, then the pointer shuffle of splitFrame
  void materializeOpenFrame(bool hdl64Order) {
    HDLFrame& f = *currentFrame;
    adoptRows(f, partial.arena, partial.metaOffset, partial.rowStart, partial.rowCount, nullptr, false,
              calibFileReportedNumLasers, partial.hasMeta);
    if (hdl64Order) {
      std::vector<pcl::PointCloud<pcl::PointXYZI>::Ptr> pts(64);
      std::vector<std::shared_ptr<PointMetaVector> > ptm(64);
      for (int i = 0; i < 64; ++i) {
        const size_t src = (size_t)kHDL64BeamLUT[i];
        if (src < f.points.size()) {
          pts[i] = f.points[src];
          ptm[i] = f.pointsMeta[src];
        } else {
          pts[i] = pcl::PointCloud<pcl::PointXYZI>::Ptr(new pcl::PointCloud<pcl::PointXYZI>);
          ptm[i] = std::shared_ptr<PointMetaVector>(new PointMetaVector);
        }
      }
      f.points = std::move(pts);
      f.pointsMeta = std::move(ptm);
    }
  }

  void applyMeta(HDLFrame& f, const vs_frame& e) {
    f.timestamp = ptime(e.timestamp_us);
    if (e.skips >= 0) f.skips = (uint8_t)e.skips;
    for (int k = 0; k < 3; ++k) {
      f.carpose->T[k] = e.carpose[k];
      f.carpose->R[k] = e.carpose[3 + k];
      f.carpose->V[k] = e.carpose[6 + k];
    }
    f.carpose->seconds_pos = e.carpose_valid ? 0 : -1;
    f.carpose->timestamp = f.timestamp;
  }

  void failBatches(const std::string& what) {
    error = what;
    std::cerr << "HDLParser: GPU decode failed: " << error << std::endl;
    // the packets of the failed batch (and of any batch behind it) are lost; the parser carries
    // on from the state before them.  Copies into the batches' arenas may still be running: wait
    // for them before the arenas go back to the pool.
    if (ctx)
      for (auto& b : inflight)
        if (b.stage >= 1) vs_sync(ctx, b.ticket, nullptr);
    inflight.clear();
    pending = 0;
    pendingWrap = false;
  }

  // ---- stage 0 -> 1: hand the filled ring (or a slice of the resident recording) to the GPU ----
  bool submitBatch(int64_t recCursorArg) {
    if (pending == 0) return true;
    if (!ensureContext()) {
      pending = 0;
      pendingWrap = false;
      return false;
    }
    Batch b;
    b.n = pending;
    b.ring = fill;
    b.recCursor = recCursorArg;
    b.packetBase = packetBase;
    b.nLasers = calibFileReportedNumLasers;
    int64_t tmin, tmax;
    if (recCursorArg >= 0) {
      const uint8_t* h = recHost.data() + VS_PCAP_PAYLOAD_OFFSET + (size_t)recCursorArg * VS_PCAP_RECORD_BYTES;
      tmin = INT64_MAX;
      tmax = INT64_MIN;
      for (int64_t i = 0; i < pending; ++i) {
        uint32_t tv[2];
        std::memcpy(tv, h + (size_t)i * VS_PCAP_RECORD_BYTES - 58, 8);
        const int64_t t = timevalToPtime(tv[0], tv[1]).us;
        tmin = std::min(tmin, t);
        tmax = std::max(tmax, t);
      }
    } else {
      const int64_t* tt = ringTimes[fill];
      tmin = *std::min_element(tt, tt + pending);
      tmax = *std::max_element(tt, tt + pending);
    }
    mark("times");
    if (!syncConfig(tmin, tmax)) {
      failBatches(error);
      return false;
    }
    mark("syncConfig");
    int rc;
    if (recCursorArg >= 0) {
      // packets come from the recording resident in HBM: no host -> device copy at all
      const uint8_t* dev = static_cast<const uint8_t*>(recDev) + VS_PCAP_PAYLOAD_OFFSET +
                           (size_t)recCursorArg * VS_PCAP_RECORD_BYTES;
      rc = vs_submit(ctx, dev, VS_PCAP_RECORD_BYTES, nullptr, pending, 0, VS_MODE_STREAMING,
                     VS_FLAG_DEVICE_INPUT | VS_FLAG_PCAP_TIMES, tmin, &carry, &b.ticket);
    } else {
      // the t_us column counts from the earliest packet of the batch
      rc = vs_submit(ctx, ringPkts[fill], VS_PACKET_BYTES, ringTimes[fill], pending, 0,
                     VS_MODE_STREAMING, 0, tmin, &carry, &b.ticket);
    }
    if (rc != VS_OK) {
      failBatches(vs_last_error(ctx));
      return false;
    }
    mark("vs_submit");
    b.stage = 1;
    inflight.push_back(std::move(b));
    packetBase += pending;
    pending = 0;
    pendingWrap = false;
    if (pipelined) fill ^= 1;
    return true;
  }

  // ---- stage 1 -> 2: frame table, HDLFrame layout on the device, device -> host copies ----------
  bool advance(Batch& b) {
    vs_result r;
    const double tw = now();
    int rc = vs_wait(ctx, b.ticket, &r);
    secWait += now() - tw;
    mark("vs_wait");
    if (rc != VS_OK) {
      failBatches(vs_last_error(ctx));
      return false;
    }
    b.frames.assign(r.frames, r.frames + r.n_frames);
    bool anyCarried = false;
    for (int l = 0; l < HDL_MAX_NUM_LASERS; ++l) anyCarried = anyCarried || openCounts[l] != 0;
    vs_layout lay;
    rc = vs_layout_frames(ctx, b.ticket, anyCarried ? openCounts : nullptr, 16, fetchMeta ? 1 : 0, &lay);
    if (rc != VS_OK) {
      failBatches(vs_last_error(ctx));
      return false;
    }
    mark("vs_layout_frames");
    b.rows.assign(lay.rows, lay.rows + lay.n_frames);
    b.arenas.resize(b.rows.size());
    b.metaOffset.assign(b.rows.size(), 0);
    for (size_t i = 0; i < b.rows.size(); ++i) {
      const size_t n = (size_t)b.rows[i].n_slots;
      if (n == 0) continue;
      const size_t metaOff = (n * sizeof(pcl::PointXYZI) + 255) & ~(size_t)255;
      b.arenas[i] = vs::Arena::acquire(metaOff + (fetchMeta ? n * sizeof(PointMeta) : 0));
      if (!b.arenas[i]) {
        failBatches("page-locked host memory for a frame could not be allocated");
        return false;
      }
      b.metaOffset[i] = metaOff;
      rc = vs_fetch_layout(ctx, b.ticket, b.rows[i].first_slot, (int64_t)n, b.arenas[i]->data(),
                           fetchMeta ? b.arenas[i]->data() + metaOff : nullptr);
      if (rc != VS_OK) {
        failBatches(vs_last_error(ctx));
        return false;
      }
    }
    mark("arenas+vs_fetch_layout");
    // raw packets: a packet is stored in the frame that is current when it arrives; the frame's
    // first packet is stored twice (reference HDLParser.cxx:999 + 1009).  Copied now: the ring
    // is handed back to the receiver as soon as this returns.
    if (storePackets) {
      const uint8_t* rawBase = ringPkts[b.ring];
      size_t rawStride = VS_PACKET_BYTES;
      if (b.recCursor >= 0) {
        rawBase = recHost.data() + VS_PCAP_PAYLOAD_OFFSET + (size_t)b.recCursor * VS_PCAP_RECORD_BYTES;
        rawStride = VS_PCAP_RECORD_BYTES;
      }
      auto rawTime = [&](int p) -> ptime {
        if (b.recCursor < 0) return ptime(ringTimes[b.ring][p]);
        uint32_t tv[2];
        std::memcpy(tv, rawBase + (size_t)p * rawStride - 58, 8);
        return timevalToPtime(tv[0], tv[1]);
      };
      b.packets.resize(b.frames.size());
      for (size_t i = 0; i < b.frames.size(); ++i) {
        const vs_frame& e = b.frames[i];
        const int first = (i == 0) ? 0 : e.start_packet + 1;
        const int last = (i + 1 < b.frames.size()) ? b.frames[i + 1].start_packet : (int)b.n - 1;
        if (last >= first) b.packets[i].reserve((size_t)(last - first) + 2);
        for (int p = first; p <= last; ++p) {
          // one allocation + one copy per stored packet: the string is built inside the pair
          const char* raw = reinterpret_cast<const char*>(rawBase) + (size_t)p * rawStride;
          if (p == e.meta_packet) b.packets[i].emplace_back(rawTime(p), std::string(raw, VS_PACKET_BYTES));
          b.packets[i].emplace_back(rawTime(p), std::string(raw, VS_PACKET_BYTES));
        }
      }
    }
    // parser state after the batch: what the next submit starts from
    carry = r.carry_out;
    const vs_frame_rows& open = b.rows.back();
    for (int l = 0; l < HDL_MAX_NUM_LASERS; ++l) openCounts[l] = open.row_count[l];  // rows == laser ids
    mark("raw packets");
    b.stage = 2;
    return true;
  }

  // ---- stage 2 -> frames: wait for the copies, adopt the rows ------------------------------------
  bool collect(Batch& b, std::deque<std::shared_ptr<HDLFrame> >* closedOut) {
    const double ts = now();
    const int rc = vs_sync(ctx, b.ticket, nullptr);
    secSync += now() - ts;
    mark("vs_sync");
    if (rc != VS_OK) {
      failBatches(vs_last_error(ctx));
      return false;
    }
    error.clear();
    for (size_t i = 0; i < b.frames.size(); ++i) {
      const vs_frame& e = b.frames[i];
      const vs_frame_rows& rw = b.rows[i];
      if (i > 0) currentFrame = newFrameShell();
      HDLFrame& f = *currentFrame;
      if (e.meta_packet >= 0) applyMeta(f, e);  // -1: carried (already applied), -2: never
      if (storePackets) {
        if (f.packets.empty())
          f.packets = std::move(b.packets[i]);
        else
          for (auto& pk : b.packets[i]) f.packets.push_back(std::move(pk));
      }
      const std::shared_ptr<vs::Arena>& arena = b.arenas[i];
      if (i == 0 && partial.total > 0 && arena) {
        // what earlier batches decoded of this frame goes into the gaps the GPU left at the head
        // of each row
        pcl::PointXYZI* dx = reinterpret_cast<pcl::PointXYZI*>(arena->data());
        PointMeta* dm = reinterpret_cast<PointMeta*>(arena->data() + b.metaOffset[0]);
        const pcl::PointXYZI* sx = reinterpret_cast<const pcl::PointXYZI*>(partial.arena->data());
        const PointMeta* sm = reinterpret_cast<const PointMeta*>(partial.arena->data() + partial.metaOffset);
        for (int r = 0; r < HDL_MAX_NUM_LASERS; ++r) {
          const uint32_t c = rw.row_carried[r];
          if (!c) continue;
          const int laser = rw.row_laser[r];
          std::memcpy(dx + rw.row_start[r], sx + partial.rowStart[laser], (size_t)c * sizeof(pcl::PointXYZI));
          if (fetchMeta) {
            if (partial.hasMeta)
              std::memcpy(dm + rw.row_start[r], sm + partial.rowStart[laser], (size_t)c * sizeof(PointMeta));
            else  // setFetchMeta(true) while this frame was open: no meta was kept for its head
              std::memset(dm + rw.row_start[r], 0, (size_t)c * sizeof(PointMeta));
          }
        }
      }
      if (e.closed) {
        adoptRows(f, arena, b.metaOffset[i], rw.row_start, rw.row_count, rw.row_laser, e.hdl64_order != 0,
                  b.nLasers, fetchMeta);
        closedOut->push_back(currentFrame);
        closedBy.push_back(b.packetBase + b.frames[i + 1].start_packet);  // packet holding the wrap
        partial = Partial();
      } else {
        // still open: its points wait in their arena for the batch that closes the frame
        partial = Partial();
        partial.arena = arena;
        partial.metaOffset = b.metaOffset[i];
        partial.hasMeta = fetchMeta;
        for (int r = 0; r < HDL_MAX_NUM_LASERS; ++r) {
          partial.rowStart[r] = rw.row_start[r];
          partial.rowCount[r] = rw.row_count[r];
        }
        partial.total = (uint64_t)rw.n_slots;
      }
    }
    return true;
  }

  // Non-pipelined decode of whatever is buffered: submit, lay out, fetch, adopt.
  bool decodePending(std::deque<std::shared_ptr<HDLFrame> >* closedOut, int64_t recCursorArg = -1) {
    if (pending == 0) return true;
    marks.clear();
    mark("start");
    bool ok = drainTo(closedOut) && submitBatch(recCursorArg) && drainTo(closedOut);
    mark("done");
    // VELOSLAM_TRACE_SLOW_MS=<ms>: where a decode that took longer than that spent its time
    if (traceSlowMs > 0 && (marks.back().second - marks.front().second) * 1e3 > traceSlowMs) {
      std::cerr << "HDLParser: slow decode #" << nDecodes << ":";
      for (size_t i = 1; i < marks.size(); ++i)
        std::cerr << ' ' << marks[i].first << ' ' << (marks[i].second - marks[i - 1].second) * 1e3 << " ms";
      std::cerr << std::endl;
    }
    ++nDecodes;
    return ok;
  }
  void mark(const char* what) {
    if (traceSlowMs > 0) marks.emplace_back(what, now());
  }
  std::vector<std::pair<const char*, double> > marks;
  double traceSlowMs = std::getenv("VELOSLAM_TRACE_SLOW_MS") ? std::atof(std::getenv("VELOSLAM_TRACE_SLOW_MS")) : 0.0;
  long nDecodes = 0;

  // finish every batch in flight
  bool drainTo(std::deque<std::shared_ptr<HDLFrame> >* closedOut) {
    while (!inflight.empty()) {
      Batch& b = inflight.front();
      if (b.stage == 1 && !advance(b)) return false;
      if (!collect(inflight.front(), closedOut)) return false;
      inflight.pop_front();
    }
    return true;
  }
  bool drain() { return ctx ? drainTo(&frames) : true; }

  // Pipelined mode, ring full: the batch submitted one ring ago has long finished on the GPU --
  // lay it out and start its device -> host copies; finish the one before it (its copies had a
  // whole ring-fill to complete); hand the ring just filled to the GPU.  Host fill, host ->
  // device copy + kernels and device -> host copy of three consecutive batches overlap.
  bool pump() {
    const double t0 = now();
    if (!inflight.empty() && inflight.back().stage == 1 && !advance(inflight.back())) return false;
    const double t1 = now();
    while (inflight.size() > 1) {
      if (!collect(inflight.front(), &frames)) return false;
      inflight.pop_front();
    }
    const double t2 = now();
    const bool ok = submitBatch(-1);
    const double t3 = now();
    secAdvance += t1 - t0;
    secCollect += t2 - t1;
    secSubmit += t3 - t2;
    ++nPumps;
    return ok;
  }
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  double secAdvance = 0, secCollect = 0, secSubmit = 0, secWait = 0, secSync = 0;
  long nPumps = 0;

  // ---- recording resident in HBM (loadRecording) ----------------------------------------------
  // records [first, first + count) of the file (count < 0: to the end); the image kept on the
  // host and in HBM is the file's global header followed by those records, so that a record's
  // place in the image is its file position minus `first` records
  bool loadRecording(const std::string& file, int64_t first = 0, int64_t count = -1) {
    unloadRecording();
    if (!ensureContext()) return false;
    FILE* f = std::fopen(file.c_str(), "rb");
    if (!f) return false;
    fseeko(f, 0, SEEK_END);
    const long long fileBytes = ftello(f);
    const long long body = fileBytes - PCAP_GLOBAL_HEADER_LEN;
    bool ok = fileBytes > PCAP_GLOBAL_HEADER_LEN && body % VS_PCAP_RECORD_BYTES == 0 && first >= 0;
    const int64_t total = ok ? body / VS_PCAP_RECORD_BYTES : 0;
    ok = ok && first < total;
    const int64_t n = !ok ? 0 : count < 0 ? total - first : std::min<int64_t>(count, total - first);
    const long long bytes = PCAP_GLOBAL_HEADER_LEN + (long long)n * VS_PCAP_RECORD_BYTES;
    if (ok) {
      recHost.resize((size_t)bytes);
      fseeko(f, 0, SEEK_SET);
      ok = std::fread(recHost.data(), 1, PCAP_GLOBAL_HEADER_LEN, f) == (size_t)PCAP_GLOBAL_HEADER_LEN;
      fseeko(f, PCAP_GLOBAL_HEADER_LEN + (off_t)first * VS_PCAP_RECORD_BYTES, SEEK_SET);
      const size_t want = (size_t)n * VS_PCAP_RECORD_BYTES;
      ok = ok && std::fread(recHost.data() + PCAP_GLOBAL_HEADER_LEN, 1, want, f) == want;
    }
    std::fclose(f);
    // every record must be a 1206-byte payload behind the 42-byte header: fixed stride
    for (int64_t i = 0; ok && i < n; ++i) {
      uint32_t len[2];
      std::memcpy(len, recHost.data() + PCAP_GLOBAL_HEADER_LEN + (size_t)i * VS_PCAP_RECORD_BYTES + 8, 8);
      ok = len[0] == VS_PACKET_BYTES + 42 && len[1] == VS_PACKET_BYTES + 42;
    }
    ok = ok && recHost[0] == 0xd4 && recHost[1] == 0xc3 && recHost[2] == 0xb2 && recHost[3] == 0xa1;
    if (ok) ok = vs_device_alloc(ctx, (uint64_t)bytes + 64, &recDev) == VS_OK;
    if (ok) ok = vs_device_upload(ctx, recDev, recHost.data(), (uint64_t)bytes) == VS_OK;
    if (!ok) {
      unloadRecording();
      return false;
    }
    recName = file;
    recPackets = n;
    recFirst = first;
    return true;
  }
  void unloadRecording() {
    if (recDev) vs_device_free(ctx, recDev);
    recDev = nullptr;
    recPackets = 0;
    recFirst = 0;
    recName.clear();
    recAlias.clear();
    std::vector<uint8_t>().swap(recHost);
  }
  bool hasRecording(const std::string& file) const {
    return recDev && (file == recName || (!recAlias.empty() && file == recAlias));
  }
  void* recDev = nullptr;
  std::vector<uint8_t> recHost;  // the same file image on the host (raw packets of HDLFrames)
  int64_t recPackets = 0;
  int64_t recFirst = 0;          // file record the image starts at (a shard of a recording)
  std::string recName, recAlias; // alias: the name readFrameInformation renamed the file to

  vs_ctx* ctx;
  int device;
  int batchPackets;
  bool storePackets;
  bool fetchMeta;
  bool pipelined;
  uint8_t* ringPkts[2];
  int64_t* ringTimes[2];
  int fill;                          // ring being filled
  int64_t pending;
  bool pendingWrap;
  int hostLastAz;
  int64_t maxPosesCtx = 0;
  vs_carry carry;                    // parser state after the last batch handed to advance()
  uint32_t openCounts[HDL_MAX_NUM_LASERS];  // currentFrame->points[laser]->size() at that point
  Partial partial;                   // the open frame's points (host), after the last collect()
  std::deque<Batch> inflight;
  int64_t packetBase = 0;            // packets handed to the GPU since unloadData
  std::vector<int64_t> closedBy;     // for each frame closed since unloadData: the closing packet

  std::deque<std::shared_ptr<HDLFrame> > frames;
  std::shared_ptr<HDLFrame> currentFrame;
  std::shared_ptr<TransformManager> transMgr;
  std::shared_ptr<HDLManager> hdlMgr;

  CalibrationFile calib;
  bool correctionsInitialized;
  int calibFileReportedNumLasers;
  int numberOfTrailingFrames;
  int applyTransform;
  int pointsSkip;
  bool shouldCropReturns;
  bool shouldCropInside;
  double cropRegion[6];
  int laserSelections[HDL_MAX_NUM_LASERS];
  unsigned int dualReturnFilter;
  bool configDirty;
  bool warnedNoCalib;
  std::string error;
  std::vector<int64_t> poseT;
  std::vector<double> poseTrv;
};

// ---------------------------------------------------------------------------------------------
HDLParser::HDLParser() {
  this->internal_ = new vsInternal;
  this->unloadData();
}
HDLParser::~HDLParser() { delete this->internal_; }

const std::string& HDLParser::getDirName() { return this->dirName; }
void HDLParser::setDirName(const std::string& filename) {
  if (filename == this->dirName) return;
  this->dirName = filename;
  this->unloadData();
}

const std::string& HDLParser::getCorrectionsFile() { return this->correctionsFile; }
void HDLParser::setCorrectionsFile(const std::string& file) {
  if (file == this->correctionsFile) return;
  CalibrationFile c;
  std::string err;
  if (!c.load(file, &err)) {
    std::cerr << "Invalid sensor configuration file" << file << std::endl;
    return;
  }
  this->flush();  // packets buffered so far were received under the old calibration
  this->internal_->calib = c;
  this->internal_->calibFileReportedNumLasers = c.n_enabled;
  this->internal_->correctionsInitialized = true;
  this->internal_->configDirty = true;
  this->correctionsFile = file;
  this->unloadData();
}

void HDLParser::setNumberOfTrailingFrames(int n) { this->internal_->numberOfTrailingFrames = n; }

void HDLParser::setLaserSelection(int x00, int x01, int x02, int x03, int x04, int x05, int x06, int x07,
                                  int x08, int x09, int x10, int x11, int x12, int x13, int x14, int x15,
                                  int x16, int x17, int x18, int x19, int x20, int x21, int x22, int x23,
                                  int x24, int x25, int x26, int x27, int x28, int x29, int x30, int x31,
                                  int x32, int x33, int x34, int x35, int x36, int x37, int x38, int x39,
                                  int x40, int x41, int x42, int x43, int x44, int x45, int x46, int x47,
                                  int x48, int x49, int x50, int x51, int x52, int x53, int x54, int x55,
                                  int x56, int x57, int x58, int x59, int x60, int x61, int x62, int x63) {
  int mask[64] = {x00, x01, x02, x03, x04, x05, x06, x07, x08, x09, x10, x11, x12, x13, x14, x15,
                  x16, x17, x18, x19, x20, x21, x22, x23, x24, x25, x26, x27, x28, x29, x30, x31,
                  x32, x33, x34, x35, x36, x37, x38, x39, x40, x41, x42, x43, x44, x45, x46, x47,
                  x48, x49, x50, x51, x52, x53, x54, x55, x56, x57, x58, x59, x60, x61, x62, x63};
  this->setLaserSelection(mask);
}
void HDLParser::setLaserSelection(int sel[64]) {
  this->flush();
  for (int i = 0; i < 64; ++i) this->internal_->laserSelections[i] = sel[i] ? 1 : 0;
  this->internal_->configDirty = true;
}
void HDLParser::getLaserSelection(int sel[64]) {
  for (int i = 0; i < 64; ++i) sel[i] = this->internal_->laserSelections[i];
}
void HDLParser::getVerticalCorrections(double v[64]) {
  for (int i = 0; i < 64; ++i) v[i] = this->internal_->calib.rows[i].vert_correction_deg;
}

unsigned int HDLParser::getDualReturnFilter() const { return this->internal_->dualReturnFilter; }
void HDLParser::setDualReturnFilter(unsigned int f) { this->internal_->dualReturnFilter = f; }
void HDLParser::setPointsSkip(int pr) {
  this->flush();
  this->internal_->pointsSkip = pr < 0 ? 0 : pr;
  this->internal_->configDirty = true;
}
void HDLParser::setCropReturns(int crop) {
  this->flush();
  this->internal_->shouldCropReturns = !!crop;
  this->internal_->configDirty = true;
}
void HDLParser::setCropInside(int crop) {
  this->flush();
  this->internal_->shouldCropInside = !!crop;
  this->internal_->configDirty = true;
}
void HDLParser::setCropRegion(double region[6]) {
  this->flush();
  std::copy(region, region + 6, this->internal_->cropRegion);
  this->internal_->configDirty = true;
}
void HDLParser::setCropRegion(double xl, double xu, double yl, double yu, double zl, double zu) {
  double r[6] = {xl, xu, yl, yu, zl, zu};
  this->setCropRegion(r);
}

int HDLParser::getNumberOfChannels() { return this->internal_->calibFileReportedNumLasers; }

void HDLParser::unloadData() {
  vsInternal* in = this->internal_;
  if (!in->inflight.empty()) {
    // batches still on the GPU belong to the data being dropped: let them finish, discard them
    std::deque<std::shared_ptr<HDLFrame> > dropped;
    in->drainTo(&dropped);
  }
  in->pending = 0;
  in->pendingWrap = false;
  in->hostLastAz = -1;
  const int skip = in->carry.firing_skip;  // unloadData does not reset firingSkip (HDLParser.cxx:478-486)
  vs_carry_init(&in->carry);
  in->carry.firing_skip = skip;
  std::memset(in->openCounts, 0, sizeof(in->openCounts));
  in->partial = vsInternal::Partial();
  in->frames.clear();
  in->closedBy.clear();
  in->packetBase = 0;
  in->currentFrame = in->newFrameShell();
}

void HDLParser::processHDLPacket(unsigned char* data, unsigned int bytesReceived, ptime t) {
  if (bytesReceived != 1206) return;  // reference HDLParser.cxx:982-985
  vsInternal* in = this->internal_;
  if (!in->correctionsInitialized) {
    // the reference decodes with indeterminate corrections here; this facade refuses
    if (!in->warnedNoCalib) {
      std::cerr << "Corrections have not been set" << std::endl;
      in->warnedNoCalib = true;
    }
    in->error = "Corrections have not been set";
    return;
  }
  if (!in->ensureContext()) return;
  if (in->pending >= in->batchPackets) {
    // cannot happen while every flush path resets `pending`; never write past the ring
    in->error = "packet ring overrun: packet dropped";
    return;
  }
  // (streaming stores instead of this cached copy were measured 20 % slower end to end)
  std::memcpy(in->ringPkts[in->fill] + (size_t)in->pending * VS_PACKET_BYTES, data, VS_PACKET_BYTES);
  in->ringTimes[in->fill][in->pending] = t.us;
  ++in->pending;
  // an azimuth decrease anywhere in the packet is the only thing that can close a frame (a
  // pipelined parser never flushes on it: it hands frames back batch by batch)
  if (!in->pipelined) {
    for (int j = 0; j < HDL_FIRING_PER_PKT; ++j) {
      const int az = data[100 * j + 2] | (data[100 * j + 3] << 8);
      if (az < in->hostLastAz) in->pendingWrap = true;
      in->hostLastAz = az;
    }
  }
  if (in->pending >= in->batchPackets) {
    if (in->pipelined)
      in->pump();
    else
      in->decodePending(&in->frames);
  }
}

void HDLParser::flush() {
  vsInternal* in = this->internal_;
  if (in->pending != 0) {
    if (in->pipelined)
      in->pump();
    else
      in->decodePending(&in->frames);
  }
  in->drain();
}

const std::deque<std::shared_ptr<HDLFrame> >& HDLParser::getAllFrames() {
  // pipelined: frames appear when their batch has come back; nothing is forced
  if (this->internal_->pendingWrap && !this->internal_->pipelined) this->flush();
  return this->internal_->frames;
}
void HDLParser::clearAllFrames() { this->internal_->frames.clear(); }
std::shared_ptr<HDLFrame> HDLParser::createHDLFrame() { return this->internal_->createHDLFrame(); }

bool HDLParser::getFrame(std::shared_ptr<HDLFrame>& dest, const std::string& filename,
                         int64_t& startPos, const int& skip) {
  vsInternal* in = this->internal_;
  this->unloadData();
  const bool resident = in->hasRecording(filename);
  vtkPacketFileReader reader;
  if (!resident && !reader.open(filename)) {
    std::cerr << "failed to open packets file: " << filename << std::endl;
    return false;
  }
  if (!in->correctionsInitialized) {
    std::cerr << "Corrections have not been set" << std::endl;
    return false;
  }
  if (!in->ensureContext()) return false;
  if (!resident) reader.setFilePosition(&startPos);
  in->carry.firing_skip = skip;
  const unsigned char* data = nullptr;
  unsigned int len = 0;
  ptime t;
  std::deque<std::shared_ptr<HDLFrame> > closed;
  bool eof = false;
  int64_t cursor = resident ? (startPos - PCAP_GLOBAL_HEADER_LEN) / VS_PCAP_RECORD_BYTES - in->recFirst : 0;
  if (resident && (cursor < 0 || cursor >= in->recPackets)) {
    in->error = "getFrame: the frame starts outside the resident part of the recording";
    std::cerr << "HDLParser: " << in->error << std::endl;
    return false;
  }
  while (closed.empty() && !eof) {
    // one chunk: enough for a rotation of either sensor, decoded in one launch
    const int chunk = std::min(in->batchPackets, 512);
    if (resident) {
      // the rotation is decoded straight out of the recording in HBM
      const int64_t left = in->recPackets - cursor;
      if (left <= 0) break;
      in->pending = std::min<int64_t>(chunk, left);
      if (!in->decodePending(&closed, cursor)) return false;
      cursor += std::min<int64_t>(chunk, left);
      eof = cursor >= in->recPackets;
      continue;
    }
    while (in->pending < chunk) {
      if (!reader.nextPacket(data, len, t)) {
        eof = true;
        break;
      }
      if (len != 1206) continue;
      std::memcpy(in->ringPkts[in->fill] + (size_t)in->pending * VS_PACKET_BYTES, data, VS_PACKET_BYTES);
      in->ringTimes[in->fill][in->pending] = t.us;
      ++in->pending;
    }
    if (in->pending == 0) break;
    if (!in->decodePending(&closed)) return false;
  }
  if (!closed.empty()) {
    // the reference stops at the first packet that closes a frame and hands back
    // frames.back(): the last frame that packet closed (HDLParser.cxx:529-537)
    std::shared_ptr<HDLFrame> pick = closed.front();
    for (size_t i = 1; i < closed.size() && in->closedBy[i] == in->closedBy[0]; ++i) pick = closed[i];
    dest->points = std::move(pick->points);
    dest->pointsMeta = std::move(pick->pointsMeta);
    this->unloadData();
    return true;
  }
  // end of file: force the split (HDLParser.cxx:540-543).  The reference swaps the two frame
  // objects here, which leaves the caller's own frame (e.g. HDLManager's TimeLine entry)
  // empty and hands back an object only the local shared_ptr owns; the contents are moved
  // into the caller's frame instead, as in the branch above.
  in->materializeOpenFrame(in->carry.is_hdl64 != 0);
  dest->points = std::move(in->currentFrame->points);
  dest->pointsMeta = std::move(in->currentFrame->pointsMeta);
  this->unloadData();
  return true;
}

std::vector<std::shared_ptr<HDLFrame> > HDLParser::readFrameInformation(const std::string& name,
                                                                       bool touchOnly) {
  std::vector<std::shared_ptr<HDLFrame> > result;
  vsInternal* in = this->internal_;
  if (!touchOnly && in->hasRecording(name) && in->recPackets > 0) {
    // the index of a resident recording is one pass of the segmentation kernel over HBM
    const int64_t n = in->recPackets;
    int32_t cap = (int32_t)std::min<int64_t>(12 * n + 1, std::max<int64_t>(4096, n / 16));
    std::vector<int32_t> sp, sk;
    std::vector<int64_t> ts;
    int32_t nf = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      sp.assign((size_t)cap, 0);
      sk.assign((size_t)cap, 0);
      ts.assign((size_t)cap, 0);
      const uint8_t* dev = static_cast<const uint8_t*>(in->recDev) + VS_PCAP_PAYLOAD_OFFSET;
      const int rc = vs_read_frame_information(in->ctx, dev, VS_PCAP_RECORD_BYTES, nullptr, n,
                                               VS_FLAG_DEVICE_INPUT | VS_FLAG_PCAP_TIMES, sp.data(),
                                               sk.data(), ts.data(), cap, &nf);
      if (rc != VS_OK) {
        in->error = vs_last_error(in->ctx);
        std::cerr << "HDLParser: GPU frame index failed: " << in->error << std::endl;
        return result;
      }
      if (nf <= cap) break;
      cap = nf;
    }
    const ptime filenameTime(ts[0]);
    for (int32_t i = 0; i < nf; ++i) {
      std::shared_ptr<HDLFrame> f(new HDLFrame);
      f->fileStartPos = PCAP_GLOBAL_HEADER_LEN + (in->recFirst + (int64_t)sp[i]) * VS_PCAP_RECORD_BYTES;
      f->skips = (uint8_t)sk[i];
      f->isOnHardDrive = true;
      f->timestamp = ptime(ts[i]);
      f->filenameTime = filenameTime;
      result.push_back(f);
    }
    // files are named after their first packet's time (reference HDLParser.cxx:1080-1085, 1150-1158)
    std::string stem = name;
    const size_t slash = stem.find_last_of('/');
    const std::string dir = slash == std::string::npos ? std::string(".") : stem.substr(0, slash);
    if (slash != std::string::npos) stem = stem.substr(slash + 1);
    const size_t dot = stem.find_last_of('.');
    if (dot != std::string::npos) stem = stem.substr(0, dot);
    ptime nameTime;
    if (!from_iso_string(stem, &nameTime) && name == in->recName && in->recFirst == 0) {
      const std::string newname = dir + "/" + to_iso_string(filenameTime) + ".pcap";
      if (std::rename(name.c_str(), newname.c_str()) == 0) {
        in->recAlias = newname;
        std::cout << "The original filename: '" << name << "' was converted to: '" << newname << '\''
                  << std::endl;
      }
    }
    return result;
  }
  vtkPacketFileReader reader;
  if (!reader.open(name)) {
    std::cerr << "Failed to open packet file: " << name << std::endl << reader.getLastError() << std::endl;
    return result;
  }
  // file names are the ISO string of the first packet's time; others get renamed below
  std::string stem = name;
  const size_t slash = stem.find_last_of('/');
  const std::string dir = slash == std::string::npos ? std::string(".") : stem.substr(0, slash);
  if (slash != std::string::npos) stem = stem.substr(slash + 1);
  const size_t dot = stem.find_last_of('.');
  if (dot != std::string::npos) stem = stem.substr(0, dot);
  ptime nameTime;
  const bool nameIsTime = from_iso_string(stem, &nameTime);
  if (!nameIsTime)
    std::cout << "Filename is not a valid ptime string, I'll convert it into a ptime string.\n";

  const unsigned char* data = nullptr;
  unsigned int len = 0;
  ptime packetTime, filenameTime;
  unsigned int lastAzimuth = 0;
  int64_t lastFilePosition = 0;
  result.push_back(std::shared_ptr<HDLFrame>(new HDLFrame));
  reader.getFilePosition(&lastFilePosition);
  result.back()->fileStartPos = lastFilePosition;
  result.back()->skips = 0;
  result.back()->isOnHardDrive = true;
  while (reader.nextPacket(data, len, packetTime)) {
    if (len != 1206) continue;
    if (filenameTime.is_special()) {
      filenameTime = packetTime;
      result.back()->timestamp = packetTime;
      if (touchOnly) break;
    }
    for (int i = 0; i < HDL_FIRING_PER_PKT; ++i) {
      const unsigned int rot = data[100 * i + 2] | (data[100 * i + 3] << 8);
      if (rot < lastAzimuth) {
        result.push_back(std::shared_ptr<HDLFrame>(new HDLFrame));
        result.back()->fileStartPos = lastFilePosition;
        result.back()->skips = (uint8_t)i;
        result.back()->isOnHardDrive = true;
        result.back()->timestamp = packetTime;
      }
      lastAzimuth = rot;
    }
    reader.getFilePosition(&lastFilePosition);
  }
  for (auto& f : result) f->filenameTime = filenameTime;
  reader.close();
  if (!nameIsTime && !filenameTime.is_special()) {
    const std::string newname = dir + "/" + to_iso_string(filenameTime) + ".pcap";
    if (std::rename(name.c_str(), newname.c_str()) == 0)
      std::cout << "The original filename: '" << name << "' was converted to: '" << newname << '\''
                << std::endl;
  }
  return result;
}

void HDLParser::setHDLManager(std::shared_ptr<HDLManager> p) { this->internal_->hdlMgr = p; }

std::shared_ptr<TransformManager> HDLParser::getTransformMgr() const { return this->internal_->transMgr; }
void HDLParser::setTransformMgr(std::shared_ptr<TransformManager> mgr) {
  this->flush();
  this->internal_->transMgr = mgr;
}
int HDLParser::getApplyTransform() { return this->internal_->applyTransform; }
void HDLParser::setApplyTransform(int apply) { this->internal_->applyTransform = apply; }

bool HDLParser::loadRecording(const std::string& pcapfile) { return this->internal_->loadRecording(pcapfile); }
bool HDLParser::loadRecordingRange(const std::string& pcapfile, int64_t firstRecord, int64_t nRecords) {
  return this->internal_->loadRecording(pcapfile, firstRecord, nRecords);
}
void HDLParser::recordingRange(int64_t* firstRecord, int64_t* nRecords) const {
  *firstRecord = this->internal_->recFirst;
  *nRecords = this->internal_->recPackets;
}
void HDLParser::unloadRecording() { this->internal_->unloadRecording(); }
bool HDLParser::hasRecording(const std::string& pcapfile) const { return this->internal_->hasRecording(pcapfile); }

void HDLParser::setDevice(int d) { this->internal_->device = d; }

// Everything that is slow the first time, done before the first packet instead of under it:
// the CUDA context and its buffers, the first launch of every kernel of the path (a throw-away
// batch of empty packets with a scratch carry: the parser's own state is untouched), and
// `frames` page-locked arenas of `pointsPerFrame` points in the pool.
bool HDLParser::prepare(int frames, size_t pointsPerFrame) {
  vsInternal* in = this->internal_;
  if (!in->ensureContext()) return false;
  if (in->correctionsInitialized && in->inflight.empty() && in->pending == 0) {
    if (!in->syncConfig(0, 0)) return false;
    const int n = std::min(in->batchPackets, 64);
    std::memset(in->ringPkts[in->fill], 0, (size_t)n * VS_PACKET_BYTES);
    for (int i = 0; i < n; ++i) in->ringTimes[in->fill][i] = i;
    vs_carry scratch;
    vs_carry_init(&scratch);
    uint64_t ticket = 0;
    vs_result r;
    vs_layout lay;
    if (vs_submit(in->ctx, in->ringPkts[in->fill], VS_PACKET_BYTES, in->ringTimes[in->fill], n, 0,
                  VS_MODE_STREAMING, 0, 0, &scratch, &ticket) != VS_OK ||
        vs_wait(in->ctx, ticket, &r) != VS_OK ||
        vs_layout_frames(in->ctx, ticket, nullptr, 16, 1, &lay) != VS_OK ||
        vs_sync(in->ctx, ticket, nullptr) != VS_OK) {
      in->error = vs_last_error(in->ctx);
      return false;
    }
  }
  std::vector<std::shared_ptr<vs::Arena> > hold;
  for (int i = 0; i < frames; ++i) {
    const size_t metaOff = (pointsPerFrame * sizeof(pcl::PointXYZI) + 255) & ~(size_t)255;
    hold.push_back(vs::Arena::acquire(metaOff + pointsPerFrame * sizeof(PointMeta)));
    if (!hold.back()) return false;
  }
  return true;  // `hold` goes back to the pool here
}
void HDLParser::setBatchPackets(int n) {
  if (!this->internal_->ctx && n > 0) this->internal_->batchPackets = n;
}
void HDLParser::setPipelined(bool on) {
  if (!this->internal_->ctx) this->internal_->pipelined = on;
}
void HDLParser::setFetchMeta(bool on) {
  this->flush();
  this->internal_->fetchMeta = on;
}
void HDLParser::setStorePackets(bool s) { this->internal_->storePackets = s; }
const std::string& HDLParser::lastError() const { return this->internal_->error; }
void HDLParser::resetPipelineStats() {
  vsInternal* in = this->internal_;
  in->secAdvance = in->secCollect = in->secSubmit = in->secWait = in->secSync = 0;
  in->nPumps = 0;
}
std::string HDLParser::pipelineStats() const {
  const vsInternal* in = this->internal_;
  char buf[256];
  std::snprintf(buf, sizeof(buf),
                "{\"pumps\": %ld, \"advance_s\": %.4f, \"collect_s\": %.4f, \"submit_s\": %.4f, "
                "\"in_vs_wait_s\": %.4f, \"in_vs_sync_s\": %.4f}",
                in->nPumps, in->secAdvance, in->secCollect, in->secSubmit, in->secWait, in->secSync);
  return buf;
}
