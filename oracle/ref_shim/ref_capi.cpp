/*
 * ref_capi.cpp -- C interface over the REFERENCE's own HDLParser / TransformManager.
 *
 * TEST INFRASTRUCTURE.  This translation unit includes /root/reference/HDLParser.cxx verbatim
 * (so that the parser's private state can be observed) and is linked with the reference's
 * TransformManager.cxx, type_defs.cxx, HDLFrame.cxx, vtkPacketFileWriter.cxx and
 * CoordiTran.cpp, all compiled where they lie against the stand-in headers of this directory
 * (oracle/build_ref.py).  Output: oracle/_ref/libvelo_ref.so.  Nothing here restates the
 * reference: it only marshals arguments.
 */
#include "HDLParser.cxx"  // found through -I /root/reference

#include <cstdint>
#include <cstring>
#include <limits>

#include "vtkPacketFileWriter.h"

namespace {

const int64_t kNone = std::numeric_limits<int64_t>::min();

struct RefParser : public HDLParser {
  vsInternal* in() { return this->internal_; }
  void unload() { this->unloadData(); }
};

inline ptime us_to_ptime(int64_t us) { return ptime::from_us(us); }
inline int64_t ptime_to_us(const ptime& t) { return t.is_special() ? kNone : t.us(); }

}  // namespace

struct vr_parser {
  RefParser parser;
  boost::shared_ptr<TransformManager> tm;
  std::deque<boost::shared_ptr<HDLFrame> > frames;  // snapshot handed out by vr_num_frames
  vr_parser() : tm(new TransformManager) { parser.setTransformMgr(tm); }
};

struct vr_frame_info {
  int64_t timestamp_us;
  int32_t skips;
  int32_t n_lasers;
  int32_t n_points;
  int32_t n_packets;
  int32_t is_hdl64_order;
  int32_t pad;
  double carpose_TRV[9];
  double carpose_seconds_pos;
};

#include "CoordiTran.h"
#include "TimeSolver.h"

extern "C" {

vr_parser* vr_create(void) { return new vr_parser; }
void vr_destroy(vr_parser* h) { delete h; }

void vr_set_corrections_file(vr_parser* h, const char* path) { h->parser.setCorrectionsFile(path); }
int32_t vr_num_channels(vr_parser* h) { return h->parser.getNumberOfChannels(); }
void vr_set_laser_selection(vr_parser* h, const int32_t sel[64]) {
  int s[64];
  for (int i = 0; i < 64; ++i) s[i] = sel[i];
  h->parser.setLaserSelection(s);
}
void vr_set_points_skip(vr_parser* h, int32_t n) { h->parser.setPointsSkip(n); }
void vr_set_crop(vr_parser* h, int32_t crop_returns, int32_t crop_inside, const double region[6]) {
  h->parser.setCropReturns(crop_returns);
  h->parser.setCropInside(crop_inside);
  double r[6];
  for (int i = 0; i < 6; ++i) r[i] = region[i];
  h->parser.setCropRegion(r);
}

void vr_clear_poses(vr_parser* h) { h->tm->clearTransforms(); }
void vr_add_pose(vr_parser* h, int64_t t_us, const double T[3], const double R[3], const double V[3]) {
  boost::shared_ptr<PoseTransform> p(new PoseTransform);
  for (int i = 0; i < 3; ++i) {
    p->T[i] = T[i];
    p->R[i] = R[i];
    p->V[i] = V[i];
  }
  p->timestamp = us_to_ptime(t_us);
  p->seconds_pos = 0;
  h->tm->addTransform(p);
}
int32_t vr_num_poses(vr_parser* h) { return h->tm->getNumberOfTransforms(); }
int32_t vr_interpolate(vr_parser* h, int64_t t_us, double out[9], double* seconds_pos) {
  PoseTransform tr;
  ptime t = us_to_ptime(t_us);
  const bool ok = h->tm->interpolateTransform(t, &tr);
  for (int i = 0; i < 3; ++i) {
    out[i] = tr.T[i];
    out[3 + i] = tr.R[i];
    out[6 + i] = tr.V[i];
  }
  *seconds_pos = tr.seconds_pos;
  return ok ? 1 : 0;
}
void vr_pose_matrix(const double TRV[9], double out[12]) {
  PoseTransform p;
  for (int i = 0; i < 3; ++i) {
    p.T[i] = TRV[i];
    p.R[i] = TRV[3 + i];
    p.V[i] = TRV[6 + i];
  }
  Eigen::Affine3d a = p.getMatrix();
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) out[4 * r + c] = a(r, c);
}

void vr_unload(vr_parser* h) { h->parser.unload(); }
void vr_get_state(vr_parser* h, int32_t out[4]) {
  out[0] = h->parser.in()->lastAzimuth;
  out[1] = h->parser.in()->firingSkip;
  out[2] = h->parser.in()->frameMetaInited ? 1 : 0;
  out[3] = h->parser.in()->isHDL64Data ? 1 : 0;
}

void vr_process_packets(vr_parser* h, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us, const int32_t* lengths) {
  for (int64_t i = 0; i < n; ++i)
    h->parser.processHDLPacket(const_cast<unsigned char*>(data + i * stride),
                               lengths ? (unsigned)lengths[i] : 1206u, us_to_ptime(t_us[i]));
}
/* The reference's consumer loop, PacketConsumer::handleSensorData (HDLSource.cxx:209-225), for n
 * packets: processHDLPacket, then -- as the reference does after EVERY packet --
 *     if (getAllFrames().size()) { hdlMgr->addFrame(getAllFrames().back()); clearAllFrames(); }
 * so frames leave the parser as they close instead of piling up.  addFrame stores the pointer
 * (HDLManager.cxx:176-192); here the frame is counted and released, which is what happens to it
 * once the manager's cache evicts it.  out[0] += frames taken, out[1] += their points. */
void vr_consume_packets(vr_parser* h, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us, int64_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    h->parser.processHDLPacket(const_cast<unsigned char*>(data + i * stride), 1206u, us_to_ptime(t_us[i]));
    if (h->parser.getAllFrames().size()) {
      boost::shared_ptr<HDLFrame> f = h->parser.getAllFrames().back();
      int64_t pts = 0;
      for (auto& c : f->points)
        if (c) pts += (int64_t)c->points.size();
      out[0] += 1;
      out[1] += pts;
      h->parser.clearAllFrames();
    }
  }
}
void vr_split_frame(vr_parser* h) { h->parser.in()->splitFrame(); }

int32_t vr_num_frames(vr_parser* h) {
  h->frames = h->parser.getAllFrames();
  return (int32_t)h->frames.size();
}
void vr_clear_frames(vr_parser* h) {
  h->parser.clearAllFrames();
  h->frames.clear();
}
int64_t vr_open_frame_points(vr_parser* h) {
  int64_t n = 0;
  for (auto& c : h->parser.in()->currentFrame->points) n += (int64_t)c->points.size();
  return n;
}

static const HDLFrame* frame_at(vr_parser* h, int32_t f) {
  if (f < 0 || f >= (int32_t)h->frames.size()) return nullptr;
  return h->frames[f].get();
}
int32_t vr_frame_get_info(vr_parser* h, int32_t f, vr_frame_info* out) {
  const HDLFrame* fr = frame_at(h, f);
  if (!fr) return 0;
  out->timestamp_us = ptime_to_us(fr->timestamp);
  out->skips = (int32_t)fr->skips;
  out->n_lasers = (int32_t)fr->points.size();
  int64_t n = 0;
  for (auto& c : fr->points)
    if (c) n += (int64_t)c->points.size();
  out->n_points = (int32_t)n;
  out->n_packets = (int32_t)fr->packets.size();
  out->is_hdl64_order = 0;
  out->pad = 0;
  for (int i = 0; i < 3; ++i) {
    out->carpose_TRV[i] = fr->carpose->T[i];
    out->carpose_TRV[3 + i] = fr->carpose->R[i];
    out->carpose_TRV[6 + i] = fr->carpose->V[i];
  }
  out->carpose_seconds_pos = fr->carpose->seconds_pos;
  return 1;
}
int32_t vr_frame_laser_counts(vr_parser* h, int32_t f, int32_t* counts) {
  const HDLFrame* fr = frame_at(h, f);
  if (!fr) return 0;
  for (size_t i = 0; i < fr->points.size(); ++i)
    counts[i] = fr->points[i] ? (int32_t)fr->points[i]->points.size() : 0;
  return (int32_t)fr->points.size();
}
int32_t vr_frame_points(vr_parser* h, int32_t f, float* xyzi, uint16_t* azimuth, float* distance) {
  const HDLFrame* fr = frame_at(h, f);
  if (!fr) return 0;
  int64_t k = 0;
  for (size_t l = 0; l < fr->points.size(); ++l) {
    if (!fr->points[l]) continue;
    const auto& pts = fr->points[l]->points;
    const auto& meta = *fr->pointsMeta[l];
    for (size_t i = 0; i < pts.size(); ++i, ++k) {
      xyzi[4 * k + 0] = pts[i].x;
      xyzi[4 * k + 1] = pts[i].y;
      xyzi[4 * k + 2] = pts[i].z;
      xyzi[4 * k + 3] = pts[i].intensity;
      azimuth[k] = meta[i].azimuth;
      distance[k] = meta[i].distance;
    }
  }
  return (int32_t)k;
}

/* HDLParser::readFrameInformation on a pcap file (the file name must be an ISO time string or
 * the reference renames the file). */
int32_t vr_read_frame_information(vr_parser* h, const char* filename, int64_t* file_pos,
                                  int32_t* skips, int64_t* timestamp_us, int32_t cap) {
  std::vector<boost::shared_ptr<HDLFrame> > v = h->parser.readFrameInformation(filename, false);
  for (size_t i = 0; i < v.size() && (int32_t)i < cap; ++i) {
    file_pos[i] = *(reinterpret_cast<long long*>(&v[i]->fileStartPos));
    skips[i] = (int32_t)v[i]->skips;
    timestamp_us[i] = ptime_to_us(v[i]->timestamp);
  }
  return (int32_t)v.size();
}
/* HDLParser::getFrame: the decoded frame becomes frame 0 of the snapshot. */
int32_t vr_get_frame(vr_parser* h, const char* filename, int64_t file_pos, int32_t skip) {
  boost::shared_ptr<HDLFrame> dest(new HDLFrame);
  fpos_t pos;
  std::memset(&pos, 0, sizeof(pos));
  NUM_TO_FPOS_T(pos, file_pos);
  const bool ok = h->parser.getFrame(dest, filename, pos, skip);
  h->frames.clear();
  if (ok) h->frames.push_back(dest);
  return ok ? 1 : 0;
}
int32_t vr_num_snapshot_frames(vr_parser* h) { return (int32_t)h->frames.size(); }

/* vtkPacketFileWriter::writePacket for n payloads. */
int32_t vr_write_pcap(const char* filename, const uint8_t* data, int64_t n, int64_t stride,
                      const int64_t* t_us) {
  vtkPacketFileWriter w;
  if (!w.open(filename)) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (!w.writePacket(data + i * stride, 1206, us_to_ptime(t_us[i]))) return 0;
  w.close();
  return 1;
}

/* ---- N3: TimeSolver.cxx / CoordiTran.cpp / INSSource.cxx:300-326 ------------------------- */
void vr_set_fake_now(int64_t us) { boost::shim_clock::fake_now_us() = us; }
void* vr_ts_create(void) { return new TimeSolver; }
void vr_ts_destroy(void* ts) { delete static_cast<TimeSolver*>(ts); }
int64_t vr_ts_hdl(void* ts, uint32_t microsec_to_hour) {
  return ptime_to_us(static_cast<TimeSolver*>(ts)->calcTimestamp(microsec_to_hour));
}
int64_t vr_ts_ins(void* ts, const InsPVA* rec) {
  return ptime_to_us(static_cast<TimeSolver*>(ts)->calcTimestamp(rec));
}
void vr_llh2enu(const double llh[3], const double orgxyz[3], double enu[3]) {
  double a[3] = {llh[0], llh[1], llh[2]}, o[3] = {orgxyz[0], orgxyz[1], orgxyz[2]};
  llh2enu(a, o, enu);
}
void vr_llh2xyz(const double llh[3], double xyz[3]) {
  double a[3] = {llh[0], llh[1], llh[2]};
  llh2xyz(a, xyz);
}
int32_t vr_sizeof_inspva(void) { return (int32_t)sizeof(InsPVA); }

} /* extern "C" */
