/*
 * velo_oracle.cpp -- CPU oracle for the VeloSLAM ingest hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see velo_oracle.h).  A dependency-free C++17 restatement
 * of what the reference does for: packet decode + per-laser calibration, rotation
 * segmentation into frames, pose-timeline bracket lookup + Euler/translation lerp and the
 * per-packet rigid transform.  Each function cites the reference file:line it follows
 * (paths are relative to /root/reference).  Bug-compatible on purpose.
 *
 * Third-party arithmetic that is not under /root/reference and is restated here:
 *   - Boost.DateTime ptime / time_duration (version unpinned by the reference's CMake):
 *     microsecond-resolution integer arithmetic  -> int64 microseconds.
 *   - boost::circular_buffer<shared_ptr<T>>(5) as used by TimeLine.h -> RingOf5 below.
 *   - Eigen3 (unpinned) Affine3d::rotate(AngleAxisd) / translation() as used by
 *     type_defs.h:134-146 -> AngleAxis::toRotationMatrix (Rodrigues form, as published in
 *     Eigen 3.x Geometry/AngleAxis.h) followed by plain 3x3 products.
 *   - PCL PointXYZI / PointCloud: containers only, no arithmetic.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off (the reference is x86-64 SSE2 code with no
 * FMA contraction; keep mul and add separate so doubles match before the float cast).
 */
#include "velo_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace vo {

typedef int64_t usec_t;
static const usec_t NOT_A_DATE_TIME = std::numeric_limits<int64_t>::min();

/* type_defs.h:16-20 */
static const int kNumRotAngles = 36001;
static const int kLaserPerFiring = 32;
static const int kMaxNumLasers = 64;
static const int kFiringPerPkt = 12;
static const int kMaxPtsPerLaser = 2200;

/* type_defs.h:25 TO_RADIUS, HDLParser.cxx:59 HDL_Grabber_toRadians: (x * M_PI) / 180 */
static inline double toRadians(double x) { return (x * M_PI) / 180.0; }

/* ---------------------------------------------------------------------------------
 * PoseTransform (type_defs.h:86-147, ctor type_defs.cxx:47-57)
 * ------------------------------------------------------------------------------- */
struct PoseTransform {
  double T[3];
  double R[3];
  double V[3];
  usec_t timestamp;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  double seconds_pos;

  PoseTransform() {
    for (int i = 0; i < 3; ++i) {
      T[i] = 0;
      R[i] = 0;
      V[i] = 0;
    }
    timestamp = NOT_A_DATE_TIME;
    week_number = 0;
    milliseconds = week_number_pos = 0;
    seconds_pos = -1; /* -1 == not a valid pose */
  }
  /* type_defs.h:102-131: component-wise on T, R, V; the result is a fresh default pose
   * (timestamp not_a_date_time, seconds_pos -1). */
  PoseTransform plus(const PoseTransform& d) const {
    PoseTransform r;
    for (int i = 0; i < 3; ++i) {
      r.T[i] = T[i] + d.T[i];
      r.R[i] = R[i] + d.R[i];
      r.V[i] = V[i] + d.V[i];
    }
    return r;
  }
  PoseTransform minus(const PoseTransform& d) const {
    PoseTransform r;
    for (int i = 0; i < 3; ++i) {
      r.T[i] = T[i] - d.T[i];
      r.R[i] = R[i] - d.R[i];
      r.V[i] = V[i] - d.V[i];
    }
    return r;
  }
  PoseTransform times(double ratio) const {
    PoseTransform r;
    for (int i = 0; i < 3; ++i) {
      r.T[i] = T[i] * ratio;
      r.R[i] = R[i] * ratio;
      r.V[i] = V[i] * ratio;
    }
    return r;
  }
};

/* Eigen::Affine3d restated as [L | t] */
struct Affine3 {
  double L[3][3];
  double t[3];
};

/* Eigen 3.x AngleAxis<double>::toRotationMatrix() for a unit axis. */
static void angleAxisMatrix(double angle, const double axis[3], double res[3][3]) {
  const double s = std::sin(angle);
  const double c = std::cos(angle);
  const double sin_axis[3] = {s * axis[0], s * axis[1], s * axis[2]};
  const double cos1_axis[3] = {(1.0 - c) * axis[0], (1.0 - c) * axis[1], (1.0 - c) * axis[2]};
  double tmp;
  tmp = cos1_axis[0] * axis[1];
  res[0][1] = tmp - sin_axis[2];
  res[1][0] = tmp + sin_axis[2];
  tmp = cos1_axis[0] * axis[2];
  res[0][2] = tmp + sin_axis[1];
  res[2][0] = tmp - sin_axis[1];
  tmp = cos1_axis[1] * axis[2];
  res[1][2] = tmp - sin_axis[0];
  res[2][1] = tmp + sin_axis[0];
  res[0][0] = cos1_axis[0] * axis[0] + c;
  res[1][1] = cos1_axis[1] * axis[1] + c;
  res[2][2] = cos1_axis[2] * axis[2] + c;
}

static void matmul3(const double a[3][3], const double b[3][3], double out[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}

/* type_defs.h:134-146 getMatrix(): identity, then rotate(Y,R0), rotate(X,R1), rotate(Z,R2)
 * (rotate post-multiplies the linear part), then translation = T. */
static Affine3 poseMatrix(const PoseTransform& p) {
  static const double UY[3] = {0, 1, 0}, UX[3] = {1, 0, 0}, UZ[3] = {0, 0, 1};
  double L[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double Rm[3][3], tmp[3][3];
  angleAxisMatrix(toRadians(p.R[0]), UY, Rm);
  matmul3(L, Rm, tmp);
  std::memcpy(L, tmp, sizeof(L));
  angleAxisMatrix(toRadians(p.R[1]), UX, Rm);
  matmul3(L, Rm, tmp);
  std::memcpy(L, tmp, sizeof(L));
  angleAxisMatrix(toRadians(p.R[2]), UZ, Rm);
  matmul3(L, Rm, tmp);
  std::memcpy(L, tmp, sizeof(L));
  Affine3 a;
  std::memcpy(a.L, L, sizeof(L));
  a.t[0] = p.T[0];
  a.t[1] = p.T[1];
  a.t[2] = p.T[2];
  return a;
}

/* type_defs.h:160-166 transformPoint: rows summed left to right, translation last. */
static void transformPoint(double pt0[3], const Affine3& m) {
  const double px = pt0[0], py = pt0[1], pz = pt0[2];
  pt0[0] = m.L[0][0] * px + m.L[0][1] * py + m.L[0][2] * pz + m.t[0];
  pt0[1] = m.L[1][0] * px + m.L[1][1] * py + m.L[1][2] * pz + m.t[1];
  pt0[2] = m.L[2][0] * px + m.L[2][1] * py + m.L[2][2] * pz + m.t[2];
}

/* ---------------------------------------------------------------------------------
 * boost::circular_buffer<shared_ptr<T>> with capacity 5, restricted to what TimeLine uses.
 * ------------------------------------------------------------------------------- */
template <typename P>
struct RingOf5 {
  std::deque<P> d;
  static const size_t cap = 5;
  size_t size() const { return d.size(); }
  void clear() { d.clear(); }
  P& operator[](size_t i) { return d[i]; }
  P& back() { return d.back(); }
  void push_back(const P& v) { /* full: overwrites the oldest (front) element */
    if (d.size() == cap) d.pop_front();
    d.push_back(v);
  }
  void push_front(const P& v) { /* full: the last element is removed */
    if (d.size() == cap) d.pop_back();
    d.push_front(v);
  }
  void insert(size_t pos, const P& v) {
    /* full: the first element is overwritten; full and pos==begin(): nothing inserted */
    if (d.size() == cap) {
      if (pos == 0) return;
      d.pop_front();
      --pos;
    }
    d.insert(d.begin() + pos, v);
  }
};

/* ---------------------------------------------------------------------------------
 * TimeLine<T> (TimeLine.h): time-bucketed container.  Only addData (140-226),
 * getBoundaryData (384-468), getAll (498-508), calcNewAvgIntervalAndRearrange (536-552).
 * ------------------------------------------------------------------------------- */
template <typename T_>
class TimeLine {
 public:
  typedef std::shared_ptr<T_> Ptr;
  TimeLine() : interval(0), intervalFinalized(false), total_num(0) {
    startTime = maxTime = NOT_A_DATE_TIME;
  }
  size_t size() const { return total_num; }
  void clear() { /* TimeLine.h:105-112 (intervalFinalized is NOT reset there) */
    timeline.clear();
    buffer.clear();
    startTime = maxTime = NOT_A_DATE_TIME;
    interval = 0;
    total_num = 0;
  }

  void addData(Ptr data) {
    const usec_t t = data->timestamp;
    if (timeline.empty()) {
      startTime = t;
      maxTime = t;
      timeline.push_back(std::vector<Ptr>());
      timeline.back().push_back(data);
      buffer.push_back(data);
      ++total_num;
    } else if (timeline.size() == 1) {
      if (t == startTime) {
        timeline.back().back() = data;
        buffer[0] = data;
        return;
      } else if (t < startTime) {
        interval = (startTime - t) * 0.95;
        timeline.insert(timeline.begin(), std::vector<Ptr>());
        timeline[0].push_back(data);
        buffer.push_front(data);
        maxTime = startTime;
        startTime = t;
        ++total_num;
        return;
      } else {
        interval = (t - startTime) * 0.95;
        timeline.push_back(std::vector<Ptr>());
        timeline.back().push_back(data);
        buffer.push_back(data);
        maxTime = t;
        ++total_num;
      }
    } else {
      if ((!intervalFinalized) && total_num == 10) calcNewAvgIntervalAndRearrange();
      if (t > buffer.back()->timestamp) {
        buffer.push_back(data);
      } else {
        int cursor = 0;
        while (buffer[cursor]->timestamp < t) ++cursor;
        if (buffer[cursor]->timestamp == t) {
          buffer[cursor] = data;
        } else {
          buffer.insert(cursor, data);
        }
      }
      int index = (int)((t - startTime) / interval); /* long / double -> double -> int */
      if (t >= startTime) {
        while ((int)timeline.size() <= index) timeline.push_back(std::vector<Ptr>());
        size_t insertPos = 0;
        while (insertPos < timeline[index].size() && timeline[index][insertPos]->timestamp < t)
          ++insertPos;
        if (insertPos == timeline[index].size()) {
          timeline[index].push_back(data);
          maxTime = t;
          ++total_num;
        } else if (timeline[index][insertPos]->timestamp == t) {
          timeline[index][insertPos] = data;
        } else {
          timeline[index].insert(timeline[index].begin() + insertPos, data);
          ++total_num;
        }
      } else {
        index = (int)std::floor((t - startTime) / interval);
        while ((index++) != 0) timeline.insert(timeline.begin(), std::vector<Ptr>());
        timeline.front().push_back(data);
        startTime = t;
        ++total_num;
      }
    }
  }

  std::pair<Ptr, Ptr> getBoundaryData(usec_t t) {
    if (timeline.empty()) return std::make_pair(Ptr(), Ptr());
    Ptr forward, backward;
    if (timeline.size() == 1) {
      forward = timeline[0][0];
      return std::make_pair(forward, backward);
    }
    if (t <= startTime) {
      forward = timeline[0][0];
      if (timeline[0].size() > 1)
        backward = timeline[0][1];
      else
        backward = timeline[1][0];
      return std::make_pair(forward, backward);
    }
    if (t >= maxTime) {
      backward = buffer.back();
      forward = buffer[buffer.size() - 2];
      return std::make_pair(forward, backward);
    }
    if (t > buffer[0]->timestamp) {
      int index = 1;
      while (buffer[index]->timestamp < t) ++index;
      return std::make_pair(buffer[index - 1], buffer[index]);
    }
    int index = (int)((t - startTime) / interval);
    if (timeline[index].size() != 0) {
      if (timeline[index][0]->timestamp <= t) {
        forward = timeline[index][0];
        for (size_t i = 1; i < timeline[index].size(); ++i) {
          if (timeline[index][i]->timestamp < t) {
            forward = timeline[index][i];
          } else {
            backward = timeline[index][i];
            break;
          }
        }
        if (!backward) {
          size_t cursor = index + 1;
          while (cursor != timeline.size() && timeline[cursor].empty()) ++cursor;
          if (cursor != timeline.size()) backward = timeline[cursor].front();
        }
        if (timeline[index][0]->timestamp == t) {
          int cursor = index - 1;
          while (cursor >= 0 && timeline[cursor].empty()) --cursor;
          if (cursor != -1) {
            Ptr anotherPossible = timeline[cursor].back();
            if (backward) {
              const usec_t diff_f = t - backward->timestamp;
              const usec_t diff_b = anotherPossible->timestamp - t;
              if (diff_f > diff_b) {
                backward = forward;
                forward = anotherPossible;
              }
            } else {
              backward = forward;
              forward = anotherPossible;
            }
          }
        }
      } else {
        backward = timeline[index][0];
        int cursor = index - 1;
        while (cursor >= 0 && timeline[cursor].empty()) --cursor;
        forward = timeline[cursor].back();
      }
    } else {
      int cursor = index - 1;
      while (cursor >= 0 && timeline[cursor].empty()) --cursor;
      forward = timeline[cursor].back();
      while (timeline[++index].empty()) {
      }
      backward = timeline[index].front();
    }
    return std::make_pair(forward, backward);
  }

  std::vector<Ptr> getAll() {
    std::vector<Ptr> result;
    if (total_num == 0) return result;
    for (size_t i = 0; i < timeline.size(); ++i)
      if (!timeline[i].empty()) result.insert(result.end(), timeline[i].begin(), timeline[i].end());
    return result;
  }

 private:
  void calcNewAvgIntervalAndRearrange() {
    /* long / size_t: the long is converted to unsigned long, integer division */
    interval = (double)((uint64_t)(maxTime - startTime) / (uint64_t)total_num);
    std::vector<Ptr> vec = getAll();
    timeline.clear();
    buffer.clear();
    for (size_t i = 0; i < vec.size(); ++i) {
      int index = (int)((vec[i]->timestamp - startTime) / interval);
      while ((int)timeline.size() <= index) timeline.push_back(std::vector<Ptr>());
      timeline[index].push_back(vec[i]);
      buffer.push_back(vec[i]);
    }
    intervalFinalized = true;
  }

  std::vector<std::vector<Ptr>> timeline;
  RingOf5<Ptr> buffer;
  usec_t startTime, maxTime;
  double interval;
  bool intervalFinalized;
  size_t total_num;
};

/* ---------------------------------------------------------------------------------
 * TransformManager (TransformManager.cxx:73-79, 149-177)
 * ------------------------------------------------------------------------------- */
class TransformManager {
 public:
  void clearTransforms() { transforms.clear(); }
  int getNumberOfTransforms() { return (int)transforms.size(); }
  void addTransform(std::shared_ptr<PoseTransform> trans) { transforms.addData(trans); }

  bool interpolateTransform(usec_t t, PoseTransform* trans) {
    trans->timestamp = t;
    auto bound = transforms.getBoundaryData(t);
    if ((!bound.first) && (!bound.second)) {
      return false;
    } else if (!bound.second) {
      PoseTransform& fore = *bound.first;
      /* long / float: the long is converted to float, float division (1e6f) */
      double sec = (float)(t - fore.timestamp) / 1e6f;
      for (int i = 0; i < 3; ++i) {
        trans->V[i] = fore.V[i];
        trans->R[i] = fore.R[i];
        trans->T[i] = fore.T[i] + fore.V[i] * sec;
      }
      return true; /* seconds_pos untouched: still "invalid" for the parser */
    } else {
      PoseTransform& fore = *bound.first;
      PoseTransform& back = *bound.second;
      const usec_t diff = t - fore.timestamp;
      double ratio = double(diff) / (back.timestamp - fore.timestamp);
      *trans = fore.plus(back.minus(fore).times(ratio));
      trans->seconds_pos = 0;
      return true;
    }
  }

 private:
  TimeLine<PoseTransform> transforms;
};

/* ---------------------------------------------------------------------------------
 * Wire layout (HDLParser.cxx:61-87, #pragma pack(1)); read through memcpy-free byte
 * accessors so the oracle has no alignment assumptions.
 * ------------------------------------------------------------------------------- */
static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static const uint16_t BLOCK_0_TO_31 = 0xeeff;

struct FiringView {
  const uint8_t* p; /* 100 bytes */
  uint16_t blockIdentifier() const { return rd16(p); }
  uint16_t rotationalPosition() const { return rd16(p + 2); }
  uint16_t distance(int dsr) const { return rd16(p + 4 + 3 * dsr); }
  uint8_t intensity(int dsr) const { return p[4 + 3 * dsr + 2]; }
};

/* HDLParser.cxx:89-100 */
struct LaserCorrection {
  double azimuthCorrection;
  double verticalCorrection;
  double distanceCorrection;
  double verticalOffsetCorrection;
  double horizontalOffsetCorrection;
  double sinVertCorrection;
  double cosVertCorrection;
  double sinVertOffsetCorrection;
  double cosVertOffsetCorrection;
};

/* HDLParser.cxx:133-145 */
static double HDL32AdjustTimeStamp(int firingblock, int dsr) {
  return (firingblock * 46.08) + (dsr * 1.152);
}
static double VLP16AdjustTimeStamp(int firingblock, int dsr, int firingwithinblock) {
  return (firingblock * 110.592) + (dsr * 2.304) + (firingwithinblock * 55.296);
}

struct PointXYZI {
  float x, y, z, intensity;
};
struct PointMeta { /* type_defs.h:168-176 (flag bytes are never written by the parser) */
  uint16_t azimuth;
  float distance;
};

/* HDLFrame.h:13-47, ctor HDLFrame.cxx:9-16 */
struct Frame {
  usec_t timestamp = NOT_A_DATE_TIME;
  std::vector<std::vector<PointXYZI>> points;
  std::vector<std::vector<PointMeta>> pointsMeta;
  std::vector<std::pair<usec_t, std::string>> packets;
  PoseTransform carpose;
  bool isInMemory = false;
  int skips = -1; /* uint8_t in the reference, uninitialised by the ctor */
  bool hdl64Order = false;
};

struct TraceRec {
  int32_t packet;
  uint8_t block, dsr, laser, intensity;
  int32_t frame;
  float x, y, z;
  uint16_t azimuth, raw_distance;
  uint32_t tadj_us;
};

/* ---------------------------------------------------------------------------------
 * HDLParser::vsInternal (HDLParser.cxx:150-281) restated.
 * ------------------------------------------------------------------------------- */
class Parser {
 public:
  Parser() {
    /* HDLParser.cxx:154-190 */
    firingSkip = 0;
    lastAzimuth = -1;
    pointsSkip = 0;
    shouldCropReturns = false;
    shouldCropInside = false;
    for (int i = 0; i < 6; ++i) cropRegion[i] = 0.0;
    correctionsInitialized = false;
    calibFileReportedNumLasers = 64;
    laserSelections.assign(kMaxNumLasers, true);
    isDualReturnData = false;
    isHDL64Data = false;
    frameMetaInited = false;
    static const int lut[64] = {38, 39, 42, 43, 32, 33, 36, 37, 40, 41, 46, 47, 50, 51, 54, 55,
                                44, 45, 48, 49, 52, 53, 58, 59, 62, 63, 34, 35, 56, 57, 60, 61,
                                6,  7,  10, 11, 0,  1,  4,  5,  8,  9,  14, 15, 18, 19, 22, 23,
                                12, 13, 16, 17, 20, 21, 26, 27, 30, 31, 2,  3,  24, 25, 28, 29};
    std::memcpy(HDL64BeamLUT, lut, sizeof(lut));
    std::memset(laser_corrections_, 0, sizeof(laser_corrections_));
    initLookUpTables();
    transMgr = std::make_shared<TransformManager>();
    traceOn = false;
    framesClosed = 0;
    packetCounter = 0;
    unloadData(); /* HDLParser::HDLParser, HDLParser.cxx:284-288 */
  }

  /* HDLParser.cxx:755-768 */
  void initLookUpTables() {
    if (cos_lookup_table_.size() == 0 || sin_lookup_table_.size() == 0) {
      cos_lookup_table_.resize(kNumRotAngles);
      sin_lookup_table_.resize(kNumRotAngles);
      for (int i = 0; i < kNumRotAngles; i++) {
        double rad = toRadians(i / 100.0);
        cos_lookup_table_[i] = std::cos(rad);
        sin_lookup_table_[i] = std::sin(rad);
      }
    }
  }

  /* HDLParser.cxx:771-858 minus the XML parsing: the five raw values per <px> item */
  void setCalibration(const double* rot, const double* vert, const double* dist_cm,
                      const double* voff_cm, const double* hoff_cm, int n_rows, int n_enabled) {
    calibFileReportedNumLasers = n_enabled;
    for (int index = 0; index < n_rows && index < kMaxNumLasers; ++index) {
      LaserCorrection& c = laser_corrections_[index];
      c.azimuthCorrection = rot[index];
      c.verticalCorrection = vert[index];
      c.distanceCorrection = dist_cm[index] / 100.0;
      c.verticalOffsetCorrection = voff_cm[index] / 100.0;
      c.horizontalOffsetCorrection = hoff_cm[index] / 100.0;
      c.cosVertCorrection = std::cos(toRadians(c.verticalCorrection));
      c.sinVertCorrection = std::sin(toRadians(c.verticalCorrection));
    }
    for (int i = 0; i < 64; i++) {
      LaserCorrection correction = laser_corrections_[i];
      laser_corrections_[i].sinVertOffsetCorrection =
          correction.verticalOffsetCorrection * correction.sinVertCorrection;
      laser_corrections_[i].cosVertOffsetCorrection =
          correction.verticalOffsetCorrection * correction.cosVertCorrection;
    }
    correctionsInitialized = true;
    unloadData(); /* setCorrectionsFile, HDLParser.cxx:471-474 */
  }

  /* HDLParser.cxx:478-486 */
  void unloadData() {
    lastAzimuth = -1;
    isDualReturnData = false;
    isHDL64Data = false;
    frames.clear();
    currentFrame = createHDLFrame();
  }

  /* HDLParser.cxx:553-584 */
  std::shared_ptr<Frame> createHDLFrame() {
    std::shared_ptr<Frame> f(new Frame);
    f->points.resize(calibFileReportedNumLasers);
    for (int i = 0; i < calibFileReportedNumLasers; ++i) f->points[i].reserve(kMaxPtsPerLaser);
    f->pointsMeta.resize(calibFileReportedNumLasers);
    for (int i = 0; i < calibFileReportedNumLasers; ++i) f->pointsMeta[i].reserve(kMaxPtsPerLaser);
    f->isInMemory = true;
    frameMetaInited = false;
    return f;
  }

  /* HDLParser.cxx:867-897 (splitCounter is never set non-zero anywhere) */
  void splitFrame() {
    if (isHDL64Data) {
      std::vector<std::vector<PointXYZI>> re_pts(64);
      std::vector<std::vector<PointMeta>> re_ptm(64);
      for (int i = 0; i < 64; ++i) {
        /* the reference indexes points[HDL64BeamLUT[i]] unchecked; a calibration that
         * reports < 64 lasers together with 0xddff blocks is undefined behaviour there.
         * The oracle leaves such rows empty. */
        const size_t src = (size_t)HDL64BeamLUT[i];
        if (src < currentFrame->points.size()) {
          re_pts[i] = std::move(currentFrame->points[src]);
          re_ptm[i] = std::move(currentFrame->pointsMeta[src]);
        }
      }
      currentFrame->points = std::move(re_pts);
      currentFrame->pointsMeta = std::move(re_ptm);
      currentFrame->hdl64Order = true;
    }
    frames.push_back(currentFrame);
    ++framesClosed;
    currentFrame = createHDLFrame();
  }

  /* HDLParser.cxx:587-752 */
  void pushFiringData(const unsigned char laserId, const unsigned char rawLaserId,
                      unsigned short azimuth, const usec_t timestamp, const unsigned int rawtime,
                      const FiringView& fv, int dsr, const LaserCorrection* correction,
                      const Affine3* geotransform, int firingBlock, unsigned int tadj) {
    (void)rawLaserId;
    (void)timestamp;
    (void)rawtime;
    azimuth %= 36000;
    const short intensity = fv.intensity(dsr);

    double cosAzimuth, sinAzimuth;
    if (correction->azimuthCorrection == 0) {
      cosAzimuth = cos_lookup_table_[azimuth];
      sinAzimuth = sin_lookup_table_[azimuth];
    } else {
      double azimuthInRadians =
          toRadians((static_cast<double>(azimuth) / 100.0) - correction->azimuthCorrection);
      cosAzimuth = std::cos(azimuthInRadians);
      sinAzimuth = std::sin(azimuthInRadians);
    }

    double distanceM = fv.distance(dsr) * 0.002 + correction->distanceCorrection;
    double xyDistance = distanceM * correction->cosVertCorrection;

    double pos[3] = {
        xyDistance * sinAzimuth - correction->horizontalOffsetCorrection * cosAzimuth,
        xyDistance * cosAzimuth + correction->horizontalOffsetCorrection * sinAzimuth,
        distanceM * correction->sinVertCorrection + correction->verticalOffsetCorrection};

    if (shouldCropReturns) {
      /* the reference names this flag "pointOutsideOfBox"; it is true when INSIDE */
      bool inBox = pos[0] >= cropRegion[0] && pos[0] <= cropRegion[1] &&
                   pos[1] >= cropRegion[2] && pos[1] <= cropRegion[3] &&
                   pos[2] >= cropRegion[4] && pos[2] <= cropRegion[5];
      if ((inBox && !shouldCropInside) || (!inBox && shouldCropInside)) return;
    }

    if (geotransform) transformPoint(pos, *geotransform);
    PointXYZI p;
    p.x = pos[0];
    p.y = pos[1];
    p.z = pos[2];
    p.intensity = intensity;
    PointMeta m;
    m.azimuth = azimuth;
    m.distance = distanceM;
    /* points[laserId] is unchecked in the reference; out-of-range ids are UB there and
     * dropped here. */
    if ((size_t)laserId < currentFrame->points.size()) {
      currentFrame->points[laserId].push_back(p);
      currentFrame->pointsMeta[laserId].push_back(m);
      if (traceOn) {
        TraceRec r;
        r.packet = packetCounter;
        r.block = (uint8_t)firingBlock;
        r.dsr = (uint8_t)dsr;
        r.laser = laserId;
        r.intensity = (uint8_t)intensity;
        r.frame = framesClosed;
        r.x = p.x;
        r.y = p.y;
        r.z = p.z;
        r.azimuth = azimuth;
        r.raw_distance = fv.distance(dsr);
        r.tadj_us = tadj;
        trace.push_back(r);
      }
    }
  }

  /* HDLParser.cxx:900-977 */
  void processFiring(const FiringView& firingData, int hdl64offset, int firingBlock,
                     int azimuthDiff, usec_t timestamp, unsigned int rawtime,
                     const Affine3* geotransform) {
    const bool dual = (lastAzimuth == firingData.rotationalPosition()) && (!isHDL64Data);
    if (dual && !isDualReturnData) isDualReturnData = true;

    for (int dsr = 0; dsr < kLaserPerFiring; dsr++) {
      unsigned char rawLaserId = static_cast<unsigned char>(dsr + hdl64offset);
      unsigned char laserId = rawLaserId;
      unsigned short azimuth = firingData.rotationalPosition();

      int firingWithinBlock = 0;
      if (calibFileReportedNumLasers == 16) {
        if (laserId >= 16) {
          laserId -= 16;
          firingWithinBlock = 1;
        }
      }

      double timestampadjustment = 0.0;
      double blockdsr0 = 0.0;
      double nextblockdsr0 = 1.0;
      if (calibFileReportedNumLasers == 32) {
        timestampadjustment = HDL32AdjustTimeStamp(firingBlock, dsr);
        nextblockdsr0 = HDL32AdjustTimeStamp(firingBlock + 1, 0);
        blockdsr0 = HDL32AdjustTimeStamp(firingBlock, 0);
      } else if (calibFileReportedNumLasers == 16) {
        timestampadjustment = VLP16AdjustTimeStamp(firingBlock, laserId, firingWithinBlock);
        nextblockdsr0 = VLP16AdjustTimeStamp(firingBlock + 1, 0, 0);
        blockdsr0 = VLP16AdjustTimeStamp(firingBlock, 0, 0);
      }
      int azimuthadjustment = (int)std::round(
          azimuthDiff * ((timestampadjustment - blockdsr0) / (nextblockdsr0 - blockdsr0)));
      timestampadjustment = std::round(timestampadjustment);

      if (firingData.distance(dsr) != 0.0 && laserSelections[laserId]) {
        pushFiringData(laserId, rawLaserId, (unsigned short)(azimuth + azimuthadjustment),
                       timestamp + (usec_t)timestampadjustment,
                       rawtime + static_cast<unsigned int>(timestampadjustment), firingData, dsr,
                       &(laser_corrections_[dsr + hdl64offset]), geotransform, firingBlock,
                       static_cast<unsigned int>(timestampadjustment));
      }
    }
  }

  /* HDLParser.cxx:980-1062 */
  void processHDLPacket(const uint8_t* data, std::size_t bytesReceived, usec_t timestamp) {
    if (bytesReceived != 1206) return;

    std::shared_ptr<PoseTransform> transform(new PoseTransform);
    const uint32_t rawtime = rd32(data + 1200);

    transMgr->interpolateTransform(timestamp, transform.get());
    if (!frameMetaInited) {
      currentFrame->carpose = *transform; /* memcpy of the whole PoseTransform */
      currentFrame->timestamp = timestamp;
      currentFrame->skips = firingSkip;
      currentFrame->packets.push_back(
          std::make_pair(timestamp, std::string(reinterpret_cast<const char*>(data), bytesReceived)));
      frameMetaInited = true;
    }
    transform->timestamp = timestamp;
    std::shared_ptr<Affine3> geotransform;
    if (transform->seconds_pos != -1) {
      /* reprojectToFrameBeginning, HDLParser.cxx:1057-1062 */
      for (int i = 0; i < 3; ++i) transform->T[i] -= currentFrame->carpose.T[i];
      geotransform = std::shared_ptr<Affine3>(new Affine3(poseMatrix(*transform)));
    }
    currentFrame->packets.push_back(
        std::make_pair(timestamp, std::string(reinterpret_cast<const char*>(data), bytesReceived)));

    int firingBlock = firingSkip;
    firingSkip = 0;

    std::vector<int> diffs(kFiringPerPkt - 1);
    for (int i = 0; i < kFiringPerPkt - 1; ++i) {
      int localDiff = (36000 + rd16(data + 100 * (i + 1) + 2) - rd16(data + 100 * i + 2)) % 36000;
      diffs[i] = localDiff;
    }
    std::nth_element(diffs.begin(), diffs.begin() + kFiringPerPkt / 2, diffs.end());
    int azimuthDiff = diffs[kFiringPerPkt / 2];

    for (; firingBlock < kFiringPerPkt; ++firingBlock) {
      FiringView firingData{data + 100 * firingBlock};
      int hdl64offset = (firingData.blockIdentifier() == BLOCK_0_TO_31) ? 0 : 32;
      isHDL64Data |= (hdl64offset > 0);

      if (firingData.rotationalPosition() < lastAzimuth) {
        firingSkip = firingBlock;
        splitFrame();
      }

      if (pointsSkip == 0 || firingBlock % (pointsSkip + 1) == 0) {
        processFiring(firingData, hdl64offset, firingBlock, azimuthDiff, timestamp, rawtime,
                      geotransform.get());
      }
      lastAzimuth = firingData.rotationalPosition();
    }
    ++packetCounter;
  }

  /* HDLParser.cxx:505-544 over an in-memory array instead of a pcap file */
  bool getFrame(const uint8_t* data, int64_t n, int64_t stride, const usec_t* t_us,
                int64_t startPacket, int skip) {
    unloadData();
    if (!correctionsInitialized) return false;
    firingSkip = skip;
    packetCounter = (int32_t)startPacket;
    for (int64_t p = startPacket; p < n; ++p) {
      processHDLPacket(data + p * stride, 1206, t_us[p]);
      if (frames.size()) return true;
    }
    splitFrame();
    return true;
  }

  /* state */
  std::deque<std::shared_ptr<Frame>> frames;
  std::shared_ptr<Frame> currentFrame;
  bool frameMetaInited;
  std::shared_ptr<TransformManager> transMgr;
  bool isDualReturnData;
  bool isHDL64Data;
  int HDL64BeamLUT[64];
  int lastAzimuth;
  int firingSkip;
  std::vector<double> cos_lookup_table_;
  std::vector<double> sin_lookup_table_;
  LaserCorrection laser_corrections_[kMaxNumLasers];
  int calibFileReportedNumLasers;
  bool correctionsInitialized;
  int pointsSkip;
  bool shouldCropReturns;
  bool shouldCropInside;
  double cropRegion[6];
  std::vector<bool> laserSelections;

  /* oracle-only instrumentation */
  bool traceOn;
  std::vector<TraceRec> trace;
  int32_t framesClosed;
  int32_t packetCounter;
};

}  // namespace vo

struct vo_parser {
  vo::Parser p;
};

extern "C" {

vo_parser* vo_create(void) { return new vo_parser; }
void vo_destroy(vo_parser* h) { delete h; }

void vo_set_calibration(vo_parser* h, const double* rot, const double* vert, const double* dist_cm,
                        const double* voff_cm, const double* hoff_cm, int n_rows, int n_enabled) {
  h->p.setCalibration(rot, vert, dist_cm, voff_cm, hoff_cm, n_rows, n_enabled);
}
void vo_set_laser_selection(vo_parser* h, const int32_t sel[64]) {
  for (int i = 0; i < 64; ++i) h->p.laserSelections[i] = sel[i]; /* HDLParser.cxx:367-373 */
}
void vo_set_points_skip(vo_parser* h, int32_t s) { h->p.pointsSkip = s; }
void vo_set_crop(vo_parser* h, int32_t crop_returns, int32_t crop_inside, const double region[6]) {
  h->p.shouldCropReturns = !!crop_returns;
  h->p.shouldCropInside = !!crop_inside;
  for (int i = 0; i < 6; ++i) h->p.cropRegion[i] = region[i];
}

void vo_clear_poses(vo_parser* h) { h->p.transMgr->clearTransforms(); }
void vo_add_pose(vo_parser* h, int64_t t_us, const double T[3], const double R[3], const double V[3]) {
  std::shared_ptr<vo::PoseTransform> p(new vo::PoseTransform);
  for (int i = 0; i < 3; ++i) {
    p->T[i] = T[i];
    p->R[i] = R[i];
    p->V[i] = V[i];
  }
  p->timestamp = t_us;
  p->seconds_pos = 0;
  h->p.transMgr->addTransform(p);
}
int32_t vo_num_poses(vo_parser* h) { return h->p.transMgr->getNumberOfTransforms(); }
int32_t vo_interpolate(vo_parser* h, int64_t t_us, double out[9], double* seconds_pos) {
  vo::PoseTransform tr;
  bool ok = h->p.transMgr->interpolateTransform(t_us, &tr);
  for (int i = 0; i < 3; ++i) {
    out[i] = tr.T[i];
    out[3 + i] = tr.R[i];
    out[6 + i] = tr.V[i];
  }
  *seconds_pos = tr.seconds_pos;
  return ok ? 1 : 0;
}
void vo_pose_matrix(const double TRV[9], double out[12]) {
  vo::PoseTransform p;
  for (int i = 0; i < 3; ++i) {
    p.T[i] = TRV[i];
    p.R[i] = TRV[3 + i];
    p.V[i] = TRV[6 + i];
  }
  vo::Affine3 a = vo::poseMatrix(p);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) out[4 * r + c] = a.L[r][c];
    out[4 * r + 3] = a.t[r];
  }
}

void vo_unload(vo_parser* h) {
  h->p.unloadData();
  h->p.trace.clear();
  h->p.framesClosed = 0;
  h->p.packetCounter = 0;
}
void vo_set_firing_skip(vo_parser* h, int32_t s) { h->p.firingSkip = s; }
void vo_get_state(vo_parser* h, int32_t out[4]) {
  out[0] = h->p.lastAzimuth;
  out[1] = h->p.firingSkip;
  out[2] = h->p.frameMetaInited ? 1 : 0;
  out[3] = h->p.isHDL64Data ? 1 : 0;
}

void vo_process_packet(vo_parser* h, const uint8_t* data, uint32_t len, int64_t t_us) {
  h->p.processHDLPacket(data, len, t_us);
}
void vo_process_packets(vo_parser* h, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us) {
  for (int64_t i = 0; i < n; ++i) h->p.processHDLPacket(data + i * stride, 1206, t_us[i]);
}
/* The consumer loop of HDLSource.cxx:209-225 around the restated parser: frames leave as they
 * close (the last one of the list is counted, the list is cleared), after every packet. */
void vo_consume_packets(vo_parser* h, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us, int64_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    h->p.processHDLPacket(data + i * stride, 1206, t_us[i]);
    if (!h->p.frames.empty()) {
      out[0] += 1;
      for (const auto& row : h->p.frames.back()->points) out[1] += (int64_t)row.size();
      h->p.frames.clear();
    }
  }
}
void vo_split_frame(vo_parser* h) { h->p.splitFrame(); }

int32_t vo_num_frames(vo_parser* h) { return (int32_t)h->p.frames.size(); }
void vo_clear_frames(vo_parser* h) { h->p.frames.clear(); }

int32_t vo_frame_get_info(vo_parser* h, int32_t f, vo_frame_info* out) {
  if (f < 0 || f >= (int32_t)h->p.frames.size()) return 0;
  const vo::Frame& fr = *h->p.frames[f];
  out->timestamp_us = fr.timestamp;
  out->skips = fr.skips;
  out->n_lasers = (int32_t)fr.points.size();
  int64_t n = 0;
  for (auto& v : fr.points) n += (int64_t)v.size();
  out->n_points = (int32_t)n;
  out->n_packets = (int32_t)fr.packets.size();
  out->is_hdl64_order = fr.hdl64Order ? 1 : 0;
  out->pad = 0;
  for (int i = 0; i < 3; ++i) {
    out->carpose_TRV[i] = fr.carpose.T[i];
    out->carpose_TRV[3 + i] = fr.carpose.R[i];
    out->carpose_TRV[6 + i] = fr.carpose.V[i];
  }
  out->carpose_seconds_pos = fr.carpose.seconds_pos;
  return 1;
}
int32_t vo_frame_laser_counts(vo_parser* h, int32_t f, int32_t* counts) {
  if (f < 0 || f >= (int32_t)h->p.frames.size()) return 0;
  const vo::Frame& fr = *h->p.frames[f];
  for (size_t i = 0; i < fr.points.size(); ++i) counts[i] = (int32_t)fr.points[i].size();
  return (int32_t)fr.points.size();
}
int32_t vo_frame_points(vo_parser* h, int32_t f, float* xyzi, uint16_t* azimuth, float* distance) {
  if (f < 0 || f >= (int32_t)h->p.frames.size()) return 0;
  const vo::Frame& fr = *h->p.frames[f];
  int64_t k = 0;
  for (size_t l = 0; l < fr.points.size(); ++l) {
    for (size_t i = 0; i < fr.points[l].size(); ++i, ++k) {
      xyzi[4 * k + 0] = fr.points[l][i].x;
      xyzi[4 * k + 1] = fr.points[l][i].y;
      xyzi[4 * k + 2] = fr.points[l][i].z;
      xyzi[4 * k + 3] = fr.points[l][i].intensity;
      azimuth[k] = fr.pointsMeta[l][i].azimuth;
      distance[k] = fr.pointsMeta[l][i].distance;
    }
  }
  return (int32_t)k;
}
int64_t vo_open_frame_points(vo_parser* h) {
  int64_t n = 0;
  for (auto& v : h->p.currentFrame->points) n += (int64_t)v.size();
  return n;
}

void vo_trace_enable(vo_parser* h, int32_t on) {
  h->p.traceOn = !!on;
  if (!on) h->p.trace.clear();
}
int64_t vo_trace_size(vo_parser* h) { return (int64_t)h->p.trace.size(); }
void vo_trace_fetch(vo_parser* h, int32_t* packet, uint8_t* block, uint8_t* dsr, uint8_t* laser,
                    int32_t* frame, float* x, float* y, float* z, uint8_t* intensity,
                    uint16_t* azimuth, uint16_t* raw_distance, uint32_t* tadj_us) {
  const auto& t = h->p.trace;
  for (size_t i = 0; i < t.size(); ++i) {
    packet[i] = t[i].packet;
    block[i] = t[i].block;
    dsr[i] = t[i].dsr;
    laser[i] = t[i].laser;
    frame[i] = t[i].frame;
    x[i] = t[i].x;
    y[i] = t[i].y;
    z[i] = t[i].z;
    intensity[i] = t[i].intensity;
    azimuth[i] = t[i].azimuth;
    raw_distance[i] = t[i].raw_distance;
    tadj_us[i] = t[i].tadj_us;
  }
}

/* HDLParser.cxx:1065-1160 over an in-memory array (file positions -> packet indices; the
 * file-rename side effect and filenameTime are I/O and not restated). */
int32_t vo_read_frame_information(const uint8_t* data, int64_t n, int64_t stride,
                                  const int64_t* t_us, int32_t* start_packet, int32_t* skips,
                                  int64_t* timestamp_us, int32_t cap) {
  int32_t count = 0;
  auto push = [&](int64_t pkt, int sk, int64_t ts) {
    if (count < cap) {
      start_packet[count] = (int32_t)pkt;
      skips[count] = sk;
      timestamp_us[count] = ts;
    }
    ++count;
  };
  unsigned int lastAzimuth = 0;
  bool first = true;
  push(0, 0, vo::NOT_A_DATE_TIME);
  for (int64_t p = 0; p < n; ++p) {
    const uint8_t* pkt = data + p * stride;
    if (first) {
      first = false;
      if (cap > 0) timestamp_us[0] = t_us[p];
    }
    for (int i = 0; i < vo::kFiringPerPkt; ++i) {
      unsigned int rot = vo::rd16(pkt + 100 * i + 2);
      if (rot < lastAzimuth) push(p, i, t_us[p]);
      lastAzimuth = rot;
    }
  }
  return count;
}

int32_t vo_get_frame(vo_parser* h, const uint8_t* data, int64_t n, int64_t stride,
                     const int64_t* t_us, int64_t start_packet, int32_t skip) {
  h->p.trace.clear();
  h->p.framesClosed = 0;
  return h->p.getFrame(data, n, stride, t_us, start_packet, skip) ? 1 : 0;
}

/* ---- N3: geodesy, TimeSolver, INS record -> pose ------------------------------------------- */
namespace vo {
const double kWgsA = 6378137.0000;  /* semi-major axis, CoordiTran.cpp:58 */
const double kWgsB = 6356752.3142;  /* semi-minor axis, CoordiTran.cpp:59 */
}

void vo_llh2xyz(const double llh[3], double xyz[3]) {
  /* CoordiTran.cpp:51-80, operation order kept */
  const double phi = llh[0], lambda = llh[1], h = llh[2];
  const double a = vo::kWgsA, b = vo::kWgsB;
  const double e = std::sqrt(1 - (b / a) * (b / a));
  const double sinphi = std::sin(phi), cosphi = std::cos(phi);
  const double coslam = std::cos(lambda), sinlam = std::sin(lambda);
  const double tan2phi = (std::tan(phi)) * (std::tan(phi));
  const double tmp = 1 - e * e;
  const double tmpden = std::sqrt(1 + tmp * tan2phi);
  xyz[0] = (a * coslam) / tmpden + h * coslam * cosphi;
  xyz[1] = (a * sinlam) / tmpden + h * sinlam * cosphi;
  const double tmp2 = std::sqrt(1 - e * e * sinphi * sinphi);
  xyz[2] = (a * tmp * sinphi) / tmp2 + h * sinphi;
}

void vo_xyz2llh(const double xyz[3], double llh[3]) {
  /* CoordiTran.cpp:82-150: closed-form (Heikkinen-style) inverse */
  const double pi = 3.141592653589793;
  const double x = xyz[0], y = xyz[1], z = xyz[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double a = vo::kWgsA, b = vo::kWgsB;
  const double e = std::sqrt(1 - (b / a) * (b / a));
  const double b2 = b * b;
  const double e2 = e * e;
  const double ep = e * (a / b);
  const double r = std::sqrt(x2 + y2);
  const double r2 = r * r;
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
  const double U = std::sqrt(tmp + z2);
  const double V = std::sqrt(tmp + (1 - e2) * z2);
  const double zo = (b2 * z) / (a * V);
  const double height = U * (a * V - b2) / (a * V);
  const double lat = std::atan((z + ep * ep * zo) / r);
  const double temp = std::atan(y / x);
  double lon;
  if (x >= 0)
    lon = temp;
  else if ((x < 0) & (y >= 0))
    lon = pi + temp;
  else
    lon = temp - pi;
  llh[0] = lat;
  llh[1] = lon;
  llh[2] = height;
}

void vo_llh2enu(const double llh[3], const double orgxyz[3], double enu[3]) {
  /* llh2enu (CoordiTran.cpp:271-276) = llh2xyz then xyz2enu (:152-187) */
  double xyz[3] = {0, 0, 0};
  vo_llh2xyz(llh, xyz);
  double dif[3], orgllh[3];
  for (int i = 0; i < 3; ++i) dif[i] = xyz[i] - orgxyz[i];
  vo_xyz2llh(orgxyz, orgllh);
  const double phi = orgllh[0], lam = orgllh[1];
  const double sinphi = std::sin(phi), cosphi = std::cos(phi);
  const double sinlam = std::sin(lam), coslam = std::cos(lam);
  const double R[3][3] = {{-sinlam, coslam, 0},
                          {-sinphi * coslam, -sinphi * sinlam, cosphi},
                          {cosphi * coslam, cosphi * sinlam, sinphi}};
  enu[0] = enu[1] = enu[2] = 0;
  for (int i = 0; i < 3; ++i) {
    enu[0] = enu[0] + R[0][i] * dif[i];
    enu[1] = enu[1] + R[1][i] * dif[i];
    enu[2] = enu[2] + R[2][i] * dif[i];
  }
}

void vo_ts_init(vo_time_solver* s) {
  s->base_us = 0;
  s->last_report = 0;  /* TimeSolver.cxx:8 */
  s->inited = 0;
}

int64_t vo_ts_hdl(vo_time_solver* s, uint32_t gps, int64_t now_us) {
  /* TimeSolver.cxx:34-49 */
  if (!s->inited) {
    /* hdlHourTime = now truncated to the hour, hdlOffset = now - hdlHourTime - gps:
     * their sum is all that is ever used */
    s->base_us = now_us - (int64_t)gps;
    s->inited = 1;
  }
  if (s->last_report > gps) s->base_us += 3600ll * 1000000ll;  /* one hour wrapped */
  s->last_report = gps;
  return s->base_us + (int64_t)gps;
}

int64_t vo_ts_ins(const vo_ins_pva* d, int64_t now_us) {
  /* TimeSolver.cxx:20-33; time_duration(h, 0, 0, frac) truncates the double to ticks (us) */
  const int64_t hour_us = 3600ll * 1000000ll;
  const int64_t p = (int64_t)(d->week_number * 168) * hour_us + (int64_t)(double(d->milliseconds) * 1e3);
  const int64_t ins = (int64_t)(d->week_number_pos * 168) * hour_us + (int64_t)(double(d->seconds_pos) * 1e6);
  return now_us + (ins - p);
}

void vo_ins_pose(const vo_ins_pva* d, const double orgxyz[3], double trv[9]) {
  /* INSSource.cxx:305-317: TO_RADIUS(x) = x * M_PI / 180 (type_defs.h) */
  const double in[3] = {d->LLH[0] * M_PI / 180, d->LLH[1] * M_PI / 180, d->LLH[2]};
  double out[3] = {0, 0, 0};
  vo_llh2enu(in, orgxyz, out);
  for (int k = 0; k < 3; ++k) {
    trv[k] = out[k];
    trv[3 + k] = d->Eulr[k];
    trv[6 + k] = d->V[k];
  }
}

} /* extern "C" */
