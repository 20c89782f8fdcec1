// facade_driver.cpp -- drives the C++ facade the way the reference's consumers do and dumps the
// frames it produces, so that pytest can compare them with the oracle.
//
//   facade_driver stream  <calib.xml> <packets.bin> <times.bin> <poses.bin|-> <out.bin> [batch]
//       per packet: processHDLPacket(); getAllFrames(); clearAllFrames()   (HDLSource.cxx:209-225)
//   facade_driver offline <calib.xml> <file.pcap> <poses.bin|-> <out.bin>
//       readFrameInformation() then getFrame() for every index entry     (HDLManager.cxx:103-117,195-211)
//
// packets.bin: n x 1206 bytes; times.bin: n x int64; poses.bin: n x (int64 t_us + 9 doubles).
// out.bin: per frame { int64 timestamp_us, int32 skips, int32 n_lasers, int32 n_packets,
// int32 carpose_valid, double carpose[9], int32 counts[n_lasers], then per point
// float x,y,z,intensity, uint16 azimuth, float distance } preceded by int32 n_frames.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "VeloSLAM.h"

static std::vector<char> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static void dumpFrame(std::ofstream& os, const HDLFrame& f) {
  const int64_t ts = f.timestamp.us;
  const int32_t skips = f.skips, nl = (int32_t)f.points.size(), np = (int32_t)f.packets.size();
  const int32_t valid = f.carpose->seconds_pos != -1 ? 1 : 0;
  os.write((const char*)&ts, 8);
  os.write((const char*)&skips, 4);
  os.write((const char*)&nl, 4);
  os.write((const char*)&np, 4);
  os.write((const char*)&valid, 4);
  double cp[9];
  for (int k = 0; k < 3; ++k) {
    cp[k] = f.carpose->T[k];
    cp[3 + k] = f.carpose->R[k];
    cp[6 + k] = f.carpose->V[k];
  }
  os.write((const char*)cp, sizeof(cp));
  for (int l = 0; l < nl; ++l) {
    const int32_t c = f.points[l] ? (int32_t)f.points[l]->points.size() : 0;
    os.write((const char*)&c, 4);
  }
  for (int l = 0; l < nl; ++l) {
    if (!f.points[l]) continue;
    for (size_t i = 0; i < f.points[l]->points.size(); ++i) {
      const pcl::PointXYZI& p = f.points[l]->points[i];
      const PointMeta& m = (*f.pointsMeta[l])[i];
      os.write((const char*)&p.x, 4);
      os.write((const char*)&p.y, 4);
      os.write((const char*)&p.z, 4);
      os.write((const char*)&p.intensity, 4);
      os.write((const char*)&m.azimuth, 2);
      os.write((const char*)&m.distance, 4);
    }
  }
}

static std::shared_ptr<TransformManager> loadPoses(const std::string& path) {
  std::shared_ptr<TransformManager> tm(new TransformManager);
  if (path == "-") return tm;
  std::vector<char> raw = slurp(path);
  const size_t rec = 8 + 9 * 8;
  for (size_t off = 0; off + rec <= raw.size(); off += rec) {
    std::shared_ptr<PoseTransform> p(new PoseTransform);
    int64_t t;
    double v[9];
    std::memcpy(&t, raw.data() + off, 8);
    std::memcpy(v, raw.data() + off + 8, sizeof(v));
    for (int k = 0; k < 3; ++k) {
      p->T[k] = v[k];
      p->R[k] = v[3 + k];
      p->V[k] = v[6 + k];
    }
    p->timestamp = ptime(t);
    p->seconds_pos = 0;
    tm->addTransform(p);
  }
  return tm;
}

// host-only modes (no GPU needed):
//   interp   <poses.bin> <queries.bin(int64)> <out.bin>   9 doubles + int32 found + int32 valid
//   index    <file.pcap> <out.bin>                        int32 n, then n x (int64 pos, int32 skips, int64 ts)
//   writepcap <packets.bin> <times.bin> <out.pcap>
//   calib    <db.xml> <out.bin>                           int32 n_enabled, int32 n_rows, 64 x 5 doubles
static int hostModes(int argc, char** argv) {
  const std::string mode = argv[1];
  if (mode == "interp" && argc >= 5) {
    std::shared_ptr<TransformManager> tm = loadPoses(argv[2]);
    std::vector<char> q = slurp(argv[3]);
    std::ofstream os(argv[4], std::ios::binary);
    for (size_t off = 0; off + 8 <= q.size(); off += 8) {
      int64_t t;
      std::memcpy(&t, q.data() + off, 8);
      PoseTransform out;
      ptime pt(t);
      const int32_t found = tm->interpolateTransform(pt, &out) ? 1 : 0;
      const int32_t valid = out.seconds_pos != -1 ? 1 : 0;
      double v[9];
      for (int k = 0; k < 3; ++k) {
        v[k] = out.T[k];
        v[3 + k] = out.R[k];
        v[6 + k] = out.V[k];
      }
      os.write((const char*)v, sizeof(v));
      os.write((const char*)&found, 4);
      os.write((const char*)&valid, 4);
    }
    return 0;
  }
  if (mode == "index" && argc >= 4) {
    HDLParser parser;
    std::vector<std::shared_ptr<HDLFrame> > index = parser.readFrameInformation(argv[2]);
    std::ofstream os(argv[3], std::ios::binary);
    const int32_t n = (int32_t)index.size();
    os.write((const char*)&n, 4);
    for (auto& f : index) {
      const int64_t pos = f->fileStartPos, ts = f->timestamp.us;
      const int32_t sk = f->skips;
      os.write((const char*)&pos, 8);
      os.write((const char*)&sk, 4);
      os.write((const char*)&ts, 8);
    }
    return 0;
  }
  if (mode == "writepcap" && argc >= 5) {
    std::vector<char> pk = slurp(argv[2]);
    std::vector<char> tm = slurp(argv[3]);
    vtkPacketFileWriter w;
    if (!w.open(argv[4])) return 1;
    for (size_t i = 0; i < pk.size() / 1206; ++i) {
      int64_t t;
      std::memcpy(&t, tm.data() + 8 * i, 8);
      if (!w.writePacket((const unsigned char*)pk.data() + 1206 * i, 1206, ptime(t))) return 1;
    }
    w.close();
    return 0;
  }
  if (mode == "calib" && argc >= 4) {
    HDLParser parser;
    parser.setCorrectionsFile(argv[2]);
    std::ofstream os(argv[3], std::ios::binary);
    const int32_t ne = parser.getNumberOfChannels();
    double vert[64];
    parser.getVerticalCorrections(vert);
    os.write((const char*)&ne, 4);
    os.write((const char*)vert, sizeof(vert));
    return 0;
  }
  return -1;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::cerr << "usage: see the header comment" << std::endl;
    return 2;
  }
  const int hm = hostModes(argc, argv);
  if (hm >= 0) return hm;
  if (argc < 6) return 2;
  const std::string mode = argv[1];
  HDLParser parser;
  if (mode == "stream") {
    if (argc < 7) return 2;
    if (argc > 7) parser.setBatchPackets(std::atoi(argv[7]));
    parser.setCorrectionsFile(argv[2]);
    parser.setTransformMgr(loadPoses(argv[5]));
    std::vector<char> pk = slurp(argv[3]);
    std::vector<char> tm = slurp(argv[4]);
    const size_t n = pk.size() / 1206;
    std::vector<std::shared_ptr<HDLFrame> > all;
    for (size_t i = 0; i < n; ++i) {
      int64_t t;
      std::memcpy(&t, tm.data() + 8 * i, 8);
      parser.processHDLPacket((unsigned char*)pk.data() + 1206 * i, 1206, ptime(t));
      std::deque<std::shared_ptr<HDLFrame> > fr = parser.getAllFrames();
      if (fr.size()) {
        for (auto& f : fr) all.push_back(f);
        parser.clearAllFrames();
      }
    }
    if (!parser.lastError().empty()) {
      std::cerr << "facade error: " << parser.lastError() << std::endl;
      return 1;
    }
    std::ofstream os(argv[6], std::ios::binary);
    const int32_t nf = (int32_t)all.size();
    os.write((const char*)&nf, 4);
    for (auto& f : all) dumpFrame(os, *f);
    return 0;
  }
  if (mode == "offline") {
    parser.setCorrectionsFile(argv[2]);
    std::shared_ptr<TransformManager> tmgr = loadPoses(argv[4]);
    parser.setTransformMgr(tmgr);
    std::vector<std::shared_ptr<HDLFrame> > index = parser.readFrameInformation(argv[3]);
    std::ofstream os(argv[5], std::ios::binary);
    const int32_t nf = (int32_t)index.size();
    os.write((const char*)&nf, 4);
    for (auto& f : index) {
      tmgr->interpolateTransform(f->timestamp, f->carpose.get());   // HDLManager::loadOffline
      if (!parser.getFrame(f, argv[3], f->fileStartPos, f->skips)) return 1;
      dumpFrame(os, *f);
    }
    return 0;
  }
  return 2;
}
