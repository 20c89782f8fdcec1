// facade_driver.cpp -- drives the C++ facade the way the reference's consumers do and dumps the
// frames it produces, so that pytest can compare them with the oracle.
//
//   facade_driver stream  <calib.xml> <packets.bin> <times.bin> <poses.bin|-> <out.bin> [batch] [pipelined 0|1]
//       per packet: processHDLPacket(); getAllFrames(); clearAllFrames()   (HDLSource.cxx:209-225)
//   facade_driver bench   <calib.xml> <packets.bin> <times.bin> <poses.bin|-> <batch> <passes>
//                         <store_packets 0|1> <fetch_meta 0|1> <pipelined 0|1> [cuda device]
//       the same consumer loop over the packet array `passes` times (+1 untimed warm-up pass),
//       prints one JSON line with points, frames and seconds
//   facade_driver offline <calib.xml> <file.pcap> <poses.bin|-> <out.bin>
//       readFrameInformation() then getFrame() for every index entry     (HDLManager.cxx:103-117,195-211)
//
// packets.bin: n x 1206 bytes; times.bin: n x int64; poses.bin: n x (int64 t_us + 9 doubles).
// out.bin: per frame { int64 timestamp_us, int32 skips, int32 n_lasers, int32 n_packets,
// int32 carpose_valid, double carpose[9], int32 counts[n_lasers], then per point
// float x,y,z,intensity, uint16 azimuth, float distance } preceded by int32 n_frames.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <algorithm>
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

#include "VeloSLAM.h"
#include "CalibrationFile.h"

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
//   arena    <out.txt>                                    zero-copy adoption checks (FrameArena.h)
static int hostModes(int argc, char** argv) {
  const std::string mode = argv[1];
  if (mode == "arena" && argc >= 3) {
    std::ofstream rep(argv[2]);
    std::weak_ptr<vs::Arena> watch;
    {
      std::shared_ptr<vs::Arena> a = vs::Arena::heap(1 << 16);
      watch = a;
      pcl::PointXYZI* rec = reinterpret_cast<pcl::PointXYZI*>(a->data());
      for (int i = 0; i < 100; ++i) {
        rec[i].x = (float)i;
        rec[i].y = 2.f * i;
        rec[i].z = -1.f * i;
        rec[i].intensity = (float)(i % 7);
      }
      pcl::PointCloud<pcl::PointXYZI> cloud;
      vs::adopt(cloud.points, a, rec + 10, 50);
      a.reset();  // the vector keeps the arena alive
      rep << "alive_while_adopted " << (watch.expired() ? 0 : 1) << "\n";
      rep << "zero_copy " << ((void*)cloud.points.data() == (void*)(rec + 10) ? 1 : 0) << "\n";
      rep << "size " << cloud.points.size() << "\n";
      rep << "first " << cloud.points.front().x << " last " << cloud.points.back().x << "\n";
      // an ordinary vector in every other respect
      std::vector<pcl::PointXYZI, vs::ArenaAllocator<pcl::PointXYZI> > copy = cloud.points;
      rep << "copy_is_heap " << ((void*)copy.data() != (void*)cloud.points.data() ? 1 : 0) << "\n";
      pcl::PointXYZI extra = {1000.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < 200; ++i) cloud.points.push_back(extra);
      rep << "grown_size " << cloud.points.size() << " kept " << cloud.points[49].x << " new "
          << cloud.points[249].x << "\n";
      cloud.points.resize(300);
      rep << "resized_zero " << cloud.points[299].x + cloud.points[299].intensity << "\n";
      // moving the vector moves the adoption
      PointMetaVector m1, m2;
      std::shared_ptr<vs::Arena> b = vs::Arena::heap(4096);
      PointMeta* pm = reinterpret_cast<PointMeta*>(b->data());
      for (int i = 0; i < 20; ++i) {
        pm[i].azimuth = (unsigned short)(100 * i);
        pm[i].distance = 0.5f * i;
      }
      vs::adopt(m1, b, pm, 20);
      m2 = std::move(m1);
      rep << "moved " << ((void*)m2.data() == (void*)pm ? 1 : 0) << " " << m2[19].azimuth << "\n";
      vs::adopt(m2, b, pm, 0);
      rep << "empty " << m2.size() << "\n";
    }
    rep << "released " << (watch.expired() ? 1 : 0) << "\n";
    return 0;
  }
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
  if (mode == "frontend" && argc >= 5) {
    // TimeSolver / CoordiTran of the facade on the host:
    //   frontend <gps.bin (u32)> <ins.bin (104-byte records)> <out.bin>
    // out: int64 per packet, then per INS record int64 t + 3 doubles ENU
    std::vector<char> g = slurp(argv[2]);
    std::vector<char> ins = slurp(argv[3]);
    std::ofstream os(argv[4], std::ios::binary);
    int64_t clock = 1467331234567890ll;
    TimeSolver ts;
    ts.setClock([&clock]() { return clock; });
    for (size_t i = 0; i + 4 <= g.size(); i += 4) {
      uint32_t v;
      std::memcpy(&v, g.data() + i, 4);
      const int64_t t = ts.calcTimestamp(v).us;
      clock += 553;
      os.write((const char*)&t, 8);
    }
    double org[3] = {-2781621.9891904, 4672106.75052387, 18.8910392};
    static_assert(sizeof(InsPVA) == 104, "INSPVA layout");
    for (size_t i = 0; i + sizeof(InsPVA) <= ins.size(); i += sizeof(InsPVA)) {
      InsPVA rec;
      std::memcpy(&rec, ins.data() + i, sizeof(rec));
      clock = 1467331200000000ll + 10000ll * (int64_t)(i / sizeof(InsPVA));
      const int64_t t = ts.calcTimestamp(&rec).us;
      double llh[3] = {TO_RADIUS(rec.LLH[0]), TO_RADIUS(rec.LLH[1]), rec.LLH[2]}, enu[3];
      llh2enu(llh, org, enu);
      os.write((const char*)&t, 8);
      os.write((const char*)enu, sizeof(enu));
    }
    return 0;
  }
  if (mode == "ins_host" && argc >= 6) {
    // INSSource over loopback: ins_host <port> <n_expected> <out.bin> <out.insmeta>
    // out: per pose int64 t + 9 doubles, in timeline order; then the count re-read from the meta file
    const int port = std::atoi(argv[2]);
    const uint64_t want = (uint64_t)std::atoll(argv[3]);
    INSSource src(port);
    std::shared_ptr<TransformManager> tm(new TransformManager);
    std::shared_ptr<TimeSolver> ts(new TimeSolver);
    int64_t clock = 1467331200000000ll;
    ts->setClock([&clock]() { return clock; });   // frozen clock: t = clock + (pose - sent)
    src.setTimeSolver(ts);
    src.setTransformManager(tm);
    src.setOutputFile(argv[5]);
    src.start();
    if (!src.isRunning()) return 1;
    std::cout << "ready" << std::endl;
    for (int i = 0; i < 1000 && src.posesReceived() < want; ++i) usleep(10000);
    src.stop();
    std::vector<int64_t> t;
    std::vector<double> trv;
    tm->snapshot(&t, &trv);
    std::ofstream os(argv[4], std::ios::binary);
    for (size_t i = 0; i < t.size(); ++i) {
      os.write((const char*)&t[i], 8);
      os.write((const char*)&trv[9 * i], 72);
    }
    TransformManager again;
    again.loadFromMetaFile(argv[5]);
    std::cerr << "poses " << src.posesReceived() << " timeline " << t.size() << " meta "
              << again.getNumberOfTransforms() << std::endl;
    return 0;
  }
  if (mode == "udp_host" && argc >= 5) {
    // HDLSource receive path without a GPU: udp_host <port> <n_expected> <out.bin>
    // out: per packet int64 time + the first 8 payload bytes + uint32 length
    const int port = std::atoi(argv[2]);
    const size_t want = (size_t)std::atoll(argv[3]);
    std::vector<char> rec;
    HDLSource src(port);
    std::shared_ptr<TimeSolver> ts(new TimeSolver);
    ts->setClock([]() { return (int64_t)1467331234567890ll; });
    src.setTimeSolver(ts);
    size_t got = 0;
    src.setPacketCallback([&](const unsigned char* d, unsigned int len, ptime t) {
      const int64_t us = t.us;
      rec.insert(rec.end(), (const char*)&us, (const char*)&us + 8);
      rec.insert(rec.end(), (const char*)d, (const char*)d + 8);
      rec.insert(rec.end(), (const char*)&len, (const char*)&len + 4);
      ++got;
    });
    src.start();
    if (!src.isRunning()) return 1;
    std::cout << "ready" << std::endl;
    for (int i = 0; i < 1000; ++i) {  // up to 10 s
      uint64_t r, d, c;
      src.getCounters(&r, &d, &c);
      if (c >= want) break;
      usleep(10000);
    }
    src.stop();
    uint64_t r, d, c;
    src.getCounters(&r, &d, &c);
    std::cerr << "received " << r << " dropped " << d << " consumed " << c << std::endl;
    std::ofstream os(argv[4], std::ios::binary);
    os.write(rec.data(), (std::streamsize)rec.size());
    return 0;
  }
  if (mode == "manager_host" && argc >= 4) {
    // HDLManager host logic without a GPU: time queries, cache, hard-drive buffers written as
    // pcap files, .hdlmeta round trip.  argv[2]: scratch directory, argv[3]: report file.
    std::ofstream rep(argv[3]);
    HDLManager mgr(100);
    mgr.setBufferDir(argv[2], false);
    const int64_t t0 = 1467331200000000ll;
    std::vector<std::shared_ptr<HDLFrame> > made;
    for (int i = 0; i < 10; ++i) {
      std::shared_ptr<HDLFrame> f(new HDLFrame);
      f->timestamp = ptime(t0 + 100000ll * i);
      f->isInMemory = true;
      f->points.resize(1);
      f->points[0] = pcl::PointCloud<pcl::PointXYZI>::Ptr(new pcl::PointCloud<pcl::PointXYZI>);
      f->points[0]->points.resize(5);
      f->carpose->T[0] = i;
      f->skips = (uint8_t)(i % 12);
      for (int k = 0; k < 3; ++k) {  // three raw packets per frame
        std::string raw(1206, (char)(i * 3 + k));
        f->packets.push_back(std::make_pair(ptime(t0 + 100000ll * i + 288 * k), raw));
      }
      made.push_back(f);
    }
    // out-of-order insertion keeps the timeline sorted
    mgr.addFrame(made[1]);
    mgr.addFrame(made[0]);
    for (int i = 2; i < 10; ++i) mgr.addFrame(made[i]);
    rep << "frames " << mgr.getNumberOfFrames() << "\n";
    ptime q(t0 + 300000);
    HDLFramePtr at = mgr.getFrameAt(q);
    rep << "at3 " << (at ? (int)at->carpose->T[0] : -1) << "\n";
    ptime qn(t0 + 449999);
    HDLFramePtr near = mgr.getFrameNear(qn);
    rep << "near4 " << (near ? (int)near->carpose->T[0] : -1) << "\n";
    ptime qm(t0 + 123);
    rep << "at_missing " << (mgr.getFrameAt(qm) ? 1 : 0) << "\n";
    ptime a(t0 + 200000), b(t0 + 500000);
    rep << "range " << mgr.getRangeBetween(a, b).size() << "\n";
    rep << "recent " << (int)mgr.getRecentFrame()->carpose->T[0] << "\n";
    // a cache of 4 frames: older frames are cleared unless an end user still holds them
    {
      HDLManager small(4);
      std::vector<std::shared_ptr<HDLFrame> > fr;
      for (int i = 0; i < 10; ++i) {
        std::shared_ptr<HDLFrame> f(new HDLFrame);
        f->timestamp = ptime(t0 + 100000ll * i);
        f->isInMemory = true;
        f->points.resize(1);
        fr.push_back(f);
      }
      for (int i = 0; i < 4; ++i) small.addFrame(fr[i]);
      ptime t1(t0 + 100000);
      HDLFramePtr held = small.getFrameAt(t1);
      for (int i = 4; i < 10; ++i) small.addFrame(fr[i]);
      int cleared = 0;
      for (auto& f : fr) cleared += f->points.empty() ? 1 : 0;
      rep << "cleared " << cleared << "\n";
      rep << "held1_alive " << ((held && !fr[1]->points.empty()) ? 1 : 0) << "\n";
      ptime t2(t0 + 200000);
      rep << "cleared_not_on_disk " << (small.getFrameAt(t2) ? 1 : 0) << "\n";
    }
    // hard-drive buffers: 3 frames per pcap file
    HDLManager disk(100);
    disk.setBufferDir(argv[2], false);
    disk.setBufferSize(3);
    disk.startSwaping();
    std::vector<std::shared_ptr<HDLFrame> > made2;
    for (int i = 0; i < 7; ++i) {
      std::shared_ptr<HDLFrame> f(new HDLFrame);
      f->timestamp = ptime(t0 + 100000ll * i);
      f->isInMemory = true;
      for (int k = 0; k < 3; ++k) {
        std::string raw(1206, (char)(i * 3 + k));
        f->packets.push_back(std::make_pair(ptime(t0 + 100000ll * i + 288 * k), raw));
      }
      made2.push_back(f);
      disk.addFrame(f);
    }
    disk.flushFileBuffer();
    for (int i = 0; i < 7; ++i)
      rep << "disk " << i << " " << (made2[i]->isOnHardDrive ? 1 : 0) << " " << made2[i]->fileStartPos
          << " " << to_iso_string(made2[i]->filenameTime) << "\n";
    rep << "savemeta " << (disk.saveHDLMeta() ? 1 : 0) << "\n";
    HDLManager again(100);
    again.setBufferDir(argv[2], false);
    rep << "loadmeta " << (again.loadHDLMeta() ? 1 : 0) << " " << again.getNumberOfFrames() << "\n";
    std::vector<std::shared_ptr<HDLFrame> > meta = again.getAllFrameMeta();
    for (size_t i = 0; i < meta.size(); ++i)
      rep << "meta " << i << " " << meta[i]->timestamp.us << " " << meta[i]->fileStartPos << " "
          << (int)meta[i]->skips << " " << (meta[i]->isOnHardDrive ? 1 : 0) << "\n";
    return 0;
  }
  return -1;
}

int main(int argc, char** argv) {
  if (argc < 3) {
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
    if (argc > 8) parser.setPipelined(std::atoi(argv[8]) != 0);
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
    if (argc > 8 && std::atoi(argv[8]) != 0) {
      // pipelined: whole batches still sit in the rings / on the GPU; everything that closed
      // comes out now (the open frame stays open, as with the per-rotation flush)
      parser.flush();
      for (auto& f : parser.getAllFrames()) all.push_back(f);
      parser.clearAllFrames();
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
  if (mode == "bench") {
    if (argc < 11) return 2;
    const int batch = std::atoi(argv[6]), passes = std::atoi(argv[7]);
    parser.setBatchPackets(batch);
    parser.setStorePackets(std::atoi(argv[8]) != 0);
    parser.setFetchMeta(std::atoi(argv[9]) != 0);
    parser.setPipelined(std::atoi(argv[10]) != 0);
    if (argc > 11) parser.setDevice(std::atoi(argv[11]));
    parser.setCorrectionsFile(argv[2]);
    parser.setTransformMgr(loadPoses(argv[5]));
    std::vector<char> pk = slurp(argv[3]);
    std::vector<char> tm = slurp(argv[4]);
    const size_t n = pk.size() / 1206;
    if (n < 2) return 2;
    std::vector<int64_t> t(n);
    std::memcpy(t.data(), tm.data(), 8 * n);
    const int64_t span = t[n - 1] - t[0] + (t[n - 1] - t[0]) / (int64_t)(n - 1);
    uint64_t points = 0, frames = 0;
    double touch = 0.0;
    auto consume = [&](bool count) {
      // what HDLSource's consumer does with a finished rotation (HDLSource.cxx:220-224): take the
      // frames, hand them on, clear the parser's list
      std::deque<std::shared_ptr<HDLFrame> > fr = parser.getAllFrames();
      if (fr.empty()) return;
      for (auto& f : fr) {
        if (count) {
          points += f->numberOfPoints();
          ++frames;
        }
        for (auto& c : f->points)
          if (c && !c->points.empty()) touch += c->points.front().x + c->points.back().z;
      }
      parser.clearAllFrames();
    };
    auto onePass = [&](int pass, bool count) {
      const int64_t shift = span * pass;
      for (size_t i = 0; i < n; ++i) {
        parser.processHDLPacket((unsigned char*)pk.data() + 1206 * i, 1206, ptime(t[i] + shift));
        if (parser.getAllFrames().size()) consume(count);
      }
    };
    onePass(0, false);
    parser.flush();
    consume(false);
    // the floor of any per-packet API on this host: the same packets copied one by one into a
    // ring of the same size by a bare loop (no parser, no GPU)
    double floorSec = 0.0;
    {
      std::vector<char> ring((size_t)batch * 1206);
      std::vector<int64_t> rt((size_t)batch);
      const auto f0 = std::chrono::steady_clock::now();
      size_t pend = 0;
      for (size_t i = 0; i < n; ++i) {
        std::memcpy(ring.data() + pend * 1206, pk.data() + 1206 * i, 1206);
        rt[pend] = t[i];
        if (++pend == (size_t)batch) pend = 0;
      }
      floorSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - f0).count();
      touch += ring[17] + (double)rt[0] * 0.0;
    }
    parser.resetPipelineStats();
    const auto t0 = std::chrono::steady_clock::now();
    for (int p = 1; p <= passes; ++p) onePass(p, true);
    parser.flush();
    consume(true);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!parser.lastError().empty()) {
      std::cerr << "facade error: " << parser.lastError() << std::endl;
      return 1;
    }
    std::printf("{\"points\": %llu, \"frames\": %llu, \"seconds\": %.6f, \"packets\": %llu, \"passes\": %d, "
                "\"batch\": %d, \"pinned_pool_bytes\": %llu, \"touch\": %.3f, \"memcpy_floor_seconds_per_pass\": %.6f, "
                "\"host\": %s}\n",
                (unsigned long long)points, (unsigned long long)frames, sec,
                (unsigned long long)(n * (size_t)passes), passes, batch,
                (unsigned long long)vs::Arena::pooledBytes(), touch, floorSec, parser.pipelineStats().c_str());
    return 0;
  }
  if (mode == "latency" && argc >= 9) {
    // BASELINE.json configs[4], the C ABI called from C++ (no interpreter in the loop):
    // rotation-sized batches out of page-locked host memory, vs_submit + vs_wait back to back
    //   facade_driver latency <calib.xml> <packets.bin> <times.bin> <poses.bin> <packets/rotation>
    //                         <device> <with_points 0|1>
    CalibrationFile cal;
    std::string err;
    if (!cal.load(argv[2], &err)) {
      std::cerr << err << std::endl;
      return 1;
    }
    std::vector<char> pk = slurp(argv[3]);
    std::vector<char> tm = slurp(argv[4]);
    const int64_t rot = std::atoll(argv[6]);
    const int64_t n = (int64_t)(pk.size() / 1206), nRot = n / rot;
    const bool withPoints = std::atoi(argv[8]) != 0;
    if (rot < 1 || nRot < 20) return 2;
    std::shared_ptr<TransformManager> poses = loadPoses(argv[5]);
    std::vector<int64_t> pt;
    std::vector<double> trv;
    poses->snapshot(&pt, &trv);
    vs_ctx* ctx = nullptr;
    if (vs_create(std::atoi(argv[7]), 512, (int64_t)pt.size() + 8, 1, &ctx) != VS_OK) {
      std::cerr << vs_last_error(nullptr) << std::endl;
      return 1;
    }
    uint8_t* hp = nullptr;
    int64_t* ht = nullptr;
    void* cols[7] = {nullptr};
    const size_t colBytes[7] = {4, 4, 4, 1, 1, 2, 2};
    bool ok = vs_set_calibration(ctx, cal.rows, cal.n_rows, cal.n_enabled) == VS_OK &&
              (pt.empty() || vs_set_poses(ctx, pt.data(), trv.data(), (int64_t)pt.size()) == VS_OK) &&
              vs_host_alloc((uint64_t)n * 1206, (void**)&hp) == VS_OK &&
              vs_host_alloc((uint64_t)n * 8, (void**)&ht) == VS_OK;
    for (int c = 0; c < 7 && ok; ++c) ok = vs_host_alloc(512 * 384 * colBytes[c], &cols[c]) == VS_OK;
    if (!ok) {
      std::cerr << "setup failed: " << vs_last_error(ctx) << std::endl;
      return 1;
    }
    std::memcpy(hp, pk.data(), (size_t)n * 1206);
    std::memcpy(ht, tm.data(), (size_t)n * 8);
    vs_carry carry;
    vs_carry_init(&carry);
    std::vector<double> lat;
    uint64_t points = 0;
    int graphLaunches = 0;
    const auto w0 = std::chrono::steady_clock::now();
    for (int64_t r = 0; r < nRot; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      uint64_t ticket = 0;
      vs_result res;
      if (vs_submit(ctx, hp + r * rot * 1206, 1206, ht + r * rot, rot, 0, 0, 0, ht[0], &carry, &ticket) != VS_OK ||
          vs_wait(ctx, ticket, &res) != VS_OK) {
        std::cerr << "batch failed: " << vs_last_error(ctx) << std::endl;
        return 1;
      }
      if (withPoints && res.n_points > 0 &&
          vs_fetch_points(ctx, ticket, 0, res.n_points, (float*)cols[0], (float*)cols[1], (float*)cols[2],
                          (uint8_t*)cols[3], (uint8_t*)cols[4], (uint16_t*)cols[5], (uint16_t*)cols[6],
                          nullptr) != VS_OK) {
        std::cerr << "fetch failed: " << vs_last_error(ctx) << std::endl;
        return 1;
      }
      lat.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      carry = res.carry_out;
      points += (uint64_t)res.n_points;
      graphLaunches = res.reserved;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    lat.erase(lat.begin(), lat.begin() + 10);   // the first rotations warm the path up
    std::sort(lat.begin(), lat.end());
    auto pct = [&](double q) { return lat[(size_t)std::min<double>(lat.size() - 1, q * (lat.size() - 1) + 0.5)]; };
    std::printf("{\"p50_ms\": %.6f, \"p99_ms\": %.6f, \"max_ms\": %.6f, \"points_per_s\": %.1f, "
                "\"rotations\": %lld, \"graph_launches\": %d}\n",
                pct(0.50), pct(0.99), lat.back(), (double)points / sec, (long long)nRot, graphLaunches);
    for (int c = 0; c < 7; ++c) vs_host_free(cols[c]);
    vs_host_free(hp);
    vs_host_free(ht);
    vs_destroy(ctx);
    return 0;
  }
  if (mode == "online") {
    // BASELINE.json configs[4] for ONE stream: a paced sender thread plays the packet file over
    // loopback UDP at the sensor's rate; HDLManager::startOnline receives, stamps (TimeSolver),
    // decodes on the GPU (per-rotation flush) and stores the frames.
    //   facade_driver online <calib.xml> <packets.bin> <poses.bin|-> <seconds> <device> <port> [packets_per_s]
    if (argc < 8) return 2;
    const double seconds = std::atof(argv[5]);
    const int device = std::atoi(argv[6]), port = std::atoi(argv[7]);
    const double rate = argc > 8 ? std::atof(argv[8]) : 3472.0;
    std::vector<char> pk = slurp(argv[3]);
    const size_t have = pk.size() / 1206;
    const size_t n = std::min(have, (size_t)(seconds * rate));
    HDLManager mgr(40);
    const std::string scratch = std::string("/tmp/vs_online_") + std::to_string(getpid());
    mkdir(scratch.c_str(), 0777);
    mgr.setBufferDir(scratch, false);
    mgr.setCalibFile(argv[2]);
    mgr.setPorts(port, port + 1);
    {
      std::shared_ptr<TransformManager> poses = loadPoses(argv[4]);
      std::vector<int64_t> pt;
      std::vector<double> trv;
      poses->snapshot(&pt, &trv);
      for (size_t i = 0; i < pt.size(); ++i) {
        std::shared_ptr<PoseTransform> p(new PoseTransform);
        for (int k = 0; k < 3; ++k) {
          p->T[k] = trv[9 * i + k];
          p->R[k] = trv[9 * i + 3 + k];
          p->V[k] = trv[9 * i + 6 + k];
        }
        p->timestamp = ptime(pt[i]);
        p->seconds_pos = 0;
        mgr.getTransformMgr()->addTransform(p);
      }
    }
    mgr.getTimeSolver()->setClock([]() { return (int64_t)1467331200000000ll; });  // == synth.T0_US
    mgr.startOnline();
    HDLSource& src = *mgr.getHDLSource();
    if (!src.isRunning()) return 1;
    src.getHDLParser()->setDevice(device);
    src.getHDLParser()->setStorePackets(true);
    if (!src.getHDLParser()->prepare(48)) {  // the manager caches 40 frames
      std::cerr << "prepare failed: " << src.getHDLParser()->lastError() << std::endl;
      return 1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::thread sender([&]() {
      const int tx = socket(AF_INET, SOCK_DGRAM, 0);
      sockaddr_in to;
      std::memset(&to, 0, sizeof(to));
      to.sin_family = AF_INET;
      to.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
      to.sin_port = htons((uint16_t)port);
      for (size_t i = 0; i < n; ++i) {
        const auto due = t0 + std::chrono::nanoseconds((int64_t)(1e9 * (double)i / rate));
        auto now = std::chrono::steady_clock::now();
        if (due - now > std::chrono::microseconds(200)) std::this_thread::sleep_until(due - std::chrono::microseconds(100));
        while (std::chrono::steady_clock::now() < due) {
        }
        sendto(tx, pk.data() + 1206 * i, 1206, 0, reinterpret_cast<sockaddr*>(&to), sizeof(to));
      }
      close(tx);
    });
    sender.join();
    for (int i = 0; i < 500; ++i) {  // let the consumer drain
      uint64_t r, d, c;
      src.getCounters(&r, &d, &c);
      if (c + d >= n) break;
      usleep(10000);
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    mgr.stopOnline();
    uint64_t r, d, c;
    src.getCounters(&r, &d, &c);
    if (!src.getHDLParser()->lastError().empty()) {
      std::cerr << "facade error: " << src.getHDLParser()->lastError() << std::endl;
      return 1;
    }
    std::vector<double> proc, arr;
    src.getFrameLatencies(&proc, &arr);
    auto pct = [](std::vector<double> v, double q) {
      if (v.empty()) return 0.0;
      std::sort(v.begin(), v.end());
      return v[std::min(v.size() - 1, (size_t)(q * (double)(v.size() - 1) + 0.5))];
    };
    // the first second of the stream (10 rotations) is start-up: threads, first touches
    const double arrMaxAll = arr.empty() ? 0.0 : *std::max_element(arr.begin(), arr.end());
    if (proc.size() > 20) {
      proc.erase(proc.begin(), proc.begin() + 10);
      arr.erase(arr.begin(), arr.begin() + 10);
    }
    uint64_t points = 0;
    for (auto& f : mgr.getAllFrameMeta()) points += f->numberOfPoints();
    // rotations that took more than 5 ms (index after the start-up ones, ms): where a hiccup sits
    std::string slow = "[";
    for (size_t i = 0; i < proc.size(); ++i)
      if (proc[i] > 5000.0) {
        char buf[64];
        std::snprintf(buf, sizeof(buf), "%s[%zu, %.1f]", slow.size() > 1 ? ", " : "", i + 10, proc[i] / 1e3);
        slow += buf;
      }
    slow += "]";
    std::printf("{\"slow_rotations\": %s, ", slow.c_str());
    std::printf("\"packets_sent\": %llu, \"received\": %llu, \"dropped\": %llu, \"consumed\": %llu, "
                "\"frames\": %d, \"seconds\": %.3f, \"packets_per_s\": %.1f, "
                "\"rotation_p50_ms\": %.4f, \"rotation_p99_ms\": %.4f, \"rotation_max_ms\": %.4f, "
                "\"from_arrival_p50_ms\": %.4f, \"from_arrival_p99_ms\": %.4f, "
                "\"from_arrival_max_ms_incl_startup\": %.4f, \"points_cached\": %llu}\n",
                (unsigned long long)n, (unsigned long long)r, (unsigned long long)d, (unsigned long long)c,
                mgr.getNumberOfFrames(), wall, (double)n / wall, pct(proc, 0.5) / 1e3, pct(proc, 0.99) / 1e3,
                pct(proc, 1.0) / 1e3, pct(arr, 0.5) / 1e3, pct(arr, 0.99) / 1e3, arrMaxAll / 1e3,
                (unsigned long long)points);
    return 0;
  }
  if (mode == "udp") {
    // the whole online path: UDP -> HDLSource -> TimeSolver -> HDLParser (GPU) -> HDLManager
    //   facade_driver udp <calib.xml> <port> <poses.bin|-> <out.bin> <n_expected>
    if (argc < 7) return 2;
    const size_t want = (size_t)std::atoll(argv[6]);
    // HDLManager::startOnline wires HDLSource / INSSource / TimeSolver / TransformManager
    HDLManager mgr(1000);
    const std::string scratch = std::string(argv[5]) + ".dir";
    mkdir(scratch.c_str(), 0777);
    mgr.setBufferDir(scratch, false);   // stopOnline writes the .hdlmeta here
    mgr.setCalibFile(argv[2]);
    mgr.setPorts(std::atoi(argv[3]), std::atoi(argv[3]) + 1);
    {
      std::shared_ptr<TransformManager> poses = loadPoses(argv[4]);
      std::vector<int64_t> pt;
      std::vector<double> trv;
      poses->snapshot(&pt, &trv);
      for (size_t i = 0; i < pt.size(); ++i) {
        std::shared_ptr<PoseTransform> p(new PoseTransform);
        for (int k = 0; k < 3; ++k) {
          p->T[k] = trv[9 * i + k];
          p->R[k] = trv[9 * i + 3 + k];
          p->V[k] = trv[9 * i + 6 + k];
        }
        p->timestamp = ptime(pt[i]);
        p->seconds_pos = 0;
        mgr.getTransformMgr()->addTransform(p);
      }
    }
    mgr.getTimeSolver()->setClock([]() { return (int64_t)1467331200000000ll; });  // == synth.T0_US
    mgr.startOnline();
    HDLSource& src = *mgr.getHDLSource();
    if (!src.isRunning()) return 1;
    std::cout << "ready" << std::endl;
    for (int i = 0; i < 3000; ++i) {  // up to 30 s
      uint64_t r, d, c;
      src.getCounters(&r, &d, &c);
      if (c >= want) break;
      usleep(10000);
    }
    mgr.stopOnline();
    uint64_t r, d, c;
    src.getCounters(&r, &d, &c);
    std::cerr << "received " << r << " dropped " << d << " consumed " << c << std::endl;
    if (!src.getHDLParser()->lastError().empty()) {
      std::cerr << "facade error: " << src.getHDLParser()->lastError() << std::endl;
      return 1;
    }
    std::vector<std::shared_ptr<HDLFrame> > all = mgr.getAllFrameMeta();
    std::ofstream os(argv[5], std::ios::binary);
    const int32_t nf = (int32_t)all.size();
    os.write((const char*)&nf, 4);
    for (auto& f : all) dumpFrame(os, *f);
    return 0;
  }
  if (mode == "manager_shards" && argc >= 7) {
    // HDLManager::setDevices + loadOffline: the recording spread over several CUDA contexts
    //   facade_driver manager_shards <calib.xml> <file.pcap> <poses.bin|-> <out.bin> <dev,dev,...>
    // every frame comes back through getRangeBetween (all shards decode at once)
    HDLManager mgr(1 << 20);
    mgr.setCalibFile(argv[2]);
    std::shared_ptr<TransformManager> poses = loadPoses(argv[4]);
    std::vector<int64_t> pt;
    std::vector<double> trv;
    poses->snapshot(&pt, &trv);
    for (size_t i = 0; i < pt.size(); ++i) {
      std::shared_ptr<PoseTransform> p(new PoseTransform);
      for (int k = 0; k < 3; ++k) {
        p->T[k] = trv[9 * i + k];
        p->R[k] = trv[9 * i + 3 + k];
        p->V[k] = trv[9 * i + 6 + k];
      }
      p->timestamp = ptime(pt[i]);
      p->seconds_pos = 0;
      mgr.getTransformMgr()->addTransform(p);
    }
    std::vector<int> devs;
    for (const char* c = argv[6]; *c;) {
      devs.push_back(std::atoi(c));
      while (*c && *c != ',') ++c;
      if (*c == ',') ++c;
    }
    mgr.setDevices(devs);
    const auto t0 = std::chrono::steady_clock::now();
    mgr.loadOffline("-", argv[3]);
    const auto t1 = std::chrono::steady_clock::now();
    std::cerr << "shards " << mgr.getNumberOfShards() << std::endl;
    std::vector<std::shared_ptr<HDLFrame> > meta = mgr.getAllFrameMeta();
    if (meta.empty()) return 1;
    ptime a = meta.front()->timestamp, b = meta.back()->timestamp;
    std::vector<HDLFramePtr> all = mgr.getRangeBetween(a, b);
    const auto t2 = std::chrono::steady_clock::now();
    std::cerr << "load_ms " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " range_ms "
              << std::chrono::duration<double, std::milli>(t2 - t1).count() << std::endl;
    if (all.size() != meta.size()) return 1;
    std::ofstream os(argv[5], std::ios::binary);
    const int32_t nf = (int32_t)all.size();
    os.write((const char*)&nf, 4);
    for (auto& f : all) {
      if (!f) return 1;
      dumpFrame(os, *f);
    }
    return 0;
  }
  if (mode == "manager" || mode == "manager_file") {
    // HDLManager::loadOffline on a packet file, then every frame through getFrameAt
    //   facade_driver manager <calib.xml> <file.pcap> <poses.bin|-> <out.bin>
    // manager: rotations are decoded out of the recording resident in HBM;
    // manager_file: the recording is dropped after the index, getFrame re-reads the file.
    HDLManager mgr(8);
    mgr.setCalibFile(argv[2]);
    std::shared_ptr<TransformManager> poses = loadPoses(argv[4]);
    std::vector<int64_t> pt;
    std::vector<double> trv;
    poses->snapshot(&pt, &trv);
    for (size_t i = 0; i < pt.size(); ++i) {
      std::shared_ptr<PoseTransform> p(new PoseTransform);
      for (int k = 0; k < 3; ++k) {
        p->T[k] = trv[9 * i + k];
        p->R[k] = trv[9 * i + 3 + k];
        p->V[k] = trv[9 * i + 6 + k];
      }
      p->timestamp = ptime(pt[i]);
      p->seconds_pos = 0;
      mgr.getTransformMgr()->addTransform(p);
    }
    mgr.loadOffline("-", argv[3]);  // "-": no INS text file, the poses above stay
    const bool resident = mgr.getParser()->hasRecording(argv[3]);
    std::cerr << "resident " << (resident ? 1 : 0) << std::endl;
    if (mode == "manager_file") mgr.getParser()->unloadRecording();
    std::vector<std::shared_ptr<HDLFrame> > meta = mgr.getAllFrameMeta();
    std::ofstream os(argv[5], std::ios::binary);
    const int32_t nf = (int32_t)meta.size();
    os.write((const char*)&nf, 4);
    for (auto& m : meta) {
      ptime t = m->timestamp;
      HDLFramePtr f = mgr.getFrameAt(t);
      if (!f) return 1;
      dumpFrame(os, *f);
    }
    if (!mgr.getParser()->lastError().empty()) {
      std::cerr << "facade error: " << mgr.getParser()->lastError() << std::endl;
      return 1;
    }
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
