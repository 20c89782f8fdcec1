// HDLManager.cpp -- see HDLManager.h.  Host logic only; every decode goes through HDLParser
// (C ABI -> sm_100a kernels).
#include "HDLManager.h"

#include "../../include/veloslam_b200.h"

#include <dirent.h>
#include <sys/stat.h>
#include <sys/time.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <thread>

namespace {
ptime localTimeNow() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return ptime((int64_t)tv.tv_sec * 1000000ll + tv.tv_usec);
}
bool isDirectory(const std::string& d) {
  struct stat st;
  return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
std::string parentPath(const std::string& f) {
  const size_t slash = f.find_last_of('/');
  return slash == std::string::npos ? std::string(".") : f.substr(0, slash);
}

// .hdlmeta record (reference HDLFrame.cxx:160-190): timestamp, filenameTime, fileStartPos,
// skips, isOnHardDrive, then the car pose as in .insmeta (type_defs.cxx:4-33)
void writePoseRecord(std::ofstream& os, const PoseTransform& p) {
  for (int i = 0; i < 3; ++i) {
    os.write(reinterpret_cast<const char*>(p.T + i), sizeof(double));
    os.write(reinterpret_cast<const char*>(p.R + i), sizeof(double));
    os.write(reinterpret_cast<const char*>(p.V + i), sizeof(double));
  }
  os.write(reinterpret_cast<const char*>(&p.timestamp.us), sizeof(int64_t));
  os.write(reinterpret_cast<const char*>(&p.week_number), sizeof(p.week_number));
  os.write(reinterpret_cast<const char*>(&p.milliseconds), sizeof(p.milliseconds));
  os.write(reinterpret_cast<const char*>(&p.week_number_pos), sizeof(p.week_number_pos));
  os.write(reinterpret_cast<const char*>(&p.seconds_pos), sizeof(p.seconds_pos));
}
bool readPoseRecord(std::ifstream& is, PoseTransform& p) {
  for (int i = 0; i < 3; ++i) {
    is.read(reinterpret_cast<char*>(p.T + i), sizeof(double));
    is.read(reinterpret_cast<char*>(p.R + i), sizeof(double));
    is.read(reinterpret_cast<char*>(p.V + i), sizeof(double));
  }
  is.read(reinterpret_cast<char*>(&p.timestamp.us), sizeof(int64_t));
  is.read(reinterpret_cast<char*>(&p.week_number), sizeof(p.week_number));
  is.read(reinterpret_cast<char*>(&p.milliseconds), sizeof(p.milliseconds));
  is.read(reinterpret_cast<char*>(&p.week_number_pos), sizeof(p.week_number_pos));
  is.read(reinterpret_cast<char*>(&p.seconds_pos), sizeof(p.seconds_pos));
  return (bool)is;
}
void writeFrameRecord(std::ofstream& os, const HDLFrame& f) {
  os.write(reinterpret_cast<const char*>(&f.timestamp.us), sizeof(int64_t));
  os.write(reinterpret_cast<const char*>(&f.filenameTime.us), sizeof(int64_t));
  os.write(reinterpret_cast<const char*>(&f.fileStartPos), sizeof(f.fileStartPos));
  os.write(reinterpret_cast<const char*>(&f.skips), sizeof(f.skips));
  os.write(reinterpret_cast<const char*>(&f.isOnHardDrive), sizeof(f.isOnHardDrive));
  writePoseRecord(os, *f.carpose);
}
bool readFrameRecord(std::ifstream& is, HDLFrame& f) {
  is.read(reinterpret_cast<char*>(&f.timestamp.us), sizeof(int64_t));
  is.read(reinterpret_cast<char*>(&f.filenameTime.us), sizeof(int64_t));
  is.read(reinterpret_cast<char*>(&f.fileStartPos), sizeof(f.fileStartPos));
  is.read(reinterpret_cast<char*>(&f.skips), sizeof(f.skips));
  is.read(reinterpret_cast<char*>(&f.isOnHardDrive), sizeof(f.isOnHardDrive));
  return readPoseRecord(is, *f.carpose);
}
}  // namespace

HDLManager::HDLManager(int capacity)
    : hardDriveBuffer(&hardDriveBuffer1), cacheCounter(0), hasNewData(false), bufferSize(100),
      maxCacheSize((size_t)capacity), bufferDirName("/tmp/"), isUsingBuffer1(true),
      writerIdle(true), fileBufferMode(false), packetWriter(new vtkPacketFileWriter),
      transMgr(new TransformManager), hdlParser(new HDLParser), timeSolver(new TimeSolver),
      hdlPort_(2368), insPort_(6777), metaSerial(0) {
  hdlParser->setTransformMgr(transMgr);
}

void HDLManager::setPorts(int hdlPort, int insPort) {
  hdlPort_ = hdlPort;
  insPort_ = insPort;
}

void HDLManager::startOnline(bool shouldSwap) {
  this->loadHDLMeta();
  this->loadINSMeta();
  if (!insSrc) {
    insSrc.reset(new INSSource(insPort_));
    insSrc->setTimeSolver(timeSolver);
    insSrc->setTransformManager(transMgr);
  }
  if (!hdlSrc) {
    hdlSrc.reset(new HDLSource(hdlPort_));
    hdlSrc->setHDLManager(this);
    hdlSrc->setTimeSolver(timeSolver);
    hdlSrc->setTransformManager(transMgr);
    if (!calibFile_.empty()) hdlSrc->setCorrectionsFile(calibFile_);
  }
  insSrc->start();
  hdlSrc->start();
  if (shouldSwap) startSwaping();
}

void HDLManager::stopOnline() {
  if (hdlSrc) hdlSrc->stop();
  if (insSrc) insSrc->stop();
  if (fileBufferMode) {
    this->flushFileBuffer();
    this->stopSwaping();
  }
  this->saveHDLMeta();
}

HDLManager::~HDLManager() { delete packetWriter; }

void HDLManager::loadOffline(const std::string& insTxt, const std::string& pcapfile) {
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    frames.clear();
  }
  transMgr->loadFromTxtFile(insTxt, true);
  std::cout << "Read " << transMgr->getNumberOfTransforms() << " transforms.\n";
  shards_.clear();
  shardFile_.clear();
  // (a sharded attempt that fails may already have renamed the file after its first packet)
  std::string file = pcapfile;
  if (devices_.size() > 1 && this->loadOfflineSharded(pcapfile, &file)) return;
  // the whole recording goes to HBM once; when the file is not a fixed-stride packet file the
  // parser falls back to reading it record by record, like the reference
  hdlParser->loadRecording(file);
  auto frameVec = hdlParser->readFrameInformation(file);
  /* because readFrameInformation() can't determine carpose for each frame
   * we need to ask transform manager */
  for (auto& f : frameVec) {
    transMgr->interpolateTransform(f->timestamp, f->carpose.get());
    this->addFrame(f);
  }
  std::cout << "Read " << this->getNumberOfFrames() << " frames." << std::endl;
  this->setBufferDir(parentPath(file), false);
}

void HDLManager::setDevices(const std::vector<int>& cudaDevices) {
  devices_ = cudaDevices;
  if (!devices_.empty()) hdlParser->setDevice(devices_[0]);
}

// One range of records per GPU.  The reference's index (HDLParser.cxx:1080-1149) is a scan of
// every firing block against the azimuth before it, so a range needs one record of lead-in and
// nothing else; a rotation that starts in a range is decoded by that range's GPU out of the
// kShardTail records kept past the range end.  false: not a fixed-stride packet file (the
// caller then takes the one-GPU path, which also handles the record-by-record fallback).
bool HDLManager::loadOfflineSharded(const std::string& pcapfile, std::string* resolved) {
  // the file is named after its first packet (touch renames it when it is not)
  std::string file = pcapfile;
  auto head = hdlParser->readFrameInformation(pcapfile, true);
  if (head.empty()) return false;
  struct stat st;
  if (stat(file.c_str(), &st) != 0) {
    file = parentPath(pcapfile) + "/" + to_iso_string(head[0]->filenameTime) + ".pcap";
    if (stat(file.c_str(), &st) != 0) return false;
  }
  *resolved = file;
  const int64_t body = (int64_t)st.st_size - PCAP_GLOBAL_HEADER_LEN;
  if (body <= 0 || body % VS_PCAP_RECORD_BYTES != 0) return false;
  const int64_t n = body / VS_PCAP_RECORD_BYTES;
  const int world = (int)devices_.size();
  std::vector<Shard> shards((size_t)world);
  std::vector<int64_t> lead((size_t)world, 0);
  for (int g = 0; g < world; ++g) {
    if (vs_shard_range(n, world, g, 1, &shards[g].first, &lead[g], &shards[g].end) != VS_OK) return false;
    if (shards[g].end == shards[g].first) continue;   // fewer records than GPUs
    if (g == 0) {
      shards[g].parser = hdlParser;
    } else {
      shards[g].parser.reset(new HDLParser);
      shards[g].parser->setTransformMgr(transMgr);
      if (!calibFile_.empty()) shards[g].parser->setCorrectionsFile(calibFile_);
    }
    shards[g].parser->setDevice(devices_[g]);
  }
  std::vector<std::vector<std::shared_ptr<HDLFrame> > > index((size_t)world);
  std::vector<char> ok((size_t)world, 1);
  std::vector<std::thread> workers;
  for (int g = 0; g < world; ++g) {
    if (!shards[g].parser) continue;
    workers.emplace_back([&, g] {
      HDLParser& p = *shards[g].parser;
      const int64_t lo = shards[g].first - lead[g];
      ok[g] = p.loadRecordingRange(file, lo, shards[g].end - lo + kShardTail);
      if (ok[g]) {
        index[g] = p.readFrameInformation(file);
        ok[g] = !index[g].empty() && p.lastError().empty();
      }
    });
  }
  for (auto& w : workers) w.join();
  for (int g = 0; g < world; ++g)
    if (!ok[g]) {
      std::cerr << "HDLManager: range " << g << " of " << file << " (records " << shards[g].first << ".."
                << shards[g].end << ", device " << devices_[g] << ") could not be loaded or indexed: "
                << shards[g].parser->lastError() << "; falling back to one GPU" << std::endl;
      hdlParser->unloadRecording();
      return false;
    }
  shards_.swap(shards);
  shardFile_ = file;
  // the frames of a range are the ones that start inside it: the lead-in record's wraps belong
  // to the range before, the tail's to the range after
  for (int g = 0; g < world; ++g)
    for (auto& f : index[g]) {
      const int64_t rec = (f->fileStartPos - PCAP_GLOBAL_HEADER_LEN) / VS_PCAP_RECORD_BYTES;
      if (rec < shards_[g].first || rec >= shards_[g].end) continue;
      f->filenameTime = head[0]->filenameTime;
      transMgr->interpolateTransform(f->timestamp, f->carpose.get());
      this->addFrame(f);
    }
  std::cout << "Read " << this->getNumberOfFrames() << " frames on " << world << " GPUs." << std::endl;
  this->setBufferDir(parentPath(file), false);
  return true;
}

HDLParser* HDLManager::parserFor(const HDLFrame& frame, const std::string& pcap) {
  if (shards_.empty() || pcap != shardFile_) return hdlParser.get();
  const int64_t rec = (frame.fileStartPos - PCAP_GLOBAL_HEADER_LEN) / VS_PCAP_RECORD_BYTES;
  for (auto& s : shards_)
    if (s.parser && rec >= s.first && rec < s.end) return s.parser.get();
  return hdlParser.get();
}

void HDLManager::touchPcap(const std::string& pcapfile) { hdlParser->readFrameInformation(pcapfile, true); }

int HDLManager::getNumberOfFrames() {
  std::lock_guard<std::mutex> lock(framesMutex);
  return (int)frames.size();
}
int HDLManager::getNumberOfTransforms() { return transMgr->getNumberOfTransforms(); }

void HDLManager::scanBufferDir() {
  DIR* d = opendir(bufferDirName.c_str());
  if (!d) return;
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    const size_t dot = name.find_last_of('.');
    if (dot == std::string::npos) continue;
    const std::string ext = name.substr(dot);
    if (ext == HDL_META_EXT_NAME)
      hdlMetaNames.insert(bufferDirName + name);
    else if (ext == INS_META_EXT_NAME)
      insMetaNames.insert(bufferDirName + name);
    else if (ext == ".pcap")
      bufferFileNames.insert(name);
  }
  closedir(d);
}

void HDLManager::switchBuffer() {
  {
    std::lock_guard<std::mutex> lock(writerMutex);
    hardDriveBuffer = isUsingBuffer1 ? &hardDriveBuffer2 : &hardDriveBuffer1;
    isUsingBuffer1 = !isUsingBuffer1;
    writerIdle = false;
  }
  // the reference wakes its writer thread here; this facade writes in place
  writePackets();
}

void HDLManager::setCalibFile(std::string filename) {
  calibFile_ = filename;
  if (hdlSrc) hdlSrc->setCorrectionsFile(filename);
  hdlParser->setCorrectionsFile(filename);
  for (auto& s : shards_)
    if (s.parser && s.parser != hdlParser) s.parser->setCorrectionsFile(filename);
}

void HDLManager::addFrame(std::shared_ptr<HDLFrame> frame) {
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    frames.addData(frame);
    hasNewData = true;
  }
  cond_.notify_one();
  const bool toDisk = fileBufferMode && !frame->isOnHardDrive;
  if (!toDisk) {
    pushCache(frame);
    return;
  }
  bool full;
  {
    std::lock_guard<std::mutex> lock(writerMutex);
    hardDriveBuffer->push_back(frame);
    full = writerIdle && hardDriveBuffer->size() == bufferSize;
  }
  if (full) switchBuffer();
}

// A frame in memory is handed out as it is; one that only exists on disk is decoded again --
// out of HBM when its recording is resident, else from the pcap file (HDLManager.cxx:195-211).
HDLFramePtr HDLManager::prepareFrame(std::shared_ptr<HDLFrame> frame) {
  if (!frame) return HDLFramePtr();
  if (frame->isInMemory) return HDLFramePtr(frame.get());
  if (!frame->isOnHardDrive) return HDLFramePtr();
  const std::string pcap = bufferDirName + to_iso_string(frame->filenameTime) + ".pcap";
  if (!parserFor(*frame, pcap)->getFrame(frame, pcap, frame->fileStartPos, frame->skips)) return HDLFramePtr();
  frame->isInMemory = true;
  pushCache(frame);
  return HDLFramePtr(frame.get());
}

HDLFramePtr HDLManager::waitForFrame(int64_t micro) {
  std::shared_ptr<HDLFrame> newest;
  {
    std::unique_lock<std::mutex> lock(framesMutex);
    cond_.wait_for(lock, std::chrono::microseconds(micro), [this] { return hasNewData; });
    if (!hasNewData) return HDLFramePtr();
    hasNewData = false;
    newest = frames.back();
  }
  return prepareFrame(newest);
}

HDLFramePtr HDLManager::getRecentFrame() {
  std::shared_ptr<HDLFrame> newest;
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    newest = frames.back();
  }
  return prepareFrame(newest);
}
// the timeline is read under framesMutex (addFrame inserts into it from the consumer thread);
// the shared_ptr is copied out before the decode
HDLFramePtr HDLManager::getFrameAt(ptime& t) {
  std::shared_ptr<HDLFrame> f;
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    f = frames.getExactDataAt(t);
  }
  return prepareFrame(f);
}
HDLFramePtr HDLManager::getFrameNear(ptime& t) {
  std::shared_ptr<HDLFrame> f;
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    f = frames.getNearestData(t);
  }
  return prepareFrame(f);
}

std::vector<std::shared_ptr<HDLFrame> > HDLManager::getAllFrameMeta() {
  std::lock_guard<std::mutex> lock(framesMutex);
  return frames.getAll();
}

std::vector<HDLFramePtr> HDLManager::getRangeBetween(ptime& a, ptime& b) {
  // inclusive on both ends (HDLManager.h:148, TimeLine.h:316-382)
  std::vector<std::shared_ptr<HDLFrame> > vec;
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    for (const auto& f : frames.items())
      if (f->timestamp >= a && f->timestamp <= b) vec.push_back(f);
  }
  std::vector<HDLFramePtr> result(vec.size());
  if (shards_.size() < 2) {
    for (size_t i = 0; i < vec.size(); ++i) result[i] = prepareFrame(vec[i]);
    return result;
  }
  // a recording spread over several GPUs: every GPU decodes the frames it holds, at once
  std::vector<std::vector<size_t> > perParser;
  std::vector<HDLParser*> parsers;
  for (size_t i = 0; i < vec.size(); ++i) {
    const std::string pcap = bufferDirName + to_iso_string(vec[i]->filenameTime) + ".pcap";
    HDLParser* p = parserFor(*vec[i], pcap);
    size_t k = 0;
    while (k < parsers.size() && parsers[k] != p) ++k;
    if (k == parsers.size()) {
      parsers.push_back(p);
      perParser.emplace_back();
    }
    perParser[k].push_back(i);
  }
  std::vector<std::thread> workers;
  for (size_t k = 0; k < parsers.size(); ++k)
    workers.emplace_back([&, k] {
      for (size_t i : perParser[k]) result[i] = prepareFrame(vec[i]);
    });
  for (auto& w : workers) w.join();
  return result;
}

void HDLManager::setBufferSize(size_t n) { bufferSize = n; }
size_t HDLManager::getBufferSize() { return bufferSize; }

bool HDLManager::setBufferDir(std::string dirname, bool shouldCreateSubDir) {
  if (!isDirectory(dirname)) {
    std::cerr << "Name of directory invalid" << std::endl;
    return false;
  }
  if (dirname.back() != '/') dirname.append("/");
  if (shouldCreateSubDir) {
    dirname += to_iso_string(localTimeNow()) + "/";
    if (mkdir(dirname.c_str(), 0777) != 0) {
      std::cerr << "Failed to create sub directory inside: " << dirname << std::endl;
      return false;
    }
  }
  std::lock_guard<std::mutex> lock(writerMutex);
  this->bufferDirName = dirname;
  return true;
}

void HDLManager::resetBufferDir() {
  std::lock_guard<std::mutex> lock(writerMutex);
  this->bufferDirName = "/tmp/";
}

void HDLManager::setFileBufferMode(bool m) {
  std::lock_guard<std::mutex> lock(framesMutex);
  fileBufferMode = m;
}

void HDLManager::flushFileBuffer() { switchBuffer(); }

// One pcap file per filled buffer, named after its first packet; every frame of the buffer gets
// the file's name time and the byte offset of its first record (HDLManager.cxx:318-371).
bool HDLManager::writePackets() {
  Buffer* full = nullptr;
  std::string path;
  ptime stamp;
  {
    std::lock_guard<std::mutex> lock(writerMutex);
    full = isUsingBuffer1 ? &hardDriveBuffer2 : &hardDriveBuffer1;  // the one NOT being filled
    const bool nothing = full->empty() || full->front()->packets.empty();
    if (nothing) {
      writerIdle = true;
      return false;
    }
    stamp = full->front()->packets.front().first;
    path = bufferDirName + to_iso_string(stamp) + ".pcap";
  }
  if (packetWriter->isOpen()) packetWriter->close();
  packetWriter->open(path);
  int64_t recordPos = PCAP_GLOBAL_HEADER_LEN;  // fpos of the next record, kept by hand
  for (const std::shared_ptr<HDLFrame>& frame : *full) {
    for (const auto& pkt : frame->packets)
      packetWriter->writePacket(reinterpret_cast<const unsigned char*>(pkt.second.data()),
                                (unsigned int)pkt.second.size(), pkt.first);
    frame->filenameTime = stamp;
    frame->fileStartPos = recordPos;
    frame->isOnHardDrive = true;
    recordPos += (int64_t)frame->packets.size() * PCAP_PACKET_LEN;
    std::lock_guard<std::mutex> lock(cacheMutex);
    cache.push_back(frame.get());
    ++cacheCounter;
  }
  packetWriter->close();
  bufferFileNames.insert(path);
  full->clear();
  updateCacheSize();
  {
    std::lock_guard<std::mutex> lock(writerMutex);
    writerIdle = true;
  }
  return true;
}

void HDLManager::startSwaping() {
  fileBufferMode = true;
  writerIdle = true;
}
void HDLManager::stopSwaping() {}

// the cache is touched by the HDLSource consumer thread (addFrame) and by user threads
// (prepareFrame, cleanCache): one mutex guards the deque and its counter
void HDLManager::pushCache(std::shared_ptr<HDLFrame>& frame) {
  std::lock_guard<std::mutex> lock(cacheMutex);
  cache.push_back(frame.get());
  ++cacheCounter;
  updateCacheSizeLocked();
}

// Evict from the front until the cache fits; a frame an end user still holds (count != 0) goes
// back to the end, at most ten times per call (HDLManager.cxx:400-421).
void HDLManager::updateCacheSize() {
  std::lock_guard<std::mutex> lock(cacheMutex);
  updateCacheSizeLocked();
}
void HDLManager::updateCacheSizeLocked() {
  for (int putBacks = 0; cacheCounter > (int)maxCacheSize && putBacks < 10 && !cache.empty();) {
    HDLFrame* oldest = cache.front();
    cache.pop_front();
    if (oldest->count != 0) {
      cache.push_back(oldest);
      ++putBacks;
      continue;
    }
    oldest->clear();
    --cacheCounter;
  }
}

void HDLManager::cleanCache() {
  std::lock_guard<std::mutex> lock(cacheMutex);
  const size_t tmp = maxCacheSize;
  maxCacheSize = 0;
  updateCacheSizeLocked();
  maxCacheSize = tmp;
}

bool HDLManager::saveHDLMeta() {
  const std::string filename = this->bufferDirName + to_iso_string(localTimeNow()) + "-" +
                               std::to_string(metaSerial++) + HDL_META_EXT_NAME;
  std::ofstream ofs(filename, std::ios::binary);
  if (!ofs) return false;
  std::vector<std::shared_ptr<HDLFrame> > all;
  {
    std::lock_guard<std::mutex> lock(framesMutex);
    all = frames.getAll();
  }
  for (const auto& f : all) writeFrameRecord(ofs, *f);
  return true;
}

bool HDLManager::saveINSMeta() {
  const std::string filename = this->bufferDirName + to_iso_string(localTimeNow()) + "-" +
                               std::to_string(metaSerial++) + INS_META_EXT_NAME;
  return transMgr->writeToMetaFile(filename);
}

bool HDLManager::loadHDLMeta() {
  scanBufferDir();
  if (hdlMetaNames.empty()) return false;
  for (const auto& name : hdlMetaNames) {
    std::ifstream ifs(name, std::ios::binary);
    while (true) {
      std::shared_ptr<HDLFrame> item(new HDLFrame);
      if (!readFrameRecord(ifs, *item)) break;
      std::lock_guard<std::mutex> lock(framesMutex);
      frames.addData(item);
    }
  }
  return true;
}

bool HDLManager::loadINSMeta() {
  scanBufferDir();
  if (insMetaNames.empty()) return false;
  for (const auto& f : insMetaNames) transMgr->loadFromMetaFile(f);
  return true;
}
