#include "HDLSource.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

#include "HDLManager.h"

namespace {
const int kSlotBytes = 1500;   // like the reference's rxBuffer: larger than a packet so that
                               // oversized datagrams are noticed (HDLSource.cxx:375-378)
const int kRingSlots = 4096;   // 1.2 s of HDL-64E traffic
}  // namespace

class HDLSource::vsInternal {
 public:
  vsInternal()
      : parser(new HDLParser), timeSolver(new TimeSolver), hdlMgr(nullptr), sock(-1),
        running(false), head(0), tail(0), received(0), dropped(0), consumed(0),
        ring((size_t)kRingSlots * kSlotBytes), lengths(kRingSlots) {}

  static int64_t nowUs() {
    return std::chrono::duration_cast<std::chrono::microseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  // HDLSource.cxx:209-225
  void handleSensorData(const unsigned char* data, unsigned int length, int64_t arrivedUs = 0) {
    if (length != 1206) return;
    std::lock_guard<std::mutex> lock(parserMutex);
    const int64_t t0 = nowUs();
    uint32_t gps;
    std::memcpy(&gps, data + 1200, 4);
    const ptime timestamp = timeSolver->calcTimestamp(gps);
    if (callback) {
      callback(data, length, timestamp);
      return;
    }
    parser->processHDLPacket(const_cast<unsigned char*>(data), length, timestamp);
    std::deque<std::shared_ptr<HDLFrame> > fr = parser->getAllFrames();
    if (fr.size()) {
      // the reference adds getAllFrames().back(): with its per-packet parser at most one frame is
      // ever waiting.  A packet that closes several frames (or a pipelined parser handing back a
      // whole batch) leaves more than one here; none of them is dropped.
      if (hdlMgr)
        for (auto& f : fr) hdlMgr->addFrame(f);
      parser->clearAllFrames();
      // per rotation: packet that closed it handed to the parser -> frame in the manager, and the
      // same from the packet's arrival on the socket (includes the wait in the ring)
      const int64_t t1 = nowUs();
      std::lock_guard<std::mutex> ll(latencyMutex);
      frameLatencyUs.push_back((double)(t1 - t0));
      arrivalLatencyUs.push_back(arrivedUs ? (double)(t1 - arrivedUs) : 0.0);
    }
  }

  void receiveLoop() {
    while (running.load()) {
      const uint64_t h = head.load(std::memory_order_relaxed);
      // Ring full (consumer kRingSlots packets behind): slot h % kRingSlots is the oldest
      // unconsumed packet, possibly being read right now -- receive into the scratch buffer and
      // drop the newest packet instead, as a full socket buffer would.
      const bool full = h - tail.load(std::memory_order_acquire) >= (uint64_t)kRingSlots;
      unsigned char* slot = full ? scratch : ring.data() + (size_t)(h % kRingSlots) * kSlotBytes;
      const ssize_t n = recv(sock, slot, kSlotBytes, 0);  // SO_RCVTIMEO wakes it up to re-check
      if (n <= 0) continue;
      ++received;
      if (full) {
        ++dropped;
        continue;
      }
      lengths[h % kRingSlots] = (unsigned int)n;
      arrived[h % kRingSlots] = nowUs();
      head.store(h + 1, std::memory_order_release);
      {
        std::lock_guard<std::mutex> lock(wakeMutex);
      }
      wake.notify_one();
    }
  }

  void consumeLoop() {
    while (true) {
      uint64_t t = tail.load(std::memory_order_relaxed);
      if (t == head.load(std::memory_order_acquire)) {
        if (!running.load()) break;
        std::unique_lock<std::mutex> lock(wakeMutex);
        wake.wait_for(lock, std::chrono::milliseconds(20));
        continue;
      }
      handleSensorData(ring.data() + (size_t)(t % kRingSlots) * kSlotBytes, lengths[t % kRingSlots],
                       arrived[t % kRingSlots]);
      ++consumed;
      tail.store(t + 1, std::memory_order_release);
    }
  }

  std::shared_ptr<HDLParser> parser;
  std::shared_ptr<TimeSolver> timeSolver;
  HDLManager* hdlMgr;
  std::function<void(const unsigned char*, unsigned int, ptime)> callback;
  std::mutex parserMutex;  // hold this when running parser code or modifying its internals
  int sock;
  std::atomic<bool> running;
  std::atomic<uint64_t> head, tail;
  std::atomic<uint64_t> received, dropped, consumed;
  std::vector<unsigned char> ring;
  unsigned char scratch[1500];  // where a packet lands when the ring is full (then dropped)
  std::vector<unsigned int> lengths;
  std::vector<int64_t> arrived = std::vector<int64_t>(kRingSlots);  // steady-clock arrival time per slot
  std::mutex latencyMutex;
  std::vector<double> frameLatencyUs, arrivalLatencyUs;
  std::thread receiver, consumer;
  std::mutex wakeMutex;
  std::condition_variable wake;
};

HDLSource::HDLSource(int _port) : sensorPort(_port), internal_(new vsInternal) {}
HDLSource::~HDLSource() {
  this->stop();
  delete internal_;
}

void HDLSource::start() {
  vsInternal* in = internal_;
  if (in->running.load()) return;
  in->sock = socket(AF_INET, SOCK_DGRAM, 0);
  if (in->sock < 0) {
    std::cerr << "HDLSource: cannot create the UDP socket" << std::endl;
    return;
  }
  int one = 1, rcvbuf = 8 << 20;
  setsockopt(in->sock, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  setsockopt(in->sock, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
  timeval tv = {0, 100000};  // 100 ms: lets stop() end the receiver
  setsockopt(in->sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  sockaddr_in addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = htons((uint16_t)sensorPort);
  if (bind(in->sock, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    std::cerr << "HDLSource: cannot bind UDP port " << sensorPort << std::endl;
    close(in->sock);
    in->sock = -1;
    return;
  }
  in->running.store(true);
  in->consumer = std::thread(&vsInternal::consumeLoop, in);
  in->receiver = std::thread(&vsInternal::receiveLoop, in);
}

void HDLSource::stop() {
  vsInternal* in = internal_;
  if (!in->running.exchange(false)) return;
  if (in->receiver.joinable()) in->receiver.join();
  in->wake.notify_all();
  if (in->consumer.joinable()) in->consumer.join();  // drains what is still in the ring
  if (in->sock >= 0) close(in->sock);
  in->sock = -1;
  std::lock_guard<std::mutex> lock(in->parserMutex);
  if (!in->callback) in->parser->flush();
}

bool HDLSource::isRunning() const { return internal_->running.load(); }

const std::string& HDLSource::getCorrectionsFile() { return this->CorrectionsFile; }
void HDLSource::setCorrectionsFile(const std::string& filename) {
  if (filename == this->CorrectionsFile) return;
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setCorrectionsFile(filename);
  this->CorrectionsFile = filename;
}
void HDLSource::setLaserSelection(int sel[64]) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setLaserSelection(sel);
}
void HDLSource::getLaserSelection(int sel[64]) { internal_->parser->getLaserSelection(sel); }
void HDLSource::setCropReturns(int c) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setCropReturns(c);
}
void HDLSource::setCropInside(int c) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setCropInside(c);
}
void HDLSource::setCropRegion(double r[6]) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setCropRegion(r);
}
void HDLSource::setCropRegion(double xl, double xu, double yl, double yu, double zl, double zu) {
  double r[6] = {xl, xu, yl, yu, zl, zu};
  this->setCropRegion(r);
}
void HDLSource::getVerticalCorrections(double a[64]) { internal_->parser->getVerticalCorrections(a); }
unsigned int HDLSource::getDualReturnFilter() const { return internal_->parser->getDualReturnFilter(); }
void HDLSource::setDualReturnFilter(unsigned int f) { internal_->parser->setDualReturnFilter(f); }
int HDLSource::getNumberOfChannels() { return internal_->parser->getNumberOfChannels(); }

void HDLSource::setHDLManager(HDLManager* hp) { internal_->hdlMgr = hp; }
void HDLSource::setTimeSolver(std::shared_ptr<TimeSolver> solver) { internal_->timeSolver = solver; }
void HDLSource::setTransformManager(std::shared_ptr<TransformManager> mgr) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->parser->setTransformMgr(mgr);
}
std::shared_ptr<HDLParser> HDLSource::getHDLParser() { return internal_->parser; }

void HDLSource::getCounters(uint64_t* received, uint64_t* dropped, uint64_t* consumed) const {
  if (received) *received = internal_->received.load();
  if (dropped) *dropped = internal_->dropped.load();
  if (consumed) *consumed = internal_->consumed.load();
}
void HDLSource::getFrameLatencies(std::vector<double>* processUs, std::vector<double>* fromArrivalUs) const {
  std::lock_guard<std::mutex> lock(internal_->latencyMutex);
  if (processUs) *processUs = internal_->frameLatencyUs;
  if (fromArrivalUs) *fromArrivalUs = internal_->arrivalLatencyUs;
}
void HDLSource::setPacketCallback(std::function<void(const unsigned char*, unsigned int, ptime)> cb) {
  std::lock_guard<std::mutex> lock(internal_->parserMutex);
  internal_->callback = cb;
}
