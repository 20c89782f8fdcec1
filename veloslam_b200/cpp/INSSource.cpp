#include "INSSource.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>

#include "CoordiTran.h"

class INSSource::vsInternal {
 public:
  explicit vsInternal(int p)
      : port(p), sock(-1), running(false), poses(0), timeSolver(new TimeSolver) {
    // the reference's default origin (INSSource.cxx:334)
    origin[0] = -2781621.9891904;
    origin[1] = 4672106.75052387;
    origin[2] = 18.8910392;
  }
  int port, sock;
  std::atomic<bool> running;
  std::atomic<uint64_t> poses;
  std::shared_ptr<TimeSolver> timeSolver;
  std::shared_ptr<TransformManager> transformMgr;
  double origin[3];
  std::string outputFile;
  std::ofstream writer;
  std::mutex mutex;  // origin / managers / writer vs. the receive thread
  std::thread thread;
};

INSSource::INSSource(int port) : internal_(new vsInternal(port)) {}
INSSource::~INSSource() {
  this->stop();
  delete internal_;
}

std::shared_ptr<PoseTransform> INSSource::calcTransform(InsPVA const* data) {
  vsInternal* in = internal_;
  double input[3] = {TO_RADIUS(data->LLH[0]), TO_RADIUS(data->LLH[1]), data->LLH[2]};
  double enu[3] = {0, 0, 0};
  llh2enu(input, in->origin, enu);
  std::shared_ptr<PoseTransform> trans(new PoseTransform);
  for (int k = 0; k < 3; ++k) {
    trans->T[k] = enu[k];
    trans->R[k] = data->Eulr[k];
    trans->V[k] = data->V[k];
  }
  trans->week_number = data->week_number;
  trans->milliseconds = data->milliseconds;
  trans->week_number_pos = data->week_number_pos;
  trans->seconds_pos = data->seconds_pos;
  trans->timestamp = in->timeSolver->calcTimestamp(data);
  return trans;
}

void INSSource::start() {
  vsInternal* in = internal_;
  if (in->running.load()) return;
  in->sock = socket(AF_INET, SOCK_DGRAM, 0);
  if (in->sock < 0) {
    std::cerr << "INSSource: cannot create the UDP socket" << std::endl;
    return;
  }
  int one = 1;
  setsockopt(in->sock, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  timeval tv = {0, 100000};
  setsockopt(in->sock, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  sockaddr_in addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = htons((uint16_t)in->port);
  if (bind(in->sock, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    std::cerr << "INSSource: cannot bind UDP port " << in->port << std::endl;
    close(in->sock);
    in->sock = -1;
    return;
  }
  if (!in->outputFile.empty()) {
    in->writer.open(in->outputFile, std::ios::binary);
    if (!in->writer) std::cerr << "Failed to write to INS file: " << in->outputFile << std::endl;
  }
  in->running.store(true);
  in->thread = std::thread([this, in]() {
    unsigned char buf[1500];
    while (in->running.load()) {
      const ssize_t n = recv(in->sock, buf, sizeof(buf), 0);
      if (n < (ssize_t)sizeof(uint16_t)) continue;
      uint16_t id;
      std::memcpy(&id, buf, 2);
      if (id != INSPVA || n < (ssize_t)sizeof(InsPVA)) continue;  // RAWINS, BESTGPSPOS: ignored
      InsPVA rec;
      std::memcpy(&rec, buf, sizeof(rec));
      std::lock_guard<std::mutex> lock(in->mutex);
      std::shared_ptr<PoseTransform> trans = this->calcTransform(&rec);
      if (in->writer.is_open()) TransformManager::writePoseRecord(in->writer, *trans);
      if (in->transformMgr) in->transformMgr->addTransform(trans);
      ++in->poses;
    }
  });
}

void INSSource::stop() {
  vsInternal* in = internal_;
  if (!in->running.exchange(false)) return;
  if (in->thread.joinable()) in->thread.join();
  if (in->sock >= 0) close(in->sock);
  in->sock = -1;
  if (in->writer.is_open()) in->writer.close();
}

bool INSSource::isRunning() const { return internal_->running.load(); }
uint64_t INSSource::posesReceived() const { return internal_->poses.load(); }

void INSSource::setOutputFile(const std::string& filename) {
  std::lock_guard<std::mutex> lock(internal_->mutex);
  internal_->outputFile = filename;
}
void INSSource::setTransformManager(std::shared_ptr<TransformManager> mgr) {
  std::lock_guard<std::mutex> lock(internal_->mutex);
  internal_->transformMgr = mgr;
}
void INSSource::setTimeSolver(std::shared_ptr<TimeSolver> solver) {
  std::lock_guard<std::mutex> lock(internal_->mutex);
  internal_->timeSolver = solver;
}
void INSSource::setOrigin(double org[3]) {
  std::lock_guard<std::mutex> lock(internal_->mutex);
  for (int k = 0; k < 3; ++k) internal_->origin[k] = org[k];
}
