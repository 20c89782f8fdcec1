#include "TransformManager.h"

#include <algorithm>

#include <fstream>

TransformManager::TransformManager() : version_(1) { originLLH[0] = originLLH[1] = originLLH[2] = 0; }
TransformManager::~TransformManager() {}

int TransformManager::getNumberOfTransforms() {
  std::unique_lock<std::mutex> lock(mutex_);
  return (int)transforms.size();
}

void TransformManager::clearTransforms() {
  std::unique_lock<std::mutex> lock(mutex_);
  transforms.clear();
  ++version_;
}

void TransformManager::addTransform(std::shared_ptr<PoseTransform> trans) {
  std::unique_lock<std::mutex> lock(mutex_);
  transforms.addData(trans);
  ++version_;
}

namespace {
// .insmeta record: T/R/V interleaved per axis, timestamp, week, ms, week_pos, seconds_pos
// (reference type_defs.cxx:4-33 with ptime stored as int64 microseconds)
void writePose(std::ostream& os, const PoseTransform& p) {
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
bool readPose(std::ifstream& is, PoseTransform& p) {
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
}  // namespace

void TransformManager::writePoseRecord(std::ostream& os, const PoseTransform& p) { writePose(os, p); }

bool TransformManager::loadFromMetaFile(std::string filename, bool clearOldData) {
  std::ifstream ifs(filename, std::ios::binary);
  if (!ifs) return false;
  if (clearOldData) clearTransforms();
  while (true) {
    std::shared_ptr<PoseTransform> p(new PoseTransform);
    if (!readPose(ifs, *p)) break;
    addTransform(p);
  }
  return true;
}

bool TransformManager::loadFromTxtFile(std::string filename, bool clearOldData) {
  std::ifstream ifs(filename);
  if (!ifs) return false;
  if (clearOldData) clearTransforms();
  double v;
  long long sec, usec;
  std::shared_ptr<PoseTransform> trans(new PoseTransform);
  while (ifs >> trans->T[0] >> trans->T[1] >> trans->R[2] >> trans->R[0] >> trans->R[1] >> v >> sec >> usec) {
    trans->R[0] = TO_DEGREE(trans->R[0]);
    trans->R[1] = TO_DEGREE(trans->R[1]);
    trans->R[2] = -TO_DEGREE(trans->R[2]);
    // timevalToPtime: +8 hours (reference type_defs.cxx:69-72)
    trans->timestamp = ptime(sec * 1000000ll + usec + 8ll * 3600 * 1000000);
    trans->seconds_pos = (float)((trans->timestamp.us / 1000) % 604800000ll) / 1000.0f;
    if (trans->seconds_pos < 0) trans->seconds_pos = 0;
    addTransform(trans);
    trans = std::shared_ptr<PoseTransform>(new PoseTransform);
  }
  return true;
}

bool TransformManager::writeToMetaFile(const std::string& filename) {
  std::ofstream ofs(filename, std::ios::binary);
  if (!ofs) return false;
  std::unique_lock<std::mutex> lock(mutex_);
  for (const auto& p : transforms.items()) writePose(ofs, *p);
  return true;
}

bool TransformManager::interpolateTransform(ptime& t, PoseTransform* trans) {
  trans->timestamp = t;
  std::pair<std::shared_ptr<PoseTransform>, std::shared_ptr<PoseTransform> > bound;
  {
    std::unique_lock<std::mutex> lock(mutex_);
    bound = transforms.getBoundaryData(t);
  }
  if ((!bound.first) && (!bound.second)) {
    return false;
  } else if (!bound.second) {
    PoseTransform& fore = *(bound.first);
    double sec = (float)(t - fore.timestamp).total_microseconds() / 1e6f;
    for (int i = 0; i < 3; ++i) {
      trans->V[i] = fore.V[i];
      trans->R[i] = fore.R[i];
      trans->T[i] = fore.T[i] + fore.V[i] * sec;
    }
    return true;
  } else {
    PoseTransform& fore = *(bound.first);
    PoseTransform& back = *(bound.second);
    time_duration diff = t - fore.timestamp;
    double ratio = double(diff.total_microseconds()) / (back.timestamp - fore.timestamp).total_microseconds();
    *trans = fore + ((back - fore) * ratio);
    trans->seconds_pos = 0;
    return true;
  }
}

void TransformManager::setOriginLLH(const double LLH[3]) {
  originLLH[0] = TO_RADIUS(LLH[0]);
  originLLH[1] = TO_RADIUS(LLH[1]);
  originLLH[2] = LLH[2];
}

uint64_t TransformManager::version() {
  std::unique_lock<std::mutex> lock(mutex_);
  return version_;
}

void TransformManager::snapshotWindow(int64_t tmin_us, int64_t tmax_us, std::vector<int64_t>* t_us,
                                      std::vector<double>* trv) {
  std::unique_lock<std::mutex> lock(mutex_);
  const auto& items = transforms.items();
  const size_t n = items.size();
  size_t a = 0, b = n;  // [a, b)
  if (n >= 2) {
    auto lower = [&](int64_t t) {
      return (size_t)(std::lower_bound(items.begin(), items.end(), t,
                                       [](const std::shared_ptr<PoseTransform>& p, int64_t v) {
                                         return p->timestamp.us < v;
                                       }) - items.begin());
    };
    const size_t lo = std::min(std::max<size_t>(lower(tmin_us), 1), n - 1);
    const size_t hi = std::min(std::max<size_t>(lower(tmax_us), 1), n - 1);
    a = lo - 1;
    b = hi + 1;
  }
  t_us->resize(b - a);
  trv->resize((b - a) * 9);
  for (size_t i = a; i < b; ++i) {
    (*t_us)[i - a] = items[i]->timestamp.us;
    for (int k = 0; k < 3; ++k) {
      (*trv)[9 * (i - a) + k] = items[i]->T[k];
      (*trv)[9 * (i - a) + 3 + k] = items[i]->R[k];
      (*trv)[9 * (i - a) + 6 + k] = items[i]->V[k];
    }
  }
}

void TransformManager::snapshot(std::vector<int64_t>* t_us, std::vector<double>* trv) {
  std::unique_lock<std::mutex> lock(mutex_);
  const auto& items = transforms.items();
  t_us->resize(items.size());
  trv->resize(items.size() * 9);
  for (size_t i = 0; i < items.size(); ++i) {
    (*t_us)[i] = items[i]->timestamp.us;
    for (int k = 0; k < 3; ++k) {
      (*trv)[9 * i + k] = items[i]->T[k];
      (*trv)[9 * i + 3 + k] = items[i]->R[k];
      (*trv)[9 * i + 6 + k] = items[i]->V[k];
    }
  }
}
