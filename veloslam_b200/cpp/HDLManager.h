// HDLManager.h -- frame store of the drop-in facade with the reference's HDLManager interface
// (/root/reference/HDLManager.h:94-249).
//
// Keeps the reference's behaviour for: the TimeLine<HDLFrame> index and its queries
// (getFrameAt / getFrameNear / getRangeBetween / getRecentFrame / getAllFrameMeta), the
// in-memory frame cache (pushCache / updateCacheSize / cleanCache, HDLManager.cxx:400-427), the
// two hard-drive buffers that are written out as pcap files with the frames' file locators
// (addFrame / switchBuffer / writePackets, :170-193, :318-371), lazy re-decode of frames that
// are only on disk (prepareFrame, :195-211) and the .hdlmeta / .insmeta persistence
// (:429-467, HDLFrame.cxx:160-190).
//
// B200-native differences:
//  * loadOffline() uploads the packet file to HBM once (HDLParser::loadRecording): the frame
//    index is one segmentation pass on the GPU and every later prepareFrame() decodes its
//    rotation out of HBM -- the reference re-opens and re-reads the pcap file per frame.
//  * setDevices({0, 1, ...}) before loadOffline(): the recording is split into one range of
//    records per GPU (vs_shard_range, the split bench.py's recording legs use), each GPU keeps
//    and indexes its own range on its own thread, the per-GPU indices are concatenated into
//    the one TimeLine, prepareFrame() decodes a rotation on the GPU that holds it and
//    getRangeBetween() decodes on all of them at once.
//  * writePackets() runs synchronously when a buffer fills (the reference runs it on a
//    boost::thread).
//  * the sources are created by startOnline() (ports: setPorts) instead of in the constructor,
//    so that offline users open no sockets.
//  * ptime is int64 microseconds and fpos_t an int64 byte offset (type_defs.h), so meta
//    files are not byte-compatible with ones the reference wrote.
#ifndef VELOSLAM_B200_HDLMANAGER_H
#define VELOSLAM_B200_HDLMANAGER_H

#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "HDLFrame.h"
#include "HDLParser.h"
#include "HDLSource.h"
#include "INSSource.h"
#include "TimeLine.h"
#include "TransformManager.h"
#include "vtkPacketFile.h"

// boost::intrusive_ptr<HDLFrame> of the reference: counts end users in HDLFrame::count so the
// cache never clears a frame somebody still holds (HDLFrame.cxx:211-219)
class HDLFramePtr {
 public:
  HDLFramePtr() : p_(nullptr) {}
  explicit HDLFramePtr(HDLFrame* p) : p_(p) { if (p_) intrusive_ptr_add_ref(p_); }
  HDLFramePtr(const HDLFramePtr& o) : p_(o.p_) { if (p_) intrusive_ptr_add_ref(p_); }
  HDLFramePtr& operator=(const HDLFramePtr& o) {
    if (o.p_) intrusive_ptr_add_ref(o.p_);
    if (p_) intrusive_ptr_release(p_);
    p_ = o.p_;
    return *this;
  }
  ~HDLFramePtr() { if (p_) intrusive_ptr_release(p_); }
  HDLFrame* get() const { return p_; }
  HDLFrame* operator->() const { return p_; }
  HDLFrame& operator*() const { return *p_; }
  explicit operator bool() const { return p_ != nullptr; }

 private:
  HDLFrame* p_;
};

class HDLManager {
 public:
  typedef std::vector<std::shared_ptr<HDLFrame> > Buffer;

  HDLManager(int capacity = 200);  // 600 hdl frames ~= 1 minute
  virtual ~HDLManager();

  // online: HDLSource (sensor UDP port) -> this manager, INSSource (INS UDP port) -> the pose
  // timeline, one TimeSolver for both (reference HDLManager.cxx:43-100)
  void startOnline(bool shouldSwap = false);
  void stopOnline();
  void setPorts(int hdlPort, int insPort);   // before startOnline; defaults 2368 / 6777
  std::shared_ptr<HDLSource> getHDLSource() const { return hdlSrc; }
  std::shared_ptr<INSSource> getINSSource() const { return insSrc; }
  std::shared_ptr<TimeSolver> getTimeSolver() const { return timeSolver; }

  void loadOffline(const std::string& insTxt, const std::string& pcapfile);
  // CUDA devices loadOffline() spreads a recording over (default: the parser's one device).
  // The same device may be named more than once (one context each).
  void setDevices(const std::vector<int>& cudaDevices);
  int getNumberOfShards() const { return (int)shards_.size(); }
  /* the purpose of 'touch' is rename the file if necessary */
  void touchPcap(const std::string& pcapfile);

  bool loadHDLMeta();
  bool loadINSMeta();
  bool saveHDLMeta();
  bool saveINSMeta();

  int getNumberOfFrames();
  int getNumberOfTransforms();

  void addFrame(std::shared_ptr<HDLFrame> frame);
  HDLFramePtr prepareFrame(std::shared_ptr<HDLFrame> frame);
  /* blocks up to `micro` microseconds for a frame added since the last call */
  HDLFramePtr waitForFrame(int64_t micro = 100000);
  HDLFramePtr getRecentFrame();
  HDLFramePtr getFrameAt(ptime& t);
  HDLFramePtr getFrameNear(ptime& t);
  /* meta only: do not touch the points of these */
  std::vector<std::shared_ptr<HDLFrame> > getAllFrameMeta();
  std::vector<HDLFramePtr> getRangeBetween(ptime& a, ptime& b);

  void setBufferSize(size_t n);
  size_t getBufferSize();
  bool setBufferDir(std::string dirname, bool shouldCreateSubDir = true);
  void resetBufferDir();
  void setFileBufferMode(bool m = true);
  void flushFileBuffer();
  bool writePackets();
  void startSwaping();
  void stopSwaping();

  void pushCache(std::shared_ptr<HDLFrame>& frame);
  void updateCacheSize();
  void cleanCache();
  void switchBuffer();

  void setCalibFile(std::string filename);

  // the parser / pose timeline this manager coordinates (the reference keeps them private)
  std::shared_ptr<HDLParser> getParser() const { return hdlParser; }
  std::shared_ptr<TransformManager> getTransformMgr() const { return transMgr; }
  const std::string& getBufferDir() const { return bufferDirName; }

 protected:
  TimeLine<HDLFrame> frames;
  std::deque<HDLFrame*> cache;
  Buffer hardDriveBuffer1, hardDriveBuffer2;
  Buffer* hardDriveBuffer;
  int cacheCounter;

 private:
  HDLManager(const HDLManager&);
  void operator=(const HDLManager&);
  void scanBufferDir();

  void updateCacheSizeLocked();  // caller holds cacheMutex
  // one range of the recording per GPU: frames that start in records [first, end) are indexed
  // and decoded by `parser`, which holds [first - 1, end + kShardTail) so that the wrap test
  // of its first record and the end of its last rotation need nothing from a neighbour
  struct Shard {
    int64_t first, end;
    std::shared_ptr<HDLParser> parser;
  };
  static const int64_t kShardTail = 2048;   // records; > one rotation of either sensor at 5 Hz
  bool loadOfflineSharded(const std::string& pcapfile, std::string* resolved);
  HDLParser* parserFor(const HDLFrame& frame, const std::string& pcap);
  std::vector<int> devices_;
  std::vector<Shard> shards_;
  std::string shardFile_;
  std::mutex framesMutex;
  std::mutex cacheMutex;         // cache + cacheCounter (consumer thread and user threads)
  std::condition_variable cond_;
  bool hasNewData;
  size_t bufferSize;
  size_t maxCacheSize;
  std::string bufferDirName;
  std::set<std::string> bufferFileNames, hdlMetaNames, insMetaNames;
  bool isUsingBuffer1;
  bool writerIdle;
  bool fileBufferMode;
  vtkPacketFileWriter* packetWriter;
  std::mutex writerMutex;
  std::shared_ptr<TransformManager> transMgr;
  std::shared_ptr<HDLParser> hdlParser;
  std::shared_ptr<HDLSource> hdlSrc;
  std::shared_ptr<INSSource> insSrc;
  std::shared_ptr<TimeSolver> timeSolver;
  int hdlPort_, insPort_;
  std::string calibFile_;
  int metaSerial;  // keeps meta file names of one process distinct within a microsecond
};

#endif
