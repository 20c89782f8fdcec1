// vtkPacketFile.h -- dependency-free pcap reader / writer with the reference's class names
// (vtkPacketFileReader.h:68-209, vtkPacketFileWriter.{h,cxx}).  Classic little-endian
// microsecond pcap: 24-byte global header, then 16 + 42 + 1206 = 1264-byte records; the reader
// strips the 42-byte Ethernet/IP/UDP header and converts the record timeval with the
// reference's timevalToPtime rule (+8 hours, type_defs.cxx:69-72).
#ifndef VELOSLAM_B200_VTKPACKETFILE_H
#define VELOSLAM_B200_VTKPACKETFILE_H

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "type_defs.h"

#define PCAP_GLOBAL_HEADER_LEN 24
#define PCAP_PACKET_LEN 1264

class vtkPacketFileReader {
 public:
  vtkPacketFileReader() : file_(nullptr) {}
  ~vtkPacketFileReader() { close(); }
  bool open(const std::string& filename);
  bool isOpen() const { return file_ != nullptr; }
  void close();
  const std::string& getLastError() const { return lastError_; }
  const std::string& getFileName() const { return fileName_; }
  void getFilePosition(int64_t* position);
  void setFilePosition(const int64_t* position);
  // payload after the 42-byte network header, its length and the packet time
  bool nextPacket(const unsigned char*& data, unsigned int& dataLength, ptime& t);

 private:
  FILE* file_;
  std::string fileName_, lastError_;
  std::vector<unsigned char> buf_;
};

class vtkPacketFileWriter {
 public:
  vtkPacketFileWriter() : file_(nullptr) {}
  ~vtkPacketFileWriter() { close(); }
  bool open(const std::string& filename);
  bool isOpen() const { return file_ != nullptr; }
  void close();
  const std::string& GetLastError() const { return lastError_; }
  const std::string& GetFileName() const { return fileName_; }
  // 1206-byte lidar payloads (and 512-byte position payloads) behind a fabricated header;
  // t is the time the READER will report (the +8 h rule is undone here)
  bool writePacket(const unsigned char* data, unsigned int dataLength, ptime t);

 private:
  FILE* file_;
  std::string fileName_, lastError_;
};

// ptime <-> pcap record timestamp (reference type_defs.cxx:69-72)
ptime timevalToPtime(uint32_t tv_sec, uint32_t tv_usec);

#endif
