#include "vtkPacketFile.h"

#include <cstring>

namespace {
const int64_t kTzShiftUs = 8ll * 3600 * 1000000;  // "we work at GMT+8" (type_defs.cxx:71)
const unsigned char kLidarHeader[42] = {
    0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x60, 0x76, 0x88, 0x00, 0x00, 0x00, 0x08, 0x00,  // Ethernet
    0x45, 0x00, 0x04, 0xd2, 0x00, 0x00, 0x40, 0x00, 0xff, 0x11, 0xb4, 0xaa,              // IPv4
    0xc0, 0xa8, 0x01, 0xc8, 0xff, 0xff, 0xff, 0xff,                                      // 192.168.1.200 -> broadcast
    0x09, 0x40, 0x09, 0x40, 0x04, 0xbe, 0x00, 0x00};                                     // UDP 2368 -> 2368
const unsigned char kPositionHeader[42] = {
    0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x60, 0x76, 0x88, 0x00, 0x00, 0x00, 0x08, 0x00,
    0x45, 0x00, 0x04, 0xd2, 0x00, 0x00, 0x40, 0x00, 0xff, 0x11, 0xb4, 0xaa,
    0xc0, 0xa8, 0x01, 0xc8, 0xff, 0xff, 0xff, 0xff,
    0x20, 0x74, 0x20, 0x74, 0x02, 0x08, 0x00, 0x00};                                     // UDP 8308 -> 8308
}  // namespace

ptime timevalToPtime(uint32_t tv_sec, uint32_t tv_usec) {
  return ptime((int64_t)tv_sec * 1000000ll + (int64_t)tv_usec + kTzShiftUs);
}

bool vtkPacketFileReader::open(const std::string& filename) {
  if (filename == fileName_ && file_) return true;
  close();
  FILE* f = std::fopen(filename.c_str(), "rb");
  if (!f) {
    lastError_ = filename + ": cannot open";
    return false;
  }
  unsigned char gh[PCAP_GLOBAL_HEADER_LEN];
  if (std::fread(gh, 1, sizeof(gh), f) != sizeof(gh) ||
      !(gh[0] == 0xd4 && gh[1] == 0xc3 && gh[2] == 0xb2 && gh[3] == 0xa1)) {
    lastError_ = filename + ": not a little-endian microsecond pcap file";
    std::fclose(f);
    return false;
  }
  file_ = f;
  fileName_ = filename;
  return true;
}

void vtkPacketFileReader::close() {
  if (file_) {
    std::fclose(file_);
    file_ = nullptr;
    fileName_.clear();
  }
}

void vtkPacketFileReader::getFilePosition(int64_t* position) {
  *position = file_ ? (int64_t)ftello(file_) : -1;
}
void vtkPacketFileReader::setFilePosition(const int64_t* position) {
  if (file_) fseeko(file_, (off_t)*position, SEEK_SET);
}

bool vtkPacketFileReader::nextPacket(const unsigned char*& data, unsigned int& dataLength, ptime& t) {
  if (!file_) return false;
  uint32_t rh[4];
  if (std::fread(rh, 4, 4, file_) != 4) {
    close();
    return false;
  }
  buf_.resize(rh[2]);
  if (rh[2] && std::fread(buf_.data(), 1, rh[2], file_) != rh[2]) {
    close();
    return false;
  }
  const unsigned int bytesToSkip = 42;
  dataLength = rh[3] >= bytesToSkip ? rh[3] - bytesToSkip : 0;
  data = buf_.data() + (rh[2] >= bytesToSkip ? bytesToSkip : rh[2]);
  t = timevalToPtime(rh[0], rh[1]);
  return true;
}

bool vtkPacketFileWriter::open(const std::string& filename) {
  close();
  FILE* f = std::fopen(filename.c_str(), "wb");
  if (!f) {
    lastError_ = filename + ": cannot create";
    return false;
  }
  const uint32_t gh[6] = {0xa1b2c3d4u, (4u << 16) | 2u, 0u, 0u, 65535u, 1u /* DLT_EN10MB */};
  std::fwrite(gh, 4, 6, f);
  file_ = f;
  fileName_ = filename;
  return true;
}

void vtkPacketFileWriter::close() {
  if (file_) {
    std::fclose(file_);
    file_ = nullptr;
    fileName_.clear();
  }
}

bool vtkPacketFileWriter::writePacket(const unsigned char* data, unsigned int dataLength, ptime t) {
  if (!file_) return false;
  const unsigned char* header;
  if (dataLength == 1206)
    header = kLidarHeader;
  else if (dataLength == 554 - 42)
    header = kPositionHeader;
  else
    return false;
  const int64_t us = t.us - kTzShiftUs;
  const uint32_t rh[4] = {(uint32_t)(us / 1000000), (uint32_t)(us % 1000000), dataLength + 42,
                          dataLength + 42};
  std::fwrite(rh, 4, 4, file_);
  std::fwrite(header, 1, 42, file_);
  std::fwrite(data, 1, dataLength, file_);
  return true;
}
