// TimeSolver.h -- sensor clock -> absolute time, with the reference's interface
// (/root/reference/TimeSolver.h, TimeSolver.cxx:20-49) for callers that stamp single packets /
// INS records as they arrive; packet arrays go through vs_solve_packet_times (one 1-bit scan
// on the GPU) and INS logs through vs_poses_from_ins.
//
// The reference reads boost's microsec_clock::local_time(); here the clock is a member that
// tests can replace (setClock).
#ifndef VELOSLAM_B200_TIMESOLVER_H
#define VELOSLAM_B200_TIMESOLVER_H

#include <cstdint>
#include <functional>

#include "type_defs.h"

// NovAtel INSPVA record (reference type_defs.h:39-58)
struct InsPVA {
  uint16_t message_id;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  double seconds_pos;
  double LLH[3];
  double V[3];
  double Eulr[3];
  int32_t ins_status;
};

class TimeSolver {
 public:
  TimeSolver();
  ~TimeSolver() {}
  ptime calcTimestamp(InsPVA const* data);      // INS version
  ptime calcTimestamp(uint32_t microsecToHour); // HDL version
  void setClock(std::function<int64_t()> nowUs) { now_ = nowUs; }

 private:
  std::function<int64_t()> now_;
  int64_t hdlBaseUs_;  // hdlHourTime + hdlOffset: only their sum is ever used
  bool hdlInited_;
  uint32_t lastHdlReport_;
};

#endif
