/*
 * velo_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free CPU restatement of the
 * reference's ingest hot path (victl/VeloSLAM: HDLParser.cxx, TransformManager.cxx,
 * TimeLine.h, type_defs.h).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (libveloslam_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pin: the reference ships no golden vectors or known-answer tests for this path
 * (SURVEY.md section 4).  The restatement is pinned three ways:
 *   1. oracle/_ref -- the reference's own HDLParser.cxx / TransformManager.cxx / TimeLine.h /
 *      type_defs.* / HDLFrame.cxx / vtkPacketFileWriter.cxx compiled verbatim, where they lie
 *      under /root/reference, against stand-in third-party headers (oracle/ref_shim, recipe
 *      oracle/build_ref.py) and compared output-for-output, bit-exact
 *      (tests/test_oracle_vs_ref.py);
 *   2. golden fixtures generated from oracle/_ref and committed under tests/golden/
 *      (tests/golden/make_golden.py, tests/test_golden.py);
 *   3. hand-derived known-answer tests (tests/test_oracle_kat.py).
 * What (1) cannot pin is the third-party arithmetic itself (Eigen's AngleAxis/3x3 product
 * rounding order, Boost's time arithmetic): those are restated in oracle/ref_shim.
 *
 * Time: boost::posix_time::ptime (microsecond resolution) is restated as
 * int64 microseconds since the Unix epoch.
 */
#ifndef VELO_ORACLE_H
#define VELO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_parser vo_parser;

typedef struct vo_frame_info {
  int64_t timestamp_us;     /* HDLFrame::timestamp; INT64_MIN when never initialised */
  int32_t skips;            /* HDLFrame::skips; -1 when never initialised            */
  int32_t n_lasers;         /* points.size()                                         */
  int32_t n_points;         /* sum over lasers                                       */
  int32_t n_packets;        /* packets.size() (first packet of a frame is doubled)   */
  int32_t is_hdl64_order;   /* 1 when splitFrame applied HDL64BeamLUT                */
  int32_t pad;
  double  carpose_TRV[9];   /* carpose T[3], R[3], V[3]                              */
  double  carpose_seconds_pos; /* -1 == invalid pose (type_defs.cxx:56)              */
} vo_frame_info;

vo_parser* vo_create(void);
void       vo_destroy(vo_parser*);

/* Calibration as the five raw db.xml values per laser (HDLParser.cxx:822-839) plus the
 * number of enabled_ items equal to 1 (HDLParser.cxx:785-799). */
void vo_set_calibration(vo_parser*, const double* rot_deg, const double* vert_deg,
                        const double* dist_cm, const double* voff_cm, const double* hoff_cm,
                        int n_rows, int n_enabled);
void vo_set_laser_selection(vo_parser*, const int32_t sel[64]);
void vo_set_points_skip(vo_parser*, int32_t points_skip);
void vo_set_crop(vo_parser*, int32_t crop_returns, int32_t crop_inside, const double region[6]);

/* TransformManager */
void    vo_clear_poses(vo_parser*);
void    vo_add_pose(vo_parser*, int64_t t_us, const double T[3], const double R[3], const double V[3]);
int32_t vo_num_poses(vo_parser*);
/* interpolateTransform: returns its bool; out_TRV = T,R,V; *seconds_pos = -1 or 0 */
int32_t vo_interpolate(vo_parser*, int64_t t_us, double out_TRV[9], double* seconds_pos);
/* PoseTransform::getMatrix of a TRV triple: out = 3x4 row-major [L | t] */
void    vo_pose_matrix(const double TRV[9], double out[12]);

/* Parser state */
void vo_unload(vo_parser*);                       /* HDLParser::unloadData            */
void vo_set_firing_skip(vo_parser*, int32_t s);   /* getFrame sets firingSkip = skip  */
void vo_get_state(vo_parser*, int32_t out[4]);    /* lastAzimuth, firingSkip, frameMetaInited, isHDL64Data */

/* Streaming decode */
void vo_process_packet(vo_parser*, const uint8_t* data, uint32_t len, int64_t t_us);
void vo_consume_packets(vo_parser*, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us, int64_t* out_frames_points);
void vo_process_packets(vo_parser*, const uint8_t* data, int64_t n, int64_t stride,
                        const int64_t* t_us);
void vo_split_frame(vo_parser*);                  /* splitFrame(), used by getFrame's tail */

/* Completed frames (HDLParser::getAllFrames order) */
int32_t vo_num_frames(vo_parser*);
void    vo_clear_frames(vo_parser*);
int32_t vo_frame_get_info(vo_parser*, int32_t f, vo_frame_info* out);
int32_t vo_frame_laser_counts(vo_parser*, int32_t f, int32_t* counts);
/* laser-major concatenation in the frame's final laser order */
int32_t vo_frame_points(vo_parser*, int32_t f, float* xyzi, uint16_t* azimuth, float* distance);
/* number of points currently sitting in the open (not yet split) frame */
int64_t vo_open_frame_points(vo_parser*);

/* Emission trace: one record per pushed point, in emission (stream) order */
void    vo_trace_enable(vo_parser*, int32_t on);
int64_t vo_trace_size(vo_parser*);
void    vo_trace_fetch(vo_parser*, int32_t* packet, uint8_t* block, uint8_t* dsr, uint8_t* laser,
                       int32_t* frame, float* x, float* y, float* z, uint8_t* intensity,
                       uint16_t* azimuth, uint16_t* raw_distance, uint32_t* tadj_us);

/* Offline index (HDLParser::readFrameInformation) over an in-memory packet array.
 * Returns the number of frames; fills up to cap entries. */
int32_t vo_read_frame_information(const uint8_t* data, int64_t n, int64_t stride,
                                  const int64_t* t_us, int32_t* start_packet,
                                  int32_t* skips, int64_t* timestamp_us, int32_t cap);
/* HDLParser::getFrame over an in-memory packet array: decode from (start_packet, skip)
 * until the first split (or force a split at the end).  The frame is appended to the
 * parser's frame list.  Returns 1 on success. */
int32_t vo_get_frame(vo_parser*, const uint8_t* data, int64_t n, int64_t stride,
                     const int64_t* t_us, int64_t start_packet, int32_t skip);

/* ---- online ingest front end --------------------------------------- */

/* CoordiTran.cpp:51-80 (llh2xyz), :82-150 (xyz2llh), :152-187 + :271-276 (xyz2enu, llh2enu):
 * WGS-84 geodetic (radians, metres) -> ECEF -> local ENU about an ECEF origin. */
void vo_llh2xyz(const double llh[3], double xyz[3]);
void vo_xyz2llh(const double xyz[3], double llh[3]);
void vo_llh2enu(const double llh[3], const double orgxyz[3], double enu[3]);

/* TimeSolver::calcTimestamp(uint32_t microsecToHour) (TimeSolver.cxx:34-49).  The state the
 * reference keeps in hdlHourTime + hdlOffset is one number, base_us; now_us is the local clock
 * it reads at the very first packet. */
typedef struct vo_time_solver {
  int64_t  base_us;
  uint32_t last_report;
  int32_t  inited;
} vo_time_solver;
void    vo_ts_init(vo_time_solver*);
int64_t vo_ts_hdl(vo_time_solver*, uint32_t microsec_to_hour, int64_t now_us);

/* NovAtel INSPVA record as the reference lays it out (type_defs.h:39-58, natural alignment). */
typedef struct vo_ins_pva {
  uint16_t message_id;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  uint32_t pad0;
  double   seconds_pos;
  double   LLH[3];
  double   V[3];
  double   Eulr[3];
  int32_t  ins_status;
  int32_t  pad1;
} vo_ins_pva;
/* TimeSolver::calcTimestamp(InsPVA const*) (TimeSolver.cxx:20-33).  insInited is never set, so
 * the offset is re-taken from the clock on every call and the result is
 * now + (time of pose - time of packet send); the epoch cancels. */
int64_t vo_ts_ins(const vo_ins_pva* rec, int64_t now_us);
/* PacketConsumer::calcTransform (INSSource.cxx:300-326) without the timestamp: T = ENU of the
 * position, R = Eulr, V = V. */
void    vo_ins_pose(const vo_ins_pva* rec, const double orgxyz[3], double trv[9]);

#ifdef __cplusplus
}
#endif
#endif
