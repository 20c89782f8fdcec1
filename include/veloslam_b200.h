/*
 * veloslam_b200.h -- C ABI of the B200-native Velodyne ingest hot path.
 *
 * Drop-in boundary for victl/VeloSLAM's packet decode -> calibrated points -> rotation
 * segmentation -> per-packet motion compensation.  The reference has no FFI layer: its
 * boundary is the public C++ API of HDLParser (HDLParser.h:84-147) and TransformManager
 * (TransformManager.h:80-123).  Each entry point below names the reference interface it
 * replaces; the C++ facade in veloslam_b200/cpp/ keeps the reference's class and method
 * names on top of this ABI.
 *
 * Plain C: pointers and sizes only, no exceptions, no C++/torch types.  Every function
 * returns a vs_status (0 == ok); vs_last_error() gives the message of the last failure on
 * a context.  There is no CPU fallback: without a CUDA device vs_create() fails.
 *
 * Threading (same contract as the reference's parserMutex, HDLSource.cxx:214): calls on
 * one context must be externally serialised; different contexts are independent.
 *
 * Time is int64 microseconds since the Unix epoch (boost::posix_time::ptime restated).
 */
#ifndef VELOSLAM_B200_H
#define VELOSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VS_API __attribute__((visibility("default")))
#else
#define VS_API
#endif

#define VS_PACKET_BYTES 1206      /* HDLParser.cxx:982 */
#define VS_BLOCKS_PER_PACKET 12   /* type_defs.h:19 HDL_FIRING_PER_PKT */
#define VS_RETURNS_PER_BLOCK 32   /* type_defs.h:17 HDL_LASER_PER_FIRING */
#define VS_MAX_LASERS 64          /* type_defs.h:18 HDL_MAX_NUM_LASERS */
#define VS_PCAP_RECORD_BYTES 1264 /* vtkPacketFileReader.h:57-66: 16 + 42 + 1206 */
#define VS_PCAP_PAYLOAD_OFFSET 82 /* 24-byte global header + 16 + 42 */
#define VS_TIME_NONE INT64_MIN    /* ptime not_a_date_time */

typedef enum vs_status {
  VS_OK = 0,
  VS_ERR_INVALID_ARG = 1,
  VS_ERR_NOT_CALIBRATED = 2, /* decode before vs_set_calibration (HDLParser.cxx:512-516) */
  VS_ERR_CUDA = 3,
  VS_ERR_CAPACITY = 4,       /* batch/pose/frame capacity exceeded */
  VS_ERR_NO_DEVICE = 5,
  VS_ERR_HALO = 6,           /* halo too short to resolve the carried state */
  VS_ERR_STATE = 7           /* bad ticket / slot busy */
} vs_status;

typedef enum vs_mode {
  /* HDLParser::processHDLPacket as called per packet by the online consumer
   * (HDLSource.cxx:209-225), bug-compatible (SURVEY.md F4 a-d). */
  VS_MODE_STREAMING = 0,
  /* HDLParser::readFrameInformation + getFrame (HDLParser.cxx:1065-1160, 505-544): every
   * block decoded, frame origin = pose at the packet that holds the frame's first block. */
  VS_MODE_OFFLINE = 1
} vs_mode;

/* Flags for vs_submit*(). */
#define VS_FLAG_DEVICE_INPUT 1u   /* packets / times are device pointers (already in HBM) */
#define VS_FLAG_PCAP_TIMES 2u     /* take packet times from the pcap record headers
                                     (ts_sec + 8 h, ts_usec: type_defs.cxx:69-72); pkts must
                                     then point at the first record's payload, stride 1264 */

/* Per-point deskew extension (NOT a reference behaviour -- the reference
 * applies one Euler-interpolated pose per packet and re-bases only the translation).  Every
 * point is transformed with the pose at its own time t_packet + firing offset: quaternion
 * slerp of the rotation and lerp of the translation inside the packet's pose bracket, and the
 * result is expressed in the frame origin's coordinates (rotation and translation re-based).
 * Needs >= 2 poses (else ignored).  Exact semantics: oracle/deskew_port.py. */
#define VS_FLAG_DESKEW_PER_POINT 4u
/* The caller takes the frame table from vs_frame_table_rows_device (multi-GPU exchange out of
 * HBM): vs_wait does not assemble the vs_frame list (vs_result.frames == NULL, n_frames and
 * carry_out are exact), which for the thousands of frames of a long recording is most of its
 * host time.  vs_layout_frames is not available for such a batch. */
#define VS_FLAG_NO_FRAME_LIST 8u

/* One <px> item of the calibration db.xml in its raw units (HDLParser.cxx:818-832). */
typedef struct vs_laser_corr {
  double rot_correction_deg;          /* rotCorrection_          */
  double vert_correction_deg;         /* vertCorrection_         */
  double dist_correction_cm;          /* distCorrection_         */
  double vert_offset_correction_cm;   /* vertOffsetCorrection_   */
  double horiz_offset_correction_cm;  /* horizOffsetCorrection_  */
} vs_laser_corr;

/* HDLParser::setLaserSelection / setPointsSkip / setCropReturns / setCropInside /
 * setCropRegion (HDLParser.cxx:367-455). */
typedef struct vs_filters {
  uint64_t laser_mask;     /* bit i == laserSelections[i] */
  int32_t  points_skip;
  int32_t  crop_returns;
  int32_t  crop_inside;
  int32_t  reserved;
  double   crop_region[6]; /* xl, xu, yl, yu, zl, zu */
} vs_filters;

/* The parser state that survives a packet (HDLParser.cxx:196-216), passed by value between
 * batches; vs_carry_init() gives the state after HDLParser::unloadData(). */
typedef struct vs_carry {
  int32_t last_azimuth;        /* lastAzimuth, -1 == none yet             */
  int32_t firing_skip;         /* firingSkip                              */
  int32_t frame_meta_inited;   /* frameMetaInited of the open frame       */
  int32_t is_hdl64;            /* sticky isHDL64Data                      */
  double  origin_T[3];         /* currentFrame->carpose->T (frame origin) */
  int64_t frame_timestamp_us;  /* open frame: HDLFrame::timestamp         */
  int32_t frame_skips;         /* open frame: HDLFrame::skips             */
  int32_t frame_carpose_valid; /* open frame: carpose->seconds_pos != -1  */
  double  frame_carpose[9];    /* open frame: carpose T, R, V             */
  int64_t frames_closed;       /* running totals since vs_carry_init      */
  int64_t points_emitted;
  int64_t packets_seen;
} vs_carry;

/* One frame (HDLFrame.h:13-47) as an index into the batch's point columns.  Points of
 * frame f are the contiguous range [first_point, first_point + n_points) in emission
 * order (packet, block, laser-in-block) -- the order in which the reference pushes them;
 * laser_counts gives the size of each HDLFrame::points[laser] list. */
typedef struct vs_frame {
  int64_t  first_point;
  int64_t  n_points;
  int64_t  timestamp_us;     /* VS_TIME_NONE when the reference never initialises it */
  int32_t  start_packet;     /* packet (index in the submitted array) and block where   */
  int32_t  start_block;      /* the frame begins; start_packet -1: an earlier batch      */
  int32_t  meta_packet;      /* packet that set timestamp/carpose/skips; -1 carry, -2 none */
  int32_t  skips;            /* HDLFrame::skips; -1 when never initialised              */
  int32_t  closed;           /* 1: closed by a wrap inside this batch                   */
  int32_t  hdl64_order;      /* 1: reference re-orders lasers by HDL64BeamLUT at close  */
  int32_t  carpose_valid;    /* carpose->seconds_pos != -1                              */
  int32_t  reserved;
  double   carpose[9];       /* T, R (deg), V                                           */
  uint32_t laser_counts[VS_MAX_LASERS]; /* by laser id as pushed (before HDL64BeamLUT)  */
} vs_frame;

/* Result of one batch.  Column pointers are DEVICE pointers owned by the context and stay
 * valid until the slot's next submit.  frames is a host array owned by the context. */
typedef struct vs_result {
  int64_t n_packets;          /* packets decoded (halo excluded)                   */
  int64_t n_points;           /* emitted points                                    */
  int32_t n_frames;           /* closed frames + 1 (the last entry is the open one) */
  int32_t n_closed;
  const float*    x;          /* metres, after the per-packet rigid transform      */
  const float*    y;
  const float*    z;
  const uint8_t*  intensity;
  const uint8_t*  laser;      /* HDLFrame::points row the reference pushes into    */
  const uint16_t* azimuth;    /* PointMeta::azimuth (adjusted, mod 36000)          */
  const uint16_t* distance;   /* raw 2 mm units                                    */
  const uint32_t* t_us;       /* (packet time - t_base) + rounded firing offset    */
  const vs_frame* frames;
  vs_carry carry_out;
  int64_t t_base_us;
  int64_t first_upper_block;  /* first decoded 0xddff block (packet*12+block), -1 none */
  float   gpu_ms;             /* device time of the batch's kernels (CUDA events); a batch  */
                              /* issued as a CUDA graph: of the whole graph, copies included */
  float   decode_ms;          /* device time of the decode kernel alone (0 for a graph)     */
  int32_t n_kernel_launches;
  int32_t reserved;           /* diagnostics: batches this slot has issued as one CUDA graph */
} vs_result;

typedef struct vs_ctx vs_ctx;

/* Library / build info: "veloslam_b200 <version> sm_100a". */
VS_API const char* vs_version(void);

/* Replaces `new HDLParser` + `new TransformManager` (HDLParser.cxx:284-288).
 * max_batch_packets bounds one submit (halo included); n_slots in {1,2} result slots. */
VS_API int vs_create(int device, int64_t max_batch_packets, int64_t max_poses, int n_slots, vs_ctx** out);
VS_API void vs_destroy(vs_ctx* ctx);
VS_API const char* vs_last_error(vs_ctx* ctx);

/* HDLParser::setCorrectionsFile -> loadCorrectionsFile (HDLParser.cxx:458-475, 771-858):
 * n_rows calibration rows (id_ == index) and the count of enabled_ items equal to 1. */
VS_API int vs_set_calibration(vs_ctx* ctx, const vs_laser_corr* corr, int n_rows, int n_lasers_enabled);
VS_API int vs_set_filters(vs_ctx* ctx, const vs_filters* f);
/* Firing time of each (block, return) slot inside a packet, microseconds after the packet
 * time, for sensors whose table the reference does not have (HDL-64E: the reference gives all
 * 384 returns the packet time): 12 x 32 values, row-major.  HDL-32E / VLP-16 use
 * the reference's own table (HDLParser.cxx:133-137) and reject this call.  Used by
 * VS_FLAG_DESKEW_PER_POINT (pose per point; the t_us column then includes the offset).  Call
 * after vs_set_calibration, which resets the table. */
VS_API int vs_set_firing_offsets(vs_ctx* ctx, const uint16_t* off_us);

/* TransformManager::addTransform / clearTransforms as an immutable snapshot per batch
 * (TransformManager.cxx:61-79): n poses sorted by strictly increasing time, trv = n x 9
 * doubles (T[3], R[3] degrees, V[3]).  n == 0 clears. */
VS_API int vs_set_poses(vs_ctx* ctx, const int64_t* t_us, const double* trv, int64_t n);
/* TransformManager::interpolateTransform (TransformManager.cxx:149-177) on the host copy of
 * the snapshot.  Returns VS_OK; *found = its bool, *valid = (seconds_pos != -1). */
VS_API int vs_interpolate(vs_ctx* ctx, int64_t t_us, double out_trv[9], int32_t* found, int32_t* valid);

VS_API void vs_carry_init(vs_carry* c);

/* HDLParser::processHDLPacket for a batch of n packets (HDLParser.cxx:489, 980-1055).
 * pkts: n records of VS_PACKET_BYTES payload at `stride` bytes (1206 dense, 1264 = pcap
 * records); pkt_time_us: n packet times (ignored with VS_FLAG_PCAP_TIMES).  The first
 * n_halo packets only rebuild parser state (multi-GPU shards) and emit nothing; with
 * n_halo > 0 the carry-in is ignored and the state is resolved from the halo.
 * t_base_us: origin of the t_us column (VS_TIME_NONE: time of the first decoded packet,
 * host input only).  Host buffers are copied asynchronously; they must stay valid until
 * vs_wait returns. */
VS_API int vs_submit(vs_ctx* ctx, const uint8_t* pkts, int64_t stride, const int64_t* pkt_time_us,
              int64_t n, int64_t n_halo, int mode, uint32_t flags, int64_t t_base_us,
              const vs_carry* carry_in, uint64_t* ticket);
/* Replaces HDLParser::getAllFrames (HDLParser.cxx:495-498): blocks until the batch is done. */
VS_API int vs_wait(vs_ctx* ctx, uint64_t ticket, vs_result* out);
/* Copy point columns [first, first+count) of a finished batch to host arrays (any may be
 * NULL).  Synchronous. */
VS_API int vs_fetch_points(vs_ctx* ctx, uint64_t ticket, int64_t first, int64_t count, float* x,
                    float* y, float* z, uint8_t* intensity, uint8_t* laser, uint16_t* azimuth,
                    uint16_t* distance, uint32_t* t_us);

/* ---- HDLFrame layout built on the device ------------------------------------------------------
 * The reference appends every emitted point to currentFrame->points[laserId] and
 * currentFrame->pointsMeta[laserId] (HDLParser.cxx:733-751) and, when a frame is closed on
 * HDL-64 data, permutes the rows by HDL64BeamLUT (splitFrame, :880-893).  vs_layout_frames()
 * produces exactly those lists for a finished batch as two device arrays of ready-made records
 * -- pcl::PointXYZI {x, y, z, intensity} and PointMeta {azimuth, distance (metres), 3 flag bytes}
 * (type_defs.h:168-176) -- in which every frame is one contiguous run of slots and every row of a
 * frame one contiguous run inside it, rows in the order the closed frame holds them.  A caller
 * builds points[row] from one copy per frame instead of scattering point by point. */
typedef struct vs_frame_rows {
  int64_t  first_slot;                  /* the frame's first slot in the layout arrays          */
  int64_t  n_slots;                     /* its points, carried ones included                    */
  uint32_t row_start[VS_MAX_LASERS];    /* slot of row r, relative to first_slot                */
  uint32_t row_count[VS_MAX_LASERS];    /* elements of row r (the carried ones come first)      */
  uint32_t row_carried[VS_MAX_LASERS];  /* leading elements of row r that earlier batches hold: */
                                        /* left unwritten, the caller copies them in            */
  int32_t  row_laser[VS_MAX_LASERS];    /* laser id pushed into row r (HDL64BeamLUT[r] when the */
                                        /* frame was closed on HDL-64 data, else r)             */
} vs_frame_rows;

typedef struct vs_layout {
  const void* xyzi;           /* device: n_slots records of xyzi_stride bytes                    */
  const void* meta;           /* device: n_slots PointMeta records of 12 bytes; NULL if not asked */
  int64_t n_slots;            /* n_points of the batch + carried points                          */
  const vs_frame_rows* rows;  /* host, one entry per vs_result.frames entry; owned by the context */
  int32_t n_frames;
  int32_t xyzi_stride;
  int32_t n_kernel_launches;
  int32_t reserved;
} vs_layout;

/* Lay a finished batch (after vs_wait) out as HDLFrames.  carried_counts (64 entries by laser
 * id, or NULL): sizes of the open frame's rows as earlier batches left them
 * (currentFrame->points[laser]->size()); the batch's first frame continues that frame, so its
 * rows leave that many leading slots free.  xyzi_stride: 16 (x, y, z, intensity) or 32
 * (pcl::PointXYZI with its padding, data[3] = 1).  The kernels run asynchronously on the
 * ticket's stream; the arrays stay valid until the slot's next submit. */
VS_API int vs_layout_frames(vs_ctx* ctx, uint64_t ticket, const uint32_t* carried_counts,
                            int xyzi_stride, int with_meta, vs_layout* out);
/* Enqueue the copy of slots [first_slot, first_slot + n_slots) to host memory (page-locked for
 * a truly asynchronous copy; either pointer may be NULL).  Returns at once. */
VS_API int vs_fetch_layout(vs_ctx* ctx, uint64_t ticket, int64_t first_slot, int64_t n_slots,
                           void* xyzi_host, void* meta_host);
/* Wait for everything enqueued for the ticket (layout kernels, vs_fetch_layout copies);
 * *layout_ms (may be NULL) = device time of the layout kernels. */
VS_API int vs_sync(vs_ctx* ctx, uint64_t ticket, float* layout_ms);

/* HDLParser::readFrameInformation (HDLParser.cxx:1065-1160) over a packet array: fills up
 * to cap entries (start packet, start block == skips, timestamp) and returns the count in
 * *n_frames.  Equivalent to a VS_MODE_OFFLINE submit that skips the decode. */
VS_API int vs_read_frame_information(vs_ctx* ctx, const uint8_t* pkts, int64_t stride,
                              const int64_t* pkt_time_us, int64_t n, uint32_t flags,
                              int32_t* start_packet, int32_t* skips, int64_t* timestamp_us,
                              int32_t cap, int32_t* n_frames);

/* ---- packet-range sharding of a long recording across GPUs ----------------------
 * The caller being sharded is HDLManager::loadOffline (HDLManager.cxx:103-117) and the frame
 * index it builds: rank g of `world` decodes packets [first, end) of the recording, preceded by
 * n_halo packets (>= one rotation) that only rebuild the parser state (vs_submit's n_halo).  No
 * data-path collective: each rank turns its frame table into rows of int64 (vs_frame_table_rows),
 * the ranks all-gather the rows (NCCL / MPI / shared memory -- transport is the caller's), and
 * vs_stitch_frame_tables() merges the rotation that straddles each shard boundary into one
 * global frame.  Plain host code: no context, no device. */
VS_API int vs_shard_range(int64_t n_packets, int32_t world, int32_t rank, int64_t halo,
                          int64_t* first, int64_t* n_halo, int64_t* end);

#define VS_FRAME_ROW_COLS 10
/* One exchanged row: n_points, first_point (in the rank's own columns), start_packet (GLOBAL
 * packet index, -1: continues the previous rank's open frame), start_block, timestamp_us, skips,
 * closed, hdl64_order, meta_packet (global; -1 carried, -2 none), rank. */
VS_API int vs_frame_table_rows(const vs_frame* frames, int32_t n_frames, int32_t rank,
                               int64_t first_packet, int64_t n_halo, int64_t* rows);

typedef struct vs_frame_segment {  /* the part of a global frame one rank holds */
  int32_t rank;
  int32_t reserved;
  int64_t first_point;
  int64_t n_points;
} vs_frame_segment;

typedef struct vs_global_frame {
  int64_t n_points;
  int64_t start_packet;        /* global packet index of the frame's first block (-1: stream start) */
  int64_t timestamp_us;
  int32_t start_block;
  int32_t skips;
  int32_t closed;
  int32_t hdl64_order;
  int32_t first_segment;       /* its segments: segs[first_segment .. first_segment + n_segments) */
  int32_t n_segments;
  int32_t timestamp_mismatch;  /* the two sides of a shard boundary disagree on the frame's meta */
  int32_t reserved;
} vs_global_frame;

/* vs_frame_table_rows on the device, for a batch in flight or finished: enqueues, behind the
 * batch's kernels on its stream, a kernel (k_table_rows) that writes the batch's frame table as
 * exchange rows into DEVICE memory -- row 0 = {n_rows, 0, ...}, rows 1..n_rows the table, same
 * columns and values as vs_frame_table_rows gives for vs_wait's frames -- so that the
 * all-gather between ranks reads HBM and no host pass sits between decode and collective.
 * d_rows: (cap_rows + 1) x VS_FRAME_ROW_COLS int64.  A batch with more frames than cap_rows
 * writes row 0 = {-n_rows, ...} only.  Does not synchronise; errors of the batch itself
 * (halo, capacity, time range) are still reported by vs_wait. */
VS_API int vs_frame_table_rows_device(vs_ctx* ctx, uint64_t ticket, int32_t rank, int64_t first_packet,
                               int64_t* d_rows, int64_t cap_rows);

/* rows: the ranks' tables in rank order, rows_per_rank[g] rows of VS_FRAME_ROW_COLS each; table g
 * starts at row g * rank_stride_rows (the layout a fixed-size all-gather leaves), or right behind
 * table g-1 when rank_stride_rows is 0.  The open (last) frame of rank g and the first frame of
 * rank g+1 are the same rotation: merged.  Fills at most frame_cap / seg_cap entries; the counts
 * come back in *n_frames / *n_segs (VS_ERR_CAPACITY when an array was too short). */
VS_API int vs_stitch_frame_tables(const int64_t* rows, const int32_t* rows_per_rank, int32_t world,
                                  int64_t rank_stride_rows, vs_global_frame* frames, int32_t frame_cap,
                                  vs_frame_segment* segs, int32_t seg_cap, int32_t* n_frames,
                                  int32_t* n_segs);

/* Page-locked host memory for packet rings / result buffers (cudaHostAlloc / cudaFreeHost),
 * so that callers above the ABI need no CUDA headers. */
VS_API int vs_host_alloc(uint64_t bytes, void** out);
VS_API void vs_host_free(void* p);

/* Device buffers for callers above the ABI that keep a whole recording resident in HBM and
 * decode rotations out of it on demand (HDLManager::loadOffline / prepareFrame,
 * HDLManager.cxx:103-117, 195-211: the reference re-reads the pcap file for every frame it
 * re-decodes).  Pointers returned here are what VS_FLAG_DEVICE_INPUT expects. */
VS_API int vs_device_alloc(vs_ctx* ctx, uint64_t bytes, void** out_dev);
VS_API void vs_device_free(vs_ctx* ctx, void* dev);
VS_API int vs_device_upload(vs_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);

/* ---- online front end in batch form ------------------------------- */

/* State of TimeSolver::calcTimestamp(uint32_t) (TimeSolver.cxx:34-49): hdlHourTime + hdlOffset
 * folded into one number, lastHdlReport and hdlInited.  Zero-initialise before the first call. */
typedef struct vs_time_solver {
  int64_t  base_us;
  uint32_t last_report;
  int32_t  inited;
} vs_time_solver;

/* TimeSolver::calcTimestamp(uint32_t microsecToHour) for n packets (HDLSource.cxx:216-217):
 * packet time = base + whole hours wrapped so far + the packet's gpsTimestamp field (payload
 * bytes 1200-1203); the hour wraps are a 1-bit scan over the array, done on the GPU.
 * now_us is the local clock the reference reads at the very first packet (ignored once
 * state->inited).  With VS_FLAG_DEVICE_INPUT pkts and out_time_us are device pointers (the
 * times can go straight into vs_submit), else host pointers.  state is updated in place. */
VS_API int vs_solve_packet_times(vs_ctx* ctx, const uint8_t* pkts, int64_t stride, int64_t n,
                          uint32_t flags, int64_t now_us, vs_time_solver* state,
                          int64_t* out_time_us);

/* NovAtel INSPVA record as the reference lays it out (type_defs.h:39-58). */
typedef struct vs_ins_pva {
  uint16_t message_id;
  uint16_t week_number;
  uint32_t milliseconds;
  uint32_t week_number_pos;
  uint32_t pad0;
  double   seconds_pos;
  double   llh[3];       /* latitude, longitude (degrees), height (m) */
  double   v[3];
  double   eulr[3];      /* roll, pitch, yaw (degrees) */
  int32_t  ins_status;
  int32_t  pad1;
} vs_ins_pva;

/* INSSource's PacketConsumer::calcTransform (INSSource.cxx:300-326) for n records: T = ENU of
 * the position about the ECEF origin (llh2enu, CoordiTran.cpp:271-276), R = eulr, V = v, and
 * TimeSolver::calcTimestamp(InsPVA const*) (TimeSolver.cxx:20-33) = arrival + (time of pose -
 * time of packet send), arrival_us being the local clock at reception of each record.
 * Host arrays in and out (out_trv = n x 9); the result is what vs_set_poses takes. */
VS_API int vs_poses_from_ins(vs_ctx* ctx, const vs_ins_pva* recs, int64_t n, const double origin_xyz[3],
                      const int64_t* arrival_us, int64_t* out_t_us, double* out_trv);

/* The CUDA stream the context launches on (cudaStream_t), for callers that time or order
 * work against it. */
VS_API void* vs_stream(vs_ctx* ctx);
/* Each result slot launches on its own stream (ticket t uses slot t % n_slots). */
VS_API void* vs_slot_stream(vs_ctx* ctx, int slot);

#ifdef __cplusplus
}
#endif
#endif
