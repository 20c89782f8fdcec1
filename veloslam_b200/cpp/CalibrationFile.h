// CalibrationFile.h -- reader for the Velodyne db.xml calibration (boost-serialization XML),
// the fields the reference's loadCorrectionsFile reads (HDLParser.cxx:771-858).
#ifndef VELOSLAM_B200_CALIBRATIONFILE_H
#define VELOSLAM_B200_CALIBRATIONFILE_H

#include <string>

#include "../../include/veloslam_b200.h"

struct CalibrationFile {
  vs_laser_corr rows[VS_MAX_LASERS];
  int n_rows;     // highest id_ + 1
  int n_enabled;  // enabled_ items equal to 1
  // false when the file cannot be read or holds no <px> item
  bool load(const std::string& filename, std::string* error);
};

#endif
