// CoordiTran.h -- WGS-84 geodesy of the drop-in facade with the reference's function names
// (/root/reference/CoordiTran.h, CoordiTran.cpp:51-187, 271-276): geodetic (radians, metres)
// <-> ECEF <-> local ENU about an ECEF origin.  Host functions for callers that convert single
// INS records (INSSource.cxx:300-326); arrays go through vs_poses_from_ins on the GPU.
#ifndef VELOSLAM_B200_COORDITRAN_H
#define VELOSLAM_B200_COORDITRAN_H

void llh2xyz(double llh[3], double xyz[3]);
void xyz2llh(double xyz[3], double llh[3]);
void xyz2enu(double xyz[3], double orgxyz[3], double enu[3]);
void llh2enu(double llh[3], double orgxyz[3], double enu[3]);

#endif
