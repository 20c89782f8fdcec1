// VeloSLAM.h -- umbrella header of the drop-in facade (the reference ships one of the same name).
#ifndef VELOSLAM_B200_VELOSLAM_H
#define VELOSLAM_B200_VELOSLAM_H
#include "CoordiTran.h"
#include "HDLFrame.h"
#include "HDLManager.h"
#include "HDLParser.h"
#include "HDLSource.h"
#include "INSSource.h"
#include "TimeLine.h"
#include "TimeSolver.h"
#include "TransformManager.h"
#include "type_defs.h"
#include "vtkPacketFile.h"
#endif
