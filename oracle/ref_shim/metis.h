// metis.h — the reference needs a METIS built with 32-bit idx_t (its arrays go straight into the Exodus C API,
// SURVEY.md §8 a15).  The only METIS in this image is the CUDA toolkit's libmetis_static.a (idx_t = int64,
// real_t = float); metis_shim.cpp converts and calls THAT library, so the partition is the real METIS answer.
// TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include <cstdint>
typedef int32_t idx_t;
typedef float real_t;
enum { METIS_OK = 1, METIS_ERROR_INPUT = -2, METIS_ERROR_MEMORY = -3, METIS_ERROR = -4 };
extern "C" int shim_METIS_PartMeshDual(idx_t *ne, idx_t *nn, idx_t *eptr, idx_t *eind, idx_t *vwgt, idx_t *vsize, idx_t *ncommon,
                                       idx_t *nparts, real_t *tpwgts, idx_t *options, idx_t *objval, idx_t *epart, idx_t *npart);
#define METIS_PartMeshDual shim_METIS_PartMeshDual
