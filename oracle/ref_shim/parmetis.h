// parmetis.h — ParMETIS does not exist in this image.  IO::getMatrix (the only caller, ExodusIO.hpp:919) asks for
// nparts = #ranks; on ONE rank every k-way partitioner answers "all elements in part 0", which is what this
// stand-in returns.  Anything else aborts loudly.  TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include "metis.h"
#include "mpi.h"
inline int ParMETIS_V3_PartMeshKway(idx_t *elmdist, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *nparts,
                                    real_t *, real_t *, idx_t *, idx_t *edgecut, idx_t *part, MPI_Comm *) {
    if (*nparts != 1) shim_mpi_needs_ranks("ParMETIS_V3_PartMeshKway with nparts > 1");
    for (idx_t i = 0; i < elmdist[1] - elmdist[0]; ++i) part[i] = 0;
    *edgecut = 0;
    return METIS_OK;
}
