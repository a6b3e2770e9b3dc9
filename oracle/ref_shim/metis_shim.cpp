// 32-bit idx_t front of the bundled 64-bit METIS (see metis.h).  TEST INFRASTRUCTURE.
#include <cstdint>
#include <vector>
extern "C" int METIS_PartMeshDual(int64_t *ne, int64_t *nn, int64_t *eptr, int64_t *eind, int64_t *vwgt, int64_t *vsize,
                                  int64_t *ncommon, int64_t *nparts, float *tpwgts, int64_t *options, int64_t *objval,
                                  int64_t *epart, int64_t *npart);
extern "C" int shim_METIS_PartMeshDual(int32_t *ne, int32_t *nn, int32_t *eptr, int32_t *eind, int32_t *vwgt, int32_t *vsize,
                                       int32_t *ncommon, int32_t *nparts, float *tpwgts, int32_t *options, int32_t *objval,
                                       int32_t *epart, int32_t *npart) {
    if (vwgt || vsize || options) return -2;             // the reference passes NULL for all three (ExodusIO.hpp:1597-1603)
    int64_t ne_ = *ne, nn_ = *nn, nc = *ncommon, np = *nparts, obj = 0;
    std::vector<int64_t> ep(eptr, eptr + *ne + 1), ei(eind, eind + eptr[*ne]), pe((size_t)*ne), pn((size_t)*nn);
    int rc = 1;
    if (np > 1) rc = METIS_PartMeshDual(&ne_, &nn_, ep.data(), ei.data(), nullptr, nullptr, &nc, &np, tpwgts, nullptr, &obj, pe.data(), pn.data());
    for (int64_t i = 0; i < ne_; ++i) epart[i] = (int32_t)pe[(size_t)i];
    for (int64_t i = 0; i < nn_; ++i) npart[i] = (int32_t)pn[(size_t)i];
    *objval = (int32_t)obj;
    return rc;
}
