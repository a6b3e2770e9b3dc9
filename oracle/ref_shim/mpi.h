// mpi.h — SINGLE-RANK stand-in for MPI, just enough for /root/reference/ExodusIO.hpp to compile and for its
// 1-rank code paths to run (TEST INFRASTRUCTURE: part of the recipe that runs the reference's own assemble /
// decompose / writeSolution code here; see oracle/ref_shim/README.md).  Anything that would need a second
// rank aborts loudly.
#pragma once
#include <cstdio>
#include <cstdlib>

typedef int MPI_Comm;
typedef int MPI_Info;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef long MPI_Aint;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
struct shim_mpi_win { void *base; };
typedef shim_mpi_win *MPI_Win;
typedef MPI_Win MPI_Window;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_LOCK_SHARED 1
#define MPI_INT 1
#define MPI_LONG 2
#define MPI_LONG_LONG 3
#define MPI_UNSIGNED_LONG 4
#define MPI_UNSIGNED_LONG_LONG 5
#define MPI_DOUBLE 6
#define MPI_BYTE 7
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

[[noreturn]] inline void shim_mpi_needs_ranks(const char *what) {
    std::fprintf(stderr, "ref_shim: %s called — this stand-in only runs the reference on ONE rank\n", what);
    std::abort();
}
inline int MPI_Win_allocate(MPI_Aint size, int, MPI_Info, MPI_Comm, void *baseptr, MPI_Win *win) {
    *win = new shim_mpi_win{std::calloc(size > 0 ? (size_t)size : 1, 1)};
    *(void **)baseptr = (*win)->base;
    return MPI_SUCCESS;
}
inline int MPI_Win_create(void *base, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win *win) { *win = new shim_mpi_win{nullptr}; (void)base; return MPI_SUCCESS; }
inline int MPI_Win_free(MPI_Win *win) { if (*win) { std::free((*win)->base); delete *win; *win = nullptr; } return MPI_SUCCESS; }
inline int MPI_Win_lock(int, int, int, MPI_Win) { return MPI_SUCCESS; }
inline int MPI_Win_unlock(int, MPI_Win) { return MPI_SUCCESS; }
inline int MPI_Alloc_mem(MPI_Aint size, MPI_Info, void *baseptr) { *(void **)baseptr = std::calloc(size > 0 ? (size_t)size : 1, 1); return MPI_SUCCESS; }
inline int MPI_Get(void *, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { shim_mpi_needs_ranks("MPI_Get"); }
// IO::getMatrix posts Isends to its OWN rank that nobody receives, through request slots of a vector sized
// 3*(ranks-1) = 0 (ExodusIO.hpp:1224-1235).  A real MPI buffers such small eager sends; here a send
// to rank 0 is dropped and the (dangling) request slot is left untouched.
inline int MPI_Isend(const void *, int, MPI_Datatype, int dest, int, MPI_Comm, MPI_Request *) {
    if (dest != 0) shim_mpi_needs_ranks("MPI_Isend to another rank");
    return MPI_SUCCESS;
}
inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { shim_mpi_needs_ranks("MPI_Irecv"); }
inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { shim_mpi_needs_ranks("MPI_Send"); }
inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { shim_mpi_needs_ranks("MPI_Recv"); }
inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { shim_mpi_needs_ranks("MPI_Probe"); }
inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { shim_mpi_needs_ranks("MPI_Get_count"); }
inline int MPI_Waitall(int count, MPI_Request *, MPI_Status *) { if (count > 0) shim_mpi_needs_ranks("MPI_Waitall"); return MPI_SUCCESS; }
