// common.cuh — shared declarations of libheat_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>

#include "../../include/heat_b200.h"

namespace heat {

void set_error(const char *fmt, ...);
void count_launch();

#define HEAT_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            heat::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 100;                                                                          \
        }                                                                                        \
    } while (0)

// placed right after every <<<>>> launch: counts it (heat_kernel_launches) and checks for launch errors
#define HEAT_LAUNCHED()                                                                          \
    do {                                                                                         \
        heat::count_launch();                                                                    \
        HEAT_CUDA(cudaGetLastError());                                                           \
    } while (0)

#define HEAT_TRY(call)                                                                           \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        if (rc__ != 0) return rc__;                                                              \
    } while (0)

#define HEAT_FAIL(code, ...)                                                                     \
    do {                                                                                         \
        heat::set_error(__VA_ARGS__);                                                            \
        return (code);                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device buffer (cudaMalloc owned)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return 101;
        }
        n = count;
        return 0;
    }
};

constexpr int kSellRowsPerLane = 2;                      // R
constexpr int kSellChunk = 32 * kSellRowsPerLane;        // C: rows per SELL slice (one warp)
constexpr int kSellDictCap = 64;                         // most distinct col-row offsets a slice may have to be byte-indexed
constexpr int kMaxPartials = 4096;
constexpr int kMaxDevices = 64;                          // per-device caches (function attributes, SM counts)                       // per-launch block partials capacity

// one entry per slice in SpMV processing order (interior slices first, then boundary slices; natural
// order on one GPU): everything a warp needs to start streaming the slice, in ONE 16-byte load
// base64 = entry offset / 64 (values: 8 B per entry), cbase64 = byte offset of the slice's column data / 64 in the
// column stream, w_mode = entries per row | column mode << 24 (0: 1-byte table index, 1: int16 col-row delta,
// 2: int32 column id), s = slice id
struct SliceMeta { uint32_t base64; uint32_t cbase64; int32_t w_mode; int32_t s; };
enum { kColModeU8 = 0, kColModeI16 = 1, kColModeI32 = 2 };

// scalar slots of the CG state (device doubles)
enum {
    S_RZ0 = 0, S_RZ1 = 1,        // gamma = r.z, ping-pong by iteration parity
    S_PAP0 = 2, S_PAP1 = 3,      // delta = p.Ap (classical) / w.u (single-reduce), ping-pong
    S_ALPHA0 = 4, S_ALPHA1 = 5,  // alpha ping-pong (single-reduce variant)
    S_RR = 6,                    // ||r||^2 (current)
    S_RR0 = 7,                   // ||r0||^2
    S_TOL2 = 8,                  // tol^2
    S_TMP0 = 9, S_TMP1 = 10, S_TMP2 = 11,
    S_COUNT = 16
};
enum { I_ITERS = 0, I_STATUS = 1, I_COUNTER = 2, I_COUNTER2 = 3, I_COUNT = 8 };

}  // namespace heat

namespace heat { struct IluState; }   // ilu.cu: ILU(0) factors + level schedules of one matrix
struct PeerMatrixState;    // comm.cu: CUDA-IPC mappings + push plans of one matrix (peer-memory halo path)

struct heat_vector {
    heat_ctx *ctx = nullptr;
    int64_t n_owned = 0, n_ghost = 0;
    heat::DevBuf<double> d;     // [n_owned + n_ghost]
};

struct HaloPlan {
    int n_neighbors = 0;
    std::vector<int32_t> nbr_rank;
    std::vector<int64_t> send_ptr;   // [n_neighbors+1]
    std::vector<int32_t> send_idx;   // local owned row ids, concatenated by neighbour
    std::vector<int64_t> recv_ptr;   // [n_neighbors+1] offsets into the ghost segment
    heat::DevBuf<int32_t> d_send_idx;
    heat::DevBuf<double> d_send_buf;
};

struct heat_matrix {
    heat_ctx *ctx = nullptr;
    int op_mode = 0;
    int64_t n_global = 0, nnz_global = 0;
    int64_t n_owned = 0, n_ghost = 0, nnz = 0;
    int32_t max_row_len = 0;
    // local CSR (kept for export / parity; columns are LOCAL ids: owned first, then ghosts)
    heat::DevBuf<int64_t> row_ptr;
    heat::DevBuf<int32_t> col;
    heat::DevBuf<double> val;
    // SELL-C (C = kSellChunk, R rows per lane), the SpMV format
    int64_t n_slices = 0, sell_padded = 0;
    heat::DevBuf<int64_t> slice_ptr;     // [n_slices+1] entry offsets
    heat::DevBuf<int32_t> sell_col;
    heat::DevBuf<double> sell_val;
    // compact column stream, chosen PER SLICE (sell.cu): a slice with <= kSellDictCap distinct col-row offsets stores
    // a 1-byte index into its offset table (col = row + sell_tab[slice][idx]); any other slice whose offsets fit 16
    // bits stores int16 deltas (col = row + delta) — every mesh of moderate bandwidth; if some slice fits neither the
    // matrix keeps the int32 stream (sell_col).  sell_cmode = widest entry of the stream: 1 (all slices table-indexed),
    // 2 (table-indexed and int16 slices mixed; slice_cptr / slice_mode say which and where), 4 (int32: sell_col)
    int sell_cmode = 4;
    heat::DevBuf<int64_t> slice_cptr;    // [n_slices+1] byte offsets into sell_idx8 (mixed streams only)
    heat::DevBuf<uint8_t> slice_mode;    // [n_slices] kColMode* (mixed streams only)
    heat::DevBuf<uint8_t> sell_rowlen;   // stored entries per row — only for matrices assembled straight into SELL (no CSR
                                         // until sell_to_csr rebuilds one for an export or ILU)
    heat::DevBuf<uint8_t> sell_idx8;
    heat::DevBuf<int32_t> sell_tab;
    int sell_tpad = 0;
    heat::DevBuf<int32_t> slices_interior, slices_boundary;   // slice id lists (multi-GPU overlap)
    heat::DevBuf<int32_t> slices_all;                         // interior list followed by boundary list
    heat::DevBuf<heat::SliceMeta> slice_meta;                 // [n_slices] in slices_all order (natural order without ghosts)
    int64_t n_int_slices = 0, n_bnd_slices = 0;
    heat::DevBuf<double> dinv;           // 1/diag (owned)
    heat::DevBuf<double> diag;
    // maps
    std::vector<int64_t> owned_gids, ghost_gids, red2orig_owned;
    std::vector<int32_t> ghost_owner;
    bool owned_contiguous = true;        // owned gids = gid0 .. gid0+n_owned-1
    int64_t gid0 = 0;
    heat::DevBuf<int64_t> d_owned_gids;  // only when !owned_contiguous
    HaloPlan halo;
    // solver workspace (lazily allocated)
    heat::DevBuf<double> w_r, w_r2, w_p, w_p2, w_ap, w_s, w_u, w_u2, w_t, w_w, w_w2;
    heat::DevBuf<double> h_x, h_b;       // staging of heat_solve_host (x with ghosts, b)
    heat::DevBuf<double> h_x2, h_b2;     // second staging set of heat_solve_host_batch (copies overlap the neighbouring solves)
    PeerMatrixState *peer = nullptr;     // non-null once the peer-memory halo path is set up
    heat::IluState *ilu = nullptr;       // non-null once HEAT_PREC_ILU0 has been set up
    heat::DevBuf<double> partials;       // [2 * kMaxPartials * 4]
    heat::DevBuf<double> scal;           // [S_COUNT]
    heat::DevBuf<int> iscal;             // [I_COUNT]
    double cheb_lmax_est = 0.0;          // lambda_max(D^-1 A) estimated by 10 power iterations, once per matrix
    double assemble_ms = 0.0;
    double assemble_fill_ms = 0.0;       // the matrix-fill kernel alone (cube_sell_kernel / values_kernel), CUDA events
    double asm_phase_ms[4] = {0, 0, 0, 0};   // explicit-mesh path: node->element sort, pattern count, pattern fill, values
};

struct ExoFile;   // exodus.hpp

struct HostMesh {
    bool valid = false;
    bool is_cube = false;
    bool cube_explicit = false;
    int nx = 0, ny = 0, nz = 0;
    int64_t num_nodes = 0, num_elem = 0;
    int num_dim = 3, npe = 0;
    std::vector<double> x, y, z;
    std::vector<int32_t> conn;                       // 0-based [num_elem][npe]
    std::map<int64_t, std::vector<int64_t>> nodesets; // id -> 0-based nodes (nodeSetMap, ExodusIO.hpp:2088)
    std::string elem_type;
};

struct heat_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t comm_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // heat_solve_host: b goes up beside the set-up SpMV
    cudaEvent_t ev_copy = nullptr;
    cudaStream_t copy_out_stream = nullptr;              // heat_solve_host_batch: results go down while the next system is solved
    cudaEvent_t ev_in[2] = {}, ev_done[2] = {}, ev_out[2] = {};   // per staging set: inputs up, solve done, result down
    cudaEvent_t wait_before_rhs = nullptr;   // if set, solve_device waits for it before it first reads b
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_halo = nullptr, ev_pack = nullptr;
    int rank = 0, nranks = 1;
    void *nccl_comm = nullptr;           // ncclComm_t
    // peer-memory path (peer.cuh): IPC-mapped arenas of all ranks of the node
    bool peer_enabled = false;
    unsigned long long *peer_arena = nullptr;                 // my arena (device)
    unsigned long long *peer_arena_of[8] = {};                // every rank's arena as mapped here
    unsigned long long peer_red_seq = 0, peer_halo_epoch = 0; // monotonic, identical on all ranks
    HostMesh mesh;
    ExoFile *read_file = nullptr;        // readFID  (ExodusIO.hpp:2082)
    ExoFile *write_file = nullptr;       // writeFID (ExodusIO.hpp:2083)
    std::string write_path;
    bool printed_time_zero = false;      // printedTimeZero (ExodusIO.hpp:2084)
    int out_word_size = 8;               // Exodus word size of the output file; the reference: sizeof(real_t) (:104-105)
    bool out_largest_id = false;         // output field of a node in several nodesets: largest id (:1983-1989) instead of the RHS's
    // state cached by assemble for writeSolution (ExodusIO.hpp:2086-2098)
    std::vector<double> node_bc;         // NaN for DOF nodes, else lowest nodeset id
    std::vector<double> node_bc_hi;      // highest nodeset id (reference output rule, :1983-1989)
    int64_t n_global = 0;
    std::vector<int64_t> owned_gids;     // reduced global ids of this rank's rows (globalIDMap keys, :2092-2098)
};

// node_bc / node_bc_hi of the context's mesh from its nodesets (api.cu); host only
namespace heat { void build_node_bc(heat_ctx *ctx); }

