// exodusII.h — stand-in for the SEACAS Exodus-II C API as /root/reference/ExodusIO.hpp calls it (42 entry
// points).  Reads come from a flat dump of the mesh made by oracle/ref_shim/dump_exo.py with
// scipy.io.netcdf_file (independent of the product's netCDF reader); writes are recorded and flushed to
// "<file>.shimdump" by ex_close so tests can compare what the reference WOULD have written.
// TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
#pragma once
#include <cstdint>

#define MAX_STR_LENGTH 32L
#define MAX_LINE_LENGTH 80L
#define EX_READ 0x0002
#define EX_WRITE 0x0001
#define EX_CLOBBER 0x0008

typedef void void_int;
typedef int64_t ex_entity_id;
enum ex_entity_type { EX_NODAL = 14, EX_ELEM_BLOCK = 1, EX_NODE_SET = 2, EX_SIDE_SET = 3, EX_ELEM_MAP = 4, EX_NODE_MAP = 5 };
enum ex_inquiry { EX_INQ_QA = 8, EX_INQ_INFO = 15, EX_INQ_NS_PROP = 27, EX_INQ_SS_PROP = 28 };

typedef struct ex_init_params {
    char title[MAX_LINE_LENGTH + 1];
    int64_t num_dim, num_nodes, num_edge, num_edge_blk, num_face, num_face_blk, num_elem, num_elem_blk, num_node_sets,
        num_edge_sets, num_face_sets, num_side_sets, num_elem_sets, num_node_maps, num_edge_maps, num_face_maps, num_elem_maps,
        num_assembly, num_blob;
} ex_init_params;

extern "C" {
int ex_open(const char *path, int mode, int *cpu_ws, int *io_ws, float *version);
int ex_create(const char *path, int mode, int *cpu_ws, int *io_ws);
int ex_close(int exoid);
int ex_get_init_ext(int exoid, ex_init_params *p);
int ex_get_id_map(int exoid, ex_entity_type t, void_int *map);
int ex_get_ids(int exoid, ex_entity_type t, void_int *ids);
int ex_get_set_param(int exoid, ex_entity_type t, ex_entity_id id, void_int *n, void_int *ndf);
int ex_get_set(int exoid, ex_entity_type t, ex_entity_id id, void_int *entries, void_int *extra);
int ex_get_set_dist_fact(int exoid, ex_entity_type t, ex_entity_id id, void *df);
int ex_get_block(int exoid, ex_entity_type t, ex_entity_id id, char *elem_type, void_int *nelem, void_int *npe, void_int *nedge,
                 void_int *nface, void_int *nattr);
int ex_get_elem_conn(int exoid, ex_entity_id id, void_int *conn);
int ex_get_coord(int exoid, void *x, void *y, void *z);
int ex_get_coord_names(int exoid, char **names);
int ex_get_map(int exoid, void_int *map);
int ex_get_node_num_map(int exoid, void_int *map);
int ex_inquire(int exoid, int req, void_int *ret_int, float *ret_float, char *ret_char);
int ex_get_prop_names(int exoid, ex_entity_type t, char **names);
int ex_get_prop_array(int exoid, ex_entity_type t, const char *name, void_int *values);
int ex_get_prop(int exoid, ex_entity_type t, ex_entity_id id, const char *name, void_int *value);
int ex_get_side_set_node_list(int exoid, ex_entity_id id, void_int *cnt, void_int *list);
int ex_get_qa(int exoid, char *qa[][4]);
int ex_get_info(int exoid, char **info);

int ex_put_init(int exoid, const char *title, int64_t ndim, int64_t nnodes, int64_t nelem, int64_t nblk, int64_t nns, int64_t nss);
int ex_put_coord(int exoid, const void *x, const void *y, const void *z);
int ex_put_coord_names(int exoid, char **names);
int ex_put_map(int exoid, const void_int *map);
int ex_put_block(int exoid, ex_entity_type t, ex_entity_id id, const char *elem_type, int64_t nelem, int64_t npe, int64_t nedge,
                 int64_t nface, int64_t nattr);
int ex_put_conn(int exoid, ex_entity_type t, ex_entity_id id, const void_int *conn, const void_int *e, const void_int *f);
int ex_put_set_param(int exoid, ex_entity_type t, ex_entity_id id, int64_t n, int64_t ndf);
int ex_put_set(int exoid, ex_entity_type t, ex_entity_id id, const void_int *entries, const void_int *extra);
int ex_put_set_dist_fact(int exoid, ex_entity_type t, ex_entity_id id, const void *df);
int ex_put_prop_names(int exoid, ex_entity_type t, int n, char **names);
int ex_put_prop_array(int exoid, ex_entity_type t, const char *name, const void_int *values);
int ex_put_prop(int exoid, ex_entity_type t, ex_entity_id id, const char *name, ex_entity_id value);
int ex_put_qa(int exoid, int n, char *qa[][4]);
int ex_put_info(int exoid, int n, char **info);
int ex_put_node_num_map(int exoid, const void_int *map);
int ex_put_variable_param(int exoid, ex_entity_type t, int n);
int ex_put_variable_names(int exoid, ex_entity_type t, int n, char **names);
int ex_put_time(int exoid, int step, const void *t);
int ex_put_nodal_var(int exoid, int step, int var, int64_t n, const void *vals);
}
