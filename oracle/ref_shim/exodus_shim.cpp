// exodus_shim.cpp — see exodusII.h.  TEST INFRASTRUCTURE (oracle/ref_shim/README.md).
//
// Container format shared with dump_exo.py and the tests: a sequence of records
//     "<name> <dtype> <count>\n" followed by count items of raw little-endian data,  dtype in {i32, f64, f32, str}.
// Reading side: $REF_SHIM_DUMP_DIR/<basename of the .exo>.dump  (ints are int32 as in every file under the
// reference's data/, floats are handed out in the word size ex_open was given: the reference asks for
// sizeof(real_t) = 4, which is how it ends up writing float32 results).
#include "exodusII.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

struct Rec { std::string dtype; std::vector<char> raw; size_t count = 0; };
struct File {
    bool writing = false;
    std::string path;
    int cpu_ws = 8;
    std::map<std::string, Rec> rec;                 // read side: dump contents; write side: what was put
    std::vector<std::string> order;                 // write side: record order
};
std::map<int, File> g_files;
int g_next = 100;

bool load(const std::string &dump, File &f) {
    FILE *fp = std::fopen(dump.c_str(), "rb");
    if (!fp) return false;
    char name[256], dtype[16];
    unsigned long long count;
    while (std::fscanf(fp, "%255s %15s %llu", name, dtype, &count) == 3) {
        std::fgetc(fp);                                                    // the newline
        Rec r;
        r.dtype = dtype; r.count = (size_t)count;
        const size_t sz = (r.dtype == "f64") ? 8 : (r.dtype == "str") ? 1 : 4;
        r.raw.resize(r.count * sz);
        if (r.count && std::fread(r.raw.data(), 1, r.raw.size(), fp) != r.raw.size()) { std::fclose(fp); return false; }
        f.rec[name] = std::move(r);
    }
    std::fclose(fp);
    return true;
}
const Rec *get(int id, const std::string &name) {
    auto it = g_files.find(id);
    if (it == g_files.end()) return nullptr;
    auto r = it->second.rec.find(name);
    return r == it->second.rec.end() ? nullptr : &r->second;
}
int64_t scalar(int id, const std::string &name, int64_t dflt = 0) {
    const Rec *r = get(id, name);
    if (!r || r->count == 0) return dflt;
    int32_t v; std::memcpy(&v, r->raw.data(), 4);
    return v;
}
int copy_i32(int id, const std::string &name, void *dst) {
    const Rec *r = get(id, name);
    if (!r) return 1;
    if (dst && r->count) std::memcpy(dst, r->raw.data(), r->count * 4);
    return 0;
}
int copy_real(int id, const std::string &name, void *dst) {       // dump holds f64; hand out cpu_ws-sized reals
    const Rec *r = get(id, name);
    if (!r) return 1;
    if (!dst) return 0;
    const double *s = reinterpret_cast<const double *>(r->raw.data());
    if (g_files[id].cpu_ws == 4) for (size_t i = 0; i < r->count; ++i) static_cast<float *>(dst)[i] = (float)s[i];
    else std::memcpy(dst, s, r->count * 8);
    return 0;
}
int index_of(int id, const char *ids_name, int64_t want) {         // 1-based position of an entity id
    const Rec *r = get(id, ids_name);
    if (!r) return 0;
    for (size_t i = 0; i < r->count; ++i) { int32_t v; std::memcpy(&v, r->raw.data() + 4 * i, 4); if (v == want) return (int)i + 1; }
    return 0;
}
const char *prefix_of(ex_entity_type t) { return t == EX_NODE_SET ? "ns" : t == EX_SIDE_SET ? "ss" : "eb"; }

// ---- write side -------------------------------------------------------------------------------------
void put(int id, const std::string &name, const char *dtype, const void *data, size_t count) {
    File &f = g_files[id];
    Rec r;
    r.dtype = dtype; r.count = count;
    const size_t sz = (r.dtype == "f64") ? 8 : (r.dtype == "str") ? 1 : 4;
    r.raw.assign(static_cast<const char *>(data), static_cast<const char *>(data) + count * sz);
    if (!f.rec.count(name)) f.order.push_back(name);
    f.rec[name] = std::move(r);
}
void put_i32(int id, const std::string &name, int64_t v) { const int32_t x = (int32_t)v; put(id, name, "i32", &x, 1); }
void put_real(int id, const std::string &name, const void *data, size_t count) {
    put(id, name, g_files[id].cpu_ws == 4 ? "f32" : "f64", data, count);
}
int next_index(int id, const std::string &counter) {               // running count of blocks / sets written
    const int64_t n = scalar(id, counter, 0) + 1;
    put_i32(id, counter, n);
    return (int)n;
}

}  // namespace

extern "C" {

int ex_open(const char *path, int, int *cpu_ws, int *, float *version) {
    const char *dir = std::getenv("REF_SHIM_DUMP_DIR");
    std::string base = path;
    const size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    File f;
    f.path = path; f.cpu_ws = cpu_ws ? *cpu_ws : 8;
    if (!load(std::string(dir ? dir : ".") + "/" + base + ".dump", f)) return -1;
    if (version) *version = 8.03f;
    g_files[g_next] = std::move(f);
    return g_next++;
}
int ex_create(const char *path, int, int *cpu_ws, int *) {
    File f;
    f.writing = true; f.path = path; f.cpu_ws = cpu_ws ? *cpu_ws : 8;
    g_files[g_next] = std::move(f);
    return g_next++;
}
int ex_close(int id) {
    auto it = g_files.find(id);
    if (it == g_files.end()) return 1;
    if (it->second.writing) {
        FILE *fp = std::fopen((it->second.path + ".shimdump").c_str(), "wb");
        if (!fp) return 1;
        for (const std::string &n : it->second.order) {
            const Rec &r = it->second.rec[n];
            std::fprintf(fp, "%s %s %llu\n", n.c_str(), r.dtype.c_str(), (unsigned long long)r.count);
            if (!r.raw.empty()) std::fwrite(r.raw.data(), 1, r.raw.size(), fp);
        }
        std::fclose(fp);
    }
    g_files.erase(it);
    return 0;
}
int ex_get_init_ext(int id, ex_init_params *p) {
    if (!g_files.count(id)) return 1;
    std::memset(p, 0, sizeof(*p));
    if (const Rec *t = get(id, "title")) std::memcpy(p->title, t->raw.data(), t->count < (size_t)MAX_LINE_LENGTH ? t->count : (size_t)MAX_LINE_LENGTH);
    p->num_dim = scalar(id, "num_dim"); p->num_nodes = scalar(id, "num_nodes"); p->num_elem = scalar(id, "num_elem");
    p->num_elem_blk = scalar(id, "num_el_blk"); p->num_node_sets = scalar(id, "num_node_sets"); p->num_side_sets = scalar(id, "num_side_sets");
    return 0;
}
int ex_get_id_map(int id, ex_entity_type t, void_int *map) { return copy_i32(id, t == EX_NODE_MAP ? "node_num_map" : "elem_map", map); }
int ex_get_node_num_map(int id, void_int *map) { return copy_i32(id, "node_num_map", map); }
int ex_get_map(int id, void_int *map) { return copy_i32(id, "elem_map", map); }
int ex_get_ids(int id, ex_entity_type t, void_int *ids) { return copy_i32(id, std::string(prefix_of(t)) + "_ids", ids); }
int ex_get_set_param(int id, ex_entity_type t, ex_entity_id sid, void_int *n, void_int *ndf) {
    const std::string p = prefix_of(t);
    const int k = index_of(id, (p + "_ids").c_str(), sid);
    if (!k) return 1;
    const Rec *e = get(id, p + std::to_string(k) + "_entries"), *d = get(id, p + std::to_string(k) + "_df");
    *static_cast<int32_t *>(n) = e ? (int32_t)e->count : 0;
    *static_cast<int32_t *>(ndf) = d ? (int32_t)d->count : 0;
    return 0;
}
int ex_get_set(int id, ex_entity_type t, ex_entity_id sid, void_int *entries, void_int *extra) {
    const std::string p = prefix_of(t);
    const int k = index_of(id, (p + "_ids").c_str(), sid);
    if (!k) return 1;
    if (copy_i32(id, p + std::to_string(k) + "_entries", entries)) return 1;
    if (extra) copy_i32(id, p + std::to_string(k) + "_extra", extra);
    return 0;
}
int ex_get_set_dist_fact(int id, ex_entity_type t, ex_entity_id sid, void *df) {
    const std::string p = prefix_of(t);
    const int k = index_of(id, (p + "_ids").c_str(), sid);
    return k ? copy_real(id, p + std::to_string(k) + "_df", df) : 1;
}
int ex_get_block(int id, ex_entity_type, ex_entity_id bid, char *elem_type, void_int *nelem, void_int *npe, void_int *nedge,
                 void_int *nface, void_int *nattr) {
    const int k = index_of(id, "eb_ids", bid);
    if (!k) return 1;
    const std::string b = "eb" + std::to_string(k);
    if (const Rec *t = get(id, b + "_type")) { std::memcpy(elem_type, t->raw.data(), t->count); elem_type[t->count] = 0; }
    *static_cast<int32_t *>(nelem) = (int32_t)scalar(id, b + "_nelem");
    *static_cast<int32_t *>(npe) = (int32_t)scalar(id, b + "_npe");
    *static_cast<int32_t *>(nedge) = 0; *static_cast<int32_t *>(nface) = 0; *static_cast<int32_t *>(nattr) = 0;
    return 0;
}
int ex_get_elem_conn(int id, ex_entity_id bid, void_int *conn) {
    const int k = index_of(id, "eb_ids", bid);
    return k ? copy_i32(id, "eb" + std::to_string(k) + "_conn", conn) : 1;
}
int ex_get_coord(int id, void *x, void *y, void *z) {
    copy_real(id, "coordx", x); copy_real(id, "coordy", y);
    if (z) copy_real(id, "coordz", z);
    return 0;
}
int ex_get_coord_names(int id, char **names) {
    const int64_t nd = scalar(id, "num_dim");
    for (int64_t i = 0; i < nd; ++i) { names[i][0] = (char)('x' + i); names[i][1] = 0; }
    return 0;
}
int ex_inquire(int id, int req, void_int *ret_int, float *, char *) {
    int32_t v = 0;
    if (req == EX_INQ_NS_PROP) v = scalar(id, "num_node_sets") > 0 ? 1 : 0;     // the ID property
    else if (req == EX_INQ_SS_PROP) v = scalar(id, "num_side_sets") > 0 ? 1 : 0;
    else v = 0;                                                                  // QA / info records are not carried
    *static_cast<int32_t *>(ret_int) = v;
    return 0;
}
int ex_get_prop_names(int id, ex_entity_type t, char **names) {     // one property ("ID") iff the file has sets of that type
    const int64_t have = scalar(id, t == EX_NODE_SET ? "num_node_sets" : t == EX_SIDE_SET ? "num_side_sets" : "num_el_blk");
    if (have > 0 && names && names[0]) std::strcpy(names[0], "ID");
    return 0;
}
int ex_get_prop_array(int id, ex_entity_type t, const char *, void_int *values) { return copy_i32(id, std::string(prefix_of(t)) + "_ids", values); }
int ex_get_prop(int, ex_entity_type, ex_entity_id sid, const char *, void_int *value) { *static_cast<int32_t *>(value) = (int32_t)sid; return 0; }
int ex_get_side_set_node_list(int, ex_entity_id, void_int *, void_int *) { return 0; }
int ex_get_qa(int, char *[][4]) { return 0; }
int ex_get_info(int, char **) { return 0; }

int ex_put_init(int id, const char *title, int64_t ndim, int64_t nnodes, int64_t nelem, int64_t nblk, int64_t nns, int64_t nss) {
    put(id, "title", "str", title, std::strlen(title));
    put_i32(id, "num_dim", ndim); put_i32(id, "num_nodes", nnodes); put_i32(id, "num_elem", nelem);
    put_i32(id, "num_el_blk", nblk); put_i32(id, "num_node_sets", nns); put_i32(id, "num_side_sets", nss);
    return 0;
}
int ex_put_coord(int id, const void *x, const void *y, const void *z) {
    const size_t n = (size_t)scalar(id, "num_nodes");
    put_real(id, "coordx", x, n); put_real(id, "coordy", y, n);
    if (z) put_real(id, "coordz", z, n);
    return 0;
}
int ex_put_coord_names(int, char **) { return 0; }
int ex_put_map(int id, const void_int *map) { put(id, "elem_map", "i32", map, (size_t)scalar(id, "num_elem")); return 0; }
int ex_put_node_num_map(int id, const void_int *map) { put(id, "node_num_map", "i32", map, (size_t)scalar(id, "num_nodes")); return 0; }
int ex_put_block(int id, ex_entity_type, ex_entity_id bid, const char *elem_type, int64_t nelem, int64_t npe, int64_t, int64_t, int64_t) {
    const std::string b = "eb" + std::to_string(next_index(id, "eb_written"));
    put_i32(id, b + "_id", bid); put(id, b + "_type", "str", elem_type, std::strlen(elem_type));
    put_i32(id, b + "_nelem", nelem); put_i32(id, b + "_npe", npe);
    return 0;
}
int ex_put_conn(int id, ex_entity_type, ex_entity_id bid, const void_int *conn, const void_int *, const void_int *) {
    const int64_t nb = scalar(id, "eb_written");
    for (int64_t k = nb; k >= 1; --k) {
        const std::string b = "eb" + std::to_string(k);
        if (scalar(id, b + "_id", -1) == bid) { put(id, b + "_conn", "i32", conn, (size_t)(scalar(id, b + "_nelem") * scalar(id, b + "_npe"))); return 0; }
    }
    return 1;
}
int ex_put_set_param(int id, ex_entity_type t, ex_entity_id sid, int64_t n, int64_t ndf) {
    const std::string p = prefix_of(t);
    const std::string s = p + std::to_string(next_index(id, p + "_written"));
    put_i32(id, s + "_id", sid); put_i32(id, s + "_n", n); put_i32(id, s + "_ndf", ndf);
    return 0;
}
int ex_put_set(int id, ex_entity_type t, ex_entity_id sid, const void_int *entries, const void_int *extra) {
    const std::string p = prefix_of(t);
    for (int64_t k = scalar(id, p + "_written"); k >= 1; --k) {
        const std::string s = p + std::to_string(k);
        if (scalar(id, s + "_id", -1) != sid) continue;
        put(id, s + "_entries", "i32", entries, (size_t)scalar(id, s + "_n"));
        if (extra) put(id, s + "_extra", "i32", extra, (size_t)scalar(id, s + "_n"));
        return 0;
    }
    return 1;
}
int ex_put_set_dist_fact(int id, ex_entity_type t, ex_entity_id sid, const void *df) {
    const std::string p = prefix_of(t);
    for (int64_t k = scalar(id, p + "_written"); k >= 1; --k) {
        const std::string s = p + std::to_string(k);
        if (scalar(id, s + "_id", -1) == sid) { put_real(id, s + "_df", df, (size_t)scalar(id, s + "_ndf")); return 0; }
    }
    return 1;
}
int ex_put_prop_names(int, ex_entity_type, int, char **) { return 0; }
int ex_put_prop_array(int, ex_entity_type, const char *, const void_int *) { return 0; }
int ex_put_prop(int, ex_entity_type, ex_entity_id, const char *, ex_entity_id) { return 0; }
int ex_put_qa(int, int, char *[][4]) { return 0; }
int ex_put_info(int, int, char **) { return 0; }
int ex_put_variable_param(int id, ex_entity_type, int n) { put_i32(id, "num_nod_var", n); return 0; }
int ex_put_variable_names(int id, ex_entity_type, int n, char **names) {
    for (int i = 0; i < n; ++i) put(id, "name_nod_var" + std::to_string(i + 1), "str", names[i], std::strlen(names[i]));
    return 0;
}
int ex_put_time(int id, int step, const void *t) {
    put_real(id, "time" + std::to_string(step), t, 1);
    if (step > scalar(id, "num_time_steps")) put_i32(id, "num_time_steps", step);
    return 0;
}
int ex_put_nodal_var(int id, int step, int var, int64_t n, const void *vals) {
    put_real(id, "vals_nod_var" + std::to_string(var) + "_step" + std::to_string(step), vals, (size_t)n);
    return 0;
}

}  // extern "C"
