// exodus.cpp — netCDF-classic reader/writer + the Exodus-II subset of the heat path, and the C-ABI
// entry points that own files: heat_open / heat_create / heat_decompose / heat_write_solution /
// heat_nodal_field (IO::open, IO::create, IO::decompose, IO::writeSolution of ExodusIO.hpp).
#include "exodus.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "comm.cuh"
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// big-endian helpers
// ------------------------------------------------------------------------------------------------
namespace {

struct Reader {
    const uint8_t *p; size_t n, pos = 0; bool ok = true; int version = 2;
    uint32_t u32() {
        if (pos + 4 > n) { ok = false; return 0; }
        uint32_t v = ((uint32_t)p[pos] << 24) | ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3];
        pos += 4; return v;
    }
    uint64_t u64() { uint64_t hi = u32(), lo = u32(); return (hi << 32) | lo; }
    int64_t count() { return version == 5 ? (int64_t)u64() : (int64_t)u32(); }     // NON_NEG
    std::string name() {
        int64_t len = count();
        if (!ok || len < 0 || pos + (size_t)len > n) { ok = false; return ""; }
        std::string s((const char *)p + pos, (size_t)len);
        pos += ((size_t)len + 3) & ~(size_t)3;
        return s;
    }
};

struct Writer {
    std::vector<uint8_t> b;
    void u32(uint32_t v) { b.push_back(v >> 24); b.push_back(v >> 16); b.push_back(v >> 8); b.push_back(v); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void name(const std::string &s) {
        u32((uint32_t)s.size());
        b.insert(b.end(), s.begin(), s.end());
        while (b.size() & 3) b.push_back(0);
    }
    void bytes(const std::vector<uint8_t> &r) { b.insert(b.end(), r.begin(), r.end()); while (b.size() & 3) b.push_back(0); }
};

void swap_copy(uint8_t *dst, const void *src, size_t n, int sz) {      // host (little-endian) <-> big-endian
    const uint8_t *s = (const uint8_t *)src;
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < sz; ++k) dst[i * sz + k] = s[i * sz + (sz - 1 - k)];
}

int read_atts(Reader &r, std::vector<NcAtt> &atts) {
    uint32_t tag = r.u32();
    int64_t n = r.count();
    if (!r.ok) return 1;
    if (tag == 0 && n == 0) return 0;
    if (tag != 0x0C) return 1;
    for (int64_t i = 0; i < n; ++i) {
        NcAtt a;
        a.name = r.name();
        a.type = (int)r.u32();
        a.nelems = r.count();
        const size_t esz = (size_t)nc_type_size(a.type);
        // overflow-safe: nelems * esz and pos + bytes are file-controlled (CDF-5 counts are 64-bit)
        if (!r.ok || a.nelems < 0 || esz == 0 || r.pos > r.n || (uint64_t)a.nelems > (uint64_t)(r.n - r.pos) / esz) return 1;
        size_t bytes = (size_t)a.nelems * esz;
        a.raw.assign(r.p + r.pos, r.p + r.pos + bytes);
        r.pos += (bytes + 3) & ~(size_t)3;
        atts.push_back(std::move(a));
    }
    return 0;
}

void write_atts(Writer &w, const std::vector<NcAtt> &atts) {
    if (atts.empty()) { w.u32(0); w.u32(0); return; }
    w.u32(0x0C); w.u32((uint32_t)atts.size());
    for (const NcAtt &a : atts) { w.name(a.name); w.u32((uint32_t)a.type); w.u32((uint32_t)a.nelems); w.bytes(a.raw); }
}

}  // namespace

int nc_type_size(int t) {
    switch (t) {
        case NC_BYTE: case NC_CHAR: case NC_UBYTE: return 1;
        case NC_SHORT: case NC_USHORT: return 2;
        case NC_INT: case NC_FLOAT: case NC_UINT: return 4;
        default: return 8;
    }
}

int NcFile::dim_id(const std::string &n) const {
    for (size_t i = 0; i < dims.size(); ++i) if (dims[i].name == n) return (int)i;
    return -1;
}
int64_t NcFile::dim_len(const std::string &n, int64_t dflt) const {
    int id = dim_id(n);
    if (id < 0) return dflt;
    return dims[(size_t)id].unlimited ? numrecs : dims[(size_t)id].len;
}
const NcVar *NcFile::var(const std::string &n) const {
    for (const NcVar &v : vars) if (v.name == n) return &v;
    return nullptr;
}
NcVar *NcFile::var(const std::string &n) {
    for (NcVar &v : vars) if (v.name == n) return &v;
    return nullptr;
}
const NcAtt *NcFile::gatt(const std::string &n) const {
    for (const NcAtt &a : gatts) if (a.name == n) return &a;
    return nullptr;
}
int NcFile::add_dim(const std::string &n, int64_t len, bool unlimited) {
    int id = dim_id(n);
    if (id >= 0) { dims[(size_t)id].len = len; dims[(size_t)id].unlimited = unlimited; return id; }
    NcDim d; d.name = n; d.len = len; d.unlimited = unlimited;
    dims.push_back(d);
    return (int)dims.size() - 1;
}
void NcFile::set_dim(const std::string &n, int64_t len) { add_dim(n, len, false); }
NcVar &NcFile::add_var(const std::string &n, int type, const std::vector<std::string> &dim_names) {
    remove_var(n);
    NcVar v; v.name = n; v.type = type;
    for (const std::string &d : dim_names) {
        int id = dim_id(d);
        v.dimids.push_back(id);
        if (id >= 0 && dims[(size_t)id].unlimited) v.is_record = true;
    }
    vars.push_back(std::move(v));
    return vars.back();
}
void NcFile::remove_var(const std::string &n) {
    vars.erase(std::remove_if(vars.begin(), vars.end(), [&](const NcVar &v) { return v.name == n; }), vars.end());
}
int64_t NcFile::var_elems_per_record(const NcVar &v) const {
    int64_t e = 1;
    for (size_t i = 0; i < v.dimids.size(); ++i) {
        const NcDim &d = dims[(size_t)v.dimids[i]];
        if (d.unlimited) continue;
        e *= d.len;
    }
    return e;
}
std::vector<double> NcFile::get_doubles(const NcVar &v) const {
    const int sz = nc_type_size(v.type);
    const size_t n = v.raw.size() / (size_t)sz;
    std::vector<double> out(n);
    for (size_t i = 0; i < n; ++i) {
        uint8_t tmp[8];
        for (int k = 0; k < sz; ++k) tmp[k] = v.raw[i * sz + (sz - 1 - k)];
        if (v.type == NC_DOUBLE) { double d; memcpy(&d, tmp, 8); out[i] = d; }
        else if (v.type == NC_FLOAT) { float f; memcpy(&f, tmp, 4); out[i] = f; }
        else if (v.type == NC_INT) { int32_t q; memcpy(&q, tmp, 4); out[i] = q; }
        else if (v.type == NC_SHORT) { int16_t q; memcpy(&q, tmp, 2); out[i] = q; }
        else if (v.type == NC_INT64) { int64_t q; memcpy(&q, tmp, 8); out[i] = (double)q; }
        else out[i] = (double)(int8_t)tmp[0];
    }
    return out;
}
std::vector<int64_t> NcFile::get_ints(const NcVar &v) const {
    const int sz = nc_type_size(v.type);
    const size_t n = v.raw.size() / (size_t)sz;
    std::vector<int64_t> out(n);
    for (size_t i = 0; i < n; ++i) {
        uint8_t tmp[8];
        for (int k = 0; k < sz; ++k) tmp[k] = v.raw[i * sz + (sz - 1 - k)];
        if (v.type == NC_INT) { int32_t q; memcpy(&q, tmp, 4); out[i] = q; }
        else if (v.type == NC_INT64) { int64_t q; memcpy(&q, tmp, 8); out[i] = q; }
        else if (v.type == NC_SHORT) { int16_t q; memcpy(&q, tmp, 2); out[i] = q; }
        else if (v.type == NC_DOUBLE) { double d; memcpy(&d, tmp, 8); out[i] = (int64_t)d; }
        else if (v.type == NC_FLOAT) { float f; memcpy(&f, tmp, 4); out[i] = (int64_t)f; }
        else out[i] = (int8_t)tmp[0];
    }
    return out;
}
std::string NcFile::get_att_string(const std::vector<NcAtt> &atts, const std::string &n) const {
    for (const NcAtt &a : atts)
        if (a.name == n && a.type == NC_CHAR) {
            std::string s((const char *)a.raw.data(), a.raw.size());
            size_t z = s.find('\0');
            if (z != std::string::npos) s.resize(z);
            return s;
        }
    return "";
}
void NcFile::put_doubles(NcVar &v, const double *p, size_t n) { v.type = NC_DOUBLE; v.raw.resize(n * 8); swap_copy(v.raw.data(), p, n, 8); }
void NcFile::put_floats(NcVar &v, const double *p, size_t n) {
    v.type = NC_FLOAT;
    v.raw.resize(n * 4);
    for (size_t i = 0; i < n; ++i) { const float f = (float)p[i]; swap_copy(v.raw.data() + 4 * i, &f, 1, 4); }
}
void NcFile::put_ints(NcVar &v, const int32_t *p, size_t n) { v.type = NC_INT; v.raw.resize(n * 4); swap_copy(v.raw.data(), p, n, 4); }
void NcFile::put_chars(NcVar &v, const char *p, size_t n) { v.type = NC_CHAR; v.raw.assign((const uint8_t *)p, (const uint8_t *)p + n); }
NcAtt NcFile::make_att_string(const std::string &n, const std::string &val) {
    NcAtt a; a.name = n; a.type = NC_CHAR; a.nelems = (int64_t)val.size(); a.raw.assign(val.begin(), val.end()); return a;
}
NcAtt NcFile::make_att_int(const std::string &n, int32_t val) {
    NcAtt a; a.name = n; a.type = NC_INT; a.nelems = 1; a.raw.resize(4); swap_copy(a.raw.data(), &val, 1, 4); return a;
}
NcAtt NcFile::make_att_float(const std::string &n, float val) {
    NcAtt a; a.name = n; a.type = NC_FLOAT; a.nelems = 1; a.raw.resize(4); swap_copy(a.raw.data(), &val, 1, 4); return a;
}

int nc_read(const std::string &path, NcFile &out) {
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) HEAT_FAIL(60, "ex_open: cannot open '%s'", path.c_str());
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> buf((size_t)(fsz > 0 ? fsz : 0));
    size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), fp);
    fclose(fp);
    if (got != buf.size() || buf.size() < 8) HEAT_FAIL(60, "ex_open: short read on '%s'", path.c_str());
    if (buf[0] != 'C' || buf[1] != 'D' || buf[2] != 'F' || (buf[3] != 1 && buf[3] != 2 && buf[3] != 5)) {
        if (buf[0] == 0x89 && buf[1] == 'H' && buf[2] == 'D' && buf[3] == 'F')
            HEAT_FAIL(61, "ex_open: '%s' is netCDF-4/HDF5; only netCDF classic (CDF-1/2/5) is supported", path.c_str());
        HEAT_FAIL(61, "ex_open: '%s' is not a netCDF classic file", path.c_str());
    }
    Reader r{buf.data(), buf.size()};
    r.pos = 4; r.version = buf[3];
    out = NcFile();
    out.version = r.version;
    int64_t numrecs = r.count();
    bool streaming = (r.version == 5) ? (numrecs == -1) : (numrecs == 0xFFFFFFFFll);
    // dims
    {
        uint32_t tag = r.u32();
        int64_t n = r.count();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0A) HEAT_FAIL(62, "netCDF header: bad dim_list tag in '%s'", path.c_str());
            for (int64_t i = 0; i < n; ++i) {
                NcDim d; d.name = r.name(); d.len = r.count(); d.unlimited = (d.len == 0);
                out.dims.push_back(d);
            }
        }
    }
    if (read_atts(r, out.gatts)) HEAT_FAIL(62, "netCDF header: bad global attributes in '%s'", path.c_str());
    struct Loc { int64_t vsize, begin; };
    std::vector<Loc> locs;
    {
        uint32_t tag = r.u32();
        int64_t n = r.count();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0B) HEAT_FAIL(62, "netCDF header: bad var_list tag in '%s'", path.c_str());
            for (int64_t i = 0; i < n; ++i) {
                NcVar v; v.name = r.name();
                int64_t nd = r.count();
                for (int64_t k = 0; k < nd; ++k) {
                    int id = (int)r.count();
                    if (id < 0 || id >= (int)out.dims.size()) HEAT_FAIL(62, "netCDF header: bad dimid in '%s'", path.c_str());
                    v.dimids.push_back(id);
                }
                if (read_atts(r, v.atts)) HEAT_FAIL(62, "netCDF header: bad attributes of '%s'", v.name.c_str());
                v.type = (int)r.u32();
                Loc l;
                l.vsize = r.count();
                l.begin = (r.version == 1) ? (int64_t)r.u32() : (int64_t)r.u64();
                v.is_record = !v.dimids.empty() && out.dims[(size_t)v.dimids[0]].unlimited;
                out.vars.push_back(std::move(v));
                locs.push_back(l);
            }
        }
    }
    if (!r.ok) HEAT_FAIL(62, "netCDF header truncated in '%s'", path.c_str());
    // record size
    int64_t recsize = 0; int nrecvars = 0;
    for (size_t i = 0; i < out.vars.size(); ++i)
        if (out.vars[i].is_record) { recsize += locs[i].vsize; ++nrecvars; }
    if (nrecvars == 1)      // a lone record variable is not padded
        for (size_t i = 0; i < out.vars.size(); ++i)
            if (out.vars[i].is_record) recsize = out.var_elems_per_record(out.vars[i]) * nc_type_size(out.vars[i].type);
    if (streaming) {
        int64_t first = -1;
        for (size_t i = 0; i < out.vars.size(); ++i)
            if (out.vars[i].is_record && (first < 0 || locs[i].begin < first)) first = locs[i].begin;
        numrecs = (first >= 0 && recsize > 0) ? ((int64_t)buf.size() - first) / recsize : 0;
    }
    out.numrecs = numrecs;
    for (size_t i = 0; i < out.vars.size(); ++i) {
        NcVar &v = out.vars[i];
        const int64_t per = out.var_elems_per_record(v) * nc_type_size(v.type);
        if (!v.is_record) {
            if (locs[i].begin < 0 || (size_t)(locs[i].begin + per) > buf.size()) HEAT_FAIL(63, "variable '%s' runs past the end of '%s'", v.name.c_str(), path.c_str());
            v.raw.assign(buf.begin() + locs[i].begin, buf.begin() + locs[i].begin + per);
        } else {
            v.raw.resize((size_t)(per * numrecs));
            for (int64_t rr = 0; rr < numrecs; ++rr) {
                const int64_t off = locs[i].begin + rr * recsize;
                if (off < 0 || (size_t)(off + per) > buf.size()) HEAT_FAIL(63, "record variable '%s' runs past the end of '%s'", v.name.c_str(), path.c_str());
                memcpy(v.raw.data() + rr * per, buf.data() + off, (size_t)per);
            }
        }
    }
    return 0;
}

// file layout of a CDF-2 image of `f`: header bytes, fixed-size variables, then the record section
namespace {
struct NcLayout {
    std::vector<int64_t> vsize, begin;
    int64_t rec_start = 0, recsize = 0;
    std::vector<uint8_t> header;
};
void nc_layout(const NcFile &f, NcLayout &L) {
    auto padded = [](int64_t b) { return (b + 3) & ~(int64_t)3; };
    L.vsize.assign(f.vars.size(), 0); L.begin.assign(f.vars.size(), 0);
    int nrecvars = 0;
    for (size_t i = 0; i < f.vars.size(); ++i) {
        L.vsize[i] = padded(f.var_elems_per_record(f.vars[i]) * nc_type_size(f.vars[i].type));
        nrecvars += f.vars[i].is_record;
    }
    auto header = [&](Writer &w) {
        w.b.push_back('C'); w.b.push_back('D'); w.b.push_back('F'); w.b.push_back(2);
        w.u32((uint32_t)f.numrecs);
        if (f.dims.empty()) { w.u32(0); w.u32(0); }
        else {
            w.u32(0x0A); w.u32((uint32_t)f.dims.size());
            for (const NcDim &d : f.dims) { w.name(d.name); w.u32(d.unlimited ? 0u : (uint32_t)d.len); }
        }
        write_atts(w, f.gatts);
        if (f.vars.empty()) { w.u32(0); w.u32(0); }
        else {
            w.u32(0x0B); w.u32((uint32_t)f.vars.size());
            for (size_t i = 0; i < f.vars.size(); ++i) {
                const NcVar &v = f.vars[i];
                w.name(v.name);
                w.u32((uint32_t)v.dimids.size());
                for (int id : v.dimids) w.u32((uint32_t)id);
                write_atts(w, v.atts);
                w.u32((uint32_t)v.type);
                w.u32((uint32_t)(L.vsize[i] > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)L.vsize[i]));
                w.u64((uint64_t)L.begin[i]);
            }
        }
    };
    Writer probe;
    header(probe);                                       // offsets are fixed-width: size is final
    int64_t off = (int64_t)probe.b.size();
    for (size_t i = 0; i < f.vars.size(); ++i)
        if (!f.vars[i].is_record) { L.begin[i] = off; off += L.vsize[i]; }
    L.rec_start = off; L.recsize = 0;
    for (size_t i = 0; i < f.vars.size(); ++i)
        if (f.vars[i].is_record) { L.begin[i] = off; off += L.vsize[i]; L.recsize += L.vsize[i]; }
    if (nrecvars == 1)                                   // a lone record variable is not padded
        for (size_t i = 0; i < f.vars.size(); ++i)
            if (f.vars[i].is_record) L.recsize = f.var_elems_per_record(f.vars[i]) * nc_type_size(f.vars[i].type);
    Writer w;
    header(w);
    L.header.swap(w.b);
}
}  // namespace

int nc_write(const std::string &path, const NcFile &f) {
    // always CDF-2 (64-bit offset), like every file under the reference's data/
    NcLayout L;
    nc_layout(f, L);
    std::vector<uint8_t> b(L.header);
    b.resize((size_t)(L.rec_start + L.recsize * f.numrecs), 0);
    for (size_t i = 0; i < f.vars.size(); ++i) {
        const NcVar &v = f.vars[i];
        const int64_t per = f.var_elems_per_record(v) * nc_type_size(v.type);
        if (!v.is_record) {
            memcpy(b.data() + L.begin[i], v.raw.data(), std::min<size_t>(v.raw.size(), (size_t)per));
        } else {
            for (int64_t r = 0; r < f.numrecs; ++r) {
                if ((size_t)((r + 1) * per) > v.raw.size()) break;
                memcpy(b.data() + L.begin[i] + r * L.recsize, v.raw.data() + r * per, (size_t)per);
            }
        }
    }
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) HEAT_FAIL(64, "ex_create: cannot open '%s' for writing", path.c_str());
    size_t put = fwrite(b.data(), 1, b.size(), fp);
    fclose(fp);
    if (put != b.size()) HEAT_FAIL(64, "short write on '%s'", path.c_str());
    return 0;
}

// Rewrites records [r0, r1) of every record variable and the record count IN PLACE.  The file on disk
// must have been written by nc_write() from the same dimension / variable set (the header has the same
// size and offsets; only numrecs changes), which is what a sequence of ex_put_var calls amounts to.
int nc_update_records(const std::string &path, const NcFile &f, int64_t r0, int64_t r1) {
    NcLayout L;
    nc_layout(f, L);
    FILE *fp = fopen(path.c_str(), "r+b");
    if (!fp) HEAT_FAIL(64, "ex_put_var: cannot reopen '%s'", path.c_str());
    bool ok = fseek(fp, 0, SEEK_SET) == 0 && fwrite(L.header.data(), 1, 8, fp) == 8;      // magic + numrecs
    std::vector<uint8_t> rec((size_t)L.recsize);
    for (int64_t r = r0; ok && r < r1 && r < f.numrecs; ++r) {
        std::fill(rec.begin(), rec.end(), 0);
        for (size_t i = 0; i < f.vars.size(); ++i) {
            const NcVar &v = f.vars[i];
            if (!v.is_record) continue;
            const int64_t per = f.var_elems_per_record(v) * nc_type_size(v.type);
            if ((size_t)((r + 1) * per) <= v.raw.size())
                memcpy(rec.data() + (L.begin[i] - L.rec_start), v.raw.data() + r * per, (size_t)per);
        }
        ok = fseek(fp, (long)(L.rec_start + r * L.recsize), SEEK_SET) == 0 && fwrite(rec.data(), 1, rec.size(), fp) == rec.size();
    }
    fclose(fp);
    if (!ok) HEAT_FAIL(64, "short write on '%s'", path.c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Exodus subset
// ------------------------------------------------------------------------------------------------
void exo_close(ExoFile *f) { delete f; }

int exo_read_mesh(const ExoFile &f, HostMesh &m) {
    const NcFile &nc = f.nc;
    m = HostMesh();
    m.num_nodes = nc.dim_len("num_nodes");
    m.num_dim = (int)nc.dim_len("num_dim", 3);
    m.num_elem = nc.dim_len("num_elem");
    const int64_t nblk = nc.dim_len("num_el_blk");
    if (const NcVar *cx = nc.var("coordx")) {
        m.x = nc.get_doubles(*cx);
        if (const NcVar *cy = nc.var("coordy")) m.y = nc.get_doubles(*cy); else m.y.assign((size_t)m.num_nodes, 0.0);
        if (const NcVar *cz = nc.var("coordz")) m.z = nc.get_doubles(*cz);
    } else if (const NcVar *c = nc.var("coord")) {
        std::vector<double> all = nc.get_doubles(*c);
        const size_t N = (size_t)m.num_nodes;
        if (m.num_dim < 1 || m.num_dim > 3 || all.size() < N * (size_t)m.num_dim)
            HEAT_FAIL(65, "'%s': coord holds %zu values, %d x %zu expected", f.path.c_str(), all.size(), m.num_dim, N);
        m.x.assign(all.begin(), all.begin() + N);
        if (m.num_dim > 1) m.y.assign(all.begin() + N, all.begin() + 2 * N); else m.y.assign(N, 0.0);
        if (m.num_dim > 2) m.z.assign(all.begin() + 2 * N, all.begin() + 3 * N);
    } else if (m.num_nodes > 0) {
        HEAT_FAIL(65, "'%s' has no coordinates", f.path.c_str());
    }
    if (m.num_nodes > 0 && ((int64_t)m.x.size() < m.num_nodes || (int64_t)m.y.size() < m.num_nodes ||
                            (!m.z.empty() && (int64_t)m.z.size() < m.num_nodes)))
        HEAT_FAIL(65, "'%s': coordinate arrays shorter than num_nodes = %lld", f.path.c_str(), (long long)m.num_nodes);
    m.npe = 0;
    for (int64_t b = 1; b <= nblk; ++b) {
        const NcVar *cv = nc.var("connect" + std::to_string(b));
        if (!cv) continue;                                   // NULL block (eb_status == 0)
        if (cv->dimids.size() < 2) HEAT_FAIL(66, "'%s': connect%lld is not a [elements][nodes] table", f.path.c_str(), (long long)b);
        const int npe_b = (int)nc.dims[(size_t)cv->dimids[1]].len;
        if (m.npe == 0) m.npe = npe_b;
        if (npe_b != m.npe) HEAT_FAIL(66, "'%s': element blocks with %d and %d nodes per element; mixed meshes are not supported", f.path.c_str(), m.npe, npe_b);
        std::vector<int64_t> c = nc.get_ints(*cv);
        for (int64_t v : c) m.conn.push_back((int32_t)(v - 1));
        m.elem_type = nc.get_att_string(cv->atts, "elem_type");    // last block wins (ExodusIO.hpp:1604, D11)
    }
    if ((int64_t)m.conn.size() != m.num_elem * (int64_t)m.npe) HEAT_FAIL(66, "'%s': connectivity size mismatch", f.path.c_str());
    for (int32_t v : m.conn)
        if (v < 0 || v >= m.num_nodes) HEAT_FAIL(66, "'%s': connectivity entry out of range", f.path.c_str());
    const int64_t nns = nc.dim_len("num_node_sets");
    if (nns > 0) {
        const NcVar *ids = nc.var("ns_prop1");
        if (!ids) HEAT_FAIL(67, "'%s': nodesets without ns_prop1", f.path.c_str());
        std::vector<int64_t> idv = nc.get_ints(*ids);
        if ((int64_t)idv.size() < nns) HEAT_FAIL(67, "'%s': ns_prop1 holds %zu ids for %lld nodesets", f.path.c_str(), idv.size(), (long long)nns);
        for (int64_t s = 1; s <= nns; ++s) {
            std::vector<int64_t> &dst = m.nodesets[idv[(size_t)s - 1]];     // ExodusIO.hpp:176-191
            if (const NcVar *nv = nc.var("node_ns" + std::to_string(s)))
                for (int64_t g : nc.get_ints(*nv)) {
                    if (g < 1 || g > m.num_nodes) HEAT_FAIL(67, "'%s': nodeset node out of range", f.path.c_str());
                    dst.push_back(g - 1);
                }
        }
    }
    m.valid = true;
    return 0;
}

namespace heat {
int metis_part_mesh_dual(int64_t ne, int64_t nn, int npe, const int32_t *conn, int64_t ncommon, int64_t nparts,
                         int64_t *objval, int64_t *epart, int64_t *npart);
}

extern "C" int heat_open(heat_ctx *ctx, const char *path, int read_only) {
    if (!ctx || !path) HEAT_FAIL(2, "heat_open: null argument");
    (void)read_only;                                       // EX_READ / EX_WRITE: the input is never modified here
    ExoFile *f = new ExoFile();
    f->path = path;
    int rc = nc_read(path, f->nc);
    if (!rc) rc = exo_read_mesh(*f, ctx->mesh);
    if (rc) { delete f; return rc; }
    exo_close(ctx->read_file);
    ctx->read_file = f;
    return 0;
}

extern "C" int heat_create(heat_ctx *ctx, const char *path) {
    if (!ctx || !path) HEAT_FAIL(2, "heat_create: null argument");
    FILE *fp = fopen(path, "wb");                          // EX_CLOBBER
    if (!fp) HEAT_FAIL(64, "ex_create: cannot create '%s'", path);
    fclose(fp);
    exo_close(ctx->write_file);
    ctx->write_file = new ExoFile();
    ctx->write_file->path = path;
    ctx->write_file->writable = true;
    ctx->printed_time_zero = false;
    return 0;
}

static int ncommon_for(const std::string &et, int64_t *ncommon) {
    if (et.compare(0, 5, "TETRA") == 0) *ncommon = 3;            // ExodusIO.hpp:1604-1613
    else if (et.compare(0, 3, "TRI") == 0) *ncommon = 2;
    else if (et.compare(0, 3, "HEX") == 0) *ncommon = 4;
    else if (et.compare(0, 3, "TET") == 0) *ncommon = 3;         // "TET4" spelling (not accepted by the reference)
    else HEAT_FAIL(68, "Currently unsupported element type for mesh: %s", et.c_str());
    return 0;
}

extern "C" int heat_decompose_partition(heat_ctx *ctx, int partitions, int64_t *objval, int64_t *epart, int64_t *npart) {
    if (!ctx || !ctx->mesh.valid || ctx->mesh.is_cube) HEAT_FAIL(4, "heat_decompose: needs a mesh from heat_open/heat_mesh_set");
    if (partitions < 1 || !epart || !npart) HEAT_FAIL(2, "heat_decompose: bad arguments");
    const HostMesh &m = ctx->mesh;
    int64_t ncommon = 1, obj = 0;
    HEAT_TRY(ncommon_for(m.elem_type, &ncommon));
    std::fill(epart, epart + m.num_elem, 0);
    std::fill(npart, npart + m.num_nodes, 0);
    if (partitions > 1)
        HEAT_TRY(heat::metis_part_mesh_dual(m.num_elem, m.num_nodes, m.npe, m.conn.data(), ncommon, partitions, &obj, epart, npart));
    if (objval) *objval = obj;
    return 0;
}

static std::vector<std::string> dim_names_of(const NcFile &nc, const NcVar &v) {
    std::vector<std::string> out;
    for (int id : v.dimids) out.push_back(nc.dims[(size_t)id].name);
    return out;
}

// IO::decompose (ExodusIO.hpp:1496-1969): METIS dual-graph partition of the elements, then a copy of
// the mesh into the output file with ONE ELEMENT BLOCK PER (non-empty) PARTITION, block ids from 0.
extern "C" int heat_decompose(heat_ctx *ctx, int partitions) {
    if (!ctx || !ctx->read_file) HEAT_FAIL(4, "heat_decompose: no input file (readFID == -1)");
    if (!ctx->write_file) HEAT_FAIL(4, "heat_decompose: call heat_create first");
    const HostMesh &m = ctx->mesh;
    std::vector<int64_t> epart((size_t)m.num_elem), npart((size_t)m.num_nodes);
    int64_t obj = 0;
    HEAT_TRY(heat_decompose_partition(ctx, partitions, &obj, epart.data(), npart.data()));
    std::vector<std::vector<int64_t>> bin((size_t)partitions);
    for (int64_t e = 0; e < m.num_elem; ++e) bin[(size_t)epart[(size_t)e]].push_back(e);     // :1639-1665
    int64_t numparts = 0;
    for (auto &b : bin) numparts += !b.empty();

    const NcFile &in = ctx->read_file->nc;
    NcFile out;
    out.version = 2;
    out.gatts = in.gatts;
    // dimensions: everything that is not per-block; per-block ones are rebuilt
    for (const NcDim &d : in.dims) {
        if (d.name.compare(0, 13, "num_el_in_blk") == 0 || d.name.compare(0, 14, "num_nod_per_el") == 0 ||
            d.name.compare(0, 14, "num_att_in_blk") == 0 || d.name.compare(0, 14, "num_edg_per_el") == 0 ||
            d.name.compare(0, 14, "num_fac_per_el") == 0 || d.name == "num_nod_var" || d.name == "num_elem_var" ||
            d.name == "num_glo_var")
            continue;
        out.add_dim(d.name, d.name == "num_el_blk" ? numparts : d.len, d.unlimited);
    }
    if (out.dim_id("time_step") < 0) out.add_dim("time_step", 0, true);
    if (out.dim_id("num_el_blk") < 0) out.add_dim("num_el_blk", numparts);
    if (out.dim_id("len_name") < 0) out.add_dim("len_name", 33);
    auto copy_var = [&](const std::string &name) {
        const NcVar *v = in.var(name);
        if (!v) return;
        NcVar &nv = out.add_var(name, v->type, dim_names_of(in, *v));
        nv.atts = v->atts; nv.raw = v->raw;
        if (nv.is_record) nv.raw.clear();
    };
    copy_var("time_whole");
    if (!out.var("time_whole")) out.add_var("time_whole", NC_DOUBLE, {"time_step"});
    for (const char *n : {"coordx", "coordy", "coordz", "coord", "coor_names", "elem_map", "ns_status", "ns_names",
                          "ss_status", "ss_names", "qa_records", "info_records", "node_num_map"})
        copy_var(n);
    // ex_get_map / ex_get_node_num_map hand out the identity 1..N when the file stores no map, and the reference
    // writes whatever it got (ExodusIO.hpp:1742-1745, :1962-1966): the output always carries both maps
    auto identity_map = [&](const char *name, const char *dim, int64_t len) {
        if (out.var(name) || len <= 0 || out.dim_id(dim) < 0) return;
        std::vector<int32_t> ids((size_t)len);
        for (int64_t i = 0; i < len; ++i) ids[(size_t)i] = (int32_t)(i + 1);
        NcVar &v = out.add_var(name, NC_INT, {dim});
        NcFile::put_ints(v, ids.data(), ids.size());
    };
    identity_map("elem_map", "num_elem", m.num_elem);
    identity_map("node_num_map", "num_nodes", m.num_nodes);
    for (const NcVar &v : in.vars) {
        const std::string &n = v.name;
        if (n.compare(0, 7, "ns_prop") == 0 || n.compare(0, 7, "ss_prop") == 0 || n.compare(0, 7, "node_ns") == 0 ||
            n.compare(0, 12, "dist_fact_ns") == 0 || n.compare(0, 7, "elem_ss") == 0 || n.compare(0, 7, "side_ss") == 0 ||
            n.compare(0, 12, "dist_fact_ss") == 0)
            copy_var(n);
    }
    // element blocks: one per non-empty partition, id = running index from 0 (:1768, :1785, D12)
    {
        std::vector<int32_t> status((size_t)numparts, 1), ids((size_t)numparts);
        for (int64_t p = 0; p < numparts; ++p) ids[(size_t)p] = (int32_t)p;
        NcVar &st = out.add_var("eb_status", NC_INT, {"num_el_blk"});
        NcFile::put_ints(st, status.data(), status.size());
        NcVar &pr = out.add_var("eb_prop1", NC_INT, {"num_el_blk"});
        NcFile::put_ints(pr, ids.data(), ids.size());
        pr.atts.push_back(NcFile::make_att_string("name", "ID"));
        const int64_t len_name = out.dim_len("len_name", 33);
        std::vector<char> names((size_t)(numparts * len_name), 0);
        NcVar &nm = out.add_var("eb_names", NC_CHAR, {"num_el_blk", "len_name"});
        NcFile::put_chars(nm, names.data(), names.size());
    }
    int64_t blk = 0;
    for (int p = 0; p < partitions; ++p) {
        if (bin[(size_t)p].empty()) continue;
        ++blk;
        const std::string sb = std::to_string(blk);
        out.add_dim("num_el_in_blk" + sb, (int64_t)bin[(size_t)p].size());
        out.add_dim("num_nod_per_el" + sb, m.npe);
        std::vector<int32_t> conn;
        conn.reserve(bin[(size_t)p].size() * (size_t)m.npe);
        for (int64_t e : bin[(size_t)p])
            for (int k = 0; k < m.npe; ++k) conn.push_back(m.conn[(size_t)(e * m.npe + k)] + 1);    // :1776
        NcVar &cv = out.add_var("connect" + sb, NC_INT, {"num_el_in_blk" + sb, "num_nod_per_el" + sb});
        NcFile::put_ints(cv, conn.data(), conn.size());
        cv.atts.push_back(NcFile::make_att_string("elem_type", m.elem_type));
    }
    // The reference creates its output with cpu/io word size sizeof(real_t) whatever the input holds (ExodusIO.hpp:
    // 104-105): with a METIS whose real_t is float every floating-point record (coordinates, distribution factors,
    // times) is float32.  Here the word size is the context's (8 unless heat_ctx_set_output asked for 4), and every
    // floating-point variable copied from the input is brought to it.
    for (NcVar &v : out.vars) {
        if (v.type != NC_DOUBLE && v.type != NC_FLOAT) continue;
        if (nc_type_size(v.type) == ctx->out_word_size) continue;
        const std::vector<double> d = out.get_doubles(v);
        if (ctx->out_word_size == 4) NcFile::put_floats(v, d.data(), d.size());
        else NcFile::put_doubles(v, d.data(), d.size());
    }
    for (NcAtt &a : out.gatts)
        if (a.name == "floating_point_word_size") a = NcFile::make_att_int("floating_point_word_size", ctx->out_word_size);
    if (!out.gatt("floating_point_word_size")) out.gatts.push_back(NcFile::make_att_int("floating_point_word_size", ctx->out_word_size));
    out.numrecs = 0;
    ctx->write_file->nc = std::move(out);
    ctx->write_file->have_results = false;
    return nc_write(ctx->write_file->path, ctx->write_file->nc);
}

// The scatter half of writeSolution (ExodusIO.hpp:1981-1989 + :2045-2055): solution in reduced-id order -> dense nodal
// array, nodeset nodes = their id.  Pure host code.
static int scatter_reduced(heat_ctx *ctx, const double *xg, int64_t n, double *field) {
    const HostMesh &m = ctx->mesh;
    if (m.is_cube) {
        if (n != (int64_t)(m.nx - 2) * m.ny * m.nz) HEAT_FAIL(2, "nodal field: the mesh has %lld unknowns, the vector %lld", (long long)((int64_t)(m.nx - 2) * m.ny * m.nz), (long long)n);
        int64_t r = 0;
        for (int64_t g = 0; g < m.num_nodes; ++g) {
            const int i = (int)(g % m.nx);
            field[g] = (i == 0) ? 1000.0 : (i == m.nx - 1) ? 100.0 : xg[(size_t)r++];
        }
        return 0;
    }
    if ((int64_t)ctx->node_bc.size() != m.num_nodes) heat::build_node_bc(ctx);
    int64_t ndof = 0;
    for (int64_t g = 0; g < m.num_nodes; ++g) ndof += std::isnan(ctx->node_bc[(size_t)g]) ? 1 : 0;
    if (ndof != n) HEAT_FAIL(2, "nodal field: the mesh has %lld unknowns, the vector %lld", (long long)ndof, (long long)n);
    int64_t r = 0;
    for (int64_t g = 0; g < m.num_nodes; ++g) {
        const double bc = ctx->node_bc[(size_t)g];
        // nodeset node = its id: the one the RHS used (lowest, D2 fixed) or, on request, the largest as the reference writes
        field[g] = std::isnan(bc) ? xg[(size_t)r++] : (ctx->out_largest_id ? ctx->node_bc_hi[(size_t)g] : bc);
    }
    return 0;
}

extern "C" int heat_scatter_nodal_field(heat_ctx *ctx, const double *x_reduced_host, int64_t n_global, double *field_host,
                                        int64_t num_nodes) {
    if (!ctx || !x_reduced_host || !field_host) HEAT_FAIL(2, "heat_scatter_nodal_field: null argument");
    const HostMesh &m = ctx->mesh;
    if (!m.valid) HEAT_FAIL(4, "heat_scatter_nodal_field: no mesh");
    if (num_nodes != m.num_nodes) HEAT_FAIL(2, "heat_scatter_nodal_field: field must hold num_nodes = %lld values", (long long)m.num_nodes);
    if (!m.is_cube) heat::build_node_bc(ctx);          // from the CURRENT mesh: no assemble call need have come before
    return scatter_reduced(ctx, x_reduced_host, n_global, field_host);
}

// dense nodal field on rank 0 (ExodusIO.hpp:1981-1989 + :2045-2055); other ranks send their rows
extern "C" int heat_nodal_field(heat_ctx *ctx, const heat_vector *X, double *field, int64_t num_nodes) {
    if (!ctx || !X) HEAT_FAIL(2, "heat_nodal_field: null argument");
    const HostMesh &m = ctx->mesh;
    if (!m.valid) HEAT_FAIL(4, "heat_nodal_field: no mesh");
    HEAT_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> xl((size_t)X->n_owned);
    if (X->n_owned)
        HEAT_CUDA(cudaMemcpyAsync(xl.data(), X->d.p, sizeof(double) * (size_t)X->n_owned, cudaMemcpyDeviceToHost, ctx->stream));
    HEAT_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<double> xg;
    HEAT_TRY(heat::comm_gather_reduced(ctx, xl, ctx->n_global, xg));           // rank 0: x in reduced-id order
    if (ctx->rank != 0) return 0;
    if (!field || num_nodes != m.num_nodes) HEAT_FAIL(2, "heat_nodal_field: field must hold num_nodes = %lld values", (long long)m.num_nodes);
    return scatter_reduced(ctx, xg.data(), (int64_t)xg.size(), field);
}

// IO::writeSolution (ExodusIO.hpp:1972-2070): nodal variable "Steady-State Heat Solution",
// step timestep+1, time (real_t)timestep; the first call also writes a BC-only frame at step 1, time 0.
extern "C" int heat_write_solution(heat_ctx *ctx, const heat_vector *X, int timestep) {
    if (!ctx || !X) HEAT_FAIL(2, "heat_write_solution: null argument");
    if (timestep < 0) HEAT_FAIL(2, "heat_write_solution: negative timestep");
    const HostMesh &m = ctx->mesh;
    std::vector<double> field;
    if (ctx->rank == 0) field.resize((size_t)m.num_nodes);
    HEAT_TRY(heat_nodal_field(ctx, X, ctx->rank == 0 ? field.data() : nullptr, m.num_nodes));
    if (ctx->rank != 0) return 0;
    return heat_write_nodal_field(ctx, field.data(), m.num_nodes, timestep);
}

// the file half of writeSolution (ExodusIO.hpp:2027-2069): ex_put_variable_param / ex_put_variable_names on
// the first call, then ex_put_time + ex_put_var of one dense nodal array.  Host only (rank 0).
extern "C" int heat_write_nodal_field(heat_ctx *ctx, const double *field_host, int64_t num_nodes, int timestep) {
    if (!ctx || !field_host) HEAT_FAIL(2, "heat_write_nodal_field: null argument");
    if (timestep < 0) HEAT_FAIL(2, "heat_write_nodal_field: negative timestep");
    const HostMesh &m = ctx->mesh;
    if (!m.valid || num_nodes != m.num_nodes) HEAT_FAIL(2, "heat_write_nodal_field: field must hold num_nodes = %lld values", (long long)m.num_nodes);
    const double *field = field_host;
    if (!ctx->write_file) HEAT_FAIL(4, "heat_write_solution: no output file (call heat_create + heat_decompose)");
    ExoFile &wf = *ctx->write_file;
    NcFile &nc = wf.nc;
    if (nc.dim_id("num_nodes") < 0) HEAT_FAIL(4, "heat_write_solution: output mesh not written (call heat_decompose)");
    const int64_t N = m.num_nodes;
    bool layout_changed = false, first_call = false;
    if (!ctx->printed_time_zero) {                              // :2034-2040
        layout_changed = true; first_call = true;
        nc.add_dim("num_nod_var", 1);
        if (nc.dim_id("len_name") < 0) nc.add_dim("len_name", 33);
        const int64_t len_name = nc.dim_len("len_name", 33);
        std::vector<char> nm((size_t)len_name, 0);
        const char *var_name = "Steady-State Heat Solution";    // :2032
        strncpy(nm.data(), var_name, (size_t)len_name - 1);
        NcVar &nv = nc.add_var("name_nod_var", NC_CHAR, {"num_nod_var", "len_name"});
        NcFile::put_chars(nv, nm.data(), nm.size());
        const int real_type = ctx->out_word_size == 4 ? NC_FLOAT : NC_DOUBLE;
        NcVar &vv = nc.add_var("vals_nod_var1", real_type, {"time_step", "num_nodes"});
        vv.is_record = true;
        NcVar *tw = nc.var("time_whole");
        if (!tw) { tw = &nc.add_var("time_whole", real_type, {"time_step"}); }
        tw->is_record = true; tw->type = real_type; tw->raw.clear();
        nc.numrecs = 0;
        ctx->printed_time_zero = true;
    }
    NcVar *tw = nc.var("time_whole");
    NcVar *vv = nc.var("vals_nod_var1");
    const int64_t step = (int64_t)timestep + 1;                 // 1-based Exodus step (:2043, :2056)
    const int64_t ws = nc_type_size(vv->type);                  // 8, or 4 when the file is written as the reference does (D6)
    if (step > nc.numrecs) {
        tw->raw.resize((size_t)(step * ws), 0);
        vv->raw.resize((size_t)(step * N * ws), 0);
        nc.numrecs = step;
    }
    if (first_call && step > 1) {
        // the reference's first call also writes a boundary-condition-only frame as step 1, time 0 (:2036-2039):
        // nodeset nodes hold their id, unknowns 0 — it survives whenever the first timestep written is not 0
        if ((int64_t)ctx->node_bc.size() != N && !m.is_cube) heat::build_node_bc(ctx);
        const std::vector<double> &bcv = ctx->out_largest_id ? ctx->node_bc_hi : ctx->node_bc;
        std::vector<double> frame((size_t)N, 0.0);
        if ((int64_t)bcv.size() == N)
            for (int64_t g = 0; g < N; ++g) frame[(size_t)g] = std::isnan(bcv[(size_t)g]) ? 0.0 : bcv[(size_t)g];
        if (ws == 8) {
            swap_copy(vv->raw.data(), frame.data(), (size_t)N, 8);
        } else {
            for (int64_t g = 0; g < N; ++g) { const float f = (float)frame[(size_t)g]; swap_copy(vv->raw.data() + 4 * g, &f, 1, 4); }
        }
    }
    const double t = (double)timestep;
    if (ws == 8) {
        swap_copy(tw->raw.data() + (step - 1) * 8, &t, 1, 8);
        swap_copy(vv->raw.data() + (step - 1) * N * 8, field, (size_t)N, 8);
    } else {
        const float tf = (float)t;                              // real_t t = (real_t) timestep (:2042)
        swap_copy(tw->raw.data() + (step - 1) * 4, &tf, 1, 4);
        uint8_t *dst = vv->raw.data() + (step - 1) * N * 4;
        for (int64_t g = 0; g < N; ++g) { const float f = (float)field[g]; swap_copy(dst + 4 * g, &f, 1, 4); }
    }
    wf.steps_written = nc.numrecs;
    // the first result call lays the file out again (new dimension + variables: ex_put_variable_param);
    // later calls only touch the record they write, like ex_put_var (:2056)
    if (layout_changed) return nc_write(wf.path, nc);
    return nc_update_records(wf.path, nc, step - 1, step);
}
