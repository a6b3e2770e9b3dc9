// exodus.hpp — minimal Exodus-II on netCDF-classic (CDF-1/2/5 read, CDF-2 write), host only.
// Stands in for the SEACAS-Exodus C API the reference calls (60+ ex_* sites, ExodusIO.hpp:93..2077);
// neither libexodus nor libnetcdf exists in this image.  Every file under the reference's data/ is
// classic 64-bit-offset ("CDF\x02"), big-endian.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

enum NcType { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6,
              NC_UBYTE = 7, NC_USHORT = 8, NC_UINT = 9, NC_INT64 = 10, NC_UINT64 = 11 };

struct NcDim { std::string name; int64_t len = 0; bool unlimited = false; };
struct NcAtt { std::string name; int type = NC_CHAR; int64_t nelems = 0; std::vector<uint8_t> raw; };   // raw = big-endian
struct NcVar {
    std::string name;
    std::vector<int> dimids;
    std::vector<NcAtt> atts;
    int type = NC_INT;
    bool is_record = false;
    std::vector<uint8_t> raw;      // big-endian payload, records concatenated WITHOUT padding
};

struct NcFile {
    int version = 2;
    int64_t numrecs = 0;
    std::vector<NcDim> dims;
    std::vector<NcAtt> gatts;
    std::vector<NcVar> vars;

    int dim_id(const std::string &n) const;
    int64_t dim_len(const std::string &n, int64_t dflt = 0) const;
    const NcVar *var(const std::string &n) const;
    NcVar *var(const std::string &n);
    const NcAtt *gatt(const std::string &n) const;
    int add_dim(const std::string &n, int64_t len, bool unlimited = false);
    void set_dim(const std::string &n, int64_t len);
    NcVar &add_var(const std::string &n, int type, const std::vector<std::string> &dim_names);
    void remove_var(const std::string &n);
    int64_t var_elems_per_record(const NcVar &v) const;      // product of non-record dims
    // typed access (host endian)
    std::vector<double> get_doubles(const NcVar &v) const;
    std::vector<int64_t> get_ints(const NcVar &v) const;
    std::string get_att_string(const std::vector<NcAtt> &atts, const std::string &n) const;
    static void put_doubles(NcVar &v, const double *p, size_t n);
    static void put_floats(NcVar &v, const double *p, size_t n);      // stored as NC_FLOAT (rounded to float32)
    static void put_ints(NcVar &v, const int32_t *p, size_t n);
    static void put_chars(NcVar &v, const char *p, size_t n);
    static NcAtt make_att_string(const std::string &n, const std::string &val);
    static NcAtt make_att_int(const std::string &n, int32_t val);
    static NcAtt make_att_float(const std::string &n, float val);
};

int nc_type_size(int type);
// 0 on success; message via heat::set_error
int nc_read(const std::string &path, NcFile &out);
int nc_write(const std::string &path, const NcFile &f);
int nc_update_records(const std::string &path, const NcFile &f, int64_t r0, int64_t r1);   // in place, same layout

struct ExoFile {
    std::string path;
    bool writable = false;
    NcFile nc;
    bool have_results = false;     // name_nod_var / vals_nod_var1 defined
    int64_t steps_written = 0;
};

struct HostMesh;
int exo_read_mesh(const ExoFile &f, HostMesh &m);      // what open()+assemble()'s reads need
void exo_close(ExoFile *f);
