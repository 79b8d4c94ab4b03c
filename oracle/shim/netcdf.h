// TEST INFRASTRUCTURE ONLY -- stand-in for <netcdf.h> (the library is absent from the image) so that the reference's
// CDFIO.cpp compiles where it lies (oracle/Makefile.ref) and reads TAU face-based mesh and solution files.  Covers the nine
// calls CDFIO.cpp makes, over NetCDF-3 "classic" files (CDF-1 and the 64-bit-offset CDF-2) with fixed-size variables, from
// the published file format: magic, numrecs, dimension list, global attributes, variable list (name, dimension ids,
// attributes, type, size, offset), big-endian data.  Values are converted to the requested type as the library does.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define NC_NOWRITE 0
#define NC_NOERR 0
#define NC_ENOTNC (-51)
#define NC_EBADDIM (-46)
#define NC_ENOTVAR (-49)
#define NC_EBADID (-33)
#define NC_EEDGE (-57)

namespace nc_shim
{
struct Var
{
    std::string name;
    std::vector<int> dims;
    int type = 0;
    uint64_t begin = 0;
};
struct File
{
    std::vector<unsigned char> bytes;
    std::vector<std::pair<std::string, size_t>> dims;
    std::vector<Var> vars;
    size_t pos = 0;
    bool ok = true;
    uint32_t u32()
    {
        if (pos + 4 > bytes.size())
        {
            ok = false;
            return 0;
        }
        const uint32_t v = (uint32_t(bytes[pos]) << 24) | (uint32_t(bytes[pos + 1]) << 16) | (uint32_t(bytes[pos + 2]) << 8) | bytes[pos + 3];
        pos += 4;
        return v;
    }
    uint64_t u64()
    {
        const uint64_t hi = u32();
        return (hi << 32) | u32();
    }
    std::string name()
    {
        const uint32_t n = u32();
        if (pos + n > bytes.size())
        {
            ok = false;
            return "";
        }
        std::string s(reinterpret_cast<const char*>(&bytes[pos]), n);
        pos += (n + 3u) & ~3u;
        return s;
    }
    static size_t width(int type) { return type == 1 || type == 2 ? 1 : type == 3 ? 2 : type == 6 ? 8 : 4; }
    void attributes()
    {
        const uint32_t tag = u32(), n = u32();
        if (tag == 0)
            return;
        for (uint32_t i = 0; i < n && ok; ++i)
        {
            name();
            const uint32_t type = u32(), cnt = u32();
            pos += (size_t(cnt) * width(int(type)) + 3u) & ~size_t(3);
        }
    }
    bool parse()
    {
        if (bytes.size() < 8 || bytes[0] != 'C' || bytes[1] != 'D' || bytes[2] != 'F' || (bytes[3] != 1 && bytes[3] != 2))
            return false;
        const bool wide = bytes[3] == 2;
        pos = 4;
        u32(); /* numrecs */
        uint32_t tag = u32(), n = u32();
        for (uint32_t i = 0; tag != 0 && i < n && ok; ++i)
        {
            std::string nm = name();
            dims.emplace_back(nm, size_t(u32()));
        }
        attributes();
        tag = u32();
        n = u32();
        for (uint32_t i = 0; tag != 0 && i < n && ok; ++i)
        {
            Var v;
            v.name = name();
            const uint32_t nd = u32();
            for (uint32_t d = 0; d < nd; ++d) v.dims.push_back(int(u32()));
            attributes();
            v.type = int(u32());
            u32(); /* vsize */
            v.begin = wide ? u64() : u32();
            vars.push_back(v);
        }
        return ok;
    }
    template <typename T>
    int read(int varid, size_t count, T* out)
    {
        if (varid < 0 || size_t(varid) >= vars.size())
            return NC_ENOTVAR;
        Var const& v = vars[size_t(varid)];
        size_t total = 1;
        for (int d : v.dims) total *= dims[size_t(d)].second;
        if (count > total)
            return NC_EEDGE;
        const size_t w = width(v.type);
        if (v.begin + count * w > bytes.size())
            return NC_EEDGE;
        for (size_t i = 0; i < count; ++i)
        {
            const unsigned char* p = &bytes[v.begin + i * w];
            if (v.type == 6)
            {
                uint64_t b = 0;
                for (int k = 0; k < 8; ++k) b = (b << 8) | p[k];
                double x;
                std::memcpy(&x, &b, 8);
                out[i] = T(x);
            }
            else if (v.type == 5)
            {
                const uint32_t b = (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3];
                float x;
                std::memcpy(&x, &b, 4);
                out[i] = T(x);
            }
            else if (v.type == 4)
                out[i] = T(int32_t((uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]));
            else if (v.type == 3)
                out[i] = T(int16_t((uint16_t(p[0]) << 8) | p[1]));
            else
                out[i] = T(int8_t(p[0]));
        }
        return NC_NOERR;
    }
};
inline std::map<int, std::unique_ptr<File>>& files()
{
    static std::map<int, std::unique_ptr<File>> f;
    return f;
}
inline File* get(int ncid)
{
    auto it = files().find(ncid);
    return it == files().end() ? nullptr : it->second.get();
}
} // namespace nc_shim

inline const char* nc_strerror(int e)
{
    switch (e)
    {
    case NC_NOERR: return "No error";
    case NC_ENOTNC: return "Not a NetCDF-3 classic file (or it could not be opened)";
    case NC_EBADDIM: return "Invalid dimension id or name";
    case NC_ENOTVAR: return "Variable not found";
    case NC_EBADID: return "Not a netcdf id";
    case NC_EEDGE: return "Start+count exceeds dimension bound";
    }
    return "Unknown error";
}
inline int nc_open(const char* path, int, int* ncid)
{
    std::unique_ptr<nc_shim::File> f(new nc_shim::File());
    FILE* fp = std::fopen(path, "rb");
    if (!fp)
        return NC_ENOTNC;
    std::fseek(fp, 0, SEEK_END);
    const long n = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    f->bytes.resize(size_t(n > 0 ? n : 0));
    const size_t got = f->bytes.empty() ? 0 : std::fread(f->bytes.data(), 1, f->bytes.size(), fp);
    std::fclose(fp);
    if (got != f->bytes.size() || !f->parse())
        return NC_ENOTNC;
    static int next = 65536;
    *ncid = next++;
    nc_shim::files()[*ncid] = std::move(f);
    return NC_NOERR;
}
inline int nc_close(int ncid) { return nc_shim::files().erase(ncid) ? NC_NOERR : NC_EBADID; }
inline int nc_inq_dimid(int ncid, const char* name, int* id)
{
    nc_shim::File* f = nc_shim::get(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t i = 0; i < f->dims.size(); ++i)
        if (f->dims[i].first == name)
        {
            *id = int(i);
            return NC_NOERR;
        }
    return NC_EBADDIM;
}
inline int nc_inq_dimlen(int ncid, int id, size_t* len)
{
    nc_shim::File* f = nc_shim::get(ncid);
    if (!f)
        return NC_EBADID;
    if (id < 0 || size_t(id) >= f->dims.size())
        return NC_EBADDIM;
    *len = f->dims[size_t(id)].second;
    return NC_NOERR;
}
inline int nc_inq_varid(int ncid, const char* name, int* id)
{
    nc_shim::File* f = nc_shim::get(ncid);
    if (!f)
        return NC_EBADID;
    for (size_t i = 0; i < f->vars.size(); ++i)
        if (f->vars[i].name == name)
        {
            *id = int(i);
            return NC_NOERR;
        }
    return NC_ENOTVAR;
}
namespace nc_shim
{
template <typename T>
inline int get_all(int ncid, int varid, T* out)
{
    File* f = get(ncid);
    if (!f)
        return NC_EBADID;
    if (varid < 0 || size_t(varid) >= f->vars.size())
        return NC_ENOTVAR;
    size_t total = 1;
    for (int d : f->vars[size_t(varid)].dims) total *= f->dims[size_t(d)].second;
    return f->read(varid, total, out);
}
} // namespace nc_shim
inline int nc_get_var_double(int ncid, int varid, double* out) { return nc_shim::get_all(ncid, varid, out); }
inline int nc_get_var_int(int ncid, int varid, int* out) { return nc_shim::get_all(ncid, varid, out); }
/* CDFIO.cpp only asks for whole arrays (start 0, count = the dimensions) */
inline int nc_get_vara_int(int ncid, int varid, const size_t* start, const size_t* count, int* out)
{
    nc_shim::File* f = nc_shim::get(ncid);
    if (!f)
        return NC_EBADID;
    if (varid < 0 || size_t(varid) >= f->vars.size())
        return NC_ENOTVAR;
    nc_shim::Var const& v = f->vars[size_t(varid)];
    size_t total = 1;
    for (size_t d = 0; d < v.dims.size(); ++d)
    {
        if (start[d] != 0 || count[d] != f->dims[size_t(v.dims[d])].second)
            return NC_EEDGE;
        total *= count[d];
    }
    return f->read(varid, total, out);
}
