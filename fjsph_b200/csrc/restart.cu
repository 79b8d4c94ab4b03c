// restart.cu — checkpoint / resume of a simulation.
//
// A raw-binary mirror of the reference's `<prefix>_particles.h5` restart data (HDF5 is not available in this image):
//   per-particle field set        reference src/H5IO.cpp:395-523   (Write_Zone_Data: the dataset names are kept)
//   inlet back / buffer tables    reference src/H5IO.cpp:525-538   (Write_Inlet_Data)
//   simulation attributes         reference src/H5IO.cpp:915-962   (current time, frame, particle index to add, block
//                                                                   and point counts, previous frame time)
// The reference restores pn = pnp1 from the file and recomputes the frozen terms (FJSPH.cpp:189-212, H5IO.cpp Read_HDF5);
// so does fjsph_read_restart: both time levels are uploaded from the one stored level, the neighbour lists and the
// frozen terms are rebuilt by the next step.  Layout (little endian):
//   "FJSPHB2R" | u32 version | u32 dim | i64 n | i64 bound_points | i64 next_part_id | i32 frame | u32 n_blocks |
//   u32 sizeof(FjsphParams) | FjsphParams | blocks | u32 n_datasets | datasets
//   block   = i64 first, second | i32 is_fluid, bound_solver, no_slip, block_type, fixed_vel_or_dynamic |
//             u32 n_times | f64 times[] | u32 n_vels | f64 vels[] | f64 insert_norm[3], insconst, delete_norm[3], delconst,
//             aero_norm[3], aeroconst | u32 n_back | u32 n_buf | i64 back[] | i64 buffer[n_back][n_buf]
//   dataset = u32 name_len | name | u32 type (0 f64, 1 i32, 2 i64) | u64 count | data
// One dataset beyond the reference's set: "Curvature" (SPHPart::curve), which find_timestep reads for the surface-tension
// limit before the first resumed step has recomputed it (Integration.cpp:393-396); a reader may ignore it.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "engine.cuh"

namespace
{
const char kMagic[8] = {'F', 'J', 'S', 'P', 'H', 'B', '2', 'R'};
constexpr uint32_t kVersion = 1;

struct Writer
{
    FILE* f;
    bool ok = true;
    void raw(const void* p, size_t n)
    {
        if (ok && n && fwrite(p, 1, n, f) != n)
            ok = false;
    }
    template <class T>
    void put(const T& v)
    {
        raw(&v, sizeof(T));
    }
    template <class T>
    void dataset(const char* name, uint32_t type, const std::vector<T>& v)
    {
        put<uint32_t>(uint32_t(strlen(name)));
        raw(name, strlen(name));
        put<uint32_t>(type);
        put<uint64_t>(uint64_t(v.size()));
        raw(v.data(), v.size() * sizeof(T));
    }
};
struct Reader
{
    FILE* f;
    bool ok = true;
    void raw(void* p, size_t n)
    {
        if (ok && n && fread(p, 1, n, f) != n)
            ok = false;
    }
    template <class T>
    T get()
    {
        T v{};
        raw(&v, sizeof(T));
        return v;
    }
};
const char* kAxis[3] = {"x", "y", "z"};
} // namespace

extern "C" int fjsph_write_restart(FjsphEngine* e, const char* path, int32_t frame)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e);
    if (!path || e->n_owned <= 0)
    {
        fj_set_error("write_restart: no path or no particles");
        return FJSPH_ERR_INVALID;
    }
    if (e->slab.on && e->slab.world > 1)
    {
        fj_set_error("write_restart: gather the slabs on the host first (fjsph_download_state per rank)");
        return FJSPH_ERR_INVALID;
    }
    const size_t n = size_t(e->n_owned);
    std::vector<double> xi(3 * n), v(3 * n), acc(3 * n), cellV(3 * n), p(n), rho(n), Rrho(n), m(n), cellRho(n), cellP(n), curve(n);
    std::vector<int32_t> b(n);
    std::vector<int64_t> pid(n), cid(n);
    FjsphStateView s;
    std::memset(&s, 0, sizeof(s));
    s.n = int64_t(n);
    s.curve = curve.data();
    s.xi = xi.data();
    s.v = v.data();
    s.acc = acc.data();
    s.cellV = cellV.data();
    s.p = p.data();
    s.rho = rho.data();
    s.Rrho = Rrho.data();
    s.m = m.data();
    s.cellRho = cellRho.data();
    s.cellP = cellP.data();
    s.b = b.data();
    s.part_id = pid.data();
    s.cellID = cid.data();
    int st = fjsph_download_state(e, 1, &s);
    if (st)
        return st;
    /* written beside the target and renamed over it: a crash mid-write never damages the checkpoint that is there */
    const std::string tmp_path = std::string(path) + ".tmp";
    FILE* f = fopen(tmp_path.c_str(), "wb");
    if (!f)
    {
        fj_set_error("write_restart: cannot open \"%s\"", tmp_path.c_str());
        return FJSPH_ERR_IO;
    }
    Writer w{f};
    w.raw(kMagic, 8);
    w.put<uint32_t>(kVersion);
    w.put<uint32_t>(3);
    w.put<int64_t>(int64_t(n));
    w.put<int64_t>(e->bound_points);
    w.put<int64_t>(e->next_part_id);
    w.put<int32_t>(frame);
    w.put<uint32_t>(uint32_t(e->blocks.size()));
    w.put<uint32_t>(uint32_t(sizeof(FjsphParams)));
    w.put(e->P);
    for (const HostBlock& B : e->blocks)
    {
        w.put(B.first);
        w.put(B.second);
        for (int32_t q : {B.is_fluid, B.bound_solver, B.no_slip, B.block_type, B.fixed_vel_or_dynamic}) w.put(q);
        w.put<uint32_t>(uint32_t(B.times.size()));
        w.raw(B.times.data(), B.times.size() * sizeof(double));
        w.put<uint32_t>(uint32_t(B.vels.size()));
        w.raw(B.vels.data(), B.vels.size() * sizeof(double));
        w.raw(B.insert_norm, sizeof(B.insert_norm));
        w.put(B.insconst);
        w.raw(B.delete_norm, sizeof(B.delete_norm));
        w.put(B.delconst);
        w.raw(B.aero_norm, sizeof(B.aero_norm));
        w.put(B.aeroconst);
        const uint32_t nb = uint32_t(B.back.size()), nf = nb ? uint32_t(B.buffer[0].size()) : 0u;
        w.put(nb);
        w.put(nf);
        w.raw(B.back.data(), B.back.size() * sizeof(int64_t));
        for (const std::vector<int64_t>& row : B.buffer) w.raw(row.data(), row.size() * sizeof(int64_t));
    }
    w.put<uint32_t>(4 * 3 + 9 + 1); /* 4 vectors by component, 9 scalars / flags / ids, the curvature */
    auto comp = [&](const std::vector<double>& a, int d) {
        std::vector<double> c(n);
        for (size_t i = 0; i < n; ++i) c[i] = a[3 * i + size_t(d)];
        return c;
    };
    for (int d = 0; d < 3; ++d) w.dataset((std::string("Position coordinate ") + kAxis[d]).c_str(), 0, comp(xi, d));
    for (int d = 0; d < 3; ++d) w.dataset((std::string("Velocity ") + kAxis[d]).c_str(), 0, comp(v, d));
    for (int d = 0; d < 3; ++d) w.dataset((std::string("Acceleration ") + kAxis[d]).c_str(), 0, comp(acc, d));
    w.dataset("Pressure", 0, p);
    w.dataset("Density", 0, rho);
    w.dataset("Density gradient", 0, Rrho);
    w.dataset("Mass", 0, m);
    w.dataset("Boundary condition", 1, b);
    w.dataset("Particle ID", 2, pid);
    w.dataset("Cell ID", 2, cid);
    for (int d = 0; d < 3; ++d) w.dataset((std::string("Cell velocity ") + kAxis[d]).c_str(), 0, comp(cellV, d));
    w.dataset("Cell density", 0, cellRho);
    w.dataset("Cell pressure", 0, cellP);
    w.dataset("Curvature", 0, curve);
    const bool ok = w.ok && fflush(f) == 0;
    const bool closed = fclose(f) == 0;
    if (!ok || !closed || std::rename(tmp_path.c_str(), path) != 0)
    {
        std::remove(tmp_path.c_str());
        fj_set_error("write_restart: short write to \"%s\"", path);
        return FJSPH_ERR_IO;
    }
    return FJSPH_OK;
}

static int read_restart_impl(FjsphEngine* e, const char* path, int32_t* frame, FILE*& f);

// A damaged or truncated file is an error return, never an exception across the C boundary (the driver overwrites its
// only checkpoint every frame: every length read from the file is bounded before anything is sized by it).
extern "C" int fjsph_read_restart(FjsphEngine* e, const char* path, int32_t* frame)
{
    FILE* f = nullptr;
    try
    {
        return read_restart_impl(e, path, frame, f);
    }
    catch (const std::exception& ex)
    {
        if (f)
            fclose(f);
        fj_set_error("read_restart: \"%s\" could not be read (%s)", path ? path : "(null)", ex.what());
        return FJSPH_ERR_IO;
    }
}

static int read_restart_impl(FjsphEngine* e, const char* path, int32_t* frame, FILE*& f)
{
    cudaSetDevice(e->device);
    fj_halo_wait(e);
    f = path ? fopen(path, "rb") : nullptr;
    if (!f)
    {
        fj_set_error("read_restart: cannot open \"%s\"", path ? path : "(null)");
        return FJSPH_ERR_IO;
    }
    Reader r{f};
    char magic[8];
    r.raw(magic, 8);
    const uint32_t version = r.get<uint32_t>(), dim = r.get<uint32_t>();
    if (!r.ok || std::memcmp(magic, kMagic, 8) != 0 || version != kVersion || dim != 3)
    {
        fclose(f);
        f = nullptr;
        fj_set_error("read_restart: \"%s\" is not a version-%u 3D restart file of this engine", path, kVersion);
        return FJSPH_ERR_IO;
    }
    const int64_t n = r.get<int64_t>(), bound_points = r.get<int64_t>(), next_part_id = r.get<int64_t>();
    const int32_t fr = r.get<int32_t>();
    const uint32_t n_blocks = r.get<uint32_t>(), psize = r.get<uint32_t>();
    if (!r.ok || psize != sizeof(FjsphParams) || n <= 0 || n > e->cap || n_blocks > 4096u)
    {
        fclose(f);
        f = nullptr;
        fj_set_error("read_restart: %lld particles / a %u-byte parameter block do not fit this engine (capacity %lld, %zu bytes)",
                     (long long)n, psize, (long long)e->cap, sizeof(FjsphParams));
        return FJSPH_ERR_CAPACITY;
    }
    FjsphParams P = r.get<FjsphParams>();
    struct Blk
    {
        FjsphBlock b;
        std::vector<double> times, vels;
        std::vector<int64_t> back, buffer;
    };
    std::vector<Blk> blocks(n_blocks);
    for (Blk& B : blocks)
    {
        std::memset(&B.b, 0, sizeof(B.b));
        B.b.first = r.get<int64_t>();
        B.b.second = r.get<int64_t>();
        B.b.is_fluid = r.get<int32_t>();
        B.b.bound_solver = r.get<int32_t>();
        B.b.no_slip = r.get<int32_t>();
        B.b.block_type = r.get<int32_t>();
        B.b.fixed_vel_or_dynamic = r.get<int32_t>();
        const uint32_t n_times = r.get<uint32_t>();
        if (!r.ok || n_times > 65536u)
        {
            r.ok = false;
            break;
        }
        B.times.resize(n_times);
        r.raw(B.times.data(), B.times.size() * sizeof(double));
        const uint32_t n_vels = r.get<uint32_t>();
        if (!r.ok || n_vels != 3u * std::max(1u, n_times)) /* one velocity per time stamp, or the single constant one */
        {
            r.ok = false;
            break;
        }
        B.vels.resize(n_vels);
        r.raw(B.vels.data(), B.vels.size() * sizeof(double));
        r.raw(B.b.insert_norm, sizeof(B.b.insert_norm));
        B.b.insconst = r.get<double>();
        r.raw(B.b.delete_norm, sizeof(B.b.delete_norm));
        B.b.delconst = r.get<double>();
        r.raw(B.b.aero_norm, sizeof(B.b.aero_norm));
        B.b.aeroconst = r.get<double>();
        const uint32_t nb = r.get<uint32_t>(), nf = r.get<uint32_t>();
        if (!r.ok || uint64_t(nb) * nf > uint64_t(n) * 8u)
        {
            r.ok = false;
            break;
        }
        B.back.resize(nb);
        r.raw(B.back.data(), B.back.size() * sizeof(int64_t));
        B.buffer.resize(size_t(nb) * nf);
        r.raw(B.buffer.data(), B.buffer.size() * sizeof(int64_t));
        B.b.n_times = int32_t(B.times.size());
        B.b.times = B.times.empty() ? nullptr : B.times.data();
        B.b.vels = B.vels.empty() ? nullptr : B.vels.data();
        B.b.n_back = int32_t(nb);
        B.b.n_buf = int32_t(nf);
        B.b.back = B.back.empty() ? nullptr : B.back.data();
        B.b.buffer = B.buffer.empty() ? nullptr : B.buffer.data();
    }
    const size_t N = size_t(n);
    std::vector<double> xi(3 * N), v(3 * N), acc(3 * N), cellV(3 * N), p(N), rho(N), Rrho(N), m(N), cellRho(N), cellP(N),
        curve(N, 0.0);
    std::vector<int32_t> b(N);
    std::vector<int64_t> pid(N), cid(N);
    /* one bit per required dataset: a file holding one dataset twice and lacking another is not complete */
    uint32_t seen_bits = 0;
    enum { D_XI = 0, D_V = 3, D_ACC = 6, D_CELLV = 9, D_P = 12, D_RHO, D_RRHO, D_M, D_CRHO, D_CP, D_B, D_PID, D_CID, D_COUNT };
    const uint32_t nds = r.ok ? r.get<uint32_t>() : 0u;
    for (uint32_t k = 0; k < nds && r.ok; ++k)
    {
        const uint32_t len = r.get<uint32_t>();
        if (!r.ok || len > 256)
        {
            r.ok = false;
            break;
        }
        std::string name(len, ' ');
        r.raw(&name[0], len);
        const uint32_t type = r.get<uint32_t>();
        const uint64_t count = r.get<uint64_t>();
        if (!r.ok || count != uint64_t(N))
        {
            r.ok = false;
            break;
        }
        std::vector<double> tmp;
        auto vec_comp = [&](std::vector<double>& dst, const char* prefix, int bit0) {
            for (int d = 0; d < 3; ++d)
                if (name == std::string(prefix) + kAxis[d] && type == 0)
                {
                    tmp.resize(N);
                    r.raw(tmp.data(), N * sizeof(double));
                    for (size_t i = 0; i < N; ++i) dst[3 * i + size_t(d)] = tmp[i];
                    seen_bits |= 1u << (bit0 + d);
                    return true;
                }
            return false;
        };
        auto scalar = [&](std::vector<double>& dst, const char* nm, int bit) {
            if (name == nm && type == 0)
            {
                r.raw(dst.data(), N * sizeof(double));
                seen_bits |= 1u << bit;
                return true;
            }
            return false;
        };
        if (vec_comp(xi, "Position coordinate ", D_XI) || vec_comp(v, "Velocity ", D_V) || vec_comp(acc, "Acceleration ", D_ACC) ||
            vec_comp(cellV, "Cell velocity ", D_CELLV) || scalar(p, "Pressure", D_P) || scalar(rho, "Density", D_RHO) ||
            scalar(Rrho, "Density gradient", D_RRHO) || scalar(m, "Mass", D_M) || scalar(cellRho, "Cell density", D_CRHO) ||
            scalar(cellP, "Cell pressure", D_CP))
            continue;
        if (name == "Curvature" && type == 0)
        {
            r.raw(curve.data(), N * sizeof(double));
            continue;
        }
        if (name == "Boundary condition" && type == 1)
        {
            r.raw(b.data(), N * sizeof(int32_t));
            seen_bits |= 1u << D_B;
        }
        else if (name == "Particle ID" && type == 2)
        {
            r.raw(pid.data(), N * sizeof(int64_t));
            seen_bits |= 1u << D_PID;
        }
        else if (name == "Cell ID" && type == 2)
        {
            r.raw(cid.data(), N * sizeof(int64_t));
            seen_bits |= 1u << D_CID;
        }
        else if (type > 2u || fseek(f, long(count * (type == 1 ? 4u : 8u)), SEEK_CUR) != 0) /* unknown dataset: skip it */
            r.ok = false;
    }
    fclose(f);
    f = nullptr;
    if (!r.ok || seen_bits != (1u << D_COUNT) - 1u)
    {
        fj_set_error("read_restart: \"%s\" is truncated or lacks datasets (%d of %d found)", path, __builtin_popcount(seen_bits),
                     int(D_COUNT));
        return FJSPH_ERR_IO;
    }
    int st = fjsph_set_params(e, &P);
    if (st)
        return st;
    FjsphStateView s;
    std::memset(&s, 0, sizeof(s));
    s.n = n;
    s.xi = xi.data();
    s.v = v.data();
    s.acc = acc.data();
    s.cellV = cellV.data();
    s.p = p.data();
    s.rho = rho.data();
    s.Rrho = Rrho.data();
    s.m = m.data();
    s.cellRho = cellRho.data();
    s.cellP = cellP.data();
    s.b = b.data();
    s.part_id = pid.data();
    s.cellID = cid.data();
    s.curve = curve.data();
    st = fjsph_upload_state(e, &s, bound_points); /* both time levels: pn = pnp1, as Read_HDF5 leaves them */
    if (st)
        return st;
    if (n_blocks)
    {
        std::vector<FjsphBlock> fb;
        for (Blk& B : blocks) fb.push_back(B.b);
        st = fjsph_set_blocks(e, int32_t(fb.size()), fb.data());
        if (st)
            return st;
    }
    e->next_part_id = next_part_id;
    if (frame)
        *frame = fr;
    return FJSPH_OK;
}
