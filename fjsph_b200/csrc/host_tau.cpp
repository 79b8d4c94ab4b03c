// host_tau.cpp -- TAU mesh + solution ingestion for the aero-mesh containment lookup, without the NetCDF library.
//
// Restates what FJSPH.cpp:76-78 runs for a 3D TAU case:
//   TAU::Read_tau_mesh_FACE, Get_Element, Get_Coordinates, Place_Faces    reference src/CDFIO.cpp:1228-1356,391-460,462-628,1103-1226
//   TAU::Read_SOLUTION, Average_Point_Data_to_Cell, KahanSum              reference src/CDFIO.cpp:655-822,112-155,70-110
// The mesh file is the FACE-based one FJSPH's Cell2Face converter writes: dimensions no_of_elements / no_of_faces /
// no_of_points / no_of_surfaceelements (+ no_of_triangles, no_of_quadrilaterals), variables points_of_triangles,
// points_of_quadrilaterals, points_xc/yc/zc, left_element_of_faces, right_element_of_faces (negative = boundary marker,
// kept as written).  Faces are the triangles followed by the quadrilaterals, which STAY four-cornered (the
// containment test then takes the reference's three-edge form, csrc/mesh.cu); a cell's faces come in face order; the
// cell values are Kahan-summed means over the cell's distinct vertices in ascending order (a std::set in the reference)
// of the point data density / x,y,z_velocity / pressure, and the cell centre is the same mean of the vertices.
// The files are read as NetCDF-3 "classic" (CDF-1, or CDF-2 with 64-bit offsets): magic, record count, dimension list,
// global attributes, variable list (name, dimension ids, attributes, type, size, offset), big-endian arrays of fixed
// size.  NetCDF-4 files (an HDF5 container) are reported as such -- `nccopy -k classic` converts them.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <exception>
#include <fstream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"
#include "host_mesh.h"

void fj_set_error(const char* fmt, ...);

namespace
{
struct TauError
{
    std::string msg;
};

class ClassicFile
{
  public:
    explicit ClassicFile(const std::string& path_) : path(path_)
    {
        std::ifstream fin(path, std::ios::binary);
        if (!fin.is_open())
            throw TauError{"cannot open " + path};
        raw.assign(std::istreambuf_iterator<char>(fin), std::istreambuf_iterator<char>());
        if (raw.size() >= 4 && std::memcmp(raw.data(), "\x89HDF", 4) == 0)
            throw TauError{path + " is a NetCDF-4 (HDF5) file; this reader takes NetCDF-3 classic files (nccopy -k classic)"};
        if (raw.size() < 8 || std::memcmp(raw.data(), "CDF", 3) != 0 || (raw[3] != 1 && raw[3] != 2))
            throw TauError{path + " is not a NetCDF-3 classic file"};
        const bool wide = raw[3] == 2;
        at = 4;
        word(); /* number of records: only fixed-size variables are read */
        uint32_t tag = word(), n = word();
        for (uint32_t i = 0; tag != 0 && i < n; ++i)
        {
            Dim d;
            d.name = text();
            d.len = word();
            dims.push_back(d);
        }
        skip_attributes();
        tag = word();
        n = word();
        for (uint32_t i = 0; tag != 0 && i < n; ++i)
        {
            Var v;
            v.name = text();
            const uint32_t nd = word();
            v.count = 1;
            for (uint32_t k = 0; k < nd; ++k)
            {
                const uint32_t id = word();
                if (id >= dims.size())
                    throw TauError{path + ": variable " + v.name + " names a dimension that does not exist"};
                if (dims[id].len == 0)
                    throw TauError{path + ": variable " + v.name + " is a record variable (unlimited dimension), not supported"};
                v.count *= dims[id].len;
            }
            skip_attributes();
            v.type = int(word());
            word(); /* vsize */
            v.begin = wide ? ((uint64_t(word()) << 32) | word()) : word();
            vars.push_back(v);
        }
    }
    bool dim(const char* name, size_t& len) const
    {
        for (const Dim& d : dims)
            if (d.name == name)
            {
                len = d.len;
                return true;
            }
        return false;
    }
    size_t need_dim(const char* name) const
    {
        size_t len = 0;
        if (!dim(name, len))
            throw TauError{path + ": no dimension \"" + name + "\""};
        return len;
    }
    bool has_variable(const char* name) const
    {
        for (const Var& v : vars)
            if (v.name == name)
                return true;
        return false;
    }
    // a whole variable, converted to T as the library's nc_get_var_<type> does
    template <typename T>
    std::vector<T> variable(const char* name, size_t expect) const
    {
        for (const Var& v : vars)
        {
            if (v.name != name)
                continue;
            if (v.count != expect)
                throw TauError{path + ": variable \"" + name + "\" holds " + std::to_string(v.count) + " values, expected " + std::to_string(expect)};
            const size_t w = v.type == 6 ? 8 : v.type == 3 ? 2 : (v.type == 1 || v.type == 2) ? 1 : 4;
            if (v.type < 1 || v.type > 6 || v.begin + uint64_t(v.count) * w > raw.size())
                throw TauError{path + ": variable \"" + name + "\" runs past the end of the file"};
            std::vector<T> out(v.count);
            const unsigned char* p = reinterpret_cast<const unsigned char*>(raw.data()) + v.begin;
            for (size_t i = 0; i < v.count; ++i, p += w)
            {
                uint64_t bits = 0;
                for (size_t k = 0; k < w; ++k) bits = (bits << 8) | p[k];
                if (v.type == 6)
                {
                    double x;
                    std::memcpy(&x, &bits, 8);
                    out[i] = T(x);
                }
                else if (v.type == 5)
                {
                    const uint32_t b32 = uint32_t(bits);
                    float x;
                    std::memcpy(&x, &b32, 4);
                    out[i] = T(x);
                }
                else if (v.type == 4)
                    out[i] = T(int32_t(uint32_t(bits)));
                else if (v.type == 3)
                    out[i] = T(int16_t(uint16_t(bits)));
                else
                    out[i] = T(int8_t(uint8_t(bits)));
            }
            return out;
        }
        throw TauError{path + ": no variable \"" + name + "\""};
    }

  private:
    struct Dim
    {
        std::string name;
        size_t len = 0;
    };
    struct Var
    {
        std::string name;
        size_t count = 0;
        int type = 0;
        uint64_t begin = 0;
    };
    std::string path;
    std::vector<char> raw;
    std::vector<Dim> dims;
    std::vector<Var> vars;
    size_t at = 0;
    uint32_t word()
    {
        if (at + 4 > raw.size())
            throw TauError{path + ": header runs past the end of the file"};
        const unsigned char* p = reinterpret_cast<const unsigned char*>(raw.data()) + at;
        at += 4;
        return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]);
    }
    std::string text()
    {
        const uint32_t n = word();
        if (at + n > raw.size())
            throw TauError{path + ": header runs past the end of the file"};
        std::string s(raw.data() + at, n);
        at += (size_t(n) + 3) & ~size_t(3);
        return s;
    }
    void skip_attributes()
    {
        const uint32_t tag = word(), n = word();
        for (uint32_t i = 0; tag != 0 && i < n; ++i)
        {
            text();
            const uint32_t type = word(), cnt = word();
            const size_t w = type == 6 ? 8 : type == 3 ? 2 : (type == 1 || type == 2) ? 1 : 4;
            at += (size_t(cnt) * w + 3) & ~size_t(3);
        }
    }
};

// the compensated sum of KahanSum (CDFIO.cpp:70-110), over the cell's vertices in ascending order
struct Kahan
{
    double sum = 0.0, c = 0.0;
    void add(double v)
    {
        const double y = v - c;
        const double t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
};
} // namespace

extern "C" int fjsph_tau_read(const char* mesh_file, const char* solution_file, double scale, FjsphFoamMesh** out)
{
    if (!mesh_file || !out)
    {
        fj_set_error("tau_read: need the mesh file and an output pointer");
        return FJSPH_ERR_INVALID;
    }
    try
    {
        std::unique_ptr<FjsphFoamMesh> M(new FjsphFoamMesh());
        const ClassicFile mesh(mesh_file);
        const size_t n_elem = mesh.need_dim("no_of_elements"), n_face = mesh.need_dim("no_of_faces"),
                     n_pnts = mesh.need_dim("no_of_points"), n_surf = mesh.need_dim("no_of_surfaceelements");
        /* faces: the triangles, then the quadrilaterals (Get_Element appends, CDFIO.cpp:1309-1318) */
        M->face_ptr.push_back(0);
        size_t n_in = 0;
        for (const auto& kind : {std::make_pair("no_of_triangles", "points_of_triangles"),
                                 std::make_pair("no_of_quadrilaterals", "points_of_quadrilaterals")})
        {
            size_t nf = 0;
            if (!mesh.dim(kind.first, nf))
                continue;
            const size_t per = mesh.need_dim(kind.first[6] == 't' ? "points_per_triangle" : "points_per_quadrilateral");
            const std::vector<int> pts = mesh.variable<int>(kind.second, nf * per);
            for (size_t f = 0; f < nf; ++f)
            {
                for (size_t k = 0; k < per; ++k)
                {
                    const int v = pts[f * per + k];
                    if (v < 0 || size_t(v) >= n_pnts)
                        throw TauError{std::string(kind.second) + " names point " + std::to_string(v) + " of " + std::to_string(n_pnts)};
                    M->face_vtx.push_back(v);
                }
                M->face_ptr.push_back(int64_t(M->face_vtx.size()));
            }
            n_in += nf;
        }
        if (n_in != n_face)
            throw TauError{"Mismatch of number of faces to that defined: " + std::to_string(n_face) + " faces, " + std::to_string(n_in) + " read"};
        /* coordinates, scaled (CDFIO.cpp:1326-1337) */
        const std::vector<double> xc = mesh.variable<double>("points_xc", n_pnts), yc = mesh.variable<double>("points_yc", n_pnts),
                                  zc = mesh.variable<double>("points_zc", n_pnts);
        M->verts.resize(3 * n_pnts);
        for (size_t i = 0; i < n_pnts; ++i)
        {
            M->verts[3 * i] = xc[i] * scale;
            M->verts[3 * i + 1] = yc[i] * scale;
            M->verts[3 * i + 2] = zc[i] * scale;
        }
        /* Place_Faces: (left, right) per face, every face into its cells in face order */
        const std::vector<int> left = mesh.variable<int>("left_element_of_faces", n_face),
                               right = mesh.variable<int>("right_element_of_faces", n_face);
        std::vector<std::vector<size_t>> cFaces(n_elem);
        size_t n_boundary = 0;
        M->leftright.resize(2 * n_face);
        for (size_t f = 0; f < n_face; ++f)
        {
            if (left[f] < 0 || size_t(left[f]) >= n_elem || right[f] >= int(n_elem))
                throw TauError{"face " + std::to_string(f) + " names a cell outside the mesh"};
            M->leftright[2 * f] = left[f];
            M->leftright[2 * f + 1] = right[f];
            cFaces[size_t(left[f])].push_back(f);
            if (right[f] >= 0)
                cFaces[size_t(right[f])].push_back(f);
            else
                ++n_boundary;
        }
        if (n_boundary != n_surf)
            throw TauError{"Mismatch of number of surface faces identified, and the number given. Identified: " +
                           std::to_string(n_boundary) + "  Given: " + std::to_string(n_surf)};
        /* point data of the solution file (Read_SOLUTION); without one the cells carry zeros */
        std::vector<double> u(n_pnts, 0.0), v(n_pnts, 0.0), w(n_pnts, 0.0), pr(n_pnts, 0.0), rho(n_pnts, 0.0);
        if (solution_file && solution_file[0])
        {
            const ClassicFile sol(solution_file);
            if (sol.need_dim("no_of_points") != n_pnts)
                throw TauError{"Solution file does not have the same number of vertices as the mesh."};
            rho = sol.variable<double>("density", n_pnts);
            u = sol.variable<double>("x_velocity", n_pnts);
            v = sol.variable<double>("y_velocity", n_pnts);
            w = sol.variable<double>("z_velocity", n_pnts);
            pr = sol.variable<double>("pressure", n_pnts);
        }
        /* Average_Point_Data_to_Cell */
        M->cell_ptr.push_back(0);
        M->cCentre.assign(3 * n_elem, 0.0);
        M->cVel.assign(3 * n_elem, 0.0);
        M->cP.assign(n_elem, 0.0);
        M->cRho.assign(n_elem, 0.0);
        std::vector<size_t> elem;
        for (size_t c = 0; c < n_elem; ++c)
        {
            elem.clear();
            for (size_t f : cFaces[c])
            {
                M->cell_faces.push_back(int64_t(f));
                for (int64_t k = M->face_ptr[f]; k < M->face_ptr[f + 1]; ++k) elem.push_back(size_t(M->face_vtx[size_t(k)]));
            }
            M->cell_ptr.push_back(int64_t(M->cell_faces.size()));
            std::sort(elem.begin(), elem.end());
            elem.erase(std::unique(elem.begin(), elem.end()), elem.end());
            if (elem.empty())
                throw TauError{"cell " + std::to_string(c) + " has no faces"};
            const double nv = double(elem.size());
            Kahan s[8];
            for (size_t i : elem)
            {
                for (int d = 0; d < 3; ++d) s[d].add(M->verts[3 * i + size_t(d)]);
                s[3].add(u[i]);
                s[4].add(v[i]);
                s[5].add(w[i]);
                s[6].add(pr[i]);
                s[7].add(rho[i]);
            }
            for (int d = 0; d < 3; ++d)
            {
                M->cCentre[3 * c + size_t(d)] = s[d].sum / nv;
                M->cVel[3 * c + size_t(d)] = s[3 + d].sum / nv;
            }
            M->cP[c] = s[6].sum / nv;
            M->cRho[c] = s[7].sum / nv;
        }
        *out = M.release();
        return FJSPH_OK;
    }
    catch (const TauError& e)
    {
        fj_set_error("tau_read: %s", e.msg.c_str());
        return FJSPH_ERR_IO;
    }
    catch (const std::exception& e)
    {
        fj_set_error("tau_read: %s", e.what());
        return FJSPH_ERR_IO;
    }
}

// TAU 2D meshes (the reference's -DSIMDIM=2 build): TAU::Read_tau_mesh_EDGE + TAU::Read_SOLUTION (reference src/CDFIO.cpp:
// 992-1097, 828-990, 462-600, 655-822; FJSPH.cpp:85-91) on the edge-based mesh file FJSPH's Cell2Edge writes and a TAU
// solution file.  Faces are EDGES (points_of_element_edges); the plane of the mesh is given by the coordinate variable the
// file LACKS (points_xc, points_yc or points_zc: exactly one must be missing); `vertices_in_use` maps every mesh point to
// its point in the (3D, two-layer) solution file; `offset_axis` (the para's "2D offset vector", 1 = x, 2 = y, 3 = z) picks
// the two velocity components, independently of the coordinates -- as in the reference.  The result is in the ABI's 3D
// shape with z = 0 (verts, cCentre, cVel [n][3]), ready for fjsph_upload_mesh on a 2D engine.
extern "C" int fjsph_tau_read_edge(const char* mesh_file, const char* solution_file, double scale, int32_t offset_axis,
                                   FjsphFoamMesh** out)
{
    if (!mesh_file || !out)
    {
        fj_set_error("tau_read_edge: need the mesh file and an output pointer");
        return FJSPH_ERR_INVALID;
    }
    try
    {
        std::unique_ptr<FjsphFoamMesh> M(new FjsphFoamMesh());
        const ClassicFile mesh(mesh_file);
        const size_t n_elem = mesh.need_dim("no_of_elements"), n_edge = mesh.need_dim("no_of_edges"),
                     per = mesh.need_dim("points_per_edge"), n_pnts = mesh.need_dim("no_of_points");
        /* the number of boundary edges: "no_of_surfaceelements" (what Read_tau_mesh_EDGE asks for, CDFIO.cpp:1040-1041), or --
           files of the layout the reference's own Examples/RAE2822 ships, which its current reader stops at -- the wall and
           far-field edge counts */
        size_t n_surf = 0;
        if (!mesh.dim("no_of_surfaceelements", n_surf))
        {
            size_t n_wall = 0, n_far = 0;
            if (!mesh.dim("no_of_wall_edges", n_wall) || !mesh.dim("no_of_farfield_edges", n_far))
                n_surf = mesh.need_dim("no_of_surfaceelements"); /* throws with the reference's diagnosis */
            else
                n_surf = n_wall + n_far;
        }
        if (per != 2)
            throw TauError{"points_per_edge is " + std::to_string(per) + ", expected 2"};
        const std::vector<int> pts = mesh.variable<int>("points_of_element_edges", n_edge * per);
        M->face_ptr.push_back(0);
        for (size_t f = 0; f < n_edge; ++f)
        {
            for (size_t k = 0; k < per; ++k)
            {
                const int v = pts[f * per + k];
                if (v < 0 || size_t(v) >= n_pnts)
                    throw TauError{"points_of_element_edges names point " + std::to_string(v) + " of " + std::to_string(n_pnts)};
                M->face_vtx.push_back(v);
            }
            M->face_ptr.push_back(int64_t(M->face_vtx.size()));
        }
        const std::vector<int> used = mesh.variable<int>("vertices_in_use", n_pnts);
        /* Get_Coordinates, 2D: the variable that is absent names the ignored dimension */
        const bool hx = mesh.has_variable("points_xc"), hy = mesh.has_variable("points_yc"), hz = mesh.has_variable("points_zc");
        if (int(!hx) + int(!hy) + int(!hz) > 1)
            throw TauError{"More than one dimension was not aquired, meaning something went wrong."};
        if (hx && hy && hz)
            throw TauError{"The ignored dimension was not found."};
        const std::vector<double> c0 = mesh.variable<double>(!hx ? "points_yc" : "points_xc", n_pnts),
                                  c1 = mesh.variable<double>(!hz ? "points_yc" : "points_zc", n_pnts);
        M->verts.assign(3 * n_pnts, 0.0);
        for (size_t i = 0; i < n_pnts; ++i)
        {
            M->verts[3 * i] = c0[i] * scale;
            M->verts[3 * i + 1] = c1[i] * scale;
        }
        /* Place_Edges */
        const std::vector<int> left = mesh.variable<int>("left_element_of_edges", n_edge),
                               right = mesh.variable<int>("right_element_of_edges", n_edge);
        std::vector<std::vector<size_t>> cFaces(n_elem);
        size_t n_boundary = 0;
        M->leftright.resize(2 * n_edge);
        for (size_t f = 0; f < n_edge; ++f)
        {
            if (left[f] < 0 || size_t(left[f]) >= n_elem || right[f] >= int(n_elem))
                throw TauError{"edge " + std::to_string(f) + " names a cell outside the mesh"};
            M->leftright[2 * f] = left[f];
            M->leftright[2 * f + 1] = right[f];
            cFaces[size_t(left[f])].push_back(f);
            if (right[f] >= 0)
                cFaces[size_t(right[f])].push_back(f);
            else
                ++n_boundary;
        }
        if (n_boundary != n_surf)
            throw TauError{"Mismatch of number of surface faces identified, and the number given. Identified: " +
                           std::to_string(n_boundary) + "  Given: " + std::to_string(n_surf)};
        /* Read_SOLUTION, 2D: point data taken at vertices_in_use */
        std::vector<double> u(n_pnts, 0.0), w(n_pnts, 0.0), pr(n_pnts, 0.0), rho(n_pnts, 0.0);
        if (solution_file && solution_file[0])
        {
            const ClassicFile sol(solution_file);
            const size_t sol_pts = sol.need_dim("no_of_points");
            const char *n0 = nullptr, *n1 = nullptr;
            if (offset_axis == 1)
                n0 = "y_velocity", n1 = "z_velocity";
            else if (offset_axis == 2)
                n0 = "x_velocity", n1 = "z_velocity";
            else if (offset_axis == 3)
                n0 = "x_velocity", n1 = "y_velocity";
            else
                throw TauError{"velocities do not have the same number of vertices as the mesh (2D offset vector is " +
                               std::to_string(offset_axis) + ", expected 1, 2 or 3)"};
            const std::vector<double> su = sol.variable<double>(n0, sol_pts), sw = sol.variable<double>(n1, sol_pts),
                                      sp = sol.variable<double>("pressure", sol_pts), sr = sol.variable<double>("density", sol_pts);
            for (size_t i = 0; i < n_pnts; ++i)
            {
                if (used[i] < 0 || size_t(used[i]) >= sol_pts)
                    throw TauError{"vertices_in_use names point " + std::to_string(used[i]) + " of the solution's " + std::to_string(sol_pts)};
                const size_t k = size_t(used[i]);
                u[i] = su[k];
                w[i] = sw[k];
                pr[i] = sp[k];
                rho[i] = sr[k];
            }
        }
        /* Average_Point_Data_to_Cell */
        M->cell_ptr.push_back(0);
        M->cCentre.assign(3 * n_elem, 0.0);
        M->cVel.assign(3 * n_elem, 0.0);
        M->cP.assign(n_elem, 0.0);
        M->cRho.assign(n_elem, 0.0);
        std::vector<size_t> elem;
        for (size_t c = 0; c < n_elem; ++c)
        {
            elem.clear();
            for (size_t f : cFaces[c])
            {
                M->cell_faces.push_back(int64_t(f));
                for (int64_t k = M->face_ptr[f]; k < M->face_ptr[f + 1]; ++k) elem.push_back(size_t(M->face_vtx[size_t(k)]));
            }
            M->cell_ptr.push_back(int64_t(M->cell_faces.size()));
            std::sort(elem.begin(), elem.end());
            elem.erase(std::unique(elem.begin(), elem.end()), elem.end());
            if (elem.empty())
                throw TauError{"cell " + std::to_string(c) + " has no edges"};
            const double nv = double(elem.size());
            Kahan s[6];
            for (size_t i : elem)
            {
                s[0].add(M->verts[3 * i]);
                s[1].add(M->verts[3 * i + 1]);
                s[2].add(u[i]);
                s[3].add(w[i]);
                s[4].add(pr[i]);
                s[5].add(rho[i]);
            }
            M->cCentre[3 * c] = s[0].sum / nv;
            M->cCentre[3 * c + 1] = s[1].sum / nv;
            M->cVel[3 * c] = s[2].sum / nv;
            M->cVel[3 * c + 1] = s[3].sum / nv;
            M->cP[c] = s[4].sum / nv;
            M->cRho[c] = s[5].sum / nv;
        }
        *out = M.release();
        return FJSPH_OK;
    }
    catch (const TauError& e)
    {
        fj_set_error("tau_read_edge: %s", e.msg.c_str());
        return FJSPH_ERR_IO;
    }
    catch (const std::exception& e)
    {
        fj_set_error("tau_read_edge: %s", e.what());
        return FJSPH_ERR_IO;
    }
}
