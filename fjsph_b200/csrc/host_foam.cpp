// host_foam.cpp — OpenFOAM case ingestion for the aero-mesh containment lookup.
//
// Restates FOAM::Read_FOAM for ASCII and binary cases (library free), feeding fjsph_upload_mesh:
//   Read_Header / Read_Preamble / Read_Patch / Read_Boundary   reference src/FOAMIO.cpp:346-536
//   ascii::Read_{Label,Scalar,Vector,Face}_Data                reference src/FOAMIO.cpp:21-112
//   binary::Read_{Label,Scalar,Vector,Face}_Data               reference src/FOAMIO.cpp:113-342
//   Read_Points / Read_Faces / Read_Label_Field                reference src/FOAMIO.cpp:685-865
//   Read_polyMesh / Post_Process                               reference src/FOAMIO.cpp:538-683,867-903
//   Read_Solution                                              reference src/FOAMIO.cpp:905-941
// Post_Process: boundary faces get leftright.second = -1 (patch of type wall: inner wall) or -2 (any other patch: outer
// boundary), faces with more than 3 vertices are fanned into triangles (0, j+1, j+2) -- which is what makes Crossings3D
// complete --, cell -> faces lists follow the triangle order, and the cell "centre" is the reference's:
// the mean of the cell's sorted vertex list AFTER std::unique without erase (FOAMIO.cpp:652-664), i.e. duplicates in the
// tail still count.  It is only the seed of the k-nearest-centre search, so it is kept bit for bit.
// Deviations, each where the reference is undefined: cells.cRho is never filled by the reference although FindCell reads
// it (Q8) -> filled with the rho_fill argument; the cell count is taken from the owner AND the neighbour file (the
// reference keeps only the neighbour file's maximum, which is 0 for a one-cell mesh).
// Binary files: each file's own header says whether it is binary and how wide its labels and scalars are (`arch
// "LSB;label=32;scalar=64"`; 32 / 64 when the header has no arch entry, which the reference leaves uninitialised); the raw
// list starts one byte after the line holding its size (the opening bracket), faces are a faceCompactList (offsets, then
// the vertex labels).  Little-endian only, like the reference, which reads the bytes as they are.
#include <algorithm>
#include <exception>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"

void fj_set_error(const char* fmt, ...);

#include "host_mesh.h"

namespace
{
struct FoamError
{
    std::string msg;
};

struct Format /* what Read_Header takes from a FoamFile dictionary */
{
    bool binary = false;
    int label_bits = 32, scalar_bits = 64;
};

size_t file_bytes(const std::string& file)
{
    std::ifstream f(file, std::ifstream::binary | std::ifstream::ate);
    return f.is_open() ? size_t(std::max<std::streamoff>(f.tellg(), 0)) : 0;
}

// FoamFile header + the element count that follows it (Read_Preamble)
size_t preamble(std::ifstream& fin, const std::string& file, const char* exp_class, Format& fmt, const char* binary_class = nullptr)
{
    if (!fin.is_open())
        throw FoamError{"cannot open " + file};
    std::string line, cls;
    while (line.find("FoamFile") == std::string::npos)
        if (!std::getline(fin, line))
            throw FoamError{file + ": no FoamFile header"};
    line.clear();
    while (line.find('}') == std::string::npos)
    {
        if (!std::getline(fin, line))
            throw FoamError{file + ": unterminated FoamFile header"};
        if (line.find("format") != std::string::npos)
            fmt.binary = line.find("ascii") == std::string::npos;
        if (line.find("arch") != std::string::npos)
        {
            const size_t p1 = line.find("label"), p2 = line.find("scalar");
            if (p1 != std::string::npos && p1 + 8 <= line.size())
                fmt.label_bits = std::atoi(line.substr(p1 + 6, 2).c_str());
            if (p2 != std::string::npos && p2 + 9 <= line.size())
                fmt.scalar_bits = std::atoi(line.substr(p2 + 7, 2).c_str());
            if ((fmt.label_bits != 32 && fmt.label_bits != 64) || (fmt.scalar_bits != 32 && fmt.scalar_bits != 64))
                throw FoamError{file + ": unsupported arch entry " + line};
        }
        if (line.find("class") != std::string::npos)
        {
            std::istringstream iss(line);
            std::string tmp;
            iss >> tmp >> cls;
            if (!cls.empty() && cls.back() == ';')
                cls.pop_back();
        }
    }
    const char* want = (fmt.binary && binary_class) ? binary_class : exp_class;
    if (cls != want)
        throw FoamError{"File " + file + " is class \"" + cls + "\" and should be \"" + want + "\""};
    if (!std::getline(fin, line))
        throw FoamError{file + ": no data"};
    while (line.find("//") == 0 || line.empty() || line.find("dimensions") != std::string::npos)
        if (!std::getline(fin, line))
            throw FoamError{file + ": no data"};
    if (line.find("internalField") != std::string::npos)
    {
        if (line.find("nonuniform") == std::string::npos)
            throw FoamError{file + ": a uniform internalField carries no per-cell list (the reference reads a list)"};
        std::getline(fin, line);
    }
    size_t n = 0;
    std::istringstream iss(line);
    if (!(iss >> n))
        throw FoamError{file + ": expected the list size, found \"" + line + "\""};
    /* an entry takes at least one byte of the file: a damaged count must not size a vector of its own making */
    if (n > file_bytes(file))
        throw FoamError{file + ": the list size " + std::to_string(n) + " exceeds the file"};
    return n;
}

// the raw list of a binary file: reopened in binary mode one byte past the size line (FOAMIO.cpp:702-706)
void reopen_binary(std::ifstream& fin, const std::string& file)
{
    const std::streamoff pos = fin.tellg();
    fin.close();
    fin.open(file, std::ifstream::binary);
    fin.seekg(pos + 1);
}
template <typename Raw, typename Out>
void read_raw(std::ifstream& fin, const std::string& file, size_t count, Out* dst)
{
    std::vector<Raw> buf(count);
    fin.read(reinterpret_cast<char*>(buf.data()), std::streamsize(count * sizeof(Raw)));
    if (size_t(fin.gcount()) != count * sizeof(Raw))
        throw FoamError{file + ": binary list ends early"};
    for (size_t i = 0; i < count; ++i) dst[i] = Out(buf[i]);
}
template <typename Out>
void read_raw_scalars(std::ifstream& fin, const std::string& file, const Format& fmt, size_t count, Out* dst)
{
    if (fmt.scalar_bits == 32)
        read_raw<float>(fin, file, count, dst);
    else
        read_raw<double>(fin, file, count, dst);
}
template <typename Out>
void read_raw_labels(std::ifstream& fin, const std::string& file, const Format& fmt, size_t count, Out* dst)
{
    if (fmt.label_bits == 32)
        read_raw<int32_t>(fin, file, count, dst);
    else
        read_raw<int64_t>(fin, file, count, dst);
}

void read_vectors(const std::string& file, const char* cls, std::vector<double>& out)
{
    std::ifstream fin(file);
    Format fmt;
    const size_t n = preamble(fin, file, cls, fmt);
    if (fmt.binary)
    {
        reopen_binary(fin, file);
        out.assign(3 * n, 0.0);
        read_raw_scalars(fin, file, fmt, 3 * n, out.data());
        return;
    }
    std::string line;
    std::getline(fin, line); /* the opening bracket */
    out.assign(3 * n, 0.0);
    for (size_t i = 0; i < n; ++i)
    {
        if (!std::getline(fin, line))
            throw FoamError{file + ": list ends early"};
        line.erase(std::remove(line.begin(), line.end(), '('), line.end());
        line.erase(std::remove(line.begin(), line.end(), ')'), line.end());
        std::istringstream iss(line);
        iss >> out[3 * i] >> out[3 * i + 1] >> out[3 * i + 2];
    }
}
void read_scalars(const std::string& file, std::vector<double>& out)
{
    std::ifstream fin(file);
    Format fmt;
    const size_t n = preamble(fin, file, "volScalarField", fmt);
    if (fmt.binary)
    {
        reopen_binary(fin, file);
        out.assign(n, 0.0);
        read_raw_scalars(fin, file, fmt, n, out.data());
        return;
    }
    std::string line;
    std::getline(fin, line);
    out.assign(n, 0.0);
    for (size_t i = 0; i < n; ++i)
    {
        if (!std::getline(fin, line))
            throw FoamError{file + ": list ends early"};
        std::istringstream iss(line);
        iss >> out[i];
    }
}
void read_labels(const std::string& file, std::vector<int>& out, size_t& n_cells)
{
    std::ifstream fin(file);
    Format fmt;
    const size_t n = preamble(fin, file, "labelList", fmt);
    if (fmt.binary)
    {
        reopen_binary(fin, file);
        out.assign(n, 0);
        read_raw_labels(fin, file, fmt, n, out.data());
        for (size_t i = 0; i < n; ++i)
            if (out[i] + 1 > int(n_cells))
                n_cells = size_t(out[i] + 1);
        return;
    }
    std::string line;
    std::getline(fin, line);
    out.assign(n, 0);
    for (size_t i = 0; i < n; ++i)
    {
        if (!std::getline(fin, line))
            throw FoamError{file + ": list ends early"};
        std::istringstream iss(line);
        iss >> out[i];
        if (out[i] + 1 > int(n_cells))
            n_cells = size_t(out[i] + 1);
    }
}
void read_faces(const std::string& file, std::vector<std::vector<size_t>>& faces)
{
    std::ifstream fin(file);
    Format fmt;
    const size_t n = preamble(fin, file, "faceList", fmt, "faceCompactList");
    if (fmt.binary)
    {
        /* faceCompactList (FOAMIO.cpp:225-340): n = faces + 1 offsets, then ")", the label count, "(" and the labels */
        if (n == 0)
            throw FoamError{file + ": empty offset list"};
        reopen_binary(fin, file);
        std::vector<int64_t> index(n);
        read_raw_labels(fin, file, fmt, n, index.data());
        std::string interim;
        char ch = '0';
        while (ch != '(')
        {
            if (!fin.get(ch))
                throw FoamError{file + ": no vertex label list after the face offsets"};
            interim.push_back(ch);
        }
        for (char drop : {'\n', '\r', '(', ')'}) interim.erase(std::remove(interim.begin(), interim.end(), drop), interim.end());
        const long n_labels = std::atol(interim.c_str());
        if (n_labels < 0 || size_t(n_labels) > file_bytes(file))
            throw FoamError{file + ": the vertex label count " + std::to_string(n_labels) + " exceeds the file"};
        if (index[0] != 0 || index[n - 1] != n_labels)
            throw FoamError{file + ": face offsets do not match the " + std::to_string(n_labels) + " vertex labels"};
        std::vector<int64_t> labels(size_t(n_labels), 0);
        read_raw_labels(fin, file, fmt, size_t(n_labels), labels.data());
        faces.assign(n - 1, {});
        for (size_t i = 0; i + 1 < n; ++i)
        {
            if (index[i + 1] < index[i] + 3 || index[i + 1] - index[i] > 64)
                throw FoamError{file + ": malformed face " + std::to_string(i)};
            for (int64_t k = index[i]; k < index[i + 1]; ++k) faces[i].push_back(size_t(labels[size_t(k)]));
        }
        return;
    }
    std::string line;
    std::getline(fin, line);
    faces.assign(n, {});
    for (size_t i = 0; i < n; ++i)
    {
        if (!std::getline(fin, line))
            throw FoamError{file + ": list ends early"};
        std::istringstream sl(line);
        size_t np = 0;
        sl >> np;
        const size_t l = line.find('('), r = line.find(')');
        if (l == std::string::npos || r == std::string::npos || np < 3 || np > 64)
            throw FoamError{file + ": malformed face \"" + line + "\""};
        std::istringstream iss(line.substr(l + 1, r - l - 1));
        faces[i].resize(np);
        for (size_t j = 0; j < np; ++j) iss >> faces[i][j];
    }
}
// constant/polyMesh/boundary: (nFaces, startFace, is wall) per patch (Read_Boundary, Read_Patch)
void read_boundary(const std::string& file, std::vector<std::pair<size_t, size_t>>& patches, std::vector<int>& walls)
{
    std::ifstream fin(file);
    if (!fin.is_open())
        throw FoamError{"Failed to open boundary file " + file};
    std::string line;
    while (line.find("FoamFile") == std::string::npos)
        if (!std::getline(fin, line))
            throw FoamError{file + ": no FoamFile header"};
    line.clear();
    while (line.find('}') == std::string::npos)
        if (!std::getline(fin, line))
            throw FoamError{file + ": unterminated header"};
    std::getline(fin, line);
    while (line.empty() || line.find("//") != std::string::npos)
        if (!std::getline(fin, line))
            throw FoamError{file + ": no patch list"};
    size_t n = 0;
    std::istringstream iss(line);
    iss >> n;
    for (size_t k = 0; k < n; ++k)
    {
        size_t nf = 0, sf = 0;
        int wall = 0;
        std::string name;
        /* the patch name precedes its dictionary; the opening "(" of the list and blank lines may come first */
        do
        {
            if (!std::getline(fin, name))
                throw FoamError{file + ": patch list ends early"};
        } while (name.find_first_not_of(" \t\r(") == std::string::npos);
        line.clear();
        while (line.find('}') == std::string::npos)
        {
            if (!std::getline(fin, line))
                throw FoamError{file + ": unterminated patch"};
            if (line.find("type") != std::string::npos)
                wall = line.find("wall") != std::string::npos ? 1 : 0;
            else if (line.find("nFaces") != std::string::npos)
            {
                std::istringstream is2(line);
                std::string t;
                is2 >> t >> nf;
            }
            else if (line.find("startFace") != std::string::npos)
            {
                std::istringstream is2(line);
                std::string t;
                is2 >> t >> sf;
            }
        }
        patches.emplace_back(nf, sf);
        walls.push_back(wall);
    }
}
} // namespace

extern "C" int fjsph_foam_read(const char* foam_dir, const char* solution_dir, int buoyant, double rho_fill, FjsphFoamMesh** out)
{
    if (!foam_dir || !out)
    {
        fj_set_error("foam_read: need the case directory and an output pointer");
        return FJSPH_ERR_INVALID;
    }
    try
    {
        const std::string dir = foam_dir, poly = dir + "/constant/polyMesh/";
        std::vector<std::pair<size_t, size_t>> patches;
        std::vector<int> walls, left, right;
        read_boundary(poly + "boundary", patches, walls);
        std::unique_ptr<FjsphFoamMesh> M(new FjsphFoamMesh());
        read_vectors(poly + "points", "vectorField", M->verts);
        std::vector<std::vector<size_t>> faces_;
        read_faces(poly + "faces", faces_);
        size_t n_cells = 0;
        read_labels(poly + "owner", left, n_cells);
        read_labels(poly + "neighbour", right, n_cells);
        /* Post_Process: the neighbour file stops at the internal faces; patch faces follow in patch order */
        if (right.size() != left.size())
            for (size_t k = 0; k < walls.size(); ++k)
            {
                if (patches[k].first > faces_.size()) /* a damaged nFaces entry must not size the list */
                    throw FoamError{"patch " + std::to_string(k) + " claims " + std::to_string(patches[k].first) + " of the " +
                                    std::to_string(faces_.size()) + " faces"};
                right.insert(right.end(), patches[k].first, walls[k] == 1 ? -1 : -2);
            }
        if (left.size() != faces_.size() || right.size() != faces_.size())
            throw FoamError{"Mismatch of number of faces (" + std::to_string(faces_.size()) + "), owner size (" +
                            std::to_string(left.size()) + ") and neighbour + patch size (" + std::to_string(right.size()) + ")"};
        const size_t n_pts = M->verts.size() / 3;
        if (n_cells > 2 * faces_.size()) /* every face touches at most two cells: a damaged label names a cell far outside */
            throw FoamError{"the owner / neighbour files name cell " + std::to_string(n_cells - 1) + ", the mesh has " +
                            std::to_string(faces_.size()) + " faces"};
        std::vector<std::vector<size_t>> cFaces(n_cells);
        M->face_ptr.push_back(0);
        for (size_t f = 0; f < faces_.size(); ++f)
        {
            for (size_t v : faces_[f])
                if (v >= n_pts)
                    throw FoamError{"face " + std::to_string(f) + " names point " + std::to_string(v) + " of " + std::to_string(n_pts)};
            if (left[f] < 0 || size_t(left[f]) >= n_cells || right[f] >= int(n_cells))
                throw FoamError{"face " + std::to_string(f) + " names a cell outside the mesh"};
            const size_t ntri = faces_[f].size() > 3 ? faces_[f].size() - 2 : 1;
            if (ntri > 1)
                M->n_quads_split++;
            for (size_t j = 0; j < ntri; ++j)
            {
                const size_t t = M->leftright.size() / 2;
                if (ntri > 1)
                    for (size_t v : {faces_[f][0], faces_[f][j + 1], faces_[f][j + 2]}) M->face_vtx.push_back(int64_t(v));
                else
                    for (size_t v : faces_[f]) M->face_vtx.push_back(int64_t(v));
                M->face_ptr.push_back(int64_t(M->face_vtx.size()));
                M->leftright.push_back(left[f]);
                M->leftright.push_back(right[f]);
                cFaces[size_t(left[f])].push_back(t);
                if (right[f] >= 0)
                    cFaces[size_t(right[f])].push_back(t);
            }
        }
        M->cell_ptr.push_back(0);
        M->cCentre.assign(3 * n_cells, 0.0);
        for (size_t c = 0; c < n_cells; ++c)
        {
            std::vector<size_t> verts;
            for (size_t t : cFaces[c])
            {
                M->cell_faces.push_back(int64_t(t));
                for (int64_t k = M->face_ptr[t]; k < M->face_ptr[t + 1]; ++k) verts.push_back(size_t(M->face_vtx[size_t(k)]));
            }
            M->cell_ptr.push_back(int64_t(M->cell_faces.size()));
            if (verts.empty())
                continue;
            std::sort(verts.begin(), verts.end());
            (void)std::unique(verts.begin(), verts.end()); /* no erase: FOAMIO.cpp:652-654 */
            double s[3] = {0.0, 0.0, 0.0};
            for (size_t v : verts)
                for (int d = 0; d < 3; ++d) s[d] += M->verts[3 * v + size_t(d)];
            for (int d = 0; d < 3; ++d) M->cCentre[3 * c + size_t(d)] = s[d] / double(verts.size());
        }
        if (solution_dir && solution_dir[0])
        {
            const std::string sol = dir + "/" + solution_dir;
            read_scalars(sol + (buoyant ? "/p_rgh" : "/p"), M->cP);
            read_vectors(sol + "/U", "volVectorField", M->cVel);
            if (M->cP.size() != n_cells || M->cVel.size() != 3 * n_cells)
                throw FoamError{"Mismatch between solution size (" + std::to_string(M->cP.size()) + " pressures, " +
                                std::to_string(M->cVel.size() / 3) + " velocities) and mesh size (" + std::to_string(n_cells) + ")"};
        }
        else
        {
            M->cP.assign(n_cells, 0.0);
            M->cVel.assign(3 * n_cells, 0.0);
        }
        M->cRho.assign(n_cells, rho_fill);
        *out = M.release();
        return FJSPH_OK;
    }
    catch (const FoamError& e)
    {
        fj_set_error("foam_read: %s", e.msg.c_str());
        return FJSPH_ERR_IO;
    }
    catch (const std::exception& e) /* e.g. a damaged binary header asking for more memory than there is */
    {
        fj_set_error("foam_read: %s", e.what());
        return FJSPH_ERR_IO;
    }
}

extern "C" int fjsph_foam_view(const FjsphFoamMesh* M, FjsphMesh* view)
{
    if (!M || !view)
        return FJSPH_ERR_INVALID;
    view->n_verts = int64_t(M->verts.size() / 3);
    view->verts = M->verts.data();
    view->n_faces = int64_t(M->leftright.size() / 2);
    view->face_ptr = M->face_ptr.data();
    view->face_vtx = M->face_vtx.data();
    view->leftright = M->leftright.data();
    view->n_cells = int64_t(M->cP.size());
    view->cell_ptr = M->cell_ptr.data();
    view->cell_faces = M->cell_faces.data();
    view->cCentre = M->cCentre.data();
    view->cVel = M->cVel.data();
    view->cP = M->cP.data();
    view->cRho = M->cRho.data();
    return FJSPH_OK;
}

extern "C" void fjsph_foam_free(FjsphFoamMesh* M) { delete M; }
