// host_case.cpp — host-side case front end: block files (bmap) -> shape blocks -> particles in FJSPH's order.
//
// Restates, for 2D and 3D builds selected at run time (the reference selects with -DSIMDIM):
//   read_shapes_bmap                     reference src/shapes/shapes.cpp:408-625
//   ShapeBlock::check_input / _post      reference src/shapes/shapes.cpp:21-166
//   Line/Plane, Square/Cube, Circle/Sphere, Cylinder (Hollow | Solid), Inlet (Square | Circle), Coordinates
//                                        reference src/shapes/{line,square,circle,cylinder,inlet,coordinates}.cpp
//   GetRotationMat                       reference src/Geometry.h:14-42
//   Generate_Points, Check_Intersection, Init_Particles, get_boundary_velocity
//                                        reference src/Init.cpp:26-38,61-268,270-496
// The perturbation stream is the reference's: std::default_random_engine (default seed) through
// std::uniform_real_distribution<double>(0, eps*dx), drawn in the reference's order, one engine per block.
// Reference behaviour kept on purpose: the solid cylinder's inner loop never runs (`kk > nk`, cylinder.cpp:378,490),
// local `rotmat` variables shadow the member so only "Rotation angles" rotates a block (cylinder.cpp:196-241,
// inlet.cpp:141-213), the hydrostatic initialisation always measures height along y (Init.cpp:480-493), an Arc / Arch
// block's counts default to 0, not -1, so its thickness / length / spacing fallbacks never fire and the 2D generator walks
// the block file's "i-direction count" (arc.cpp:446-495,546-554), the 3D centre + start + end form states a -90 degree arc
// whatever the end point (arc.cpp:206-245) and so generates only its straights.
// JSON block files are read too (read_json below: the slice of nlohmann/json's behaviour shapes.cpp relies on).
#include <algorithm>
#include <exception>
#include <iterator>
#include <map>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"

void fj_set_error(const char* fmt, ...);

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_4
#define M_PI_4 0.78539816339744830962
#endif

namespace
{
constexpr double DEFV = 9999999.0; /* default_val, VarDefs.h:195 */
constexpr double MEPS = std::numeric_limits<double>::epsilon();
const char* WS = " \n\r\t\f\v";

enum ShapeType /* shape_type, VarDefs.h:116-127 */
{
    linePlane = 0,
    squareCube,
    circleSphere,
    cylinderT,
    arcSection,
    coordDef,
    inletZone,
    hollowT,
    solidT
};

struct V3
{
    double c[3];
    V3() : c{0.0, 0.0, 0.0} {}
    V3(double a, double b, double d) : c{a, b, d} {}
    double& operator[](int i) { return c[i]; }
    double operator[](int i) const { return c[i]; }
};
V3 constant(double v, int dim) { return dim == 3 ? V3(v, v, v) : V3(v, v, 0.0); }
V3 operator+(const V3& a, const V3& b) { return V3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
V3 operator-(const V3& a, const V3& b) { return V3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
V3 operator*(const V3& a, double s) { return V3(a[0] * s, a[1] * s, a[2] * s); }
V3 operator*(double s, const V3& a) { return a * s; }
V3 operator/(const V3& a, double s) { return V3(a[0] / s, a[1] / s, a[2] / s); }
double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
double norm(const V3& a) { return std::sqrt(dot(a, a)); }
V3 normalized(const V3& a) /* Eigen: the zero vector stays zero */
{
    const double n2 = dot(a, a);
    return n2 > 0.0 ? a / std::sqrt(n2) : a;
}
V3 cross(const V3& a, const V3& b) { return V3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
bool same(const V3& a, const V3& b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
/* the perturbation vector; a function call so that the draws happen in the order the compiler gives the
   reference's StateVecD(unif(re), unif(re), unif(re)) constructor call */
V3 make3(double a, double b, double d) { return V3(a, b, d); }
V3 make2(double a, double b) { return V3(a, b, 0.0); }

struct M3
{
    double m[3][3];
    M3() { set_identity(); }
    void set_identity()
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m[i][j] = (i == j) ? 1.0 : 0.0;
    }
    bool is_identity() const
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                if (m[i][j] != ((i == j) ? 1.0 : 0.0))
                    return false;
        return true;
    }
};
V3 operator*(const M3& A, const V3& v)
{
    return V3(A.m[0][0] * v[0] + A.m[0][1] * v[1] + A.m[0][2] * v[2], A.m[1][0] * v[0] + A.m[1][1] * v[1] + A.m[1][2] * v[2],
              A.m[2][0] * v[0] + A.m[2][1] * v[1] + A.m[2][2] * v[2]);
}

// GetRotationMat (Geometry.h:14-42): 3D = AngleAxis(a0, X) * AngleAxis(-a1, Y) * AngleAxis(a2, Z), which Eigen
// evaluates as a quaternion product turned into a matrix (Quaternion::toRotationMatrix); 2D = [[c,-s],[s,c]].
struct Quat
{
    double w, x, y, z;
};
Quat axis_quat(double angle, int axis)
{
    Quat q{std::cos(0.5 * angle), 0.0, 0.0, 0.0};
    const double s = std::sin(0.5 * angle);
    (axis == 0 ? q.x : axis == 1 ? q.y : q.z) = s;
    return q;
}
Quat qmul(const Quat& a, const Quat& b)
{
    return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
M3 rotation_matrix(const V3& angles, int dim)
{
    M3 R;
    if (dim == 3)
    {
        const Quat q = qmul(qmul(axis_quat(angles[0], 0), axis_quat(-angles[1], 1)), axis_quat(angles[2], 2));
        const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
        const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
        const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
        R.m[0][0] = 1 - (tyy + tzz);
        R.m[0][1] = txy - twz;
        R.m[0][2] = txz + twy;
        R.m[1][0] = txy + twz;
        R.m[1][1] = 1 - (txx + tzz);
        R.m[1][2] = tyz - twx;
        R.m[2][0] = txz - twy;
        R.m[2][1] = tyz + twx;
        R.m[2][2] = 1 - (txx + tyy);
    }
    else
    {
        R.m[0][0] = std::cos(angles[0]);
        R.m[0][1] = -std::sin(angles[0]);
        R.m[1][0] = std::sin(angles[0]);
        R.m[1][1] = std::cos(angles[0]);
    }
    return R;
}

struct Block /* ShapeBlock, shapes.h:10-116 */
{
    int dim = 3;
    bool write_data = false, no_slip = false, particle_order = false;
    std::string name, shape, subshape, filename, position_filename, solver_name, inlet_bc_type;
    int bound_type = -1, sub_bound_type = -1, fixed_vel_or_dynamic = 0, bound_solver = FJSPH_PRESSURE_G;
    int ni = 0, nj = 0, nk = 0;
    size_t npts = 0, ntimes = 0;
    double insconst = DEFV, delconst = DEFV, aeroconst = DEFV;
    double dx = -1, radius = -1, arc_start = -1, arc_end = -1, arclength = DEFV, thickness = -1, length = -1, sstraight = 0,
           estraight = 0;
    double vmag = 0, press = 0, dens = 1000, nu = -1, rho_rest = 1000, gamma = 7, speedOfSound = -1, backgroundP = 0,
           renorm_vol = -1;
    V3 stretch, insert_norm{1, 0, 0}, delete_norm{1, 0, 0}, aero_norm{1, 0, 0}, normal{1, 0, 0}, angles;
    M3 rotmat;
    V3 start, end, right, mid, centre, vel;
    std::vector<size_t> back;
    std::vector<std::vector<size_t>> buffer;
    std::vector<unsigned> bc;
    std::vector<int> intersect;
    std::vector<double> times;
    std::vector<V3> pos, vels, coords;

    explicit Block(int d) : dim(d)
    {
        stretch = constant(1.0, d);
        start = end = right = mid = centre = constant(DEFV, d);
    }
    bool defined(const V3& v) const /* check_vector, Var.h:38-46 */
    {
        return v[0] != DEFV && v[1] != DEFV && (dim == 2 || v[2] != DEFV);
    }
};

struct Ctx /* what the shapes read from SIM */
{
    int dim;
    double scale, rho_rest, speed_sound, gam, press_pipe, nu;
};

std::string ltrim(const std::string& s)
{
    const size_t a = s.find_first_not_of(WS);
    return a == std::string::npos ? "" : s.substr(a);
}
std::string rtrim(const std::string& s)
{
    const size_t b = s.find_last_not_of(WS);
    return b == std::string::npos ? "" : s.substr(0, b + 1);
}
// Get_String & co (IOFunctions.h:32-131): the key is the left-trimmed text before the FIRST ':', compared exactly
bool key_value(const std::string& line, const char* key, std::string& value)
{
    const size_t pos = line.find(':');
    if (pos == std::string::npos || ltrim(line.substr(0, pos)) != key)
        return false;
    value = ltrim(rtrim(line.substr(pos + 1)));
    return true;
}
void get_string(const std::string& line, const char* key, std::string& out)
{
    std::string v;
    if (key_value(line, key, v))
        out = v;
}
template <class T>
void get_number(const std::string& line, const char* key, T& out)
{
    std::string v;
    if (key_value(line, key, v))
    {
        std::istringstream iss(v);
        T t;
        if (iss >> t)
            out = t;
    }
}
void get_vector(const std::string& line, const char* key, V3& out, int dim)
{
    std::string v;
    if (!key_value(line, key, v))
        return;
    std::istringstream iss(v);
    std::string item;
    /* IOFunctions.h:133-218: SIMDIM tokens; a missing token leaves the previous token's text in place (getline on an
       exhausted stream does not touch its string) and is read again; text that is not a number reads as 0 */
    for (int d = 0; d < dim; ++d)
    {
        std::getline(iss, item, ',');
        std::istringstream is2(item);
        double t = 0.0;
        if (!(is2 >> t))
            t = 0.0;
        out[d] = t;
    }
}

int shape_type_of(const std::string& shape, int dim)
{
    if (shape == (dim == 2 ? "Line" : "Plane"))
        return linePlane;
    if (shape == (dim == 2 ? "Square" : "Cube"))
        return squareCube;
    if (shape == (dim == 2 ? "Circle" : "Sphere"))
        return circleSphere;
    if (shape == (dim == 2 ? "Arc" : "Arch"))
        return arcSection;
    if (shape == "Cylinder")
        return cylinderT;
    if (shape == "Inlet")
        return inletZone;
    if (shape == "Coordinates")
        return coordDef;
    return -1;
}

// ---------------------------------------------------------------- common checks (shapes.cpp:21-166)
bool common_check(Block& b, const Ctx& C, std::string& err)
{
    if (C.scale != 1.0)
    {
        for (V3* v : {&b.start, &b.end, &b.right, &b.mid, &b.centre})
            if (b.defined(*v))
                *v = *v * C.scale;
        for (double* s : {&b.aeroconst, &b.insconst, &b.delconst})
            if (*s != DEFV)
                *s *= C.scale;
    }
    if (!b.position_filename.empty() && b.ntimes != 0)
    {
        if (b.times.empty())
            err += "Block \"" + b.name + "\" has no time information. ";
        if (b.pos.empty() && b.vels.empty())
            err += "Block \"" + b.name + "\" position or velocity data has not been ingested properly. ";
    }
    if (b.bound_type == -1)
        err += "Block \"" + b.name + "\" shape has not been correctly defined. ";
    if (!b.solver_name.empty())
    {
        if (b.solver_name == "DBC")
            b.bound_solver = FJSPH_DBC;
        else if (b.solver_name == "Pressure-Gradient")
            b.bound_solver = FJSPH_PRESSURE_G;
        else if (b.solver_name == "Ghost")
            b.bound_solver = FJSPH_GHOST;
        /* anything else: the reference warns and keeps Pressure-Gradient */
    }
    if (!b.inlet_bc_type.empty()) /* only the JSON reader fills it (shapes.cpp:104-121,292); anything else keeps the default */
    {
        if (b.inlet_bc_type == "Fixed Velocity")
            b.fixed_vel_or_dynamic = 0;
        else if (b.inlet_bc_type == "Dynamic")
            b.fixed_vel_or_dynamic = 1;
    }
    if (b.press != 0)
    {
        const double Bc = C.rho_rest * (C.speed_sound * C.speed_sound) / C.gam;
        b.dens = std::pow(((b.press - C.press_pipe) / Bc + 1.0), C.gam) * C.rho_rest; /* sic: exponent gam, shapes.cpp:151-156 */
    }
    else
        b.dens = C.rho_rest;
    if (b.nu < 0)
        b.nu = C.nu;
    return err.empty();
}
void post_check(Block& b, double& gs)
{
    gs = std::max(b.dx, gs);
    b.npts = b.npts > 1 ? b.npts : 1;
}
int ceil_i(double v) { return static_cast<int>(std::ceil(v)); }

// rotation set-up shared by Cylinder and Inlet: only "Rotation angles" reaches the member matrix
void rotation_from_angles(Block& b, bool is_inlet)
{
    if (norm(b.angles) != 0)
    {
        b.angles = b.angles * (M_PI / 180.0);
        b.rotmat = rotation_matrix(b.angles, b.dim);
        b.normal = b.rotmat * V3(1, 0, 0);
        if (is_inlet)
            b.insert_norm = b.normal;
    }
    else if (!same(b.normal, V3(1, 0, 0)))
    {
        b.normal = normalized(b.normal); /* the matrix built here is a shadowing local in the reference */
        if (b.dim == 2)
            b.angles[0] = std::atan2(b.normal[1], b.normal[0]);
        if (is_inlet)
            b.insert_norm = b.normal;
    }
    else if (is_inlet && !same(b.insert_norm, V3(1, 0, 0)))
    {
        b.normal = normalized(b.insert_norm);
        if (b.dim == 2)
            b.angles[0] = std::atan2(b.normal[1], b.normal[0]) - M_PI / 2.0;
    }
}

// ---------------------------------------------------------------- Line / Plane (line.cpp)
void line_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    if (!b.defined(b.start))
        err += "Block \"" + b.name + "\" starting position has not been correctly defined. ";
    if (!b.defined(b.end))
        err += "Block \"" + b.name + "\" ending position has not been correctly defined. ";
    if (b.dim == 3 && !b.defined(b.right))
        err += "Block \"" + b.name + "\" right position has not been correctly defined. ";
    if (!err.empty())
        return;
    if (b.dx < 0 && (b.ni < 0 || (b.dim == 3 && b.nj < 0)))
        b.dx = gs;
    if (b.dim == 2)
    {
        if (b.ni > 0 && b.dx < 0)
            b.dx = norm(b.end - b.start) / double(b.ni);
    }
    else if (b.ni > 0 && b.nj > 0)
        b.dx = std::min(norm(b.right - b.start) / double(b.ni), norm(b.end - b.right) / double(b.nj));
    if (b.thickness < 0)
    {
        if (b.nk < 0)
            err += "Block \"" + b.name + "\" line thickness has not been correctly defined. ";
    }
    else if (b.nk < 0)
        b.nk = b.particle_order ? ceil_i(b.thickness / b.dx / std::sqrt(3.0) * 2.0) : ceil_i(b.thickness / b.dx);
    b.nk = b.nk > 1 ? b.nk : 1;
    if (b.ni > 0)
    {
        if (b.dim == 2)
            b.npts = size_t(b.ni) * b.nk;
        else
        {
            if (b.nj < 1)
            {
                b.nj = ceil_i(norm(b.end - b.right) / gs);
                b.nj = b.nj > 1 ? b.nj : 1;
            }
            b.npts = size_t(b.ni) * b.nj * b.nk;
        }
    }
    else if (b.dim == 2)
    {
        b.ni = ceil_i(norm(b.end - b.start) / gs);
        b.ni = b.ni > 1 ? b.ni : 1;
        b.npts = size_t(b.ni) * b.nk;
    }
    else
    {
        b.ni = ceil_i(norm(b.right - b.start) / gs);
        b.nj = b.particle_order ? ceil_i(norm(b.end - b.right) / gs / std::sqrt(3.0) * 2.0) : ceil_i(norm(b.end - b.right) / gs);
        b.ni = b.ni > 1 ? b.ni : 1;
        b.nj = b.nj > 1 ? b.nj : 1;
        b.npts = size_t(b.ni) * b.nj * b.nk;
    }
    post_check(b, gs);
}
void line_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    std::vector<V3>& pts = b.coords;
    if (b.dim == 2)
    {
        const V3 delta = normalized(b.end - b.start) * gs;
        const V3 nrm(delta[1], -delta[0], 0.0);
        for (int ii = 0; ii < b.ni; ++ii)
            for (int jj = 0; jj < b.nk; ++jj)
            {
                V3 p = b.particle_order ? delta * double(ii + 0.5 * (jj % 2)) + 0.5 * nrm * std::sqrt(3.0) * double(jj)
                                        : delta * double(ii) + nrm * double(jj);
                p = p + constant(unif(re), 2);
                pts.push_back(p + b.start);
            }
        return;
    }
    const V3 di = normalized(b.right - b.start) * gs;
    const V3 dj = normalized(b.end - b.right) * gs;
    const V3 nrm = normalized(cross(dj, di)) * gs;
    for (int jj = 0; jj < b.nj; ++jj)
        for (int ii = 0; ii < b.ni; ++ii)
            for (int kk = 0; kk < b.nk; ++kk)
            {
                V3 p = b.particle_order ? 0.5 * (di * double(2 * ii + ((jj + kk) % 2)) +
                                                 dj * (std::sqrt(3.0) * (double(jj) + double(kk % 2) / 3)) +
                                                 nrm * 2 * std::sqrt(6.0) / 3.0 * double(kk))
                                        : di * double(ii) + dj * double(jj) + nrm * double(kk);
                p = p + constant(unif(re), 3);
                pts.push_back(p + b.start);
            }
}

// ---------------------------------------------------------------- Square / Cube (square.cpp)
void lattice_counts(const Block& b, double gs, int& ni, int& nj, int& nk)
{
    if (b.particle_order)
    {
        ni = ceil_i((b.end[0] - b.start[0]) / gs);
        nj = ceil_i((b.end[1] - b.start[1]) / gs / std::sqrt(3.0) * 2.0);
        nk = b.dim == 3 ? ceil_i((b.end[2] - b.start[2]) / gs / std::sqrt(6.0) * 3.0) : 1;
    }
    else
    {
        ni = ceil_i((b.end[0] - b.start[0]) / gs);
        nj = ceil_i((b.end[1] - b.start[1]) / gs);
        nk = b.dim == 3 ? ceil_i((b.end[2] - b.start[2]) / gs) : 1;
    }
    ni = ni > 1 ? ni : 1;
    nj = nj > 1 ? nj : 1;
    nk = nk > 1 ? nk : 1;
}
V3 lattice_point(const Block& b, int i, int j, int k)
{
    if (b.dim == 2)
        return b.particle_order ? 0.5 * V3(double(2 * i + (j % 2)), std::sqrt(3.0) * double(j), 0.0) : V3(double(i), double(j), 0.0);
    return b.particle_order ? 0.5 * V3(double(2 * i + ((j + k) % 2)), std::sqrt(3.0) * (double(j) + double(k % 2) / 3.0),
                                       2.0 / 3.0 * std::sqrt(6.0) * double(k))
                            : V3(double(i), double(j), double(k));
}
void square_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    if (!b.defined(b.start))
        err += "Block \"" + b.name + "\" starting position has not been correctly defined. ";
    if (!b.defined(b.end))
        err += "Block \"" + b.name + "\" ending position has not been correctly defined. ";
    if (!err.empty())
        return;
    if (b.dx < 0 && (b.ni < 0 || b.nj < 0))
        b.dx = gs;
    if (b.ni > 0 && b.nj > 0 && (b.dim == 2 || b.nk > 0) && b.dx < 0)
    {
        const V3 dist = b.end - b.start;
        const double di = dist[0] / double(b.ni), dj = dist[1] / double(b.nj);
        b.dx = b.dim == 2 ? std::min(di, dj) : std::min(di, std::min(dj, dist[1] / double(b.nk))); /* sic: dist[1], square.cpp:54 */
    }
    lattice_counts(b, gs, b.ni, b.nj, b.nk);
    b.npts = size_t(b.ni) * b.nj * (b.dim == 3 ? b.nk : 1);
    post_check(b, gs);
}
void square_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    const int nk = b.dim == 3 ? b.nk : 1;
    for (int k = 0; k < nk; ++k)
        for (int j = 0; j < b.nj; ++j)
            for (int i = 0; i < b.ni; ++i)
            {
                V3 p = lattice_point(b, i, j, k) * gs;
                p = p + constant(unif(re), b.dim);
                b.coords.push_back(p + b.start);
            }
}

// ---------------------------------------------------------------- Circle / Sphere (circle.cpp)
void circle_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    if (!b.defined(b.centre))
        err += "Block \"" + b.name + "\" centre position has not been correctly defined. ";
    if (b.radius < 0)
        err += "Block \"" + b.name + "\" radius has not been correctly defined. ";
    if (!err.empty())
        return;
    b.start = b.centre - constant(b.radius, b.dim);
    b.end = b.centre + constant(b.radius, b.dim);
    if (b.dx < 0)
        b.dx = (b.ni < 0) ? gs : (2.0 * b.radius) / double(b.ni);
    int ni, nj, nk;
    lattice_counts(b, gs, ni, nj, nk); /* local counts: the member ni, nj, nk are left alone (circle.cpp:46-81) */
    long np = long(ni) * nj * (b.dim == 3 ? nk : 1);
    np = np > 1 ? np : 1;
    b.npts = size_t(std::ceil(double(np) * (b.dim == 2 ? M_PI / 4.0 : M_PI / 6.0)));
    post_check(b, gs);
}
void ring_points(std::vector<V3>& pts, std::default_random_engine& re, std::uniform_real_distribution<double>& unif, double gs,
                 double radius, double x_layer, const M3& rot, const V3& centre, int dim)
{
    /* concentric rings (circle.cpp:98-129 in 2D, inlet.cpp create_radial_disk in 3D) */
    for (double rad = radius; rad > 0.99 * gs; rad -= gs)
    {
        double dtheta = std::atan(gs / rad);
        const int ncirc = int(std::floor(std::fabs(2.0 * M_PI / dtheta)));
        dtheta = 2.0 * M_PI / double(ncirc);
        for (double theta = 0.0; theta < 2 * M_PI - 0.5 * dtheta; theta += dtheta)
        {
            V3 p = dim == 2 ? V3(rad * std::sin(theta), rad * std::cos(theta), 0.0)
                            : V3(x_layer, rad * std::sin(theta), rad * std::cos(theta));
            p = p + (dim == 2 ? make2(unif(re), unif(re)) : make3(unif(re), unif(re), unif(re)));
            pts.push_back(rot * p + centre);
        }
    }
    V3 p = dim == 2 ? V3() : V3(x_layer, 0.0, 0.0);
    p = p + (dim == 2 ? make2(unif(re), unif(re)) : make3(unif(re), unif(re), unif(re)));
    pts.push_back(rot * p + centre);
}
void circle_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    if (b.dim == 2)
    {
        ring_points(b.coords, re, unif, gs, b.radius, 0.0, b.rotmat, b.centre, 2);
        return;
    }
    const double rsq = b.radius * b.radius;
    int ni, nj, nk;
    lattice_counts(b, gs, ni, nj, nk);
    for (int k = 0; k < nk; ++k)
        for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i)
            {
                V3 p = b.rotmat * (lattice_point(b, i, j, k) * gs);
                p = p + make3(unif(re), unif(re), unif(re));
                p = p + b.start;
                const V3 d = p - b.centre;
                if (dot(d, d) > rsq)
                    continue;
                b.coords.push_back(p);
            }
}

// ---------------------------------------------------------------- Arc / Arch (arc.cpp)
// x of A x = b for a 2x2 A as Eigen's colPivHouseholderQr().solve() computes it (arc.cpp:73,230): the column of larger
// norm first, one Householder reflection, back-substitution over the pivots above the rank threshold.
void qr_solve2(const double A[2][2], const double rhs[2], double x[2])
{
    const double eps = std::numeric_limits<double>::epsilon();
    double q[2][2] = {{A[0][0], A[0][1]}, {A[1][0], A[1][1]}};
    double c[2] = {rhs[0], rhs[1]};
    int perm[2] = {0, 1};
    const double n0 = q[0][0] * q[0][0] + q[1][0] * q[1][0], n1 = q[0][1] * q[0][1] + q[1][1] * q[1][1];
    const double maxcol = std::sqrt(std::max(n0, n1));
    const double thr = (maxcol * eps / 2.0) * (maxcol * eps / 2.0);
    int nzp = 2;
    if (n1 > n0)
    {
        std::swap(q[0][0], q[0][1]);
        std::swap(q[1][0], q[1][1]);
        std::swap(perm[0], perm[1]);
    }
    if (std::max(n0, n1) < thr * 2.0)
        nzp = 0;
    /* reflection that zeroes q[1][0] */
    const double c0 = q[0][0], tail = q[1][0] * q[1][0];
    double beta = c0, tau = 0.0, ess = 0.0;
    if (tail > (std::numeric_limits<double>::min)())
    {
        beta = std::sqrt(c0 * c0 + tail);
        if (c0 >= 0.0)
            beta = -beta;
        ess = q[1][0] / (c0 - beta);
        tau = (beta - c0) / beta;
    }
    q[0][0] = beta;
    {
        const double t = (q[0][1] + ess * q[1][1]) * tau;
        q[0][1] -= t;
        q[1][1] -= t * ess;
    }
    if (nzp == 2 && q[1][1] * q[1][1] < thr)
        nzp = 1;
    if (nzp > 0)
    {
        const double t = (c[0] + ess * c[1]) * tau;
        c[0] -= t;
        c[1] -= t * ess;
    }
    /* the second reflection of a 2x2 is the identity (no tail) */
    double y[2] = {0.0, 0.0};
    if (nzp == 2)
    {
        y[1] = c[1] / q[1][1];
        y[0] = (c[0] - q[0][1] * y[1]) / q[0][0];
    }
    else if (nzp == 1)
        y[0] = c[0] / q[0][0];
    x[perm[0]] = y[0];
    x[perm[1]] = y[1];
}

// get_arclength_centrepoint (arc.cpp:27-57 in 2D, 176-250 in 3D): radius and the two angles, in degrees
bool arc_from_centre(Block& b, std::string& err)
{
    const V3 d1 = b.start - b.centre, d2 = b.end - b.centre;
    const double r = dot(d1, d1);
    if (std::fabs(dot(d2, d2) - r) > 0.001)
    {
        err += "Block \"" + b.name + "\": arc points are not correctly defined, the ending radius differs from the starting radius. ";
        return false;
    }
    b.radius = std::sqrt(r);
    if (b.dim == 2)
    {
        b.arc_start = std::atan2(d1[1], d1[0]) * 180 / M_PI;
        b.arc_end = std::atan2(d2[1], d2[0]) * 180 / M_PI;
        return true;
    }
    /* 3D: coordinates of the end point in the plane basis (d1, v).  As in the reference, the right-hand side of the
       2x2 system is v itself (so the solution is (0, 1) whatever the end point), and it keeps its x and y components
       when another pair of axes is chosen for the matrix: the stated angles are 90 and atan2(a, b) degrees. */
    const V3 e1 = normalized(d1), e2 = normalized(d2);
    const V3 w = normalized(cross(e2, e1));
    const V3 v = normalized(cross(w, e1));
    const double maxx = std::fabs(e1[0]) + std::fabs(v[0]), maxy = std::fabs(e1[1]) + std::fabs(v[1]),
                 maxz = std::fabs(e1[2]) + std::fabs(v[2]);
    double m[2][2] = {{e1[0], v[0]}, {e1[1], v[1]}};
    const double rhs[2] = {v[0], v[1]};
    if (maxx < maxy && maxx < maxz)
    {
        m[0][0] = e1[1], m[0][1] = v[1], m[1][0] = e1[2], m[1][1] = v[2];
    }
    else if (maxy < maxx && maxy < maxz)
    {
        m[0][0] = e1[0], m[0][1] = v[0], m[1][0] = e1[2], m[1][1] = v[2];
    }
    double ab[2];
    qr_solve2(m, rhs, ab);
    b.arc_start = std::atan2(1.0, 0.0) * 180 / M_PI;
    b.arc_end = std::atan2(ab[0], ab[1]) * 180 / M_PI;
    b.right = w;
    return true;
}

void arc_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    const int dim = b.dim;
    bool has_config = false, arc_defined = false;
    if (dim == 2 && b.defined(b.centre))
    {
        if (b.arc_start >= 0 && b.arc_end >= 0 && b.radius > 0)
            has_config = true;
        else if (b.arc_start >= 0 && b.ni > 0 && b.radius > 0)
            has_config = true;
        else if (b.arc_start >= 0 && b.arclength != DEFV && b.radius > 0)
            has_config = arc_defined = true;
    }
    if (!has_config && b.defined(b.centre) && b.defined(b.start) && b.defined(b.end))
    {
        has_config = true;
        if (!arc_from_centre(b, err))
            return;
    }
    if (!has_config && dim == 3)
    {
        if (b.defined(b.centre) && b.defined(b.start) && b.defined(b.mid) && b.arclength != DEFV)
            has_config = arc_defined = true;
    }
    else if (!has_config && b.defined(b.start) && b.defined(b.end) && b.defined(b.mid))
    {
        /* 2D, three points on the arc (get_arclength_midpoint, arc.cpp:59-78): the centre from the two chord equations */
        has_config = true;
        const V3 &s = b.start, &e = b.end, &mp = b.mid;
        const V3 r1(s[0] * s[0], s[1] * s[1], 0), r2(mp[0] * mp[0], mp[1] * mp[1], 0), r3(e[0] * e[0], e[1] * e[1], 0);
        double m[2][2] = {{(s[0] - e[0]), (s[1] - e[1])}, {(s[0] - mp[0]), (s[1] - mp[1])}};
        const double rhs[2] = {r1[0] - r3[0] + r1[1] - r3[1], r1[0] - r2[0] + r1[1] - r2[1]};
        for (auto& row : m)
            for (double& a : row) a *= 2.0;
        double cxy[2];
        qr_solve2(m, rhs, cxy);
        b.centre = V3(cxy[0], cxy[1], 0.0);
        if (!arc_from_centre(b, err))
            return;
    }
    if (!has_config && b.defined(b.centre) && b.defined(b.start) && b.arclength != DEFV)
    {
        if (dim == 2 || b.defined(b.right))
            has_config = arc_defined = true;
    }
    if (!has_config)
        err += "Block \"" + b.name + "\" arc geometry has not been sufficiently defined. ";
    if (b.dx < 0)
        b.dx = (b.ni < 0) ? gs : (2.0 * b.radius) / double(b.ni);
    if (b.thickness < 0)
    {
        if (b.nk < 0)
            err += "Block \"" + b.name + "\" arc thickness has not been correctly defined. ";
    }
    else if (b.nk < 0)
        b.nk = b.particle_order ? ceil_i(b.thickness / gs / std::sqrt(3.0) * 2.0) : ceil_i(b.thickness / gs);
    if (!err.empty())
        return;

    const double dtheta = gs / b.radius;
    if (arc_defined)
    {
        if (dim == 3)
        {
            /* get_arc_end (arc.cpp:149-174): the radius from the start point; the start must lie in the arch plane */
            const V3 u = b.start - b.centre;
            b.radius = norm(u);
            if (std::fabs(dot(b.right, b.start) - dot(b.right, b.centre)) > 1e-4)
            {
                err += "Block \"" + b.name + "\": points do not exist upon the plane defined by the provided normal. ";
                return;
            }
            const V3 uu = u / b.radius, w = normalized(b.right);
            const V3 v = normalized(cross(uu, w));
            const double alen = b.arclength * M_PI / 180.0;
            b.end = b.radius * (std::cos(alen) * uu + std::sin(alen) * v) + b.centre;
        }
    }
    else
    {
        if (b.arc_end < 0 && b.ni > 0)
            b.arc_end = 180.0 / M_PI * (b.arc_start * M_PI / 180 + double(b.ni) * dtheta);
        b.arclength = b.arc_end - b.arc_start;
    }
    /* a local count: the member ni (what the 2D generator walks) stays what the block file gave (arc.cpp:493-495) */
    int ni = ceil_i((std::fabs(b.arclength) * M_PI / 180) / dtheta);
    ni = ni > 1 ? ni : 1;
    const int smax = b.sstraight > 0 ? ceil_i(b.sstraight / gs) : 0;
    const int emax = b.estraight > 0 ? ceil_i(b.estraight / gs) : 0;
    long np = long(ni + smax + emax) * b.nk;
    if (dim == 3)
    {
        if (b.length < 0)
        {
            if (b.nj < 0)
            {
                err += "Block \"" + b.name + "\" arch length has not been correctly defined. ";
                return;
            }
        }
        else if (b.nj < 0)
        {
            b.nj = b.particle_order ? ceil_i(b.length / gs / std::sqrt(6.0) * 3.0) : ceil_i(b.length / gs);
            b.nj = b.nj > 1 ? b.nj : 1;
        }
        np *= b.nj;
    }
    b.npts = np > 0 ? size_t(np) : 0;
    post_check(b, gs);
}

void arc_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    const double dtheta = gs / b.radius;
    const int smax = ceil_i(b.sstraight / gs), emax = ceil_i(b.estraight / gs);
    if (b.dim == 2)
    {
        /* make_arc (arc.cpp:80-145): nk rings inwards from the radius, each a start straight, ni arc points, an end straight */
        const double theta0 = b.arc_start * M_PI / 180.0;
        const int ni = b.ni, nk = b.nk;
        const V3 svec(std::cos(theta0), std::sin(theta0), 0.0), snormal(svec[1], -svec[0], 0.0);
        const V3 evec(std::cos(theta0 + double(ni - 1) * dtheta), std::sin(theta0 + double(ni - 1) * dtheta), 0.0);
        const V3 enormal(-evec[1], evec[0], 0.0);
        for (int jj = 0; jj < nk; ++jj)
        {
            if (b.sstraight > 0)
                for (int ii = smax; ii > 0; --ii)
                {
                    V3 p = svec * (b.radius - double(jj) * gs) + snormal * double(ii) * gs;
                    p = p + (make2(unif(re), unif(re)) + b.centre);
                    b.coords.push_back(p);
                }
            for (int ii = 0; ii < ni; ++ii)
            {
                double theta = theta0 + double(ii) * dtheta, r = b.radius;
                if (b.particle_order)
                {
                    theta += 0.5 * dtheta * (jj % 2);
                    r -= 0.5 * std::sqrt(3.0) * double(jj) * gs;
                }
                else
                    r -= double(jj) * gs;
                V3 p(std::cos(theta) * r, std::sin(theta) * r, 0.0);
                p = p + make2(unif(re), unif(re));
                p = p + b.centre;
                b.coords.push_back(p);
            }
            if (b.estraight > 0)
                for (int ii = 1; ii <= emax; ++ii)
                {
                    V3 p = evec * (b.radius - double(jj) * gs) + enormal * double(ii) * gs;
                    p = p + (make2(unif(re), unif(re)) + b.centre);
                    b.coords.push_back(p);
                }
        }
        return;
    }
    /* make_arch (arc.cpp:280-384): plane basis (u, v) with u towards the start point, w along the archway; nj slices of
       nk layers outwards from the radius.  The arc count comes from the signed arc length (arc.cpp:567-569). */
    const int nrad = ceil_i((b.arclength * M_PI / 180.0) / dtheta), nthick = b.nk, nlong = b.nj;
    const V3 u = normalized(b.start - b.centre), w = normalized(b.right);
    const V3 v = normalized(cross(u, w));
    const V3 evec = std::cos(double(nrad - 1) * dtheta) * u + std::sin(double(nrad - 1) * dtheta) * v;
    const V3 enorm = cross(evec, w);
    for (int kk = 0; kk < nlong; ++kk)
        for (int jj = 0; jj < nthick; ++jj)
        {
            double r = b.radius, l = double(kk) * gs, doffset = 0;
            if (b.particle_order)
            {
                r += 1.0 / 3.0 * std::sqrt(6.0) * double(jj) * gs;
                l = 0.5 * std::sqrt(3.0) * (double(kk) + double(jj % 2) / 3.0) * gs;
                doffset = 0.5 * dtheta * ((jj + kk) % 2);
            }
            else
                r += double(jj) * gs;
            if (b.sstraight > 0)
                for (int ii = smax; ii > 0; --ii)
                {
                    const double dist = double(ii) * gs + doffset;
                    V3 p = u * r - v * dist + l * w;
                    p = p + (make3(unif(re), unif(re), unif(re)) + b.centre);
                    b.coords.push_back(p);
                }
            for (int ii = 0; ii < nrad; ++ii)
            {
                double theta = double(ii) * dtheta;
                const double la = double(kk) * gs; /* the arc keeps the grid offset along w in HCP order too (arc.cpp:352) */
                if (b.particle_order)
                    theta += doffset;
                const double a = std::cos(theta) * r, bb = std::sin(theta) * r;
                V3 p = a * u + bb * v + la * w;
                p = p + make3(unif(re), unif(re), unif(re));
                p = p + b.centre;
                b.coords.push_back(p);
            }
            if (b.estraight > 0)
                for (int ii = 1; ii <= emax; ++ii)
                {
                    const double dist = double(ii) * gs + doffset;
                    V3 p = evec * r + enorm * dist + l * w;
                    p = p + (make3(unif(re), unif(re), unif(re)) + b.centre);
                    b.coords.push_back(p);
                }
        }
}

// ---------------------------------------------------------------- Cylinder (cylinder.cpp)
void cylinder_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    const bool d3 = b.dim == 3;
    int has_config = 0;
    auto centre_radius_frame = [&]() {
        b.start = b.centre;
        b.start[1] -= b.radius;
        b.end = b.centre;
        b.end[1] += b.radius;
        if (d3)
        {
            b.start[2] -= b.radius;
            b.end[2] += b.radius;
            b.right = b.centre;
            b.right[1] += b.radius;
            b.right[2] -= b.radius;
        }
    };
    auto centre_right_end_frame = [&]() {
        const V3 v = b.right - b.centre, u = b.end - b.right;
        b.start = b.centre - v - u;
        b.right = b.start + 2.0 * v;
        b.end = b.right + 2.0 * u;
        b.radius = norm(v);
    };
    if (b.subshape == "Hollow")
    {
        b.sub_bound_type = hollowT;
        if (b.thickness < 0 && b.nk < 0)
            err += "Cylinder block \"" + b.name + "\" thickness or wall count has not been defined. ";
        if (b.length < 0 && b.nj < 0)
            err += "Cylinder block \"" + b.name + "\" length or j-count has not been defined. ";
        if (b.defined(b.centre))
        {
            if (b.radius > 0)
            {
                has_config = 3;
                centre_radius_frame();
            }
            else if (d3 && b.defined(b.right) && b.defined(b.end))
            {
                has_config = 4;
                centre_right_end_frame();
            }
        }
        else if (b.defined(b.start) && b.defined(b.end) && (!d3 || b.defined(b.right)))
        {
            has_config = 1;
            b.centre = 0.5 * (b.start + b.end);
            const V3 v = b.right - b.start;
            b.radius = 0.5 * norm(v);
            b.right = b.centre + 0.5 * v;
        }
    }
    else if (b.subshape == "Solid")
    {
        b.sub_bound_type = solidT;
        if (b.length < 0 || b.nk < 0)
            err += "Cylinder block \"" + b.name + "\" length or k-count has not been defined. ";
        if (b.defined(b.start) && b.defined(b.right) && b.defined(b.end))
        {
            has_config = 1;
            b.centre = 0.5 * (b.start + b.end);
            b.radius = 0.5 * norm(b.right - b.start);
        }
        else if (b.defined(b.centre))
        {
            if (b.radius > 0)
            {
                has_config = 3;
                centre_radius_frame();
            }
            if (b.defined(b.right) && b.defined(b.end))
            {
                has_config = 4;
                centre_right_end_frame();
            }
        }
        else if (b.defined(b.start) && b.ni > 0 && b.nj > 0)
            has_config = 2;
    }
    else
        err += "Cylinder block \"" + b.name + "\" subtype not defined appropriately: choose Hollow or Solid. ";
    if (has_config == 0)
        err += "Cylinder block \"" + b.name + "\" geometry has not been sufficiently defined. ";
    if (!err.empty())
        return;
    rotation_from_angles(b, false);
    if (has_config == 1)
    {
        const V3 ab = normalized(b.end - b.start);
        if (d3)
            b.normal = normalized(cross(ab, normalized(b.right - b.start)));
        else
        {
            b.angles[0] = std::atan2(ab[1], ab[0]);
            /* normal = R * (1,0) with the local R = [[c, s], [-s, c]] */
            b.normal = V3(std::cos(b.angles[0]), -std::sin(b.angles[0]), 0.0);
        }
    }
    if (has_config == 2)
    {
        b.centre = b.start;
        b.radius = 0.5 * gs * b.ni;
        b.centre[1] += b.radius;
        if (d3)
        {
            b.right = b.start;
            b.right[1] += gs * b.ni;
            b.end = b.right;
            b.end[2] += gs * b.nj;
            b.centre[2] += b.radius;
        }
        else
        {
            b.end = b.start;
            b.end[1] += gs * b.ni;
        }
        if (!b.rotmat.is_identity())
        {
            b.centre = b.rotmat * (b.centre - b.start) + b.start;
            b.end = b.rotmat * (b.end - b.start) + b.start;
            if (d3)
                b.right = b.rotmat * (b.right - b.start) + b.start;
        }
    }
    else if ((has_config == 3 || has_config == 4) && !b.rotmat.is_identity())
    {
        b.start = b.rotmat * (b.start - b.centre) + b.centre;
        b.end = b.rotmat * (b.end - b.centre) + b.centre;
        if (d3)
            b.right = b.rotmat * (b.right - b.centre) + b.centre;
    }
    b.dx = gs;
    if (b.sub_bound_type == hollowT)
    {
        if (b.thickness < 0)
            b.thickness = b.particle_order ? double(b.nk) * gs * std::sqrt(3.0) * 2.0 : double(b.nk) * gs;
        else if (b.nk < 0)
        {
            b.nk = b.particle_order ? ceil_i(b.thickness / (gs * std::sqrt(3.0) * 2.0)) : ceil_i(b.thickness / gs);
            b.nk = b.nk > 1 ? b.nk : 1;
        }
        if (d3)
        {
            const double dtheta = gs / b.radius;
            b.ni = ceil_i((2 * M_PI) / dtheta);
            b.ni = b.ni > 1 ? b.ni : 1;
        }
        b.nj = int(std::ceil(b.length / gs) + 1);
        b.nj = b.nj > 1 ? b.nj : 1;
        b.npts = size_t(b.nj) * b.nk * (d3 ? b.ni : 2);
    }
    else
    {
        b.ni = int(std::ceil(2.0 * b.radius / gs));
        b.nj = int(std::ceil(b.length / gs));
        if (d3)
            b.nk = int(std::ceil(2.0 * b.radius / gs));
        b.ni = b.ni > 1 ? b.ni : 1;
        b.nj = b.nj > 1 ? b.nj : 1;
        b.nk = b.nk > 1 ? b.nk : 1;
        b.npts = d3 ? size_t(b.nj * b.nk * M_PI_4 * b.ni) * b.ni : size_t(b.ni) * b.nj;
    }
    post_check(b, gs);
}
void cylinder_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    if (b.sub_bound_type != hollowT)
        return; /* Solid: the reference's loops `for (ii = 0; ii > ni; ++ii)` / `for (kk = 0; kk > nk; ++kk)` never run */
    if (b.dim == 2)
    {
        const V3 nrm = normalized(b.normal);
        const V3 left(nrm[1], -nrm[0], 0.0);
        const double r = b.radius;
        for (int side = 0; side < 2; ++side)
        {
            const double sg = side == 0 ? 1.0 : -1.0;
            for (int ii = 0; ii < b.nj; ii++)
                for (int kk = 0; kk < b.nk; kk++)
                {
                    V3 p = b.particle_order ? gs * (nrm * double(-ii + 0.5 * (kk % 2)) + sg * 0.5 * left * std::sqrt(3.0) * double(kk)) +
                                                  sg * r * left
                                            : gs * (nrm * double(-ii) + sg * left * double(kk)) + sg * r * left;
                    p = p + make2(unif(re), unif(re));
                    b.coords.push_back(p + b.centre);
                }
        }
        return;
    }
    const double dtheta = 2 * M_PI / double(b.ni);
    for (int jj = 0; jj < b.nj; jj++)
        for (int kk = 0; kk < b.nk; kk++)
        {
            double r = b.radius, l = double(jj) * gs, doffset = 0;
            if (b.particle_order)
            {
                r += 1.0 / 3.0 * std::sqrt(6.0) * double(kk) * gs;
                l = 0.5 * std::sqrt(3) * (double(jj) + double(kk % 2) / 3.0) * gs;
                doffset = 0.5 * dtheta * ((kk + jj) % 2);
            }
            else
                r += double(kk) * gs;
            for (int ii = 0; ii < b.ni; ii++)
            {
                double theta = double(ii) * dtheta;
                if (b.particle_order)
                    theta += doffset;
                V3 p(-l, std::cos(theta) * r, std::sin(theta) * r);
                p = p + make3(unif(re), unif(re), unif(re));
                b.coords.push_back(b.rotmat * p + b.centre);
            }
        }
}

// ---------------------------------------------------------------- Inlet (inlet.cpp:7-305, 307-575)
void inlet_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    const bool d3 = b.dim == 3;
    int has_config = 0;
    if (d3)
    {
        if (b.subshape == "Square")
        {
            b.sub_bound_type = squareCube;
            if (b.defined(b.start) && b.defined(b.right) && b.defined(b.end))
                has_config = 1;
            if (b.defined(b.start) && b.ni > 0 && b.nj > 0)
                has_config = 2;
        }
        else if (b.subshape == "Circle")
        {
            b.sub_bound_type = circleSphere;
            if (b.defined(b.centre))
            {
                if (b.defined(b.right) && b.defined(b.end))
                {
                    has_config = 4;
                    const V3 v = b.right - b.centre, u = b.end - b.right;
                    b.start = b.centre - v - u;
                    b.right = b.start + 2.0 * v;
                    b.end = b.right + 2.0 * u;
                }
                if (b.radius > 0)
                {
                    has_config = 3;
                    b.start = b.centre;
                    b.start[1] -= b.radius;
                    b.start[2] -= b.radius;
                    b.end = b.centre;
                    b.end[1] += b.radius;
                    b.end[2] += b.radius;
                    b.right = b.centre;
                    b.right[1] += b.radius;
                    b.right[2] -= b.radius;
                }
            }
            else if (b.defined(b.start) && b.defined(b.right) && b.defined(b.end))
            {
                has_config = 1;
                b.centre = 0.5 * (b.start + b.end);
                const V3 v = b.right - b.start;
                b.radius = 0.5 * norm(v);
                b.right = b.centre + 0.5 * v;
            }
        }
        else
            err += "Inlet block \"" + b.name + "\" sub shape type has not been correctly defined: choose Square or Circle. ";
    }
    else if (b.defined(b.start) && b.defined(b.end))
    {
        has_config = 1;
        b.centre = 0.5 * (b.start + b.end);
        b.radius = norm(b.centre - b.start);
    }
    else if (b.defined(b.centre) && b.radius > 0)
    {
        has_config = 3;
        b.start = b.centre;
        b.start[1] -= b.radius;
        b.end = b.centre;
        b.end[1] += b.radius;
    }
    if (has_config == 0)
        err += "Inlet block \"" + b.name + "\" geometry not sufficiently defined. ";
    if (b.particle_order)
        err += "Inlet block \"" + b.name + "\": HCP ordering makes the reference index its 4-entry buffer rows with 5 "
               "(Init.cpp:358 against inlet.cpp:491); not supported. ";
    if (!err.empty())
        return;
    b.dx = gs;
    rotation_from_angles(b, true);
    if (has_config == 1)
    {
        const V3 ab = normalized(b.end - b.start);
        if (d3)
            b.normal = normalized(cross(ab, normalized(b.right - b.start)));
        else
        {
            b.angles[0] = std::atan2(ab[1], ab[0]);
            /* normal = R * (0,1) with the local R = [[c, s], [-s, c]] */
            b.normal = V3(std::sin(b.angles[0]), std::cos(b.angles[0]), 0.0);
        }
        b.insert_norm = b.normal;
        const V3 test = b.start - (b.length - 0.01 * gs) * b.insert_norm;
        b.insconst = dot(b.insert_norm, test);
    }
    else if (has_config == 2)
    {
        if (d3)
        {
            b.right = b.start;
            b.right[1] += gs * b.ni;
            b.end = b.right;
            b.end[2] += gs * b.nj;
        }
        else
        {
            b.end = b.start;
            b.end[1] += gs * b.ni;
        }
        if (!b.rotmat.is_identity())
        {
            b.end = b.rotmat * (b.end - b.start) + b.start;
            if (d3)
                b.right = b.rotmat * (b.right - b.start) + b.start;
        }
    }
    else if (!b.rotmat.is_identity())
    {
        b.start = b.rotmat * (b.start - b.centre) + b.centre;
        b.end = b.rotmat * (b.end - b.centre) + b.centre;
        if (d3)
            b.right = b.rotmat * (b.right - b.centre) + b.centre;
    }
    const double xlength = d3 ? norm(b.right - b.start) : norm(b.end - b.start);
    b.ni = ceil_i(xlength / gs);
    if (d3)
        b.nj = ceil_i(norm(b.end - b.right) / gs);
    b.nk = ceil_i(b.length / gs);
    b.ni = b.ni > 1 ? b.ni : 1;
    b.nk = b.nk > 1 ? b.nk : 1;
    const int nBuff = 4;
    if (d3)
    {
        b.nj = b.nj > 1 ? b.nj : 1;
        if (b.sub_bound_type == circleSphere)
        {
            if (!b.defined(b.centre))
                b.centre = 0.5 * (b.start + b.end);
            if (b.radius < 0)
                b.radius = norm(b.centre - b.start) * std::cos(M_PI_4);
            b.npts = size_t(b.ni * b.nj * M_PI_4 * (b.nk + nBuff + 1));
        }
        else
            b.npts = size_t(b.ni) * b.nj * (b.nk + nBuff + 1);
    }
    else
        b.npts = size_t(b.ni) * (b.nk + nBuff);
    if (b.vmag != 0)
        b.vel = b.vmag * b.insert_norm;
    post_check(b, gs);
}
std::vector<V3> lattice_disk(const Block& b, double dx, int kk)
{
    std::vector<V3> pts;
    std::uniform_real_distribution<double> unif(0.0, MEPS * dx);
    std::default_random_engine re;
    const double rsq = b.radius * b.radius;
    for (int jj = 0; jj < b.nj; ++jj)
        for (int ii = 0; ii < b.ni; ++ii)
        {
            V3 p = V3(-double(kk), double(ii), double(jj)) * dx;
            const double a = p[1] - b.radius, c = p[2] - b.radius;
            if ((a * a + c * c) > rsq)
                continue;
            p = p + make3(unif(re), unif(re), unif(re));
            pts.push_back(b.rotmat * p + b.start);
        }
    return pts;
}
void inlet_generate(Block& b, double gs)
{
    std::uniform_real_distribution<double> unif(0.0, MEPS * gs);
    std::default_random_engine re;
    std::vector<V3>& pts = b.coords;
    const int nBuff = 4;
    if (b.dim == 2)
    {
        const V3 delta = (b.end - b.start) / double(b.ni);
        const V3 nrm(delta[1], -delta[0], 0.0);
        int jj = 0;
        for (jj = 0; jj < b.nk; ++jj)
            for (int ii = 0; ii < b.ni; ++ii)
            {
                V3 p = delta * double(ii) + nrm * double(-jj);
                p = p + constant(unif(re), 2);
                pts.push_back(p + b.start);
                b.bc.push_back(FJSPH_PIPE);
            }
        for (int ii = 0; ii < b.ni; ++ii)
        {
            V3 p = delta * double(ii) + nrm * double(-jj);
            p = p + constant(unif(re), 2);
            pts.push_back(p + b.start);
            b.bc.push_back(FJSPH_BACK);
            b.back.push_back(pts.size() - 1);
        }
        b.buffer.assign(size_t(b.ni), std::vector<size_t>(nBuff));
        size_t buff = 0;
        for (jj = b.nk + 1; jj <= b.nk + nBuff; ++jj)
        {
            for (int ii = 0; ii < b.ni; ++ii)
            {
                V3 p = delta * double(ii) + nrm * double(-jj);
                p = p + constant(unif(re), 2);
                pts.push_back(p + b.start);
                b.bc.push_back(FJSPH_BUFFER);
                b.buffer[size_t(ii)][buff] = pts.size() - 1;
            }
            buff++;
        }
        return;
    }
    if (b.sub_bound_type == squareCube)
    {
        int kk = 0;
        for (kk = 0; kk < b.nk; ++kk)
            for (int jj = 0; jj < b.nj; ++jj)
                for (int ii = 0; ii < b.ni; ++ii)
                {
                    V3 p = V3(double(-kk), double(ii), double(jj)) * gs;
                    p = p + make3(unif(re), unif(re), unif(re));
                    pts.push_back(b.rotmat * p + b.start);
                    b.bc.push_back(FJSPH_PIPE);
                }
        for (int jj = 0; jj < b.nj; ++jj)
            for (int ii = 0; ii < b.ni; ++ii)
            {
                V3 p = V3(double(-kk), double(jj), double(ii)) * gs; /* sic: (jj, ii) swapped, inlet.cpp:488 */
                p = p + make3(unif(re), unif(re), unif(re));
                pts.push_back(b.rotmat * p + b.start);
                b.bc.push_back(FJSPH_BACK);
                b.back.push_back(pts.size() - 1);
            }
        b.buffer.assign(b.back.size(), std::vector<size_t>(nBuff));
        size_t buff = 0;
        for (kk = b.nk + 1; kk <= b.nk + nBuff; ++kk)
        {
            size_t col = 0;
            for (int jj = 0; jj < b.nj; ++jj)
                for (int ii = 0; ii < b.ni; ++ii)
                {
                    V3 p = V3(double(-kk), double(ii), double(jj)) * gs;
                    p = p + make3(unif(re), unif(re), unif(re));
                    pts.push_back(b.rotmat * p + b.start);
                    b.bc.push_back(FJSPH_BUFFER);
                    b.buffer[col++][buff] = pts.size() - 1;
                }
            buff++;
        }
        return;
    }
    /* Circle: one disk per layer, each from a fresh engine (create_disk, inlet.cpp:370-445) */
    auto disk = [&](int kk) {
        if (!b.particle_order)
            return lattice_disk(b, gs, kk);
        std::vector<V3> d;
        std::uniform_real_distribution<double> u2(0.0, MEPS * gs);
        std::default_random_engine r2;
        ring_points(d, r2, u2, gs, b.radius, -gs * kk, b.rotmat, b.centre, 3);
        return d;
    };
    int kk = 0;
    for (kk = 0; kk < b.nk; ++kk)
        for (const V3& p : disk(kk))
        {
            pts.push_back(p);
            b.bc.push_back(FJSPH_PIPE);
        }
    for (const V3& p : disk(kk))
    {
        b.back.push_back(pts.size());
        pts.push_back(p);
        b.bc.push_back(FJSPH_BACK);
    }
    b.buffer.assign(b.back.size(), std::vector<size_t>(nBuff));
    size_t buff = 0;
    for (kk = b.nk + 1; kk <= b.nk + nBuff; ++kk)
    {
        const std::vector<V3> d = disk(kk);
        for (size_t ii = 0; ii < d.size() && ii < b.buffer.size(); ii++)
        {
            b.buffer[ii][buff] = pts.size();
            pts.push_back(d[ii]);
            b.bc.push_back(FJSPH_BUFFER);
        }
        buff++;
    }
}

// ---------------------------------------------------------------- Coordinates (coordinates.cpp)
void coord_check(Block& b, const Ctx& C, double& gs, std::string& err)
{
    common_check(b, C, err);
    if (b.filename.empty())
    {
        if (b.coords.empty())
            err += "Block \"" + b.name + "\" coordinates have not been ingested properly. ";
        else
            b.npts = b.coords.size();
    }
    post_check(b, gs);
}

void check_block(Block& b, const Ctx& C, double& gs, std::string& err)
{
    std::string e;
    switch (b.bound_type)
    {
    case linePlane: line_check(b, C, gs, e); break;
    case squareCube: square_check(b, C, gs, e); break;
    case circleSphere: circle_check(b, C, gs, e); break;
    case arcSection: arc_check(b, C, gs, e); break;
    case cylinderT: cylinder_check(b, C, gs, e); break;
    case inletZone: inlet_check(b, C, gs, e); break;
    case coordDef: coord_check(b, C, gs, e); break;
    default: e = "unsupported shape. ";
    }
    err += e;
}

// ---------------------------------------------------------------- block file (shapes.cpp:408-625)
struct Shapes
{
    std::vector<std::unique_ptr<Block>> block;
    size_t total_points = 0;
};

bool read_bmap(const std::string& path, const Ctx& C, double& gs, Shapes& out, std::string& err)
{
    std::ifstream fin(path);
    if (!fin.is_open())
    {
        err = path + " file missing";
        return false;
    }
    const int dim = C.dim;
    std::string line, shape_name;
    /* pass 1: one block per "block end", typed by the last "Shape" seen */
    while (std::getline(fin, line))
    {
        const size_t hash = line.find('#');
        if (hash != std::string::npos)
            line = line.substr(0, hash);
        get_string(line, "Shape", shape_name);
        if (line.find("block end") != std::string::npos)
        {
            std::unique_ptr<Block> b(new Block(dim));
            b->bound_type = shape_type_of(shape_name, dim);
            if (b->bound_type < 0)
                err += "Unrecognised boundary shape, \"" + shape_name + "\". ";
            out.block.push_back(std::move(b));
            shape_name.clear();
        }
    }
    if (!err.empty())
        return false;
    fin.clear();
    fin.seekg(0);
    size_t ib = 0;
    const size_t nblocks = out.block.size();
    while (ib < nblocks && std::getline(fin, line))
    {
        const size_t hash = line.find('#');
        if (hash != std::string::npos)
            line = line.substr(0, hash);
        Block& b = *out.block[ib];
        get_string(line, "Name", b.name);
        get_string(line, "Shape", b.shape);
        get_string(line, "Sub-shape", b.subshape);
        get_string(line, "Boundary solver", b.solver_name);
        get_number(line, "Write surface data (0/1)", b.write_data);
        get_number(line, "Fixed velocity or dynamic inlet BC (0/1)", b.fixed_vel_or_dynamic);
        get_vector(line, "Aerodynamic entry normal", b.aero_norm, dim);
        get_vector(line, "Deletion normal", b.delete_norm, dim);
        get_vector(line, "Insertion normal", b.insert_norm, dim);
        get_number(line, "Aerodynamic entry plane constant", b.aeroconst);
        get_number(line, "Deletion plane constant", b.delconst);
        get_number(line, "Insertion plane constant", b.insconst);
        get_number(line, "Pipe depth", b.thickness);
        get_number(line, "i-direction count", b.ni);
        get_number(line, "j-direction count", b.nj);
        get_number(line, "k-direction count", b.nk);
        get_vector(line, "Stretching factor", b.stretch, dim);
        get_vector(line, "Normal vector", b.normal, dim);
        get_vector(line, "Rotation angles", b.angles, dim);
        get_number(line, "Rotation angle", b.angles[0]);
        get_vector(line, "Start coordinate", b.start, dim);
        get_vector(line, "End coordinate", b.end, dim);
        get_vector(line, "Right coordinate", b.right, dim);
        get_vector(line, "Midpoint coordinate", b.mid, dim);
        get_vector(line, "Centre coordinate", b.centre, dim);
        get_vector(line, "Arch normal", b.right, dim);
        get_number(line, "Radius", b.radius);
        get_number(line, "Length", b.length);
        get_number(line, "Arc start (degree)", b.arc_start);
        get_number(line, "Arc end (degree)", b.arc_end);
        get_number(line, "Arc length (degree)", b.arclength);
        get_number(line, "Starting straight length", b.sstraight);
        get_number(line, "Ending straight length", b.estraight);
        get_number(line, "Particle spacing", b.dx);
        get_number(line, "Particle ordering (0=grid,1=HCP)", b.particle_order);
        get_number(line, "Wall thickness", b.thickness);
        get_number(line, "Wall radial particle count", b.nk);
        get_number(line, "Wall is no slip (0/1)", b.no_slip);
        get_string(line, "Coordinate filename", b.filename);
        get_vector(line, "Starting velocity", b.vel, dim);
        get_number(line, "Starting jet velocity", b.vmag);
        get_number(line, "Starting pressure", b.press);
        get_number(line, "Starting density", b.dens);
        get_number(line, "Cole EOS gamma", b.gamma);
        get_number(line, "Speed of sound", b.speedOfSound);
        get_number(line, "Resting density", b.rho_rest);
        get_number(line, "Volume to target", b.renorm_vol);
        get_string(line, "Time data filename", b.position_filename);
        if (line.find("Coordinate data:") != std::string::npos)
        {
            std::string tmp;
            std::getline(fin, tmp);
            size_t npts = 0;
            fin >> npts;
            b.npts = npts;
            b.coords.assign(npts, V3());
            for (size_t ii = 0; ii < npts; ++ii)
                for (int d = 0; d < dim; ++d) fin >> b.coords[ii][d];
        }
        const bool tpos = line.find("Time position data") != std::string::npos;
        const bool tvel = line.find("Time velocity data") != std::string::npos;
        if (tpos || tvel)
        {
            std::string tmp;
            std::getline(fin, tmp);
            size_t nt = 0;
            std::istringstream iss(tmp);
            iss >> nt;
            b.ntimes = nt;
            b.times.assign(nt, 0.0);
            std::vector<V3>& dst = tpos ? b.pos : b.vels;
            dst.assign(nt, V3());
            for (size_t ii = 0; ii < nt; ++ii)
            {
                std::getline(fin, tmp);
                std::istringstream is2(tmp);
                is2 >> b.times[ii];
                for (int d = 0; d < dim; ++d) is2 >> dst[ii][d];
            }
        }
        if (line.find("block end") != std::string::npos)
            ib++;
    }
    for (auto& bp : out.block) check_block(*bp, C, gs, err);
    if (!err.empty())
        return false;
    for (auto& bp : out.block) out.total_points += bp->npts;
    return true;
}

// ---------------------------------------------------------------- JSON block file (shapes.cpp:229-395)
// The reference reads these with nlohmann/json; what it relies on is restated here: one object per file, its members the
// blocks, visited in KEY ORDER (the library's object is a std::map, so blocks come sorted by name, not in file order), a
// repeated key keeps its last value, and typed reads as strict as the library's -- a string only from a string, a bool
// only from true / false, a number from a number or a boolean (integers and reals kept apart, converted by a cast), a
// vector only from an array of exactly `dim` numbers -- where the reference exits on a mismatch this returns the error.
struct JVal
{
    enum Kind
    {
        Null,
        Bool,
        Int,
        Real,
        Str,
        Arr,
        Obj
    } kind = Null;
    bool b = false;
    long long i = 0;
    double f = 0.0;
    std::string s;
    std::vector<JVal> arr;
    std::map<std::string, JVal> obj;
    const JVal* find(const char* key) const
    {
        const auto it = obj.find(key);
        return it == obj.end() ? nullptr : &it->second;
    }
};
struct JsonError
{
    std::string msg;
};
class JsonReader
{
  public:
    explicit JsonReader(const std::string& text) : t(text) {}
    JVal document()
    {
        JVal v = value();
        space();
        if (p != t.size())
            bad("text after the document");
        return v;
    }

  private:
    const std::string& t;
    size_t p = 0;
    [[noreturn]] void bad(const char* what) const { throw JsonError{"JSON parse error at byte " + std::to_string(p) + ": " + what}; }
    void space()
    {
        while (p < t.size() && std::strchr(" \t\n\r", t[p])) ++p;
    }
    bool take(char c)
    {
        space();
        if (p < t.size() && t[p] == c)
        {
            ++p;
            return true;
        }
        return false;
    }
    std::string quoted()
    {
        std::string out;
        for (++p; p < t.size() && t[p] != '"'; ++p)
        {
            if (t[p] != '\\')
            {
                out.push_back(t[p]);
                continue;
            }
            if (++p >= t.size())
                break;
            switch (t[p])
            {
            case 'n': out.push_back('\n'); break;
            case 't': out.push_back('\t'); break;
            case 'r': out.push_back('\r'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'u': bad("\\u escapes are not supported");
            default: out.push_back(t[p]);
            }
        }
        if (p >= t.size())
            bad("unterminated string");
        ++p;
        return out;
    }
    JVal number()
    {
        const size_t a = p;
        bool integral = true;
        if (t[p] == '-')
            ++p;
        const size_t first_digit = p;
        while (p < t.size() && std::isdigit(static_cast<unsigned char>(t[p]))) ++p;
        if (p == first_digit)
            bad("invalid literal");
        if (p < t.size() && t[p] == '.')
        {
            integral = false;
            for (++p; p < t.size() && std::isdigit(static_cast<unsigned char>(t[p])); ++p) {}
        }
        if (p < t.size() && (t[p] == 'e' || t[p] == 'E'))
        {
            integral = false;
            ++p;
            if (p < t.size() && (t[p] == '+' || t[p] == '-'))
                ++p;
            while (p < t.size() && std::isdigit(static_cast<unsigned char>(t[p]))) ++p;
        }
        const std::string tok = t.substr(a, p - a);
        JVal v;
        v.kind = integral ? JVal::Int : JVal::Real;
        if (integral)
            v.i = std::strtoll(tok.c_str(), nullptr, 10);
        else
            v.f = std::strtod(tok.c_str(), nullptr);
        return v;
    }
    JVal value()
    {
        space();
        if (p >= t.size())
            bad("unexpected end of input");
        JVal v;
        if (take('{'))
        {
            v.kind = JVal::Obj;
            if (take('}'))
                return v;
            do
            {
                space();
                if (p >= t.size() || t[p] != '"')
                    bad("expected a member name");
                const std::string key = quoted();
                if (!take(':'))
                    bad("expected ':'");
                v.obj[key] = value();
            } while (take(','));
            if (!take('}'))
                bad("expected ',' or '}'");
            return v;
        }
        if (take('['))
        {
            v.kind = JVal::Arr;
            if (take(']'))
                return v;
            do
                v.arr.push_back(value());
            while (take(','));
            if (!take(']'))
                bad("expected ',' or ']'");
            return v;
        }
        if (t[p] == '"')
        {
            v.kind = JVal::Str;
            v.s = quoted();
            return v;
        }
        for (const char* word : {"true", "false", "null"})
            if (t.compare(p, std::strlen(word), word) == 0)
            {
                p += std::strlen(word);
                v.kind = word[0] == 'n' ? JVal::Null : JVal::Bool;
                v.b = word[0] == 't';
                return v;
            }
        return number();
    }
};

struct JsonBlock /* typed reads of one block's members (get_var, shapes.cpp:229-282) */
{
    const JVal& o;
    const std::string& file;
    int dim;
    [[noreturn]] void mismatch(const char* key, const char* want) const
    {
        throw JsonError{"An error occured trying to read a parameter from the JSON file. File: " + file + " Parameter: " + key +
                        " Error: type must be " + want};
    }
    double as_real(const JVal& v, const char* key) const
    {
        if (v.kind == JVal::Int)
            return static_cast<double>(v.i);
        if (v.kind == JVal::Real)
            return v.f;
        if (v.kind == JVal::Bool)
            return v.b ? 1.0 : 0.0;
        mismatch(key, "number");
    }
    void get(const char* key, std::string& out) const
    {
        if (const JVal* v = o.find(key))
        {
            if (v->kind != JVal::Str)
                mismatch(key, "string");
            out = v->s;
        }
    }
    void get(const char* key, bool& out) const
    {
        if (const JVal* v = o.find(key))
        {
            if (v->kind != JVal::Bool)
                mismatch(key, "boolean");
            out = v->b;
        }
    }
    void get(const char* key, int& out) const
    {
        if (const JVal* v = o.find(key))
            out = v->kind == JVal::Int ? static_cast<int>(v->i) : static_cast<int>(as_real(*v, key));
    }
    void get(const char* key, double& out) const
    {
        if (const JVal* v = o.find(key))
            out = as_real(*v, key);
    }
    void get(const char* key, std::vector<double>& out) const
    {
        if (const JVal* v = o.find(key))
        {
            if (v->kind != JVal::Arr)
                mismatch(key, "array");
            out.clear();
            for (const JVal& e : v->arr) out.push_back(as_real(e, key));
        }
    }
    void get(const char* key, V3& out) const /* taken only when the array has exactly dim entries */
    {
        std::vector<double> tmp;
        get(key, tmp);
        if (int(tmp.size()) == dim)
            for (int d = 0; d < dim; ++d) out[d] = tmp[size_t(d)];
    }
    void get(const char* key, std::vector<V3>& out) const /* rows are read as dim numbers each, as the reference does */
    {
        const JVal* v = o.find(key);
        if (!v)
            return;
        if (v->kind != JVal::Arr)
            mismatch(key, "array");
        if (v->arr.empty())
            return;
        std::vector<V3> rows;
        for (const JVal& r : v->arr)
        {
            if (r.kind != JVal::Arr)
                mismatch(key, "array");
            if (int(r.arr.size()) < dim)
                throw JsonError{"JSON file " + file + ": a row of \"" + key + "\" has fewer than " + std::to_string(dim) + " numbers"};
            V3 x;
            for (int d = 0; d < dim; ++d) x[d] = as_real(r.arr[size_t(d)], key);
            rows.push_back(x);
        }
        out = rows;
    }
};

bool read_json(const std::string& path, const Ctx& C, double& gs, Shapes& out, std::string& err)
{
    std::ifstream fin(path);
    if (!fin.is_open())
    {
        err = path + " file missing";
        return false;
    }
    const std::string text((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
    const int dim = C.dim;
    try
    {
        const JVal doc = JsonReader(text).document();
        if (doc.kind != JVal::Obj)
            throw JsonError{"JSON file " + path + ": the document must be an object of blocks"};
        for (const auto& member : doc.obj) /* key order */
        {
            if (member.second.kind != JVal::Obj && member.second.kind != JVal::Null)
                throw JsonError{"JSON file " + path + ": block \"" + member.first + "\" is not an object"};
            const JsonBlock J{member.second, path, dim};
            std::unique_ptr<Block> bp(new Block(dim));
            Block& b = *bp;
            std::string shape_name;
            J.get("Shape", shape_name);
            b.bound_type = shape_type_of(shape_name, dim);
            if (b.bound_type < 0)
                err += "Unrecognised boundary shape, \"" + shape_name + "\". ";
            b.filename = path; /* shapes.cpp:379: the JSON file's own name, until "Coordinate filename" replaces it */
            b.name = member.first;
            J.get("Shape", b.shape);
            J.get("Sub-shape", b.subshape);
            J.get("Boundary solver", b.solver_name);
            J.get("Fixed velocity or dynamic inlet BC", b.inlet_bc_type); /* "Fixed Velocity" or "Dynamic" */
            J.get("Aerodynamic entry normal", b.aero_norm);
            J.get("Deletion normal", b.delete_norm);
            J.get("Insertion normal", b.insert_norm);
            J.get("Aerodynamic entry plane constant", b.aeroconst);
            J.get("Deletion plane constant", b.delconst);
            J.get("Insertion plane constant", b.insconst);
            J.get("Pipe depth", b.thickness);
            J.get("i-direction count", b.ni);
            J.get("j-direction count", b.nj);
            J.get("k-direction count", b.nk);
            J.get("Stretching factor", b.stretch);
            J.get("Normal vector", b.normal);
            J.get("Rotation angles (degree)", b.angles);
            J.get("Rotation angle (degree)", b.angles[0]);
            J.get("Start coordinates", b.start);
            J.get("End coordinates", b.end);
            J.get("Right coordinates", b.right);
            J.get("Midpoint coordinates", b.mid);
            J.get("Centre coordinates", b.centre);
            J.get("Arch normal", b.right);
            J.get("Radius", b.radius);
            J.get("Length", b.length);
            J.get("Arc start (degree)", b.arc_start);
            J.get("Arc end (degree)", b.arc_end);
            J.get("Arc length (degree)", b.arclength);
            J.get("Start straight length", b.sstraight);
            J.get("End straight length", b.estraight);
            J.get("Particle spacing", b.dx);
            J.get("Particle ordering (Grid/HCP)", b.particle_order);
            J.get("Wall thickness", b.thickness);
            J.get("Wall radial particle count", b.nk);
            J.get("Wall is no-slip", b.no_slip);
            J.get("Start velocity", b.vel);
            J.get("Start jet velocity", b.vmag);
            J.get("Start pressure", b.press);
            J.get("Start density", b.dens);
            J.get("Cole EOS gamma", b.gamma);
            J.get("Speed of sound", b.speedOfSound);
            J.get("Rest density", b.rho_rest);
            J.get("Volume to target", b.renorm_vol);
            J.get("Coordinate filename", b.filename);
            J.get("Coordinate data", b.coords);
            J.get("Time data filename", b.position_filename);
            J.get("Time data", b.times);
            J.get("Position data", b.pos);
            if (b.bound_type >= 0)
                check_block(b, C, gs, err); /* each block is checked as it is read (shapes.cpp:354) */
            out.block.push_back(std::move(bp));
        }
    }
    catch (const JsonError& e)
    {
        err = e.msg;
        return false;
    }
    if (!err.empty())
        return false;
    for (auto& bp : out.block) out.total_points += bp->npts;
    return true;
}

// read_shapes_JSON for a file whose extension is ".json" in any case, read_shapes_bmap otherwise (Init.cpp:276-287)
bool read_blocks(const std::string& path, const Ctx& C, double& gs, Shapes& out, std::string& err)
{
    const size_t dot = path.find_last_of("./");
    std::string ext = (dot != std::string::npos && path[dot] == '.') ? path.substr(dot) : std::string();
    for (char& c : ext) c = char(std::tolower(static_cast<unsigned char>(c)));
    return ext == ".json" ? read_json(path, C, gs, out, err) : read_bmap(path, C, gs, out, err);
}

void generate_points(Shapes& S, double gs, const Ctx& C)
{
    size_t total = 0;
    for (auto& bp : S.block)
    {
        Block& b = *bp;
        if (b.bound_type != coordDef)
            b.coords.clear(); /* every generator assigns its points: coordinate data given to another shape is dropped */
        switch (b.bound_type)
        {
        case linePlane: line_generate(b, gs); break;
        case squareCube: square_generate(b, gs); break;
        case circleSphere: circle_generate(b, gs); break;
        case arcSection: arc_generate(b, gs); break;
        case cylinderT: cylinder_generate(b, gs); break;
        case inletZone: inlet_generate(b, gs); break;
        default: break;
        }
        b.npts = b.coords.size();
        total += b.npts;
        /* use_global_gas_law = 1 (Var.h:382): block gas-law values are the global ones */
        b.rho_rest = C.rho_rest;
        b.gamma = C.gam;
        b.speedOfSound = C.speed_sound;
        b.backgroundP = C.press_pipe;
    }
    S.total_points = total;
}

// ---------------------------------------------------------------- Check_Intersection (Init.cpp:61-225)
// radius_search(tree, q, searchDist) = every point p of the tree with |p - q|^2 < searchDist (nanoflann, strict).
struct PointGrid
{
    const std::vector<V3>& pts;
    double cell, inv;
    V3 lo;
    int n[3];
    std::vector<int> start, items;
    PointGrid(const std::vector<V3>& p, double radius, int dim) : pts(p), cell(radius), inv(1.0 / radius)
    {
        n[0] = n[1] = n[2] = 1;
        if (pts.empty())
            return;
        V3 hi = pts[0];
        lo = pts[0];
        for (const V3& q : pts)
            for (int d = 0; d < dim; ++d)
            {
                lo[d] = std::min(lo[d], q[d]);
                hi[d] = std::max(hi[d], q[d]);
            }
        for (int d = 0; d < dim; ++d) n[d] = std::max(1, std::min(1024, int((hi[d] - lo[d]) * inv) + 1));
        inv = 1.0 / cell;
        std::vector<int> count(size_t(n[0]) * n[1] * n[2] + 1, 0);
        for (const V3& q : pts) count[size_t(key(q)) + 1]++;
        for (size_t k = 1; k < count.size(); ++k) count[k] += count[k - 1];
        start = count;
        items.resize(pts.size());
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (size_t i = 0; i < pts.size(); ++i) items[size_t(fill[size_t(key(pts[i]))]++)] = int(i);
    }
    int coord(double v, int d) const { return std::max(0, std::min(n[d] - 1, int(std::floor((v - lo[d]) * inv)))); }
    int key(const V3& q) const { return (coord(q[2], 2) * n[1] + coord(q[1], 1)) * n[0] + coord(q[0], 0); }
    template <class F>
    void search(const V3& q, double r2, F&& hit) const
    {
        if (pts.empty())
            return;
        /* cells may be wider than `cell` when an axis was clamped to 1024 cells: sweep by coordinate range */
        const double r = std::sqrt(r2);
        int a[3], b[3];
        for (int d = 0; d < 3; ++d)
        {
            a[d] = coord(q[d] - r, d);
            b[d] = coord(q[d] + r, d);
        }
        for (int z = a[2]; z <= b[2]; ++z)
            for (int y = a[1]; y <= b[1]; ++y)
                for (int x = a[0]; x <= b[0]; ++x)
                {
                    const size_t k = size_t((z * n[1] + y) * n[0] + x);
                    for (int s = start[k]; s < start[k + 1]; ++s)
                    {
                        const V3 d = pts[size_t(items[size_t(s)])] - q;
                        if (dot(d, d) < r2)
                            hit(size_t(items[size_t(s)]));
                    }
                }
    }
};

void check_intersection(double dx, int dim, Shapes& bound, Shapes& fluid)
{
    const double sd = (0.9 * dx) * (0.9 * dx);
    for (auto& b : bound.block) b->intersect.assign(b->npts, 0);
    for (size_t id = 0; id < bound.block.size(); ++id)
    {
        PointGrid tree(bound.block[id]->coords, 0.9 * dx, dim);
        for (size_t ii = id; ii < bound.block.size(); ++ii)
        {
            Block& other = *bound.block[ii];
            for (size_t jj = 0; jj < other.coords.size(); ++jj)
                if (other.intersect[jj] == 0)
                    tree.search(other.coords[jj], sd, [&](size_t m) {
                        if (m != jj) /* sic: also compared across different blocks, Init.cpp:95-96 */
                            bound.block[id]->intersect[m] = 1;
                    });
        }
    }
    for (auto& b : fluid.block) b->intersect.assign(b->npts, 0);
    for (size_t id = 0; id < fluid.block.size(); ++id)
    {
        Block& me = *fluid.block[id];
        PointGrid tree(me.coords, 0.9 * dx, dim);
        for (auto& bb : bound.block)
            for (size_t jj = 0; jj < bb->coords.size(); ++jj)
                if (bb->intersect[jj] == 0)
                    tree.search(bb->coords[jj], sd, [&](size_t m) { me.intersect[m] = 1; });
        for (size_t ii = id + 1; ii < fluid.block.size(); ++ii)
        {
            Block& other = *fluid.block[ii];
            for (size_t jj = 0; jj < other.npts; ++jj)
                if (other.intersect[jj] == 0)
                    tree.search(other.coords[jj], sd, [&](size_t m) { me.intersect[m] = 1; });
        }
    }
    for (auto& fb : fluid.block)
        for (size_t bID = 0; bID < fb->back.size(); bID++)
        {
            int does = fb->intersect[fb->back[bID]] ? 1 : 0;
            if (!does)
                for (size_t id : fb->buffer[bID])
                    if (fb->intersect[id])
                        does = 1;
            if (does)
            {
                fb->intersect[fb->back[bID]] = 1;
                for (size_t id : fb->buffer[bID]) fb->intersect[id] = 1;
            }
        }
    for (Shapes* S : {&bound, &fluid})
    {
        size_t tot = 0;
        for (auto& b : S->block)
        {
            b->npts = size_t(std::count(b->intersect.begin(), b->intersect.end(), 0));
            tot += b->npts;
        }
        S->total_points = tot;
    }
}

} // namespace

// ---------------------------------------------------------------- the case object behind the C ABI
struct FjsphCase
{
    int dim = 3;
    FjsphParams params;
    int init_hydro = 0;
    double hydro_height = -1.0;
    int max_frames = -1;            /* "SPH frame count" (IO.cpp:375) */
    long long max_points = -1;      /* "SPH maximum particle count" (IO.cpp:428) */
    std::string output_prefix, restart_prefix; /* IO.cpp:373,355 */
    std::string foam_dir, foam_sol, tau_mesh, tau_bmap, tau_sol; /* IO.cpp:352-354,359-360 */
    std::string vlm_file;                                        /* IO.cpp:366 */
    double scale = 1.0, angle_alpha = 0.0;     /* IO.cpp:356-357 */
    int foam_buoyant = 0;                      /* IO.cpp:364 */
    int offset_axis = -1;                      /* IO.cpp:358; -1 = not in the deck: Var.h:99-103 (0 in 3D, 2 in 2D) */
    int64_t bound_points = 0;
    int n_bound_blocks = 0;
    std::vector<double> xi, v, rho, p, m;
    std::vector<int32_t> b;
    struct Limit
    {
        std::string name;
        FjsphBlock blk;
        std::vector<double> times, vels;
        std::vector<int64_t> back, buffer;
    };
    std::vector<Limit> limits;
};

namespace
{
void emit(FjsphCase& c, const V3& x, const V3& vel, double dens, double mass, double press, int bflag)
{
    for (int d = 0; d < c.dim; ++d)
    {
        c.xi.push_back(x[d]);
        c.v.push_back(vel[d]);
    }
    c.rho.push_back(dens);
    c.m.push_back(mass);
    c.p.push_back(press);
    c.b.push_back(bflag);
}

// get_boundary_velocity (Init.cpp:26-38)
void boundary_velocity(Block& b)
{
    if (b.ntimes != 0)
    {
        b.vels.assign(b.ntimes - 1, V3());
        for (size_t j = 0; j + 1 < b.ntimes; ++j) b.vels[j] = (b.pos[j + 1] - b.pos[j]) / (b.times[j + 1] - b.times[j]);
    }
}

void fill_common(FjsphCase::Limit& L, const Block& b, int dim, bool is_fluid)
{
    std::memset(&L.blk, 0, sizeof(L.blk));
    L.name = b.name;
    L.blk.is_fluid = is_fluid ? 1 : 0;
    L.blk.bound_solver = b.bound_solver;
    L.blk.no_slip = b.no_slip ? 1 : 0;
    L.blk.block_type = b.bound_type;
    L.blk.fixed_vel_or_dynamic = b.fixed_vel_or_dynamic;
    /* bound_block's constructor (Var.h:781-815): planes unset; only fluid blocks get theirs copied (Init.cpp:455-461) */
    for (int d = 0; d < 3; ++d) L.blk.insert_norm[d] = L.blk.delete_norm[d] = L.blk.aero_norm[d] = DEFV;
    L.blk.insconst = L.blk.delconst = L.blk.aeroconst = DEFV;
    (void)dim;
}

// Init_Particles (Init.cpp:270-496)
int init_particles(FjsphCase& c, Shapes& bound, Shapes& fluid)
{
    const FjsphParams& P = c.params;
    const int dim = c.dim;
    int64_t part_id = 0;
    for (auto& bp : bound.block)
    {
        Block& b = *bp;
        c.limits.emplace_back();
        FjsphCase::Limit& L = c.limits.back();
        fill_common(L, b, dim, false);
        L.blk.first = part_id;
        for (size_t ii = 0; ii < b.coords.size(); ii++)
            if (!b.intersect[ii])
            {
                emit(c, b.coords[ii], b.vel, b.dens, P.bnd_mass, b.press, FJSPH_BOUND);
                part_id++;
            }
        if (!b.times.empty())
        {
            if (!b.pos.empty() && b.vels.empty())
                boundary_velocity(b);
            else if (b.vels.empty())
            {
                fj_set_error("No velocity or position data available for boundary block \"%s\" even though times were defined.",
                             b.name.c_str());
                return FJSPH_ERR_INVALID;
            }
            L.times = b.times;
            for (const V3& u : b.vels)
                for (int d = 0; d < 3; ++d) L.vels.push_back(u[d]);
            /* position data give ntimes - 1 velocities, yet the solvers index vels[ntimes - 1] once the last time
               stamp has passed (Newmark_Beta.cpp:75-81, out of bounds in the reference): the wall then stands still */
            L.vels.resize(3 * b.ntimes, 0.0);
            L.blk.n_times = int32_t(b.ntimes);
        }
        else
        {
            L.blk.n_times = 0;
            for (int d = 0; d < 3; ++d) L.vels.push_back(b.vel[d]);
        }
        L.blk.second = part_id;
    }
    c.n_bound_blocks = int(bound.block.size());
    c.bound_points = part_id;
    for (auto& bp : fluid.block)
    {
        Block& b = *bp;
        c.limits.emplace_back();
        FjsphCase::Limit& L = c.limits.back();
        fill_common(L, b, dim, true);
        L.blk.first = part_id;
        if (b.bound_type == inletZone)
        {
            const size_t nBuff = 4;
            size_t ii = 0;
            while (ii < b.bc.size() && b.bc[ii] == FJSPH_PIPE)
            {
                if (!b.intersect[ii])
                {
                    emit(c, b.coords[ii], b.vel, b.dens, P.bnd_mass, b.press, FJSPH_PIPE); /* sic: bnd_mass, Init.cpp:364 */
                    part_id++;
                }
                ii++;
            }
            std::vector<char> skip(b.back.size(), 0);
            for (size_t k = 0; k < b.back.size(); k++) skip[k] = b.intersect[b.back[k]] ? 1 : 0;
            for (size_t k = 0; k < b.back.size(); k++)
                if (!skip[k])
                {
                    emit(c, b.coords[b.back[k]], b.vel, b.dens, P.sim_mass, b.press, FJSPH_BACK);
                    L.back.push_back(part_id);
                    L.buffer.insert(L.buffer.end(), nBuff, 0);
                    part_id++;
                }
            for (size_t f = 0; f < nBuff; f++)
            {
                size_t col = 0;
                for (size_t k = 0; k < b.back.size(); k++)
                    if (!skip[k])
                    {
                        emit(c, b.coords[b.buffer[k][f]], b.vel, b.dens, P.sim_mass, b.press, FJSPH_BUFFER);
                        L.buffer[col * nBuff + f] = part_id;
                        part_id++;
                        col++;
                    }
            }
        }
        else
        {
            for (size_t ii = 0; ii < b.coords.size(); ii++)
                if (!b.intersect[ii])
                {
                    emit(c, b.coords[ii], b.vel, b.dens, P.sim_mass, b.press, FJSPH_FREE);
                    part_id++;
                }
        }
        L.blk.second = part_id;
        if (!b.times.empty())
        {
            L.times = b.times;
            for (const V3& u : b.vels)
                for (int d = 0; d < 3; ++d) L.vels.push_back(u[d]);
            L.vels.resize(3 * b.ntimes, 0.0);
            L.blk.n_times = int32_t(b.ntimes);
        }
        else
        {
            L.blk.n_times = 0;
            L.vels.assign(3, 0.0);
        }
        for (int d = 0; d < 3; ++d)
        {
            L.blk.insert_norm[d] = b.insert_norm[d];
            L.blk.delete_norm[d] = b.delete_norm[d];
            L.blk.aero_norm[d] = b.aero_norm[d];
        }
        L.blk.insconst = b.insconst;
        L.blk.delconst = b.delconst;
        L.blk.aeroconst = b.aeroconst;
    }
    if (c.init_hydro)
    {
        /* Init.cpp:480-493: height along y whatever the dimension */
        const size_t n = c.rho.size();
        for (size_t i = 0; i < n; ++i)
        {
            const double y = c.xi[i * size_t(dim) + 1];
            const double press = std::max(0.0, -P.rho_rest * P.grav[1] * (c.hydro_height - y));
            double dens;
            if (P.pressure_rel == 0)
                dens = P.rho_rest * std::pow(((press - P.press_back) / P.B) + 1.0, 1.0 / P.gam);
            else
                dens = (press - P.press_back) / (P.speed_sound * P.speed_sound) + P.rho_rest;
            c.p[i] = press;
            c.rho[i] = dens;
        }
    }
    for (FjsphCase::Limit& L : c.limits)
    {
        L.blk.times = L.times.empty() ? nullptr : L.times.data();
        L.blk.vels = L.vels.empty() ? nullptr : L.vels.data();
        L.blk.n_back = int32_t(L.back.size());
        L.blk.n_buf = L.back.empty() ? 0 : 4;
        L.blk.back = L.back.empty() ? nullptr : L.back.data();
        L.blk.buffer = L.buffer.empty() ? nullptr : L.buffer.data();
    }
    return FJSPH_OK;
}

std::string dir_of(const std::string& path)
{
    const size_t s = path.find_last_of('/');
    return s == std::string::npos ? "" : path.substr(0, s + 1);
}
std::string resolve(const std::string& name, const std::string& para_dir)
{
    if (name.empty() || name[0] == '/')
        return name;
    std::ifstream probe(name);
    if (probe.is_open())
        return name; /* the reference opens names relative to the working directory */
    return para_dir + name;
}
} // namespace

// GetInput (IO.cpp:305-723) for the keys the path reads + Init_Particles.  dim = the SIMDIM of the build the deck is for.
static int case_read_impl(const char* para_path, int dim, FjsphCase** out);
// nothing thrown inside (a block file asking for more memory than there is, say) crosses the C boundary
extern "C" int fjsph_case_read(const char* para_path, int dim, FjsphCase** out)
{
    try
    {
        return case_read_impl(para_path, dim, out);
    }
    catch (const std::exception& e)
    {
        fj_set_error("case_read: %s", e.what());
        return FJSPH_ERR_IO;
    }
}
static int case_read_impl(const char* para_path, int dim, FjsphCase** out)
{
    if (!para_path || !out || (dim != 2 && dim != 3))
    {
        fj_set_error("case_read: need a para file, dim 2 or 3 and an output pointer");
        return FJSPH_ERR_INVALID;
    }
    std::unique_ptr<FjsphCase> c(new FjsphCase());
    c->dim = dim;
    int st = fjsph_default_params(&c->params, dim);
    if (st)
        return st;
    char fluid_file[1024] = "", bound_file[1024] = "";
    st = fjsph_read_para(para_path, &c->params, fluid_file, bound_file, 1024);
    if (st)
        return st;
    double scale = 1.0;
    {
        /* case-level keys GetInput reads that are not part of FjsphParams (IO.cpp:356,387-388) */
        std::ifstream fin(para_path);
        std::string line;
        while (std::getline(fin, line))
        {
            const size_t hash = line.find('#');
            if (hash != std::string::npos)
                line = line.substr(0, hash);
            get_number(line, "Grid scale", scale);
            get_number(line, "Init hydrostatic pressure (0/1)", c->init_hydro);
            get_number(line, "Hydrostatic height", c->hydro_height);
            get_number(line, "SPH frame count", c->max_frames);
            get_number(line, "SPH maximum particle count", c->max_points);
            get_string(line, "Output files prefix", c->output_prefix);
            get_string(line, "SPH restart prefix", c->restart_prefix);
            get_string(line, "OpenFOAM input directory", c->foam_dir);
            get_string(line, "VLM definition filename", c->vlm_file);
            get_string(line, "OpenFOAM solution directory", c->foam_sol);
            get_number(line, "OpenFOAM buoyant (0/1)", c->foam_buoyant);
            get_string(line, "Primary grid face filename", c->tau_mesh);
            get_string(line, "Boundary mapping filename", c->tau_bmap);
            get_string(line, "Restart-data prefix", c->tau_sol);
            get_number(line, "Angle alpha (degree)", c->angle_alpha);
            get_number(line, "2D offset vector (0 / x=1,y=2,z=3)", c->offset_axis);
        }
    }
    if (c->offset_axis < 0)
        c->offset_axis = dim == 2 ? 2 : 0;
    if (dim == 3)
        c->offset_axis = 0; /* IO.cpp:686-692: a 3D build ignores the 2D setting */
    else if (c->offset_axis < 1 || c->offset_axis > 3)
    {
        fj_set_error(c->offset_axis == 0 ? "Offset axis has not been defined." : "2D offset axis option out of bounds");
        return FJSPH_ERR_INVALID; /* IO.cpp:694-705 */
    }
    /* aero source, IO.cpp:464-499: a mesh named in the deck couples the aero model to it (meshInfl) */
    c->scale = scale;
    if (!c->tau_mesh.empty())
    {
        /* a TAU mesh wins over an OpenFOAM case (IO.cpp:465-533); it needs its boundary map and its solution file */
        if (c->tau_bmap.empty())
        {
            fj_set_error("Input TAU bmap file not defined.");
            return FJSPH_ERR_INVALID;
        }
        if (c->tau_bmap == "(thisfile)")
            c->tau_bmap = para_path;
        if (c->tau_sol.empty())
        {
            fj_set_error("Input TAU solution file not defined.");
            return FJSPH_ERR_INVALID;
        }
        /* like the block files: a name the working directory does not hold is looked for beside the para file */
        c->tau_bmap = resolve(c->tau_bmap, dir_of(para_path));
        c->tau_mesh = resolve(c->tau_mesh, dir_of(para_path));
        c->tau_sol = resolve(c->tau_sol, dir_of(para_path));
        /* TAU::Read_BMAP (CDFIO.cpp:234-315), the part the time step sees: the map may restate the angle of attack, and
           gravity is turned by it -- g_z cos(alpha), and g_x = -g_Y sin(alpha) as the reference writes it in 3D too */
        std::ifstream bm(c->tau_bmap);
        if (!bm.is_open())
        {
            fj_set_error("Couldn't open the boundary map file. Attempted path: %s", c->tau_bmap.c_str());
            return FJSPH_ERR_IO;
        }
        std::string line;
        while (std::getline(bm, line))
        {
            line = ltrim(line);
            if (!line.empty() && line[0] == '#')
                continue;
            get_number(line, "Angle alpha (degree)", c->angle_alpha);
        }
        const double alpha = c->angle_alpha * (M_PI / 180.0); /* `angle_alpha *= M_PI / 180.0`, CDFIO.cpp:219,301 */
        const double g1 = c->params.grav[1];
        c->params.grav[dim - 1] = c->params.grav[dim - 1] * std::cos(alpha);
        c->params.grav[0] = -g1 * std::sin(alpha);
        c->foam_dir.clear();
        c->params.asource = 1;
    }
    if (!c->foam_dir.empty())
    {
        if (c->foam_sol.empty())
        {
            fj_set_error("OpenFOAM solution directory not defined.");
            return FJSPH_ERR_INVALID;
        }
        c->foam_dir = resolve(c->foam_dir, dir_of(para_path));
        c->params.asource = 1;
    }
    /* IO.cpp:465-477: with no mesh named, a 3D deck that names a VLM definition takes the vortex-lattice aero source.  The
       engine has no such source (out of scope) and says so at fjsph_create; the deck must not run on a constant free stream
       unnoticed */
    if (c->tau_mesh.empty() && c->foam_dir.empty() && dim == 3 && !c->vlm_file.empty())
        c->params.asource = 2;
    st = fjsph_set_values(&c->params);
    if (st)
        return st;
    if (c->init_hydro && c->hydro_height < 0)
    {
        fj_set_error("Hydrostatic initialisation requested but no hydrostatic height given."); /* IO.cpp:596-602 */
        return FJSPH_ERR_INVALID;
    }
    const FjsphParams& P = c->params;
    Ctx C{dim, scale, P.rho_rest, P.speed_sound, P.gam, P.press_pipe, P.nu};
    const std::string pdir = dir_of(para_path);
    Shapes bound, fluid;
    std::string err;
    double gs = P.dx; /* Init_Particles passes a local copy of svar.dx the checks may raise (Init.cpp:272-290) */
    if (!bound_file[0])
    {
        /* the reference insists on both block files (IO.cpp:572-585); a case without walls names an empty one, as
           Examples/Droplet does */
        fj_set_error("Input boundary definition filename is not set in \"%s\"", para_path);
        return FJSPH_ERR_IO;
    }
    if (!read_blocks(resolve(bound_file, pdir), C, gs, bound, err))
    {
        fj_set_error("boundary blocks: %s", err.c_str());
        return FJSPH_ERR_IO;
    }
    if (!fluid_file[0])
    {
        fj_set_error("Input fluid definition filename is not set in \"%s\"", para_path);
        return FJSPH_ERR_IO;
    }
    if (!read_blocks(resolve(fluid_file, pdir), C, gs, fluid, err))
    {
        fj_set_error("fluid blocks: %s", err.c_str());
        return FJSPH_ERR_IO;
    }
    generate_points(bound, P.dx, C);
    generate_points(fluid, P.dx, C);
    check_intersection(P.dx, dim, bound, fluid);
    st = init_particles(*c, bound, fluid);
    if (st)
        return st;
    *out = c.release();
    return FJSPH_OK;
}

extern "C" void fjsph_case_free(FjsphCase* c) { delete c; }
extern "C" int64_t fjsph_case_count(const FjsphCase* c) { return c ? int64_t(c->rho.size()) : 0; }
extern "C" int64_t fjsph_case_bound_points(const FjsphCase* c) { return c ? c->bound_points : 0; }
extern "C" int32_t fjsph_case_num_blocks(const FjsphCase* c) { return c ? int32_t(c->limits.size()) : 0; }
extern "C" int32_t fjsph_case_dim(const FjsphCase* c) { return c ? c->dim : 0; }
// the para's "2D offset vector" (IO.cpp:358): which axis a 2D case ignores -- picks the velocity components of a TAU
// solution (fjsph_tau_read_edge); 0 in a 3D case
extern "C" int32_t fjsph_case_offset_axis(const FjsphCase* c) { return c ? c->offset_axis : 0; }
// the run-control keys of GetInput the frame loop reads (FJSPH.cpp:262-330): frame count, particle capacity, prefixes
extern "C" int fjsph_case_io(const FjsphCase* c, int32_t* max_frames, int64_t* max_points, char* output_prefix,
                             char* restart_prefix, int32_t cap)
{
    if (!c)
        return FJSPH_ERR_INVALID;
    if (max_frames)
        *max_frames = c->max_frames;
    if (max_points)
        *max_points = c->max_points;
    if (output_prefix && cap > 0)
        std::snprintf(output_prefix, size_t(cap), "%s", c->output_prefix.c_str());
    if (restart_prefix && cap > 0)
        std::snprintf(restart_prefix, size_t(cap), "%s", c->restart_prefix.c_str());
    return FJSPH_OK;
}
// the TAU mesh and solution files the deck couples to ("" when it names none) and its "Grid scale": hand them to
// fjsph_tau_read, then fjsph_upload_mesh
extern "C" int fjsph_case_tau(const FjsphCase* c, char* mesh_file, char* solution_file, double* scale, int32_t cap)
{
    if (!c)
        return FJSPH_ERR_INVALID;
    if (mesh_file && cap > 0)
        std::snprintf(mesh_file, size_t(cap), "%s", c->tau_mesh.c_str());
    if (solution_file && cap > 0)
        std::snprintf(solution_file, size_t(cap), "%s", c->tau_sol.c_str());
    if (scale)
        *scale = c->scale;
    return FJSPH_OK;
}
// the OpenFOAM case the deck couples to ("" when it names none): hand it to fjsph_foam_read, then fjsph_upload_mesh
extern "C" int fjsph_case_foam(const FjsphCase* c, char* foam_dir, char* solution_dir, int32_t* buoyant, int32_t cap)
{
    if (!c)
        return FJSPH_ERR_INVALID;
    if (foam_dir && cap > 0)
        std::snprintf(foam_dir, size_t(cap), "%s", c->foam_dir.c_str());
    if (solution_dir && cap > 0)
        std::snprintf(solution_dir, size_t(cap), "%s", c->foam_sol.c_str());
    if (buoyant)
        *buoyant = c->foam_buoyant;
    return FJSPH_OK;
}
extern "C" int fjsph_case_params(const FjsphCase* c, FjsphParams* out)
{
    if (!c || !out)
        return FJSPH_ERR_INVALID;
    *out = c->params;
    return FJSPH_OK;
}
extern "C" int fjsph_case_block(const FjsphCase* c, int32_t i, FjsphBlock* out, char* name, int32_t name_cap)
{
    if (!c || !out || i < 0 || size_t(i) >= c->limits.size())
    {
        fj_set_error("case_block: index out of range");
        return FJSPH_ERR_INVALID;
    }
    *out = c->limits[size_t(i)].blk;
    if (name && name_cap > 0)
        std::snprintf(name, size_t(name_cap), "%s", c->limits[size_t(i)].name.c_str());
    return FJSPH_OK;
}
// xi and v are [n][dim] (dim = fjsph_case_dim); part_id is 0..n-1 in the emitted order (Init.cpp:298-475)
extern "C" int fjsph_case_state(const FjsphCase* c, FjsphStateView* s)
{
    if (!c || !s)
        return FJSPH_ERR_INVALID;
    const size_t n = c->rho.size();
    if (s->n != int64_t(n))
    {
        fj_set_error("case_state: view holds %lld particles, the case %zu", (long long)s->n, n);
        return FJSPH_ERR_INVALID;
    }
    if (s->xi)
        std::memcpy(s->xi, c->xi.data(), c->xi.size() * sizeof(double));
    if (s->v)
        std::memcpy(s->v, c->v.data(), c->v.size() * sizeof(double));
    if (s->rho)
        std::memcpy(s->rho, c->rho.data(), n * sizeof(double));
    if (s->p)
        std::memcpy(s->p, c->p.data(), n * sizeof(double));
    if (s->m)
        std::memcpy(s->m, c->m.data(), n * sizeof(double));
    if (s->b)
        std::memcpy(s->b, c->b.data(), n * sizeof(int32_t));
    if (s->part_id)
        for (size_t i = 0; i < n; ++i) s->part_id[i] = int64_t(i);
    return FJSPH_OK;
}
