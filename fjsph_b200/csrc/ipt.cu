// ipt.cu — implicit particle tracking on the device: the particles update_data hands over at a delete plane
// (Integration.cpp:151-169) followed cell by cell through the aero mesh until they leave it, pass max_x or fail.
//
// Replaces:
//   IPT::Integrate, BFD1, BFD2, AeroForce, Terminate_Particle   reference src/IPT.cpp:735-1107
//   FindFace, CheckCellFace                                     reference src/Containment.cpp:896-1079
//   Cross_Plane (2D and 3D), MollerTrumbore,
//   RayNormalIntersection (2D and 3D)                           reference src/Geometry.cpp:399-478, 579-744
//
// The reference runs `#pragma omp parallel for` over the converted particles, each a serial march; here one thread
// marches one particle (the trajectories are independent and short: a few to a few hundred cells), reading the mesh
// arrays of fjsph_upload_mesh (mesh.cu) plus the two the tracker alone needs: the owner cell of every face and its
// fourth corner.  This file is compiled with -fmad=false and every expression keeps the reference's order of
// operations, so a trajectory differs from the CPU's only through pow() and log10() (GetCd, the convergence test).
//
// Kept as the reference has them (notes Q9-Q11): MollerTrumbore's e2 = v2 - v1, which makes the
// tracker accept only the half (v0, v1, v3) of a parallelogram face -- a particle leaving through the other half is
// stopped and counted as failed; pnp1.t advanced once more after a particle that ends in its very first step.
// The product's own: max_steps bounds the march (failed = 2), a containing cell outside the mesh counts as "no face".
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.cuh"

namespace
{
constexpr int TPB = 64;
constexpr double MEPS = 2.220446049250313e-16;

struct V3
{
    double x, y, z;
};
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm(const V3& a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 normalized(const V3& a)
{
    const double z = dot(a, a);
    return z > 0.0 ? a / sqrt(z) : a;
}
__device__ __forceinline__ V3 cross(const V3& a, const V3& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct IptMesh
{
    int dim, n_cells;
    const double4* __restrict__ fx;  // 3 per face: {v0.xyz, v1.x} {v1.yz, v2.xy} {v2.z, vlast.xyz}
    const double4* __restrict__ fq;  // {face[3].xyz, vertex count}
    const int* __restrict__ fown;    // leftright.first
    const int* __restrict__ fmark;   // leftright.second
    const int* __restrict__ cell_ptr;
    const int* __restrict__ cell_faces;
    const double4* __restrict__ cc;  // {centre.xyz, cRho}
    const double4* __restrict__ cvp; // {cVel.xyz, cP}
};

struct Part /* IPTPart, Var.h:645-813 */
{
    long long part_id;
    int going, failed;
    double t, dt;
    int faceID; /* -1: c_no_face */
    V3 faceV;
    double faceRho;
    int cellID;
    V3 cellV;
    double cellRho;
    double acc;
    V3 v, xi;
    double d;
};

// sign of | p 1; a 1; b 1; c 1 | as mesh.cu evaluates it: -det3 of the edge vectors
__device__ __forceinline__ double det4_sign_arg(const V3& p, const V3& a, const V3& b, const V3& c)
{
    const V3 u = a - p, v = b - p, w = c - p;
    const double det3 = u.x * (v.y * w.z - v.z * w.y) - u.y * (v.x * w.z - v.z * w.x) + u.z * (v.x * w.y - v.y * w.x);
    return -det3;
}

// Geometry.cpp:666-714
__device__ bool moller_trumbore(const V3& v0, const V3& v1, const V3& v2, const V3& o, const V3& dir, double& dt, double& denom)
{
    const V3 e1 = v1 - v0;
    const V3 e2 = v2 - v1;
    const V3 h = cross(dir, e2);
    const double a = dot(e1, h);
    if (a < MEPS && a > -MEPS)
    {
        dt = MEPS;
        denom = MEPS;
        return false;
    }
    const double f = 1.0 / a;
    const V3 s = o - v0;
    const double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0)
    {
        dt = 9999999;
        denom = -1;
        return false;
    }
    const V3 q = cross(s, e1);
    const double v = f * dot(dir, q);
    if (v < 0.0 || u + v > 1.0)
    {
        dt = 9999999;
        denom = -1;
        return false;
    }
    dt = f * dot(e2, q);
    denom = dt > MEPS ? 1.0 : -1.0;
    return true;
}

// Cross_Plane + RayNormalIntersection of one face of the cell; false when the segment misses the face's plane
__device__ bool face_ahead(const IptMesh& M, int f, int cell, const V3& testp, const V3& rayp, const V3& origin, const V3& dir,
                           int current_face, bool& skip, double& dt, double& denom)
{
    const double4 a = M.fx[3 * f], b = M.fx[3 * f + 1];
    skip = false;
    if (M.dim == 2)
    {
        const double f0x = a.x, f0y = a.y, f1x = a.w, f1y = b.x;
        /* Geometry.cpp:399-440 */
        double vol = testp.x * (f0y - f1y) - testp.y * (f0x - f1x) + (f0x * f1y - f1x * f0y);
        const int flag1 = vol < 0.0;
        vol = rayp.x * (f0y - f1y) - rayp.y * (f0x - f1x) + (f0x * f1y - f1x * f0y);
        const int flag2 = vol < 0.0;
        if (flag1 == flag2)
            return false;
        if (f == current_face)
        {
            skip = true;
            return true;
        }
        /* Geometry.cpp:444-478 */
        double nx = f0y - f1y, ny = f1x - f0x;
        const double fcx = (f0x + f1x) / 2.0, fcy = (f0y + f1y) / 2.0;
        const double4 c = M.cc[cell];
        const double cdx = fcx - c.x, cdy = fcy - c.y;
        if (nx * cdx + ny * cdy < 0)
        {
            nx = -1.0 * nx;
            ny = -1.0 * ny;
        }
        double tx = 0.0, ty = 0.0;
        {
            const double ax = f0x - origin.x, ay = f0y - origin.y;
            if (ax * ax + ay * ay > tx * tx + ty * ty)
            {
                tx = ax;
                ty = ay;
            }
            const double bx = f1x - origin.x, by = f1y - origin.y;
            if (bx * bx + by * by > tx * tx + ty * ty)
            {
                tx = bx;
                ty = by;
            }
        }
        denom = dir.x * nx + dir.y * ny;
        dt = (tx * nx + ty * ny) / denom;
        return true;
    }
    const double4 c = M.fx[3 * f + 2];
    const V3 f0 = {a.x, a.y, a.z}, f1 = {a.w, b.x, b.y}, f2 = {b.z, b.w, c.x};
    /* Geometry.cpp:579-621 */
    const int flag1 = det4_sign_arg(testp, f0, f1, f2) < 0.0;
    const int flag2 = det4_sign_arg(rayp, f0, f1, f2) < 0.0;
    if (flag1 == flag2)
        return false;
    if (f == current_face)
    {
        skip = true;
        return true;
    }
    /* Geometry.cpp:717-744 */
    const double4 q = M.fq[f];
    if (q.w == 3.0)
        moller_trumbore(f0, f1, f2, origin, dir, dt, denom);
    else if (!moller_trumbore(f0, f1, f2, origin, dir, dt, denom))
    {
        const V3 f3 = {q.x, q.y, q.z};
        moller_trumbore(f0, f3, f2, origin, dir, dt, denom);
    }
    return true;
}

// Containment.cpp:944-1079 (CheckCellFace fused into the loop: the faces are visited in the same order)
__device__ double find_face(const IptMesh& M, const Part& pn, Part& pnp1)
{
    V3 testv = 0.5 * (pnp1.v + pn.v);
    if (norm(testv) < 1e-10)
        testv = pnp1.cellV;
    const V3 dirn = normalized(testv);
    const V3 testp = pn.xi - 1e3 * dirn;
    const V3 rayp = pn.xi + 1e3 * dirn;

    int nextface = pn.faceID;
    double mindist = 1e6;
    int hasintersect = 0;
    if (pn.cellID >= 0 && pn.cellID < M.n_cells)
        for (int k = M.cell_ptr[pn.cellID]; k < M.cell_ptr[pn.cellID + 1]; ++k)
        {
            const int f = M.cell_faces[k];
            bool skip;
            double dt = 0.0, denom = 0.0;
            if (!face_ahead(M, f, pn.cellID, testp, rayp, pn.xi, testv, pn.faceID, skip, dt, denom) || skip)
                continue;
            if (denom > 0)
            {
                if (dt < mindist)
                {
                    nextface = f;
                    mindist = dt > MEPS ? dt : MEPS;
                }
                hasintersect = 1;
            }
        }
    if (hasintersect == 0 || nextface < 0)
    {
        pnp1.xi = pn.xi;
        pnp1.v = pn.v;
        pnp1.going = 0;
        return 0.0;
    }
    pnp1.faceID = nextface;
    const int own = M.fown[nextface];
    const int newcell = (own == pn.cellID) ? M.fmark[nextface] : own;
    pnp1.cellID = newcell;
    if (newcell > -1)
    {
        const double4 cv = M.cvp[newcell];
        pnp1.cellV = {cv.x, cv.y, cv.z};
        pnp1.cellRho = M.cc[newcell].w;
    }
    return mindist;
}

// Aero.h:10-13
__device__ __forceinline__ double get_cd(double Re)
{
    return (1.0 + 0.197 * pow(Re, 0.63) + 2.6e-04 * pow(Re, 1.38)) * (24.0 / (Re + 0.00001));
}
// IPT.cpp:829-835
__device__ __forceinline__ double aero_force(const V3& Vdiff, const Part& pi, double mu_g, double rho_l)
{
    const double Re = 2.0 * pi.faceRho * norm(Vdiff) * pi.d / mu_g;
    const double Cd = get_cd(Re);
    return norm(Vdiff) * (3.0 * Cd * pi.faceRho) / (4.0 * pi.d * rho_l);
}
// IPT.cpp:837-849
__device__ void bfd1(const V3& g, double mu_g, double rho_l, double dt, const Part& pn, Part& pnp1)
{
    const V3 Vdiff = pnp1.faceV - pnp1.v;
    const double res = aero_force(Vdiff, pnp1, mu_g, rho_l);
    pnp1.acc = res;
    pnp1.v = ((pn.v + (dt * res) * pnp1.faceV) + dt * g) / (1.0 + dt * res);
    pnp1.xi = pn.xi + (0.5 * dt) * (pnp1.v + pn.v);
}
// IPT.cpp:851-869
__device__ void bfd2(const V3& g, double mu_g, double rho_l, double dt, double dtm1, const Part& pnm1, const Part& pn, Part& pnp1)
{
    const V3 Vdiff = pnp1.faceV - pnp1.v;
    const double res = aero_force(Vdiff, pnp1, mu_g, rho_l);
    pnp1.acc = res;
    pnp1.v = ((dt + dtm1) * ((dt * dtm1) * (res * pnp1.faceV + g) + (dt + dtm1) * pn.v) - (dt * dt) * pnm1.v) /
             (dtm1 * ((2 * dt + dtm1) + dt * (dt + dtm1) * res));
    pnp1.xi = pn.xi + (0.5 * pnp1.dt) * (pnp1.v + pn.v);
}

__device__ void point_out(const Part& p, FjsphIptPoint& q)
{
    q.part_id = p.part_id;
    q.cellID = p.cellID;
    q.faceID = p.faceID;
    q.going = p.going;
    q.failed = p.failed;
    q.t = p.t;
    q.dt = p.dt;
    q.acc = p.acc;
    q.cellRho = p.cellRho;
    q.xi[0] = p.xi.x, q.xi[1] = p.xi.y, q.xi[2] = p.xi.z;
    q.v[0] = p.v.x, q.v[1] = p.v.y, q.v[2] = p.v.z;
    q.cellV[0] = p.cellV.x, q.cellV[1] = p.cellV.y, q.cellV[2] = p.cellV.z;
}

struct Recorder
{
    FjsphIptPoint* out; /* this particle's row of the record table, or NULL */
    long long cap;
    int count;
    __device__ void push(const Part& p)
    {
        if (out && count < cap)
            point_out(p, out[count]);
        count++;
    }
};

// IPT::Integrate, IPT.cpp:871-1107: one thread, one particle
__global__ void k_ipt_integrate(IptMesh M, FjsphIptSettings S, int n, const FjsphDeleted* __restrict__ in, FjsphIptPoint* last,
                                int* n_steps, FjsphIptPoint* records, long long record_cap, int* n_records,
                                unsigned long long* tallies /* success, failed */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    /* IPTPart(SPHPart const&, time, diam, area), Var.h:737-763 */
    Part pnp1;
    pnp1.part_id = in[i].part_id;
    pnp1.going = 1;
    pnp1.failed = 0;
    pnp1.t = in[i].t;
    pnp1.dt = 0.0;
    pnp1.faceID = -1;
    pnp1.faceV = {0.0, 0.0, 0.0};
    pnp1.faceRho = 0.0;
    const long long cell_in = in[i].cellID;
    pnp1.cellID = (cell_in >= 0 && cell_in < M.n_cells) ? int(cell_in) : -3;
    pnp1.cellV = {in[i].cellV[0], in[i].cellV[1], M.dim == 2 ? 0.0 : in[i].cellV[2]};
    pnp1.cellRho = in[i].cellRho;
    pnp1.acc = 0.0;
    pnp1.v = {in[i].v[0], in[i].v[1], M.dim == 2 ? 0.0 : in[i].v[2]};
    pnp1.xi = {in[i].xi[0], in[i].xi[1], M.dim == 2 ? 0.0 : in[i].xi[2]};
    pnp1.d = S.diam;
    Part pn = pnp1, pnm1 = pnp1;

    const V3 g = {S.grav[0], S.grav[1], M.dim == 2 ? 0.0 : S.grav[2]};
    Recorder R = {records ? records + (long long)i * record_cap : nullptr, record_cap, 0};
    R.push(pnp1);
    const unsigned min_iter = 3, max_iter = unsigned(S.max_subits > 0 ? S.max_subits : 0);
    long long steps = 0;
    int outcome = 0;
    bool first = true;
    for (;;)
    {
        if (!first)
        {
            if (pnp1.d < 1.0e-5)
                pnp1.v = pnp1.cellV;
            else if (pnp1.d < 100.0e-6)
                pnp1.v = (1.0 - (pnp1.d - 1e-05) / (9e-5)) * pnp1.cellV + ((pnp1.d - 1e-05) / (9e-5)) * pnp1.v;
        }
        pnp1.dt = find_face(M, pn, pnp1);
        pnp1.faceV = 0.5 * (pnp1.cellV + pn.cellV);
        pnp1.faceRho = 0.5 * (pnp1.cellRho + pn.cellRho);

        unsigned iter = 0;
        double error = 1.0;
        while ((error > -7.0 || iter < min_iter) && iter < max_iter)
        {
            const double dt = pnp1.dt;
            if (S.eq_order == 2)
            {
                const double dtm1 = first ? pnp1.dt : (iter > 0 ? pn.dt : pnp1.dt);
                bfd2(g, S.mu_g, S.rho_rest, dt, dtm1, pnm1, pn, pnp1);
            }
            else
                bfd1(g, S.mu_g, S.rho_rest, dt, pn, pnp1);
            const double dt_temp = find_face(M, pn, pnp1);
            if (first || double(iter) > S.n_relax)
                pnp1.dt = (1.0 - S.relax) * dt_temp + S.relax * pnp1.dt;
            else
                pnp1.dt = dt_temp;
            pnp1.faceV = 0.5 * (pnp1.cellV + pn.cellV);
            pnp1.faceRho = 0.5 * (pnp1.cellRho + pn.cellRho);
            error = log10(fabs(pnp1.dt - dt));
            iter++;
        }
        steps++;

        bool ended = false;
        if (pnp1.going == 0 || norm(pnp1.v) > 1e3 || norm(pnp1.xi - pn.xi) > S.max_length)
        {
            pnp1.failed = 1;
            pnp1.going = 0; /* Terminate_Particle: a failed particle's last state is not recorded */
            outcome = 2;
            ended = true;
        }
        else if (pnp1.cellID < 0 || pnp1.xi.x > S.max_x)
        {
            pnp1.going = 0;
            R.push(pnp1);
            outcome = 1;
            ended = true;
        }
        if (ended && !first)
            break;
        if (S.eq_order == 2)
            pnm1 = pn;
        pn = pnp1;
        pnp1.t += pnp1.dt;
        if (ended)
            break;
        if (S.record)
            R.push(pnp1);
        first = false;
        if (steps >= S.max_steps)
        {
            pnp1.failed = 2;
            pnp1.going = 0;
            outcome = 2;
            break;
        }
    }
    if (last)
        point_out(pnp1, last[i]);
    if (n_steps)
        n_steps[i] = int(steps);
    if (n_records)
        n_records[i] = R.count;
    atomicAdd(&tallies[outcome == 1 ? 0 : 1], 1ull);
}

template <class T>
struct DevBuf
{
    T* p = nullptr;
    ~DevBuf()
    {
        if (p)
            cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)); }
};
} // namespace

extern "C" int fjsph_ipt_integrate(FjsphEngine* e, const FjsphIptSettings* s, int64_t n, const FjsphDeleted* in, FjsphIptPoint* last,
                                   int32_t* n_steps, FjsphIptPoint* records, int64_t record_cap, int32_t* n_records,
                                   int64_t* n_success, int64_t* n_failed)
{
    if (!e || !s || n < 0 || (n > 0 && !in) || (records && record_cap <= 0) || n > 0x7fffffff)
    {
        fj_set_error("ipt_integrate: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    if (s->eq_order != 1 && s->eq_order != 2)
    {
        fj_set_error("Equation order not 1 or 2. Please choose between these."); /* IO.cpp:668-672 */
        return FJSPH_ERR_INVALID;
    }
    if (!(s->diam > 0.0) || !(s->mu_g > 0.0) || !(s->rho_rest > 0.0) || s->max_steps <= 0)
    {
        fj_set_error("ipt_integrate: diam, mu_g, rho_rest and max_steps must be positive");
        return FJSPH_ERR_INVALID;
    }
    const DeviceMesh& D = e->mesh;
    if (!D.loaded)
    {
        fj_set_error("ipt_integrate needs a mesh: call fjsph_upload_mesh first (Integration.cpp:151: Asource != constVel)");
        return FJSPH_ERR_STATE;
    }
    if (n_success)
        *n_success = 0;
    if (n_failed)
        *n_failed = 0;
    if (n == 0)
        return FJSPH_OK;
    cudaSetDevice(e->device);
    const size_t nn = size_t(n), cap = records ? size_t(record_cap) : 0;
    DevBuf<FjsphDeleted> d_in;
    DevBuf<FjsphIptPoint> d_last, d_rec;
    DevBuf<int> d_steps, d_nrec;
    DevBuf<unsigned long long> d_tally;
    FJ_CUDA(d_in.alloc(nn));
    FJ_CUDA(d_last.alloc(nn));
    FJ_CUDA(d_rec.alloc(nn * cap));
    FJ_CUDA(d_steps.alloc(nn));
    FJ_CUDA(d_nrec.alloc(nn));
    FJ_CUDA(d_tally.alloc(2));
    FJ_CUDA(cudaMemcpyAsync(d_in.p, in, nn * sizeof(FjsphDeleted), cudaMemcpyHostToDevice, e->stream));
    FJ_CUDA(cudaMemsetAsync(d_tally.p, 0, 2 * sizeof(unsigned long long), e->stream));
    if (cap)
        FJ_CUDA(cudaMemsetAsync(d_rec.p, 0, nn * cap * sizeof(FjsphIptPoint), e->stream));
    IptMesh M;
    M.dim = D.dim;
    M.n_cells = D.n_cells;
    M.fx = D.fx;
    M.fq = D.fq;
    M.fown = D.fown;
    M.fmark = D.fmark;
    M.cell_ptr = D.cell_ptr;
    M.cell_faces = D.cell_faces;
    M.cc = D.cc;
    M.cvp = D.cvp;
    {
        KScope ks(e, "ipt_integrate");
        k_ipt_integrate<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(M, *s, int(n), d_in.p, d_last.p, d_steps.p, cap ? d_rec.p : nullptr,
                                                                  (long long)record_cap, d_nrec.p, d_tally.p);
        FJ_CUDA(cudaGetLastError());
    }
    unsigned long long tally[2] = {0, 0};
    if (last)
        FJ_CUDA(cudaMemcpyAsync(last, d_last.p, nn * sizeof(FjsphIptPoint), cudaMemcpyDeviceToHost, e->stream));
    if (n_steps)
        FJ_CUDA(cudaMemcpyAsync(n_steps, d_steps.p, nn * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if (n_records)
        FJ_CUDA(cudaMemcpyAsync(n_records, d_nrec.p, nn * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if (cap)
        FJ_CUDA(cudaMemcpyAsync(records, d_rec.p, nn * cap * sizeof(FjsphIptPoint), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaMemcpyAsync(tally, d_tally.p, sizeof(tally), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (n_success)
        *n_success = int64_t(tally[0]);
    if (n_failed)
        *n_failed = int64_t(tally[1]);
    return FJSPH_OK;
}
