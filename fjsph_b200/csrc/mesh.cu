// mesh.cu — aero-mesh containment on the device: which CFD cell holds each free surface particle.
//
// Replaces, for the aero source "meshInfl" (reference src/Resid.cpp:471-523):
//   Crossings3D      reference src/Geometry.cpp:485-575   ray / triangle crossing by signed tetrahedron volumes
//   CheckCell        reference src/Containment.cpp:385-420 crossing parity of a +x ray over the cell's faces
//   FindCell         reference src/Containment.cpp:579-820 previous cell, then the 5 / 500 nearest cell centres,
//                                                          then boundary faces (inner wall / outer boundary / lost)
//   FirstCell        reference src/Containment.cpp:425-573 PIPE -> FREE transition (150 nearest cell centres; 20 in 2D)
// and in the 2D build (-DSIMDIM=2: faces are EDGES of two vertices, z = 0 throughout):
//   Crossings2D      reference src/Geometry.cpp:354-399   the +x ray of the crossings test against one edge
//   get_line_intersection reference src/Geometry.cpp:312-341 segment point -> cell centre against a boundary edge (its
//                                                          denominator test is one-sided, as in the reference)
// The reference finds nearest cell centres with a second nanoflann KD-tree (FJSPH.cpp:148); here the centres are
// binned into a uniform grid on the host once per mesh and a thread walks Chebyshev shells of bins around its
// particle.  "The first of the k nearest centres whose cell contains the point" is evaluated as: the containing
// cell of smallest (distance, index), accepted if fewer than k centres are closer (ties by cell index).
// The 4x4 determinants are evaluated as -det3 of the edge vectors without FMA contraction (only their sign is
// used), so the device and the CPU oracle take the same branch on the same bits.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.cuh"

namespace
{
constexpr int TPB = 128;

struct MeshView
{
    int n_cells;
    int dim;                          // 3, or 2 (edges: v0 and v1 only)
    const double4* __restrict__ fx;   // 3 per face: {v0.xyz, v1.x} {v1.yz, v2.xy} {v2.z, vlast.xyz}
    const int* __restrict__ fmark;    // leftright.second: neighbour cell, -1 inner wall, -2 outer boundary
    const int* __restrict__ cell_ptr;
    const int* __restrict__ cell_faces;
    const double4* __restrict__ cc;   // {centre.xyz, cRho}
    const double4* __restrict__ cvp;  // {cVel.xyz, cP}
    double ox, oy, oz, bin, inv_bin;
    double hx, hy, hz;                // upper corner of the binned box
    int bx, by, bz;
    const int* __restrict__ bin_start;
    const int* __restrict__ bin_cells;
};

struct V3
{
    double x, y, z;
};

// -det3[(a-p); (b-p); (c-p)], no contraction: the sign of | p 1; a 1; b 1; c 1 |
__device__ __forceinline__ double det4_sign_arg(const V3& p, const V3& a, const V3& b, const V3& c)
{
    const double ux = __dsub_rn(a.x, p.x), uy = __dsub_rn(a.y, p.y), uz = __dsub_rn(a.z, p.z);
    const double vx = __dsub_rn(b.x, p.x), vy = __dsub_rn(b.y, p.y), vz = __dsub_rn(b.z, p.z);
    const double wx = __dsub_rn(c.x, p.x), wy = __dsub_rn(c.y, p.y), wz = __dsub_rn(c.z, p.z);
    const double t0 = __dmul_rn(ux, __dsub_rn(__dmul_rn(vy, wz), __dmul_rn(vz, wy)));
    const double t1 = __dmul_rn(uy, __dsub_rn(__dmul_rn(vx, wz), __dmul_rn(vz, wx)));
    const double t2 = __dmul_rn(uz, __dsub_rn(__dmul_rn(vx, wy), __dmul_rn(vy, wx)));
    return -__dadd_rn(__dsub_rn(t0, t1), t2);
}

__device__ int crossings3d(const MeshView& M, int f, const V3& testp, const V3& rayp)
{
    const double4 a = M.fx[3 * f], b = M.fx[3 * f + 1], c = M.fx[3 * f + 2];
    const V3 f0 = {a.x, a.y, a.z}, f1 = {a.w, b.x, b.y}, f2 = {b.z, b.w, c.x}, fl = {c.y, c.z, c.w};
    const int flag1 = det4_sign_arg(testp, f0, f1, f2) < 0.0;
    const int flag2 = det4_sign_arg(rayp, f0, f1, f2) < 0.0;
    if (flag1 == flag2)
        return 0;
    /* Q6: edges (last,0), (0,1), (1,2) */
    const int flag3 = det4_sign_arg(testp, fl, f0, rayp) < 0.0;
    if ((det4_sign_arg(testp, f0, f1, rayp) < 0.0) != flag3)
        return 0;
    if ((det4_sign_arg(testp, f1, f2, rayp) < 0.0) != flag3)
        return 0;
    return 1;
}

// Crossings2D (Geometry.cpp:354-399), products rounded one by one as the CPU evaluates them
__device__ int crossings2d(const MeshView& M, int f, const V3& p)
{
    const double4 a = M.fx[3 * f], b = M.fx[3 * f + 1];
    const double v0x = a.x, v0y = a.y, v1x = a.w, v1y = b.x;
    const int yflag0 = (v0y >= p.y), yflag1 = (v1y >= p.y);
    if (yflag0 == yflag1)
        return 0;
    const double lhs = __dmul_rn(__dsub_rn(v1y, p.y), __dsub_rn(v1x, v0x));
    const double rhs = __dmul_rn(__dsub_rn(v1x, p.x), __dsub_rn(v1y, v0y));
    return (int(lhs >= rhs) == yflag1) ? 1 : 0;
}

// get_line_intersection (Geometry.cpp:312-341): segment p -> c against the edge of face f
__device__ int line_intersection2d(const MeshView& M, int f, const V3& p, const V3& c)
{
    const double4 a = M.fx[3 * f], b = M.fx[3 * f + 1];
    const double e1x = a.x, e1y = a.y, e2x = a.w, e2y = b.x;
    const double sx = __dsub_rn(c.x, p.x), sy = __dsub_rn(c.y, p.y);
    const double rx = __dsub_rn(e2x, e1x), ry = __dsub_rn(e2y, e1y);
    const double denom = __dadd_rn(__dmul_rn(-rx, sy), __dmul_rn(sx, ry));
    if (denom < 2.220446049250313e-16) /* MEPSILON: collinear, or any negative denominator */
        return 0;
    const double dx = __dsub_rn(p.x, e1x), dy = __dsub_rn(p.y, e1y);
    const double u = __ddiv_rn(__dadd_rn(__dmul_rn(-sy, dx), __dmul_rn(sx, dy)), denom);
    const double t = __ddiv_rn(__dsub_rn(__dmul_rn(rx, dy), __dmul_rn(ry, dx)), denom);
    return (u > 0.0 && u < 1.0 && t > 0.0 && t < 1.0) ? 1 : 0;
}

// the containment test of one face (ray along +x from the point) and the boundary test (segment point -> cell centre)
__device__ __forceinline__ int face_contains_ray(const MeshView& M, int f, const V3& p)
{
    if (M.dim == 2)
        return crossings2d(M, f, p);
    const V3 rayp = {p.x + 1e+5, p.y, p.z};
    return crossings3d(M, f, p, rayp);
}
__device__ __forceinline__ int face_cut_by_segment(const MeshView& M, int f, const V3& p, const V3& rayp)
{
    return (M.dim == 2) ? line_intersection2d(M, f, p, rayp) : crossings3d(M, f, p, rayp);
}

__device__ bool check_cell(const MeshView& M, int cell, const V3& p)
{
    if (cell < 0 || cell >= M.n_cells) /* Q7 */
        return false;
    unsigned line_flag = 0, inside = 0;
    for (int k = M.cell_ptr[cell]; k < M.cell_ptr[cell + 1]; ++k)
        if (face_contains_ray(M, M.cell_faces[k], p))
        {
            inside = !inside;
            if (line_flag)
                break; /* convex assumption */
            line_flag = 1;
        }
    return inside != 0;
}

__device__ __forceinline__ double centre_d2(const MeshView& M, int c, const V3& p)
{
    /* nanoflann L2_Simple order */
    const double4 q = M.cc[c];
    const double dx = __dsub_rn(p.x, q.x), dy = __dsub_rn(p.y, q.y), dz = __dsub_rn(p.z, q.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ bool before(double d2a, int ia, double d2b, int ib)
{
    return d2a < d2b || (d2a == d2b && ia < ib);
}

struct Search
{
    int bi, bj, bk; // bin of the (clamped) point
    int s_end;      // cube of shells <= s_end has been established as covering the k nearest
    double dq;      // Chebyshev distance from the point to the binned box (0 inside)
};

template <class F>
__device__ void for_cells_in_shells(const MeshView& M, const Search& S, int s_lo, int s_hi, F&& f)
{
    for (int k = max(S.bk - s_hi, 0); k <= min(S.bk + s_hi, M.bz - 1); ++k)
        for (int j = max(S.bj - s_hi, 0); j <= min(S.bj + s_hi, M.by - 1); ++j)
            for (int i = max(S.bi - s_hi, 0); i <= min(S.bi + s_hi, M.bx - 1); ++i)
            {
                const int cheb = max(max(abs(i - S.bi), abs(j - S.bj)), abs(k - S.bk));
                if (cheb < s_lo)
                    continue;
                const int b = (k * M.by + j) * M.bx + i;
                for (int t = M.bin_start[b]; t < M.bin_start[b + 1]; ++t) f(M.bin_cells[t]);
            }
}

// Containing cell among the k nearest centres (-1 if none); S describes the scanned cube afterwards.
__device__ int find_containing(const MeshView& M, const V3& p, int k, Search& S)
{
    S.bi = min(max(int(floor((p.x - M.ox) * M.inv_bin)), 0), M.bx - 1);
    S.bj = min(max(int(floor((p.y - M.oy) * M.inv_bin)), 0), M.by - 1);
    S.bk = min(max(int(floor((p.z - M.oz) * M.inv_bin)), 0), M.bz - 1);
    S.dq = fmax(fmax(fmax(M.ox - p.x, p.x - M.hx), fmax(M.oy - p.y, p.y - M.hy)), fmax(fmax(M.oz - p.z, p.z - M.hz), 0.0));
    const int s_max = max(max(M.bx, M.by), M.bz);
    double best_d2 = 1e300;
    int best = -1;
    int s = 0;
    for (;; ++s)
    {
        for_cells_in_shells(M, S, s, s, [&](int c) {
            const double d2 = centre_d2(M, c, p);
            if (before(d2, c, best_d2, best < 0 ? 0x7fffffff : best) && check_cell(M, c, p))
            {
                best_d2 = d2;
                best = c;
            }
        });
        /* every centre within rc of the point lies in the scanned cube */
        const double rc = double(s) * M.bin - S.dq;
        if (best >= 0 && rc > 0.0 && best_d2 <= rc * rc)
            break;
        if (s >= s_max)
            break;
        if (rc > 0.0)
        {
            int cnt = 0;
            const double rc2 = rc * rc;
            for_cells_in_shells(M, S, 0, s, [&](int c) { cnt += centre_d2(M, c, p) <= rc2; });
            if (cnt >= k)
                break;
        }
    }
    S.s_end = s;
    if (best < 0)
        return -1;
    int rank = 0;
    for_cells_in_shells(M, S, 0, s, [&](int c) { rank += before(centre_d2(M, c, p), c, best_d2, best); });
    return rank < k ? best : -1;
}

// next centre after (last_d2, last) in ascending (distance, index) order inside the scanned cube; -1 when exhausted
__device__ int next_nearest(const MeshView& M, const Search& S, const V3& p, double& last_d2, int last)
{
    double nd2 = 1e300;
    int nc = -1;
    for_cells_in_shells(M, S, 0, S.s_end, [&](int c) {
        const double d2 = centre_d2(M, c, p);
        if (before(last_d2, last, d2, c) && before(d2, c, nd2, nc < 0 ? 0x7fffffff : nc))
        {
            nd2 = d2;
            nc = c;
        }
    });
    if (nc >= 0)
        last_d2 = nd2;
    return nc;
}

__device__ __forceinline__ void take_cell(Level& L, int i, const MeshView& M, int cell)
{
    L.cellID[i] = cell;
    L.CV[i] = M.cvp[cell];
    double4 th = L.TH[i];
    th.w = M.cc[cell].w;
    L.TH[i] = th;
}

// FindCell: internal[] carries `internal` in its low byte and ipt_n_failed above it
__global__ void __launch_bounds__(TPB)
    k_find_cell(Level L, const int* __restrict__ blk, const int* __restrict__ oidx, int n_bound_blocks, int n, MeshView M,
                double lam_cutoff, unsigned* __restrict__ del_by_caller, int* __restrict__ n_del)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] < n_bound_blocks)
        return;
    if (L.b[i] != FJSPH_FREE || L.NP[i].w > lam_cutoff)
    {
        L.cellID[i] = -3;
        return;
    }
    const double4 x = L.P0[i];
    const V3 p = {x.x, x.y, x.z};
    int flags = L.internal[i];
    const int prev = L.cellID[i];
    if (check_cell(M, prev, p))
    {
        take_cell(L, i, M, prev);
        L.internal[i] = flags & 0xFF; /* ipt_n_failed = 0 */
        return;
    }
    Search S;
    const int cell = find_containing(M, p, 500, S);
    if (cell >= 0)
    {
        take_cell(L, i, M, cell);
        L.internal[i] = 0; /* ipt_n_failed = 0, internal = 0 */
        return;
    }
    /* across a boundary?  rays from the point to the 500 nearest cell centres against their boundary faces */
    unsigned cross = 0;
    bool del = false;
    double last_d2 = -1.0;
    int last = -1;
    for (int t = 0; t < 500; ++t)
    {
        last = next_nearest(M, S, p, last_d2, last);
        if (last < 0)
            break;
        const double4 c = M.cc[last];
        const V3 rayp = {c.x, c.y, c.z};
        for (int k = M.cell_ptr[last]; k < M.cell_ptr[last + 1]; ++k)
        {
            const int f = M.cell_faces[k];
            const int mark = M.fmark[f];
            if (mark < 0 && face_cut_by_segment(M, f, p, rayp))
            {
                cross = !cross;
                if (mark == -1)
                {
                    flags = (flags & ~0xFF) | 1; /* internal = 1 */
                    break;
                }
                else if (mark == -2)
                {
                    del = true;
                    break;
                }
            }
        }
    }
    if (cross == 0)
    {
        if ((flags >> 8) > 10)
            del = true;
        else
            flags += 1 << 8; /* ipt_n_failed++ */
    }
    L.internal[i] = flags;
    if (del)
    {
        del_by_caller[oidx[i]] = 1u;
        atomicAdd(n_del, 1);
    }
}

// Check_Pipe_Outlet with a mesh (Containment.cpp:822-847): PIPE -> FREE past the aero plane, then FirstCell
__global__ void __launch_bounds__(TPB)
    k_pipe_outlet_mesh(Level L, const int* __restrict__ blk, const int* __restrict__ oidx, int block, double nx, double ny,
                       double nz, double aeroconst, int n, MeshView M, double lam_cutoff,
                       unsigned* __restrict__ del_by_caller, int* __restrict__ counters /* [0] deleted, [1] not found */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || blk[i] != block || L.b[i] != FJSPH_PIPE)
        return;
    double4 x = L.P0[i];
    if (!(x.x * nx + x.y * ny + x.z * nz > aeroconst))
        return;
    L.b[i] = FJSPH_FREE;
    if (!(L.NP[i].w < lam_cutoff))
        return;
    const V3 p = {x.x, x.y, x.z};
    Search S;
    const int k_first = (M.dim == 2) ? 20 : 150; /* Containment.cpp:431-436 */
    const int cell = find_containing(M, p, k_first, S);
    if (cell >= 0)
    {
        take_cell(L, i, M, cell);
        return;
    }
    unsigned cross = 0;
    bool del = false;
    double last_d2 = -1.0;
    int last = -1;
    double4 v = L.P1[i];
    for (int t = 0; t < k_first; ++t)
    {
        last = next_nearest(M, S, p, last_d2, last);
        if (last < 0)
            break;
        const double4 c = M.cc[last];
        const V3 rayp = {c.x, c.y, c.z};
        for (int k = M.cell_ptr[last]; k < M.cell_ptr[last + 1]; ++k)
        {
            const int f = M.cell_faces[k];
            const int mark = M.fmark[f];
            if (mark < 0 && face_cut_by_segment(M, f, p, rayp))
            {
                cross = !cross;
                if (mark == -1)
                {
                    /* reflect the velocity off the wall and put the particle on its plane (Containment.cpp:521-543) */
                    const double4 a = M.fx[3 * f], b = M.fx[3 * f + 1], cc = M.fx[3 * f + 2];
                    const double r1x = a.w - a.x, r1y = b.x - a.y, r1z = b.y - a.z;
                    const double r2x = b.z - a.x, r2y = b.w - a.y, r2z = cc.x - a.z;
                    double qx = r1y * r2z - r1z * r2y, qy = r1z * r2x - r1x * r2z, qz = r1x * r2y - r1y * r2x;
                    if (M.dim == 2) /* Containment.cpp:540-544: norm = (-r1.y, r1.x) */
                    {
                        qx = -r1y;
                        qy = r1x;
                        qz = 0.0;
                    }
                    const double qq = qx * qx + qy * qy + qz * qz;
                    if (qq > 0.0)
                    {
                        const double inv = 1.0 / sqrt(qq);
                        qx *= inv;
                        qy *= inv;
                        qz *= inv;
                    }
                    const double vd = 2.0 * (v.x * qx + v.y * qy + v.z * qz);
                    v.x -= vd * qx;
                    v.y -= vd * qy;
                    v.z -= vd * qz;
                    const double plane = qx * a.w + qy * b.x + qz * b.y;
                    const double dist = (plane - (x.x * qx + x.y * qy + x.z * qz)) / (qx * qx + qy * qy + qz * qz);
                    x.x += dist * qx;
                    x.y += dist * qy;
                    x.z += dist * qz;
                }
                else if (mark == -2)
                    del = true;
            }
        }
    }
    L.P0[i] = x;
    L.P1[i] = v;
    if (cross == 0)
        atomicAdd(&counters[1], 1);
    if (del)
    {
        del_by_caller[oidx[i]] = 1u;
        atomicAdd(&counters[0], 1);
    }
}

MeshView view_of(const DeviceMesh& D)
{
    MeshView M;
    M.dim = D.dim;
    M.n_cells = D.n_cells;
    M.fx = D.fx;
    M.fmark = D.fmark;
    M.cell_ptr = D.cell_ptr;
    M.cell_faces = D.cell_faces;
    M.cc = D.cc;
    M.cvp = D.cvp;
    M.ox = D.ox;
    M.oy = D.oy;
    M.oz = D.oz;
    M.hx = D.hx;
    M.hy = D.hy;
    M.hz = D.hz;
    M.bin = D.bin;
    M.inv_bin = 1.0 / D.bin;
    M.bx = D.bx;
    M.by = D.by;
    M.bz = D.bz;
    M.bin_start = D.bin_start;
    M.bin_cells = D.bin_cells;
    return M;
}

template <class T>
int to_device(const std::vector<T>& h, T** d)
{
    if (*d)
        cudaFree(*d);
    *d = nullptr;
    FJ_CUDA(cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    FJ_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FJSPH_OK;
}

} // namespace

void fj_free_mesh(FjsphEngine* e)
{
    DeviceMesh& D = e->mesh;
    void* ptrs[] = {D.fx, D.fmark, D.fown, D.fq, D.cell_ptr, D.cell_faces, D.cc, D.cvp, D.bin_start, D.bin_cells, D.counters};
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    D = DeviceMesh();
}

extern "C" int fjsph_upload_mesh(FjsphEngine* e, const FjsphMesh* m)
{
    const int dim = e->P.dim; /* 2: faces are edges, every z of verts / cCentre / cVel must be 0 */
    cudaSetDevice(e->device);
    if (!m || m->n_cells <= 0 || m->n_faces <= 0 || !m->verts || !m->face_ptr || !m->face_vtx || !m->leftright ||
        !m->cell_ptr || !m->cell_faces || !m->cCentre || !m->cVel || !m->cP || !m->cRho)
    {
        fj_set_error("upload_mesh: every array of the MESH is required (the OpenFOAM reader never fills cRho, Q8)");
        return FJSPH_ERR_INVALID;
    }
    if (m->n_cells > 0x7fffffff || m->n_faces > 0x7fffffff / 3)
    {
        fj_set_error("upload_mesh: mesh too large for 32-bit indices");
        return FJSPH_ERR_CAPACITY;
    }
    const size_t nf = size_t(m->n_faces), nc = size_t(m->n_cells);
    std::vector<double4> fx(3 * nf);
    std::vector<int> fmark(nf), fown(nf);
    std::vector<double4> fq(nf);
    for (size_t f = 0; f < nf; ++f)
    {
        const int64_t a = m->face_ptr[f], b = m->face_ptr[f + 1];
        if (b - a < dim)
        {
            fj_set_error("upload_mesh: face %zu has fewer than %d vertices", f, dim);
            return FJSPH_ERR_INVALID;
        }
        const int64_t id[4] = {m->face_vtx[a], m->face_vtx[a + 1], m->face_vtx[dim == 2 ? a + 1 : a + 2], m->face_vtx[b - 1]};
        double v[4][3];
        for (int k = 0; k < 4; ++k)
        {
            if (id[k] < 0 || id[k] >= m->n_verts)
            {
                fj_set_error("upload_mesh: vertex index out of range in face %zu", f);
                return FJSPH_ERR_INVALID;
            }
            for (int d = 0; d < 3; ++d) v[k][d] = m->verts[3 * id[k] + d];
            if (dim == 2 && v[k][2] != 0.0)
            {
                fj_set_error("upload_mesh: SIMDIM=2 mesh with a non-zero z coordinate (vertex %lld)", (long long)id[k]);
                return FJSPH_ERR_INVALID;
            }
        }
        fx[3 * f] = make_double4(v[0][0], v[0][1], v[0][2], v[1][0]);
        fx[3 * f + 1] = make_double4(v[1][1], v[1][2], v[2][0], v[2][1]);
        fx[3 * f + 2] = make_double4(v[2][2], v[3][0], v[3][1], v[3][2]);
        fmark[f] = m->leftright[2 * f + 1];
        fown[f] = m->leftright[2 * f];
        if (fown[f] < 0 || fown[f] >= m->n_cells || fmark[f] >= m->n_cells)
        {
            fj_set_error("upload_mesh: face %zu names a cell outside the mesh", f);
            return FJSPH_ERR_INVALID;
        }
        /* face[3] (not the last vertex: they differ on polygons of five and more corners) */
        const int64_t i3 = (b - a >= 4) ? m->face_vtx[a + 3] : id[3];
        if (i3 < 0 || i3 >= m->n_verts)
        {
            fj_set_error("upload_mesh: vertex index out of range in face %zu", f);
            return FJSPH_ERR_INVALID;
        }
        fq[f] = make_double4(m->verts[3 * i3], m->verts[3 * i3 + 1], m->verts[3 * i3 + 2], double(b - a));
    }
    if (m->cell_ptr[0] != 0 || m->cell_ptr[nc] < 0 || m->cell_ptr[nc] > 0x7fffffff)
    {
        fj_set_error("upload_mesh: cell_ptr is not a CSR offset array");
        return FJSPH_ERR_INVALID;
    }
    for (size_t c = 0; c < nc; ++c)
        if (m->cell_ptr[c + 1] < m->cell_ptr[c])
        {
            fj_set_error("upload_mesh: cell_ptr decreases at cell %zu", c);
            return FJSPH_ERR_INVALID;
        }
    std::vector<int> cptr(nc + 1), cfaces(size_t(m->cell_ptr[nc]));
    for (size_t c = 0; c <= nc; ++c) cptr[c] = int(m->cell_ptr[c]);
    for (size_t k = 0; k < cfaces.size(); ++k)
    {
        if (m->cell_faces[k] < 0 || m->cell_faces[k] >= m->n_faces)
        {
            fj_set_error("upload_mesh: cell_faces[%zu] is not a face of the mesh", k);
            return FJSPH_ERR_INVALID;
        }
        cfaces[k] = int(m->cell_faces[k]);
    }
    std::vector<double4> cc(nc), cvp(nc);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t c = 0; c < nc; ++c)
    {
        cc[c] = make_double4(m->cCentre[3 * c], m->cCentre[3 * c + 1], m->cCentre[3 * c + 2], m->cRho[c]);
        cvp[c] = make_double4(m->cVel[3 * c], m->cVel[3 * c + 1], m->cVel[3 * c + 2], m->cP[c]);
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = std::min(lo[d], m->cCentre[3 * c + d]);
            hi[d] = std::max(hi[d], m->cCentre[3 * c + d]);
        }
    }
    // bins over the cell centres: about 4 centres per bin
    DeviceMesh& D = e->mesh;
    const double ext[3] = {std::max(hi[0] - lo[0], 1e-300), std::max(hi[1] - lo[1], 1e-300), std::max(hi[2] - lo[2], 1e-300)};
    double bin = dim == 2 ? std::sqrt(ext[0] * ext[1] / double(nc) * 4.0) : std::cbrt(ext[0] * ext[1] * ext[2] / double(nc) * 4.0);
    if (!(bin > 0.0) || !std::isfinite(bin))
        bin = std::max(ext[0], std::max(ext[1], ext[2]));
    int nb[3];
    for (int d = 0; d < 3; ++d)
    {
        nb[d] = std::max(1, std::min(1024, int(std::floor(ext[d] / bin)) + 1));
    }
    std::vector<int> bstart(size_t(nb[0]) * nb[1] * nb[2] + 1, 0), bcells(nc);
    auto bin_of = [&](size_t c) {
        int ijk[3];
        for (int d = 0; d < 3; ++d)
            ijk[d] = std::min(std::max(int(std::floor((m->cCentre[3 * c + d] - lo[d]) / bin)), 0), nb[d] - 1);
        return (size_t(ijk[2]) * nb[1] + ijk[1]) * nb[0] + ijk[0];
    };
    for (size_t c = 0; c < nc; ++c) bstart[bin_of(c) + 1]++;
    for (size_t b = 0; b + 1 < bstart.size(); ++b) bstart[b + 1] += bstart[b];
    {
        std::vector<int> fill(bstart.begin(), bstart.end() - 1);
        for (size_t c = 0; c < nc; ++c) bcells[size_t(fill[bin_of(c)]++)] = int(c); /* ascending cell index per bin */
    }
    int st;
    if ((st = to_device(fx, &D.fx)) || (st = to_device(fmark, &D.fmark)) || (st = to_device(fown, &D.fown)) ||
        (st = to_device(fq, &D.fq)) || (st = to_device(cptr, &D.cell_ptr)) ||
        (st = to_device(cfaces, &D.cell_faces)) || (st = to_device(cc, &D.cc)) || (st = to_device(cvp, &D.cvp)) ||
        (st = to_device(bstart, &D.bin_start)) || (st = to_device(bcells, &D.bin_cells)))
        return st;
    if (!D.counters)
        FJ_CUDA(cudaMalloc(&D.counters, 4 * sizeof(int)));
    D.dim = dim;
    D.n_cells = int(nc);
    D.n_faces = int(nf);
    D.ox = lo[0];
    D.oy = lo[1];
    D.oz = lo[2];
    D.bin = bin;
    D.bx = nb[0];
    D.by = nb[1];
    D.bz = nb[2];
    D.hx = lo[0] + bin * nb[0];
    D.hy = lo[1] + bin * nb[1];
    D.hz = lo[2] + bin * nb[2];
    D.loaded = true;
    return FJSPH_OK;
}

// get_aero_velocity, aero source meshInfl (Resid.cpp:480-523): FindCell, erase the escaped particles from both time
// levels, then redo the neighbour list and the prestep.
int fj_aero_velocity_mesh(FjsphEngine* e)
{
    DeviceMesh& D = e->mesh;
    if (!D.loaded)
    {
        fj_set_error("aero source meshInfl needs a mesh: call fjsph_upload_mesh first");
        return FJSPH_ERR_STATE;
    }
    const bool slabs = e->slab.on && e->slab.world > 1;
    const int n = int(e->n_owned);
    unsigned* d_del = reinterpret_cast<unsigned*>(e->key);
    FJ_CUDA(cudaMemsetAsync(d_del, 0, size_t(n) * sizeof(unsigned), e->stream));
    FJ_CUDA(cudaMemsetAsync(D.counters, 0, 4 * sizeof(int), e->stream));
    {
        KScope ks(e, "find_cell", 1);
        k_find_cell<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1], e->blk, e->oidx, e->n_bound_blocks, n, view_of(D),
                                                             e->P.lam_cutoff, d_del, D.counters);
    }
    FJ_CUDA(cudaGetLastError());
    int h_cnt[4];
    FJ_CUDA(cudaMemcpyAsync(h_cnt, D.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (slabs)
    {
        /* Slab decomposition (the mesh is replicated, every rank looks its own particles up): the ranks erase together.
           The escaped particles are handed to the re-decomposition, which drops them while it re-makes the owned and
           ghost sets and the global counts; then, as on one GPU, the lists and the prestep are redone. */
        double any = double(h_cnt[0]);
        int st = fj_allreduce(e, FJSPH_COMM_SUM, &any, 1);
        if (st)
            return st;
        if (any > 0.0)
        {
            e->mesh_deleted += h_cnt[0];
            e->slab.del_by_caller = d_del; /* e->key is not written again before fj_redecompose has read it */
            e->skin_valid = false;
            st = fj_build_neighbours(e);
            if (st)
                return st;
            st = fj_prestep(e, nullptr); /* the caller exchanges the prestep's ghost fields next (frozen_terms) */
            if (st)
                return st;
        }
        return FJSPH_OK;
    }
    if (h_cnt[0] > 0)
    {
        int n_del = 0;
        int st = fj_delete_flagged(e, d_del, true, &n_del);
        if (st)
            return st;
        e->mesh_deleted += n_del;
        st = fj_build_neighbours(e);
        if (st)
            return st;
        st = fj_prestep(e, nullptr);
        if (st)
            return st;
    }
    return FJSPH_OK;
}

// Check_Pipe_Outlet with the mesh: one launch per fluid block that defines an aero plane
int fj_pipe_outlet_mesh(FjsphEngine* e)
{
    e->x_moved = true; /* a particle reflected off an inner wall is put on the wall's plane (k_pipe_outlet_mesh) */
    DeviceMesh& D = e->mesh;
    if (!D.loaded)
    {
        fj_set_error("aero source meshInfl needs a mesh: call fjsph_upload_mesh first");
        return FJSPH_ERR_STATE;
    }
    if (e->slab.on && e->slab.world > 1)
    {
        for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
            if (e->blocks[bl].aeroconst != 9999999.0)
            {
                fj_set_error("pipe blocks with an aero entry plane are not available with slab decomposition yet");
                return FJSPH_ERR_INVALID;
            }
        return FJSPH_OK;
    }
    const int n = int(e->n_owned);
    unsigned* d_del = reinterpret_cast<unsigned*>(e->key);
    bool any = false;
    for (size_t bl = size_t(e->n_bound_blocks); bl < e->blocks.size(); ++bl)
    {
        const HostBlock& B = e->blocks[bl];
        if (B.aeroconst == 9999999.0)
            continue;
        if (!any)
        {
            FJ_CUDA(cudaMemsetAsync(d_del, 0, size_t(n) * sizeof(unsigned), e->stream));
            FJ_CUDA(cudaMemsetAsync(D.counters, 0, 4 * sizeof(int), e->stream));
            any = true;
        }
        KScope ks(e, "pipe_outlet", 1);
        k_pipe_outlet_mesh<<<fj_blocks(n, TPB), TPB, 0, e->stream>>>(e->lv[1], e->blk, e->oidx, int(bl), B.aero_norm[0],
                                                                    B.aero_norm[1], B.aero_norm[2], B.aeroconst, n,
                                                                    view_of(D), e->P.lam_cutoff, d_del, D.counters);
    }
    if (!any)
        return FJSPH_OK;
    FJ_CUDA(cudaGetLastError());
    int h_cnt[4];
    FJ_CUDA(cudaMemcpyAsync(h_cnt, D.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FJ_CUDA(cudaStreamSynchronize(e->stream));
    if (h_cnt[1] > 0)
    {
        fj_set_error("first containing cell not found for %d particle(s) leaving the pipe (the reference exits here, "
                     "Containment.cpp:563-571)", h_cnt[1]);
        return FJSPH_ERR_STATE;
    }
    if (h_cnt[0] > 0)
    {
        int n_del = 0;
        int st = fj_delete_flagged(e, d_del, true, &n_del);
        if (st)
            return st;
        e->mesh_deleted += n_del;
        return fj_build_neighbours(e);
    }
    return FJSPH_OK;
}
