// host_mesh.h -- the MESH arrays a mesh reader (host_foam.cpp, host_tau.cpp) hands to fjsph_upload_mesh through
// fjsph_foam_view: vertices, faces as CSR vertex lists, (left, right) cells per face, cells as CSR face lists, and the
// per-cell centre, velocity, pressure and density (reference MESH, Var.h:396-451).
#pragma once
#include <cstdint>
#include <vector>

struct FjsphFoamMesh
{
    std::vector<double> verts, cCentre, cVel, cP, cRho;
    std::vector<int64_t> face_ptr, face_vtx, cell_ptr, cell_faces;
    std::vector<int32_t> leftright;
    int64_t n_quads_split = 0;
};
