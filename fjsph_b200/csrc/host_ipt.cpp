// host_ipt.cpp — host-side settings of the implicit particle tracker (ipt.cu).
//
// Mirrors:
//   IPT_SETT defaults                          reference src/Var.h:313-337
//   the IPT keys of GetInput's para parser      reference src/IO.cpp:447-453, checks at IO.cpp:666-680
//   ipt_diam / ipt_area, max_x *= scale         reference src/IO.cpp:29,126-127
//   cells.maxlength                             reference src/CDFIO.cpp:867-898,931 (edges), 1117-1183,1214 (faces)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/fjsph_b200.h"

void fj_set_error(const char* fmt, ...);

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

extern "C" int fjsph_ipt_default_settings(const FjsphParams* p, FjsphIptSettings* s)
{
    if (!p || !s)
    {
        fj_set_error("ipt_default_settings: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    *s = FjsphIptSettings();
    s->eq_order = 2;
    s->max_subits = p->max_subits;
    s->record = 1; /* streak_out = 1 */
    s->max_steps = 100000; /* far beyond any mesh crossing; bounds what a particle caught between two cells can cost */
    s->relax = 0.6;
    s->n_relax = 5;
    s->max_x = 9999999;
    s->max_length = 0.0;
    s->diam = std::pow((6.0 * p->sim_mass) / (M_PI * p->rho_rest), 1.0 / 3.0);
    s->area = M_PI * s->diam * s->diam / 4.0;
    for (int d = 0; d < 3; ++d) s->grav[d] = p->grav[d];
    s->mu_g = p->mu_g;
    s->rho_rest = p->rho_rest;
    return FJSPH_OK;
}

extern "C" int fjsph_read_para_ipt(const char* path, double scale, int32_t* using_ipt, FjsphIptSettings* s)
{
    if (!path || !using_ipt || !s)
    {
        fj_set_error("read_para_ipt: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    std::ifstream fin(path);
    if (!fin.is_open())
    {
        fj_set_error("could not open SPH parameter file \"%s\"", path);
        return FJSPH_ERR_IO;
    }
    auto trim = [](std::string t) {
        const size_t a = t.find_first_not_of(" \t\r\n"), b = t.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
    };
    int use = 0, order = s->eq_order;
    unsigned part_out = 0, streak_out = 1, cells_out = 0;
    double max_x = 9999999, max_x_sph = 9999999;
    std::string line;
    while (std::getline(fin, line))
    {
        const size_t hash = line.find('#');
        if (hash != std::string::npos)
            line = line.substr(0, hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos)
            continue;
        const std::string key = trim(line.substr(0, colon));
        std::istringstream val(trim(line.substr(colon + 1)));
        if (key == "Transition to IPT (0/1)")
            val >> use;
        else if (key == "Velocity equation order (1/2)")
            val >> order;
        else if (key == "SPH tracking conversion x coordinate")
            val >> max_x_sph;
        else if (key == "Maximum x trajectory coordinate")
            val >> max_x;
        else if (key == "Particle scatter output (0/1/2)")
            val >> part_out;
        else if (key == "Particle streak output (0/1/2)")
            val >> streak_out;
        else if (key == "Particle cell intersection output (0/1/2)")
            val >> cells_out;
    }
    if (use)
    {
        if (order > 2 || order < 1) /* the reference exits here (IO.cpp:668-672) */
        {
            fj_set_error("Equation order not 1 or 2. Please choose between these.");
            return FJSPH_ERR_INVALID;
        }
        if (max_x < max_x_sph) /* IO.cpp:674-679: a warning, and no tracking */
            use = 0;
    }
    *using_ipt = use;
    s->eq_order = order;
    s->max_x = max_x * scale;
    s->record = (streak_out == 1 || cells_out == 1) ? 1 : 0;
    (void)part_out; /* the scatter file is an output format of the host */
    return FJSPH_OK;
}

extern "C" int fjsph_mesh_max_length(const FjsphMesh* m, int32_t dim, double* max_length)
{
    if (!m || !max_length || !m->verts || !m->face_ptr || !m->face_vtx || (dim != 2 && dim != 3))
    {
        fj_set_error("mesh_max_length: bad arguments");
        return FJSPH_ERR_INVALID;
    }
    auto dist = [&](int64_t a, int64_t b) {
        double s = 0.0;
        for (int d = 0; d < dim; ++d)
        {
            const double t = m->verts[3 * a + d] - m->verts[3 * b + d];
            s += t * t;
        }
        return std::sqrt(s);
    };
    double longest = 0.0;
    for (int64_t f = 0; f < m->n_faces; ++f)
    {
        const int64_t a = m->face_ptr[f], k = m->face_ptr[f + 1] - a;
        if (k < dim)
        {
            fj_set_error("mesh_max_length: face %lld has fewer than %d vertices", (long long)f, dim);
            return FJSPH_ERR_INVALID;
        }
        for (int64_t j = 0; j < std::min<int64_t>(k, 4); ++j)
            if (m->face_vtx[a + j] < 0 || m->face_vtx[a + j] >= m->n_verts)
            {
                fj_set_error("mesh_max_length: vertex index out of range in face %lld", (long long)f);
                return FJSPH_ERR_INVALID;
            }
        const int64_t* v = m->face_vtx + a;
        double e;
        if (dim == 2)
            e = dist(v[0], v[1]);
        else if (k == 3)
            e = std::max(dist(v[0], v[1]), std::max(dist(v[0], v[2]), dist(v[1], v[2])));
        else
            e = std::max(dist(v[0], v[2]), dist(v[1], v[3]));
        longest = std::max(longest, e);
    }
    /* the readers leave a multiple of it: maxedge *= 5.0 for faces (CDFIO.cpp:1214), maxedge *= 4.0 for edges (CDFIO.cpp:931) */
    *max_length = longest * (dim == 2 ? 4.0 : 5.0);
    return FJSPH_OK;
}
