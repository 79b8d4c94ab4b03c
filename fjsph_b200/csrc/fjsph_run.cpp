// fjsph_run.cpp — command-line driver: FJSPH's main() (reference src/FJSPH.cpp:29-356) on top of the C ABI.
//
//   fjsph_b200_run <para file> [--frames N] [--device K] [--restart file.fjr] [--out prefix] [--quiet]
//
// GetInput + Init_Particles (fjsph_case_read) -> engine -> integrate_no_update at t = 0 (FJSPH.cpp:183) -> the frame
// loop `while (stept + 0.1 dt_min < frame_dt) integrate` (FJSPH.cpp:262-283) with the reference's per-step table
// (Integration.cpp:250-265), `<prefix>_frame.info` (FJSPH.cpp:228-243,286-300), one ASCII Tecplot zone file per frame
// (the fallback of AsciiIO.cpp; TECIO / HDF5 are not available here) and a restart file after every frame
// (csrc/restart.cu).  Everything numerical happens behind fjsph_step; this file is host bookkeeping only.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"

static int fail(const char* what)
{
    std::fprintf(stderr, "ERROR: %s: %s\n", what, fjsph_last_error());
    return 1;
}

static int write_frame(FjsphEngine* e, const std::string& prefix, int frame, double time, double rho_rest)
{
    const int64_t n = fjsph_count(e);
    std::vector<double> xi(3 * n), v(3 * n), p(n), rho(n);
    std::vector<int32_t> b(n), surf(n);
    std::vector<int64_t> pid(n);
    FjsphStateView s;
    std::memset(&s, 0, sizeof(s));
    s.n = n;
    s.xi = xi.data();
    s.v = v.data();
    s.p = p.data();
    s.rho = rho.data();
    s.b = b.data();
    s.surf = surf.data();
    s.part_id = pid.data();
    if (fjsph_download_state(e, 1, &s))
        return 1;
    char name[1024];
    std::snprintf(name, sizeof(name), "%s_frame_%05d.dat", prefix.c_str(), frame);
    FILE* f = std::fopen(name, "w");
    if (!f)
        return 1;
    std::fprintf(f, "TITLE=\"%s\"\n", prefix.c_str());
    std::fprintf(f, "VARIABLES=\"X\" \"Y\" \"Z\" \"V-x\" \"V-y\" \"V-z\" \"Pressure\" \"Density\" \"Density variation\" "
                    "\"Boundary condition\" \"Surface flag\" \"Particle ID\"\n");
    std::fprintf(f, "ZONE T=\"Particles\", I=%lld, DATAPACKING=POINT, STRANDID=1, SOLUTIONTIME=%.7g\n", (long long)n, time);
    for (int64_t i = 0; i < n; ++i)
        std::fprintf(f, "%3.7e %3.7e %3.7e %3.7e %3.7e %3.7e %3.7e %3.7e %3.7e %d %d %lld\n", xi[3 * i], xi[3 * i + 1],
                     xi[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2], p[i], rho[i], 100.0 * (rho[i] / rho_rest - 1.0),
                     b[i], surf[i], (long long)pid[i]);
    std::fclose(f);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s <para file> [--frames N] [--device K] [--restart file] [--out prefix] [--quiet]\n", argv[0]);
        return 2;
    }
    const char* para = argv[1];
    int frames_cli = -1, device = 0;
    bool quiet = false;
    std::string restart_file, prefix_cli;
    for (int a = 2; a < argc; ++a)
    {
        const std::string k = argv[a];
        if (k == "--quiet")
            quiet = true;
        else if (a + 1 < argc && k == "--frames")
            frames_cli = std::atoi(argv[++a]);
        else if (a + 1 < argc && k == "--device")
            device = std::atoi(argv[++a]);
        else if (a + 1 < argc && k == "--restart")
            restart_file = argv[++a];
        else if (a + 1 < argc && k == "--out")
            prefix_cli = argv[++a];
        else
        {
            std::fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 2;
        }
    }
    FjsphCase* c = nullptr;
    if (fjsph_case_read(para, 3, &c))
        return fail("reading the case");
    FjsphParams P;
    fjsph_case_params(c, &P);
    int32_t max_frames = -1;
    int64_t max_points = -1;
    char out_prefix[512] = "", rst_prefix[512] = "";
    fjsph_case_io(c, &max_frames, &max_points, out_prefix, rst_prefix, 512);
    std::string prefix = !prefix_cli.empty() ? prefix_cli : (out_prefix[0] ? out_prefix : "fjsph_b200");
    if (frames_cli >= 0)
        max_frames = frames_cli;
    if (max_frames < 0)
    {
        std::fprintf(stderr, "ERROR: SPH frame count has not been defined (para key or --frames).\n");
        return 1;
    }
    /* aero mesh named by the deck, read on the host before anything touches the device, as FJSPH.cpp:70-100 does
       (FOAM::Read_FOAM, or TAU::Read_tau_mesh_FACE + Read_SOLUTION; Read_BMAP is part of fjsph_case_read); uploaded for
       the containment lookup once the particles are on the device */
    FjsphFoamMesh* aero_mesh = nullptr;
    {
        char fdir[1024] = "", fsol[1024] = "", tmesh[1024] = "", tsol[1024] = "";
        int32_t buoyant = 0;
        double tscale = 1.0;
        FjsphMesh view;
        fjsph_case_foam(c, fdir, fsol, &buoyant, 1024);
        fjsph_case_tau(c, tmesh, tsol, &tscale, 1024);
        if (tmesh[0])
        {
            if (fjsph_tau_read(tmesh, tsol, tscale, &aero_mesh) || fjsph_foam_view(aero_mesh, &view))
                return fail("reading the TAU mesh");
            std::printf("TAU mesh: %lld cells, %lld faces\n", (long long)view.n_cells, (long long)view.n_faces);
        }
        else if (fdir[0])
        {
            if (fjsph_foam_read(fdir, fsol, buoyant, P.rho_g, &aero_mesh) || fjsph_foam_view(aero_mesh, &view))
                return fail("reading the OpenFOAM case");
            std::printf("OpenFOAM mesh: %lld cells, %lld triangles\n", (long long)view.n_cells, (long long)view.n_faces);
        }
    }
    const int64_t n0 = fjsph_case_count(c), nb0 = fjsph_case_bound_points(c);
    const int64_t capacity = max_points > n0 ? max_points : 2 * n0 + 100000;
    FjsphEngine* e = nullptr;
    if (fjsph_create(&P, device, capacity, &e))
        return fail("creating the engine");
    int32_t frame = 0;
    if (restart_file.empty())
    {
        std::vector<double> xi(3 * n0), v(3 * n0), rho(n0), p(n0), m(n0);
        std::vector<int32_t> b(n0);
        std::vector<int64_t> pid(n0);
        FjsphStateView s;
        std::memset(&s, 0, sizeof(s));
        s.n = n0;
        s.xi = xi.data();
        s.v = v.data();
        s.rho = rho.data();
        s.p = p.data();
        s.m = m.data();
        s.b = b.data();
        s.part_id = pid.data();
        if (fjsph_case_state(c, &s) || fjsph_upload_state(e, &s, nb0))
            return fail("uploading the particles");
        const int32_t nblk = fjsph_case_num_blocks(c);
        std::vector<FjsphBlock> blocks(nblk);
        for (int32_t i = 0; i < nblk; ++i) fjsph_case_block(c, i, &blocks[i], nullptr, 0);
        if (nblk && fjsph_set_blocks(e, nblk, blocks.data()))
            return fail("setting the blocks");
    }
    else if (fjsph_read_restart(e, restart_file.c_str(), &frame))
        return fail("reading the restart file");
    if (aero_mesh)
    {
        FjsphMesh view;
        if (fjsph_foam_view(aero_mesh, &view) || fjsph_upload_mesh(e, &view))
            return fail("uploading the aero mesh");
        fjsph_foam_free(aero_mesh);
    }
    fjsph_get_params(e, &P);
    std::printf("Starting counts:\nBoundary: %lld  Sim: %lld\n\n", (long long)nb0, (long long)(fjsph_count(e) - nb0));

    const std::string info_name = prefix + "_frame.info";
    FILE* info = std::fopen(info_name.c_str(), restart_file.empty() ? "w" : "a");
    if (!info)
    {
        std::fprintf(stderr, "ERROR: cannot open %s\n", info_name.c_str());
        return 1;
    }
    FjsphStepStats st;
    std::memset(&st, 0, sizeof(st));
    const auto t1 = std::chrono::high_resolution_clock::now();
    auto seconds = [&]() { return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count(); };
    long long deleted = 0;
    if (restart_file.empty())
    {
        if (write_frame(e, prefix, 0, P.current_time, P.rho_rest))
            return fail("writing frame 0");
        if (fjsph_integrate_no_update(e, &st)) /* populate the force vectors, FJSPH.cpp:183 */
            return fail("first integrate_no_update");
        std::fprintf(info, "Frame: %u\nTotal Points: %lld Boundary Points: %lld Fluid Points: %lld\n", 0u,
                     (long long)fjsph_count(e), (long long)nb0, (long long)(fjsph_count(e) - nb0));
        std::fprintf(info, "Sim Time:  %.7g Comp Time: %.6e Error: %.6f Sub-iterations: %d\n", P.current_time, seconds(), 0.0, 0);
        std::fprintf(info, "Deleted particles: %lld Internal collisions: %d\n", deleted, 0);
    }
    double error = 0.0;
    for (int32_t fr = frame + 1; fr < max_frames; ++fr)
    {
        int stepits = 0;
        double stept = 0.0;
        while (stept + 0.1 * P.delta_t_min < P.frame_time_interval)
        {
            if (!quiet && stepits % 50 == 0)
                std::printf("\nTime      | Timestep | CFL  | RMS error | its | dRho (%%) | Max-F     | Max-Af    | Max Shift | Step time (ms)| \n");
            const auto s0 = std::chrono::high_resolution_clock::now();
            if (fjsph_step(e, &st))
                return fail("integrate");
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - s0).count();
            fjsph_get_params(e, &P);
            error = st.rms_error;
            deleted += st.n_del;
            if (!quiet)
                std::printf("%9.3e | %7.2e | %4.2f | %9.4f | %3d | %8.3f | %9.3e | %9.3e | %9.3e | %14ld|\n", P.current_time - st.dt,
                            st.dt, st.cfl_ratio, st.rms_error, st.iterations, st.maxRho_pc, st.maxf, st.maxAf, st.maxShift, long(ms));
            stept += st.dt;
            ++stepits;
        }
        const int64_t n = fjsph_count(e);
        std::fprintf(info, "\nFrame: %u\nTotal Points: %lld Boundary Points: %lld Fluid Points: %lld\n", unsigned(fr), (long long)n,
                     (long long)nb0, (long long)(n - nb0));
        std::fprintf(info, "Sim Time:  %.7g Comp Time: %.6e Error: %.6f Sub-iterations: %d\n", P.current_time, seconds(), error, stepits);
        std::fprintf(info, "Deleted particles: %lld Internal collisions: %d\n", deleted, 0);
        std::fflush(info);
        std::printf("Frame: %d  Sim Time: %.7g  Compute Time: %.3f  Error: %.5f\n", fr, P.current_time, seconds(), error);
        std::printf("Boundary particles:  %lld Sim particles: %lld Deleted particles: %lld\n", (long long)nb0, (long long)(n - nb0), deleted);
        if (n - nb0 == 0)
        {
            std::printf("No more points in the simulation space. Ending....\n");
            break;
        }
        if (write_frame(e, prefix, fr, P.current_time, P.rho_rest))
            return fail("writing a frame");
        /* march the frame time forward (FJSPH.cpp:326-327) BEFORE the checkpoint, so that a resumed run clamps its
           first steps to the end of the NEXT frame (find_timestep, Integration.cpp:433-440) */
        P.last_frame_time += P.frame_time_interval;
        if (fjsph_set_params(e, &P))
            return fail("set_params");
        if (fjsph_write_restart(e, (prefix + "_particles.fjr").c_str(), fr))
            return fail("writing the restart file");
    }
    std::fclose(info);
    std::printf("Simulation complete!\nTime taken:\t%.3f seconds\nTotal simulation time:\t%.7g seconds\n", seconds(), P.current_time);
    fjsph_destroy(e);
    fjsph_case_free(c);
    return 0;
}
