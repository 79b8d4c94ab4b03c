// fjsph_run.cpp — command-line driver: FJSPH's main() (reference src/FJSPH.cpp:29-356) on top of the C ABI.
//
//   fjsph_b200_run <para file> [--frames N] [--device K] [--restart file.fjr] [--out prefix] [--quiet] [--ranks N]
//
// --ranks N (N > 1): one process per GPU (forked here; rank r on device r), the case cut into N x-slabs of equal
// particle count, ghosts and migration over NCCL send/recv, the step's scalars by NCCL all-reduce on device buffers
// (comm_nccl.cu, include/fjsph_b200_nccl.h).  Every rank writes its own frame files (<prefix>_r<rank>_frame_*.dat); rank 0
// writes <prefix>_frame.info with the global counts.  Decks with inlet tables and restart files are single-GPU for now.
//
// GetInput + Init_Particles (fjsph_case_read) -> engine -> integrate_no_update at t = 0 (FJSPH.cpp:183) -> the frame
// loop `while (stept + 0.1 dt_min < frame_dt) integrate` (FJSPH.cpp:262-283) with the reference's per-step table
// (Integration.cpp:250-265), `<prefix>_frame.info` (FJSPH.cpp:228-243,286-300), one ASCII Tecplot zone file per frame
// (the fallback of AsciiIO.cpp; TECIO / HDF5 are not available here) and a restart file after every frame
// (csrc/restart.cu).  Everything numerical happens behind fjsph_step; this file is host bookkeeping only.
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fjsph_b200.h"
#include "../../include/fjsph_b200_nccl.h"

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

/* The tracker's side of the frame loop: Integration.cpp:151-169 hands the erased particles to IPT::Integrate at every step,
   Terminate_Particle queues their time records in iptdata, IPT::Write_Data (IPT.cpp:650-733) writes and clears the queue at
   every frame.  ASCII streaks only (ASCII::Write_Streaks / Write_Point, IPT.cpp:180-190,205-224); the reference's header
   leaves out the line break after its VARIABLES line (IPT.cpp:52-68), which is put in here. */
struct IptOutput
{
    FILE* streaks = nullptr;
    std::vector<FjsphIptPoint> queue;  /* the records of the particles ended since the last frame, one after the other */
    std::vector<int32_t> lengths;      /* records per particle */
    long long n_tracked = 0, n_success = 0, n_failed = 0;
    static constexpr int64_t CAP = 4096; /* records kept per particle */
    ~IptOutput()
    {
        if (streaks)
            std::fclose(streaks);
    }
    int open(const std::string& name, bool append, int offset_axis)
    {
        streaks = std::fopen(name.c_str(), append ? "a" : "w");
        if (!streaks)
        {
            std::fprintf(stderr, "Couldn't open the IPT streaks output file.\n");
            return 1;
        }
        if (!append)
        {
            const char* xyz = offset_axis == 1 ? "\"Y\", \"Z\"" : offset_axis == 2 ? "\"X\", \"Z\"" : offset_axis == 3 ? "\"X\", \"Y\""
                                                                                                      : "\"X\", \"Y\", \"Z\"";
            std::fprintf(streaks, "TITLE = \"IPT Streaks\"\nVARIABLES = %s, \"t\", \"dt\", \"v\", \"a\", \"ptID\", \"Cell_V\", \"Cell_Rho\", "
                                  "\"Cell_ID\"\n", xyz);
        }
        return 0;
    }
    int follow(FjsphEngine* e, const FjsphIptSettings& S)
    {
        int64_t n = 0;
        if (fjsph_take_deleted(e, nullptr, 0, &n))
            return 1;
        if (n == 0)
            return 0;
        std::vector<FjsphDeleted> in(static_cast<size_t>(n));
        if (fjsph_take_deleted(e, in.data(), n, &n))
            return 1;
        /* two passes over batches of particles: the first counts the records of every track, the second keeps them in a
           table as wide as the longest one (the marches cost little next to a step) */
        int64_t ok = 0, bad = 0;
        const int64_t BATCH = 65536;
        for (int64_t at = 0; at < n; at += BATCH)
        {
            const int64_t m = std::min(BATCH, n - at);
            std::vector<int32_t> n_rec(static_cast<size_t>(m));
            int64_t ok1 = 0, bad1 = 0;
            if (fjsph_ipt_integrate(e, &S, m, in.data() + at, nullptr, nullptr, nullptr, 0, n_rec.data(), &ok1, &bad1))
                return 1;
            ok += ok1;
            bad += bad1;
            const int64_t cap = std::min<int64_t>(*std::max_element(n_rec.begin(), n_rec.end()), CAP);
            std::vector<FjsphIptPoint> rec(static_cast<size_t>(m) * size_t(cap));
            if (fjsph_ipt_integrate(e, &S, m, in.data() + at, nullptr, nullptr, rec.data(), cap, n_rec.data(), nullptr, nullptr))
                return 1;
            for (int64_t i = 0; i < m; ++i)
            {
                const int32_t k = int32_t(std::min<int64_t>(n_rec[size_t(i)], cap));
                queue.insert(queue.end(), rec.begin() + i * cap, rec.begin() + i * cap + k);
                lengths.push_back(k);
            }
        }
        n_tracked += n;
        n_success += ok;
        n_failed += bad;
        return 0;
    }
    void write(int dim, double scale, bool streak_out)
    {
        size_t at = 0;
        for (const int32_t k : lengths)
        {
            if (streak_out && k > 0 && streaks)
            {
                std::fprintf(streaks, "ZONE T=\"Particle %lld\"\nI= %d, J=1, K=1, DATAPACKING=POINT\n", (long long)queue[at].part_id, k);
                for (int32_t j = 0; j < k; ++j)
                {
                    const FjsphIptPoint& q = queue[at + size_t(j)];
                    for (int d = 0; d < dim; ++d) std::fprintf(streaks, " %2.7e", q.xi[d] / scale);
                    std::fprintf(streaks, " %2.7e %2.7e %2.7e %2.7e %lld %2.7e %2.7e %6lld\n", q.t, q.dt,
                                 std::sqrt(q.v[0] * q.v[0] + q.v[1] * q.v[1] + q.v[2] * q.v[2]), q.acc, (long long)q.part_id,
                                 std::sqrt(q.cellV[0] * q.cellV[0] + q.cellV[1] * q.cellV[1] + q.cellV[2] * q.cellV[2]), q.cellRho,
                                 (long long)q.cellID);
                }
            }
            at += size_t(k);
        }
        if (streaks)
            std::fflush(streaks);
        queue.clear();
        lengths.clear();
    }
};

static int run(int argc, char** argv, int rank, int world, const char* nccl_id);

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s <para file> [--frames N] [--device K] [--restart file] [--out prefix] [--quiet] [--ranks N]\n",
                     argv[0]);
        return 2;
    }
    int ranks = 1;
    for (int a = 2; a + 1 < argc; ++a)
        if (std::string(argv[a]) == "--ranks")
            ranks = std::atoi(argv[a + 1]);
    if (ranks <= 1)
        return run(argc, argv, 0, 1, nullptr);
    /* one process per GPU.  Rank 0 makes the NCCL id AFTER the fork (ncclGetUniqueId starts a bootstrap thread, which a
       fork would not carry over) and hands it to the others through pipes made before it. */
    std::vector<int> rd(size_t(ranks), -1), wr(size_t(ranks), -1);
    for (int r = 1; r < ranks; ++r)
    {
        int fd[2];
        if (pipe(fd) != 0)
        {
            std::perror("pipe");
            return 1;
        }
        rd[size_t(r)] = fd[0];
        wr[size_t(r)] = fd[1];
    }
    std::vector<pid_t> kids;
    for (int r = 0; r < ranks; ++r)
    {
        const pid_t pid = fork();
        if (pid < 0)
        {
            std::perror("fork");
            return 1;
        }
        if (pid == 0)
        {
            char id[FJSPH_NCCL_ID_BYTES];
            if (r == 0)
            {
                if (fjsph_nccl_unique_id(id))
                {
                    std::fprintf(stderr, "ERROR: %s\n", fjsph_nccl_last_error());
                    _exit(1);
                }
                for (int q = 1; q < ranks; ++q)
                    if (write(wr[size_t(q)], id, sizeof(id)) != ssize_t(sizeof(id)))
                        _exit(1);
            }
            else if (read(rd[size_t(r)], id, sizeof(id)) != ssize_t(sizeof(id)))
                _exit(1);
            const int rc = run(argc, argv, r, ranks, id);
            std::fflush(nullptr);
            _exit(rc);
        }
        kids.push_back(pid);
    }
    int worst = 0;
    for (pid_t k : kids)
    {
        int status = 0;
        waitpid(k, &status, 0);
        const int rc = WIFEXITED(status) ? WEXITSTATUS(status) : 1;
        worst = std::max(worst, rc);
    }
    return worst;
}

static int run(int argc, char** argv, int rank, int world, const char* nccl_id)
{
    const char* para = argv[1];
    int frames_cli = -1, device = 0, dim = 3;
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
        else if (a + 1 < argc && k == "--dim") /* the SIMDIM of the build the deck was written for (makefile: 2D / 3D targets) */
            dim = std::atoi(argv[++a]);
        else if (a + 1 < argc && k == "--restart")
            restart_file = argv[++a];
        else if (a + 1 < argc && k == "--out")
            prefix_cli = argv[++a];
        else if (a + 1 < argc && k == "--ranks")
            ++a;
        else
        {
            std::fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 2;
        }
    }
    FjsphCase* c = nullptr;
    if (dim != 2 && dim != 3)
    {
        std::fprintf(stderr, "--dim must be 2 or 3\n");
        return 2;
    }
    if (fjsph_case_read(para, dim, &c))
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
    double tscale = 1.0, mesh_max_length = 0.0;
    {
        char fdir[1024] = "", fsol[1024] = "", tmesh[1024] = "", tsol[1024] = "";
        int32_t buoyant = 0;
        FjsphMesh view;
        fjsph_case_foam(c, fdir, fsol, &buoyant, 1024);
        fjsph_case_tau(c, tmesh, tsol, &tscale, 1024);
        if (tmesh[0])
        {
            /* FJSPH.cpp:72-91: the face-based mesh in 3D, the edge-based one in 2D */
            if ((dim == 2 ? fjsph_tau_read_edge(tmesh, tsol, tscale, fjsph_case_offset_axis(c), &aero_mesh)
                          : fjsph_tau_read(tmesh, tsol, tscale, &aero_mesh)) ||
                fjsph_foam_view(aero_mesh, &view))
                return fail("reading the TAU mesh");
            std::printf("TAU mesh: %lld cells, %lld faces\n", (long long)view.n_cells, (long long)view.n_faces);
            /* cells.maxlength, the tracker's bound on one step; only the TAU readers set it (CDFIO.cpp:867-898, 1117-1183) */
            if (fjsph_mesh_max_length(&view, dim, &mesh_max_length))
                return fail("measuring the TAU mesh");
        }
        else if (fdir[0])
        {
            if (fjsph_foam_read(fdir, fsol, buoyant, P.rho_g, &aero_mesh) || fjsph_foam_view(aero_mesh, &view))
                return fail("reading the OpenFOAM case");
            std::printf("OpenFOAM mesh: %lld cells, %lld triangles\n", (long long)view.n_cells, (long long)view.n_faces);
        }
    }
    const int64_t n_case = fjsph_case_count(c), nb_case = fjsph_case_bound_points(c);
    int64_t n0 = n_case, nb0 = nb_case;
    FjsphNcclComm* comm = nullptr;
    if (world > 1)
    {
        device = rank; /* one process per GPU */
        if (!restart_file.empty())
        {
            std::fprintf(stderr, "ERROR: --restart is single-GPU (gather the slabs first)\n");
            return 1;
        }
        if (fjsph_nccl_create(nccl_id, rank, world, device, &comm))
        {
            std::fprintf(stderr, "ERROR: %s\n", fjsph_nccl_last_error());
            return 1;
        }
        prefix += "_r" + std::to_string(rank);
    }
    FjsphEngine* e = nullptr;
    int32_t frame = 0;
    if (restart_file.empty())
    {
        std::vector<double> xi(3 * n_case), v(3 * n_case), rho(n_case), p(n_case), m(n_case);
        std::vector<int32_t> b(n_case);
        std::vector<int64_t> pid(n_case);
        FjsphStateView s;
        std::memset(&s, 0, sizeof(s));
        s.n = n_case;
        s.xi = xi.data();
        s.v = v.data();
        s.rho = rho.data();
        s.p = p.data();
        s.m = m.data();
        s.b = b.data();
        s.part_id = pid.data();
        if (fjsph_case_state(c, &s))
            return fail("reading the particles of the case");
        if (dim == 2) /* the case holds [n][2] vectors; the engine's view is [n][3] with z = 0 */
            for (int64_t i = n_case - 1; i >= 0; --i)
            {
                const size_t k = size_t(i);
                const double x0 = xi[2 * k], x1 = xi[2 * k + 1], v0 = v[2 * k], v1 = v[2 * k + 1];
                xi[3 * k] = x0, xi[3 * k + 1] = x1, xi[3 * k + 2] = 0.0;
                v[3 * k] = v0, v[3 * k + 1] = v1, v[3 * k + 2] = 0.0;
            }
        const int32_t nblk = fjsph_case_num_blocks(c);
        std::vector<FjsphBlock> blocks(nblk);
        for (int32_t i = 0; i < nblk; ++i) fjsph_case_block(c, i, &blocks[i], nullptr, 0);
        double x_lo = -1e300, x_hi = 1e300;
        if (world > 1)
        {
            /* x-slabs of equal particle count: cuts at the quantiles of x, half-way between the two particles either side;
               every rank keeps its particles in the case's order, so every block stays one contiguous range */
            for (const FjsphBlock& B : blocks)
                if (B.n_back > 0)
                {
                    std::fprintf(stderr, "ERROR: decks with inlet tables run on one GPU with this driver (--ranks 1)\n");
                    return 1;
                }
            std::vector<double> xs(static_cast<size_t>(n_case), 0.0);
            for (int64_t i = 0; i < n_case; ++i) xs[size_t(i)] = xi[3 * size_t(i)];
            std::sort(xs.begin(), xs.end());
            auto cut = [&](int r) {
                const size_t k = size_t(n_case) * size_t(r) / size_t(world);
                return 0.5 * (xs[k - 1] + xs[k]);
            };
            if (rank > 0)
                x_lo = cut(rank);
            if (rank < world - 1)
                x_hi = cut(rank + 1);
            std::vector<int64_t> keep;
            for (int64_t i = 0; i < n_case; ++i)
                if (xi[3 * size_t(i)] >= x_lo && xi[3 * size_t(i)] < x_hi)
                    keep.push_back(i);
            std::vector<int64_t> before(static_cast<size_t>(n_case) + 1, 0); /* kept particles with a smaller case index */
            {
                size_t k = 0;
                for (int64_t i = 0; i <= n_case; ++i)
                {
                    while (k < keep.size() && keep[k] < i) ++k;
                    before[size_t(i)] = int64_t(k);
                }
            }
            for (FjsphBlock& B : blocks)
            {
                B.first = before[size_t(B.first)];
                B.second = before[size_t(B.second)];
            }
            n0 = int64_t(keep.size());
            nb0 = before[size_t(nb_case)];
            for (size_t k = 0; k < keep.size(); ++k)
            {
                const size_t i = size_t(keep[k]);
                for (int d = 0; d < 3; ++d)
                {
                    xi[3 * k + size_t(d)] = xi[3 * i + size_t(d)];
                    v[3 * k + size_t(d)] = v[3 * i + size_t(d)];
                }
                rho[k] = rho[i];
                p[k] = p[i];
                m[k] = m[i];
                b[k] = b[i];
                pid[k] = pid[i];
            }
            s.n = n0;
        }
        const int64_t capacity = (world == 1 && max_points > n0) ? max_points : 2 * n0 + 100000;
        if (fjsph_create(&P, device, capacity, &e))
            return fail("creating the engine");
        if (fjsph_upload_state(e, &s, nb0))
            return fail("uploading the particles");
        if (nblk && fjsph_set_blocks(e, nblk, blocks.data()))
            return fail("setting the blocks");
        if (world > 1 && fjsph_nccl_attach(comm, e, x_lo, x_hi))
        {
            std::fprintf(stderr, "ERROR: %s\n", fjsph_nccl_last_error());
            return 1;
        }
    }
    else
    {
        const int64_t capacity = max_points > n0 ? max_points : 2 * n0 + 100000;
        if (fjsph_create(&P, device, capacity, &e))
            return fail("creating the engine");
        if (fjsph_read_restart(e, restart_file.c_str(), &frame))
            return fail("reading the restart file");
    }
    /* particle counts of the whole domain for the frame table (a rank's own counts change with migration) */
    auto global_counts = [&](long long& n_all, long long& n_bound) -> int {
        double cnt[2] = {double(fjsph_count(e)), double(nb0)};
        if (comm && fjsph_nccl_allreduce_host(comm, cnt, 2, FJSPH_COMM_SUM))
            return 1;
        n_all = (long long)cnt[0];
        n_bound = (long long)cnt[1];
        return 0;
    };
    if (aero_mesh)
    {
        FjsphMesh view;
        if (fjsph_foam_view(aero_mesh, &view) || fjsph_upload_mesh(e, &view))
            return fail("uploading the aero mesh");
        fjsph_foam_free(aero_mesh);
    }
    fjsph_get_params(e, &P);
    /* particle tracking downstream of the delete planes (FJSPH.cpp:206-210, Integration.cpp:151-169): only with an aero mesh */
    FjsphIptSettings ipt;
    int32_t using_ipt = 0;
    if (fjsph_ipt_default_settings(&P, &ipt) || fjsph_read_para_ipt(para, tscale, &using_ipt, &ipt))
        return fail("reading the particle-tracking settings");
    ipt.max_length = mesh_max_length;
    using_ipt = using_ipt && P.asource != 0;
    if (using_ipt && mesh_max_length == 0.0 && rank == 0)
        std::printf("WARNING: cells.maxlength is 0 (only the TAU readers set it, CDFIO.cpp:931,1214): as in FJSPH, every tracked "
                    "particle fails its first step on this mesh.\n");
    IptOutput tracks;
    if (using_ipt && tracks.open(prefix + "_IPT_streaks.dat", !restart_file.empty(), fjsph_case_offset_axis(c)))
        return 1;
    long long n_all = 0, n_bound = 0;
    if (global_counts(n_all, n_bound))
        return 1;
    if (rank == 0)
        std::printf("Starting counts:\nBoundary: %lld  Sim: %lld%s\n\n", n_bound, n_all - n_bound,
                    world > 1 ? (" (" + std::to_string(world) + " x-slabs, one GPU each)").c_str() : "");
    quiet = quiet || rank != 0;

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
        std::fprintf(info, "Frame: %u\nTotal Points: %lld Boundary Points: %lld Fluid Points: %lld\n", 0u, n_all, n_bound,
                     n_all - n_bound);
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
            if (using_ipt && st.n_del > 0 && tracks.follow(e, ipt))
                return fail("tracking the deleted particles");
            if (!quiet)
                std::printf("%9.3e | %7.2e | %4.2f | %9.4f | %3d | %8.3f | %9.3e | %9.3e | %9.3e | %14ld|\n", P.current_time - st.dt,
                            st.dt, st.cfl_ratio, st.rms_error, st.iterations, st.maxRho_pc, st.maxf, st.maxAf, st.maxShift, long(ms));
            stept += st.dt;
            ++stepits;
        }
        if (global_counts(n_all, n_bound))
            return 1;
        const long long n = n_all;
        std::fprintf(info, "\nFrame: %u\nTotal Points: %lld Boundary Points: %lld Fluid Points: %lld\n", unsigned(fr), n, n_bound,
                     n - n_bound);
        std::fprintf(info, "Sim Time:  %.7g Comp Time: %.6e Error: %.6f Sub-iterations: %d\n", P.current_time, seconds(), error, stepits);
        std::fprintf(info, "Deleted particles: %lld Internal collisions: %d\n", deleted, 0);
        std::fflush(info);
        if (rank == 0)
        {
            std::printf("Frame: %d  Sim Time: %.7g  Compute Time: %.3f  Error: %.5f\n", fr, P.current_time, seconds(), error);
            std::printf("Boundary particles:  %lld Sim particles: %lld Deleted particles: %lld\n", n_bound, n - n_bound, deleted);
        }
        if (n - n_bound == 0)
        {
            std::printf("No more points in the simulation space. Ending....\n");
            break;
        }
        if (write_frame(e, prefix, fr, P.current_time, P.rho_rest))
            return fail("writing a frame");
        if (using_ipt) /* IPT::Write_Data, FJSPH.cpp:319-320 */
            tracks.write(dim, tscale, ipt.record != 0);
        /* march the frame time forward (FJSPH.cpp:326-327) BEFORE the checkpoint, so that a resumed run clamps its
           first steps to the end of the NEXT frame (find_timestep, Integration.cpp:433-440) */
        P.last_frame_time += P.frame_time_interval;
        if (fjsph_set_params(e, &P))
            return fail("set_params");
        if (world == 1 && fjsph_write_restart(e, (prefix + "_particles.fjr").c_str(), fr))
            return fail("writing the restart file");
    }
    std::fclose(info);
    if (using_ipt)
        std::printf("Particle tracking: %lld particles followed, %lld left the mesh or passed the end plane, %lld failed\n",
                    tracks.n_tracked, tracks.n_success, tracks.n_failed);
    if (rank == 0)
        std::printf("Simulation complete!\nTime taken:\t%.3f seconds\nTotal simulation time:\t%.7g seconds\n", seconds(), P.current_time);
    if (comm && rank == 0)
        std::printf("NCCL transport: %lld device all-reduces, %lld host all-reduces, %lld + %lld neighbour exchanges\n",
                    (long long)(fjsph_nccl_calls(comm, FJSPH_COMM_SUM_DEV) + fjsph_nccl_calls(comm, FJSPH_COMM_MAX_DEV)),
                    (long long)(fjsph_nccl_calls(comm, FJSPH_COMM_SUM) + fjsph_nccl_calls(comm, FJSPH_COMM_MAX)),
                    (long long)fjsph_nccl_calls(comm, FJSPH_COMM_SENDRECV_DEV_ASYNC),
                    (long long)fjsph_nccl_calls(comm, FJSPH_COMM_SENDRECV_DEV));
    fjsph_destroy(e);
    if (comm)
        fjsph_nccl_destroy(comm);
    fjsph_case_free(c);
    return 0;
}
