// run_particles.x -- the program of optimize/run_particles.f90 over the C-ABI of libsemic_b200.so.
//   ./run_particles.x <prefix> <i0> <i1>
// Reads &driver from ./optimization.namelist, the forcing/validation tables, then for every particle i in [i0,i1] with
// a namelist <prefix>NNNNNN.nml writes the cost to <prefix>NNNNNN.out (run_particles.f90:104-178) -- the file protocol
// of optimize/tools.py:46-63.  All particles of the call are ONE ensemble launch (semic_ensemble_cost); with several
// GPUs visible the grid points are sharded over them, the shards run CONCURRENTLY (one host thread per device) and the
// per-particle sums are added on the host in device order (the NCCL path of the library is for one process per GPU,
// see semic_b200/driver.py and bench.py).
// Deviation: every particle starts from the same initial state; the reference carries the state of
// particle i-1 into particle i.
#include <chrono>
#include <cmath>
#include <sys/stat.h>
#include <thread>

#include "driver_util.h"

using namespace drv;

static std::string unquote(std::string s)            // read(arg,*) prefix: list-directed input strips quotes (tools.py:58)
{
    if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
    return s;
}

int main(int argc, char **argv)
{
    const char *prog = "run_particles";
    std::printf(" " GREEN "run_particles:" RESET " start particle calculation.\n");
    if (argc < 4) die("usage: run_particles.x <prefix> <i0> <i1>");
    const std::string prefix = unquote(argv[1]);
    const int i0 = std::atoi(argv[2]), i1 = std::atoi(argv[3]);
    const DriverNml d = read_driver_namelist("optimization.namelist", 2);                             // :61-64
    const int nx = d.nx, ntime = d.ntime;
    const std::vector<double> forc = read_table(d.input_forcing, ntime, SEMIC_NFORCING, nx, prog, "forcing");     // :70
    const std::vector<double> vali = read_table(d.validation, ntime, SEMIC_NOUT, nx, prog, "validation");         // :71

    // parameter sets: par keeps its contents from one particle's namelist to the next, as in the reference (:117)
    semic_par par;
    blank_par(par);
    par.nx = nx;
    std::vector<semic_par> pars;
    std::vector<int> ids;
    char name[512];
    for (int i = i0; i <= i1; ++i) {
        std::snprintf(name, sizeof name, "%s%06d.nml", prefix.c_str(), i);                            // :108
        struct stat st;
        if (stat(name, &st) != 0) {
            std::printf(" file '%s' not found, exiting loop\n", name);                                // :110-113
            break;
        }
        check(semic_surface_physics_par_load(&par, name), "surface_physics_par_load");                // :117
        if (i % 100 == 0) std::printf(" " GREEN "run_particle:" RESET " Load file %s\n", name);
        pars.push_back(par);
        ids.push_back(i);
    }

    if (!pars.empty()) {
        semic_bnd bnd;
        check(semic_surface_boundary_define(&bnd, &pars[0].boundary[0][0], SEMIC_NBOUNDARY, SEMIC_STRLEN), "surface_boundary_define");
        int ndev = semic_device_count();
        if (ndev < 1) die(std::string("no CUDA device: ") + semic_last_error());
        if (ndev > nx) ndev = nx;
        std::vector<double> sumsq(pars.size(), 0.0);
        std::vector<std::vector<double>> part(ndev, std::vector<double>(pars.size(), 0.0));
        auto run_shard = [&](int dev) {                              // contiguous point shard [lo, hi) on GPU `dev`
            const int lo = (int)((long long)nx * dev / ndev), hi = (int)((long long)nx * (dev + 1) / ndev), n = hi - lo;
            if (n == 0) return;
            check(semic_set_device(dev), "set_device");
            semic_state *now = nullptr;
            check(semic_surface_alloc(&now, n), "surface_alloc");                                     // :74
            int *mask = semic_state_mask(now);
            double *hsnow = semic_state_field(now, SEMIC_F_HSNOW), *hice = semic_state_field(now, SEMIC_F_HICE);
            double *alb = semic_state_field(now, SEMIC_F_ALB), *tsurf = semic_state_field(now, SEMIC_F_TSURF);
            double *alb_snow = semic_state_field(now, SEMIC_F_ALB_SNOW);
            for (int i = 0; i < n; ++i) {                                                             // :77-82
                mask[i] = d.loi_mask[lo + i];
                hsnow[i] = 1.0;
                hice[i] = 0.0;
                alb[i] = vali[(size_t)SEMIC_O_ALB * nx + lo + i];
                tsurf[i] = vali[(size_t)SEMIC_O_STT * nx + lo + i];
                alb_snow[i] = vali[(size_t)SEMIC_O_ALB * nx + lo + i];
            }
            check(semic_state_upload(now, SEMIC_FIELDS_ALL), "state upload");
            semic_series *fs = nullptr, *vs = nullptr;
            check(semic_series_alloc(&fs, n, SEMIC_NFORCING, ntime), "series_alloc");
            check(semic_series_alloc(&vs, n, SEMIC_NOUT, ntime), "series_alloc");
            check(semic_series_upload(fs, 0, ntime, forc.data() + lo, nx), "series_upload");
            check(semic_series_upload(vs, 0, ntime, vali.data() + lo, nx), "series_upload");
            float ms = 0.f;
            check(semic_ensemble_cost(now, pars.data(), (int)pars.size(), &bnd, fs, vs, ntime, d.nloop, nullptr, part[dev].data(), &ms),
                  "semic_ensemble_cost");                                                             // :124-174
            semic_series_free(fs); semic_series_free(vs);
            check(semic_surface_dealloc(now), "surface_dealloc");                                     // :181
        };
        {   // bring every device's context up first (in parallel), so that the figure below is the work itself
            std::vector<std::thread> th;
            for (int dev = 0; dev < ndev; ++dev)
                th.emplace_back([dev]() { check(semic_set_device(dev), "set_device"); check(semic_device_synchronize(), "device_synchronize"); });
            for (auto &t : th) t.join();
        }
        const auto t_begin = std::chrono::steady_clock::now();
        if (ndev == 1) {
            run_shard(0);
        } else {
            std::vector<std::thread> th;
            for (int dev = 0; dev < ndev; ++dev) th.emplace_back(run_shard, dev);
            for (auto &t : th) t.join();
        }
        for (int dev = 0; dev < ndev; ++dev)                          // fixed order: reproducible sums
            for (size_t p = 0; p < pars.size(); ++p) sumsq[p] += part[dev][p];
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        std::printf(" " GREEN "run_particle:" RESET " %zu particles x %d points x %d days x %d loops on %d GPU(s): %.3f s "
                    "(upload, ensemble kernel, download)\n", pars.size(), nx, ntime, d.nloop, ndev, secs);
        for (size_t p = 0; p < pars.size(); ++p) {
            std::snprintf(name, sizeof name, "%s%06d.out", prefix.c_str(), ids[p]);                   // :119-120
            std::FILE *f = std::fopen(name, "w");
            if (!f) die(std::string("cannot open ") + name);
            std::fprintf(f, " %24.16E\n", std::sqrt(sumsq[p]));                                       // :175-176 (NaN prints as NaN; tools.py:62 maps it)
            std::fclose(f);
        }
    }
    std::printf(" " GREEN "run_particle:" RESET " end particle calculation.\n");
    return 0;
}
