// example.x -- the program of example/example.f90 over the C-ABI of libsemic_b200.so.
//   ./example.x example.namelist
// Same namelist (&driver + &surface_physics), same forcing/validation files, same output file (8*nx columns per day:
// tsurf alb swnet smb melt acc shf lhf, example.f90:117-122).  The nloop x ntime day loop (example.f90:88-127) runs as
// ONE launch of the persistent kernel (semic_run) on device-resident forcing instead of nloop*ntime library calls.
#include <chrono>

#include "driver_util.h"

using namespace drv;

int main(int argc, char **argv)
{
    const char *prog = "example";
    std::printf(" " GREEN "example:" RESET " start example program.\n");
    if (argc < 2 || !argv[1][0]) die("Namelist file is missing.");                                    // example.f90:38-42
    const std::string nml = argv[1];
    std::printf(" " GREEN "example:" RESET " read namelist from file: %s\n", nml.c_str());
    const DriverNml d = read_driver_namelist(nml, 2);                                                 // :51-53
    const int nx = d.nx, ntime = d.ntime;

    const std::vector<double> forc = read_table(d.input_forcing, ntime, SEMIC_NFORCING, nx, prog, "forcing");      // :60
    const std::vector<double> vali = read_table(d.validation, ntime, SEMIC_NOUT, nx, prog, "validation");          // :61

    semic_par par;
    blank_par(par);
    par.nx = nx;                                                                                      // :56
    check(semic_surface_physics_par_load(&par, nml.c_str()), "surface_physics_par_load");             // :67
    check(semic_print_param(&par), "print_param");                                                    // :68

    semic_state *now = nullptr;
    check(semic_surface_alloc(&now, nx), "surface_alloc");                                            // :71
    int *mask = semic_state_mask(now);
    for (int i = 0; i < nx; ++i) mask[i] = d.loi_mask[i];                                             // :74
    const float alb0 = 0.8f;                                           // :77,:79 assign the single-precision literal 0.8
    struct { int f; double v; } init[] = {{SEMIC_F_HSNOW, 5.0}, {SEMIC_F_HICE, 0.0}, {SEMIC_F_ALB, (double)alb0},
                                          {SEMIC_F_TSURF, 260.0}, {SEMIC_F_ALB_SNOW, (double)alb0}};
    for (auto &q : init) {
        double *a = semic_state_field(now, q.f);
        for (int i = 0; i < nx; ++i) a[i] = q.v;
    }
    check(semic_state_upload(now, SEMIC_FIELDS_ALL), "state upload");

    semic_bnd bnd;
    check(semic_surface_boundary_define(&bnd, &par.boundary[0][0], SEMIC_NBOUNDARY, SEMIC_STRLEN), "surface_boundary_define");   // :82
    check(semic_print_boundary_opt(&bnd), "print_boundary_opt");

    semic_series *fs = nullptr, *vs = nullptr, *os = nullptr;
    check(semic_series_alloc(&fs, nx, SEMIC_NFORCING, ntime), "series_alloc");
    check(semic_series_alloc(&vs, nx, SEMIC_NOUT, ntime), "series_alloc");
    check(semic_series_alloc(&os, nx, SEMIC_NOUT, ntime), "series_alloc");
    check(semic_series_upload(fs, 0, ntime, forc.data(), nx), "series_upload");
    check(semic_series_upload(vs, 0, ntime, vali.data(), nx), "series_upload");

    std::printf(" " GREEN "example:" RESET " Write output to file " ULINE "%s" RESET "\n", d.output.c_str());
    semic_run_opts opts;
    std::memset(&opts, 0, sizeof opts);
    opts.out_days = ntime;                                                                            // :117 k == nloop
    float ms = 0.f;
    const auto t0 = std::chrono::steady_clock::now();
    check(semic_run(now, &par, &bnd, fs, vs, 0, d.nloop * ntime, os, &opts, &ms), "semic_run");       // :88-127
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    check(semic_state_download(now, SEMIC_FIELDS_ALL), "state download");

    std::vector<double> out((size_t)ntime * SEMIC_NOUT * nx);
    check(semic_series_download(os, 0, ntime, out.data(), nx, 0, nx), "series_download");
    std::FILE *f = std::fopen(d.output.c_str(), "w");
    if (!f) die("cannot open " + d.output);
    std::fprintf(f, " # tsurf(K)      alb      hsnow(m)      smb(m/s)     melt(m/s)      acc(m/s)\n");       // :86
    for (int day = 0; day < ntime; ++day) {
        const double *row = &out[(size_t)day * SEMIC_NOUT * nx];
        for (int k = 0; k < SEMIC_NOUT * nx; ++k) std::fprintf(f, "%s%24.16E", k ? " " : "", row[k]);
        std::fputc('\n', f);
    }
    std::fclose(f);

    semic_series_free(fs); semic_series_free(vs); semic_series_free(os);
    check(semic_surface_dealloc(now), "surface_dealloc");                                             // :132
    std::printf(" total time for surface_physics: %d %.6f s (kernel %.3f ms)\n", d.nloop, wall, ms);  // :134
    std::printf(" " GREEN "example:" RESET " end example program.\n");
    return 0;
}
