/*
 * ref_harness.c -- glue between the ctypes binding / bench.py and the MECHANICALLY GENERATED
 * reference build (oracle/_ref/semic_ref.c, made by f90_to_c.py from /root/reference/*.f90).
 * TEST INFRASTRUCTURE ONLY.  No model arithmetic here: objects are reached through the
 * reflection tables the generator emits (component name -> offset), and the day loop below only
 * copies rows in and out around the generated `surface_energy_and_mass_balance`, the way the
 * reference's own drivers do (example/example.f90:94-122, optimize/run_particles.f90:130-157).
 */
#include "f90_rt.h"

extern const f_typeinfo ref_types[];
void ref_surface_energy_and_mass_balance(void *now, void *par, void *bnd, int *day, int *year);

static const f_typeinfo *find_type(const char *type)
{
    const f_typeinfo *t;
    for (t = ref_types; t->name; ++t) if (strcmp(t->name, type) == 0) return t;
    return 0;
}

size_t ref_sizeof(const char *type)
{
    const f_typeinfo *t = find_type(type);
    return t ? t->size : 0;
}

void *ref_new(const char *type)
{
    const f_typeinfo *t = find_type(type);
    const f_compinfo *c;
    char *p;
    if (!t) return 0;
    p = (char *)calloc(1, t->size);
    /* character components start blank, as a Fortran program would leave them after `= ""` */
    for (c = t->comps; c && c->name; ++c) {
        if (c->kind == F_K_S) memset(p + c->offset, ' ', (size_t)c->charlen);
    }
    return p;
}

void ref_free(void *p) { free(p); }

/* address of component `comp` of an object of derived type `type`; kind/charlen/size returned */
void *ref_comp(const char *type, void *obj, const char *comp, int *kind, int *charlen)
{
    const f_typeinfo *t = find_type(type);
    const f_compinfo *c;
    if (!t) return 0;
    for (c = t->comps; c->name; ++c) {
        if (strcmp(c->name, comp) == 0) {
            if (kind) *kind = c->kind;
            if (charlen) *charlen = c->charlen;
            return (char *)obj + c->offset;
        }
    }
    return 0;
}

/* name of the i-th component (NULL past the end): lets the binding enumerate a type */
const char *ref_comp_name(const char *type, int i)
{
    const f_typeinfo *t = find_type(type);
    int k = 0;
    const f_compinfo *c;
    if (!t) return 0;
    for (c = t->comps; c->name; ++c, ++k) if (k == i) return c->name;
    return 0;
}

static double *arr(void *now, const char *name)
{
    f_arr_d *d = (f_arr_d *)ref_comp("surface_state_class", now, name, 0, 0);
    return d ? d->a : 0;
}

static int flag(void *bnd, const char *name)
{
    int *p = (int *)ref_comp("boundary_opt_class", bnd, name, 0, 0);
    return p ? *p : 0;
}

/*
 * Day loop around the generated time step.  Same row conventions as oracle_run():
 *   forcing [nwin][9][ldf] columns sf rf swd lwd wind sp rhoa qq t2m; day d uses row d % nwin
 *   vali    [nwin][8][ldf] columns stt alb swnet smb melt acc shf lhf, or NULL
 *   out     [nout][8][ldo] tsurf alb swd*(1-alb) smb melt acc shf lhf of the LAST out_days days
 */
int ref_run(void *now, void *par, void *bnd, int nx, const double *forcing, const double *vali, int nwin, int ldf,
            int ndays, double *out, int nout, int ldo, int out_days)
{
    static const char *fname[9] = {"sf", "rf", "swd", "lwd", "wind", "sp", "rhoa", "qq", "t2m"};
    static const char *oname[8] = {"tsurf", "alb", 0, "smb", "melt", "acc", "shf", "lhf"};
    static const struct { const char *flag, *field; int col; } ovr[5] = {
        {"tsurf", "tsurf", 0}, {"alb", "alb", 1}, {"melt", "melt", 4}, {"acc", "acc", 5}, {"smb", "smb", 3}};
    double *f[9], *o[8], *ov[5];
    int use[5];
    int d, v, i, year = 0;
    for (v = 0; v < 9; ++v) if (!(f[v] = arr(now, fname[v]))) return -1;
    for (v = 0; v < 8; ++v) o[v] = oname[v] ? arr(now, oname[v]) : 0;
    for (v = 0; v < 5; ++v) { ov[v] = arr(now, ovr[v].field); use[v] = vali && flag(bnd, ovr[v].flag); }
    for (d = 0; d < ndays; ++d) {
        const double *fr = forcing + (size_t)(d % nwin) * 9u * (size_t)ldf;
        int day = d % nwin + 1;
        for (v = 0; v < 9; ++v) memcpy(f[v], fr + (size_t)v * (size_t)ldf, sizeof(double) * (size_t)nx);
        for (v = 0; v < 5; ++v) {
            if (use[v])
                memcpy(ov[v], vali + ((size_t)(d % nwin) * 8u + (size_t)ovr[v].col) * (size_t)ldf, sizeof(double) * (size_t)nx);
        }
        ref_surface_energy_and_mass_balance(now, par, bnd, &day, &year);
        if (out && d >= ndays - out_days) {
            double *row = out + (size_t)((d - (ndays - out_days)) % nout) * 8u * (size_t)ldo;
            const double *swd = fr + 2 * (size_t)ldf;
            for (v = 0; v < 8; ++v) {
                if (v == 2) { for (i = 0; i < nx; ++i) row[2 * (size_t)ldo + i] = swd[i] * (1.0f - o[1][i]); }
                else memcpy(row + (size_t)v * (size_t)ldo, o[v], sizeof(double) * (size_t)nx);
            }
        }
    }
    return 0;
}
